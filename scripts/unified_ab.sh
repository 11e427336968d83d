#!/bin/bash
# A/B the unified-memory policy with the unmodified harness: per-iteration time of the
# gpu_unified rows, with and without cudaMemAdvise.
OUT=$PWD/gpurun_out/${1:-unified_ab}; mkdir -p $OUT; export OMP_NUM_THREADS=$(nproc)
BIN=$PWD/gpu-blas-offload-benchmark_b200/lib/gpu-blob-b200
cd $OUT
B200_ADVISE=all $BIN -i 10 -s 1 -d 384 --no_cpu -o $OUT/advise > $OUT/advise.log 2>&1
B200_ADVISE=none $BIN -i 10 -s 1 -d 384 --no_cpu -o $OUT/noadvise > $OUT/noadvise.log 2>&1
for d in 1024 2048 4096; do
  B200_ADVISE=all $BIN -i 10 -s $d -d $d --no_cpu -o $OUT/advise_$d > /dev/null 2>&1
  B200_ADVISE=none $BIN -i 10 -s $d -d $d --no_cpu -o $OUT/noadvise_$d > /dev/null 2>&1
done
python - <<'PY'
import csv,statistics,os
def col(p):
    return [float(r['Total Seconds'])*1e5 for r in csv.DictReader(open(p)) if r['Device']=='gpu_unified']
for f in ('dgemm_square_square_M=N=K.csv','sgemm_square_square_M=N=K.csv','dgemv_square_vector_M=N.csv','sgemv_square_vector_M=N.csv'):
    a=col('advise/'+f); b=col('noadvise/'+f)
    q=lambda x: (statistics.median(x), sorted(x)[int(len(x)*0.9)], max(x))
    print(f[:5], 'advise   median/p90/max us/iter = %.0f %.0f %.0f' % q(a), '| noadvise = %.0f %.0f %.0f' % q(b))
    for d in (1024,2048,4096):
        a=col(f'advise_{d}/'+f); b=col(f'noadvise_{d}/'+f)
        print('   d=%d advise %.0f us/iter | noadvise %.0f' % (d, a[0], b[0]))
PY
