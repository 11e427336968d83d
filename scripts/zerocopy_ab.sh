#!/bin/bash
# A/B the Always-mode zero-copy limits with the unmodified harness (gpu_offloadAlways rows).
# usage: zerocopy_ab.sh <tag> <upper-dim> "<gemv limits>"
OUT=$PWD/gpurun_out/${1:-zc_ab}; D=${2:-448}; LIMS=${3:-"0 4194304 16777216 67108864"}
mkdir -p $OUT; export OMP_NUM_THREADS=$(nproc)
BIN=$PWD/gpu-blas-offload-benchmark_b200/lib/gpu-blob-b200
cd $OUT
for lim in $LIMS; do
  B200_ZEROCOPY_GEMV_BYTES=$lim $BIN -i 10 -s $D -d $D --no_cpu -o $OUT/zc_$lim > $OUT/zc_$lim.log 2>&1 || tail -3 $OUT/zc_$lim.log
done
python - $LIMS <<'PY'
import csv,sys,glob,os
for f in ('dgemv_square_vector_M=N.csv','sgemv_square_vector_M=N.csv','dgemv_tall-thin_vector_M_N=32.csv','dgemv_short-wide_vector_M=32_N.csv'):
    for lim in sys.argv[1:]:
        rows=[r for r in csv.DictReader(open(f'zc_{lim}/{f}')) if r['Device']=='gpu_offloadAlways']
        for r in rows: print(f"{f[:-4]:34s} M={r['M']:>5s} N={r['N']:>5s} limit {int(lim):9d}: {float(r['Total Seconds'])*1e5:8.0f} us/iter")
PY
