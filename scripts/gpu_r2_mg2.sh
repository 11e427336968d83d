#!/bin/bash
# Round-2 multi-GPU session 2: team tests, harness CSVs with B200_NGPUS, N = 2 bench.
TAG=${1:-r2b}
NG=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export OMP_NUM_THREADS=$(nproc)
timeout -k 10 1500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $OUT/pytest_multi.txt 2>&1; echo "pytest_multi rc=$?" > $OUT/box.txt
BLOB=$PWD/gpu-blas-offload-benchmark_b200/lib/gpu-blob-b200
for ng in 1 $NG; do
  for d in 4096 8192; do
    mkdir -p $OUT/harness_ng${ng}_$d
    ( cd $OUT/harness_ng${ng}_$d && B200_NGPUS=$ng timeout -k 10 600 $BLOB -i 10 -s $d -d $d --no_cpu -o csv > log.txt 2>&1 ); echo "harness ng=$ng d=$d rc=$?" >> $OUT/box.txt
  done
done
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $NG --steps 3 --warmup 3 > $OUT/bench_n${NG}_dgemm.json 2> $OUT/bench_n${NG}_dgemm.err
echo "bench rc=$?" >> $OUT/box.txt
tail -30 $OUT/pytest_multi.txt | cut -c1-250
for ng in 1 $NG; do for d in 4096 8192; do echo "== ng=$ng d=$d"; grep -h "gpu" $OUT/harness_ng${ng}_$d/csv/dgemm_square_square_M=N=K.csv $OUT/harness_ng${ng}_$d/csv/sgemm_square_square_M=N=K.csv $OUT/harness_ng${ng}_$d/csv/dgemv_square_vector_M=N.csv; done; done
cut -c1-400 $OUT/bench_n${NG}_dgemm.json; tail -3 $OUT/bench_n${NG}_dgemm.err
cat $OUT/box.txt
