#!/bin/bash
# Run the UNMODIFIED GPU-BLOB harness with (a) the reference cuBLAS backend and
# (b) the B200 backend on the same box: dense sweep of the threshold region and
# sampled large points.  Usage: scripts/gpu_sweep.sh <tag> <dense-upper> "<points>"
TAG=$1; DENSE=${2:-256}; POINTS=${3:-"512 1024 2048 4096 8192"}
OUT=$PWD/gpurun_out/$TAG
mkdir -p $OUT
export OMP_NUM_THREADS=$(nproc)
B200=$PWD/gpu-blas-offload-benchmark_b200/lib/gpu-blob-b200
CUBLAS=$PWD/oracle/_ref/gpu-blob-cublas
cd $OUT
for impl in cublas b200; do
  BIN=$CUBLAS; [ $impl = b200 ] && BIN=$B200
  timeout 1200 $BIN -i 10 -s 1 -d $DENSE -o $OUT/${impl}_dense > $OUT/${impl}_dense.log 2>&1
  echo "$impl dense rc=$?"
  for d in $POINTS; do
    timeout 900 $BIN -i 10 -s $d -d $d -o $OUT/${impl}_d$d > $OUT/${impl}_d$d.log 2>&1
    echo "$impl d=$d rc=$?"
  done
done
for d in $POINTS; do
  for k in dgemm sgemm dgemv sgemv; do
    f=${k}_square_square_M=N=K.csv; [ ${k:1} = gemv ] && f=${k}_square_vector_M=N.csv
    echo "--- $k d=$d"; paste -d'|' <(cut -d, -f1,9 cublas_d$d/$f) <(cut -d, -f1,9 b200_d$d/$f)
  done
done
