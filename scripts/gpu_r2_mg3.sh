#!/bin/bash
# Round-2 multi-GPU session 3: team tests + Once-mode trace of the 4096^3 slab problem.
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export OMP_NUM_THREADS=$(nproc)
timeout -k 10 1500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $OUT/pytest_multi.txt 2>&1; echo "pytest_multi rc=$?" > $OUT/box.txt
P=gpu-blas-offload-benchmark_b200/lib/b200_parity
for mode in once always; do
  for d in 4096 6144; do
    B200_MG_TRACE=1 B200_NGPUS=2 timeout 300 $P dgemm $mode $d $d $d 10 /dev/null > $OUT/trace_${mode}_$d.txt 2>&1
  done
done
B200_NGPUS=2 timeout 300 $P dgemm once 4096 4096 4096 10 /dev/null > $OUT/notrace_once_4096.txt 2>&1
B200_NGPUS=1 timeout 300 $P dgemm once 4096 4096 4096 10 /dev/null > $OUT/notrace_once_4096_ng1.txt 2>&1
for shape in "4096 2048 4096" "4096 4096 4096" "6144 3072 6144" "8192 4096 8192"; do
  timeout 120 python scripts/quick_gemm.py dgemm $shape 10 >> $OUT/quick.txt 2>&1
done
tail -15 $OUT/pytest_multi.txt | cut -c1-250
for f in $OUT/trace_*.txt $OUT/notrace*.txt; do echo "== $f"; cat $f | cut -c1-200; done
cat $OUT/quick.txt $OUT/box.txt
