#!/bin/bash
# Round-2 session 4: DGEMM TMA-store epilogue (parity + A/B), Unified-mode guard, full suite, shape sweep.
TAG=${1:-r2k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/box.txt 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_dgemm_streamk.py tests/test_gpu_unified_guard.py -x -q -s -m gpu > $OUT/pytest_new.txt 2>&1
echo "pytest_new rc=$?" >> $OUT/box.txt
grep -v "^$" $OUT/pytest_new.txt | tail -15 | cut -c1-300
for st in on off; do
  for shape in "8192 8192 32" "2048 2048 32" "8192 8192 512" "2048 2048 2048" "8192 512 512" "2048 2048 128" "4096 4096 4096"; do
    echo -n "dgemm_tma_store=$st " >> $OUT/ab_dgemm.txt
    B200_DGEMM_TMA_STORE=$st timeout 120 python scripts/quick_gemm.py dgemm $shape 10 >> $OUT/ab_dgemm.txt 2>&1
  done
done
cat $OUT/ab_dgemm.txt
timeout -k 10 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_all.txt 2>&1
echo "pytest_all rc=$?" >> $OUT/box.txt
tail -4 $OUT/pytest_all.txt | cut -c1-300
timeout 900 python scripts/shape_sweep.py > $OUT/shape_sweep.md 2> $OUT/shape_sweep.err
echo "shape_sweep rc=$?" >> $OUT/box.txt
cat $OUT/shape_sweep.md | cut -c1-230
cat $OUT/box.txt
