#!/bin/bash
# Round-2 session 5: stage-release race fix in the TMA DGEMM kernel: regression test, full suite, perf check.
TAG=${1:-r2m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -k 10 900 python -m pytest tests/test_gpu_dgemm_streamk.py tests/test_gpu_always_pipeline.py -x -q -m gpu > $OUT/pytest_fix.txt 2>&1
echo "pytest_fix rc=$?" > $OUT/box.txt
tail -4 $OUT/pytest_fix.txt | cut -c1-300
timeout 300 python scripts/debug_dgemm_tma_store.py 2>&1 | grep -v "no dgemm_tma_store" | tee $OUT/chains.txt | cut -c1-250
for shape in "8192 8192 8192" "4096 4096 4096" "2048 2048 2048" "8192 8192 512" "8192 8192 32" "512 512 8192"; do
  timeout 120 python scripts/quick_gemm.py dgemm $shape 10 >> $OUT/dgemm_perf.txt 2>&1
done
cat $OUT/dgemm_perf.txt
timeout -k 10 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_all.txt 2>&1
echo "pytest_all rc=$?" >> $OUT/box.txt
tail -4 $OUT/pytest_all.txt | cut -c1-300
cat $OUT/box.txt
