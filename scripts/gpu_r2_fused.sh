#!/bin/bash
# Round-2: fused 3xTF32 operand split (converter warps) vs the pre-split pass: parity, then A/B timing.
TAG=${1:-r2h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/box.txt 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_sgemm_tc.py tests/test_gpu_sgemm_streamk.py -x -q -m gpu > $OUT/pytest_sgemm.txt 2>&1
echo "pytest_sgemm rc=$?" >> $OUT/box.txt
tail -5 $OUT/pytest_sgemm.txt | cut -c1-300
for f in on off; do
  B200_SGEMM_FUSED=$f timeout 300 python scripts/size_sweep.py sgemm 512 768 1024 1280 1536 1792 2048 2304 2560 3072 4096 5120 6144 8192 >> $OUT/sweep.txt 2>&1
  for shape in "8192 8192 512" "8192 8192 32" "512 512 8192" "8192 512 512" "512 8192 512" "2048 2048 128" "2048 2048 32" "16384 4096 16384"; do
    B200_SGEMM_FUSED=$f timeout 120 python scripts/quick_gemm.py sgemm $shape 10 >> $OUT/quick_$f.txt 2>&1
  done
done
cat $OUT/sweep.txt; for f in on off; do echo "== fused=$f"; cat $OUT/quick_$f.txt; done
timeout -k 10 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_all.txt 2>&1
echo "pytest_all rc=$?" >> $OUT/box.txt
tail -5 $OUT/pytest_all.txt | cut -c1-300
cat $OUT/box.txt
