#!/bin/bash
# round-1 session k: split-K / skinny-shape fixes (A/B against the previous dispatch through
# env overrides), Unified prefetch lanes A/B, ncu of the K=32 shapes.
OUT=gpurun_out/r1k; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_all.txt 2>&1
echo "pytest all rc=$?"; tail -3 $OUT/pytest_all.txt
timeout 900 python scripts/shape_sweep.py --md $OUT/shape_sweep.md > $OUT/shape_sweep.log 2>&1
echo "shape sweep rc=$?"
{
for sh in "512 512 8192" "256 256 4096" "384 384 6144" "128 128 2048" "32 32 8192" "32 32 2048" "8192 32 32" "32 8192 32"; do
  for tile in auto 32 128; do
    if [ $tile = auto ]; then timeout 120 python scripts/quick_gemm.py dgemm $sh 20
    else B200_DGEMM_TILE=$tile timeout 120 python scripts/quick_gemm.py dgemm $sh 20; fi
  done
  B200_DGEMM_MIN=48 timeout 120 python scripts/quick_gemm.py dgemm $sh 20
done
} > $OUT/dgemm_tile_ab.txt 2>&1
cat $OUT/dgemm_tile_ab.txt
for lanes in 3 1 3 1; do
  KINDS=gemm PF_LANES=$lanes REPS=30 timeout 600 python scripts/unified_phases.py 1024 2048 2>&1 | grep -A10 "^---" | grep "^---\|pf_to_gpu\|timed" >> $OUT/unified_lanes.txt
done
cat $OUT/unified_lanes.txt
# ncu: K = 32 shapes (why ~2 TB/s?) -- one launch each
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pair -c 1 -o $OUT/sgemm_k32 python scripts/quick_gemm.py sgemm 8192 8192 32 3 > $OUT/ncu_sgemm_k32.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dgemm_tma -c 1 -o $OUT/dgemm_k32 python scripts/quick_gemm.py dgemm 8192 8192 32 3 > $OUT/ncu_dgemm_k32.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pair -s 2 -c 1 -o $OUT/sgemm_8192_pair python scripts/quick_gemm.py sgemm 8192 8192 8192 2 > $OUT/ncu_sgemm_pair.log 2>&1
ls -la $OUT
