#!/bin/bash
# round-1 session j: pair-kernel SGEMM parity + timing, per-problem-type roofline sweep,
# Unified-mode phase timing.
OUT=gpurun_out/r1j; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sgemm_tc.py -x -q -m gpu > $OUT/pytest_sgemm.txt 2>&1
echo "sgemm pytest rc=$?"; tail -5 $OUT/pytest_sgemm.txt
for p in on off; do
  B200_SGEMM_PAIR=$p timeout 300 python scripts/size_sweep.py sgemm 1024 2048 3072 4096 6144 8192 12288 16384 >> $OUT/sgemm_pair_ab.txt 2>&1
done
cat $OUT/sgemm_pair_ab.txt
timeout 900 python scripts/shape_sweep.py --md $OUT/shape_sweep.md > $OUT/shape_sweep.log 2>&1
echo "shape sweep rc=$?"
REPS=12 timeout 600 python scripts/unified_phases.py 1024 2048 4096 > $OUT/unified_phases.txt 2>&1
echo "unified rc=$?"; cat $OUT/unified_phases.txt
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_all.txt 2>&1
echo "pytest all rc=$?"; tail -3 $OUT/pytest_all.txt
