#!/bin/bash
# One GPU-box session: probe, smoke, parity tests, bench, baselines.  Everything
# lands in gpurun_out/ (merged back by gpurun).  Usage: scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export OMP_NUM_THREADS=$(nproc)
{
  echo "== box"; nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,pcie.link.gen.current,pcie.link.width.current --format=csv
  ls /root/reference 2>&1 | head -2
} > $OUT/box.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "smoke rc=$?" >> $OUT/box.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/box.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/box.txt
tail -5 $OUT/smoke.txt; tail -15 $OUT/pytest.txt; cat $OUT/bench.json; tail -5 $OUT/bench.err; cat $OUT/box.txt
