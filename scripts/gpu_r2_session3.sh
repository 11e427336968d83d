#!/bin/bash
# Round-2 session 3 (2 GPUs): full test suite incl. the team tests, SGEMM TMA-store epilogue A/B, 2-GPU SGEMM step.
TAG=${1:-r2j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/box.txt 2>&1
timeout -k 10 1800 python -m pytest tests -x -q -m gpu > $OUT/pytest_all.txt 2>&1
echo "pytest_all rc=$?" >> $OUT/box.txt
tail -4 $OUT/pytest_all.txt | cut -c1-300
for st in on off; do
  for rg in 8 32; do
    for shape in "8192 8192 32" "2048 2048 32" "8192 8192 512" "2048 2048 2048" "4096 4096 4096" "8192 8192 8192"; do
      echo -n "tma_store=$st raster=$rg " >> $OUT/ab.txt
      B200_SGEMM_TMA_STORE=$st B200_SGEMM_RASTER=$rg timeout 120 python scripts/quick_gemm.py sgemm $shape 10 >> $OUT/ab.txt 2>&1
    done
  done
done
cat $OUT/ab.txt
for wl in sgemm dgemm; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --dim 16384 --steps 3 --warmup 3 > $OUT/bench_n2_${wl}_16384.json 2> $OUT/bench_n2_${wl}_16384.err
  echo "bench n2 $wl rc=$?" >> $OUT/box.txt
done
python - <<PY
import json
for f in ("bench_n2_sgemm_16384", "bench_n2_dgemm_16384"):
    try:
        d = json.loads([l for l in open("$OUT/%s.json" % f) if l.startswith("{")][-1])
        print(f, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["parity"], d["kernel"], d["clocks"])
    except Exception as ex:
        print(f, "unreadable", ex)
PY
cat $OUT/box.txt
