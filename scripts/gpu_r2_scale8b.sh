#!/bin/bash
# Round-2 final 8-GPU check: BASELINE config 5 through bench.py exactly as the driver launches it.
OUT=gpurun_out/${1:-r2v}; mkdir -p $OUT
for wl in dgemm sgemm; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload $wl --steps 5 --warmup 3 > $OUT/bench_n8_$wl.json 2> $OUT/bench_n8_$wl.err
  echo "bench n=8 $wl rc=$?"
done
python - <<PY
import json
for f in ("bench_n8_dgemm", "bench_n8_sgemm"):
    try:
        d = json.loads([l for l in open("$OUT/%s.json" % f) if l.startswith("{")][-1])
        print(f, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["parity"]["max_rel_err_sampled_blocks"], d["parity"]["ok"], d["kernel"], d["clocks"], d["roofline"]["frac"])
    except Exception as ex:
        print(f, "unreadable", ex)
PY
