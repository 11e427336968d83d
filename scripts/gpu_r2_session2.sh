#!/bin/bash
# Round-2 session: full GPU test suite, tall-K kernel A/B, bench lines, ncu captures of the changed kernels.
TAG=${1:-r2i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/box.txt 2>&1
timeout -k 10 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_all.txt 2>&1
echo "pytest_all rc=$?" >> $OUT/box.txt
tail -4 $OUT/pytest_all.txt | cut -c1-300
for t in on off; do
  for kind in dgemm sgemm; do
    for K in 1024 2048 8192 32768; do
      B200_TALLK=$t timeout 120 python scripts/quick_gemm.py $kind 32 32 $K 20 >> $OUT/tallk_$t.txt 2>&1
    done
  done
done
for t in on off; do echo "== tallk=$t"; cat $OUT/tallk_$t.txt; done
# bench lines (own arm), then ncu: launch list + full capture for the SGEMM workload (fused kernel)
timeout 900 python bench.py > $OUT/bench_dgemm.json 2> $OUT/bench_dgemm.err; echo "bench dgemm rc=$?" >> $OUT/box.txt
timeout 900 python bench.py --workload sgemm --no-extras > $OUT/bench_sgemm.json 2> $OUT/bench_sgemm.err; echo "bench sgemm rc=$?" >> $OUT/box.txt
SPECS="sgemm:sgemm_3xtf32" timeout 1200 bash scripts/gpu_profile_all.sh $TAG/prof >> $OUT/box.txt 2>&1
# full captures of the K = 32 shapes and the tall-K kernel
Q="python scripts/quick_gemm.py"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgemm_3xtf32 -s 3 -c 1 -o $OUT/sgemm_k32 $Q sgemm 8192 8192 32 5 > $OUT/ncu_sgemm_k32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -s 3 -c 1 -o $OUT/dgemm_k32 $Q dgemm 8192 8192 32 5 > $OUT/ncu_dgemm_k32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tallk -s 3 -c 1 -o $OUT/dgemm_tallk $Q dgemm 32 32 8192 5 > $OUT/ncu_dgemm_tallk.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgemm_3xtf32 -s 3 -c 1 -o $OUT/sgemm_2048 $Q sgemm 2048 2048 2048 5 > $OUT/ncu_sgemm_2048.log 2>&1
ls -la $OUT $OUT/prof 2>/dev/null | head -40
python - <<PY
import json
for f in ("bench_dgemm", "bench_sgemm"):
    try:
        d = json.loads([l for l in open("$OUT/%s.json" % f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["kernel"], d.get("gpu_baseline"))
        for k, v in (d.get("extras") or {}).items():
            if isinstance(v, dict) and "gflops" in v: print("   ", k, v["gflops"], v["kernel"], round(v["roofline"]["frac"], 3), v["e2e"]["value"])
    except Exception as ex:
        print(f, "unreadable", ex)
PY
cat $OUT/box.txt
