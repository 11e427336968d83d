#!/bin/bash
# evidence pass: tests, smoke, bench lines of the four workloads, reference arm, shape sweep, engine
# traces, ncu launch lists + full captures (every command individually bounded).  usage: gpu_evidence.sh <tag>
OUT=gpurun_out/${1:-r1p}; mkdir -p $OUT
export OMP_NUM_THREADS=$(nproc)
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_all.txt 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_all.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.txt
timeout 600 python bench.py > $OUT/bench_dgemm.json 2> $OUT/bench_dgemm.err; echo "bench rc=$?"; cat $OUT/bench_dgemm.json
for wl in sgemm dgemv sgemv; do
  timeout 300 python bench.py --workload $wl --no-extras > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "bench $wl rc=$?"; cat $OUT/bench_$wl.json
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>&1
timeout 400 python scripts/shape_sweep.py --md $OUT/shape_sweep.md > $OUT/shape_sweep.log 2>&1; echo "shape sweep rc=$?"
B200_ENGINE_TRACE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-extras > $OUT/engine_trace_bench.json 2> $OUT/engine_trace.txt
B200_ENGINE_TRACE=1 timeout 200 python bench.py --workload sgemm --steps 3 --warmup 3 --no-extras > $OUT/engine_trace_bench_sgemm.json 2> $OUT/engine_trace_sgemm.txt
timeout 900 bash scripts/gpu_profile_all.sh ${1:-r1p}_prof > $OUT/profile_all.log 2>&1; cat $OUT/profile_all.log
