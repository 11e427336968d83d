#!/bin/bash
# round-1 session q: Always-mode pipeline retune check + harness sweep vs the cuBLAS backend
OUT=gpurun_out/r1q; mkdir -p $OUT
export OMP_NUM_THREADS=$(nproc)
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_all.txt 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_all.txt
B200_ENGINE_TRACE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-extras > $OUT/bench_dgemm.json 2> $OUT/engine_trace.txt; cat $OUT/bench_dgemm.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'])"
grep -c gemm $OUT/engine_trace.txt; tail -4 $OUT/engine_trace.txt
timeout 1500 bash scripts/gpu_sweep.sh r1q_sweep 320 "512 1024 2048 4096 8192" > $OUT/sweep.log 2>&1; grep "rc=" $OUT/sweep.log
