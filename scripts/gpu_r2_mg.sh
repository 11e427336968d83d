#!/bin/bash
# Round-2 multi-GPU session (run under gpurun --gpus 2): team tests, the whole GPU suite, bench at
# N = 1 and N = 2 (BASELINE config 5 through torchrun).  Everything lands in gpurun_out/$TAG.
TAG=${1:-r2a}
NG=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export OMP_NUM_THREADS=$(nproc)
{
  echo "== box"; nproc; free -g | head -2
  nvidia-smi --query-gpu=index,name,memory.total,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv
  nvidia-smi topo -m
} > $OUT/box.txt 2>&1
timeout -k 10 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $OUT/pytest_multi.txt 2>&1; echo "pytest_multi rc=$?" >> $OUT/box.txt
timeout -k 10 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_multi.py > $OUT/pytest_all.txt 2>&1; echo "pytest_all rc=$?" >> $OUT/box.txt
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "smoke rc=$?" >> $OUT/box.txt
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench_n1 rc=$?" >> $OUT/box.txt
for wl in dgemm dgemv sgemm; do
  timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $NG --steps 3 --warmup 3 --workload $wl > $OUT/bench_n${NG}_$wl.json 2> $OUT/bench_n${NG}_$wl.err
  echo "bench_n${NG}_$wl rc=$?" >> $OUT/box.txt
done
B200_MG_WAIT=kernel timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $NG --steps 3 --warmup 3 > $OUT/bench_n${NG}_dgemm_waitkernel.json 2> $OUT/bench_n${NG}_dgemm_waitkernel.err
echo "bench_waitkernel rc=$?" >> $OUT/box.txt
tail -25 $OUT/pytest_multi.txt; tail -8 $OUT/pytest_all.txt; tail -3 $OUT/smoke.txt
for f in $OUT/bench_n*.json; do echo "--- $f"; cut -c1-1500 $f; done
for f in $OUT/bench_n*.err; do echo "--- $f"; tail -5 $f; done
cat $OUT/box.txt
