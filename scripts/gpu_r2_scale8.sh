#!/bin/bash
# Round-2 scaling session on an 8-GPU box: BASELINE config 5 (32768^3 strong) at N = 8 and 4 through
# torchrun, SGEMM / DGEMV at N = 8, and the unmodified harness with B200_NGPUS=8.
TAG=${1:-r2d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export OMP_NUM_THREADS=$(nproc)
{ echo "== box"; nproc; free -g | head -2; nvidia-smi topo -m | head -12; } > $OUT/box.txt 2>&1
run() {  # n workload steps extra...
  local n=$1 wl=$2 st=$3; shift 3
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $n --steps $st --warmup 3 --workload $wl "$@" > $OUT/bench_n${n}_$wl.json 2> $OUT/bench_n${n}_$wl.err
  echo "bench n=$n $wl rc=$?" >> $OUT/box.txt
}
run 8 dgemm 5
run 4 dgemm 3
run 8 sgemm 5
run 8 dgemv 10
BLOB=$PWD/gpu-blas-offload-benchmark_b200/lib/gpu-blob-b200
for d in 8192 16384; do
  mkdir -p $OUT/harness_ng8_$d
  ( cd $OUT/harness_ng8_$d && B200_NGPUS=8 timeout -k 10 400 $BLOB -i 10 -s $d -d $d --no_cpu -o csv > log.txt 2>&1 ); echo "harness ng=8 d=$d rc=$?" >> $OUT/box.txt
done
for f in $OUT/bench_n*.json; do echo "--- $f"; grep '^{' $f | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'), 'parity', d['parity']['max_rel_err_sampled_blocks'], d['parity']['ok'], 'frac', round(d['roofline']['frac'],4), d['clocks'])
"; tail -2 ${f%.json}.err; done
for d in 8192 16384; do echo "== ng=8 d=$d"; grep -h "gpu" $OUT/harness_ng8_$d/csv/dgemm_square_square_M=N=K.csv $OUT/harness_ng8_$d/csv/sgemm_square_square_M=N=K.csv $OUT/harness_ng8_$d/csv/dgemv_square_vector_M=N.csv; tail -3 $OUT/harness_ng8_$d/log.txt; done
cat $OUT/box.txt
