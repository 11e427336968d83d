#!/bin/bash
# ncu evidence per workload: plain run, launch list, one full capture of the dominant kernel.
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for spec in ${SPECS:-"dgemm:dgemm_tma" "sgemm:sgemm_3xtf32" "dgemv:gemv_n" "sgemv:gemv_n"}; do
  wl=${spec%%:*}; kr=${spec##*:}
  CMD="python bench.py --steps 2 --warmup 3 --no-extras --workload $wl"
  $CMD > $OUT/${wl}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${wl}_launches.csv $CMD > $OUT/${wl}_ncu_launch.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$kr -s 3 -c 1 -o $OUT/${wl}_prof $CMD > $OUT/${wl}_ncu_full.log 2>&1
  echo "$wl rc=$?"
done
