#!/bin/bash
# ncu evidence for the bench command: launch list + one full capture of the
# dominant kernel.  Usage: scripts/gpu_profile.sh <tag> <kernel-regex> [bench args...]
TAG=$1; KREGEX=$2; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-extras $*"
$CMD > $OUT/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launch.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 1 -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
echo "rc=$?"; tail -3 $OUT/plain.log | cut -c1-400; grep -v "^==PROF" $OUT/launches.csv | tail -12
