#!/bin/bash
# GPU-side kernel durations (ncu launch list) for small problems, where a python/ctypes
# timing loop is host-bound.  usage: kernel_durations.sh <tag> <kind> "<sizes>" [ENV=val ...]
TAG=$1; KIND=$2; SIZES=$3; shift 3
OUT=gpurun_out/$TAG; mkdir -p $OUT
for cfg in "$@"; do
  env $cfg python scripts/size_sweep.py $KIND $SIZES > $OUT/plain_$cfg.log 2>&1 &&
  env $cfg ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/dur_$cfg.csv python scripts/size_sweep.py $KIND $SIZES > /dev/null 2>&1
  python - "$OUT/dur_$cfg.csv" "$cfg" "$SIZES" <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit() and ('b200::' in r[4])]
sizes=sys.argv[3].split()
# 25 launches per size (5 warm-up + 20 timed), possibly several kernels per call
per=len(rows)//len(sizes)
out=[]
for i,s in enumerate(sizes):
    chunk=rows[i*per:(i+1)*per]
    tail=chunk[len(chunk)//2:]
    calls=20*len(tail)/ (len(chunk)*0.8) if chunk else 1
    tot=sum(float(r[-1]) for r in tail)/1e3
    ncalls=len(tail)/(per/25)
    out.append(f"{s}:{tot/ncalls:.2f}us")
print(sys.argv[2] or "default", " ".join(out))
PY
done
