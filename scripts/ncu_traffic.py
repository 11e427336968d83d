#!/usr/bin/env python
"""profiles/r1_traffic.json from the `ncu --set full` captures of the four bench workloads.
usage: ncu_traffic.py <dir with {dgemm,sgemm,dgemv,sgemv}_prof.ncu-rep and *_plain.log> > profiles/r1_traffic.json"""
import csv, io, json, os, subprocess, sys
d = sys.argv[1]
out = {}
for wl in ("dgemm", "sgemm", "dgemv", "sgemv"):
    rep = os.path.join(d, f"{wl}_prof.ncu-rep")
    if not os.path.exists(rep):
        continue
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                                                      capture_output=True, text=True).stdout)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    def to_bytes(key):
        u, v = rec[key]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    label = None
    plain = os.path.join(d, f"{wl}_plain.log")
    if os.path.exists(plain):
        for ln in open(plain):
            if ln.startswith("{"):
                label = json.loads(ln).get("kernel")
    u, dur = rec["gpu__time_duration.sum"]
    out[wl] = {"kernel_label": label, "kernel": rec["Kernel Name"][1][:80],
               "dram_read_bytes": to_bytes("dram__bytes_read.sum"), "dram_write_bytes": to_bytes("dram__bytes_write.sum"),
               "duration_under_ncu": f"{dur} {u}", "source": f"profiles/ (ncu --set full of `python bench.py --steps 2 --warmup 3 --no-extras --workload {wl}`, one launch)"}
print(json.dumps(out, indent=1))
