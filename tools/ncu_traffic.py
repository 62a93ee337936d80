#!/usr/bin/env python
"""Writes profiles/traffic.json from an `ncu --set full` capture of the E-step and M-step kernels:
DRAM bytes (read + write) per launch. Usage: ncu_traffic.py report.ncu-rep WORKLOAD NSEQ"""
import csv, json, subprocess, sys, os, io
rep, workload, nseq = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
def scale(u): return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
res = {"workload": workload, "nseq": nseq, "source": os.path.basename(rep)}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(m); tot += float(r[i]) * scale(units[i])
    key = "estep_bytes" if "k_estep" in name else "mstep_bytes" if "k_mstep" in name else None
    if key and key not in res:
        res[key] = tot; res[key.replace("bytes", "kernel")] = name.split("(")[0]
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
print(res)
