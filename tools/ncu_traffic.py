#!/usr/bin/env python
"""Writes profiles/traffic.json from an `ncu --set full` capture of ONE iteration's E-step and M-step kernels:
DRAM bytes (read + write) per iteration, summed over the kernels of each phase (the pruned E-step is three kernels:
k_emasked, k_ebound, k_eexact; the dense one k_estep_packed). Also prints a per-kernel table (duration, DRAM bytes, shared
wavefronts, issue / LSU utilisation). Usage: ncu_traffic.py report.ncu-rep WORKLOAD NSEQ [summary.csv]"""
import csv, json, subprocess, sys, os, io
rep, workload, nseq = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
def scale(u): return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
def val(r, m):
    i = hdr.index(m); return float(r[i].replace(",", "")) * scale(units[i])
res = {"workload": workload, "nseq": nseq, "source": os.path.basename(rep), "estep_bytes": 0.0, "mstep_bytes": 0.0, "kernels": {}}
COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
table = [["kernel"] + COLS]
seen = set()
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
    if name in seen:
        continue                                   # first launch of every kernel = one iteration
    seen.add(name)
    tot = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    phase = "estep" if any(k in name for k in ("k_estep", "k_ebound", "k_eexact", "k_emasked")) else "mstep" if "k_mstep" in name else None
    if phase:
        res[phase + "_bytes"] += tot
        res["kernels"][name] = {"dram_bytes": tot, "ms": val(r, "gpu__time_duration.sum") / (1e6 if units[hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond") else 1.0)}
    table.append([name] + [r[hdr.index(c)] if c in hdr else "" for c in COLS])
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
if len(sys.argv) > 4:
    with open(sys.argv[4], "w", newline="") as f:
        csv.writer(f).writerows(table)
