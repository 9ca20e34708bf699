#!/usr/bin/env python
"""Turn one `ncu --set full` report holding the hot kernels of a substep into the tracked artifacts under profiles/:
  profiles/<tag>_hot_kernels.md                      one row per captured launch (duration, DRAM bytes, occupancy, issue, hit rates, stalls)
  profiles/<tag>_ncu_full_<kernel>.details.csv       the details page of one representative launch per kernel
  profiles/ncu_traffic.json                          dram__bytes_read.sum + dram__bytes_write.sum per launch (mean), read by bench.py
usage (here, no GPU needed): python tools/profiles_from_rep.py gpurun_out/x_hot.ncu-rep r02"""
import csv, io, json, re, subprocess, sys, collections

rep, tag = sys.argv[1], sys.argv[2]
def page(p, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", p, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))
raw = page("raw")
h, units, data = raw[0], raw[1], raw[2:]
col = {n: i for i, n in enumerate(h)}
def short(n):
    n = re.sub(r"\(.*", "", n); n = n.replace("void ", "")
    return n.split("::")[-1].split("<")[0]
want = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "DRAM rd MB", 1.0), ("dram__bytes_write.sum", "DRAM wr MB", 1.0),
        ("launch__registers_per_thread", "regs", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %", 1),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", 1), ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1),
        ("smsp__issue_inst0.avg.pct_of_peak_sustained_active", "no-issue %", 1)]
def unit_scale(name):
    u = units[col[name]]
    return {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
lines = [f"# ncu --set full, hot kernels of one substep ({rep}); per-launch values, cold-cache and serialised by the profiler\n",
         "| id | kernel | grid x block | " + " | ".join(w[1] for w in want) + " | top stalls (pc samples) |", "|---" * (4 + len(want)) + "|"]
stall_cols = [n for n in h if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued")]
traffic = collections.defaultdict(list); rep_id = {}
for r in data:
    name = short(r[col["Kernel Name"]])
    vals = []
    for n, label, _ in want:
        if n not in col: vals.append("-"); continue
        v = float(r[col[n]].replace(",", "") or 0)
        if "bytes" in n or "duration" in n: v *= unit_scale(n)
        vals.append(f"{v:.2f}" if v < 1000 else f"{v:.0f}")
    st = sorted(((float(r[col[n]].replace(",", "") or 0), n.replace("smsp__pcsamp_warps_issue_stalled_", "")) for n in stall_cols), reverse=True)
    tot = sum(s for s, _ in st) or 1
    stalls = ", ".join(f"{n} {100*s/tot:.0f}%" for s, n in st[:4])
    lines.append(f"| {r[col['ID']]} | {name} | {r[col['launch__grid_size']]} x {r[col['launch__block_size']]} | " + " | ".join(vals) + f" | {stalls} |")
    b = (float(r[col["dram__bytes_read.sum"]].replace(",", "")) * unit_scale("dram__bytes_read.sum") + float(r[col["dram__bytes_write.sum"]].replace(",", "")) * unit_scale("dram__bytes_write.sum")) * 1e6
    traffic[name].append(b); rep_id.setdefault(name, r[col["ID"]])
open(f"profiles/{tag}_hot_kernels.md", "w").write("\n".join(lines) + "\n")
det = page("details")
dh = det[0]; idc = dh.index("ID")
for name, i in rep_id.items():
    with open(f"profiles/{tag}_ncu_full_{name}.details.csv", "w", newline="") as f:
        w = csv.writer(f); w.writerow(dh)
        for r in det[1:]:
            if r[idc] == i: w.writerow(r)
keymap = {"mg_cycle_kernel": "mg_cycle_upper", "mg_cluster_kernel": "mg_cluster", "g2p_tile_kernel": "g2p_advect", "p2g_gather_kernel": "p2g_gather", "p2g_xrow_kernel": "p2g_gather"}
tj = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the launches of one substep) from the `ncu --set full` capture named in each entry, bench workload (512^3 tank, 16.8 M particles); bench.py copies the matching entry into roofline.traffic"}
for name, bs in traffic.items():
    tj[keymap.get(name, name)] = {"bytes": int(sum(bs) / len(bs)), "launches": len(bs), "capture": f"profiles/{tag}_ncu_full_{name}.details.csv"}
if "mg_cycle_upper" in tj: tj["mg_cycle"] = tj["mg_cycle_upper"]
json.dump(tj, open("profiles/ncu_traffic.json", "w"), indent=1)
print(open(f"profiles/{tag}_hot_kernels.md").read())
