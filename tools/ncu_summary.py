"""Print the headline metrics and the hottest SASS lines of one kernel from an ncu report (run here, no GPU needed)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(det)))
h = rows[0]
keys = ['Duration', 'Registers Per Thread', 'Achieved Occupancy', 'Theoretical Occupancy', 'Issue Slots Busy', 'Executed Ipc Active', 'No Eligible',
        'Bank Conflicts', 'DRAM Throughput', 'L1/TEX Hit Rate', 'L2 Hit Rate', 'Warp Cycles Per Issued Instruction', 'Block Limit', 'Dynamic Shared Memory Per Block',
        'Compute (SM) Throughput', 'Memory Throughput', 'Eligible Warps', 'Issued Warp', 'Active Warps', 'Mem Pipes Busy', 'Max Bandwidth', 'Mem Busy', 'Shared Memory']
for r in rows[1:]:
    d = dict(zip(h, r))
    n = d.get('Metric Name', '')
    if any(k in n for k in keys):
        print(f"{n:55s} {d.get('Metric Value'):>14s} {d.get('Metric Unit')}")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
if len(rr) > 2:
    hh = rr[0]
    for name in hh:
        if any(k in name for k in ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__average_warps_issue_stalled', 'smsp__pcsamp_warps_issue_stalled']):
            if name.endswith('_ratio') or 'pct' in name or name.endswith('.sum') :
                v = rr[2][hh.index(name)]
                try:
                    if float(v.replace(',', '')) != 0: print(f"{name:80s} {v}")
                except ValueError:
                    pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ia = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "warp-instructions executed", sum(int(r[iex]) for r in data))
stall_cols = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
agg = {hdr[j]: sum(int(r[j] or 0) for r in data) for j in stall_cols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in top:
    r = data[i]
    st = {hdr[j]: int(r[j]) for j in stall_cols if r[j] not in ('', '0')}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:5d} {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}% exec {r[iex]:>9s}  {r[ia].strip()[:58]:58s} {st}")
