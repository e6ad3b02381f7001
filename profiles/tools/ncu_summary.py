#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline metrics, stall reasons, instruction mix per tile-frame, hottest SASS lines.
   ncu_summary.py report.ncu-rep TILES FRAMES"""
import csv, collections, subprocess, sys, io
rep, tiles, frames = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
def page(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout
det = page("--page", "details")
for line in det.splitlines():
    if any(k in line for k in ("Duration", "Executed Ipc Active", "Issue Slots Busy", "No Eligible", "Eligible Warps", "Registers Per", "DRAM Throughput", "Warp Cycles Per Issued", "L2 Hit", "L1/TEX Hit", "Dynamic Shared", "Achieved Occupancy")):
        print(line.rstrip())
rows = list(csv.reader(io.StringIO(page("--page", "raw", "--csv"))))
d = dict(zip(rows[0], rows[2]))
st = {}
for k, v in d.items():
    if 'issue_stalled' in k and k.endswith('.ratio'):
        try:
            if float(v) > 0.15:
                st[k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')] = round(float(v), 2)
        except ValueError:
            pass
print("stalls (warp cycles per issue):", st)
print("dram read / write bytes:", d.get('dram__bytes_read.sum'), d.get('dram__bytes_write.sum'))
rows = list(csv.reader(io.StringIO(page("--page", "source", "--csv", "--print-source", "sass"))))
hdr = rows[1]; isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
data = rows[2:]
cnt = collections.Counter()
for r in data:
    op = r[isrc].strip().split()
    if not op:
        continue
    o = op[1] if op[0].startswith('@') else op[0]
    cnt[o.split('.')[0]] += int(r[iex])
tf = tiles * frames
print("instructions per tile-frame:", {o: round(c / tf, 1) for o, c in cnt.most_common(24)})
print("total per tile-frame", round(sum(cnt.values()) / tf, 1))
tot = sum(int(r[isamp]) for r in data)
print("samples", tot)
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:int(sys.argv[4]) if len(sys.argv) > 4 else 24]
for i in sorted(top):
    r = data[i]
    print(i, r[isrc].strip()[:90], r[isamp], r[iex])
