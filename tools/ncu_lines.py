#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` (SASS) dump by CUDA source line using nvdisasm -g line info (dev tool).
usage: ncu_lines.py <src.csv> <cubin> <mangled kernel name> [top]"""
import csv, re, subprocess, sys, collections
src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(['nvdisasm', '-g', cubin], capture_output=True, text=True).stdout
m0 = re.search(r'\.section\s+\.text\.' + re.escape(kname) + r'\b', dis)
sec = dis[m0.end():]
nxt = sec.find('.section')
if nxt > 0: sec = sec[:nxt]
line_of = {}
cur = None
for l in sec.splitlines():
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m: cur = int(m.group(1)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);', l)
    if m: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ai = hdr.index('Address'); si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
wi = hdr.index('L1 Wavefronts Shared'); wx = hdr.index('L1 Wavefronts Shared Excessive')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = int(rows[2][ai], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0, collections.Counter()])
tot = 0
for r in rows[2:]:
    off = int(r[ai], 16) - base
    ln = line_of.get(off)
    a = agg[ln]
    s = int(r[si] or 0); a[0] += s; tot += s
    a[1] += int(r[ii] or 0); a[2] += int(r[wi] or 0); a[3] += int(r[wx] or 0)
    for i in stalls:
        v = int(r[i] or 0)
        if v: a[4][hdr[i]] += v
print('total samples', tot)
import os
srcl = open('/root/repo/poissonrecon_gpu_b200/csrc/' + os.path.basename(cubin).split('.')[0] + '.cu').read().splitlines()
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ', '.join(f'{k[6:]}={v}' for k, v in a[4].most_common(3))
    text = srcl[ln - 1].strip()[:70] if (srcl and ln) else ''
    print(f'{ln!s:>5} {100*a[0]/tot:5.1f}% inst={a[1]:>11} smemwf={a[2]:>11} exc={a[3]:>10} | {st} | {text}')
