"""Joins an `ncu --page source --csv` SASS table with `nvdisasm -g -c` line info and aggregates executed
instructions and stall samples per source line.  Usage: ncu_by_line.py <src.csv> <dis.txt> <kernel-substr> [top]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis_txt, ksub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# per-instruction line numbers of the chosen kernel, in order
lines, cur, active = [], None, False
for l in open(dis_txt):
    if l.startswith('//---') and '.text.' in l:
        active = ksub in l
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s', l):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
print('sass instructions: disasm', len(lines), 'ncu', len(body))
agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for i, r in enumerate(body):
    key = lines[i] if i < len(lines) else None
    a = agg[key]
    a[0] += int(r[ix['Instructions Executed']] or 0)
    a[1] += int(r[ix['# Samples']] or 0)
    a[2] += 1
    for s in stall_cols:
        v = int(r[ix[s]] or 0)
        if v:
            a[3][s] += v
tot_i = sum(a[0] for a in agg.values())
tot_s = sum(a[1] for a in agg.values())
print('total warp-insts', tot_i, 'samples', tot_s)
srcs = {}
def srcline(k):
    if k is None:
        return ''
    f, n = k
    if f not in srcs:
        try:
            srcs[f] = open('/root/repo/moshpp_b200/csrc/' + f).read().split('\n')
        except Exception:
            srcs[f] = []
    return srcs[f][n - 1].strip()[:90] if n - 1 < len(srcs[f]) else ''
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
    print(f'{str(k):34s} samp {100*a[1]/tot_s:5.1f}% inst {100*a[0]/tot_i:5.1f}% sass {a[2]:5d} {[(s.replace("stall_",""),v) for s,v in st]} | {srcline(k)}')

# ---- optional phase summary: NCU_PHASES="name:lo-hi,name:lo-hi" groups lines of mosh2_device.cuh
import os
ph = os.environ.get('NCU_PHASES')
if ph:
    groups = []
    for item in ph.split(','):
        name, rng = item.split(':')
        lo, hi = rng.split('-')
        groups.append((name, int(lo), int(hi)))
    tot = defaultdict(lambda: [0, 0, defaultdict(int)])
    for k, a in agg.items():
        nm = 'other'
        if k and k[0] == 'mosh2_device.cuh':
            for name, lo, hi in groups:
                if lo <= k[1] <= hi:
                    nm = name
                    break
        tot[nm][0] += a[0]; tot[nm][1] += a[1]
        for s_, v in a[3].items():
            tot[nm][2][s_] += v
    print('--- phases')
    for nm, a in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        st = sorted(a[2].items(), key=lambda kv: -kv[1])[:4]
        print(f'{nm:14s} samp {100*a[1]/tot_s:5.1f}% inst {100*a[0]/tot_i:5.1f}%  {[(s_.replace("stall_",""), round(100*v/max(1,a[1]))) for s_, v in st]}')
