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
