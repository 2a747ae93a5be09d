#!/usr/bin/env python
"""Hot SASS instructions of one kernel from `ncu --page source --csv`, joined with the source lines of
`nvdisasm -g -c` output (optional): samples, stall reasons, executions.
usage: ncu_source_hot.py source.csv [all.sass kernel-substring] [min-percent]"""
import csv, re, sys
f = sys.argv[1]
rows = list(csv.reader(open(f)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) >= 30]
H = {k: i for i, k in enumerate(hdr)}
def I(r, k):
    try: return int(float(r[H[k]] or 0))
    except Exception: return 0
off2line = {}
if len(sys.argv) > 3:
    sass = open(sys.argv[2]).read().split('\n')
    start = [i for i, l in enumerate(sass) if l.startswith('.text.') and sys.argv[3] in l][0]
    line = None
    for s in sass[start + 1:]:
        if s.strip().startswith('.section'): break
        m = re.search(r'//## File "(.*?)", line (\d+)', s)
        if m: line = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', s)
        if m: off2line[int(m.group(1), 16)] = line
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
base = int(data[0][0], 16)
tot = sum(I(r, '# Samples') for r in data)
stall = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
print('instructions', len(data), 'samples', tot, 'executed', sum(I(r, 'Instructions Executed') for r in data))
cur = None
for r in data:
    off = int(r[0], 16) - base
    fl = off2line.get(off)
    if fl and fl[0] == 'hss_engine.cu': cur = fl[1]
    s = I(r, '# Samples')
    if s > minp / 100 * tot:
        st = sorted(((k[6:], I(r, k)) for k in stall), key=lambda kv: -kv[1])[:3]
        print('%05x L%-5s %5.2f%% exec %9d  %-58s %s' % (off, cur, 100 * s / tot, I(r, 'Instructions Executed'),
              r[H['Source']].strip()[:58], ' '.join('%s=%d' % kv for kv in st)))
