"""Aggregate `ncu --page source --csv --print-source sass` output by opcode (stall samples)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break  # further launches of the same kernel follow: keep the first
    if len(r) == len(hdr) and r[ix['# Samples']].isdigit():
        data.append(r)
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('instructions', len(data), 'total samples', tot)
agg = collections.Counter(); cnt = collections.Counter(); ex = collections.Counter()
stall = collections.defaultdict(collections.Counter)
KEYS = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
for r in data:
    src = r[ix['Source']].split()
    op = src[1] if src[0].startswith('@') else src[0]
    s = int(r[ix['# Samples']] or 0)
    agg[op] += s; cnt[op] += 1; ex[op] += int(r[ix['Instructions Executed']] or 0)
    for k in KEYS:
        stall[op][k] += int(r[ix[k]] or 0)
for op, s in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
    print(f"{op:24s} samples {s:7d} ({100*s/tot:5.1f}%) static {cnt[op]:5d} exec {ex[op]:11d}",
          dict((k[6:], v) for k, v in stall[op].most_common(4)))
print()
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
    print(r[ix['Address']][-6:], r[ix['# Samples']].rjust(6), r[ix['Source']][:100],
          dict((k[6:], int(r[ix[k]] or 0)) for k in KEYS if int(r[ix[k]] or 0) > 0))
