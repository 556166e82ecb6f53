"""Summarise an `ncu --page source --csv` dump: top instructions by stall samples, grouped stall reasons."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
samp = ix['# Samples']; src = ix['Source']; ex = ix['Instructions Executed']
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
tot = sum(int(r[samp] or 0) for r in body)
print('total samples', tot, 'instructions', len(body), 'inst executed', sum(int(r[ex] or 0) for r in body))
agg = collections.Counter()
for r in body:
    for i in stall_cols:
        try: agg[hdr[i]] += int(r[i] or 0)
        except ValueError: pass
print({k: v for k, v in agg.most_common(12)})
# opcode histogram by executed count
ops = collections.Counter(); ops_s = collections.Counter()
for r in body:
    op = r[src].split()[0] if not r[src].strip().startswith('@') else r[src].split()[1]
    op = op.split('.')[0]
    ops[op] += int(r[ex] or 0); ops_s[op] += int(r[samp] or 0)
print('opcode: executed / samples')
for k, v in ops.most_common(25): print(f'  {k:10s} {v:10d} {ops_s[k]:6d}')
order = sorted(range(len(body)), key=lambda i: -int(body[i][samp] or 0))[:top]
for i in sorted(order):
    r = body[i]
    st = {hdr[c][6:]: int(r[c]) for c in stall_cols if r[c] not in ('', '0')}
    print(i, r[samp], r[ex], r[src].strip()[:70], st)
