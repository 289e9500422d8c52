"""Top stall sites of one kernel from `ncu --page source --csv` (SASS view): samples per instruction with the dominant reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
print('total samples', tot)
top = sorted(enumerate(data), key=lambda ir: -int(ir[1][idx['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i, r in sorted(top):
    s = int(r[idx['# Samples']] or 0)
    reasons = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {100.0 * s / tot:5.1f}%  {r[idx['Source']].strip()[:70]:70s} {' '.join(f'{n}:{c}' for c, n in [(c, n) for n, c in reasons] if n)}")
