"""Summarise an ncu --page source CSV: per-instruction stall samples of the hottest region (helper, not a test)."""
import csv, sys
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60
rows = list(csv.reader(open(path)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]; end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
ia, isrc, iss, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
data = []
for r in rows[hi + 1:end]:
    if len(r) > iex and r[ia].startswith('0x'):
        try: data.append((int(r[ia], 16), r[isrc].strip(), int(r[iss] or 0), int(r[iex] or 0)))
        except Exception: pass
base = data[0][0]
tot = sum(d[2] for d in data)
print('total samples', tot, 'instructions', len(data))
mode = sys.argv[3] if len(sys.argv) > 3 else 'range'
if mode == 'top':
    for d in sorted(data, key=lambda d: -d[2])[:top]:
        print(f"{d[0]-base:6x} {d[2]:7d} {100*d[2]/tot:5.1f}% exec {d[3]:9d}  {d[1][:70]}")
else:
    lo, hi_ = int(sys.argv[4], 16), int(sys.argv[5], 16)
    acc = 0
    for d in data:
        if lo <= d[0] - base <= hi_:
            acc += d[2]
            print(f"{d[0]-base:6x} {d[2]:7d} exec {d[3]:9d}  {d[1][:70]}")
    print('range samples', acc, f'{100*acc/tot:.1f}%')
