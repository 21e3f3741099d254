"""Top stall-sample instructions of one kernel from an ncu source-page CSV:
    ncu -i X.ncu-rep --page source --csv --print-source sass > /tmp/src.csv ; python scripts/ncu_hot.py /tmp/src.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[idx['# Samples']]) for r in data)
print('kernel', rows[0][1][:100], '| total samples', tot, '| instrs', len(data))
for i, r in enumerate(data):
    s = int(r[idx['# Samples']])
    if s > tot * thr / 100:
        print(f"{i:5d} {s:6d} {100 * s / tot:5.1f}%  {r[idx['Source']].strip()[:88]:88s} exec {r[idx['Instructions Executed']]:>9s} shwf {r[idx['L1 Wavefronts Shared']]:>9s}/{r[idx['L1 Wavefronts Shared Ideal']]}")
