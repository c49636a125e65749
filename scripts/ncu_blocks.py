"""summarises an `ncu --page source --csv` dump: blocks of SASS with equal execution counts"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hdr = rows[1]
ia = hdr.index('Instructions Executed'); it = hdr.index('Thread Instructions Executed'); isrc = hdr.index('Source'); iss = hdr.index('# Samples')
data = [(int(r[ia]), int(r[it]), r[isrc].strip(), int(r[iss])) for r in rows[2:] if len(r) > it]
tot = sum(d[0] for d in data); ts = sum(d[3] for d in data)
print('total warp inst', tot, 'n sass', len(data), 'samples', ts)
blocks = []; cur = None
for i, d in enumerate(data):
    if cur and abs(d[0] - cur['c']) <= 0.02 * max(cur['c'], 1):
        cur['n'] += 1; cur['w'] += d[0]; cur['t'] += d[1]; cur['s'] += d[3]; cur['end'] = i
    else:
        cur = {'start': i, 'end': i, 'c': d[0], 'n': 1, 'w': d[0], 't': d[1], 's': d[3]}; blocks.append(cur)
for b in blocks:
    if b['w'] > tot * thr or b['s'] > ts * thr:
        print(f"{b['start']:5d}-{b['end']:5d} n={b['n']:4d} exec={b['c']:10d} inst={b['w']/tot*100:5.1f}% act={b['t']/max(b['w'],1):5.1f} samples={b['s']/ts*100:5.1f}%  {data[b['start']][2][:60]}")
