"""Summarise an `ncu --page source --print-source cuda --csv` export: top source lines by stall samples / instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
secs = []; cur = None; fname = None
for r in rows:
    if r and r[0] == 'File Name': fname = r[1]; continue
    if r and r[0] == 'Line No': cur = dict(file=fname, hdr=r, data=[]); secs.append(cur); continue
    if cur is not None and r: cur['data'].append(r)
def iv(x):
    try: return float(x)
    except: return 0.0
allrows = []
for s in secs:
    ix = {h: i for i, h in enumerate(s['hdr'])}
    if '# Samples' not in ix: continue
    for r in s['data']:
        if len(r) < len(s['hdr']): continue
        allrows.append((s['file'].split('/')[-1], r[0], r[1].strip()[:110], iv(r[ix['# Samples']]), iv(r[ix['Instructions Executed']]), {h: iv(r[i]) for h, i in ix.items() if h.startswith('stall_') and 'Not Issued' not in h}))
S = sum(a[3] for a in allrows); I = sum(a[4] for a in allrows)
print('total samples', S, 'instructions', I)
tot = {}
for a in allrows:
    for k, v in a[5].items(): tot[k] = tot.get(k, 0) + v
ts = sum(tot.values())
print('stalls:', ', '.join('%s %.1f%%' % (k[6:], 100 * v / ts) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]))
for a in sorted(allrows, key=lambda x: -x[3])[:top]:
    st = sorted(a[5].items(), key=lambda x: -x[1])[:2]
    print('%5.1f%% smp %5.1f%% ins  %s:%s  %s   [%s]' % (100 * a[3] / S, 100 * a[4] / I, a[0], a[1], a[2], ', '.join('%s %.0f' % (k[6:], v) for k, v in st)))
