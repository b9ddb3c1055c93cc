"""Attribute ncu SASS-level stall samples to CUDA source lines.
usage: ncu_sass_lines.py <ncu source-page csv (sass)> <nvdisasm -g -c output> <mangled kernel name> [top]
The ncu csv lists the kernel's SASS instructions in order; nvdisasm -g annotates the same instructions with
'//## File "...", line N' markers, so the i-th instruction of both listings is the same instruction."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
sass = open(sys.argv[2]).read().split('\n')
kern = sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
# --- nvdisasm: instruction -> (file, line)
start = next(i for i, l in enumerate(sass) if l.startswith('.text.' + kern + ':'))
lines = []
cur = ('?', 0)
for l in sass[start + 1:]:
    if l.startswith('//-----') or l.strip().startswith('.section'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): lines.append(cur)
# --- ncu rows (last kernel section)
secs = []; c = None
for r in rows:
    if r and r[0] == 'Kernel Name': c = []; secs.append(c); continue
    if c is not None: c.append(r)
sec = secs[-1]; hdr = sec[0]; data = [r for r in sec[1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
def fv(x):
    try: return float(x)
    except: return 0.0
print('ncu instructions', len(data), 'nvdisasm instructions', len(lines))
agg = {}
for i, r in enumerate(data):
    key = lines[i] if i < len(lines) else ('?', 0)
    a = agg.setdefault(key, dict(s=0.0, n=0.0, st={}))
    a['s'] += fv(r[ix['# Samples']]); a['n'] += fv(r[ix['Instructions Executed']])
    for h, j in ix.items():
        if h.startswith('stall_') and 'Not Issued' not in h: a['st'][h[6:]] = a['st'].get(h[6:], 0) + fv(r[j])
S = sum(a['s'] for a in agg.values()); I = sum(a['n'] for a in agg.values())
src = {}
for key, a in sorted(agg.items(), key=lambda x: -x[1]['s'])[:top]:
    f, ln = key
    if f not in src:
        try: src[f] = open('/root/repo/subrosadg_b200/csrc/' + f).read().split('\n')
        except Exception: src[f] = []
    text = src[f][ln - 1].strip()[:95] if 0 < ln <= len(src[f]) else ''
    st = ', '.join('%s %.0f%%' % (k, 100 * v / max(a['s'], 1)) for k, v in sorted(a['st'].items(), key=lambda x: -x[1])[:2])
    print('%5.1f%% smp %5.1f%% ins %s:%d  %s  [%s]' % (100 * a['s'] / S, 100 * a['n'] / I, f, ln, text, st))
