#!/usr/bin/env python
"""Per-CUDA-source-line summary of `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`:
stall samples, instructions, shared-memory wavefronts (actual / ideal), global sectors, top stall reasons.
usage: python tools/ncu_src.py dump.csv [top_n] [kernel_instance] [function-name substring]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
filt = sys.argv[4] if len(sys.argv) > 4 else ""
tables, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        tables.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
lines, kidx, seen = {}, 0, set()
for t in tables:
    fn = t["rows"][0][1] if t["rows"] and t["rows"][0][0] == "Function Name" else "?"
    hdr = t["rows"][1] if len(t["rows"]) > 1 else []
    col = {h: i for i, h in enumerate(hdr)}
    if "Line No" not in col or filt not in fn:
        continue
    if (fn, t["file"]) in seen:
        kidx += 1
        seen = set()
    seen.add((fn, t["file"]))
    if kidx != want:
        continue
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in t["rows"][2:]:
        if len(r) < len(hdr) or not r[col["Line No"]].strip().isdigit():
            continue

        def num(name):
            try:
                return float(r[col[name]])
            except Exception:
                return 0.0
        k = (t["file"].split("/")[-1], int(r[col["Line No"]]))
        d = lines.setdefault(k, dict(src=r[1].strip()[:100], samples=0, inst=0, shw=0, shi=0, gl=0, gli=0, stalls={}))
        d["samples"] += num("# Samples"); d["inst"] += num("Instructions Executed")
        d["shw"] += num("L1 Wavefronts Shared"); d["shi"] += num("L1 Wavefronts Shared Ideal")
        d["gl"] += num("L2 Theoretical Sectors Global"); d["gli"] += num("L2 Theoretical Sectors Global Ideal")
        for s in stall_cols:
            d["stalls"][s] = d["stalls"].get(s, 0) + num(s)
tot = sum(d["samples"] for d in lines.values()) or 1
toti = sum(d["inst"] for d in lines.values()) or 1
totw = sum(d["shw"] for d in lines.values()) or 1
alls = {}
for d in lines.values():
    for k, v in d["stalls"].items():
        alls[k] = alls.get(k, 0) + v
print(f"samples {tot:.0f}  warp-inst {toti:.4g}  shared wavefronts {totw:.4g} (ideal {sum(d['shi'] for d in lines.values()):.4g})  "
      f"global sectors {sum(d['gl'] for d in lines.values()):.4g} (ideal {sum(d['gli'] for d in lines.values()):.4g})")
print("stalls: " + ", ".join(f"{k[6:]} {100 * v / max(sum(alls.values()), 1):.1f}%" for k, v in sorted(alls.items(), key=lambda kv: -kv[1])[:8]))
for k, d in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:3]
    ss = " ".join(f"{n[6:]}:{v:.0f}" for n, v in st if v > 0)
    print(f"{k[0]}:{k[1]:<4d} {100 * d['samples'] / tot:5.1f}% inst {100 * d['inst'] / toti:4.1f}% shw {100 * d['shw'] / totw:4.1f}% (x{d['shw'] / max(d['shi'], 1):.1f}) "
          f"gsec {d['gl']:.2g}/{d['gli']:.2g} | {ss} | {d['src']}")
