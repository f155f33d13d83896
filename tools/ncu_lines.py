"""Per-source-line instruction / stall-sample totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_lines.py file.csv [top]"""
import csv, sys, collections, os
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name",) or hdr is None: continue
    if r[0] == "": continue
    try:
        ie = int(r[hdr.index("Instructions Executed")]); sm = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    out.append((ie, sm, cur, int(r[0]), r[1].strip()))
tot = sum(o[0] for o in out); tots = sum(o[1] for o in out)
print("total warp instructions %d, samples %d" % (tot, tots))
perfile = collections.Counter()
for o in out: perfile[o[2]] += o[0]
print({k: "%.1f%%" % (100.0 * v / tot) for k, v in perfile.items()})
for ie, sm, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%5.2f%% instr %5.2f%% smpl  %s:%d  %s" % (100.0 * ie / tot, 100.0 * sm / max(tots, 1), f, ln, src[:110]))
