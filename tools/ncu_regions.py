"""SASS-level attribution from `ncu --page source --csv --print-source cuda,sass`: each SASS instruction counted once,
summed per source line; the row loop (highest execution count) reported separately.
usage: python tools/ncu_regions.py both.csv [top]"""
import csv, sys, collections, os
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur = None; hdr = None; line = None
agg = collections.Counter(); cnt = collections.Counter(); srcs = {}; seen = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0] != "":
        line = (cur, int(r[0])); srcs[line] = r[1].strip(); continue
    a = r[2]
    if not a.startswith("0x") or a in seen: continue
    v = int(r[ie]) if r[ie].isdigit() else 0
    seen[a] = (v, line, r[3])
tot = sum(v for v, _, _ in seen.values())
mx = max(v for v, _, _ in seen.values())
loop = [x for x in seen.values() if x[0] >= 0.9 * mx]
print("SASS instructions %d, executed %.3f G warp-instr; row loop: %d instr x %.2f M = %.1f%%" % (
    len(seen), tot / 1e9, len(loop), mx / 1e6, 100.0 * sum(x[0] for x in loop) / tot))
print("per row-loop trip: total/mx = %.1f warp-instr (=> x35 trips per evaluation at 5 views)" % (tot / mx))
for v, l, t in seen.values():
    if v < 0.9 * mx: agg[l] += v; cnt[l] += 1
print("outside the row loop: %.1f%%" % (100.0 * sum(agg.values()) / tot))
for l, v in agg.most_common(top):
    print("%5.2f%% %4d sass  %s:%d  %s" % (100 * v / tot, cnt[l], l[0], l[1], srcs[l][:100]))
