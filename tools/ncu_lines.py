#!/usr/bin/env python
"""Per-source-line warp-sample table of one kernel in an .ncu-rep (captured with --import-source on; the object file
must come from the same build).   python tools/ncu_lines.py report.ncu-rep object.o [top]"""
import collections, csv, io, re, subprocess, sys, tempfile, os, glob

rep, obj = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, rows = rows[1], rows[2:]
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cub = glob.glob(os.path.join(tmp, "*.cubin"))[0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
cur, ins = None, []
for l in dis.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((cur, m.group(2)))
assert len(ins) == len(rows), (len(ins), len(rows))
cols = {k: h.index(k) for k in ("# Samples", "Instructions Executed", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_selected", "stall_not_selected", "stall_no_inst", "stall_branch_resolving")}
agg = collections.defaultdict(lambda: collections.Counter())
for (loc, _), r in zip(ins, rows):
    for k, i in cols.items():
        agg[loc][k] += int(r[i])
tot = sum(a["# Samples"] for a in agg.values())
print("total samples", tot)
print("%-28s %7s %6s %9s  long_sb wait short math sel notsel noinst branch" % ("line", "samples", "%", "executed"))
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    print("%-28s %7d %5.1f%% %9d  %6d %5d %5d %4d %4d %5d %5d %5d" % ("%s:%d" % loc if loc else "?", a["# Samples"], 100.0 * a["# Samples"] / tot, a["Instructions Executed"],
          a["stall_long_sb"], a["stall_wait"], a["stall_short_sb"], a["stall_math"], a["stall_selected"], a["stall_not_selected"], a["stall_no_inst"], a["stall_branch_resolving"]))

# call sites of the hottest inlined line (e.g. a wait helper): nearest following distinct source line
if len(sys.argv) > 4:
    hot = int(sys.argv[4])
    print("\ninstances of line", hot)
    for i, (loc, txt) in enumerate(ins):
        n = int(rows[i][cols["# Samples"]])
        if loc and loc[1] == hot and n > 40:
            nxt = [ins[j][0][1] for j in range(i + 1, min(len(ins), i + 80)) if ins[j][0] and ins[j][0][0] == loc[0] and abs(ins[j][0][1] - hot) > 8]
            prv = [ins[j][0][1] for j in range(max(0, i - 80), i) if ins[j][0] and ins[j][0][0] == loc[0] and abs(ins[j][0][1] - hot) > 8]
            print("  sass %5d samples %5d executed %8s  prev %s next %s" % (i, n, rows[i][cols["Instructions Executed"]], prv[-2:], nxt[:2]))
