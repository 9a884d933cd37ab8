#!/usr/bin/env python
"""Development tool: fold an `ncu --page source --csv` SASS listing onto source lines using nvdisasm's line table.
    python tools/ncu_by_line.py <report.ncu-rep> <kernel-substring> [lib.so]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, pat = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym-acas2d_b200/csrc/libacas2d_b200.so")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
kname = rows[0][1]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = [r for r in rows[2:] if len(r) == len(hdr)]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
# locate the function: mangled name contains template args; match via c++filt
secs = re.split(r"\n//-+ \.text\.", sass)
best = None
for s in secs[1:]:
    mangled = s.split(" ", 1)[0]
    dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    def norm(x):
        x = re.sub(r"\((int|bool)\)", "", x).replace(" ", "").replace("void", "")
        x = x.replace("<unnamed>::", "").replace("(anonymousnamespace)::", "").replace("acas2d::", "")
        x = x.replace("false", "0").replace("true", "1")
        return x.split("(DevParams")[0]
    if norm(dem) == norm(kname):
        best = s
        break
if best is None:
    print("kernel not found in", lib, "for", kname); sys.exit(1)
lines = []
cur = ("?", 0)
for ln in best.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        lines.append(cur)
if len(lines) != len(ins):
    print(f"warning: {len(lines)} SASS instructions in the library vs {len(ins)} in the report (rebuilt since?)")
agg = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), r in zip(lines, ins):
    a = agg[(f, l)]
    a[0] += int(r[ix["# Samples"]] or 0)
    a[1] += int(r[ix["Instructions Executed"]] or 0)
    a[2] += 1
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print(f"{kname[:100]}\n total samples {tot_s}, warp instructions {tot_i}")
print(f"{'file:line':34s}{'samples%':>9s}{'instr%':>8s}{'instr':>11s}{'sass':>6s}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[: int(os.environ.get("TOP", "45"))]:
    print(f"{f + ':' + str(l):34s}{100 * a[0] / max(tot_s, 1):9.1f}{100 * a[1] / max(tot_i, 1):8.1f}{a[1]:11d}{a[2]:6d}")
