#!/usr/bin/env python
"""Development tool: ncu_by_line folded further into source regions given as  name=file:lo-hi  arguments."""
import os, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    name, rest = a.split("=")
    f, rng = rest.split(":")
    lo, hi = rng.split("-")
    regions.append((name, f, int(lo), int(hi)))
env = dict(os.environ, TOP="100000")
out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_by_line.py"), rep, pat], capture_output=True, text=True, env=env).stdout
b = collections.defaultdict(lambda: [0.0, 0.0, 0])
for ln in out.splitlines()[3:]:
    f, s, i, n, k = ln.split()
    fn, l = f.rsplit(":", 1); l = int(l)
    key = fn
    for name, rf, lo, hi in regions:
        if fn == rf and lo <= l <= hi:
            key = name; break
    x = b[key]; x[0] += float(s); x[1] += float(i); x[2] += int(n)
print(out.splitlines()[1])
for k, v in sorted(b.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} samples {v[0]:5.1f}%  instr {v[1]:5.1f}%  {v[2]}")
