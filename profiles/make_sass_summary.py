#!/usr/bin/env python
"""Regenerates profiles/r02_sass_summary.txt: per kernel of the shipped library, how many SASS instructions of
each Blackwell-specific kind it contains (`cuobjdump -sass csrc/libacas2d_b200.so`).

    python profiles/make_sass_summary.py

UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS = mbarrier arrive / try_wait, UTCHMMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, LDGSTS = cp.async, DFMA/DMUL/DADD = the float64 flag chain,
MUFU = special-function unit, REDG/ATOMG = the episode counters, LD/ST .SYS and MEMBAR.SYS = the peer-memory exchange.
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gym-acas2d_b200", "csrc", "libacas2d_b200.so")
KINDS = ["UBLKCP", "SYNCS", "UTCHMMA", "LDTM", "STTM", "UTCBAR", "LDGSTS", "LDS", "STS", "SHFL", "DFMA", "DMUL", "DADD",
         "MUFU", "F2F", "ATOMG", "REDG", "MEMBAR", "STG", "LDG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("acas2d::", "")
            name = re.sub(r"^void ", "", re.sub(r"\(.*", "", name))
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            per[name]["total"] += 1
            for k in KINDS:
                if op == k or op.startswith(k + "."):
                    per[name][k] += 1
    out = [f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} -- instruction counts per kernel (static SASS, sm_100a)"]
    out += ["# " + ln for ln in __doc__.strip().splitlines()[4:]] + [""]
    head = f"{'kernel':<58}{'total':>7}" + "".join(f"{k:>8}" for k in KINDS)
    out.append(head)
    tot = collections.Counter()
    for fn, c in per.items():
        out.append(f"{fn[:57]:<58}{c['total']:>7}" + "".join(f"{c[k]:>8}" for k in KINDS))
        tot.update(c)
    out.append(f"{'ALL KERNELS':<58}{tot['total']:>7}" + "".join(f"{tot[k]:>8}" for k in KINDS))
    dst = os.path.join(ROOT, "profiles", "r02_sass_summary.txt")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[-1:]))
    print("wrote", dst)


if __name__ == "__main__":
    main()
