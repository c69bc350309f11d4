#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell path: counts of UTCHMMA (tcgen05.mma), UTMALDG (TMA loads), LDTM / STTM
(tcgen05.ld / st), UTCBAR (tcgen05.commit) and HMMA (legacy mma.sync) in every kernel of librbm_b200.so.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt      (cuobjdump -sass; no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "recommender-baseline-model_b200", "librbm_b200.so")
OPS = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "HMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*", "", cur)
            order.append(cur)
            continue
        if cur:
            for op in OPS:
                if re.search(r"\b%s\b" % op, line.split("/*")[1] if "/*" in line and line.strip().startswith("/*") else line):
                    counts[cur][op] += 1
    print("# SASS summary of %s (arch: %s)" % (os.path.relpath(LIB, ROOT), ", ".join(arch)))
    print("# %-78s %s" % ("kernel", " ".join("%8s" % o for o in OPS)))
    tot = collections.Counter()
    for k in order:
        c = counts[k]
        if any(c[o] for o in OPS):
            print("%-80s %s" % (k[:80], " ".join("%8d" % c[o] for o in OPS)))
        tot.update(c)
    print("%-80s %s" % ("TOTAL", " ".join("%8d" % tot[o] for o in OPS)))
    hm = [k for k in order if counts[k]["HMMA"]]
    print("\n# kernels still on the legacy mma.sync path (HMMA): %d" % len(hm))
    for k in hm:
        print("#   " + k[:110])


if __name__ == "__main__":
    sys.exit(main())
