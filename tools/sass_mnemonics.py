#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that identify the hardware paths in use (FP64 FMA, FP64 tensor-core MMA,
bulk-TMA copies + mbarrier, named barriers, shared/global atomics, popcount): cuobjdump -sass of the built library ->
profiles/<round>_sass_mnemonics.txt.  Usage: python tools/sass_mnemonics.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = ("DFMA", "DMMA", "UBLKCP", "SYNCS", "BAR", "ATOMS", "ATOMG", "ATOM", "REDG", "RED", "POPC", "SHFL", "LDS", "LDG", "MUFU.RSQ64H")


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
    so = os.path.join(ROOT, "okvis_b200", "csrc", "libokvis_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    counts = collections.defaultdict(collections.Counter)
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.x]*)", line)
        if not (m and fn):
            continue
        op = m.group(1)
        for c in CLASSES:
            if op == c or op.startswith(c + "."):
                counts[fn][c + ("" if c != "BAR" else "")] += 1
                if c == "BAR" and not op.startswith("BAR.SYNC.DEFER_BLOCKING"):
                    counts[fn]["BAR(named: " + op + ")"] += 1
                break
    out = os.path.join(ROOT, "profiles", rnd + "_sass_mnemonics.txt")
    with open(out, "w") as f:
        for fn in sorted(counts):
            for c in sorted(counts[fn]):
                f.write("%s %s %d\n" % (fn, c, counts[fn][c]))
    tot = collections.Counter()
    for fn in counts:
        tot.update(counts[fn])
    print(out, dict(tot))


if __name__ == "__main__":
    main()
