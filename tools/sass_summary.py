#!/usr/bin/env python
"""profiles/rNN_sass_mnemonics.txt: per-kernel counts of the SASS mnemonics that prove which
hardware paths libllc.so uses (cuobjdump -sass; runs without a GPU).
    python tools/sass_summary.py profiles/r02_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lifelong-clip_b200", "libllc.so")
PAT = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|SYNCS|HMMA\.\d+|"
                 r"MUFU\.\w+|ELECT|NANOSLEEP)")


def main(out_path):
    sass = subprocess.check_output(["cuobjdump", "-sass", LIB], text=True)
    cur, counts = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur:
            for op in PAT.findall(line):
                counts[cur][op] += 1
    names = subprocess.check_output(["c++filt"], input="\n".join(counts), text=True).splitlines()
    out = ["# SASS mnemonic counts per kernel of lifelong-clip_b200/libllc.so (cuobjdump -sass, "
           "sm_100a build of this tree)",
           "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st (TMEM), "
           "UTMALDG/UTMASTG/UTMAPF = TMA load/store/prefetch,",
           "# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (only the "
           "L > 256 attention fallback)", ""]
    tot = collections.Counter()
    for (fn, c), name in sorted(zip(counts.items(), names), key=lambda kv: -sum(kv[0][1].values())):
        name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
        out += [name, "    " + "  ".join(f"{k}={v}" for k, v in sorted(c.items()))]
        tot.update(c)
    out += ["", "TOTAL  " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items()))]
    with open(out_path, "w") as f:
        f.write("\n".join(out) + "\n")
    print(out[-1])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles",
                                                            "r02_sass_mnemonics.txt"))
