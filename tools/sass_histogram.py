"""SASS opcode histogram per kernel of the built libafb200.so (the evidence the profiling recipe asks for:
UTC*MMA = tcgen05.mma, UTMALDG/UTMASTG = TMA, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit ...).

  python tools/sass_histogram.py [path/to/libafb200.so] > profiles/rNN_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "spatiotemporal-deepfake-detection-for-live-video-calls_b200", "libafb200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCMMA", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOM", "SYNCS",
       "HMMA", "LDGSTS", "RED", "ATOM", "FFMA", "LDS", "STS", "LDG", "STG", "ELECT", "ACQBULK", "UCGABAR", "CCTL")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            kernels[cur][m.group(1) + m.group(2)] += 1
    names = demangle(list(kernels))
    print("SASS opcode histogram of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(LIB, ROOT))
    total = collections.Counter()
    for k, c in kernels.items():
        base = collections.Counter()
        for op, n in c.items():
            base[op.split(".")[0]] += n
        total.update(base)
        short = names.get(k, k).replace("(anonymous namespace)::", "")
        short = re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", short)
        print("\n== %s   (%d instructions)" % (short, sum(c.values())))
        print("   " + "  ".join("%s=%d" % (op, base[op]) for op in KEY if base.get(op)))
        full = [(op, n) for op, n in sorted(c.items()) if op.split(".")[0] in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "RED", "UBLKCP")]
        if full:
            print("   " + "  ".join("%s x%d" % (op, n) for op, n in full))
    print("\n== whole library: " + "  ".join("%s=%d" % (op, total[op]) for op in KEY if total.get(op)))


if __name__ == "__main__":
    main()
