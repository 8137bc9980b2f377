"""One-screen SASS census of libhctr_b200.so: per kernel, how many tcgen05 MMA (UTCHMMA, `.2CTA` = cta_group::2), TMA tensor
loads (UTMALDG), TMEM loads (LDTM), bulk copies (UBLKCP), MUFU and fp64 instructions the shipped cubins contain - so that the
"hand-written sm_100a" claim can be checked without rebuilding. usage: python scripts/sass_census.py > profiles/r2_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "handwritten-chinese-ocr-samples_b200", "libhctr_b200.so")
PATTERNS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "LDTM", "UBLKCP", "UTCBAR", "SYNCS", "REDG", "MUFU.EX2", "DFMA", "HMNMX2", "REDUX"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    counts = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("hctr::", "")
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        c = counts[name]
        c["instructions"] += 1
        for p in PATTERNS:
            if op == p or op.startswith(p + ".") or (p == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                c[p] += 1
    print("# SASS census of %s (%s; %d kernels)" % (os.path.relpath(LIB, ROOT), ", ".join(arch), len(counts)))
    print("%-78s %7s " % ("kernel", "instr") + " ".join("%9s" % p[-9:] for p in PATTERNS))
    tot = collections.Counter()
    for k, c in counts.items():
        tot.update(c)
        print("%-78s %7d " % (k[:78], c["instructions"]) + " ".join("%9d" % c[p] for p in PATTERNS))
    print("%-78s %7d " % ("TOTAL", tot["instructions"]) + " ".join("%9d" % tot[p] for p in PATTERNS))


if __name__ == "__main__":
    main()
