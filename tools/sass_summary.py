#!/usr/bin/env python
"""Instruction histogram per kernel of libdiffpose_b200.so (cuobjdump -sass, runs without a GPU):
the Blackwell-native mnemonics (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk, SYNCS = mbarrier, F2FP = fp32->fp16 pack, FFMA2/FADD2/FMUL2 = packed fp32 pairs) and the
total instruction count of each kernel.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffpose_nw_b200", "libdiffpose_b200.so")
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "F2FP", "FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU", "LDS", "STS",
       "LDG", "STG", "BAR", "SHFL", "DFMA", "DADD", "DMUL", "HMMA", "WARPSYNC", "NANOSLEEP", "FENCE", "ACQBULK"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(anonymous namespace\)::", "", kern)
            kern = re.sub(r"\(.*", "", kern)
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and kern:
            op, mods = m.group(1), m.group(2)
            hist[kern][op] += 1
            hist[kern]["_total"] += 1
            if op in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "F2FP", "SYNCS", "UTCBAR"):
                hist[kern][op + mods] += 1
    print(f"# SASS instruction histogram of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); static counts, one line per kernel")
    print("# tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM, tcgen05.commit = UTCBAR, cp.async.bulk = UBLKCP, mbarrier = SYNCS")
    for k, h in hist.items():
        print(f"\n{k}: {h['_total']} instructions")
        print("   " + "  ".join(f"{op}={h[op]}" for op in KEY if h[op]))
        det = sorted((n, c) for n, c in h.items() if "." in n)
        if det:
            print("   " + "  ".join(f"{n}={c}" for n, c in det))
    tot = collections.Counter()
    for h in hist.values():
        tot.update({op: h[op] for op in KEY})
    print("\nwhole library: " + "  ".join(f"{op}={tot[op]}" for op in KEY if tot[op]))


if __name__ == "__main__":
    main()
