#!/usr/bin/env python
"""Counts the SASS mnemonics that prove what the shipped kernels are made of (TMA = UTMALDG, packed FMA = FFMA2, mbarrier =
SYNCS, spills = STL/LDL ...), per kernel of ssim_b200/lib/libssim_cuda.so.  Runs on the build host (cuobjdump, no GPU):

    python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ssim_b200", "lib", "libssim_cuda.so")
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "PRMT", "LDS", "STS", "STG", "LDG", "SYNCS",
         "VOTE", "NANOSLEEP", "ATOMG", "RED", "MEMBAR", "BAR", "USETMAXREG", "STL", "LDL", "HMMA", "UTCHMMA", "I2F", "F2F", "DADD", "BRA"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["(all)"] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("library: %s   cubin architectures: %s" % (os.path.relpath(lib, ROOT), ", ".join(arch)))
    print("git: %s" % subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip())
    for (name, c), pretty in zip(kernels.items(), demangled):
        print("\n%s\n  instructions: %d" % (pretty, c["(all)"]))
        print("  " + "  ".join("%s %d" % (k, c[k]) for k in WATCH if c[k]))
        if not (c["STL"] or c["LDL"]):
            print("  no local-memory spills (STL/LDL absent)")


if __name__ == "__main__":
    main()
