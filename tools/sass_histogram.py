#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass): the evidence that the hot kernels are tcgen05 / TMEM / TMA code.
   python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "selenite_lite_b200", "lib", "libselenite_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEY = ["UTCIMMA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "IMMA", "HMMA", "FFMA", "FFMA2", "FADD2", "FMUL2", "I2FP", "F2I", "PRMT", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "MEMBAR", "ATOM", "RED"]
kern = None; hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "")
        kern = re.sub(r"\(.*", "", kern).replace("sl::", "")
        hist[kern] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1; hist[kern]["_total"] += 1
print("SASS opcode counts per kernel of %s (sm_100a)\n" % os.path.relpath(lib, ROOT))
print("%-44s %7s  %s" % ("kernel", "instrs", "  ".join("%s" % k for k in KEY)))
for k, h in hist.items():
    if h["_total"] < 200:
        continue
    print("%-44s %7d  %s" % (k[:44], h["_total"], "  ".join("%*d" % (len(n), h[n]) for n in KEY)))
