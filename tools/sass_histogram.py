#!/usr/bin/env python
"""tools/sass_histogram.py [libyagi_b200.so] -- per-kernel SASS opcode histogram (cuobjdump -sass), so a reader can
see which pipes each kernel uses without disassembling the library: FFMA2/FADD2/FMUL2 (packed f32x2), UBLKCP /
UTMALDG / UTMASTG (TMA), SYNCS (mbarrier), UTC*MMA / LDTM / STTM (tcgen05 + tensor memory), LDGSTS (cp.async)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "yagi_b200", "lib", "libyagi_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
dem = {}
hist = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
names = list(hist)
out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
KEY = ("FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "LDS", "STS", "LDG", "STG",
       "LDTM", "STTM", "SHFL", "BAR")
print("SASS opcode histogram of %s (sm_100a)" % os.path.relpath(so, ROOT))
for mangled, nice in zip(names, out):
    h = hist[mangled]
    nice = re.sub(r"\(anonymous namespace\)::", "", nice)
    nice = re.sub(r"\(.*\)$", "", nice)
    total = sum(h.values())
    utc = {k: v for k, v in h.items() if k.startswith("UTC")}
    keyed = ", ".join("%s %d" % (k, h[k]) for k in KEY if h.get(k)) + ("".join(", %s %d" % kv for kv in sorted(utc.items())))
    print("\n%s\n  %d instructions; %s" % (nice, total, keyed))
    print("  top: " + ", ".join("%s %d" % kv for kv in h.most_common(12)))
