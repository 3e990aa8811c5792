"""Per-kernel SASS opcode histogram of the built library (profiles/r0N_sass_histogram.md): python tools/sass_histogram.py"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "mswe-gnn_b200/csrc/libswe_gnn_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
for k, h in sorted(hist.items(), key=lambda kv: -sum(kv[1].values())):
    print(f"{sum(h.values()):6d}  UTCHMMA {h['UTCHMMA']:4d}  LDTM {h['LDTM']:3d}  STTM {h['STTM']:3d}  UBLKCP {h['UBLKCP']:3d}  {k[:90]}")
