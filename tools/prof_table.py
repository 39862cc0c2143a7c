"""Per-step self-CUDA-time table from a tools/profile_mp.py log (3 profiled steps).  Usage: python tools/prof_table.py LOG [top]"""
import re
import sys

rows, hdr = [], None
for line in open(sys.argv[1]):
    if line.startswith("-"):
        continue
    if "Name" in line and "Self CUDA" in line:
        hdr = re.split(r"\s{2,}", line.strip())
        continue
    parts = re.split(r"\s{2,}", line.strip())
    if hdr and len(parts) >= 8:
        rows.append(parts)


def us(s):
    for suf, m in (("ms", 1e3), ("us", 1.0), ("s", 1e6)):
        if s.endswith(suf):
            return float(s[: -len(suf)]) * m
    return 0.0


i_self, i_calls = hdr.index("Self CUDA"), hdr.index("# of Calls")
out = [(us(r[i_self]) / 3, int(r[i_calls]) / 3, r[0][:100]) for r in rows if us(r[i_self]) > 0]
out.sort(reverse=True)
kern = [o for o in out if o[2].startswith("void ") or "::" in o[2].split("(")[0] and not o[2].startswith(("aten::", "autograd::"))]
for o in kern[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%9.1f us/step  x%5.1f  %s" % o)
print("kernels total per step: %.1f us" % sum(o[0] for o in kern))
