#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel in an object file.
    python tools/sass_hist.py smilecode_b200/build/attn_tma.o ILi8ELi3ELb1ELb1ELi2E [top]
"""
import collections, re, subprocess, sys
obj, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
want = False
hist = collections.Counter()
for ln in out.splitlines():
    if "Function :" in ln:
        want = sub in ln
        continue
    if not want:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", ln)
    if m:
        hist[m.group(1)] += 1
print("total", sum(hist.values()))
print("  ".join(f"{k}:{v}" for k, v in hist.most_common(top)))
