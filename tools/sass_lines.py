#!/usr/bin/env python
"""Static SASS instruction count per source line for one kernel (nvdisasm -g line info).
    python tools/sass_lines.py obj.o <kernel-substring> [--cold a-b,c-d ...]   (line ranges of attn_tma.cu counted as cold)
"""
import collections, glob, os, re, subprocess, sys, tempfile
obj, sub = os.path.abspath(sys.argv[1]), sys.argv[2]
cold = []
if "--cold" in sys.argv:
    for rg in sys.argv[sys.argv.index("--cold") + 1].split(","):
        a, b = rg.split("-"); cold.append((int(a), int(b)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
cubin = sorted(glob.glob(tmp + "/*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
want, cur = False, None
per = collections.Counter()
ops = collections.Counter()
def _cold(f,l):
    return (f == 'common.cuh' and 37 <= l <= 80) or (f == 'attn_tma.cu' and any(a <= l <= b for a, b in cold))
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        want = sub in m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        # inlined frames: keep the outermost attn_tma.cu line when present ("inlined at" chain)
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3)); continue
    if want and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        per[cur[:2] if cur else ("?", 0)] += 1
        mm = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", ln)
        if mm and cur and not _cold(cur[0], cur[1]):
            ops[mm.group(1)] += 1
tot = sum(per.values())
hot = 0
for (f, l), n in per.items():
    is_cold = f == "common.cuh" and l >= 37 and l <= 80
    if f == "attn_tma.cu" and any(a <= l <= b for a, b in cold):
        is_cold = True
    if not is_cold:
        hot += n
print(f"total {tot}  hot {hot}  (hot/3 = {hot/3:.0f} per plane step)")
print("  ".join(f"{k}:{v/3:.0f}" for k, v in ops.most_common(45)))
if "-v" in sys.argv:
    for (f, l), n in sorted(per.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        print(f"{n:5d}  {f}:{l}")
