#!/usr/bin/env python
"""Join the ncu SASS page (executed counts, stall samples) with nvdisasm line info:
per-source-line executed warp instructions for one kernel.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep smilecode_b200/build/attn_tma.o <kernel-substring> [top]
"""
import collections, csv, io, re, subprocess, sys
rep, obj, sub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
subprocess.run(["cuobjdump", "-xelf", "all", __import__("os").path.abspath(obj)], cwd="/tmp", capture_output=True)
import glob, os
cubins = sorted(glob.glob("/tmp/*.cubin"), key=os.path.getmtime)[-1:]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubins[0]], capture_output=True, text=True).stdout
# parse: function sections ".text.<mangled>:" ; line markers "//## File "...", line N" ; instructions "/*0010*/  OP ..."
line_of = {}
cur_fn, cur_line, want = None, None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        cur_fn = m.group(1); want = sub in cur_fn; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and want:
        line_of[int(m.group(1), 16)] = cur_line
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, iex, ismp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg = collections.defaultdict(lambda: [0, 0])
tot = 0
for r in rows[2:]:
    if len(r) <= iex or not r[iex].isdigit():
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    key = line_of.get(a - base, ("?", 0))
    agg[key][0] += int(r[iex]); agg[key][1] += int(r[ismp] or 0); tot += int(r[iex])
src = {}
print(f"total executed {tot}")
for (f, l), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src and f != "?":
        for root in ("smilecode_b200/csrc/", ""):
            try:
                src[f] = open(root + f).read().splitlines(); break
            except OSError:
                pass
    text = src.get(f, [""] * (l + 1))[l - 1].strip()[:90] if f in src and l > 0 else ""
    print(f"{n:11d} {100*n/tot:5.1f}%  smp {s:5d}  {f}:{l}  {text}")
