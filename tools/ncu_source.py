#!/usr/bin/env python
"""Per-opcode and hottest-instruction summary from the ncu source page of a --set full capture.
    python tools/ncu_source.py gpurun_out/x.ncu-rep [top_n]
"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(io.StringIO(out)))
# first row: kernel name; second: header
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter(); stall = collections.Counter(); tot = 0; tots = 0
inst = []
for r in rows[2:]:
    if len(r) <= iex or not r[iex].isdigit():
        continue
    n, s = int(r[iex]), int(r[ismp] or 0)
    src = r[isrc].strip()
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    op = op.split(".")[0]
    ops[op] += n; stall[op] += s; tot += n; tots += s
    inst.append((s, n, src))
print(f"total warp-instructions executed {tot}, stall samples {tots}")
print("opcode            executed   share   samples share")
for op, n in ops.most_common(28):
    print(f"  {op:12s} {n:12d} {100*n/tot:6.1f}% {stall[op]:8d} {100*stall[op]/max(tots,1):5.1f}%")
print("hottest instructions by stall samples:")
for s, n, src in sorted(inst, reverse=True)[:topn]:
    print(f"  {s:6d} {n:10d}  {src[:100]}")
