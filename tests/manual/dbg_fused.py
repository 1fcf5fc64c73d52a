import sys, subprocess, os
if len(sys.argv) == 1:
    for shp in ["1,8,6,12", "1,12,24,64", "1,8,16,32", "1,40,19,80", "2,3,9,36"]:
        for mode in ["attn", "fused"]:
            r = subprocess.run([sys.executable, __file__, shp, mode], capture_output=True, text=True, timeout=100)
            print(shp, mode, (r.stdout.strip().splitlines() or ["-"])[-1], "|", (r.stderr.strip().splitlines() or ["ok"])[-1][:150], flush=True)
    sys.exit(0)
import torch
sys.path.insert(0, '.')
from smilecode_b200 import ops
from oracle import modet_oracle as orc
torch.manual_seed(0)
B, D, H, W = (int(x) for x in sys.argv[1].split(","))
q = torch.randn(B, D, H, W, 6); k = torch.randn(B, D, H, W, 6); rpb = torch.randn(1, 3, 3, 3) * 0.5
w = orc.modet_attention(q, k, rpb, 1, 1.0)
if sys.argv[2] == "attn":
    out = ops.modet_attention(q.cuda(), k.cuda(), rpb.cuda(), 1, 1.0)
    torch.cuda.synchronize()
    print("attn err", float((out.cpu() - w).abs().max()))
else:
    flow = torch.randn(B, 3, D, H, W) * 2; mov = torch.rand(B, 1, D, H, W)
    f_ref = orc.warp_trilinear(flow, w) + w
    m_ref = orc.warp_trilinear(mov, f_ref)
    f, m = ops.modet_fused(q.cuda(), k.cuda(), rpb.cuda(), flow.cuda(), mov.cuda(), 1.0, 1.0)
    torch.cuda.synchronize()
    print("fused err", float((f.cpu() - f_ref).abs().max()), float((m.cpu() - m_ref).abs().max()))
