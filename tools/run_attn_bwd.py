import sys, torch
sys.path.insert(0, ".")
from smilecode_b200 import ops
dev = torch.device("cuda"); g = torch.Generator(device=dev).manual_seed(7)
S = (160, 192, 160)
q = torch.randn(1, *S, 6, device=dev, generator=g); k = torch.randn(1, *S, 6, device=dev, generator=g)
rpb = torch.randn(1, 3, 3, 3, device=dev, generator=g) * 0.5
G = torch.randn(1, 3, *S, device=dev, generator=g)
for _ in range(2):
    ops.modet_attention_bwd(G, q, k, rpb, 1, 1.0)
torch.cuda.synchronize()
src = torch.randn(1, 8, *S, device=dev, generator=g); flow = torch.randn(1, 3, *S, device=dev, generator=g)
G8 = torch.randn(1, 8, *S, device=dev, generator=g)
ops.warp3d_bwd(G8, src, flow)
torch.cuda.synchronize()
