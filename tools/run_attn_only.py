import sys, torch
sys.path.insert(0, ".")
from smilecode_b200 import ops
dev = torch.device("cuda"); g = torch.Generator(device=dev).manual_seed(7)
S = (160, 192, 160)
q = torch.randn(1, *S, 6, device=dev, generator=g); k = torch.randn(1, *S, 6, device=dev, generator=g)
rpb = torch.randn(1, 3, 3, 3, device=dev, generator=g) * 0.5
for _ in range(4):
    ops.modet_attention(q, k, rpb, 1, 1.0)
torch.cuda.synchronize()
