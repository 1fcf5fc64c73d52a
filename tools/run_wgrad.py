import sys, torch
sys.path.insert(0, ".")
from smilecode_b200 import ops
dev = torch.device("cuda"); g = torch.Generator(device=dev).manual_seed(7)
S = (160, 192, 160)
x = torch.randn(2, 8, *S, device=dev, generator=g); dy = torch.randn(2, 8, *S, device=dev, generator=g)
w = torch.randn(8, 8, 3, 3, 3, device=dev, generator=g)
for _ in range(3):
    ops.conv3d_bwd(dy, x, w, need_x=False)
torch.cuda.synchronize()
