import sys, torch
sys.path.insert(0, ".")
from smilecode_b200 import models
from smilecode_b200.synth import make_pair, randomize_weights
S=(160,192,160)
model = models.ModeT(S, head_dim=6, num_heads=[8,4,2,1,1], scale=1); randomize_weights(model, seed=1234); model=model.cuda().eval()
for B in (1,2,4):
    m,f = make_pair(S, batch=B, seed=24); m,f=m.cuda(),f.cuda()
    with torch.no_grad():
        for _ in range(3): model(m,f)
        torch.cuda.synchronize()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): model(m,f)
        b.record(); torch.cuda.synchronize()
    ms=a.elapsed_time(b)/10
    print(f"B={B}: {ms:.3f} ms/forward -> {B*1e3/ms:.1f} pairs/s, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
