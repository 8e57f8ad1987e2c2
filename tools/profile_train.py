"""Kernel table of one training step (Imagen.forward -> backward -> Adam) of the base U-Net used by bench.py's training_step row."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from minimagen_b200.Imagen import Imagen
from minimagen_b200.Unet import Unet

dev = torch.device("cuda")
tb = int(os.environ.get("TB", 8))
cfg = dict(dim=128, dim_mults=(1, 2, 4), num_resnet_blocks=(1, 2, 2), layer_attns=(False, False, True),
           layer_cross_attns=(False, True, True), memory_efficient=True, text_embed_dim=768)
torch.manual_seed(0)
with torch.device(dev):
    u = Unet(**cfg)
im = Imagen(unets=u, text_encoder_name="t5_base", image_sizes=(64,), timesteps=1000, cond_drop_prob=0.1).to(dev).train()
g = torch.Generator().manual_seed(3)
imgs = torch.rand(tb, 3, 64, 64, generator=g).to(dev)
te = torch.randn(tb, 16, 768, generator=g).to(dev)
tm = torch.ones(tb, 16, dtype=torch.bool, device=dev)
opt = torch.optim.Adam(im.parameters(), lr=1e-4)

def one():
    opt.zero_grad(set_to_none=True)
    loss = im(imgs, text_embeds=te, text_masks=tm, unet_number=1)
    loss.backward()
    opt.step()
    return loss

for _ in range(3):
    one()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    one()
e1.record(); torch.cuda.synchronize()
print(f"training step b={tb}: {e0.elapsed_time(e1) / 5:.2f} ms (CUDA events, 5 steps)")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        one()
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        r = rows.setdefault(ev.name, [0.0, 0])
        r[0] += ev.device_time / 3e3
        r[1] += 1 / 3
tot = sum(r[0] for r in rows.values())
print(f"# GPU kernel time {tot:.2f} ms/step")
for name, (ms, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{ms:9.3f} ms {n:7.1f} launches  {name[:150]}")
