"""cProfile of the host side of eager training steps (where do the ~30 ms of Python go?)."""
import os, sys, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minimagen_b200.Imagen import Imagen
from minimagen_b200.Unet import Unet
dev = torch.device("cuda"); tb = int(os.environ.get("TB", 8))
cfg = dict(dim=128, dim_mults=(1, 2, 4), num_resnet_blocks=(1, 2, 2), layer_attns=(False, False, True),
           layer_cross_attns=(False, True, True), memory_efficient=True, text_embed_dim=768)
torch.manual_seed(0)
with torch.device(dev):
    u = Unet(**cfg)
im = Imagen(unets=u, text_encoder_name="t5_base", image_sizes=(64,), timesteps=1000, cond_drop_prob=0.1).to(dev).train()
g = torch.Generator().manual_seed(3)
imgs = torch.rand(tb, 3, 64, 64, generator=g).to(dev); te = torch.randn(tb, 16, 768, generator=g).to(dev)
tm = torch.ones(tb, 16, dtype=torch.bool, device=dev)
opt = torch.optim.Adam(u.parameters(), lr=1e-4)
def one():
    opt.zero_grad(set_to_none=True)
    loss = im(imgs, text_embeds=te, text_masks=tm, unet_number=1)
    loss.backward()
    opt.step()
for _ in range(3): one()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): one()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); ps = pstats.Stats(pr, stream=s).sort_stats("tottime"); ps.print_stats(32); print(s.getvalue()[:6000])
s = io.StringIO(); ps = pstats.Stats(pr, stream=s).sort_stats("cumulative"); ps.print_stats(28); print(s.getvalue()[:5000])
