"""Training step (Imagen.forward -> backward -> Adam) eager vs captured in one CUDA graph; stock PyTorch (oracle/restatement.py,
bench baseline leg) eager and graphed beside it.  Usage: TB=8 python tools/train_graph.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
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


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3


def graphed(step, opt):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    opt.zero_grad(set_to_none=True)
    with torch.cuda.graph(gr):
        loss = step(zero=False)
    return gr, loss


# ---- this library
opt = torch.optim.Adam(im.parameters(), lr=1e-4, capturable=True)
def mine(zero=True):
    if zero:
        opt.zero_grad(set_to_none=True)
    loss = im(imgs, text_embeds=te, text_masks=tm, unet_number=1)
    loss.backward()
    opt.step()
    return loss
ev, wall = timed(mine)
print(f"b={tb} minimagen_b200 eager   : {ev:7.2f} ms/step (device)  {wall:7.2f} ms wall", flush=True)
try:
    gr, loss = graphed(mine, opt)
    ev, wall = timed(gr.replay)
    print(f"b={tb} minimagen_b200 graphed : {ev:7.2f} ms/step (device)  {wall:7.2f} ms wall   loss {float(loss):.4f}", flush=True)
    l0 = float(loss)
    for _ in range(20):
        gr.replay()
    print(f"      loss after 20 more graphed steps on the same batch: {float(loss):.4f} (was {l0:.4f})", flush=True)
except Exception as ex:
    import traceback
    print("graph capture of the library step failed:", type(ex).__name__, str(ex)[:400], flush=True)
    traceback.print_exc(limit=-12)

# ---- stock PyTorch on the same weights / shapes
from oracle import restatement as R
sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in u.state_dict().items()}
leaves = [v for v in sd.values() if v.requires_grad]
topt = torch.optim.Adam(leaves, lr=1e-4, capturable=True)
tt = torch.randint(0, 1000, (tb,), generator=g).to(dev)
xin, tgt = torch.randn(tb, 3, 64, 64, generator=g).to(dev), torch.randn(tb, 3, 64, 64, generator=g).to(dev)
for name, dt in (("fp32/tf32", None), ("fp16 autocast", torch.float16)):
    def ref(zero=True):
        if zero:
            topt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=dt or torch.float16, enabled=dt is not None):
            pred = R.unet_forward(sd, cfg, xin, tt, text_embeds=te, text_mask=tm)
        l_ = F.mse_loss(pred.float(), tgt)
        l_.backward()
        topt.step()
        return l_
    ev, wall = timed(ref)
    print(f"b={tb} stock PyTorch {name:13s} eager  : {ev:7.2f} ms/step (device)  {wall:7.2f} ms wall", flush=True)
    try:
        gr2, _ = graphed(ref, topt)
        ev, wall = timed(gr2.replay)
        print(f"b={tb} stock PyTorch {name:13s} graphed: {ev:7.2f} ms/step (device)  {wall:7.2f} ms wall", flush=True)
        del gr2
    except Exception as ex:
        print(f"graph capture of the stock PyTorch step ({name}) failed:", type(ex).__name__, str(ex)[:300], flush=True)
