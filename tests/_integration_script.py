"""Runs in a SUBPROCESS of tests/test_integration.py (it rewires sys.modules['minimagen*']).

1. INTEGRATION.md's drop-in snippet verbatim;
2. a Training Directory written the way the reference's `training.create_directory` / `save_training_info` /
   `training.py:389` write it, loaded back by the REFERENCE's own `generate.load_minimagen` (generate.py:79-121), must
   come back as minimagen_b200 classes with the checkpoint weights, and sample.
Prints one JSON line.  argv[1]: 'emu' (CPU: torch emulation of the ops interface) or 'native' (CUDA)."""
import json
import os
import sys
import tempfile
import time
from argparse import Namespace

import torch

mode = sys.argv[1] if len(sys.argv) > 1 else "emu"
res = {}

# ---- 1. the documented snippet, verbatim
import minimagen_b200; minimagen_b200.install_as_minimagen()   # registers minimagen.Unet / .Imagen / .diffusion_model / ...
from minimagen.generate import load_minimagen, sample_and_save   # the reference's own file, now building B200 classes

import minimagen.generate as gen
import minimagen_b200.Imagen
import minimagen_b200.Unet
res["generate_file"] = gen.__file__
res["generate_uses_b200_classes"] = (gen.Imagen is minimagen_b200.Imagen.Imagen and gen.Unet is minimagen_b200.Unet.Unet)

if mode == "emu":
    import minimagen_b200.ops as ops_mod
    from emu_ops import EmuOps
    ops_mod.set_ops(EmuOps())
    dev = "cpu"
else:
    dev = "cuda"

# ---- 2. Training Directory -> reference's load_minimagen -> sample
from minimagen import training                                   # the reference's own training.py (imports the aliases)
res["training_uses_b200_unet"] = training.Unet is minimagen_b200.Unet
from minimagen.Unet import Unet, BaseTest, SuperTest
from minimagen.Imagen import Imagen

unets_params = [training.get_default_args(BaseTest), training.get_default_args(SuperTest)]
imagen_params = dict(image_sizes=(16, 32), timesteps=25, cond_drop_prob=0.15, text_encoder_name="t5_small")
torch.manual_seed(0)
src = Imagen(unets=[Unet(**p) for p in unets_params], **imagen_params)
with tempfile.TemporaryDirectory() as tmp:
    tdir = os.path.join(tmp, "training_run")
    ts = "20260101_000000"
    cm = training.create_directory(tdir)
    args = Namespace(RESTART_DIRECTORY=None, BATCH_SIZE=2, TIMESTEPS=25)
    training.save_training_info(args, ts, unets_params, imagen_params, training.get_model_size(src), cm)
    with cm("state_dicts"):
        for i in range(len(src.unets)):                         # training.py:389 file naming
            torch.save(src.unets[i].state_dict(), f"unet_{i}_state_{ts}.pth")
    t0 = time.perf_counter()
    loaded = load_minimagen(tdir)
    res["load_s"] = time.perf_counter() - t0
res["loaded_type"] = type(loaded).__module__ + "." + type(loaded).__name__
res["unet_types"] = sorted({type(u).__module__ for u in loaded.unets})
sd_a, sd_b = src.unets[1].state_dict(), loaded.unets[1].state_dict()
res["weights_equal"] = list(sd_a) == list(sd_b) and all(torch.equal(sd_a[k].cpu(), sd_b[k].cpu()) for k in sd_a)

g = torch.Generator().manual_seed(1)
te = torch.randn(2, 7, 512, generator=g)
tm = torch.ones(2, 7, dtype=torch.bool)
bank = {}


def noise_fn(kind, shape, step):
    key = (kind, step, tuple(shape))
    if key not in bank:
        bank[key] = torch.randn(*shape, generator=g)
    return bank[key]


outs = []
for m in (src, loaded):
    m = m.to(dev)
    m.noise_fn = noise_fn
    outs.append(m.sample(text_embeds=te.to(dev), text_masks=tm.to(dev), cond_scale=2.).cpu())
# ---- 3. the reference's own training loop (training.MinimagenTrain, the body of train.py) driving the B200 classes:
#         imagen(images, text_embeds=, text_masks=, unet_number=) -> loss.backward() -> optimizer.step(), its checkpointing,
#         its validation pass; then the reference's load_minimagen reads what it wrote
torch.manual_seed(0)
trn = Imagen(unets=[Unet(**p) for p in unets_params], **imagen_params).to(dev)
w_before = [p.detach().clone() for p in trn.parameters()]
gb = torch.Generator().manual_seed(7)
mk = lambda: dict(image=torch.rand(2, 3, 32, 32, generator=gb).to(dev), encoding=torch.randn(2, 7, 512, generator=gb).to(dev),
                  mask=torch.ones(2, 7, dtype=torch.bool).to(dev))
train_dl, valid_dl = [mk(), mk(), mk()], [mk()]
with tempfile.TemporaryDirectory() as tmp2:
    tdir2 = os.path.join(tmp2, "training_run")
    cm2 = training.create_directory(tdir2)
    targs = Namespace(RESTART_DIRECTORY=None, ACCUM_ITER=1, CHCKPT_NUM=2, EPOCHS=1)
    training.save_training_info(targs, ts, unets_params, imagen_params, training.get_model_size(trn), cm2)
    opt = torch.optim.Adam(trn.parameters(), lr=1e-3)
    training.MinimagenTrain(ts, targs, list(trn.unets), trn, train_dl, valid_dl, cm2, opt, timeout=300)
    res["train_files"] = sorted(os.listdir(os.path.join(tdir2, "state_dicts")))
    res["train_progress_lines"] = len(open(os.path.join(tdir2, "training_progess.txt")).read().splitlines())
    trained = load_minimagen(tdir2)
trn._reset_unets_all_one_device()
res["train_weights_changed"] = any(not torch.equal(a, b.detach()) for a, b in zip(w_before, trn.parameters()))
res["trained_loaded_type"] = type(trained).__module__

res["sample_shape"] = list(outs[1].shape)
res["sample_equal"] = bool(torch.allclose(outs[0], outs[1], atol=1e-6))
res["sample_finite"] = bool(torch.isfinite(outs[1]).all())
print(json.dumps(res))
