"""N > 1 path on CPU: two gloo ranks shard the batch of `Imagen.sample(distributed=True)`, each runs the whole (tiny)
cascade stage on its shard through the torch emulation of the ops, and ONE all-gather assembles the images.  The result
must equal the single-process full-batch run (noise is a function of the global sample index)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(g):
    from minimagen_b200.Imagen import Imagen
    from minimagen_b200.Unet import Unet
    u = Unet(**g["cfg"]).eval()
    u.load_state_dict(g["state_dict"])
    im = Imagen(unets=u, text_encoder_name="t5_small", image_sizes=(64,), timesteps=25, cond_drop_prob=0.15).eval()
    im.unets[0].load_state_dict(g["state_dict"])
    im.use_cuda_graph = False
    return im


def _noise_bank(total_b):
    gen = torch.Generator().manual_seed(123)
    bank = {}

    def get(kind, step):
        key = (kind, step)
        if key not in bank:
            bank[key] = torch.randn(total_b, 3, 64, 64, generator=gen)
        return bank[key]
    # materialise deterministically in a fixed order so every process sees the same numbers
    get("init", -1)
    for s in range(24, -1, -1):
        get("step", s)
    return bank


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import minimagen_b200.ops as ops_mod
    from emu_ops import EmuOps
    ops_mod.set_ops(EmuOps())
    g = torch.load(os.path.join(ROOT, "tests", "golden", "sample_loop.pt"), map_location="cpu", weights_only=False)
    im = _build(g)
    B = 4
    bank = _noise_bank(B)
    per = B // world
    im.noise_fn = lambda kind, shape, step: bank[(kind, step)][rank * per:(rank + 1) * per]
    gen = torch.Generator().manual_seed(7)
    te = torch.randn(B, 9, 512, generator=gen)
    tm = torch.ones(B, 9, dtype=torch.bool)
    tm[1, 4:] = False
    out = im.sample(text_embeds=te, text_masks=tm, cond_scale=3., distributed=True)
    assert out.shape == (B, 3, 64, 64)
    if rank == 0:
        torch.save(out, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gloo_sampling_matches_single_process(tmp_path, emu):
    port = 29600 + (os.getpid() % 200)
    out_path = str(tmp_path / "dist_out.pt")
    mp.spawn(_worker, args=(2, port, out_path), nprocs=2, join=True)
    dist_out = torch.load(out_path)
    # single process, full batch, same noise bank
    from conftest import load_golden
    g = load_golden("sample_loop.pt")
    im = _build(g)
    bank = _noise_bank(4)
    im.noise_fn = lambda kind, shape, step: bank[(kind, step)]
    gen = torch.Generator().manual_seed(7)
    te = torch.randn(4, 9, 512, generator=gen)
    tm = torch.ones(4, 9, dtype=torch.bool)
    tm[1, 4:] = False
    full = im.sample(text_embeds=te, text_masks=tm, cond_scale=3.)
    assert torch.allclose(dist_out, full, atol=1e-5)
