"""The oracle is only trusted once pinned: restatement == golden vectors (bit exact, CPU fp32) and, when the reference
tree is present (build container), restatement == the unmodified reference on fresh inputs."""
import pytest
import torch

from conftest import load_golden
from oracle import reference, restatement as R


@pytest.mark.parametrize("name,lowres", [("unet_tiny_base.pt", False), ("unet_tiny_sr.pt", True)])
def test_restatement_matches_golden_unet(name, lowres):
    g = load_golden(name)
    inp = g["inputs"]
    kw = {k: v for k, v in inp.items() if k not in ("x", "time")}
    with torch.no_grad():
        assert torch.equal(R.unet_forward(g["state_dict"], g["cfg"], inp["x"], inp["time"], **kw), g["out_cond"])
        assert torch.equal(R.unet_forward(g["state_dict"], g["cfg"], inp["x"], inp["time"], cond_drop_prob=1., **kw),
                           g["out_null"])
        kw2 = dict(kw, text_mask=None)
        assert torch.equal(R.unet_forward(g["state_dict"], g["cfg"], inp["x"], inp["time"], **kw2), g["out_nomask"])
        cfg3 = R.cfg_combine(g["out_cond"], g["out_null"], 3.)
        assert torch.equal(cfg3, g["out_cfg3"])


@pytest.mark.parametrize("T", [25, 1000])
def test_restatement_matches_golden_step(T):
    g = load_golden("ddpm_step.pt")[T]
    tabs = R.ddpm_tables(T)
    for k, v in tabs.items():
        assert torch.equal(v, g["tables"][k]), k
    out = R.p_sample_step(tabs, g["x"], g["t"], g["eps"], g["noise"])
    assert torch.equal(out, g["out"])


def test_quantile_rank_is_fp32_arithmetic():
    """timestep-index / percentile-rank work is integer work: must be exact.  n = 3*1024^2 gives weight 0.25 because
    torch computes 0.9*(n-1) in fp32."""
    from minimagen_b200.Imagen import quantile_rank
    ranks = load_golden("ddpm_step.pt")["ranks"]
    assert ranks[3 * 1024 * 1024] == (2831154, 2831155, 0.25)
    for n, expect in ranks.items():
        assert quantile_rank(n, 0.9) == expect
        # and torch.quantile really behaves like sorted[lo] lerp sorted[hi] with that weight
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 12288, generator=g).abs()
    lo, hi, w = quantile_rank(12288, 0.9)
    srt = x.sort(dim=-1).values
    assert torch.equal(torch.lerp(srt[:, lo], srt[:, hi], torch.tensor(w)), torch.quantile(x, 0.9, dim=-1))


def test_sample_loop_golden_matches_restatement():
    g = load_golden("sample_loop.pt")
    tabs = R.ddpm_tables(g["timesteps"])
    img = g["x_T"]
    with torch.no_grad():
        for i in range(3):
            t = torch.full((2,), g["timesteps"] - 1 - i)
            kw = dict(text_embeds=g["text_embeds"], text_mask=g["text_mask"])
            e = R.unet_forward(g["state_dict"], g["cfg"], img, t, **kw)
            n = R.unet_forward(g["state_dict"], g["cfg"], img, t, cond_drop_prob=1., **kw)
            img = R.p_sample_step(tabs, img, t, R.cfg_combine(e, n, g["cond_scale"]), g["noises"][i])
            assert torch.equal(img, g["traj"][i])


@pytest.mark.skipif(not reference.available(), reason="reference tree only exists in the build container")
def test_restatement_matches_live_reference():
    reference.load()
    from minimagen.Unet import Unet as RUnet
    cfgs = [
        (dict(dim=32, dim_mults=(1, 2), attend_at_middle=True, text_embed_dim=768), 32, False),
        (dict(dim=32, dim_mults=(1, 2), lowres_cond=True, memory_efficient=True, num_resnet_blocks=(1, 2),
              layer_attns=(False, True), layer_cross_attns=(False, True)), 32, True),
    ]
    for cfg, s, lowres in cfgs:
        torch.manual_seed(0)
        r = RUnet(**cfg).eval()
        g = torch.Generator().manual_seed(7)
        x = torch.randn(2, 3, s, s, generator=g)
        te = torch.randn(2, 20, cfg.get("text_embed_dim", 512), generator=g)
        tm = torch.ones(2, 20, dtype=torch.bool)
        tm[1, 5:] = False
        kw = dict(text_embeds=te, text_mask=tm)
        if lowres:
            kw.update(lowres_cond_img=torch.randn(2, 3, s, s, generator=g), lowres_noise_times=torch.tensor([200, 3]))
        t = torch.tensor([999, 0])
        with torch.no_grad():
            for cdp in (0., 1.):
                assert torch.equal(r(x, t, cond_drop_prob=cdp, **kw),
                                   R.unet_forward(r.state_dict(), cfg, x, t, cond_drop_prob=cdp, **kw))
