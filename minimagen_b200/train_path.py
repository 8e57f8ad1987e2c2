"""Unet.forward under autograd: the training side of the hot path (reference Unet.py:355-634, layers.py).

Same network, same parameters and the same kernels as the sampling path (minimagen_b200/Unet.py `_forward_dev`), but expressed
over plain fp32 NHWC tensors through the autograd Functions of `minimagen_b200.autograd`, so that `loss.backward()` reaches
every parameter.  torch itself only keeps the tape and does shape glue (cat / slice / permute / where / small vector adds);
every convolution, normalisation, attention, linear layer and their gradients run in the library's kernels.
"""
import torch
import torch.nn.functional as F

from .autograd import AttentionFn, Conv2dFn, GroupNormSiluFn, LayerNormFn, LinearFn, Upsample2xFn
from .helpers import exists, prob_mask_like
from .layers import (Attention, Conv2d, CrossAttention, Identity, Parallel, ResnetBlock, TransformerBlock, _UpsampleSeq)
from .ops import get_ops

F32 = torch.float32


def _conv(m, x):
    return Conv2dFn.apply(x, m.weight, m.bias, m.stride[0], m.padding[0])


def _linear(m, rows):
    return LinearFn.apply(rows, m.weight, m.bias)


def _ln(norm, x, pre_gelu=False):
    """layers.LayerNorm (gamma + zero beta buffer), nn.LayerNorm (weight / bias) or ChanLayerNorm (g) over the last dim."""
    if hasattr(norm, 'gamma'):
        return LayerNormFn.apply(x, norm.gamma, norm.beta, 1e-5, pre_gelu)
    if hasattr(norm, 'g'):
        return LayerNormFn.apply(x, norm.g, None, norm.eps, pre_gelu)
    return LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps, pre_gelu)


def _block(blk, x, scale_shift=None):
    """Block.forward (layers.py:131-145)"""
    gn = blk.groupnorm
    h = GroupNormSiluFn.apply(x, gn.weight, gn.bias, scale_shift, gn.num_groups, gn.eps)
    return _conv(blk.project, h)


def _attention_core(att, x_tokens, ctx_tokens, multi_query):
    """CrossAttention.forward (layers.py:220-251) / multi-query Attention.forward (layers.py:52-104) on [B, n, C] tokens."""
    B, n, C = x_tokens.shape
    xn = _ln(att.norm, x_tokens)
    q = LinearFn.apply(xn.reshape(B * n, C), att.to_q.weight, None).reshape(B, n, -1) * att.scale
    src = xn if multi_query else ctx_tokens
    m = src.shape[1]
    kv = LinearFn.apply(src.reshape(B * m, src.shape[2]), att.to_kv.weight, None).reshape(B, m, -1)
    k, v = kv.chunk(2, dim=-1)
    o = AttentionFn.apply(q, k, v, att.null_kv, att.heads)
    y = LinearFn.apply(o.reshape(B * n, -1), att.to_out[0].weight, None).reshape(B, n, C)
    return _ln(att.to_out[1], y)


def _resnet(rb, x, t, c):
    """ResnetBlock.forward (layers.py:417-439)"""
    scale_shift = None
    if exists(rb.time_mlp) and exists(t):
        scale_shift = _linear(rb.time_mlp[1], F.silu(t))             # [B, 2*dim_out] = [scale | shift]
    h = _block(rb.block1, x)
    if exists(rb.cross_attn):
        B, H, W, C = h.shape
        h = _attention_core(rb.cross_attn.fn, h.reshape(B, H * W, C), c, False).reshape(B, H, W, C) + h
    h = _block(rb.block2, h, scale_shift)
    res = x if isinstance(rb.res_conv, Identity) else _conv(rb.res_conv, x)
    return h + res


def _transformer(tb, x):
    """TransformerBlock.forward (layers.py:496-499): x = attn(x) + x ; x = ff(x) + x"""
    if isinstance(tb, Identity):
        return x
    B, H, W, C = x.shape
    x = _attention_core(tb.attn.fn, x.reshape(B, H * W, C), None, True).reshape(B, H, W, C) + x
    ln1, conv1, _, ln2, conv2 = tb.ff
    rows = x.reshape(B * H * W, C)
    h = LinearFn.apply(_ln(ln1, rows), conv1.weight.reshape(conv1.out_channels, C), None)
    h = LinearFn.apply(_ln(ln2, h, pre_gelu=True), conv2.weight.reshape(C, conv1.out_channels), None)
    return h.reshape(B, H, W, C) + x


def _down(mod, x):
    if isinstance(mod, Parallel):                       # Unet.py:233-234: conv3x3(x) + conv1x1(x)
        return _conv(mod.fns[0], x) + _conv(mod.fns[1], x)
    return _conv(mod, x)


def _time_branch(unet, times, hiddens_seq, cond_seq, tokens_seq):
    B = times.shape[0]
    pos = hiddens_seq[0](times)                          # SinusoidalPosEmb kernel (no parameters)
    hid = F.silu(_linear(hiddens_seq[1], pos))
    return _linear(cond_seq[0], hid), _linear(tokens_seq[0], hid).reshape(B, unet.num_time_tokens, unet.cond_dim)


def unet_forward_train(unet, x, time, *, lowres_cond_img=None, lowres_noise_times=None, text_embeds=None, text_mask=None,
                       cond_drop_prob=0., cond_keep=None):
    """Unet.forward (Unet.py:355-472) with autograd.  x: (b, c, s, s) NCHW fp32 -> predicted noise (b, c_out, s, s)."""
    B = x.shape[0]
    device = x.device
    # --- conditioning (Unet.py:508-634)
    t, time_tokens = _time_branch(unet, time, unet.to_time_hiddens, unet.to_time_cond, unet.to_time_tokens)
    if unet.lowres_cond:
        lt, ltok = _time_branch(unet, lowres_noise_times, unet.to_lowres_time_hiddens, unet.to_lowres_time_cond,
                                unet.to_lowres_time_tokens)
        t = t + lt
        time_tokens = torch.cat((time_tokens, ltok), dim=-2)
    c = time_tokens
    if exists(text_embeds):
        L, E = text_embeds.shape[1], text_embeds.shape[2]
        tok = _linear(unet.text_to_cond, text_embeds.to(F32).reshape(B * L, E)).reshape(B, L, -1)[:, :unet.max_text_len]
        rem = unet.max_text_len - tok.shape[1]
        if rem > 0:
            tok = F.pad(tok, (0, 0, 0, rem))
        keep = (cond_keep.to(device=device, dtype=torch.bool) if exists(cond_keep)
                else prob_mask_like((B,), 1 - cond_drop_prob, device=device))
        keep_embed = keep[:, None, None]
        if exists(text_mask):
            tm = text_mask[:, :unet.max_text_len]
            if rem > 0:
                tm = F.pad(tm, (0, rem), value=False)
            keep_embed = tm[:, :, None] & keep_embed
        tok = torch.where(keep_embed, tok, unet.null_text_embed.to(tok.dtype))
        pooled = tok.mean(dim=-2)
        ln, lin1, _, lin2 = unet.to_text_non_attn_cond
        hid = _linear(lin2, F.silu(_linear(lin1, _ln(ln, pooled))))
        hid = torch.where(keep[:, None], hid, unet.null_text_hidden.to(hid.dtype))
        t = t + hid
        c = torch.cat((time_tokens, tok), dim=-2)
    c = _ln(unet.norm_cond, c)

    # --- body, NHWC
    if exists(lowres_cond_img):
        x = torch.cat((x, lowres_cond_img), dim=1)
    h = x.to(F32).permute(0, 2, 3, 1).contiguous()
    h = torch.cat([_conv(cv, h) for cv in unet.init_conv.convs], dim=-1)

    hiddens = []
    for pre_down, init_block, resnet_blocks, attn_block, post_down in unet.downs:
        if exists(pre_down):
            h = _down(pre_down, h)
        h = _resnet(init_block, h, t, c)
        for rb in resnet_blocks:
            h = _resnet(rb, h, t, None)
            hiddens.append(h)
        h = _transformer(attn_block, h)
        hiddens.append(h)
        if exists(post_down):
            h = _down(post_down, h)

    h = _resnet(unet.mid_block1, h, t, c)
    if exists(unet.mid_attn):
        att = unet.mid_attn.fn.fn
        Bh, H, W, C = h.shape
        h = _attention_core(att, h.reshape(Bh, H * W, C), None, True).reshape(Bh, H, W, C) + h
    h = _resnet(unet.mid_block2, h, t, c)

    skip = lambda cur: torch.cat((cur, hiddens.pop() * unet.skip_connect_scale), dim=-1)
    for init_block, resnet_blocks, attn_block, upsample in unet.ups:
        h = _resnet(init_block, skip(h), t, c)
        for rb in resnet_blocks:
            h = _resnet(rb, skip(h), t, None)
        h = _transformer(attn_block, h)
        if isinstance(upsample, _UpsampleSeq):
            h = _conv(upsample[1], Upsample2xFn.apply(h))

    h = _resnet(unet.final_res_block, h, t, None)
    out = _conv(unet.final_conv, h)
    return out.permute(0, 3, 1, 2).contiguous()
