"""Stand-in for `einops_exts.torch.EinopsToAndFrom` (reference call sites: minimagen/layers.py:403, :492;
minimagen/Unet.py:272).  The wrapped module MUST be stored as attribute `fn`: the reference's state_dict keys are
`...cross_attn.fn.*`, `...attn.fn.*`, `mid_attn.fn.fn.*`.  TEST INFRASTRUCTURE ONLY."""
from einops import rearrange
from torch import nn


class EinopsToAndFrom(nn.Module):
    def __init__(self, from_einops, to_einops, fn):
        super().__init__()
        self.from_einops = from_einops
        self.to_einops = to_einops
        self.fn = fn

    def forward(self, x, **kwargs):
        shape = x.shape
        reconstitute_kwargs = dict(tuple(zip(self.from_einops.split(' '), shape)))
        x = rearrange(x, f'{self.from_einops} -> {self.to_einops}')
        x = self.fn(x, **kwargs)
        x = rearrange(x, f'{self.to_einops} -> {self.from_einops}', **reconstitute_kwargs)
        return x
