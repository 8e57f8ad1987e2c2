"""Stand-in for the un-installable `einops-exts==0.0.3` (pinned in the reference's requirements.txt).

TEST INFRASTRUCTURE ONLY: lets the unmodified reference at /root/reference import in this container so that it
can act as the parity oracle.  Call sites it serves (reference file:line):
  rearrange_many  minimagen/layers.py:228
  repeat_many     minimagen/layers.py:65, :232
  check_shape     minimagen/Imagen.py:611
All three are thin wrappers over einops; the semantics below are exact by construction.
"""
from einops import rearrange, repeat


def check_shape(tensor, pattern, **kwargs):
    return rearrange(tensor, f"{pattern} -> {pattern}", **kwargs)


def rearrange_many(tensors, pattern, **kwargs):
    return (rearrange(t, pattern, **kwargs) for t in tensors)


def repeat_many(tensors, pattern, **kwargs):
    return (repeat(t, pattern, **kwargs) for t in tensors)
