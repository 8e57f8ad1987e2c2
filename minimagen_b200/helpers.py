"""Small host-side helpers with the names the reference's callers import from `minimagen.helpers`
(reference: minimagen/helpers.py).  Only `extract` / `prob_mask_like` / `right_pad_dims_to` / normalisation touch
tensors, and none of them is on the per-step path of this package (the step kernels do the schedule gathers)."""
from contextlib import contextmanager
from functools import wraps

import torch


def exists(val):
    return val is not None


def default(val, d):
    if exists(val):
        return val
    return d() if callable(d) else d


def cast_tuple(val, length=None):
    """list -> tuple; scalar -> tuple repeated `length` (1 if None) times; checks the length (helpers.py:9-23)."""
    if isinstance(val, list):
        val = tuple(val)
    out = val if isinstance(val, tuple) else ((val,) * default(length, 1))
    if exists(length):
        assert len(out) == length
    return out


def identity(t, *args, **kwargs):
    return t


def maybe(fn):
    @wraps(fn)
    def inner(x):
        return fn(x) if exists(x) else x
    return inner


def eval_decorator(fn):
    """Run `fn` with the model in eval mode, restoring the previous mode afterwards (helpers.py:35-46)."""
    def inner(model, *args, **kwargs):
        was_training = model.training
        model.eval()
        try:
            return fn(model, *args, **kwargs)
        finally:
            model.train(was_training)
    return inner


def module_device(module):
    return next(module.parameters()).device


@contextmanager
def null_context(*args, **kwargs):
    yield


def extract(a, t, x_shape):
    """a[t] reshaped to (b, 1, 1, ...) to broadcast against x_shape (helpers.py:56-67)."""
    b = t.shape[0]
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


def log(t, eps=1e-12):
    return torch.log(t.clamp(min=eps))


def normalize_neg_one_to_one(img):
    return img * 2 - 1


def unnormalize_zero_to_one(normed_img):
    return (normed_img + 1) * 0.5


def prob_mask_like(shape, prob, device):
    """Boolean keep-mask with P(True) = prob; deterministic (no RNG draw) for prob in {0, 1} (helpers.py:121-135)."""
    if prob == 1:
        return torch.ones(shape, device=device, dtype=torch.bool)
    if prob == 0:
        return torch.zeros(shape, device=device, dtype=torch.bool)
    return torch.zeros(shape, device=device).float().uniform_(0, 1) < prob


def right_pad_dims_to(x, t):
    pad = x.ndim - t.ndim
    return t if pad <= 0 else t.view(*t.shape, *((1,) * pad))


def resize_image_to(image, target_image_size, clamp_range=None, pad_mode='reflect'):
    """Inter-stage resize of the cascade (helpers.py:138-164 -> resize_right.resize).  SURVEY.md 8f-1 ("next" row):
    not on the per-step hot path and the third-party source is absent, so it is not rebuilt yet."""
    if image.shape[-1] == target_image_size:
        return image
    raise NotImplementedError(
        "resize_image_to (resize_right cubic resize between cascade stages) is a SURVEY.md 8f 'next' row; "
        "pass `lowres_cond_img` at the target resolution directly")
