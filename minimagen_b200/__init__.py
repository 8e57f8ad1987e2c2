"""minimagen_b200 -- B200-native (sm_100a) implementation of MinImagen's U-Net denoising hot path.

Drop-in module layout (same names as the reference package `minimagen`):
    minimagen_b200.Unet             Unet, Base, Super, BaseTest, SuperTest
    minimagen_b200.Imagen           Imagen
    minimagen_b200.diffusion_model  GaussianDiffusion
    minimagen_b200.layers, .helpers, .t5

`install_as_minimagen()` registers these modules under the reference's package name so that the reference's own
`train.py` / `inference.py` / `minimagen.generate` import them unchanged.
"""
import sys

__version__ = "0.1.0"


def install_as_minimagen():
    """Alias this package as `minimagen` in sys.modules (only the hot-path modules; `minimagen.training` /
    `minimagen.generate` stay the reference's own files and import the aliased classes)."""
    import importlib
    pkg = sys.modules[__name__]
    sys.modules.setdefault('minimagen', pkg)
    for name in ('Unet', 'Imagen', 'diffusion_model', 'layers', 'helpers', 't5'):
        mod = importlib.import_module(f'{__name__}.{name}')
        sys.modules[f'minimagen.{name}'] = mod
    return pkg
