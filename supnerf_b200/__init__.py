"""supnerf_b200 — B200-native (sm_100a) implementation of SUP-NeRF's object-centric volumetric render
hot path, behind the reference's own Python call API.

    from supnerf_b200 import renderer, utils            # drop-ins for src/renderer.py, src/utils.py (render half)
    from supnerf_b200.models import CodeNeRF, AutoRFMix, SUPNeRF, AutoRF

Everything computes in hand-written CUDA kernels reached through the C ABI of libsupnerf_b200.so
(include/supnerf_b200.h); there is no CPU or eager-PyTorch fallback."""
from . import _lib, ops  # noqa: F401
from .models import AutoRF, AutoRFMix, CodeNeRF, SUPNeRF, get_default_precision, set_default_precision  # noqa: F401
from . import losses, parallel, refine, renderer, scene, synthetic, utils  # noqa: F401,E402

__all__ = ["renderer", "utils", "ops", "parallel", "synthetic", "losses", "refine", "scene", "CodeNeRF", "AutoRFMix", "SUPNeRF", "AutoRF", "set_default_precision",
           "get_default_precision"]
