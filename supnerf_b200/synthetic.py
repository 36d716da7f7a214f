"""Seeded synthetic workloads for the render hot path: camera, object pose/size, roi, targets,
latents and random-init decoder weights with the reference's ``state_dict`` names.

Input generation only — no rendering arithmetic lives here.  ``bench.py``, ``__graft_entry__.smoke()``, the tests
and the CPU oracle all draw their inputs from this one module so that every arm sees identical bits.
Citations are ``file:line`` into ``/root/reference/src``."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


def init_codenerf_state(shape_blocks=2, texture_blocks=1, W=256, num_xyz_freq=10, num_dir_freq=4, latent_dim=256,
                        seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Random-init weights with the reference's parameter names/shapes and ``nn.Linear``'s default
    init, created in the reference's registration order (model_codenerf.py:22-37) under ``seed``."""
    g_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def add(name, fin, fout):
        lin = torch.nn.Linear(fin, fout)
        sd[name + ".weight"] = lin.weight.detach().to(dtype)
        sd[name + ".bias"] = lin.bias.detach().to(dtype)

    d_xyz, d_dir = 3 + 6 * num_xyz_freq, 3 + 6 * num_dir_freq
    add("encoding_xyz.0", d_xyz, W)
    for j in range(1, shape_blocks + 1):
        add(f"shape_latent_layer_{j}.0", latent_dim, W)
        add(f"shape_layer_{j}.0", W, W)
    add("encoding_shape", W, W)
    add("sigma.0", W, 1)
    add("encoding_viewdir.0", W + d_dir, W)
    for j in range(1, texture_blocks + 1):
        add(f"texture_latent_layer_{j}.0", latent_dim, W)
        add(f"texture_layer_{j}.0", W, W)
    add("rgb.0", W, W // 2)
    add("rgb.2", W // 2, 3)
    torch.random.set_rng_state(g_state)
    return sd


def init_autorf_state(shape_blocks=5, texture_blocks=5, latent_dim=128, num_xyz_freq=10, num_dir_freq=4,
                      seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Decoder half of model_autorf.py:123-150 (no image encoder), registration order preserved."""
    g_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def add(name, fin, fout):
        lin = torch.nn.Linear(fin, fout)
        sd[name + ".weight"] = lin.weight.detach().to(dtype)
        sd[name + ".bias"] = lin.bias.detach().to(dtype)

    d_xyz, d_dir = 3 + 6 * num_xyz_freq, 3 + 6 * num_dir_freq
    add("encoding_xyz.0", d_xyz, latent_dim)
    for j in range(shape_blocks - 1):
        add(f"shape_layer_{j}.0", latent_dim, latent_dim)
    add("sigma.0", latent_dim, 1)
    for j in range(texture_blocks - 2):
        add(f"texture_layer_{j}.0", latent_dim, latent_dim)
    add(f"texture_layer_{texture_blocks - 2}.0", latent_dim + d_dir, latent_dim)
    add("rgb.0", latent_dim + d_dir, 3)
    torch.random.set_rng_state(g_state)
    return sd


NUSC_K = [[1266.4, 0.0, 816.27], [0.0, 1266.4, 491.5], [0.0, 0.0, 1.0]]
WLH_MEAN = [1.9446588, 4.641784, 1.7103361]
WLH_STD = [0.1611075, 0.3961748, 0.20885137]


def synthetic_object(seed: int, im_sz: int, margin: int = 5) -> Dict[str, object]:
    """One synthetic car: K, cam_pose (camera->object, 3x4), wlh, roi, targets."""
    rng = np.random.RandomState(seed)
    K = np.asarray(NUSC_K, dtype=np.float32)
    wlh = (np.asarray(WLH_MEAN) + np.asarray(WLH_STD) * rng.randn(3)).astype(np.float32)
    yaw = rng.uniform(-math.pi, math.pi)
    base = np.asarray([[0, -1, 0], [0, 0, -1], [1, 0, 0]], dtype=np.float64)  # utils.py:1337-1339
    cy, sy = math.cos(yaw), math.sin(yaw)
    Rz = np.asarray([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], dtype=np.float64)
    R = base @ Rz  # object -> camera
    depth = rng.uniform(8.0, 40.0)
    lat = rng.uniform(-0.25, 0.25) * depth
    t = np.asarray([lat, rng.uniform(0.5, 1.5), depth], dtype=np.float64)
    c2o = np.concatenate([R.T, (-R.T @ t)[:, None]], 1).astype(np.float32)  # data_nuscenes.py:479-481
    # roi: square around the projected 3-D box
    w, l, h = wlh.astype(np.float64)
    corners = np.asarray([[sx * l / 2, sy_ * w / 2, sz * h / 2] for sx in (-1, 1) for sy_ in (-1, 1) for sz in (-1, 1)])
    cam = (R @ corners.T).T + t
    uv = (K.astype(np.float64) @ cam.T).T
    uv = uv[:, :2] / uv[:, 2:3]
    x0, y0 = uv.min(0) - margin
    x1, y1 = uv.max(0) + margin
    side = max(x1 - x0, y1 - y0, 8.0)
    cxm, cym = (x0 + x1) / 2, (y0 + y1) / 2
    roi = np.asarray([cxm - side / 2, cym - side / 2, cxm + side / 2, cym + side / 2]).astype(np.int32)
    trng = torch.Generator().manual_seed(seed)
    img = torch.rand(im_sz, im_sz, 3, generator=trng)
    mask = -torch.ones(im_sz, im_sz, 1)
    q = im_sz // 8
    mask[q:im_sz - q, q:im_sz - q] = 0
    mask[2 * q:im_sz - 2 * q, 2 * q:im_sz - 2 * q] = 1
    return dict(K=torch.from_numpy(K), cam_pose=torch.from_numpy(c2o), wlh=wlh, roi=torch.from_numpy(roi),
                img=img, mask_occ=mask)


def synthetic_latents(seed: int, B: int, D: int = 256) -> Tuple[Tensor, Tensor]:
    """randn(B,D)/sqrt(D/2) (trainer_unified_nuscenes.py:443)."""
    g = torch.Generator().manual_seed(1000 + seed)
    s = torch.randn(B, D, generator=g) / math.sqrt(D / 2)
    t = torch.randn(B, D, generator=g) / math.sqrt(D / 2)
    return s, t


def pose_estimator_state(template, seed=0):
    """Deterministic weights for the pose-estimator half of SUPNeRF (img_encoder.*, pose_layer_*, regress_layer_*, out_delta_layer):
    every entry of `template` (a state_dict: key -> tensor, only shapes / dtypes are read) whose key belongs to that half is drawn
    from ONE seeded generator in key order -- conv / linear weights He-scaled, norm weights near 1, running statistics near (0, 1)
    -- so that the reference module (tools/make_golden.py) and the drop-in (tests) hold identical parameters without shipping
    49 M floats.  Decoder entries are not touched (init_codenerf_state covers them)."""
    g = torch.Generator().manual_seed(1000 + seed)
    out = {}
    for k, v in template.items():
        if not k.startswith(("img_encoder.", "pose_layer_", "regress_layer_", "out_delta_layer")):
            continue
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros(shape, dtype=torch.int64)
        elif k.endswith("running_mean"):
            out[k] = 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("running_var"):
            out[k] = 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif len(shape) == 4:      # conv: He, fan_out
            out[k] = torch.randn(shape, generator=g) * (2.0 / (shape[0] * shape[2] * shape[3])) ** 0.5
        elif len(shape) == 2:      # linear
            out[k] = torch.randn(shape, generator=g) * (1.0 / shape[1]) ** 0.5
        elif ".bn" in k or "downsample.1" in k:
            out[k] = (1.0 + 0.1 * torch.randn(shape, generator=g)) if k.endswith("weight") else 0.1 * torch.randn(shape, generator=g)
        else:                      # linear bias
            out[k] = 0.05 * torch.randn(shape, generator=g)
    return out
