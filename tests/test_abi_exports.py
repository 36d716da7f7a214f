"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares; the Python API mirrors the reference's names and signatures.  No compute calls (no GPU here)."""
import ctypes
import inspect
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "supnerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from supnerf_b200 import _lib
    from supnerf_b200.build import build
    build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/supnerf_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(syms), set(_lib.SIGNATURES) ^ set(syms)
    assert _lib.load().snb_abi_version() == 1


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import supnerf_b200 as snb
    m = snb.CodeNeRF()
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 4, 3), torch.zeros(2, 4, 3), torch.zeros(1, 256), torch.zeros(1, 256))
    with pytest.raises(RuntimeError):
        snb.renderer.NeRFRenderer().volume_render(torch.zeros(2, 4), torch.zeros(2, 4, 3), torch.zeros(2, 4))


REF_SIGNATURES = {
    # name: positional parameter names, copied from the reference's defs (utils.py / renderer.py line cited)
    "utils.get_rays": ["K", "c2w", "roi", "uv_steps"],  # utils.py:107
    "utils.get_rays_specified": ["K", "c2w", "x_vec", "y_vec"],  # :138
    "utils.sample_from_rays": ["ro", "vd", "near", "far", "N_samples", "z_fixed"],  # :154
    "utils.sample_from_rays_v2": ["rays", "n_samples"],  # :170
    "utils.volume_rendering": ["sigmas", "rgbs", "z_vals"],  # :187
    "utils.volume_rendering2": ["sigmas", "rgbs", "z_vals"],  # :202
    "utils.volume_rendering_batch": ["sigmas", "rgbs", "z_vals"],  # :220
    "utils.ray_box_intersection": ["ray_o", "ray_d", "aabb_min", "aabb_max"],  # :236
    "utils.ray_box_intersection_tensor": ["ray_o", "ray_d", "aabb_min", "aabb_max"],  # :283
    "utils.prepare_pixel_samples": ["img", "mask_occ", "cam_pose", "obj_diag", "K", "roi", "n_rays", "n_samples",
                                    "shapenet_obj_cood", "sym_aug", "im_sz"],  # :330
    "utils.render_rays": ["model", "device", "img", "mask_occ", "cam_pose", "obj_diag", "K", "roi", "n_samples", "shapecode",
                          "texturecode", "shapenet_obj_cood", "sym_aug", "kitti2nusc", "n_rays"],  # :380
    "utils.render_rays_v2": ["model", "device", "img", "mask_occ", "cam_pose", "obj_diag", "K", "roi", "n_samples", "shapecode",
                             "texturecode", "shapenet_obj_cood", "sym_aug", "kitti2nusc", "im_sz", "n_rays"],  # :435
    "utils.render_rays_specified": ["model", "device", "img", "mask_occ", "cam_pose", "obj_diag", "K", "roi", "x_vec", "y_vec",
                                    "n_samples", "shapecode", "texturecode", "shapenet_obj_cood", "sym_aug", "kitti2nusc"],  # :504
    "utils.render_full_img": ["model", "device", "cam_pose", "obj_sz", "K", "roi", "n_samples", "shapecode", "texturecode",
                              "shapenet_obj_cood", "out_depth", "debug_occ", "kitti2nusc"],  # :554
    "renderer.volume_rendering3": ["sigmas", "rgbs", "z_vals", "white_bkgd"],  # renderer.py:355
    "renderer.render_rays_v3": ["model", "device", "img", "mask_occ", "cam_pose", "obj_wlh", "K", "roi", "n_samples", "shapecode",
                                "texturecode", "shapenet_obj_cood", "sym_aug", "kitti2nusc", "im_sz", "n_rays", "adjust_scale"],  # :382
    "renderer.NeRFRenderer.__init__": ["self", "n_samples", "noise_std", "white_bkgd"],  # :16
    "renderer.NeRFRenderer.sample_from_ray": ["self", "rays"],
    "renderer.NeRFRenderer.volume_render": ["self", "sigmas", "rgbs", "z_vals"],
    "renderer.NeRFRenderer.volume_render_batch": ["self", "sigmas", "rgbs", "z_vals"],
    "renderer.NeRFRenderer.prepare_sampled_rays": ["self", "rays_o", "viewdir", "obj_sz"],
    "renderer.NeRFRenderer.render_rays": ["self", "model", "device", "img", "mask_occ", "cam_pose", "obj_sz", "K", "roi", "shapecode",
                                          "texturecode", "kitti2nusc", "im_sz", "n_rays"],  # :117
    "renderer.NeRFRenderer.render_rays_specified": ["self", "model", "device", "img", "mask_occ", "cam_pose", "obj_sz", "K", "roi",
                                                    "x_vec", "y_vec", "shapecode", "texturecode", "kitti2nusc"],  # :169
    "renderer.NeRFRenderer.prepare_pixel_samples": ["self", "img", "mask_occ", "cam_pose", "obj_sz", "K", "roi", "n_rays", "im_sz"],
    "renderer.NeRFRenderer.render_full_img": ["self", "model", "device", "cam_pose", "obj_sz", "K", "roi", "shapecode", "texturecode",
                                              "out_depth", "debug_occ", "kitti2nusc"],  # :238
}


def test_python_api_mirrors_reference_signatures():
    import supnerf_b200 as snb
    for name, params in REF_SIGNATURES.items():
        obj = snb
        for part in name.split("."):
            obj = getattr(obj, part)
        got = list(inspect.signature(obj).parameters)
        assert got == params, (name, got)


def test_state_dict_keys_match_reference_weight_abi():
    import supnerf_b200 as snb
    from oracle import oracle
    for ctor, init in ((lambda: snb.CodeNeRF(), lambda: oracle.init_codenerf_state()),
                       (lambda: snb.AutoRFMix(3, 1, 256), lambda: oracle.init_codenerf_state(3, 1)),
                       (lambda: snb.SUPNeRF(3, 1, 3, 3, 256), lambda: oracle.init_codenerf_state(3, 1)),
                       (lambda: snb.AutoRF(), lambda: oracle.init_autorf_state())):
        m, sd = ctor(), init()
        assert list(m.state_dict().keys()) == list(sd.keys())
        assert all(m.state_dict()[k].shape == sd[k].shape for k in sd)
        m.load_state_dict(sd)


def test_reference_checkpoints_load_strict_and_round_trip():
    """The checkpoint ABI: a reference state_dict -- including the image-encoder / pose-head entries of AutoRFMix and
    SUPNeRF that are not on this path -- loads with the default strict=True (optimizer_nuscenes.py:1795-1796) and state_dict()
    gives every entry back unchanged.  Keys / shapes come from tests/golden/state_dict_keys.json (tools/make_golden.py)."""
    import json
    import torch
    import supnerf_b200 as snb
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    cases = {"CodeNeRF": (snb.CodeNeRF, ()), "AutoRFMix_3_1_256": (snb.AutoRFMix, (3, 1, 256)), "AutoRF": (snb.AutoRF, ()),
             "SUPNeRF_3_1_3_3_256": (snb.SUPNeRF, (3, 1, 3, 3, 256))}
    g = torch.Generator().manual_seed(0)
    for name, (cls, args) in cases.items():
        sd = {}
        for k, (shape, dtype) in keys[name].items():
            dt = getattr(torch, dtype)
            sd[k] = torch.randn(shape, generator=g).to(dt) if dt.is_floating_point else torch.zeros(shape, dtype=dt)
        m = cls(*args)
        res = m.load_state_dict(sd)
        assert not res.missing_keys and not res.unexpected_keys
        out = m.state_dict()
        assert set(out) == set(sd) and all(torch.equal(out[k], sd[k]) for k in sd), name
        for k, p_ in m.named_parameters():
            assert torch.equal(p_.detach(), sd[k])
    with pytest.raises(RuntimeError):
        snb.CodeNeRF().load_state_dict({"bogus.weight": torch.zeros(1)})


def test_shipped_artefacts_do_not_name_batch_memcpy_entry_points():
    """a statically linked CUDA runtime embeds the names of every runtime entry point, including the batched memcpy
    calls the GPU pool refuses.  The library links the shared runtime; no binary that travels to the GPU box may carry the names."""
    import subprocess
    from supnerf_b200 import _lib
    from supnerf_b200.build import build
    build()
    banned = re.compile(rb"cu(da)?Memcpy(3D)?BatchAsync")
    tracked = subprocess.run(["git", "ls-files"], cwd=ROOT, capture_output=True, text=True).stdout.split()
    paths = [_lib.LIB_PATH] + [os.path.join(ROOT, p) for p in tracked]
    for p in paths:
        if not os.path.isfile(p) or p.endswith("test_abi_exports.py") or os.path.basename(p) in ("ADVICE.md", "VERDICT.md"):
            continue
        with open(p, "rb") as f:
            assert not banned.search(f.read()), p
    with open(_lib.LIB_PATH, "rb") as f:
        assert b"libcudart.so" in f.read()  # dynamic runtime


def test_result_file_formats_round_trip(tmp_path):
    """The refine loops' result files (optimizer_nuscenes.py:1463-1476 codes+poses.pth, :1405-1409 cross_eval.pth): same keys."""
    import torch
    from supnerf_b200 import scene
    p = str(tmp_path / "codes+poses.pth")
    vals = dict(num_obj=2, optimized_shapecodes=torch.zeros(2, 3, 256), optimized_texturecodes=torch.zeros(2, 3, 256), optimized_poses=torch.zeros(2, 3, 3, 4),
                psnr_eval={"a": [1.0]}, ssim_eval={}, depth_err_mean={}, lidar_pts_cnt={}, R_eval={}, T_eval={})
    scene.save_opts_w_pose(p, **vals)
    got = scene.load_result(p)
    assert list(got.keys()) == ["num_obj", "optimized_shapecodes", "optimized_texturecodes", "optimized_poses", "psnr_eval", "ssim_eval",
                                "depth_err_mean", "lidar_pts_cnt", "R_eval", "T_eval"]
    q = str(tmp_path / "cross_eval.pth")
    scene.save_cross_eval(q, {"i": [torch.ones(2, 2)]}, {}, {}, [0, 5, 10])
    assert list(scene.load_result(q).keys()) == ["psnr_eval_mat_per_ins", "depth_eval_mat_per_ins", "cnt_lidar_pts_per_ins", "CODE_SAVE_ITERS_"]
