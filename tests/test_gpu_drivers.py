"""Every remaining public driver of SURVEY 8(b) (row a9) through the drop-in package on the GPU, against outputs of the UNMODIFIED
reference (tests/golden/drivers.npz, made by tools/make_golden.py:golden_drivers): renderer.render_rays_v3, utils.render_rays /
render_rays_specified / prepare_pixel_samples / render_full_img, NeRFRenderer.render_rays_specified / prepare_pixel_samples /
render_full_img.  fp32 back end; tolerance: per-tensor max|a-b|/max|b| <= 1e-5 for renders, samples AND gradients (north star).
The pose / latent gradients are sums with heavy cancellation where the reference's own fp32 value is up to 5e-3 away from the
exact one: for those the fixture tests/golden/drivers_truth64.npz holds the same reference drivers run in float64, and
conftest.parity passes a gradient that is within 1e-5 of the fp32 reference OR no further from the fp64 truth than 1.5 x the
fp32 reference itself.  Every measured error lands in profiles/parity_r2.json."""
import numpy as np
import pytest
import torch

from conftest import T, load_golden, parity, parity_ok, rel_err
from oracle import oracle
from test_gpu_parity import DEV, forced_rand_like, model_from_state, snb

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ctx():
    g = load_golden("drivers")
    g.update(load_golden("drivers_truth64"))
    S = snb()
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=int(g["seed"]))
    m = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
    m.precision = "fp32"
    m.requires_grad_(False)
    return g, S, m


def _leaves(g):
    return (T(g["cam_pose"], device=DEV).requires_grad_(), T(g["shapecode"], device=DEV).requires_grad_(),
            T(g["texturecode"], device=DEV).requires_grad_())


def _check_render(got, g, pre):
    rgb, dep, acc, tgt, occ = got
    assert parity_ok("tgt", tgt, g[pre + "_tgt"], 1e-6) and torch.equal(occ.cpu(), T(g[pre + "_occ"]))
    for name, t in (("rgb", rgb), ("depth", dep), ("acc", acc)):
        parity(pre + "_" + name, t, g[pre + "_" + name], TOL)


def _check_grads(cam, shp, tex, g, pre):
    for name, t in (("g_cam", cam), ("g_shp", shp), ("g_tex", tex)):
        parity(pre + "_" + name, t.grad, g[pre + "_" + name], TOL, truth=g[pre + "_" + name + "64"])


def test_render_rays_v3_full_grid_and_random_subset(ctx):
    g, S, m = ctx
    cam, shp, tex = _leaves(g)
    with forced_rand_like(T(g["v3_jitter"])):
        out = S.renderer.render_rays_v3(m, DEV, T(g["img"]), T(g["mask_occ"]), cam, g["wlh"], T(g["K"]), T(g["roi"]), 64, shp, tex, 1, 0,
                                        im_sz=8, n_rays=None, adjust_scale=1.1)
    _check_render(out, g, "v3")
    loss = oracle.refine_losses(out[0], out[2], out[3], out[4])[0]
    parity("v3_loss", loss, g["v3_loss"], TOL)
    loss.backward()
    _check_grads(cam, shp, tex, g, "v3")
    cam, shp, tex = _leaves(g)
    np.random.seed(71)
    with forced_rand_like(T(g["v3s_jitter"])):
        out = S.renderer.render_rays_v3(m, DEV, T(g["img"]), T(g["mask_occ"]), cam, g["wlh"], T(g["K"]), T(g["roi"]), 64, shp, tex, 0, 0,
                                        im_sz=8, n_rays=20)
    _check_render(out, g, "v3s")


def test_renderer_class_drivers_on_a_full_resolution_crop(ctx):
    g, S, m = ctx
    R = S.renderer.NeRFRenderer(n_samples=16)
    img, mask, roi = T(g["img_s"]), T(g["mask_s"]), T(g["roi_s"])
    cam, shp, tex = _leaves(g)
    with forced_rand_like(T(g["rs_jitter"])):
        out = R.render_rays_specified(m, DEV, img, mask, cam, g["wlh"], T(g["K"]), roi, g["x_vec"], g["y_vec"], shp, tex)
    _check_render(out, g, "rs")
    oracle.refine_losses(out[0], out[2], out[3], out[4])[0].backward()
    _check_grads(cam, shp, tex, g, "rs")
    np.random.seed(73)
    with forced_rand_like(T(g["rp_jitter"])):
        xyz, vd, zv, tgt, occ = R.prepare_pixel_samples(img, mask, T(g["cam_pose"], device=DEV), g["wlh"], T(g["K"]), roi, 40)
    assert parity_ok("xyz", xyz, g["rp_xyz"], TOL) and parity_ok("vd", vd, g["rp_viewdir"], TOL) and parity_ok("zv", zv, g["rp_z_vals"], TOL)
    assert parity_ok("tgt", tgt, g["rp_tgt"], 1e-6) and torch.equal(occ.cpu(), T(g["rp_occ"]))
    with torch.no_grad(), forced_rand_like(T(g["rf_jitter"])):
        im, dep = R.render_full_img(m, DEV, T(g["cam_pose"], device=DEV), g["wlh"], T(g["K"]), roi, T(g["shapecode"], device=DEV),
                                    T(g["texturecode"], device=DEV), out_depth=True)
    assert tuple(im.shape) == (10, 12, 3) and tuple(dep.shape) == (10, 12)
    assert parity_ok("im", im, g["rf_img"], TOL) and parity_ok("dep", dep, g["rf_depth"], TOL)


def test_utils_drivers_on_a_full_resolution_crop(ctx):
    g, S, m = ctx
    img, mask, roi, diag = T(g["img_s"]), T(g["mask_s"]), T(g["roi_s"]), np.float32(g["obj_diag"])
    cam, shp, tex = _leaves(g)
    np.random.seed(75)
    torch.manual_seed(75)
    out = S.utils.render_rays(m, DEV, img, mask, cam, diag, T(g["K"]), roi, 16, shp, tex, 1, 0, n_rays=50)
    _check_render(out, g, "ur")
    oracle.refine_losses(out[0], out[2], out[3], out[4])[0].backward()
    _check_grads(cam, shp, tex, g, "ur")
    cam, shp, tex = _leaves(g)
    torch.manual_seed(76)
    out = S.utils.render_rays_specified(m, DEV, img, mask, cam, diag, T(g["K"]), roi, g["x_vec"], g["y_vec"], 16, shp, tex, 1, 0)
    _check_render(out, g, "us")
    np.random.seed(77)
    torch.manual_seed(77)
    xyz, vd, zv, tgt, occ = S.utils.prepare_pixel_samples(img, mask, T(g["cam_pose"], device=DEV), diag, T(g["K"]), roi, 40, 16, 1, 0)
    assert parity_ok("xyz", xyz, g["up_xyz"], TOL) and parity_ok("vd", vd, g["up_viewdir"], TOL) and parity_ok("zv", zv, g["up_z_vals"], TOL)
    assert parity_ok("tgt", tgt, g["up_tgt"], 1e-6) and torch.equal(occ.cpu(), T(g["up_occ"]))
    torch.manual_seed(78)
    with torch.no_grad():
        im, dep = S.utils.render_full_img(m, DEV, T(g["cam_pose"], device=DEV), g["wlh"], T(g["K"]), roi, 16, T(g["shapecode"], device=DEV),
                                          T(g["texturecode"], device=DEV), 1, out_depth=True)
    assert parity_ok("im", im, g["uf_img"], TOL) and parity_ok("dep", dep, g["uf_depth"], TOL)


def test_autorf_decoder_fp32_back_end_golden():
    """Row a7: AutoRF.forward (model_autorf.py:156-186) on the fp32 back end against the reference's outputs and gradients."""
    g = load_golden("autorf_decoder")
    S = snb()
    sd = oracle.init_autorf_state(seed=int(g["seed"]))
    m = model_from_state(S.AutoRF, sd)
    m.precision = "fp32"
    ins = [T(g[k], device=DEV).requires_grad_() for k in ("xyz", "viewdir", "shapecode", "texturecode")]
    sig, rgbs = m(*ins)
    parity("sigmas", sig, g["sigmas"], TOL)
    parity("rgbs", rgbs, g["rgbs"], TOL)
    ((sig * T(g["up_sigma"], device=DEV)).sum() + (rgbs * T(g["up_rgb"], device=DEV)).sum()).backward()
    # fp64 truth of the same computation: the oracle (pinned to this fixture by tests/test_oracle_golden.py) in float64
    sd64 = {k: v.double().requires_grad_() for k, v in sd.items()}
    in64 = [T(g[k]).double().requires_grad_() for k in ("xyz", "viewdir", "shapecode", "texturecode")]
    s64, c64 = oracle.autorf_decoder(sd64, *in64)
    ((s64 * T(g["up_sigma"]).double()).sum() + (c64 * T(g["up_rgb"]).double()).sum()).backward()
    for t, t64, k in zip(ins, in64, ("g_xyz", "g_viewdir", "g_shapecode", "g_texturecode")):
        parity(k, t.grad, g[k], TOL, truth=t64.grad)
    for k, p_ in m.named_parameters():
        if "gw_" + k in g:
            parity("gw_" + k, p_.grad, g["gw_" + k], TOL, truth=sd64[k].grad)   # sums over every sample


@pytest.mark.parametrize("blocks,B,n,S_", [((1, 2), 1, 40, 8), ((2, 3), 2, 24, 8), ((5, 5), 3, 16, 4)])
def test_autorf_decoder_block_counts_vs_oracle_all_grads(blocks, B, n, S_):
    """AutoRF with the smallest legal block counts (shape_blocks 1: the sigma head mixes the latent itself; texture_blocks 2: the
    3-way mix takes the texture latent itself), a mid-size and the default 5/5, several objects per call: outputs, input gradients
    and every weight gradient against the CPU oracle (itself pinned to the reference by tests/golden/autorf_decoder.npz)."""
    S = snb()
    sd = oracle.init_autorf_state(shape_blocks=blocks[0], texture_blocks=blocks[1], seed=90 + blocks[0])
    g = torch.Generator().manual_seed(blocks[0] * 10 + B)
    xyz = (torch.rand(B * n, S_, 3, generator=g) - 0.5) * 1.6
    vd = torch.nn.functional.normalize(torch.randn(B * n, 1, 3, generator=g), dim=-1).repeat(1, S_, 1)
    shp, tex = torch.randn(B, 128, generator=g) * 0.3, torch.randn(B, 128, generator=g) * 0.3
    up_s, up_c = torch.randn(B * n, S_, 1, generator=g), torch.randn(B * n, S_, 3, generator=g)
    sdr = {k: v.clone().requires_grad_() for k, v in sd.items()}
    ins = [t.clone().requires_grad_() for t in (xyz, vd, shp, tex)]
    sig, rgbs = oracle.autorf_decoder(sdr, *ins, shape_blocks=blocks[0], texture_blocks=blocks[1])
    ((sig * up_s).sum() + (rgbs * up_c).sum()).backward()
    m = model_from_state(S.AutoRF, sd, shape_blocks=blocks[0], texture_blocks=blocks[1])
    gin = [t.to(DEV).requires_grad_() for t in (xyz, vd, shp, tex)]
    sig2, rgbs2 = m(*gin)
    parity("sigmas", sig2, sig, TOL)
    parity("rgbs", rgbs2, rgbs, TOL)
    ((sig2 * up_s.to(DEV)).sum() + (rgbs2 * up_c.to(DEV)).sum()).backward()
    sd64 = {k: v.double().requires_grad_() for k, v in sd.items()}
    in64 = [t.double().requires_grad_() for t in (xyz, vd, shp, tex)]
    s64, c64 = oracle.autorf_decoder(sd64, *in64, shape_blocks=blocks[0], texture_blocks=blocks[1])
    ((s64 * up_s.double()).sum() + (c64 * up_c.double()).sum()).backward()
    for a, b, c, name in zip(gin, ins, in64, ("xyz", "viewdir", "shape", "texture")):
        parity("g_" + name, a.grad, b.grad, TOL, truth=c.grad)
    for k, p_ in m.named_parameters():
        parity("gw_" + k, p_.grad, sdr[k].grad, TOL, truth=sd64[k].grad)


def test_kitti2nusc_rotation_and_symmetric_augmentation(ctx):
    """The frame fix-ups no shipped caller enables but the API carries: kitti2nusc=True through NeRFRenderer.render_rays
    (renderer.py:153-163) and, together with the y-flip of sym_aug (python `random` seeded so that the flip is taken), through
    utils.render_rays_v2 (utils.py:474-489)."""
    import random
    g, S, m = ctx
    R = S.renderer.NeRFRenderer(n_samples=16)
    cam, shp, tex = _leaves(g)
    with forced_rand_like(T(g["k2n_jitter"])):
        out = R.render_rays(m, DEV, T(g["img"]), T(g["mask_occ"]), cam, g["wlh"], T(g["K"]), T(g["roi"]), shp, tex, kitti2nusc=True, im_sz=8)
    _check_render(out, g, "k2n")
    oracle.refine_losses(out[0], out[2], out[3], out[4])[0].backward()
    _check_grads(cam, shp, tex, g, "k2n")
    cam, shp, tex = _leaves(g)
    random.seed(2)
    torch.manual_seed(80)
    out = S.utils.render_rays_v2(m, DEV, T(g["img"]), T(g["mask_occ"]), cam, np.float32(g["obj_diag"]), T(g["K"]), T(g["roi"]), 16, shp, tex,
                                 1, 1, kitti2nusc=True, im_sz=8, n_rays=None)
    _check_render(out, g, "sym")
    oracle.refine_losses(out[0], out[2], out[3], out[4])[0].backward()
    _check_grads(cam, shp, tex, g, "sym")
