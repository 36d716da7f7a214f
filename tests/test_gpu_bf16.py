"""bf16 / tcgen05 back end of the decoder (SNB_PREC_BF16) against the fp32 CPU oracle.
Tolerance (north star): per-tensor max|a-b|/max|b| <= 2e-2 for renders and gradients."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import T, load_golden, parity, parity_ok, rel_err
from oracle import oracle
from test_gpu_parity import DEV, forced_rand_like, model_from_state, snb

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _case(B, n, S_, seed):
    g = torch.Generator().manual_seed(seed)
    xyz = (torch.rand(B * n, S_, 3, generator=g) - 0.5) * 1.6
    vd = torch.nn.functional.normalize(torch.randn(B * n, 1, 3, generator=g), dim=-1).repeat(1, S_, 1)
    shp, tex = oracle.synthetic_latents(seed, B)
    up_s, up_c = torch.randn(B * n, S_, 1, generator=g), torch.randn(B * n, S_, 3, generator=g)
    return xyz, vd, shp, tex, up_s, up_c


def test_tc_forward_layer_by_layer():
    """Every layer's post-activation output (debug dump of the fused kernel) against the oracle's activations."""
    S = snb()
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=5)
    xyz, vd, shp, tex, _, _ = _case(2, 16, 16, 5)   # 256 rows / object, 4 tiles
    with torch.no_grad():
        sig, rgbs, acts = oracle.codenerf_decoder(sd, xyz, vd, shp, tex, return_acts=True)
    m = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
    m.precision = "bf16"
    M = xyz.shape[0] * xyz.shape[1]
    dbg = torch.zeros(len(acts), M, 256, device=DEV)
    lib = S._lib.load()
    hnd = m._handle(dbg.device).h
    lib.snb_tc_set_debug(hnd, ctypes.c_void_p(dbg.data_ptr()))
    try:
        with torch.no_grad():
            sig2, rgbs2 = m(xyz.to(DEV), vd.to(DEV), shp.to(DEV), tex.to(DEV))
        torch.cuda.synchronize()
    finally:
        lib.snb_tc_set_debug(hnd, None)
    errs = []
    for i, a in enumerate(acts):
        w = a.shape[-1]
        errs.append(parity("layer_%d_activations" % i, dbg[i, :, :w], a.reshape(M, w), TOL))
    print("per-layer rel err:", ["%.2e" % e for e in errs])
    assert parity_ok("sig2", sig2, sig, TOL) and parity_ok("rgbs2", rgbs2, rgbs, TOL)


@pytest.mark.parametrize("blocks,B,n,S_", [((2, 1), 1, 128, 16), ((3, 1), 4, 32, 8), ((5, 3), 2, 16, 16), ((5, 5), 2, 16, 16)])
def test_tc_decoder_fwd_bwd_vs_oracle(blocks, B, n, S_):
    S = snb()
    sd = oracle.init_codenerf_state(shape_blocks=blocks[0], texture_blocks=blocks[1], seed=blocks[0])
    xyz, vd, shp, tex, up_s, up_c = _case(B, n, S_, blocks[0] * 7 + B)
    ins = [t.clone().requires_grad_() for t in (xyz, vd, shp, tex)]
    sig, rgbs = oracle.codenerf_decoder(sd, *ins)
    ((sig * up_s).sum() + (rgbs * up_c).sum()).backward()
    # the same arithmetic with the kernel's bf16 rounding points (CPU emulation): separates kernel bugs from bf16 noise.
    # Per-sample RANDOM upstream gradients make every reduction cancel, so ReLU units whose sign flips under bf16
    # rounding dominate: the emulation itself sits 3-20 % from the fp32 oracle here; the realistic
    # render + loss case below is the one held to the 2e-2 budget.
    ine = [t.clone().requires_grad_() for t in (xyz, vd, shp, tex)]
    sig_e, rgbs_e = oracle.codenerf_decoder_bf16(sd, *ine)
    ((sig_e * up_s).sum() + (rgbs_e * up_c).sum()).backward()
    m = model_from_state(S.CodeNeRF, sd, shape_blocks=blocks[0], texture_blocks=blocks[1])
    m.precision = "bf16"
    m.requires_grad_(False)  # refine mode: frozen weights (the bf16 back end produces no weight gradients)
    gin = [t.to(DEV).requires_grad_() for t in (xyz, vd, shp, tex)]
    sig2, rgbs2 = m(*gin)
    assert parity_ok("sig2", sig2, sig, TOL) and parity_ok("rgbs2", rgbs2, rgbs, TOL)
    assert parity_ok("sig2", sig2, sig_e, 5e-3) and parity_ok("rgbs2", rgbs2, rgbs_e, 5e-3)
    ((sig2 * up_s.to(DEV)).sum() + (rgbs2 * up_c.to(DEV)).sum()).backward()
    names = ("xyz", "viewdir", "shape", "texture")
    # kernel vs the bf16 emulation (same rounding points): the kernel-correctness check, held to the bf16 budget; kernel vs the fp32
    # oracle: at most 1.25 x what the emulation of the prescribed rounding itself shows on these adversarial inputs (recorded)
    for a, b, c, n_ in zip(gin, ins, ine, names):
        floor = rel_err(c.grad, b.grad)
        parity("g_%s_vs_bf16_emulation" % n_, a.grad, c.grad, TOL, floor=floor, floor_slack=1.0)
        parity("g_%s_vs_fp32_oracle" % n_, a.grad, b.grad, TOL, floor=floor)
    # the same decoder under a SMOOTH upstream gradient (identical for every sample: what a loss over rendered pixels looks like):
    # reductions are well conditioned and the plain 2e-2 budget against the fp32 oracle applies
    us, uc = torch.tensor(0.7), torch.tensor([0.3, -0.5, 0.9])
    ins_s = [t.clone().requires_grad_() for t in (xyz, vd, shp, tex)]
    sig_s, rgbs_s = oracle.codenerf_decoder(sd, *ins_s)
    ((sig_s * us).sum() + (rgbs_s * uc).sum()).backward()
    gin_s = [t.to(DEV).requires_grad_() for t in (xyz, vd, shp, tex)]
    sig4, rgbs4 = m(*gin_s)
    ((sig4 * us.to(DEV)).sum() + (rgbs4 * uc.to(DEV)).sum()).backward()
    for a, b, n_ in zip(gin_s[2:], ins_s[2:], names[2:]):   # the reductions over samples: latent gradients
        parity("smooth_upstream_g_%s" % n_, a.grad, b.grad, TOL)
    ine_s = [t.clone().requires_grad_() for t in (xyz, vd, shp, tex)]
    sig_es, rgbs_es = oracle.codenerf_decoder_bf16(sd, *ine_s)
    ((sig_es * us).sum() + (rgbs_es * uc).sum()).backward()
    for a, b, c, n_ in zip(gin_s[:2], ins_s[:2], ine_s[:2], names[:2]):
        # PER-SAMPLE gradients (no reduction): a single ReLU unit whose sign flips under bf16 rounding changes one sample's
        # gradient at full size (the 2^9 PE frequency multiplies it), so max|a-b|/max|b| sits at 10-20 % for ANY bf16 MLP
        # (the emulation's own figure is recorded); the quantities the path returns are their sums over samples (pose gradient),
        # held to 2e-2 in the render tests below
        parity("smooth_upstream_g_%s" % n_, a.grad, b.grad, TOL, floor=rel_err(c.grad, b.grad), floor_slack=1.5)
    # latents only (no pose gradient requested): the shorter backward program must give the same latent gradients
    gin2 = [t.to(DEV) for t in (xyz, vd)] + [t.to(DEV).requires_grad_() for t in (shp, tex)]
    sig3, rgbs3 = m(*gin2)
    ((sig3 * up_s.to(DEV)).sum() + (rgbs3 * up_c.to(DEV)).sum()).backward()
    assert parity_ok("gin2_2__grad", gin2[2].grad, gin[2].grad, 1e-4) and parity_ok("gin2_3__grad", gin2[3].grad, gin[3].grad, 1e-4)


def test_tc_weight_grads_fail_loudly_where_unsupported():
    """Architectures the two-tile kernels do not cover (more than 4 latent slots) have no bf16 weight gradients: loud error."""
    S = snb()
    sd = oracle.init_codenerf_state(shape_blocks=5, texture_blocks=3, seed=1)
    m = model_from_state(S.CodeNeRF, sd, shape_blocks=5, texture_blocks=3)
    m.precision = "bf16"
    xyz, vd, shp, tex, _, _ = _case(1, 8, 16, 1)
    with pytest.raises(RuntimeError):
        sig, rgbs = m(xyz.to(DEV), vd.to(DEV), shp.to(DEV), tex.to(DEV))
        (sig.sum() + rgbs.sum()).backward()
    m.requires_grad_(False)
    # 100 rows of ONE object: padded to the 128-row tile internally
    xs, vs = xyz[:, :10].to(DEV), vd[:, :10].to(DEV)          # 8 rays x 10 samples = 80 rows
    sig, rgbs = m(xs, vs, shp.to(DEV), tex.to(DEV))
    assert sig.shape == (8, 10, 1) and rgbs.shape == (8, 10, 3) and bool(torch.isfinite(sig).all())
    sig_o, rgb_o = oracle.codenerf_decoder(sd, xs.cpu(), vs.cpu(), shp, tex)
    assert parity_ok("sig", sig, sig_o, TOL) and parity_ok("rgbs", rgbs, rgb_o, TOL)
    with pytest.raises(RuntimeError):  # batched latents: 50 rows per object cannot be tile-aligned
        s2, t2 = oracle.synthetic_latents(2, 2)
        m(xs, vs, s2.to(DEV), t2.to(DEV))   # 40 rows per object


@pytest.mark.parametrize("blocks,B,n,S_", [((3, 1), 2, 64, 16), ((2, 1), 1, 256, 8)])
def test_tc_training_mode_weight_grads_vs_oracle(blocks, B, n, S_):
    """bf16 TRAINING mode (weights require grad => SNB_PREC_BF16_TRAIN): every weight / bias gradient from the tcgen05
    weight-gradient kernel (MN-major operands straight from the saved operand tiles), the heads and the latent layers,
    against the CPU emulation of the kernel's bf16 rounding points and the fp32 oracle.  A smooth upstream gradient
    (the same for every sample) keeps the reductions well conditioned so the 2e-2 budget applies."""
    S = snb()
    sd = oracle.init_codenerf_state(shape_blocks=blocks[0], texture_blocks=blocks[1], seed=blocks[0] + 10)
    xyz, vd, shp, tex, _, _ = _case(B, n, S_, blocks[0] * 5 + B)
    up_s, up_c = torch.tensor(0.7), torch.tensor([0.3, -0.5, 0.9])

    def run_oracle(fn):
        sdr = {k: v.clone().requires_grad_() for k, v in sd.items()}
        ins = [t.clone().requires_grad_() for t in (xyz, vd, shp, tex)]
        sig, rgbs = fn(sdr, *ins)
        ((sig * up_s).sum() + (rgbs * up_c).sum()).backward()
        return sdr, ins
    sd32, in32 = run_oracle(oracle.codenerf_decoder)
    sde, ine = run_oracle(oracle.codenerf_decoder_bf16)
    m = model_from_state(S.CodeNeRF, sd, shape_blocks=blocks[0], texture_blocks=blocks[1])
    m.precision = "bf16"
    gin = [t.to(DEV).requires_grad_() for t in (xyz, vd, shp, tex)]
    sig2, rgbs2 = m(*gin)
    ((sig2 * up_s.to(DEV)).sum() + (rgbs2 * up_c.to(DEV)).sum()).backward()
    for k, p_ in m.named_parameters():
        assert p_.grad is not None, k
        floor = rel_err(sde[k].grad, sd32[k].grad)
        parity("gw_%s_vs_bf16_emulation" % k, p_.grad, sde[k].grad, TOL, floor=floor, floor_slack=1.0)
        parity("gw_%s_vs_fp32_oracle" % k, p_.grad, sd32[k].grad, TOL, floor=floor)
    for a, b, c, name in zip(gin, in32, ine, ("xyz", "viewdir", "shape", "texture")):
        floor = rel_err(c.grad, b.grad)
        parity("g_%s_vs_bf16_emulation" % name, a.grad, c.grad, TOL, floor=floor, floor_slack=1.0)
        parity("g_%s_vs_fp32_oracle" % name, a.grad, b.grad, TOL, floor=floor)
    # frozen-weight mode must give the same input gradients (same arithmetic, shorter program)
    m.requires_grad_(False)
    gin2 = [t.to(DEV).requires_grad_() for t in (xyz, vd, shp, tex)]
    sig3, rgbs3 = m(*gin2)
    assert torch.equal(sig3, sig2) and torch.equal(rgbs3, rgbs2)
    ((sig3 * up_s.to(DEV)).sum() + (rgbs3 * up_c.to(DEV)).sum()).backward()
    for a, b in zip(gin2, gin):
        assert parity_ok("a_grad", a.grad, b.grad, 1e-4)


def test_tc_render_c1_full_size_end_to_end():
    """Config 1 at full size (64x64 rays x 64 samples, CodeNeRF() defaults) through NeRFRenderer.render_rays in bf16
    mode: renders and pose / latent gradients against the fp32 oracle (14 tiles per CTA: ring wrap, TMEM ping-pong)."""
    S = snb()
    obj = oracle.synthetic_object(31, im_sz=64)
    sd = oracle.init_codenerf_state(seed=31)
    shp, tex = oracle.synthetic_latents(31, 1)
    jit = torch.rand(4096, 64, generator=torch.Generator().manual_seed(31))
    cam_o = obj["cam_pose"].clone().requires_grad_()
    s_o, t_o = shp.clone().requires_grad_(), tex.clone().requires_grad_()
    rgb_o, dep_o, acc_o, _ = oracle.render_rays_box(sd, obj["K"], cam_o, obj["wlh"], obj["roi"], 64, 64, s_o, t_o, jit)
    tgt, occ = obj["img"].reshape(-1, 3), obj["mask_occ"].reshape(-1, 1)
    oracle.refine_losses(rgb_o, acc_o, tgt, occ)[0].backward()
    m = model_from_state(S.CodeNeRF, sd)
    m.precision = "bf16"
    m.requires_grad_(False)
    cam = obj["cam_pose"].to(DEV).requires_grad_()
    s_g, t_g = shp.to(DEV).requires_grad_(), tex.to(DEV).requires_grad_()
    R = S.renderer.NeRFRenderer(n_samples=64)
    with forced_rand_like(jit):
        rgb, dep, acc, tg, oc = R.render_rays(m, DEV, obj["img"], obj["mask_occ"], cam, obj["wlh"], obj["K"].to(DEV), obj["roi"],
                                              s_g, t_g, im_sz=64)
    oracle.refine_losses(rgb, acc, tg, oc)[0].backward()
    errs = dict(rgb=parity("rgb", rgb, rgb_o, TOL), depth=parity("depth", dep, dep_o, TOL), acc=parity("acc", acc, acc_o, TOL),
                g_pose=parity("g_pose", cam.grad, cam_o.grad, TOL), g_shape=parity("g_shape", s_g.grad, s_o.grad, TOL),
                g_texture=parity("g_texture", t_g.grad, t_o.grad, TOL))
    print(errs)


def test_tc_training_mode_through_fused_render():
    """NeRFRenderer.render_rays with trainable weights in bf16 mode (fused render, training precision selected automatically):
    weight gradients against the fp32 back end of the same library on the same inputs."""
    S = snb()
    obj = oracle.synthetic_object(37, im_sz=32)
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=37)
    shp0, tex0 = oracle.synthetic_latents(37, 1)
    jit = torch.rand(1024, 64, generator=torch.Generator().manual_seed(37))
    R = S.renderer.NeRFRenderer(n_samples=64)
    grads = {}
    for prec in ("fp32", "bf16"):
        m = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
        m.precision = prec
        cam = obj["cam_pose"].to(DEV).requires_grad_()
        shp, tex = shp0.to(DEV).requires_grad_(), tex0.to(DEV).requires_grad_()
        with forced_rand_like(jit):
            rgb, dep, acc, tgt, occ = R.render_rays(m, DEV, obj["img"], obj["mask_occ"], cam, obj["wlh"], obj["K"].to(DEV), obj["roi"],
                                                    shp, tex, im_sz=32)
        oracle.refine_losses(rgb, acc, tgt, occ)[0].backward()
        grads[prec] = {k: p_.grad.clone() for k, p_ in m.named_parameters() if p_.grad is not None}
        grads[prec].update(cam=cam.grad, shp=shp.grad, tex=tex.grad)
    assert set(grads["bf16"]) == set(grads["fp32"]) and len(grads["bf16"]) >= 28 + 3
    errs = {k: parity("g_" + k, grads["bf16"][k], grads["fp32"][k], TOL) for k in grads["fp32"]}
    print(errs)


def test_tc_fused_render_any_ray_count():
    """With miss-ray compaction the decoder's row count is padded to the 128-row tile on the device, so the fused bf16 box render
    accepts ray counts whose N * S is not a multiple of 128 (n_rays = 333 random rays x 50 samples); checked against the fp32
    back end on the same rays."""
    S = snb()
    obj = oracle.synthetic_object(43, im_sz=32)
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=43)
    shp0, tex0 = oracle.synthetic_latents(43, 1)
    jit = torch.rand(333, 50, generator=torch.Generator().manual_seed(43))
    R = S.renderer.NeRFRenderer(n_samples=50)
    res = {}
    for prec in ("fp32", "bf16"):
        m = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
        m.precision = prec
        m.requires_grad_(False)
        cam = obj["cam_pose"].to(DEV).requires_grad_()
        shp, tex = shp0.to(DEV).requires_grad_(), tex0.to(DEV).requires_grad_()
        np.random.seed(3)
        with forced_rand_like(jit):
            rgb, dep, acc, tgt, occ = R.render_rays(m, DEV, obj["img"], obj["mask_occ"], cam, obj["wlh"], obj["K"].to(DEV), obj["roi"],
                                                    shp, tex, im_sz=32, n_rays=333)
        assert rgb.shape == (333, 3)
        oracle.refine_losses(rgb, acc, tgt, occ)[0].backward()
        res[prec] = [rgb, dep, acc, cam.grad, shp.grad, tex.grad]
    errs = [parity(n_, a, b, TOL) for n_, a, b in zip(("rgb", "depth", "acc", "g_pose", "g_shape", "g_texture"), res["bf16"], res["fp32"])]
    print(errs)


@pytest.mark.parametrize("im,S_,label", [(128, 64, "C2 object"), (512, 128, "C4 object")])
def test_full_size_properties_bf16(im, S_, label):
    """BASELINE configs at FULL size on one GPU (C2: 128x128 rays x 64 samples; C4: 512x512 rays x 128 samples = 33.5 M samples):
    size-independent properties -- hit mask bit-exact with the oracle's slab test, miss rays render the reference's closed form
    (acc = 1, depth = diag/2), transmittance in [0, 1], and renders + pose/latent gradients of a strided ray subsample equal
    to the fp32 oracle within the bf16 budget (the gradients of the loss restricted to the subsample)."""
    S = snb()
    obj = oracle.synthetic_object(71, im_sz=im)
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=71)
    m = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
    m.precision = "bf16"
    m.requires_grad_(False)
    shp0, tex0 = oracle.synthetic_latents(71, 1)
    n = im * im
    torch.manual_seed(71)
    torch.cuda.manual_seed(71)
    jit_dev = torch.rand(n, S_, device=DEV)          # drawn on the device (33.5 M values for C4); the oracle gets the same rows
    R = S.renderer.NeRFRenderer(n_samples=S_)
    cam = obj["cam_pose"].to(DEV).requires_grad_()
    shp, tex = shp0.to(DEV).requires_grad_(), tex0.to(DEV).requires_grad_()
    with forced_rand_like(jit_dev):
        rgb, dep, acc, tgt, occ = R.render_rays(m, DEV, obj["img"], obj["mask_occ"], cam, obj["wlh"], obj["K"].to(DEV), obj["roi"], shp, tex,
                                                im_sz=im)
    ids = torch.arange(0, n, max(n // 256, 1))
    sel = torch.zeros(n, 1, device=DEV)
    sel[ids.to(DEV)] = 1.0
    oracle.refine_losses(rgb, acc, tgt, occ * sel)[0].backward()      # loss over the subsample only
    # oracle on the subsample
    cam_o = obj["cam_pose"].clone().requires_grad_()
    s_o, t_o = shp0.clone().requires_grad_(), tex0.clone().requires_grad_()
    o_rgb, o_dep, o_acc, o_hit = oracle.render_rays_box(sd, obj["K"], cam_o, obj["wlh"], obj["roi"], im, S_, s_o, t_o, jit_dev[ids.to(DEV)].cpu(),
                                                        ray_ids=ids)
    tg, oc = obj["img"].reshape(-1, 3)[ids], obj["mask_occ"].reshape(-1, 1)[ids]
    oracle.refine_losses(o_rgb, o_acc, tg, oc)[0].backward()
    errs = dict(rgb=parity("rgb", rgb[ids.to(DEV)], o_rgb, TOL, rows=n * S_), depth=parity("depth", dep[ids.to(DEV)], o_dep, TOL),
                acc=parity("acc", acc[ids.to(DEV)], o_acc, TOL), g_pose=parity("g_pose", cam.grad, cam_o.grad, TOL),
                g_shape=parity("g_shape", shp.grad, s_o.grad, TOL), g_texture=parity("g_texture", tex.grad, t_o.grad, TOL))
    print(label, errs)
    # hit mask of ALL rays, bit-exact (slab test on the CPU oracle)
    ro, vd = oracle.get_rays(obj["K"], obj["cam_pose"], obj["roi"], uv_steps=[im, im])
    diag, half = oracle.box_constants(obj["wlh"])
    hb = torch.from_numpy(half)
    _, _, hit = oracle.ray_box_intersection(ro / (diag / 2), vd, -hb.expand_as(ro), hb.expand_as(ro))
    rays_o, viewdir = S.utils.get_rays(obj["K"].to(DEV), obj["cam_pose"].to(DEV), obj["roi"], uv_steps=[im, im])
    with forced_rand_like(jit_dev):
        hit_gpu = R.prepare_sampled_rays(rays_o, viewdir, obj["wlh"])[3]
    assert torch.equal(hit_gpu.cpu(), hit)
    miss = ~hit
    assert 0 < int(miss.sum()) < n
    assert torch.all(acc.detach().cpu()[miss] == 1.0)
    assert torch.allclose(dep.detach().cpu()[miss], torch.full((int(miss.sum()),), float(diag) / 2), rtol=1e-5)
    a = acc.detach()
    assert bool(((a >= 0) & (a <= 1.0 + 1e-6)).all()) and bool(torch.isfinite(rgb).all())


def test_fused_render_empty_ray_set():
    """n_rays = 0 (an empty random subset): every entry point returns empty tensors and zero gradients without launching."""
    S = snb()
    obj = oracle.synthetic_object(5, im_sz=16)
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=5)
    for prec in ("fp32", "bf16"):
        m = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
        m.precision = prec
        m.requires_grad_(False)
        shp0, tex0 = oracle.synthetic_latents(5, 1)
        cam = obj["cam_pose"].to(DEV).requires_grad_()
        shp, tex = shp0.to(DEV).requires_grad_(), tex0.to(DEV).requires_grad_()
        R = S.renderer.NeRFRenderer(n_samples=64)
        rgb, dep, acc, tgt, occ = R.render_rays(m, DEV, obj["img"], obj["mask_occ"], cam, obj["wlh"], obj["K"].to(DEV), obj["roi"], shp, tex,
                                                im_sz=16, n_rays=0)
        assert rgb.shape == (0, 3) and dep.shape == (0,) and acc.shape == (0,) and tgt.shape == (0, 3)
        (rgb.sum() + acc.sum()).backward()
        assert float(cam.grad.abs().sum()) == 0.0 and float(shp.grad.abs().sum()) == 0.0


def test_render_full_img_bf16_ragged_rows():
    """NeRFRenderer.render_full_img (renderer.py:238-294: row-chunked decode, chunk = roi side) in bf16 mode on a crop whose
    chunk rows x samples is not a multiple of the 128-row tile (37 x 50): padded internally; image equals the fp32 back end."""
    S = snb()
    obj = oracle.synthetic_object(47, im_sz=8)
    r = obj["roi"].clone()
    roi = torch.tensor([int(r[0]), int(r[1]), int(r[0]) + 37, int(r[1]) + 37], dtype=r.dtype)
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=47)
    shp, tex = oracle.synthetic_latents(47, 1)
    imgs = {}
    for prec in ("fp32", "bf16"):
        m = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
        m.precision = prec
        m.requires_grad_(False)
        R = S.renderer.NeRFRenderer(n_samples=50)
        torch.manual_seed(1); torch.cuda.manual_seed(1)
        with torch.no_grad():
            img, dep = R.render_full_img(m, DEV, obj["cam_pose"].to(DEV), obj["wlh"], obj["K"].to(DEV), roi, shp.to(DEV), tex.to(DEV),
                                         out_depth=True)
        assert img.shape == (37, 37, 3) and dep.shape == (37, 37)
        imgs[prec] = (img, dep)
    assert parity_ok("imgs_bf16_0", imgs["bf16"][0], imgs["fp32"][0], TOL) and parity_ok("imgs_bf16_1", imgs["bf16"][1], imgs["fp32"][1], TOL)


@pytest.mark.parametrize("B,n,S_", [(1, 150, 64), (4, 64, 16), (1, 16, 8), (3, 96, 8)])
def test_cta_group2_kernels_equal_cta_group1_kernels(B, n, S_):
    """The cta_group::2 decoder kernels (one M = 256 MMA over the CTA pair, half weight stages per SM; the default wherever every
    object owns a multiple of 256 rows) against the cta_group::1 kernels on the same inputs: same MMAs in the same K order, so
    sigma / rgb and the per-sample gradients must agree to the last bits; the latent gradients are sums of atomics (order differs).
    Covers an odd number of tile pairs (75), several objects per launch (256- and 768-row objects: super tiles never straddle
    two objects) and a launch of a single tile (half of one CTA pair's super tile)."""
    S = snb()
    lib = S._lib.load()
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=11)
    xyz, vd, shp, tex, up_s, up_c = _case(B, n, S_, 40 + B)
    m = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
    m.precision = "bf16"
    m.requires_grad_(False)
    outs = []
    try:
        for mode in (0, 1):
            lib.snb_tc_set_cg2(m._handle(xyz.to(DEV).device).h, mode)
            ins = [t.to(DEV).requires_grad_() for t in (xyz, vd, shp, tex)]
            sig, rgbs = m(*ins)
            ((sig * up_s.to(DEV)).sum() + (rgbs * up_c.to(DEV)).sum()).backward()
            torch.cuda.synchronize()
            outs.append((sig.detach(), rgbs.detach(), [t.grad for t in ins]))
    finally:
        lib.snb_tc_set_cg2(m._handle(xyz.to(DEV).device).h, -1)
    (s0, c0, g0), (s1, c1, g1) = outs
    assert torch.equal(s0, s1) and torch.equal(c0, c1)
    assert parity_ok("g1_0", g1[0], g0[0], 1e-6) and parity_ok("g1_1", g1[1], g0[1], 1e-6)      # d xyz, d viewdir: per sample, no reduction
    assert parity_ok("g1_2", g1[2], g0[2], 1e-5) and parity_ok("g1_3", g1[3], g0[3], 1e-5)      # latent gradients: atomically accumulated sums


@pytest.mark.parametrize("n_obj,im,S_,fused", [(3, 32, 64, False), (2, 16, 16, False), (5, 24, 128, False), (3, 32, 64, True), (2, 16, 16, True),
                                               (5, 24, 128, True)])
def test_batched_render_equals_per_object_render(n_obj, im, S_, fused):
    """NeRFRenderer.render_rays_batch + losses.refine_loss_batch (csrc/render_batch.cu: ONE launch set for all objects, compositing
    on the compact rows) against the per-object fused render of the same library: hit masks and hit-ray renders bit-identical (same
    per-row arithmetic), miss rays to 1e-6 (their single sample in closed form instead of S samples an ulp of z apart), gradients to
    summation order -- and against the fp32 CPU oracle within the bf16 budget.  fused = the north star's K1: the decoder's forward
    computes every row's stratified sample from its ray (no sampler kernel, no coordinates in HBM on the forward path)."""
    S = snb()
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=61)
    m = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
    m.precision = "bf16"
    m.requires_grad_(False)
    R = S.renderer.NeRFRenderer(n_samples=S_)
    objs = [oracle.synthetic_object(200 + 7 * i, im_sz=im) for i in range(n_obj)]
    lat = [oracle.synthetic_latents(200 + 7 * i, 1) for i in range(n_obj)]
    n = im * im
    jit = torch.rand(n_obj, n, S_, generator=torch.Generator().manual_seed(61))
    # per-object reference path of the same library
    per = []
    for i, o in enumerate(objs):
        cam = o["cam_pose"].to(DEV).requires_grad_()
        shp, tex = lat[i][0].to(DEV).requires_grad_(), lat[i][1].to(DEV).requires_grad_()
        with forced_rand_like(jit[i]):
            rgb, dep, acc, tgt, occ = R.render_rays(m, DEV, o["img"], o["mask_occ"], cam, o["wlh"], o["K"].to(DEV), o["roi"], shp, tex, im_sz=im)
        loss = S.losses.refine_loss(rgb, acc, tgt, occ, 0.1)[0]
        loss.backward()
        per.append((rgb.detach(), dep.detach(), acc.detach(), loss.detach(), cam.grad, shp.grad, tex.grad))
    # batched
    cams = torch.stack([o["cam_pose"] for o in objs]).to(DEV).requires_grad_()
    shps = torch.cat([l[0] for l in lat]).to(DEV).requires_grad_()
    texs = torch.cat([l[1] for l in lat]).to(DEV).requires_grad_()
    rgb, dep, acc, tgt, occ = R.render_rays_batch(m, DEV, [o["img"] for o in objs], [o["mask_occ"] for o in objs], cams,
                                                  [o["wlh"] for o in objs], torch.stack([o["K"] for o in objs]), [o["roi"] for o in objs],
                                                  shps, texs, im_sz=im, jitter=jit.to(DEV), fused_sampler=fused)
    losses, parts = S.losses.refine_loss_batch(rgb, acc, tgt, occ, 0.1)
    losses.sum().backward()
    assert rgb.shape == (n_obj, n, 3) and dep.shape == (n_obj, n) and losses.shape == (n_obj,)
    for i, o in enumerate(objs):
        ro, vd = oracle.get_rays(o["K"], o["cam_pose"], o["roi"], uv_steps=[im, im])
        diag, half = oracle.box_constants(o["wlh"])
        hb = torch.from_numpy(half)
        _, _, hit = oracle.ray_box_intersection(ro / (diag / 2), vd, -hb.expand_as(ro), hb.expand_as(ro))
        hit = hit.to(DEV)
        assert 0 < int(hit.sum()) < n
        for a, b_ in zip((rgb[i], dep[i], acc[i]), per[i][:3]):
            assert torch.equal(a[hit], b_[hit])                      # hit rays: the same bits
        parity("obj%d_miss_rays_rgb" % i, rgb[i][~hit], per[i][0][~hit], 1e-6)
        parity("obj%d_miss_rays_depth" % i, dep[i][~hit], per[i][1][~hit], 1e-6)
        assert bool((acc[i][~hit] == 1.0).all())
        parity("obj%d_loss" % i, losses[i], per[i][3], 1e-6)
        parity("obj%d_g_shape" % i, shps.grad[i], per[i][5][0], 1e-4)
        parity("obj%d_g_texture" % i, texs.grad[i], per[i][6][0], 1e-4)
        parity("obj%d_g_pose" % i, cams.grad[i], per[i][4], 1e-3)
    # object 0 against the fp32 CPU oracle (bf16 budget)
    o = objs[0]
    cam_o = o["cam_pose"].clone().requires_grad_()
    s_o, t_o = lat[0][0].clone().requires_grad_(), lat[0][1].clone().requires_grad_()
    rgb_o, dep_o, acc_o, _ = oracle.render_rays_box(sd, o["K"], cam_o, o["wlh"], o["roi"], im, S_, s_o, t_o, jit[0])
    oracle.refine_losses(rgb_o, acc_o, o["img"].reshape(-1, 3), o["mask_occ"].reshape(-1, 1))[0].backward()
    parity("obj0_rgb_vs_oracle", rgb[0], rgb_o, TOL)
    parity("obj0_depth_vs_oracle", dep[0], dep_o, TOL)
    parity("obj0_g_pose_vs_oracle", cams.grad[0], cam_o.grad, TOL)
    parity("obj0_g_shape_vs_oracle", shps.grad[0], s_o.grad[0], TOL)
    parity("obj0_g_texture_vs_oracle", texs.grad[0], t_o.grad[0], TOL)


def test_packed_weight_staleness_and_single_input_gradient():
    """(a) weight edits through ``p.data`` do not bump autograd's version counter, so the packed bf16 images stay as they
    were until ``model.invalidate_packed()``; (b) a backward whose forward's weights were re-packed in between fails loudly instead
    of mixing old masks with new W^T images; (c) asking for d xyz WITHOUT d viewdir (or the reverse) works in bf16 mode."""
    S = snb()
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=3)
    m = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
    m.precision = "bf16"
    m.requires_grad_(False)
    xyz, vd, shp, tex, _, _ = _case(1, 32, 16, 3)
    xyz, vd, shp, tex = xyz.to(DEV), vd.to(DEV), shp.to(DEV), tex.to(DEV)
    with torch.no_grad():
        s0, c0 = m(xyz, vd, shp, tex)
        m.rgb[2].weight.data.mul_(2.0)          # the head is read from the live fp32 pointer: visible at once
        m.encoding_shape.weight.data.mul_(1.5)     # a tensor-core layer: its packed image is stale until invalidated
        s1, c1 = m(xyz, vd, shp, tex)
        assert torch.equal(s1, s0)
        m.invalidate_packed()
        s2, c2 = m(xyz, vd, shp, tex)
        assert not torch.equal(s2, s0)
    # (c) one input gradient only
    x = xyz.clone().requires_grad_()
    sig, rgbs = m(x, vd, shp, tex)
    (sig.sum() + rgbs.sum()).backward()
    gx = x.grad.clone()
    x2, v2 = xyz.clone().requires_grad_(), vd.clone().requires_grad_()
    sig, rgbs = m(x2, v2, shp, tex)
    (sig.sum() + rgbs.sum()).backward()
    assert torch.equal(gx, x2.grad)
    # (b) re-pack between a forward and its backward
    m2 = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
    m2.precision = "bf16"
    x3 = xyz.clone().requires_grad_()
    sig, rgbs = m2(x3, vd, shp, tex)
    with torch.no_grad():
        m2.shape_layer_1[0].weight.mul_(1.01)      # an optimiser-style in-place update (bumps the version)
    m2(xyz, vd, shp, tex)                           # second forward re-packs
    with pytest.raises(RuntimeError):
        (sig.sum() + rgbs.sum()).backward()


def test_batched_render_full_c2_size_16_objects():
    """configs[1] at FULL size through ONE launch set (16 objects x 128x128 rays x 64 samples = the bench's step): two of the objects
    against their per-object renders (hit rays bit-identical, gradients to summation order), one against the fp32 CPU oracle on a
    strided ray subsample (bf16 budget), every object's hit mask against the oracle's slab test, losses finite."""
    S = snb()
    B, im, S_ = 16, 128, 64
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=0)
    m = model_from_state(S.AutoRFMix, sd, 3, 1, 256)
    m.precision = "bf16"
    m.requires_grad_(False)
    R = S.renderer.NeRFRenderer(n_samples=S_)
    objs = [oracle.synthetic_object(100 + i, im_sz=im) for i in range(B)]
    lat = [oracle.synthetic_latents(100 + i, 1) for i in range(B)]
    n = im * im
    torch.manual_seed(5)
    torch.cuda.manual_seed(5)
    jit = torch.rand(B, n, S_, device=DEV)
    cams = torch.stack([o["cam_pose"] for o in objs]).to(DEV).requires_grad_()
    shps = torch.cat([l[0] for l in lat]).to(DEV).requires_grad_()
    texs = torch.cat([l[1] for l in lat]).to(DEV).requires_grad_()
    batch = R.make_batch(DEV, torch.stack([o["img"] for o in objs]), torch.stack([o["mask_occ"] for o in objs]), [o["wlh"] for o in objs],
                         torch.stack([o["K"] for o in objs]), [o["roi"] for o in objs], im)
    rgb, dep, acc = R.render_batch(m, batch, cams, shps, texs, jitter=jit)
    # loss over a strided subsample of object 3's rays only (so that its gradients can be compared with the oracle's)
    ids = torch.arange(0, n, n // 256)
    sel = torch.zeros(B, n, 1, device=DEV)
    sel[3, ids.to(DEV)] = 1.0
    losses, parts = S.losses.refine_loss_batch(rgb, acc, batch.rgb_tgt, batch.occ_pixels * sel, 0.1)
    losses[3].backward()
    assert bool(torch.isfinite(parts).all())
    o = objs[3]
    cam_o = o["cam_pose"].clone().requires_grad_()
    s_o, t_o = lat[3][0].clone().requires_grad_(), lat[3][1].clone().requires_grad_()
    o_rgb, o_dep, o_acc, _ = oracle.render_rays_box(sd, o["K"], cam_o, o["wlh"], o["roi"], im, S_, s_o, t_o, jit[3][ids.to(DEV)].cpu(), ray_ids=ids)
    oracle.refine_losses(o_rgb, o_acc, o["img"].reshape(-1, 3)[ids], o["mask_occ"].reshape(-1, 1)[ids])[0].backward()
    parity("obj3_rgb_vs_oracle", rgb[3][ids.to(DEV)], o_rgb, TOL, rows=B * n * S_)
    parity("obj3_depth_vs_oracle", dep[3][ids.to(DEV)], o_dep, TOL)
    parity("obj3_g_pose_vs_oracle", cams.grad[3], cam_o.grad, TOL)
    parity("obj3_g_shape_vs_oracle", shps.grad[3], s_o.grad[0], TOL)
    parity("obj3_g_texture_vs_oracle", texs.grad[3], t_o.grad[0], TOL)
    assert float(cams.grad[0].abs().max()) == 0.0 and float(shps.grad[7].abs().max()) == 0.0   # objects outside the loss: zero gradients
    for i in (0, 9):   # per-object render of the same library on the same jitter
        with forced_rand_like(jit[i]):
            r1, d1, a1, _, _ = R.render_rays(m, DEV, objs[i]["img"], objs[i]["mask_occ"], objs[i]["cam_pose"].to(DEV), objs[i]["wlh"],
                                             objs[i]["K"].to(DEV), objs[i]["roi"], lat[i][0].to(DEV), lat[i][1].to(DEV), im_sz=im)
        ro, vd = oracle.get_rays(objs[i]["K"], objs[i]["cam_pose"], objs[i]["roi"], uv_steps=[im, im])
        diag, half = oracle.box_constants(objs[i]["wlh"])
        hb = torch.from_numpy(half)
        _, _, hit = oracle.ray_box_intersection(ro / (diag / 2), vd, -hb.expand_as(ro), hb.expand_as(ro))
        hit = hit.to(DEV)
        assert torch.equal(rgb[i].detach()[hit], r1.detach()[hit]) and torch.equal(dep[i].detach()[hit], d1.detach()[hit])
        parity("obj%d_miss_rays_rgb" % i, rgb[i].detach()[~hit], r1.detach()[~hit], 1e-6)
        assert bool((acc[i].detach()[~hit] == 1.0).all())


def test_tc_training_mode_c5_size_weight_grads_vs_fp32_oracle():
    """bf16 TRAINING mode at config C5's real size (8 objects x 1024 rays x 64 samples = 524 288 rows through one decoder call +
    volume_rendering_batch + the trainer's losses, trainer_unified_nuscenes.py:120-146): EVERY weight / bias gradient and the code
    gradients against the fp32 CPU ORACLE (not the repo's own fp32 back end), bf16 budget."""
    S = snb()
    B, n, S_ = 8, 1024, 64
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=5)
    g = torch.Generator().manual_seed(5)
    xyz = (torch.rand(B, n, S_, 3, generator=g) - 0.5) * 1.2
    vd = torch.nn.functional.normalize(torch.randn(B, n, 1, 3, generator=g), dim=-1).repeat(1, 1, S_, 1)
    z = torch.rand(B, S_, generator=g).sort(-1).values * 4 + 8
    tgt, occ = torch.rand(B, n, 3, generator=g), torch.randint(-1, 2, (B, n, 1), generator=g).float()
    shp0, tex0 = oracle.synthetic_latents(5, B)

    def losses_of(sig, rgbs, zv, tg, oc, comp):
        rgb, dep, acc = comp(sig.reshape(B, n, S_, 1), rgbs.reshape(B, n, S_, 3), zv)
        den = torch.sum(torch.abs(oc), dim=[-2, -1]) + 1e-9
        l_rgb = (torch.sum((rgb - tg) ** 2 * torch.abs(oc), dim=[-2, -1]) / den).mean()
        l_occ = (torch.sum(torch.exp(-oc * (0.5 - acc.unsqueeze(-1))) * torch.abs(oc), dim=[-2, -1]) / den).mean()
        return l_rgb + 0.1 * l_occ
    # fp32 CPU oracle
    sdr = {k: v.clone().requires_grad_() for k, v in sd.items()}
    s_o, t_o = shp0.clone().requires_grad_(), tex0.clone().requires_grad_()
    sig, rgbs = oracle.codenerf_decoder(sdr, xyz.reshape(-1, S_, 3), vd.reshape(-1, S_, 3), s_o, t_o)
    losses_of(sig, rgbs, z, tgt, occ, lambda a, b, c: oracle.composite(a.squeeze(-1), b, c.unsqueeze(1).expand(B, n, S_), False)).backward()
    # bf16 training mode on the GPU
    m = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
    m.precision = "bf16"
    s_g, t_g = shp0.to(DEV).requires_grad_(), tex0.to(DEV).requires_grad_()
    sig2, rgbs2 = m(xyz.reshape(-1, S_, 3).to(DEV), vd.reshape(-1, S_, 3).to(DEV), s_g, t_g)
    losses_of(sig2, rgbs2, z.to(DEV), tgt.to(DEV), occ.to(DEV), S.utils.volume_rendering_batch).backward()
    parity("g_shapecode", s_g.grad, s_o.grad, TOL, rows=B * n * S_)
    parity("g_texturecode", t_g.grad, t_o.grad, TOL)
    for k, p_ in m.named_parameters():
        parity("gw_" + k, p_.grad, sdr[k].grad, TOL)


def test_class_default_decoder_5_5_blocks_on_the_tensor_cores():
    """SUPNeRF()'s class defaults (5 shape + 5 texture blocks, 842 880 MAC / sample) are beyond the two-tile kernels' 4 latent slots:
    the one-tile tcgen05 kernels take them -- per-object fused render with miss-ray compaction and the batched launch set alike --
    in bf16 and (frozen weights) in the fp32-grade split precision.  Batched vs per-object: the same per-row arithmetic (hit rays
    bit-identical); against the fp32 CPU oracle within each mode's budget."""
    S = snb()
    n_obj, im, S_ = 2, 16, 32
    sd = oracle.init_codenerf_state(shape_blocks=5, texture_blocks=5, seed=77)
    objs = [oracle.synthetic_object(260 + 7 * i, im_sz=im) for i in range(n_obj)]
    lat = [oracle.synthetic_latents(260 + 7 * i, 1) for i in range(n_obj)]
    n = im * im
    jit = torch.rand(n_obj, n, S_, generator=torch.Generator().manual_seed(77))
    R = S.renderer.NeRFRenderer(n_samples=S_)
    for prec, tol in (("bf16", TOL), ("fp32", 1e-5)):
        m = model_from_state(S.SUPNeRF, sd, 5, 5, 3, 3, 256)
        m.precision = prec
        m.requires_grad_(False)
        per = []
        for i, o in enumerate(objs):
            cam = o["cam_pose"].to(DEV).requires_grad_()
            shp, tex = lat[i][0].to(DEV).requires_grad_(), lat[i][1].to(DEV).requires_grad_()
            with forced_rand_like(jit[i]):
                rgb, dep, acc, tgt, occ = R.render_rays(m, DEV, o["img"], o["mask_occ"], cam, o["wlh"], o["K"].to(DEV), o["roi"], shp, tex, im_sz=im)
            S.losses.refine_loss(rgb, acc, tgt, occ, 0.1)[0].backward()
            per.append((rgb.detach(), dep.detach(), cam.grad, shp.grad))
        cams = torch.stack([o["cam_pose"] for o in objs]).to(DEV).requires_grad_()
        shps = torch.cat([l[0] for l in lat]).to(DEV).requires_grad_()
        texs = torch.cat([l[1] for l in lat]).to(DEV).requires_grad_()
        rgb, dep, acc, tgt, occ = R.render_rays_batch(m, DEV, [o["img"] for o in objs], [o["mask_occ"] for o in objs], cams,
                                                      [o["wlh"] for o in objs], torch.stack([o["K"] for o in objs]), [o["roi"] for o in objs],
                                                      shps, texs, im_sz=im, jitter=jit.to(DEV))
        S.losses.refine_loss_batch(rgb, acc, tgt, occ, 0.1)[0].sum().backward()
        for i, o in enumerate(objs):
            ro, vd = oracle.get_rays(o["K"], o["cam_pose"], o["roi"], uv_steps=[im, im])
            diag, half = oracle.box_constants(o["wlh"])
            hb = torch.from_numpy(half)
            hit = oracle.ray_box_intersection(ro / (diag / 2), vd, -hb.expand_as(ro), hb.expand_as(ro))[2].to(DEV)
            assert torch.equal(rgb[i][hit], per[i][0][hit]) and torch.equal(dep[i][hit], per[i][1][hit])
            parity("%s_obj%d_g_shape_vs_per_object" % (prec, i), shps.grad[i], per[i][3][0], 1e-4)
            cam_o = o["cam_pose"].clone().requires_grad_()
            s_o, t_o = lat[i][0].clone().requires_grad_(), lat[i][1].clone().requires_grad_()
            rgb_o, dep_o, acc_o, _ = oracle.render_rays_box(sd, o["K"], cam_o, o["wlh"], o["roi"], im, S_, s_o, t_o, jit[i])
            oracle.refine_losses(rgb_o, acc_o, o["img"].reshape(-1, 3), o["mask_occ"].reshape(-1, 1))[0].backward()
            parity("%s_obj%d_rgb_vs_oracle" % (prec, i), rgb[i], rgb_o, tol)
            parity("%s_obj%d_depth_vs_oracle" % (prec, i), dep[i], dep_o, tol)
            parity("%s_obj%d_g_shape_vs_oracle" % (prec, i), shps.grad[i], s_o.grad[0], tol if prec == "bf16" else 1e-4)


def test_graphed_batch_step_equals_the_call_by_call_step():
    """renderer.GraphedBatchStep (H2D of the inputs from pinned host buffers, batched render, losses, backward, D2H of the result as one
    CUDA graph) against the same step issued call by call, under the same jitter; and it follows the CONTENT of the host buffers."""
    S = snb()
    n_obj, im, S_ = 3, 16, 32
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=91)
    m = model_from_state(S.SUPNeRF, sd, 3, 1, 3, 3, 256)
    m.precision = "bf16"
    m.requires_grad_(False)
    R = S.renderer.NeRFRenderer(n_samples=S_)
    objs = [oracle.synthetic_object(300 + 7 * i, im_sz=im) for i in range(n_obj)]
    lat = [oracle.synthetic_latents(300 + 7 * i, 1) for i in range(n_obj)]
    pin = lambda t: t.contiguous().float().pin_memory()   # noqa: E731
    h_img, h_mask = pin(torch.stack([o["img"] for o in objs])), pin(torch.stack([o["mask_occ"] for o in objs]))
    h_cam, h_K = pin(torch.stack([o["cam_pose"] for o in objs])), pin(torch.stack([o["K"] for o in objs]))
    h_shp, h_tex = pin(torch.cat([l[0] for l in lat])), pin(torch.cat([l[1] for l in lat]))
    wlhs, rois = [o["wlh"] for o in objs], [o["roi"] for o in objs]
    jit = torch.rand(n_obj, im * im, S_, generator=torch.Generator().manual_seed(91)).to(DEV)
    step = S.renderer.GraphedBatchStep(R, m, DEV, h_img, h_mask, h_cam, wlhs, h_K, rois, h_shp, h_tex, im_sz=im, jitter=jit)

    def eager():
        cam, shp, tex = [t.to(DEV).requires_grad_() for t in (h_cam, h_shp, h_tex)]
        rgb, dep, acc, tgt, occ = R.render_rays_batch(m, DEV, h_img, h_mask, cam, wlhs, h_K, rois, shp, tex, im_sz=im, jitter=jit)
        loss, parts = S.losses.refine_loss_batch(rgb, acc, tgt, occ, 0.1)
        loss.backward(gradient=torch.ones(n_obj, device=DEV))
        return torch.cat([parts[:, :1], cam.grad.reshape(n_obj, 12), shp.grad, tex.grad], 1).cpu()
    for rep in range(2):
        got = step.run()
        torch.cuda.synchronize()
        ref = eager()
        parity("rep%d_loss" % rep, got[:, 0], ref[:, 0], 1e-6)
        parity("rep%d_g_pose" % rep, got[:, 1:13], ref[:, 1:13], 1e-4)
        parity("rep%d_g_codes" % rep, got[:, 13:], ref[:, 13:], 1e-5)
        assert float(got[:, 0].abs().min()) > 0
        h_cam[:, :, 3] += 0.05          # new content in the same pinned buffers: the next replay must see it
        h_shp.mul_(0.9)
