#!/usr/bin/env python
"""Config C3: the SUP-NeRF test-time pose/shape/texture refinement loop of optimizer_nuscenes.py:684-769
(`optimize_objs_w_pose_unified`): per object 50 iterations of  axis-angle -> cam2opt -> utils.render_rays_v2 -> the two losses ->
backward -> AdamW step on (shapecode, texturecode, rot_vec, trans_vec)  with the lrs of jsonfiles/supnerf.nusc.vehicle.car.json
(0.02 / 0.02 / 0.01 / 0.01).  Objects are independent: rank r owns objects r, r+G, ... (object-parallel, no collective).
Reports ms per refine iteration.  The pytorch3d axis-angle map sits one step upstream of the path (unpinned):
this harness supplies its own Rodrigues formula.
  python tools/refine_bench.py [--objects 32] [--iters 50] [--im 32]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/refine_bench.py"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import supnerf_b200 as snb  # noqa: E402
from supnerf_b200 import synthetic  # noqa: E402


def axis_angle_to_matrix(v):
    """Rodrigues: R = I + sin(t)/t [v]x + (1 - cos(t))/t^2 [v]x^2  (what pytorch3d.transforms.axis_angle_to_matrix evaluates)."""
    t = torch.sqrt((v * v).sum() + 1e-20)
    zero = torch.zeros((), device=v.device)
    Kx = torch.stack([torch.stack([zero, -v[2], v[1]]), torch.stack([v[2], zero, -v[0]]), torch.stack([-v[1], v[0], zero])])
    return torch.eye(3, device=v.device) + torch.sin(t) / t * Kx + (1 - torch.cos(t)) / (t * t) * (Kx @ Kx)


def matrix_to_axis_angle(R):
    R = R.double()
    cos = ((torch.trace(R) - 1) / 2).clamp(-1, 1)
    t = torch.acos(cos)
    w = torch.stack([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return (w / (2 * torch.sin(t).clamp_min(1e-12)) * t).float()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--objects", type=int, default=32)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--im", type=int, default=32)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--streams", type=int, default=1, help="with --graph: refine this rank's objects side by side over this many CUDA streams (refine.run_objects)")
    ap.add_argument("--batch", action="store_true", help="supnerf_b200.refine.BatchRefiner: this rank's objects as ONE launch set per iteration (one CUDA graph)")
    ap.add_argument("--eager", action="store_true", help="with --batch: do not capture (one launch per kernel: for ncu launch lists)")
    ap.add_argument("--graph", action="store_true", help="supnerf_b200.refine.ObjectRefiner: one CUDA graph per iteration, no host sync")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    sd = synthetic.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=3)
    model = snb.SUPNeRF(3, 1, 3, 3, 256)
    model.load_state_dict(sd)
    model = model.to(dev)
    model.precision = a.precision
    model.requires_grad_(False)
    mine = snb.parallel.object_shard(a.objects, rank, world)
    objs = []
    for i in mine:
        o = synthetic.synthetic_object(300 + i, im_sz=a.im)
        s, t = synthetic.synthetic_latents(300 + i, 1)
        # the loop optimises the OBJECT pose (opt_cam_pose false): cam2opt = [R^T | -R^T t]
        c2o = o["cam_pose"]
        R_obj = c2o[:, :3].t().contiguous()
        t_obj = -(R_obj @ c2o[:, 3:])
        o.update(shapecode=s.to(dev).requires_grad_(), texturecode=t.to(dev).requires_grad_(),
                 rot_vec=matrix_to_axis_angle(R_obj).to(dev).requires_grad_(), trans_vec=t_obj.reshape(3).to(dev).requires_grad_(),
                 img=o["img"].to(dev), mask=o["mask_occ"].to(dev), K=o["K"].to(dev),
                 obj_diag=np.linalg.norm(o["wlh"]).astype(np.float32))
        objs.append(o)

    def refine(o, iters):
        opt = torch.optim.AdamW([{"params": o["shapecode"], "lr": 0.02}, {"params": o["texturecode"], "lr": 0.02},
                                 {"params": o["rot_vec"], "lr": 0.01}, {"params": o["trans_vec"], "lr": 0.01}])
        loss = None
        for _ in range(iters):
            opt.zero_grad()
            rot = axis_angle_to_matrix(o["rot_vec"]).t()                       # optimizer_nuscenes.py:684-699
            cam2opt = torch.cat((rot, -rot @ o["trans_vec"].unsqueeze(-1)), dim=-1)
            rgb, dep, acc, tgt, occ = snb.utils.render_rays_v2(model, dev, o["img"], o["mask"], cam2opt, o["obj_diag"], o["K"], o["roi"],
                                                              a.samples, o["shapecode"], o["texturecode"], 1, 0, im_sz=a.im, n_rays=None)
            loss = snb.losses.refine_loss(rgb, acc, tgt, occ, 0.1)[0]          # :729-736
            loss.backward()
            opt.step()
        return loss

    if a.graph or a.batch:
        refiners = []
        for o in objs:
            torch.manual_seed(1000 + len(refiners))
            refiners.append(snb.refine.ObjectRefiner(model, dev, o["img"], o["mask"], o["K"], o["roi"], o["obj_diag"], o["shapecode"],
                                                     o["texturecode"], o["rot_vec"], o["trans_vec"], n_samples=a.samples, im_sz=a.im,
                                                     max_iters=a.iters + 8))
            if a.graph:
                refiners[-1].capture()

        def refine(o, iters):   # noqa: F811
            return refiners[[id(x) for x in objs].index(id(o))].run(iters)[0]

    l0 = float(snb.losses.refine_loss(*[t for t in _first_render(model, dev, objs[0], a)], 0.1)[0]) if objs else 0.0
    if not a.graph:
        for o in objs[:1]:
            refine(o, 3)   # warm-up (packs the weights, fills the allocator)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if a.batch:
        bat = snb.refine.BatchRefiner(refiners)
        if not a.eager:
            bat.capture()      # warms up and restores the initial state itself
        else:
            bat.run(3)
        torch.cuda.synchronize()
    e0.record()
    if a.batch:
        last = list(bat.run(a.iters)[:, 0])
    elif a.graph and a.streams > 1:
        last = [l[0] for l in snb.refine.run_objects(refiners, a.iters, a.streams)]
    else:
        last = [refine(o, a.iters) for o in objs]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        n_local = len(objs)
        print(json.dumps({"config": "C3: SUP-NeRF refine loop, %d objects over %d GPU(s) (object-parallel), %d iterations, render %dx%d rays x %d samples"
                                    % (a.objects, world, a.iters, a.im, a.im, a.samples),
                          "n_gpus": world, "ms_total": round(ms, 2), "ms_per_refine_iteration": round(ms / max(n_local * a.iters, 1), 4),
                          "objects_per_gpu": n_local, "rays_per_s": round(a.objects * a.iters * a.im * a.im / (ms / 1e3), 1),
                          "loss_first_object": {"before": round(l0, 5), "after": round(float(last[0].detach()), 5)}, "precision": a.precision,
                          "mode": ("one CUDA graph per iteration (refine.ObjectRefiner)" + (", objects side by side over %d streams (refine.run_objects)" % a.streams if a.streams > 1 else "")) if a.graph else "refine.BatchRefiner: one launch set per iteration for the rank's objects" if a.batch else "eager reference-API loop (render_rays_v2 + torch AdamW)"}))
    if world > 1:
        torch.distributed.destroy_process_group()


def _first_render(model, dev, o, a):
    with torch.no_grad():
        rot = axis_angle_to_matrix(o["rot_vec"]).t()
        cam2opt = torch.cat((rot, -rot @ o["trans_vec"].unsqueeze(-1)), dim=-1)
        st = torch.random.get_rng_state()
        rgb, dep, acc, tgt, occ = snb.utils.render_rays_v2(model, dev, o["img"], o["mask"], cam2opt, o["obj_diag"], o["K"], o["roi"],
                                                          a.samples, o["shapecode"], o["texturecode"], 1, 0, im_sz=a.im, n_rays=None)
        torch.random.set_rng_state(st)
    return rgb, acc, tgt, occ


if __name__ == "__main__":
    main()
