"""Multi-GPU sharding of the render hot path.  One process per GPU (``torch.distributed``; NCCL over
NVLink on the B200 box, gloo in the CPU tests).  The reference has no equivalent: it only wraps the model in
``nn.DataParallel`` (trainer_nerf_nuscenes.py:24-60), so nothing here mirrors a reference signature.

* **object-parallel** (configs C2/C3): objects are independent (own pose, latents, rays, loss, optimiser state):
  rank r takes objects r, r+G, ...  No collective on the data path.
* **ray-sharded** (config C4): one object, contiguous ray tiles per rank, weights / pose / latents replicated.  Loss
  denominators depend only on the input mask, so every rank evaluates the GLOBAL denominator locally; each rank's partial
  ``d cam_pose (12) | d shapecode (D) | d texturecode (D) | loss (1)`` go into ONE flat fp32 buffer and ONE
  ``all_reduce(sum)``; afterwards every rank holds identical gradients and applies the identical optimiser step.
"""
import math

import numpy as np
import torch
import torch.distributed as dist

from . import utils as U

TILE_ROWS = 128  # the tcgen05 decoder works on 128-sample tiles; shards keep ray*sample rows tile-aligned


def object_shard(n_objects, rank, world):
    """Indices of the objects rank `rank` owns: r, r+G, r+2G, ... (object-parallel)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_objects, world))


def ray_shard(n_rays, n_samples, rank, world):
    """[start, end) of the contiguous ray tile of `rank`.  Tile boundaries are multiples of 128/gcd(128, S) rays so every
    shard's N*S sample rows are a whole number of 128-row decoder tiles (except the tail shard, which takes the rest)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    align = TILE_ROWS // math.gcd(TILE_ROWS, int(n_samples))
    units = -(-n_rays // align)                      # ceil
    per = -(-units // world)
    start = min(rank * per * align, n_rays)
    end = min((rank + 1) * per * align, n_rays)
    return start, end


def ray_shard_indices(n_rays, n_samples, rank, world, device=None, tile=TILE_ROWS):
    """Ray ids of `rank` in the INTERLEAVED layout: the rays are cut into tiles of 128 rays (128 x S sample rows = S whole
    128-row decoder tiles for any S) and rank r owns tiles r, r+G, r+2G, ...
    With miss-ray compaction a rank's decoder work is proportional to the rays of its shard that HIT the object box, and a
    contiguous split of a row-major crop gives the ranks holding the middle rows about twice the hits of the outer ones
    (512x512 rays on 4 GPUs: 12.1 ms per step against 7.8 ms for a balanced split); interleaved tiles sample the crop uniformly.
    `tile` = rays per tile (default 128; 1 = ray i on rank i % G, "strided": neighbouring pixels go to different ranks, so the hit
    counts -- hence the decoder rows after compaction -- agree to a few rays; miss-ray compaction makes tile alignment of the
    shards irrelevant, the decoder pads each shard's rows itself).
    -> sorted int64 tensor of ray ids (the last tile may be partial)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    tile = max(1, int(tile))
    n_tiles = -(-int(n_rays) // tile)
    if rank >= n_tiles:
        return torch.empty(0, dtype=torch.int64, device=device)
    mine = torch.arange(rank, n_tiles, world, dtype=torch.int64, device=device)
    ids = (mine[:, None] * tile + torch.arange(tile, dtype=torch.int64, device=device)[None, :]).reshape(-1)
    return ids[ids < n_rays]


def refine_loss_sharded(rgb_rays, acc_rays, rgb_tgt, occ_pixels, occ_all, loss_occ_coef=0.1):
    """This rank's PARTIAL of the refine loss (optimizer_nuscenes.py:729-736): local numerators over the GLOBAL
    denominator sum|occ_all| + 1e-9 (a function of the input mask only, so every rank evaluates it locally), so that the
    partials of all ranks add up to the single-GPU loss.  Runs the fused loss kernel (losses.refine_loss); CUDA only."""
    from . import losses
    den = torch.sum(torch.abs(occ_all)) + 1e-9
    return losses.refine_loss(rgb_rays, acc_rays, rgb_tgt, occ_pixels, loss_occ_coef, den=den)[0]


def allreduce_grads(params, loss=None, group=None, scale=None, collective=True, comm=None):
    """Sum the gradients of `params` (cam_pose, shapecode, texturecode, ...) and, if given, the partial loss over all ranks
    with ONE all_reduce of one flat fp32 buffer (2.1 KB for 12 + 256 + 256 + 1 floats): one gather kernel (torch.cat) in, the
    collective, and the parameters' ``.grad`` become VIEWS of the reduced buffer (no copy back).  Returns the global loss
    (a 0-dim tensor) or None."""
    flat = [_flat_view(p.grad if p.grad is not None else torch.zeros_like(p)).float() for p in params]
    if loss is not None:
        flat.append(loss.detach().reshape(1).float())
    buf = torch.cat(flat)
    if collective and comm is not None and comm.world > 1:
        comm.allreduce_(buf)        # the C-ABI collective (snb_allreduce_grads) on this process' own NCCL communicator
    elif collective and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    if scale is not None:
        buf.mul_(scale)          # one kernel over the flat buffer (data-parallel averaging)
    off = 0
    for p in params:
        n = p.numel()
        g = _unflat_view(buf[off:off + n], p)
        p.grad = g if g.dtype == p.dtype else g.to(p.dtype)
        off += n
    return buf[off] if loss is not None else None


def allreduce_weight_grads(model, group=None, average=True):
    """Data-parallel training: the batch is split by objects over the ranks, every rank back-propagates
    its own objects, then ONE all_reduce(sum) over the flat buffer of all weight gradients (decoder: 0.7-1.1 M floats) and a
    scale by 1/G -- what DistributedDataParallel's single bucket would do.  Per-instance latent codes stay rank-local (each
    object lives on exactly one rank).  Parameters without a gradient are skipped consistently on every rank (they must be the
    same set everywhere)."""
    params = [p for p in model.parameters() if p.grad is not None]
    if not params:
        return 0
    g = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    allreduce_grads(params, None, group, scale=(1.0 / g) if (average and g > 1) else None)
    return sum(p.numel() for p in params)


class NcclComm:
    """An NCCL communicator of this process' own (ncclCommInitRank over the ranks of `group`, the unique id broadcast through
    torch.distributed), for the C-ABI all-reduce ``snb_allreduce_grads(handle, ncclComm_t, flat, n, stream)``: the collective then is
    ONE call into libsupnerf_b200.so on the caller's stream, with no torch.distributed bookkeeping on the step's critical path.
    Uses the libnccl.so.2 torch already loaded."""

    def __init__(self, device, group=None):
        import ctypes
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("NcclComm needs an initialised torch.distributed process group (for the rendezvous)")
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        try:
            self._nccl = ctypes.CDLL("libnccl.so.2")
        except OSError:
            import glob
            import os
            cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2"))
            if not cands:
                raise
            self._nccl = ctypes.CDLL(cands[0])

        class UniqueId(ctypes.Structure):
            _fields_ = [("internal", ctypes.c_byte * 128)]
        uid = UniqueId()
        if self.rank == 0:
            rc = self._nccl.ncclGetUniqueId(ctypes.byref(uid))
            if rc != 0:
                raise RuntimeError("ncclGetUniqueId failed: %d" % rc)
        buf = torch.frombuffer(bytearray(bytes(uid.internal)), dtype=torch.uint8).clone().to(self.device)
        dist.broadcast(buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(buf.cpu().numpy().tobytes())
        ctypes.memmove(ctypes.byref(uid), raw, 128)
        self.comm = ctypes.c_void_p()
        self._nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
        with torch.cuda.device(self.device):
            rc = self._nccl.ncclCommInitRank(ctypes.byref(self.comm), self.world, uid, self.rank)
        if rc != 0:
            raise RuntimeError("ncclCommInitRank failed: %d" % rc)

    def allreduce_(self, flat, handle=None):
        """In-place sum of the flat fp32 CUDA tensor over the ranks, on the current stream (snb_allreduce_grads)."""
        from . import _lib
        from ._lib import check, on_device, ptr, stream_ptr
        if flat.dtype != torch.float32 or not flat.is_cuda or not flat.is_contiguous():
            raise ValueError("NcclComm.allreduce_: a contiguous fp32 CUDA tensor is required")
        lib = _lib.load()
        with on_device(flat.device):
            check(lib.snb_allreduce_grads(handle, self.comm, ptr(flat), flat.numel(), stream_ptr()), "snb_allreduce_grads")
        return flat

    def destroy(self):
        import ctypes
        if getattr(self, "comm", None):
            self._nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
            self._nccl.ncclCommDestroy(self.comm)
            self.comm = None


def _flat_view(g):
    """1-D view of a dense gradient in MEMORY order (no copy): contiguous tensors as they are, channels-last 4-D tensors through
    their NHWC permutation -- the same on every rank, so the element order of a flat bucket matches across ranks."""
    if g.is_contiguous():
        return g.reshape(-1)
    if g.dim() == 4 and g.is_contiguous(memory_format=torch.channels_last):
        return g.permute(0, 2, 3, 1).reshape(-1)
    return g.contiguous().reshape(-1)


def _unflat_view(flat, like):
    """The inverse of _flat_view: a view of `flat` with the shape AND strides of `like` (fused optimisers require them equal)."""
    if like.is_contiguous():
        return flat.view(like.shape)
    if like.dim() == 4 and like.is_contiguous(memory_format=torch.channels_last):
        n, c, h, w = like.shape
        return flat.view(n, h, w, c).permute(0, 3, 1, 2)
    return flat.view(like.shape)


class BucketedGradAllReduce:
    """Data-parallel gradient exchange of the joint training step (config C5) OVERLAPPED with the rest of the backward pass.
    The parameters are cut into buckets in the order their gradients become ready (caller-supplied lists: e.g. decoder + pose head
    + encoder heads first, the three ``layer4`` branches -- 80 % of the 49 M parameters -- next, the encoder trunk last).  A
    post-accumulate hook counts a bucket's gradients; when the last one has arrived the bucket is gathered into one flat fp32
    buffer (ONE torch.cat), all-reduced (sum) and scaled by 1/G on a side stream while autograd keeps producing the remaining
    buckets on the main stream.  ``finish()`` makes the main stream wait for the side stream and re-points every ``p.grad`` at its
    slice of the reduced buffer (views, no copy).  World size 1: no collective, gradients untouched."""

    def __init__(self, buckets, group=None, average=True):
        self.buckets = [[p for p in b if p.requires_grad] for b in buckets]
        self.buckets = [b for b in self.buckets if b]
        self.group, self.average = group, average
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self._bucket_of, self._pending, self._flat = {}, [0] * len(self.buckets), [None] * len(self.buckets)
        self._handles = []
        dev = self.buckets[0][0].device
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.events = []
        for i, b in enumerate(self.buckets):
            for p in b:
                self._bucket_of[p] = i
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook))
        self.reset()

    def reset(self):
        self._pending = [len(b) for b in self.buckets]
        self._flat = [None] * len(self.buckets)
        self.events = []

    def _hook(self, p):
        i = self._bucket_of[p]
        self._pending[i] -= 1
        if self._pending[i] == 0 and self.world > 1:
            self._launch(i)

    def _launch(self, i):
        b = self.buckets[i]
        if self.stream is not None:
            main = torch.cuda.current_stream(b[0].device)
            self.stream.wait_stream(main)
            ctx = torch.cuda.stream(self.stream)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            e0 = torch.cuda.Event(enable_timing=True) if self.stream is not None else None
            if e0 is not None:
                e0.record()
            flat = torch.cat([_flat_view(p.grad) for p in b])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                flat.mul_(1.0 / self.world)
            if e0 is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                self.events.append((i, e0, e1))
            for p in b:
                if self.stream is not None:
                    p.grad.record_stream(self.stream)
            self._flat[i] = flat

    def finish(self):
        """After backward(): wait for the exchanges, point the gradients at the reduced buffers.  -> floats exchanged."""
        if self.world == 1:
            self.reset()
            return 0
        for i, n in enumerate(self._pending):
            if n == 0 and self._flat[i] is None:
                self._launch(i)
            elif n != 0:   # a parameter of this bucket got no gradient this step: exchange what exists, consistently on every rank
                raise RuntimeError("bucket %d: %d parameters received no gradient" % (i, n))
        if self.stream is not None:
            torch.cuda.current_stream(self.buckets[0][0].device).wait_stream(self.stream)
        total = 0
        for b, flat in zip(self.buckets, self._flat):
            off = 0
            for p in b:
                n = p.numel()
                p.grad = _unflat_view(flat[off:off + n], p)
                off += n
            total += off
        events = self.events
        self.reset()
        self.last_events = events
        return total

    def allreduce_ms(self):
        """Per-bucket duration of gather + all-reduce + scale on the side stream of the last finished step (after a synchronize)."""
        return {i: e0.elapsed_time(e1) for i, e0, e1 in getattr(self, "last_events", [])}

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def render_rays_sharded(renderer, model, device, img, mask_occ, cam_pose, obj_sz, K, roi, shapecode, texturecode, im_sz=64,
                        rank=None, world=None, layout="contiguous", jitter="torch", seed=None):
    """Ray-sharded NeRFRenderer.render_rays (renderer.py:117-167 semantics, `n_rays=None`): renders only this rank's rays --
    one contiguous ray tile (layout="contiguous", ray_shard) or every G-th 128-ray tile (layout="interleaved",
    ray_shard_indices: balanced decoder work under miss-ray compaction).  The (N,S) jitter is drawn in full on every rank
    (same generator state => same numbers as the single-GPU call) and sliced, so the union of the shards is bit-identical to
    the unsharded render.  jitter="counter": only the shard's rows are filled, with the counter-based draw of ops.jitter_fill
    (a pure function of (seed, ray id, sample): the union of the shards equals the world=1 call with the same seed bit for bit;
    `seed` defaults to one draw from torch's CPU generator, identical on every rank when the ranks seed it identically).
    -> rgb, depth, acc, rgb_tgt, occ_pixels of the shard, plus occ_all (N,1) for the global loss denominator."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    device = torch.device(device)
    px, py = U._pixel_grid_on(device, roi, [im_sz, im_sz])
    img, mask_occ = U._resize_targets(img, mask_occ, im_sz)
    n = px.numel()
    occ_all = mask_occ.reshape(-1, 1).to(device, non_blocking=True)
    if layout in ("interleaved", "strided"):
        sel = ray_shard_indices(n, renderer.n_samples, rank, world, device=device, tile=TILE_ROWS if layout == "interleaved" else 1)
    elif layout == "contiguous":
        a, b = ray_shard(n, renderer.n_samples, rank, world)
        sel = slice(a, b)
    else:
        raise ValueError("layout must be 'contiguous' or 'interleaved'")
    if jitter == "torch":
        jit = torch.rand_like(torch.empty(n, renderer.n_samples, device=device))[sel].contiguous()
    elif jitter == "counter":
        from . import ops
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        ids = sel if torch.is_tensor(sel) else torch.arange(sel.start, sel.stop, dtype=torch.int64, device=device)
        jit = ops.jitter_fill(seed, ids.numel(), renderer.n_samples, device, ray_ids=ids)
    else:
        raise ValueError("jitter must be 'torch' or 'counter'")
    rgb_tgt = img.reshape(-1, 3).to(device, non_blocking=True)[sel]
    rgb, dep, acc = renderer._render_fused(model, device, px[sel].contiguous(), py[sel].contiguous(), K, cam_pose, obj_sz, shapecode,
                                           texturecode, jitter=jit)
    return rgb, dep, acc, rgb_tgt, occ_all[sel], occ_all


class RayShard:
    """One rank's share of a ray-sharded object, prepared ONCE (config C4's refine-style loop re-renders the same crop every
    step): the shard's pixel coordinates, targets and mask, the global loss denominator and the flat gradient buffer are built
    here, so a step is   jitter rows (counter-based) -> fused render of the shard -> partial loss over the global denominator ->
    backward -> ONE all_reduce(sum) of [d cam_pose | d shapecode | d texturecode | loss]   with no per-step index arithmetic."""

    def __init__(self, renderer, model, device, img, mask_occ, obj_sz, K, roi, im_sz, rank=None, world=None, layout="interleaved",
                 group=None, comm=None):
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.renderer, self.model, self.device, self.group = renderer, model, torch.device(device), group
        self.comm = comm      # optional NcclComm: the gradient all-reduce then goes through the C ABI (snb_allreduce_grads)
        self.obj_sz, self.K, self.layout = obj_sz, K.to(self.device), layout
        dev = self.device
        px, py = U._pixel_grid_on(dev, roi, [im_sz, im_sz])
        img, mask_occ = U._resize_targets(img, mask_occ, im_sz)
        self.n_rays, S = px.numel(), renderer.n_samples
        if layout == "interleaved":
            self.ids = ray_shard_indices(self.n_rays, S, self.rank, self.world, device=dev)
        elif layout == "strided":
            self.ids = ray_shard_indices(self.n_rays, S, self.rank, self.world, device=dev, tile=1)
        elif layout == "contiguous":
            a, b = ray_shard(self.n_rays, S, self.rank, self.world)
            self.ids = torch.arange(a, b, dtype=torch.int64, device=dev)
        else:
            raise ValueError("layout must be 'contiguous', 'interleaved' or 'strided'")
        occ_all = mask_occ.reshape(-1, 1).to(dev)
        self.den = torch.sum(torch.abs(occ_all)) + 1e-9          # global denominator: a function of the input mask only
        self.px, self.py = px[self.ids].contiguous(), py[self.ids].contiguous()
        self.rgb_tgt = img.reshape(-1, 3).to(dev)[self.ids].contiguous()
        self.occ = occ_all[self.ids].contiguous()
        self.flat = None

    def step(self, cam_pose, shapecode, texturecode, seed, loss_occ_coef=0.1, events=None):
        """-> (global loss, rgb, depth, acc of the shard); afterwards cam_pose.grad / shapecode.grad / texturecode.grad hold the
        all-reduced gradients (views of one flat buffer).  `events`: optional callable(name) the bench uses to mark phases."""
        from . import losses, ops
        mark = events if events is not None else (lambda name: None)
        jit = ops.jitter_fill(seed, self.ids.numel(), self.renderer.n_samples, self.device, ray_ids=self.ids)
        mark("jitter")
        rgb, dep, acc = self.renderer._render_fused(self.model, self.device, self.px, self.py, self.K, cam_pose, self.obj_sz, shapecode,
                                                    texturecode, jitter=jit)
        part = losses.refine_loss(rgb, acc, self.rgb_tgt, self.occ, loss_occ_coef, den=self.den)[0]
        mark("forward")
        cam_pose.grad = shapecode.grad = texturecode.grad = None
        part.backward()
        mark("backward")
        # a RayShard built with world=1 (the one-GPU anchor of a multi-rank run) must not enter a collective the other ranks skip
        loss = allreduce_grads([cam_pose, shapecode, texturecode], part, self.group, collective=self.world > 1, comm=self.comm)
        mark("allreduce")
        return loss, rgb, dep, acc


def gather_rays(t, n_rays, n_samples, group=None, layout="contiguous"):
    """Optional: assemble the full (N, ...) image from the per-rank shards (an all_gather of 20 B/ray; only when the caller
    wants the whole render on every rank).  `layout` as in render_rays_sharded."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    if layout in ("interleaved", "strided"):
        ids = [ray_shard_indices(n_rays, n_samples, r, world, device=t.device, tile=TILE_ROWS if layout == "interleaved" else 1) for r in range(world)]
        counts = [int(i.numel()) for i in ids]
    else:
        sizes = [ray_shard(n_rays, n_samples, r, world) for r in range(world)]
        counts = [b - a for a, b in sizes]
    width = max(counts)
    pad = torch.zeros((width,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    if layout in ("interleaved", "strided"):
        full = torch.empty((n_rays,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        for o, i, c in zip(outs, ids, counts):
            full[i] = o[:c]
        return full
    return torch.cat([o[:c] for o, c in zip(outs, counts)], 0)
