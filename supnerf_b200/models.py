"""Drop-in decoder modules: same constructor arguments, parameter names (state_dict keys) and
``forward(xyz, viewdir, shape_latent, texture_latent) -> (sigmas (N,S,1), rgbs (N,S,3))`` as the
reference's CodeNeRF (model_codenerf.py:13-63), AutoRFMix (model_autorf.py:190-250), SUPNeRF
(model_supnerf.py:165-269, decoder half) and AutoRF (model_autorf.py:123-186).  ``forward`` runs the
sm_100a kernels through the C ABI; parameters stay ordinary ``nn.Parameter``s so checkpoints made by the
reference load with ``load_state_dict`` and optimisers/autograd see the usual ``.grad``s.

The image encoder / pose head of AutoRF*/SUPNeRF are outside this path (out of scope); their
checkpoint keys are accepted and ignored by ``load_state_dict(strict=False)``."""
import torch
import torch.nn as nn

from . import ops

_DEFAULT_PRECISION = "fp32"


def set_default_precision(p):
    """'fp32' (SIMT FFMA, 1e-5 parity mode) or 'bf16' (tcgen05 tensor-core mode, 2e-2)."""
    global _DEFAULT_PRECISION
    if p not in ops.PREC:
        raise ValueError(p)
    _DEFAULT_PRECISION = p


def get_default_precision():
    return _DEFAULT_PRECISION


class _DecoderBase(nn.Module):
    _arch = 0

    def _init_common(self, shape_blocks, texture_blocks, W, latent_dim, num_xyz_freq, num_dir_freq):
        self.shape_blocks = shape_blocks
        self.texture_blocks = texture_blocks
        self.num_xyz_freq = num_xyz_freq
        self.num_dir_freq = num_dir_freq
        self._W = W
        self._latent_dim = latent_dim
        self.precision = None  # None -> module default
        self._handles = {}

    def _decoder_layers(self):
        raise NotImplementedError

    def _weights(self):
        layers = self.__dict__.get("_layer_cache")
        # the Linear modules are fixed after construction; their Parameters are read fresh every call.  The cache is validated
        # against this module's own first layer: nn.DataParallel's replicas shallow-copy __dict__
        if layers is None or layers[0] is not self._modules["encoding_xyz"][0]:
            layers = self._decoder_layers()
            self.__dict__["_layer_cache"] = layers
        out = []
        for lin in layers:
            out += [lin.weight, lin.bias]
        return out

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica.__dict__.pop("_layer_cache", None)     # replicas own their parameters and their C handles
        replica.__dict__["_handles"] = {}
        return replica

    def invalidate_packed(self):
        """bf16 mode: re-pack the weight images at the next call (after ``p.data`` edits autograd's version counter cannot see)."""
        for h in self._handles.values():
            h.invalidate_packed()

    def _handle(self, device):
        key = (device.type, device.index)
        if key not in self._handles:
            self._handles[key] = ops.DecoderHandle(self._arch, self.shape_blocks, self.texture_blocks, self._W,
                                                   self._latent_dim, self.num_xyz_freq, self.num_dir_freq)
        return self._handles[key]

    def forward(self, xyz, viewdir, shape_latent, texture_latent):
        if not xyz.is_cuda:
            raise RuntimeError("supnerf_b200 decoders run on CUDA only (no CPU fallback)")
        lead = xyz.shape[:-1]
        n_rows = xyz.numel() // 3
        if n_rows % shape_latent.shape[0] != 0:
            raise ValueError("number of samples must be a multiple of the number of objects")
        prec = self.precision or _DEFAULT_PRECISION
        if self.encoding_xyz[0].weight.device != xyz.device:
            raise RuntimeError("decoder weights live on %s, inputs on %s" % (self.encoding_xyz[0].weight.device, xyz.device))
        sigma, rgb = ops.decoder(self._handle(xyz.device), prec, xyz.reshape(-1, 3), viewdir.reshape(-1, 3),
                                 shape_latent, texture_latent, self._weights())
        return sigma.reshape(*lead, 1), rgb.reshape(*lead, 3)

    # ---- checkpoints: the reference saves the WHOLE module (optimizer_nuscenes.py:1795-1796 loads saved['model_params'] with the
    # default strict=True); AutoRFMix / SUPNeRF checkpoints also carry the image encoder / pose head, which are not on this path.
    # Those entries are kept verbatim (and re-emitted by state_dict()) so that a reference checkpoint loads and round-trips.
    _OFFPATH_PREFIXES = ("img_encoder.", "pose_layer_", "regress_layer_", "out_delta_layer", "out_wlh_layer", "pose_encoder",
                         "encoder.", "wlh_layer")

    def load_state_dict(self, state_dict, strict=True, assign=False):
        own = set(super().state_dict().keys())
        mine, off = {}, {}
        for k, v in state_dict.items():
            if k in own or not k.startswith(self._OFFPATH_PREFIXES):
                mine[k] = v          # unknown keys outside the known off-path families still raise under strict=True
            else:
                off[k] = v.detach().clone() if torch.is_tensor(v) else v
        self.__dict__["_offpath_state"] = off
        return super().load_state_dict(mine, strict=strict, assign=assign)

    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        prefix = kwargs.get("prefix", args[1] if len(args) > 1 else "")
        for k, v in self.__dict__.get("_offpath_state", {}).items():
            sd[prefix + k] = v
        return sd

    def __deepcopy__(self, memo):  # handles are per-instance C objects
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_layer_cache":
                continue
            setattr(new, k, {} if k == "_handles" else copy.deepcopy(v, memo))
        return new


class _CodeNeRFFamily(_DecoderBase):
    _arch = 0

    def _build(self, shape_blocks, texture_blocks, W, latent_dim, num_xyz_freq, num_dir_freq):
        # registration order == reference order (so seeded init and state_dict order agree): model_codenerf.py:22-37
        d_xyz, d_viewdir = 3 + 6 * num_xyz_freq, 3 + 6 * num_dir_freq
        self.encoding_xyz = nn.Sequential(nn.Linear(d_xyz, W), nn.ReLU())
        for j in range(shape_blocks):
            setattr(self, f"shape_latent_layer_{j+1}", nn.Sequential(nn.Linear(latent_dim, W), nn.ReLU()))
            setattr(self, f"shape_layer_{j+1}", nn.Sequential(nn.Linear(W, W), nn.ReLU()))
        self.encoding_shape = nn.Linear(W, W)
        self.sigma = nn.Sequential(nn.Linear(W, 1), nn.Softplus())
        self.encoding_viewdir = nn.Sequential(nn.Linear(W + d_viewdir, W), nn.ReLU())
        for j in range(texture_blocks):
            setattr(self, f"texture_latent_layer_{j+1}", nn.Sequential(nn.Linear(latent_dim, W), nn.ReLU()))
            setattr(self, f"texture_layer_{j+1}", nn.Sequential(nn.Linear(W, W), nn.ReLU()))
        self.rgb = nn.Sequential(nn.Linear(W, W // 2), nn.ReLU(), nn.Linear(W // 2, 3))

    def _decoder_layers(self):
        ls = [self.encoding_xyz[0]]
        for j in range(1, self.shape_blocks + 1):
            ls += [getattr(self, f"shape_latent_layer_{j}")[0], getattr(self, f"shape_layer_{j}")[0]]
        ls += [self.encoding_shape, self.sigma[0], self.encoding_viewdir[0]]
        for j in range(1, self.texture_blocks + 1):
            ls += [getattr(self, f"texture_latent_layer_{j}")[0], getattr(self, f"texture_layer_{j}")[0]]
        ls += [self.rgb[0], self.rgb[2]]
        return ls


class CodeNeRF(_CodeNeRFFamily):
    def __init__(self, shape_blocks=2, texture_blocks=1, W=256, num_xyz_freq=10, num_dir_freq=4, latent_dim=256):
        super().__init__()
        self._init_common(shape_blocks, texture_blocks, W, latent_dim, num_xyz_freq, num_dir_freq)
        self._build(shape_blocks, texture_blocks, W, latent_dim, num_xyz_freq, num_dir_freq)


class AutoRFMix(_CodeNeRFFamily):
    def __init__(self, shape_blocks=5, texture_blocks=5, latent_dim=128, num_xyz_freq=10, num_dir_freq=4,
                 norm_layer_type='BatchNorm2d'):
        super().__init__()
        self._init_common(shape_blocks, texture_blocks, latent_dim, latent_dim, num_xyz_freq, num_dir_freq)
        self._build(shape_blocks, texture_blocks, latent_dim, latent_dim, num_xyz_freq, num_dir_freq)


class SUPNeRF(_CodeNeRFFamily):
    """model_supnerf.py:165-269.  The decoder is the package's kernel path; the pose estimator half (``img_encoder``,
    ``pose_layer_j`` / ``regress_layer_j`` / ``out_delta_layer``; SURVEY 8f rank 2, pose_estimator.py) is materialised on first
    use -- ``encode_img`` / ``pose_update`` / ``materialize_pose_estimator()`` -- or when a checkpoint carrying its keys is
    loaded and then used, so decoder-only callers (the refine loops of this path) do not build 49 M encoder parameters."""

    def __init__(self, shape_blocks=5, texture_blocks=5, pose_blocks=3, regress_blocks=3, latent_dim=256, pose_dim=16,
                 num_xyz_freq=10, num_dir_freq=4, norm_layer_type='BatchNorm2d', pose_shortcut=False, pred_wlh=False):
        super().__init__()
        self._init_common(shape_blocks, texture_blocks, latent_dim, latent_dim, num_xyz_freq, num_dir_freq)
        self._build(shape_blocks, texture_blocks, latent_dim, latent_dim, num_xyz_freq, num_dir_freq)
        self.pose_blocks, self.regress_blocks = pose_blocks, regress_blocks
        self.pose_shortcut, self.pred_wlh = pose_shortcut, pred_wlh
        self._pose_dim, self._norm_layer_type = pose_dim, norm_layer_type

    def has_pose_estimator(self):
        return "img_encoder" in self._modules

    def materialize_pose_estimator(self):
        """Build ``img_encoder`` + the pose head (same registration order and initialisers as model_supnerf.py:169-216) on the
        decoder's device; entries of a previously loaded reference checkpoint that belong to them are moved in."""
        if self.has_pose_estimator():
            return self
        from . import pose_estimator as pe
        norm = nn.InstanceNorm2d if self._norm_layer_type == "InstanceNorm2d" else nn.BatchNorm2d
        enc = pe.ImgEncoder((3, 4, 6, 3), num_classes=self._latent_dim, norm_layer=norm, pred_wlh=self.pred_wlh)
        self.img_encoder = enc
        mods = dict(self._modules)                       # the reference registers the encoder first (state_dict order)
        self._modules.clear()
        self._modules["img_encoder"] = mods.pop("img_encoder")
        self._modules.update(mods)
        pe.build_pose_head(self, self.pose_blocks, self.regress_blocks, self._latent_dim, self._pose_dim)
        ref = self.encoding_xyz[0].weight
        for name in ["img_encoder", "out_delta_layer"] + [f"pose_layer_{j}" for j in range(self.pose_blocks)] + \
                [f"regress_layer_{j}" for j in range(self.regress_blocks)]:
            self._modules[name].to(device=ref.device).train(self.training)
        off = self.__dict__.get("_offpath_state") or {}
        mine = {k: v for k, v in off.items() if k.startswith(("img_encoder.", "pose_layer_", "regress_layer_", "out_delta_layer"))}
        if mine:
            own = super(_DecoderBase, self).state_dict()
            missing = [k for k in own if k.startswith(tuple(set(k2.split(".")[0] + "." for k2 in mine))) and k not in mine and "num_batches_tracked" not in k]
            if missing:
                raise RuntimeError("checkpoint lacks pose-estimator entries: %s" % missing[:5])
            with torch.no_grad():
                for k, v in mine.items():
                    own[k].copy_(v)
                    del off[k]
        return self

    def encode_img(self, img):
        """model_supnerf.py:218-224 -> (shape_feat, texture_feat, pose_feat, box_uv_pred, box_wlh_pred | None)."""
        self.materialize_pose_estimator()
        out = self.img_encoder(img, self.pose_shortcut)
        return out if self.pred_wlh else out + (None,)

    def encode_img_fast(self, img):
        """``encode_img`` under bf16 autocast with channels-last activations (cuDNN tensor-core convolutions; the B200 way to run
        this library stage).  Heads return fp32."""
        self.materialize_pose_estimator()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = self.img_encoder(img.contiguous(memory_format=torch.channels_last), self.pose_shortcut)
        out = tuple(t.float() for t in out)
        return out if self.pred_wlh else out + (None,)

    def pose_update(self, im_feat, box_uv_src):
        """model_supnerf.py:226-239."""
        self.materialize_pose_estimator()
        from . import pose_estimator as pe
        return pe.pose_update(self, im_feat, box_uv_src)


class AutoRF(_DecoderBase):
    _arch = 1

    def __init__(self, shape_blocks=5, texture_blocks=5, latent_dim=128, num_xyz_freq=10, num_dir_freq=4,
                 norm_layer_type='BatchNorm2d'):
        super().__init__()
        self._init_common(shape_blocks, texture_blocks, latent_dim, latent_dim, num_xyz_freq, num_dir_freq)
        self.precision = "fp32"   # the AutoRF chain (W = 128, feature mixing after every layer) runs on the fp32 back end only
        d_xyz, d_viewdir = 3 + 6 * num_xyz_freq, 3 + 6 * num_dir_freq
        self.encoding_xyz = nn.Sequential(nn.Linear(d_xyz, latent_dim), nn.ReLU())
        for j in range(shape_blocks - 1):
            setattr(self, f"shape_layer_{j}", nn.Sequential(nn.Linear(latent_dim, latent_dim), nn.ReLU()))
        self.sigma = nn.Sequential(nn.Linear(latent_dim, 1), nn.Softplus())
        for j in range(texture_blocks - 2):
            setattr(self, f"texture_layer_{j}", nn.Sequential(nn.Linear(latent_dim, latent_dim), nn.ReLU()))
        setattr(self, f"texture_layer_{texture_blocks-2}", nn.Sequential(nn.Linear(latent_dim + d_viewdir, latent_dim), nn.ReLU()))
        self.rgb = nn.Sequential(nn.Linear(latent_dim + d_viewdir, 3), nn.Sigmoid())

    def _decoder_layers(self):
        ls = [self.encoding_xyz[0]]
        ls += [getattr(self, f"shape_layer_{j}")[0] for j in range(self.shape_blocks - 1)]
        ls += [self.sigma[0]]
        ls += [getattr(self, f"texture_layer_{j}")[0] for j in range(self.texture_blocks - 1)]
        ls += [self.rgb[0]]
        return ls
