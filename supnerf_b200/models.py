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
        if layers is None:   # the Linear modules are fixed after construction; their Parameters are read fresh every call
            layers = self._decoder_layers()
            self.__dict__["_layer_cache"] = layers
        out = []
        for lin in layers:
            out += [lin.weight, lin.bias]
        return out

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
    def __init__(self, shape_blocks=5, texture_blocks=5, pose_blocks=3, regress_blocks=3, latent_dim=256, pose_dim=16,
                 num_xyz_freq=10, num_dir_freq=4, norm_layer_type='BatchNorm2d', pose_shortcut=False, pred_wlh=False):
        super().__init__()
        self._init_common(shape_blocks, texture_blocks, latent_dim, latent_dim, num_xyz_freq, num_dir_freq)
        self._build(shape_blocks, texture_blocks, latent_dim, latent_dim, num_xyz_freq, num_dir_freq)
        self.pose_blocks, self.regress_blocks = pose_blocks, regress_blocks
        self.pose_shortcut, self.pred_wlh = pose_shortcut, pred_wlh


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
