"""Pose-estimator forward: the other half of SUP-NeRF's joint training step and the feed-forward stage
before every refine loop.  Mirrors, with the same names, argument meaning and ``state_dict`` keys:

* ``ImgEncoder`` (model_supnerf.py:16-152): ResNet-34 trunk (BasicBlock [3, 4, 6, 3]) with THREE ``layer4`` branches
  (shape / texture / pose), average pool, one linear head per branch and ``fc_uv`` (the 8 projected box corners regressed
  directly from the pose code); ``pose_shortcut`` subtracts the pose branch from the other two; ``pred_wlh`` adds a fourth branch;
* ``SUPNeRF.encode_img`` / ``SUPNeRF.pose_update`` (model_supnerf.py:218-239), attached to the drop-in ``SUPNeRF`` by
  ``models.py`` (materialised on first use so that decoder-only users do not pay for 49 M encoder parameters);
* ``corners_of_box_batch`` / ``view_points_batch`` / ``normalize_by_roi`` (utils.py:1032-1147, 1175-1197);
* ``pose_regress`` and ``joint_training_losses`` (trainer_unified_nuscenes.py:27-195: ``ParallelModel.forward``).

This row is NOT a hand-written kernel path: the convolutions run through cuDNN (``torch.nn.Conv2d``; bf16 autocast + channels-last
in ``encode_img_fast``) on the caller's stream -- library code, stated as such in DESIGN.md.  The render half of the joint step
(decoder + compositing + losses, forward and backward with every weight gradient) is the package's own tcgen05 path.
The axis-angle <-> matrix maps are pytorch3d's in the reference (``rot_trans``), a dependency that is not part of the reference
tree (parity unpinned there); ``refine.axis_angle_to_matrix`` / ``matrix_to_axis_angle_batch`` restate Rodrigues' formula."""
import torch
import torch.nn as nn


def _conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1, bias=False)


class _BasicBlock(nn.Module):
    """conv3x3-norm-relu-conv3x3-norm (+ optional 1x1 projection of the identity) -- the residual block of ResNet-18/34
    (He et al. 2015, fig. 5 left); attribute names follow torchvision's so that reference checkpoints load key for key."""
    expansion = 1

    def __init__(self, inplanes, planes, stride, downsample, norm_layer):
        super().__init__()
        self.conv1 = _conv3x3(inplanes, planes, stride)
        self.bn1 = norm_layer(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv3x3(planes, planes)
        self.bn2 = norm_layer(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + identity)


class ImgEncoder(nn.Module):
    """model_supnerf.py:16-152 with ``block = BasicBlock`` (the only block SUPNeRF instantiates, :172-176)."""

    def __init__(self, layers=(3, 4, 6, 3), num_classes=128, norm_layer=None, pred_wlh=False):
        super().__init__()
        norm_layer = norm_layer or nn.BatchNorm2d
        self._norm_layer = norm_layer
        self.inplanes = 64
        self.pred_wlh = pred_wlh
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = norm_layer(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(64, layers[0], 1)
        self.layer2 = self._make_layer(128, layers[1], 2)
        self.layer3 = self._make_layer(256, layers[2], 2)
        self.layer4_shape = self._make_layer(512, layers[3], 2)
        self.inplanes = 256          # each layer4 branch starts from layer3's 256 channels again (model_supnerf.py:55-58)
        self.layer4_texture = self._make_layer(512, layers[3], 2)
        self.inplanes = 256
        self.layer4_pose = self._make_layer(512, layers[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc_shape = nn.Linear(512, num_classes)
        self.fc_texture = nn.Linear(512, num_classes)
        self.fc_pose = nn.Linear(512, num_classes)
        self.fc_uv = nn.Linear(num_classes, 16)
        if pred_wlh:
            self.inplanes = 256
            self.layer4_wlh = self._make_layer(512, layers[3], 2)
            self.fc_wlh = nn.Sequential(nn.Linear(512, num_classes), nn.ReLU(), nn.Linear(num_classes, 3))
        for m in self.modules():    # model_supnerf.py:69-74
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, planes, blocks, stride):
        norm = self._norm_layer
        down = None
        if stride != 1 or self.inplanes != planes:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride, bias=False), norm(planes))
        seq = [_BasicBlock(self.inplanes, planes, stride, down, norm)]
        self.inplanes = planes
        seq += [_BasicBlock(planes, planes, 1, None, norm) for _ in range(1, blocks)]
        return nn.Sequential(*seq)

    def forward(self, x, pose_shortcut=False):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer3(self.layer2(self.layer1(x)))
        x_shape, x_texture, x_pose = self.layer4_shape(x), self.layer4_texture(x), self.layer4_pose(x)
        if pose_shortcut:
            x_shape = x_shape - x_pose
            x_texture = x_texture - x_pose
        f_shape = self.fc_shape(torch.flatten(self.avgpool(x_shape), 1))
        f_texture = self.fc_texture(torch.flatten(self.avgpool(x_texture), 1))
        f_pose = self.fc_pose(torch.flatten(self.avgpool(x_pose), 1))
        uv = self.fc_uv(f_pose)
        if self.pred_wlh:
            wlh = self.fc_wlh(torch.flatten(self.avgpool(self.layer4_wlh(x)), 1))
            return f_shape, f_texture, f_pose, uv, wlh
        return f_shape, f_texture, f_pose, uv


def build_pose_head(module, pose_blocks, regress_blocks, latent_dim, pose_dim=16):
    """pose_layer_j / regress_layer_j / out_delta_layer (model_supnerf.py:199-216), registered on `module` in the reference's order."""
    W = latent_dim
    setattr(module, "pose_layer_0", nn.Sequential(nn.Linear(pose_dim, W), nn.ReLU(inplace=True)))
    for j in range(1, pose_blocks):
        setattr(module, f"pose_layer_{j}", nn.Sequential(nn.Linear(W, W), nn.ReLU(inplace=True)))
    setattr(module, "regress_layer_0", nn.Sequential(nn.Linear(latent_dim + W, W), nn.ReLU(inplace=True)))
    for j in range(1, regress_blocks):
        setattr(module, f"regress_layer_{j}", nn.Sequential(nn.Linear(W, W), nn.ReLU(inplace=True)))
    module.out_delta_layer = nn.Linear(W, 6)


def pose_update(module, im_feat, box_uv_src):
    """model_supnerf.py:226-239."""
    pose_feat = module.pose_layer_0(box_uv_src)
    for j in range(1, module.pose_blocks):
        pose_feat = getattr(module, f"pose_layer_{j}")(pose_feat)
    delta = module.regress_layer_0(torch.cat([im_feat, pose_feat], -1))
    for j in range(1, module.regress_blocks):
        delta = getattr(module, f"regress_layer_{j}")(delta)
    return module.out_delta_layer(delta)


# ---------------------------------------------------------------------------------------------------------------------
# box-corner projection (utils.py:1032-1147, 1175-1197)
# ---------------------------------------------------------------------------------------------------------------------
_SIGNS = {}


def _corner_signs(device, dtype, is_kitti):
    key = (str(device), dtype, bool(is_kitti))
    s = _SIGNS.get(key)
    if s is None:
        x = [1, 1, 1, 1, -1, -1, -1, -1]
        if is_kitti:
            y, z = [-2, -2, 0, 0, -2, -2, 0, 0], [1, -1, -1, 1, 1, -1, -1, 1]
        else:
            y, z = [1, -1, -1, 1, 1, -1, -1, 1], [1, 1, -1, -1, 1, 1, -1, -1]
        s = torch.tensor([x, y, z], device=device, dtype=dtype)
        _SIGNS[key] = s
    return s


def corners_of_box_batch(obj_pose_batch, wlh_batch, is_kitti=False, scale=1.0):
    """utils.py:1110-1147 -> (B, 3, 8): the 8 corners of every box in the frame `obj_pose_batch` (B, 3, 4) maps into."""
    w, l, h = wlh_batch[:, 0], wlh_batch[:, 1], wlh_batch[:, 2]
    s = _corner_signs(wlh_batch.device, wlh_batch.dtype, is_kitti)
    ext = torch.stack([l, h, w] if is_kitti else [l, w, h], 1)                  # (B, 3): extent along x, y, z
    corners = ext.unsqueeze(-1) / 2 * s.unsqueeze(0) * scale                    # same evaluation order as the reference
    corners = torch.matmul(obj_pose_batch[:, :, :3], corners)
    return corners + obj_pose_batch[:, :, 3:4]


def view_points_batch(points, view, normalize):
    """utils.py:1032-1073: points (B, 3, n), view (B, 3|4, 3|4) -> (B, 3, n)."""
    assert view.shape[1] <= 4 and view.shape[2] <= 4 and points.shape[1] == 3
    bsize, n = view.shape[0], points.shape[2]
    viewpad = torch.eye(4, device=points.device, dtype=torch.float32).repeat(bsize, 1, 1)
    viewpad[:, :view.shape[1], :view.shape[2]] = view
    pts = torch.cat([points, torch.ones((bsize, 1, n), dtype=torch.float32, device=points.device)], dim=1)
    pts = torch.matmul(viewpad, pts)[:, :3, :]
    if normalize:
        pts = pts / pts[:, 2:3, :]
    return pts


def normalize_by_roi(pts_batch, roi_batch, need_square=True):
    """utils.py:1175-1197."""
    w = roi_batch[:, 2] - roi_batch[:, 0]
    h = roi_batch[:, 3] - roi_batch[:, 1]
    cx = (roi_batch[:, 2] + roi_batch[:, 0]) / 2
    cy = (roi_batch[:, 3] + roi_batch[:, 1]) / 2
    pts = pts_batch - torch.stack([cx, cy], 1).unsqueeze(-1)
    if need_square:
        dim = torch.maximum(w, h)
        return pts / dim.view(-1, 1, 1), dim
    return pts / torch.stack([w, h], 1).unsqueeze(-1), None


# ---------------------------------------------------------------------------------------------------------------------
# axis-angle maps (pytorch3d.transforms in the reference: not part of its tree, parity unpinned -- SURVEY §8c)
# ---------------------------------------------------------------------------------------------------------------------
def axis_angle_to_matrix_batch(v):
    """Rodrigues: R = I + sin(t)/t [v]x + (1 - cos t)/t^2 [v]x^2, batched (B, 3) -> (B, 3, 3)."""
    t = torch.sqrt((v * v).sum(-1) + 1e-20)
    z = torch.zeros_like(t)
    kx = torch.stack([torch.stack([z, -v[:, 2], v[:, 1]], -1), torch.stack([v[:, 2], z, -v[:, 0]], -1),
                      torch.stack([-v[:, 1], v[:, 0], z], -1)], -2)
    eye = torch.eye(3, device=v.device, dtype=v.dtype).expand_as(kx)
    a = (torch.sin(t) / t).view(-1, 1, 1)
    b = ((1 - torch.cos(t)) / (t * t)).view(-1, 1, 1)
    return eye + a * kx + b * (kx @ kx)


def matrix_to_axis_angle_batch(R):
    """(B, 3, 3) -> (B, 3): angle from the trace, axis from the antisymmetric part (angles away from pi)."""
    cos = ((R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2] - 1) / 2).clamp(-1, 1)
    t = torch.acos(cos)
    w = torch.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], -1)
    return w / (2 * torch.sin(t).clamp_min(1e-12)).unsqueeze(-1) * t.unsqueeze(-1)


def pose_regress(model, im_feat_batch, src_pose_batch, tgt_uv_batch, wlh_batch, roi_batch, K_batch, K_inv=None):
    """trainer_unified_nuscenes.py:150-195 -> (loss (B, 8), pred_pose_batch (B, 3, 4))."""
    src_uv = view_points_batch(corners_of_box_batch(src_pose_batch.detach(), wlh_batch), K_batch, normalize=True)
    src_uv_norm, dim_batch = normalize_by_roi(src_uv[:, :2, :], roi_batch, need_square=True)
    bsize = src_uv.shape[0]
    delta = model.pose_update(im_feat_batch, src_uv_norm.reshape(bsize, -1))
    # un-normalise to the expected scope (the network output is assumed around (-1, 1)): out-of-place form of :167-169
    scale = torch.cat([torch.full_like(delta[:, :3], 2 * torch.pi), dim_batch.unsqueeze(-1).expand(-1, 2), torch.ones_like(delta[:, 5:])], 1)
    shift = torch.cat([torch.zeros_like(delta[:, :5]), torch.ones_like(delta[:, 5:])], 1)
    delta = delta * scale + shift
    pred_R = axis_angle_to_matrix_batch(matrix_to_axis_angle_batch(src_pose_batch[:, :, :3]) + delta[:, :3])
    src_pose_uv = torch.matmul(K_batch, src_pose_batch[:, :, 3:])
    pred_u = src_pose_uv[:, 0] / src_pose_uv[:, 2] + delta[:, 3:4]
    pred_v = src_pose_uv[:, 1] / src_pose_uv[:, 2] + delta[:, 4:5]
    pred_Z = src_pose_batch[:, 2, 3:] * delta[:, 5:]
    pred_T = torch.cat([pred_u * pred_Z, pred_v * pred_Z, pred_Z], dim=1).unsqueeze(-1)
    pred_T = torch.matmul(torch.linalg.inv(K_batch) if K_inv is None else K_inv, pred_T)
    pred_pose = torch.cat([pred_R, pred_T], dim=2)
    pred_uv = view_points_batch(corners_of_box_batch(pred_pose, wlh_batch), K_batch, normalize=True)
    loss = torch.sqrt(torch.sum((pred_uv[:, :2, :] - tgt_uv_batch) ** 2, dim=-2))
    return loss, pred_pose


class PoseRegress3(nn.Module):
    """The three pose-regress iterations of the joint step (trainer_unified_nuscenes.py:93-118) as ONE module over the model's
    pose head, so that they can be captured in a CUDA graph (torch.cuda.make_graphed_callables): ~100 small kernels per iteration
    that are launch-bound when issued eagerly.  K_inv is passed in (torch.linalg.inv synchronises, which a capture forbids)."""

    def __init__(self, model):
        super().__init__()
        self.pose_blocks, self.regress_blocks = model.pose_blocks, model.regress_blocks
        for name in [f"pose_layer_{j}" for j in range(model.pose_blocks)] + [f"regress_layer_{j}" for j in range(model.regress_blocks)] + ["out_delta_layer"]:
            setattr(self, name, getattr(model, name))

    def pose_update(self, im_feat, box_uv_src):
        return pose_update(self, im_feat, box_uv_src)

    def forward(self, posecode, src_pose, tgt_uv, wlh, roi, K, K_inv):
        pose, losses = src_pose, []
        for _ in range(3):
            loss_i, pose = pose_regress(self, posecode, pose, tgt_uv, wlh, roi, K, K_inv)
            losses.append(loss_i.mean())
        return losses[0], losses[1], losses[2], pose


def joint_training_losses(model, hpams, img_in_batch, shapecode_batch, texturecode_batch, xyz_batch, viewdir_batch, z_vals_batch,
                          rgb_tgt_batch, occ_pixels_batch, src_pose_batch, tgt_uv_batch, roi_batch, K_batch, wlh_batch_aug,
                          tgt_uv_batch_aug, enc_active=True, im_enc_rate=1.0, encode=None, regress3=None):
    """ParallelModel.forward (trainer_unified_nuscenes.py:27-148) without the ``pred_wlh`` branch: common image encoding, direct
    corner regression, code consistency, three pose-regress iterations, then the NeRF sub-network (the package's decoder +
    ``volume_rendering_batch`` kernels) and the rgb / occupancy losses.  ``enc_active`` replaces the reference's
    ``random.uniform(0, 1) < im_enc_rate`` draw (the caller draws it).  ``encode``: the image-encoding callable
    (default ``model.encode_img``; the bench passes the bf16 channels-last one).  ``regress3``: optional callable
    (posecode, src_pose, tgt_uv, wlh, roi, K) -> (loss1, loss2, loss3, pose3) standing for the three pose-regress iterations
    (the bench passes a CUDA-graphed PoseRegress3).
    -> (losses_all, loss_total, shapecode_batch, texturecode_batch, pred_pose_batch3, pred_uv_batch_direct)"""
    from . import utils as U
    losses_all = {}
    shapecode, texturecode, posecode, pred_uv_direct, _ = (encode or model.encode_img)(img_in_batch)
    pred_uv_direct = pred_uv_direct.float().view(-1, 2, 8)
    dim_batch = torch.maximum(roi_batch[:, 2] - roi_batch[:, 0], roi_batch[:, 3] - roi_batch[:, 1])
    centre = torch.stack([(roi_batch[:, 0] + roi_batch[:, 2]) / 2, (roi_batch[:, 1] + roi_batch[:, 3]) / 2], 1).unsqueeze(-1)
    pred_uv_direct = pred_uv_direct * (dim_batch.view(-1, 1, 1) / 2) + centre
    loss_total = 0.
    losses_all["loss_pose_direct"] = torch.sqrt(torch.sum((pred_uv_direct[:, :2, :] - tgt_uv_batch) ** 2, dim=-2)).mean()
    if enc_active:
        loss_total = loss_total + hpams["loss_pose_coef"] * losses_all["loss_pose_direct"]
    shapecode, texturecode, posecode = shapecode.float(), texturecode.float(), posecode.float()
    losses_all["loss_code"] = torch.mean((shapecode - shapecode_batch) ** 2 + (texturecode - texturecode_batch) ** 2)
    if enc_active:
        if im_enc_rate < 1.0:
            loss_total = loss_total + hpams["loss_code_coef"] * losses_all["loss_code"]
        shapecode_batch = (shapecode_batch + shapecode) / 2
        texturecode_batch = (texturecode_batch + texturecode) / 2
    if regress3 is not None:
        l1, l2, l3, pose = regress3(posecode, src_pose_batch, tgt_uv_batch_aug, wlh_batch_aug, roi_batch, K_batch)
        iters = [l1, l2, l3]
    else:
        pose = src_pose_batch
        iters = []
        for _ in range(3):
            loss_i, pose = pose_regress(model, posecode, pose, tgt_uv_batch_aug, wlh_batch_aug, roi_batch, K_batch)
            iters.append(loss_i.mean())
    for i, l in enumerate(iters):
        losses_all["loss_pose_iter%d" % (i + 1)] = l
    if enc_active:
        loss_total = loss_total + hpams["loss_pose_coef"] * (iters[0] + iters[1] + iters[2]) / 3
    sigmas, rgbs = model(xyz_batch.flatten(0, 1), viewdir_batch.flatten(0, 1), shapecode_batch, texturecode_batch)
    b = img_in_batch.shape[0]
    n, s, _ = sigmas.shape
    rgb_rays, depth_rays, acc = U.volume_rendering_batch(sigmas.view(b, n // b, s, -1), rgbs.view(b, n // b, s, -1), z_vals_batch)
    den = torch.sum(torch.abs(occ_pixels_batch), dim=[-2, -1]) + 1e-9
    loss_rgb = torch.sum((rgb_rays - rgb_tgt_batch) ** 2 * torch.abs(occ_pixels_batch), dim=[-2, -1]) / den
    losses_all["loss_rgb"] = loss_rgb.mean()
    loss_occ = torch.sum(torch.exp(-occ_pixels_batch * (0.5 - acc.unsqueeze(-1))) * torch.abs(occ_pixels_batch), dim=[-2, -1]) / den
    losses_all["loss_occ"] = loss_occ.mean()
    losses_all["loss_reg"] = (torch.norm(shapecode_batch, dim=-1) + torch.norm(texturecode_batch, dim=-1)).mean()
    loss_total = loss_total + losses_all["loss_rgb"] + hpams["loss_occ_coef"] * losses_all["loss_occ"]
    losses_all["loss_total"] = loss_total
    return losses_all, loss_total, shapecode_batch, texturecode_batch, pose, pred_uv_direct[:, :2, :]
