"""CPU fp32 restatement of the PillarNeXt-B hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every function cites the reference file:line it follows (paths relative to /root/reference).
All functions are functional: parameters come from a state dict ``sd`` whose keys/shapes are
the reference's own, so the same dict drives the reference modules,
this oracle and the CUDA product.  Plain torch CPU ops, differentiable, so
``torch.autograd`` on this file is also the gradient oracle.

Pinned (in the build container) against the reference's own reader/neck/head/loss files by
oracle/make_golden.py + tests/test_oracle_vs_reference.py.  Backbone: parity UNPINNED (spconv
absent); ``sparse_resnet_dense`` and ``sparse_resnet_gather`` must agree with each other.
"""
import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- helpers
# QUANT=True makes the restatement "bf16-faithful": weights and every stored activation are rounded to bf16
# (round-to-nearest-even, straight-through for autograd) at exactly the points where the CUDA product stores
# bf16, all arithmetic staying fp32.  The product must match THIS variant tightly (only accumulation order
# differs) and the plain fp32 variant within bf16 noise.
QUANT = False


def q(x):
    if not QUANT:
        return x
    return x + (x.to(torch.bfloat16).to(torch.float32) - x).detach()


def grid_size_xy(voxel_size, pc_range):
    """pillar_encoder.py:87-89 -- float64 numpy, round to int64. Returns (Gx, Gy, Gz)."""
    vs = np.array(voxel_size, dtype=np.float64)
    pr = np.array(pc_range, dtype=np.float64)
    g = (pr[3:] - pr[:3]) / vs
    return np.round(g).astype(np.int64)


def bn_train(x, weight, bias, eps, dims):
    """Training-mode batch norm over ``dims`` (nn.BatchNorm1d/2d forward, biased variance).
    Returns y, batch_mean, batch_var_unbiased, n."""
    n = 1
    for d in dims:
        n *= x.shape[d]
    mean = x.mean(dim=dims, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=dims, keepdim=True)
    shape = [1] * x.dim()
    cdim = [d for d in range(x.dim()) if d not in dims][0]
    shape[cdim] = -1
    y = (x - mean) / torch.sqrt(var + eps) * weight.view(shape) + bias.view(shape)
    unb = var.flatten() * (n / max(n - 1, 1))
    return y, mean.flatten().detach(), unb.detach(), n


def bn_eval(x, weight, bias, rm, rv, eps, cdim):
    shape = [1] * x.dim()
    shape[cdim] = -1
    return (x - rm.view(shape)) / torch.sqrt(rv.view(shape) + eps) * weight.view(shape) + bias.view(shape)


def _bn(x, sd, pfx, eps, dims, cdim, train, stats, momentum):
    if train:
        y, m, v, n = bn_train(x, sd[pfx + "weight"], sd[pfx + "bias"], eps, dims)
        if stats is not None:
            stats[pfx + "running_mean"] = (1 - momentum) * sd[pfx + "running_mean"] + momentum * m
            stats[pfx + "running_var"] = (1 - momentum) * sd[pfx + "running_var"] + momentum * v
        return y
    return bn_eval(x, sd[pfx + "weight"], sd[pfx + "bias"], sd[pfx + "running_mean"],
                   sd[pfx + "running_var"], eps, cdim)


# ----------------------------------------------------------------------------- V1-V3 voxelizer
def voxelize(points, voxel_size, pc_range):
    """PillarNet.forward, pillar_encoder.py:78-125.

    points [N, 1+5] fp32 = (batch_idx, x, y, z, intensity, time).
    Returns dict(features [Nv,10], coords [P,3] int32 (b, yi, xi), unq_inv [Nv] int64,
                 grid (Gy, Gx) np.int64[2], keep [N] bool, pillar_of_point [N] int64 (-1 dropped)).
    """
    points = points.detach().to(torch.float32)
    g = grid_size_xy(voxel_size, pc_range)                                   # :87-89
    vs = torch.from_numpy(np.array(voxel_size)).to(torch.float32)            # :91-92 (fp32 cast)
    pr = torch.from_numpy(np.array(pc_range)).to(torch.float32)              # :93
    pc = (points[:, 1:4] - pr[:3].view(-1, 3)) / vs.view(-1, 3)              # :95-96 true fp32 division
    keep = (pc[:, 0] >= 0) & (pc[:, 0] < float(g[0])) & (pc[:, 1] >= 0) & (pc[:, 1] < float(g[1]))  # :98-101
    pts = points[keep]                                                       # :103
    pcl = pc[keep].long()                                                    # :104,106 trunc
    bidx = pts[:, 0:1].long()                                                # :107
    pidx = torch.cat((bidx, pcl[:, :2]), dim=1)                              # :109  (b, xi, yi)
    unq, unq_inv = torch.unique(pidx, return_inverse=True, dim=0)            # :110  sorted lexicographically
    unq = unq.int()                                                          # :111
    n_p = unq.shape[0]
    mean = torch.zeros(n_p, 3).index_add_(0, unq_inv, pts[:, 1:4])           # :113-114 scatter_mean
    cnt = torch.zeros(n_p).index_add_(0, unq_inv, torch.ones(pts.shape[0])).clamp(min=1)
    mean = mean / cnt.view(-1, 1)
    f_cluster = pts[:, 1:4] - mean[unq_inv]                                  # :116
    f_center = pts[:, 1:3] - (pcl[:, :2].to(torch.float32) * vs[:2].unsqueeze(0)
                              + vs[:2].unsqueeze(0) / 2 + pr[:2].unsqueeze(0))   # :119-120
    features = torch.cat([pts[:, 1:], f_cluster, f_center], dim=-1)          # :123
    pop = torch.full((points.shape[0],), -1, dtype=torch.long)
    pop[keep] = unq_inv
    return dict(features=features, coords=unq[:, [0, 2, 1]].contiguous(), unq_inv=unq_inv,
                grid=g[[1, 0]], keep=keep, pillar_of_point=pop)              # :125


def scatter_max(x, idx, n):
    idx2 = idx.view(-1, 1).expand_as(x)
    out = torch.full((n, x.shape[1]), float("-inf"), dtype=x.dtype)
    return out.scatter_reduce(0, idx2, x, reduce="amax", include_self=True)


# ----------------------------------------------------------------------------- P1-P3 PFN
def pfn_forward(features, unq_inv, n_pillars, sd, prefix="reader.", train=True, stats=None):
    """PFNLayer.forward x2 + final scatter_max: pillar_encoder.py:35-50, 174-182.
    BN1d eps=1e-3 momentum=0.01 (:33)."""
    x = features
    for li, last in ((0, False), (1, True)):
        p = "%spfn_layers.%d." % (prefix, li)
        x = x @ sd[p + "linear.weight"].t()                                  # :37
        x = _bn(x, sd, p + "norm.", 1e-3, [0], 1, train, stats, 0.01)        # :38
        x = F.relu(x)                                                        # :39
        fmax = scatter_max(x, unq_inv, n_pillars)                            # :43
        xmax = fmax[unq_inv]                                                 # :44
        x = xmax if last else torch.cat([x, xmax], dim=1)                    # :46-50
    return scatter_max(x, unq_inv, n_pillars)                                # :180


def reader_forward(points, sd, voxel_size, pc_range, prefix="reader.", train=True, stats=None):
    """PillarFeatureNet.forward, pillar_encoder.py:174-182 -> (feat_max[P,64], coords[P,3] i32, grid)."""
    v = voxelize(points, voxel_size, pc_range)
    feat = pfn_forward(v["features"], v["unq_inv"], v["coords"].shape[0], sd, prefix, train, stats)
    return feat, v["coords"], v["grid"]


# ----------------------------------------------------------------------------- B1-B4 backbone
def _spw(w):
    """spconv-2.x weight layout [Cout, kH, kW, Cin] ([3P-unverified]) -> KCRS."""
    return w.permute(0, 3, 1, 2).contiguous()


def _masked_bn(x, mask, sd, pfx, train, stats):
    """nn.BatchNorm1d(eps=1e-3, momentum=0.01) applied to ``.features`` = ACTIVE sites only
    (sparse_conv.py:31,36; sparse_resnet.py:46).  x [B,C,H,W], mask [B,1,H,W] in {0,1}."""
    m = mask.bool().expand_as(x)
    c = x.shape[1]
    feats = x.permute(1, 0, 2, 3)[mask.bool().expand(-1, c, -1, -1).permute(1, 0, 2, 3)].view(c, -1).t()
    y = _bn(feats, sd, pfx, 1e-3, [0], 1, train, stats, 0.01)
    out = torch.zeros_like(x)
    out.permute(1, 0, 2, 3)[m.permute(1, 0, 2, 3)] = y.t().reshape(-1)
    return out


def sparse_resnet_dense(feat, coords, grid_hw, batch_size, sd, strides=(1, 2, 2, 2),
                        prefix="backbone.", train=True, stats=None):
    """SparseResNet.forward, sparse_resnet.py:61-68, restated on a dense canvas + active mask.

    SparseConv2d(bias=False) == dense conv on the zero-filled canvas, active set dilates:
    mask_out = max_pool2d(mask_in, 3, stride, 1)     (sparse_conv.py:28-29, use_subm=False :53-54)
    SubMConv2d == dense conv * mask                   (sparse_conv.py:25-26, 50-51)
    BN over active sites only; ReLU; residual         (sparse_conv.py:33-39, 55-63)
    Returns dense [B,256,H/8,W/8] (x.dense(), :68) and the final mask.
    """
    H, W = int(grid_hw[0]), int(grid_hw[1])
    c_in = feat.shape[1]
    b, yy, xx = coords[:, 0].long(), coords[:, 1].long(), coords[:, 2].long()
    canvas = torch.zeros(batch_size, H, W, c_in, dtype=feat.dtype)
    canvas = canvas.index_put((b, yy, xx), feat)
    x = canvas.permute(0, 3, 1, 2)
    mask = torch.zeros(batch_size, 1, H, W)
    mask[b, 0, yy, xx] = 1.0
    for s in range(4):
        p = "%sblocks.%d." % (prefix, s)
        x = F.conv2d(x, _spw(sd[p + "0.conv.weight"]), stride=strides[s], padding=1)
        mask = F.max_pool2d(mask, 3, strides[s], 1)
        x = F.relu(_masked_bn(x, mask, sd, p + "0.norm.", train, stats))
        for j in (1, 2):
            q = "%s%d." % (p, j)
            idt = x
            o = F.conv2d(x, _spw(sd[q + "block1.conv.weight"]), padding=1) * mask
            o = F.relu(_masked_bn(o, mask, sd, q + "block1.norm.", train, stats))
            o = F.conv2d(o, _spw(sd[q + "conv2.weight"]), padding=1) * mask
            o = _masked_bn(o, mask, sd, q + "norm2.", train, stats)
            x = F.relu(o + idt)
    x = F.conv2d(x, _spw(sd[prefix + "mapping.0.weight"])) * mask            # :43-48 (1x1: set unchanged)
    x = F.relu(_masked_bn(x, mask, sd, prefix + "mapping.1.", train, stats))
    return x, mask


def _site_index(coords, shape_bhw):
    B, H, W = shape_bhw
    grid = torch.full((B, H, W), -1, dtype=torch.long)
    grid[coords[:, 0], coords[:, 1], coords[:, 2]] = torch.arange(coords.shape[0])
    return grid


def _gather_conv(feat, in_coords, in_shape, out_coords, w, stride, pad=1):
    """out[i] = sum_{ky,kx} W[:,ky,kx,:] @ in[(y*stride+ky-pad, x*stride+kx-pad)]  (gather-GEMM)."""
    B, H, W = in_shape
    grid = _site_index(in_coords, in_shape)
    kh, kw = w.shape[1], w.shape[2]
    out = torch.zeros(out_coords.shape[0], w.shape[0], dtype=feat.dtype)
    ob, oy, ox = out_coords[:, 0], out_coords[:, 1], out_coords[:, 2]
    for ky in range(kh):
        for kx in range(kw):
            iy, ix = oy * stride + ky - pad, ox * stride + kx - pad
            ok = (iy >= 0) & (iy < H) & (ix >= 0) & (ix < W)
            src = torch.full_like(iy, -1)
            src[ok] = grid[ob[ok], iy[ok], ix[ok]]
            sel = src >= 0
            if sel.any():
                out[sel] += feat[src[sel]] @ q(w[:, ky, kx, :]).t()
    return q(out)


def _dilate_sites(coords, in_shape, stride):
    B, H, W = in_shape
    m = torch.zeros(B, 1, H, W)
    m[coords[:, 0], 0, coords[:, 1], coords[:, 2]] = 1.0
    m = F.max_pool2d(m, 3, stride, 1)
    oc = torch.nonzero(m[:, 0] > 0)                       # sorted (b, y, x)
    return oc, (B, m.shape[2], m.shape[3])


def sparse_resnet_gather(feat, coords, grid_hw, batch_size, sd, strides=(1, 2, 2, 2),
                         prefix="backbone.", train=True, stats=None):
    """Second, independent restatement of sparse_resnet.py:61-68 in gather-GEMM-scatter form
    (feature matrix over active sites + per-offset index_select/mm), as spconv executes it.
    Returns (features [N4,256], coords [N4,3] (b,y,x), (B,H4,W4))."""
    c = coords.long()
    shape = (batch_size, int(grid_hw[0]), int(grid_hw[1]))
    x = q(feat)
    for s in range(4):
        p = "%sblocks.%d." % (prefix, s)
        oc, oshape = _dilate_sites(c, shape, strides[s])
        x = _gather_conv(x, c, shape, oc, sd[p + "0.conv.weight"], strides[s])
        c, shape = oc, oshape
        x = q(F.relu(_bn(x, sd, p + "0.norm.", 1e-3, [0], 1, train, stats, 0.01)))
        for j in (1, 2):
            bq = "%s%d." % (p, j)
            idt = x
            o = _gather_conv(x, c, shape, c, sd[bq + "block1.conv.weight"], 1)
            o = q(F.relu(_bn(o, sd, bq + "block1.norm.", 1e-3, [0], 1, train, stats, 0.01)))
            o = _gather_conv(o, c, shape, c, sd[bq + "conv2.weight"], 1)
            o = _bn(o, sd, bq + "norm2.", 1e-3, [0], 1, train, stats, 0.01)
            x = q(F.relu(o + idt))
    x = q(x @ q(sd[prefix + "mapping.0.weight"][:, 0, 0, :]).t())
    x = q(F.relu(_bn(x, sd, prefix + "mapping.1.", 1e-3, [0], 1, train, stats, 0.01)))
    return x, c, shape


def densify(feat, coords, shape):
    """SparseConvTensor.dense(): zero-filled NCHW (sparse_resnet.py:68)."""
    B, H, W = shape
    out = torch.zeros(B, H, W, feat.shape[1], dtype=feat.dtype)
    out = out.index_put((coords[:, 0], coords[:, 1], coords[:, 2]), feat)
    return out.permute(0, 3, 1, 2).contiguous()


# ----------------------------------------------------------------------------- N1 neck
def _convblock2d(x, sd, pfx, train, stats, padding):
    """ConvBlock: conv(bias=False) + BatchNorm2d(eps 1e-5, mom 0.1) + ReLU, conv.py:14-34."""
    x = q(F.conv2d(x, q(sd[pfx + "conv.conv.weight"]), padding=padding))
    x = _bn(x, sd, pfx + "norm.", 1e-5, [0, 2, 3], 1, train, stats, 0.1)
    return q(F.relu(x))


def aspp_forward(x, sd, prefix="neck.", train=True, stats=None):
    """ASPPNeck._forward, aspp.py:19-32 (checkpointing :34-38 does not change values)."""
    idt = x
    o = _convblock2d(x, sd, prefix + "pre_conv.block1.", train, stats, 1)    # conv.py:44-51
    o = _convblock2d(o, sd, prefix + "pre_conv.block2.", train, stats, 1)
    x = q(F.relu(o + idt))
    w = q(sd[prefix + "weight"])
    br = [x, q(F.conv2d(x, q(sd[prefix + "conv1x1.weight"])))]
    for d in (1, 6, 12, 18):
        br.append(q(F.conv2d(x, w, padding=d, dilation=d)))                  # :22-29 shared weight
    x = torch.cat(br, dim=1)
    return _convblock2d(x, sd, prefix + "post_conv.", train, stats, 0)       # :30-31


# ----------------------------------------------------------------------------- H1 head
def centerhead_forward(x, sd, tasks, common_heads, prefix="head.", train=True, stats=None):
    """CenterHead.forward / SepHead.forward, centerhead.py:128-136, 53-59.
    tasks: list of class-name lists; common_heads: ordered dict name -> (channels, num_conv)."""
    p = prefix + "shared_conv."
    x = q(F.conv2d(x, q(sd[p + "0.weight"]), sd[p + "0.bias"], padding=1))   # :108-114
    x = q(F.relu(_bn(x, sd, p + "1.", 1e-5, [0, 2, 3], 1, train, stats, 0.1)))
    rets = []
    for t, names in enumerate(tasks):
        tq = "%stasks.%d." % (prefix, t)
        y = q(F.conv_transpose2d(x, q(sd[tq + "deblock.conv.conv.weight"]), stride=2))   # :26-27
        y = q(F.relu(_bn(y, sd, tq + "deblock.norm.", 1e-5, [0, 2, 3], 1, train, stats, 0.1)))
        heads = list(common_heads.keys()) + ["hm"]                           # :121-122 order
        ret = {}
        for h in heads:
            r = tq + h + "."
            z = q(F.conv2d(y, q(sd[r + "0.weight"]), sd[r + "0.bias"], padding=1))   # :36-38
            z = q(F.relu(_bn(z, sd, r + "1.", 1e-5, [0, 2, 3], 1, train, stats, 0.1)))
            ret[h] = F.conv2d(z, q(sd[r + "3.weight"]), sd[r + "3.bias"], padding=1)  # :44-46 (fp32 output)
        rets.append(ret)
    return rets


# ----------------------------------------------------------------------------- L1 loss
def _gather_feat(feat, ind):
    """_transpose_and_gather_feat, centerloss.py:113-128: feat [B,C,H,W], ind [B,M] -> [B,M,C]."""
    B, C = feat.shape[0], feat.shape[1]
    f = feat.permute(0, 2, 3, 1).reshape(B, -1, C)
    return f.gather(1, ind.unsqueeze(2).expand(-1, -1, C))


def fast_focal_loss(out, target, ind, mask, cat):
    """FastFocalLoss.forward, centerloss.py:17-37."""
    mask = mask.float()
    gt = torch.pow(1 - target, 4)
    neg_loss = (torch.pow(out, 2) * gt * torch.log(1 - out)).sum()
    pos_pred = _gather_feat(out, ind).gather(2, cat.unsqueeze(2))
    num_pos = mask.sum()
    pos_loss = (torch.log(pos_pred) * torch.pow(1 - pos_pred, 2) * mask.unsqueeze(2)).sum()
    if num_pos == 0:
        return -neg_loss
    return -(pos_loss + neg_loss) / num_pos


def reg_loss(output, mask, ind, target):
    """RegLoss.forward, centerloss.py:53-61 (NaN targets replaced by the prediction :56-57)."""
    pred = _gather_feat(output, ind)
    mask = mask.float().unsqueeze(2)
    target = torch.where(torch.isnan(target), pred.detach(), target)
    loss = F.l1_loss(pred * mask, target * mask, reduction="none")
    loss = loss / (mask.sum() + 1e-4)
    return loss.transpose(2, 0).sum(dim=2).sum(dim=1)


def diou_aligned(pred_boxes, gt_boxes):
    """bbox3d_overlaps_diou, centerloss.py:139-176 (axis-aligned; yaw ignored)."""
    def corners(center, dim):
        cn = torch.tensor([[-0.5, -0.5], [-0.5, 0.5], [0.5, 0.5], [0.5, -0.5]], dtype=torch.float32)
        return dim.view(-1, 1, 2) * cn.view(1, 4, 2) + center.view(-1, 1, 2)
    q = corners(pred_boxes[:, :2], pred_boxes[:, 3:5])
    g = corners(gt_boxes[:, :2], gt_boxes[:, 3:5])
    inter_max, inter_min = torch.minimum(q[:, 2], g[:, 2]), torch.maximum(q[:, 0], g[:, 0])
    out_max, out_min = torch.maximum(q[:, 2], g[:, 2]), torch.minimum(q[:, 0], g[:, 0])
    vp = pred_boxes[:, 3] * pred_boxes[:, 4] * pred_boxes[:, 5]
    vg = gt_boxes[:, 3] * gt_boxes[:, 4] * gt_boxes[:, 5]
    ih = torch.minimum(pred_boxes[:, 2] + 0.5 * pred_boxes[:, 5], gt_boxes[:, 2] + 0.5 * gt_boxes[:, 5]) - \
        torch.maximum(pred_boxes[:, 2] - 0.5 * pred_boxes[:, 5], gt_boxes[:, 2] - 0.5 * gt_boxes[:, 5])
    ih = torch.clamp(ih, min=0)
    inter = torch.clamp(inter_max - inter_min, min=0)
    vi = inter[:, 0] * inter[:, 1] * ih
    vu = vg + vp - vi
    idiag = torch.pow(gt_boxes[:, 0:3] - pred_boxes[:, 0:3], 2).sum(-1)
    oh = torch.maximum(gt_boxes[:, 2] + 0.5 * gt_boxes[:, 5], pred_boxes[:, 2] + 0.5 * pred_boxes[:, 5]) - \
        torch.minimum(gt_boxes[:, 2] - 0.5 * gt_boxes[:, 5], pred_boxes[:, 2] - 0.5 * pred_boxes[:, 5])
    oh = torch.clamp(oh, min=0)
    outer = torch.clamp(out_max - out_min, min=0)
    odiag = outer[:, 0] ** 2 + outer[:, 1] ** 2 + oh ** 2
    return torch.clamp(vi / vu - idiag / odiag, min=-1.0, max=1.0)


def iou_reg_loss(box_pred, mask, ind, box_gt):
    """IouRegLoss.forward, centerloss.py:103-110."""
    if mask.sum() == 0:
        return box_pred.sum() * 0
    m = mask.bool()
    pb = _gather_feat(box_pred, ind)
    iou = diou_aligned(pb[m], box_gt[m])
    return (1.0 - iou).sum() / (m.sum() + 1e-4)


def iou_loss(iou_pred, mask, ind, box_pred, box_gt):
    """IouLoss, centerloss.py:64-87: L1 between the `iou` head at the object centres and 2*IoU3d(decoded box, gt) - 1
    (the target carries no gradient: the boxes are detached by the caller, centerhead.py:211-212)."""
    from oracle import predict_oracle as PO
    if mask.sum() == 0:
        return iou_pred.sum() * 0
    m = mask.bool()
    pred = _gather_feat(iou_pred, ind)[m]          # [B, 1, H, W] -> [K, 1]
    pb = _gather_feat(box_pred, ind)[m]            # [B, 7, H, W] -> [K, 7]
    gt = box_gt[m]
    tgt = torch.tensor([[float(PO.aligned_iou3d(a.numpy(), b.numpy()))] for a, b in zip(pb.detach(), gt)], dtype=pred.dtype)
    tgt = 2 * tgt - 1
    return F.l1_loss(pred, tgt.reshape(pred.shape), reduction="sum") / (mask.sum() + 1e-4)


def center_loss(example, preds, weight, code_weights, with_reg_iou, voxel_size, pc_range, out_size_factor, with_iou=False):
    """CenterHead.loss, centerhead.py:142-229 (`with_iou`: the Waymo `iou` head branch, :210-215).
    NOTE: like the reference (:146, :138-140) this REPLACES preds[t]['hm'] by its clamped sigmoid
    (out of place here so autograd stays valid)."""
    total = None
    rets = []
    for t, pd in enumerate(preds):
        hm = torch.clamp(torch.sigmoid(pd["hm"]), min=1e-4, max=1 - 1e-4)    # :138-140
        hm_loss = fast_focal_loss(hm, example["hm"][t], example["ind"][t], example["mask"][t], example["cat"][t])
        anno = torch.cat((pd["reg"], pd["height"], pd["dim"], pd["vel"], pd["rot"]), dim=1)   # :154-155
        box_loss = reg_loss(anno, example["mask"][t], example["ind"][t], example["anno_box"][t])
        loc_loss = (box_loss * box_loss.new_tensor(code_weights)).sum()      # :161
        loss = hm_loss + weight * loc_loss                                   # :163
        ret = dict(hm_loss=hm_loss.detach(), loc_loss=loc_loss.detach(), loc_loss_elem=box_loss.detach(),
                   num_positive=example["mask"][t].float().sum())
        if with_reg_iou or with_iou:
            bdim = torch.exp(torch.clamp(pd["dim"], min=-5, max=5)).permute(0, 2, 3, 1)      # :172-174
            brot = pd["rot"].permute(0, 2, 3, 1)
            brot = torch.atan2(brot[..., 0:1], brot[..., 1:2])               # :177-179
            breg = pd["reg"].permute(0, 2, 3, 1)
            bhei = pd["height"].permute(0, 2, 3, 1)
            B, H, W, _ = bdim.shape
            ys, xs = torch.meshgrid(torch.arange(0, H), torch.arange(0, W), indexing="ij")   # :193
            xs = xs.view(1, H, W, 1).to(bdim) + breg[..., 0:1]
            ys = ys.view(1, H, W, 1).to(bdim) + breg[..., 1:2]
            xs = xs * out_size_factor[t] * voxel_size[0] + pc_range[0]       # :201-204
            ys = ys * out_size_factor[t] * voxel_size[1] + pc_range[1]
            boxes = torch.cat([xs, ys, bhei, bdim, brot], dim=3).permute(0, 3, 1, 2)          # :206-209
            if with_iou:
                il = iou_loss(pd["iou"], example["mask"][t], example["ind"][t], boxes.detach(), example["gt_boxes"][t])
                loss = loss + il                                             # :214
                ret["iou_loss"] = il.detach()
            if with_reg_iou:
                irl = iou_reg_loss(boxes, example["mask"][t], example["ind"][t], example["gt_boxes"][t])
                loss = loss + weight * irl                                   # :221
                ret["iou_reg_loss"] = irl.detach()
        ret["loss"] = loss
        rets.append(ret)
        total = loss if total is None else total + loss
    return total, rets


# ----------------------------------------------------------------------------- whole detector
def detector_forward(points, sd, cfg, batch_size, train=True, stats=None, backbone="gather"):
    """SingleStageDetector._forward, single_stage.py:22-33: reader -> backbone -> neck -> head."""
    feat, coords, grid = reader_forward(points, sd, cfg["voxel_size"], cfg["pc_range"], "reader.", train, stats)
    if backbone == "gather":
        f4, c4, shp = sparse_resnet_gather(feat, coords, grid, batch_size, sd, cfg["strides"], "backbone.", train, stats)
        x = densify(f4, c4, shp)
    else:
        x, _ = sparse_resnet_dense(feat, coords, grid, batch_size, sd, cfg["strides"], "backbone.", train, stats)
    x = aspp_forward(x, sd, "neck.", train, stats)
    return centerhead_forward(x, sd, cfg["tasks"], cfg["common_heads"], "head.", train, stats)
