"""CenterPoint losses on the head outputs (row L1 of SURVEY.md section 8a).

Same formulas as reference det3d/models/heads/centerhead.py:142-229 and det3d/models/loss/centerloss.py:8-176
(FastFocalLoss, RegLoss, IouRegLoss/DIoU) but WITHOUT the reference's host syncs (`if num_pos == 0`,
`if mask.sum() == 0`, `.cpu()` log values: centerloss.py:35,77,104; centerhead.py:166-169): the empty cases are
folded into the arithmetic with identical results.  Round-1 status: these are small elementwise/gather torch
ops on the fp32 head maps (HBM-bound, ~1.1 M elements per frame and task); the fused CUDA loss is row F2.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F


def _gather_feat(feat, ind):
    """_transpose_and_gather_feat (centerloss.py:113-128): feat [B,C,H,W], ind [B,M] -> [B,M,C]."""
    B, C = feat.shape[0], feat.shape[1]
    f = feat.permute(0, 2, 3, 1).reshape(B, -1, C)
    return f.gather(1, ind.unsqueeze(2).expand(-1, -1, C))


def fast_focal_loss(out, target, ind, mask, cat):
    """centerloss.py:17-37."""
    mask = mask.float()
    gt = torch.pow(1 - target, 4)
    neg_loss = (torch.pow(out, 2) * gt * torch.log(1 - out)).sum()
    pos_pred = _gather_feat(out, ind).gather(2, cat.unsqueeze(2))
    num_pos = mask.sum()
    pos_loss = (torch.log(pos_pred) * torch.pow(1 - pos_pred, 2) * mask.unsqueeze(2)).sum()
    # num_pos == 0  ->  pos_loss == 0 and the reference returns -neg_loss: same value as dividing by 1
    return -(pos_loss + neg_loss) / num_pos.clamp(min=1.0)


def reg_loss(output, mask, ind, target):
    """centerloss.py:53-61."""
    pred = _gather_feat(output, ind)
    mask = mask.float().unsqueeze(2)
    target = torch.where(torch.isnan(target), pred.detach(), target)
    loss = F.l1_loss(pred * mask, target * mask, reduction="none")
    loss = loss / (mask.sum() + 1e-4)
    return loss.transpose(2, 0).sum(dim=2).sum(dim=1)


def diou_aligned(p, g):
    """bbox3d_overlaps_diou (centerloss.py:139-176): axis-aligned, yaw ignored. p, g [N,7]."""
    qmin, qmax = p[:, :2] - 0.5 * p[:, 3:5], p[:, :2] + 0.5 * p[:, 3:5]
    gmin, gmax = g[:, :2] - 0.5 * g[:, 3:5], g[:, :2] + 0.5 * g[:, 3:5]
    inter = torch.clamp(torch.minimum(qmax, gmax) - torch.maximum(qmin, gmin), min=0)
    outer = torch.clamp(torch.maximum(qmax, gmax) - torch.minimum(qmin, gmin), min=0)
    vp = p[:, 3] * p[:, 4] * p[:, 5]
    vg = g[:, 3] * g[:, 4] * g[:, 5]
    ih = torch.clamp(torch.minimum(p[:, 2] + 0.5 * p[:, 5], g[:, 2] + 0.5 * g[:, 5]) -
                     torch.maximum(p[:, 2] - 0.5 * p[:, 5], g[:, 2] - 0.5 * g[:, 5]), min=0)
    vi = inter[:, 0] * inter[:, 1] * ih
    vu = vg + vp - vi
    idiag = torch.pow(g[:, 0:3] - p[:, 0:3], 2).sum(-1)
    oh = torch.clamp(torch.maximum(g[:, 2] + 0.5 * g[:, 5], p[:, 2] + 0.5 * p[:, 5]) -
                     torch.minimum(g[:, 2] - 0.5 * g[:, 5], p[:, 2] - 0.5 * p[:, 5]), min=0)
    odiag = outer[:, 0] ** 2 + outer[:, 1] ** 2 + oh ** 2
    return torch.clamp(vi / vu - idiag / odiag, min=-1.0, max=1.0)


def iou_reg_loss(box_pred, mask, ind, box_gt):
    """centerloss.py:103-110; the empty-mask early-out is folded in (0 / 1e-4 == 0)."""
    m = mask.bool()
    pb = _gather_feat(box_pred, ind)
    one = torch.ones((), dtype=pb.dtype, device=pb.device)
    mm = m.unsqueeze(2)
    pb = torch.where(mm, pb, one)        # unmasked rows -> harmless unit boxes (keeps 0/0 out of the graph)
    gb = torch.where(mm, box_gt, one)
    iou = diou_aligned(pb.reshape(-1, pb.shape[2]), gb.reshape(-1, gb.shape[2])).view(m.shape)
    return ((1.0 - iou) * m.float()).sum() / (m.float().sum() + 1e-4)


def center_loss(example, preds_dicts, class_names, weight, code_weights, with_reg_iou, voxel_size, pc_range,
                out_size_factor):
    """CenterHead.loss, centerhead.py:142-229. Returns (total_loss, list of per-task OrderedDict logs).
    Like the reference, preds_dict['hm'] is replaced by its clamped sigmoid (:146)."""
    rets = []
    total = None
    for t, pd in enumerate(preds_dicts):
        pd["hm"] = torch.clamp(torch.sigmoid(pd["hm"]), min=1e-4, max=1 - 1e-4)          # :138-140
        ind, mask = example["ind"][t], example["mask"][t]
        hm_loss = fast_focal_loss(pd["hm"], example["hm"][t], ind, mask, example["cat"][t])
        pd["anno_box"] = torch.cat((pd["reg"], pd["height"], pd["dim"], pd["vel"], pd["rot"]), dim=1)   # :154-155
        box_loss = reg_loss(pd["anno_box"], mask, ind, example["anno_box"][t])
        loc_loss = (box_loss * box_loss.new_tensor(code_weights)).sum()                  # :161
        loss = hm_loss + weight * loc_loss                                               # :163
        ret = OrderedDict()
        ret.update({"task": class_names[t], "loss": loss, "hm_loss": hm_loss.detach(), "loc_loss": loc_loss.detach(),
                    "loc_loss_elem": box_loss.detach(), "num_positive": mask.float().sum()})
        if with_reg_iou:
            bdim = torch.exp(torch.clamp(pd["dim"], min=-5, max=5)).permute(0, 2, 3, 1)  # :172-174
            brot = pd["rot"].permute(0, 2, 3, 1)
            brot = torch.atan2(brot[..., 0:1], brot[..., 1:2])                           # :177-179
            breg = pd["reg"].permute(0, 2, 3, 1)
            bhei = pd["height"].permute(0, 2, 3, 1)
            B, H, W, _ = bdim.shape
            ys = torch.arange(0, H, device=bdim.device, dtype=bdim.dtype).view(1, H, 1, 1)
            xs = torch.arange(0, W, device=bdim.device, dtype=bdim.dtype).view(1, 1, W, 1)
            xs = (xs + breg[..., 0:1]) * out_size_factor[t] * voxel_size[0] + pc_range[0]   # :198-204
            ys = (ys + breg[..., 1:2]) * out_size_factor[t] * voxel_size[1] + pc_range[1]
            boxes = torch.cat([xs, ys, bhei, bdim, brot], dim=3).permute(0, 3, 1, 2)        # :206-209
            irl = iou_reg_loss(boxes, mask, ind, example["gt_boxes"][t])
            loss = loss + weight * irl                                                   # :221
            ret.update({"iou_reg_loss": irl.detach()})
            ret["loss"] = loss
        rets.append(ret)
        total = loss if total is None else total + loss
    return total, rets


def iou_head_loss(out, off, B, H, W, ind, mask, gt_boxes, out_size_factor, voxel_size, pc_range):
    """IouLoss of the Waymo `iou` head (centerloss.py:64-87, called from centerhead.py:210-215) on the fused head's
    channels-last output `out` [B*H*W, npad] (column offsets `off`).  Sync-free: the object rows are gathered with
    torch (differentiable w.r.t. `out`), the detached target 2*IoU3d(decoded box, gt)-1 comes from the libpnx
    aligned rotated-IoU kernel, absent objects are masked instead of compacted (0 / 1e-4 = 0 when a frame set has none)."""
    from . import ops
    npad = out.shape[1]
    rows = out.view(B, H * W, npad)
    g = torch.gather(rows, 1, ind.unsqueeze(-1).expand(-1, -1, npad))                  # [B, M, npad]
    pred = g[..., off["iou"]]
    with torch.no_grad():
        x = (ind % W).to(g.dtype) + g[..., off["reg"]]
        y = torch.div(ind, W, rounding_mode="floor").to(g.dtype) + g[..., off["reg"] + 1]
        x = x * out_size_factor * voxel_size[0] + pc_range[0]                          # centerhead.py:201-204
        y = y * out_size_factor * voxel_size[1] + pc_range[1]
        dim = torch.exp(torch.clamp(g[..., off["dim"]:off["dim"] + 3], min=-5, max=5))
        rot = torch.atan2(g[..., off["rot"]], g[..., off["rot"] + 1])
        boxes = torch.cat([x.unsqueeze(-1), y.unsqueeze(-1), g[..., off["height"]:off["height"] + 1], dim, rot.unsqueeze(-1)], -1)
        target = 2 * ops.aligned_iou3d(boxes.reshape(-1, 7), gt_boxes.reshape(-1, 7).to(boxes.dtype)).view(pred.shape) - 1
    m = mask.bool()
    l1 = torch.where(m, (pred - target).abs(), torch.zeros_like(pred))
    return l1.sum() / (m.sum().to(pred.dtype) + 1e-4)
