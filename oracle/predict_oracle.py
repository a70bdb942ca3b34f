"""CPU restatement of the detection post-processing (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Row F1 of SURVEY.md section 8f: CenterHead.predict / post_processing (reference det3d/models/heads/centerhead.py:231-384),
rotate_nms_pcdet (det3d/core/bbox/box_torch_ops.py:5-31) and the rotated BEV IoU + greedy NMS of the reference's native
extension (det3d/core/iou3d_nms/src/iou3d_cpu.cpp:62-238, iou3d_nms_kernel.cu:280-324, iou3d_nms.cpp:113-159).

Pinning (tests/test_predict_cpu.py, build container only): ``iou_bev`` against the reference's own iou3d_cpu.cpp compiled
into oracle/_ref (oracle/build_ref_iou.py); ``predict`` against the reference's own CenterHead.predict executed on CPU
with its CUDA-only ``nms_gpu`` call replaced by ``nms_rotated`` from this file.  fp32 arithmetic throughout (numpy
float32 scalars) so threshold decisions match the native code.
"""
import math

import numpy as np
import torch

F = np.float32
EPS = F(1e-8)        # iou3d_cpu.cpp:40
MARGIN = F(1e-2)     # iou3d_cpu.cpp:78


def _cross2(ax, ay, bx, by):
    return F(F(ax * by) - F(ay * bx))


def _cross3(p1, p2, p0):
    """iou3d_cpu.cpp:63-65"""
    return F(F(F(p1[0] - p0[0]) * F(p2[1] - p0[1])) - F(F(p2[0] - p0[0]) * F(p1[1] - p0[1])))


def _check_rect_cross(p1, p2, q1, q2):
    """iou3d_cpu.cpp:67-73"""
    return (min(p1[0], p2[0]) <= max(q1[0], q2[0]) and min(q1[0], q2[0]) <= max(p1[0], p2[0]) and
            min(p1[1], p2[1]) <= max(q1[1], q2[1]) and min(q1[1], q2[1]) <= max(p1[1], p2[1]))


def _check_in_box2d(box, p):
    """iou3d_cpu.cpp:75-86"""
    cx, cy = box[0], box[1]
    ang = F(-box[6])
    c, s = F(math.cos(ang)), F(math.sin(ang))
    rx = F(F(F(p[0] - cx) * c) + F(F(p[1] - cy) * F(-s)))
    ry = F(F(F(p[0] - cx) * s) + F(F(p[1] - cy) * c))
    return abs(rx) < F(F(box[3] / F(2)) + MARGIN) and abs(ry) < F(F(box[4] / F(2)) + MARGIN)


def _intersection(p1, p0, q1, q0):
    """iou3d_cpu.cpp:88-116 -> point or None"""
    if not _check_rect_cross(p0, p1, q0, q1):
        return None
    s1 = _cross3(q0, p1, p0)
    s2 = _cross3(p1, q1, p0)
    s3 = _cross3(p0, q1, q0)
    s4 = _cross3(q1, p1, q0)
    if not (F(s1 * s2) > 0 and F(s3 * s4) > 0):
        return None
    s5 = _cross3(q1, p1, p0)
    if abs(F(s5 - s1)) > EPS:
        x = F(F(F(s5 * q0[0]) - F(s1 * q1[0])) / F(s5 - s1))
        y = F(F(F(s5 * q0[1]) - F(s1 * q1[1])) / F(s5 - s1))
    else:
        a0, b0, c0 = F(p0[1] - p1[1]), F(p1[0] - p0[0]), F(F(p0[0] * p1[1]) - F(p1[0] * p0[1]))
        a1, b1, c1 = F(q0[1] - q1[1]), F(q1[0] - q0[0]), F(F(q0[0] * q1[1]) - F(q1[0] * q0[1]))
        D = F(F(a0 * b1) - F(a1 * b0))
        x = F(F(F(b0 * c1) - F(b1 * c0)) / D)
        y = F(F(F(a1 * c0) - F(a0 * c1)) / D)
    return (x, y)


def _corners(box):
    """iou3d_cpu.cpp:133-163: axis-aligned corners rotated about the centre by the heading."""
    cx, cy = box[0], box[1]
    hx, hy = F(box[3] / F(2)), F(box[4] / F(2))
    c, s = F(math.cos(box[6])), F(math.sin(box[6]))
    pts = [(F(cx - hx), F(cy - hy)), (F(cx + hx), F(cy - hy)), (F(cx + hx), F(cy + hy)), (F(cx - hx), F(cy + hy))]
    out = []
    for (px, py) in pts:
        nx = F(F(F(F(px - cx) * c) + F(F(py - cy) * F(-s))) + cx)
        ny = F(F(F(F(px - cx) * s) + F(F(py - cy) * c)) + cy)
        out.append((nx, ny))
    out.append(out[0])
    return out


def box_overlap(box_a, box_b):
    """iou3d_cpu.cpp:127-227: intersection polygon = edge crossings + contained corners, sorted by angle, shoelace."""
    a = [F(v) for v in box_a]
    b = [F(v) for v in box_b]
    ca, cb = _corners(a), _corners(b)
    pts = []
    sx, sy = F(0), F(0)
    for i in range(4):
        for j in range(4):
            p = _intersection(ca[i + 1], ca[i], cb[j + 1], cb[j])
            if p is not None:
                sx, sy = F(sx + p[0]), F(sy + p[1])
                pts.append(p)
    for k in range(4):
        if _check_in_box2d(a, cb[k]):
            sx, sy = F(sx + cb[k][0]), F(sy + cb[k][1])
            pts.append(cb[k])
        if _check_in_box2d(b, ca[k]):
            sx, sy = F(sx + ca[k][0]), F(sy + ca[k][1])
            pts.append(ca[k])
    cnt = len(pts)
    if cnt == 0:
        return F(0)
    cx, cy = F(sx / F(cnt)), F(sy / F(cnt))
    # bubble sort with the reference's comparator (atan2 descending swaps), iou3d_cpu.cpp:207-217
    ang = [F(math.atan2(F(p[1] - cy), F(p[0] - cx))) for p in pts]
    for j in range(cnt - 1):
        for i in range(cnt - j - 1):
            if ang[i] > ang[i + 1]:
                pts[i], pts[i + 1] = pts[i + 1], pts[i]
                ang[i], ang[i + 1] = ang[i + 1], ang[i]
    area = F(0)
    for k in range(cnt - 1):
        area = F(area + _cross2(F(pts[k][0] - pts[0][0]), F(pts[k][1] - pts[0][1]),
                                F(pts[k + 1][0] - pts[0][0]), F(pts[k + 1][1] - pts[0][1])))
    return F(abs(area) / F(2.0))


def iou_bev(box_a, box_b):
    """iou3d_cpu.cpp:229-237"""
    sa = F(F(box_a[3]) * F(box_a[4]))
    sb = F(F(box_b[3]) * F(box_b[4]))
    so = box_overlap(box_a, box_b)
    return F(so / max(F(F(sa + sb) - so), EPS))


def aligned_iou3d(box_a, box_b):
    """iou3d_nms_utils.py:45-87 (boxes_aligned_iou3d_gpu) for one pair of (x, y, z, dx, dy, dz, heading) boxes."""
    a = [F(v) for v in box_a]
    b = [F(v) for v in box_b]
    a_max, a_min = F(a[2] + F(a[5] / F(2))), F(a[2] - F(a[5] / F(2)))
    b_max, b_min = F(b[2] + F(b[5] / F(2))), F(b[2] - F(b[5] / F(2)))
    oh = max(F(min(a_max, b_max) - max(a_min, b_min)), F(0))
    o3 = F(box_overlap(a, b) * oh)
    va, vb = F(F(a[3] * a[4]) * a[5]), F(F(b[3] * b[4]) * b[5])
    return F(o3 / max(F(F(va + vb) - o3), F(1e-6)))


def nms_rotated(boxes, thresh):
    """Greedy NMS over boxes already sorted by descending score (iou3d_nms_kernel.cu:280-324 mask + iou3d_nms.cpp:138-154
    sweep): box i, if not removed, removes every later box j with iou_bev(i, j) > thresh.  Returns kept indices."""
    n = len(boxes)
    removed = np.zeros(n, dtype=bool)
    keep = []
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        for j in range(i + 1, n):
            if not removed[j] and iou_bev(boxes[i], boxes[j]) > F(thresh):
                removed[j] = True
    return keep


def rotate_nms_pcdet(boxes, scores, thresh, pre_maxsize=None, post_max_size=None):
    """box_torch_ops.py:5-31 (torch tensors in, LongTensor of selected indices out)."""
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    b = boxes[order].contiguous().numpy().astype(np.float32)
    keep = nms_rotated(b, thresh) if len(b) else []
    selected = order[torch.tensor(keep, dtype=torch.long)] if len(keep) else order[:0]
    if post_max_size is not None:
        selected = selected[:post_max_size]
    return selected


def decode(preds_dict, out_size_factor, voxel_size, pc_range):
    """centerhead.py:247-304 for one task.  preds_dict: head name -> [B, c, H, W] fp32.
    Returns boxes [B, H*W, 9] (x, y, z, dx, dy, dz, vx, vy, yaw), hm [B, H*W, C] (sigmoid), iou [B, H*W]."""
    p = {k: v.permute(0, 2, 3, 1).contiguous() for k, v in preds_dict.items()}
    hm = torch.sigmoid(p["hm"])
    dim = torch.exp(p["dim"])
    rot = torch.atan2(p["rot"][..., 0:1], p["rot"][..., 1:2])
    B, H, W, C = hm.shape
    if "iou" in p:
        iou = (p["iou"].squeeze(-1) + 1) * 0.5
    else:
        iou = torch.ones((B, H, W), dtype=dim.dtype)
    ys, xs = torch.meshgrid([torch.arange(0, H), torch.arange(0, W)], indexing="ij")
    ys = ys.view(1, H, W).repeat(B, 1, 1).float().view(B, -1, 1) + p["reg"].reshape(B, H * W, 2)[:, :, 1:2]
    xs = xs.view(1, H, W).repeat(B, 1, 1).float().view(B, -1, 1) + p["reg"].reshape(B, H * W, 2)[:, :, 0:1]
    xs = xs * out_size_factor * voxel_size[0] + pc_range[0]
    ys = ys * out_size_factor * voxel_size[1] + pc_range[1]
    boxes = torch.cat([xs, ys, p["height"].reshape(B, H * W, 1), dim.reshape(B, H * W, 3), p["vel"].reshape(B, H * W, 2),
                       rot.reshape(B, H * W, 1)], dim=2)
    return boxes, hm.reshape(B, H * W, C), iou.reshape(B, H * W)


def post_processing(boxes, hm, iou, rectifier, score_threshold, post_center_range, nms_iou_threshold, pre_max, post_max):
    """centerhead.py:332-384 for one task: list (per frame) of dicts box3d_lidar [K,9], scores [K], label_preds [K]."""
    out = []
    pcr = torch.tensor(post_center_range, dtype=hm.dtype)
    for i in range(hm.shape[0]):
        box_preds, hm_preds, iou_preds = boxes[i], hm[i], iou[i].view(-1)
        scores, labels = torch.max(hm_preds, dim=-1)
        mask = (scores > score_threshold) & (box_preds[..., :3] >= pcr[:3]).all(1) & (box_preds[..., :3] <= pcr[3:]).all(1)
        box_preds, scores, labels = box_preds[mask], scores[mask], labels[mask]
        iou_p = torch.clamp(iou_preds[mask], min=0., max=1.)
        rect = torch.tensor(rectifier).to(hm_preds)
        scores = torch.pow(scores, 1 - rect[labels]) * torch.pow(iou_p, rect[labels])
        sel_b, sel_s, sel_l = [torch.zeros((0, 9))], [torch.zeros((0,))], [torch.zeros((0,), dtype=torch.int64)]
        for c in range(hm_preds.shape[-1]):
            m = labels == c
            sc, lc, bc = scores[m], labels[m], box_preds[m]
            sel = rotate_nms_pcdet(bc[:, [0, 1, 2, 3, 4, 5, -1]], sc, nms_iou_threshold[c], pre_max, post_max)
            sel_b.append(bc[sel]); sel_s.append(sc[sel]); sel_l.append(lc[sel])
        out.append(dict(box3d_lidar=torch.cat(sel_b), scores=torch.cat(sel_s), label_preds=torch.cat(sel_l)))
    return out


def predict(preds_dicts, num_classes, test_cfg, rectifier, tokens=None):
    """centerhead.py:231-330.  test_cfg: dict with post_center_limit_range, score_threshold, out_size_factor (per task),
    voxel_size, pc_range, nms = dict(nms_iou_threshold (per task, per class), nms_pre_max_size, nms_post_max_size)."""
    rets = []
    for t, pd in enumerate(preds_dicts):
        boxes, hm, iou = decode(pd, test_cfg["out_size_factor"][t], test_cfg["voxel_size"], test_cfg["pc_range"])
        rets.append(post_processing(boxes, hm, iou, rectifier[t], test_cfg["score_threshold"],
                                    test_cfg["post_center_limit_range"], test_cfg["nms"]["nms_iou_threshold"][t],
                                    test_cfg["nms"]["nms_pre_max_size"], test_cfg["nms"]["nms_post_max_size"]))
    out = []
    for i in range(len(rets[0])):
        flag, labs = 0, []
        for t, nc in enumerate(num_classes):
            labs.append(rets[t][i]["label_preds"] + flag)
            flag += nc
        out.append(dict(box3d_lidar=torch.cat([r[i]["box3d_lidar"] for r in rets]),
                        scores=torch.cat([r[i]["scores"] for r in rets]), label_preds=torch.cat(labs),
                        token=tokens[i] if tokens else None))
    return out
