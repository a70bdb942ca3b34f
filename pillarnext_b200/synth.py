"""Synthetic nuScenes-/Waymo-shaped inputs: seeded point clouds, GT boxes and
CenterPoint label tensors in the reference's collate format.  No dataset, devkit or network needed.

Label assignment restates reference det3d/datasets/pipelines/assign.py:23-116 + center_utils.py:12-60
(gaussian heat-map splat, ind/mask/cat/anno_box/gt_boxes) for the synthetic boxes only; batching
restates det3d/datasets/loader/collate.py:6-35.
"""
import numpy as np
import torch

NUSC = dict(
    voxel_size=[0.075, 0.075, 8], pc_range=[-50.4, -50.4, -5.0, 50.4, 50.4, 3.0],
    tasks=[["car"], ["truck", "construction_vehicle"], ["bus", "trailer"], ["barrier"],
           ["motorcycle", "bicycle"], ["pedestrian", "traffic_cone"]],
    common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
    strides=[1, 2, 2, 2], weight=0.25, code_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2, 1.0, 1.0],
    out_size_factor=[4, 4, 4, 4, 4, 4], head_strides=[2, 2, 2, 2, 2, 2], with_reg_iou=True,
    intensity_max=255.0, xy_extent=54.0, z_range=(-5.0, 3.0),
)
NUSC["post_processing"] = dict(          # configs/experiments/nusc_det_pp18_aspp_iou_sp.yaml:39-50
    post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], score_threshold=0.1,
    nms=dict(nms_pre_max_size=1000, nms_post_max_size=83,
             nms_iou_threshold=[[0.2], [0.2, 0.2], [0.2, 0.2], [0.2], [0.2, 0.2], [0.2, 0.2]]),
    pc_range=NUSC["pc_range"], voxel_size=NUSC["voxel_size"], out_size_factor=NUSC["out_size_factor"])
NUSC["frames_per_gpu"] = 6               # docs/RUN.md:9

# BASELINE.json configs[3]/[4]: the reference's Waymo experiment (configs/experiments/waymo_det_pp18_aspp_iou_car_sp.yaml:
# 2 tasks, `iou` head, weight 1, rectifier, nms_pre_max_size 4096) at the benchmark shape 0.1 m / +-75.2 m -> 1504^2 BEV
# (the reference yaml itself says 0.075 m / +-76.8 m -> 2048^2 = WAYMO_REF below).
WAYMO_BENCH = dict(
    voxel_size=[0.1, 0.1, 20], pc_range=[-75.2, -75.2, -10.0, 75.2, 75.2, 10.0],
    tasks=[["vehicle"], ["pedestrian", "cyclist"]],
    common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2), "iou": (1, 2)},
    strides=[1, 2, 2, 2], weight=1.0, code_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2, 1.0, 1.0],
    out_size_factor=[4, 4], head_strides=[2, 2], with_reg_iou=True, rectifier=[[0.68], [0.71, 0.65]],
    intensity_max=1.0, xy_extent=80.0, z_range=(-3.0, 5.0), frames_per_gpu=3,          # docs/RUN.md:34
)
WAYMO_BENCH["post_processing"] = dict(
    post_center_limit_range=[-80.0, -80.0, -10.0, 80.0, 80.0, 10.0], score_threshold=0.1,
    nms=dict(nms_pre_max_size=4096, nms_post_max_size=500, nms_iou_threshold=[[0.7], [0.2, 0.25]]),
    pc_range=WAYMO_BENCH["pc_range"], voxel_size=WAYMO_BENCH["voxel_size"], out_size_factor=WAYMO_BENCH["out_size_factor"])
WAYMO_REF = dict(WAYMO_BENCH)
WAYMO_REF.update(voxel_size=[0.075, 0.075, 20], pc_range=[-76.8, -76.8, -10.0, 76.8, 76.8, 10.0])
WAYMO_REF["post_processing"] = dict(WAYMO_BENCH["post_processing"], pc_range=WAYMO_REF["pc_range"], voxel_size=WAYMO_REF["voxel_size"])

# bench.py --config: BASELINE.json configs[1..4]
BENCH_CONFIGS = {
    "nusc": dict(cfg=NUSC, points=30000, sweeps=10, rings=32, label="PillarNeXt-B nuScenes-shape synthetic (BASELINE.json configs[1]): "
                 "%d pts/frame, 0.075 m pillars, 1344^2 BEV, 6 tasks"),
    "waymo180k": dict(cfg=WAYMO_BENCH, points=180000, sweeps=1, rings=64, label="PillarNeXt-B Waymo-shape synthetic (BASELINE.json "
                      "configs[3]): %d pts/frame, 0.1 m pillars, 1504^2 BEV, 2 tasks + iou head"),
    "waymo540k": dict(cfg=WAYMO_BENCH, points=540000, sweeps=3, rings=64, label="PillarNeXt-B Waymo 3-frame multi-sweep concat "
                      "(BASELINE.json configs[4]): %d pts/frame, 0.1 m pillars, 1504^2 BEV, 2 tasks + iou head"),
}


def tiny_config(grid=64, tasks=None):
    """Small grid with the same structure (grid divisible by 8) for oracle-speed parity tests."""
    vs = 0.5
    half = grid * vs / 2
    cfg = dict(NUSC)
    cfg.update(voxel_size=[vs, vs, 8], pc_range=[-half, -half, -5.0, half, half, 3.0], xy_extent=half * 1.08)
    if tasks is not None:
        cfg.update(tasks=tasks, out_size_factor=[4] * len(tasks), head_strides=[2] * len(tasks))
    cfg["post_processing"] = dict(post_center_limit_range=[-half * 1.2, -half * 1.2, -10.0, half * 1.2, half * 1.2, 10.0],
                                  score_threshold=0.1, pc_range=cfg["pc_range"], voxel_size=cfg["voxel_size"],
                                  out_size_factor=cfg["out_size_factor"],
                                  nms=dict(nms_pre_max_size=1000, nms_post_max_size=83,
                                           nms_iou_threshold=[[0.2] * len(t) for t in cfg["tasks"]]))
    return cfg


def make_frame(seed, n_points, cfg=NUSC, kind="uniform", sweeps=1, rings=32):
    """[n,5] fp32 (x,y,z,intensity,time).  'uniform': x,y ~ U(+-extent) (some out of range);
    'lidar': `rings` beams hitting a ground plane + box clutter; sweeps > 1 with rings = 64 (Waymo multi-sweep concat,
    det3d/datasets/waymo/waymo.py:49-67): every sweep is the same scan rigidly shifted, time = 0.1 * sweep."""
    g = np.random.default_rng(1234 + seed)
    ext = cfg["xy_extent"]
    if kind == "uniform":
        xy = g.uniform(-ext, ext, size=(n_points, 2))
        z = g.uniform(cfg["z_range"][0], cfg["z_range"][1], size=(n_points, 1))
    else:
        n_ring = rings
        per = n_points // n_ring
        elev = np.deg2rad(np.linspace(-30.0, -0.6, n_ring))
        rng_ = np.clip(1.84 / np.tan(-elev), 1.5, ext * 1.1)
        pts = []
        for r in range(n_ring):
            az = g.uniform(0, 2 * np.pi, per)
            rr = rng_[r] + g.normal(0, 0.02 * rng_[r], per)
            pts.append(np.stack([rr * np.cos(az), rr * np.sin(az), -1.84 + g.normal(0, 0.02, per)], 1))
        p = np.concatenate(pts, 0)
        rest = n_points - p.shape[0]
        if rest > 0:
            c = g.uniform(-ext * 0.6, ext * 0.6, size=(20, 2))
            k = g.integers(0, 20, rest)
            q = np.concatenate([c[k] + g.normal(0, 1.0, (rest, 2)), g.uniform(-1.8, 0.5, (rest, 1))], 1)
            p = np.concatenate([p, q], 0)
        xy, z = p[:, :2], p[:, 2:3]
    inten = g.uniform(0, cfg["intensity_max"], size=(n_points, 1))
    sw = g.integers(0, sweeps, size=(n_points, 1))
    if rings == 64 and sweeps > 1:                                   # Waymo: past sweeps pose-aligned = rigid shift of the scan
        xy = xy + sw * g.normal(0, 0.4, size=(1, 2))
        t = sw.astype(np.float64) * 0.1
        inten = np.tanh(inten * 3.0)                                 # waymo_convert.py:31
    else:
        t = sw.astype(np.float64) * 0.05
    return np.concatenate([xy, z, inten, t], 1).astype(np.float32)


def collate_points(frames):
    """collate.py:15-22: prepend the frame index as a float column."""
    out = [np.pad(f, ((0, 0), (1, 0)), mode="constant", constant_values=i) for i, f in enumerate(frames)]
    return torch.tensor(np.concatenate(out, 0))


def make_gt(seed, n_boxes, cfg=NUSC):
    """Random GT: boxes [n, 9] (x,y,z,dx,dy,dz,vx,vy,yaw) + class names drawn from the tasks."""
    g = np.random.default_rng(4321 + seed)
    names_all = [n for t in cfg["tasks"] for n in t]
    pr = cfg["pc_range"]
    xy = g.uniform(pr[0] * 0.95, pr[3] * 0.95, size=(n_boxes, 2))
    z = g.uniform(-2.0, 0.0, size=(n_boxes, 1))
    dim = np.exp(g.uniform(np.log(0.5), np.log(6.0), size=(n_boxes, 3)))
    vel = g.normal(0, 2.0, size=(n_boxes, 2))
    vel[g.uniform(size=n_boxes) < 0.15] = np.nan                      # nuScenes has NaN velocities
    yaw = g.uniform(-np.pi, np.pi, size=(n_boxes, 1))
    names = [names_all[i] for i in g.integers(0, len(names_all), n_boxes)]
    return np.concatenate([xy, z, dim, vel, yaw], 1).astype(np.float32), names


def _gaussian_radius(det_size, min_overlap):
    h, w = det_size
    b1 = h + w
    c1 = w * h * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + np.sqrt(b1 ** 2 - 4 * c1)) / 2
    b2 = 2 * (h + w)
    c2 = (1 - min_overlap) * w * h
    r2 = (b2 + np.sqrt(b2 ** 2 - 16 * c2)) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (h + w)
    c3 = (min_overlap - 1) * w * h
    r3 = (b3 + np.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
    return min(r1, r2, r3)


def _draw_gaussian(hm, center, radius):
    d = 2 * radius + 1
    sigma = d / 6
    m = (d - 1.0) / 2.0
    y, x = np.ogrid[-m:m + 1, -m:m + 1]
    gk = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
    gk[gk < np.finfo(gk.dtype).eps * gk.max()] = 0
    cx, cy = int(center[0]), int(center[1])
    H, W = hm.shape
    left, right = min(cx, radius), min(W - cx, radius + 1)
    top, bottom = min(cy, radius), min(H - cy, radius + 1)
    mh = hm[cy - top:cy + bottom, cx - left:cx + right]
    mg = gk[radius - top:radius + bottom, radius - left:radius + right]
    if min(mg.shape) > 0 and min(mh.shape) > 0:
        np.maximum(mh, mg, out=mh)


def assign_labels(gt_boxes, gt_names, cfg=NUSC, max_objs=500, gaussian_overlap=0.1, min_radius=2):
    """assign.py:23-116 for one frame -> dict of per-task lists (hm, anno_box, ind, mask, cat, gt_boxes)."""
    tasks = cfg["tasks"]
    vs = np.array(cfg["voxel_size"], dtype=np.float64)
    pr = np.array(cfg["pc_range"], dtype=np.float64)
    osf = np.array(cfg["out_size_factor"])
    grid = np.round((pr[3:] - pr[:3]) / vs).astype(np.int64)
    name2id = {n: (ti, ni) for ti, t in enumerate(tasks) for ni, n in enumerate(t)}
    hms, annos, inds, masks, cats, gtbs = [], [], [], [], [], []
    for ti, t in enumerate(tasks):
        fm = grid[:2] // osf[ti]
        hms.append(np.zeros((len(t), fm[1], fm[0]), dtype=np.float32))
        annos.append(np.zeros((max_objs, 10), dtype=np.float32))
        inds.append(np.zeros((max_objs,), dtype=np.int64))
        masks.append(np.zeros((max_objs,), dtype=np.uint8))
        cats.append(np.zeros((max_objs,), dtype=np.int64))
        gtbs.append(np.zeros((max_objs, 7), dtype=np.float32))
    nums = np.zeros(len(tasks), dtype=np.int64)
    for k, name in enumerate(gt_names):
        if name not in name2id:
            continue
        ti, ci = name2id[name]
        bx = gt_boxes[k]
        sx = bx[3] / vs[0] / osf[ti]
        sy = bx[4] / vs[1] / osf[ti]
        if not (sx > 0 and sy > 0):
            continue
        radius = max(min_radius, int(_gaussian_radius((sy, sx), gaussian_overlap)))
        cx = (bx[0] - pr[0]) / vs[0] / osf[ti]
        cy = (bx[1] - pr[1]) / vs[1] / osf[ti]
        ct = np.array([cx, cy], dtype=np.float32)
        cti = ct.astype(np.int32)
        if not (0 <= cti[0] < hms[ti].shape[2] and 0 <= cti[1] < hms[ti].shape[1]):
            continue
        _draw_gaussian(hms[ti][ci], ct, radius)
        j = nums[ti]
        if j >= max_objs:
            continue
        cats[ti][j] = ci
        inds[ti][j] = cti[1] * hms[ti].shape[2] + cti[0]
        masks[ti][j] = 1
        annos[ti][j] = np.concatenate((ct - cti, bx[2:3], np.log(bx[3:6]), bx[6:8], np.sin(bx[8:9]), np.cos(bx[8:9])))
        gtbs[ti][j] = np.concatenate((bx[0:6], bx[8:9]))
        nums[ti] += 1
    return dict(hm=hms, anno_box=annos, ind=inds, mask=masks, cat=cats, gt_boxes=gtbs)


def make_gt_batch(seeds, n_boxes, cfg=NUSC):
    """Raw ground truth of a batch for the GPU label assignment (ops.assign_labels): gt_boxes [B, N, 9] fp32 and
    gt_cls [B, N] int32 (index into the flattened class list of cfg['tasks'])."""
    names_all = [n for t in cfg["tasks"] for n in t]
    boxes, cls = [], []
    for s in seeds:
        b, names = make_gt(s, n_boxes, cfg)
        boxes.append(b)
        cls.append(np.array([names_all.index(n) for n in names], dtype=np.int32))
    return torch.tensor(np.stack(boxes)), torch.tensor(np.stack(cls))


def make_batch(seeds, n_points, cfg=NUSC, kind="uniform", n_boxes=40, sweeps=1, with_labels=True, rings=32):
    """A collated `example` dict in the reference's format (collate.py): points [sumN, 6] + per-task label lists."""
    frames = [make_frame(s, n_points, cfg, kind, sweeps, rings) for s in seeds]
    ex = {"points": collate_points(frames), "token": ["synthetic_%d" % s for s in seeds]}
    if with_labels:
        labs = [assign_labels(*make_gt(s, n_boxes, cfg), cfg=cfg) for s in seeds]
        for key in ("hm", "anno_box", "ind", "mask", "cat", "gt_boxes"):
            ex[key] = [torch.stack([torch.tensor(l[key][t]) for l in labs]) for t in range(len(cfg["tasks"]))]
    return ex
