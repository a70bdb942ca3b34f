"""In-repo detection mAP (row X2): the nuScenes detection-benchmark definition, restated from its published description
because the reference delegates evaluation to devkits that are not installable here (det3d/datasets/nuscenes/nusc.py:
204-212 calls nuscenes-devkit; waymo.py:124 writes a .bin for the Waymo tools).

  * a prediction matches the closest not-yet-matched ground-truth box of the same class whose centre is within d metres
    in the ground plane (d in {0.5, 1, 2, 4}); predictions are visited in descending score order over ALL frames;
  * precision is interpolated (monotone envelope) at 101 recall points; AP = mean over recall in (0.1, 1] of
    max(precision - 0.1, 0) / 0.9 (recall and precision below 10 % are clipped, as in the benchmark);
  * mAP = mean over classes and the four thresholds, reported in percent like the reference's README.
The same function scores the product's detections and the CPU oracle's (tests/test_map_gpu.py): the accuracy gate of
BASELINE.json is the DIFFERENCE of the two on held-out synthetic scenes."""
import numpy as np

DIST_THS = (0.5, 1.0, 2.0, 4.0)


def _ap(scores, tp, n_gt, min_recall=0.1, min_precision=0.1):
    if n_gt == 0 or len(scores) == 0:
        return 0.0
    order = np.argsort(-np.asarray(scores), kind="stable")
    tp = np.asarray(tp, dtype=np.float64)[order]
    ctp, cfp = np.cumsum(tp), np.cumsum(1.0 - tp)
    rec, prec = ctp / n_gt, ctp / (ctp + cfp)
    rec_pts = np.linspace(0, 1, 101)
    p = np.interp(rec_pts, rec, prec, right=0.0)                    # nuScenes: interpolate, 0 beyond the reached recall
    p = p[round(100 * min_recall) + 1:]
    p = np.clip(p - min_precision, 0, None)
    return float(p.mean() / (1.0 - min_precision))


def detection_map(gts, detections, class_names, dist_ths=DIST_THS):
    """gts: token -> {"boxes" [M, >=7] (x, y, ...), "names" list[str]};
    detections: token -> {"box3d_lidar" [K, 9], "scores" [K], "label_preds" [K]} (the reference's validation_step output,
    single_stage.py:47-59; tensors or arrays).  Returns {"mAP": percent, "per_class": {name: AP percent}}."""
    def arr(x):
        return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)

    per_class = {}
    for ci, cname in enumerate(class_names):
        aps = []
        gt_xy = {t: np.asarray(g["boxes"], dtype=np.float64)[[i for i, n in enumerate(g["names"]) if n == cname]][:, :2]
                 for t, g in gts.items()}
        n_gt = sum(len(v) for v in gt_xy.values())
        preds = []
        for t, d in detections.items():
            lab = arr(d["label_preds"])
            sel = np.nonzero(lab == ci)[0]
            bx, sc = arr(d["box3d_lidar"]), arr(d["scores"])
            preds += [(float(sc[i]), t, float(bx[i, 0]), float(bx[i, 1])) for i in sel]
        preds.sort(key=lambda r: -r[0])
        for th in dist_ths:
            taken = {t: np.zeros(len(v), dtype=bool) for t, v in gt_xy.items()}
            tp = []
            for _, t, x, y in preds:
                g = gt_xy.get(t)
                hit = False
                if g is not None and len(g):
                    dist = np.hypot(g[:, 0] - x, g[:, 1] - y)
                    dist[taken[t]] = np.inf
                    j = int(np.argmin(dist))
                    if dist[j] < th:
                        taken[t][j] = True
                        hit = True
                tp.append(1.0 if hit else 0.0)
            aps.append(_ap([p[0] for p in preds], tp, n_gt))
        per_class[cname] = 100.0 * float(np.mean(aps)) if n_gt else float("nan")
    vals = [v for v in per_class.values() if v == v]
    return {"mAP": float(np.mean(vals)) if vals else 0.0, "per_class": per_class}
