"""Row F3: GPU label assignment (pnx_assign_labels through ops.assign_labels) against the numpy restatement of the
reference's AssignLabel + collate (pillarnext_b200/synth.assign_labels, itself checked against the reference's own
det3d/datasets/pipelines/assign.py in tests/test_oracle_cpu.py::test_oracle_and_synth_vs_live_reference):
ind / mask / cat bit-exact, heat-maps / anno_box / gt_boxes 1e-6."""
import numpy as np
import pytest
import torch

from pillarnext_b200 import ops, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg_name,n_boxes", [("nusc", 40), ("waymo", 120), ("tiny", 25)])
def test_assign_labels_matches_numpy_reference(cfg_name, n_boxes):
    cfg = {"nusc": synth.NUSC, "waymo": synth.WAYMO_BENCH, "tiny": synth.tiny_config(128, [["car"], ["truck", "construction_vehicle"]])}[cfg_name]
    seeds = [7, 8, 9]
    boxes, cls = synth.make_gt_batch(seeds, n_boxes, cfg)
    # edge cases: an ignored object, a degenerate size, an out-of-range centre, a centre just left of the map (ct in (-1, 0):
    # the reference's int truncation accepts it), NaN velocity stays NaN in anno_box
    cls[0, 0] = -1
    boxes[0, 1, 3] = 0.0
    boxes[1, 0, 0] = cfg["pc_range"][3] + 5.0
    boxes[1, 1, 0] = cfg["pc_range"][0] - 0.3 * cfg["voxel_size"][0] * cfg["out_size_factor"][0]
    got = ops.assign_labels(boxes.cuda(), cls.cuda(), cfg["tasks"], cfg["voxel_size"], cfg["pc_range"], cfg["out_size_factor"])
    names_all = [n for t in cfg["tasks"] for n in t]
    want = []
    for b in range(len(seeds)):
        names = [names_all[c] if c >= 0 else "ignored_class" for c in cls[b].tolist()]
        want.append(synth.assign_labels(boxes[b].numpy(), names, cfg=cfg))
    for t in range(len(cfg["tasks"])):
        for key in ("ind", "mask", "cat"):
            w = torch.stack([torch.tensor(want[b][key][t]) for b in range(len(seeds))])
            assert torch.equal(got[key][t].cpu(), w), (t, key)
        assert int(got["mask"][t].sum()) > 0 or cfg_name == "tiny"
        for key, tol in (("hm", 1e-6), ("anno_box", 2e-6), ("gt_boxes", 0.0)):
            w = torch.stack([torch.tensor(want[b][key][t]) for b in range(len(seeds))])
            g = got[key][t].cpu()
            assert g.shape == w.shape and g.dtype == w.dtype, (t, key, g.shape, w.shape)
            both_nan = torch.isnan(g) & torch.isnan(w)
            d = torch.where(both_nan, torch.zeros_like(g), (g - w).abs())
            assert not torch.isnan(d).any() and d.max().item() <= tol, (t, key, d.max().item())
        assert got["hm"][t].max().item() == 1.0 or int(got["mask"][t].sum()) == 0
