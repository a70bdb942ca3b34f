"""Host-side runtime around the hot path (rows F4 / X2), CPU only: the hydra stand-in on this repo's config tree, the
synthetic scene dataset + collate contract, the mAP evaluator's known answers, the checkpoint format."""
import os

import numpy as np
import torch

from pillarnext_b200 import evaluate, hydra_lite, scenes, trainer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "configs", "experiments")


def test_compose_resolve_override_instantiate():
    cfg = hydra_lite.main(CFG, "synth_det_pp18_aspp", ["trainer.max_epochs=3", "dataloader.train.batch_size=2", "+load_from=x.pth"])
    assert cfg.trainer.max_epochs == 3 and cfg.scheduler.epochs == 3                      # interpolation after the override
    assert cfg.dataloader.train.batch_size == 2 and cfg.load_from == "x.pth"
    assert cfg.model.head.tasks == [["car"], ["truck", "pedestrian"]] and cfg.model.backbone.num_input_features == 64
    assert cfg.model.head.out_size_factor == [4, 4] and cfg.data.val_dataset.cfg.pc_range == cfg.model.reader.pc_range
    model = hydra_lite.instantiate(cfg.model)
    from pillarnext_b200 import modules
    assert isinstance(model, modules.SingleStageDetector) and len(model.head.tasks) == 2
    opt = hydra_lite.instantiate(dict(cfg.optimizer, fused=False), params=model.parameters())
    sched = hydra_lite.instantiate(dict(cfg.scheduler, _recursive_=False), optimizer=opt, steps_per_epoch=10)
    assert isinstance(opt, torch.optim.AdamW) and isinstance(sched, torch.optim.lr_scheduler.OneCycleLR)
    part = hydra_lite.instantiate({"_target_": "torch.zeros", "_partial_": True, "dtype": None})
    assert part(3).shape == (3,)


def test_scene_dataset_and_collate_contract():
    cfg = hydra_lite.main(CFG, "synth_det_pp18_aspp")
    ds = hydra_lite.instantiate(cfg.data.val_dataset)
    a, b = ds[0], ds[0]
    assert np.array_equal(a["points"], b["points"])                                        # reproducible
    assert a["points"].dtype == np.float32 and a["points"].shape[1] == 5
    boxes = a["gt_boxes_raw"]
    inside = 0
    for bx in boxes:                                                                       # every box owns a cluster of returns
        d = np.hypot(a["points"][:, 0] - bx[0], a["points"][:, 1] - bx[1])
        inside += int((d < 0.75 * max(bx[3], bx[4])).sum() >= 15)
    assert inside >= len(boxes) - 1
    batch = scenes.collate([ds[0], ds[1], ds[2]])
    assert batch["points"].shape[1] == 6 and set(batch["points"][:, 0].tolist()) == {0.0, 1.0, 2.0}
    assert batch["gt_boxes_raw"].shape[:2] == batch["gt_classes"].shape and batch["gt_classes"].dtype == torch.int32
    assert len(batch["token"]) == 3 and ds.ground_truth()[batch["token"][0]]["names"]


def test_map_known_answers():
    names = ["car", "pedestrian"]
    gts = {"a": {"boxes": np.array([[0.0, 0, 0], [10.0, 0, 0], [0.0, 10, 0]]), "names": ["car", "car", "pedestrian"]},
           "b": {"boxes": np.array([[5.0, 5, 0]]), "names": ["car"]}}

    def det(boxes, scores, labels):
        b = np.zeros((len(boxes), 9))
        b[:, :2] = boxes
        return {"box3d_lidar": torch.tensor(b), "scores": torch.tensor(scores), "label_preds": torch.tensor(labels)}

    perfect = {"a": det([[0, 0], [10, 0], [0, 10]], [0.9, 0.8, 0.7], [0, 0, 1]), "b": det([[5, 5]], [0.6], [0])}
    assert abs(evaluate.detection_map(gts, perfect, names)["mAP"] - 100.0) < 1e-6
    # 0.8 m off: misses only the 0.5 m threshold -> 75 %
    off = {"a": det([[0.8, 0], [10.8, 0], [0.8, 10]], [0.9, 0.8, 0.7], [0, 0, 1]), "b": det([[5.8, 5]], [0.6], [0])}
    assert abs(evaluate.detection_map(gts, off, names)["mAP"] - 75.0) < 1e-6
    # wrong class -> 0 for that class; a duplicate of a matched box is a false positive, not a second match
    r = evaluate.detection_map(gts, {"a": det([[0, 0], [0, 0.1], [0, 10]], [0.9, 0.8, 0.7], [0, 0, 0]), "b": det([[5, 5]], [0.95], [0])}, names)
    assert r["per_class"]["pedestrian"] == 0.0 and 0.0 < r["per_class"]["car"] < 100.0
    # one of three cars found, at the top score: recall stops at 1/3 -> AP = (1/3 - 0.1) / 0.9 of the recall axis at precision 1
    one = evaluate.detection_map(gts, {"a": det([[0, 0]], [0.9], [0])}, ["car"])
    assert abs(one["mAP"] - 100.0 * (33 - 10) / 90) < 1.0


def test_checkpoint_format_roundtrip(tmp_path):
    net = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.BatchNorm1d(4))
    opt = torch.optim.AdamW(net.parameters())
    f = str(tmp_path / "epoch_1.pth")
    trainer.save_checkpoint(net, f, opt, None, dict(epoch=1, iter=10))
    ck = torch.load(f, weights_only=False)
    assert set(ck) >= {"meta", "state_dict", "optimizer"} and ck["meta"] == dict(epoch=1, iter=10)
    # a DDP-saved checkpoint ("module." prefix) loads strict=True like the reference's load_checkpoint
    torch.save({"state_dict": {"module." + k: v for k, v in ck["state_dict"].items()}}, f)
    net2 = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.BatchNorm1d(4))
    trainer.load_checkpoint(net2, f, strict=True)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
