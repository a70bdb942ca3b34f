"""Rows F4 + X2 on the GPU: the training entry point end to end, and the accuracy gate of BASELINE.json.

`tools/train.py` (the reference's flow: compose config -> datasets/loaders -> instantiate(cfg.model) -> AdamW + OneCycleLR
-> Trainer.fit) trains PillarNeXt-B for a few hundred iterations on synthetic scenes whose boxes are real point clusters;
`tools/test.py` reloads the checkpoint (strict) and evaluates.  The accuracy gate: the SAME trained weights are run by the
fp32 CPU oracle (reference algorithm) on the held-out scenes, both sets of detections are scored by the in-repo
nuScenes-style mAP, and |mAP(product) - mAP(oracle)| must stay within 0.1 mAP point for the fp32-grade path; the bf16
production path's difference is reported and bounded more loosely."""
import os
import sys

import pytest
import torch

from oracle import pillarnext_oracle as O
from oracle import predict_oracle as P
from pillarnext_b200 import functional as Fn
from pillarnext_b200 import hydra_lite, scenes
from pillarnext_b200.trainer import example_to_device

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def oracle_detections(cfg, sd, dataset, batch=4):
    head = cfg.model.head
    ocfg = dict(voxel_size=list(cfg.model.reader.voxel_size), pc_range=list(cfg.model.reader.pc_range), strides=[1, 2, 2, 2],
                tasks=[list(t) for t in head.tasks], common_heads={k: tuple(v) for k, v in head.common_heads.items()})
    pp = cfg.model.post_processing
    tcfg = dict(post_center_limit_range=list(pp.post_center_limit_range), score_threshold=pp.score_threshold,
                out_size_factor=list(pp.out_size_factor), voxel_size=list(pp.voxel_size)[:2], pc_range=list(pp.pc_range)[:2],
                nms=dict(nms_iou_threshold=[list(t) for t in pp.nms.nms_iou_threshold], nms_pre_max_size=pp.nms.nms_pre_max_size,
                         nms_post_max_size=pp.nms.nms_post_max_size))
    dets = {}
    with torch.no_grad():
        for i0 in range(0, len(dataset), batch):
            ex = scenes.collate([dataset[i] for i in range(i0, min(len(dataset), i0 + batch))])
            preds = O.detector_forward(ex["points"], sd, ocfg, len(ex["token"]), train=False)
            outs = P.predict([{k: v.contiguous() for k, v in pd.items()} for pd in preds], [len(t) for t in head.tasks], tcfg,
                             [list(r) for r in head.rectifier], tokens=ex["token"])
            for o in outs:
                dets[o["token"]] = o
    return dets


def test_train_entry_point_and_map_gate(tmp_path):
    import importlib.util
    import train as train_tool
    spec = importlib.util.spec_from_file_location("pnx_tools_test", os.path.join(ROOT, "tools", "test.py"))
    test_tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(test_tool)
    work = str(tmp_path)
    overrides = ["trainer.max_epochs=6", "data.train_dataset.n_frames=640", "data.val_dataset.n_frames=24", "dataloader.train.num_workers=0",
                 "dataloader.val.num_workers=0", "dataloader.train.batch_size=4"]
    tr = train_tool.main(["--config-name", "synth_det_pp18_aspp", "--work-dir", work] + overrides)
    assert tr.epoch == 6 and os.path.exists(os.path.join(work, "epoch_6.pth"))
    map_train_end = tr.last_eval["mAP"]
    # tools/test.py: strict reload + validation epoch reproduces the number
    res = test_tool.main(["--config-name", "synth_det_pp18_aspp", "--work-dir", work, "+load_from=" + os.path.join(work, "epoch_6.pth")] + overrides)
    assert abs(res["mAP"] - map_train_end) < 1e-6
    cfg = hydra_lite.main(os.path.join(ROOT, "configs", "experiments"), "synth_det_pp18_aspp", overrides)
    val = hydra_lite.instantiate(cfg.data.val_dataset)
    model = tr.model.eval()
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    m_oracle = val.evaluation(oracle_detections(cfg, sd, val))["mAP"]

    def product_map():
        dets = {}
        with torch.no_grad():
            for i0 in range(0, len(val), 4):
                ex = scenes.collate([val[i] for i in range(i0, min(len(val), i0 + 4))])
                dets.update(model(example_to_device(ex, torch.device("cuda"))))
        return val.evaluation(dets)["mAP"]

    m_bf16 = product_map()
    with Fn.precision("split"):
        m_split = product_map()
    msg = "mAP on %d held-out scenes: oracle (fp32 CPU) %.3f | product fp32-grade %.3f | product bf16 %.3f" % (len(val), m_oracle, m_split, m_bf16)
    print(msg)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        open(os.path.join(out, "map_gate.txt"), "w").write(msg + "\n")
    assert m_oracle > 5.0, "the detector did not learn the synthetic task: " + msg     # the gate is meaningless at ~0 mAP
    assert abs(m_split - m_oracle) <= 0.1, msg
    assert abs(m_bf16 - m_oracle) <= 1.5, msg
