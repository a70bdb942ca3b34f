"""Drop-in check at the CONFIG level, CPU only, build container only.

hydra/omegaconf are absent, so pillarnext_b200/hydra_lite.py -- a restatement of what `hydra.main` + `OmegaConf.resolve` +
`instantiate` do with the reference's experiment files (defaults lists with `@package` targets, `_self_`, `${a.b[1]}`
interpolation, `_target_` / `_recursive_`) -- is used.  The reference's OWN yaml files are read from /root/reference/configs and
`instantiate(cfg.model)` (tools/train.py:53, tools/test.py:49) is replayed against this repo's `det3d` package: the
`_target_` strings must resolve, the constructors must accept the reference's kwargs (lists as given by yaml), and the
state-dict keys/shapes of reader / neck / head must equal those of the reference's own modules built from the same
composed config.
"""
import importlib
import os
import re

import pytest
import torch
import yaml

from oracle import reference_loader

CFG_ROOT = os.path.join(reference_loader.REF_ROOT, "configs")


from pillarnext_b200 import hydra_lite

_instantiate = hydra_lite.instantiate


def compose(experiment):
    # only the model-side defaults are composed: the dataset/dataloader groups need packages that are out of scope
    out = hydra_lite.compose(os.path.join(CFG_ROOT, "experiments"), experiment, only_groups=["models/"])
    out["model"] = hydra_lite.resolve(out["model"], out)      # the data group is not composed, so only the model node is resolved
    return out


@pytest.mark.skipif(not os.path.isdir(CFG_ROOT), reason="reference configs not present (build container only)")
@pytest.mark.parametrize("experiment,n_tasks,has_iou", [("nusc_det_pp18_aspp_iou_sp", 6, False),
                                                         ("waymo_det_pp18_aspp_iou_car_sp", 2, True)])
def test_reference_experiment_instantiates_our_modules(experiment, n_tasks, has_iou):
    cfg = compose(experiment)
    m = cfg["model"]
    assert m["_target_"] == "det3d.models.detectors.single_stage.SingleStageDetector"
    assert m["backbone"]["num_input_features"] == 64 and len(m["head"]["tasks"]) == n_tasks      # interpolations resolved
    assert m["post_processing"]["out_size_factor"] == [4] * n_tasks
    model = _instantiate(m)                                           # tools/train.py:53 `instantiate(cfg.model)`
    from pillarnext_b200 import modules
    assert isinstance(model, modules.SingleStageDetector) and isinstance(model.head, modules.CenterHead)
    assert model.post_processing["nms"]["nms_pre_max_size"] >= model.post_processing["nms"]["nms_post_max_size"]
    assert ("iou" in model.head.tasks[0].heads) == has_iou
    assert [float(r) for r in model.head.rectifier[0]] == [float(r) for r in m["head"]["rectifier"][0]]
    # the reference's own reader / neck / head built from the same composed node: identical parameter/buffer names+shapes
    ref = reference_loader.load_reference()
    kw = lambda n: {k: v for k, v in m[n].items() if not k.startswith("_")}
    ref_parts = dict(reader=ref.PillarFeatureNet(**kw("reader")), neck=ref.ASPPNeck(**kw("neck")),
                     head=ref.CenterHead(**kw("head")))
    ours = model.state_dict()
    for part, mod in ref_parts.items():
        want = {part + "." + k: tuple(v.shape) for k, v in mod.state_dict().items()}
        got = {k: tuple(v.shape) for k, v in ours.items() if k.startswith(part + ".")}
        assert got == want, (part, sorted(set(got) ^ set(want))[:6])
    n_params = sum(p.numel() for p in model.parameters())
    assert n_params == (10_379_782 if experiment.startswith("nusc") else n_params)             # SURVEY section 8a count
