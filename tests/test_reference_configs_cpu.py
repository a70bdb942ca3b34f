"""Drop-in check at the CONFIG level, CPU only, build container only.

hydra/omegaconf are absent, so a ~60-line restatement of what `hydra.main` + `OmegaConf.resolve` + `instantiate` do with
the reference's experiment files is used (defaults lists with `@package` targets, `_self_`, `${a.b[1]}` interpolation,
`_target_` / `_recursive_`).  The reference's OWN yaml files are read from /root/reference/configs and
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


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _load(path):
    """One config file with its `defaults` list composed (hydra semantics restricted to what the reference uses)."""
    with open(path) as fh:
        node = yaml.safe_load(fh) or {}
    defaults = node.pop("defaults", [])
    out, self_done = {}, False
    for d in defaults:
        if d == "_self_":
            _merge(out, node)
            self_done = True
            continue
        (key, name), = d.items() if isinstance(d, dict) else ((d, None),)
        group, _, pkg = key.partition("@")
        fpath = os.path.normpath(os.path.join(os.path.dirname(path), group, name if name else "")) if name else \
            os.path.normpath(os.path.join(os.path.dirname(path), group))
        sub = _load(fpath + ".yaml")
        if not pkg:                      # default package = the group's last path component
            pkg = os.path.basename(group) if name else ""
        tgt = out
        for part in [p for p in pkg.split(".") if p]:
            tgt = tgt.setdefault(part, {})
        _merge(tgt, sub)
    if not self_done:
        _merge(out, node)
    return out


def _lookup(root, expr):
    cur = root
    for part in re.findall(r"[^.\[\]]+", expr):
        cur = cur[int(part)] if isinstance(cur, list) else cur[part]
    return cur


def _resolve(node, root):
    if isinstance(node, dict):
        return {k: _resolve(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root) for v in node]
    if isinstance(node, str):
        m = re.fullmatch(r"\$\{([^}]+)\}", node.strip())
        if m:
            return _resolve(_lookup(root, m.group(1)), root)
    return node


def _instantiate(node):
    if isinstance(node, dict) and "_target_" in node:
        kw = {k: v for k, v in node.items() if k not in ("_target_", "_recursive_", "_partial_")}
        if node.get("_recursive_", True):
            kw = {k: _instantiate(v) for k, v in kw.items()}
        mod, _, cls = node["_target_"].rpartition(".")
        return getattr(importlib.import_module(mod), cls)(**kw)
    if isinstance(node, dict):
        return {k: _instantiate(v) for k, v in node.items()}
    return node


def compose(experiment):
    path = os.path.join(CFG_ROOT, "experiments", experiment + ".yaml")
    with open(path) as fh:
        raw = yaml.safe_load(fh)
    # only the model-side defaults are composed: the dataset/dataloader groups need packages that are out of scope
    raw["defaults"] = [d for d in raw["defaults"] if d == "_self_" or "models/" in next(iter(d))]
    node = {k: v for k, v in raw.items()}
    defaults = node.pop("defaults")
    out = {}
    for d in defaults:
        if d == "_self_":
            _merge(out, node)
            continue
        (key, name), = d.items()
        group, _, pkg = key.partition("@")
        sub = _load(os.path.normpath(os.path.join(os.path.dirname(path), group, name)) + ".yaml")
        tgt = out
        for part in pkg.split("."):
            tgt = tgt.setdefault(part, {})
        _merge(tgt, sub)
    out["model"] = _resolve(out["model"], out)      # the data group is not composed, so only the model node is resolved
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
