"""A minimal hydra / OmegaConf stand-in for the reference's entry points (row F4 of SURVEY.md section 8f).

The reference's tools/train.py and tools/test.py are driven by hydra (`@hydra.main(config_path=..., config_name=...)`,
`OmegaConf.resolve`, `hydra.utils.instantiate`; tools/train.py:16-19,44-68).  hydra and omegaconf are not installable here
(no network), so the subset of their semantics that the reference's config tree uses is restated on PyYAML:

  * defaults lists: `- group: name`, `- ../group@package: name`, `- _self_`                    (compose)
  * `# @package x` headers of group files (the reference's optimizer / scheduler / trainer / dataloader groups)
  * `${a.b[1]}` interpolations, resolved against the root after composition                    (resolve)
  * `_target_` / `_partial_` / `_recursive_: False` instantiation with keyword overrides        (instantiate)
  * dotted command-line overrides `a.b=c` and `+a.b=c`                                          (apply_overrides)

It is host-side configuration plumbing only: nothing here touches the GPU path."""
import functools
import importlib
import os
import re

import yaml


class Cfg(dict):
    """dict with attribute access (what the reference code expects from a DictConfig: cfg.model, cfg.trainer.max_epochs)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(node):
    if isinstance(node, dict):
        return Cfg({k: _wrap(v) for k, v in node.items()})
    if isinstance(node, list):
        return [_wrap(v) for v in node]
    return node


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _package_header(path):
    with open(path) as fh:
        for line in fh:
            m = re.match(r"#\s*@package\s+(\S+)", line)
            if m:
                return m.group(1)
            if line.strip() and not line.startswith("#"):
                break
    return None


def load(path):
    """One config file with its `defaults` list composed.  Returns (node, package header or None)."""
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
        base = os.path.normpath(os.path.join(os.path.dirname(path), group, name)) if name else \
            os.path.normpath(os.path.join(os.path.dirname(path), group))
        sub, header = load(base + ".yaml")
        if not pkg:
            pkg = header if header is not None else (os.path.basename(group) if name else "")
        if pkg == "_global_":
            pkg = ""
        tgt = out
        for part in [p for p in pkg.split(".") if p]:
            tgt = tgt.setdefault(part, {})
        _merge(tgt, sub)
    if not self_done:
        _merge(out, node)
    return out, _package_header(path)


def lookup(root, expr):
    cur = root
    for part in re.findall(r"[^.\[\]]+", expr):
        cur = cur[int(part)] if isinstance(cur, list) else cur[part]
    return cur


def resolve(node, root=None):
    """Replace `${path}` strings by the value they point to (recursively)."""
    root = node if root is None else root
    if isinstance(node, dict):
        return {k: resolve(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [resolve(v, root) for v in node]
    if isinstance(node, str):
        m = re.fullmatch(r"\$\{([^}]+)\}", node.strip())
        if m:
            return resolve(lookup(root, m.group(1)), root)
    return node


def apply_overrides(cfg, overrides):
    for ov in overrides or []:
        key, _, val = ov.lstrip("+").partition("=")
        cur = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = yaml.safe_load(val)
    return cfg


def compose(config_dir, config_name, overrides=None, only_groups=None):
    """hydra.compose for the reference's experiment files.  only_groups: keep only defaults whose group path contains one
    of these substrings (e.g. ['models/'] when the dataset packages are not importable)."""
    path = os.path.join(config_dir, config_name + ("" if config_name.endswith(".yaml") else ".yaml"))
    if only_groups is not None:
        with open(path) as fh:
            raw = yaml.safe_load(fh)
        raw["defaults"] = [d for d in raw.get("defaults", []) if d == "_self_" or any(g in next(iter(d)) for g in only_groups)]
        tmp_dir = os.path.dirname(path)
        node = {k: v for k, v in raw.items() if k != "defaults"}
        out = {}
        for d in raw["defaults"]:
            if d == "_self_":
                _merge(out, node)
                continue
            (key, name), = d.items()
            group, _, pkg = key.partition("@")
            sub, header = load(os.path.normpath(os.path.join(tmp_dir, group, name)) + ".yaml")
            pkg = pkg or header or os.path.basename(group)
            tgt = out
            for part in [p for p in pkg.split(".") if p and p != "_global_"]:
                tgt = tgt.setdefault(part, {})
            _merge(tgt, sub)
        cfg = out
    else:
        cfg, _ = load(path)
    apply_overrides(cfg, overrides)
    return cfg


def instantiate(node, **kwargs):
    """hydra.utils.instantiate: `_target_` class/function called with the node's keys (+ kwargs); `_partial_: True`
    returns functools.partial; `_recursive_: False` passes nested nodes through un-instantiated."""
    if isinstance(node, dict) and "_target_" in node:
        kw = {k: v for k, v in node.items() if k not in ("_target_", "_recursive_", "_partial_")}
        if node.get("_recursive_", True):
            kw = {k: instantiate(v) for k, v in kw.items()}
        else:
            kw = {k: _wrap(v) for k, v in kw.items()}
        kw.update(kwargs)
        mod, _, name = node["_target_"].rpartition(".")
        fn = getattr(importlib.import_module(mod), name)
        return functools.partial(fn, **kw) if node.get("_partial_", False) else fn(**kw)
    if isinstance(node, dict):
        return _wrap({k: instantiate(v) for k, v in node.items()})
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    return node


def main(config_dir, config_name, argv=None):
    """What `@hydra.main` hands to the decorated function: the composed, override-applied, resolved config."""
    cfg = compose(config_dir, config_name, overrides=[a for a in (argv or []) if "=" in a])
    return _wrap(resolve(cfg))
