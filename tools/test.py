#!/usr/bin/env python
"""Evaluation entry point with the reference's flow (tools/test.py:16-66): compose the config, build the validation
loader and the model, `load_checkpoint(cfg.load_from, strict=True)`, Trainer.val_epoch() -> dataset.evaluation().

  python tools/test.py --config-name synth_det_pp18_aspp +load_from=gpurun_out/work/epoch_2.pth
"""
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from pillarnext_b200 import hydra_lite  # noqa: E402
from pillarnext_b200.loader import build_dataloader  # noqa: E402
from pillarnext_b200.trainer import Trainer  # noqa: E402
import train as _train  # noqa: E402


def main(argv=None):
    args, overrides = _train.parse(sys.argv[1:] if argv is None else argv)
    cfg, distributed, logger = _train.setup(args, overrides)
    val_loader = build_dataloader(hydra_lite.instantiate(cfg.data.val_dataset), **cfg.dataloader.val)
    model = _train.build_model(cfg, distributed)
    trainer = Trainer(model, val_dataloader=val_loader, logger=logger, work_dir=args.work_dir, **cfg.trainer)
    trainer.load_checkpoint(cfg.load_from, strict=True)
    result, _ = trainer.val_epoch()
    if distributed:
        dist.destroy_process_group()
    return result


if __name__ == "__main__":
    main()
