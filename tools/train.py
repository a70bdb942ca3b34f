#!/usr/bin/env python
"""Training entry point with the reference's flow (tools/train.py:16-79): compose the experiment config (hydra-style
defaults / interpolation / dotted overrides through pillarnext_b200.hydra_lite), NCCL process group from the torchrun
environment, datasets + loaders, `instantiate(cfg.model)` (the det3d `_target_` classes of this repo), SyncBatchNorm when
`model.sync_batchnorm`, AdamW + OneCycleLR, Trainer.fit().

  python tools/train.py --config-name synth_det_pp18_aspp trainer.max_epochs=2 dataloader.train.batch_size=4
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train.py --config-name synth_det_pp18_aspp
"""
import argparse
import logging
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pillarnext_b200 import hydra_lite  # noqa: E402
from pillarnext_b200.loader import build_dataloader  # noqa: E402
from pillarnext_b200.trainer import Trainer  # noqa: E402


def parse(argv):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config-path", default=os.path.join(ROOT, "configs", "experiments"))
    ap.add_argument("--config-name", required=True)
    ap.add_argument("--work-dir", default=os.path.join(ROOT, "gpurun_out", "work"))
    args, overrides = ap.parse_known_args(argv)
    return args, overrides


def setup(args, overrides):
    cfg = hydra_lite.main(args.config_path, args.config_name, overrides)
    distributed = int(os.environ.get("WORLD_SIZE", "1")) > 1
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if distributed:
        dist.init_process_group(backend="nccl", init_method="env://", device_id=torch.device("cuda", local_rank))
    logging.basicConfig(level=logging.INFO if int(os.environ.get("RANK", "0")) == 0 else logging.ERROR,
                        format="%(asctime)s %(message)s")
    return cfg, distributed, logging.getLogger("pillarnext_b200")


def build_model(cfg, distributed):
    model = hydra_lite.instantiate(cfg.model)           # sync_batchnorm is consumed by SingleStageDetector (enable_sync_batchnorm)
    if distributed and cfg.model.get("sync_batchnorm", False):
        from pillarnext_b200.modules import enable_sync_batchnorm
        enable_sync_batchnorm(model)
    return model.cuda()


def main(argv=None):
    args, overrides = parse(sys.argv[1:] if argv is None else argv)
    cfg, distributed, logger = setup(args, overrides)
    train_dataset = hydra_lite.instantiate(cfg.data.train_dataset)
    train_loader = build_dataloader(train_dataset, **cfg.dataloader.train)
    val_loader = None
    if "val_dataset" in cfg.data:
        val_loader = build_dataloader(hydra_lite.instantiate(cfg.data.val_dataset), **cfg.dataloader.val)
    model = build_model(cfg, distributed)
    optimizer = hydra_lite.instantiate(cfg.optimizer, params=model.parameters())
    sched = dict(cfg.scheduler)
    sched["_recursive_"] = False
    lr_scheduler = hydra_lite.instantiate(sched, optimizer=optimizer, steps_per_epoch=len(train_loader))
    trainer = Trainer(model, train_loader, val_loader, optimizer, lr_scheduler, logger=logger, work_dir=args.work_dir, **cfg.trainer)
    if "resume_from" in cfg:
        trainer.resume(cfg.resume_from)
    if "load_from" in cfg:
        trainer.load_checkpoint(cfg.load_from)
    trainer.fit()
    if distributed:
        dist.destroy_process_group()
    return trainer


if __name__ == "__main__":
    main()
