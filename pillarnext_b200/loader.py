"""build_dataloader of the reference (det3d/datasets/loader/build_loader.py:7-26): DistributedSampler when torch.distributed
is initialised, the dataset's collate, pinned memory for the non-blocking host->device copies of the trainer."""
import torch
import torch.distributed as dist

from .scenes import collate


def build_dataloader(dataset, batch_size, num_workers=0, shuffle=False, **_):
    sampler = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        sampler = torch.utils.data.distributed.DistributedSampler(dataset, shuffle=shuffle)
        shuffle = False
    return torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, sampler=sampler, num_workers=num_workers,
                                       collate_fn=collate, pin_memory=torch.cuda.is_available(), drop_last=False)
