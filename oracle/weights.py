"""Seeded parameter sets keyed like the reference state dict (TEST INFRASTRUCTURE).

Weights come from explicit seeded draws (not from module init order) so that the reference modules,
the oracle and the CUDA product can all be driven by the same values on any machine."""
import torch


def randomize_state_dict(template, seed):
    """template: name -> tensor (shapes from a model's state_dict()). Returns fp32 CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(template.keys()):
        shape = tuple(template[k].shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros(shape, dtype=torch.long)
        elif k.endswith("running_mean"):
            out[k] = torch.randn(shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            out[k] = torch.rand(shape, generator=g) + 0.5
        elif ".norm" in k or k.endswith(".1.weight") or k.endswith(".1.bias") or "norm2" in k:
            # BatchNorm affine parameters (1-D)
            out[k] = (torch.rand(shape, generator=g) + 0.5) if k.endswith("weight") else torch.randn(shape, generator=g) * 0.1
        elif k.endswith("bias"):
            out[k] = torch.randn(shape, generator=g) * 0.1
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if k.endswith("deblock.conv.conv.weight"):      # ConvTranspose2d [Cin, Cout, 2, 2]
                fan_in = shape[0]
            out[k] = torch.randn(shape, generator=g) * (1.5 / max(fan_in, 1) ** 0.5)
    return out
