"""Synthetic head outputs and the post-processing config shared by the F1 tests (TEST INFRASTRUCTURE)."""
import torch

def test_cfg():
    return dict(post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], score_threshold=0.1,
                out_size_factor=[4, 4], voxel_size=[0.075, 0.075], pc_range=[-50.4, -50.4],
                nms=dict(nms_iou_threshold=[[0.2], [0.2, 0.2]], nms_pre_max_size=1000, nms_post_max_size=83))


def fake_preds(B, H, W, classes, seed):
    """Head outputs with a sparse set of confident, separated peaks (so the NMS input is small and not degenerate)."""
    g = torch.Generator().manual_seed(seed)
    pd = dict(reg=torch.rand(B, 2, H, W, generator=g), height=torch.randn(B, 1, H, W, generator=g),
              dim=torch.randn(B, 3, H, W, generator=g) * 0.3 + 0.5, rot=torch.randn(B, 2, H, W, generator=g),
              vel=torch.randn(B, 2, H, W, generator=g), hm=torch.full((B, classes, H, W), -6.0))
    n = 60
    for b in range(B):
        ys = torch.randint(0, H, (n,), generator=g)
        xs = torch.randint(0, W, (n,), generator=g)
        cs = torch.randint(0, classes, (n,), generator=g)
        pd["hm"][b, cs, ys, xs] = torch.randn(n, generator=g) * 1.5 + 0.5
    return pd



OFFS = dict(reg=0, height=2, dim=3, rot=6, vel=8, hm=10)


def to_rows(pd, npad=16):
    """Head-name -> [B, c, H, W] dict to the channels-last matrix [B*H*W, npad] the fused head writes."""
    B, _, H, W = pd["hm"].shape
    out = torch.zeros(B * H * W, npad)
    for k, o in OFFS.items():
        v = pd[k].permute(0, 2, 3, 1).reshape(B * H * W, -1)
        out[:, o:o + v.shape[1]] = v
    return out.contiguous()


