"""Thin host wrappers: torch device buffers + current stream -> libpnx C-ABI calls.

torch is used here for device memory, streams and dtype plumbing only; all compute is in libpnx.
"""
import numpy as np
import torch

from ._lib import check, lib, ptr, sm_count, stream

# ---- instrumentation used by bench.py: kernel-launch counter and optional per-call CUDA-event timing
LAUNCHES = 0            # number of libpnx kernels launched so far (each wrapper adds its own kernel count)
PROFILE = None          # when a list: (kind, flops, bytes, start_event, end_event) per tensor-core GEMM call


def _count(n):
    global LAUNCHES
    LAUNCHES += n


class _Timed:
    def __init__(self, kind, flops, nbytes, tag=""):
        self.rec = None
        if PROFILE is not None:
            self.rec = [kind, flops, nbytes, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), tag]

    def __enter__(self):
        if self.rec is not None:
            self.rec[3].record()

    def __exit__(self, *a):
        if self.rec is not None:
            self.rec[4].record()
            PROFILE.append(tuple(self.rec))


# --------------------------------------------------------------------------------- per-step zero arena
class ZeroArena:
    """The small zero-initialised accumulators of one training step (BatchNorm statistics, reduction buffers, the fp32
    weight-gradient accumulators: ~130 tensors, ~45 MB) carved out of ONE buffer that is cleared by a single memset at
    the start of the step, instead of one fill kernel each.  Inactive (plain torch.zeros) until begin_step() is called;
    the views are only valid until the next begin_step()."""

    def __init__(self):
        self.buf, self.off, self.need, self.active = None, 0, 0, False

    def begin_step(self, device):
        cap = self.buf.numel() if self.buf is not None else 0
        want = max(self.need, self.off)
        if self.buf is None or self.buf.device != device or want > cap:
            self.buf = torch.empty(max(int(want * 1.25), 64 << 20), dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.off, self.need, self.active = 0, 0, True

    def zeros(self, shape, dtype, device):
        n = 1
        for s in (shape if isinstance(shape, (tuple, list)) else (shape,)):
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        if not self.active or self.buf is None or self.buf.device != device:
            return torch.zeros(shape, dtype=dtype, device=device)
        start = (self.off + 255) // 256 * 256
        self.need = max(self.need, start + nbytes)
        if start + nbytes > self.buf.numel():       # first step(s): grow at the next begin_step, plain allocation now
            self.off = start + nbytes
            return torch.zeros(shape, dtype=dtype, device=device)
        self.off = start + nbytes
        return self.buf[start:start + nbytes].view(dtype).view(shape)


ARENA = ZeroArena()


def zeros(shape, dtype, device):
    return ARENA.zeros(shape, dtype, device)


# --------------------------------------------------------------------------------- geometry
def grid_size_xy(voxel_size, pc_range):
    """Same float64 arithmetic as the reference (pillar_encoder.py:87-89)."""
    vs = np.array(voxel_size, dtype=np.float64)
    pr = np.array(pc_range, dtype=np.float64)
    g = (pr[3:] - pr[:3]) / vs
    return np.round(g).astype(np.int64)


def _f32(x):
    return float(np.float32(x))


class Voxels:
    """Result of the voxelizer: fixed-capacity device buffers + device counts {P, Nv}."""

    def __init__(self):
        self.P = None
        self.Nv = None

    def sync_counts(self):
        if self.P is None:
            if getattr(self, "status", None) is not None:
                c = torch.cat([self.counts, self.status]).cpu()
                self.check_order(c[2])
            else:
                c = self.counts.cpu()
            self.P, self.Nv = int(c[0]), int(c[1])
        return self.P, self.Nv

    def check_order(self, status_value=None):
        """Raise if the frame-tiled voxelizer found the points not grouped by frame (status_value: the already fetched
        host copy of self.status; fetched here, with a synchronisation, when omitted)."""
        if getattr(self, "status", None) is None:
            return
        bad = int(self.status.item()) if status_value is None else int(status_value)
        if bad:
            raise RuntimeError("pnx_voxelize_frames: %d point(s) lie outside their frame's range -- the points are not grouped "
                               "by ascending batch index; use voxelize(..., frame_sorted=False)" % bad)


# The frame-tiled kernels (pnx_voxelize_frames) are bit-exact but, measured on the B200, NOT faster than the global-bitmap
# ones at any batch size (256 nuScenes frames: 416 us vs 290 us; ncu: instruction-issue bound, 139 M + 96 M warp
# instructions -- every slice re-processes all points of its frame and the bitmap sweeps cost as much as the points), so
# they only run on request (frame_sorted="force"); None = never chosen automatically.
FRAME_TILED_MIN_CTAS = None


def voxelize(points, batch, voxel_size, pc_range, buckets=True, frame_sorted=False):
    """V1-V2 index generation (pnx_voxelize) and, with buckets=True, the CSR grouping the PFN needs
    (pnx_bucketize). points [N,6] fp32 cuda (b,x,y,z,i,t).
    frame_sorted=True: the points are grouped by frame in ascending batch index (what collate produces) -> the
    frame-tiled kernels (pnx_voxelize_frames: bitmap slices in shared memory) MAY be used: "force" = whenever the geometry
    allows, True = when FRAME_TILED_MIN_CTAS says so (never, today: see there).  The order is verified on the device:
    `v.status` (int32 [1]) is non-zero when it did not hold and the outputs must be discarded -- check it at the next
    host synchronisation (Voxels.check_order(), modules.build_pyramid does) and re-run with frame_sorted=False."""
    assert points.is_cuda and points.dtype == torch.float32 and points.dim() == 2 and points.shape[1] == 6, \
        "points must be a CUDA float32 [N, 6] tensor (batch_idx, x, y, z, intensity, time)"
    points = points.contiguous()
    if points.data_ptr() % 16:
        points = points.clone()
    n = points.shape[0]
    g = grid_size_xy(voxel_size, pc_range)
    gx, gy = int(g[0]), int(g[1])
    dev = points.device
    L = lib()
    words = L.pnx_voxelize_bitmap_words(batch, gx, gy)
    cap_p = min(n, batch * gx * gy)
    v = Voxels()
    v.points, v.n, v.batch, v.gx, v.gy, v.cap_p = points, n, batch, gx, gy, cap_p
    v.min_x, v.min_y = _f32(pc_range[0]), _f32(pc_range[1])
    v.vs_x, v.vs_y = _f32(voxel_size[0]), _f32(voxel_size[1])
    i32 = dict(dtype=torch.int32, device=dev)
    v.bitmap = torch.empty(words, **i32)
    v.blockcnt = torch.empty(L.pnx_blockcnt_size(words // 32), **i32)
    v.inblk = torch.empty(words, dtype=torch.int16, device=dev)
    v.blockpref = torch.empty(words // 32 + 1, **i32)
    cell = torch.empty(max(n, 1), **i32)
    v.pillar_of_point = torch.empty(max(n, 1), **i32)
    v.coords = torch.empty(max(cap_p, 1), 3, **i32)
    bucket_cnt = torch.empty(2 * (cap_p + 1), **i32) if buckets else None
    v.counts = torch.zeros(2, **i32)
    v.status = None
    n_cta = L.pnx_voxelize_frames_supported(batch, gx, gy) if (frame_sorted and n > 0) else 0
    if n_cta > 0 and (frame_sorted == "force" or (FRAME_TILED_MIN_CTAS is not None and n_cta >= FRAME_TILED_MIN_CTAS)):
        scratch = torch.empty(L.pnx_voxelize_frames_scratch(batch), **i32)
        v.status = scratch[0:1]
        _count(4)
        check(L.pnx_voxelize_frames(ptr(points), n, batch, v.min_x, v.min_y, v.vs_x, v.vs_y, gx, gy, ptr(v.bitmap), ptr(v.inblk),
                                    ptr(v.blockcnt), ptr(v.blockpref), ptr(cell), ptr(v.pillar_of_point), ptr(v.coords), cap_p,
                                    ptr(bucket_cnt) if buckets else None, ptr(v.counts), ptr(scratch), stream()))
    else:
        _count(5 if n > 0 else 3)
        check(L.pnx_voxelize(ptr(points), n, batch, v.min_x, v.min_y, v.vs_x, v.vs_y, gx, gy, ptr(v.bitmap), ptr(v.inblk),
                             ptr(v.blockcnt), ptr(v.blockpref), ptr(cell), ptr(v.pillar_of_point), ptr(v.coords), cap_p,
                             ptr(bucket_cnt) if buckets else None, ptr(v.counts), stream()))
    if buckets:
        scratch = torch.empty(cap_p // 2048 + 4, **i32)
        v.bucket_off = torch.empty(cap_p + 1, **i32)
        bucket_tmp = torch.empty(max(n, 1), **i32)
        v.bucket_pts = torch.empty(max(n, 1), **i32)
        _count(5 if n > 0 else 3)
        check(L.pnx_bucketize(ptr(v.pillar_of_point), n, cap_p, ptr(bucket_cnt), ptr(scratch), ptr(v.bucket_off),
                              ptr(bucket_tmp), ptr(v.bucket_pts), ptr(v.counts), stream()))
    return v


def pillar_mean(v):
    """Per-pillar mean of (x, y, z) (pnx_pfn_mean; pillar_encoder.py:113-114 scatter_mean): [cap_p, 3] fp32, rows >= P
    undefined.  Needs the CSR buckets (voxelize(..., buckets=True))."""
    mean = torch.empty(max(v.cap_p, 1), 3, dtype=torch.float32, device=v.points.device)
    _count(1)
    check(lib().pnx_pfn_mean(ptr(v.points), ptr(v.bucket_off), ptr(v.bucket_pts), ptr(v.counts), v.cap_p, ptr(mean), stream()))
    return mean


# --------------------------------------------------------------------------------- BatchNorm stats
def bn_finalize(stats, channels, count_ptr, count_mult, gamma, beta, eps, momentum, running_mean, running_var,
                want_saved=True):
    dev = stats.device
    scale = torch.empty(channels, dtype=torch.float32, device=dev)
    shift = torch.empty_like(scale)
    mean = torch.empty_like(scale) if want_saved else None
    invstd = torch.empty_like(scale) if want_saved else None
    _count(1)
    check(lib().pnx_bn_finalize(ptr(stats), channels, ptr(count_ptr) if count_ptr is not None else None,
                                int(count_mult), ptr(gamma), ptr(beta), float(eps), float(momentum),
                                ptr(running_mean) if running_mean is not None else None,
                                ptr(running_var) if running_var is not None else None,
                                ptr(scale), ptr(shift), ptr(mean) if want_saved else None,
                                ptr(invstd) if want_saved else None, stream()))
    return scale, shift, mean, invstd


def bn_eval_affine(gamma, beta, rm, rv, eps):
    c = gamma.shape[0]
    scale = torch.empty(c, dtype=torch.float32, device=gamma.device)
    shift = torch.empty_like(scale)
    _count(1)
    check(lib().pnx_bn_eval_affine(c, ptr(gamma), ptr(beta), ptr(rm), ptr(rv), float(eps), ptr(scale), ptr(shift),
                                   stream()))
    return scale, shift


# --------------------------------------------------------------------------------- PFN forward
def _sync_stats(stats, count):
    """SyncBatchNorm (reference tools/train.py:55-56): batch statistics over the points of ALL ranks.
    stats fp64 [2C] local sums, count = device int tensor [1] -> (global stats, global count as device int32 [1])."""
    import torch.distributed as dist
    pack = torch.cat([stats, count.to(torch.float64)])
    dist.all_reduce(pack)
    return pack[:-1], pack[-1:].round().to(torch.int32)


def pfn_forward(v, w0, bn0, w1, bn1, training, eps=1e-3, momentum=0.01, want_f32=True, sync=False):
    """P1-P3. bn0/bn1 = (gamma, beta, running_mean, running_var). Returns dict with feat (fp32 [capP,64]),
    feat_bf16, and everything the backward needs.  sync: SyncBatchNorm statistics (all-reduced over the ranks)."""
    L = lib()
    dev = v.points.device
    f32 = dict(dtype=torch.float32, device=dev)
    n, cap_p = max(v.n, 1), max(v.cap_p, 1)
    out = {}
    mean = torch.empty(cap_p, 3, **f32)
    y0 = torch.empty(n, 32, **f32)
    y1 = torch.empty(n, 64, **f32)
    x0max = torch.empty(cap_p, 32, **f32)
    stats = torch.zeros(2 * 32 + 2 * 64, dtype=torch.float64, device=dev)
    s0, s1 = stats[:64], stats[64:]
    nv_ptr = v.counts[1:2]
    tr = 1 if training else 0
    _count(5)
    check(L.pnx_pfn_mean(ptr(v.points), ptr(v.bucket_off), ptr(v.bucket_pts), ptr(v.counts), v.cap_p, ptr(mean),
                         stream()))
    check(L.pnx_pfn_lin0(ptr(v.points), ptr(v.bucket_pts), ptr(v.pillar_of_point), ptr(v.coords), ptr(mean),
                         ptr(v.counts), v.n, v.min_x, v.min_y, v.vs_x, v.vs_y, ptr(w0), ptr(y0), ptr(s0), tr,
                         stream()))
    sync = bool(sync and training)
    s0f, s1f, nv_bn = s0, s1, nv_ptr
    if sync:
        s0f, nv_bn = _sync_stats(s0, nv_ptr)
    if training:
        sc0, sh0, m0, i0 = bn_finalize(s0f, 32, nv_bn, 1, bn0[0], bn0[1], eps, momentum, bn0[2], bn0[3])
    else:
        sc0, sh0 = bn_eval_affine(bn0[0], bn0[1], bn0[2], bn0[3], eps)
        m0 = i0 = None
    check(L.pnx_pfn_max0(ptr(y0), ptr(v.bucket_off), ptr(v.counts), v.cap_p, ptr(sc0), ptr(sh0), ptr(x0max),
                         stream()))
    check(L.pnx_pfn_lin1(ptr(y0), ptr(x0max), ptr(v.bucket_pts), ptr(v.pillar_of_point), ptr(v.counts), v.n,
                         ptr(sc0), ptr(sh0), ptr(w1), ptr(y1), ptr(s1), tr, stream()))
    if sync:
        s1f, _ = _sync_stats(s1, nv_ptr)
    if training:
        sc1, sh1, m1, i1 = bn_finalize(s1f, 64, nv_bn, 1, bn1[0], bn1[1], eps, momentum, bn1[2], bn1[3])
    else:
        sc1, sh1 = bn_eval_affine(bn1[0], bn1[1], bn1[2], bn1[3], eps)
        m1 = i1 = None
    feat = torch.empty(cap_p, 64, **f32) if want_f32 else None
    feat_bf16 = torch.empty(cap_p, 64, dtype=torch.bfloat16, device=dev)
    check(L.pnx_pfn_max1(ptr(y1), ptr(v.bucket_off), ptr(v.counts), v.cap_p, ptr(sc1), ptr(sh1),
                         ptr(feat) if want_f32 else None, ptr(feat_bf16), stream()))
    out.update(feat=feat, feat_bf16=feat_bf16, mean=mean, y0=y0, y1=y1, x0max=x0max, sc0=sc0, sh0=sh0, sc1=sc1,
               sh1=sh1, mean0=m0, invstd0=i0, mean1=m1, invstd1=i1, bn_count=nv_bn if sync else None)
    return out


def pfn_backward(v, fwd, dfeat, w1, gamma0, gamma1):
    """Backward of pfn_forward (training mode). dfeat fp32 [>=P, 64] contiguous.
    Returns dW0 [32,10], dW1 [64,64], dgamma0, dbeta0, dgamma1, dbeta1 (fp32)."""
    dev = dfeat.device
    n, cap_p = max(v.n, 1), max(v.cap_p, 1)
    assert dfeat.dtype == torch.float32 and dfeat.is_contiguous() and dfeat.shape[0] >= min(v.P or 0, cap_p)
    if dfeat.shape[0] < cap_p:
        full = torch.zeros(cap_p, 64, dtype=torch.float32, device=dev)
        full[:dfeat.shape[0]] = dfeat
        dfeat = full
    argq1 = torch.empty(cap_p, 64, dtype=torch.int32, device=dev)
    d_x0 = torch.empty(n, 32, dtype=torch.float32, device=dev)
    dxm = torch.empty(n, 32, dtype=torch.float32, device=dev)
    red = torch.zeros(64 + 128, dtype=torch.float64, device=dev)
    dW0 = torch.zeros(32, 10, dtype=torch.float64, device=dev)      # fp64 accumulators (cross-CTA atomics, order-insensitive)
    dW1 = torch.zeros(64, 64, dtype=torch.float64, device=dev)
    _count(4)
    bn_count = fwd.get("bn_count")

    def run(phases, red_buf):
        check(lib().pnx_pfn_backward(ptr(v.points), ptr(v.bucket_off), ptr(v.bucket_pts), ptr(v.pillar_of_point),
                                     ptr(v.coords), ptr(v.counts), v.n, v.cap_p, v.min_x, v.min_y, v.vs_x, v.vs_y,
                                     ptr(fwd["mean"]), ptr(fwd["y0"]), ptr(fwd["y1"]), ptr(fwd["x0max"]), ptr(fwd["feat"]),
                                     ptr(dfeat), ptr(w1), ptr(fwd["sc0"]), ptr(fwd["sh0"]), ptr(fwd["mean0"]),
                                     ptr(fwd["invstd0"]), ptr(gamma0), ptr(fwd["sc1"]), ptr(fwd["sh1"]), ptr(fwd["mean1"]),
                                     ptr(fwd["invstd1"]), ptr(gamma1), ptr(argq1), ptr(d_x0), ptr(dxm), ptr(red_buf), ptr(dW0),
                                     ptr(dW1), phases, ptr(bn_count) if bn_count is not None else None, stream()))

    if bn_count is None:
        run(0, red)
    else:
        # SyncBatchNorm backward: the two per-channel sums of each BN are all-reduced between the phase that produces
        # them and the phase that applies them; dgamma / dbeta stay local (DDP averages parameter gradients)
        import torch.distributed as dist
        glob = torch.zeros_like(red)
        run(1, red)
        glob[64:] = red[64:]
        dist.all_reduce(glob[64:])
        run(2, glob)
        run(4, red)
        glob[:64] = red[:64]
        dist.all_reduce(glob[:64])
        run(8, glob)
    r = red.float()
    return dW0.float(), dW1.float(), r[32:64], r[0:32], r[128:192], r[64:128]


# --------------------------------------------------------------------------------- sites / rulebook
class Level:
    """Active-site set of one backbone level: bitmap (b,u,v), inblk (u16 per word), blockpref, coords [n,3], count."""
    pass


def level_from_bitmap(bm, blockpref, count, batch, U, V, inblk=None):
    """Wrap an existing bitmap (e.g. the voxelizer's level 0); inblk is computed here when not given."""
    lv = Level()
    lv.bm, lv.blockpref, lv.count, lv.batch, lv.U, lv.V = bm, blockpref, count, batch, U, V
    if inblk is None:
        inblk = torch.empty(bm.numel(), dtype=torch.int16, device=bm.device)
        _count(1)
        check(lib().pnx_sites_inblock(ptr(bm), bm.numel(), ptr(inblk), stream()))
    lv.inblk = inblk
    lv.n = None
    lv.coords = None
    return lv


def level_from_mask_words(bm, batch, U, V):
    """Level from a raw bitmap (words padded to whole blocks): block counts, scan and in-block prefix."""
    assert bm.numel() % 32 == 0
    blocks = bm.numel() // 32
    inblk = torch.empty(bm.numel(), dtype=torch.int16, device=bm.device)
    _count(1)
    check(lib().pnx_sites_inblock(ptr(bm), bm.numel(), ptr(inblk), stream()))
    cnt = torch.zeros(lib().pnx_blockcnt_size(blocks), dtype=torch.int32, device=bm.device)
    # block counts = in-block prefix of the last word + its popcount (host-side torch plumbing; test helper only)
    w = bm.view(-1, 32)
    pc = torch.zeros_like(w)
    x = w.clone().to(torch.int64) & 0xFFFFFFFF
    for _ in range(32):
        pc += (x & 1).to(torch.int32)
        x >>= 1
    cnt[:blocks] = pc.sum(1)
    so = (blocks + 3) // 4 * 4
    nsb = (blocks + 1023) // 1024
    padded = torch.zeros(nsb * 1024, dtype=torch.int32, device=bm.device)
    padded[:blocks] = cnt[:blocks]
    cnt[so:so + nsb] = padded.view(nsb, 1024).sum(1)
    blockpref = torch.empty(blocks + 1, dtype=torch.int32, device=bm.device)
    count = torch.empty(1, dtype=torch.int32, device=bm.device)
    _count(1)
    check(lib().pnx_scan_blocks(ptr(cnt), blocks, ptr(blockpref), ptr(count), stream()))
    return level_from_bitmap(bm, blockpref, count, batch, U, V, inblk=inblk)


def level_dilate(src, stride):
    L = lib()
    uo, vo = L.pnx_sites_out_dim(src.U, stride), L.pnx_sites_out_dim(src.V, stride)
    words = (src.batch * uo * ((vo + 31) // 32) + 31) // 32 * 32
    dev = src.bm.device
    bm = torch.empty(words, dtype=torch.int32, device=dev)
    inblk = torch.empty(words, dtype=torch.int16, device=dev)
    cnt = torch.empty(L.pnx_blockcnt_size(words // 32), dtype=torch.int32, device=dev)
    blockpref = torch.empty(words // 32 + 1, dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    _count(2)
    check(L.pnx_sites_dilate(ptr(src.bm), src.batch, src.U, src.V, stride, ptr(bm), ptr(inblk), ptr(cnt), stream()))
    check(L.pnx_scan_blocks(ptr(cnt), words // 32, ptr(blockpref), ptr(count), stream()))
    return level_from_bitmap(bm, blockpref, count, src.batch, uo, vo, inblk=inblk)


def level_coords(lv, n):
    lv.n = n
    lv.coords = torch.empty(max(n, 1), 3, dtype=torch.int32, device=lv.bm.device)
    _count(1)
    check(lib().pnx_sites_coords(ptr(lv.bm), ptr(lv.blockpref), ptr(lv.inblk), lv.batch, lv.U, lv.V, ptr(lv.coords), n,
                                 stream()))
    return lv.coords


def nbr_table(dst, src, stride, transposed):
    nbr = torch.empty(max(dst.n, 1), 9, dtype=torch.int32, device=dst.bm.device)
    _count(1)
    check(lib().pnx_nbr_table(ptr(dst.coords), ptr(dst.count), dst.n, ptr(src.bm), ptr(src.blockpref), ptr(src.inblk),
                              src.batch, src.U, src.V, stride, 1 if transposed else 0, ptr(nbr), stream()))
    return nbr


def scatter_dense(feat, lv, channels, canvas=None, gather=False):
    """feat [n,C] bf16 <-> canvas [B, V(y), U(x), C] bf16."""
    if canvas is None:
        canvas = torch.zeros(lv.batch, lv.V, lv.U, channels, dtype=torch.bfloat16, device=feat.device)
    _count(1)
    check(lib().pnx_scatter_dense(ptr(feat), ptr(lv.coords), ptr(lv.count), lv.n, channels, lv.batch, lv.U, lv.V,
                                  ptr(canvas), 1 if gather else 0, stream()))
    return canvas


# --------------------------------------------------------------------------------- implicit GEMM
def _bnr_args(bnr, cout):
    """Trailing C-ABI arguments of the fused BatchNorm-backward reduce (None: off).  bnr: functional.BNInfo."""
    if bnr is None:
        return (None, 0, None, None, None, None, None, 0)
    assert bnr.raw.dtype == torch.bfloat16 and bnr.C == cout and bnr.red.numel() == 2 * cout
    return (ptr(bnr.raw), bnr.raw.stride(0), ptr(bnr.scale), ptr(bnr.shift), ptr(bnr.mean), ptr(bnr.invstd), ptr(bnr.red), cout)


def bnr_eligible(cout, out_fp32=False, shuffle=False):
    """The fused reduce lives in the bf16 staged (coalesced) store of the GEMM epilogues: 64-column granularity."""
    return (not out_fp32) and (not shuffle) and cout % 64 == 0


def pick_block_n(cout):
    for bn in (256, 192, 128, 64, 32, 16):
        if cout % bn == 0:
            return bn
    raise ValueError("Cout %d is not a multiple of 16" % cout)


def igemm(A, M, w_packed, taps, cin, cout, out, *, lda=None, ldc=None, nbr=None, dense=None, bias=None, stats=None,
          stats_mod=None, shuffle=False, relu=False, block_n=None, addend=None, segs=None, a_lo_off=0, bnr=None):
    """out[m, :cout] = sum_t A[nbr(m,t)] @ W[t]^T.  w_packed [taps, cout, cin] bf16.
    dense = (Hout, Wout, Hin, Win, kw, mul, dil, pad) or None.
    segs: fp32-grade split mode -- list of (A piece, W piece) K segments in execution order; A rows hold piece p at
    column p*a_lo_off + c, w_packed is [taps, cout, pieces*cin] (see include/pnx.h "split rows")."""
    assert A.dtype == torch.bfloat16 and w_packed.dtype == torch.bfloat16 and w_packed.is_contiguous()
    nseg, seg_code = 1, 0
    if segs:
        nseg = len(segs)
        for i, (ap, wp) in enumerate(segs):
            seg_code |= ((ap << 2) | wp) << (4 * i)
    w_pieces = (max(wp for _, wp in segs) + 1) if segs else 1
    assert tuple(w_packed.shape) == (taps, cout, cin * w_pieces), (tuple(w_packed.shape), (taps, cout, cin), segs)
    lda = A.stride(0) if lda is None else lda
    ldc = out.stride(-2) if ldc is None else ldc
    bn = block_n or pick_block_n(cout)
    d = dense or (0, 0, 0, 0, 1, 1, 1, 0)
    out_fp32 = 1 if out.dtype == torch.float32 else 0
    assert out.dtype in (torch.float32, torch.bfloat16)
    sC = stats.numel() // 2 if stats is not None else 0
    _count(1)
    add_f32 = 1 if (addend is not None and addend.dtype == torch.float32) else 0
    with _Timed("igemm", 2.0 * M * taps * cin * cout * nseg, 2.0 * M * (cin * min(taps, 2) + cout) + 2.0 * taps * cin * cout,
                "M%d_T%d_K%d_N%d_bn%d%s" % (M, taps, cin, cout, bn, "_tbl" if nbr is not None else ("_dense" if dense else ""))):
      check(lib().pnx_igemm(ptr(A), lda, A.shape[0], M, taps, cin, ptr(w_packed), cout, bn, ptr(nbr) if nbr is not None else None,
                          1 if dense else 0, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], ptr(out), ldc, out_fp32,
                          ptr(bias) if bias is not None else None, ptr(stats) if stats is not None else None, sC,
                          stats_mod or (sC if sC else 1), 1 if shuffle else 0, 1 if relu else 0,
                          ptr(addend) if addend is not None else None, addend.stride(0) if addend is not None else 0,
                          int(nseg), int(a_lo_off), int(seg_code), add_f32, *_bnr_args(bnr, cout), sm_count(), stream()))
    return out


WIN_BASE_OFF = 0   # measured on B200: the 128B swizzle is applied on absolute smem address bits, so row-shifted
                   # descriptors need NO matrix base offset (tests/test_igemm_win_gpu.py checks both conventions)


def win_eligible(dense, n_cols, k_cols, out_fp32=False, shuffle=False):
    """Dense 3x3 stride-1 dilation-1 'same' conv whose image rows tile into 128-pixel chunks with <= 15 % waste."""
    if dense is None or out_fp32 or shuffle:
        return False
    Ho, Wo, Hi, Wi, kw, mul, dil, pad = dense
    if not (kw == 3 and mul == 1 and dil == 1 and pad == 1 and Ho == Hi and Wo == Wi):
        return False
    if n_cols % 64 or k_cols % 64:
        return False
    xc = (Wo + 127) // 128
    return xc * 128 <= 1.15 * Wo


def conv3x3_win(A, B, H, W, w_packed, cin, cout, out, *, bias=None, stats=None, relu=False, block_n=None, base_off=None, bnr=None):
    """Dense 3x3 'same' conv with TMA-folded im2col (pnx_conv3x3_win). A bf16 rows [B*H*W, >=cin]; out bf16."""
    assert A.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and tuple(w_packed.shape) == (9, cout, cin)
    if block_n:
        bn = block_n
    elif cin == 64 and cout % 128 == 0:
        bn = 128            # weights-stationary variant: the nine [128 x 64] tiles stay in shared memory
    else:
        bn = 192 if cout % 192 == 0 else (128 if cout % 128 == 0 else 64)
    M = B * H * W
    _count(1)
    with _Timed("igemm_win", 2.0 * M * 9 * cin * cout, 2.0 * M * (cin * 3 + cout), "M%d_K%d_N%d_bn%d" % (M, cin, cout, bn)):
      check(lib().pnx_conv3x3_win(ptr(A), A.stride(0), B, H, W, cin, ptr(w_packed), cout, bn, ptr(out), out.stride(0),
                                ptr(bias) if bias is not None else None, ptr(stats) if stats is not None else None,
                                stats.numel() // 2 if stats is not None else 0, 1 if relu else 0,
                                WIN_BASE_OFF if base_off is None else base_off, *_bnr_args(bnr, cout), sm_count(), stream()))
    return out


DETERMINISTIC = False     # set through functional.set_deterministic


def set_deterministic(on):
    """Library side of functional.set_deterministic: ordered split-K reduction in pnx_wgrad (per-split slabs added in split
    order) and fp64 accumulation of the BatchNorm statistics in the GEMM epilogues (pnx_set_deterministic)."""
    global DETERMINISTIC
    prev, DETERMINISTIC = DETERMINISTIC, bool(on)
    lib().pnx_set_deterministic(1 if on else 0)
    return prev


def wgrad(X, x_channels, Y, y_channels, M, taps, dW, *, nbr=None, dense=None, shuffle=False, gathered=None):
    """dW[t, x, y] += sum_m X[m, x] * Y[g(m,t), y]; X direct, Y gathered; dW fp32 [taps, x_channels, y_channels]."""
    assert X.dtype == torch.bfloat16 and Y.dtype == torch.bfloat16 and dW.dtype == torch.float32
    assert tuple(dW.shape) == (taps, x_channels, y_channels) and dW.is_contiguous()
    if x_channels != 64 and x_channels % 128 != 0:
        # the kernel tiles X in 128-channel blocks (or one 64-channel block): split e.g. 448 = 384 + 64 column slices of X
        # (seven sibling heads of a Waymo task); each part accumulates into its own zeroed buffer, added to its dW rows
        assert x_channels % 64 == 0 and x_channels > 64, "x_channels must be a multiple of 64"
        main = x_channels // 128 * 128
        for x0, x1 in ((0, main), (main, x_channels)):
            part = torch.zeros(taps, x1 - x0, y_channels, dtype=torch.float32, device=dW.device)
            wgrad(X[:, x0:x1], x1 - x0, Y, y_channels, M, taps, part, nbr=nbr, dense=dense, shuffle=shuffle, gathered=gathered)
            dW[:, x0:x1] += part
        return dW
    d = dense or (0, 0, 0, 0, 1, 1, 1, 0)
    if gathered is None:
        gathered = nbr is not None or dense is not None or shuffle
    partials = None
    if DETERMINISTIC and M > 0:
        splits = lib().pnx_wgrad_splits(x_channels, y_channels, taps, M, sm_count())
        partials = torch.empty(splits * taps * x_channels * y_channels, dtype=torch.float32, device=dW.device)
    _count(2 if partials is not None else 1)
    with _Timed("wgrad", 2.0 * M * taps * x_channels * y_channels, 2.0 * M * (x_channels + taps * y_channels),
                "M%d_T%d_X%d_Y%d" % (M, taps, x_channels, y_channels)):
      check(lib().pnx_wgrad(ptr(X), X.stride(0), x_channels, ptr(Y), Y.stride(0), Y.shape[0], y_channels, 1 if gathered else 0, M,
                          taps, ptr(nbr) if nbr is not None else None, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7],
                          1 if shuffle else 0, ptr(dW), ptr(partials) if partials is not None else None, sm_count(), stream()))
    return dW


# --------------------------------------------------------------------------------- row-wise kernels
def bn_apply(x, M, C, scale, shift, y, res=None, relu=True):
    _count(1)
    check(lib().pnx_bn_apply(ptr(x), x.stride(0), M, C, ptr(scale), ptr(shift), ptr(res) if res is not None else None,
                             res.stride(0) if res is not None else 8, 1 if relu else 0, ptr(y), y.stride(0), stream()))
    return y


def bn_bwd(dy, y, x, M, C, mean, invstd, gamma, count, relu, dx, dres=None, dres_accumulate=False, affine=None, red=None):
    """Returns red (fp64 [2C]: sum g = dbeta, sum g*xhat = dgamma); writes dx (and dres).
    y=None (no residual): the ReLU mask is recomputed from x with affine=(scale, shift) of the forward.
    red given: the reduce pass already happened in the epilogue of the GEMM that produced dy (which is then the gated
    g, so the apply pass runs without a ReLU mask)."""
    L = lib()
    fs, fh = (affine if affine is not None else (None, None))
    yp, ys = (ptr(y), y.stride(0)) if y is not None else (None, 8)
    if red is None:
        red = zeros(2 * C, torch.float64, dy.device)
        _count(1)
        check(L.pnx_bn_bwd_reduce(ptr(dy), dy.stride(0), yp, ys, ptr(x), x.stride(0), M, C, ptr(mean), ptr(invstd),
                                  1 if relu else 0, ptr(fs) if fs is not None else None, ptr(fh) if fh is not None else None,
                                  ptr(red), stream()))
    else:
        relu = False
    _count(1)
    check(L.pnx_bn_bwd_apply(ptr(dy), dy.stride(0), yp, ys, ptr(x), x.stride(0), M, C, ptr(mean), ptr(invstd),
                             ptr(gamma), ptr(red), float(max(count, 1)), None, 1 if relu else 0,
                             ptr(fs) if fs is not None else None, ptr(fh) if fh is not None else None, ptr(dx),
                             dx.stride(0), ptr(dres) if dres is not None else None,
                             dres.stride(0) if dres is not None else 8, 1 if dres_accumulate else 0, stream()))
    return red


def add_rows(a, b, M, C):
    _count(1)
    check(lib().pnx_add_rows(ptr(a), a.stride(0), ptr(b), b.stride(0), M, C, stream()))
    return a


def add_relu(a, b, M, C, y):
    _count(1)
    check(lib().pnx_add_relu(ptr(a), a.stride(0), ptr(b), b.stride(0), M, C, ptr(y), y.stride(0), stream()))
    return y


def relu_bwd(dy, y, M, C, g, accumulate=False):
    _count(1)
    check(lib().pnx_relu_bwd(ptr(dy), dy.stride(0), ptr(y), y.stride(0), M, C, ptr(g), g.stride(0),
                             1 if accumulate else 0, stream()))
    return g


# --------------------------------------------------------------------------------- fp32-grade split rows
# A "split" activation is a bf16 tensor [M, P*C]: piece 0 = bf16(v) in columns [0, C), piece 1 = bf16(v - piece 0) in
# [C, 2C), ... (include/pnx.h, "split rows"; P = split_pieces(), 2 or 3).  The wrappers take (tensor-or-view, piece
# offset in elements) per operand.
SEGMENTS = {2: [(1, 1), (1, 0), (0, 1), (0, 0)],                       # (A piece, W piece), smallest products first
            3: [(2, 0), (0, 2), (1, 1), (1, 0), (0, 1), (0, 0)]}       # 2^-16-sized terms, 2^-8-sized terms, hi*hi


def split_pieces(set_to=None):
    if set_to is not None:
        lib().pnx_split_set_pieces(int(set_to))
    return lib().pnx_split_get_pieces()


def split_segments():
    return SEGMENTS[split_pieces()]


def to_pieces(p):
    """fp32 [..., K] -> bf16 [..., P*K] = [piece 0 | piece 1 | ...] (packed weights of the split mode)."""
    out, rem = [], p.float()
    for _ in range(split_pieces()):
        q = rem.to(torch.bfloat16)
        out.append(q)
        rem = rem - q.float()
    return torch.cat(out, -1).contiguous()


def rows_split(x, C=None, out=None, lo=None):
    """fp32 rows [M, >=C] -> split rows.  out/lo: write into an existing buffer (view) with the lo half at +lo."""
    assert x.dtype == torch.float32 and x.stride(-1) == 1
    M = x.shape[0]
    C = x.shape[1] if C is None else C
    if out is None:
        out = torch.empty(M, split_pieces() * C, dtype=torch.bfloat16, device=x.device)
        lo = C
    _count(1)
    check(lib().pnx_rows_split(ptr(x), x.stride(0), M, C, ptr(out), out.stride(0), int(lo), stream()))
    return out


def rows_merge(x, C, lo, out=None, accumulate=False):
    """split rows -> fp32 [M, C] (hi + lo)."""
    M = x.shape[0]
    if out is None:
        out = torch.empty(M, C, dtype=torch.float32, device=x.device)
    _count(1)
    check(lib().pnx_rows_merge(ptr(x), x.stride(0), int(lo), M, C, ptr(out), out.stride(0), 1 if accumulate else 0, stream()))
    return out


def bn_apply_split(x, M, C, scale, shift, y, y_lo, res=None, res_lo=0, relu=True):
    _count(1)
    check(lib().pnx_bn_apply_split(ptr(x), x.stride(0), M, C, ptr(scale), ptr(shift), ptr(res) if res is not None else None,
                                   res.stride(0) if res is not None else 16, int(res_lo) if res is not None else 16, 1 if relu else 0,
                                   ptr(y), y.stride(0), int(y_lo), stream()))
    return y


def bn_bwd_reduce_split(dy, dy_lo, y, y_lo, x, M, C, mean, invstd, relu, affine):
    L = lib()
    part = torch.empty(L.pnx_bn_bwd_reduce_split_scratch(C), dtype=torch.float64, device=dy.device)
    red = torch.empty(2 * C, dtype=torch.float64, device=dy.device)
    fs, fh = affine if affine is not None else (None, None)
    _count(2)
    check(L.pnx_bn_bwd_reduce_split(ptr(dy), dy.stride(0), int(dy_lo), ptr(y) if y is not None else None,
                                    y.stride(0) if y is not None else 16, int(y_lo) if y is not None else 16, ptr(x), x.stride(0), M, C,
                                    ptr(mean), ptr(invstd), 1 if relu else 0, ptr(fs) if fs is not None else None,
                                    ptr(fh) if fh is not None else None, ptr(part), ptr(red), stream()))
    return red


def bn_bwd_apply_split(dy, dy_lo, y, y_lo, x, M, C, mean, invstd, gamma, red, count, relu, affine, dx, dx_lo, dres=None, dres_lo=0):
    fs, fh = affine if affine is not None else (None, None)
    _count(1)
    check(lib().pnx_bn_bwd_apply_split(ptr(dy), dy.stride(0), int(dy_lo), ptr(y) if y is not None else None,
                                       y.stride(0) if y is not None else 16, int(y_lo) if y is not None else 16, ptr(x), x.stride(0), M, C,
                                       ptr(mean), ptr(invstd), ptr(gamma), ptr(red), float(max(count, 1)), 1 if relu else 0,
                                       ptr(fs) if fs is not None else None, ptr(fh) if fh is not None else None, ptr(dx), dx.stride(0),
                                       int(dx_lo), ptr(dres) if dres is not None else None,
                                       dres.stride(0) if dres is not None else 16, int(dres_lo) if dres is not None else 16, stream()))
    return dx


def add_relu_split(a, a_lo, b, b_lo, M, C, y, y_lo):
    _count(1)
    check(lib().pnx_add_relu_split(ptr(a), a.stride(0), int(a_lo), ptr(b), b.stride(0), int(b_lo), M, C, ptr(y), y.stride(0),
                                   int(y_lo), stream()))
    return y


def relu_bwd_split(dy_f32, y, y_lo, M, C, g, g_lo):
    _count(1)
    check(lib().pnx_relu_bwd_split(ptr(dy_f32), dy_f32.stride(0), ptr(y), y.stride(0), int(y_lo), M, C, ptr(g), g.stride(0),
                                   int(g_lo), stream()))
    return g


def wgrad_split(X, x_lo, x_channels, Y, y_lo, y_channels, M, taps, dW, **kw):
    """Weight gradient of split operands: one launch per (X piece, Y piece) segment, smallest products first, all
    accumulating (fp32 red.global.add) into the same dW."""
    for xp, yp in split_segments():
        wgrad(X[:, xp * x_lo:xp * x_lo + x_channels], x_channels, Y[:, yp * y_lo:yp * y_lo + y_channels], y_channels, M, taps, dW, **kw)
    return dW


# --------------------------------------------------------------------------------- F3: label assignment
_cls_tables = {}


def pack_weights(table, n, total):
    """pnx_pack_weights: rewrite all bf16 packed operands described by the device table (functional._repack_all)."""
    _count(1)
    check(lib().pnx_pack_weights(ptr(table), int(n), int(total), stream()))


def assign_labels(gt_boxes, gt_cls, tasks, voxel_size, pc_range, out_size_factor, max_objs=500, gaussian_overlap=0.1,
                  min_radius=2):
    """AssignLabel + collate on the GPU (pnx_assign_labels): gt_boxes [B, N, 9] fp32 cuda, gt_cls [B, N] int32 cuda
    (index into the flattened class list of `tasks`, < 0 = ignored) -> dict of per-task lists hm / anno_box / ind /
    mask / cat / gt_boxes in the reference's collate format (det3d/datasets/pipelines/assign.py, loader/collate.py)."""
    assert gt_boxes.is_cuda and gt_boxes.dtype == torch.float32 and gt_boxes.dim() == 3 and gt_boxes.shape[2] == 9
    assert gt_cls.is_cuda and gt_cls.dtype == torch.int32 and tuple(gt_cls.shape) == tuple(gt_boxes.shape[:2])
    gt_boxes, gt_cls = gt_boxes.contiguous(), gt_cls.contiguous()
    dev = gt_boxes.device
    B, N = gt_cls.shape
    key = (tuple(tuple(t) for t in tasks), dev)
    if key not in _cls_tables:
        ct = [ti for ti, t in enumerate(tasks) for _ in t]
        ci = [ni for t in tasks for ni, _ in enumerate(t)]
        _cls_tables[key] = (torch.tensor(ct, dtype=torch.int32, device=dev), torch.tensor(ci, dtype=torch.int32, device=dev))
    cls_task, cls_id = _cls_tables[key]
    g = grid_size_xy(voxel_size, pc_range)
    out = {k: [] for k in ("hm", "anno_box", "ind", "mask", "cat", "gt_boxes")}
    L = lib()
    obj = torch.empty(B, max(N, 1), 4, dtype=torch.int32, device=dev)
    for t, task in enumerate(tasks):
        osf = int(out_size_factor[t])
        W, H = int(g[0]) // osf, int(g[1]) // osf
        hm = torch.zeros(B, len(task), H, W, dtype=torch.float32, device=dev)
        anno = torch.zeros(B, max_objs, 10, dtype=torch.float32, device=dev)
        ind = torch.zeros(B, max_objs, dtype=torch.int64, device=dev)
        mask = torch.zeros(B, max_objs, dtype=torch.uint8, device=dev)
        cat = torch.zeros(B, max_objs, dtype=torch.int64, device=dev)
        gtb = torch.zeros(B, max_objs, 7, dtype=torch.float32, device=dev)
        _count(2)
        check(L.pnx_assign_labels(ptr(gt_boxes), ptr(gt_cls), B, N, ptr(cls_task), ptr(cls_id), cls_task.numel(), t,
                                  float(voxel_size[0]), float(voxel_size[1]), float(pc_range[0]), float(pc_range[1]), osf,
                                  float(gaussian_overlap), int(min_radius), int(max_objs), len(task), H, W, ptr(hm), ptr(anno),
                                  ptr(ind), ptr(mask), ptr(cat), ptr(gtb), ptr(obj), stream()))
        for k, v in (("hm", hm), ("anno_box", anno), ("ind", ind), ("mask", mask), ("cat", cat), ("gt_boxes", gtb)):
            out[k].append(v)
    return out


# --------------------------------------------------------------------------------- F1: decode + rotated NMS
def _host_floats(vals, n):
    import ctypes
    a = (ctypes.c_float * n)(*([float(v) for v in vals] + [0.0] * (n - len(vals))))
    return a, ctypes.addressof(a)


def det_postprocess(out, B, H, W, C, offs, osf, voxel_size, pc_range, score_thr, post_center_range, rectifier, nms_thr,
                    pre_max, post_max, label_offset=0):
    """One task of CenterHead.predict (centerhead.py:231-384): `out` = channels-last head output [B*H*W, ld] fp32,
    offs = column offsets (reg, height, dim, rot, vel, hm, iou or -1).  Returns det_box [B*C, post_max, 9], det_score
    [B*C, post_max], det_label [B*C, post_max] (int64, + label_offset) and keep_count [B*C] -- all on the device, no
    host synchronisation.  Segment s = frame * C + class.  The candidate ordering is one device sort (torch.sort of
    the 64-bit keys the decode kernel builds); decode, IoU mask, greedy sweep and gather are libpnx kernels."""
    import ctypes
    assert out.dtype == torch.float32 and out.is_cuda and out.dim() == 2 and out.shape[0] == B * H * W
    dev = out.device
    M, nseg = B * H * W, B * C
    offs_a = (ctypes.c_int * 7)(*[int(o) for o in offs])
    r6, r6p = _host_floats(post_center_range, 6)
    rc, rcp = _host_floats(rectifier, 8)
    nt, ntp = _host_floats(nms_thr, 8)
    common = (ptr(out), out.stride(0), B, H, W, C, ctypes.addressof(offs_a), float(osf), float(voxel_size[0]),
              float(voxel_size[1]), float(pc_range[0]), float(pc_range[1]), float(score_thr), r6p, rcp)
    keys = torch.empty(M, dtype=torch.int64, device=dev)
    seg_count = torch.empty(nseg, dtype=torch.int32, device=dev)
    L = lib()
    _count(1)
    check(L.pnx_det_keys(*common, ptr(keys), ptr(seg_count), stream()))
    order = torch.sort(keys)[1]
    seg_start = (torch.cumsum(seg_count, 0, dtype=torch.int32) - seg_count).contiguous()
    col_blocks = (pre_max + 63) // 64
    mask = torch.empty(nseg * pre_max * col_blocks, dtype=torch.int64, device=dev)
    keep = torch.empty(nseg * post_max, dtype=torch.int32, device=dev)
    keep_count = torch.empty(nseg, dtype=torch.int32, device=dev)
    _count(2)
    check(L.pnx_det_nms(*common, ntp, ptr(order), ptr(seg_start), ptr(seg_count), int(pre_max), int(post_max), ptr(mask),
                        ptr(keep), ptr(keep_count), stream()))
    det_box = torch.empty(nseg, post_max, 9, dtype=torch.float32, device=dev)
    det_score = torch.empty(nseg, post_max, dtype=torch.float32, device=dev)
    det_label = torch.empty(nseg, post_max, dtype=torch.int64, device=dev)
    _count(1)
    check(L.pnx_det_gather(*common, ptr(order), ptr(seg_start), ptr(keep), ptr(keep_count), int(post_max), int(label_offset),
                           ptr(det_box), ptr(det_score), ptr(det_label), stream()))
    return det_box, det_score, det_label, keep_count


def aligned_iou3d(boxes_a, boxes_b):
    """[n, 7] x [n, 7] fp32 (x, y, z, dx, dy, dz, heading) -> [n] aligned 3-D IoU (iou3d_nms_utils.py:45-87)."""
    assert boxes_a.shape == boxes_b.shape and boxes_a.shape[-1] == 7 and boxes_a.is_cuda
    a, b = boxes_a.contiguous().float(), boxes_b.contiguous().float()
    out = torch.empty(a.shape[0], dtype=torch.float32, device=a.device)
    _count(1)
    check(lib().pnx_aligned_iou3d(ptr(a), ptr(b), a.shape[0], ptr(out), stream()))
    return out
