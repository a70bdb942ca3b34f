// loss.cu -- fused CenterPoint loss (row L1 of SURVEY.md section 8a): forward value AND the gradient w.r.t. the
// head output in one pass per task, no host synchronisation.
//
// Same arithmetic as reference det3d/models/heads/centerhead.py:142-229 + det3d/models/loss/centerloss.py:
//   FastFocalLoss (:17-37)  on p = clamp(sigmoid(hm), 1e-4, 1-1e-4) (centerhead.py:138-140)
//   RegLoss       (:53-61)  L1 on (reg2, height1, dim3, vel2, rot2) gathered at `ind`, NaN targets ignored
//   IouRegLoss    (:103-110) + bbox3d_overlaps_diou (:139-176): axis-aligned DIoU of the decoded box
// The reference runs ~150 elementwise/gather kernels per task and syncs the host 3-4 times (`.cpu()`,
// `if num_pos == 0`); here: one dense kernel over the heat map (negative focal term + its gradient, also zeroing
// the rest of the gradient tensor), one kernel over the <= B*500 positive slots, one finalise kernel for all tasks.
// The head output is the channels-last fp32 matrix [B*H*W, npad] written by the last head GEMM
// (columns: reg2 | height1 | dim3 | rot2 | vel2 | hm C | zero padding).
#include "pnx_common.cuh"

namespace {

struct LossTask {
  const float* out;        // [B*H*W, npad]
  float* dout;             // same shape, fully written
  const float* hm_gt;      // [B, C, H, W]
  const float* anno;       // [B, M, 10]  (reg2, height1, dim3, vel2, rot2)
  const long long* ind;    // [B, M]
  const unsigned char* mask;  // [B, M]
  const long long* cat;    // [B, M]
  const float* gt_boxes;   // [B, M, 7]
  int B, H, W, npad, C, M;
  int off_reg, off_height, off_dim, off_rot, off_vel, off_hm;
  float sx, sy, ox, oy;    // xs = (col + reg_x) * sx + ox   (sx = out_size_factor * voxel_x, ox = pc_range[0])
  float weight;
  float code_w[10];
  int with_iou;
  double* acc;             // [16]: 0 neg, 1 pos, 2 npos, 3..12 sum|pred-tgt|*mask per code, 13 sum(1-diou)
};

__device__ __forceinline__ float block_sum(float v, float* sred) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? sred[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in thread 0 (warp 0)
}

__device__ __forceinline__ float count_pos(const LossTask& t, float* sred) {
  float c = 0.f;
  for (int i = threadIdx.x; i < t.B * t.M; i += blockDim.x) c += t.mask[i] ? 1.f : 0.f;
  __shared__ float s_np;
  float tot = block_sum(c, sred);
  if (threadIdx.x == 0) s_np = tot;
  __syncthreads();
  return s_np;
}

// ---- dense part: negative focal term + gradient; zero-fills every other column of dout
__global__ void __launch_bounds__(256) loss_dense_kernel(LossTask t) {
  __shared__ float sred[8];
  const float npos = count_pos(t, sred);
  const float scale = -1.f / fmaxf(npos, 1.f);  // d(loss)/d(neg_sum) (centerloss.py:35-37)
  const long long npix = (long long)t.B * t.H * t.W;
  const int hw = t.H * t.W;
  float neg = 0.f;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(pix / hw);
    const int rem = (int)(pix - (long long)b * hw);
    const float* o = t.out + pix * t.npad;
    float* d = t.dout + pix * t.npad;
    for (int k = 0; k < t.npad; k += 4) *reinterpret_cast<float4*>(d + k) = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < t.C; ++c) {
      const float x = o[t.off_hm + c];
      const float ps = 1.f / (1.f + __expf(-x));
      const bool clamped = ps < 1e-4f || ps > 1.f - 1e-4f;
      const float p = fminf(fmaxf(ps, 1e-4f), 1.f - 1e-4f);
      const float g = 1.f - t.hm_gt[((long long)b * t.C + c) * hw + rem];
      const float g4 = g * g * g * g;
      const float l1p = __logf(1.f - p);
      neg += p * p * g4 * l1p;
      // d/dx [p^2 g4 log(1-p)] = g4 p^2 (2 (1-p) log(1-p) - p)     (zero where the clamp is active)
      d[t.off_hm + c] = clamped ? 0.f : scale * g4 * p * p * (2.f * (1.f - p) * l1p - p);
    }
  }
  const float tot = block_sum(neg, sred);
  if (threadIdx.x == 0) atomicAdd(&t.acc[0], (double)tot);
}

// ---- forward-mode dual numbers over the six box variables (x, y, z, dx, dy, dz) for the DIoU gradient
struct D6 {
  float v, d[6];
};
__device__ __forceinline__ D6 dvar(float v, int i) { D6 r; r.v = v; for (int k = 0; k < 6; ++k) r.d[k] = k == i ? 1.f : 0.f; return r; }
__device__ __forceinline__ D6 dcst(float v) { D6 r; r.v = v; for (int k = 0; k < 6; ++k) r.d[k] = 0.f; return r; }
__device__ __forceinline__ D6 operator+(const D6& a, const D6& b) { D6 r; r.v = a.v + b.v; for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] + b.d[k]; return r; }
__device__ __forceinline__ D6 operator-(const D6& a, const D6& b) { D6 r; r.v = a.v - b.v; for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] - b.d[k]; return r; }
__device__ __forceinline__ D6 operator*(const D6& a, const D6& b) { D6 r; r.v = a.v * b.v; for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k]; return r; }
__device__ __forceinline__ D6 operator/(const D6& a, const D6& b) { D6 r; r.v = a.v / b.v; for (int k = 0; k < 6; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) / b.v; return r; }
__device__ __forceinline__ D6 dscale(const D6& a, float s) { D6 r; r.v = a.v * s; for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] * s; return r; }
// torch.minimum/maximum: gradient to the selected operand (ties: split evenly, like torch)
__device__ __forceinline__ D6 dmin(const D6& a, const D6& b) { if (a.v < b.v) return a; if (b.v < a.v) return b; return dscale(a + b, 0.5f); }
__device__ __forceinline__ D6 dmax(const D6& a, const D6& b) { if (a.v > b.v) return a; if (b.v > a.v) return b; return dscale(a + b, 0.5f); }
__device__ __forceinline__ D6 dclamp_min0(const D6& a) { return a.v >= 0.f ? a : dcst(0.f); }   // torch.clamp(min=0): grad 1 at the boundary

// ---- positive slots: focal positive term, L1 regression, DIoU; gradients accumulated with atomics (two objects
//      may share a pixel)
__global__ void __launch_bounds__(128) loss_pos_kernel(LossTask t) {
  __shared__ float sred[8];
  const float npos = count_pos(t, sred);
  const float scale_f = -1.f / fmaxf(npos, 1.f);
  const float inv_np = 1.f / (npos + 1e-4f);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float pos = 0.f, el[10], iou_l = 0.f;
#pragma unroll
  for (int c = 0; c < 10; ++c) el[c] = 0.f;
  bool slot = i < t.B * t.M && t.mask[i];
  if (slot && (t.ind[i] < 0 || t.ind[i] >= (long long)t.H * t.W || t.cat[i] < 0 || t.cat[i] >= t.C)) {
    // a label built for another feature-map size / class list (the reference's gather raises an index error): skip the
    // slot instead of writing outside the head matrix, and count it in acc[15] (res[15] = number of bad slots)
    atomicAdd(&t.acc[15], 1.0);
    slot = false;
  }
  if (slot) {
    const int b = i / t.M;
    const long long pix = (long long)b * t.H * t.W + t.ind[i];
    const float* o = t.out + pix * t.npad;
    float* d = t.dout + pix * t.npad;
    // focal, positive location (centerloss.py:29-33)
    {
      const int c = (int)t.cat[i];
      const float ps = 1.f / (1.f + __expf(-o[t.off_hm + c]));
      const bool clamped = ps < 1e-4f || ps > 1.f - 1e-4f;
      const float p = fminf(fmaxf(ps, 1e-4f), 1.f - 1e-4f);
      const float lp = __logf(p);
      pos = lp * (1.f - p) * (1.f - p);
      // d/dx [log p (1-p)^2] = (1-p)^3 - 2 p (1-p)^2 log p
      if (!clamped) atomicAdd(&d[t.off_hm + c], scale_f * ((1.f - p) * (1.f - p) * (1.f - p) - 2.f * p * (1.f - p) * (1.f - p) * lp));
    }
    // regression (centerloss.py:53-61); anno order: reg2, height1, dim3, vel2, rot2 (centerhead.py:154-155)
    const int col[10] = {t.off_reg, t.off_reg + 1, t.off_height, t.off_dim, t.off_dim + 1, t.off_dim + 2,
                         t.off_vel, t.off_vel + 1, t.off_rot, t.off_rot + 1};
    const float* a = t.anno + (long long)i * 10;
#pragma unroll
    for (int c = 0; c < 10; ++c) {
      const float pr = o[col[c]];
      const float tg = a[c];
      if (!(tg != tg)) {  // NaN target -> replaced by the prediction: zero loss, zero gradient
        const float diff = pr - tg;
        el[c] = fabsf(diff);
        const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        if (sg != 0.f) atomicAdd(&d[col[c]], sg * t.code_w[c] * t.weight * inv_np);
      }
    }
    if (t.with_iou) {
      // decode (centerhead.py:172-209): xs = (col + reg_x) * osf * voxel_x + pc_min_x, dim = exp(clamp(dim, -5, 5))
      const int rem = (int)t.ind[i];
      const int yy = rem / t.W, xx = rem - yy * t.W;
      D6 px = dvar(((float)xx + o[t.off_reg]) * t.sx + t.ox, 0);
      D6 py = dvar(((float)yy + o[t.off_reg + 1]) * t.sy + t.oy, 1);
      D6 pz = dvar(o[t.off_height], 2);
      float dl[3], dv[3];
      bool dcl[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        dl[k] = o[t.off_dim + k];
        dcl[k] = dl[k] < -5.f || dl[k] > 5.f;
        dv[k] = __expf(fminf(fmaxf(dl[k], -5.f), 5.f));
      }
      D6 pdx = dvar(dv[0], 3), pdy = dvar(dv[1], 4), pdz = dvar(dv[2], 5);
      const float* g = t.gt_boxes + (long long)i * 7;
      const D6 gx = dcst(g[0]), gy = dcst(g[1]), gz = dcst(g[2]), gdx = dcst(g[3]), gdy = dcst(g[4]), gdz = dcst(g[5]);
      const D6 h = dcst(0.5f);
      // bbox3d_overlaps_diou (centerloss.py:139-176)
      D6 ix = dclamp_min0(dmin(px + h * pdx, gx + h * gdx) - dmax(px - h * pdx, gx - h * gdx));
      D6 iy = dclamp_min0(dmin(py + h * pdy, gy + h * gdy) - dmax(py - h * pdy, gy - h * gdy));
      D6 ih = dclamp_min0(dmin(pz + h * pdz, gz + h * gdz) - dmax(pz - h * pdz, gz - h * gdz));
      D6 oxx = dclamp_min0(dmax(px + h * pdx, gx + h * gdx) - dmin(px - h * pdx, gx - h * gdx));
      D6 oyy = dclamp_min0(dmax(py + h * pdy, gy + h * gdy) - dmin(py - h * pdy, gy - h * gdy));
      D6 ohh = dclamp_min0(dmax(gz + h * gdz, pz + h * pdz) - dmin(gz - h * gdz, pz - h * pdz));
      D6 vi = ix * iy * ih;
      D6 vu = gdx * gdy * gdz + pdx * pdy * pdz - vi;
      D6 ex = gx - px, ey = gy - py, ez = gz - pz;
      D6 idiag = ex * ex + ey * ey + ez * ez;
      D6 odiag = oxx * oxx + oyy * oyy + ohh * ohh;
      D6 diou = vi / vu - idiag / odiag;
      const bool cl = diou.v < -1.f || diou.v > 1.f;
      const float dval = fminf(fmaxf(diou.v, -1.f), 1.f);
      iou_l = 1.f - dval;
      if (!cl) {
        const float s = -t.weight * inv_np;  // d(weight * sum(1-diou)/(npos+1e-4)) / d(diou)
        atomicAdd(&d[t.off_reg], s * diou.d[0] * t.sx);
        atomicAdd(&d[t.off_reg + 1], s * diou.d[1] * t.sy);
        atomicAdd(&d[t.off_height], s * diou.d[2]);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (!dcl[k]) atomicAdd(&d[t.off_dim + k], s * diou.d[3 + k] * dv[k]);
      }
    }
  }
  float v = block_sum(pos, sred);
  if (threadIdx.x == 0 && v != 0.f) atomicAdd(&t.acc[1], (double)v);
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    v = block_sum(el[c], sred);
    if (threadIdx.x == 0 && v != 0.f) atomicAdd(&t.acc[3 + c], (double)v);
  }
  v = block_sum(iou_l, sred);
  if (threadIdx.x == 0 && v != 0.f) atomicAdd(&t.acc[13], (double)v);
  if (blockIdx.x == 0 && threadIdx.x == 0) t.acc[2] = (double)npos;
}

// res[task][16]: 0 loss, 1 hm_loss, 2 loc_loss, 3 iou_reg_loss, 4 num_positive, 5..14 loc_loss_elem; total[0] = sum of losses
__global__ void loss_finalize_kernel(const double* __restrict__ acc, int n_tasks, const float* __restrict__ weights,
                                     const float* __restrict__ code_w, const int* __restrict__ with_iou,
                                     float* __restrict__ res, float* __restrict__ total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float tot = 0.f;
  for (int t = 0; t < n_tasks; ++t) {
    const double* a = acc + t * 16;
    const float npos = (float)a[2];
    const float hm = -(float)(a[1] + a[0]) / fmaxf(npos, 1.f);
    float loc = 0.f;
    for (int c = 0; c < 10; ++c) {
      const float e = (float)a[3 + c] / (npos + 1e-4f);
      res[t * 16 + 5 + c] = e;
      loc += e * code_w[c];
    }
    float loss = hm + weights[t] * loc;
    float iou = 0.f;
    if (with_iou[t]) {
      iou = (float)a[13] / (npos + 1e-4f);
      loss += weights[t] * iou;
    }
    res[t * 16 + 0] = loss;
    res[t * 16 + 1] = hm;
    res[t * 16 + 2] = loc;
    res[t * 16 + 3] = iou;
    res[t * 16 + 4] = npos;
    res[t * 16 + 15] = (float)a[15];   // label slots skipped because ind / cat were out of range (0 for valid labels)
    tot += loss;
  }
  total[0] = tot;
}

}  // namespace

// Contract: include/pnx.h (pnx_center_loss_task / pnx_center_loss_finalize).  acc [16] fp64 must be zeroed.
extern "C" int pnx_center_loss_task(const float* out, float* dout, const float* hm_gt, const float* anno,
                                    const long long* ind, const unsigned char* mask, const long long* cat,
                                    const float* gt_boxes, int B, int H, int W, int npad, int C, int M, int off_reg,
                                    int off_height, int off_dim, int off_rot, int off_vel, int off_hm, float sx,
                                    float sy, float ox, float oy, float weight, const float* code_w_host, int with_iou,
                                    double* acc, cudaStream_t stream) {
  PNX_CHECK_ARG(B > 0 && H > 0 && W > 0 && npad % 4 == 0 && C >= 1 && M >= 0, "shapes");
  PNX_CHECK_ARG(off_hm + C <= npad, "hm columns exceed npad");
  LossTask t;
  t.out = out; t.dout = dout; t.hm_gt = hm_gt; t.anno = anno; t.ind = ind; t.mask = mask; t.cat = cat; t.gt_boxes = gt_boxes;
  t.B = B; t.H = H; t.W = W; t.npad = npad; t.C = C; t.M = M;
  t.off_reg = off_reg; t.off_height = off_height; t.off_dim = off_dim; t.off_rot = off_rot; t.off_vel = off_vel; t.off_hm = off_hm;
  t.sx = sx; t.sy = sy; t.ox = ox; t.oy = oy; t.weight = weight; t.with_iou = with_iou; t.acc = acc;
  for (int c = 0; c < 10; ++c) t.code_w[c] = code_w_host[c];
  const long long npix = (long long)B * H * W;
  int blocks = (int)((npix + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  loss_dense_kernel<<<blocks, 256, 0, stream>>>(t);
  PNX_CHECK_LAUNCH();
  if (B * M > 0) {
    loss_pos_kernel<<<(B * M + 127) / 128, 128, 0, stream>>>(t);
    PNX_CHECK_LAUNCH();
  }
  return PNX_OK;
}

extern "C" int pnx_center_loss_finalize(const double* acc, int n_tasks, const float* weights_dev,
                                        const float* code_w_dev, const int* with_iou_dev, float* res, float* total,
                                        cudaStream_t stream) {
  loss_finalize_kernel<<<1, 32, 0, stream>>>(acc, n_tasks, weights_dev, code_w_dev, with_iou_dev, res, total);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}
