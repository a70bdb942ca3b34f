// wgrad.cu -- weight-gradient of every convolution on the path as a tcgen05 GEMM over rows (sm_100a).
//
//   dW[t, x, y] += sum_m  X[ix(m,t), x] * Y[iy(m,t), y]          fp32, accumulated with red.global.add
// X / Y are the two row-major bf16 operands of the convolution at tap t: the layer input gathered
// through the neighbour table (or the dense geometry) and the output gradient read directly -- in
// either role, so the 128-wide MMA M side can always be the wider channel count.
// Both operands are MN-major for the MMA (the reduction index K = row m is the slow one in memory):
// tiles are staged as [64 rows x 128 B] blocks per 64-channel group with the 128-byte swizzle and
// described to tcgen05.mma with MN-major descriptors (LBO = block stride, SBO = 8-row group stride).
// Grid = (x groups, taps, K splits); each CTA owns one fp32 accumulator set in TMEM for its K range.
// This is the backward of: spconv SparseConv2d/SubMConv2d (reference sparse_conv.py:25-29,50-51),
// nn.Conv2d/F.conv2d (aspp.py:19-32, conv.py:9-10, centerhead.py:35-46,108-114), nn.ConvTranspose2d
// (centerhead.py:26-27) -- autograd derives these in the reference (trainer.py:94-108 loss.backward()).
#include "pnx_common.cuh"

namespace {

struct WgradParams {
  const __nv_bfloat16* X;
  const __nv_bfloat16* Y;
  long long ldx, ldy;
  int M, T;
  int gather_x, gather_y;  // which operand goes through the neighbour map (0/1)
  const int* nbr;          // [M, T] or null
  int dense, Hout, Wout, Hin, Win, kw, mul, dil, pad;
  int shuffle;             // ConvTranspose k2s2: tap q selects output pixel (2y+q/2, 2x+q%2) of the gathered operand
  float* dW;               // [T, X_total, Y_total]
  int X_total, Y_total;
  int x_dup;               // X has only 64 channels: second MN atom aliases the first (rows 64..127 ignored)
  int rows_per_split;
};

constexpr int kThreads = 320;
constexpr int kKS = 64;               // rows (K) per stage
constexpr uint32_t kBlk = kKS * 128;  // bytes of one [64 rows x 64 ch] block
constexpr int kLag = 2;

template <int NY, int XB>
struct WCfg {
  static constexpr int kXBlocks = XB * 2;
  static constexpr int kYBlocks = NY / 64;
  static constexpr uint32_t kStageBytes = (kXBlocks + kYBlocks) * kBlk;
  static constexpr int kStagesRaw = (192 * 1024) / (int)kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr size_t kSmem = 1024 + (size_t)kStages * kStageBytes + 256;
};

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return pnx::make_smem_desc_sw128(addr, lbo, sbo);
}

__device__ __forceinline__ int map_row(const WgradParams& p, int m, int t) {
  if (p.nbr) return p.nbr[(size_t)m * p.T + t];
  const int hw = p.Hout * p.Wout;
  const int b = m / hw, rem = m - b * hw;
  const int y = rem / p.Wout, x = rem - y * p.Wout;
  if (p.shuffle) return (b * 2 * p.Hout + 2 * y + (t >> 1)) * (2 * p.Wout) + 2 * x + (t & 1);
  const int r = t / p.kw, s = t - r * p.kw;
  const int yi = y * p.mul + r * p.dil - p.pad, xi = x * p.mul + s * p.dil - p.pad;
  return (yi >= 0 && yi < p.Hin && xi >= 0 && xi < p.Win) ? (b * p.Hin + yi) * p.Win + xi : -1;
}

template <int NY, int XB>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(WgradParams p) {
  using C = WCfg<NY, XB>;
  constexpr int kStages = C::kStages;
  constexpr int kXBlocks = C::kXBlocks, kYBlocks = C::kYBlocks;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * C::kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* done = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xg = blockIdx.x, t = blockIdx.y, split = blockIdx.z;
  const int m_begin = split * p.rows_per_split;
  const int m_end = min(p.M, m_begin + p.rows_per_split);
  const int num_k = (m_end - m_begin + kKS - 1) / kKS;

  if (warp == 0 && pnx::elect_one()) {
    for (int s = 0; s < kStages; ++s) {
      pnx::mbar_init(&full[s], 4);
      pnx::mbar_init(&empty[s], 1);
    }
    pnx::mbar_init(done, 1);
    pnx::fence_barrier_init();
  }
  if (warp == 1) pnx::tmem_alloc<512>(tmem_slot);
  pnx::tc_fence_before();
  __syncthreads();
  pnx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (num_k > 0) {
    if (warp == 1) {
      if (pnx::elect_one()) {
        constexpr uint32_t idesc = pnx::make_idesc_bf16(128, NY, 1, 1);
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t x_lbo = p.x_dup ? 0u : kBlk;
        for (int kc = 0; kc < num_k; ++kc) {
          pnx::mbar_wait(&full[stage], phase);
          pnx::tc_fence_after();
          const uint32_t sx = pnx::smem_u32(smem + (size_t)stage * C::kStageBytes);
          const uint32_t sy = sx + kXBlocks * kBlk;
#pragma unroll
          for (int k = 0; k < kKS / 16; ++k) {
            const uint64_t dy = make_desc_mn_sw128(sy + k * 2048, kBlk, 1024);
#pragma unroll
            for (int xb = 0; xb < XB; ++xb) {
              const uint64_t dx = make_desc_mn_sw128(sx + xb * 2 * kBlk + k * 2048, x_lbo, 1024);
              pnx::umma_f16(tmem_base + xb * NY, dx, dy, idesc, (kc > 0 || k > 0) ? 1u : 0u);
            }
          }
          pnx::umma_commit(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        pnx::umma_commit(done);
      }
    } else if (warp >= 2 && warp < 6) {
      const int ptid = threadIdx.x - 64;
      const int sub_row = ptid >> 3, chunk = ptid & 7;
      const int x_ch0 = xg * XB * 128;
      int stage = 0, arr_stage = 0, pending = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < num_k; ++kc) {
        pnx::mbar_wait(&empty[stage], phase ^ 1);
        const uint32_t sx = pnx::smem_u32(smem + (size_t)stage * C::kStageBytes);
        const uint32_t sy = sx + kXBlocks * kBlk;
#pragma unroll
        for (int j = 0; j < kKS / 16; ++j) {
          const int r = j * 16 + sub_row;
          const int m = m_begin + kc * kKS + r;
          int ix = -1, iy = -1;
          if (m < m_end) {
            const int g = map_row(p, m, t);
            ix = p.gather_x ? g : m;
            iy = p.gather_y ? g : m;
            if (ix < 0 || iy < 0) ix = iy = -1;
          }
          const uint32_t off = r * 128 + ((chunk ^ (r & 7)) << 4);
          const __nv_bfloat16* xs = p.X + (size_t)(ix < 0 ? 0 : ix) * p.ldx + x_ch0 + chunk * 8;
          const __nv_bfloat16* ys = p.Y + (size_t)(iy < 0 ? 0 : iy) * p.ldy + chunk * 8;
          const uint32_t nb = ix < 0 ? 0u : 16u;
          const int nxb = p.x_dup ? 1 : kXBlocks;
          for (int b = 0; b < nxb; ++b) pnx::cp_async16(sx + b * kBlk + off, xs + b * 64, nb);
#pragma unroll
          for (int b = 0; b < kYBlocks; ++b) pnx::cp_async16(sy + b * kBlk + off, ys + b * 64, nb);
        }
        pnx::cp_async_commit();
        if (pending == kLag) {
          pnx::cp_async_wait<kLag>();
          pnx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) pnx::mbar_arrive(&full[arr_stage]);
          if (++arr_stage == kStages) arr_stage = 0;
        } else {
          ++pending;
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      pnx::cp_async_wait<0>();
      pnx::fence_proxy_async_smem();
      __syncwarp();
      for (; pending > 0; --pending) {
        if (lane == 0) pnx::mbar_arrive(&full[arr_stage]);
        if (++arr_stage == kStages) arr_stage = 0;
      }
    } else if (warp >= 6) {
      const int quarter = warp & 3;
      pnx::mbar_wait(done, 0);
      pnx::tc_fence_after();
      const int xrow_local = quarter * 32 + lane;
#pragma unroll
      for (int xb = 0; xb < XB; ++xb) {
        const int xch = xg * XB * 128 + xb * 128 + xrow_local;
        const bool ok = xch < p.X_total && !(p.x_dup && xrow_local >= 64);
        float* dst = p.dW + ((size_t)t * p.X_total + (ok ? xch : 0)) * p.Y_total;
#pragma unroll
        for (int cb = 0; cb < NY / 32; ++cb) {
          uint32_t r[32];
          pnx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + xb * NY + cb * 32, r);
          pnx::tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int k = 0; k < 32; ++k) atomicAdd(dst + cb * 32 + k, __uint_as_float(r[k]));
          }
        }
      }
    }
  }
  pnx::tc_fence_before();
  __syncthreads();
  pnx::tc_fence_after();
  if (warp == 1) pnx::tmem_dealloc<512>(tmem_base);
}

template <int NY, int XB>
int launch_wgrad(const WgradParams& p, int x_groups, int splits, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    PNX_CUDA(cudaFuncSetAttribute(wgrad_kernel<NY, XB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)WCfg<NY, XB>::kSmem));
    attr_set = true;
  }
  dim3 grid(x_groups, p.T, splits);
  wgrad_kernel<NY, XB><<<grid, kThreads, WCfg<NY, XB>::kSmem, stream>>>(p);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

}  // namespace

// Contract: include/pnx.h (pnx_wgrad).  dW must be zeroed (or hold the value to accumulate into).
extern "C" int pnx_wgrad(const void* X, long long ldx, int x_channels, int gather_x, const void* Y, long long ldy,
                         int y_channels, int gather_y, int M, int taps, const int* nbr, int dense, int Hout, int Wout,
                         int Hin, int Win, int kw, int mul, int dil, int pad, int shuffle, float* dW, int sm_count,
                         cudaStream_t stream) {
  PNX_CHECK_ARG(M >= 0, "M");
  if (M == 0) return PNX_OK;
  PNX_CHECK_ARG(taps >= 1 && taps <= 9, "taps");
  PNX_CHECK_ARG(x_channels % 64 == 0 && y_channels % 64 == 0, "channel counts must be multiples of 64");
  PNX_CHECK_ARG(y_channels <= 256, "y_channels <= 256 (put the wider operand on X)");
  PNX_CHECK_ARG(x_channels == 64 || x_channels % 128 == 0, "x_channels must be 64 or a multiple of 128");
  PNX_CHECK_ARG(gather_x + gather_y <= 1, "at most one gathered operand");
  PNX_CHECK_ARG(!(gather_x + gather_y) || nbr || dense || shuffle, "gather needs a table or geometry");
  PNX_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0, "ldx/ldy % 8");
  if (sm_count <= 0) sm_count = 148;
  WgradParams p;
  p.X = (const __nv_bfloat16*)X; p.Y = (const __nv_bfloat16*)Y;
  p.ldx = ldx; p.ldy = ldy; p.M = M; p.T = taps;
  p.gather_x = gather_x; p.gather_y = gather_y; p.nbr = nbr; p.dense = dense;
  p.Hout = Hout > 0 ? Hout : 1; p.Wout = Wout > 0 ? Wout : 1; p.Hin = Hin; p.Win = Win;
  p.kw = kw > 0 ? kw : 1; p.mul = mul; p.dil = dil; p.pad = pad; p.shuffle = shuffle;
  p.dW = dW; p.X_total = x_channels; p.Y_total = y_channels;
  p.x_dup = x_channels == 64 ? 1 : 0;
  // one CTA owns XB*128 X channels; XB=2 only when the TMEM (512 columns) and smem budgets allow
  const int xb = (x_channels % 256 == 0 && y_channels <= 256) ? 2 : 1;
  const int x_groups = x_channels == 64 ? 1 : x_channels / (128 * xb);
  int splits = (2 * sm_count) / (x_groups * taps);
  const int max_splits = (M + 4 * kKS - 1) / (4 * kKS);  // at least 4 K-chunks per CTA
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.rows_per_split = ((M + splits - 1) / splits + kKS - 1) / kKS * kKS;
  splits = (M + p.rows_per_split - 1) / p.rows_per_split;
#define PNX_WG(NYV, XBV) \
  if (y_channels == NYV && xb == XBV) return launch_wgrad<NYV, XBV>(p, x_groups, splits, stream);
  PNX_WG(64, 1) PNX_WG(64, 2) PNX_WG(128, 1) PNX_WG(128, 2) PNX_WG(192, 1) PNX_WG(192, 2) PNX_WG(256, 1) PNX_WG(256, 2)
#undef PNX_WG
  pnx_set_error("pnx_wgrad: unsupported shape x=%d y=%d", x_channels, y_channels);
  return PNX_ERR_ARG;
}
