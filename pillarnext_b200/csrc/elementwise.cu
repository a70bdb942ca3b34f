// elementwise.cu -- HBM-bound row-wise kernels around the convolutions (bf16 rows x channels):
//   BatchNorm apply (+residual)(+ReLU)                     forward of sparse_conv.py:33-39,55-63, conv.py:29-34,44-51
//   BatchNorm backward: reduce (sum g, sum g*xhat) + apply  (autograd of the same lines in the reference)
// 128-bit loads/stores, one thread = 8 consecutive channels of one row.
#include "pnx_common.cuh"

namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 t = __bfloat1622float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pnx::pack_bf16x2(f[0], f[1]), pnx::pack_bf16x2(f[2], f[3]), pnx::pack_bf16x2(f[4], f[5]),
                    pnx::pack_bf16x2(f[6], f[7]));
}

// Thread layout of the row-wise kernels: 256 threads = (256 / cg) rows x cg channel groups of 8 channels; a thread
// keeps its channel group for the whole grid-stride loop, so per-channel coefficients live in registers and the
// loop has no integer division.
template <int U>
__global__ void __launch_bounds__(256)
    bn_apply_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, long long M, int C,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    const __nv_bfloat16* __restrict__ res, long long ldr, int relu,
                    __nv_bfloat16* __restrict__ y, long long ldy) {
  const int cg = C >> 3;
  const int rpb = 256 / cg;
  const int my_cg = threadIdx.x % cg, my_row = threadIdx.x / cg;
  if (my_row >= rpb) return;
  const int c0 = my_cg << 3;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = scale[c0 + k];
    sh[k] = shift[c0 + k];
  }
  // U rows per iteration: U independent 16-byte loads in flight per operand (memory-level parallelism)
  const long long step = rpb;  // the CTA sweeps 4*rpb contiguous rows per iteration
  for (long long m0 = (long long)blockIdx.x * rpb * U + my_row; m0 < M; m0 += (long long)gridDim.x * rpb * U) {
    uint4 xin[U], rin[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = m0 + u * step;
      if (m < M) {
        xin[u] = *reinterpret_cast<const uint4*>(x + m * ldx + c0);
        if (res) rin[u] = *reinterpret_cast<const uint4*>(res + m * ldr + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = m0 + u * step;
      if (m >= M) break;
      float v[8], r[8];
      unpack8(xin[u], v);
      if (res) unpack8(rin[u], r);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float o = fmaf(v[k], sc[k], sh[k]);
        if (res) o += r[k];
        if (relu) o = fmaxf(o, 0.f);
        v[k] = o;
      }
      *reinterpret_cast<uint4*>(y + m * ldy + c0) = pack8(v);
    }
  }
}

// g = dy * (y > 0 if relu);  red[0:C] += sum g ; red[C:2C] += sum g * xhat,  xhat = (x - mean) * invstd
// Register diet (round 2, ncu: 128 registers -> 24 % occupancy -> 4.7 TB/s): the threads accumulate the RAW moments
// sum g and sum g*x; mean / invstd enter once per block in the fp64 tail (sum g*xhat = invstd * (sum g*x - mean * sum g)),
// and the affine of the recomputed ReLU mask is only resident in the kMask == 2 instantiation.
// kMask: 0 no ReLU | 1 mask from the stored output y | 2 mask recomputed from x (y = x*fscale + fshift)
template <int kThreads, int U, int kMask>
__global__ void __launch_bounds__(kThreads, 2048 / kThreads >= 3 ? 3 : 1)
    bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                         const __nv_bfloat16* __restrict__ y, long long ldy,
                         const __nv_bfloat16* __restrict__ x, long long ldx, long long M, int C,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ fscale, const float* __restrict__ fshift,
                         double* __restrict__ red) {
  extern __shared__ float sred[];  // [kThreads][16]
  const int cg = C >> 3;
  const int my_cg = threadIdx.x % cg;
  const int rows_per_block = kThreads / cg;
  const int my_row = threadIdx.x / cg;
  const int c0 = my_cg << 3;
  float sg[8], sgx[8], fs[8], fh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sg[k] = sgx[k] = 0.f;
    if (kMask == 2) {
      fs[k] = fscale[c0 + k];
      fh[k] = fshift[c0 + k];
    }
  }
  if (my_row < rows_per_block) {
    const long long step = rows_per_block;
    for (long long m0 = (long long)blockIdx.x * rows_per_block * U + my_row; m0 < M;
         m0 += (long long)gridDim.x * rows_per_block * U) {
      uint4 gin[U], xin[U], yin[kMask == 1 ? U : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long m = m0 + u * step;
        if (m < M) {
          gin[u] = *reinterpret_cast<const uint4*>(dy + m * lddy + c0);
          xin[u] = *reinterpret_cast<const uint4*>(x + m * ldx + c0);
          if (kMask == 1) yin[u] = *reinterpret_cast<const uint4*>(y + m * ldy + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (m0 + u * step >= M) break;
        float g[8], yy[8], xx[8];
        unpack8(gin[u], g);
        unpack8(xin[u], xx);
        if (kMask == 1) unpack8(yin[u], yy);
        if (kMask == 2) {
#pragma unroll
          for (int k = 0; k < 8; ++k) yy[k] = fmaf(xx[k], fs[k], fh[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float gg = (kMask != 0 && !(yy[k] > 0.f)) ? 0.f : g[k];
          sg[k] += gg;
          sgx[k] = fmaf(gg, xx[k], sgx[k]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sred[threadIdx.x * 16 + k] = sg[k];
    sred[threadIdx.x * 16 + 8 + k] = sgx[k];
  }
  __syncthreads();
  // thread c < C reduces both moments of one channel over the block's rows (fp64) and centres the second one
  for (int c = threadIdx.x; c < C; c += kThreads) {
    const int g_ = c >> 3, k = c & 7;
    double ag = 0.0, agx = 0.0;
    for (int r = 0; r < rows_per_block; ++r) {
      ag += (double)sred[(r * cg + g_) * 16 + k];
      agx += (double)sred[(r * cg + g_) * 16 + 8 + k];
    }
    atomicAdd(&red[c], ag);
    atomicAdd(&red[C + c], (double)invstd[c] * (agx - (double)mean[c] * ag));
  }
}

// dx = gamma*invstd * (g - sum_g/n - xhat * sum_gx/n) ; optional dres (+)= g        (same thread layout as bn_apply)
// Folded to three per-channel coefficients, dx = A*g + B*x + D with A = gamma*invstd, B = -A*invstd*sum_gx/n,
// D = -A*sum_g/n - B*mean (round 2: 125 registers / 24 % occupancy / 3.7 TB/s before).  kMask as above.
template <int U, int kMask>
__global__ void __launch_bounds__(256, 3)
    bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ y,
                        long long ldy, const __nv_bfloat16* __restrict__ x, long long ldx, long long M, int C,
                        const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ gamma, const double* __restrict__ red, float inv_n,
                        const int* __restrict__ count_dev,
                        const float* __restrict__ fscale, const float* __restrict__ fshift,
                        __nv_bfloat16* __restrict__ dx, long long lddx, __nv_bfloat16* __restrict__ dres,
                        long long lddres, int dres_accumulate) {
  const int cg = C >> 3;
  const int rpb = 256 / cg;
  const int my_cg = threadIdx.x % cg, my_row = threadIdx.x / cg;
  if (my_row >= rpb) return;
  const int c0 = my_cg << 3;
  float cA[8], cB[8], cD[8], fs[8], fh[8];
  if (count_dev) inv_n = 1.f / (float)max(*count_dev, 1);   // SyncBatchNorm: population of all ranks, known on the device only
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + k;
    if (kMask == 2) {
      fs[k] = fscale[c];
      fh[k] = fshift[c];
    }
    const float is = invstd[c], mu = mean[c];
    const float a = gamma[c] * is;
    const float sgn = (float)red[c] * inv_n, sgxn = (float)red[C + c] * inv_n;
    cA[k] = a;
    cB[k] = -a * is * sgxn;
    cD[k] = -a * sgn - cB[k] * mu;
  }
  const long long step = rpb;
  for (long long mm = (long long)blockIdx.x * rpb * U + my_row; mm < M; mm += (long long)gridDim.x * rpb * U) {
    uint4 gin[U], xin[U], yin[kMask == 1 ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = mm + u * step;
      if (m < M) {
        gin[u] = *reinterpret_cast<const uint4*>(dy + m * lddy + c0);
        xin[u] = *reinterpret_cast<const uint4*>(x + m * ldx + c0);
        if (kMask == 1) yin[u] = *reinterpret_cast<const uint4*>(y + m * ldy + c0);
      }
    }
#pragma unroll
   for (int u = 0; u < U; ++u) {
    const long long m = mm + u * step;
    if (m >= M) break;
    float g[8], yy[8], xx[8], o[8];
    unpack8(gin[u], g);
    unpack8(xin[u], xx);
    if (kMask == 1) unpack8(yin[u], yy);
    if (kMask == 2) {
#pragma unroll
      for (int k = 0; k < 8; ++k) yy[k] = fmaf(xx[k], fs[k], fh[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float gg = (kMask != 0 && !(yy[k] > 0.f)) ? 0.f : g[k];
      g[k] = gg;
      o[k] = fmaf(cA[k], gg, fmaf(cB[k], xx[k], cD[k]));
    }
    *reinterpret_cast<uint4*>(dx + m * lddx + c0) = pack8(o);
    if (dres) {
      if (dres_accumulate) {
        float a[8];
        unpack8(*reinterpret_cast<const uint4*>(dres + m * lddres + c0), a);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] += a[k];
      }
      *reinterpret_cast<uint4*>(dres + m * lddres + c0) = pack8(g);
    }
   }
  }
}

// a[m, 0:C] += b[m, 0:C]  (bf16)
__global__ void add_rows_kernel(__nv_bfloat16* __restrict__ a, long long lda, const __nv_bfloat16* __restrict__ b,
                                long long ldb, long long M, int C) {
  const int cg = C >> 3;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = M * cg;
  for (; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / cg;
    const int c0 = (int)(t - m * cg) << 3;
    float x[8], y[8];
    unpack8(*reinterpret_cast<const uint4*>(a + m * lda + c0), x);
    unpack8(*reinterpret_cast<const uint4*>(b + m * ldb + c0), y);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] += y[k];
    *reinterpret_cast<uint4*>(a + m * lda + c0) = pack8(x);
  }
}

// y = relu(a + b)   (BasicBlock tail, reference conv.py:48-50)
__global__ void add_relu_kernel(const __nv_bfloat16* __restrict__ a, long long lda, const __nv_bfloat16* __restrict__ b,
                                long long ldb, long long M, int C, __nv_bfloat16* __restrict__ y, long long ldy) {
  const int cg = C >> 3;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = M * cg;
  for (; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / cg;
    const int c0 = (int)(t - m * cg) << 3;
    float x[8], z[8];
    unpack8(*reinterpret_cast<const uint4*>(a + m * lda + c0), x);
    unpack8(*reinterpret_cast<const uint4*>(b + m * ldb + c0), z);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = fmaxf(x[k] + z[k], 0.f);
    *reinterpret_cast<uint4*>(y + m * ldy + c0) = pack8(x);
  }
}
// g (+)= dy * (y > 0)
__global__ void relu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ y,
                                long long ldy, long long M, int C, __nv_bfloat16* __restrict__ g, long long ldg,
                                int accumulate) {
  const int cg = C >> 3;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = M * cg;
  for (; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / cg;
    const int c0 = (int)(t - m * cg) << 3;
    float d[8], yy[8], acc[8];
    unpack8(*reinterpret_cast<const uint4*>(dy + m * lddy + c0), d);
    unpack8(*reinterpret_cast<const uint4*>(y + m * ldy + c0), yy);
    if (accumulate) unpack8(*reinterpret_cast<const uint4*>(g + m * ldg + c0), acc);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = yy[k] > 0.f ? d[k] : 0.f;
      d[k] = accumulate ? v + acc[k] : v;
    }
    *reinterpret_cast<uint4*>(g + m * ldg + c0) = pack8(d);
  }
}

// grid for the (rows x channel-group) layout: enough blocks for ~16 waves, each thread streaming >= 4 rows
inline int tune_env(const char* name, int dflt) {   // developer knob (tools/bench_ew.py sweeps); unset in production
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
inline int row_blocks(long long M, int C, int waves) {
  const int rpb = 256 / (C / 8);
  long long b = (M + (long long)rpb * 4 - 1) / ((long long)rpb * 4);
  const long long cap = 148LL * waves;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
#define PNX_DISPATCH_U(u, ...)            \
  switch (u) {                            \
    case 1: { constexpr int U = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int U = 2; __VA_ARGS__; } break; \
    default: { constexpr int U = 4; __VA_ARGS__; } break; \
  }

// kMask dispatch of the BatchNorm backward kernels (0 none | 1 stored y | 2 recomputed from x)
#define PNX_DISPATCH_MASK(mask, ...)      \
  switch (mask) {                         \
    case 0: { constexpr int K = 0; __VA_ARGS__; } break; \
    case 1: { constexpr int K = 1; __VA_ARGS__; } break; \
    default: { constexpr int K = 2; __VA_ARGS__; } break; \
  }
inline int ew_blocks(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  long long cap = 148LL * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

extern "C" int pnx_bn_apply(const void* x, long long ldx, long long M, int C, const float* scale, const float* shift,
                            const void* res, long long ldr, int relu, void* y, long long ldy, cudaStream_t stream) {
  PNX_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldr % 8 == 0, "C/ld % 8");
  if (M == 0) return PNX_OK;
  PNX_CHECK_ARG(C <= 2048, "C <= 2048");
  constexpr int u = 1;  // rows in flight per thread; measured (tools/bench_ew.py): 6.2 TB/s at U=1, slower unrolled
  PNX_DISPATCH_U(u, bn_apply_kernel<U><<<row_blocks(M, C, 16), 256, 0, stream>>>(
                        (const __nv_bfloat16*)x, ldx, M, C, scale, shift, (const __nv_bfloat16*)res, ldr, relu,
                        (__nv_bfloat16*)y, ldy));
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

extern "C" int pnx_bn_bwd_reduce(const void* dy, long long lddy, const void* y, long long ldy, const void* x,
                                 long long ldx, long long M, int C, const float* mean, const float* invstd, int relu,
                                 const float* fscale, const float* fshift, double* red, cudaStream_t stream) {
  PNX_CHECK_ARG(!relu || y || (fscale && fshift), "relu backward needs y or the forward affine (scale, shift)");
  PNX_CHECK_ARG(C % 8 == 0 && C <= 2048, "C % 8 == 0 and C <= 2048");
  if (M == 0) return PNX_OK;
  constexpr int kT = 256;
  PNX_CHECK_ARG(C / 8 <= kT, "C <= 2048");
  const int rows_per_block = kT / (C / 8);
  long long nb = (M + rows_per_block * 8 - 1) / (rows_per_block * 8);
  static const int rw = tune_env("PNX_BN_BWD_REDUCE_WAVES", 3);
  if (nb > 148 * rw) nb = 148 * rw;
  if (nb < 1) nb = 1;
  static const int u = tune_env("PNX_BN_BWD_REDUCE_U", 4);
  const int mask = !relu ? 0 : (y ? 1 : 2);
  PNX_DISPATCH_MASK(mask, PNX_DISPATCH_U(u, bn_bwd_reduce_kernel<kT, U, K><<<(int)nb, kT, kT * 16 * sizeof(float), stream>>>(
                        (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)y, ldy, (const __nv_bfloat16*)x, ldx, M,
                        C, mean, invstd, fscale, fshift, red)));
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

extern "C" int pnx_bn_bwd_apply(const void* dy, long long lddy, const void* y, long long ldy, const void* x,
                                long long ldx, long long M, int C, const float* mean, const float* invstd,
                                const float* gamma, const double* red, double count, const int* count_dev, int relu,
                                const float* fscale, const float* fshift, void* dx, long long lddx, void* dres,
                                long long lddres, int dres_accumulate, cudaStream_t stream) {
  PNX_CHECK_ARG(!relu || y || (fscale && fshift), "relu backward needs y or the forward affine (scale, shift)");
  PNX_CHECK_ARG(C % 8 == 0, "C % 8");
  if (M == 0) return PNX_OK;
  PNX_CHECK_ARG(C <= 2048, "C <= 2048");
  static const int u = tune_env("PNX_BN_BWD_APPLY_U", 2);
  // 3 resident CTAs per SM (80 registers): one full wave of long-lived CTAs beats 16 waves of short ones here, the
  // per-CTA coefficient set-up (6 per-channel vectors, two of them fp64) is not free (3.49 -> 3.01 ms per step)
  static const int waves = tune_env("PNX_BN_BWD_APPLY_WAVES", 3);
  const int mask = !relu ? 0 : (y ? 1 : 2);
  PNX_DISPATCH_MASK(mask, PNX_DISPATCH_U(u, bn_bwd_apply_kernel<U, K><<<row_blocks(M, C, waves), 256, 0, stream>>>(
                        (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)y, ldy, (const __nv_bfloat16*)x, ldx, M,
                        C, mean, invstd, gamma, red, (float)(1.0 / count), count_dev, fscale, fshift, (__nv_bfloat16*)dx,
                        lddx, (__nv_bfloat16*)dres, lddres, dres_accumulate)));
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

extern "C" int pnx_add_rows(void* a, long long lda, const void* b, long long ldb, long long M, int C,
                            cudaStream_t stream) {
  PNX_CHECK_ARG(C % 8 == 0, "C % 8");
  if (M == 0) return PNX_OK;
  add_rows_kernel<<<ew_blocks(M * (C / 8), 256), 256, 0, stream>>>((__nv_bfloat16*)a, lda, (const __nv_bfloat16*)b,
                                                                   ldb, M, C);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

extern "C" int pnx_add_relu(const void* a, long long lda, const void* b, long long ldb, long long M, int C, void* y,
                            long long ldy, cudaStream_t stream) {
  PNX_CHECK_ARG(C % 8 == 0, "C % 8");
  if (M == 0) return PNX_OK;
  add_relu_kernel<<<ew_blocks(M * (C / 8), 256), 256, 0, stream>>>((const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)b,
                                                                   ldb, M, C, (__nv_bfloat16*)y, ldy);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

extern "C" int pnx_relu_bwd(const void* dy, long long lddy, const void* y, long long ldy, long long M, int C, void* g,
                            long long ldg, int accumulate, cudaStream_t stream) {
  PNX_CHECK_ARG(C % 8 == 0, "C % 8");
  if (M == 0) return PNX_OK;
  relu_bwd_kernel<<<ew_blocks(M * (C / 8), 256), 256, 0, stream>>>((const __nv_bfloat16*)dy, lddy,
                                                                   (const __nv_bfloat16*)y, ldy, M, C,
                                                                   (__nv_bfloat16*)g, ldg, accumulate);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}
