// stencil.cu -- "pointwise then stencil" form of the head's final 3x3 convolutions (64 -> c, c <= 3 per head;
// reference det3d/models/heads/centerhead.py:44-46).  A 3x3 conv with very few output channels is gather-bound
// as an implicit GEMM (every 384-channel input row is read 9 times for 16 outputs).  It is algebraically
//     out[m, j] = bias[j] + sum_t Z[m + off_t, t*16 + j],     Z = y . Wz^T   (one 1x1 GEMM, N = 9*16, no gather)
// so the tensor-core GEMM reads y once and this kernel sums nine 16-float vectors per pixel (zero padding at the
// image border).  The backward is the mirrored gather dZ[m', t*16 + j] = dout[m' - off_t, j].
#include "pnx_common.cuh"

namespace {

// out [M,16] fp32 ; Z [M, ldz] fp32 ; one thread = one pixel x 4 output channels
__global__ void tap_gather_sum_kernel(const float* __restrict__ Z, long long ldz, const float* __restrict__ bias, int B,
                                      int H, int W, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long M = (long long)B * H * W;
  const long long m = t >> 2;
  const int q = (int)(t & 3);
  if (m >= M) return;
  const int hw = H * W;
  const int b = (int)(m / hw);
  const int rem = (int)(m - (long long)b * hw);
  const int y = rem / W, x = rem - y * W;
  float4 acc = *reinterpret_cast<const float4*>(bias + q * 4);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int yi = y + r - 1;
    if (yi < 0 || yi >= H) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int xi = x + s - 1;
      if (xi < 0 || xi >= W) continue;
      const long long src = ((long long)b * H + yi) * W + xi;
      const float4 z = __ldg(reinterpret_cast<const float4*>(Z + src * ldz + (r * 3 + s) * 16 + q * 4));
      acc.x += z.x; acc.y += z.y; acc.z += z.z; acc.w += z.w;
    }
  }
  *reinterpret_cast<float4*>(out + m * 16 + q * 4) = acc;
}

// dZ [M, ldz] bf16 (columns >= 144 zero) ; dout [M,16] fp32 ; one thread = one pixel x one 16-byte chunk of dZ
// lo_off > 0 (fp32-grade split mode): dZ rows are [hi(nz) | lo(nz)] pairs, lo = bf16(v - hi) at column lo_off + c
__global__ void tap_scatter_kernel(const float* __restrict__ dout, int B, int H, int W, __nv_bfloat16* __restrict__ dZ,
                                   long long ldz, int nz, long long lo_off, int pieces) {
  const int chunks = nz >> 3;  // 8 bf16 per chunk
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long M = (long long)B * H * W;
  const long long m = t / chunks;
  const int c = (int)(t - m * chunks);
  if (m >= M) return;
  uint4 o = make_uint4(0u, 0u, 0u, 0u), ol = make_uint4(0u, 0u, 0u, 0u), ol2 = make_uint4(0u, 0u, 0u, 0u);
  const int tap = c >> 1, half = c & 1;  // chunk c covers columns c*8 .. c*8+7 = tap (c/2), outputs half*8 ..
  if (tap < 9) {
    const int hw = H * W;
    const int b = (int)(m / hw);
    const int rem = (int)(m - (long long)b * hw);
    const int y = rem / W, x = rem - y * W;
    const int r = tap / 3, s = tap - r * 3;
    const int yo = y - (r - 1), xo = x - (s - 1);  // the output pixel that read this Z entry through tap (r,s)
    if (yo >= 0 && yo < H && xo >= 0 && xo < W) {
      const float* d = dout + (((long long)b * H + yo) * W + xo) * 16 + half * 8;
      const float4 a = __ldg(reinterpret_cast<const float4*>(d)), bq = __ldg(reinterpret_cast<const float4*>(d + 4));
      o = make_uint4(pnx::pack_bf16x2(a.x, a.y), pnx::pack_bf16x2(a.z, a.w), pnx::pack_bf16x2(bq.x, bq.y),
                     pnx::pack_bf16x2(bq.z, bq.w));
      if (lo_off > 0) {
        const float v[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
        float l[8], l2[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          l[k] = v[k] - pnx::bf16_round(v[k]);
          l2[k] = l[k] - pnx::bf16_round(l[k]);
        }
        ol = make_uint4(pnx::pack_bf16x2(l[0], l[1]), pnx::pack_bf16x2(l[2], l[3]), pnx::pack_bf16x2(l[4], l[5]),
                        pnx::pack_bf16x2(l[6], l[7]));
        ol2 = make_uint4(pnx::pack_bf16x2(l2[0], l2[1]), pnx::pack_bf16x2(l2[2], l2[3]), pnx::pack_bf16x2(l2[4], l2[5]),
                         pnx::pack_bf16x2(l2[6], l2[7]));
      }
    }
  }
  *reinterpret_cast<uint4*>(dZ + m * ldz + c * 8) = o;
  if (lo_off > 0) *reinterpret_cast<uint4*>(dZ + m * ldz + lo_off + c * 8) = ol;
  if (lo_off > 0 && pieces > 2) *reinterpret_cast<uint4*>(dZ + m * ldz + 2 * lo_off + c * 8) = ol2;
}

}  // namespace

extern "C" int pnx_tap_gather_sum(const float* Z, long long ldz, const float* bias16, int B, int H, int W, float* out,
                                  cudaStream_t stream) {
  PNX_CHECK_ARG(ldz >= 144 && ldz % 4 == 0, "ldz");
  const long long threads = (long long)B * H * W * 4;
  if (threads == 0) return PNX_OK;
  tap_gather_sum_kernel<<<pnx_cdiv(threads, 256), 256, 0, stream>>>(Z, ldz, bias16, B, H, W, out);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

extern "C" int pnx_split_get_pieces(void);

extern "C" int pnx_tap_scatter(const float* dout, int B, int H, int W, void* dZ, long long ldz, int nz, long long lo_off,
                               cudaStream_t stream) {
  PNX_CHECK_ARG(nz >= 144 && nz % 8 == 0 && ldz >= nz && ldz % 8 == 0, "nz/ldz");
  const int pieces = lo_off > 0 ? pnx_split_get_pieces() : 1;
  PNX_CHECK_ARG(lo_off == 0 || (lo_off >= nz && (pieces - 1) * lo_off + nz <= ldz && lo_off % 8 == 0), "lo_off");
  const long long threads = (long long)B * H * W * (nz / 8);
  if (threads == 0) return PNX_OK;
  tap_scatter_kernel<<<pnx_cdiv(threads, 256), 256, 0, stream>>>(dout, B, H, W, (__nv_bfloat16*)dZ, ldz, nz, lo_off, pieces);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}
