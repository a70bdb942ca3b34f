// stencil.cu -- "pointwise then stencil" form of the head's final 3x3 convolutions (64 -> c, c <= 3 per head;
// reference det3d/models/heads/centerhead.py:44-46).  A 3x3 conv with very few output channels is gather-bound
// as an implicit GEMM (every 384-channel input row is read 9 times for 16 outputs).  It is algebraically
//     out[m, j] = bias[j] + sum_t Z[m + off_t, t*cpt + j],    Z = y . Wz^T   (one 1x1 GEMM, N = 9*cpt, no gather; cpt =
//     output channels per tap rounded up to 4: 12 for the reference's heads -> N = 108 padded to 128, instead of 9*16 -> 192)
// so the tensor-core GEMM reads y once and this kernel sums nine 16-float vectors per pixel (zero padding at the
// image border).  The backward is the mirrored gather dZ[m', t*16 + j] = dout[m' - off_t, j].
#include "pnx_common.cuh"

namespace {

// out [M,16] fp32 ; Z [M, ldz] fp32, column = tap * cpt + j (cpt = channels per tap, a multiple of 4, <= 16);
// one thread = one pixel x 4 output channels
__global__ void tap_gather_sum_kernel(const float* __restrict__ Z, long long ldz, int cpt, const float* __restrict__ bias, int B,
                                      int H, int W, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long M = (long long)B * H * W;
  const long long m = t >> 2;
  const int q = (int)(t & 3);
  if (m >= M) return;
  float4 acc = *reinterpret_cast<const float4*>(bias + q * 4);
  if (q * 4 < cpt) {
    const int hw = H * W;
    const int b = (int)(m / hw);
    const int rem = (int)(m - (long long)b * hw);
    const int y = rem / W, x = rem - y * W;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yi = y + r - 1;
      if (yi < 0 || yi >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int xi = x + s - 1;
        if (xi < 0 || xi >= W) continue;
        const long long src = ((long long)b * H + yi) * W + xi;
        const float4 z = __ldg(reinterpret_cast<const float4*>(Z + src * ldz + (r * 3 + s) * cpt + q * 4));
        acc.x += z.x; acc.y += z.y; acc.z += z.z; acc.w += z.w;
      }
    }
  }
  *reinterpret_cast<float4*>(out + m * 16 + q * 4) = acc;
}

// dZ [M, ldz] bf16 (columns >= 9*cpt zero) ; dout [M,16] fp32 ; one thread = one pixel x one 16-byte chunk of dZ (two
// groups of 4 columns; a group never straddles a tap because cpt % 4 == 0).
// lo_off > 0 (fp32-grade split mode): dZ rows hold `pieces` bf16 pieces per value, piece q at column q * lo_off + c
__global__ void __launch_bounds__(256) tap_scatter_kernel(const float* __restrict__ dout, int B, int H, int W,
                                                         __nv_bfloat16* __restrict__ dZ, long long ldz, int nz, int cpt,
                                                         long long lo_off, int pieces, int slots_log2) {
  // 256 threads = (256 >> slots_log2) pixels x 16 or 32 chunk slots (a row has nz / 8 = 16 or 24 chunks): pixel and chunk come from the thread index
  // without a division, the pixel's (b, y, x) from two 32-bit ones.  (The first version derived everything from one flat
  // 64-bit index: ~350 instructions per 16-byte store, instruction-bound at 1.45 TB/s.)
  const int chunks = nz >> 3;  // 8 bf16 per chunk
  const int c = threadIdx.x & ((1 << slots_log2) - 1);
  const long long m = (long long)blockIdx.x * (256 >> slots_log2) + (threadIdx.x >> slots_log2);
  if (c >= chunks || m >= (long long)B * H * W) return;
  const unsigned hw = (unsigned)(H * W);
  const unsigned b = (unsigned)(m / hw);            // m < 2^31 (checked by the launcher): 32-bit division
  const unsigned rem = (unsigned)m - b * hw;
  const int y = (int)(rem / (unsigned)W), x = (int)(rem - (unsigned)y * (unsigned)W);
  float v[8];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const unsigned col = (unsigned)(c * 8 + g * 4);
    const unsigned tap = col / (unsigned)cpt, j0 = col - tap * (unsigned)cpt;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tap < 9u) {
      const int r = (int)(tap / 3u), s_ = (int)tap - r * 3;
      const int yo = y - (r - 1), xo = x - (s_ - 1);  // the output pixel that read this Z entry through tap (r,s)
      if (yo >= 0 && yo < H && xo >= 0 && xo < W)
        a = __ldg(reinterpret_cast<const float4*>(dout + ((long long)(b * (unsigned)H + (unsigned)yo) * W + xo) * 16 + j0));
    }
    v[4 * g] = a.x; v[4 * g + 1] = a.y; v[4 * g + 2] = a.z; v[4 * g + 3] = a.w;
  }
  const int np = lo_off > 0 ? pieces : 1;
  for (int q = 0; q < np; ++q) {
    float pc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      pc[k] = pnx::bf16_round(v[k]);
      v[k] -= pc[k];
    }
    *reinterpret_cast<uint4*>(dZ + m * ldz + q * lo_off + c * 8) =
        make_uint4(pnx::pack_bf16x2(pc[0], pc[1]), pnx::pack_bf16x2(pc[2], pc[3]), pnx::pack_bf16x2(pc[4], pc[5]), pnx::pack_bf16x2(pc[6], pc[7]));
  }
}

}  // namespace

extern "C" int pnx_tap_gather_sum(const float* Z, long long ldz, int cpt, const float* bias16, int B, int H, int W, float* out,
                                  cudaStream_t stream) {
  PNX_CHECK_ARG(cpt >= 4 && cpt <= 16 && cpt % 4 == 0, "cpt in {4, 8, 12, 16}");
  PNX_CHECK_ARG(ldz >= 9 * cpt && ldz % 4 == 0, "ldz");
  const long long threads = (long long)B * H * W * 4;
  if (threads == 0) return PNX_OK;
  tap_gather_sum_kernel<<<pnx_cdiv(threads, 256), 256, 0, stream>>>(Z, ldz, cpt, bias16, B, H, W, out);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

extern "C" int pnx_split_get_pieces(void);

extern "C" int pnx_tap_scatter(const float* dout, int B, int H, int W, void* dZ, long long ldz, int nz, int cpt, long long lo_off,
                               cudaStream_t stream) {
  PNX_CHECK_ARG(cpt >= 4 && cpt <= 16 && cpt % 4 == 0, "cpt in {4, 8, 12, 16}");
  PNX_CHECK_ARG(nz >= 9 * cpt && nz % 8 == 0 && ldz >= nz && ldz % 8 == 0, "nz/ldz");
  const int pieces = lo_off > 0 ? pnx_split_get_pieces() : 1;
  PNX_CHECK_ARG(lo_off == 0 || (lo_off >= nz && (pieces - 1) * lo_off + nz <= ldz && lo_off % 8 == 0), "lo_off");
  const long long M = (long long)B * H * W;
  if (M == 0) return PNX_OK;
  PNX_CHECK_ARG(nz / 8 <= 32 && M < 2147483647LL, "nz <= 256 and fewer than 2^31 pixels");
  const int slots_log2 = nz / 8 <= 16 ? 4 : 5;
  tap_scatter_kernel<<<pnx_cdiv(M, 256 >> slots_log2), 256, 0, stream>>>(dout, B, H, W, (__nv_bfloat16*)dZ, ldz, nz, cpt, lo_off, pieces,
                                                                        slots_log2);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}
