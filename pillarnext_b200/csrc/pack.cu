// pack.cu -- all convolution weights of the model repacked in ONE launch.
//
// The tensor-core kernels read bf16 weights in [tap, Cout, Cin] (forward) and [tap, Cin, Cout] (data gradient, taps
// flipped for 'same' / submanifold convolutions) order; the parameters stay fp32 in the reference's own layouts
// (nn.Conv2d [Cout, Cin, kh, kw], spconv [Cout, kH, kW, Cin], nn.ConvTranspose2d [Cin, Cout, kh, kw]; state-dict
// compatible).  After every optimizer step every packed copy is stale; instead of ~2 small torch kernels
// per (weight, layout) -- ~190 launches per step -- one kernel walks a table of 4-D gather-copy descriptors:
//     dst[((a*d1 + b)*d2 + c)*d3 + d] = bf16( src[base + a*s0 + b*s1 + c*s2 + d*s3] )       (strides may be negative: flip)
#include "pnx_common.cuh"

struct PnxPackDesc {       // 96 bytes, mirrored by functional.py (_pack_desc)
  const float* src;
  __nv_bfloat16* dst;
  long long start;         // first global element index of this entry (prefix over the table)
  long long base;
  long long stride[4];
  int dim[4];
};

namespace {

__global__ void __launch_bounds__(256) pack_weights_kernel(const PnxPackDesc* __restrict__ table, int n, long long total) {
  __shared__ long long s_start[257];
  for (int i = threadIdx.x; i <= n && i <= 256; i += blockDim.x) s_start[i] = i < n ? table[i].start : total;
  __syncthreads();
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n - 1;                 // entry with start <= g < next start
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_start[mid] <= g) lo = mid; else hi = mid - 1;
    }
    const PnxPackDesc& e = table[lo];
    long long r = g - e.start;
    const int d = (int)(r % e.dim[3]); r /= e.dim[3];
    const int c = (int)(r % e.dim[2]); r /= e.dim[2];
    const int b = (int)(r % e.dim[1]);
    const int a = (int)(r / e.dim[1]);
    e.dst[g - e.start] = __float2bfloat16_rn(e.src[e.base + a * e.stride[0] + b * e.stride[1] + c * e.stride[2] + d * e.stride[3]]);
  }
}

}  // namespace

// table: device array of n PnxPackDesc (n <= 256), total = sum of the entries' element counts.
extern "C" int pnx_pack_weights(const void* table, int n, long long total, cudaStream_t stream) {
  PNX_CHECK_ARG(n >= 0 && n <= 256 && total >= 0, "at most 256 entries per call");
  if (n == 0 || total == 0) return PNX_OK;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weights_kernel<<<(int)blocks, 256, 0, stream>>>((const PnxPackDesc*)table, n, total);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}
