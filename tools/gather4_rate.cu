// Scratch microbenchmark: TMA tile::gather4 (4 arbitrary rows of a 2-D tensor per request) -- layout check and
// per-SM / chip throughput, to decide whether the row gathers of the sparse convolutions can move from
// cp.async producer warps to TMA.
#include "../pillarnext_b200/csrc/pnx_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
void pnx_set_error(const char*, ...) {}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tma_gather4(const CUtensorMap* m, uint64_t* bar, uint32_t dst, int c0, int r0, int r1,
                                            int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(pnx::smem_u32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

__global__ void check_kernel(const __grid_constant__ CUtensorMap tm, const int* idx, uint16_t* out, int tx_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < 128 * 128 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0xdeadbeefu;
  if (threadIdx.x == 0) { pnx::mbar_init(&bar, 1); pnx::fence_barrier_init(); }
  __syncthreads();
  pnx::fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    const int l = threadIdx.x;
    if (l == 0) pnx::mbar_arrive_expect_tx(&bar, tx_bytes);
    __syncwarp();
    tma_gather4(&tm, &bar, pnx::smem_u32(smem) + l * 512, 0, idx[4 * l], idx[4 * l + 1], idx[4 * l + 2], idx[4 * l + 3]);
  }
  pnx::mbar_wait(&bar, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) out[i] = ((uint16_t*)smem)[i];
}

// W warps issue gather4 requests (each lane one request of 4 rows x 128 B) into per-warp rings of S tiles of 128 rows
template <int W, int S>
__global__ void __launch_bounds__(W * 32, 1) rate_kernel(const __grid_constant__ CUtensorMap tm, int rows, int iters,
                                                         long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[W * S];
  if (threadIdx.x == 0) {
    for (int i = 0; i < W * S; ++i) pnx::mbar_init(&bar[i], 1);
    pnx::fence_barrier_init();
  }
  __syncthreads();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  uint32_t seed = (blockIdx.x * 977 + threadIdx.x) * 2654435761u + 12345u;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int s = it % S;
    if (it >= S) pnx::mbar_wait(&bar[w * S + s], ((it / S) - 1) & 1);
    if (l == 0) pnx::mbar_arrive_expect_tx(&bar[w * S + s], 128 * 128);
    __syncwarp();
    int r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      seed = seed * 1664525u + 1013904223u;
      r[j] = (int)((seed >> 8) % (uint32_t)rows);
    }
    tma_gather4(&tm, &bar[w * S + s], pnx::smem_u32(smem) + (w * S + s) * 16384 + l * 512, 0, r[0], r[1], r[2], r[3]);
  }
  for (int it = iters; it < iters + S; ++it) {
    const int s = it % S;
    if (it >= S) pnx::mbar_wait(&bar[w * S + s], ((it / S) - 1) & 1);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int W, int S>
void run_rate(const CUtensorMap& tm, int rows, const char* note) {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int smem = W * S * 16384 + 2048;
  cudaFuncSetAttribute(rate_kernel<W, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4000;
  rate_kernel<W, S><<<148, W * 32, smem>>>(tm, rows, 200, d);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_kernel<W, S><<<148, W * 32, smem>>>(tm, rows, iters, d);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double bytes_sm = (double)W * iters * 16384;
  printf("gather4 W=%d S=%d %s: %.1f B/clk/SM (SM0), chip %.2f TB/s  err=%s\n", W, S, note, bytes_sm / (double)h[0],
         bytes_sm * 148 / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main(int argc, char** argv) {
  const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  const int R = 600000;  // 77 MB of 128-byte rows: mostly L2-resident like the 9-tap re-reads of the real layers
  std::vector<uint16_t> h((size_t)R * 64);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < 64; ++c) h[(size_t)r * 64 + c] = (uint16_t)((r * 7 + c) & 0xffff);
  uint16_t* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  CUtensorMap tm;
  cuuint64_t dims[2] = {64, (cuuint64_t)R}; cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
  CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box_rows=%d -> CUresult %d\n", box_rows, (int)cr);
  if (cr != CUDA_SUCCESS) return 1;
  // ---- layout / zero-fill check
  std::vector<int> idx(128);
  for (int i = 0; i < 128; ++i) idx[i] = (int)((i * 7919u + 13) % R);
  idx[5] = -1; idx[6] = R + 3; idx[127] = R - 1; idx[0] = 0;
  int* didx; cudaMalloc(&didx, 512); cudaMemcpy(didx, idx.data(), 512, cudaMemcpyHostToDevice);
  uint16_t* dout; cudaMalloc(&dout, 128 * 128);
  cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024);
  check_kernel<<<1, 128, 20 * 1024>>>(tm, didx, dout, 128 * 128);
  cudaError_t ce = cudaDeviceSynchronize();
  printf("check kernel: %s\n", cudaGetErrorString(ce));
  if (ce != cudaSuccess) return 1;
  std::vector<uint16_t> o(128 * 64); cudaMemcpy(o.data(), dout, 128 * 128, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 64; ++c) {
      const int chunk = c >> 3, pos = (chunk ^ (r & 7)) * 8 + (c & 7);
      const uint16_t got = o[r * 64 + pos];
      const bool oob = idx[r] < 0 || idx[r] >= R;
      const uint16_t want = oob ? 0 : (uint16_t)((idx[r] * 7 + c) & 0xffff);
      if (got != want) { if (bad < 6) printf("  mismatch row %d (idx %d) col %d: got %04x want %04x\n", r, idx[r], c, got, want); ++bad; }
    }
  printf("layout check: %d mismatches (rows at r*128 B, 16-B chunks XOR (r&7); OOB rows zero)\n", bad);
  run_rate<1, 4>(tm, R, "1 warp, 4 tiles in flight");
  run_rate<1, 8>(tm, R, "1 warp, 8 tiles in flight");
  run_rate<2, 4>(tm, R, "2 warps");
  run_rate<4, 3>(tm, R, "4 warps");
  run_rate<6, 2>(tm, R, "6 warps");
  run_rate<8, 1>(tm, R, "8 warps, 1 tile each");
  run_rate<12, 1>(tm, R, "12 warps, 1 tile each");
  run_rate<4, 3>(tm, 40000, "4 warps, 5 MB table");
  run_rate<8, 1>(tm, 40000, "8 warps, 5 MB table");
  return 0;
}
