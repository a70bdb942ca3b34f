// Scratch microbenchmark: tcgen05.mma issue rate for K-major vs MN-major smem operands (no loads in the loop).
#include "../pillarnext_b200/csrc/pnx_common.cuh"
#include <cstdio>
void pnx_set_error(const char*, ...) {}
template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int a_mn, int b_mn, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { pnx::mbar_init(&bar, 1); pnx::fence_barrier_init(); }
  if (threadIdx.x < 32) pnx::tmem_alloc<512>(&slot);
  pnx::tc_fence_before(); __syncthreads(); pnx::tc_fence_after();
  pnx::fence_proxy_async_smem();
  uint32_t tm = slot;
  if (threadIdx.x == 0) {
    uint32_t sa = pnx::smem_u32(smem), sb = sa + 32 * 1024;
    uint32_t idesc = pnx::make_idesc_bf16(128, N, a_mn, b_mn);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t da = a_mn ? pnx::make_smem_desc_sw128(sa + k * 2048, 8192, 1024) : pnx::make_smem_desc_sw128(sa + k * 32, 0, 1024);
        uint64_t db = b_mn ? pnx::make_smem_desc_sw128(sb + k * 2048, 8192, 1024) : pnx::make_smem_desc_sw128(sb + k * 32, 0, 1024);
        pnx::umma_f16(tm, da, db, idesc, 1);
      }
    }
    pnx::umma_commit(&bar);
    pnx::mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  pnx::tc_fence_before(); __syncthreads(); pnx::tc_fence_after();
  if (threadIdx.x < 32) pnx::tmem_dealloc<512>(tm);
}
template <int N> void run(const char* name) {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  for (int mode = 0; mode < 4; ++mode) {
    int a = mode & 1, b = mode >> 1;
    rate_kernel<N><<<148, 128, 70 * 1024>>>(a, b, 2000, d);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%s N=%d A_%s B_%s: %.1f cycles per MMA (ideal %d)  err=%s\n", name, N, a ? "MN" : "K", b ? "MN" : "K", (double)h / (2000 * 4), 128 * N / 256,
           cudaGetErrorString(cudaGetLastError()));
  }
}
int main() { run<256>("mma"); run<64>("mma"); run<128>("mma"); return 0; }
