// rulebook.cu -- active-site sets and neighbour tables for the sparse BEV backbone (rows B1-B4).
//
// Replaces spconv's hash-table index-pair generation (called from reference
// det3d/models/utils/sparse_conv.py:25-29,50-51 and backbones/sparse_resnet.py:43-48,63-64; the
// reference rebuilds pairs for all 21 layers because no indice_key is passed).  Here the active set
// of every level is an occupancy bitmap in (b, u=x, v=y) order, a site's row index is the rank of its
// bit (popcount prefix), and a neighbour table is built ONCE per level and shared by every conv on it.
//   SparseConv2d (k3,p1,stride s): out set = { (u',v') : any in (u's+du-1, v's+dv-1) }  -> bitmap dilation
//   SubMConv2d: out set = in set.
#include "pnx_common.cuh"

namespace {

struct Level {
  int batch, U, V, vwords;
};

// rank of site (b,u,v) = blockpref[word/32] + inblk[word] + popc(bits below)   (hierarchical prefix: the dense
// per-word array is 16-bit and written by the same warp pass that produces the bitmap)
__device__ __forceinline__ int site_lookup(const uint32_t* __restrict__ bm, const int* __restrict__ blockpref,
                                           const uint16_t* __restrict__ inblk, Level L, int b, int u, int v) {
  if (u < 0 || u >= L.U || v < 0 || v >= L.V) return -1;
  int w = (b * L.U + u) * L.vwords + (v >> 5);
  uint32_t bits = bm[w];
  uint32_t bit = 1u << (v & 31);
  if (!(bits & bit)) return -1;
  return blockpref[w >> 5] + (int)inblk[w] + __popc(bits & (bit - 1u));
}

__device__ __forceinline__ int warp_incl_scan_i(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if ((int)(threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

__device__ __forceinline__ uint32_t compress_even(uint32_t x) {
  x &= 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0F0F0F0Fu;
  x = (x | (x >> 4)) & 0x00FF00FFu;
  x = (x | (x >> 8)) & 0x0000FFFFu;
  return x;
}

// one thread per OUTPUT bitmap word
__global__ void dilate_kernel(const uint32_t* __restrict__ in, Level Li, int stride, uint32_t* __restrict__ out,
                              Level Lo, int n_words_pad, uint16_t* __restrict__ inblk, int* __restrict__ blockcnt,
                              int* __restrict__ supercnt) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  int n_words = Lo.batch * Lo.U * Lo.vwords;
  uint32_t acc = 0;
  int row = w / Lo.vwords, j = w - row * Lo.vwords;
  int b = row / Lo.U, uo = row - b * Lo.U;
  if (w < n_words)
  for (int du = -1; du <= 1; ++du) {
    int ui = uo * stride + du;
    if (ui < 0 || ui >= Li.U) continue;
    const uint32_t* r = in + (size_t)(b * Li.U + ui) * Li.vwords;
    if (stride == 1) {
      uint32_t c = r[j];
      uint32_t p = j > 0 ? r[j - 1] : 0u;
      uint32_t nx = j + 1 < Li.vwords ? r[j + 1] : 0u;
      acc |= c | (c << 1) | (p >> 31) | (c >> 1) | (nx << 31);
    } else {
      uint32_t lo = 2 * j < Li.vwords ? r[2 * j] : 0u;
      uint32_t hi = 2 * j + 1 < Li.vwords ? r[2 * j + 1] : 0u;
      uint32_t p = (2 * j > 0 && 2 * j - 1 < Li.vwords) ? r[2 * j - 1] : 0u;
      uint32_t tlo = lo | ((lo >> 1) | (hi << 31)) | ((lo << 1) | (p >> 31));
      uint32_t thi = hi | (hi >> 1) | ((hi << 1) | (lo >> 31));
      acc |= compress_even(tlo) | (compress_even(thi) << 16);
    }
  }
  int v0 = j * 32;
  int valid = Lo.V - v0;
  if (valid < 32) acc &= (valid <= 0) ? 0u : ((1u << valid) - 1u);
  if (w >= n_words) acc = 0u;
  // the warp owns one 32-word block: in-block exclusive prefix + block count (no atomics)
  const int cnt = __popc(acc);
  const int incl = warp_incl_scan_i(cnt);
  if (w < n_words_pad) {
    out[w] = acc;
    inblk[w] = (uint16_t)(incl - cnt);
    if ((threadIdx.x & 31) == 31) {
      blockcnt[w >> 5] = incl;
      if (incl) atomicAdd(&supercnt[w >> 15], incl);
    }
  }
}

// in-block prefix of an existing bitmap (level 0, produced by the voxelizer)
__global__ void inblock_kernel(const uint32_t* __restrict__ bm, int n_words_pad, uint16_t* __restrict__ inblk) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int cnt = w < n_words_pad ? __popc(bm[w]) : 0;
  const int incl = warp_incl_scan_i(cnt);
  if (w < n_words_pad) inblk[w] = (uint16_t)(incl - cnt);
}

__global__ void site_coords_kernel(const uint32_t* __restrict__ bm, const int* __restrict__ blockpref,
                                   const uint16_t* __restrict__ inblk, Level L, int* __restrict__ coords, int cap) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= L.batch * L.U * L.vwords) return;
  uint32_t bits = bm[w];
  if (!bits) return;
  int idx = blockpref[w >> 5] + (int)inblk[w];
  int row = w / L.vwords, vw = w - row * L.vwords;
  int b = row / L.U, u = row - b * L.U;
  while (bits) {
    int bit = __ffs(bits) - 1;
    bits &= bits - 1;
    if (idx < cap) {
      coords[idx * 3 + 0] = b;
      coords[idx * 3 + 1] = u;
      coords[idx * 3 + 2] = vw * 32 + bit;
    }
    ++idx;
  }
}

// nbr[i][t], t = ku*3+kv : row index (in the INPUT level) feeding output site i through tap (ku,kv)
//   forward  (transposed=0): in = (u*stride + ku - 1, v*stride + kv - 1)
//   backward (transposed=1): "output" is a site of the conv's INPUT level, "in" the conv's OUTPUT level:
//                            u' = (u + 1 - ku)/stride when divisible           (dgrad gather)
__global__ void nbr_table_kernel(const int* __restrict__ dst_coords, const int* __restrict__ n_dst_ptr, int cap,
                                 const uint32_t* __restrict__ src_bm, const int* __restrict__ src_prefix,
                                 const uint16_t* __restrict__ src_inblk, Level Ls, int stride, int transposed,
                                 int* __restrict__ nbr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = min(*n_dst_ptr, cap);
  if (i >= n) return;
  int b = dst_coords[i * 3], u = dst_coords[i * 3 + 1], v = dst_coords[i * 3 + 2];
#pragma unroll
  for (int ku = 0; ku < 3; ++ku) {
#pragma unroll
    for (int kv = 0; kv < 3; ++kv) {
      int r = -1;
      if (!transposed) {
        r = site_lookup(src_bm, src_prefix, src_inblk, Ls, b, u * stride + ku - 1, v * stride + kv - 1);
      } else {
        int nu = u + 1 - ku, nv = v + 1 - kv;
        if (nu >= 0 && nv >= 0 && (nu % stride) == 0 && (nv % stride) == 0)
          r = site_lookup(src_bm, src_prefix, src_inblk, Ls, b, nu / stride, nv / stride);
      }
      nbr[(size_t)i * 9 + ku * 3 + kv] = r;
    }
  }
}

// x.dense() (sparse_resnet.py:68) into a zero-filled channels-last canvas [B, H(=V), W(=U), C] (bf16)
__global__ void scatter_dense_kernel(const uint4* __restrict__ feat, const int* __restrict__ coords,
                                     const int* __restrict__ n_ptr, int cap, int c_vec, Level L,
                                     uint4* __restrict__ out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int n = min(*n_ptr, cap);
  long long i = t / c_vec;
  int c = (int)(t - i * c_vec);
  if (i >= n) return;
  int b = coords[i * 3], u = coords[i * 3 + 1], v = coords[i * 3 + 2];
  long long pix = ((long long)b * L.V + v) * L.U + u;
  out[pix * c_vec + c] = feat[i * c_vec + c];
}
__global__ void gather_dense_kernel(const uint4* __restrict__ canvas, const int* __restrict__ coords,
                                    const int* __restrict__ n_ptr, int cap, int c_vec, Level L,
                                    uint4* __restrict__ feat) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int n = min(*n_ptr, cap);
  long long i = t / c_vec;
  int c = (int)(t - i * c_vec);
  if (i >= n) return;
  int b = coords[i * 3], u = coords[i * 3 + 1], v = coords[i * 3 + 2];
  long long pix = ((long long)b * L.V + v) * L.U + u;
  feat[i * c_vec + c] = canvas[pix * c_vec + c];
}

}  // namespace

extern "C" int pnx_sites_out_dim(int in_dim, int stride) { return (in_dim - 1) / stride + 1; }

static inline int words_pad(int batch, int u, int v) {
  long long w = (long long)batch * u * ((v + 31) / 32);
  return (int)((w + 31) / 32 * 32);
}

// SparseConv2d(k3,p1,stride) output site set: bm_out [words_pad], inblk [words_pad] u16,
// blockcnt [pnx_blockcnt_size(words_pad/32)] (block + superblock counts, input of pnx_scan_blocks).
extern "C" int pnx_sites_dilate(const uint32_t* bm_in, int batch, int u_in, int v_in, int stride, uint32_t* bm_out,
                                uint16_t* inblk, int* blockcnt, cudaStream_t stream) {
  PNX_CHECK_ARG(stride == 1 || stride == 2, "stride must be 1 or 2");
  Level Li{batch, u_in, v_in, (v_in + 31) / 32};
  int uo = pnx_sites_out_dim(u_in, stride), vo = pnx_sites_out_dim(v_in, stride);
  Level Lo{batch, uo, vo, (vo + 31) / 32};
  int n_pad = words_pad(batch, uo, vo);
  const int n_blocks = n_pad / 32;
  int* supercnt = blockcnt + (n_blocks + 3) / 4 * 4;   // layout of pnx_blockcnt_size / pnx_scan_blocks
  PNX_CUDA(cudaMemsetAsync(supercnt, 0, (size_t)((n_blocks + 1023) / 1024 + 4) * 4, stream));
  dilate_kernel<<<pnx_cdiv(n_pad, 256), 256, 0, stream>>>(bm_in, Li, stride, bm_out, Lo, n_pad, inblk, blockcnt, supercnt);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

// in-block prefix (u16 per word) of a bitmap produced elsewhere (the voxelizer's level-0 bitmap)
extern "C" int pnx_sites_inblock(const uint32_t* bm, int n_words_pad, uint16_t* inblk, cudaStream_t stream) {
  PNX_CHECK_ARG(n_words_pad % 32 == 0, "bitmap must be padded to whole 32-word blocks");
  if (n_words_pad == 0) return PNX_OK;
  inblock_kernel<<<pnx_cdiv(n_words_pad, 256), 256, 0, stream>>>(bm, n_words_pad, inblk);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

extern "C" int pnx_sites_coords(const uint32_t* bm, const int* blockpref, const uint16_t* inblk, int batch, int u,
                                int v, int* coords, int cap, cudaStream_t stream) {
  Level L{batch, u, v, (v + 31) / 32};
  int n_words = batch * u * L.vwords;
  site_coords_kernel<<<pnx_cdiv(n_words, 256), 256, 0, stream>>>(bm, blockpref, inblk, L, coords, cap);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

extern "C" int pnx_nbr_table(const int* dst_coords, const int* n_dst_ptr, int cap, const uint32_t* src_bm,
                             const int* src_blockpref, const uint16_t* src_inblk, int batch, int src_u, int src_v,
                             int stride, int transposed, int* nbr, cudaStream_t stream) {
  PNX_CHECK_ARG(stride == 1 || stride == 2, "stride must be 1 or 2");
  if (cap == 0) return PNX_OK;
  Level Ls{batch, src_u, src_v, (src_v + 31) / 32};
  nbr_table_kernel<<<pnx_cdiv(cap, 128), 128, 0, stream>>>(dst_coords, n_dst_ptr, cap, src_bm, src_blockpref, src_inblk,
                                                          Ls, stride, transposed, nbr);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

// feat [n, C] bf16 (C % 8 == 0) <-> canvas [B, V, U, C] bf16.  The canvas must be zeroed by the caller
// (cudaMemsetAsync) before scatter.
extern "C" int pnx_scatter_dense(const void* feat, const int* coords, const int* n_ptr, int cap, int channels,
                                 int batch, int u, int v, void* canvas, int gather, cudaStream_t stream) {
  PNX_CHECK_ARG(channels % 8 == 0, "channels % 8");
  if (cap == 0) return PNX_OK;
  Level L{batch, u, v, (v + 31) / 32};
  int c_vec = channels / 8;
  long long threads = (long long)cap * c_vec;
  if (gather)
    gather_dense_kernel<<<pnx_cdiv(threads, 256), 256, 0, stream>>>((const uint4*)canvas, coords, n_ptr, cap, c_vec,
                                                                    L, (uint4*)feat);
  else
    scatter_dense_kernel<<<pnx_cdiv(threads, 256), 256, 0, stream>>>((const uint4*)feat, coords, n_ptr, cap, c_vec,
                                                                     L, (uint4*)canvas);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}
