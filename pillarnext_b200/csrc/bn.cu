// bn.cu -- batch-norm statistics finalisation shared by every BN layer on the path
// (nn.BatchNorm1d eps=1e-3/momentum=0.01 in the reader+backbone: reference pillar_encoder.py:33,
//  sparse_conv.py:31,52, sparse_resnet.py:46; nn.BatchNorm2d defaults in neck/head: conv.py:27,
//  centerhead.py:41,113).  Producers accumulate per-channel sum / sum-of-squares in fp64; this kernel
// turns them into the affine (scale, shift) the apply kernels use and updates the running statistics
// exactly as torch does (biased variance to normalise, unbiased into running_var).
#include "pnx_common.cuh"

namespace {

__global__ void bn_finalize_kernel(const double* __restrict__ stats, int C, const int* __restrict__ count_ptr,
                                   long long count_host, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double n = count_ptr ? (double)(*count_ptr) * (double)count_host : (double)count_host;
  double mean = 0.0, var = 0.0;
  if (n > 0) {
    mean = stats[c] / n;
    var = stats[C + c] / n - mean * mean;
    if (var < 0) var = 0;
  }
  float invstd = (float)(1.0 / sqrt(var + (double)eps));
  float g = gamma[c], b = beta[c];
  scale[c] = g * invstd;
  shift[c] = b - (float)mean * g * invstd;
  if (mean_out) mean_out[c] = (float)mean;
  if (invstd_out) invstd_out[c] = invstd;
  if (running_mean && n > 0) {
    double unb = n > 1 ? var * (n / (n - 1.0)) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

__global__ void bn_eval_affine_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float invstd = (float)(1.0 / sqrt((double)rv[c] + (double)eps));
  scale[c] = gamma[c] * invstd;
  shift[c] = beta[c] - rm[c] * gamma[c] * invstd;
}

}  // namespace

// stats: [2*C] doubles (sum, sumsq). Effective count = (*count_ptr if non-null else 1) * count_mult.
extern "C" int pnx_bn_finalize(const double* stats, int channels, const int* count_ptr, long long count_mult,
                               const float* gamma, const float* beta, float eps, float momentum,
                               float* running_mean, float* running_var, float* scale, float* shift,
                               float* mean_out, float* invstd_out, cudaStream_t stream) {
  PNX_CHECK_ARG(channels > 0, "channels");
  bn_finalize_kernel<<<pnx_cdiv(channels, 128), 128, 0, stream>>>(stats, channels, count_ptr, count_mult, gamma,
                                                                  beta, eps, momentum, running_mean, running_var,
                                                                  scale, shift, mean_out, invstd_out);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}

extern "C" int pnx_bn_eval_affine(int channels, const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, float eps, float* scale, float* shift,
                                  cudaStream_t stream) {
  bn_eval_affine_kernel<<<pnx_cdiv(channels, 128), 128, 0, stream>>>(channels, gamma, beta, running_mean,
                                                                     running_var, eps, scale, shift);
  PNX_CHECK_LAUNCH();
  return PNX_OK;
}
