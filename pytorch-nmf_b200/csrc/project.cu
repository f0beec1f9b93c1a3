// Hoyer's sparseness projection (reference: torchnmf/nmf.py:21-49 `_proj_func`): the closest non-negative vector with a
// given L1 norm k1 and squared L2 norm k2.  The reference runs it per column from Python, a data-dependent `while True`
// with one `.item()` per round (nmf.py:31-42) and seven ATen launches per round; here ONE launch projects every slice
// x[:, j, :] of a parameter viewed as (outer, D, inner), one block per slice, the loop on the device.
#include "common.cuh"

namespace nmfb200 {

namespace {

constexpr int kProjThreads = 1024;

// Block-wide sums of three doubles in a fixed order (warp shuffles, then warp 0 over the 32 warp results).
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double (*sh)[3]) {
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                       // the previous call's readers are done with sh
  if (lane == 0) { sh[warp][0] = a; sh[warp][1] = b; sh[warp][2] = c; }
  __syncthreads();
  if (warp == 0) {
    a = sh[lane][0]; b = sh[lane][1]; c = sh[lane][2];
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { sh[32][0] = a; sh[32][1] = b; sh[32][2] = c; }
  }
  __syncthreads();
  a = sh[32][0]; b = sh[32][1]; c = sh[32][2];
}

// x viewed as (outer, D, inner); block j owns the n = outer * inner elements of slice j.  `zeroed` (n bytes per slice) is
// nmf.py:30's zero_coef.  Element arithmetic is fp32 like the reference's tensors, the reductions accumulate in double.
__global__ void __launch_bounds__(kProjThreads)
hoyer_project_kernel(float* __restrict__ x, int64_t outer, int D, int64_t inner, const float* __restrict__ k1v,
                     const float* __restrict__ k2v, unsigned char* __restrict__ zeroed_ws, int max_rounds) {
  __shared__ double sh[33][3];
  const int j = blockIdx.x;
  const int64_t n = outer * inner;
  const double k1 = (double)k1v[j], k2 = (double)k2v[j];
  unsigned char* __restrict__ zeroed = zeroed_ws + (int64_t)j * n;
  auto at = [&](int64_t e) -> float* {
    const int64_t o = e / inner;
    return x + (o * D + j) * inner + (e - o * inner);
  };

  // nmf.py:28: v = s + (k1 - s.sum()) / N
  double s0 = 0.0, z0 = 0.0, z1 = 0.0;
  for (int64_t e = threadIdx.x; e < n; e += kProjThreads) s0 += (double)*at(e);
  block_sum3(s0, z0, z1, sh);
  float shift = (float)((k1 - s0) / (double)n);     // added (then clamped at 0 from round 1 on) at the top of each round
  bool clamp = false;
  int64_t nzero = 0;

  for (int round = 0; round < max_rounds; ++round) {
    // nmf.py:32-37 (with the pending `v += shift; relu` of :47-48 applied first)
    const float m = (float)(k1 / (double)(n - nzero));
    double a = 0.0, b = 0.0, c = 0.0;
    for (int64_t e = threadIdx.x; e < n; e += kProjThreads) {
      float* p = at(e);
      float v = *p + shift;
      if (clamp) v = fmaxf(v, 0.f);
      *p = v;
      if (round == 0) zeroed[e] = 0;
      const float w = (round > 0 && zeroed[e]) ? v : v - m;
      a += (double)w * w; b += (double)w * v; c += (double)v * v;
    }
    block_sum3(a, b, c, sh);
    const float af = (float)a, bf = (float)(2.0 * b), cf = (float)(c - k2);
    const float disc = fmaxf(bf * bf - 4.f * af * cf, 0.f);
    const float alphap = (-bf + sqrtf(disc)) * 0.5f / af;

    // nmf.py:38-46: v += alphap * w; negative entries join zero_coef and are clamped
    double anyneg = 0.0, total = 0.0, count = 0.0;
    for (int64_t e = threadIdx.x; e < n; e += kProjThreads) {
      float* p = at(e);
      const float v = *p;
      unsigned char z = zeroed[e];
      const float w = z ? v : v - m;
      float nv = fmaf(alphap, w, v);
      if (nv < 0.f) { anyneg = 1.0; if (!z) { z = 1; zeroed[e] = 1; } nv = 0.f; }
      *p = nv;
      total += (double)nv; count += (double)z;
    }
    block_sum3(anyneg, total, count, sh);
    if (anyneg == 0.0) break;                          // nmf.py:41-42
    nzero = (int64_t)count;
    if (nzero >= n) break;                             // nothing left to redistribute over (the reference would divide by 0)
    shift = (float)((k1 - total) / (double)(n - nzero));   // nmf.py:46
    clamp = true;                                      // nmf.py:47
  }
}

}  // namespace

int hoyer_project(float* x, int64_t outer, int D, int64_t inner, const float* k1, const float* k2,
                  unsigned char* zeroed_ws, cudaStream_t st) {
  if (D <= 0 || outer <= 0 || inner <= 0) return 0;
  const int64_t n = outer * inner;
  // every round that does not end the loop zeroes at least one more coordinate: n rounds is the hard bound
  const int max_rounds = (int)(n < 100000 ? n + 2 : 100002);
  hoyer_project_kernel<<<D, kProjThreads, 0, st>>>(x, outer, D, inner, k1, k2, zeroed_ws, max_rounds);
  NMF_LAUNCH_CHECK();
  return 0;
}

}  // namespace nmfb200
