// NMFD (1-D convolutive NMF) contractions as im2col-free sliding GEMMs on CUDA cores, fp32.
//
//   recon : WH[b,c,l]   = sum_{r,t} W[c,r,t] H[b,r,l-t]          nmf.py:776-779 (conv1d, flipped kernel, full pad)
//   wgrad : gW[c,r,t]   = sum_{b,l} G[b,c,l] H[b,r,l-t]          autograd of the above w.r.t. W
//   dgrad : gH[b,r,j]   = sum_{c,t} W[c,r,t] G[b,c,j+t]          autograd of the above w.r.t. H
//
// NMF2D / NMF3D (nmf.py:782-942: conv2d / conv3d with flipped kernels and full padding) are the same three contractions
// with multi-indices l = (x1, x2, l), t = (t1, t2, t): the LAST axis is the sliding axis of the kernels below, the outer
// axes are loops (recon, dgrad: over the outer shifts; wgrad: over the outer positions) around the same inner product.
//
// No Toeplitz/im2col matrix is ever formed: each CTA stages one contiguous window of the shifted
// operand in shared memory and every thread slides a 4-element register window across it (one shared
// load + one float4 load per 16 FMAs).
#include "common.cuh"

namespace nmfb200 {

namespace {

constexpr int kT = 64;     // output tile edge
constexpr int kLd = 68;    // shared pitch
constexpr int kTK = 32;    // k-chunk (shifts or samples) per stage

template <int MODE>
__device__ __forceinline__ void phi_d(float v, float s, float bm2, float bm1, float& pn, float& pp) {
  if (MODE == kKL) { pn = v / (s + kEps); pp = 0.f; }
  else if (MODE == kEU) { pn = v; pp = s; }
  else if (MODE == kIS) { float r = 1.0f / (s + kEps); pp = r; pn = (r * r) * v; }
  else { float x = s + kEps; pn = powf(x, bm2) * v; pp = powf(x, bm1); }
}

template <int MODE>
__device__ __forceinline__ float loss_term_d(float v, float s, float beta) {
  if (MODE == kKL) return v * (logf(v + kEps) - logf(s + kEps)) - v + s;
  if (MODE == kEU) { float d = s - v; return 0.5f * d * d; }
  if (MODE == kIS) { float te = v + kEps, xe = s + kEps; return te / xe - logf(te) + logf(xe) - 1.0f; }
  float x = s + kEps, t = beta < 0.f ? v + kEps : v, bm = beta - 1.0f;
  return (powf(t, beta) + bm * powf(x, beta) - beta * t * powf(x, bm)) / (beta * bm);
}

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  if (w == 0) {
    t = l < 8 ? sh[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;
}

// grid (ceil(L/64) * X1 * X2, ceil(C/64), B), 256 threads; thread (ty,tx) owns c = 4ty..4ty+3, l = 4tx..4tx+3.
template <int MODE, bool LOSS>
__global__ void __launch_bounds__(256)
nmfd_recon_kernel(NmfdShape s, const float* __restrict__ V, const float* __restrict__ W,
                  const float* __restrict__ H, float beta, float* __restrict__ Pn, float* __restrict__ Pp,
                  double* __restrict__ block_partials) {
  __shared__ __align__(16) float Ws[kTK * kLd];          // Ws[tt][c]
  __shared__ float Hs[kT + kTK];                         // window of H[b,r,:]
  __shared__ double red[8];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ntl = (s.L + kT - 1) / kT;
  const int outer = blockIdx.x / ntl;                    // (x1, x2): outer position of this tile of the target
  const int x1 = outer / s.X2, x2 = outer - x1 * s.X2;
  const int l0 = (blockIdx.x - outer * ntl) * kT, c0 = blockIdx.y * kT, b = blockIdx.z;
  const int J1 = s.J1(), J2 = s.J2();
  const int64_t WI = s.w_inner(), RT = (int64_t)s.R * WI;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nouter = s.T1 * s.T2;
  for (int ro = 0; ro < s.R * nouter; ++ro) {
    const int r = ro / nouter, to = ro - r * nouter;
    const int t1 = to / s.T2, t2 = to - t1 * s.T2;
    const int j1 = x1 - t1, j2 = x2 - t2;              // outer position in H (block-uniform)
    if (j1 < 0 || j1 >= J1 || j2 < 0 || j2 >= J2) continue;
    const float* Hrow = H + ((((int64_t)b * s.R + r) * J1 + j1) * J2 + j2) * s.Lin;
    const int64_t wofs = (int64_t)r * WI + (int64_t)to * s.T;
    for (int t0 = 0; t0 < s.T; t0 += kTK) {
      __syncthreads();
      for (int idx = tid; idx < kT * kTK; idx += 256) {
        int c = idx / kTK, tt = idx - c * kTK;
        float w = 0.f;
        if (c0 + c < s.C && t0 + tt < s.T) w = W[(int64_t)(c0 + c) * RT + wofs + t0 + tt];
        Ws[tt * kLd + c] = w;
      }
      for (int i = tid; i < kT + kTK - 1; i += 256) {
        int src = l0 - t0 - (kTK - 1) + i;
        Hs[i] = (src >= 0 && src < s.Lin) ? Hrow[src] : 0.f;
      }
      __syncthreads();
      // window index for (lj, tt) is lj - tt + kTK-1
      float h[4];
      h[1] = Hs[4 * tx + kTK];
      h[2] = Hs[4 * tx + kTK + 1];
      h[3] = Hs[4 * tx + kTK + 2];
#pragma unroll
      for (int tt = 0; tt < kTK; ++tt) {
        if (tt > 0) { h[3] = h[2]; h[2] = h[1]; h[1] = h[0]; }
        h[0] = Hs[4 * tx - tt + kTK - 1];
        float4 a = *reinterpret_cast<const float4*>(&Ws[tt * kLd + 4 * ty]);
        float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], h[j], acc[i][j]);
      }
    }
  }

  const float bm2 = beta - 2.0f, bm1 = beta - 1.0f;
  float local = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = c0 + 4 * ty + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int l = l0 + 4 * tx + j;
      if (c < s.C && l < s.L) {
        int64_t off = ((((int64_t)b * s.C + c) * s.X1 + x1) * s.X2 + x2) * s.L + l;
        float v = V[off];
        if (LOSS) {
          local += loss_term_d<MODE>(v, acc[i][j], beta);
        } else {
          float pn, pp;
          phi_d<MODE>(v, acc[i][j], bm2, bm1, pn, pp);
          Pn[off] = pn;
          if (MODE != kKL) Pp[off] = pp;
        }
      }
    }
  }
  if (LOSS) {
    double tot = block_sum_d((double)local, red);
    if (tid == 0)
      block_partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
  }
}

// grid (ceil(T/64) * R * T1 * T2, ceil(C/64)); thread (ty,tx): c = 4ty.., t = 4tx..
__global__ void __launch_bounds__(256)
nmfd_wgrad_kernel(NmfdShape s, const float* __restrict__ G, const float* __restrict__ H,
                  float* __restrict__ out) {
  __shared__ __align__(16) float Gs[kTK * kLd];          // Gs[ll][c]
  __shared__ float Hs[kT + kTK];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ntt = (s.T + kT - 1) / kT;
  const int zo = blockIdx.x / ntt;                        // (r, t1, t2)
  const int nouter = s.T1 * s.T2;
  const int r = zo / nouter, to = zo - r * nouter;
  const int t1 = to / s.T2, t2 = to - t1 * s.T2;
  const int t0 = (blockIdx.x - zo * ntt) * kT, c0 = blockIdx.y * kT;
  const int J1 = s.J1(), J2 = s.J2();
  const int64_t VI = s.v_inner();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int bo = 0; bo < s.B * J1 * J2; ++bo) {            // every H line (b, j1, j2) meets the G line (b, j1 + t1, j2 + t2)
    const int b = bo / (J1 * J2), jo = bo - b * (J1 * J2);
    const int j1 = jo / J2, j2 = jo - j1 * J2;
    const float* Hrow = H + ((((int64_t)b * s.R + r) * J1 + j1) * J2 + j2) * s.Lin;
    const float* Gb = G + (int64_t)b * s.C * VI + ((int64_t)(j1 + t1) * s.X2 + (j2 + t2)) * s.L;
    for (int l0 = 0; l0 < s.L; l0 += kTK) {
      __syncthreads();
      for (int idx = tid; idx < kT * kTK; idx += 256) {
        int c = idx / kTK, ll = idx - c * kTK;
        float g = 0.f;
        if (c0 + c < s.C && l0 + ll < s.L) g = Gb[(int64_t)(c0 + c) * VI + l0 + ll];
        Gs[ll * kLd + c] = g;
      }
      for (int i = tid; i < kT + kTK - 1; i += 256) {
        int src = l0 - t0 - (kT - 1) + i;       // window index for (ll, tj) is ll - tj + 63
        Hs[i] = (src >= 0 && src < s.Lin) ? Hrow[src] : 0.f;
      }
      __syncthreads();
      float h[4];
      h[0] = Hs[62 - 4 * tx];
      h[1] = Hs[61 - 4 * tx];
      h[2] = Hs[60 - 4 * tx];
#pragma unroll
      for (int ll = 0; ll < kTK; ++ll) {
        h[3] = h[2]; h[2] = h[1]; h[1] = h[0];
        h[0] = Hs[ll + 63 - 4 * tx];
        float4 a = *reinterpret_cast<const float4*>(&Gs[ll * kLd + 4 * ty]);
        float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], h[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = c0 + 4 * ty + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int t = t0 + 4 * tx + j;
      if (c < s.C && t < s.T) out[(((int64_t)c * s.R + r) * nouter + to) * s.T + t] = acc[i][j];
    }
  }
}

// grid (ceil(Lin/JTILE) * J1 * J2, nsplit, B); block = RG x JT threads (RG*JT = 256); thread: r = 4rg.., j = 4jx..
template <int RG>
__global__ void __launch_bounds__(256)
nmfd_dgrad_kernel(NmfdShape s, const float* __restrict__ G, const float* __restrict__ W,
                  float* __restrict__ out, int nsplit) {
  constexpr int JT = 256 / RG;
  constexpr int JTILE = 4 * JT;
  constexpr int Rp = 4 * RG;
  __shared__ __align__(16) float Ws[kTK * Rp];           // Ws[tt][r]
  __shared__ float Gs[JTILE + kTK];
  const int tid = threadIdx.x;
  const int rg = tid % RG, jx = tid / RG;
  const int ntj = (s.Lin + JTILE - 1) / JTILE;
  const int outer = blockIdx.x / ntj;                    // (j1, j2): outer position of this tile of H
  const int J2 = s.J2();
  const int j1 = outer / J2, j2 = outer - j1 * J2;
  const int j0 = (blockIdx.x - outer * ntj) * JTILE, split = blockIdx.y, b = blockIdx.z;
  const int cps = (s.C + nsplit - 1) / nsplit;
  const int cbeg = split * cps, cend = min(s.C, cbeg + cps);
  const int64_t WI = s.w_inner(), RT = (int64_t)s.R * WI, VI = s.v_inner();
  const int nouter = s.T1 * s.T2;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int co = cbeg * nouter; co < cend * nouter; ++co) {
    const int c = co / nouter, to = co - c * nouter;
    const int t1 = to / s.T2, t2 = to - t1 * s.T2;
    const float* Grow = G + ((int64_t)b * s.C + c) * VI + ((int64_t)(j1 + t1) * s.X2 + (j2 + t2)) * s.L;
    const float* Wc = W + (int64_t)c * RT + (int64_t)to * s.T;
    for (int t0 = 0; t0 < s.T; t0 += kTK) {
      __syncthreads();
      for (int idx = tid; idx < kTK * Rp; idx += 256) {
        int r = idx / kTK, tt = idx - r * kTK;
        float w = 0.f;
        if (r < s.R && t0 + tt < s.T) w = Wc[(int64_t)r * WI + t0 + tt];
        Ws[tt * Rp + r] = w;
      }
      for (int i = tid; i < JTILE + kTK - 1; i += 256) {
        int src = j0 + t0 + i;                  // window index for (jj, tt) is jj + tt
        Gs[i] = (src < s.L) ? Grow[src] : 0.f;
      }
      __syncthreads();
      float g[4];
      g[1] = Gs[4 * jx];
      g[2] = Gs[4 * jx + 1];
      g[3] = Gs[4 * jx + 2];
#pragma unroll
      for (int tt = 0; tt < kTK; ++tt) {
        g[0] = g[1]; g[1] = g[2]; g[2] = g[3];
        g[3] = Gs[4 * jx + 3 + tt];
        float4 a = *reinterpret_cast<const float4*>(&Ws[tt * Rp + 4 * rg]);
        float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], g[j], acc[i][j]);
      }
    }
  }
  const int64_t HI = s.h_inner();
  float* o = out + (int64_t)split * s.B * s.R * HI;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = 4 * rg + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int jj = j0 + 4 * jx + j;
      if (r < s.R && jj < s.Lin) o[((int64_t)b * s.R + r) * HI + (int64_t)outer * s.Lin + jj] = acc[i][j];
    }
  }
}

int dgrad_rg(int R) {
  int rg = 1;
  while (4 * rg < R) rg <<= 1;
  return rg;
}

}  // namespace

int nmfd_max_blocks(const NmfdShape& s) {
  return (int)(ceil_div(s.L, kT) * s.X1 * s.X2 * ceil_div(s.C, kT) * s.B);
}

int nmfd_recon_phi(const NmfdShape& s, const float* V, const float* W, const float* H, double beta, float* Pn,
                   float* Pp, double* loss_blocks, int max_blocks, double* loss_dev, cudaStream_t st) {
  dim3 grid((unsigned)(ceil_div(s.L, kT) * s.X1 * s.X2), (unsigned)ceil_div(s.C, kT), (unsigned)s.B);
  const int nblocks = (int)(grid.x * grid.y * grid.z);
  const bool loss = loss_blocks != nullptr;
  if (loss && nblocks > max_blocks) { set_error("nmfd loss: partial buffer too small"); return 1; }
  const int mode = beta_mode(beta);
#define NMFD_GO(M)                                                                                          \
  if (loss) nmfd_recon_kernel<M, true><<<grid, 256, 0, st>>>(s, V, W, H, (float)beta, Pn, Pp, loss_blocks); \
  else nmfd_recon_kernel<M, false><<<grid, 256, 0, st>>>(s, V, W, H, (float)beta, Pn, Pp, loss_blocks);
  switch (mode) {
    case kKL: NMFD_GO(kKL); break;
    case kEU: NMFD_GO(kEU); break;
    case kIS: NMFD_GO(kIS); break;
    default: NMFD_GO(kGeneric); break;
  }
#undef NMFD_GO
  NMF_LAUNCH_CHECK();
  if (loss) return sum_partials(loss_blocks, nblocks, loss_dev, st);
  return 0;
}

int nmfd_wgrad(const NmfdShape& s, const float* G, const float* H, float* out, cudaStream_t st) {
  dim3 grid((unsigned)(ceil_div(s.T, kT) * s.R * s.T1 * s.T2), (unsigned)ceil_div(s.C, kT), 1);
  nmfd_wgrad_kernel<<<grid, 256, 0, st>>>(s, G, H, out);
  NMF_LAUNCH_CHECK();
  return 0;
}

int nmfd_dgrad_nsplit(const NmfdShape& s) {
  const int rg = dgrad_rg(s.R);
  const int jtile = 4 * (256 / rg);
  int64_t tiles = ceil_div(s.Lin, jtile) * s.J1() * s.J2() * s.B;
  int64_t ns = ceil_div(148 * 4, tiles);
  if (ns > s.C) ns = s.C;
  if (ns > 64) ns = 64;
  if (ns < 1) ns = 1;
  return (int)ns;
}

int nmfd_dgrad(const NmfdShape& s, const float* G, const float* W, float* out, int nsplit, cudaStream_t st) {
  if (s.R > 256) { set_error("nmfd_dgrad: rank must be <= 256"); return 1; }
  const int rg = dgrad_rg(s.R);
  const int jtile = 4 * (256 / rg);
  dim3 grid((unsigned)(ceil_div(s.Lin, jtile) * s.J1() * s.J2()), (unsigned)nsplit, (unsigned)s.B);
  switch (rg) {
    case 1: nmfd_dgrad_kernel<1><<<grid, 256, 0, st>>>(s, G, W, out, nsplit); break;
    case 2: nmfd_dgrad_kernel<2><<<grid, 256, 0, st>>>(s, G, W, out, nsplit); break;
    case 4: nmfd_dgrad_kernel<4><<<grid, 256, 0, st>>>(s, G, W, out, nsplit); break;
    case 8: nmfd_dgrad_kernel<8><<<grid, 256, 0, st>>>(s, G, W, out, nsplit); break;
    case 16: nmfd_dgrad_kernel<16><<<grid, 256, 0, st>>>(s, G, W, out, nsplit); break;
    case 32: nmfd_dgrad_kernel<32><<<grid, 256, 0, st>>>(s, G, W, out, nsplit); break;
    default: nmfd_dgrad_kernel<64><<<grid, 256, 0, st>>>(s, G, W, out, nsplit); break;
  }
  NMF_LAUNCH_CHECK();
  return 0;
}

}  // namespace nmfb200
