// Tensor-core path of the dense NMF factor update for sm_100a: tcgen05.mma + TMEM + TMA.
//
// One persistent, warp-specialised CTA per SM walks work items (128-row block of the row factor F,
// chunk of 128-column tiles).  Per tile (FlashAttention-shaped, but with a ratio instead of softmax):
//
//   TMA      G tile [128 c][KW] and V tile [128 m][128 c] (fp16, SWIZZLE_128B) -> shared memory
//   MMA-1    S[128 m][128 c]  = F_blk G_tile^T            tcgen05.mma SS, fp32 accumulators in TMEM
//   ratio    P = V * rcp(S*c1 + c2)  (== V / (F G^T + eps), nmf.py:65) -> fp16, written back to TMEM
//            over the S columns it was computed from (128 threads per tile, one TMEM lane each)
//   MMA-2    O[128 m][KW]   += P G_tile                    tcgen05.mma TS (A = P from TMEM, B = G tile MN-major)
//
// so neither S = WH nor P = V/(WH) ever leaves the SM (nmf.py:376-378 materialises both in HBM).
// F/G are fp16 copies of the factors scaled by a power of two; in split mode they carry hi|lo halves
// (KW = 2*Rp) and S = Fhi Ghi + Flo Ghi + Fhi Glo, O = P [Ghi|Glo] recovers ~22-bit factors.
//
// Warp roles (448 threads): 0 TMA producer | 1 MMA issuer (one lane) | 2-5 ratio warpgroup A (even
// tiles) | 6-9 ratio warpgroup B (odd tiles) | 10-13 epilogue (O -> fp32 partial numerators).
#include "tc_nmf.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

#include <vector>

#include "sm100_ptx.cuh"

namespace nmfb200 {

namespace {

constexpr int kRp = 64;              // padded rank handled by this kernel
constexpr int kTileM = 128, kTileN = 128;
constexpr int kThreads = 448;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0;        // two S/P stages: columns [0,128) and [128,256)
constexpr uint32_t kColO = 256;      // O accumulator: [256, 256 + KW)

struct TcKernelParams {
  int Mr, Nc;                 // valid rows of F / rows of G (= columns of Vm)
  int row_blocks, tiles, nchunks, tiles_per_chunk;
  float* part;                // [nchunks][Mr][ldp] fp32 numerators
  int64_t chunk_stride;
  int ldp;
  const int* exps;            // device: {v, aW, aH} power-of-two exponents of the fp16 copies
  int ef, eg;                 // which of exps[] belong to F and G
};

template <int KW, int NST>
struct SmemLayout {
  static constexpr int kFBytes = kTileM * KW * 2;
  static constexpr int kGBytes = kTileN * KW * 2;
  static constexpr int kVBytes = kTileM * kTileN * 2;
  static constexpr int kF = 0;
  static constexpr int kG = kF + 2 * kFBytes;
  static constexpr int kV = kG + NST * kGBytes;
  static constexpr int kBar = kV + NST * kVBytes;
  // barriers: f_full[2] f_empty[2] gv_full[NST] g_empty[NST] v_empty[NST] s_full[2] p_full[2] o_full o_empty
  static constexpr int kNumBars = 4 + 3 * NST + 6;
  static constexpr int kTmemPtr = kBar + 8 * kNumBars;
  static constexpr int kTotal = kTmemPtr + 16;
};

template <int KW, int NST, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
tc_contract_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmG,
                   const __grid_constant__ CUtensorMap tmV, const TcKernelParams p) {
  using L = SmemLayout<KW, NST>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw32 = ptx::smem_u32(smem_raw);
  const uint32_t sbase = (raw32 + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (sbase - raw32);
  const uint32_t sF = sbase + L::kF, sG = sbase + L::kG, sV = sbase + L::kV;
  const uint32_t bar0 = sbase + L::kBar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int B_FFULL = 0, B_FEMPTY = 2, B_GVFULL = 4, B_GEMPTY = 4 + NST, B_VEMPTY = 4 + 2 * NST,
            B_SFULL = 4 + 3 * NST, B_PFULL = B_SFULL + 2, B_OFULL = B_PFULL + 2, B_OEMPTY = B_OFULL + 1;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_al + L::kTmemPtr);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmF); ptx::prefetch_tmap(&tmG); ptx::prefetch_tmap(&tmV);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(BAR(B_FFULL + i), 1); ptx::mbar_init(BAR(B_FEMPTY + i), 1); }
    for (int i = 0; i < NST; ++i) {
      ptx::mbar_init(BAR(B_GVFULL + i), 1);
      ptx::mbar_init(BAR(B_GEMPTY + i), 1);
      ptx::mbar_init(BAR(B_VEMPTY + i), 128);
    }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(BAR(B_SFULL + i), 1); ptx::mbar_init(BAR(B_PFULL + i), 128); }
    ptx::mbar_init(BAR(B_OFULL), 1);
    ptx::mbar_init(BAR(B_OEMPTY), 128);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) {
    ptx::tmem_alloc(sbase + L::kTmemPtr, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;

  const int total_items = p.row_blocks * p.nchunks;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t it = 0, t = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const int rb = item % p.row_blocks, chunk = item / p.row_blocks;
        const int tb = chunk * p.tiles_per_chunk;
        const int te = min(p.tiles, tb + p.tiles_per_chunk);
        const uint32_t fb = it & 1;
        ptx::mbar_wait(BAR(B_FEMPTY + fb), ((it >> 1) & 1) ^ 1);
        ptx::mbar_expect_tx(BAR(B_FFULL + fb), L::kFBytes);
        for (int kb = 0; kb < KW / 64; ++kb)
          ptx::tma_load_2d(&tmF, BAR(B_FFULL + fb), sF + fb * L::kFBytes + kb * (kTileM * 128), kb * 64, rb * kTileM);
        for (int j = tb; j < te; ++j, ++t) {
          const uint32_t s = t % NST, ph = (t / NST) & 1;
          ptx::mbar_wait(BAR(B_GEMPTY + s), ph ^ 1);
          ptx::mbar_wait(BAR(B_VEMPTY + s), ph ^ 1);
          ptx::mbar_expect_tx(BAR(B_GVFULL + s), L::kGBytes + L::kVBytes);
          for (int kb = 0; kb < KW / 64; ++kb)
            ptx::tma_load_2d(&tmG, BAR(B_GVFULL + s), sG + s * L::kGBytes + kb * (kTileN * 128), kb * 64, j * kTileN);
          for (int vb = 0; vb < 2; ++vb)
            ptx::tma_load_2d(&tmV, BAR(B_GVFULL + s), sV + s * L::kVBytes + vb * (kTileM * 128), j * kTileN + vb * 64,
                             rb * kTileM);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================
    if (lane == 0) {
      constexpr uint32_t idescS = ptx::idesc_f16(kTileM, kTileN, 0, 0);
      constexpr uint32_t idescO = ptx::idesc_f16(kTileM, KW, 0, 1);
      uint32_t it = 0, t = 0;
      // S = sum over terms (F part, G part): fast: (0,0); split: (hi,hi), (lo,hi), (hi,lo)
      constexpr int kTerms = SPLIT ? 3 : 1;
      const int termF[3] = {0, 1, 0}, termG[3] = {0, 0, 1};
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const int chunk = item / p.row_blocks;
        const int tb = chunk * p.tiles_per_chunk;
        const int te = min(p.tiles, tb + p.tiles_per_chunk);
        const int n = te - tb;
        const uint32_t fb = it & 1;
        ptx::mbar_wait(BAR(B_FFULL + fb), (it >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t fbase = sF + fb * L::kFBytes;

        auto issue_S = [&](uint32_t tt) {
          const uint32_t s = tt % NST, st = tt & 1;
          ptx::mbar_wait(BAR(B_GVFULL + s), (tt / NST) & 1);
          ptx::tc_fence_after();
          const uint32_t gbase = sG + s * L::kGBytes;
          uint32_t acc = 0;
#pragma unroll
          for (int term = 0; term < kTerms; ++term) {
            const uint32_t fa = fbase + termF[term] * (kTileM * 128);
            const uint32_t ga = gbase + termG[term] * (kTileN * 128);
#pragma unroll
            for (int ks = 0; ks < kRp / 16; ++ks) {
              const uint64_t ad = ptx::smem_desc_sw128(fa + ks * 32, 16, 1024);
              const uint64_t bd = ptx::smem_desc_sw128(ga + ks * 32, 16, 1024);
              ptx::mma_ss(tmem + kColS + st * 128, ad, bd, idescS, acc);
              acc = 1;
            }
          }
          ptx::mma_commit(BAR(B_SFULL + st));
        };
        auto issue_O = [&](uint32_t tt, bool first, bool last) {
          const uint32_t s = tt % NST, st = tt & 1;
          ptx::mbar_wait(BAR(B_PFULL + st), (tt >> 1) & 1);
          if (first) ptx::mbar_wait(BAR(B_OEMPTY), (it & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t gbase = sG + s * L::kGBytes;
#pragma unroll
          for (int ks = 0; ks < kTileN / 16; ++ks) {
            // B = G tile as [K = 16 c-rows][N = KW] MN-major: 8-row groups 1024 B apart, 64-wide column blocks
            // (hi | lo) one tile-block (16 KB) apart
            const uint64_t bd = ptx::smem_desc_sw128(gbase + ks * 2048, kTileN * 128, 1024);
            ptx::mma_ts(tmem + kColO, tmem + kColS + st * 128 + ks * 8, bd, idescO, (first && ks == 0) ? 0u : 1u);
          }
          ptx::mma_commit(BAR(B_GEMPTY + s));
          if (last) ptx::mma_commit(BAR(B_OFULL));
        };

        issue_S(t);
        for (int j = 0; j < n; ++j) {
          if (j + 1 < n) issue_S(t + j + 1);
          else ptx::mma_commit(BAR(B_FEMPTY + fb));
          issue_O(t + j, j == 0, j == n - 1);
        }
        t += n;
      }
    }
  } else if (warp < 10) {
    // =========================== ratio warpgroups =======================
    const int g = (warp - 2) >> 2;             // 0: even tiles, 1: odd tiles
    const int q = warp & 3;                    // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int ev = p.exps[0], ea = p.exps[p.ef], eb = p.exps[p.eg];
    const float c1 = exp2f((float)(ev - ea - eb));     // x' = S~ * 2^(v-a-b) + eps * 2^v
    const float c2 = kEps * exp2f((float)ev);
    uint32_t t = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int chunk = item / p.row_blocks;
      const int tb = chunk * p.tiles_per_chunk;
      const int te = min(p.tiles, tb + p.tiles_per_chunk);
      const int n = te - tb;
      for (int j = 0; j < n; ++j) {
        const uint32_t tt = t + j;
        if ((int)(tt & 1) != g) continue;
        const uint32_t s = tt % NST, st = tt & 1;
        ptx::mbar_wait(BAR(B_GVFULL + s), (tt / NST) & 1);      // V tile landed (TMA -> this thread)
        ptx::mbar_wait(BAR(B_SFULL + st), (tt >> 1) & 1);       // S tile complete
        ptx::tc_fence_after();
        const uint32_t vrow = sV + s * L::kVBytes + row * 128;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t sreg[32];
          ptx::tmem_ld32(tmem + lane_addr + kColS + st * 128 + c4 * 32, sreg);
          uint4 vv[4];
          const uint32_t vsub = vrow + (c4 >> 1) * (kTileM * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t chunk16 = (uint32_t)((c4 & 1) * 4 + k) ^ (uint32_t)(row & 7);
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(vv[k].x), "=r"(vv[k].y), "=r"(vv[k].z), "=r"(vv[k].w)
                         : "r"(vsub + (chunk16 << 4)));
          }
          ptx::tc_wait_ld();
          uint32_t preg[16];
          const uint32_t* vw = reinterpret_cast<const uint32_t*>(vv);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const __half2 hv = *reinterpret_cast<const __half2*>(&vw[i]);
            const float2 vf = __half22float2(hv);
            const float x0 = fmaf(__uint_as_float(sreg[2 * i]), c1, c2);
            const float x1 = fmaf(__uint_as_float(sreg[2 * i + 1]), c1, c2);
            const float p0 = vf.x * ptx::rcp_approx(x0);
            const float p1 = vf.y * ptx::rcp_approx(x1);
            preg[i] = ptx::pack_f16x2_sat(p0, p1);
          }
          ptx::tmem_st16(tmem + lane_addr + kColS + st * 128 + c4 * 16, preg);
        }
        ptx::tc_wait_st();
        ptx::tc_fence_before();
        ptx::mbar_arrive(BAR(B_PFULL + st));
        ptx::mbar_arrive(BAR(B_VEMPTY + s));
      }
      t += n;
    }
  } else {
    // =========================== epilogue warpgroup =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float oscale = exp2f(-(float)p.exps[p.eg]);      // O = sum P * (G * 2^eg)
    uint32_t it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int rb = item % p.row_blocks, chunk = item / p.row_blocks;
      ptx::mbar_wait(BAR(B_OFULL), it & 1);
      ptx::tc_fence_after();
      const int64_t grow = (int64_t)rb * kTileM + row;
      float* dst = p.part + (int64_t)chunk * p.chunk_stride + grow * p.ldp;
#pragma unroll
      for (int c = 0; c < kRp / 32; ++c) {
        uint32_t hi[32];
        ptx::tmem_ld32(tmem + lane_addr + kColO + c * 32, hi);
        float o[32];
        if (SPLIT) {
          uint32_t lo[32];
          ptx::tmem_ld32(tmem + lane_addr + kColO + kRp + c * 32, lo);
          ptx::tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = (__uint_as_float(hi[i]) + __uint_as_float(lo[i])) * oscale;
        } else {
          ptx::tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(hi[i]) * oscale;
        }
        if (grow < p.Mr) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(dst + c * 32 + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(BAR(B_OEMPTY));
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, kTmemCols);
}

// ---- operand preparation --------------------------------------------------------------------------

__device__ __forceinline__ int pow2_exp_for(float mx) {
  // exponent a with mx * 2^a in [2^13, 2^14); 0 for an all-zero / non-finite matrix
  if (!(mx > 0.f) || !isfinite(mx)) return 0;
  int e;
  frexpf(mx, &e);               // mx = m * 2^e, m in [0.5, 1)
  return 14 - e;
}

__global__ void set_vexp_kernel(const float* __restrict__ minmax, int* __restrict__ exps) {
  if (threadIdx.x == 0 && blockIdx.x == 0) exps[0] = pow2_exp_for(minmax[1]);
}

// V (N x C fp32, ld) -> V16 (N x ldc) and Vt16 (C x ldn), both scaled by 2^exps[0].  64x64 tiles.
__global__ void __launch_bounds__(256)
v_to_f16_kernel(const float* __restrict__ V, int64_t ldv, int N, int C, __half* __restrict__ V16, int64_t ldc,
                __half* __restrict__ Vt16, int64_t ldn, const int* __restrict__ exps) {
  __shared__ float tile[64][65];
  const float sc = exp2f((float)exps[0]);
  const int n0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  for (int idx = threadIdx.x; idx < 64 * 64; idx += 256) {
    int r = idx >> 6, c = idx & 63;
    float v = (n0 + r < N && c0 + c < C) ? V[(int64_t)(n0 + r) * ldv + c0 + c] * sc : 0.f;
    tile[r][c] = v;
    if (n0 + r < N && c0 + c < C) V16[(int64_t)(n0 + r) * ldc + c0 + c] = __float2half_rn(v);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 64 * 64; idx += 256) {
    int c = idx >> 6, r = idx & 63;
    if (n0 + r < N && c0 + c < C) Vt16[(int64_t)(c0 + c) * ldn + n0 + r] = __float2half_rn(tile[r][c]);
  }
}

__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, int64_t n, unsigned int* __restrict__ bits) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) m = fmaxf(m, x[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(bits, __float_as_uint(m));
}

// fp32 factor (rows x R) -> fp16 operand copy (rows x KW): [hi(0..Rp) | lo(Rp..2Rp)], scaled by 2^a where a is
// derived from *absmax_bits; pad columns stay zero (buffer is zero-initialised once).  Writes exps[slot].
template <bool SPLIT>
__global__ void __launch_bounds__(256)
factor_to_f16_kernel(const float* __restrict__ x, int64_t rows, int R, __half* __restrict__ out, int KW,
                     const unsigned int* __restrict__ absmax_bits, int* __restrict__ exps, int slot) {
  const int a = pow2_exp_for(__uint_as_float(*absmax_bits));
  if (blockIdx.x == 0 && threadIdx.x == 0) exps[slot] = a;
  const float sc = exp2f((float)a);
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= rows * R) return;
  const int64_t row = idx / R;
  const int r = (int)(idx - row * R);
  const float xs = x[idx] * sc;
  const __half hi = __float2half_rn(xs);
  out[row * KW + r] = hi;
  if (SPLIT) out[row * KW + kRp + r] = __float2half_rn(xs - __half2float(hi));
}

// ---- host side --------------------------------------------------------------------------------------

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  return fn;
}

// 2-D fp16 row-major tensor (rows x cols, row pitch ld elements), box 64 cols x 128 rows, SWIZZLE_128B
int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return 2; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r)); return 2; }
  return 0;
}

struct Plan { int row_blocks, tiles, nchunks, tpc; };

Plan make_plan(int64_t Mr, int64_t Nc, int num_sms) {
  Plan pl;
  pl.row_blocks = (int)ceil_div(Mr, kTileM);
  pl.tiles = (int)ceil_div(Nc, kTileN);
  int best = 1;
  double best_eff = -1.0;
  for (int nch = 1; nch <= pl.tiles && nch <= 64; ++nch) {
    int tpc = (int)ceil_div(pl.tiles, nch);
    if (tpc < 4 && nch > 1) break;
    int real = (int)ceil_div(pl.tiles, tpc);
    if (real != nch) continue;
    int64_t items = (int64_t)pl.row_blocks * nch;
    double eff = (double)items / (double)(ceil_div(items, num_sms) * num_sms);
    // fewer chunks = less partial traffic: only move on for a clear win
    if (eff > best_eff + 0.03) { best_eff = eff; best = nch; }
    if (best_eff >= 0.97) break;
  }
  pl.nchunks = best;
  pl.tpc = (int)ceil_div(pl.tiles, best);
  return pl;
}

}  // namespace

struct TcState {
  int device = 0, num_sms = 148;
  int64_t N = 0, C = 0, R = 0;
  bool split = true;
  int KW = 128;
  int64_t ldc = 0, ldn = 0;
  __half *V16 = nullptr, *Vt16 = nullptr, *W16 = nullptr, *H16 = nullptr;
  float* part = nullptr;
  int64_t part_floats = 0;
  float* colsum = nullptr;       // [2][R]  0 = W, 1 = H
  float* cs_scratch = nullptr;
  int64_t cs_scratch_floats = 0;
  unsigned int* absmax = nullptr;   // [2]
  int* exps = nullptr;              // {v, aW, aH}
  CUtensorMap tmV, tmVt, tmW, tmH;
  Plan plan_w, plan_h;
  bool dirty_w = true, dirty_h = true, has_target = false;
};

bool tc_shape_supported(int64_t N, int64_t C, int64_t R) {
  return R >= 1 && R <= kRp && N >= 1 && C >= 1 && N < (1ll << 31) && C < (1ll << 31);
}

void tc_destroy(TcState* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaFree(s->V16); cudaFree(s->Vt16); cudaFree(s->W16); cudaFree(s->H16); cudaFree(s->part);
  cudaFree(s->colsum); cudaFree(s->cs_scratch); cudaFree(s->absmax); cudaFree(s->exps);
  delete s;
}

int tc_create(TcState** out, int device, int64_t N, int64_t C, int64_t R, bool split) {
  *out = nullptr;
  TcState* s = new TcState();
  s->device = device; s->N = N; s->C = C; s->R = R; s->split = split;
  s->KW = split ? 2 * kRp : kRp;
  s->ldc = round_up(C, 8);
  s->ldn = round_up(N, 8);
  cudaDeviceProp prop;
  NMF_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  s->num_sms = prop.multiProcessorCount;
  if (prop.major != 10) { delete s; set_error("the tensor-core path needs an sm_100 device"); return 1; }
  s->plan_w = make_plan(C, N, s->num_sms);
  s->plan_h = make_plan(N, C, s->num_sms);
  int64_t pw = (int64_t)s->plan_w.nchunks * C * kRp, ph = (int64_t)s->plan_h.nchunks * N * kRp;
  s->part_floats = pw > ph ? pw : ph;
  s->cs_scratch_floats = colsum_scratch_floats(N > C ? N : C, (int)R, 1);
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaMalloc(&s->V16, (size_t)N * s->ldc * 2);
  if (e == cudaSuccess) e = cudaMalloc(&s->Vt16, (size_t)C * s->ldn * 2);
  if (e == cudaSuccess) e = cudaMalloc(&s->W16, (size_t)C * s->KW * 2);
  if (e == cudaSuccess) e = cudaMalloc(&s->H16, (size_t)N * s->KW * 2);
  if (e == cudaSuccess) e = cudaMalloc(&s->part, (size_t)s->part_floats * 4);
  if (e == cudaSuccess) e = cudaMalloc(&s->colsum, 2 * R * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&s->cs_scratch, s->cs_scratch_floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&s->absmax, 2 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMalloc(&s->exps, 4 * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(s->W16, 0, (size_t)C * s->KW * 2);
  if (e == cudaSuccess) e = cudaMemset(s->H16, 0, (size_t)N * s->KW * 2);
  if (e == cudaSuccess) e = cudaMemset(s->exps, 0, 4 * sizeof(int));
  if (e != cudaSuccess) {
    tc_destroy(s);
    set_error(std::string("tc_create cudaMalloc: ") + cudaGetErrorString(e));
    return 2;
  }
  int rc = 0;
  rc |= make_tmap(&s->tmV, s->V16, N, C, s->ldc);
  rc |= make_tmap(&s->tmVt, s->Vt16, C, N, s->ldn);
  rc |= make_tmap(&s->tmW, s->W16, C, s->KW, s->KW);
  rc |= make_tmap(&s->tmH, s->H16, N, s->KW, s->KW);
  if (rc) { tc_destroy(s); return 2; }
  *out = s;
  return 0;
}

bool tc_supports_beta(const TcState*, double beta) { return beta == 1.0; }
bool tc_supports_loss(const TcState*, double) { return false; }

int tc_set_target(TcState* s, const float* V, int64_t ldv, const float* minmax_dev, cudaStream_t st) {
  set_vexp_kernel<<<1, 32, 0, st>>>(minmax_dev, s->exps);
  NMF_LAUNCH_CHECK();
  dim3 grid((unsigned)ceil_div(s->C, 64), (unsigned)ceil_div(s->N, 64));
  v_to_f16_kernel<<<grid, 256, 0, st>>>(V, ldv, (int)s->N, (int)s->C, s->V16, s->ldc, s->Vt16, s->ldn, s->exps);
  NMF_LAUNCH_CHECK();
  s->has_target = true;
  return 0;
}

void tc_mark_dirty(TcState* s, bool w, bool h) {
  if (w) s->dirty_w = true;
  if (h) s->dirty_h = true;
}

namespace {

// (re)build the fp16 operand copy + column sums of one factor; `have_absmax`: absmax[which] already holds max
int refresh_factor(TcState* s, int which, const float* x, bool have_absmax, cudaStream_t st) {
  const int64_t rows = which == 0 ? s->C : s->N;
  const int64_t n = rows * s->R;
  if (!have_absmax) {
    NMF_CUDA_CHECK(cudaMemsetAsync(s->absmax + which, 0, sizeof(unsigned int), st));
    int64_t nb = ceil_div(n, 256 * 8);
    if (nb > 1184) nb = 1184;
    absmax_kernel<<<(unsigned)nb, 256, 0, st>>>(x, n, s->absmax + which);
    NMF_LAUNCH_CHECK();
  }
  __half* out = which == 0 ? s->W16 : s->H16;
  if (s->split)
    factor_to_f16_kernel<true><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(x, rows, (int)s->R, out, s->KW,
                                                                         s->absmax + which, s->exps, 1 + which);
  else
    factor_to_f16_kernel<false><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(x, rows, (int)s->R, out, s->KW,
                                                                          s->absmax + which, s->exps, 1 + which);
  NMF_LAUNCH_CHECK();
  return factor_colsum(x, rows, (int)s->R, 1, s->cs_scratch, s->cs_scratch_floats, s->colsum + which * s->R, st);
}

int ensure_synced(TcState* s, const float* W, const float* H, cudaStream_t st) {
  if (!s->has_target) { set_error("tensor-core path: set_target has not been called"); return 3; }
  if (s->dirty_w) { int rc = refresh_factor(s, 0, W, false, st); if (rc) return rc; s->dirty_w = false; }
  if (s->dirty_h) { int rc = refresh_factor(s, 1, H, false, st); if (rc) return rc; s->dirty_h = false; }
  return 0;
}

template <int KW, int NST, bool SPLIT>
int launch_contract_t(TcState* s, int which, cudaStream_t st) {
  using L = SmemLayout<KW, NST>;
  auto kern = tc_contract_kernel<KW, NST, SPLIT>;
  static bool attr_set = false;
  const int smem = L::kTotal + 1024;     // slack so the kernel-visible base can be 1024-aligned
  if (!attr_set) {
    NMF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  TcKernelParams p{};
  const Plan& pl = which == 0 ? s->plan_w : s->plan_h;
  p.Mr = (int)(which == 0 ? s->C : s->N);
  p.Nc = (int)(which == 0 ? s->N : s->C);
  p.row_blocks = pl.row_blocks; p.tiles = pl.tiles; p.nchunks = pl.nchunks; p.tiles_per_chunk = pl.tpc;
  p.part = s->part; p.chunk_stride = (int64_t)p.Mr * kRp; p.ldp = kRp;
  p.exps = s->exps;
  p.ef = which == 0 ? 1 : 2;
  p.eg = which == 0 ? 2 : 1;
  const int items = pl.row_blocks * pl.nchunks;
  const int grid = items < s->num_sms ? items : s->num_sms;
  if (which == 0)
    kern<<<grid, kThreads, smem, st>>>(s->tmW, s->tmH, s->tmVt, p);
  else
    kern<<<grid, kThreads, smem, st>>>(s->tmH, s->tmW, s->tmV, p);
  NMF_LAUNCH_CHECK();
  return 0;
}

int launch_contract(TcState* s, int which, cudaStream_t st) {
  if (s->split) return launch_contract_t<2 * kRp, 2, true>(s, which, st);
  return launch_contract_t<kRp, 3, false>(s, which, st);
}

int tc_apply(TcState* s, int which, float* param, double gamma, double l1, double l2, cudaStream_t st) {
  const Plan& pl = which == 0 ? s->plan_w : s->plan_h;
  const int64_t rows = which == 0 ? s->C : s->N;
  NMF_CUDA_CHECK(cudaMemsetAsync(s->absmax + which, 0, sizeof(unsigned int), st));
  ApplyArgs a{};
  a.param = param; a.numel = rows * s->R; a.R = (int)s->R; a.inner = 1; a.rowlen = s->R;
  a.num = s->part; a.den = nullptr; a.nchunks = pl.nchunks; a.chunk_stride = rows * kRp; a.ldp = kRp;
  a.kl_den = s->colsum + (which == 0 ? 1 : 0) * s->R;   // W update divides by colsum(H), H update by colsum(W)
  a.out_scale = nullptr;
  a.gamma = (float)gamma; a.l1 = (float)l1; a.l2 = (float)l2;
  a.absmax_bits = s->absmax + which;
  int rc = apply_update(a, st);
  if (rc) return rc;
  return refresh_factor(s, which, param, true, st);
}

}  // namespace

int tc_update_w(TcState* s, float* W, const float* H, double beta, double gamma, double l1, double l2,
                cudaStream_t st) {
  (void)beta;
  int rc = ensure_synced(s, W, H, st);
  if (rc) return rc;
  rc = launch_contract(s, 0, st);
  if (rc) return rc;
  return tc_apply(s, 0, W, gamma, l1, l2, st);
}

int tc_update_h(TcState* s, const float* W, float* H, double beta, double gamma, double l1, double l2,
                cudaStream_t st) {
  (void)beta;
  int rc = ensure_synced(s, W, H, st);
  if (rc) return rc;
  rc = launch_contract(s, 1, st);
  if (rc) return rc;
  return tc_apply(s, 1, H, gamma, l1, l2, st);
}

int tc_w_partial(TcState* s, const float* W, const float* H, double beta, float* partial, cudaStream_t st) {
  (void)beta;
  int rc = ensure_synced(s, W, H, st);
  if (rc) return rc;
  rc = launch_contract(s, 0, st);
  if (rc) return rc;
  const int64_t CR = s->C * s->R;
  rc = reduce_chunks(s->part, s->plan_w.nchunks, s->C * kRp, s->C, (int)s->R, kRp, partial, st);
  if (rc) return rc;
  NMF_CUDA_CHECK(cudaMemcpyAsync(partial + CR, s->colsum + s->R, s->R * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int tc_contract_only(TcState* s, const float* W, const float* H, int which, double beta, cudaStream_t st) {
  (void)beta;
  int rc = ensure_synced(s, W, H, st);
  if (rc) return rc;
  return launch_contract(s, which, st);
}

int tc_loss(TcState*, const float*, const float*, double, double*, cudaStream_t) {
  set_error("tensor-core loss kernel not available");
  return 1;
}

}  // namespace nmfb200
