// Dense per-group algebra for large groups (8 <= gs <= 64) as separate, latency-tuned launches
// behind the tensor-core contraction:
//
//   partial_reduce   sum the per-CTA partial Gram matrices of a 64-channel super-block in fixed
//                    chunk order, spread over many CTAs (the single "last CTA" doing this alone was
//                    ~70 % of the contraction kernel's elapsed time -- profiles/ncu_r01c_tc.txt)
//   fwd_factor       per group: mean, covariance, S = a*cov + b*I, Cholesky S = L L^T, W = L^-1,
//                    running-statistic EMA (domains in order: one CTA owns all domains of its group)
//   bwd_coef         per (domain, group): A1 = W^T, Bm = (2a/M) sym(W^T Phi(-R W^T) W), cvec
//
// All three keep a 64x64 problem in ONE 256-thread CTA arranged 16x16, each thread owning a 4x4
// register block (the whole matrix lives in registers during the factorisation; shared memory only
// carries the broadcast column / operand panels), so the 64 sequential Cholesky steps cost two
// barriers and ~20 FMAs each instead of a shared-memory round trip per element.
//
// Reference: utils/whitening.py:47-53,57-59 (/root/reference); backward: SURVEY.md §8a.
#include "dwt_common.cuh"
#include "norm_launch.h"

namespace dwt {
namespace {

constexpr int kSB = 64;                         // super-block edge
constexpr int kNacc = kSB * kSB + kSB;          // Gram + row sums
constexpr int LDS = kSB + 1;                    // padded leading dimension in shared memory
constexpr int kMat = kSB * LDS;                 // floats per shared matrix

// ------------------------------------------------------------------------------------------
// partial_reduce: out[p][e] = sum_c partial[p][c][e]   (p = domain*SB + sb, fixed order over c)
// grid (ceil(kNacc/64), problems), 256 threads = 4 chunk-quarters x 64 elements
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) partial_reduce_kernel(const float* __restrict__ partial, int nchunks,
                                                             float* __restrict__ out) {
  __shared__ float sQ[4][64];
  const int l = threadIdx.x & 63, q = threadIdx.x >> 6;
  const int e = blockIdx.x * 64 + l, p = blockIdx.y;
  const float* base = partial + (size_t)p * nchunks * kNacc;
  const int c0 = (nchunks * q) / 4, c1 = (nchunks * (q + 1)) / 4;
  float acc = 0.f;
  if (e < kNacc) {
    int c = c0;
    for (; c + 4 <= c1; c += 4) {
      const float v0 = __ldcg(base + (size_t)(c + 0) * kNacc + e), v1 = __ldcg(base + (size_t)(c + 1) * kNacc + e);
      const float v2 = __ldcg(base + (size_t)(c + 2) * kNacc + e), v3 = __ldcg(base + (size_t)(c + 3) * kNacc + e);
      acc = (((acc + v0) + v1) + v2) + v3;
    }
    for (; c < c1; ++c) acc += __ldcg(base + (size_t)c * kNacc + e);
  }
  sQ[q][l] = acc;
  __syncthreads();
  if (q == 0 && e < kNacc) out[(size_t)p * kNacc + e] = ((sQ[0][l] + sQ[1][l]) + sQ[2][l]) + sQ[3][l];
}

// ------------------------------------------------------------------------------------------
// register-blocked helpers (256 threads as 16 x 16, thread (bi,bj) owns rows 4bi.., cols 4bj..)
// ------------------------------------------------------------------------------------------
struct Blk {
  int bi, bj;
  bool act;
  __device__ Blk(int GS) : bi(threadIdx.x >> 4), bj(threadIdx.x & 15) { act = 4 * bi < GS && 4 * bj < GS; }
};

// C = op(A) * op(B) on GS x GS matrices in shared memory (leading dimension LDS), result in registers.
// TA: use A^T ; TB: use B^T.  Structural zeros are simply stored zeros.
template <bool TA, bool TB>
__device__ __forceinline__ void mm_block(const float* A, const float* B, int GS, const Blk& t, float (&c)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int s = 0; s < 4; ++s) c[r][s] = 0.f;
  if (!t.act) return;
  for (int k = 0; k < GS; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = TA ? A[k * LDS + 4 * t.bi + r] : A[(4 * t.bi + r) * LDS + k];
#pragma unroll
    for (int s = 0; s < 4; ++s) b[s] = TB ? B[(4 * t.bj + s) * LDS + k] : B[k * LDS + 4 * t.bj + s];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = 0; s < 4; ++s) c[r][s] = fmaf(a[r], b[s], c[r][s]);
  }
}

__device__ __forceinline__ void store_block(float* M, const Blk& t, const float (&c)[4][4]) {
  if (!t.act) return;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int s = 0; s < 4; ++s) M[(4 * t.bi + r) * LDS + 4 * t.bj + s] = c[r][s];
}

// In-register right-looking Cholesky of the SPD matrix held as 4x4 blocks.  The blocks are held TRANSPOSED:
// a[r][s] is element (i = 4*bj + s, j = 4*bi + r) -- for the symmetric input the same numbers as the (bi, bj) block, so
// the caller fills `a` as if it were untransposed -- which puts the 16 owners of a COLUMN block (fixed bi) into one
// half-warp: the pivot travels by shuffle and a step needs ONE block barrier (publish the scaled column, then every
// thread applies the rank-1 update from the double-buffered sCol), not two (round 1: pivot through shared memory,
// 128 barriers for a 64 x 64 group; profiles/ncu_r01i_tc_final.txt put 42 % of the kernel there).
// On exit the entries with i >= j hold L[i][j].  False on a non-positive pivot.
__device__ __forceinline__ bool cholesky_blocked(float (&a)[4][4], int GS, const Blk& t, float (*sCol)[kSB]) {
  bool ok = true;
  const unsigned hmask = (threadIdx.x & 16) ? 0xFFFF0000u : 0x0000FFFFu;
  for (int k = 0; k < GS; ++k) {
    const int kb = k >> 2, kk = k & 3, buf = k & 1;
    if (t.bi == kb) {                               // the half-warp that owns column k (j == k  <=>  r == kk)
      float pv = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q == kk) pv = a[q][q];                  // meaningful in the diagonal thread bj == kb
      pv = __shfl_sync(hmask, pv, kb, 16);
      ok = ok && (pv > 0.f);
      const float d = sqrtf(pv), inv = 1.f / d;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int i = 4 * t.bj + s;
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (r == kk) {
            float v = a[r][s];
            v = (i > k) ? v * inv : ((i == k) ? d : v);
            a[r][s] = v;
            sCol[buf][i] = (i > k) ? v : 0.f;
          }
      }
    }
    __syncthreads();
    if (t.act) {                                    // rank-1 update of the trailing matrix (sCol is 0 for rows <= k)
      float ci[4], cj[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { ci[q] = sCol[buf][4 * t.bj + q]; cj[q] = sCol[buf][4 * t.bi + q]; }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s) a[r][s] = fmaf(-cj[r], ci[s], a[r][s]);
    }
  }
  return ok;
}

// W = L^-1 for lower-triangular L (GS x GS in shared memory, zeros above the diagonal).  Diagonal
// 4x4 blocks are inverted by one thread each, then block sizes double: W21 = -W22 (L21 W11).
__device__ __forceinline__ void tri_inverse(const float* sL, float* sW, float* sT, int GS) {
  for (int e = threadIdx.x; e < GS * LDS; e += blockDim.x) sW[e] = 0.f;
  __syncthreads();
  if ((int)threadIdx.x < GS / 4) {
    const int o = 4 * threadIdx.x;
    float l[4][4], w[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = 0; s < 4; ++s) { l[r][s] = sL[(o + r) * LDS + o + s]; w[r][s] = 0.f; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      w[j][j] = 1.f / l[j][j];
#pragma unroll
      for (int i = j + 1; i < 4; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int k = j; k < i; ++k) acc = fmaf(l[i][k], w[k][j], acc);
        w[i][j] = -acc / l[i][i];
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = 0; s < 4; ++s) sW[(o + r) * LDS + o + s] = w[r][s];
  }
  __syncthreads();
  for (int s = 4; s < GS; s <<= 1) {
    // pair p: off-diagonal block rows [2ps+s, 2ps+2s) x cols [2ps, 2ps+s)
    const int npairs = GS / (2 * s), nel = npairs * s * s;
    for (int e = threadIdx.x; e < nel; e += blockDim.x) {      // T = L21 * W11  (W11 lower-triangular)
      const int p = e / (s * s), r = (e / s) % s, c = e % s, r0 = 2 * p * s + s, c0 = 2 * p * s;
      float acc = 0.f;
      for (int k = c; k < s; ++k) acc = fmaf(sL[(r0 + r) * LDS + c0 + k], sW[(c0 + k) * LDS + c0 + c], acc);
      sT[(r0 + r) * LDS + c0 + c] = acc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nel; e += blockDim.x) {      // W21 = -W22 * T  (W22 lower-triangular)
      const int p = e / (s * s), r = (e / s) % s, c = e % s, r0 = 2 * p * s + s, c0 = 2 * p * s;
      float acc = 0.f;
      for (int k = 0; k <= r; ++k) acc = fmaf(sW[(r0 + r) * LDS + r0 + k], sT[(r0 + k) * LDS + c0 + c], acc);
      sW[(r0 + r) * LDS + c0 + c] = -acc;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// fwd_factor: grid (G), 256 threads; loops over the domains in order (EMA sequence)
//   gram  [D][SB][kNacc]  reduced moments around the pilot shift (null: take the running buffers = eval)
//   shift [D][SB*64]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fwd_factor_kernel(const float* __restrict__ gram, const float* __restrict__ shift,
                                                         const Geom gm, const FwdFin f) {
  extern __shared__ __align__(16) float dsm[];
  float* sL = dsm;                 // Cholesky factor
  float* sW = sL + kMat;           // its inverse
  float* sT = sW + kMat;           // scratch of the inverse
  float* sC = sT + kMat;           // un-shrunk covariance (for the EMA)
  __shared__ float sCol[2][kSB], sMean[kSB];
  __shared__ int sBad, sBadDom;
  const int g = blockIdx.x, GS = gm.GS, nb = kSB / GS, sb = g / nb, o = (g % nb) * GS;
  const Blk t(GS);
  const int SB = (gm.C + kSB - 1) / kSB;
  const float invM = 1.f / gm.M;
  if (threadIdx.x == 0) sBad = 0;
  for (int d = 0; d < gm.D; ++d) {
    const float* G = gram ? gram + ((size_t)d * SB + sb) * kNacc : nullptr;
    const size_t gbase = ((size_t)d * gm.G + g) * GS * GS;
    for (int i = threadIdx.x; i < GS; i += blockDim.x) {
      const float mu = G ? shift[((size_t)d * SB + sb) * kSB + o + i] + G[kSB * kSB + o + i] * invM
                         : f.rmean[d][g * GS + i];
      sMean[i] = mu;
      f.save_mean[(size_t)d * gm.C + g * GS + i] = mu;
    }
    float a[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        float v = 0.f;
        if (t.act) {
          const int i = 4 * t.bi + r, j = 4 * t.bj + s, hi = i > j ? i : j, lo = i > j ? j : i;
          float cov;
          if (G) {
            cov = G[(o + hi) * kSB + o + lo] * invM - (G[kSB * kSB + o + i] * invM) * (G[kSB * kSB + o + j] * invM);
            sC[i * LDS + j] = cov;
          } else {
            cov = f.rcov[d][(size_t)g * GS * GS + i * GS + j];
          }
          v = f.a * cov + (i == j ? f.b : 0.f);
        }
        a[r][s] = v;
      }
    // the running buffers of this domain are fetched NOW (their latency hides behind the factorisation); the previous
    // domain's stores precede this point by a block barrier, so aliased buffers still see the ordered sequence
    constexpr int kEmaPer = kSB * kSB / 256;
    float rc_old[kEmaPer], rm_old = 0.f;
    const bool ema = G != nullptr && f.update_running;
    if (ema) {
#pragma unroll
      for (int n = 0; n < kEmaPer; ++n) {
        const int e = threadIdx.x + 256 * n;
        rc_old[n] = e < GS * GS ? f.rcov[d][(size_t)g * GS * GS + e] : 0.f;
      }
      if ((int)threadIdx.x < GS) rm_old = f.rmean[d][g * GS + threadIdx.x];
    }
    if (threadIdx.x == 0) sBadDom = 0;
    __syncthreads();
    if (!cholesky_blocked(a, GS, t, sCol)) { sBad = 1; sBadDom = 1; }
    if (t.act) {                                    // a[r][s] = element (i = 4bj+s, j = 4bi+r): keep the lower triangle
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s)
          sL[(4 * t.bj + s) * LDS + 4 * t.bi + r] = (4 * t.bi + r <= 4 * t.bj + s) ? a[r][s] : 0.f;
    }
    __syncthreads();
    tri_inverse(sL, sW, sT, GS);
    for (int e = threadIdx.x; e < GS * GS; e += blockDim.x) f.save_w[gbase + e] = sW[(e / GS) * LDS + e % GS];
    if (ema && !sBadDom) {                         // a non-PD batch covariance never reaches the running buffers
      const float m = f.momentum, k = 1.f - f.momentum;
#pragma unroll
      for (int n = 0; n < kEmaPer; ++n) {
        const int e = threadIdx.x + 256 * n;
        if (e < GS * GS) f.rcov[d][(size_t)g * GS * GS + e] = m * (sC[(e / GS) * LDS + e % GS] * f.unbias) + k * rc_old[n];
      }
      if ((int)threadIdx.x < GS) f.rmean[d][g * GS + threadIdx.x] = m * sMean[threadIdx.x] + k * rm_old;
    }
    __syncthreads();      // also orders this domain's buffer writes before the next domain's reads (aliasing)
  }
  if (threadIdx.x == 0 && sBad) atomicOr(f.status, DWT_STATUS_NOT_PD);
}

// ------------------------------------------------------------------------------------------
// bwd_coef: grid (G, 1, D), 256 threads.  rgram [D][SB][kNacc] = (R = sum dy xc^T | sdz = sum dy).
// coef[d][g] = A1 | Bm | cvec  with  dx = A1 dy + Bm x + cvec   (no affine epilogue on this path)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bwd_coef_kernel(const float* __restrict__ rgram, const Geom gm, const BwdFin f,
                                                       float* __restrict__ dybar) {
  extern __shared__ __align__(16) float dsm[];
  float* sW = dsm;
  float* sR = sW + kMat;
  float* sT1 = sR + kMat;
  float* sT2 = sT1 + kMat;
  __shared__ float sSdz[kSB], sMu[kSB];
  const int g = blockIdx.x, d = blockIdx.z, GS = gm.GS, nb = kSB / GS, sb = g / nb, o = (g % nb) * GS;
  const Blk t(GS);
  const int SB = (gm.C + kSB - 1) / kSB;
  const bool train = f.mode == DWT_MODE_TRAIN;
  const float* G = rgram ? rgram + ((size_t)d * SB + sb) * kNacc : nullptr;
  const size_t gbase = ((size_t)d * gm.G + g) * GS * GS;
  float* coef = f.coef + ((size_t)d * gm.G + g) * coef_stride(GS);
  for (int e = threadIdx.x; e < GS * GS; e += blockDim.x) {
    const int i = e / GS, j = e - i * GS;
    sW[i * LDS + j] = f.save_w[gbase + e];
    sR[i * LDS + j] = (G && train) ? G[(o + i) * kSB + o + j] : 0.f;
  }
  for (int i = threadIdx.x; i < GS; i += blockDim.x) {
    sSdz[i] = (G && train) ? G[kSB * kSB + o + i] : 0.f;
    sMu[i] = f.save_mean[(size_t)d * gm.C + g * GS + i];
    if (dybar) dybar[((size_t)d * SB + sb) * kSB + o + i] = sSdz[i] / gm.M;      // mean_M dy (0 in eval mode)
  }
  __syncthreads();
  float c[4][4];
  if (train) {
    mm_block<false, true>(sR, sW, GS, t, c);               // R W^T ; P = Phi(-R W^T)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int i = 4 * t.bi + r, j = 4 * t.bj + s;
        c[r][s] = (i > j) ? -c[r][s] : ((i == j) ? -0.5f * c[r][s] : 0.f);
      }
    store_block(sT1, t, c);
    __syncthreads();
    mm_block<true, false>(sW, sT1, GS, t, c);              // T = W^T P
    store_block(sT2, t, c);
    __syncthreads();
    mm_block<false, false>(sT2, sW, GS, t, c);             // S' = T W
    store_block(sT1, t, c);
    __syncthreads();
  }
  const float sc = f.a / gm.M;
  for (int e = threadIdx.x; e < GS * GS; e += blockDim.x) {
    const int i = e / GS, j = e - i * GS;
    const float bm = train ? sc * (sT1[i * LDS + j] + sT1[j * LDS + i]) : 0.f;
    coef[e] = (j >= i) ? sW[j * LDS + i] : 0.f;            // A1 = W^T
    coef[GS * GS + e] = bm;
    sT2[i * LDS + j] = bm;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < GS; i += blockDim.x) {
    float cv = 0.f;
    if (train) {
      for (int j = i; j < GS; ++j) cv = fmaf(sW[j * LDS + i], sSdz[j] / gm.M, cv);
      for (int j = 0; j < GS; ++j) cv = fmaf(sT2[i * LDS + j], sMu[j], cv);
      cv = -cv;
    }
    coef[2 * GS * GS + i] = cv;
  }
}

constexpr size_t kFactorSmem = sizeof(float) * 4 * kMat;
constexpr size_t kCoefSmem = sizeof(float) * 4 * kMat;

}  // namespace

int dense_init() {
  cudaError_t e = cudaFuncSetAttribute(fwd_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFactorSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(bwd_coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoefSmem);
  return (int)e;
}

// partial [problems][nchunks][kNacc] -> gram [problems][kNacc]
void dense_partial_reduce(const float* partial, int nchunks, int problems, float* gram, cudaStream_t st) {
  partial_reduce_kernel<<<dim3((kNacc + 63) / 64, problems), 256, 0, st>>>(partial, nchunks, gram);
}

// gram == nullptr: eval mode (running buffers -> W)
void dense_fwd_factor(const float* gram, const float* shift, const Geom& gm, const FwdFin& fin, cudaStream_t st) {
  fwd_factor_kernel<<<gm.G, 256, kFactorSmem, st>>>(gram, shift, gm, fin);
}

// rgram == nullptr: eval mode without affine (A1 = W^T only)
void dense_bwd_coef(const float* rgram, const Geom& gm, const BwdFin& fin, float* dybar, cudaStream_t st) {
  bwd_coef_kernel<<<dim3(gm.G, 1, gm.D), 256, kCoefSmem, st>>>(rgram, gm, fin, dybar);
}

}  // namespace dwt
