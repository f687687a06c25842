// Register-resident path for small channel groups (GS = 1, 2, 4): domain batch norm
// (GS = 1) and the whitening layers of the shipped ResNet-50-DWT / LeNet (GS = 4).
//
// Each thread owns VEC consecutive pixels of all GS channels of one group, so a warp
// reads GS fully coalesced row segments; every kernel is one HBM pass:
//   stats       read x                    -> shifted moments -> mean, cov, W, running-stat EMA
//   apply       read x, write y           y = W (x - mean) [* gamma + beta] [relu]
//   bwd_reduce  read x, dout              -> R = sum dz xc^T, sum dz -> A1, Bm, cvec, dgamma, dbeta
//   bwd_apply   read x, dout, write dx    dx = A1 dz + Bm x + cvec
// i.e. 12 B/element forward + 20 B/element backward, the algorithmic minimum of SURVEY.md §8d.
//
// Work decomposition.  A "problem" is one (domain, group).  A CTA of 8 warps serves `ppc`
// consecutive groups of one domain, 8/ppc warps ("team") per problem, so that sites with
// thousands of tiny problems (domain BN at 7x7: 6144 problems of 12.5 KB) still run a few
// hundred long-lived CTAs instead of thousands of short ones whose prologue/epilogue
// round trips dominate; sites with few large problems instead split each problem over
// `nchunks` CTAs (then ppc = 1) that meet through per-CTA partials and an arrival counter.
// The small dense algebra (4x4 Cholesky, triangular inverse, backward coefficients, EMA)
// is done by ONE thread per problem, entirely in registers.
//
// Reference: utils/whitening.py:37-61, utils/batch_norm.py:54-69 (/root/reference).
#include "dwt_common.cuh"
#include "norm_launch.h"

namespace dwt {
namespace {

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = __ldg(p);
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    *p = v[0];
  }
}

// out_c = bp_c + sum_{j<=c} Wp[c][j] x_j : W is lower-triangular (Cholesky basis).
// One fixed FMA order, shared by forward apply and the backward's ReLU-mask recompute so
// both see bit-identical pre-activations.
template <int GS>
__device__ __forceinline__ void apply_group(const float (&Wp)[GS * (GS + 1) / 2], const float (&bp)[GS],
                                            const float (&x)[GS], float (&out)[GS]) {
#pragma unroll
  for (int c = 0; c < GS; ++c) {
    float acc = bp[c];
#pragma unroll
    for (int j = 0; j <= c; ++j) acc = fmaf(Wp[c * (c + 1) / 2 + j], x[j], acc);
    out[c] = acc;
  }
}

// Per-thread copy of the group's forward map: Wp = diag(gamma) W, bp = gamma (-W mean) + beta.
template <int GS, int EPI>
__device__ __forceinline__ void load_forward_map(const float* save_w_g, const float* mean_g, const float* gamma_g,
                                                 const float* beta_g, float (&Wp)[GS * (GS + 1) / 2],
                                                 float (&bp)[GS]) {
#pragma unroll
  for (int c = 0; c < GS; ++c) {
    float b = 0.f;
#pragma unroll
    for (int j = 0; j <= c; ++j) {
      float w = __ldg(save_w_g + c * GS + j);
      b = fmaf(-w, __ldg(mean_g + j), b);
      Wp[c * (c + 1) / 2 + j] = w;
    }
    bp[c] = b;
  }
  if constexpr ((EPI & DWT_EPI_AFFINE) != 0) {
#pragma unroll
    for (int c = 0; c < GS; ++c) {
      const float ga = __ldg(gamma_g + c), be = __ldg(beta_g + c);
#pragma unroll
      for (int j = 0; j <= c; ++j) Wp[c * (c + 1) / 2 + j] *= ga;
      bp[c] = fmaf(ga, bp[c], be);
    }
  }
}

// Team geometry of the calling thread.
struct Team {
  int wpp;       // warps per problem
  int team;      // team index inside the CTA
  int ttid;      // thread index inside the team
  int tthreads;  // threads per team
  int g;         // group served (may be >= G in the last CTA of a row: then !valid)
  bool valid;
  __device__ __forceinline__ Team(const Geom& gm) {
    wpp = kWarps / gm.ppc;
    tthreads = wpp * 32;
    team = threadIdx.x / tthreads;
    ttid = threadIdx.x - team * tthreads;
    g = blockIdx.y * gm.ppc + team;
    valid = g < gm.G;
  }
};

// item -> (image n, pixel-vector pv) of the flattened per-group work list
struct ItemMap {
  unsigned PV;          // pixel vectors per row
  unsigned img_stride;  // C*HW   (N*C*HW < 2^31 is checked on the host: 32-bit element offsets)
  __device__ __forceinline__ unsigned offset(unsigned item, int VEC) const {
    unsigned n = item / PV, pv = item - n * PV;
    return n * img_stride + pv * VEC;
  }
};

template <int GS, int VEC> struct Unroll {
  static constexpr int raw = 16 / (GS * VEC);                                   // ~16 floats per tensor per batch
  static constexpr int one = raw < 1 ? 1 : (raw > 8 ? 8 : raw);
  static constexpr int stats = 2 * one;                                        // ~32 floats in flight per thread
};

// ------------------------------------------------------------------------------------------
// single-thread finalize steps (GS <= 4, everything in registers)
// ------------------------------------------------------------------------------------------
template <int GS>
__device__ __forceinline__ void factor_thread(const Geom& gm, const FwdFin& f, int d, int g, const float (&mean)[GS],
                                              const float (&cov)[GS][GS], bool store_cov) {
  const size_t gbase = ((size_t)d * gm.G + g) * GS * GS;
  float L[GS][GS], W[GS][GS];
#pragma unroll
  for (int i = 0; i < GS; ++i) {
    f.save_mean[(size_t)d * gm.C + g * GS + i] = mean[i];
#pragma unroll
    for (int j = 0; j < GS; ++j) {
      if (store_cov) f.save_cov[gbase + i * GS + j] = cov[i][j];
      L[i][j] = f.a * cov[i][j] + (i == j ? f.b : 0.f);
      W[i][j] = 0.f;
    }
  }
  bool bad = false;
#pragma unroll
  for (int k = 0; k < GS; ++k) {
    bad |= !(L[k][k] > 0.f);
    L[k][k] = sqrtf(L[k][k]);
    const float inv = 1.f / L[k][k];
#pragma unroll
    for (int i = k + 1; i < GS; ++i) L[i][k] *= inv;
#pragma unroll
    for (int i = k + 1; i < GS; ++i)
#pragma unroll
      for (int j = k + 1; j <= i; ++j) L[i][j] -= L[i][k] * L[j][k];
  }
  if (bad) atomicOr(f.status, 1);
#pragma unroll
  for (int j = 0; j < GS; ++j) {
    W[j][j] = 1.f / L[j][j];
#pragma unroll
    for (int i = j + 1; i < GS; ++i) {
      float acc = 0.f;
#pragma unroll
      for (int k = j; k < i; ++k) acc = fmaf(L[i][k], W[k][j], acc);
      W[i][j] = -acc / L[i][i];
    }
  }
#pragma unroll
  for (int i = 0; i < GS; ++i)
#pragma unroll
    for (int j = 0; j < GS; ++j) f.save_w[gbase + i * GS + j] = W[i][j];
}

// EMA of the running buffers by one thread, domains in order.  With several
// domains the thread that finalizes the LAST domain of group g applies all D updates.
template <int GS>
__device__ __forceinline__ void ema_thread(const Geom& gm, const FwdFin& f, int d_self, int g,
                                           const float (&mean)[GS], const float (&cov)[GS][GS]) {
  if (!f.update_running) return;
  const float m = f.momentum, k = 1.f - f.momentum;
  if (gm.D == 1 || f.aliased == 0) {
    // this domain owns its buffers: update them directly from registers, no cross-CTA traffic
    float* rc = f.rcov[d_self] + (size_t)g * GS * GS;
    float* rm = f.rmean[d_self] + g * GS;
#pragma unroll
    for (int i = 0; i < GS; ++i) {
      rm[i] = m * mean[i] + k * rm[i];
#pragma unroll
      for (int j = 0; j < GS; ++j) rc[i * GS + j] = m * (cov[i][j] * f.unbias) + k * rc[i * GS + j];
    }
    return;
  }
  // shared buffers: the thread that finalizes the LAST domain of group g applies all D updates
  __threadfence();
  const int t = atomicAdd(f.dom_counter + g, 1);
  if (t != gm.D - 1) return;
  atomicExch(f.dom_counter + g, 0);
  __threadfence();
  if (f.aliased == 1) {
    // r' = k^D r + m * sum_d k^(D-1-d) s_d  ==  D sequential updates of one buffer
    float* rc = f.rcov[0] + (size_t)g * GS * GS;
    float* rm = f.rmean[0] + g * GS;
    float c[GS * GS], u[GS];
#pragma unroll
    for (int e = 0; e < GS * GS; ++e) c[e] = rc[e];
#pragma unroll
    for (int e = 0; e < GS; ++e) u[e] = rm[e];
    for (int d = 0; d < gm.D; ++d) {
      const float* cv = f.save_cov + ((size_t)d * gm.G + g) * GS * GS;
      const float* mu = f.save_mean + (size_t)d * gm.C + g * GS;
#pragma unroll
      for (int e = 0; e < GS * GS; ++e) c[e] = m * (__ldcg(cv + e) * f.unbias) + k * c[e];
#pragma unroll
      for (int e = 0; e < GS; ++e) u[e] = m * __ldcg(mu + e) + k * u[e];
    }
#pragma unroll
    for (int e = 0; e < GS * GS; ++e) rc[e] = c[e];
#pragma unroll
    for (int e = 0; e < GS; ++e) rm[e] = u[e];
    return;
  }
  for (int d = 0; d < gm.D; ++d) {       // mixed aliasing: plain ordered read-modify-write
    const float* cv = f.save_cov + ((size_t)d * gm.G + g) * GS * GS;
    const float* mu = f.save_mean + (size_t)d * gm.C + g * GS;
    float* rc = f.rcov[d] + (size_t)g * GS * GS;
    float* rm = f.rmean[d] + g * GS;
    for (int e = 0; e < GS * GS; ++e) rc[e] = m * (__ldcg(cv + e) * f.unbias) + k * rc[e];
    for (int e = 0; e < GS; ++e) rm[e] = m * __ldcg(mu + e) + k * rm[e];
  }
}

template <int GS>
__device__ __forceinline__ void bwd_finalize_thread(const Geom& gm, const BwdFin& f, int d, int g,
                                                    const float (&R)[GS][GS], const float (&sdz)[GS]) {
  const size_t gbase = ((size_t)d * gm.G + g) * GS * GS;
  const int c0 = g * GS;
  const bool affine = (f.epi & DWT_EPI_AFFINE) != 0, train = f.mode == DWT_MODE_TRAIN;
  float W[GS][GS], ga[GS], mu[GS];
#pragma unroll
  for (int i = 0; i < GS; ++i) {
    ga[i] = affine ? f.gamma[c0 + i] : 1.f;
    mu[i] = f.save_mean[(size_t)d * gm.C + c0 + i];
#pragma unroll
    for (int j = 0; j < GS; ++j) W[i][j] = f.save_w[gbase + i * GS + j];
  }
  float* coef = f.coef + ((size_t)d * gm.G + g) * coef_stride(GS);
  if (affine) {
#pragma unroll
    for (int i = 0; i < GS; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j <= i; ++j) s = fmaf(W[i][j], R[i][j], s);
      f.dgb_part[((size_t)d * 2 + 0) * gm.C + c0 + i] = s;
      f.dgb_part[((size_t)d * 2 + 1) * gm.C + c0 + i] = sdz[i];
    }
  }
  float Bm[GS][GS];
#pragma unroll
  for (int i = 0; i < GS; ++i)
#pragma unroll
    for (int j = 0; j < GS; ++j) Bm[i][j] = 0.f;
  if (train) {
    float P[GS][GS], T[GS][GS], S[GS][GS];
#pragma unroll
    for (int i = 0; i < GS; ++i)
#pragma unroll
      for (int j = 0; j < GS; ++j) {
        float q = 0.f;
        if (j <= i) {
#pragma unroll
          for (int k = 0; k <= j; ++k) q = fmaf(R[i][k], W[j][k], q);
          q *= -ga[i] * (i == j ? 0.5f : 1.f);
        }
        P[i][j] = q;
      }
#pragma unroll
    for (int i = 0; i < GS; ++i)
#pragma unroll
      for (int j = 0; j < GS; ++j) {
        float t = 0.f;
#pragma unroll
        for (int k = (i > j ? i : j); k < GS; ++k) t = fmaf(W[k][i], P[k][j], t);
        T[i][j] = t;
      }
#pragma unroll
    for (int i = 0; i < GS; ++i)
#pragma unroll
      for (int j = 0; j < GS; ++j) {
        float s = 0.f;
#pragma unroll
        for (int k = j; k < GS; ++k) s = fmaf(T[i][k], W[k][j], s);
        S[i][j] = s;
      }
    const float sc = f.a / gm.M;
#pragma unroll
    for (int i = 0; i < GS; ++i)
#pragma unroll
      for (int j = 0; j < GS; ++j) Bm[i][j] = sc * (S[i][j] + S[j][i]);
  }
#pragma unroll
  for (int i = 0; i < GS; ++i) {
    float c = 0.f;
#pragma unroll
    for (int j = 0; j < GS; ++j) {
      const float a1 = (j >= i) ? W[j][i] * ga[j] : 0.f;
      coef[i * GS + j] = a1;
      coef[GS * GS + i * GS + j] = Bm[i][j];
      if (train) {
        c = fmaf(a1, sdz[j] / gm.M, c);
        c = fmaf(Bm[i][j], mu[j], c);
      }
    }
    coef[2 * GS * GS + i] = -c;
  }
  if (affine && f.dgamma != nullptr) {
    if (gm.D > 1) {
      __threadfence();
      const int t = atomicAdd(f.dom_counter + g, 1);
      if (t != gm.D - 1) return;
      atomicExch(f.dom_counter + g, 0);
      __threadfence();
    }
#pragma unroll
    for (int i = 0; i < GS; ++i) {
      float sg = 0.f, sb = 0.f;
      for (int dd = 0; dd < gm.D; ++dd) {
        sg += __ldcg(f.dgb_part + ((size_t)dd * 2 + 0) * gm.C + c0 + i);
        sb += __ldcg(f.dgb_part + ((size_t)dd * 2 + 1) * gm.C + c0 + i);
      }
      f.dgamma[c0 + i] = sg;
      f.dbeta[c0 + i] = sb;
    }
  }
}

// Team-level reduction of NACC per-thread accumulators into sAcc[team][NACC] (or, when the
// problem is split over several CTAs, through the global partials + arrival counter).
// Returns true in the threads that should finalize (team thread 0 of a valid problem, and only
// in the last-arriving CTA when nchunks > 1).
template <int NACC>
__device__ __forceinline__ bool team_reduce(const Geom& gm, const Team& tm, int d, const float (&acc)[NACC],
                                            float (*sRed)[NACC], float (*sAcc)[NACC], float* partial,
                                            int* counters, int* sFlag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) sRed[warp][i] = v;
  }
  __syncthreads();
  if (tm.ttid < NACC) {
    float t = 0.f;
    for (int w = 0; w < tm.wpp; ++w) t += sRed[tm.team * tm.wpp + w][tm.ttid];
    sAcc[tm.team][tm.ttid] = t;
  }
  if (gm.nchunks > 1) {      // ppc == 1: one problem per CTA, split over gridDim.x CTAs
    float* prob = partial + ((size_t)d * gm.G + tm.g) * gm.nchunks * NACC;
    if (threadIdx.x < NACC) prob[(size_t)blockIdx.x * NACC + threadIdx.x] = sAcc[0][threadIdx.x];
    if (!arrive_is_last(counters + d * gm.G + tm.g, gm.nchunks, sFlag)) return false;
    if (threadIdx.x < NACC) {
      double s = 0.0;
      for (int c = 0; c < gm.nchunks; ++c) s += (double)__ldcg(prob + (size_t)c * NACC + threadIdx.x);
      sAcc[0][threadIdx.x] = (float)s;
    }
  }
  __syncthreads();
  return tm.valid && tm.ttid == 0;
}

// ------------------------------------------------------------------------------------------
// stats
// ------------------------------------------------------------------------------------------
template <int GS, int VEC>
__global__ void __launch_bounds__(kThreads, (GS * VEC >= 16) ? 3 : 4) small_stats_kernel(const float* __restrict__ x, const Geom gm,
                                                                const FwdFin fin, float* __restrict__ partial,
                                                                int* counters) {
  constexpr int NM = GS * (GS + 1) / 2, NACC = GS + NM, UNROLL = Unroll<GS, VEC>::stats;
  __shared__ float sK[kWarps][GS];
  __shared__ float sRed[kWarps][NACC];
  __shared__ float sAcc[kWarps][NACC];
  __shared__ int sFlag;
  const Team tm(gm);
  const int d = blockIdx.z;
  const int g = tm.valid ? tm.g : gm.G - 1;          // out-of-range teams shadow the last group, results dropped
  const float* xg = x + ((size_t)d * gm.N * gm.C + (size_t)g * GS) * gm.HW;
  const ItemMap map{(unsigned)(gm.HW / VEC), (unsigned)(gm.C * gm.HW)};
  const unsigned items = (unsigned)gm.N * map.PV, stride = gridDim.x * tm.tthreads;
  unsigned i0 = blockIdx.x * tm.tthreads + tm.ttid;

  // Pilot shift (mean of <=32 mid-image pixels of image 0, per channel) overlapped with the
  // first batch of loads: the loads do not depend on K, only the arithmetic does.
  float pv[GS];
  {
    const int np = gm.HW < 32 ? gm.HW : 32, p0 = ((gm.HW - np) / 2) & ~3, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < GS; ++c) pv[c] = (tm.ttid < 32 && lane < np) ? __ldg(xg + (size_t)c * gm.HW + p0 + lane) : 0.f;
  }
  float v[UNROLL][GS][VEC];
  bool have[UNROLL];
  auto load_batch = [&](unsigned base) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned it = base + u * stride;
      have[u] = it < items;
      if (have[u]) {
        const float* p = xg + map.offset(it, VEC);
#pragma unroll
        for (int c = 0; c < GS; ++c) load_vec<VEC>(p + (size_t)c * gm.HW, v[u][c]);
      }
    }
  };
  if (i0 < items) load_batch(i0);
  if (tm.ttid < 32) {
    const int np = gm.HW < 32 ? gm.HW : 32;
#pragma unroll
    for (int c = 0; c < GS; ++c) {
      const float t = warp_sum(pv[c]);
      if (tm.ttid == 0) sK[tm.team][c] = t / (float)np;
    }
  }
  __syncthreads();
  float K[GS], acc[NACC];
#pragma unroll
  for (int c = 0; c < GS; ++c) K[c] = sK[tm.team][c];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
  while (i0 < items) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (have[u]) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          float xs[GS];
#pragma unroll
          for (int c = 0; c < GS; ++c) { xs[c] = v[u][c][e] - K[c]; acc[c] += xs[c]; }
#pragma unroll
          for (int c = 0; c < GS; ++c)
#pragma unroll
            for (int j = 0; j <= c; ++j) acc[GS + c * (c + 1) / 2 + j] = fmaf(xs[c], xs[j], acc[GS + c * (c + 1) / 2 + j]);
        }
      }
    }
    i0 += stride * UNROLL;
    if (i0 < items) load_batch(i0);
  }
  if (!team_reduce<NACC>(gm, tm, d, acc, sRed, sAcc, partial, counters, &sFlag)) return;

  const float invM = 1.f / gm.M;
  const float* a = sAcc[tm.team];
  float mean[GS], cov[GS][GS];
#pragma unroll
  for (int i = 0; i < GS; ++i) mean[i] = K[i] + a[i] * invM;
#pragma unroll
  for (int i = 0; i < GS; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      const float cij = a[GS + i * (i + 1) / 2 + j] * invM - (a[i] * invM) * (a[j] * invM);
      cov[i][j] = cij; cov[j][i] = cij;
    }
  factor_thread<GS>(gm, fin, d, tm.g, mean, cov, true);
  ema_thread<GS>(gm, fin, d, tm.g, mean, cov);
}

// Eval mode: W and mean straight from the running buffers (whitening.py:42-43,50-53).
template <int GS>
__global__ void __launch_bounds__(kThreads) small_eval_prep_kernel(const Geom gm, const FwdFin fin) {
  const int g = blockIdx.x * kThreads + threadIdx.x, d = blockIdx.z;
  if (g >= gm.G) return;
  float mean[GS], cov[GS][GS];
#pragma unroll
  for (int i = 0; i < GS; ++i) {
    mean[i] = fin.rmean[d][g * GS + i];
#pragma unroll
    for (int j = 0; j < GS; ++j) cov[i][j] = fin.rcov[d][(size_t)g * GS * GS + i * GS + j];
  }
  factor_thread<GS>(gm, fin, d, g, mean, cov, false);
}

// ------------------------------------------------------------------------------------------
// apply
// ------------------------------------------------------------------------------------------
template <int GS, int VEC, int EPI>
__global__ void __launch_bounds__(kThreads, (GS * VEC >= 16) ? 3 : 4) small_apply_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                const Geom gm, const float* __restrict__ save_mean,
                                                                const float* __restrict__ save_w,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta) {
  constexpr int NM = GS * (GS + 1) / 2, UNROLL = Unroll<GS, VEC>::stats;
  const Team tm(gm);
  if (!tm.valid) return;
  const int g = tm.g, d = blockIdx.z;
  float Wp[NM], bp[GS];
  load_forward_map<GS, EPI>(save_w + ((size_t)d * gm.G + g) * GS * GS, save_mean + (size_t)d * gm.C + g * GS,
                            gamma + g * GS, beta + g * GS, Wp, bp);
  const size_t base = ((size_t)d * gm.N * gm.C + (size_t)g * GS) * gm.HW;
  const float* xg = x + base;
  float* yg = y + base;
  const ItemMap map{(unsigned)(gm.HW / VEC), (unsigned)(gm.C * gm.HW)};
  const unsigned items = (unsigned)gm.N * map.PV, stride = gridDim.x * tm.tthreads;
  for (unsigned i0 = blockIdx.x * tm.tthreads + tm.ttid; i0 < items; i0 += stride * UNROLL) {
    float v[UNROLL][GS][VEC];
    unsigned off[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned it = i0 + u * stride;
      if (it < items) {
        off[u] = map.offset(it, VEC);
#pragma unroll
        for (int c = 0; c < GS; ++c) load_vec<VEC>(xg + off[u] + (size_t)c * gm.HW, v[u][c]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned it = i0 + u * stride;
      if (it < items) {
        float o[GS][VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          float xi[GS], oi[GS];
#pragma unroll
          for (int c = 0; c < GS; ++c) xi[c] = v[u][c][e];
          apply_group<GS>(Wp, bp, xi, oi);
#pragma unroll
          for (int c = 0; c < GS; ++c) o[c][e] = (EPI & DWT_EPI_RELU) ? fmaxf(oi[c], 0.f) : oi[c];
        }
#pragma unroll
        for (int c = 0; c < GS; ++c) store_vec<VEC>(yg + off[u] + (size_t)c * gm.HW, o[c]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward reduce
// ------------------------------------------------------------------------------------------
template <int GS, int VEC, int EPI>
__global__ void __launch_bounds__(kThreads, 3) small_bwd_reduce_kernel(const float* __restrict__ x,
                                                                     const float* __restrict__ dout, const Geom gm,
                                                                     const BwdFin fin, const float* __restrict__ beta,
                                                                     float* __restrict__ partial, int* counters) {
  constexpr int NM = GS * (GS + 1) / 2, NACC = GS * GS + GS, UNROLL = Unroll<GS, VEC>::one;
  constexpr bool RELU = (EPI & DWT_EPI_RELU) != 0;
  __shared__ float sRed[kWarps][NACC];
  __shared__ float sAcc[kWarps][NACC];
  __shared__ int sFlag;
  const Team tm(gm);
  const int d = blockIdx.z;
  const int g = tm.valid ? tm.g : gm.G - 1;
  float Wp[NM], bp[GS], mu[GS];
  const float* mean_g = fin.save_mean + (size_t)d * gm.C + g * GS;
  if constexpr (RELU)
    load_forward_map<GS, EPI>(fin.save_w + ((size_t)d * gm.G + g) * GS * GS, mean_g, fin.gamma + g * GS,
                              beta + g * GS, Wp, bp);
#pragma unroll
  for (int c = 0; c < GS; ++c) mu[c] = __ldg(mean_g + c);
  float acc[NACC];          // R[i][j] at i*GS+j, then sdz
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
  const size_t base = ((size_t)d * gm.N * gm.C + (size_t)g * GS) * gm.HW;
  const float* xg = x + base;
  const float* gg = dout + base;
  const ItemMap map{(unsigned)(gm.HW / VEC), (unsigned)(gm.C * gm.HW)};
  const unsigned items = (unsigned)gm.N * map.PV, stride = gridDim.x * tm.tthreads;
  for (unsigned i0 = blockIdx.x * tm.tthreads + tm.ttid; i0 < items; i0 += stride * UNROLL) {
    float v[UNROLL][GS][VEC], q[UNROLL][GS][VEC];
    bool have[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned it = i0 + u * stride;
      have[u] = it < items;
      if (have[u]) {
        const unsigned off = map.offset(it, VEC);
#pragma unroll
        for (int c = 0; c < GS; ++c) {
          load_vec<VEC>(xg + off + (size_t)c * gm.HW, v[u][c]);
          load_vec<VEC>(gg + off + (size_t)c * gm.HW, q[u][c]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (have[u]) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          float xi[GS], dz[GS];
#pragma unroll
          for (int c = 0; c < GS; ++c) { xi[c] = v[u][c][e]; dz[c] = q[u][c][e]; }
          if constexpr (RELU) {
            float oi[GS];
            apply_group<GS>(Wp, bp, xi, oi);
#pragma unroll
            for (int c = 0; c < GS; ++c) dz[c] = oi[c] > 0.f ? dz[c] : 0.f;
          }
#pragma unroll
          for (int i = 0; i < GS; ++i) {
            acc[GS * GS + i] += dz[i];
#pragma unroll
            for (int j = 0; j < GS; ++j) acc[i * GS + j] = fmaf(dz[i], xi[j] - mu[j], acc[i * GS + j]);
          }
        }
      }
    }
  }
  if (!team_reduce<NACC>(gm, tm, d, acc, sRed, sAcc, partial, counters, &sFlag)) return;
  const float* a = sAcc[tm.team];
  float R[GS][GS], sdz[GS];
#pragma unroll
  for (int i = 0; i < GS; ++i) {
    sdz[i] = a[GS * GS + i];
#pragma unroll
    for (int j = 0; j < GS; ++j) R[i][j] = a[i * GS + j];
  }
  bwd_finalize_thread<GS>(gm, fin, d, tm.g, R, sdz);
}

// Backward coefficients when no reduction is needed (eval mode, no affine gradient): A1 = W^T diag(gamma).
template <int GS>
__global__ void __launch_bounds__(kThreads) small_bwd_prep_kernel(const Geom gm, const BwdFin fin) {
  const int g = blockIdx.x * kThreads + threadIdx.x, d = blockIdx.z;
  if (g >= gm.G) return;
  float R[GS][GS], sdz[GS];
#pragma unroll
  for (int i = 0; i < GS; ++i) {
    sdz[i] = 0.f;
#pragma unroll
    for (int j = 0; j < GS; ++j) R[i][j] = 0.f;
  }
  bwd_finalize_thread<GS>(gm, fin, d, g, R, sdz);
}

// ------------------------------------------------------------------------------------------
// backward apply
// ------------------------------------------------------------------------------------------
template <int GS, int VEC, int EPI>
__global__ void __launch_bounds__(kThreads, 3) small_bwd_apply_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ dout,
                                                                    float* __restrict__ dx, const Geom gm,
                                                                    const float* __restrict__ coef,
                                                                    const float* __restrict__ save_mean,
                                                                    const float* __restrict__ save_w,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta) {
  constexpr int NM = GS * (GS + 1) / 2, UNROLL = Unroll<GS, VEC>::one;
  constexpr bool RELU = (EPI & DWT_EPI_RELU) != 0;
  const Team tm(gm);
  if (!tm.valid) return;
  const int g = tm.g, d = blockIdx.z;
  float Wp[NM], bp[GS];
  if constexpr (RELU)
    load_forward_map<GS, EPI>(save_w + ((size_t)d * gm.G + g) * GS * GS, save_mean + (size_t)d * gm.C + g * GS,
                              gamma + g * GS, beta + g * GS, Wp, bp);
  // A1 upper-triangular (packed by rows), Bm symmetric (packed lower), cvec
  float A1[NM], Bm[NM], cv[GS];
  const float* cf = coef + ((size_t)d * gm.G + g) * coef_stride(GS);
#pragma unroll
  for (int i = 0; i < GS; ++i) {
    cv[i] = __ldg(cf + 2 * GS * GS + i);
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      A1[i * (i + 1) / 2 + j] = __ldg(cf + j * GS + i);             // A1[j][i], j <= i
      Bm[i * (i + 1) / 2 + j] = __ldg(cf + GS * GS + i * GS + j);   // Bm[i][j] = Bm[j][i]
    }
  }
  const size_t base = ((size_t)d * gm.N * gm.C + (size_t)g * GS) * gm.HW;
  const float* xg = x + base;
  const float* gg = dout + base;
  float* dg = dx + base;
  const ItemMap map{(unsigned)(gm.HW / VEC), (unsigned)(gm.C * gm.HW)};
  const unsigned items = (unsigned)gm.N * map.PV, stride = gridDim.x * tm.tthreads;
  for (unsigned i0 = blockIdx.x * tm.tthreads + tm.ttid; i0 < items; i0 += stride * UNROLL) {
    float v[UNROLL][GS][VEC], q[UNROLL][GS][VEC];
    unsigned off[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned it = i0 + u * stride;
      if (it < items) {
        off[u] = map.offset(it, VEC);
#pragma unroll
        for (int c = 0; c < GS; ++c) {
          load_vec<VEC>(xg + off[u] + (size_t)c * gm.HW, v[u][c]);
          load_vec<VEC>(gg + off[u] + (size_t)c * gm.HW, q[u][c]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned it = i0 + u * stride;
      if (it < items) {
        float o[GS][VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          float xi[GS], dz[GS];
#pragma unroll
          for (int c = 0; c < GS; ++c) { xi[c] = v[u][c][e]; dz[c] = q[u][c][e]; }
          if constexpr (RELU) {
            float oi[GS];
            apply_group<GS>(Wp, bp, xi, oi);
#pragma unroll
            for (int c = 0; c < GS; ++c) dz[c] = oi[c] > 0.f ? dz[c] : 0.f;
          }
#pragma unroll
          for (int i = 0; i < GS; ++i) {
            float acc = cv[i];
#pragma unroll
            for (int j = i; j < GS; ++j) acc = fmaf(A1[j * (j + 1) / 2 + i], dz[j], acc);   // A1[i][j], j >= i
#pragma unroll
            for (int j = 0; j < GS; ++j) {
              const int hi = i > j ? i : j, lo = i > j ? j : i;
              acc = fmaf(Bm[hi * (hi + 1) / 2 + lo], xi[j], acc);
            }
            o[i][e] = acc;
          }
        }
#pragma unroll
        for (int c = 0; c < GS; ++c) store_vec<VEC>(dg + off[u] + (size_t)c * gm.HW, o[c]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// launch tables
// ------------------------------------------------------------------------------------------
inline dim3 grid_of(const Geom& gm, int chunks) { return dim3(chunks, (gm.G + gm.ppc - 1) / gm.ppc, gm.D); }
inline dim3 grid_prep(const Geom& gm) { return dim3((gm.G + kThreads - 1) / kThreads, 1, gm.D); }

#define DWT_DISPATCH_GS(GS_, ...)                      \
  switch (GS_) {                                       \
    case 1: { constexpr int kGS = 1; __VA_ARGS__; break; } \
    case 2: { constexpr int kGS = 2; __VA_ARGS__; break; } \
    case 4: { constexpr int kGS = 4; __VA_ARGS__; break; } \
    default: break;                                    \
  }
#define DWT_DISPATCH_VEC(V_, ...)                      \
  if ((V_) == 4) { constexpr int kVEC = 4; __VA_ARGS__; } else { constexpr int kVEC = 1; __VA_ARGS__; }
#define DWT_DISPATCH_EPI(E_, ...)                                         \
  if ((E_) == 3) { constexpr int kEPI = 3; __VA_ARGS__; }                 \
  else if ((E_) == 1) { constexpr int kEPI = 1; __VA_ARGS__; }            \
  else { constexpr int kEPI = 0; __VA_ARGS__; }

}  // namespace

bool small_supports(int GS) { return GS == 1 || GS == 2 || GS == 4; }

void small_stats(const float* x, const Geom& gm, int vec, const FwdFin& fin, float* partial, int* counters,
                 cudaStream_t st) {
  DWT_DISPATCH_GS(gm.GS, DWT_DISPATCH_VEC(vec, (small_stats_kernel<kGS, kVEC><<<grid_of(gm, gm.nchunks), kThreads, 0, st>>>(
                                                   x, gm, fin, partial, counters))));
}

void small_eval_prep(const Geom& gm, const FwdFin& fin, cudaStream_t st) {
  DWT_DISPATCH_GS(gm.GS, (small_eval_prep_kernel<kGS><<<grid_prep(gm), kThreads, 0, st>>>(gm, fin)));
}

void small_apply(const float* x, float* y, const Geom& gm, int vec, int chunks, int epi, const float* mean,
                 const float* w, const float* gamma, const float* beta, cudaStream_t st) {
  DWT_DISPATCH_GS(gm.GS, DWT_DISPATCH_VEC(vec, DWT_DISPATCH_EPI(epi, (small_apply_kernel<kGS, kVEC, kEPI><<<grid_of(gm, chunks), kThreads, 0, st>>>(
                                                                         x, y, gm, mean, w, gamma, beta)))));
}

void small_bwd_reduce(const float* x, const float* dout, const Geom& gm, int vec, const BwdFin& fin,
                      const float* beta, float* partial, int* counters, cudaStream_t st) {
  DWT_DISPATCH_GS(gm.GS, DWT_DISPATCH_VEC(vec, DWT_DISPATCH_EPI(fin.epi, (small_bwd_reduce_kernel<kGS, kVEC, kEPI><<<grid_of(gm, gm.nchunks), kThreads, 0, st>>>(
                                                                             x, dout, gm, fin, beta, partial, counters)))));
}

void small_bwd_prep(const Geom& gm, const BwdFin& fin, cudaStream_t st) {
  DWT_DISPATCH_GS(gm.GS, (small_bwd_prep_kernel<kGS><<<grid_prep(gm), kThreads, 0, st>>>(gm, fin)));
}

void small_bwd_apply(const float* x, const float* dout, float* dx, const Geom& gm, int vec, int chunks, int epi,
                     const float* coef, const float* mean, const float* w, const float* gamma, const float* beta,
                     cudaStream_t st) {
  DWT_DISPATCH_GS(gm.GS, DWT_DISPATCH_VEC(vec, DWT_DISPATCH_EPI(epi, (small_bwd_apply_kernel<kGS, kVEC, kEPI><<<grid_of(gm, chunks), kThreads, 0, st>>>(
                                                                         x, dout, dx, gm, coef, mean, w, gamma, beta)))));
}

}  // namespace dwt
