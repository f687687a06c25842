// Register-resident path for small channel groups (GS = 1, 2, 4): domain batch norm
// (GS = 1) and the whitening layers of the shipped ResNet-50-DWT / LeNet (GS = 4).
//
// Each thread owns VEC consecutive pixels of all GS channels of one group, so a warp
// reads GS fully coalesced row segments; every kernel is one HBM pass:
//   stats       read x                    -> per-CTA shifted moments -> last CTA: mean, cov, W, EMA
//   apply       read x, write y           y = W (x - mean) [* gamma + beta] [relu]
//   bwd_reduce  read x, dout              -> R = sum dz xc^T, sum dz -> last CTA: A1, Bm, cvec, dgamma, dbeta
//   bwd_apply   read x, dout, write dx    dx = A1 dz + Bm x + cvec
// i.e. 12 B/element forward + 20 B/element backward, the algorithmic minimum of SURVEY.md §8d.
//
// Reference: utils/whitening.py:37-61, utils/batch_norm.py:54-69 (/root/reference).
#include "dwt_common.cuh"
#include "norm_launch.h"

namespace dwt {
namespace {

template <int VEC> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<4> { using type = float4; };

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

// item -> (image n, pixel-vector pv) of the flattened per-group work list
struct ItemMap {
  unsigned PV;       // pixel vectors per row
  size_t img_stride; // C*HW
  __device__ __forceinline__ size_t offset(unsigned item, int VEC) const {
    unsigned n = item / PV, pv = item - n * PV;
    return (size_t)n * img_stride + (size_t)pv * VEC;
  }
};

// ------------------------------------------------------------------------------------------
// stats
// ------------------------------------------------------------------------------------------
template <int GS, int VEC>
__global__ void __launch_bounds__(kThreads) small_stats_kernel(const float* __restrict__ x, Geom gm, FwdFin fin,
                                                                float* __restrict__ partial, int* counters) {
  constexpr int NM = GS * (GS + 1) / 2, NACC = GS + NM, UNROLL = (GS * VEC >= 16) ? 2 : 4;
  constexpr int LD = GS + 1;
  __shared__ float sK[GS];
  __shared__ float sRed[kWarps][NACC];
  __shared__ float sAcc[NACC];
  __shared__ float sMean[GS], sCov[GS * LD], sL[GS * LD], sW[GS * LD];
  __shared__ int sFlag;
  const int g = blockIdx.y, d = blockIdx.z, tid = threadIdx.x;
  const float* xg = x + ((size_t)d * gm.N * gm.C + (size_t)g * GS) * gm.HW;
  pilot_shift(xg, GS, gm.HW, sK);
  __syncthreads();
  float K[GS], s[GS], m[NM];
#pragma unroll
  for (int c = 0; c < GS; ++c) { K[c] = sK[c]; s[c] = 0.f; }
#pragma unroll
  for (int i = 0; i < NM; ++i) m[i] = 0.f;

  const ItemMap map{(unsigned)(gm.HW / VEC), (size_t)gm.C * gm.HW};
  const unsigned items = (unsigned)gm.N * map.PV, stride = gridDim.x * kThreads;
  for (unsigned i0 = blockIdx.x * kThreads + tid; i0 < items; i0 += stride * UNROLL) {
    float v[UNROLL][GS][VEC];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned it = i0 + u * stride;
      if (it < items) {
        const float* p = xg + map.offset(it, VEC);
#pragma unroll
        for (int c = 0; c < GS; ++c) load_vec<VEC>(p + (size_t)c * gm.HW, v[u][c]);
      } else {
#pragma unroll
        for (int c = 0; c < GS; ++c)
#pragma unroll
          for (int e = 0; e < VEC; ++e) v[u][c][e] = K[c];
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        float xs[GS];
#pragma unroll
        for (int c = 0; c < GS; ++c) { xs[c] = v[u][c][e] - K[c]; s[c] += xs[c]; }
#pragma unroll
        for (int c = 0; c < GS; ++c)
#pragma unroll
          for (int j = 0; j <= c; ++j) m[c * (c + 1) / 2 + j] = fmaf(xs[c], xs[j], m[c * (c + 1) / 2 + j]);
      }
  }
  // CTA reduction -> one partial row
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float v = warp_sum(i < GS ? s[i] : m[i - GS]);
    if (lane == 0) sRed[warp][i] = v;
  }
  __syncthreads();
  float* prow = partial + (((size_t)d * gm.G + g) * gm.nchunks + blockIdx.x) * NACC;
  if (tid < NACC) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += sRed[w][tid];
    prow[tid] = t;
  }
  if (!arrive_is_last(counters + d * gm.G + g, gm.nchunks, &sFlag)) return;

  reduce_partials(partial + ((size_t)d * gm.G + g) * gm.nchunks * NACC, gm.nchunks, NACC, sAcc);
  __syncthreads();
  const float invM = 1.f / gm.M;
  if (tid < GS) sMean[tid] = sK[tid] + sAcc[tid] * invM;
  if (tid < GS * GS) {
    const int i = tid / GS, j = tid % GS, hi = i > j ? i : j, lo = i > j ? j : i;
    sCov[i * LD + j] = sAcc[GS + hi * (hi + 1) / 2 + lo] * invM - (sAcc[i] * invM) * (sAcc[j] * invM);
  }
  __syncthreads();
  fwd_factor_block(gm, fin, d, g, sMean, sCov, sL, sW, true);
  fwd_ema_block(gm, fin, g, &sFlag);
}

// Eval mode: W and mean straight from the running buffers (whitening.py:42-43,50-53).
template <int GS>
__global__ void __launch_bounds__(kThreads) small_eval_prep_kernel(Geom gm, FwdFin fin) {
  constexpr int LD = GS + 1;
  __shared__ float sMean[GS], sCov[GS * LD], sL[GS * LD], sW[GS * LD];
  const int g = blockIdx.y, d = blockIdx.z, tid = threadIdx.x;
  if (tid < GS) sMean[tid] = fin.rmean[d][g * GS + tid];
  if (tid < GS * GS) sCov[(tid / GS) * LD + tid % GS] = fin.rcov[d][(size_t)g * GS * GS + tid];
  __syncthreads();
  fwd_factor_block(gm, fin, d, g, sMean, sCov, sL, sW, false);
}

// ------------------------------------------------------------------------------------------
// apply
// ------------------------------------------------------------------------------------------
template <int GS, int VEC, int EPI>
__global__ void __launch_bounds__(kThreads) small_apply_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                Geom gm, const float* __restrict__ save_mean,
                                                                const float* __restrict__ save_w,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta) {
  constexpr int NM = GS * (GS + 1) / 2, UNROLL = (GS * VEC >= 16) ? 2 : 4;
  const int g = blockIdx.y, d = blockIdx.z, tid = threadIdx.x;
  float Wp[NM], bp[GS];
  load_forward_map<GS, EPI>(save_w + ((size_t)d * gm.G + g) * GS * GS, save_mean + (size_t)d * gm.C + g * GS,
                            gamma + g * GS, beta + g * GS, Wp, bp);
  const size_t base = ((size_t)d * gm.N * gm.C + (size_t)g * GS) * gm.HW;
  const float* xg = x + base;
  float* yg = y + base;
  const ItemMap map{(unsigned)(gm.HW / VEC), (size_t)gm.C * gm.HW};
  const unsigned items = (unsigned)gm.N * map.PV, stride = gridDim.x * kThreads;
  for (unsigned i0 = blockIdx.x * kThreads + tid; i0 < items; i0 += stride * UNROLL) {
    float v[UNROLL][GS][VEC];
    size_t off[UNROLL];
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
__global__ void __launch_bounds__(kThreads) small_bwd_reduce_kernel(const float* __restrict__ x,
                                                                     const float* __restrict__ dout, Geom gm,
                                                                     BwdFin fin, const float* __restrict__ beta,
                                                                     float* __restrict__ partial, int* counters) {
  constexpr int NM = GS * (GS + 1) / 2, NACC = GS * GS + GS, LD = GS + 1;
  constexpr bool RELU = (EPI & DWT_EPI_RELU) != 0;
  __shared__ float sRed[kWarps][NACC];
  __shared__ float sAcc[NACC];
  __shared__ float sR[GS * LD], sSdz[GS], sW[GS * LD], sT1[GS * LD], sT2[GS * LD], sVec[3 * GS];
  __shared__ int sFlag;
  const int g = blockIdx.y, d = blockIdx.z, tid = threadIdx.x;
  float Wp[NM], bp[GS], mu[GS];
  const float* mean_g = fin.save_mean + (size_t)d * gm.C + g * GS;
  if constexpr (RELU)
    load_forward_map<GS, EPI>(fin.save_w + ((size_t)d * gm.G + g) * GS * GS, mean_g, fin.gamma + g * GS,
                              beta + g * GS, Wp, bp);
#pragma unroll
  for (int c = 0; c < GS; ++c) mu[c] = __ldg(mean_g + c);
  float R[GS][GS], sdz[GS];
#pragma unroll
  for (int i = 0; i < GS; ++i) {
    sdz[i] = 0.f;
#pragma unroll
    for (int j = 0; j < GS; ++j) R[i][j] = 0.f;
  }
  const size_t base = ((size_t)d * gm.N * gm.C + (size_t)g * GS) * gm.HW;
  const float* xg = x + base;
  const float* gg = dout + base;
  const ItemMap map{(unsigned)(gm.HW / VEC), (size_t)gm.C * gm.HW};
  const unsigned items = (unsigned)gm.N * map.PV, stride = gridDim.x * kThreads;
  constexpr int UNROLL = (GS * VEC >= 16) ? 1 : 2;
  for (unsigned i0 = blockIdx.x * kThreads + tid; i0 < items; i0 += stride * UNROLL) {
    float v[UNROLL][GS][VEC], q[UNROLL][GS][VEC];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned it = i0 + u * stride;
      if (it < items) {
        const size_t off = map.offset(it, VEC);
#pragma unroll
        for (int c = 0; c < GS; ++c) {
          load_vec<VEC>(xg + off + (size_t)c * gm.HW, v[u][c]);
          load_vec<VEC>(gg + off + (size_t)c * gm.HW, q[u][c]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < GS; ++c)
#pragma unroll
          for (int e = 0; e < VEC; ++e) { v[u][c][e] = mu[c]; q[u][c][e] = 0.f; }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
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
          sdz[i] += dz[i];
#pragma unroll
          for (int j = 0; j < GS; ++j) R[i][j] = fmaf(dz[i], xi[j] - mu[j], R[i][j]);
        }
      }
  }
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float t = warp_sum(i < GS * GS ? R[i / GS][i % GS] : sdz[i - GS * GS]);
    if (lane == 0) sRed[warp][i] = t;
  }
  __syncthreads();
  float* prow = partial + (((size_t)d * gm.G + g) * gm.nchunks + blockIdx.x) * NACC;
  if (tid < NACC) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += sRed[w][tid];
    prow[tid] = t;
  }
  if (!arrive_is_last(counters + d * gm.G + g, gm.nchunks, &sFlag)) return;
  reduce_partials(partial + ((size_t)d * gm.G + g) * gm.nchunks * NACC, gm.nchunks, NACC, sAcc);
  __syncthreads();
  if (tid < GS * GS) sR[(tid / GS) * LD + tid % GS] = sAcc[tid];
  if (tid < GS) sSdz[tid] = sAcc[GS * GS + tid];
  __syncthreads();
  bwd_finalize_block(gm, fin, d, g, sR, sSdz, sW, sT1, sT2, sVec, &sFlag);
}

// Backward coefficients when no reduction is needed (eval mode, no affine): A1 = W^T.
template <int GS>
__global__ void __launch_bounds__(kThreads) small_bwd_prep_kernel(Geom gm, BwdFin fin) {
  constexpr int LD = GS + 1;
  __shared__ float sR[GS * LD], sSdz[GS], sW[GS * LD], sT1[GS * LD], sT2[GS * LD], sVec[3 * GS];
  __shared__ int sFlag;
  if (threadIdx.x < GS * LD) sR[threadIdx.x] = 0.f;
  if (threadIdx.x < GS) sSdz[threadIdx.x] = 0.f;
  __syncthreads();
  bwd_finalize_block(gm, fin, blockIdx.z, blockIdx.y, sR, sSdz, sW, sT1, sT2, sVec, &sFlag);
}

// ------------------------------------------------------------------------------------------
// backward apply
// ------------------------------------------------------------------------------------------
template <int GS, int VEC, int EPI>
__global__ void __launch_bounds__(kThreads) small_bwd_apply_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ dout,
                                                                    float* __restrict__ dx, Geom gm,
                                                                    const float* __restrict__ coef,
                                                                    const float* __restrict__ save_mean,
                                                                    const float* __restrict__ save_w,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta) {
  constexpr int NM = GS * (GS + 1) / 2;
  constexpr bool RELU = (EPI & DWT_EPI_RELU) != 0;
  const int g = blockIdx.y, d = blockIdx.z, tid = threadIdx.x;
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
  const ItemMap map{(unsigned)(gm.HW / VEC), (size_t)gm.C * gm.HW};
  const unsigned items = (unsigned)gm.N * map.PV, stride = gridDim.x * kThreads;
  constexpr int UNROLL = (GS * VEC >= 16) ? 1 : 2;
  for (unsigned i0 = blockIdx.x * kThreads + tid; i0 < items; i0 += stride * UNROLL) {
    float v[UNROLL][GS][VEC], q[UNROLL][GS][VEC];
    size_t off[UNROLL];
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
inline dim3 grid_of(const Geom& gm, int chunks) { return dim3(chunks, gm.G, gm.D); }

template <int GS, int VEC>
void launch_stats(const float* x, const Geom& gm, const FwdFin& fin, float* partial, int* counters,
                  cudaStream_t st) {
  small_stats_kernel<GS, VEC><<<grid_of(gm, gm.nchunks), kThreads, 0, st>>>(x, gm, fin, partial, counters);
}

template <int GS, int VEC, int EPI>
void launch_apply(const float* x, float* y, const Geom& gm, int chunks, const float* mean, const float* w,
                  const float* gamma, const float* beta, cudaStream_t st) {
  small_apply_kernel<GS, VEC, EPI><<<grid_of(gm, chunks), kThreads, 0, st>>>(x, y, gm, mean, w, gamma, beta);
}

template <int GS, int VEC, int EPI>
void launch_bwd_reduce(const float* x, const float* dout, const Geom& gm, const BwdFin& fin, const float* beta,
                       float* partial, int* counters, cudaStream_t st) {
  small_bwd_reduce_kernel<GS, VEC, EPI><<<grid_of(gm, gm.nchunks), kThreads, 0, st>>>(x, dout, gm, fin, beta,
                                                                                      partial, counters);
}

template <int GS, int VEC, int EPI>
void launch_bwd_apply(const float* x, const float* dout, float* dx, const Geom& gm, int chunks, const float* coef,
                      const float* mean, const float* w, const float* gamma, const float* beta, cudaStream_t st) {
  small_bwd_apply_kernel<GS, VEC, EPI><<<grid_of(gm, chunks), kThreads, 0, st>>>(x, dout, dx, gm, coef, mean, w,
                                                                                 gamma, beta);
}

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
  DWT_DISPATCH_GS(gm.GS, DWT_DISPATCH_VEC(vec, (launch_stats<kGS, kVEC>(x, gm, fin, partial, counters, st))));
}

void small_eval_prep(const Geom& gm, const FwdFin& fin, cudaStream_t st) {
  DWT_DISPATCH_GS(gm.GS, (small_eval_prep_kernel<kGS><<<dim3(1, gm.G, gm.D), kThreads, 0, st>>>(gm, fin)));
}

void small_apply(const float* x, float* y, const Geom& gm, int vec, int chunks, int epi, const float* mean,
                 const float* w, const float* gamma, const float* beta, cudaStream_t st) {
  DWT_DISPATCH_GS(gm.GS, DWT_DISPATCH_VEC(vec, DWT_DISPATCH_EPI(epi, (launch_apply<kGS, kVEC, kEPI>(
                                                                         x, y, gm, chunks, mean, w, gamma, beta, st)))));
}

void small_bwd_reduce(const float* x, const float* dout, const Geom& gm, int vec, const BwdFin& fin,
                      const float* beta, float* partial, int* counters, cudaStream_t st) {
  DWT_DISPATCH_GS(gm.GS, DWT_DISPATCH_VEC(vec, DWT_DISPATCH_EPI(fin.epi, (launch_bwd_reduce<kGS, kVEC, kEPI>(
                                                                             x, dout, gm, fin, beta, partial, counters, st)))));
}

void small_bwd_prep(const Geom& gm, const BwdFin& fin, cudaStream_t st) {
  DWT_DISPATCH_GS(gm.GS, (small_bwd_prep_kernel<kGS><<<dim3(1, gm.G, gm.D), kThreads, 0, st>>>(gm, fin)));
}

void small_bwd_apply(const float* x, const float* dout, float* dx, const Geom& gm, int vec, int chunks, int epi,
                     const float* coef, const float* mean, const float* w, const float* gamma, const float* beta,
                     cudaStream_t st) {
  DWT_DISPATCH_GS(gm.GS, DWT_DISPATCH_VEC(vec, DWT_DISPATCH_EPI(epi, (launch_bwd_apply<kGS, kVEC, kEPI>(
                                                                         x, dout, dx, gm, chunks, coef, mean, w, gamma, beta, st)))));
}

}  // namespace dwt
