// Channels-last max-pool, forward + backward: the op that sits between the stem whitening site and layer1
// (`x = self.maxpool(self.relu(...))`, resnet50_dwt_mec_officehome.py:337-338; nn.MaxPool2d(3, 2, 1) at :295).
//
// Why it is here: on B200 the stock ATen kernels for this op in NHWC fp32 (max_pool_forward_nhwc 0.65 ms,
// max_pool_backward_nhwc 1.31 ms, int64 argmax map of 308 MB) cost 5 % of the whole ResNet-50-DWT step
// (profiles/launches_r02_step.md) for 1.1 GB of algorithmic traffic = 0.17 ms at the HBM peak.  Both passes here
// are pure gathers -- no atomics, deterministic:
//   forward   one CTA per output row, one thread per output float4 (4 channels of one output pixel): k*k coalesced
//             float4 loads (a warp covers whole pixels: C/4 consecutive float4), max with ATen's rule (strict >, NaN wins, first maximum
//             in row-major window order keeps the index), one float4 store + 4 index BYTES (window-local k*k index);
//   backward  one CTA per input row, one thread per INPUT float4: the <= ceil(k/s)^2 windows that contain the pixel
//             are visited, and a window's gradient is taken iff its saved index names this pixel.
// Algorithmic bytes: forward 4*(in + out) + out, backward 4*(in + out) + out  (in, out = element counts).
//
// Semantics = torch.nn.functional.max_pool2d(x, k, s, p) (dilation 1, ceil_mode False) and its autograd, bit for bit
// including the tie rule (post-ReLU windows of all zeros are common: the gradient goes to the first element).
#include "dwt_common.cuh"
#include "norm_launch.h"

namespace dwt {
namespace {

struct PoolGeom {
  int N, H, W, C4, OH, OW, k, s, p, c4shift;
};

// K > 0: the window is a compile-time constant (nn.MaxPool2d(3, 2, 1) of the ResNet stem): the loops over the window unroll
// and the divisions by the stride become shifts -- the generic version spent more time on index arithmetic than on memory
template <int K, int S, int P>
__device__ __forceinline__ PoolGeom fixed(PoolGeom g) {
  if constexpr (K > 0) { g.k = K; g.s = S; g.p = P; }
  return g;
}

__device__ __forceinline__ bool takes(float v, float best) { return (v > best) || (v != v); }

// One CTA per output row (n, oh): the W*C4 float4 of up to k input rows are re-read from L1/L2 by neighbouring
// windows; threads run over (ow, c4) with 32-bit index arithmetic only.
template <int K, int S, int P, bool POW2>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          uint8_t* __restrict__ idx, const PoolGeom g0) {
  const PoolGeom g = fixed<K, S, P>(g0);
  const int n = blockIdx.x / g.OH, oh = blockIdx.x - n * g.OH;
  const int h0 = oh * g.s - g.p;
  const float4* xn = reinterpret_cast<const float4*>(x) + (size_t)n * g.H * g.W * g.C4;
  const size_t obase = ((size_t)n * g.OH + oh) * g.OW * g.C4;
  for (int t = threadIdx.x; t < g.OW * g.C4; t += blockDim.x) {
    const int ow = POW2 ? (t >> g.c4shift) : t / g.C4, c4 = t - ow * g.C4;
    const int w0 = ow * g.s - g.p;
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    unsigned bi[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int kh = 0; kh < g.k; ++kh) {
      const int h = h0 + kh;
      if (h < 0 || h >= g.H) continue;
      const float4* xr = xn + (size_t)h * g.W * g.C4 + c4;
#pragma unroll
      for (int kw = 0; kw < g.k; ++kw) {
        const int w = w0 + kw;
        if (w < 0 || w >= g.W) continue;
        const float4 v = __ldg(xr + (size_t)w * g.C4);
        const float e[4] = {v.x, v.y, v.z, v.w};
        const unsigned code = (unsigned)(kh * g.k + kw);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (takes(e[c], best[c])) { best[c] = e[c]; bi[c] = code; }
      }
    }
    reinterpret_cast<float4*>(y)[obase + t] = make_float4(best[0], best[1], best[2], best[3]);
    reinterpret_cast<uint32_t*>(idx)[obase + t] = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
  }
}

// One CTA per input row (n, h): every input float4 gathers from the <= ceil(k/s)^2 windows that contain it.
template <int K, int S, int P, bool POW2>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ idx,
                                                          float* __restrict__ dx, const PoolGeom g0) {
  const PoolGeom g = fixed<K, S, P>(g0);
  const int n = blockIdx.x / g.H, h = blockIdx.x - n * g.H;
  // windows (oh, ow) with oh*s - p <= h <= oh*s - p + k - 1
  int oh0 = h + g.p - g.k + 1;
  oh0 = oh0 <= 0 ? 0 : (oh0 + g.s - 1) / g.s;
  int oh1 = (h + g.p) / g.s;
  if (oh1 > g.OH - 1) oh1 = g.OH - 1;
  const float4* dyn = reinterpret_cast<const float4*>(dy) + (size_t)n * g.OH * g.OW * g.C4;
  const uint32_t* ixn = reinterpret_cast<const uint32_t*>(idx) + (size_t)n * g.OH * g.OW * g.C4;
  float4* dxr = reinterpret_cast<float4*>(dx) + ((size_t)n * g.H + h) * g.W * g.C4;
  for (int t = threadIdx.x; t < g.W * g.C4; t += blockDim.x) {
    const int w = POW2 ? (t >> g.c4shift) : t / g.C4, c4 = t - w * g.C4;
    int ow0 = w + g.p - g.k + 1;
    ow0 = ow0 <= 0 ? 0 : (ow0 + g.s - 1) / g.s;
    int ow1 = (w + g.p) / g.s;
    if (ow1 > g.OW - 1) ow1 = g.OW - 1;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int oh = oh0; oh <= oh1; ++oh) {            // ascending (oh, ow): the same summation order for every run
      const unsigned kh = (unsigned)(h - (oh * g.s - g.p));
      for (int ow = ow0; ow <= ow1; ++ow) {
        const unsigned code = kh * g.k + (unsigned)(w - (ow * g.s - g.p));
        const size_t o = ((size_t)oh * g.OW + ow) * g.C4 + c4;
        const uint32_t m = __ldg(ixn + o);
        const bool h0 = (m & 0xFFu) == code, h1 = ((m >> 8) & 0xFFu) == code, h2 = ((m >> 16) & 0xFFu) == code, h3 = (m >> 24) == code;
        if (h0 || h1 || h2 || h3) {
          const float4 gq = __ldg(dyn + o);
          if (h0) acc[0] += gq.x;
          if (h1) acc[1] += gq.y;
          if (h2) acc[2] += gq.z;
          if (h3) acc[3] += gq.w;
        }
      }
    }
    dxr[t] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

}  // namespace

namespace {
int pow2_shift(int v) {
  int sh = 0;
  while ((1 << sh) < v) ++sh;
  return (1 << sh) == v ? sh : -1;
}
}  // namespace

void maxpool_fwd_launch(const float* x, float* y, uint8_t* idx, int N, int H, int W, int C, int OH, int OW, int k, int s, int p,
                        cudaStream_t st) {
  const int sh = pow2_shift(C / 4);
  const PoolGeom g{N, H, W, C / 4, OH, OW, k, s, p, sh < 0 ? 0 : sh};
  if (k == 3 && s == 2 && p == 1 && sh >= 0) maxpool_fwd_kernel<3, 2, 1, true><<<N * OH, 256, 0, st>>>(x, y, idx, g);
  else maxpool_fwd_kernel<0, 0, 0, false><<<N * OH, 256, 0, st>>>(x, y, idx, g);
}

void maxpool_bwd_launch(const float* dy, const uint8_t* idx, float* dx, int N, int H, int W, int C, int OH, int OW, int k, int s,
                        int p, cudaStream_t st) {
  const int sh = pow2_shift(C / 4);
  const PoolGeom g{N, H, W, C / 4, OH, OW, k, s, p, sh < 0 ? 0 : sh};
  if (k == 3 && s == 2 && p == 1 && sh >= 0) maxpool_bwd_kernel<3, 2, 1, true><<<N * H, 256, 0, st>>>(dy, idx, dx, g);
  else maxpool_bwd_kernel<0, 0, 0, false><<<N * H, 256, 0, st>>>(dy, idx, dx, g);
}

}  // namespace dwt
