// Tensor-core path for the two APPLY passes when groups are large (group size 8..64):
//
//   apply       y  = W (x - mean)                               one input tensor
//   bwd_apply   dx = A1 (dy - mean_M dy) + Bm (x - mean)        two input tensors
//
// Per 64-channel super-block the per-group matrices form one block-diagonal 64x64 matrix M, and a
// [64 ch x 64 px] tile of the NCHW tensor is out = M * tile: a GEMM whose B operand is the tile exactly
// as TMA lands it (rows = channels = K, pixels contiguous = MN-major), so again nothing is transposed.
// A single-pass TF32 product is NOT accurate enough here -- each output is one length-64 dot product
// whose terms can cancel by the condition number of the covariance -- so both operands are split
// x = hi + lo (hi = RN_tf32(x), lo = RN_tf32(x - hi)) and the four partial products are accumulated in
// fp32: the matrix halves are STACKED as a 128-row A operand [M_hi ; M_lo], so one M=128 MMA per B half
// yields M_hi*B and M_lo*B rows in TMEM; the epilogue adds the two row halves.
//
// CTA = 10 warps, persistent over a contiguous range of 64-pixel tiles of one (domain, super-block):
//   warp 0     TMA producer (2 boxes of 32 px x 64 ch per input tensor and stage, SWIZZLE_128B_ATOM_32B:
//              the only layout the tensor core takes for an MN-major tf32 operand)
//   warp 1     MMA issuer (tcgen05.mma kind::tf32, M=128 N=64 K=8; A resident in TMEM, B MN-major), TMEM owner
//   warps 2-5  transform: in place hi = RN(v - shift[c]), second buffer lo = RN(v - shift[c] - hi)
//   warps 6-9  epilogue: tcgen05.ld 128 lanes x 64 columns, fold the hi/lo row halves with one shuffle,
//              256-byte row stores; TMEM accumulators are double-buffered against the MMA.
//
// Reference: the grouped 1x1 convolution at utils/whitening.py:55 (/root/reference) and its backward.
#include <cuda.h>

#include "dwt_common.cuh"
#include "norm_launch.h"
#include "tc_ptx.cuh"

namespace dwt {
namespace {

using namespace tc;

constexpr int kApThreads = 320;
constexpr int kBoxPx = 32, kTilePxA = 64, kCh = 64;
constexpr int kBoxBytes = kCh * kBoxPx * 4;          // 8192
constexpr int kHalfBytes = 2 * kBoxBytes;            // hi (or lo) of one input: 2 boxes = 16 KB
constexpr int kInBytes = 2 * kHalfBytes;             // hi + lo of one input = 32 KB
constexpr int kAccCols = 64;                         // TMEM columns per accumulator / per matrix
constexpr int kTmemAlloc = 256;                      // 2 accumulators + up to 2 resident matrices
constexpr int kMaxStages = 4;

struct ApBarriers {
  uint64_t full[kMaxStages], ready[kMaxStages], empty[kMaxStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_slot;
};

struct ApplyArgs {
  const float* mats;      // per (domain, group) records
  int rec_stride;         // floats per record
  int off[2];             // offset of the matrix applied to input i inside a record
  const float* shift[2];  // per-channel shift of input i
  int shift_stride[2];    // floats per domain in shift[i]
  float* out;
};

// hi/lo split of one landed 32 px x 64 ch box by the 128 transform threads (tt = 0..127).
__device__ __forceinline__ void split_box(float* hi, float* lo, int tt, const float (&shift)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = tt + 128 * i;                    // 16-byte chunk; row = q >> 3
    float4 v = reinterpret_cast<float4*>(hi)[q];
    float e[4] = {v.x - shift[i], v.y - shift[i], v.z - shift[i], v.w - shift[i]}, h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { h[k] = round_tf32(e[k]); l[k] = round_tf32(e[k] - h[k]); }
    reinterpret_cast<float4*>(hi)[q] = make_float4(h[0], h[1], h[2], h[3]);
    reinterpret_cast<float4*>(lo)[q] = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// The A operand lives in TENSOR MEMORY for the whole kernel (it is the same for every tile; reading
// it from shared memory on every MMA was the largest shared-memory consumer of the first version).
// Row order: TMEM lane 32q + l holds channel 16q + (l & 15), the hi part for l < 16 and the lo part
// for l >= 16, so the two halves that must be added end up 16 lanes apart IN THE SAME WARP and the
// epilogue folds them with one shuffle per value -- no shared-memory staging at all.
template <int NIN>
__global__ void __launch_bounds__(kApThreads, NIN == 1 ? 2 : 1)
tc_apply_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const Geom gm,
                const ApplyArgs args) {
  constexpr int STAGES = 3;       // NIN = 1: 97 KB -> two CTAs per SM; NIN = 2: 193 KB -> one
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sStage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ ApBarriers bars;
  __shared__ float sShift[2][kCh];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int sb = blockIdx.y, d = blockIdx.z, ch0 = sb * kCh;
  const int PB = (gm.HW + kTilePxA - 1) / kTilePxA;
  const long long T = (long long)gm.N * PB;
  const int t_begin = (int)(T * blockIdx.x / gridDim.x), t_end = (int)(T * (blockIdx.x + 1) / gridDim.x);
  const int ntiles = t_end - t_begin;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.ready[s], 4); mbar_init(&bars.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars.acc_full[b], 1); mbar_init(&bars.acc_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<kTmemAlloc>(&bars.tmem_slot);
  if (tid < NIN * kCh) {
    const int i = tid / kCh, r = tid - i * kCh, c = ch0 + r;
    sShift[i][r] = (c < gm.C && args.shift[i] != nullptr) ? __ldg(args.shift[i] + (size_t)d * args.shift_stride[i] + c) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_slot;
  const uint32_t tmem_acc = tmem, tmem_mat = tmem + 2 * kAccCols;

  if (warp >= 6) {
    // epilogue warps first park the split matrices in tensor memory
    const int quad = warp & 3, i = 16 * quad + (lane & 15), GS = gm.GS, gi = i / GS, g = sb * (kCh / GS) + gi;
    const bool lo_half = lane >= 16;
    for (int m = 0; m < NIN; ++m) {
      const float* rec = args.mats + ((size_t)d * gm.G + (g < gm.G ? g : 0)) * args.rec_stride + args.off[m] + (i - gi * GS) * GS;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int k = h * 32 + c;
          float w = 0.f;
          if (g < gm.G && k / GS == gi) w = __ldg(rec + (k - gi * GS));
          const float hi = round_tf32(w);
          v[c] = lo_half ? round_tf32(w - hi) : hi;
        }
        tmem_st32(tmem_mat + m * kAccCols + ((uint32_t)(quad * 32) << 16) + h * 32, v);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int it = 0; it < ntiles; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&bars.empty[s], ph ^ 1);
        const int t = t_begin + it, n = t / PB, pb = t - n * PB;
        uint8_t* dst = sStage + (size_t)s * NIN * kInBytes;
        // the second 32-pixel box of the tile may lie entirely past the row end: do not issue it (its
        // columns are never stored, whatever the stale shared memory holds)
        const int nbox = (pb * kTilePxA + kBoxPx < gm.HW) ? 2 : 1;
        mbar_arrive_expect_tx(&bars.full[s], NIN * nbox * kBoxBytes);
#pragma unroll
        for (int i = 0; i < NIN; ++i)
          for (int j = 0; j < nbox; ++j)
            tma_load_3d(dst + i * kInBytes + j * kBoxBytes, i == 0 ? &map0 : &map1, pb * kTilePxA + j * kBoxPx, ch0,
                        d * gm.N + n, &bars.full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_tf32(128, kTilePxA, true);
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1, b = it & 1, aph = (it >> 1) & 1;
      mbar_wait(&bars.ready[s], ph);
      mbar_wait(&bars.acc_empty[b], aph ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t stage = smem_u32(sStage + (size_t)s * NIN * kInBytes);
        uint32_t first = 1;
#pragma unroll
        for (int i = 0; i < NIN; ++i) {
#pragma unroll
          for (int term = 0; term < 2; ++term) {                 // B = hi half, then lo half
            const uint32_t bbase = stage + i * kInBytes + term * kHalfBytes;
#pragma unroll
            for (int ks = 0; ks < kCh / 8; ++ks) {
              const uint64_t bdesc = make_mnmajor_sw128_32b_desc(bbase + ks * 1024, kBoxBytes, 512);
              umma_tf32_ts(tmem_acc + b * kAccCols, tmem_mat + i * kAccCols + ks * 8, bdesc, idesc, first ? 0u : 1u);
              first = 0;
            }
          }
        }
        umma_commit(&bars.empty[s]);
        umma_commit(&bars.acc_full[b]);
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ===== transform warps =====
    const int tt = tid - 64;
    float sh[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int r = 0; r < 4; ++r) sh[i][r] = sShift[i][(tt + 128 * r) >> 3];
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait(&bars.full[s], ph);
      uint8_t* st = sStage + (size_t)s * NIN * kInBytes;
#pragma unroll
      for (int i = 0; i < NIN; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          split_box(reinterpret_cast<float*>(st + i * kInBytes + j * kBoxBytes),
                    reinterpret_cast<float*>(st + i * kInBytes + kHalfBytes + j * kBoxBytes), tt, sh[i]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.ready[s]);
    }
  } else {
    // ===== epilogue warps =====
    const int quad = warp & 3;                     // TMEM lane quadrant of this warp
    const int ch = ch0 + 16 * quad + (lane & 15);  // channel whose hi (lane<16) / lo (lane>=16) row this thread reads
    for (int it = 0; it < ntiles; ++it) {
      const int b = it & 1, aph = (it >> 1) & 1;
      mbar_wait(&bars.acc_full[b], aph);
      tc_fence_after();
      float v[2][32];
      tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + b * kAccCols, v[0]);
      tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + b * kAccCols + 32, v[1]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.acc_empty[b]);          // accumulator drained: MMA may reuse it
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < 32; ++c) v[h][c] += __shfl_down_sync(0xffffffffu, v[h][c], 16);
      const int t = t_begin + it, n = t / PB, pb = t - n * PB, px0 = pb * kTilePxA;
      if (lane < 16 && ch < gm.C) {
        float* orow = args.out + ((size_t)(d * gm.N + n) * gm.C + ch) * gm.HW + px0;
        if ((gm.HW & 7) == 0) {
          // 256-bit stores (sm_100 STG.256): every instruction writes whole 32-byte sectors of this thread's row;
          // 128-bit stores left half-written sectors behind (16 sectors per request, 2x the L2 write traffic)
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8)
              if (px0 + h * 32 + 8 * c8 < gm.HW)
                asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + h * 32 + 8 * c8),
                             "f"(v[h][8 * c8]), "f"(v[h][8 * c8 + 1]), "f"(v[h][8 * c8 + 2]), "f"(v[h][8 * c8 + 3]),
                             "f"(v[h][8 * c8 + 4]), "f"(v[h][8 * c8 + 5]), "f"(v[h][8 * c8 + 6]), "f"(v[h][8 * c8 + 7]) : "memory");
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4)
              if (px0 + h * 32 + 4 * c4 < gm.HW)
                *reinterpret_cast<float4*>(orow + h * 32 + 4 * c4) =
                    make_float4(v[h][4 * c4], v[h][4 * c4 + 1], v[h][4 * c4 + 2], v[h][4 * c4 + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<kTmemAlloc>(tmem); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_ap = nullptr;

int make_map_ap(CUtensorMap* map, const float* base, const Geom& gm) {
  const cuuint64_t dims[3] = {(cuuint64_t)gm.HW, (cuuint64_t)gm.C, (cuuint64_t)gm.N * gm.D};
  const cuuint64_t strides[2] = {(cuuint64_t)gm.HW * 4, (cuuint64_t)gm.C * gm.HW * 4};
  const cuuint32_t box[3] = {kBoxPx, kCh, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return (int)g_encode_ap(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

template <int NIN> constexpr size_t ap_smem() { return (size_t)3 * NIN * kInBytes + 1024; }

}  // namespace

int tc_apply_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || fn == nullptr || q != cudaDriverEntryPointSuccess) return e == cudaSuccess ? -1 : (int)e;
  g_encode_ap = reinterpret_cast<EncodeTiledFn>(fn);
  e = cudaFuncSetAttribute(tc_apply_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ap_smem<1>());
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_apply_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ap_smem<2>());
  // two 97 KB CTAs per SM need the full shared-memory carve-out
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_apply_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_apply_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  return (int)e;
}

// y = W (x - mean): W from save_w [D][G][gs*gs], mean from save_mean [D][C]
int tc_apply(const float* x, float* y, const Geom& gm, int nctas, const float* save_mean, const float* save_w,
             cudaStream_t st) {
  CUtensorMap mx;
  cudaFree(nullptr);
  if (int rc = make_map_ap(&mx, x, gm)) return rc;
  ApplyArgs a{};
  a.mats = save_w; a.rec_stride = gm.GS * gm.GS; a.off[0] = 0; a.off[1] = 0;
  a.shift[0] = save_mean; a.shift_stride[0] = gm.C; a.shift[1] = nullptr; a.shift_stride[1] = 0;
  a.out = y;
  dim3 grid(nctas, (gm.C + kCh - 1) / kCh, gm.D);
  tc_apply_kernel<1><<<grid, kApThreads, ap_smem<1>(), st>>>(mx, mx, gm, a);
  return 0;
}

// dx = A1 (dy - dybar) + Bm (x - mean): coef [D][G][2 gs^2 + gs] = A1 | Bm | cvec, dybar [D][SB*64]
int tc_bwd_apply(const float* x, const float* dout, float* dx, const Geom& gm, int nctas, const float* coef,
                 const float* save_mean, const float* dybar, cudaStream_t st) {
  CUtensorMap mx, mg;
  cudaFree(nullptr);
  if (int rc = make_map_ap(&mg, dout, gm)) return rc;
  if (int rc = make_map_ap(&mx, x, gm)) return rc;
  ApplyArgs a{};
  a.mats = coef; a.rec_stride = coef_stride(gm.GS); a.off[0] = 0; a.off[1] = gm.GS * gm.GS;
  a.shift[0] = dybar; a.shift_stride[0] = ((gm.C + kCh - 1) / kCh) * kCh;
  a.shift[1] = save_mean; a.shift_stride[1] = gm.C;
  a.out = dx;
  dim3 grid(nctas, (gm.C + kCh - 1) / kCh, gm.D);
  tc_apply_kernel<2><<<grid, kApThreads, ap_smem<2>(), st>>>(mg, mx, gm, a);
  return 0;
}

}  // namespace dwt
