// Tensor-core path for the two CONTRACTIONS over the flattened N*H*W axis when groups are large
// (group size 8..64, i.e. BASELINE.json config 2: C=256, group_size=64):
//
//   stats       G = sum_m (x-K)(x-K)^T        per 64-channel super-block (gs x gs diagonal blocks kept)
//   bwd_reduce  R = sum_m dy (x-mean)^T
//
// 2*gs flop per 4 bytes read puts these above the FP32-core ridge but far below the TF32
// tensor ridge: on tcgen05 they are HBM-bound.  Structure of one CTA
// (320 threads, persistent over a contiguous range of [64 channels x 32 pixels] tiles):
//
//   warp 0      TMA producer: cp.async.bulk.tensor.3d (box 32 px x 64 ch x 1 image, SWIZZLE_128B)
//               into an 8-stage shared-memory ring, mbarrier complete_tx.
//   warps 2-9   transform: in place, tile <- RN_tf32(tile - shift[channel]) (zero outside the
//               tensor), row sums in registers, fence.proxy.async, arrive on the stage's "ready" barrier.
//               Round-to-NEAREST operands make the TF32 product errors zero-mean, so they average
//               out over the M ~ 1e5..1e6 samples instead of biasing the covariance (truncation would).
//   warp 1      MMA issuer: one elected thread, tcgen05.mma.cta_group::1.kind::tf32, M=64 N=64 K=8,
//               both operands K-major straight from the swizzled tile (NCHW rows ARE K-major: the
//               reference's transposing copy, whitening.py:46, disappears), fp32 accumulator in TMEM;
//               tcgen05.commit releases the stage back to the producer.
//   epilogue    tcgen05.ld the 64x64 accumulator (M=64 layout: row r -> lane 32*(r/16) + r%16),
//               per-CTA partial -> global.  The fixed-order reduction of the partials and the dense
//               algebra (Cholesky / inverse / EMA, or the backward coefficients) run as the small
//               follow-up launches of norm_dense.cu.
//
// Reference: utils/whitening.py:46-47 (/root/reference) and its autograd transpose.
#include <cuda.h>

#include "dwt_common.cuh"
#include "norm_launch.h"
#include "tc_ptx.cuh"

namespace dwt {
namespace {

using namespace tc;

constexpr int kTW = 8;                                  // transform warps (4 left the kernel transform-latency-bound: 0.74 of the HBM peak, 0.87 without the transform)
constexpr int kTT = 32 * kTW;                           // transform threads
constexpr int kPer = 512 / kTT;                         // 16-byte chunks of a tile per transform thread
constexpr int kTcThreads = 64 + kTT;
constexpr int kTilePx = 32, kTileCh = 64;
constexpr int kTileBytes = kTileCh * kTilePx * 4;       // 8192
constexpr int kStagesStats = 12, kStagesBwd = 6;      // 97 KB per CTA, two CTAs per SM: ~190 KB of loads in flight per SM
constexpr int kTmemCols = 64;

// ------------------------------------------------------------------------------------------
// shared memory carve-up
// ------------------------------------------------------------------------------------------
struct TcBarriers {
  uint64_t full[kStagesStats];
  uint64_t ready[kStagesStats];
  uint64_t empty[kStagesStats];
  uint64_t accum;
  uint32_t tmem_slot;
  int flag;
};

// Tile range of this CTA inside its (domain, super-block) problem.
struct TileRange {
  int begin, end, PB;     // tiles [begin, end), pixel blocks per image
  __device__ TileRange(const Geom& gm) {
    PB = (gm.HW + kTilePx - 1) / kTilePx;
    const long long T = (long long)gm.N * PB;
    begin = (int)(T * blockIdx.x / gridDim.x);
    end = (int)(T * (blockIdx.x + 1) / gridDim.x);
  }
};

// In-place transform of one landed tile by the kTT transform threads:
//   v <- RN_tf32(v - shift[row]) inside the tensor, 0 outside; returns per-thread row sums of (v - shift).
// Chunk q = tt + kTT*i (16-byte units): row = q >> 3, physical chunk jp = q & 7, logical chunk = jp ^ (row & 7).
// Shared memory is addressed in its own state space (LDS/STS; a generic pointer costs an address-space check
// and global-load latency class per access), all loads first.
__device__ __forceinline__ void transform_tile(uint32_t tile, int tt, const float (&shift)[kPer], int px0, int HW, int ch0,
                                               int C, float (&rowsum)[kPer]) {
  float4 v[kPer];
#pragma unroll
  for (int i = 0; i < kPer; ++i)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w)
                 : "r"(tile + 16u * (tt + kTT * i)));
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int q = tt + kTT * i, row = q >> 3, jp = q & 7, j = jp ^ (row & 7);
    const int px = px0 + 4 * j;
    const bool rowok = (ch0 + row) < C;
    float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool ok = rowok && (px + k) < HW;
      const float s = ok ? e[k] - shift[i] : 0.f;
      rowsum[i] += s;
      e[k] = round_tf32(s);
    }
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile + 16u * (tt + kTT * i)), "f"(e[0]), "f"(e[1]),
                 "f"(e[2]), "f"(e[3]) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// the contraction kernel.  TWO = false: G = sum xs xs^T (stats).  TWO = true: R = sum dy xc^T.
// ------------------------------------------------------------------------------------------
template <bool TWO>
__global__ void __launch_bounds__(kTcThreads, 2)
tc_contract_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g,
                   const float* __restrict__ x, const Geom gm, const float* __restrict__ save_mean,
                   float* __restrict__ shift_out, float* __restrict__ partial) {
  constexpr int STAGES = TWO ? kStagesBwd : kStagesStats;
  constexpr int NT = TWO ? 2 : 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ TcBarriers bars;
  __shared__ float sShift[kTileCh];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int sb = blockIdx.y, d = blockIdx.z, ch0 = sb * kTileCh;
  const TileRange tr(gm);
  const int ntiles = tr.end - tr.begin;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.ready[s], kTW); mbar_init(&bars.empty[s], 1); }
    mbar_init(&bars.accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<kTmemCols>(&bars.tmem_slot);
  // shift per channel: pilot mean (stats) or the saved batch mean (backward)
  if (tid >= 64 && tid < 128) {
    const int r = tid - 64, c = ch0 + r;
    float sh = 0.f;
    if (c < gm.C) {
      if (TWO) sh = save_mean[(size_t)d * gm.C + c];
      else {
        const int np = gm.HW < 32 ? gm.HW : 32, p0 = ((gm.HW - np) / 2) & ~3;
        const float* px = x + ((size_t)d * gm.N * gm.C + c) * gm.HW + p0;
        float a = 0.f;
        for (int k = 0; k < np; ++k) a += __ldg(px + k);
        sh = a / (float)np;
      }
    }
    sShift[r] = sh;
    if (!TWO && blockIdx.x == 0) shift_out[((size_t)d * gridDim.y + sb) * kTileCh + r] = sh;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bars.tmem_slot;

  float rowsum[kPer];
#pragma unroll
  for (int i = 0; i < kPer; ++i) rowsum[i] = 0.f;
  const int tt = tid - 64;                       // transform thread index (warps 2..9)

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int it = 0; it < ntiles; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&bars.empty[s], ph ^ 1);
        const int t = tr.begin + it, n = t / tr.PB, pb = t - n * tr.PB;
        uint8_t* dst = smem + (size_t)s * NT * kTileBytes;
        mbar_arrive_expect_tx(&bars.full[s], NT * kTileBytes);
        tma_load_3d(dst, &map_x, pb * kTilePx, ch0, d * gm.N + n, &bars.full[s]);
        if (TWO) tma_load_3d(dst + kTileBytes, &map_g, pb * kTilePx, ch0, d * gm.N + n, &bars.full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_tf32(64, 64);
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait(&bars.ready[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t xaddr = smem_u32(smem + (size_t)s * NT * kTileBytes);
        const uint64_t xdesc = make_kmajor_sw128_desc(xaddr);
        const uint64_t adesc = TWO ? make_kmajor_sw128_desc(xaddr + kTileBytes) : xdesc;   // A = dy tile (rows i)
#pragma unroll
        for (int k = 0; k < kTilePx / 8; ++k)
          umma_tf32(tmem_d, adesc + 2 * k, xdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(&bars.empty[s]);
        if (it == ntiles - 1) umma_commit(&bars.accum);
      }
      __syncwarp();
    }
  } else {
    // ===== transform warps =====
    float shift[kPer], zero[kPer], dummy[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) { shift[i] = sShift[(tt + kTT * i) >> 3]; zero[i] = 0.f; dummy[i] = 0.f; }
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait(&bars.full[s], ph);
      const int t = tr.begin + it, n = t / tr.PB, pb = t - n * tr.PB;
      const uint32_t tile = smem_u32(smem + (size_t)s * NT * kTileBytes);
      if (TWO) {
        transform_tile(tile, tt, shift, pb * kTilePx, gm.HW, ch0, gm.C, dummy);                     // xc
        transform_tile(tile + kTileBytes, tt, zero, pb * kTilePx, gm.HW, ch0, gm.C, rowsum);        // dy, sums
      } else {
        transform_tile(tile, tt, shift, pb * kTilePx, gm.HW, ch0, gm.C, rowsum);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.ready[s]);
    }
  }

  // ===== epilogue: accumulator + row sums -> this CTA's partial row =====
  const int nacc = kTileCh * kTileCh + kTileCh;
  float* prow = partial + (((size_t)d * gridDim.y + sb) * gridDim.x + blockIdx.x) * nacc;
  if (warp >= 2) {
    if (ntiles > 0 && warp < 6) {                  // four warps cover the four TMEM lane quadrants
      mbar_wait(&bars.accum, 0);
      tc_fence_after();
      const int quad = warp & 3;                   // TMEM lane quadrant this warp may access
      float v[32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + half * 32, v);
        if (lane < 16) {                           // M=64 accumulator: row = 16*quad + lane
          float* dst = prow + (quad * 16 + lane) * kTileCh + half * 32;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4)
            *reinterpret_cast<float4*>(dst + 4 * c4) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
        }
      }
      tc_fence_before();
    } else if (ntiles == 0) {
      for (int e = tt; e < kTileCh * kTileCh; e += kTT) prow[e] = 0.f;
    }
    // row sums: the 8 lanes that share a row are consecutive
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      float s = rowsum[i];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if ((tt & 7) == 0) prow[kTileCh * kTileCh + ((tt + kTT * i) >> 3)] = s;
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<kTmemCols>(tmem_d); }

}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int make_map(CUtensorMap* map, const float* base, const Geom& gm) {
  const cuuint64_t dims[3] = {(cuuint64_t)gm.HW, (cuuint64_t)gm.C, (cuuint64_t)gm.N * gm.D};
  const cuuint64_t strides[2] = {(cuuint64_t)gm.HW * 4, (cuuint64_t)gm.C * gm.HW * 4};
  const cuuint32_t box[3] = {kTilePx, kTileCh, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return (int)g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

size_t tc_smem_bytes(bool two) { return (size_t)(two ? kStagesBwd * 2 : kStagesStats) * kTileBytes + 1024; }

}  // namespace

int tc_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || fn == nullptr || q != cudaDriverEntryPointSuccess) return e == cudaSuccess ? -1 : (int)e;
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  e = cudaFuncSetAttribute(tc_contract_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(false));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(tc_contract_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(true));
  // two ~97 KB CTAs per SM need the full shared-memory carve-out
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_contract_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_contract_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess) e = (cudaError_t)dense_init();
  if (e == cudaSuccess) return tc_apply_init();
  return (int)e;
}

// The TMA/tcgen05 contraction takes group sizes that tile a 64-channel super-block, rows that TMA can
// address (16-byte strides and base) and at least one full 32-pixel box per row.
// TF32 operands are rounded to nearest, so product errors are zero-mean and shrink as 1/sqrt(M); below a
// few thousand samples per channel they do not, and the (exact fp32, FFMA) tiled kernels take the call.
bool tc_supports(const Geom& gm, int vec) {
  return gm.GS >= 8 && kTileCh % gm.GS == 0 && vec == 4 && gm.HW >= kTilePx && (long long)gm.N * gm.HW >= 4096;
}

int tc_superblocks(const Geom& gm) { return (gm.C + kTileCh - 1) / kTileCh; }

// partial: [D][SB][nchunks][64*64+64] per-CTA moments;  shift: [D][SB][64] pilot shift of every channel
int tc_stats(const float* x, const Geom& gm, int nchunks, float* shift, float* partial, cudaStream_t st) {
  CUtensorMap mx;
  bind_context();
  if (int rc = make_map(&mx, x, gm)) return rc;
  dim3 grid(nchunks, tc_superblocks(gm), gm.D);
  tc_contract_kernel<false><<<grid, kTcThreads, tc_smem_bytes(false), st>>>(mx, mx, x, gm, nullptr, shift, partial);
  return 0;
}

int tc_bwd_reduce(const float* x, const float* dout, const Geom& gm, int nchunks, const float* save_mean,
                  float* partial, cudaStream_t st) {
  CUtensorMap mx, mg;
  bind_context();
  if (int rc = make_map(&mx, x, gm)) return rc;
  if (int rc = make_map(&mg, dout, gm)) return rc;
  dim3 grid(nchunks, tc_superblocks(gm), gm.D);
  tc_contract_kernel<true><<<grid, kTcThreads, tc_smem_bytes(true), st>>>(mx, mg, x, gm, save_mean, nullptr, partial);
  return 0;
}

}  // namespace dwt
