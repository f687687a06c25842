// Paired target augmentation on the GPU: from ONE resized uint8 image per sample, both views
// the reference's data loader produces on CPU workers -- one launch, written straight into the model's input.
//
//   plain  crop -> /255 -> (v - mean) / std                                  resnet50_dwt_mec_officehome.py:526-531
//   aug    crop -> flip -> /255 -> cv2.warpAffine -> blur(k=1: identity)
//          -> (v - mean) / std                                                resnet50_dwt_mec_officehome.py:481-492,534-542
//   both from the same image                                                  utils/folder.py:127-147
//
// The affine warp restates cv2.warpAffine (INTER_LINEAR, BORDER_CONSTANT 0, CV_32FC3) bit for bit: the float32
// 2x3 matrix is widened to double and inverted in double, destination pixel (x, y) samples the source at the
// FIXED-POINT position  X = (rint((m1*y + m2)*1024) + 16 + rint(m0*x*1024)) >> 5  (likewise Y), integer part X >> 5,
// fraction (X & 31)/32, four taps blended in float32 with product weights, summed left to right, taps outside the
// image contributing 0.  Every floating-point step uses the explicitly rounded intrinsics so that nvcc cannot
// contract a multiply-add: parity with the CPU pipeline is exact, not approximate.
//
// One thread per output pixel (3 channels); the uint8 source (<= 200 KB per image) is served by L1/L2, the float
// outputs are written coalesced (NCHW: three planes; NHWC: 12 contiguous bytes per thread).
#include <stdint.h>

#include "dwt_common.cuh"
#include "norm_launch.h"

namespace dwt {
namespace {

struct AugArgs {
  const uint8_t* images;      // [B, SH, SW, 3]
  const int* crop_plain;      // [B, 2] top, left
  const int* crop_aug;        // [B, 2]
  const uint8_t* flip;        // [B]
  const float* affine;        // [B, 6] row-major 2x3
  float mean[3], stdv[3];
  float* out_plain;           // [B, 3, CR, CR] or NHWC; may be null
  float* out_aug;             // may be null
  int B, SH, SW, CR, nhwc;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ void store_pixel(float* out, const AugArgs& a, int b, int y, int x, const float (&v)[3]) {
  if (a.nhwc) {
    float* p = out + (((size_t)b * a.CR + y) * a.CR + x) * 3;
    p[0] = v[0]; p[1] = v[1]; p[2] = v[2];
  } else {
    const size_t plane = (size_t)a.CR * a.CR;
    float* p = out + (size_t)b * 3 * plane + (size_t)y * a.CR + x;
    p[0] = v[0]; p[plane] = v[1]; p[2 * plane] = v[2];
  }
}

__global__ void __launch_bounds__(256) augment_pair_kernel(const AugArgs a) {
  // Per-CTA tables instead of per-pixel IEEE divisions (an exact division is ~10 instructions and the augmented
  // view needs 15 per pixel): a uint8 has 256 values, so ToTensor (and, for the plain view, Normalize as well) is
  // a lookup of the very same correctly rounded results.  The inverse matrix is one thread's work per CTA.
  __shared__ float sLut[3][256];
  __shared__ double sM[6];
  const int pix = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y, branch = blockIdx.z;
  if ((branch == 0 ? a.out_plain : a.out_aug) == nullptr) return;               // uniform per CTA
  if (branch == 0) {
    for (int c = 0; c < 3; ++c)
      sLut[c][threadIdx.x] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)threadIdx.x, 255.f), a.mean[c]), a.stdv[c]);
  } else {
    sLut[0][threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.f);
    if (threadIdx.x == 0) {
      // cv2::invertAffineTransform in double, no contraction
      const float* Mf = a.affine + 6 * b;
      const double M0 = Mf[0], M1 = Mf[1], M2 = Mf[2], M3 = Mf[3], M4 = Mf[4], M5 = Mf[5];
      double D = __dsub_rn(__dmul_rn(M0, M4), __dmul_rn(M1, M3));
      D = D != 0.0 ? __ddiv_rn(1.0, D) : 0.0;
      const double m0 = __dmul_rn(M4, D), m4 = __dmul_rn(M0, D), m1 = __dmul_rn(M1, -D), m3 = __dmul_rn(M3, -D);
      sM[0] = m0; sM[1] = m1; sM[3] = m3; sM[4] = m4;
      sM[2] = __dsub_rn(__dmul_rn(-m0, M2), __dmul_rn(m1, M5));
      sM[5] = __dsub_rn(__dmul_rn(-m3, M2), __dmul_rn(m4, M5));
    }
  }
  __syncthreads();
  if (pix >= a.CR * a.CR) return;
  const int y = pix / a.CR, x = pix - y * a.CR;
  const uint8_t* img = a.images + (size_t)b * a.SH * a.SW * 3;
  float v[3];
  if (branch == 0) {
    const int top = clampi(a.crop_plain[2 * b], 0, a.SH - a.CR), left = clampi(a.crop_plain[2 * b + 1], 0, a.SW - a.CR);
    const uint8_t* s = img + ((size_t)(top + y) * a.SW + left + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = sLut[c][s[c]];
    store_pixel(a.out_plain, a, b, y, x, v);
    return;
  }
  const int top = clampi(a.crop_aug[2 * b], 0, a.SH - a.CR), left = clampi(a.crop_aug[2 * b + 1], 0, a.SW - a.CR);
  const bool flip = a.flip[b] != 0;
  const double m0 = sM[0], m1 = sM[1], m2 = sM[2], m3 = sM[3], m4 = sM[4], m5 = sM[5];
  // fixed-point source position (AB_BITS = 10, INTER_BITS = 5); OpenCV keeps these in 32-bit ints as well
  const int X = (__double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m1, (double)y), m2), 1024.0)) + 16 +
                 __double2int_rn(__dmul_rn(__dmul_rn(m0, (double)x), 1024.0))) >> 5;
  const int Y = (__double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m4, (double)y), m5), 1024.0)) + 16 +
                 __double2int_rn(__dmul_rn(__dmul_rn(m3, (double)x), 1024.0))) >> 5;
  const int sx = X >> 5, sy = Y >> 5;
  const float fx = (float)(X & 31) * 0.03125f, fy = (float)(Y & 31) * 0.03125f;     // exact
  const float w00 = __fmul_rn(1.f - fy, 1.f - fx), w01 = __fmul_rn(1.f - fy, fx);
  const float w10 = __fmul_rn(fy, 1.f - fx), w11 = __fmul_rn(fy, fx);
  // the four taps of the cropped (and flipped) image; 32-bit byte offsets inside this image (< 2^31 by the size check)
  const bool okx0 = (unsigned)sx < (unsigned)a.CR, okx1 = (unsigned)(sx + 1) < (unsigned)a.CR;
  const bool oky0 = (unsigned)sy < (unsigned)a.CR, oky1 = (unsigned)(sy + 1) < (unsigned)a.CR;
  const int col0 = left + (flip ? a.CR - 1 - sx : sx), dcol = flip ? -3 : 3, drow = a.SW * 3;
  const int o00 = ((top + sy) * a.SW + col0) * 3;
  const float* lut = sLut[0];
  float t[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool ok = ((k & 1) ? okx1 : okx0) && ((k >> 1) ? oky1 : oky0);
    // an out-of-image tap reads the crop origin instead (always a valid address) and is then discarded
    const int o = ok ? o00 + ((k & 1) ? dcol : 0) + ((k >> 1) ? drow : 0) : (top * a.SW + left) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) t[k][c] = ok ? lut[img[o + c]] : 0.f;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[0][c], w00), __fmul_rn(t[1][c], w01)), __fmul_rn(t[2][c], w10)),
                              __fmul_rn(t[3][c], w11));
    v[c] = __fdiv_rn(__fsub_rn(w, a.mean[c]), a.stdv[c]);
  }
  store_pixel(a.out_aug, a, b, y, x, v);
}

}  // namespace

void augment_pair_launch(const uint8_t* images, int B, int SH, int SW, int CR, const int* crop_plain, const int* crop_aug,
                         const uint8_t* flip, const float* affine, const float* mean, const float* stdv,
                         float* out_plain, float* out_aug, int nhwc, cudaStream_t st) {
  AugArgs a{};
  a.images = images; a.crop_plain = crop_plain; a.crop_aug = crop_aug; a.flip = flip; a.affine = affine;
  for (int c = 0; c < 3; ++c) { a.mean[c] = mean[c]; a.stdv[c] = stdv[c]; }
  a.out_plain = out_plain; a.out_aug = out_aug;
  a.B = B; a.SH = SH; a.SW = SW; a.CR = CR; a.nhwc = nhwc;
  dim3 grid((CR * CR + 255) / 256, B, 2);
  augment_pair_kernel<<<grid, 256, 0, st>>>(a);
}

}  // namespace dwt
