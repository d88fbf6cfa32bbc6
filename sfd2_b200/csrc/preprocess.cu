// Input leg of the extraction sweep (extract_localization.py:158-190, ImageDataset.__getitem__): the decoded uint8 image
// goes to the device as it is (3 bytes per pixel instead of 12) and ONE kernel does what the reference does on four
// DataLoader worker processes: BGR -> RGB, float32, cv2.resize(..., INTER_CUBIC) to the preset's resize_max, HWC -> CHW,
// / 255.  HBM-bound gather (16 taps x 3 channels of uint8 per output pixel, 12 bytes written).
//
// cv2's float path (imgproc/resize.cpp, resizeGeneric_<HResizeCubic<float,..>, VResizeCubic<float,..>>) restated:
//   fx = (float)((dx + 0.5) * scale_x - 0.5), sx = floor(fx), fx -= sx      (scale in double, as cv2 computes it)
//   coefficients: interpolateCubic with A = -0.75, all in float, c3 = 1 - c0 - c1 - c2
//   horizontal pass on each of the 4 source rows (taps sx-1 .. sx+2, indices clamped to the image), then the vertical
//   combination b0*H0 + b1*H1 + b2*H2 + b3*H3; no rounding or saturation (the result may leave [0, 255]).
// Every product and sum is individually rounded (no FMA contraction), left to right, like the scalar C++ code.
#include "common.cuh"

namespace sfd2 {

__device__ __forceinline__ void cubic_coeffs(float x, float (&c)[4]) {
  const float A = -0.75f;
  const float x1 = __fadd_rn(x, 1.f);
  c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), 5.f * A), x1), 8.f * A), x1), 4.f * A);
  c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, x), A + 3.f), x), x), 1.f);
  const float y = __fsub_rn(1.f, x);
  c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, y), A + 3.f), y), y), 1.f);
  c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
}

__global__ void __launch_bounds__(256)
preprocess_kernel(const uint8_t* __restrict__ src, int h, int w, int swap_rb, int hn, int wn, double scale_x, double scale_y,
                  float* __restrict__ out) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
  if (dx >= wn || dy >= hn) return;
  const size_t plane = (size_t)hn * wn;
  float* o = out + (size_t)dy * wn + dx;
  if (hn == h && wn == w) {                       // no resize: convert only
    const uint8_t* p = src + ((size_t)dy * w + dx) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * plane] = __fdiv_rn((float)p[swap_rb ? 2 - c : c], 255.f);
    return;
  }
  float fx = (float)(((double)dx + 0.5) * scale_x - 0.5);
  const int sx = (int)floorf(fx);
  fx = __fsub_rn(fx, (float)sx);
  float fy = (float)(((double)dy + 0.5) * scale_y - 0.5);
  const int sy = (int)floorf(fy);
  fy = __fsub_rn(fy, (float)sy);
  float ca[4], cb[4];
  cubic_coeffs(fx, ca);
  cubic_coeffs(fy, cb);
  int xs[4], ys[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    xs[k] = min(max(sx - 1 + k, 0), w - 1);
    ys[k] = min(max(sy - 1 + k, 0), h - 1);
  }
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const uint8_t* row = src + (size_t)ys[r] * w * 3;
    float hsum[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int ch = swap_rb ? 2 - c : c;
      float s = __fmul_rn((float)row[xs[0] * 3 + ch], ca[0]);
      s = __fadd_rn(s, __fmul_rn((float)row[xs[1] * 3 + ch], ca[1]));
      s = __fadd_rn(s, __fmul_rn((float)row[xs[2] * 3 + ch], ca[2]));
      s = __fadd_rn(s, __fmul_rn((float)row[xs[3] * 3 + ch], ca[3]));
      hsum[c] = s;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] = (r == 0) ? __fmul_rn(hsum[c], cb[0]) : __fadd_rn(acc[c], __fmul_rn(hsum[c], cb[r]));
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c * plane] = __fdiv_rn(acc[c], 255.f);
}

int launch_preprocess(const uint8_t* src, int h, int w, int swap_rb, int hn, int wn, float* out, cudaStream_t st) {
  // cv2: inv_scale = dsize / ssize (double), scale = 1 / inv_scale
  const double scale_x = 1.0 / ((double)wn / (double)w), scale_y = 1.0 / ((double)hn / (double)h);
  dim3 grid(cdiv(wn, 256), hn);
  preprocess_kernel<<<grid, 256, 0, st>>>(src, h, w, swap_rb, hn, wn, scale_x, scale_y, out);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

}  // namespace sfd2
