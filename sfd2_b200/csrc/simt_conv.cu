// CUDA-core (fp32 FMA) kernels of the convolution stack.
//
//  * conv1a_kernel   : RGB normalise (nets/extractor.py:14-17,104) fused with conv1a (3->64,
//                      nets/sfd2.py:268) - K = 27 is too thin for the tensor cores, so both
//                      precision families use this kernel; it writes fp32 (FP32 mode) or
//                      fp16 hi/lo planes (tcgen05 modes).
//  * conv_f32_kernel : generic NHWC implicit-GEMM 3x3 / 1x1 convolution, fp32 in / fp32 out,
//                      bias + ReLU + residual epilogue (SFD2_PREC_FP32 reference mode).
//  * gconv_f32_kernel: the ResBlock's grouped 3x3 (groups = 32, 8 ch/group; nets/sfd2.py:32).
//  * sta_kernel      : ConvSta 1x1 256->3 (nets/sfd2.py:303,345), always fp32.
//  * softmax65 / l2norm128 : head epilogues (nets/sfd2.py:330-333, :342).
#include <algorithm>

#include "common.cuh"

namespace sfd2 {

// ------------------------------------------------------------------------------ normalise + conv1a
// norm_kernel: (x - mean) / std with the reference's exact IEEE divisions (tvf.Normalize; u8 inputs are
// divided by 255 first), once per value, into an NHWC4 fp32 image (r, g, b, 0).  Doing it here instead of
// inside conv1a removes ~50 division instructions per value per tap-neighbourhood from the conv kernel.
template <int IMG_DTYPE>
__global__ void __launch_bounds__(256)
norm_kernel(const void* __restrict__ img, int H, int W, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  float v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float raw;
    if (IMG_DTYPE == SFD2_IMG_F32_NCHW) raw = __ldg(reinterpret_cast<const float*>(img) + (size_t)c * H * W + i);
    else raw = __fdiv_rn((float)__ldg(reinterpret_cast<const unsigned char*>(img) + (size_t)i * 3 + c), 255.0f);
    const float mean = (c == 0) ? 0.485f : (c == 1 ? 0.456f : 0.406f);
    const float stdv = (c == 0) ? 0.229f : (c == 1 ? 0.224f : 0.225f);
    v[c] = __fdiv_rn(__fsub_rn(raw, mean), stdv);
  }
  out[i] = make_float4(v[0], v[1], v[2], 0.f);
}

// conv1a: thread = 4 horizontally adjacent pixels x 16 output channels.  The 16-channel chunk is chosen per
// WARP (warp & 3), so every weight read from shared memory is a warp-uniform broadcast; lanes are consecutive
// pixel quads and write whole 32-byte sectors.  1728 FMAs per thread against 108 LDS.128 and 18 LDG.128.
template <int TC_OUT>
__global__ void __launch_bounds__(128)
conv1a_kernel(const float4* __restrict__ nimg, int H, int W, int Wp, const float* __restrict__ wt /*[27][64]*/,
              const float* __restrict__ bias, float* __restrict__ out_f32, __half* __restrict__ out_hi,
              __half* __restrict__ out_lo) {
  __shared__ __align__(16) float ws[27 * 64 + 64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) ws[i] = wt[i];
  if (threadIdx.x < 64) ws[27 * 64 + threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, cq = threadIdx.x >> 5;
  const int x = (blockIdx.x * 32 + lane) * 4;
  const int y = blockIdx.y;
  if (x >= W) return;
  float acc[4][16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[0][j] = acc[1][j] = acc[2][j] = acc[3][j] = ws[27 * 64 + cq * 16 + j];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = y + ky - 1;
    float in[6][3];   // columns x-1 .. x+4 of this row (zero padding outside the image)
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const int ix = x + q - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(nimg + (size_t)iy * W + ix);
      in[q][0] = v.x; in[q][1] = v.y; in[q][2] = v.z;
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4* wr = reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * 3 + c) * 64 + cq * 16);
        float wv[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) { const float4 t = wr[g]; wv[4 * g] = t.x; wv[4 * g + 1] = t.y; wv[4 * g + 2] = t.z; wv[4 * g + 3] = t.w; }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float a = in[kx + p][c];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[p][j] = fmaf(a, wv[j], acc[p][j]);
        }
      }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (x + p >= W) break;
    const size_t obase = ((size_t)y * Wp + x + p) * 64 + cq * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[p][j] = fmaxf(acc[p][j], 0.f);
    if (TC_OUT) {
      __align__(16) __half2 hi[8];
      __align__(16) __half2 lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        hi[j] = __floats2half2_rn(acc[p][2 * j], acc[p][2 * j + 1]);
        const float2 hf = __half22float2(hi[j]);
        lo[j] = __floats2half2_rn(acc[p][2 * j] - hf.x, acc[p][2 * j + 1] - hf.y);
      }
      uint4* ph = reinterpret_cast<uint4*>(out_hi + obase);
      uint4* pl = reinterpret_cast<uint4*>(out_lo + obase);
      ph[0] = reinterpret_cast<uint4*>(hi)[0];
      ph[1] = reinterpret_cast<uint4*>(hi)[1];
      pl[0] = reinterpret_cast<uint4*>(lo)[0];
      pl[1] = reinterpret_cast<uint4*>(lo)[1];
    } else {
      float4* o = reinterpret_cast<float4*>(out_f32 + obase);
#pragma unroll
      for (int g = 0; g < 4; ++g) o[g] = make_float4(acc[p][4 * g], acc[p][4 * g + 1], acc[p][4 * g + 2], acc[p][4 * g + 3]);
    }
  }
}

// conv1a for the tcgen05 modes: same arithmetic, but the memory side goes through shared memory both ways.
//  * the block's 3 x 258 normalised input pixels are staged once per 256-pixel row segment (coalesced float4);
//  * thread = 8 pixels (lane, lane+32, ... lane+224 of the segment) x 8 output channels (chosen per WARP, so weight
//    reads are warp-uniform broadcasts).  Every weight fetched from smem feeds 8 FMAs: at 4 px/thread the kernel
//    needed 128 B/clk/SM of shared-memory bandwidth - exactly the hardware limit (ncu: L1/TEX 78 %);
//  * fp16 hi/lo results go to 128-byte-swizzled [256 px][64 ch] staging tiles (4 bank wavefronts per 16-byte
//    store, the minimum) and leave through two TMA stores per segment (box {64 ch, 256 px, 1 row}, clipped at
//    the image edge).  Direct 16-byte global stores at a 512-byte lane stride made the first versions L1-bound.
//  * persistent: weights are staged once per block, blocks walk over row segments.
constexpr int C1_SEG = 256;
__global__ void __launch_bounds__(256, 2)
conv1a_tc_kernel(const float4* __restrict__ nimg, int H, int W, const float* __restrict__ wt /*[27][64]*/,
                 const float* __restrict__ bias, const __grid_constant__ CUtensorMap tm_hi,
                 const __grid_constant__ CUtensorMap tm_lo) {
  extern __shared__ uint8_t smem_raw1a[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw1a) + 1023) & ~(uintptr_t)1023);
  uint8_t* t_hi = base;                                   // [256 px][128 B], SWIZZLE_128B
  uint8_t* t_lo = base + C1_SEG * 128;
  float* ws = reinterpret_cast<float*>(base + 2 * C1_SEG * 128);            // [27][64] + bias[64]
  float* patch = reinterpret_cast<float*>(base + 2 * C1_SEG * 128 + 7168);    // [3 rows][3 ch][260]
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) ws[i] = wt[i];
  if (threadIdx.x < 64) ws[27 * 64 + threadIdx.x] = bias[threadIdx.x];
  const int lane = threadIdx.x & 31, co = (threadIdx.x >> 5) * 8;   // this warp's 8 output channels
  const int segs_x = (W + C1_SEG - 1) / C1_SEG, nseg = segs_x * H;
  for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    const int y = seg / segs_x, x0 = (seg - y * segs_x) * C1_SEG;
    for (int i = threadIdx.x; i < 3 * (C1_SEG + 2); i += blockDim.x) {
      const int ky = i / (C1_SEG + 2), col = i - ky * (C1_SEG + 2);
      const int iy = y + ky - 1, ix = x0 + col - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(nimg + (size_t)iy * W + ix);
      patch[(ky * 3 + 0) * 260 + col] = v.x;
      patch[(ky * 3 + 1) * 260 + col] = v.y;
      patch[(ky * 3 + 2) * 260 + col] = v.z;
    }
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = ws[27 * 64 + co + j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float in[8];
#pragma unroll
          for (int p = 0; p < 8; ++p) in[p] = patch[(ky * 3 + c) * 260 + lane + 32 * p + kx];
          const float4* wr = reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * 3 + c) * 64 + co);
          const float4 w0 = wr[0], w1 = wr[1];
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int p = 0; p < 8; ++p)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(in[p], wv[j], acc[p][j]);
        }
      }
    // the previous segment's TMA stores must have finished reading the staging tiles
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int row = lane + 32 * p;                      // pixel within the segment = row of the staging tile
      __align__(16) __half2 hi[4];
      __align__(16) __half2 lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v0 = fmaxf(acc[p][2 * j], 0.f), v1 = fmaxf(acc[p][2 * j + 1], 0.f);
        hi[j] = __floats2half2_rn(v0, v1);
        const float2 hf = __half22float2(hi[j]);
        lo[j] = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
      }
      const int chunk = (co >> 3) ^ (row & 7);            // SWIZZLE_128B
      *reinterpret_cast<uint4*>(t_hi + row * 128 + chunk * 16) = *reinterpret_cast<const uint4*>(hi);
      *reinterpret_cast<uint4*>(t_lo + row * 128 + chunk * 16) = *reinterpret_cast<const uint4*>(lo);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();                                      // tiles complete; also: everyone is done with `patch`
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(reinterpret_cast<uint64_t>(&tm_hi)), "r"((uint32_t)__cvta_generic_to_shared(t_hi)), "r"(0), "r"(x0), "r"(y) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(reinterpret_cast<uint64_t>(&tm_lo)), "r"((uint32_t)__cvta_generic_to_shared(t_lo)), "r"(0), "r"(x0), "r"(y) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int g_conv1a_mma = 1;   // SFD2_CONV1A_MMA=0: keep conv1a on the CUDA cores in the tcgen05 modes

// tm1a: store maps of the conv1a output (tcgen05 modes only): [hi, lo] with box {64 ch, 256 px, 1 row} for the
// CUDA-core kernel, then [hi, lo] with box {64 ch, 128 px, 1 row} for the tensor-core kernel.
// tc_out: 0 = fp32 output (FP32 mode), else the split of the tcgen05 mode (1 or 3).
bool conv1a_bands_ok(int tc_out) { return tc_out && g_conv1a_mma; }

int launch_conv1a(const void* img, int img_dtype, int H, int W, const Layer& L, Act out, int tc_out, float4* nimg,
                  const CUtensorMap* tm1a, int num_sms, cudaStream_t st, int y_begin, int y_end) {
  SFD2_CHECK(L.cin == 3 && L.cout == 64 && L.k == 3, SFD2_ERR_WEIGHTS, "conv1a: unexpected layer shape");
  SFD2_CHECK(img_dtype == SFD2_IMG_F32_NCHW || img_dtype == SFD2_IMG_U8_NHWC, SFD2_ERR_ARG, "unknown image dtype %d", img_dtype);
  // tcgen05 modes: one kernel normalises, builds the im2col operand and runs the MMAs (tc_conv1a.cu)
  if (tc_out && g_conv1a_mma) return launch_conv1a_mma(img, img_dtype, H, W, L, tm1a + 2, tc_out, num_sms, st, y_begin, y_end);
  SFD2_CHECK(y_begin == 0 && (y_end < 0 || y_end == H), SFD2_ERR_ARG, "conv1a: row bands need the tcgen05 kernel");
  if (img_dtype == SFD2_IMG_F32_NCHW) norm_kernel<SFD2_IMG_F32_NCHW><<<cdiv(H * W, 256), 256, 0, st>>>(img, H, W, nimg);
  else norm_kernel<SFD2_IMG_U8_NHWC><<<cdiv(H * W, 256), 256, 0, st>>>(img, H, W, nimg);
  // L.w_simt is [tap][ci][cout_pad = 64] fp32 = exactly the [27][64] table the kernels stage in smem
  dim3 grid(cdiv(W, 128), H), block(128);
  if (tc_out) {
    SFD2_CHECK(tm1a != nullptr, SFD2_ERR_ARG, "conv1a: store maps missing");
    const int smem = 1024 + 2 * C1_SEG * 128 + 7168 + 9 * 260 * 4;
    SFD2_CUDA(cudaFuncSetAttribute(conv1a_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   // per device: set on every launch (cheap)
    const int nseg = cdiv(W, C1_SEG) * H;
    conv1a_tc_kernel<<<std::min(nseg, 148 * 2), 256, smem, st>>>(nimg, H, W, L.w_simt, L.b_dev, tm1a[0], tm1a[1]);
  } else {
    conv1a_kernel<0><<<grid, block, 0, st>>>(nimg, H, W, out.Wp, L.w_simt, L.b_dev, out.f32, nullptr, nullptr);
  }
  g_launches += 2;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

// ------------------------------------------------------------------------------ generic fp32 conv
// C[pixel][co] = sum_{tap,ci} in[pixel@tap][ci] * w[tap][ci][co]   (implicit GEMM, no im2col buffer)
// block tile 128 pixels x 64 output channels, K step = 16 input channels of one tap.
constexpr int CF_BM = 128, CF_BN = 64, CF_BK = 16;

template <int KS>
__global__ void __launch_bounds__(256)
conv_f32_kernel(const float* __restrict__ in, int H, int W, int Wp_in, int Cin,
                const float* __restrict__ wt, const float* __restrict__ bias, int cout_pad,
                float* __restrict__ out, int Ho, int Wo, int Wp_out, int Cout, int out_C,
                int stride, int relu, const float* __restrict__ res) {
  __shared__ __align__(16) float As[CF_BK][CF_BM + 4];
  __shared__ __align__(16) float Bs[CF_BK][CF_BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * CF_BM;
  const int n0 = blockIdx.y * CF_BN;
  const int npix = Ho * Wo;
  // loader mapping: thread -> (pixel row, 8-channel half)
  const int lm = tid >> 1, lk = (tid & 1) * 8;
  const int lp = m0 + lm;
  const int loy = (lp < npix) ? lp / Wo : 0, lox = (lp < npix) ? lp % Wo : 0;
  // compute mapping: 16 (n) x 16 (m) threads, 8 pixels x 4 channels each
  const int tn = tid & 15, tm = tid >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  constexpr int PAD = KS / 2;
  for (int tap = 0; tap < KS * KS; ++tap) {
    const int ky = tap / KS, kx = tap % KS;
    const int iy = loy * stride + ky - PAD, ix = lox * stride + kx - PAD;
    const bool ok = (lp < npix) && iy >= 0 && iy < H && ix >= 0 && ix < W;
    const float* src = in + ((size_t)iy * Wp_in + ix) * Cin + lk;
    for (int c0 = 0; c0 < Cin; c0 += CF_BK) {
      float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
      if (ok) {
        a0 = __ldg(reinterpret_cast<const float4*>(src + c0));
        a1 = __ldg(reinterpret_cast<const float4*>(src + c0 + 4));
      }
      const float4 b = __ldg(reinterpret_cast<const float4*>(
          wt + ((size_t)tap * Cin + c0 + (tid >> 4)) * cout_pad + n0 + (tid & 15) * 4));
      __syncthreads();
      As[lk + 0][lm] = a0.x; As[lk + 1][lm] = a0.y; As[lk + 2][lm] = a0.z; As[lk + 3][lm] = a0.w;
      As[lk + 4][lm] = a1.x; As[lk + 5][lm] = a1.y; As[lk + 6][lm] = a1.z; As[lk + 7][lm] = a1.w;
      *reinterpret_cast<float4*>(&Bs[tid >> 4][(tid & 15) * 4]) = b;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < CF_BK; ++k) {
        const float4 av0 = *reinterpret_cast<const float4*>(&As[k][tm * 8]);
        const float4 av1 = *reinterpret_cast<const float4*>(&As[k][tm * 8 + 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
        const float a[8] = {av0.x, av0.y, av0.z, av0.w, av1.x, av1.y, av1.z, av1.w};
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
    }
  }
  const int co = n0 + tn * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int p = m0 + tm * 8 + i;
    if (p >= npix) continue;
    const int oy = p / Wo, ox = p % Wo;
    const size_t o = ((size_t)oy * Wp_out + ox) * out_C + co;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (co + j >= Cout) continue;
      float v = acc[i][j] + __ldg(bias + co + j);
      if (res) v += __ldg(res + o + j);
      if (relu) v = fmaxf(v, 0.f);
      out[o + j] = v;
    }
  }
}

// grouped 3x3, 256 channels, 32 groups of 8.  thread = (4 consecutive pixels) x (one group).
__global__ void __launch_bounds__(256)
gconv_f32_kernel(const float* __restrict__ in, int H, int W, int Wp, const float* __restrict__ wt /*[9][8][256]*/,
                 const float* __restrict__ bias, float* __restrict__ out, int relu) {
  extern __shared__ float ws[];  // [9][8][256] + bias[256]
  for (int i = threadIdx.x; i < 9 * 8 * 256; i += blockDim.x) ws[i] = wt[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) ws[9 * 8 * 256 + i] = bias[i];
  __syncthreads();
  const int g = threadIdx.x & 31;
  const int xq = blockIdx.x * 8 + (threadIdx.x >> 5);  // quad of 4 pixels along x
  const int y = blockIdx.y;
  const int x0 = xq * 4;
  if (x0 >= W) return;
  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = ws[9 * 8 * 256 + g * 8 + j];
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = y + ky - 1;
    if (iy < 0 || iy >= H) continue;
    // 6 input pixels x0-1 .. x0+4 of this row, 8 channels of group g
    float v[6][8];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const int ix = x0 + q - 1;
      if (ix >= 0 && ix < W) {
        const float4* s = reinterpret_cast<const float4*>(in + ((size_t)iy * Wp + ix) * 256 + g * 8);
        const float4 s0 = __ldg(s), s1 = __ldg(s + 1);
        v[q][0] = s0.x; v[q][1] = s0.y; v[q][2] = s0.z; v[q][3] = s0.w;
        v[q][4] = s1.x; v[q][5] = s1.y; v[q][6] = s1.z; v[q][7] = s1.w;
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) v[q][c] = 0.f;
      }
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const float* wrow = ws + ((ky * 3 + kx) * 8 + ci) * 256 + g * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(wrow);
        const float4 w1 = *reinterpret_cast<const float4*>(wrow + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(v[p + kx][ci], wv[j], acc[p][j]);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int x = x0 + p;
    if (x >= W) break;
    float4* o = reinterpret_cast<float4*>(out + ((size_t)y * Wp + x) * 256 + g * 8);
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = relu ? fmaxf(acc[p][j], 0.f) : acc[p][j];
    o[0] = make_float4(r[0], r[1], r[2], r[3]);
    o[1] = make_float4(r[4], r[5], r[6], r[7]);
  }
}

int launch_conv_simt(const Act& in, const Layer& L, Act out, const Act* res, cudaStream_t st) {
  SFD2_CHECK(in.f32 && out.f32, SFD2_ERR_ARG, "conv_simt(%s): fp32 buffers missing", L.name.c_str());
  SFD2_CHECK(in.C == L.cin, SFD2_ERR_ARG, "conv_simt(%s): cin %d != %d", L.name.c_str(), in.C, L.cin);
  if (L.groups == 32) {
    SFD2_CHECK(L.cin == 256 && L.cout == 256 && L.k == 3 && L.stride == 1, SFD2_ERR_WEIGHTS, "gconv shape");
    const size_t smem = (9 * 8 * 256 + 256) * sizeof(float);
    SFD2_CUDA(cudaFuncSetAttribute(gconv_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: set on every launch (cheap)
    dim3 grid(cdiv(cdiv(in.W, 4), 8), in.H);
    gconv_f32_kernel<<<grid, 256, smem, st>>>(in.f32, in.H, in.W, in.Wp, L.w_simt, L.b_dev, out.f32, L.relu);
  } else {
    SFD2_CHECK(L.groups == 1 && L.cin % CF_BK == 0, SFD2_ERR_WEIGHTS, "conv_simt(%s): unsupported", L.name.c_str());
    const int cout_pad = round_up(L.cout, 64);
    dim3 grid(cdiv(out.H * out.W, CF_BM), cout_pad / CF_BN);
    const float* r = res ? res->f32 : nullptr;
    if (L.k == 3)
      conv_f32_kernel<3><<<grid, 256, 0, st>>>(in.f32, in.H, in.W, in.Wp, L.cin, L.w_simt, L.b_dev, cout_pad,
                                                out.f32, out.H, out.W, out.Wp, L.cout, out.C, L.stride, L.relu, r);
    else
      conv_f32_kernel<1><<<grid, 256, 0, st>>>(in.f32, in.H, in.W, in.Wp, L.cin, L.w_simt, L.b_dev, cout_pad,
                                                out.f32, out.H, out.W, out.Wp, L.cout, out.C, L.stride, L.relu, r);
  }
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

// ------------------------------------------------------------------------------ ConvSta (always fp32)
// one warp per pixel (grid-stride), lane = 8 consecutive channels; each lane keeps its 24 weights in
// registers, so the only memory traffic is the coalesced read of the 256-channel pixel rows.
template <int TC_IN>
__global__ void __launch_bounds__(256)
sta_kernel(const float* __restrict__ in_f32, const __half* __restrict__ in_hi, const __half* __restrict__ in_lo,
           int H, int W, int Wp, const float* __restrict__ wt /*[256][64pad]*/, const float* __restrict__ bias,
           float* __restrict__ logits /*[H*W][3]*/) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float w[8][3];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int c = 0; c < 3; ++c) w[j][c] = __ldg(wt + (size_t)(lane * 8 + j) * 64 + c);
  const float b = (lane < 3) ? __ldg(bias + lane) : 0.f;
  // 4 pixels per iteration: 8 independent 16-byte loads per lane in flight (the kernel is pure HBM streaming)
  for (int pix0 = warp0 * 4; pix0 < H * W; pix0 += nwarps * 4) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pix = min(pix0 + u, H * W - 1);
      const int y = pix / W, x = pix - y * W;
      const size_t base = ((size_t)y * Wp + x) * 256 + lane * 8;
      if (TC_IN) {
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(in_hi + base));
        const uint4 l = (TC_IN == 1) ? __ldg(reinterpret_cast<const uint4*>(in_lo + base)) : make_uint4(0, 0, 0, 0);
        const __half* hh = reinterpret_cast<const __half*>(&h);
        const __half* ll = reinterpret_cast<const __half*>(&l);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[u][j] = __half2float(hh[j]) + (TC_IN == 1 ? __half2float(ll[j]) : 0.f);
      } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(in_f32 + base));
        const float4 c = __ldg(reinterpret_cast<const float4*>(in_f32 + base + 4));
        v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = c.x; v[u][5] = c.y; v[u][6] = c.z; v[u][7] = c.w;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float s[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) s[c] = fmaf(v[u][j], w[j][c], s[c]);
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
      const int pix = pix0 + u;
      if (lane < 3 && pix < H * W) logits[(size_t)pix * 3 + lane] = (lane == 0 ? s[0] : (lane == 1 ? s[1] : s[2])) + b;
    }
  }
}

int launch_sta(const Act& in, int tc_in, const Layer& L, float* logits, cudaStream_t st) {
  SFD2_CHECK(L.cin == 256 && L.cout == 3 && L.k == 1, SFD2_ERR_WEIGHTS, "sta layer shape");
  const int blocks = 148 * 8;
  if (tc_in == 1) sta_kernel<1><<<blocks, 256, 0, st>>>(nullptr, in.hi, in.lo, in.H, in.W, in.Wp, L.w_simt, L.b_dev, logits);
  else if (tc_in == 2) sta_kernel<2><<<blocks, 256, 0, st>>>(nullptr, in.hi, in.lo, in.H, in.W, in.Wp, L.w_simt, L.b_dev, logits);
  else sta_kernel<0><<<blocks, 256, 0, st>>>(in.f32, nullptr, nullptr, in.H, in.W, in.Wp, L.w_simt, L.b_dev, logits);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

// ------------------------------------------------------------------------------ head epilogues
// logits [npix][80] (65 valid) -> semi_norm [npix][64]: exp, / (sum65 + 1e-5), drop dustbin (sfd2.py:330-333)
__global__ void softmax65_kernel(const float* __restrict__ logits, int npix, float* __restrict__ semi) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= npix) return;
  const float* r = logits + (size_t)warp * 80;
  const float e0 = expf(r[lane]), e1 = expf(r[lane + 32]);
  const float e2 = (lane == 0) ? expf(r[64]) : 0.f;
  float s = e0 + e1 + e2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float d = s + 0.00001f;
  semi[(size_t)warp * 64 + lane] = __fdiv_rn(e0, d);
  semi[(size_t)warp * 64 + lane + 32] = __fdiv_rn(e1, d);
}

int launch_softmax65(const float* logits, int npix, float* semi, cudaStream_t st) {
  softmax65_kernel<<<cdiv(npix * 32, 256), 256, 0, st>>>(logits, npix, semi);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

// F.normalize(desc, dim=1): x / max(||x||, 1e-12), in place on [npix][128] (sfd2.py:342)
__global__ void l2norm128_kernel(float* __restrict__ desc, int npix) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= npix) return;
  float4* p = reinterpret_cast<float4*>(desc + (size_t)warp * 128) + lane;
  float4 v = *p;
  float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float n = fmaxf(sqrtf(s), 1e-12f);
  v.x = __fdiv_rn(v.x, n); v.y = __fdiv_rn(v.y, n); v.z = __fdiv_rn(v.z, n); v.w = __fdiv_rn(v.w, n);
  *p = v;
}

int launch_l2norm128(float* desc, int npix, cudaStream_t st) {
  l2norm128_kernel<<<cdiv(npix * 32, 256), 256, 0, st>>>(desc, npix);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

}  // namespace sfd2
