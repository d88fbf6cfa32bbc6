// conv1a (3 -> 64 channels, 3x3, nets/sfd2.py:268) on the tensor cores, with the RGB normalisation
// (nets/extractor.py:14-17,104) fused into its operand producer.
//
// K = 27 is too thin for a TMA-fed implicit GEMM, and on the CUDA cores the layer is bound by the FP32
// FMA pipe (3.3 GFMA per 1600x1200 image; 3-register FFMA issues every other cycle per SM sub-partition:
// 0.26 ms measured for the register-tiled kernel in simt_conv.cu).  Here the CUDA cores only build the
// im2col operand - for a 128-pixel row segment, row r of the A tile holds the 27 normalised inputs of pixel r
// (K order = tap-major, channel-minor; padded to 32 with zeros) as fp16 hi / lo planes in the same
// 128-byte-swizzled K-major layout TMA would produce - and one thread issues the tcgen05 MMAs
// (M = 128 pixels, N = 64 channels, K = 32: 2 steps per pass, 3 passes in exact mode).  The epilogue is the usual
// TMEM -> bias + ReLU -> fp16 hi/lo -> swizzled staging tile -> TMA store; the staging tile IS the A tile
// (the MMAs have finished reading it by then), which keeps the block at 54 KB of shared memory.
//
// One 128-thread block = one segment at a time, persistent over segments; the phases of a segment are a serial
// latency chain (load -> build -> MMA -> TMEM read -> store), so 4 blocks per SM overlap each other's phases and
// each block prefetches the raw pixels of its next segment into registers while its MMAs run.  The layer's floor
// is the 491 MB write of the hi + lo planes.
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace sfd2 {

using namespace ptx;

constexpr int C1M_SEG = 128;
constexpr int C1M_COLS = C1M_SEG + 2;          // segment + 1-pixel apron
constexpr int C1M_ITEMS = 9 * C1M_COLS;        // (row, channel, column) values of a segment's input patch
constexpr int C1M_PER_THREAD = (C1M_ITEMS + 127) / 128;   // 10

struct Conv1aMmaArgs {
  int H, W, split, img_dtype;
  const void* img;         // raw image: f32 NCHW in [0,1] or u8 NHWC
  const __half* w_hi;      // [64 co][64 k] fp16, k = tap*3 + c (27 used)
  const __half* w_lo;
  const float* bias;       // [64]
};

__global__ void __launch_bounds__(128, 4)
conv1a_mma_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                  const __grid_constant__ Conv1aMmaArgs a) {
  extern __shared__ uint8_t smem_raw_c1[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_c1) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA_hi = base;                 // [128 px][128 B]: im2col rows (K 0..31), then the output tile [128 px][64 ch]
  uint8_t* sA_lo = base + 16384;
  uint8_t* sB_hi = base + 32768;         // [64 co][128 B]
  uint8_t* sB_lo = base + 40960;
  float* patch = reinterpret_cast<float*>(base + 49152);   // [3 rows][3 ch][132]
  float* sbias = patch + 9 * 132;                          // [64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sbias + 64);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  // one-time setup: stage the weight tiles (swizzled), bias, TMEM, barrier
  for (int i = tid; i < 64 * 8; i += blockDim.x) {          // 64 rows x 8 chunks of 16 B
    const int n = i >> 3, j = i & 7;
    const int dst = n * 128 + ((j ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(sB_hi + dst) = __ldg(reinterpret_cast<const uint4*>(a.w_hi + n * 64) + j);
    *reinterpret_cast<uint4*>(sB_lo + dst) = __ldg(reinterpret_cast<const uint4*>(a.w_lo + n * 64) + j);
  }
  if (tid < 64) sbias[tid] = __ldg(a.bias + tid);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t idesc = make_idesc_f16(128, 64);
  const int segs_x = (a.W + C1M_SEG - 1) / C1M_SEG, nseg = segs_x * a.H;
  const size_t plane = (size_t)a.H * a.W;
  uint32_t phase = 0;
  const int r = tid;                                        // pixel of the segment = A row = TMEM lane

  // raw (un-normalised) input values of a segment's patch: item i = (ky*3 + c) * 130 + col, this thread owns
  // items tid, tid + 128, ...; out-of-image positions are flagged and become the conv's zero padding
  float raw[C1M_PER_THREAD];
  unsigned inb = 0;
  auto prefetch = [&](int seg) {
    const int y = seg / segs_x, x0 = (seg - y * segs_x) * C1M_SEG;
    inb = 0;
#pragma unroll
    for (int t = 0; t < C1M_PER_THREAD; ++t) {
      const int i = tid + 128 * t;
      const int kyc = i / C1M_COLS, col = i - kyc * C1M_COLS;
      const int ky = kyc / 3, c = kyc - ky * 3;
      const int iy = y + ky - 1, ix = x0 + col - 1;
      raw[t] = 0.f;
      if (i < C1M_ITEMS && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
        inb |= 1u << t;
        if (a.img_dtype == SFD2_IMG_F32_NCHW) raw[t] = __ldg(reinterpret_cast<const float*>(a.img) + (size_t)c * plane + (size_t)iy * a.W + ix);
        else raw[t] = (float)__ldg(reinterpret_cast<const unsigned char*>(a.img) + ((size_t)iy * a.W + ix) * 3 + c);
      }
    }
  };
  if ((int)blockIdx.x < nseg) prefetch(blockIdx.x);

  for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    const int y = seg / segs_x, x0 = (seg - y * segs_x) * C1M_SEG;
    // normalise exactly like the reference ((x - mean) / std, IEEE division; u8 / 255 first) and stage the patch
#pragma unroll
    for (int t = 0; t < C1M_PER_THREAD; ++t) {
      const int i = tid + 128 * t;
      if (i < C1M_ITEMS) {
        const int kyc = i / C1M_COLS, col = i - kyc * C1M_COLS;
        const int c = kyc % 3;
        const float mean = (c == 0) ? 0.485f : (c == 1 ? 0.456f : 0.406f);
        const float stdv = (c == 0) ? 0.229f : (c == 1 ? 0.224f : 0.225f);
        const float x = (a.img_dtype == SFD2_IMG_F32_NCHW) ? raw[t] : __fdiv_rn(raw[t], 255.0f);
        patch[kyc * 132 + col] = ((inb >> t) & 1u) ? __fdiv_rn(__fsub_rn(x, mean), stdv) : 0.f;
      }
    }
    if (tid == 0) bulk_wait_read<0>();                      // the previous segment's stores have read the A / staging tiles
    __syncthreads();
    // im2col row of pixel r: k = (ky*3 + kx)*3 + c, 27 values + 5 zeros = 4 chunks of 8 halfs per plane
    {
      __align__(16) __half hi[32];
      __align__(16) __half lo[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float v = 0.f;
        if (k < 27) {
          const int tap = k / 3, c = k - tap * 3, ky = tap / 3, kx = tap - ky * 3;
          v = patch[(ky * 3 + c) * 132 + r + kx];
        }
        hi[k] = __float2half_rn(v);
        lo[k] = __float2half_rn(v - __half2float(hi[k]));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int dst = r * 128 + ((j ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(sA_hi + dst) = reinterpret_cast<const uint4*>(hi)[j];
        if (a.split == 3) *reinterpret_cast<uint4*>(sA_lo + dst) = reinterpret_cast<const uint4*>(lo)[j];
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t da_hi = make_desc_sw128(smem_u32(sA_hi)), da_lo = make_desc_sw128(smem_u32(sA_lo));
      const uint64_t db_hi = make_desc_sw128(smem_u32(sB_hi)), db_lo = make_desc_sw128(smem_u32(sB_lo));
#pragma unroll
      for (int k = 0; k < 2; ++k) umma_f16(tmem, desc_advance_k(da_hi, k), desc_advance_k(db_hi, k), idesc, k ? 1u : 0u);
      if (a.split == 3) {
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tmem, desc_advance_k(da_hi, k), desc_advance_k(db_lo, k), idesc, 1u);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tmem, desc_advance_k(da_lo, k), desc_advance_k(db_hi, k), idesc, 1u);
      }
      umma_commit(bar);
    }
    __syncwarp();
    if (seg + (int)gridDim.x < nseg) prefetch(seg + gridDim.x);   // loads fly while the MMAs run and the epilogue drains
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
      tmem_ld_wait();
      __align__(16) __half2 hi[16];
      __align__(16) __half2 lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float v0 = fmaxf(__uint_as_float(v[2 * j]) + sbias[ch * 32 + 2 * j], 0.f);
        const float v1 = fmaxf(__uint_as_float(v[2 * j + 1]) + sbias[ch * 32 + 2 * j + 1], 0.f);
        hi[j] = __floats2half2_rn(v0, v1);
        const float2 hf = __half22float2(hi[j]);
        lo[j] = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
      }
      // the MMAs are complete (barrier above), so the A tiles are free to become the output staging tiles
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int dst = r * 128 + (((ch * 4 + g) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(sA_hi + dst) = reinterpret_cast<const uint4*>(hi)[g];
        if (a.split == 3) *reinterpret_cast<uint4*>(sA_lo + dst) = reinterpret_cast<const uint4*>(lo)[g];
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();                                        // staging complete; TMEM + patch reusable
    if (tid == 0) {
      tma_store_3d(&tm_hi, sA_hi, 0, x0, y);
      if (a.split == 3) tma_store_3d(&tm_lo, sA_lo, 0, x0, y);   // the single-pass mode never reads lo planes
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// weights of conv1a as the B operand: [64 co][64 k] fp16 hi / lo, k = tap*3 + c
int conv1a_mma_encode(Layer& L) {
  std::vector<__half> hi(64 * 64, __float2half_rn(0.f)), lo(64 * 64, __float2half_rn(0.f));
  for (int co = 0; co < 64; ++co)
    for (int c = 0; c < 3; ++c)
      for (int t = 0; t < 9; ++t) {
        const float v = L.w[((size_t)co * 3 + c) * 9 + t];
        const __half h = __float2half_rn(v);
        hi[co * 64 + t * 3 + c] = h;
        lo[co * 64 + t * 3 + c] = __float2half_rn(v - __half2float(h));
      }
  SFD2_CUDA(cudaMalloc(&L.w_hi, hi.size() * sizeof(__half)));
  SFD2_CUDA(cudaMalloc(&L.w_lo, lo.size() * sizeof(__half)));
  SFD2_CUDA(cudaMemcpy(L.w_hi, hi.data(), hi.size() * sizeof(__half), cudaMemcpyHostToDevice));
  SFD2_CUDA(cudaMemcpy(L.w_lo, lo.data(), lo.size() * sizeof(__half), cudaMemcpyHostToDevice));
  return SFD2_OK;
}

// tm1a: [hi, lo] store maps of the conv1a output with box {64 ch, 128 px, 1 row}
int launch_conv1a_mma(const void* img, int img_dtype, int H, int W, const Layer& L, const CUtensorMap* tm1a, int split,
                      int num_sms, cudaStream_t st) {
  SFD2_CHECK(L.w_hi && L.w_lo && tm1a, SFD2_ERR_ARG, "conv1a_mma: weights / store maps missing");
  Conv1aMmaArgs a{H, W, split, img_dtype, img, L.w_hi, L.w_lo, L.b_dev};
  const int smem = 1024 + 49152 + (9 * 132 + 64) * 4 + 64;
  SFD2_CUDA(cudaFuncSetAttribute(conv1a_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   // per device: set on every launch (cheap)
  const int nseg = cdiv(W, C1M_SEG) * H;
  conv1a_mma_kernel<<<std::min(nseg, 4 * num_sms), 128, smem, st>>>(tm1a[0], tm1a[1], a);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

}  // namespace sfd2
