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
// One 128-thread block owns a 128-pixel-wide column strip over a range of rows and walks DOWN it: the three
// input rows of an output row live in a ring of three row buffers, so every step normalises and splits only the
// ONE new row (390 values per block instead of 1170), once, into packed (hi, lo) half pairs - the im2col build
// is then 27 shared loads and 32 byte-permutes per pixel, no conversions.  The phases of a step are a serial
// latency chain (load -> build -> MMA -> TMEM read -> store), so 4 blocks per SM overlap each other's phases and
// each block prefetches the raw pixels of its next row into registers while its MMAs run.  The layer's floor
// is the 491 MB write of the hi + lo planes.
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace sfd2 {

using namespace ptx;

constexpr int C1M_SEG = 128;
constexpr int C1M_COLS = C1M_SEG + 2;          // segment + 1-pixel apron
constexpr int C1M_ROW_ITEMS = 3 * C1M_COLS;    // (channel, column) values of one input row of the strip
constexpr int C1M_PER_THREAD = (C1M_ROW_ITEMS + 127) / 128;   // 4
constexpr int C1M_PITCH = 132;                 // words per (row slot, channel)

struct Conv1aMmaArgs {
  int H, W, split, img_dtype;
  int rows_per_block;      // output rows a block walks through
  int y_begin, y_end;      // output rows of this launch (a row band of the image; the whole image = [0, H))
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
  uint32_t* patch = reinterpret_cast<uint32_t*>(base + 49152);   // [3 row slots][3 ch][132]: lo half << 16 | hi half
  float* sbias = reinterpret_cast<float*>(patch + 9 * C1M_PITCH);   // [64]
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
  const int segs_x = (a.W + C1M_SEG - 1) / C1M_SEG;
  const int x0 = ((int)blockIdx.x % segs_x) * C1M_SEG;
  const int ya = a.y_begin + ((int)blockIdx.x / segs_x) * a.rows_per_block, yb = min(ya + a.rows_per_block, a.y_end);
  const size_t plane = (size_t)a.H * a.W;
  uint32_t phase = 0;
  const int r = tid;                                        // pixel of the segment = A row = TMEM lane

  // One input row of the strip = 390 (channel, column) items; this thread owns items tid, tid + 128, ...
  // fetch(): raw (un-normalised) values into registers, out-of-image positions flagged (they become the conv's
  // zero padding); stash(): normalise exactly like the reference ((x - mean) / std, IEEE division; u8 / 255
  // first), split into fp16 hi / lo once, and store the packed pair into row slot `rs` of the ring.
  float raw[C1M_PER_THREAD];
  unsigned inb = 0;
  auto fetch = [&](int iy) {
    inb = 0;
#pragma unroll
    for (int t = 0; t < C1M_PER_THREAD; ++t) {
      const int i = tid + 128 * t;
      const int c = i / C1M_COLS, col = i - c * C1M_COLS;
      const int ix = x0 + col - 1;
      raw[t] = 0.f;
      if (i < C1M_ROW_ITEMS && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
        inb |= 1u << t;
        if (a.img_dtype == SFD2_IMG_F32_NCHW) raw[t] = __ldg(reinterpret_cast<const float*>(a.img) + (size_t)c * plane + (size_t)iy * a.W + ix);
        else raw[t] = (float)__ldg(reinterpret_cast<const unsigned char*>(a.img) + ((size_t)iy * a.W + ix) * 3 + c);
      }
    }
  };
  auto stash = [&](int rs) {
#pragma unroll
    for (int t = 0; t < C1M_PER_THREAD; ++t) {
      const int i = tid + 128 * t;
      if (i < C1M_ROW_ITEMS) {
        const int c = i / C1M_COLS, col = i - c * C1M_COLS;
        const float mean = (c == 0) ? 0.485f : (c == 1 ? 0.456f : 0.406f);
        const float stdv = (c == 0) ? 0.229f : (c == 1 ? 0.224f : 0.225f);
        const float x = (a.img_dtype == SFD2_IMG_F32_NCHW) ? raw[t] : __fdiv_rn(raw[t], 255.0f);
        const float v = ((inb >> t) & 1u) ? __fdiv_rn(__fsub_rn(x, mean), stdv) : 0.f;
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        patch[(rs * 3 + c) * C1M_PITCH + col] = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
      }
    }
  };
  if (ya >= yb) goto done;                                  // (uniform per block)
  // ring slot of input row iy = (iy - (ya - 1)) % 3: rows ya-1 and ya go in first, row ya+1 is fetched for the first step
  fetch(ya - 1); stash(0);
  fetch(ya);     stash(1);
  fetch(ya + 1);
  {
  int s0 = 0;                                               // slot of input row y-1; rows y, y+1 follow cyclically
  for (int y = ya; y < yb; ++y) {
    const int s1 = (s0 + 1 == 3) ? 0 : s0 + 1, s2 = (s1 + 1 == 3) ? 0 : s1 + 1;
    stash(s2);                                              // row y+1 (its slot was last read two steps ago)
    if (tid == 0) bulk_wait_read<0>();                      // the previous step's stores have read the A / staging tiles
    __syncthreads();
    // im2col row of pixel r: k = (ky*3 + kx)*3 + c, 27 values + 5 zeros = 4 chunks of 8 halfs per plane
    {
      uint32_t w[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        w[k] = 0u;
        if (k < 27) {
          const int tap = k / 3, c = k - tap * 3, ky = tap / 3, kx = tap - ky * 3;
          const int rs = (ky == 0) ? s0 : (ky == 1 ? s1 : s2);
          w[k] = patch[(rs * 3 + c) * C1M_PITCH + r + kx];
        }
      }
      __align__(16) uint32_t hi[16];
      __align__(16) uint32_t lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        hi[j] = __byte_perm(w[2 * j], w[2 * j + 1], 0x5410);   // the two low halves  (k = 2j, 2j+1)
        lo[j] = __byte_perm(w[2 * j], w[2 * j + 1], 0x7632);   // the two high halves
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
    if (y + 1 < yb) fetch(y + 2);                           // loads fly while the MMAs run and the epilogue drains
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
    __syncthreads();                                        // staging complete; TMEM + the oldest row slot reusable
    if (tid == 0) {
      tma_store_3d(&tm_hi, sA_hi, 0, x0, y);
      if (a.split == 3) tma_store_3d(&tm_lo, sA_lo, 0, x0, y);   // the single-pass mode never reads lo planes
      bulk_commit();
    }
    s0 = s1;
  }
  }
done:
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
                      int num_sms, cudaStream_t st, int y_begin, int y_end) {
  SFD2_CHECK(L.w_hi && L.w_lo && tm1a, SFD2_ERR_ARG, "conv1a_mma: weights / store maps missing");
  if (y_end < 0) y_end = H;
  SFD2_CHECK(0 <= y_begin && y_begin < y_end && y_end <= H, SFD2_ERR_ARG, "conv1a_mma: bad row band [%d, %d)", y_begin, y_end);
  // one wave of 4 resident blocks per SM: column strips x row ranges
  const int segs_x = cdiv(W, C1M_SEG);
  const int band = y_end - y_begin;
  int rows_per_block = cdiv(band * segs_x, 4 * num_sms);
  if (rows_per_block < 4) rows_per_block = std::min(4, band);   // the two extra rows a block loads amortise over its range
  const int blocks = segs_x * cdiv(band, rows_per_block);
  Conv1aMmaArgs a{H, W, split, img_dtype, rows_per_block, y_begin, y_end, img, L.w_hi, L.w_lo, L.b_dev};
  const int smem = 1024 + 49152 + (9 * C1M_PITCH + 64) * 4 + 64;
  SFD2_CUDA(cudaFuncSetAttribute(conv1a_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   // per device: set on every launch (cheap)
  conv1a_mma_kernel<<<blocks, 128, smem, st>>>(tm1a[0], tm1a[1], a);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

}  // namespace sfd2
