// Post-processing kernels: heat-map assembly, 3-round radius-4 NMS with fused candidate
// compaction, top-K selection, bilinear descriptor sampling.  All are HBM/latency-bound
// compare/gather work on CUDA cores (no tensor-core shape to exploit).
#include <math_constants.h>

#include "common.cuh"

namespace sfd2 {

// ------------------------------------------------------------------------------ heat-map
// PyTorch upsample_bilinear2d (align_corners=False) source index + lambda for one axis.
__device__ __forceinline__ void lin_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  if (src < 0.f) src = 0.f;
  i0 = min((int)floorf(src), in_size - 1);
  l1 = fminf(fmaxf(__fsub_rn(src, (float)i0), 0.f), 1.f);
  l0 = __fsub_rn(1.f, l1);
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
}

__device__ __forceinline__ float bil(float v00, float v01, float v10, float v11, float lw0, float lw1, float lh0,
                                     float lh1) {
  const float top = __fadd_rn(__fmul_rn(v00, lw0), __fmul_rn(v01, lw1));
  const float bot = __fadd_rn(__fmul_rn(v10, lw0), __fmul_rn(v11, lw1));
  return __fadd_rn(__fmul_rn(top, lh0), __fmul_rn(bot, lh1));
}

// semi [H8][W8][64] holds the softmax cells; the full-resolution score of pixel (Y, X) of the
// (8*H8) x (8*W8) map is semi[Y/8][X/8][(Y%8)*8 + X%8]  (depth-to-space, nets/sfd2.py:333-337).
__device__ __forceinline__ float score_at(const float* __restrict__ semi, int W8, int Y, int X) {
  return __ldg(semi + ((size_t)(Y >> 3) * W8 + (X >> 3)) * 64 + ((Y & 7) << 3) + (X & 7));
}

// heat[y][x] = score (bilinearly resized to (H, W) when 8*H8 != H or 8*W8 != W, extractor.py:137-138)
//              * {0.1, 0.5, 1.0}[argmax_c bilinear(sta_logits)[c]]   (sfd2.py:345-347, extractor.py:141)
__global__ void heat_kernel(const float* __restrict__ semi, int H8, int W8, const float* __restrict__ sta, int H4,
                            int W4, int use_sta, float* __restrict__ heat, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const int HS = H8 * 8, WS = W8 * 8;
  float s;
  if (HS == H && WS == W) {
    s = score_at(semi, W8, y, x);
  } else {
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    if (HS == H) { y0 = y1 = y; ly0 = 1.f; ly1 = 0.f; } else lin_src(y, __fdiv_rn((float)HS, (float)H), HS, y0, y1, ly0, ly1);
    if (WS == W) { x0 = x1 = x; lx0 = 1.f; lx1 = 0.f; } else lin_src(x, __fdiv_rn((float)WS, (float)W), WS, x0, x1, lx0, lx1);
    s = bil(score_at(semi, W8, y0, x0), score_at(semi, W8, y0, x1), score_at(semi, W8, y1, x0),
            score_at(semi, W8, y1, x1), lx0, lx1, ly0, ly1);
  }
  if (use_sta) {
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    if (H4 == H) { y0 = y1 = y; ly0 = 1.f; ly1 = 0.f; } else lin_src(y, __fdiv_rn((float)H4, (float)H), H4, y0, y1, ly0, ly1);
    if (W4 == W) { x0 = x1 = x; lx0 = 1.f; lx1 = 0.f; } else lin_src(x, __fdiv_rn((float)W4, (float)W), W4, x0, x1, lx0, lx1);
    const float* p00 = sta + ((size_t)y0 * W4 + x0) * 3;
    const float* p01 = sta + ((size_t)y0 * W4 + x1) * 3;
    const float* p10 = sta + ((size_t)y1 * W4 + x0) * 3;
    const float* p11 = sta + ((size_t)y1 * W4 + x1) * 3;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = bil(__ldg(p00 + c), __ldg(p01 + c), __ldg(p10 + c), __ldg(p11 + c), lx0, lx1, ly0, ly1);
    int cls = 0;                      // torch.max: first maximal index
    if (v[1] > v[cls]) cls = 1;
    if (v[2] > v[cls]) cls = 2;
    const float stab = (cls == 0) ? 0.1f : (cls == 1 ? 0.5f : 1.0f);
    s = __fmul_rn(s, stab);
  }
  heat[(size_t)y * W + x] = s;
}

int launch_heat(const float* semi, int H8, int W8, const float* sta, int H4, int W4, int use_sta, float* heat,
                int H, int W, cudaStream_t st) {
  dim3 block(32, 8), grid(cdiv(W, 32), cdiv(H, 8));
  heat_kernel<<<grid, block, 0, st>>>(semi, H8, W8, sta, H4, W4, use_sta, heat, H, W);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

// ------------------------------------------------------------------------------ NMS (radius 4, 3 rounds)
// Literal restatement of simple_nms (nets/extractor.py:20-35) on one smem tile with a 20-pixel
// halo: each of the 5 max-pools / dilations reaches 4 px and the dependency chain is 4+8+8.
// Out-of-image pixels hold -inf, which is max_pool2d's implicit padding value, and are never
// allowed into the mask.  Windows are clamped to the tile, so only the interior (>= 20 px from
// the tile edge) is exact - and only the interior is written.
constexpr int NT_W = 64, NT_H = 32, NHALO = 20, NR = 4;
constexpr int NS_W = NT_W + 2 * NHALO;  // 104
constexpr int NS_H = NT_H + 2 * NHALO;  // 72
constexpr int NS_N = NS_W * NS_H;

// dst = 9x9 max of src (row pass into tmp, column pass into dst)
__device__ __forceinline__ void pool9_f(const float* __restrict__ src, float* __restrict__ tmp,
                                        float* __restrict__ dst) {
  for (int i = threadIdx.x; i < NS_N; i += blockDim.x) {
    const int y = i / NS_W, x = i - y * NS_W;
    const int a = max(x - NR, 0), b = min(x + NR, NS_W - 1);
    float mx = src[y * NS_W + a];
    for (int k = a + 1; k <= b; ++k) mx = fmaxf(mx, src[y * NS_W + k]);
    tmp[i] = mx;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NS_N; i += blockDim.x) {
    const int y = i / NS_W, x = i - y * NS_W;
    const int a = max(y - NR, 0), b = min(y + NR, NS_H - 1);
    float mx = tmp[a * NS_W + x];
    for (int k = a + 1; k <= b; ++k) mx = fmaxf(mx, tmp[k * NS_W + x]);
    dst[i] = mx;
  }
  __syncthreads();
}

// dst = 9x9 dilation of the 0/1 mask src
__device__ __forceinline__ void dilate9(const unsigned char* __restrict__ src, unsigned char* __restrict__ tmp,
                                        unsigned char* __restrict__ dst) {
  for (int i = threadIdx.x; i < NS_N; i += blockDim.x) {
    const int y = i / NS_W, x = i - y * NS_W;
    const int a = max(x - NR, 0), b = min(x + NR, NS_W - 1);
    unsigned char mx = 0;
    for (int k = a; k <= b; ++k) mx |= src[y * NS_W + k];
    tmp[i] = mx;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NS_N; i += blockDim.x) {
    const int y = i / NS_W, x = i - y * NS_W;
    const int a = max(y - NR, 0), b = min(y + NR, NS_H - 1);
    unsigned char mx = 0;
    for (int k = a; k <= b; ++k) mx |= tmp[k * NS_W + x];
    dst[i] = mx;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512)
nms_kernel(const float* __restrict__ heat, int H, int W, float conf_th, int border, float* __restrict__ nms_out,
           unsigned long long* __restrict__ cand, int cap, int* __restrict__ counter) {
  extern __shared__ float sm[];
  float* s = sm;             // scores (-inf outside the image)
  float* t = sm + NS_N;      // row-pass scratch
  float* u = sm + 2 * NS_N;  // pooled scores / suppressed scores
  unsigned char* m = reinterpret_cast<unsigned char*>(sm + 3 * NS_N);  // max_mask
  unsigned char* tb = m + NS_N;                                        // dilation scratch
  unsigned char* supp = tb + NS_N;                                     // supp_mask
  const int X0 = blockIdx.x * NT_W - NHALO, Y0 = blockIdx.y * NT_H - NHALO;
  for (int i = threadIdx.x; i < NS_N; i += blockDim.x) {
    const int yy = i / NS_W;
    const int y = Y0 + yy, x = X0 + (i - yy * NS_W);
    s[i] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(heat + (size_t)y * W + x) : -CUDART_INF_F;
  }
  __syncthreads();
  pool9_f(s, t, u);                                                    // max_pool(scores)
  for (int i = threadIdx.x; i < NS_N; i += blockDim.x) m[i] = (s[i] != -CUDART_INF_F) && (s[i] == u[i]);
  __syncthreads();
  for (int round = 0; round < 2; ++round) {
    dilate9(m, tb, supp);                                              // supp_mask = max_pool(max_mask) > 0
    for (int i = threadIdx.x; i < NS_N; i += blockDim.x) {
      const float sv = s[i];
      u[i] = (supp[i] && sv != -CUDART_INF_F) ? 0.f : sv;              // supp_scores
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NS_N; i += blockDim.x) {             // row pass of max_pool(supp_scores)
      const int y = i / NS_W, x = i - y * NS_W;
      const int a = max(x - NR, 0), b = min(x + NR, NS_W - 1);
      float mx = u[y * NS_W + a];
      for (int k = a + 1; k <= b; ++k) mx = fmaxf(mx, u[y * NS_W + k]);
      t[i] = mx;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NS_N; i += blockDim.x) {             // column pass + mask update
      const int y = i / NS_W, x = i - y * NS_W;
      const int a = max(y - NR, 0), b = min(y + NR, NS_H - 1);
      float mx = t[a * NS_W + x];
      for (int k = a + 1; k <= b; ++k) mx = fmaxf(mx, t[k * NS_W + x]);
      if (s[i] != -CUDART_INF_F && u[i] == mx && !supp[i]) m[i] = 1;   // max_mask |= new_max & ~supp
    }
    __syncthreads();
  }
  // output + candidate compaction for the tile interior
  for (int i = threadIdx.x; i < NT_W * NT_H; i += blockDim.x) {
    const int ty = i / NT_W, tx = i - ty * NT_W;
    const int y = blockIdx.y * NT_H + ty, x = blockIdx.x * NT_W + tx;
    const bool in_img = (y < H && x < W);
    const int si = (ty + NHALO) * NS_W + tx + NHALO;
    const float v = (in_img && m[si]) ? s[si] : 0.f;
    if (in_img && nms_out) nms_out[(size_t)y * W + x] = v;
    const bool is_cand = in_img && (v > conf_th) && x >= border && x < W - border && y >= border && y < H - border;
    const unsigned ball = __ballot_sync(0xffffffffu, is_cand);
    if (is_cand) {
      const int lane = threadIdx.x & 31;
      const int leader = __ffs(ball) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(counter, __popc(ball));
      base = __shfl_sync(ball, base, leader);
      const int pos = base + __popc(ball & ((1u << lane) - 1));
      if (pos < cap) {
        const unsigned lin = (unsigned)(y * W + x);
        cand[pos] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - lin);
      }
    }
  }
}

int launch_nms(const float* heat, int H, int W, float conf_th, int border, float* nms_out, unsigned long long* cand,
               int cap, int* counter, cudaStream_t st) {
  const size_t smem = (size_t)3 * NS_N * sizeof(float) + 3 * NS_N;
  static bool attr = false;
  if (!attr) {
    SFD2_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  SFD2_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
  dim3 grid(cdiv(W, NT_W), cdiv(H, NT_H));
  nms_kernel<<<grid, 512, smem, st>>>(heat, H, W, conf_th, border, nms_out, cand, cap, counter);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

// ------------------------------------------------------------------------------ top-K selection
// Candidates are 64-bit keys (score bits << 32 | ~pixel index): positive floats order like their
// bit patterns, and the inverted index makes the lower pixel index win among exactly equal scores
// (the reference's order there is np.argsort-unstable, nets/extractor.py:176,323).  One CTA sorts
// them descending (bitonic) - in shared memory when they fit, otherwise in the global scratch -
// and emits the first K as (x, y), score.
constexpr int SEL_SMEM_KEYS = 16384;

__device__ __forceinline__ void bitonic_desc(unsigned long long* a, int n) {  // n = power of two
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long x = a[i], y = a[p];
          const bool desc = ((i & k) == 0);
          if (desc ? (x < y) : (x > y)) { a[i] = y; a[p] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(1024)
select_kernel(const unsigned long long* __restrict__ cand, int cap, const int* __restrict__ counter, int W, int topk,
              float* __restrict__ kpts, float* __restrict__ scores, int32_t* __restrict__ count_out,
              int* __restrict__ status, unsigned long long* __restrict__ scratch) {
  extern __shared__ unsigned long long keys[];
  int n = *counter;
  if (n > cap) {                       // more candidates than the workspace holds: report, keep what fits
    if (threadIdx.x == 0) *status = SFD2_ERR_OVERFLOW;
    n = cap;
  }
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  unsigned long long* a = (n2 <= SEL_SMEM_KEYS) ? keys : scratch;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) a[i] = (i < n) ? cand[i] : 0ull;
  __syncthreads();
  bitonic_desc(a, n2);
  const int k = (topk > 0 && topk < n) ? topk : n;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const unsigned long long key = a[i];
    const unsigned lin = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
    kpts[2 * i] = (float)(lin % (unsigned)W);
    kpts[2 * i + 1] = (float)(lin / (unsigned)W);
    scores[i] = __uint_as_float((unsigned)(key >> 32));
  }
  if (threadIdx.x == 0) *count_out = k;
}

int launch_select(unsigned long long* cand, int cap, const int* counter, int W, int topk, float* kpts, float* scores,
                  int32_t* count_out, int* status, unsigned long long* scratch, cudaStream_t st) {
  const size_t smem = (size_t)SEL_SMEM_KEYS * sizeof(unsigned long long);
  static bool attr = false;
  if (!attr) {
    SFD2_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  select_kernel<<<1, 1024, smem, st>>>(cand, cap, counter, W, topk, kpts, scores, count_out, status, scratch);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

// ------------------------------------------------------------------------------ descriptor sampling
// F.grid_sample(coarse_desc, (x/(W/2)-1, y/(H/2)-1)) bilinear, zeros padding, align_corners=False,
// then /= L2 norm (nets/extractor.py:199-208).  One warp per keypoint, 4 channels per lane.
__global__ void sample_kernel(const float* __restrict__ desc_map, int H4, int W4, int H, int W,
                              const float* __restrict__ kpts, const int32_t* __restrict__ count, int topk,
                              float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= topk) return;
  float4* o = reinterpret_cast<float4*>(out + (size_t)warp * 128) + lane;
  if (warp >= *count) { *o = make_float4(0.f, 0.f, 0.f, 0.f); return; }
  const float px = kpts[2 * warp], py = kpts[2 * warp + 1];
  // normalised grid coordinate exactly as the reference builds it (float32 ops)
  const float gx = __fsub_rn(__fdiv_rn(px, __fdiv_rn((float)W, 2.f)), 1.f);
  const float gy = __fsub_rn(__fdiv_rn(py, __fdiv_rn((float)H, 2.f)), 1.f);
  // grid_sampler_unnormalize, align_corners=False: ((g + 1) * size - 1) / 2
  const float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W4), 1.f), 2.f);
  const float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H4), 1.f), 2.f);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - fx, wy1 = iy - fy, wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy;
  const float w_nw = wx0 * wy0, w_ne = wx1 * wy0, w_sw = wx0 * wy1, w_se = wx1 * wy1;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto tap = [&](int yy, int xx, float wgt) {
    if (yy >= 0 && yy < H4 && xx >= 0 && xx < W4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(desc_map + ((size_t)yy * W4 + xx) * 128) + lane);
      acc.x = fmaf(v.x, wgt, acc.x); acc.y = fmaf(v.y, wgt, acc.y);
      acc.z = fmaf(v.z, wgt, acc.z); acc.w = fmaf(v.w, wgt, acc.w);
    }
  };
  tap(y0, x0, w_nw); tap(y0, x1, w_ne); tap(y1, x0, w_sw); tap(y1, x1, w_se);
  float ss = acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, k);
  const float nrm = sqrtf(ss);
  *o = make_float4(__fdiv_rn(acc.x, nrm), __fdiv_rn(acc.y, nrm), __fdiv_rn(acc.z, nrm), __fdiv_rn(acc.w, nrm));
}

int launch_sample(const float* desc_map, int H4, int W4, int H, int W, const float* kpts, const int32_t* count,
                  int topk, float* desc_out, cudaStream_t st) {
  if (topk <= 0) return SFD2_OK;
  sample_kernel<<<cdiv(topk * 32, 256), 256, 0, st>>>(desc_map, H4, W4, H, W, kpts, count, topk, desc_out);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

}  // namespace sfd2
