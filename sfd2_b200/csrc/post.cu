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
// allowed into the mask.  The tile arrays carry a 4-element apron of the pooling identity so the
// 9-wide windows need no clamping; windows are still cut at the tile edge, so only the interior
// (>= 20 px from the tile edge) is exact - and only the interior is written.
//
// Each 9-wide running max is computed in registers for 8 outputs at a time from 16 loaded values
// with the doubling scheme m2 -> m4 -> m8 -> m9 (45 max ops and 16 smem reads per 8 outputs).
constexpr int NT_W = 64, NT_H = 32, NHALO = 20, NR = 4;
constexpr int NS_W = NT_W + 2 * NHALO;  // 104 (13 runs of 8)
constexpr int NS_H = NT_H + 2 * NHALO;  // 72  (9 runs of 8)
constexpr int NP_W = NS_W + 2 * NR;     // 112 padded pitch
constexpr int NP_H = NS_H + 2 * NR;     // 80
constexpr int NP_N = NP_W * NP_H;
__device__ __forceinline__ int nidx(int y, int x) { return (y + NR) * NP_W + x + NR; }  // tile coords -> padded index

template <typename T> __device__ __forceinline__ T pmax(T a, T b);
template <> __device__ __forceinline__ float pmax<float>(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ unsigned char pmax<unsigned char>(unsigned char a, unsigned char b) { return a | b; }

template <typename T>
__device__ __forceinline__ void run9(const T (&v)[16], T (&o)[8]) {
  T m2[15], m4[13], m8[9];
#pragma unroll
  for (int i = 0; i < 15; ++i) m2[i] = pmax(v[i], v[i + 1]);
#pragma unroll
  for (int i = 0; i < 13; ++i) m4[i] = pmax(m2[i], m2[i + 2]);
#pragma unroll
  for (int i = 0; i < 9; ++i) m8[i] = pmax(m4[i], m4[i + 4]);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = pmax(m8[i], v[i + 8]);
}

// dst = 9x9 max of src over the tile (src/tmp/dst are padded arrays whose apron holds the identity)
template <typename T>
__device__ __forceinline__ void pool9(const T* __restrict__ src, T* __restrict__ tmp, T* __restrict__ dst) {
  // row pass: work item = (row, run of 8 columns)
  for (int it = threadIdx.x; it < NS_H * (NS_W / 8); it += blockDim.x) {
    const int y = it / (NS_W / 8), x0 = (it - y * (NS_W / 8)) * 8;
    T v[16], o[8];
    const T* p = src + nidx(y, x0 - NR);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = p[i];
    run9(v, o);
    T* q = tmp + nidx(y, x0);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = o[i];
  }
  __syncthreads();
  // column pass: work item = (run of 8 rows, column); consecutive threads take consecutive columns
  for (int it = threadIdx.x; it < (NS_H / 8) * NS_W; it += blockDim.x) {
    const int yr = it / NS_W, x = it - yr * NS_W, y0 = yr * 8;
    T v[16], o[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = tmp[nidx(y0 - NR + i, x)];
    run9(v, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[nidx(y0 + i, x)] = o[i];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512)
nms_kernel(const float* __restrict__ heat, int H, int W, float conf_th, int border, int bw, int bh,
           float* __restrict__ nms_out,
           unsigned long long* __restrict__ cand, int cap, int* __restrict__ counter) {
  extern __shared__ float sm[];
  float* s = sm;             // scores (-inf outside the image)
  float* t = sm + NP_N;      // row-pass scratch
  float* u = sm + 2 * NP_N;  // pooled scores / suppressed scores
  unsigned char* m = reinterpret_cast<unsigned char*>(sm + 3 * NP_N);  // max_mask
  unsigned char* tb = m + NP_N;                                        // dilation scratch
  unsigned char* supp = tb + NP_N;                                     // supp_mask
  const int X0 = blockIdx.x * NT_W - NHALO, Y0 = blockIdx.y * NT_H - NHALO;
  for (int i = threadIdx.x; i < NP_N; i += blockDim.x) {
    const int py = i / NP_W, px = i - py * NP_W;
    const int ty = py - NR, tx = px - NR;
    const int y = Y0 + ty, x = X0 + tx;
    const bool in_tile = ty >= 0 && ty < NS_H && tx >= 0 && tx < NS_W;
    s[i] = (in_tile && y >= 0 && y < H && x >= 0 && x < W) ? __ldg(heat + (size_t)y * W + x) : -CUDART_INF_F;
    t[i] = -CUDART_INF_F;   // aprons of the scratch arrays: identity of max
    u[i] = -CUDART_INF_F;
    m[i] = 0; tb[i] = 0; supp[i] = 0;
  }
  __syncthreads();
  pool9<float>(s, t, u);                                               // max_pool(scores)
  for (int it = threadIdx.x; it < NS_H * NS_W; it += blockDim.x) {
    const int y = it / NS_W, x = it - y * NS_W, i = nidx(y, x);
    m[i] = (s[i] != -CUDART_INF_F) && (s[i] == u[i]);
  }
  __syncthreads();
  for (int round = 0; round < 2; ++round) {
    pool9<unsigned char>(m, tb, supp);                                 // supp_mask = max_pool(max_mask) > 0
    for (int it = threadIdx.x; it < NS_H * NS_W; it += blockDim.x) {
      const int y = it / NS_W, x = it - y * NS_W, i = nidx(y, x);
      const float sv = s[i];
      t[i] = (supp[i] && sv != -CUDART_INF_F) ? 0.f : sv;              // supp_scores (kept in t)
    }
    __syncthreads();
    // max_pool(supp_scores): row pass t -> u, column pass u -> compare in place
    for (int it = threadIdx.x; it < NS_H * (NS_W / 8); it += blockDim.x) {
      const int y = it / (NS_W / 8), x0 = (it - y * (NS_W / 8)) * 8;
      float v[16], o[8];
      const float* p = t + nidx(y, x0 - NR);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = p[i];
      run9(v, o);
      float* q = u + nidx(y, x0);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = o[i];
    }
    __syncthreads();
    for (int it = threadIdx.x; it < (NS_H / 8) * NS_W; it += blockDim.x) {
      const int yr = it / NS_W, x = it - yr * NS_W, y0 = yr * 8;
      float v[16], o[8];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = u[nidx(y0 - NR + i, x)];
      run9(v, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int id = nidx(y0 + i, x);
        if (s[id] != -CUDART_INF_F && t[id] == o[i] && !supp[id]) m[id] = 1;   // max_mask |= new_max & ~supp
      }
    }
    __syncthreads();
  }
  // output + candidate compaction for the tile interior
  for (int i = threadIdx.x; i < NT_W * NT_H; i += blockDim.x) {
    const int ty = i / NT_W, tx = i - ty * NT_W;
    const int y = blockIdx.y * NT_H + ty, x = blockIdx.x * NT_W + tx;
    const bool in_img = (y < H && x < W);
    const int si = nidx(ty + NHALO, tx + NHALO);
    const float v = (in_img && m[si]) ? s[si] : 0.f;
    if (in_img && nms_out) nms_out[(size_t)y * W + x] = v;
    const bool is_cand = in_img && (v > conf_th) && x >= border && x < bw - border && y >= border && y < bh - border;
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

// bw / bh: the extents the border test uses (the reference tests scaled coordinates against the ORIGINAL
// image size in its multi-scale loop, nets/extractor.py:181-182); pass W / H for the single-scale case
int launch_nms(const float* heat, int H, int W, float conf_th, int border, int bw, int bh, float* nms_out, unsigned long long* cand,
               int cap, int* counter, cudaStream_t st) {
  const size_t smem = (size_t)3 * NP_N * sizeof(float) + 3 * NP_N;
  SFD2_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: set on every launch (cheap)
  SFD2_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
  dim3 grid(cdiv(W, NT_W), cdiv(H, NT_H));
  nms_kernel<<<grid, 512, smem, st>>>(heat, H, W, conf_th, border, bw, bh, nms_out, cand, cap, counter);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

// ------------------------------------------------------------------------------ top-K selection
// Candidates are 64-bit keys (score bits << 32 | ~pixel index): positive floats order like their
// bit patterns, and the inverted index makes the lower pixel index win among exactly equal scores
// (the reference's order there is np.argsort-unstable, nets/extractor.py:176,323).  Keys are unique,
// so a candidate's position in the descending order is simply the number of keys greater than it:
// every thread owns one candidate, streams all keys through shared memory, counts, and - if its
// rank is below K - writes its (x, y), score straight to row `rank`.  No sort, no single-CTA tail;
// O(n^2) compares spread over the whole GPU (n ~ 6k: 39 M compares).
constexpr int SEL_THREADS = 256, SEL_CHUNK = 2048;

__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(const unsigned long long* __restrict__ cand, int cap, const int* __restrict__ counter, int W, int topk,
              float* __restrict__ kpts, float* __restrict__ scores, int32_t* __restrict__ count_out,
              int* __restrict__ status) {
  __shared__ unsigned long long keys[SEL_CHUNK];
  int n = *counter;
  if (n > cap) {                       // more candidates than the workspace holds: report, keep what fits
    if (blockIdx.x == 0 && threadIdx.x == 0) *status = SFD2_ERR_OVERFLOW;
    n = cap;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *count_out = (topk > 0 && topk < n) ? topk : n;
  if (blockIdx.x * SEL_THREADS >= n) return;
  const int i = blockIdx.x * SEL_THREADS + threadIdx.x;
  const unsigned long long mine = (i < n) ? cand[i] : 0xFFFFFFFFFFFFFFFFull;
  int rank = 0;
  for (int c0 = 0; c0 < n; c0 += SEL_CHUNK) {
    const int cn = min(SEL_CHUNK, n - c0);
    __syncthreads();
    for (int j = threadIdx.x; j < cn; j += SEL_THREADS) keys[j] = cand[c0 + j];
    __syncthreads();
    int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    int j = 0;
    for (; j + 4 <= cn; j += 4) {
      r0 += keys[j] > mine; r1 += keys[j + 1] > mine; r2 += keys[j + 2] > mine; r3 += keys[j + 3] > mine;
    }
    for (; j < cn; ++j) r0 += keys[j] > mine;
    rank += r0 + r1 + r2 + r3;
  }
  const int k = (topk > 0 && topk < n) ? topk : n;
  if (i < n && rank < k) {
    const unsigned lin = 0xFFFFFFFFu - (unsigned)(mine & 0xFFFFFFFFull);
    kpts[2 * rank] = (float)(lin % (unsigned)W);
    kpts[2 * rank + 1] = (float)(lin / (unsigned)W);
    scores[rank] = __uint_as_float((unsigned)(mine >> 32));
  }
}

int launch_select(unsigned long long* cand, int cap, const int* counter, int W, int topk, float* kpts, float* scores,
                  int32_t* count_out, int* status, unsigned long long* scratch, cudaStream_t st) {
  (void)scratch;
  select_kernel<<<cdiv(cap, SEL_THREADS), SEL_THREADS, 0, st>>>(cand, cap, counter, W, topk, kpts, scores, count_out, status);
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
