// Post-processing kernels: heat-map assembly, 3-round radius-4 NMS with fused candidate
// compaction, top-K selection, bilinear descriptor sampling.  All are HBM/latency-bound
// compare/gather work on CUDA cores (no tensor-core shape to exploit).
#include <math_constants.h>

#include "common.cuh"
#include "ptx.cuh"

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
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();                      // (launch_pdl) the head's maps are complete from here on
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
  SFD2_CUDA(launch_pdl(heat_kernel, grid, block, 0, st, semi, H8, W8, sta, H4, W4, use_sta, heat, H, W));
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

// ------------------------------------------------------------------------------ NMS (radius 4, 3 rounds)
// Literal restatement of simple_nms (nets/extractor.py:20-35) on one smem tile with a 20-pixel
// halo: each of the 5 max-pools / dilations reaches 4 px and the dependency chain is 4+4+4+4+4.
// Out-of-image pixels hold -inf, which is max_pool2d's implicit padding value, and are never
// allowed into the mask.  Every stage is evaluated only where the later stages still need it: the
// interior grown by a margin of 16 (first pool), 12, 8, 4 and 0 pixels, so no window ever leaves the
// tile and the total pooled area is 1.9x smaller than five full-tile passes.
//
// Each 9-wide running max is computed in registers for 8 outputs at a time from 16 loaded values
// with the doubling scheme m2 -> m4 -> m8 -> m9 (45 max ops and 16 smem reads per 8 outputs).
// Row passes map consecutive lanes to consecutive ROWS and the pitch is odd, column passes map lanes to
// consecutive columns, so every shared-memory access is conflict-free (the first version's row passes
// walked runs of 8 floats per lane = an 8-way bank conflict on each of their 24 accesses).
// supp_scores (= 0 where suppressed) is never materialised: it is re-derived from s and supp where read,
// which keeps the block at 83 KB of shared memory = 2 blocks per SM.
constexpr int NT_W = 64, NT_H = 32, NHALO = 20, NR = 4;
constexpr int NS_W = NT_W + 2 * NHALO;  // 104
constexpr int NS_H = NT_H + 2 * NHALO;  // 72
constexpr int NP_W = NS_W + 1;          // odd pitch
constexpr int NP_N = NP_W * NS_H;
constexpr int NMS_THREADS = 512;
__device__ __forceinline__ int nidx(int y, int x) { return y * NP_W + x; }  // tile coords -> smem index

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

// horizontal 9-max of `load(idx)` into tmp, for the rows / columns a following column pass with output margin
// `mg` needs: rows [NHALO-mg-4, NHALO+NT_H+mg+4), columns [NHALO-mg, NHALO+NT_W+mg).  Work item = (run of 8
// columns, row); consecutive threads take consecutive rows.
template <typename T, typename Load>
__device__ __forceinline__ void row_pass(Load load, T* __restrict__ tmp, int mg) {
  const int ylo = NHALO - mg - NR, nrows = NT_H + 2 * mg + 2 * NR;
  const int xlo = NHALO - mg, nruns = (NT_W + 2 * mg) / 8;
  for (int it = threadIdx.x; it < nrows * nruns; it += NMS_THREADS) {
    const int k = it / nrows, y = ylo + (it - k * nrows), x0 = xlo + 8 * k;
    T v[16], o[8];
    const int p = nidx(y, x0 - NR);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = load(p + i);
    run9(v, o);
    T* q = tmp + nidx(y, x0);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = o[i];
  }
}

// vertical 9-max of tmp over the interior grown by `mg`; every result goes to consume(idx, value).
// Work item = (run of 8 rows, column); consecutive threads take consecutive columns.
template <typename T, typename Consume>
__device__ __forceinline__ void col_pass(const T* __restrict__ tmp, Consume consume, int mg) {
  const int ylo = NHALO - mg, nruns = (NT_H + 2 * mg) / 8;
  const int xlo = NHALO - mg, ncols = NT_W + 2 * mg;
  for (int it = threadIdx.x; it < nruns * ncols; it += NMS_THREADS) {
    const int yr = it / ncols, x = xlo + (it - yr * ncols), y0 = ylo + 8 * yr;
    T v[16], o[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = tmp[nidx(y0 - NR + i, x)];
    run9(v, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) consume(nidx(y0 + i, x), o[i]);
  }
}

__global__ void __launch_bounds__(NMS_THREADS, 2)
nms_kernel(const float* __restrict__ heat, int H, int W, float conf_th, int border, int bw, int bh,
           float* __restrict__ nms_out,
           unsigned long long* __restrict__ cand, int cap, int* __restrict__ counter) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  extern __shared__ float sm[];
  float* s = sm;             // scores (-inf outside the image)
  float* u = sm + NP_N;      // row-pass scratch of the float pools
  unsigned char* m = reinterpret_cast<unsigned char*>(sm + 2 * NP_N);  // max_mask
  unsigned char* tb = m + NP_N;                                        // row-pass scratch of the dilations
  unsigned char* supp = tb + NP_N;                                     // supp_mask
  const int X0 = blockIdx.x * NT_W - NHALO, Y0 = blockIdx.y * NT_H - NHALO;
  for (int i = threadIdx.x; i < NS_W * NS_H; i += NMS_THREADS) {
    const int ty = i / NS_W, tx = i - ty * NS_W;
    const int y = Y0 + ty, x = X0 + tx;
    s[nidx(ty, tx)] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(heat + (size_t)y * W + x) : -CUDART_INF_F;
  }
  __syncthreads();
  // max_mask = scores == max_pool(scores), on the interior + 16
  row_pass<float>([&](int id) { return s[id]; }, u, 16);
  __syncthreads();
  col_pass<float>(u, [&](int id, float o) { const float sv = s[id]; m[id] = (sv != -CUDART_INF_F) && (sv == o); }, 16);
  __syncthreads();
#pragma unroll 1
  for (int round = 0; round < 2; ++round) {
    const int mg = round == 0 ? 12 : 4;
    // supp_mask = max_pool(max_mask) > 0, on the interior + mg
    row_pass<unsigned char>([&](int id) { return m[id]; }, tb, mg);
    __syncthreads();
    col_pass<unsigned char>(tb, [&](int id, unsigned char o) { supp[id] = o; }, mg);
    __syncthreads();
    // new_max_mask = supp_scores == max_pool(supp_scores), supp_scores = supp ? 0 : scores; on the interior + mg - 4
    auto supp_score = [&](int id) { const float sv = s[id]; return (supp[id] && sv != -CUDART_INF_F) ? 0.f : sv; };
    row_pass<float>(supp_score, u, mg - 4);
    __syncthreads();
    col_pass<float>(u, [&](int id, float o) {
      const float sv = s[id];
      const unsigned char sp = supp[id];
      const float tv = (sp && sv != -CUDART_INF_F) ? 0.f : sv;
      if (sv != -CUDART_INF_F && tv == o && !sp) m[id] = 1;           // max_mask |= new_max & ~supp
    }, mg - 4);
    __syncthreads();
  }
  // output + candidate compaction for the tile interior
  for (int i = threadIdx.x; i < NT_W * NT_H; i += NMS_THREADS) {
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
// zero_counter = false: the caller has zeroed `counter` earlier on the stream (so that no memset sits between heat_kernel and this launch)
int launch_nms(const float* heat, int H, int W, float conf_th, int border, int bw, int bh, float* nms_out, unsigned long long* cand,
               int cap, int* counter, cudaStream_t st, bool zero_counter) {
  const size_t smem = (size_t)2 * NP_N * sizeof(float) + 3 * NP_N;
  SFD2_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: set on every launch (cheap)
  if (zero_counter) SFD2_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
  dim3 grid(cdiv(W, NT_W), cdiv(H, NT_H));
  SFD2_CUDA(launch_pdl(nms_kernel, grid, dim3(NMS_THREADS), smem, st, heat, H, W, conf_th, border, bw, bh, nms_out, cand, cap, counter));
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
constexpr int SEL_THREADS = 256, SEL_CHUNK = 1024, SEL_Y = 8;

// grid = (candidate chunks of 256, SEL_Y key ranges): block (bx, by) counts, for its 256 candidates, the keys
// greater than each among the key chunks by, by + SEL_Y, ... and adds the partial ranks into `rank_buf`; the
// last block to arrive for a candidate chunk (per-chunk arrival counter) reads the complete ranks and writes the
// output rows.  The n^2 compares thus spread over ~n/256 x min(SEL_Y, n/1024) CTAs instead of n/256.
// rank_buf [cap] and arrive [gridDim.x] are zero on entry and are left zero (the finishing block resets them).
__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(const unsigned long long* __restrict__ cand, int cap, const int* __restrict__ counter, int W, int topk,
              float* __restrict__ kpts, float* __restrict__ scores, int32_t* __restrict__ count_out,
              int* __restrict__ status, int* __restrict__ rank_buf, int* __restrict__ arrive) {
  __shared__ unsigned long long keys[SEL_CHUNK];
  __shared__ int is_last;
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  int n = *counter;
  if (n > cap) {                       // more candidates than the workspace holds: report, keep what fits
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *status = SFD2_ERR_OVERFLOW;
    n = cap;
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *count_out = (topk > 0 && topk < n) ? topk : n;
  // rows beyond the count read zero (sample_kernel does the same for the descriptors): callers need no memset
  if (blockIdx.y == 0)
    for (int r = ((topk > 0 && topk < n) ? topk : n) + blockIdx.x * SEL_THREADS + threadIdx.x; r < topk; r += gridDim.x * SEL_THREADS) {
      kpts[2 * r] = 0.f; kpts[2 * r + 1] = 0.f; scores[r] = 0.f;
    }
  const int nchunks = (n + SEL_CHUNK - 1) / SEL_CHUNK;
  const int active_y = min(nchunks, (int)gridDim.y);
  if (blockIdx.x * SEL_THREADS >= n || (int)blockIdx.y >= active_y) return;
  const int i = blockIdx.x * SEL_THREADS + threadIdx.x;
  const unsigned long long mine = (i < n) ? cand[i] : 0xFFFFFFFFFFFFFFFFull;
  int rank = 0;
  for (int c0 = blockIdx.y * SEL_CHUNK; c0 < n; c0 += gridDim.y * SEL_CHUNK) {
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
  if (active_y > 1) {
    if (i < n && rank) atomicAdd(&rank_buf[i], rank);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&arrive[blockIdx.x], 1) == active_y - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    rank = (i < n) ? atomicExch(&rank_buf[i], 0) : 0;      // complete rank; leave the buffer zero for the next call
    if (threadIdx.x == 0) arrive[blockIdx.x] = 0;
  }
  const int k = (topk > 0 && topk < n) ? topk : n;
  if (i < n && rank < k) {
    const unsigned lin = 0xFFFFFFFFu - (unsigned)(mine & 0xFFFFFFFFull);
    kpts[2 * rank] = (float)(lin % (unsigned)W);
    kpts[2 * rank + 1] = (float)(lin / (unsigned)W);
    scores[rank] = __uint_as_float((unsigned)(mine >> 32));
  }
}

// scratch: >= (cap + cdiv(cap, 256)) ints, zero-initialised once by the owner (the kernel leaves it zero)
int launch_select(unsigned long long* cand, int cap, const int* counter, int W, int topk, float* kpts, float* scores,
                  int32_t* count_out, int* status, unsigned long long* scratch, cudaStream_t st) {
  int* rank_buf = reinterpret_cast<int*>(scratch);
  int* arrive = rank_buf + cap;
  dim3 grid(cdiv(cap, SEL_THREADS), SEL_Y);
  SFD2_CUDA(launch_pdl(select_kernel, grid, dim3(SEL_THREADS), 0, st, cand, cap, counter, W, topk, kpts, scores, count_out, status, rank_buf, arrive));
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
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
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
  SFD2_CUDA(launch_pdl(sample_kernel, dim3(cdiv(topk * 32, 256)), dim3(256), 0, st, desc_map, H4, W4, H, W, kpts, count, topk, desc_out));
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

}  // namespace sfd2
