// Mutual-nearest-neighbour descriptor matcher (hloc/matchers/nearest_neighbor.py:38-57,
// it_loc/matcher.py:122-130).  The N x M similarity matrix is never written to HBM: every
// tile reduces straight into per-row and per-column 64-bit keys
//     key = orderable(sim) << 32 | (0xFFFFFFFF - index)
// combined with atomicMax, so the largest similarity wins and, among exactly equal values, the
// LOWEST index (the reference's own tie order differs between topk and max).
//
// This file holds the CUDA-core fp32 version (SFD2_PREC_FP32) and the finishing kernel that
// decodes keys, applies the optional distance threshold and the mutual check.  The tcgen05
// version lives in tc_match.cu and produces the same keys.
#include "common.cuh"

namespace sfd2 {

__device__ __forceinline__ unsigned ord_f32(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
__device__ __forceinline__ unsigned long long make_key(float sim, int idx) {
  return ((unsigned long long)ord_f32(sim) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
}

constexpr int MT = 64, MK = 32;

__global__ void __launch_bounds__(256)
match_simt_kernel(const float* __restrict__ d0, int n0, const float* __restrict__ d1, int n1, int d,
                  unsigned long long* __restrict__ row_key, unsigned long long* __restrict__ col_key) {
  __shared__ float As[MK][MT + 4];
  __shared__ float Bs[MK][MT + 4];
  __shared__ unsigned long long colred[16][MT];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * MT, j0 = blockIdx.x * MT;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int k0 = 0; k0 < d; k0 += MK) {
    // each thread loads 8 floats of A and of B: row = tid/4, cols (tid%4)*8
    const int r = tid >> 2, c = (tid & 3) * 8;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k = k0 + c + q;
      As[c + q][r] = (i0 + r < n0 && k < d) ? __ldg(d0 + (size_t)(i0 + r) * d + k) : 0.f;
      Bs[c + q][r] = (j0 + r < n1 && k < d) ? __ldg(d1 + (size_t)(j0 + r) * d + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = As[k][ty * 4 + q]; b[q] = Bs[k][tx * 4 + q]; }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
    }
  }
  // row keys: reduce over this thread's 4 columns, then across the 16 threads sharing ty
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    unsigned long long best = 0ull;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + tx * 4 + q;
      if (j < n1) { const unsigned long long k = make_key(acc[p][q], j); best = (k > best) ? k : best; }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = (other > best) ? other : best;
    }
    const int i = i0 + ty * 4 + p;
    if (tx == 0 && i < n0 && best) atomicMax(row_key + i, best);
  }
  // column keys: per-thread over 4 rows, then across the 16 ty values through smem
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    unsigned long long best = 0ull;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int i = i0 + ty * 4 + p;
      if (i < n0) { const unsigned long long k = make_key(acc[p][q], i); best = (k > best) ? k : best; }
    }
    colred[ty][tx * 4 + q] = best;
  }
  __syncthreads();
  if (tid < MT) {
    unsigned long long best = 0ull;
#pragma unroll
    for (int y = 0; y < 16; ++y) { const unsigned long long k = colred[y][tid]; best = (k > best) ? k : best; }
    const int j = j0 + tid;
    if (j < n1 && best) atomicMax(col_key + j, best);
  }
}

// Second pass for the ratio tests (nearest_neighbor.py:7,10-11; it_loc/matcher.py:165-194): the best
// similarity of every row / column EXCLUDING its arg-max, i.e. topk(2)[1].  Same tiling as above; the
// excluded index comes from the first pass's keys.  Values are reduced as order-preserving uint32.
__global__ void __launch_bounds__(256)
match_second_kernel(const float* __restrict__ d0, int n0, const float* __restrict__ d1, int n1, int d,
                    const unsigned long long* __restrict__ row_key, const unsigned long long* __restrict__ col_key,
                    unsigned* __restrict__ row2, unsigned* __restrict__ col2) {
  __shared__ float As[MK][MT + 4];
  __shared__ float Bs[MK][MT + 4];
  __shared__ unsigned colred[16][MT];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * MT, j0 = blockIdx.x * MT;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int k0 = 0; k0 < d; k0 += MK) {
    const int r = tid >> 2, c = (tid & 3) * 8;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k = k0 + c + q;
      As[c + q][r] = (i0 + r < n0 && k < d) ? __ldg(d0 + (size_t)(i0 + r) * d + k) : 0.f;
      Bs[c + q][r] = (j0 + r < n1 && k < d) ? __ldg(d1 + (size_t)(j0 + r) * d + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = As[k][ty * 4 + q]; b[q] = Bs[k][tx * 4 + q]; }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
    }
  }
  int excl_j[4], excl_i[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int i = i0 + ty * 4 + p;
    excl_j[p] = (i < n0) ? (int)(0xFFFFFFFFu - (unsigned)(row_key[i] & 0xFFFFFFFFull)) : -1;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = j0 + tx * 4 + q;
    excl_i[q] = (j < n1) ? (int)(0xFFFFFFFFu - (unsigned)(col_key[j] & 0xFFFFFFFFull)) : -1;
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    unsigned best = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + tx * 4 + q;
      if (j < n1 && j != excl_j[p]) best = max(best, ord_f32(acc[p][q]));
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    const int i = i0 + ty * 4 + p;
    if (tx == 0 && i < n0 && best) atomicMax(row2 + i, best);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    unsigned best = 0u;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int i = i0 + ty * 4 + p;
      if (i < n0 && i != excl_i[q]) best = max(best, ord_f32(acc[p][q]));
    }
    colred[ty][tx * 4 + q] = best;
  }
  __syncthreads();
  if (tid < MT) {
    unsigned best = 0u;
#pragma unroll
    for (int y = 0; y < 16; ++y) best = max(best, colred[y][tid]);
    const int j = j0 + tid;
    if (j < n1 && best) atomicMax(col2 + j, best);
  }
}

int launch_match_second(const float* d0, int n0, const float* d1, int n1, int d, const unsigned long long* row_key,
                        const unsigned long long* col_key, unsigned* row2, unsigned* col2, cudaStream_t st) {
  SFD2_CUDA(cudaMemsetAsync(row2, 0, sizeof(unsigned) * (size_t)(n0 > 0 ? n0 : 1), st));
  SFD2_CUDA(cudaMemsetAsync(col2, 0, sizeof(unsigned) * (size_t)(n1 > 0 ? n1 : 1), st));
  if (n0 <= 0 || n1 <= 0) return SFD2_OK;
  dim3 grid(cdiv(n1, MT), cdiv(n0, MT));
  match_second_kernel<<<grid, 256, 0, st>>>(d0, n0, d1, n1, d, row_key, col_key, row2, col2);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

int launch_match_simt(const float* d0, int n0, const float* d1, int n1, int d, unsigned long long* row_key,
                      unsigned long long* col_key, cudaStream_t st) {
  SFD2_CUDA(cudaMemsetAsync(row_key, 0, sizeof(unsigned long long) * (size_t)(n0 > 0 ? n0 : 1), st));
  SFD2_CUDA(cudaMemsetAsync(col_key, 0, sizeof(unsigned long long) * (size_t)(n1 > 0 ? n1 : 1), st));
  if (n0 <= 0 || n1 <= 0) return SFD2_OK;
  dim3 grid(cdiv(n1, MT), cdiv(n0, MT));
  match_simt_kernel<<<grid, 256, 0, st>>>(d0, n0, d1, n1, d, row_key, col_key);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

// Lowe ratio test on (best, second-best) similarity.  mode 1 = hloc find_nn (nearest_neighbor.py:8-11):
// 2(1-s0) <= r^2 * 2(1-s1); mode 2 = it_loc (matcher.py:172-174): sqrt(2-2 s0) / (sqrt(2-2 s1) + 1e-8) <= r.
// A missing second neighbour (only one candidate) passes.
__device__ __forceinline__ bool ratio_ok(float s0, unsigned second_ord, float r, int mode) {
  if (second_ord == 0u) return true;
  const float s1 = unord_f32(second_ord);
  if (mode == 2) return sqrtf(2.f - 2.f * s0) / (sqrtf(2.f - 2.f * s1) + 1e-8f) <= r;
  return 2.f * (1.f - s0) <= (r * r) * (2.f * (1.f - s1));
}

// matches0[i] = nn12[i] if (ratio ok) and (dist ok) and (no mutual check or nn21[nn12[i]] == i with the column
// passing the same tests) else -1 ; sim0[i] = max_j sim[i][j]
__global__ void match_finish_kernel(const unsigned long long* __restrict__ row_key,
                                    const unsigned long long* __restrict__ col_key, int n0, int n1, int mutual,
                                    float dist_th, float ratio_th, int ratio_mode, const unsigned* __restrict__ row2,
                                    const unsigned* __restrict__ col2, int32_t* __restrict__ matches0,
                                    float* __restrict__ sim0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0) return;
  const unsigned long long k = row_key[i];
  if (k == 0ull) { matches0[i] = -1; sim0[i] = 0.f; return; }
  const int j = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
  const float s = unord_f32((unsigned)(k >> 32));
  bool ok = true;
  const bool plain = (ratio_mode & SFD2_MATCH_PLAIN_CODES) != 0;   // every unmatched row reads -1
  ratio_mode &= 0xFF;
  if (ratio_th > 0.f) ok = ratio_ok(s, row2[i], ratio_th, ratio_mode);
  if (ok && dist_th > 0.f) ok = (2.f * (1.f - s)) <= dist_th * dist_th;      // nearest_neighbor.py:8,12-13
  const bool row_ok = ok;   // find_nn's own mask for this row (decides whether hloc keeps its score)
  if (ok && mutual) {
    const unsigned long long kc = col_key[j];
    const int i2 = (int)(0xFFFFFFFFu - (unsigned)(kc & 0xFFFFFFFFull));
    ok = (kc != 0ull) && (i2 == i);                                   // mutual_check, :19-24
    if (ok && ratio_th > 0.f) ok = ratio_ok(unord_f32((unsigned)(kc >> 32)), col2[j], ratio_th, ratio_mode);
    if (ok && dist_th > 0.f) ok = (2.f * (1.f - unord_f32((unsigned)(kc >> 32)))) <= dist_th * dist_th;
  }
  matches0[i] = ok ? j : ((row_ok && !plain) ? -2 : -1);   // -1: rejected by the row's own tests, -2: by the mutual check
  sim0[i] = s;
}

int launch_match_finish(const unsigned long long* row_key, const unsigned long long* col_key, int n0, int n1,
                        int mutual, float dist_th, float ratio_th, int ratio_mode, const unsigned* row2,
                        const unsigned* col2, int32_t* matches0, float* sim0, cudaStream_t st) {
  if (n0 <= 0) return SFD2_OK;
  match_finish_kernel<<<cdiv(n0, 256), 256, 0, st>>>(row_key, col_key, n0, n1, mutual, dist_th, ratio_th, ratio_mode,
                                                     row2, col2, matches0, sim0);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

}  // namespace sfd2
