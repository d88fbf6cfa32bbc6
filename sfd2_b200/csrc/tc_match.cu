// Mutual-NN matcher on tcgen05 (hloc/matchers/nearest_neighbor.py:6-57, it_loc/matcher.py:122-194).
//
// sim = D0 * D1^T is ONE K-major x K-major GEMM (K = 128) per pair; the epilogue reduces every 128 x 128 accumulator
// tile straight into 64-bit arg-max keys, so the similarity matrix never exists in HBM:
//   * rows    (TMEM lane = row): thread-local running (best, index [, second best]) along the CTA's strip of tiles,
//             merged into keys[] with one atomicMax per row per strip;
//   * columns (TMEM column = column): a THRESHOLD FILTER - every thread compares its 64 values with the column's current
//             best (keys[] as of the tile's start, staged in shared memory): 1 compare + 1 vote per column.  Only columns
//             where some row of the warp reaches the threshold take the slow path: warp arg-max (redux.sync + ballot,
//             lowest row wins ties) and ONE atomicMax.  After a column has seen a few hundred rows that is a handful of
//             columns per tile.  (Round 1 ran the transposed product D1 * D0^T as a second GEMM to make the column
//             reduction thread-local: twice the tensor work, which capped the kernel at 1/6 of the pipe in exact mode.)
//   * ratio tests need the SECOND best of every row and column (topk(2), nearest_neighbor.py:7): those configurations run
//             both products with the thread-local (best, second) reduction on each - still tensor cores, no CUDA-core pass.
//   * the mutual check / thresholds / index remap run in the tail of the CTA that completes a pair's last tile.
//
// Work unit = TWO row-blocks (256 rows of the A operand, resident in shared memory) x one 128-row B tile: every B tile
// fetched from L2 feeds two accumulators, which halves the L2 -> SM traffic per MMA.  With one row-block per unit the
// kernel was bound by exactly that traffic (64 KB of B per 128 x 128 x 128 tile in exact mode = 5.8 TB/s chip-wide).
//
// One launch serves MANY pairs: a table of descriptor sets (operands) and of (set a, set b) problems; tiles of all
// problems form one list cut into contiguous per-CTA ranges.  Set sizes may live in DEVICE memory (the extractor's counts),
// so an extract -> match pipeline never synchronises with the host.
//
// Operands are the fp16 hi/lo split of the fp32 descriptors (unit-norm rows: hi + lo carries ~22 bits), written by
// match_prep_kernel from either layout ([n,128] rows or hloc's [128,n]); split==3 issues a_hi*b_hi + a_hi*b_lo + a_lo*b_hi
// into one fp32 accumulator, which reproduces the fp32 reference arg-max on real SFD2 descriptors; split==1 is single-pass.
#include <math_constants.h>

#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tc_match.cuh"

namespace sfd2 {

using namespace ptx;

__device__ __forceinline__ unsigned m_ord_f32(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float m_unord_f32(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
__device__ __forceinline__ unsigned long long m_key(float sim, int idx) {
  return ((unsigned long long)m_ord_f32(sim) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
}
__device__ __forceinline__ int m_key_idx(unsigned long long k) { return (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull)); }
__device__ __forceinline__ float m_key_sim(unsigned long long k) { return m_unord_f32((unsigned)(k >> 32)); }

// fire-and-forget 64-bit max (REDG): atomicMax() compiles to ATOMG even when the result is unused, and the warp then
// waits out the L2 round trip of every key update (~2.9 us per tile in the column path)
__device__ __forceinline__ void red_max_u64(unsigned long long* addr, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void red_max_u32(unsigned* addr, unsigned v) {
  asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}


// ------------------------------------------------------------------------------------------------ prep
// ids -> order-preserving compaction table (desc_db[db_3D_ids != -1], it_loc/localize_cv2.py:540-555): one block per
// operand with ids; remap[prow0 + k] = k-th row whose id != -1, efflen = number of such rows.
__global__ void __launch_bounds__(1024)
match_scan_ids_kernel(const MOperD* __restrict__ opers_dev, const __grid_constant__ MTabInline inl, int noper,
                      int* __restrict__ efflen, int* __restrict__ remap) {
  __shared__ int warp_sums[32];
  __shared__ int base;
  for (int o = blockIdx.x; o < noper; o += gridDim.x) {
    const MOperD op = opers_dev ? opers_dev[o] : (o ? inl.opers[1] : inl.opers[0]);
    if (!op.ids) continue;
    int n = op.cap;
    if (op.count) n = min(n, max(0, *op.count));
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int r0 = 0; r0 < n; r0 += blockDim.x) {
      const int r = r0 + threadIdx.x;
      const int valid = (r < n && op.ids[r] != -1) ? 1 : 0;
      const unsigned b = __ballot_sync(0xffffffffu, valid);
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      if (lane == 0) warp_sums[w] = __popc(b);
      __syncthreads();
      int woff = 0;
      for (int k = 0; k < w; ++k) woff += warp_sums[k];
      const int pos = base + woff + __popc(b & ((1u << lane) - 1u));
      if (valid) remap[op.prow0 + pos] = r;
      __syncthreads();
      if (threadIdx.x == blockDim.x - 1) base = pos + valid;
      __syncthreads();
    }
    if (threadIdx.x == 0) efflen[o] = base;
    __syncthreads();
  }
}

// fp32 descriptors -> fp16 hi / lo plane rows [prow][128] (rows >= the set's effective length are zero up to the next
// multiple of 128), key / second-best / done-counter reset, effective lengths.  Block = 256 threads = 32 plane rows.
__global__ void __launch_bounds__(256)
match_prep_kernel(const MOperD* __restrict__ opers_dev, const __grid_constant__ MTabInline inl, int noper, int total_prows,
                  __half* __restrict__ hi, __half* __restrict__ lo, const int* __restrict__ remap, int* __restrict__ efflen,
                  unsigned long long* __restrict__ keys, unsigned* __restrict__ sec, long long nkeys,
                  int* __restrict__ done, int nprob) {
  __shared__ float tile[128][33];
  pdl_launch_dependents();   // the matcher kernel may start its prologue; it waits (griddepcontrol.wait) before reading
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  for (long long i = gtid; i < nkeys; i += gsz) { keys[i] = 0ull; if (sec) sec[i] = 0u; }
  for (int i = gtid; i < nprob; i += gsz) done[i] = 0;
  for (int g = blockIdx.x; g < total_prows / 32; g += gridDim.x) {
    const int R0 = g * 32;
    int a = 0;
    MOperD op;
    if (opers_dev) {
      int b = noper - 1;
      while (a < b) { const int mid = (a + b + 1) >> 1; if (__ldg(&opers_dev[mid].prow0) <= R0) a = mid; else b = mid - 1; }
      op = opers_dev[a];
    } else {               // single pair: tables in the kernel parameters, static indices only (constant bank)
      a = (noper > 1 && R0 >= inl.opers[1].prow0) ? 1 : 0;
      op = a ? inl.opers[1] : inl.opers[0];
    }
    int n = op.cap;
    if (op.ids) n = efflen[a];                            // written by match_scan_ids_kernel (earlier launch)
    else if (op.count) n = min(n, max(0, *op.count));
    if (!op.ids && R0 == op.prow0 && threadIdx.x == 0) efflen[a] = n;
    const int l0 = R0 - op.prow0;                         // first local row of this group
    if (l0 >= ((n + 127) & ~127)) continue;               // beyond the last tile any kernel will touch
    if (op.layout == 0) {
      // [n][128] rows: thread -> (row = tid / 32, 4 consecutive k), 8 rows per sweep
      for (int rr = threadIdx.x >> 5; rr < 32; rr += 8) {
        const int l = l0 + rr, kq = threadIdx.x & 31;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l < n) {
          const int sr = op.ids ? remap[op.prow0 + l] : l;
          v = __ldg(reinterpret_cast<const float4*>(op.src + (size_t)sr * op.rs) + kq);
        }
        const float x[4] = {v.x, v.y, v.z, v.w};
        __align__(8) __half h[4];
        __align__(8) __half q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { h[j] = __float2half_rn(x[j]); q[j] = __float2half_rn(x[j] - __half2float(h[j])); }
        const size_t o = ((size_t)(R0 + rr) * 128 + kq * 4) / 4;
        reinterpret_cast<uint2*>(hi)[o] = *reinterpret_cast<uint2*>(h);
        reinterpret_cast<uint2*>(lo)[o] = *reinterpret_cast<uint2*>(q);
      }
    } else {
      // hloc layout [128][n] (element (row r, k) at src[k * cs + r * rs], rs == 1): coalesced along r, transposed in smem
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      const int l = l0 + lane;
      const bool ok = l < n;
      const int sr = ok ? (op.ids ? remap[op.prow0 + l] : l) : 0;
      for (int k = w; k < 128; k += 8) tile[k][lane] = ok ? __ldg(op.src + (size_t)k * op.cs + (size_t)sr * op.rs) : 0.f;
      __syncthreads();
      const int rr = threadIdx.x >> 3, k0 = (threadIdx.x & 7) * 16;
      __align__(16) __half h[16];
      __align__(16) __half q[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x = tile[k0 + j][rr];
        h[j] = __float2half_rn(x);
        q[j] = __float2half_rn(x - __half2float(h[j]));
      }
      uint4* oh = reinterpret_cast<uint4*>(hi + (size_t)(R0 + rr) * 128 + k0);
      uint4* ol = reinterpret_cast<uint4*>(lo + (size_t)(R0 + rr) * 128 + k0);
      oh[0] = reinterpret_cast<const uint4*>(h)[0]; oh[1] = reinterpret_cast<const uint4*>(h)[1];
      ol[0] = reinterpret_cast<const uint4*>(q)[0]; ol[1] = reinterpret_cast<const uint4*>(q)[1];
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------ main kernel
constexpr int TM_TILE = 128;
constexpr int TM_OP_BYTES = 128 * 128;   // 128 rows x 64 fp16 (one K half of one plane)
constexpr int TM_THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int TM_EPI_THREADS = 256;
// row-blocks per work unit: SUB = 2 halves the L2 -> SM traffic per MMA (rows-only and two-product modes, which are bound
// by it); the mutual mode is bound by its column epilogue and runs SUB = 1 (more accumulator buffers, finer work units).
// Accumulator buffers = 512 TMEM columns / (SUB x 128).
constexpr int TM_MAX_STAGES = 10;
constexpr int TM_MAX_ACC_BUFS = 4;

struct TileInfo {
  int p, pass, mt, nt;
  int a_len, b_len;            // effective rows of the A / B operand of this pass
  int a_prow, b_prow;          // first plane row of the A / B operand
  long long ka, kb;            // key offsets of the A rows / B rows
  int rb;                      // row-block identity within the launch (changes <=> the resident A operand changes)
  bool skip;
};

// Everything tm_locate needs about one problem, fetched ONCE per problem per thread: the tables live in global memory
// or in the kernel's parameter space, and a handful of dependent loads per tile (through a generic pointer into the
// parameter window they cost ~1 us each) made the single-pair call take 58 us instead of 33.
struct ProbCache {
  int p = -1;
  int tile0 = 0, ntiles = 0, tm = 0, tn = 0;
  int len_a = 0, len_b = 0, prow_a = 0, prow_b = 0;
  long long key_a = 0, key_b = 0;
};

// Linear tile -> (problem, pass, mt, nt).  Problems are laid out by their CAPACITY tile counts (host-known);
// tiles beyond the effective extents (device-side counts) are skipped identically by all three warp roles.
__device__ __forceinline__ void tm_locate(const TcMatchArgs& a, int tile, ProbCache& c, TileInfo& t) {
  if (c.p < 0 || tile < c.tile0 || tile >= c.tile0 + c.ntiles) {
    MProbD pr;
    MOperD oa, ob;
    int p = 0;
    if (a.probs) {
      int lo = 0, hi = a.nprob - 1;
      while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (__ldg(&a.probs[mid].tile0) <= tile) lo = mid; else hi = mid - 1; }
      p = lo;
      pr = a.probs[p];
      oa = a.opers[pr.a];
      ob = a.opers[pr.b];
    } else {                 // single pair: the tables are kernel parameters (constant bank, static indices)
      pr = a.inl.probs[0];
      oa = a.inl.opers[0];
      ob = a.inl.opers[1];
      if (pr.a == pr.b) ob = oa;
      else if (pr.a == 1) { oa = a.inl.opers[1]; ob = a.inl.opers[0]; }
    }
    c.p = p;
    c.tile0 = pr.tile0; c.ntiles = pr.ntiles; c.tm = pr.tm; c.tn = pr.tn;
    // host-known sizes need no round trip to the prep kernel's output
    c.len_a = (oa.count || oa.ids) ? a.efflen[pr.a] : oa.cap;
    c.len_b = (ob.count || ob.ids) ? a.efflen[pr.b] : ob.cap;
    c.prow_a = oa.prow0; c.prow_b = ob.prow0;
    c.key_a = pr.key_a; c.key_b = pr.key_b;
  }
  int r = tile - c.tile0;
  const int sub = a.sub;
  const int u0 = ((c.tm + sub - 1) / sub) * c.tn;    // units of the first product: super-blocks of a x column tiles of b
  t.p = c.p;
  t.pass = (r >= u0) ? 1 : 0;
  if (t.pass) r -= u0;
  const int tn = t.pass ? c.tm : c.tn;
  t.mt = r / tn;                                     // super-block: rows [128 sub mt, 128 sub (mt + 1)) of the A operand
  // strips start at skewed columns: CTAs working on different row-blocks of one problem at the same time then sit on
  // different column tiles, so a column's threshold is established by whoever comes first instead of being cold for all
  t.nt = (r - t.mt * tn + t.mt * 5) % tn;
  t.a_len = t.pass ? c.len_b : c.len_a; t.b_len = t.pass ? c.len_a : c.len_b;
  t.a_prow = t.pass ? c.prow_b : c.prow_a; t.b_prow = t.pass ? c.prow_a : c.prow_b;
  t.ka = t.pass ? c.key_b : c.key_a;
  t.kb = t.pass ? c.key_a : c.key_b;
  t.rb = c.tile0 + (t.pass ? u0 : 0) + t.mt * tn;     // linear index of the strip's first unit: unique per (p, pass, mt)
  t.skip = (t.mt * sub * TM_TILE >= t.a_len) || (t.nt * TM_TILE >= t.b_len);
}

// Lowe ratio test on (best, second-best) similarity.  mode 1 = hloc find_nn (nearest_neighbor.py:8-11):
// 2(1-s0) <= r^2 * 2(1-s1); mode 2 = it_loc (matcher.py:172-174): sqrt(2-2 s0) / (sqrt(2-2 s1) + 1e-8) <= r.
// A missing second neighbour (only one candidate) passes.
__device__ __forceinline__ bool tm_ratio_ok(float s0, unsigned second_ord, float r, int mode) {
  if (second_ord == 0u) return true;
  const float s1 = m_unord_f32(second_ord);
  if (mode == 2) return sqrtf(2.f - 2.f * s0) / (sqrtf(2.f - 2.f * s1) + 1e-8f) <= r;
  return 2.f * (1.f - s0) <= (r * r) * (2.f * (1.f - s1));
}

// Tail of a problem (run by the 256 epilogue threads of the CTA that completed its last tile): decode keys, ratio /
// distance tests, mutual check, index remap -> matches0 / sim0.  Same decisions as match_finish_kernel (match.cu).
__device__ void tm_finish(const TcMatchArgs& a, int p, int et) {
  MProbD pr;
  MOperD oa, ob;
  if (a.probs) { pr = a.probs[p]; oa = a.opers[pr.a]; ob = a.opers[pr.b]; }
  else {
    pr = a.inl.probs[0]; oa = a.inl.opers[0]; ob = a.inl.opers[1];
    if (pr.a == pr.b) ob = oa;
    else if (pr.a == 1) { oa = a.inl.opers[1]; ob = a.inl.opers[0]; }
  }
  const int n0 = a.efflen[pr.a];
  const int cap0 = oa.cap;
  const int* remap_b = ob.ids ? a.remap + ob.prow0 : nullptr;
  // feature_matching (it_loc/localize_cv2.py:537-538): a db image with <= 3 keypoints that have a 3-D point yields no matches
  const bool too_few = remap_b && a.efflen[pr.b] <= 3;
  const unsigned long long* rk = a.keys + pr.key_a;
  const unsigned long long* ck = a.keys + pr.key_b;
  const unsigned* rs = a.sec ? a.sec + pr.key_a : nullptr;
  const unsigned* cs = a.sec ? a.sec + pr.key_b : nullptr;
  const bool plain = (a.ratio_mode & SFD2_MATCH_PLAIN_CODES) != 0;
  const bool hloc_scores = (a.ratio_mode & SFD2_MATCH_HLOC_SCORES) != 0;
  const bool i64 = (a.ratio_mode & SFD2_MATCH_I64) != 0;
  const int rmode = a.ratio_mode & 0xFF;
  const bool ratio = a.ratio_th > 0.f;
  // the tail is latency-bound (row key -> column key of the row's arg-max): R rows per thread per sweep, all loads of
  // one level in flight together (one thread walking its rows one at a time cost 12 us for 4096 rows)
  constexpr int R = 8;
  for (int base = et; base < cap0; base += TM_EPI_THREADS * R) {
    unsigned long long k[R], kc[R];
    unsigned s2r[R], s2c[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int i = base + u * TM_EPI_THREADS;
      k[u] = (i < n0 && !too_few) ? __ldcg(rk + i) : 0ull;
      s2r[u] = (ratio && i < n0) ? __ldcg(rs + i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const bool need = k[u] != 0ull && a.mutual;
      kc[u] = need ? __ldcg(ck + m_key_idx(k[u])) : 0ull;
      s2c[u] = (need && ratio) ? __ldcg(cs + m_key_idx(k[u])) : 0u;
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int i = base + u * TM_EPI_THREADS;
      if (i >= cap0) continue;
      int m = -1;
      float s = 0.f;
      bool keep_score = false;
      if (k[u] != 0ull) {
        const int j = m_key_idx(k[u]);
        s = m_key_sim(k[u]);
        bool ok = true;
        if (ratio) ok = tm_ratio_ok(s, s2r[u], a.ratio_th, rmode);
        if (ok && a.dist_th > 0.f) ok = (2.f * (1.f - s)) <= a.dist_th * a.dist_th;     // nearest_neighbor.py:8,12-13
        const bool row_ok = ok;   // find_nn's own mask for this row (decides whether hloc keeps its score)
        keep_score = row_ok;
        if (ok && a.mutual) {
          ok = (kc[u] != 0ull) && (m_key_idx(kc[u]) == i);                                // mutual_check, :19-24
          if (ok && ratio) ok = tm_ratio_ok(m_key_sim(kc[u]), s2c[u], a.ratio_th, rmode);
          if (ok && a.dist_th > 0.f) ok = (2.f * (1.f - m_key_sim(kc[u]))) <= a.dist_th * a.dist_th;
        }
        // -1: rejected by the row's own tests, -2: only by the mutual check; matches report ORIGINAL db rows (remap)
        m = ok ? (remap_b ? remap_b[j] : j) : ((row_ok && !plain) ? -2 : -1);
      }
      // hloc's find_nn (nearest_neighbor.py:14-15): (sim + 1) / 2 where the row passed its own tests, else 0
      if (hloc_scores) s = keep_score ? (s + 1.f) * 0.5f : 0.f;
      if (i64) reinterpret_cast<long long*>(a.matches0)[pr.out_off + i] = (long long)m;
      else a.matches0[pr.out_off + i] = m;
      a.sim0[pr.out_off + i] = s;
    }
  }
}

// Column reduction of a COLD tile (thresholds not yet established: most of a warp's 64 columns have candidates):
// butterfly transpose-reduce over the warp's 32 rows.  Each step halves the columns a lane is responsible for and
// exchanges the other half with lane ^ half; after five steps lane L holds (max value, lowest row lane attaining it)
// of column L.  31 exchange steps per 32 columns instead of 5 shuffles per column, and no REDUX (a CREDUX per column
// cost ~100 cycles each: 7000 cycles per tile).
__device__ __forceinline__ void tm_butterfly32(unsigned (&v)[32], unsigned (&id)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const unsigned mv = up ? v[j + half] : v[j], mi = up ? id[j + half] : id[j];
      const unsigned ov = up ? v[j] : v[j + half], oi = up ? id[j] : id[j + half];
      const unsigned rv = __shfl_xor_sync(0xffffffffu, ov, half), ri = __shfl_xor_sync(0xffffffffu, oi, half);
      const bool take = (rv > mv) || (rv == mv && ri < mi);
      v[j] = take ? rv : mv;
      id[j] = take ? ri : mi;
    }
  }
}

template <int SUB>
__global__ void __launch_bounds__(TM_THREADS, 1)
tc_match_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                const __grid_constant__ TcMatchArgs a) {
  constexpr int TM_ACC_BUFS = 4 / SUB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nops = (a.split == 3) ? 2 : 1;
  const int a_sub_bytes = 2 * nops * TM_OP_BYTES;                  // one row-block of A: [kb][plane] x 16 KB
  const int a_slot_bytes = SUB * a_sub_bytes;                   // resident A of a unit: [sub][kb][plane]
  uint8_t* aslot = smem;
  uint8_t* bring = smem + (size_t)a.aslots * a_slot_bytes;                        // B ring: stage = one plane of one K half (16 KB)
  uint8_t* tail = bring + (size_t)a.stages * TM_OP_BYTES;
  unsigned long long* thr_key = reinterpret_cast<unsigned long long*>(tail);        // [2][128] column keys at tile start
  float* thr_sim = reinterpret_cast<float*>(tail + 2048);                           // [2][128] their similarities
  uint64_t* full = reinterpret_cast<uint64_t*>(tail + 3072);
  uint64_t* empty = full + TM_MAX_STAGES;
  uint64_t* tfull = empty + TM_MAX_STAGES;
  uint64_t* tempty = tfull + TM_MAX_ACC_BUFS;
  uint64_t* afull = tempty + TM_MAX_ACC_BUFS;
  uint64_t* aempty = afull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty + 2);
  int* last_flag = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { prefetch_tmap(&tm_hi); prefetch_tmap(&tm_lo); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < TM_ACC_BUFS; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_cta = (a.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_begin = (int)blockIdx.x * per_cta;
  const int tile_end = min(tile_begin + per_cta, a.total_tiles);
  pdl_wait();   // everything below reads what match_prep_kernel wrote (planes, effective lengths, zeroed keys)

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (one elected lane)
    if (elect_one()) {
      int stage = 0, prev_rb = -1, as = 0;
      ProbCache pc;
      uint32_t phase = 0, aph = 0u;             // bit s of aph = phase of A slot s
      TileInfo t;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        tm_locate(a, tile, pc, t);
        if (t.skip) continue;
        if (t.rb != prev_rb) {                    // new strip: load its A operand into the other slot
          if (prev_rb >= 0) as = (as + 1 == a.aslots) ? 0 : as + 1;
          mbar_wait(&aempty[as], ((aph >> as) & 1u) ^ 1u);
          aph ^= 1u << as;
          uint8_t* dst = aslot + (size_t)as * a_slot_bytes;
          mbar_expect_tx(&afull[as], (uint32_t)a_slot_bytes);
          for (int sub = 0; sub < SUB; ++sub)
            for (int kb = 0; kb < 2; ++kb) {
              const int row = t.a_prow + (t.mt * SUB + sub) * TM_TILE;     // rows past the set read zeros / a neighbour: masked later
              uint8_t* d = dst + (size_t)sub * a_sub_bytes + (kb * nops) * TM_OP_BYTES;
              tma_load_2d(d, &tm_hi, &afull[as], kb * 64, row);
              if (nops == 2) tma_load_2d(d + TM_OP_BYTES, &tm_lo, &afull[as], kb * 64, row);
            }
          prev_rb = t.rb;
        }
        for (int kb = 0; kb < 2; ++kb)
          for (int pl = 0; pl < nops; ++pl) {     // one stage = one plane of one K half of the B tile (16 KB)
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sb = bring + (size_t)stage * TM_OP_BYTES;
            mbar_expect_tx(&full[stage], (uint32_t)TM_OP_BYTES);
            tma_load_2d(sb, pl ? &tm_lo : &tm_hi, &full[stage], kb * 64, t.b_prow + t.nt * TM_TILE);
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (one elected lane)
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, 128);
      int stage = 0, buf = 0, prev_rb = -1, as = 0;
      ProbCache pc, pc2;
      uint32_t phase = 0, bphase = 0, aph = 0u;
      TileInfo t, t2;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        tm_locate(a, tile, pc, t);
        if (t.skip) continue;
        if (t.rb != prev_rb) {
          if (prev_rb >= 0) as = (as + 1 == a.aslots) ? 0 : as + 1;
          mbar_wait(&afull[as], (aph >> as) & 1u);
          aph ^= 1u << as;
          prev_rb = t.rb;
        }
        mbar_wait(&tempty[buf], bphase ^ 1);
        tc_fence_after();
        const uint32_t abase = smem_u32(aslot + (size_t)as * a_slot_bytes);
        for (int kb = 0; kb < 2; ++kb)
          for (int pl = 0; pl < nops; ++pl) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t db = make_desc_sw128(smem_u32(bring + (size_t)stage * TM_OP_BYTES));
#pragma unroll
            for (int sub = 0; sub < SUB; ++sub) {      // the B stage feeds both row-blocks of the unit
              const uint32_t dcol = tmem_base + (uint32_t)(buf * (SUB * 128) + sub * 128);
              const uint32_t sa = abase + (uint32_t)(sub * a_sub_bytes + (kb * nops) * TM_OP_BYTES);
              const uint64_t da_hi = make_desc_sw128(sa), da_lo = make_desc_sw128(sa + TM_OP_BYTES);
              if (pl == 0) {                          // b_hi: a_hi * b_hi, then (exact mode) a_lo * b_hi
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16(dcol, desc_advance_k(da_hi, k), desc_advance_k(db, k), idesc, (kb == 0 && k == 0) ? 0u : 1u);
                if (nops == 2) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) umma_f16(dcol, desc_advance_k(da_lo, k), desc_advance_k(db, k), idesc, 1u);
                }
              } else {                                // b_lo: a_hi * b_lo
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(dcol, desc_advance_k(da_hi, k), desc_advance_k(db, k), idesc, 1u);
              }
            }
            umma_commit(&empty[stage]);
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
        umma_commit(&tfull[buf]);
        // last non-skipped tile of this strip within the CTA's range: its A slot may be refilled once these MMAs are done
        bool last_of_rb = true;
        for (int nx = tile + 1; nx < tile_end; ++nx) {
          tm_locate(a, nx, pc2, t2);
          if (t2.skip) continue;
          last_of_rb = (t2.rb != t.rb);
          break;
        }
        if (last_of_rb) umma_commit(&aempty[as]);
        if (++buf == TM_ACC_BUFS) { buf = 0; bphase ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue (warps 2..9)
    // warp (q, h): TMEM lanes 32q.. (q = warp % 4, the hardware's lane-quarter rule), columns [64h, 64h + 64) of each of
    // the unit's two accumulators (row-blocks)
    const int et = (int)threadIdx.x - 64;
    const int q = warp & 3, h = (warp - 2) >> 2;
    const bool top2 = a.passes == 2;
    int buf = 0, par = 0;
    ProbCache pc;
    uint32_t bphase = 0;
    int cur_rb = -1, cur_p = -1, p_tiles = 0, cur_ntiles = 0;
    long long cur_ka = 0;
    int cur_i[SUB];
    bool cur_valid[SUB];
#pragma unroll
    for (int s_ = 0; s_ < SUB; ++s_) { cur_i[s_] = 0; cur_valid[s_] = false; }
    float rbest[SUB], rsec[SUB];
    int rbest_j[SUB];
#pragma unroll
    for (int s_ = 0; s_ < SUB; ++s_) { rbest[s_] = -CUDART_INF_F; rsec[s_] = -CUDART_INF_F; rbest_j[s_] = -1; }
    TileInfo t;

    auto flush_rows = [&]() {          // merge this thread's strip results into the row keys
#pragma unroll
      for (int s_ = 0; s_ < SUB; ++s_) {
        if (cur_rb >= 0 && cur_valid[s_] && rbest_j[s_] >= 0) {
          const unsigned long long key = m_key(rbest[s_], rbest_j[s_]);
          if (!top2) {
            red_max_u64(a.keys + cur_ka + cur_i[s_], key);
          } else {
            const unsigned long long old = atomicMax(a.keys + cur_ka + cur_i[s_], key);
            // every partial best except the final winner loses exactly one atomicMax: it is a second-best candidate
            float cand = rsec[s_];
            if (old != 0ull) cand = fmaxf(cand, fminf(m_key_sim(old), rbest[s_]));
            if (cand > -CUDART_INF_F) red_max_u32(a.sec + cur_ka + cur_i[s_], m_ord_f32(cand));
          }
        }
        rbest[s_] = -CUDART_INF_F; rsec[s_] = -CUDART_INF_F; rbest_j[s_] = -1;
      }
    };
    auto leave_problem = [&]() {       // count this CTA's units of problem cur_p; the CTA completing the problem finishes it
      if (cur_p < 0) return;
      // release: the CTA barrier orders every epilogue thread's atomics before thread 0's acq_rel RMW on the counter
      // (cumulativity); acquire: the finisher's loads (ld.global.cg, L2) come after that RMW and the second barrier
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (et == 0) {
        int old;
        asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(a.done + cur_p), "r"(p_tiles) : "memory");
        *last_flag = (old + p_tiles == cur_ntiles) ? 1 : 0;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (*last_flag && !(a.debug & 4)) tm_finish(a, cur_p, et);
      asm volatile("bar.sync 2, 256;" ::: "memory");    // last_flag is rewritten by the next problem
      p_tiles = 0;
    };

    for (int tile = tile_begin; tile < tile_end; ++tile) {
      tm_locate(a, tile, pc, t);
      if (t.p != cur_p) {
        flush_rows();
        cur_rb = -1;
        leave_problem();
        cur_p = t.p;
        cur_ntiles = pc.ntiles;
      }
      ++p_tiles;
      if (t.skip) continue;
      if (t.rb != cur_rb) {
        flush_rows();
        cur_rb = t.rb;
        cur_ka = t.ka;
#pragma unroll
        for (int s_ = 0; s_ < SUB; ++s_) {
          cur_i[s_] = (t.mt * SUB + s_) * TM_TILE + q * 32 + lane;
          cur_valid[s_] = cur_i[s_] < t.a_len;
        }
      }
      const bool want_cols = a.cols && t.pass == 0 && !(a.debug & 2);
      const int c0 = t.nt * TM_TILE;
      unsigned long long kc = 0ull;
      if (want_cols && et < 128) kc = __ldcg(a.keys + t.kb + c0 + et);      // in flight while the MMAs finish
      mbar_wait(&tfull[buf], bphase);
      tc_fence_after();
      if (want_cols) {
        if (et < 128) {
          thr_key[par * 128 + et] = kc;
          thr_sim[par * 128 + et] = kc ? m_key_sim(kc) : -3.0e38f;    // cold column: every VALID value passes (-inf = invalid never does)
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      const int cols_valid = t.b_len - c0 - h * 64;           // valid columns among this warp's 64
#pragma unroll
      for (int s_ = 0; s_ < SUB; ++s_) {
        if (a.debug & 1) break;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * (SUB * 128) + s_ * 128 + h * 64);
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr, v0);
        tmem_ld32(taddr + 32, v1);
        tmem_ld_wait();
        float f[64];
        if (cols_valid >= 64 && cur_valid[s_]) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { f[j] = __uint_as_float(v0[j]); f[32 + j] = __uint_as_float(v1[j]); }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            f[j] = (cur_valid[s_] && j < cols_valid) ? __uint_as_float(v0[j]) : -CUDART_INF_F;
            f[32 + j] = (cur_valid[s_] && 32 + j < cols_valid) ? __uint_as_float(v1[j]) : -CUDART_INF_F;
          }
        }
        // ---- rows: tree max of the values, then (only when it beats the running best) tree min of the indices attaining it
        {
          float m32[32], m16[16], m8[8], m4[4];
#pragma unroll
          for (int j = 0; j < 32; ++j) m32[j] = fmaxf(f[2 * j], f[2 * j + 1]);
#pragma unroll
          for (int j = 0; j < 16; ++j) m16[j] = fmaxf(m32[2 * j], m32[2 * j + 1]);
#pragma unroll
          for (int j = 0; j < 8; ++j) m8[j] = fmaxf(m16[2 * j], m16[2 * j + 1]);
#pragma unroll
          for (int j = 0; j < 4; ++j) m4[j] = fmaxf(m8[2 * j], m8[2 * j + 1]);
          const float cmax = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          if (top2) {
            // second best of this tile = max over everything except ONE instance of the maximum
            int amin = 64;
#pragma unroll
            for (int j = 63; j >= 0; --j) amin = (f[j] == cmax) ? j : amin;
            float s2 = -CUDART_INF_F;
#pragma unroll
            for (int j = 0; j < 64; ++j) s2 = fmaxf(s2, (j == amin) ? -CUDART_INF_F : f[j]);
            if (cmax > rbest[s_]) {           // strict: an equal value in a later tile keeps the earlier column
              rsec[s_] = fmaxf(rbest[s_], fmaxf(rsec[s_], s2));
              rbest[s_] = cmax;
              rbest_j[s_] = c0 + h * 64 + amin;
            } else {
              rsec[s_] = fmaxf(rsec[s_], cmax);
            }
          } else if (cmax > rbest[s_]) {
            int i32[32], i16[16], i8[8], i4[4];
#pragma unroll
            for (int j = 0; j < 32; ++j) i32[j] = min((f[2 * j] == cmax) ? 2 * j : 64, (f[2 * j + 1] == cmax) ? 2 * j + 1 : 64);
#pragma unroll
            for (int j = 0; j < 16; ++j) i16[j] = min(i32[2 * j], i32[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < 8; ++j) i8[j] = min(i16[2 * j], i16[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < 4; ++j) i4[j] = min(i8[2 * j], i8[2 * j + 1]);
            rbest[s_] = cmax;
            rbest_j[s_] = c0 + h * 64 + min(min(i4[0], i4[1]), min(i4[2], i4[3]));
          }
        }
        // ---- columns: threshold filter against the column's best so far; the rare candidates update the column keys
        if (want_cols && !(a.debug & 8)) {
          const float* ts = thr_sim + par * 128 + h * 64;
          const unsigned long long* tk = thr_key + par * 128 + h * 64;
          unsigned long long* gk = a.keys + t.kb + c0 + h * 64;
          unsigned m0 = 0u, m1 = 0u;                    // this thread's candidate columns (invalid entries are -inf: never >=)
#pragma unroll
          for (int j4 = 0; j4 < 16; ++j4) {
            const float4 th = *reinterpret_cast<const float4*>(ts + j4 * 4);
            const float thv[4] = {th.x, th.y, th.z, th.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = j4 * 4 + jj;
              if (f[j] >= thv[jj]) { if (j < 32) m0 |= 1u << j; else m1 |= 1u << (j - 32); }
            }
          }
          const unsigned w0 = __reduce_or_sync(0xffffffffu, m0), w1 = __reduce_or_sync(0xffffffffu, m1);
          const int hot = __popc(w0) + __popc(w1);          // columns in which some row of this warp is a candidate
          if (hot >= 12) {
            // cold tile (thresholds not established): full warp arg-max of all 64 columns by butterfly, one key update
            // per column by the lane that ends up owning it
            const int row0 = cur_i[s_] - lane;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              unsigned v[32], id[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float x = f[g * 32 + j];
                v[j] = (x > -CUDART_INF_F) ? m_ord_f32(x) : 0u;      // 0 = invalid row / column: never wins
                id[j] = (unsigned)lane;
              }
              tm_butterfly32(v, id, lane);
              if (v[0] != 0u) {
                const unsigned long long key = ((unsigned long long)v[0] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(row0 + (int)id[0]));
                if (key > tk[g * 32 + lane]) red_max_u64(gk + g * 32 + lane, key);
              }
            }
          } else if (w0 | w1) {
            // warm tile: a handful of columns have a candidate somewhere in the warp.  Walk the set bits of the warp-wide
            // mask (uniform loop), fetch that ONE column from TMEM again, and let the candidate lanes push their keys
            // (the atomic max resolves ties: lowest row wins).  Measured alternatives, all slower: a fully unrolled
            // predicated loop over the 64 columns (~650 instructions per warp per tile), a vote + branch per column
            // (2x slower end to end), four re-reads per round, REDUX-based per-column arg-max (a CREDUX costs ~100 cycles).
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
              unsigned wm = g ? w1 : w0;
              const unsigned mm = g ? m1 : m0;
              while (wm) {
                const int jb = __ffs(wm) - 1;
                wm &= wm - 1u;
                const int j = g * 32 + jb;
                const float x = __uint_as_float(tmem_ld1(taddr + (uint32_t)j));
                tmem_ld_wait();
                if ((mm >> jb) & 1u) {
                  const unsigned long long key = m_key(x, cur_i[s_]);
                  if (key > tk[j]) red_max_u64(gk + j, key);
                }
              }
            }
          }
        }
      }
      // both accumulators of the unit are consumed: the buffer may be overwritten
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
      if (want_cols) par ^= 1;
      if (++buf == TM_ACC_BUFS) { buf = 0; bphase ^= 1; }
    }
    flush_rows();
    leave_problem();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host side
constexpr size_t TM_TAIL_BYTES = 3072 + 512;      // thresholds + barriers

size_t tm_smem_bytes(int split, int sub, int aslots, int stages) {
  const int nops = split == 3 ? 2 : 1;
  return 1024 + (size_t)aslots * sub * 2 * nops * TM_OP_BYTES + (size_t)stages * TM_OP_BYTES + TM_TAIL_BYTES;
}

int tm_stages(int split, int sub, int aslots) {
  const int nops = split == 3 ? 2 : 1;
  const size_t fixed = 1024 + (size_t)aslots * sub * 2 * nops * TM_OP_BYTES + TM_TAIL_BYTES;
  int s = (int)((227 * 1024 - fixed) / (size_t)TM_OP_BYTES);
  return s > TM_MAX_STAGES ? TM_MAX_STAGES : s;
}

// work units of one pair: super-blocks (`sub` row-blocks) of the A operand x column tiles, for each product
int tm_units(int tm, int tn, int passes, int sub) {
  return ((tm + sub - 1) / sub) * tn + (passes == 2 ? ((tn + sub - 1) / sub) * tm : 0);
}

int tm_make_plane_map(CUtensorMap* tm, const __half* base, size_t rows) {
  const uint64_t dims[2] = {128, (uint64_t)rows};
  const uint64_t strides[1] = {256};
  const uint32_t box[2] = {64u, (uint32_t)TM_TILE};
  return make_tmap_f16(tm, base, 2, dims, strides, box);
}

int launch_match_prep(const MOperD* opers_dev, const MTabInline* inl, int noper, int total_prows, bool any_ids, __half* hi, __half* lo, int* remap,
                      int* efflen, unsigned long long* keys, unsigned* sec, long long nkeys, int* done, int nprob,
                      int num_sms, cudaStream_t st) {
  if (any_ids) {
    match_scan_ids_kernel<<<noper < 4 * num_sms ? noper : 4 * num_sms, 1024, 0, st>>>(opers_dev, *inl, noper, efflen, remap);
    ++g_launches;
  }
  int blocks = total_prows / 32;
  if (blocks > 8 * num_sms) blocks = 8 * num_sms;
  if (blocks < 1) blocks = 1;
  match_prep_kernel<<<blocks, 256, 0, st>>>(opers_dev, *inl, noper, total_prows, hi, lo, remap, efflen, keys, sec, nkeys, done, nprob);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

int launch_match_tc(const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, TcMatchArgs a, int num_sms, cudaStream_t st) {
  // exact mode: ONE resident A slot (sub x 64 KB); the rest of shared memory is the B ring of 16 KB stages (nine at
  // sub = 1, five at sub = 2 where each stage feeds twice the MMAs).  Single-pass mode: two A slots.
  static const int env_aslots = getenv("SFD2_TM_ASLOTS") ? atoi(getenv("SFD2_TM_ASLOTS")) : 0;
  a.aslots = env_aslots ? env_aslots : (a.split == 3 ? 1 : 2);
  static const int env_debug = getenv("SFD2_TM_DEBUG") ? atoi(getenv("SFD2_TM_DEBUG")) : 0;
  a.debug = env_debug;
  a.stages = tm_stages(a.split, a.sub, a.aslots);
  const size_t smem = tm_smem_bytes(a.split, a.sub, a.aslots, a.stages);
  auto kern = a.sub == 2 ? tc_match_kernel<2> : tc_match_kernel<1>;
  SFD2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (a.total_tiles <= 0) return SFD2_OK;
  const int grid = a.total_tiles < num_sms ? a.total_tiles : num_sms;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue overlaps match_prep_kernel's tail
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SFD2_CUDA(cudaLaunchKernelEx(&cfg, kern, tm_hi, tm_lo, a));
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

}  // namespace sfd2
