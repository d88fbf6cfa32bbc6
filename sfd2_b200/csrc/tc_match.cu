// Mutual-NN matcher on tcgen05: sim = D0 * D1^T as a K-major x K-major GEMM (K = 128) whose
// epilogue reduces every 128 x 128 accumulator tile straight into the per-row / per-column
// arg-max keys of match.cu - the similarity matrix never exists in HBM.
//
// Operands are the fp16 hi/lo split of the fp32 descriptors (unit-norm rows, so hi+lo carries
// ~22 bits): split==3 issues d0_hi*d1_hi + d0_hi*d1_lo + d0_lo*d1_hi into one fp32 accumulator,
// which reproduces the fp32 reference arg-max on real SFD2 descriptors;
// split==1 is the single-pass fp16 variant.
//
// Each CTA owns a CONTIGUOUS range of tiles in row-major order and keeps the A row-block (both K halves, hi and lo
// planes: 64 KB) resident in shared memory while it walks along the columns; only the B tiles stream through the
// stage ring.  The first version reloaded A for every tile and was bound by L2->SM bandwidth (128 KB per 1536 MMA
// cycles per SM, ncu: tensor pipe 42 % active).
#include <math_constants.h>

#include "common.cuh"
#include "ptx.cuh"

namespace sfd2 {

using namespace ptx;

__device__ __forceinline__ unsigned m_ord_f32(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long m_key(float sim, int idx) {
  return ((unsigned long long)m_ord_f32(sim) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
}

// fp32 rows [n][128] -> fp16 hi / lo rows [n_pad][128] (rows >= n are zero)
__global__ void split_rows_kernel(const float* __restrict__ src, int n, int n_pad, __half* __restrict__ hi,
                                  __half* __restrict__ lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 elements
  if (idx >= n_pad * 32) return;
  const int r = idx >> 5;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < n) v = __ldg(reinterpret_cast<const float4*>(src) + idx);
  const float x[4] = {v.x, v.y, v.z, v.w};
  __align__(8) __half h[4];
  __align__(8) __half l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __float2half_rn(x[j]);
    l[j] = __float2half_rn(x[j] - __half2float(h[j]));
  }
  reinterpret_cast<uint2*>(hi)[idx] = *reinterpret_cast<uint2*>(h);
  reinterpret_cast<uint2*>(lo)[idx] = *reinterpret_cast<uint2*>(l);
}

// both operands of a pair in one launch, plus the reset of the arg-max keys (saves two launches and two memsets)
__global__ void split_rows2_kernel(const float* __restrict__ src0, int n0, int n0p, __half* __restrict__ hi0,
                                   __half* __restrict__ lo0, const float* __restrict__ src1, int n1, int n1p,
                                   __half* __restrict__ hi1, __half* __restrict__ lo1,
                                   unsigned long long* __restrict__ key0, unsigned long long* __restrict__ key1) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 elements
  if (idx >= (n0p + n1p) * 32) return;
  const bool second = idx >= n0p * 32;
  if (second) idx -= n0p * 32;
  const float* src = second ? src1 : src0;
  const int n = second ? n1 : n0;
  __half* hi = second ? hi1 : hi0;
  __half* lo = second ? lo1 : lo0;
  unsigned long long* key = second ? key1 : key0;
  const int r = idx >> 5;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < n) {
    v = __ldg(reinterpret_cast<const float4*>(src) + idx);
    if ((idx & 31) == 0) key[r] = 0ull;
  }
  const float x[4] = {v.x, v.y, v.z, v.w};
  __align__(8) __half h[4];
  __align__(8) __half l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __float2half_rn(x[j]);
    l[j] = __float2half_rn(x[j] - __half2float(h[j]));
  }
  reinterpret_cast<uint2*>(hi)[idx] = *reinterpret_cast<uint2*>(h);
  reinterpret_cast<uint2*>(lo)[idx] = *reinterpret_cast<uint2*>(l);
}

constexpr int TM_TILE = 128;
constexpr int TM_OP_BYTES = 128 * 128;  // 128 rows x 64 fp16
constexpr int TM_THREADS = 192;
constexpr int TM_MAX_STAGES = 6;

// One launch runs BOTH products: pass 0 = D0 * D1^T reduces its rows into p[0].row_key (the row arg-max), pass 1 =
// D1 * D0^T reduces ITS rows (= the columns of the first product) into p[1].row_key (the column arg-max).  The GEMM is
// ~3 us of tensor work either way; what costs is the epilogue, and a row reduction is thread-local (one TMEM lane = one
// row) while a column reduction needs a shared-memory transpose plus 128 atomics per warp per tile - doing the cheap
// reduction twice is ~2x faster than doing both at once.  The tiles of the two passes form one list that is cut into
// contiguous per-CTA ranges, so there is one launch, one ramp and one tail.
struct TcMatchPass {
  int n0, n1, tiles_m, tiles_n;   // rows / columns of this pass's product and its tile grid
  unsigned long long* row_key;
};
struct TcMatchArgs {
  TcMatchPass p[2];
  int split, stages, stage_bytes;
  // one-to-many (pass 0 only): the B operand is a concatenation of nseg row segments, each padded to a multiple of
  // 128 rows (seg_poff = padded starts, seg_len = valid rows); row keys are then kept per (segment, row): key index
  // seg * n0 + i, column index local to the segment
  int nseg;
  const int* seg_poff;
  const int* seg_len;
};

__device__ __forceinline__ void tm_decode(const TcMatchArgs& a, int tile, int& pass, int& mt, int& nt) {
  const int t0 = a.p[0].tiles_m * a.p[0].tiles_n;
  pass = tile >= t0 ? 1 : 0;
  const int t = pass ? tile - t0 : tile;
  const int tn = a.p[pass].tiles_n;
  mt = t / tn;
  nt = t - mt * tn;
}

__global__ void __launch_bounds__(TM_THREADS, 1)
tc_match_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const __grid_constant__ TcMatchArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nops = (a.split == 3) ? 2 : 1;
  uint8_t* aslot = smem;                                       // resident A: [kb][plane] x 16 KB
  uint8_t* bring = smem + 2 * nops * TM_OP_BYTES;              // B ring: stage = [plane] x 16 KB of one K half
  uint64_t* full = reinterpret_cast<uint64_t*>(bring + (size_t)a.stages * a.stage_bytes);
  uint64_t* empty = full + TM_MAX_STAGES;
  uint64_t* tfull = empty + TM_MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* afull = tempty + 2;
  uint64_t* aempty = afull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA_hi);
    prefetch_tmap(&tmB_hi);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    mbar_init(afull, 1); mbar_init(aempty, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_tiles = a.p[0].tiles_m * a.p[0].tiles_n + a.p[1].tiles_m * a.p[1].tiles_n;
  const int per_cta = (num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_begin = (int)blockIdx.x * per_cta;
  const int tile_end = min(tile_begin + per_cta, num_tiles);

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, prev_rb = -1;
      uint32_t phase = 0, aphase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        int pass, mt, nt;
        tm_decode(a, tile, pass, mt, nt);
        const int rb = pass ? a.p[0].tiles_m + mt : mt;       // row-block id over both passes
        const int r0 = mt * TM_TILE, c0 = nt * TM_TILE;
        const CUtensorMap* ah = pass ? &tmB_hi : &tmA_hi;
        const CUtensorMap* al = pass ? &tmB_lo : &tmA_lo;
        const CUtensorMap* bh = pass ? &tmA_hi : &tmB_hi;
        const CUtensorMap* bl = pass ? &tmA_lo : &tmB_lo;
        if (rb != prev_rb) {                    // new row-block: (re)load the resident A operand
          mbar_wait(aempty, aphase ^ 1);
          mbar_expect_tx(afull, (uint32_t)(2 * nops * TM_OP_BYTES));
          for (int kb = 0; kb < 2; ++kb) {
            tma_load_2d(aslot + (kb * nops) * TM_OP_BYTES, ah, afull, kb * 64, r0);
            if (a.split == 3) tma_load_2d(aslot + (kb * nops + 1) * TM_OP_BYTES, al, afull, kb * 64, r0);
          }
          aphase ^= 1;
          prev_rb = rb;
        }
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sb = bring + (size_t)stage * a.stage_bytes;
          mbar_expect_tx(&full[stage], (uint32_t)a.stage_bytes);
          tma_load_2d(sb, bh, &full[stage], kb * 64, c0);
          if (a.split == 3) tma_load_2d(sb + TM_OP_BYTES, bl, &full[stage], kb * 64, c0);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, 128);
      int stage = 0, buf = 0, prev_rb = -1;
      uint32_t phase = 0, bphase = 0, aphase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        int pass, mt, nt;
        tm_decode(a, tile, pass, mt, nt);
        const int rb = pass ? a.p[0].tiles_m + mt : mt;
        if (rb != prev_rb) {
          mbar_wait(afull, aphase);
          aphase ^= 1;
          prev_rb = rb;
        }
        mbar_wait(&tempty[buf], bphase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(aslot + (kb * nops) * TM_OP_BYTES);
          const uint32_t sb = smem_u32(bring + (size_t)stage * a.stage_bytes);
          const uint64_t da_hi = make_desc_sw128(sa), da_lo = make_desc_sw128(sa + TM_OP_BYTES);
          const uint64_t db_hi = make_desc_sw128(sb), db_lo = make_desc_sw128(sb + TM_OP_BYTES);
          const uint32_t dcol = tmem_base + (uint32_t)(buf * 128);
#pragma unroll 1
          for (int pass_k = 0; pass_k < a.split; ++pass_k) {
            const uint64_t da = (pass_k == 2) ? da_lo : da_hi;
            const uint64_t db = (pass_k == 1) ? db_lo : db_hi;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(dcol, desc_advance_k(da, k), desc_advance_k(db, k), idesc, (kb == 0 && pass_k == 0 && k == 0) ? 0u : 1u);
          }
          umma_commit(&empty[stage]);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[buf]);
        // last tile of this row-block (or of the CTA): the resident A may be replaced once these MMAs are done
        bool last_of_rb = (tile + 1 == tile_end);
        if (!last_of_rb) {
          int p2, m2, n2;
          tm_decode(a, tile + 1, p2, m2, n2);
          last_of_rb = (p2 ? a.p[0].tiles_m + m2 : m2) != rb;
        }
        if (last_of_rb) umma_commit(aempty);
        if (++buf == 2) { buf = 0; bphase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    int buf = 0;
    uint32_t bphase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      int pass, mt, nt;
      tm_decode(a, tile, pass, mt, nt);
      const int r0 = mt * TM_TILE;
      int c0 = nt * TM_TILE;
      const int i = r0 + q * 32 + lane;
      const int rows = a.p[pass].n0;
      int n1 = a.p[pass].n1, seg = 0;
      if (pass == 0 && a.nseg > 0) {            // which segment does this column block belong to?
        int lo = 0, hi = a.nseg - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (__ldg(a.seg_poff + mid) <= c0) lo = mid; else hi = mid - 1; }
        seg = lo;
        c0 -= __ldg(a.seg_poff + seg);          // column index local to the segment
        n1 = __ldg(a.seg_len + seg);
      }
      mbar_wait(&tfull[buf], bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128);
      // the whole 128-column row of this lane in one go: four loads in flight, one wait (the epilogue, not the
      // MMAs, bounds this kernel, and each tcgen05.ld -> wait round trip used to be paid four times per tile)
      uint32_t v[4][32];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) tmem_ld32(taddr + ch * 32, v[ch]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);   // the accumulator is in registers: the MMAs of tile + 2 may start
      float rbest = -CUDART_INF_F;              // plain float compares in the loops; keys are built once per tile
      int rbest_j = -1;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int jbase = c0 + ch * 32;
        // row arg-max over this chunk's 32 columns (thread-local: one TMEM lane = one row).  A running
        // "if (s > best)" chain is 128 dependent compare/select steps per tile and made the epilogue slower than
        // the MMAs (ncu: IPC 0.9, tensor pipe 42 %); instead: tree max of the values, then tree min of the
        // indices that attain it - lowest column wins among equal values, every step is independent.
        const int cols_valid = min(32, n1 - jbase);
        float f[32];
        if (cols_valid >= 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[ch][j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = (j < cols_valid) ? __uint_as_float(v[ch][j]) : -CUDART_INF_F;
        }
        float m16[16], m8[8], m4[4];
#pragma unroll
        for (int j = 0; j < 16; ++j) m16[j] = fmaxf(f[2 * j], f[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < 8; ++j) m8[j] = fmaxf(m16[2 * j], m16[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) m4[j] = fmaxf(m8[2 * j], m8[2 * j + 1]);
        const float cmax = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        if (cols_valid > 0 && cmax > rbest) {     // strict: an equal value in a later chunk keeps the earlier column
          int i16[16], i8[8], i4[4];
#pragma unroll
          for (int j = 0; j < 16; ++j) i16[j] = min((f[2 * j] == cmax) ? 2 * j : 64, (f[2 * j + 1] == cmax) ? 2 * j + 1 : 64);
#pragma unroll
          for (int j = 0; j < 8; ++j) i8[j] = min(i16[2 * j], i16[2 * j + 1]);
#pragma unroll
          for (int j = 0; j < 4; ++j) i4[j] = min(i8[2 * j], i8[2 * j + 1]);
          rbest = cmax;
          rbest_j = jbase + min(min(i4[0], i4[1]), min(i4[2], i4[3]));
        }
      }
      if (i < rows && rbest_j >= 0) atomicMax(a.p[pass].row_key + (size_t)seg * rows + i, m_key(rbest, rbest_j));
      if (++buf == 2) { buf = 0; bphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

static size_t tm_smem_bytes(int split, int stages) {
  return (size_t)2 * (split == 3 ? 2 : 1) * TM_OP_BYTES + (size_t)stages * TM_OP_BYTES * (split == 3 ? 2 : 1) + 1024 + 256;
}

// ws_half must hold 2 * (n0_pad + n1_pad) * 128 halves (n*_pad = n* rounded up to 128)
int launch_match_tc(const float* d0, int n0, const float* d1, int n1, int d, int split, __half* ws_half,
                    unsigned long long* row_key, unsigned long long* col_key, int num_sms, cudaStream_t st) {
  SFD2_CHECK(d == 128, SFD2_ERR_ARG, "match_tc: descriptor dim must be 128 (got %d)", d);
  if (n0 <= 0 || n1 <= 0) {   // degenerate: every row is unmatched
    SFD2_CUDA(cudaMemsetAsync(row_key, 0, sizeof(unsigned long long) * (size_t)(n0 > 0 ? n0 : 1), st));
    SFD2_CUDA(cudaMemsetAsync(col_key, 0, sizeof(unsigned long long) * (size_t)(n1 > 0 ? n1 : 1), st));
    return SFD2_OK;
  }
  const int n0p = round_up(n0, TM_TILE), n1p = round_up(n1, TM_TILE);
  __half* a_hi = ws_half;
  __half* a_lo = a_hi + (size_t)n0p * 128;
  __half* b_hi = a_lo + (size_t)n0p * 128;
  __half* b_lo = b_hi + (size_t)n1p * 128;
  split_rows2_kernel<<<cdiv((n0p + n1p) * 32, 256), 256, 0, st>>>(d0, n0, n0p, a_hi, a_lo, d1, n1, n1p, b_hi, b_lo, row_key, col_key);
  ++g_launches;
  CUtensorMap tA_hi, tA_lo, tB_hi, tB_lo;
  const uint32_t box[2] = {64u, (uint32_t)TM_TILE};
  const uint64_t strides[1] = {256};
  {
    const uint64_t dims[2] = {128, (uint64_t)n0p};
    int rc = make_tmap_f16(&tA_hi, a_hi, 2, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_f16(&tA_lo, a_lo, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {128, (uint64_t)n1p};
    int rc = make_tmap_f16(&tB_hi, b_hi, 2, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_f16(&tB_lo, b_lo, 2, dims, strides, box);
    if (rc) return rc;
  }
  TcMatchArgs a{};
  a.p[0] = TcMatchPass{n0, n1, n0p / TM_TILE, n1p / TM_TILE, row_key};
  a.p[1] = TcMatchPass{n1, n0, n1p / TM_TILE, n0p / TM_TILE, col_key};
  a.split = split;
  a.stage_bytes = TM_OP_BYTES * (split == 3 ? 2 : 1);     // one K half of the B tile (hi [+ lo])
  a.stages = 4;
  const size_t smem = tm_smem_bytes(split, a.stages);
  SFD2_CUDA(cudaFuncSetAttribute(tc_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = 2 * a.p[0].tiles_m * a.p[0].tiles_n;
  const int grid = tiles < num_sms ? tiles : num_sms;
  tc_match_kernel<<<grid, TM_THREADS, smem, st>>>(tA_hi, tA_lo, tB_hi, tB_lo, a);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

// rows of `nseg` segments (unpadded starts `off`, padded starts `poff`) -> fp16 hi / lo rows in the padded layout
__global__ void split_rows_seg_kernel(const float* __restrict__ src, const int* __restrict__ off,
                                      const int* __restrict__ poff, int nseg, int total_padded,
                                      __half* __restrict__ hi, __half* __restrict__ lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 elements
  if (idx >= total_padded * 32) return;
  const int r = idx >> 5;
  int a = 0, b = nseg - 1;
  while (a < b) { const int mid = (a + b + 1) >> 1; if (__ldg(poff + mid) <= r) a = mid; else b = mid - 1; }
  const int local = r - __ldg(poff + a), len = __ldg(off + a + 1) - __ldg(off + a);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (local < len) v = __ldg(reinterpret_cast<const float4*>(src) + (size_t)(__ldg(off + a) + local) * 32 + (idx & 31));
  const float x[4] = {v.x, v.y, v.z, v.w};
  __align__(8) __half h[4];
  __align__(8) __half l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __float2half_rn(x[j]);
    l[j] = __float2half_rn(x[j] - __half2float(h[j]));
  }
  reinterpret_cast<uint2*>(hi)[idx] = *reinterpret_cast<uint2*>(h);
  reinterpret_cast<uint2*>(lo)[idx] = *reinterpret_cast<uint2*>(l);
}

// matches0[p*nq + i] = local column of segment p (or -1 / -2), sim0[p*nq + i] = best similarity
__global__ void match_finish_seg_kernel(const unsigned long long* __restrict__ row_key,
                                        const unsigned long long* __restrict__ col_key_padded,
                                        const int* __restrict__ poff, int nq, int nseg, int mutual, float dist_th,
                                        int32_t* __restrict__ matches0, float* __restrict__ sim0) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * nseg) return;
  const int p = t / nq, i = t - p * nq;
  const unsigned long long k = row_key[t];
  if (k == 0ull) { matches0[t] = -1; sim0[t] = 0.f; return; }
  const int j = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
  const unsigned o = (unsigned)(k >> 32);
  const float s = __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
  bool ok = true;
  if (dist_th > 0.f) ok = (2.f * (1.f - s)) <= dist_th * dist_th;
  const bool row_ok = ok;
  if (ok && mutual) {
    const unsigned long long kc = col_key_padded[__ldg(poff + p) + j];
    ok = (kc != 0ull) && ((int)(0xFFFFFFFFu - (unsigned)(kc & 0xFFFFFFFFull)) == i);
  }
  matches0[t] = ok ? j : (row_ok ? -2 : -1);
  sim0[t] = s;
}

// One query set against many db sets in one grouped launch (it_loc/localize_cv2.py:705: a query against the <= 50
// retrieved db images).  seg_dev: device ints [off(nseg+1) | poff(nseg+1) | len(nseg)]; P1 = total padded db rows.
int launch_match_one_to_many(const float* q, int nq, const float* db, const int* seg_dev, int nseg, int P1, int split,
                             int mutual, float dist_th, __half* ws_half, unsigned long long* row_key,
                             unsigned long long* col_key, int32_t* matches0, float* sim0, int num_sms, cudaStream_t st) {
  const int* off = seg_dev;
  const int* poff = seg_dev + (nseg + 1);
  const int* len = seg_dev + 2 * (nseg + 1);
  SFD2_CUDA(cudaMemsetAsync(row_key, 0, sizeof(unsigned long long) * (size_t)nq * nseg, st));
  SFD2_CUDA(cudaMemsetAsync(col_key, 0, sizeof(unsigned long long) * (size_t)P1, st));
  const int nqp = round_up(nq, TM_TILE);
  __half* a_hi = ws_half;
  __half* a_lo = a_hi + (size_t)nqp * 128;
  __half* b_hi = a_lo + (size_t)nqp * 128;
  __half* b_lo = b_hi + (size_t)P1 * 128;
  split_rows_kernel<<<cdiv(nqp * 32, 256), 256, 0, st>>>(q, nq, nqp, a_hi, a_lo);
  split_rows_seg_kernel<<<cdiv(P1 * 32, 256), 256, 0, st>>>(db, off, poff, nseg, P1, b_hi, b_lo);
  g_launches += 2;
  CUtensorMap tA_hi, tA_lo, tB_hi, tB_lo;
  const uint32_t box[2] = {64u, (uint32_t)TM_TILE};
  const uint64_t strides[1] = {256};
  const uint64_t dq[2] = {128, (uint64_t)nqp}, dd[2] = {128, (uint64_t)P1};
  int rc = make_tmap_f16(&tA_hi, a_hi, 2, dq, strides, box);
  if (!rc) rc = make_tmap_f16(&tA_lo, a_lo, 2, dq, strides, box);
  if (!rc) rc = make_tmap_f16(&tB_hi, b_hi, 2, dd, strides, box);
  if (!rc) rc = make_tmap_f16(&tB_lo, b_lo, 2, dd, strides, box);
  if (rc) return rc;
  TcMatchArgs a{};
  a.split = split;
  a.stage_bytes = TM_OP_BYTES * (split == 3 ? 2 : 1);     // one K half of the B tile (hi [+ lo])
  a.stages = 4;
  // pass 0: rows = query, columns = padded db segments -> row_key[seg * nq + i];
  // pass 1: rows = padded db rows, columns = query -> col_key[padded db row]
  a.p[0] = TcMatchPass{nq, P1, nqp / TM_TILE, P1 / TM_TILE, row_key};
  a.p[1] = TcMatchPass{P1, nq, P1 / TM_TILE, nqp / TM_TILE, col_key};
  a.nseg = nseg; a.seg_poff = poff; a.seg_len = len;
  const size_t smem = tm_smem_bytes(split, a.stages);
  SFD2_CUDA(cudaFuncSetAttribute(tc_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = 2 * a.p[0].tiles_m * a.p[0].tiles_n;
  const int grid = tiles < num_sms ? tiles : num_sms;
  tc_match_kernel<<<grid, TM_THREADS, smem, st>>>(tA_hi, tA_lo, tB_hi, tB_lo, a);
  ++g_launches;
  match_finish_seg_kernel<<<cdiv(nq * nseg, 256), 256, 0, st>>>(row_key, col_key, poff, nq, nseg, mutual, dist_th, matches0, sim0);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

}  // namespace sfd2
