// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
//   out[pixel][co] = sum_{tap, ci} act[pixel @ tap][ci] * w[tap][co][ci]   (+ bias, ReLU, residual)
//
// * M tile  = 128 output pixels = 8 rows x 16 columns of the output map (one TMEM lane each);
// * N       = all output channels of the layer (64 / 80 / 128 / 256), one fp32 accumulator of
//             N columns in TMEM, double-buffered so the epilogue of tile i overlaps the MMAs of i+1;
// * K loop  = (tap, 64-channel chunk).  For each step TMA brings
//               A: the box {64 ch, 16 px, 8 rows} of the NHWC activation at the tap's offset - it
//                  lands as 128 rows of 128 B under the 128-byte swizzle, i.e. already the K-major
//                  UMMA operand; out-of-image coordinates are zero-filled by TMA = the conv padding.
//                  Stride-2 layers view the (even-padded) activation as [H/2][2][W/2][2][C] and
//                  pick the tap's parity plane with a 5-D box, so no im2col buffer ever exists;
//               B: the tap's [N][64] weight slab (K-major, pre-packed per layer).
// * precision: split==1 -> one fp16 MMA per step; split==3 -> a_hi*w_hi + a_hi*w_lo + a_lo*w_hi into
//   the same fp32 accumulator (activations and weights carried as fp16 hi/lo planes, ~22 bits).
// * grouped 3x3 (groups=32, 8 ch/group): "diag" mode - for each 64-channel chunk the weight slab is
//   the 64x64 block-diagonal piece, one N=64 MMA per chunk into its own 64 accumulator columns.
//
// Warp roles (192 threads, 1 CTA/SM, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA
// issuer (one elected lane), warps 2..5 = epilogue (TMEM lane quarter = warp % 4).
#include "common.cuh"
#include "ptx.cuh"

namespace sfd2 {

using namespace ptx;

constexpr int TC_TILE_H = 8, TC_TILE_W = 16;
constexpr int TC_A_BYTES = 128 * 128;  // 128 pixels x 64 fp16
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_THREADS = 192;

struct TcConvArgs {
  int Ho, Wo, tiles_x, num_tiles;
  int taps, stride, kchunks, diag;
  int n_mma, acc_cols, tmem_cols, cout, relu, split;
  int stages, stage_bytes, b_bytes;
  const float* bias;
  __half* out_hi;
  __half* out_lo;
  float* out_f32;
  int out_Wp, out_C;
  const __half* res_hi;
  const __half* res_lo;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const __grid_constant__ TcConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * a.stage_bytes);
  uint64_t* empty = full + TC_MAX_STAGES;
  uint64_t* tfull = empty + TC_MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA_hi);
    prefetch_tmap(&tmB_hi);
    if (a.split == 3) { prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nkb = a.taps * a.kchunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int y0 = (tile / a.tiles_x) * TC_TILE_H, x0 = (tile % a.tiles_x) * TC_TILE_W;
        for (int tap = 0; tap < a.taps; ++tap) {
          const int ky = (a.taps == 9) ? tap / 3 : 1, kx = (a.taps == 9) ? tap % 3 : 1;
          for (int kc = 0; kc < a.kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * a.stage_bytes;
            uint8_t* sb = sa + (a.split == 3 ? 2 : 1) * TC_A_BYTES;
            mbar_expect_tx(&full[stage], (uint32_t)a.stage_bytes);
            const int c0 = kc * 64;
            if (a.stride == 1) {
              const int cx = x0 + kx - 1, cy = y0 + ky - 1;
              tma_load_3d(sa, &tmA_hi, &full[stage], c0, cx, cy);
              if (a.split == 3) tma_load_3d(sa + TC_A_BYTES, &tmA_lo, &full[stage], c0, cx, cy);
            } else {
              const int xp = (kx + 1) & 1, yp = (ky + 1) & 1;
              const int cx = x0 + (kx - 1 - xp) / 2, cy = y0 + (ky - 1 - yp) / 2;
              tma_load_5d(sa, &tmA_hi, &full[stage], c0, xp, cx, yp, cy);
              if (a.split == 3) tma_load_5d(sa + TC_A_BYTES, &tmA_lo, &full[stage], c0, xp, cx, yp, cy);
            }
            const int brow = a.diag ? tap * 256 + kc * 64 : tap * a.acc_cols;
            const int bcol = a.diag ? 0 : c0;
            tma_load_2d(sb, &tmB_hi, &full[stage], bcol, brow);
            if (a.split == 3) tma_load_2d(sb + a.b_bytes, &tmB_lo, &full[stage], bcol, brow);
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, a.n_mma);
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t bphase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[buf], bphase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * a.stage_bytes);
          const uint32_t sb = sa + (a.split == 3 ? 2 : 1) * TC_A_BYTES;
          const uint64_t da_hi = make_desc_sw128(sa), da_lo = make_desc_sw128(sa + TC_A_BYTES);
          const uint64_t db_hi = make_desc_sw128(sb), db_lo = make_desc_sw128(sb + a.b_bytes);
          const uint32_t dcol = tmem_base + (uint32_t)(buf * a.acc_cols + (a.diag ? (kb % a.kchunks) * 64 : 0));
          const bool first = a.diag ? (kb < a.kchunks) : (kb == 0);
#pragma unroll 1
          for (int pass = 0; pass < a.split; ++pass) {
            const uint64_t da = (pass == 2) ? da_lo : da_hi;
            const uint64_t db = (pass == 1) ? db_lo : db_hi;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(dcol, desc_advance_k(da, k), desc_advance_k(db, k), idesc, (first && pass == 0 && k == 0) ? 0u : 1u);
          }
          umma_commit(&empty[stage]);           // smem slot free once these MMAs have read it
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[buf]);               // accumulator complete -> epilogue
        if (++buf == 2) { buf = 0; bphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ty = row >> 4, tx = row & 15;
    int buf = 0;
    uint32_t bphase = 0;
    const int nchunks = (a.cout + 31) / 32;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int oy = (tile / a.tiles_x) * TC_TILE_H + ty, ox = (tile % a.tiles_x) * TC_TILE_W + tx;
      const bool valid = (oy < a.Ho) && (ox < a.Wo);
      const size_t opix = (size_t)oy * a.out_Wp + ox;
      mbar_wait(&tfull[buf], bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * a.acc_cols);
      for (int ch = 0; ch < nchunks; ++ch) {
        const int c0 = ch * 32;
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        if (valid) {
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + __ldg(a.bias + c0 + j);
          if (a.res_hi) {
            const uint4* rh = reinterpret_cast<const uint4*>(a.res_hi + opix * a.out_C + c0);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 h = __ldg(rh + g);
              const __half* hh = reinterpret_cast<const __half*>(&h);
#pragma unroll
              for (int j = 0; j < 8; ++j) x[g * 8 + j] += __half2float(hh[j]);
            }
            if (a.res_lo) {
              const uint4* rl = reinterpret_cast<const uint4*>(a.res_lo + opix * a.out_C + c0);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 l = __ldg(rl + g);
                const __half* ll = reinterpret_cast<const __half*>(&l);
#pragma unroll
                for (int j = 0; j < 8; ++j) x[g * 8 + j] += __half2float(ll[j]);
              }
            }
          }
          if (a.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
          }
          if (a.out_f32) {
            float* o = a.out_f32 + opix * a.out_C + c0;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (c0 + g * 4 < a.out_C)
                *reinterpret_cast<float4*>(o + g * 4) = make_float4(x[g * 4], x[g * 4 + 1], x[g * 4 + 2], x[g * 4 + 3]);
          } else {
            __align__(16) __half hi[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) hi[j] = __float2half_rn(x[j]);
            uint4* oh = reinterpret_cast<uint4*>(a.out_hi + opix * a.out_C + c0);
#pragma unroll
            for (int g = 0; g < 4; ++g) oh[g] = reinterpret_cast<const uint4*>(hi)[g];
            if (a.out_lo) {
              __align__(16) __half lo[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) lo[j] = __float2half_rn(x[j] - __half2float(hi[j]));
              uint4* ol = reinterpret_cast<uint4*>(a.out_lo + opix * a.out_C + c0);
#pragma unroll
              for (int g = 0; g < 4; ++g) ol[g] = reinterpret_cast<const uint4*>(lo)[g];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
      if (++buf == 2) { buf = 0; bphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

// ------------------------------------------------------------------------------------ host side
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  PFN_encodeTiled enc = get_encode_tiled();
  SFD2_CHECK(enc != nullptr, SFD2_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                         bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SFD2_CHECK(r == CUDA_SUCCESS, SFD2_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank);
  return SFD2_OK;
}

static void split_f16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// Pack a folded layer's weights into the K-major fp16 hi/lo slabs the kernel's B operand reads.
int tc_encode_weights(Layer& L) {
  const int taps = L.k * L.k;
  const bool diag = (L.groups == 32);
  SFD2_CHECK(L.groups == 1 || diag, SFD2_ERR_WEIGHTS, "tc weights(%s): groups=%d unsupported", L.name.c_str(), L.groups);
  if (diag) SFD2_CHECK(L.cin == 256 && L.cout == 256, SFD2_ERR_WEIGHTS, "tc diag shape");
  else SFD2_CHECK(L.cin % 64 == 0, SFD2_ERR_WEIGHTS, "tc weights(%s): cin %% 64", L.name.c_str());
  L.cout_tc = diag ? 256 : round_up(L.cout, 16);
  const int cols = diag ? 64 : L.cin;
  const size_t rows = (size_t)taps * L.cout_tc;
  std::vector<__half> hi(rows * cols, __float2half_rn(0.f)), lo(rows * cols, __float2half_rn(0.f));
  const int cpg = L.cin / L.groups;
  for (int t = 0; t < taps; ++t)
    for (int o = 0; o < L.cout; ++o)
      for (int r = 0; r < cpg; ++r) {
        const float v = L.w[((size_t)o * cpg + r) * taps + t];
        size_t col;
        if (diag) { const int ci = (o / 8) * 8 + r; col = (size_t)(ci - 64 * (o / 64)); }
        else col = (size_t)r;
        split_f16(v, hi[((size_t)t * L.cout_tc + o) * cols + col], lo[((size_t)t * L.cout_tc + o) * cols + col]);
      }
  SFD2_CUDA(cudaMalloc(&L.w_hi, hi.size() * sizeof(__half)));
  SFD2_CUDA(cudaMalloc(&L.w_lo, lo.size() * sizeof(__half)));
  SFD2_CUDA(cudaMemcpy(L.w_hi, hi.data(), hi.size() * sizeof(__half), cudaMemcpyHostToDevice));
  SFD2_CUDA(cudaMemcpy(L.w_lo, lo.data(), lo.size() * sizeof(__half), cudaMemcpyHostToDevice));
  const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  const uint64_t strides[1] = {(uint64_t)cols * 2};
  const uint32_t box[2] = {64u, (uint32_t)(diag ? 64 : L.cout_tc)};
  int rc = make_tmap_f16(&L.tm_w_hi, L.w_hi, 2, dims, strides, box);
  if (rc) return rc;
  return make_tmap_f16(&L.tm_w_lo, L.w_lo, 2, dims, strides, box);
}

// Activation tensor maps: [0] stride-1 view {C, W, H}, box {64,16,8};
//                         [1] stride-2 view {C, 2, Wp/2, 2, Hp/2}, box {64,1,16,1,8}.
int tc_make_act_maps(const Act& t, const __half* base, CUtensorMap* s1, CUtensorMap* s2) {
  {
    const uint64_t dims[3] = {(uint64_t)t.C, (uint64_t)t.W, (uint64_t)t.H};
    const uint64_t str[2] = {(uint64_t)t.C * 2, (uint64_t)t.Wp * t.C * 2};
    const uint32_t box[3] = {64u, (uint32_t)TC_TILE_W, (uint32_t)TC_TILE_H};
    int rc = make_tmap_f16(s1, base, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[5] = {(uint64_t)t.C, 2, (uint64_t)t.Wp / 2, 2, (uint64_t)t.Hp / 2};
    const uint64_t str[4] = {(uint64_t)t.C * 2, (uint64_t)t.C * 4, (uint64_t)t.Wp * t.C * 2, (uint64_t)t.Wp * t.C * 4};
    const uint32_t box[5] = {64u, 1u, (uint32_t)TC_TILE_W, 1u, (uint32_t)TC_TILE_H};
    return make_tmap_f16(s2, base, 5, dims, str, box);
  }
}

int launch_conv_tc(const Act& in, const Layer& L, Act out, const Act* res, float* out_f32, int split, int num_sms,
                   cudaStream_t st) {
  SFD2_CHECK(in.tm != nullptr && in.hi != nullptr, SFD2_ERR_ARG, "conv_tc(%s): input has no tensor maps", L.name.c_str());
  SFD2_CHECK(in.C == L.cin && in.C % 64 == 0, SFD2_ERR_ARG, "conv_tc(%s): cin %d", L.name.c_str(), in.C);
  SFD2_CHECK(split == 1 || split == 3, SFD2_ERR_ARG, "conv_tc: split must be 1 or 3");
  const bool diag = (L.groups == 32);
  TcConvArgs a{};
  a.Ho = out.H; a.Wo = out.W;
  a.tiles_x = cdiv(out.W, TC_TILE_W);
  a.num_tiles = a.tiles_x * cdiv(out.H, TC_TILE_H);
  a.taps = L.k * L.k; a.stride = L.stride; a.diag = diag ? 1 : 0;
  a.kchunks = L.cin / 64;
  a.n_mma = diag ? 64 : L.cout_tc;
  a.acc_cols = L.cout_tc;
  int tc = 32;
  while (tc < 2 * a.acc_cols) tc <<= 1;
  SFD2_CHECK(tc <= 512, SFD2_ERR_ARG, "conv_tc(%s): accumulator too wide", L.name.c_str());
  a.tmem_cols = tc;
  a.cout = L.cout; a.relu = L.relu; a.split = split;
  a.b_bytes = a.n_mma * 128;
  a.stage_bytes = (TC_A_BYTES + a.b_bytes) * (split == 3 ? 2 : 1);
  const int smem_max = 227 * 1024;
  int stages = (smem_max - 2048) / a.stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  SFD2_CHECK(stages >= 2, SFD2_ERR_ARG, "conv_tc(%s): stage too large", L.name.c_str());
  a.stages = stages;
  a.bias = L.b_dev;
  a.out_hi = out_f32 ? nullptr : out.hi;
  a.out_lo = (out_f32 || split == 1) ? nullptr : out.lo;
  a.out_f32 = out_f32;
  a.out_Wp = out.Wp; a.out_C = out.C;
  a.res_hi = res ? res->hi : nullptr;
  a.res_lo = (res && split == 3) ? res->lo : nullptr;
  const size_t smem = (size_t)stages * a.stage_bytes + 1024 + 256;
  SFD2_CUDA(cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const CUtensorMap* tmA = in.tm + (L.stride == 2 ? 2 : 0);
  const int grid = a.num_tiles < num_sms ? a.num_tiles : num_sms;
  tc_conv_kernel<<<grid, TC_THREADS, smem, st>>>(tmA[0], tmA[1], L.tm_w_hi, L.tm_w_lo, a);
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

}  // namespace sfd2
