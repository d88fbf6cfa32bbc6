// Descriptor head evaluated ONLY where descriptors are sampled (single-pass modes: `mixed`, `fast`).
//
// The reference computes the dense 128-channel descriptor map (convDb o convDa.3, merged into one 3x3 conv "headD",
// nets/sfd2.py:340-342) on all H/4 x W/4 pixels, L2-normalises it per pixel and then bilinearly samples it at the <= topk
// keypoints (nets/extractor.py:190-208): 4 taps per keypoint = at most 16 384 of the 120 000 pixels of a 1600x1200 image.
// Here the head runs after NMS / selection, as a gathered implicit GEMM on exactly those tap pixels:
//
//     rows[4k + t][0..127] = normalize( sum_{tap, ci} act[pixel(k, t) @ tap][ci] * w[tap][co][ci] + bias )
//
// * M tile = 128 tap pixels (32 keypoints x 4 taps), N = 128, K loop = (64-channel chunk, filter tap) in the SAME order and
//   with the same weight slabs as the dense tc_conv_kernel, so every row is bit-identical to the dense map's pixel;
// * A operand: no TMA box can describe 128 unrelated pixels, so four producer warps gather them with cp.async (16-byte
//   pieces, zero-fill for filter taps outside the map = the conv's zero padding) straight into the 128-byte-swizzled
//   K-major layout the UMMA descriptor expects; three stages are in flight per thread;
// * B operand: the layer's packed [tap][128][Cin] fp16 slabs by TMA, as in the dense kernel;
// * epilogue: TMEM -> + bias -> L2 normalisation over the row's 128 channels (F.normalize, sfd2.py:342) -> fp32 row.
// sample_rows_kernel then blends the four rows of a keypoint with the bilinear weights and normalises again
// (extractor.py:206-208), in the arithmetic order of sample_kernel.
//
// Executed work: 4 topk x 128 x 9 Cin MACs instead of H/4 W/4 x 128 x 9 Cin (7x less at 1600x1200, topk 4096); the
// algorithmic FLOP count the roofline uses stays the reference's dense one.
#include "common.cuh"
#include "ptx.cuh"

namespace sfd2 {

using namespace ptx;

int g_sparse_desc = 1;    // SFD2_SPARSE_DESC=0: dense descriptor head + sample_kernel in every mode

constexpr int DS_THREADS = 192;      // warp 0: weight TMA, warp 1: MMA issuer, warps 2..5: gather producers, then epilogue
constexpr int DS_STAGES = 6;
constexpr int DS_STAGE_BYTES = 32768;   // A 16 KB (128 rows x 64 ch) + B 16 KB (128 co x 64 ch)
constexpr int DS_LAG = 3;            // cp.async groups in flight per producer thread

struct DescSparseArgs {
  const __half* act;       // convDa0 output, hi plane, [H4][Wp4][C]
  int H4, W4, Wp4, C, kchunks;
  int H, W;                // image extents (grid_sample's coordinate transform)
  const float* kpts;       // [topk][2]
  const int32_t* count;
  int topk, tap_rows;      // rows per filter tap in the packed weights (= padded Cout = 128)
  const float* bias;       // [128]
  float* rows;             // [4 * topk][128]
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// desc-map coordinates of a keypoint, exactly as sample_kernel / F.grid_sample compute them (extractor.py:199-206)
__device__ __forceinline__ void ds_coords(float px, float py, int H, int W, int H4, int W4, float& ix, float& iy) {
  const float gx = __fsub_rn(__fdiv_rn(px, __fdiv_rn((float)W, 2.f)), 1.f);
  const float gy = __fsub_rn(__fdiv_rn(py, __fdiv_rn((float)H, 2.f)), 1.f);
  ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W4), 1.f), 2.f);
  iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H4), 1.f), 2.f);
}

__global__ void __launch_bounds__(DS_THREADS, 1)
tc_desc_sparse_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ DescSparseArgs a) {
  extern __shared__ uint8_t smem_raw_ds[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_ds) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + DS_STAGES * DS_STAGE_BYTES);
  uint64_t* empty = full + DS_STAGES;
  uint64_t* tfull = empty + DS_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();                                                  // (launch_pdl) the selection's count / keypoints are final from here on
  const int n = min(max(*a.count, 0), a.topk);
  const int row0 = blockIdx.x * 128;
  if (row0 >= 4 * n) return;                                   // (uniform) nothing sampled in this tile
  if (warp == 0 && lane == 0) prefetch_tmap(&tmB);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < DS_STAGES; ++i) { mbar_init(&full[i], 128 + 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 128);
  if (threadIdx.x < 128) sbias[threadIdx.x] = __ldg(a.bias + threadIdx.x);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nsteps = a.kchunks * 9;

  if (warp == 0) {
    // ------------------------------------------------------------ weight slabs by TMA
    if (elect_one()) {
      for (int i = 0; i < nsteps; ++i) {
        const int s = i % DS_STAGES;
        const uint32_t ph = (uint32_t)((i / DS_STAGES) & 1);
        mbar_wait(&empty[s], ph ^ 1u);
        const int kc = i / 9, tap = i - kc * 9;
        mbar_expect_tx(&full[s], 16384u);
        tma_load_2d(smem + (size_t)s * DS_STAGE_BYTES + 16384, &tmB, &full[s], kc * 64, tap * a.tap_rows);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, 128);
      for (int i = 0; i < nsteps; ++i) {
        const int s = i % DS_STAGES;
        mbar_wait(&full[s], (uint32_t)((i / DS_STAGES) & 1));
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * DS_STAGE_BYTES);
        const uint64_t da = make_desc_sw128(sa), db = make_desc_sw128(sa + 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base, desc_advance_k(da, k), desc_advance_k(db, k), idesc, (i == 0 && k == 0) ? 0u : 1u);
        umma_commit(&empty[s]);
      }
      umma_commit(tfull);
    }
  } else {
    // ------------------------------------------------------------ gather producers (thread = tile row), then epilogue
    const int r = (warp & 3) * 32 + lane;                         // tile row = TMEM lane; a warp reads the lane quarter warp % 4
    const int row = row0 + r;
    const int k = row >> 2, t = row & 3;
    int px = -1000000, py = -1000000;                             // tap pixel in the descriptor map; invalid rows stay far outside
    if (k < n) {
      float ix, iy;
      ds_coords(a.kpts[2 * k], a.kpts[2 * k + 1], a.H, a.W, a.H4, a.W4, ix, iy);
      const int x = (int)floorf(ix) + (t & 1), y = (int)floorf(iy) + (t >> 1);
      if (x >= 0 && x < a.W4 && y >= 0 && y < a.H4) { px = x; py = y; }     // a tap outside the map contributes zero (zeros padding)
    }
    const uint32_t dst_row = (uint32_t)(r * 128);
    const int sw = r & 7;
    auto issue = [&](int i) {
      const int s = i % DS_STAGES;
      mbar_wait(&empty[s], (uint32_t)(((i / DS_STAGES) & 1) ^ 1));
      const int kc = i / 9, tap = i - kc * 9;
      const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
      const bool ok = (yy >= 0 && yy < a.H4 && xx >= 0 && xx < a.W4);
      const __half* src = ok ? a.act + ((size_t)yy * a.Wp4 + xx) * a.C + kc * 64 : a.act;
      const uint32_t dst = smem_u32(smem + (size_t)s * DS_STAGE_BYTES) + dst_row;
#pragma unroll
      for (int j = 0; j < 8; ++j) cp_async16(dst + (uint32_t)((j ^ sw) << 4), src + j * 8, ok ? 16 : 0);
      cp_async_commit();
    };
    auto publish = [&](int i) {          // stage i's bytes have landed: make them visible to the tensor core, signal the MMA warp
      fence_proxy_async();
      mbar_arrive(&full[i % DS_STAGES]);
    };
    for (int i = 0; i < nsteps; ++i) {
      issue(i);
      if (i >= DS_LAG) { cp_async_wait<DS_LAG>(); publish(i - DS_LAG); }
    }
    cp_async_wait<0>();
    for (int i = max(nsteps - DS_LAG, 0); i < nsteps; ++i) publish(i);
    // ---- epilogue: this thread's row = TMEM lane r
    mbar_wait(tfull, 0u);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float x[128];
    float ssp[2] = {0.f, 0.f};      // the dense epilogue sums chunks {0, 2} and {1, 3} in two warps and adds the partials: same order here
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t v[32];
      tmem_ld32(taddr + ch * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float tv = __uint_as_float(v[j]) + sbias[ch * 32 + j];
        x[ch * 32 + j] = tv;
        ssp[ch & 1] += tv * tv;
      }
    }
    const float ss = ssp[0] + ssp[1];
    // one reciprocal per pixel, then multiplies - exactly the dense head's epilogue (tc_conv.cu, epi_fn == 1)
    const float scale = (px >= 0) ? __frcp_rn(fmaxf(sqrtf(ss), 1e-12f)) : 0.f;
    if (row < 4 * a.topk) {
      float4* o = reinterpret_cast<float4*>(a.rows + (size_t)row * 128);
#pragma unroll
      for (int g = 0; g < 32; ++g) o[g] = make_float4(x[4 * g] * scale, x[4 * g + 1] * scale, x[4 * g + 2] * scale, x[4 * g + 3] * scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 128);
}

// bilinear blend of a keypoint's four tap rows + L2 normalisation: sample_kernel's arithmetic on the gathered rows
__global__ void sample_rows_kernel(const float* __restrict__ rows, int H4, int W4, int H, int W, const float* __restrict__ kpts,
                                   const int32_t* __restrict__ count, int topk, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= topk) return;
  float4* o = reinterpret_cast<float4*>(out + (size_t)warp * 128) + lane;
  if (warp >= *count) { *o = make_float4(0.f, 0.f, 0.f, 0.f); return; }
  float ix, iy;
  ds_coords(kpts[2 * warp], kpts[2 * warp + 1], H, W, H4, W4, ix, iy);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - fx, wy1 = iy - fy, wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy;
  const float wgt[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
  const int xs[4] = {x0, x1, x0, x1}, ys[4] = {y0, y0, y1, y1};
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (ys[t] >= 0 && ys[t] < H4 && xs[t] >= 0 && xs[t] < W4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(rows + ((size_t)warp * 4 + t) * 128) + lane);
      acc.x = fmaf(v.x, wgt[t], acc.x); acc.y = fmaf(v.y, wgt[t], acc.y);
      acc.z = fmaf(v.z, wgt[t], acc.z); acc.w = fmaf(v.w, wgt[t], acc.w);
    }
  }
  float ss = acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, k);
  const float nrm = sqrtf(ss);
  *o = make_float4(__fdiv_rn(acc.x, nrm), __fdiv_rn(acc.y, nrm), __fdiv_rn(acc.z, nrm), __fdiv_rn(acc.w, nrm));
}

// in: convDa0's output (hi plane is read), L: the merged descriptor head; rows: [4 * topk][128] scratch
int launch_desc_sparse(const Act& in, const Layer& L, int H, int W, const float* kpts, const int32_t* count, int topk,
                       float* rows, float* desc_out, cudaStream_t st) {
  SFD2_CHECK(in.hi && L.w_hi && L.k == 3 && L.stride == 1 && L.groups == 1 && L.cout == 128 && L.cout_tc == 128 && in.C % 64 == 0 &&
                 in.C == L.cin,
             SFD2_ERR_ARG, "desc_sparse: unsupported head shape (cin %d cout %d)", L.cin, L.cout);
  if (topk <= 0) return SFD2_OK;
  DescSparseArgs a{};
  a.act = in.hi; a.H4 = in.H; a.W4 = in.W; a.Wp4 = in.Wp; a.C = in.C; a.kchunks = in.C / 64;
  a.H = H; a.W = W; a.kpts = kpts; a.count = count; a.topk = topk; a.tap_rows = L.cout_tc;
  a.bias = L.b_dev; a.rows = rows;
  const size_t smem = 1024 + (size_t)DS_STAGES * DS_STAGE_BYTES + 256 + 512;
  SFD2_CUDA(cudaFuncSetAttribute(tc_desc_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SFD2_CUDA(launch_pdl(tc_desc_sparse_kernel, dim3(cdiv(4 * topk, 128)), dim3(DS_THREADS), smem, st, L.tm_w_hi, a));
  ++g_launches;
  SFD2_CUDA(launch_pdl(sample_rows_kernel, dim3(cdiv(topk * 32, 256)), dim3(256), 0, st, rows, in.H, in.W, H, W, kpts, count, topk, desc_out));
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

}  // namespace sfd2
