// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
//   out[pixel][co] = sum_{tap, ci} act[pixel @ tap][ci] * w[tap][co][ci]   (+ bias, ReLU, residual)
//
// * M tile  = 128 output pixels = 8 rows x 16 columns of the output map (one TMEM lane each); 16 rows x 8 columns
//             for the stride-1 3x3 layers, whose nine taps read ONE halo box per chunk (TcConvArgs::halo);
// * N       = all output channels of the layer (64 / 80 / 128 / 192 / 256), one fp32 accumulator of
//             N columns in TMEM, double-buffered so the epilogue of tile i overlaps the MMAs of i+1;
// * K loop  = (tap, 64-channel chunk).  For each step TMA brings
//               A: the box {64 ch, 16 px, 8 rows} of the NHWC activation at the tap's offset - it
//                  lands as 128 rows of 128 B under the 128-byte swizzle, i.e. already the K-major
//                  UMMA operand; out-of-image coordinates are zero-filled by TMA = the conv padding.
//                  Stride-2 layers view the (even-padded) activation as [H/2][2][W/2][2][C] and
//                  pick the tap's parity plane with a 5-D box, so no im2col buffer ever exists;
//               B: the tap's [N][64] weight slab (K-major, pre-packed per layer).
// * precision: split==1 -> one fp16 MMA per step; split==3 -> a_hi*w_hi + a_hi*w_lo + a_lo*w_hi
//   (activations and weights carried as fp16 hi/lo planes, ~22 bits); the two correction products of the 3x3
//   layers accumulate in their own TMEM columns (TcConvArgs::corr) and are added in the epilogue.
// * grouped 3x3 (groups=32, 8 ch/group): "diag" mode - for each 64-channel chunk the weight slab is
//   the 64x64 block-diagonal piece, one N=64 MMA per chunk into its own 64 accumulator columns; in exact
//   mode the slab is [w_hi 16 | w_lo 16] per K step and the MMAs are N = 32 / N = 16 on the non-zero blocks
//   only (TcConvArgs::cat).
// * work units = (tile pair, channel pass), dealt round-robin to the clusters (`unit` in the kernel).
//
// * weight multicast: CTAs run as clusters of 2 neighbouring tiles; each CTA fetches HALF of every
//   weight slab and TMA-multicasts it into both CTAs' shared memory, halving the L2->SM traffic of
//   the B operand (the kernel is L2-bandwidth bound otherwise).  A stage is released to the
//   producers only when BOTH CTAs' MMAs have consumed it (multicast tcgen05.commit).
// * CTA pairs (template PAIR, TcConvArgs::cg2): the exact-mode per-tap-ring layers run ONE
//   tcgen05.mma.cta_group::2 of M = 256 over the cluster instead - each CTA keeps only its half of the slab.
//   SLIM is the compile-time-specialised instantiation of that for the 1x1 layers.
//
// Warp roles (320 threads, 1 CTA/SM, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA
// issuer (one elected lane), warps 2..9 = epilogue: TMEM lane quarter = warp % 4, and the two warps
// of a quarter take alternate 32-channel chunks (the epilogue of a chunk is a serial chain of
// tcgen05.ld -> convert -> st.shared -> proxy fence -> TMA store, so two chains per scheduler
// overlap each other's latencies; with one warp per quarter the 1x1 layers were bound by it).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace sfd2 {

using namespace ptx;

constexpr int TC_TILE_H = 8, TC_TILE_W = 16;
constexpr int TC_A_BYTES = 128 * 128;  // 128 pixels x 64 fp16
constexpr int TC_MAX_STAGES = 12;   // ring slots; the 8 KB slabs of the grouped layers need the depth: a slab feeds only
                                    // 4-8 N=64 MMAs (~130-260 cycles) while an L2 -> smem TMA round trip is ~1000
constexpr int TC_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int TC_EPI_WARPS = 8;

struct TcConvArgs {
  int Ho, Wo, tiles_x, num_tiles;   // this launch covers tiles [tile_begin, num_tiles) (row-major over the tile grid)
  int tile_begin, oob_tile;         // oob_tile: a tile index below the image (all padding), for the odd tile of a cluster
  int taps, stride, kchunks, diag;
  int n_mma, acc_cols, tmem_cols, cout, relu, split;
  int nsplit;      // halo mode: the output channels are processed in nsplit passes of n_mma channels each, so that
                   // main + correction accumulators of BOTH TMEM buffers fit (N=256: 2 x (128 + 128) x 2 = 512 columns)
                   // and the epilogue of one pass overlaps the MMAs of the next; the small halo A tile is re-fetched
  int tap_rows;    // rows per tap in the packed weights (= the layer's padded Cout)
  int ncat;        // per-tap ring, exact mode, N <= 64 with a correction accumulator: see the MMA issuer
  int cat;         // grouped layers, exact mode ("diag-cat"): the weight slab of a (tap, 64-channel chunk) is the
                   // N-concatenation [w_hi | w_lo] (128 rows), so ONE N=128 MMA per K step yields a_hi*w_hi (columns
                   // 0..63 of the chunk's 128 accumulator columns) and a_hi*w_lo (columns 64..127); a_lo*w_hi is an
                   // N=64 MMA on the slab's first 64 rows into columns 64..127.  An SMEM-operand MMA at M=128 costs
                   // >= 53 cycles whatever N is, 64 at N = 128 (tools/mma_rate_probe.py), so two MMAs per K step instead of three cut the
                   // layer's tensor time by a third.  The 256 output channels run as two channel passes of two
                   // chunks (2 x 128 columns per buffer, double-buffered); chunks are independent, nothing is re-read.
  int corr;        // 1: the hi*lo / lo*hi passes accumulate in their own TMEM region (added in the epilogue)
  int nbuf;        // accumulator buffers (2 = epilogue overlaps the next tile's MMAs)
  int buf_stride;  // TMEM columns per buffer = acc_cols * (1 + corr)
  int stages, stage_bytes, b_bytes;
  // "halo" staging for stride-1 3x3 layers: the tile is 16 rows x 8 px and ONE {64 ch, 10 px, 18 rows} box per
  // 64-channel chunk serves all nine taps - each tap's A operand is the same smem tile read through a UMMA
  // descriptor whose start address is shifted by (ky*10 + kx) pixel rows and whose 8-row-group stride is one
  // halo row (1280 B).  The swizzle is a function of the absolute smem address, so the shifted views decode
  // correctly (profiles/r1_umma_halo_probe.log).  A and B then use separate rings.
  int halo, a_slots, a_slot_bytes, b_stages;
  // geometry of the A box in split-ring ("halo") mode: 3x3 layers {10 px, 18 rows} at origin (x0-1, y0-1), nine taps;
  // 1x1 layers run through the same path with a plain {8 px, 16 rows} box, one tap: their A operand comes from HBM and
  // is used once, the weights come from L2 and are reused by every tile, so A gets a ring as deep as a whole tile
  // (4 chunks) and the weight slabs a shallow one - the combined (A + B) stages left room for only two in flight
  int hw, hoff, a_plane_bytes, a_plane_off;
  int tile_w, tile_h, epi_rows, ring_bytes;
  int dbg_nob;     // experiment: skip weight reloads (timing only)
  int mc;          // cluster size (1 or 2): with 2, each CTA loads half of every weight slab and multicasts it
  int cg2;         // per-tap ring only: the cluster is a CTA PAIR running tcgen05.mma.cta_group::2 (M = 256 = both CTAs' tiles):
                   // each CTA keeps only ITS half of every weight slab (no multicast), so a stage is A + B/2 - a third less
                   // shared-memory fill per tile for the 256-wide 1x1 layers and three stages instead of two.  The even CTA
                   // issues for both; loads of both CTAs complete on its `full` barriers, its commits arrive on both CTAs'
                   // `empty` / `tfull`, and both CTAs' epilogue warps arrive on its `tempty`
  int stiles;      // staging tiles per epilogue warp (1; CTA-pair layers have the room for 2, or 3 with a residual: the TMA
                   // store of chunk i drains - and the residual of chunk i+1 arrives - while chunk i+1 is converted)
  int iters;       // tiles per CTA (same for every CTA so cluster peers stay in lock step)
  const float* bias;
  int out_mode;    // 0: fp16 hi plane only, 1: fp16 hi + lo planes, 2: fp32
  int epi_fn;      // fp32 head epilogues: 0 none, 1 = L2-normalise the pixel's channels (F.normalize, sfd2.py:342),
                   // 2 = exp / (sum_65 exp + 1e-5), channels 0..63 (sfd2.py:330-333); both need a first pass over TMEM
  int has_res;     // residual planes to add: 0 none, 1 hi, 2 hi + lo
  // fused ConvSta (1x1 256 -> 3, nets/sfd2.py:303,345) on this layer's OUTPUT (rb2c3 = out4): the epilogue already
  // holds every output pixel's 256 channels in registers chunk by chunk, so the three dot products cost 96 FMAs per
  // chunk and save re-reading the 123 MB activation in a separate kernel.  Computed in fp32 on the value the
  // planes carry (hi + lo), like the standalone sta_kernel.
  float* sta_out;       // [Ho*Wo][3]; NULL = not fused
  float sta_b[3];
};

// ConvSta weights [256 ci][3] travel as a kernel parameter: every lane reads the same element at the same time, which
// is exactly what the constant bank behind __grid_constant__ parameters serves in one broadcast; shared memory has no
// 3 KB to spare (rb2c3's two 96 KB stages + staging fill the 227 KB to within 400 bytes).
struct TcStaW { float w[256 * 3]; };

constexpr int TC_HALO_W = 10, TC_HALO_H = 18;
constexpr int TC_HALO_BYTES = TC_HALO_W * TC_HALO_H * 128;           // 23040
constexpr int TC_HALO_SLOT = 23 * 1024;                                // per plane, 1024-aligned
constexpr int TC_STAGING_BYTES = TC_EPI_WARPS * 4096;   // one (32 px x 128 B) staging tile per epilogue warp
constexpr int TC_BIAS_BYTES = 1024 + 128;       // up to 288 floats (256 + 32-column over-read)
constexpr int TC_BAR_BYTES = 512;               // mbarriers + TMEM slot
constexpr int TC_EXCH_BYTES = 2 * TC_EPI_WARPS * 32 * 4;   // head epilogues: partial-sum exchange between paired warps

// PAIR: the instantiations that contain the cta_group::2 instructions - such a kernel can only be launched as clusters of two
// (a cluster-of-one launch fails with "cluster misconfiguration"), so the single-CTA / multicast path is its own instantiation
template <bool PAIR, bool SLIM>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo,
               const __grid_constant__ CUtensorMap tmR_hi, const __grid_constant__ CUtensorMap tmR_lo,
               const __grid_constant__ TcConvArgs a, const __grid_constant__ TcStaW sw) {
  pdl_launch_dependents();       // the next layer's prologue may overlap this layer's last wave (it waits below)
  // SLIM: the instantiation for the exact-mode 1x1 pair layers (6 of the 17 launches of an image, all epilogue-bound).  Its
  // configuration is fixed at compile time, so the halo path, the grouped / head / fp32 epilogues and the single-CTA variants
  // drop out: the generic kernel is ~5800 instructions and a quarter of the epilogue warps' stall samples in these layers
  // were instruction-fetch misses (`no_inst`) on the branches around the variants they never take
  const int k_halo = SLIM ? 0 : a.halo, k_taps = SLIM ? 1 : a.taps, k_stride = SLIM ? 1 : a.stride, k_split = SLIM ? 3 : a.split;
  const int k_diag = SLIM ? 0 : a.diag, k_cat = SLIM ? 0 : a.cat, k_ncat = SLIM ? 0 : a.ncat, k_nsplit = SLIM ? 1 : a.nsplit;
  const int k_corr = SLIM ? 0 : a.corr, k_epi_fn = SLIM ? 0 : a.epi_fn, k_out_mode = SLIM ? 1 : a.out_mode;
  const int k_dbg_nob = SLIM ? 0 : a.dbg_nob, k_mc = SLIM ? 2 : a.mc;
  const bool k_cg2 = SLIM ? true : (PAIR && a.cg2);
  const uint32_t crank = (k_mc > 1) ? cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << k_mc) - 1u);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + (size_t)a.ring_bytes;                        // 1024-aligned
  const int staging_bytes = a.stiles * TC_STAGING_BYTES;
  float* sbias = reinterpret_cast<float*>(staging + staging_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + staging_bytes + TC_BIAS_BYTES);
  uint64_t* empty = full + TC_MAX_STAGES;
  uint64_t* tfull = empty + TC_MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* resbar = tempty + 2;                                         // [8 epilogue warps]
  uint64_t* fullA = resbar + 8;                                          // halo mode: A-tile ring
  uint64_t* emptyA = fullA + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(emptyA + 4);
  uint64_t* resbar_x = emptyA + 5;                                       // [8 warps][2]: residual barriers of staging tiles 1, 2
  float* exch = reinterpret_cast<float*>(staging + staging_bytes + TC_BIAS_BYTES + TC_BAR_BYTES);   // heads only: [2][8 warps][32]

  // The warp index goes through a shuffle so that the compiler KNOWS it is warp-uniform: the producer and MMA-issuer
  // loops below are run by the whole warp (ring indices, phases, descriptors stay in uniform registers) and only the
  // TMA / MMA / commit instructions themselves sit behind elect.sync.  With the loops inside a single elected lane
  // every descriptor went through vector registers + R2UR and the issue loop took ~300 cycles per tap - as long as
  // the four single-pass MMAs it issues (ncu source view: 80 dependent instructions per tap).
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA_hi);
    prefetch_tmap(&tmB_hi);
    if (k_split == 3) { prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < TC_MAX_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], k_cg2 ? 1u : (uint32_t)k_mc); }
    for (int i = 0; i < 4; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], k_cg2 ? 2 * TC_EPI_WARPS : TC_EPI_WARPS); }
    for (int i = 0; i < TC_EPI_WARPS; ++i) { mbar_init(&resbar[i], 1); mbar_init(&resbar_x[2 * i], 1); mbar_init(&resbar_x[2 * i + 1], 1); }
    fence_barrier_init();
  }
  if (warp == 2) { if (k_cg2) tmem_alloc_cg2(tmem_slot, (uint32_t)a.tmem_cols); else tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols); }
  for (int i = threadIdx.x; i < (int)(TC_BIAS_BYTES / sizeof(float)); i += blockDim.x)
    sbias[i] = (i < ((a.cout + 31) / 32) * 32) ? __ldg(a.bias + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  if (k_mc > 1) cluster_sync_all();   // peers' barriers are initialised before anyone multicasts into them
  tc_fence_after();
  pdl_wait();                          // everything above touched only weights / parameters; activations from here on
  const uint32_t tmem_base = *tmem_slot;
  const int nkb = k_taps * a.kchunks;
  // Unit of iteration `it` = (tile, channel pass nh).  Units are dealt round-robin to the CLUSTERS: cluster-unit
  // v = cluster + it * #clusters covers tile pair v / nsplit (tile = pair * mc + rank: cluster peers consume the same weight
  // slabs in lock step, which is what the multicast needs) and channel pass v % nsplit.  With the two passes of a wide
  // exact-mode layer as separate units, 950 tiles on 74 clusters take 13 half-tile rounds instead of 7 whole ones.  A
  // cluster whose FIRST tile is out of range stops; otherwise an out-of-range tile is processed as an all-padding dummy.
  const int ncl = (int)gridDim.x / k_mc, cl = ((int)blockIdx.x - (int)crank) / k_mc;
  auto unit = [&](int it, int& tile, int& nh) -> bool {
    const int v = cl + it * ncl;
    const int tp = (k_nsplit == 2) ? (v >> 1) : v;
    nh = (k_nsplit == 2) ? (v & 1) : 0;
    tile = a.tile_begin + tp * k_mc + (int)crank;
    const bool ok = a.tile_begin + tp * k_mc < a.num_tiles;
    if (tile >= a.num_tiles) tile = a.oob_tile;
    return ok;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (k_halo) {
      const int planes = (k_split == 3) ? 2 : 1;
      uint8_t* bring = smem + (size_t)a.a_slots * a.a_slot_bytes;
      int sa = 0, sb = 0, nb_loaded = 0;
      uint32_t pha = 0, phb = 0;
      for (int it = 0; it < a.iters; ++it) {
        int tile, nh;
        if (!unit(it, tile, nh)) break;
        const int y0 = (tile / a.tiles_x) * a.tile_h, x0 = (tile % a.tiles_x) * a.tile_w;
        for (int kc = k_cat ? nh * 2 : 0; kc < (k_cat ? nh * 2 + 2 : a.kchunks); ++kc) {
          mbar_wait(&emptyA[sa], pha ^ 1);
          uint8_t* slot = smem + (size_t)sa * a.a_slot_bytes;
          if (elect_one()) {
            if (k_cg2) {    // both CTAs' halo tiles complete on the leader's barrier
              if (crank == 0) mbar_expect_tx(&fullA[sa], 2u * (uint32_t)(planes * a.a_plane_bytes));
              tma_load_3d_cg2(slot, &tmA_hi, &fullA[sa], kc * 64, x0 - a.hoff, y0 - a.hoff);
              if (planes == 2) tma_load_3d_cg2(slot + a.a_plane_off, &tmA_lo, &fullA[sa], kc * 64, x0 - a.hoff, y0 - a.hoff);
            } else {
            mbar_expect_tx(&fullA[sa], (uint32_t)(planes * a.a_plane_bytes));
            tma_load_3d(slot, &tmA_hi, &fullA[sa], kc * 64, x0 - a.hoff, y0 - a.hoff);
            if (planes == 2) tma_load_3d(slot + a.a_plane_off, &tmA_lo, &fullA[sa], kc * 64, x0 - a.hoff, y0 - a.hoff);
            }
          }
          __syncwarp();
          if (++sa == a.a_slots) { sa = 0; pha ^= 1; }
          for (int tap = 0; tap < k_taps; ++tap) {
            const int brow = k_cat ? (tap * 4 + kc) * 128 : (k_diag ? tap * 256 + kc * 64 : tap * a.tap_rows + nh * a.n_mma);
            const int bcol = k_diag ? 0 : kc * 64;
            for (int pl = 0; pl < (k_cat ? 1 : planes); ++pl) {
              mbar_wait(&empty[sb], phb ^ 1);
              uint8_t* bs = bring + (size_t)sb * a.b_bytes;
              const CUtensorMap* tb = pl ? &tmB_lo : &tmB_hi;
              if (elect_one()) {
                if (k_dbg_nob && nb_loaded >= a.b_stages) {
                  mbar_arrive(&full[sb]);            // EXPERIMENT (SFD2_TC_DEBUG_NOB): no weight traffic after the first fill
                } else {
                if (k_cg2) {     // my half of the slab's rows stays here; the pair's MMA reads both halves
                  if (crank == 0) mbar_expect_tx(&full[sb], 2u * (uint32_t)a.b_bytes);
                  tma_load_2d_cg2(bs, tb, &full[sb], bcol, brow + (int)crank * (a.n_mma / 2));
                } else if (k_mc > 1) {
                  mbar_expect_tx(&full[sb], (uint32_t)a.b_bytes);
                  const int ro = (int)crank * (a.n_mma / 2);
                  tma_load_2d_mc(bs + ro * 128, tb, &full[sb], bcol, brow + ro, cmask);
                } else {
                  mbar_expect_tx(&full[sb], (uint32_t)a.b_bytes);
                  tma_load_2d(bs, tb, &full[sb], bcol, brow);
                }
                }
              }
              ++nb_loaded;
              __syncwarp();
              if (++sb == a.b_stages) { sb = 0; phb ^= 1; }
            }
          }
        }
      }
    } else {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < a.iters; ++it) {
        int tile, nh;
        if (!unit(it, tile, nh)) break;
        const int y0 = (tile / a.tiles_x) * a.tile_h, x0 = (tile % a.tiles_x) * a.tile_w;
        for (int tap = 0; tap < k_taps; ++tap) {
          const int ky = (k_taps == 9) ? tap / 3 : 1, kx = (k_taps == 9) ? tap % 3 : 1;
          for (int kc = 0; kc < a.kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * a.stage_bytes;
            uint8_t* sb = sa + (k_split == 3 ? 2 : 1) * TC_A_BYTES;
            const int c0 = kc * 64;
            if (k_cg2) {
              if (elect_one()) {
                // both CTAs' bytes complete on the leader's barrier (its own arrival is this expect_tx)
                if (crank == 0) mbar_expect_tx(&full[stage], 2u * (uint32_t)a.stage_bytes);
                if (k_stride == 1) {
                  const int cx = x0 + kx - 1, cy = y0 + ky - 1;
                  tma_load_3d_cg2(sa, &tmA_hi, &full[stage], c0, cx, cy);
                  if (k_split == 3) tma_load_3d_cg2(sa + TC_A_BYTES, &tmA_lo, &full[stage], c0, cx, cy);
                } else {
                  const int xp = (kx + 1) & 1, yp = (ky + 1) & 1;
                  const int cx = x0 + (kx - 1 - xp) / 2, cy = y0 + (ky - 1 - yp) / 2;
                  tma_load_5d_cg2(sa, &tmA_hi, &full[stage], c0, xp, cx, yp, cy);
                  if (k_split == 3) tma_load_5d_cg2(sa + TC_A_BYTES, &tmA_lo, &full[stage], c0, xp, cx, yp, cy);
                }
                const int brow2 = tap * a.tap_rows + (int)crank * (a.n_mma / 2);   // my half of the slab's rows, kept here
                tma_load_2d_cg2(sb, &tmB_hi, &full[stage], c0, brow2);
                if (k_split == 3) tma_load_2d_cg2(sb + a.b_bytes, &tmB_lo, &full[stage], c0, brow2);
              }
              __syncwarp();
              if (++stage == a.stages) { stage = 0; phase ^= 1; }
              continue;
            }
            if (elect_one()) {
            mbar_expect_tx(&full[stage], (uint32_t)a.stage_bytes);
            if (k_stride == 1) {
              const int cx = x0 + kx - 1, cy = y0 + ky - 1;
              tma_load_3d(sa, &tmA_hi, &full[stage], c0, cx, cy);
              if (k_split == 3) tma_load_3d(sa + TC_A_BYTES, &tmA_lo, &full[stage], c0, cx, cy);
            } else {
              const int xp = (kx + 1) & 1, yp = (ky + 1) & 1;
              const int cx = x0 + (kx - 1 - xp) / 2, cy = y0 + (ky - 1 - yp) / 2;
              tma_load_5d(sa, &tmA_hi, &full[stage], c0, xp, cx, yp, cy);
              if (k_split == 3) tma_load_5d(sa + TC_A_BYTES, &tmA_lo, &full[stage], c0, xp, cx, yp, cy);
            }
            const int brow = k_diag ? tap * 256 + kc * 64 : tap * a.tap_rows;
            const int bcol = k_diag ? 0 : c0;
            if (k_mc > 1) {   // my half of the slab, delivered to both CTAs
              const int hrows = a.n_mma / 2;
              const int ro = (int)crank * hrows;
              tma_load_2d_mc(sb + ro * 128, &tmB_hi, &full[stage], bcol, brow + ro, cmask);
              if (k_split == 3) tma_load_2d_mc(sb + a.b_bytes + ro * 128, &tmB_lo, &full[stage], bcol, brow + ro, cmask);
            } else {
              tma_load_2d(sb, &tmB_hi, &full[stage], bcol, brow);
              if (k_split == 3) tma_load_2d(sb + a.b_bytes, &tmB_lo, &full[stage], bcol, brow);
            }
            }
            __syncwarp();
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // (whole warp runs the loops; tcgen05.mma / commit come from the one lane elect.sync picks - always the same)
    if (k_halo) {
      const bool pair = k_cg2;
      const uint32_t idesc = make_idesc_f16(pair ? 256 : 128, k_cat ? 16 : a.n_mma);
      const uint32_t idesc_cat = make_idesc_f16(128, 32);
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
        if (k_cg2) umma_f16_cg2(d, da, db, idesc, acc); else umma_f16(d, da, db, idesc, acc);
      };
      auto commit = [&](uint64_t* bar, bool both) {     // both: the barrier exists in both CTAs of a multicast / pair cluster
        if (k_cg2) umma_commit_cg2(bar, cmask);
        else if (both && k_mc > 1) umma_commit_mc(bar, cmask);
        else umma_commit(bar);
      };
      const uint32_t bring = smem_u32(smem + (size_t)a.a_slots * a.a_slot_bytes);
      int sa = 0, sb = 0, buf = 0;
      uint32_t pha = 0, phb = 0, bphase = 0;
      for (int it = 0; it < a.iters; ++it) {
        int tile, nh;
        if (!unit(it, tile, nh) || (pair && crank != 0)) break;      // CTA pair: the even CTA issues for both
        mbar_wait(&tempty[buf], bphase ^ 1);
        tc_fence_after();
        for (int kc = k_cat ? nh * 2 : 0; kc < (k_cat ? nh * 2 + 2 : a.kchunks); ++kc) {
          mbar_wait(&fullA[sa], pha);
          tc_fence_after();
          const uint32_t abase = smem_u32(smem + (size_t)sa * a.a_slot_bytes);
          const uint32_t dcol = tmem_base + (uint32_t)(buf * a.buf_stride + (k_cat ? (kc & 1) * 128 : (k_diag ? kc * 64 : 0)));
          const uint32_t ccol = k_cat ? dcol + 64u : (k_corr ? dcol + (uint32_t)a.acc_cols : dcol);
          for (int tap = 0; tap < k_taps; ++tap) {
            const uint32_t off = (uint32_t)(((tap / 3) * a.hw + (tap % 3)) * 128);
            const uint64_t da_hi = make_desc_sw128_sbo(abase + off, (uint32_t)a.hw * 128);
            const uint64_t da_lo = make_desc_sw128_sbo(abase + (uint32_t)a.a_plane_off + off, (uint32_t)a.hw * 128);
            const bool first = (tap == 0) && (k_diag || kc == 0);
            if (k_cat) {
              // K step k of the chunk = input channels 16k.. = two whole groups, whose 16 outputs are the only non-zero rows
              // of the block-diagonal slab: rows [32k, 32k+32) of the slab hold [w_hi 16 | w_lo 16] of exactly those.
              // a_hi x both (N = 32: main | correction columns of the 16 channels), then a_lo x w_hi (N = 16) into the
              // correction columns - which the N = 32 MMA of this tap has already initialised when tap == 0
              mbar_wait(&full[sb], phb);
              tc_fence_after();
              const uint64_t dbc = make_desc_sw128(bring + (uint32_t)(sb * a.b_bytes));
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)       // 32 slab rows = 4096 B further on; + the K offset inside the 128-byte rows
                  umma_f16(dcol + 32u * k, desc_advance_k(da_hi, k), desc_advance_k(dbc, k) + (uint64_t)(k * (4096 >> 4)), idesc_cat, first ? 0u : 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16(dcol + 32u * k + 16u, desc_advance_k(da_lo, k), desc_advance_k(dbc, k) + (uint64_t)(k * (4096 >> 4)), idesc, 1u);
                if (k_mc > 1) umma_commit_mc(&empty[sb], cmask); else umma_commit(&empty[sb]);
              }
              __syncwarp();
              if (++sb == a.b_stages) { sb = 0; phb ^= 1; }
              continue;
            }
            // weight slab, hi plane: main product, then (exact mode) a_lo * w_hi into the correction accumulator
            mbar_wait(&full[sb], phb);
            tc_fence_after();
            uint64_t db = make_desc_sw128(bring + (uint32_t)(sb * a.b_bytes));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma(dcol, desc_advance_k(da_hi, k), desc_advance_k(db, k), (first && k == 0) ? 0u : 1u);
              if (k_split == 3) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  mma(ccol, desc_advance_k(da_lo, k), desc_advance_k(db, k), (k_corr && first && k == 0) ? 0u : 1u);
              }
              commit(&empty[sb], true);
            }
            __syncwarp();
            if (++sb == a.b_stages) { sb = 0; phb ^= 1; }
            if (k_split == 3) {                 // weight slab, lo plane: a_hi * w_lo
              mbar_wait(&full[sb], phb);
              tc_fence_after();
              db = make_desc_sw128(bring + (uint32_t)(sb * a.b_bytes));
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  mma(ccol, desc_advance_k(da_hi, k), desc_advance_k(db, k), 1u);
                commit(&empty[sb], true);
              }
              __syncwarp();
              if (++sb == a.b_stages) { sb = 0; phb ^= 1; }
            }
          }
          if (elect_one()) commit(&emptyA[sa], false);   // halo tile free once all nine taps have read it
          __syncwarp();
          if (++sa == a.a_slots) { sa = 0; pha ^= 1; }
        }
        if (elect_one()) commit(&tfull[buf], false);
        __syncwarp();
        if (++buf == a.nbuf) { buf = 0; bphase ^= 1; }
      }
    } else {
      const uint32_t idesc = make_idesc_f16(k_cg2 ? 256 : 128, a.n_mma);
      const uint32_t idesc_ncat = make_idesc_f16(128, 2 * a.n_mma);
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t bphase = 0;
      for (int it = 0; it < a.iters; ++it) {
        int tile, nh;
        if (!unit(it, tile, nh) || (k_cg2 && crank != 0)) break;     // CTA pair: the even CTA issues for both
        mbar_wait(&tempty[buf], bphase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * a.stage_bytes);
          const uint32_t sb = sa + (k_split == 3 ? 2 : 1) * TC_A_BYTES;
          const uint64_t da_hi = make_desc_sw128(sa), da_lo = make_desc_sw128(sa + TC_A_BYTES);
          const uint64_t db_hi = make_desc_sw128(sb), db_lo = make_desc_sw128(sb + a.b_bytes);
          const uint32_t dcol = tmem_base + (uint32_t)(buf * a.buf_stride + (k_diag ? (kb % a.kchunks) * 64 : 0));
          const bool first = k_diag ? (kb < a.kchunks) : (kb == 0);
          if (elect_one()) {
          if (k_ncat) {
            // narrow layers (N <= 64): an SMEM-operand MMA at M = 128 costs >= 53 cycles whatever N is and 64 at N = 128
            // (tools/mma_rate_probe.py), so a_hi x w_hi and a_hi x w_lo go out as ONE MMA over the adjacent [w_hi | w_lo]
            // slabs of the stage - it fills the main columns and, right behind them, the correction columns
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(dcol, desc_advance_k(da_hi, k), desc_advance_k(db_hi, k), idesc_ncat, (first && k == 0) ? 0u : 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(dcol + (uint32_t)a.acc_cols, desc_advance_k(da_lo, k), desc_advance_k(db_hi, k), idesc, 1u);
          } else
#pragma unroll 1
          for (int pass = 0; pass < k_split; ++pass) {
            const uint64_t da = (pass == 2) ? da_lo : da_hi;
            const uint64_t db = (pass == 1) ? db_lo : db_hi;
            // the tensor core truncates (round-toward-zero) every time it adds into the fp32 accumulator, so
            // the two small correction passes go to their own accumulator: truncation there is relative to a
            // ~2^-11 smaller magnitude, and the main accumulator sees 3x fewer additions.
            const bool to_corr = k_corr && pass > 0;
            const uint32_t d = to_corr ? dcol + (uint32_t)a.acc_cols : dcol;
            const int first_pass = to_corr ? 1 : 0;
            if (k_cg2) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_cg2(d, desc_advance_k(da, k), desc_advance_k(db, k), idesc, (first && pass == first_pass && k == 0) ? 0u : 1u);
            } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d, desc_advance_k(da, k), desc_advance_k(db, k), idesc, (first && pass == first_pass && k == 0) ? 0u : 1u);
            }
          }
          // smem slot free once these MMAs have read it (in both CTAs when the slab is multicast / the MMA spans the pair)
          if (k_cg2) umma_commit_cg2(&empty[stage], cmask);
          else if (k_mc > 1) umma_commit_mc(&empty[stage], cmask); else umma_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) { if (k_cg2) umma_commit_cg2(&tfull[buf], cmask); else umma_commit(&tfull[buf]); }   // accumulator complete -> epilogue
        __syncwarp();
        if (++buf == a.nbuf) { buf = 0; bphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Per 32-channel chunk: TMEM -> registers -> (+bias, +residual, ReLU, fp16 hi/lo split) -> this warp's
    // swizzled staging tile in smem -> ONE TMA store per plane (box {32 ch, 16 px, 2 rows}).  TMA clips
    // partial tiles and writes full lines; the threads never touch global memory (per-lane 16-byte
    // stores at a 512-byte stride cost ~8k LSU cycles per tile in the first version and bounded every 1x1
    // layer).  Residual tiles arrive the same way (TMA load into the staging tile, one own chunk ahead).
    // Warp (q, h) owns TMEM lanes 32q.. and the chunks h, h + 2, h + 4, ... of every channel pass.
    const int ew = warp - 2;
    const int q = warp & 3;
    const int h = ew >> 2;
    const int nst = a.stiles;
    uint8_t* st = staging + ew * nst * 4096;     // this warp's staging tile(s); `st` = the tile of the current chunk
    uint8_t* const st0 = st;
    auto wres_of = [&](int ti) -> uint64_t* { return ti == 0 ? resbar + ew : resbar_x + 2 * ew + (ti - 1); };
    uint64_t* wres = resbar + ew;
    int ti = 0;                                  // staging tile of the current chunk (rotates when nst > 1)
    uint32_t rph = 0u;                           // residual-barrier phase per tile (bit ti)
    int buf = 0;
    uint32_t bphase = 0;
    const int nchunks = (k_nsplit > 1) ? a.n_mma / 32 : (a.cout + 31) / 32;   // per channel pass
    const int nstore = (k_epi_fn == 2) ? 2 : nchunks;   // the softmax head writes channels 0..63 only
    const int r = lane;                          // row of this warp's 32-pixel box (2 tile rows x 16 px)
    auto tile_xy = [&](int it, int& x0, int& y0, int& nh) -> bool {
      int tile;
      const bool ok = unit(it, tile, nh);
      x0 = (tile % a.tiles_x) * a.tile_w;
      y0 = (tile / a.tiles_x) * a.tile_h + a.epi_rows * q;
      return ok;
    };
    auto issue_res = [&](int it, int ch, int tdst) {  // lane 0 only (layers with a residual run a single channel pass)
      int x0, y0, nh;
      if (ch >= nstore || it >= a.iters || !tile_xy(it, x0, y0, nh)) return;
      uint64_t* wr = wres_of(tdst);
      uint8_t* sd = st0 + tdst * 4096;
      if (k_dbg_nob & 8) { mbar_arrive(wr); return; }
      mbar_expect_tx(wr, a.has_res == 2 ? 4096u : 2048u);
      tma_load_3d(sd, &tmR_hi, wr, ch * 32, x0, y0);
      if (a.has_res == 2) tma_load_3d(sd + 2048, &tmR_lo, wr, ch * 32, x0, y0);
    };
    if (a.has_res && lane == 0) issue_res(0, h, 0);  // the residual of this warp's first chunk
    for (int it = 0; it < a.iters; ++it) {
      int x0, y0, nh;
      if (!tile_xy(it, x0, y0, nh)) break;
      const int cbase = nh * a.n_mma;             // first output channel of this pass (0 unless nsplit > 1)
      mbar_wait(&tfull[buf], bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * a.buf_stride);
      if (k_epi_fn) {
        // Head epilogues (fp32 output, <= 128 channels = at most two own chunks per warp): every pixel needs a reduction
        // over ALL its channels before anything can be written (sum x^2 for F.normalize, sfd2.py:342; sum_65 exp for the
        // detector, sfd2.py:330-333).  Each warp reads only its own chunks, keeps them in registers, and the two warps
        // of a lane quarter exchange their partial sums through shared memory (named barrier per quarter) - the
        // accumulator is read exactly once and handed back to the MMA issuer before any store is staged.
        float xo[2][32];
        float part = 0.f;
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int ch = h + 2 * ci;
          if (ch < nchunks) {
            uint32_t v[32];
            tmem_ld32(taddr + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) xo[ci][j] = __uint_as_float(v[j]);
            if (k_corr) {
              tmem_ld32(taddr + a.acc_cols + ch * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) xo[ci][j] += __uint_as_float(v[j]);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float t = xo[ci][j] + sbias[ch * 32 + j];
              xo[ci][j] = t;
              if (ch * 32 + j < a.cout) part += (k_epi_fn == 1) ? t * t : expf(t);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (k_cg2) mbar_arrive_cluster(&tempty[buf], 0); else mbar_arrive(&tempty[buf]); }   // this warp's share of the accumulator is in registers
        float* ex = exch + ((it & 1) * TC_EPI_WARPS + ew) * 32;    // double-buffered by tile parity
        ex[lane] = part;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // the two warps of this lane quarter
        const float acc = part + exch[((it & 1) * TC_EPI_WARPS + (ew ^ 4)) * 32 + lane];
        // one reciprocal per pixel, then multiplies (<= 1 ulp from the reference's per-element division)
        const float row_scale = __frcp_rn((k_epi_fn == 1) ? fmaxf(sqrtf(acc), 1e-12f) : (acc + 0.00001f));
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int ch = h + 2 * ci;
          if (ch < nstore) {
            if (lane == 0) bulk_wait_read<0>();     // this warp's previous store has read the staging tile
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float t = (k_epi_fn == 1) ? xo[ci][j] * row_scale : expf(xo[ci][j]) * row_scale;
              xo[ci][j] = a.relu ? fmaxf(t, 0.f) : t;
            }
#pragma unroll
            for (int g = 0; g < 8; ++g)               // fp32 rows of 128 B, SWIZZLE_128B
              *reinterpret_cast<float4*>(st + r * 128 + ((g ^ (r & 7)) << 4)) =
                  make_float4(xo[ci][g * 4], xo[ci][g * 4 + 1], xo[ci][g * 4 + 2], xo[ci][g * 4 + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tmO_hi, st, cbase + ch * 32, x0, y0);
              bulk_commit();
            }
          }
        }
        if (++buf == a.nbuf) { buf = 0; bphase ^= 1; }
        continue;
      }
      // fused ConvSta: each warp of a pair sums its own chunks; the pair's two partial sums meet in global memory
      // (atomicAdd on a zeroed map: two addends, so the result does not depend on their order)
      float sta0 = h ? 0.f : a.sta_b[0], sta1 = h ? 0.f : a.sta_b[1], sta2 = h ? 0.f : a.sta_b[2];
      for (int ch = h; ch < nstore; ch += 2) {
        if (k_dbg_nob & 16) break;                 // experiment: null epilogue (timing only)
        const int c0 = ch * 32;
        // accumulator column of channel c0: diag-cat keeps [main 16 | correction 16] per 16 channels, 2 columns per channel
        const uint32_t tcol = k_cat ? (uint32_t)(2 * c0) : (uint32_t)c0;
        const uint32_t coff = (uint32_t)a.acc_cols;
        uint32_t v[32];
        float x[32];
        st = st0 + ti * 4096;
        wres = wres_of(ti);
        if (a.has_res && nst == 3 && lane == 0) {
          // three tiles: the store of the previous chunk may still be draining, the one before it has been read - its
          // tile takes the residual of the NEXT chunk now, a whole chunk ahead of its use
          bulk_wait_read<1>();
          if (ch + 2 < nstore) issue_res(it, ch + 2, (ti + 1) % 3); else issue_res(it + 1, h, (ti + 1) % 3);
        }
        tmem_ld32(taddr + tcol, v);
        if (!a.has_res) {                         // the store that last used this tile must have finished reading it
          if (lane == 0 && !(k_dbg_nob & 2)) { if (nst > 1) bulk_wait_read<1>(); else bulk_wait_read<0>(); }   // (with a residual, issue_res waited already)
          __syncwarp();
        }
        tmem_ld_wait();
        if (k_cat) {
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]) + __uint_as_float(v[j + 16]);
          tmem_ld32(taddr + tcol + 32u, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) x[16 + j] = __uint_as_float(v[j]) + __uint_as_float(v[j + 16]);
        } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        if (k_corr) {
          tmem_ld32(taddr + tcol + coff, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] += __uint_as_float(v[j]);
        }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b = *reinterpret_cast<const float4*>(sbias + cbase + c0 + g * 4);
          x[g * 4] += b.x; x[g * 4 + 1] += b.y; x[g * 4 + 2] += b.z; x[g * 4 + 3] += b.w;
        }
        const int sw64 = (r >> 1) & 3;            // SWIZZLE_64B: 16-byte chunk index ^= address bits [7,9)
        if (a.has_res) {
          mbar_wait(wres, (rph >> ti) & 1u);
          rph ^= 1u << ti;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 h = *reinterpret_cast<const uint4*>(st + r * 64 + ((g ^ sw64) << 4));
            const __half* hh = reinterpret_cast<const __half*>(&h);
#pragma unroll
            for (int j = 0; j < 8; ++j) x[g * 8 + j] += __half2float(hh[j]);
          }
          if (a.has_res == 2) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 l = *reinterpret_cast<const uint4*>(st + 2048 + r * 64 + ((g ^ sw64) << 4));
              const __half* ll = reinterpret_cast<const __half*>(&l);
#pragma unroll
              for (int j = 0; j < 8; ++j) x[g * 8 + j] += __half2float(ll[j]);
            }
          }
          __syncwarp();                           // everyone has read the residual before it is overwritten
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
        }
        if (k_out_mode == 2) {                    // fp32 rows of 128 B, SWIZZLE_128B
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(st + r * 128 + ((g ^ (r & 7)) << 4)) =
                make_float4(x[g * 4], x[g * 4 + 1], x[g * 4 + 2], x[g * 4 + 3]);
        } else {
          __align__(16) __half hi[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) hi[j] = __float2half_rn(x[j]);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(st + r * 64 + ((g ^ sw64) << 4)) = reinterpret_cast<const uint4*>(hi)[g];
          if (k_out_mode == 1) {
            __align__(16) __half lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) lo[j] = __float2half_rn(x[j] - __half2float(hi[j]));
#pragma unroll
            for (int g = 0; g < 4; ++g)
              *reinterpret_cast<uint4*>(st + 2048 + r * 64 + ((g ^ sw64) << 4)) = reinterpret_cast<const uint4*>(lo)[g];
          }
          if (a.sta_out) {
            // the three ConvSta dot products on x (fp32; hi + lo carries 22 of its 24 bits, and the 3-class argmax
            // downstream is insensitive at that level - the standalone sta_kernel reads the planes instead)
            const float4* w4 = reinterpret_cast<const float4*>(sw.w) + ((cbase + c0) * 3) / 4;
#pragma unroll
            for (int g = 0; g < 8; ++g) {             // 4 channels x 3 classes = 3 float4 of the [ci][3] table
              const float4 wa = w4[3 * g], wb = w4[3 * g + 1], wc = w4[3 * g + 2];
              const float x0v = x[4 * g], x1v = x[4 * g + 1], x2v = x[4 * g + 2], x3v = x[4 * g + 3];
              sta0 = fmaf(x0v, wa.x, sta0); sta1 = fmaf(x0v, wa.y, sta1); sta2 = fmaf(x0v, wa.z, sta2);
              sta0 = fmaf(x1v, wa.w, sta0); sta1 = fmaf(x1v, wb.x, sta1); sta2 = fmaf(x1v, wb.y, sta2);
              sta0 = fmaf(x2v, wb.z, sta0); sta1 = fmaf(x2v, wb.w, sta1); sta2 = fmaf(x2v, wc.x, sta2);
              sta0 = fmaf(x3v, wc.y, sta0); sta1 = fmaf(x3v, wc.z, sta1); sta2 = fmaf(x3v, wc.w, sta2);
            }
          }
        }
        fence_proxy_async();                      // generic-proxy smem writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          if (!(k_dbg_nob & 4)) {
          tma_store_3d(&tmO_hi, st, cbase + c0, x0, y0);
          if (k_out_mode == 1) tma_store_3d(&tmO_lo, st + 2048, cbase + c0, x0, y0);
          }
          bulk_commit();
          if (a.has_res && nst != 3) {            // one tile: refill it with the residual of this warp's next chunk
            if (!(k_dbg_nob & 2)) bulk_wait_read<0>();
            if (ch + 2 < nstore) issue_res(it, ch + 2, 0); else issue_res(it + 1, h, 0);
          }
        }
        if (nst > 1 && ++ti == nst) ti = 0;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (k_cg2) mbar_arrive_cluster(&tempty[buf], 0); else mbar_arrive(&tempty[buf]); }
      if (++buf == a.nbuf) { buf = 0; bphase ^= 1; }
      if (a.sta_out) {                            // (nsplit == 1 here) this lane's pixel: TMEM lane q*32 + r of the tile
        const int py = y0 + r / a.tile_w, px = x0 + r % a.tile_w;
        if (py < a.Ho && px < a.Wo) {
          float* o = a.sta_out + ((size_t)py * a.Wo + px) * 3;
          atomicAdd(o, sta0); atomicAdd(o + 1, sta1); atomicAdd(o + 2, sta2);
        }
      }
    }
    if (lane == 0) bulk_wait_all();               // all output bytes are in global memory before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (k_mc > 1) cluster_sync_all();   // no CTA leaves while a peer may still arrive on its barriers
  if (warp == 2) { if (k_cg2) tmem_dealloc_cg2(tmem_base, (uint32_t)a.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols); }
}

// ------------------------------------------------------------------------------------ host side
int g_fuse_sta = 1;       // SFD2_FUSE_STA=0: run ConvSta as its own kernel (sta_kernel) instead of in rb2c3's epilogue
int g_tc_multicast = 1;   // SFD2_TC_MULTICAST=0 in the environment disables the 2-CTA weight multicast
int g_tc_slim = 1;        // SFD2_TC_SLIM=0: the 1x1 pair layers run the generic pair instantiation
int g_tc_stiles = 1;      // SFD2_TC_STILES=0: one staging tile per epilogue warp also in the CTA-pair layers
int g_tc_cg2 = 2;         // SFD2_TC_CG2: 0 = no CTA-pair MMAs, 1 = the 1x1 layers, 2 = every per-tap-ring layer that qualifies
int g_tc_pdl = 1;         // SFD2_TC_PDL=0: launch the conv layers without programmatic dependent launch

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
  return make_tmap(tm, base, rank, dims, strides_bytes, box, 0, 128);
}

// is_f32: element type fp32 instead of fp16; swizzle: 64 or 128 (bytes)
int make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int is_f32, int swizzle) {
  PFN_encodeTiled enc = get_encode_tiled();
  SFD2_CHECK(enc != nullptr, SFD2_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const CUresult r = enc(tm, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                         const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
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
  if (diag) {   // diag-cat slabs: row (tap*4 + kc)*128 + k*32 + j = w_hi row (tap, kc*64 + 16k + j) for j < 16, w_lo row (.., 16k + j - 16) else
    std::vector<__half> cat((size_t)taps * 4 * 128 * 64);
    for (int t = 0; t < taps; ++t)
      for (int kc = 0; kc < 4; ++kc)
        for (int r = 0; r < 128; ++r) {
          const int k = r >> 5, j = r & 31;
          const std::vector<__half>& src = j < 16 ? hi : lo;
          memcpy(&cat[(((size_t)t * 4 + kc) * 128 + r) * 64], &src[((size_t)t * 256 + kc * 64 + 16 * k + (j & 15)) * 64], 64 * sizeof(__half));
        }
    SFD2_CUDA(cudaMalloc(&L.w_cat, cat.size() * sizeof(__half)));
    SFD2_CUDA(cudaMemcpy(L.w_cat, cat.data(), cat.size() * sizeof(__half), cudaMemcpyHostToDevice));
    const uint64_t cdims[2] = {64, (uint64_t)taps * 4 * 128};
    const uint64_t cstr[1] = {128};
    const uint32_t cbox[2] = {64u, 128u}, chalf[2] = {64u, 64u};
    int rc = make_tmap_f16(&L.tm_w_cat, L.w_cat, 2, cdims, cstr, cbox);
    if (!rc) rc = make_tmap_f16(&L.tm_w_cat_half, L.w_cat, 2, cdims, cstr, chalf);
    if (rc) return rc;
  }
  const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  const uint64_t strides[1] = {(uint64_t)cols * 2};
  const uint32_t box[2] = {64u, (uint32_t)(diag ? 64 : L.cout_tc)};
  int rc = make_tmap_f16(&L.tm_w_hi, L.w_hi, 2, dims, strides, box);
  if (rc) return rc;
  rc = make_tmap_f16(&L.tm_w_lo, L.w_lo, 2, dims, strides, box);
  if (rc) return rc;
  // half-slab boxes for the 2-CTA multicast path (each CTA fetches n_mma/2 rows)
  const uint32_t hbox[2] = {64u, box[1] / 2};
  rc = make_tmap_f16(&L.tm_w_hi_half, L.w_hi, 2, dims, strides, hbox);
  if (rc) return rc;
  rc = make_tmap_f16(&L.tm_w_lo_half, L.w_lo, 2, dims, strides, hbox);
  if (rc) return rc;
  // quarter-slab boxes: channel-split layers (two passes of Cout/2) under multicast
  const uint32_t qbox[2] = {64u, box[1] / 4};
  rc = make_tmap_f16(&L.tm_w_hi_quarter, L.w_hi, 2, dims, strides, qbox);
  if (rc) return rc;
  return make_tmap_f16(&L.tm_w_lo_quarter, L.w_lo, 2, dims, strides, qbox);
}

// Activation tensor maps: s1   stride-1 view {C, W, H}, box {64,16,8}          (per-tap loads, 1x1 layers)
//                         s2   stride-2 view {C, 2, Wp/2, 2, Hp/2}, box {64,1,16,1,8}
//                         halo stride-1 view {C, W, H}, box {64,10,18}          (one load per chunk, 3x3 layers)
int tc_make_act_maps(const Act& t, const __half* base, CUtensorMap* s1, CUtensorMap* s2, CUtensorMap* halo, CUtensorMap* t8x16) {
  if (t8x16) {   // {64 ch, 8 px, 16 rows}: the A box of the 1x1 layers in split-ring mode
    const uint64_t dims[3] = {(uint64_t)t.C, (uint64_t)t.W, (uint64_t)t.H};
    const uint64_t str[2] = {(uint64_t)t.C * 2, (uint64_t)t.Wp * t.C * 2};
    const uint32_t box[3] = {64u, 8u, 16u};
    int rc = make_tmap_f16(t8x16, base, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)t.C, (uint64_t)t.W, (uint64_t)t.H};
    const uint64_t str[2] = {(uint64_t)t.C * 2, (uint64_t)t.Wp * t.C * 2};
    const uint32_t box[3] = {64u, (uint32_t)TC_HALO_W, (uint32_t)TC_HALO_H};
    int rc = make_tmap_f16(halo, base, 3, dims, str, box);
    if (rc) return rc;
  }
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

// Epilogue-side view of an activation plane: {C, W, H}, box {32 ch, 16 px, 2 rows}.  fp16 planes use the
// 64-byte swizzle (64-byte rows in the staging tile), fp32 head outputs the 128-byte swizzle.
// box_w x box_h = 16 x 2 (8 x 16 tiles) or 8 x 4 (16 x 8 "halo" tiles): one epilogue warp's 32 pixels
int tc_make_store_map(CUtensorMap* tm, const void* base, int C, int W, int H, int Wp, int is_f32, int box_w) {
  const uint64_t es = is_f32 ? 4 : 2;
  const uint64_t dims[3] = {(uint64_t)C, (uint64_t)W, (uint64_t)H};
  const uint64_t str[2] = {(uint64_t)C * es, (uint64_t)Wp * C * es};
  const uint32_t box[3] = {32u, (uint32_t)box_w, (uint32_t)(32 / box_w)};
  return make_tmap(tm, base, 3, dims, str, box, is_f32, is_f32 ? 128 : 64);
}

int g_tc_nsplit = 1;      // SFD2_TC_NSPLIT=0: keep wide layers in one channel pass (single-buffered accumulators)
int g_tc_split1x1 = 0;    // SFD2_TC_SPLIT1X1=1: 1x1 layers with split A / B rings (measured: c1 63 -> 68 us)
int g_tc_diagcat = 1;     // SFD2_TC_DIAGCAT=0: grouped layers in exact mode as three N=64 MMAs per K step (no [w_hi | w_lo] slabs)
int g_tc_halo = 1;        // SFD2_TC_HALO=0 falls back to per-tap A loads for the stride-1 3x3 layers

// out_f32_map: NULL for fp16 hi/lo outputs, else two maps {16x2 boxes, 8x4 boxes} of the fp32 output
int launch_conv_tc(const Act& in, const Layer& L, Act out, const Act* res, const CUtensorMap* out_f32_map, int split,
                   int num_sms, cudaStream_t st, int epi_fn, const Layer* sta, float* sta_out, int* rows_done, int rows_avail) {
  SFD2_CHECK(in.tm != nullptr && in.hi != nullptr, SFD2_ERR_ARG, "conv_tc(%s): input has no tensor maps", L.name.c_str());
  SFD2_CHECK(in.C == L.cin && in.C % 64 == 0, SFD2_ERR_ARG, "conv_tc(%s): cin %d", L.name.c_str(), in.C);
  SFD2_CHECK(split == 1 || split == 3, SFD2_ERR_ARG, "conv_tc: split must be 1 or 3");
  const bool diag = (L.groups == 32);
  TcConvArgs a{};
  a.Ho = out.H; a.Wo = out.W;
  const bool split1 = g_tc_split1x1 && L.k == 1 && L.stride == 1 && !diag;
  a.halo = ((g_tc_halo && L.k == 3 && L.stride == 1) || split1) ? 1 : 0;
  a.hw = split1 ? 8 : TC_HALO_W;
  a.hoff = split1 ? 0 : 1;
  a.a_plane_bytes = split1 ? 8 * 16 * 128 : TC_HALO_BYTES;
  a.a_plane_off = split1 ? 8 * 16 * 128 : TC_HALO_SLOT;
  a.tile_w = a.halo ? 8 : TC_TILE_W;
  a.tile_h = a.halo ? 16 : TC_TILE_H;
  a.epi_rows = a.halo ? 4 : 2;
  a.tiles_x = cdiv(out.W, a.tile_w);
  a.num_tiles = a.tiles_x * cdiv(out.H, a.tile_h);
  a.oob_tile = a.num_tiles;
  if (rows_done) {
    SFD2_CHECK(*rows_done % a.tile_h == 0 && !sta_out, SFD2_ERR_ARG, "conv_tc(%s): bad row band", L.name.c_str());
    const int tr0 = *rows_done / a.tile_h;
    const int tr1 = rows_avail >= out.H ? cdiv(out.H, a.tile_h) : rows_avail / a.tile_h;
    if (tr1 <= tr0) return SFD2_OK;
    a.tile_begin = tr0 * a.tiles_x;
    a.num_tiles = tr1 * a.tiles_x;
    *rows_done = tr1 * a.tile_h;
  }
  const int n_range = a.num_tiles - a.tile_begin;
  a.taps = L.k * L.k; a.stride = L.stride; a.diag = diag ? 1 : 0;
  a.kchunks = L.cin / 64;
  a.tap_rows = L.cout_tc;
  a.nsplit = 1;
  a.n_mma = diag ? 64 : L.cout_tc;
  a.acc_cols = L.cout_tc;
  // wide exact-mode 3x3 layers: two channel passes of Cout/2 so that (main + correction) x 2 buffers fit in TMEM
  // (measured: conv3a 177 -> 152 us, conv3b 295 -> 276 us; N = 192 -> 2 x 96 was slower, so only Cout = 256 is split)
  if (a.halo && !diag && split == 3 && L.k == 3 && L.cout_tc == 256 && L.cout == 256 && g_tc_nsplit) {
    a.nsplit = 2;
    a.n_mma = L.cout_tc / 2;
    a.acc_cols = a.n_mma;
  }
  // (block-diagonal grouped layers add mostly exact zeros, so their single accumulator is already accurate)
  a.corr = (split == 3 && L.k == 3 && !diag) ? 1 : 0;
  a.buf_stride = a.acc_cols * (1 + a.corr);
  if (a.halo && diag && split == 3 && g_tc_diagcat && L.w_cat) {   // see TcConvArgs::cat
    a.cat = 1; a.nsplit = 2; a.n_mma = 128; a.acc_cols = 128; a.corr = 1;
    a.buf_stride = 256;
  }
  a.ncat = (g_tc_diagcat && !a.halo && !diag && split == 3 && a.corr && a.nsplit == 1 && a.n_mma == 64 && a.acc_cols == 64) ? 1 : 0;
  // the epilogue reads 32-column chunks, so an 80-wide accumulator (headP) is over-read by 16 columns:
  // keep that inside the allocation
  const int pass_ch = (a.nsplit > 1) ? a.n_mma : round_up(L.cout, 32);   // channels the epilogue reads per pass
  int over = a.corr * a.acc_cols + pass_ch - a.buf_stride;
  if (over < 0) over = 0;
  SFD2_CHECK(a.buf_stride + over <= 512, SFD2_ERR_ARG, "conv_tc(%s): accumulator too wide", L.name.c_str());
  a.nbuf = (2 * a.buf_stride + over <= 512) ? 2 : 1;
  int tc = 32;
  while (tc < a.nbuf * a.buf_stride + over) tc <<= 1;
  a.tmem_cols = tc;
  a.cout = L.cout; a.relu = L.relu; a.split = split;
  // CTA pair (see TcConvArgs::cg2): per-tap-ring layers whose slab halves are whole swizzle atoms, enough tiles for a cluster
  const bool can_mc = g_tc_multicast && n_range >= 2 && (a.n_mma / 2) % 8 == 0;
  a.cg2 = (g_tc_cg2 && split == 3 && can_mc && (!a.halo || g_tc_cg2 > 2) && !diag && !a.ncat && a.n_mma % 32 == 0 && (g_tc_cg2 > 1 || L.k == 1)) ? 1 : 0;
  a.b_bytes = (a.cg2 ? a.n_mma / 2 : a.n_mma) * 128;
  a.stage_bytes = (TC_A_BYTES + a.b_bytes) * (split == 3 ? 2 : 1);
  const int smem_max = 227 * 1024;
  const bool fuse_sta = sta && sta_out;
  SFD2_CHECK(!fuse_sta || (!out_f32_map && L.cout == 256 && sta->cin == 256 && sta->cout == 3 && sta->k == 1 && sta->w.size() == 768),
             SFD2_ERR_ARG, "conv_tc(%s): ConvSta can only be fused into a 256-channel fp16-plane layer", L.name.c_str());
  // alignment slack, staging, bias, barriers
  // CTA-pair 1x1 layers with a residual are bound by their epilogue, not by the operand ring: two 64 KB stages, and the shared
  // memory the halved weight slabs free goes to extra staging tiles (TcConvArgs::stiles)
  // (measured with the slim instantiation: without a residual, three operand stages + one staging tile beat two + two,
  //  rb.c1 49-50 us against 56-57; with a residual two stages + three staging tiles win, rb.c3 70-72 against 78-85)
  a.stiles = (a.cg2 && L.k == 1 && !out_f32_map && g_tc_stiles && res) ? 3 : 1;
  const int smem_fixed = 1024 + a.stiles * TC_STAGING_BYTES + TC_BIAS_BYTES + TC_BAR_BYTES + ((out_f32_map && epi_fn) ? TC_EXCH_BYTES : 0);
  int stages = (smem_max - smem_fixed) / a.stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  SFD2_CHECK(stages >= 2, SFD2_ERR_ARG, "conv_tc(%s): stage too large", L.name.c_str());
  a.stages = stages;
  a.ring_bytes = stages * a.stage_bytes;
  if (a.halo) {
    a.a_slot_bytes = (split == 3 ? 2 : 1) * a.a_plane_off;
    a.a_slots = 2;
    if (split1)   // a whole tile of A in flight if two weight slabs still fit beside it
      while (a.a_slots < 4 && a.a_slots < a.kchunks && (a.a_slots + 1) * a.a_slot_bytes + 2 * a.b_bytes <= smem_max - smem_fixed) ++a.a_slots;
    int bs = (smem_max - smem_fixed - a.a_slots * a.a_slot_bytes) / a.b_bytes;
    const int max_bs = getenv("SFD2_TC_BSTAGES") ? atoi(getenv("SFD2_TC_BSTAGES")) : TC_MAX_STAGES;
    if (bs > max_bs) bs = max_bs;
    if (bs > TC_MAX_STAGES) bs = TC_MAX_STAGES;
    SFD2_CHECK(bs >= 2, SFD2_ERR_ARG, "conv_tc(%s): weight ring does not fit", L.name.c_str());
    a.b_stages = bs;
    a.ring_bytes = a.a_slots * a.a_slot_bytes + bs * a.b_bytes;
  }
  a.bias = L.b_dev;
  a.out_mode = out_f32_map ? 2 : (split == 3 ? 1 : 0);
  a.epi_fn = out_f32_map ? epi_fn : 0;
  a.has_res = res ? (split == 3 ? 2 : 1) : 0;
  a.sta_out = fuse_sta ? sta_out : nullptr;
  static const TcStaW kNoSta{};
  TcStaW* swp = nullptr;
  TcStaW sw_local;
  if (fuse_sta) {          // OIHW [3][256][1][1] -> [ci][3]
    for (int ci = 0; ci < 256; ++ci)
      for (int c = 0; c < 3; ++c) sw_local.w[ci * 3 + c] = sta->w[(size_t)c * 256 + ci];
    swp = &sw_local;
  }
  const TcStaW& sw = swp ? *swp : kNoSta;
  for (int i = 0; i < 3; ++i) a.sta_b[i] = fuse_sta ? sta->b[i] : 0.f;
  SFD2_CHECK(!fuse_sta || a.nsplit == 1, SFD2_ERR_ARG, "conv_tc(%s): fused ConvSta needs a single channel pass", L.name.c_str());
  SFD2_CHECK(out_f32_map || out.tm_st, SFD2_ERR_ARG, "conv_tc(%s): output has no store maps", L.name.c_str());
  SFD2_CHECK(!res || res->tm_st, SFD2_ERR_ARG, "conv_tc(%s): residual has no store maps", L.name.c_str());
  SFD2_CHECK(L.cout <= 256, SFD2_ERR_ARG, "conv_tc(%s): cout > 256", L.name.c_str());
  const int so = a.halo ? 2 : 0;   // store maps: [hi, lo] with 16x2 boxes, then [hi, lo] with 8x4 boxes
  const CUtensorMap& o_hi = out_f32_map ? out_f32_map[a.halo] : out.tm_st[so];
  const CUtensorMap& o_lo = out_f32_map ? out_f32_map[a.halo] : out.tm_st[so + 1];
  const CUtensorMap& r_hi = res ? res->tm_st[so] : o_hi;       // residual boxes have the epilogue warps' pixel shape
  const CUtensorMap& r_lo = res ? res->tm_st[so + 1] : o_lo;
  const size_t smem = (size_t)a.ring_bytes + smem_fixed;
  const CUtensorMap* tmA = in.tm + (a.halo ? (split1 ? 6 : 4) : (L.stride == 2 ? 2 : 0));
  // multicast needs an even number of 1024-byte-aligned half slabs and at least one full cluster of work
  a.mc = can_mc ? 2 : 1;
  {
    static const int nob = getenv("SFD2_TC_DEBUG_NOB") ? atoi(getenv("SFD2_TC_DEBUG_NOB")) : 0;
    a.dbg_nob = (nob == 1 && a.cat) || (nob == 2 && a.halo) ? 1 : 0;     // 1: grouped layers only, 2: every halo-mode layer
    if (a.dbg_nob) a.mc = 1;
    if (nob == 4) a.dbg_nob = 2;
    if (nob == 8) a.dbg_nob = 4;                              // 8: epilogue computes but issues no output stores (timing only)
    if (nob == 16) a.dbg_nob = 8;
    if (nob == 32) a.dbg_nob = 16 | 8;                        // 32: null epilogue - accumulators are handed back untouched (timing only)                             // 16: residual tiles are not loaded (timing only)                              // 4: epilogue does not wait for its TMA stores to drain (RACY, timing only)
  }
  const bool slim = g_tc_slim && a.cg2 && a.mc == 2 && !a.halo && L.k == 1 && L.stride == 1 && split == 3 && !diag && !a.cat && !a.ncat &&
                    a.nsplit == 1 && !a.corr && a.epi_fn == 0 && a.out_mode == 1 && a.dbg_nob == 0;
  auto kern = slim ? tc_conv_kernel<true, true> : (a.cg2 ? tc_conv_kernel<true, false> : tc_conv_kernel<false, false>);
  SFD2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = n_range < num_sms ? n_range : num_sms;
  if (a.mc > 1) grid = (grid / 2) * 2 > 0 ? ((grid + 1) / 2) * 2 : 2;
  if (a.mc > 1 && grid > num_sms) grid -= 2;
  a.iters = cdiv(cdiv(n_range, a.mc) * a.nsplit, grid / a.mc);   // cluster-units per cluster (see `unit` in the kernel)
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)a.mc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_tc_pdl ? 2 : 1;
  // weight-slab boxes: full slab = n_mma rows; with multicast each CTA fetches n_mma/2 rows
  const int box_rows = (a.mc > 1) ? a.n_mma / 2 : a.n_mma;
  const int full_rows = a.cat ? 128 : (diag ? 64 : L.cout_tc);
  const int mi = (box_rows == full_rows) ? 0 : (box_rows * 2 == full_rows ? 1 : 2);
  SFD2_CHECK(box_rows == full_rows || box_rows * 2 == full_rows || box_rows * 4 == full_rows, SFD2_ERR_ARG,
             "conv_tc(%s): no weight map for %d-row boxes", L.name.c_str(), box_rows);
  const CUtensorMap& wb_hi = a.cat ? (mi == 0 ? L.tm_w_cat : L.tm_w_cat_half)
                                   : (mi == 0 ? L.tm_w_hi : (mi == 1 ? L.tm_w_hi_half : L.tm_w_hi_quarter));
  const CUtensorMap& wb_lo = mi == 0 ? L.tm_w_lo : (mi == 1 ? L.tm_w_lo_half : L.tm_w_lo_quarter);
  if (getenv("SFD2_DEBUG_LAUNCH"))
    fprintf(stderr, "conv_tc %-8s out %dx%d halo %d tiles %d grid %d mc %d cg2 %d stiles %d stages %d/%d smem %zu iters %d nsplit %d\n", L.name.c_str(), out.H, out.W,
            a.halo, n_range, grid, a.mc, a.cg2, a.stiles, a.stages, a.b_stages, smem, a.iters, a.nsplit);
  {
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA[0], tmA[1], wb_hi, wb_lo, o_hi, o_lo, r_hi, r_lo, a, sw);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("conv_tc(%s): launch failed: %s (grid %d, cluster %d, pair %d, %zu B smem, %d tiles)", L.name.c_str(), cudaGetErrorString(e),
                grid, a.mc, a.cg2, smem, n_range);
      return SFD2_ERR_CUDA;
    }
  }
  ++g_launches;
  SFD2_CUDA(cudaGetLastError());
  return SFD2_OK;
}

}  // namespace sfd2
