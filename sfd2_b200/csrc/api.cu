// C ABI of libsfd2_b200.so (see include/sfd2_b200.h): context, weight blob parsing, workspace
// management and the per-image kernel schedule of the extract path, plus the matcher entry points.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "common.cuh"
#include "tc_match.cuh"

namespace sfd2 {

static thread_local char g_err[1024] = "";
thread_local long long g_launches = 0;
int g_band_layers = 2;    // SFD2_BAND_LAYERS: how many of conv1b / conv2a / conv2b follow conv1a band by band (0..3)
int g_post_pdl = 1;       // SFD2_POST_PDL=0: heat / nms / select / descriptor kernels launched without programmatic stream serialization
int g_host_bands = 4;     // SFD2_HOST_BANDS: row bands of a single host image's upload (<= 1: one copy, conv1a after it)

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace sfd2

using namespace sfd2;

// ---- weight blob (written by sfd2_b200/weights.py) -------------------------------------------
//   char magic[8] = "SFD2W001"; uint32 n_layers; uint32 reserved;
//   n_layers x { char name[16]; int32 cin, cout, k, stride, groups, relu; uint64 w_off, b_off; }
//   float data (offsets in bytes from the start of the blob; weights OIHW, BN already folded)
#pragma pack(push, 1)
struct BlobLayer {
  char name[16];
  int32_t cin, cout, k, stride, groups, relu;
  uint64_t w_off, b_off;
};
#pragma pack(pop)

enum ActId { A1A, A1B, A2A, A2B, A3A, A3B, T1, T2, BA, BB, PA, DA, NUM_ACTS };

// per-image extract workspace (activations, TMA views, head buffers, candidate list).  Buffers are grow-only:
// every buffer keeps its own byte capacity, a new image size only re-allocates the buffers it outgrows and
// re-encodes the tensor maps (host-side, no device sync), so a stream of mixed portrait / landscape / multi-scale
// sizes settles after the largest one has been seen instead of freeing and re-allocating ~40 buffers per call.
struct Ws {
  int wsH = 0, wsW = 0;          // shape the dims / tensor maps below are encoded for
  bool maps_tc = false;          // tensor maps valid for (wsH, wsW)
  bool zero_f32 = false, zero_tc = false;   // padded planes must be re-zeroed on the next use (shape changed)
  Act acts[NUM_ACTS];
  size_t cap_f32[NUM_ACTS] = {}, cap_hi[NUM_ACTS] = {}, cap_lo[NUM_ACTS] = {};
  CUtensorMap maps[NUM_ACTS][8];
  CUtensorMap st_maps[NUM_ACTS][4];
  CUtensorMap map_1a[4];                    // conv1a output rows: [hi, lo] box {64 ch, 256 px, 1 row}, [hi, lo] box {64, 128, 1}
  CUtensorMap map_logits[2], map_desc[2], map_semi[2];   // fp32 head outputs (TMA store views: 16x2 and 8x4 boxes)
  int H2 = 0, W2 = 0, H4 = 0, W4 = 0, H8 = 0, W8 = 0;
  float4* nimg = nullptr;  // normalised image, NHWC4 fp32
  float *logits = nullptr, *semi = nullptr, *descmap = nullptr, *sta = nullptr, *heat = nullptr, *nmsdbg = nullptr;
  float* drows = nullptr; size_t cap_drows = 0;     // [4 * topk][128] tap descriptors of the sparse descriptor head
  unsigned long long *cand = nullptr, *scratch = nullptr;
  size_t cap_nimg = 0, cap_logits = 0, cap_semi = 0, cap_descmap = 0, cap_sta = 0, cap_heat = 0, cap_nmsdbg = 0,
         cap_cand = 0, cap_scratch = 0;
  int cap = 0;
  int *counter = nullptr, *status = nullptr;
};

struct sfd2_ctx {
  int device = 0, num_sms = 148;
  std::vector<Layer> layers;
  std::map<std::string, int> lidx;
  cudaStream_t stream = nullptr;  // used by the *_host entry points
  static constexpr int kMaxStreams = 4;
  Ws ws[kMaxStreams];             // per-image workspaces: consecutive images of a batch rotate through nstreams of them
  cudaStream_t aux[kMaxStreams] = {};   // internal streams so one image's kernel tails overlap the others'
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxStreams] = {};
  int nstreams = 2;
  cudaStream_t copy_stream = nullptr;          // H2D of batched host inputs
  std::vector<cudaEvent_t> img_ready;
  int debug_flags = 0;
  int cand_cap_override = 0;     // SFD2_CAND_CAP (test hook): cap the NMS candidate list to exercise the overflow report
  int last_prec = -1;
  // host-API staging
  void* img_dev = nullptr; size_t img_cap = 0;
  float *kp_dev = nullptr, *sc_dev = nullptr, *de_dev = nullptr; int32_t* cnt_dev = nullptr; size_t out_cap = 0;
  // matcher workspace, CUDA-core fp32 mode
  unsigned long long *row_key = nullptr, *col_key = nullptr; size_t key_cap = 0;
  unsigned *row2 = nullptr, *col2 = nullptr;   // second-best similarities (ratio tests)
  // matcher workspace, grouped tcgen05 path (grow-only; tensor maps cover the whole plane allocation)
  struct MatchWs {
    __half *hi = nullptr, *lo = nullptr; size_t cap_hi = 0, cap_lo = 0;
    CUtensorMap tm_hi, tm_lo; bool maps_ok = false;
    unsigned long long* keys = nullptr; size_t cap_keys = 0;
    unsigned* sec = nullptr; size_t cap_sec = 0;
    int *remap = nullptr, *efflen = nullptr, *done = nullptr; size_t cap_remap = 0, cap_efflen = 0, cap_done = 0;
    uint8_t* tab = nullptr; size_t cap_tab = 0;
    static constexpr int kRing = 32;
    struct Slot { void* host = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; } ring[kRing];
    unsigned next = 0;
  } mws;
  float *m_d0 = nullptr, *m_d1 = nullptr; size_t m_d0_cap = 0, m_d1_cap = 0;
  int32_t* m_out = nullptr; float* m_sim = nullptr; size_t m_out_cap = 0;
  long long launches = 0;
  // optional per-launch CUDA-event timing (sfd2_profile / sfd2_profile_read)
  bool prof_on = false;
  struct ProfRec { std::string label; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;

  const Layer& L(const char* n) const { return layers[lidx.at(n)]; }
};

static const char* kLayerNames[] = {"conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "rb0c1", "rb0c2",
                                    "rb0c3", "rb1c1", "rb1c2", "rb1c3", "rb2c1", "rb2c2", "rb2c3", "convPa0",
                                    "headP", "convDa0", "headD", "sta"};

static int upload_simt(Layer& L) {
  const int taps = L.k * L.k, cpg = L.cin / L.groups;
  const int cp = round_up(L.cout, 64);
  L.cout_pad = cp;
  std::vector<float> w((size_t)taps * cpg * cp, 0.f), b(cp < 128 ? 128 : cp, 0.f);
  for (int o = 0; o < L.cout; ++o)
    for (int r = 0; r < cpg; ++r)
      for (int t = 0; t < taps; ++t) w[((size_t)t * cpg + r) * cp + o] = L.w[((size_t)o * cpg + r) * taps + t];
  for (int o = 0; o < L.cout; ++o) b[o] = L.b[o];
  SFD2_CUDA(cudaMalloc(&L.w_simt, w.size() * sizeof(float)));
  SFD2_CUDA(cudaMalloc(&L.b_dev, b.size() * sizeof(float)));
  SFD2_CUDA(cudaMemcpy(L.w_simt, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  SFD2_CUDA(cudaMemcpy(L.b_dev, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
  return SFD2_OK;
}

static void free_layer(Layer& L) {
  cudaFree(L.w_simt); cudaFree(L.b_dev); cudaFree(L.w_hi); cudaFree(L.w_lo); cudaFree(L.w_cat);
  L.w_simt = L.b_dev = nullptr; L.w_hi = L.w_lo = L.w_cat = nullptr;
}

static void free_workspace(Ws& w) {
  for (int i = 0; i < NUM_ACTS; ++i) {
    cudaFree(w.acts[i].f32); cudaFree(w.acts[i].hi); cudaFree(w.acts[i].lo);
    w.acts[i] = Act();
    w.cap_f32[i] = w.cap_hi[i] = w.cap_lo[i] = 0;
  }
  cudaFree(w.nimg); w.nimg = nullptr;
  cudaFree(w.logits); cudaFree(w.semi); cudaFree(w.descmap); cudaFree(w.sta); cudaFree(w.heat); cudaFree(w.nmsdbg);
  cudaFree(w.cand); cudaFree(w.scratch); cudaFree(w.counter); cudaFree(w.status); cudaFree(w.drows);
  w.drows = nullptr; w.cap_drows = 0;
  w.logits = w.semi = w.descmap = w.sta = w.heat = w.nmsdbg = nullptr;
  w.cand = w.scratch = nullptr; w.counter = w.status = nullptr;
  w.cap_nimg = w.cap_logits = w.cap_semi = w.cap_descmap = w.cap_sta = w.cap_heat = w.cap_nmsdbg = w.cap_cand = w.cap_scratch = 0;
  w.wsH = w.wsW = 0; w.maps_tc = false; w.zero_f32 = w.zero_tc = false;
}

// grow-only device buffer: re-allocates only when `bytes` exceeds the capacity (contents are then undefined)
template <typename T>
static int reserve(T*& p, size_t& cap, size_t bytes, bool* grew = nullptr) {
  if (grew) *grew = false;
  if (bytes <= cap && p) return SFD2_OK;
  if (p) { cudaFree(p); p = nullptr; cap = 0; }
  SFD2_CUDA(cudaMalloc(&p, bytes));
  cap = bytes;
  if (grew) *grew = true;
  return SFD2_OK;
}

static int ensure_workspace(const sfd2_ctx* c, Ws& w, int H, int W, int prec, int topk) {
  int rc;
  if (g_sparse_desc && (prec == SFD2_PREC_TC_MIXED || prec == SFD2_PREC_TC_FAST) &&
      (rc = reserve(w.drows, w.cap_drows, (size_t)4 * topk * SFD2_DESC_DIM * sizeof(float))))
    return rc;
  const bool want_tc = (prec != SFD2_PREC_FP32);
  if (w.wsH != H || w.wsW != W) {
    w.wsH = H; w.wsW = W;
    w.maps_tc = false;
    w.H2 = conv_out(H, 2); w.W2 = conv_out(W, 2);
    w.H4 = conv_out(w.H2, 2); w.W4 = conv_out(w.W2, 2);
    w.H8 = conv_out(w.H4, 2); w.W8 = conv_out(w.W4, 2);
    const int dims[NUM_ACTS][3] = {{H, W, 64}, {w.H2, w.W2, 64}, {w.H2, w.W2, 128}, {w.H4, w.W4, 128},
                                   {w.H4, w.W4, 256}, {w.H4, w.W4, 256}, {w.H4, w.W4, 256}, {w.H4, w.W4, 256},
                                   {w.H4, w.W4, 256}, {w.H4, w.W4, 256}, {w.H8, w.W8, c->L("convPa0").cout}, {w.H4, w.W4, c->L("convDa0").cout}};
    for (int i = 0; i < NUM_ACTS; ++i) {
      Act& a = w.acts[i];
      a.H = dims[i][0]; a.W = dims[i][1]; a.C = dims[i][2];
      a.Hp = round_up(a.H, 2); a.Wp = round_up(a.W, 2);
    }
    // planes that exist already hold another shape's data where this shape's zero padding must be
    w.zero_f32 = w.zero_tc = true;
    const size_t n8 = (size_t)w.H8 * w.W8, n4 = (size_t)w.H4 * w.W4;
    if ((rc = reserve(w.nimg, w.cap_nimg, (size_t)H * W * sizeof(float4)))) return rc;
    if ((rc = reserve(w.logits, w.cap_logits, n8 * 80 * sizeof(float)))) return rc;
    if ((rc = reserve(w.semi, w.cap_semi, n8 * 64 * sizeof(float)))) return rc;
    if ((rc = reserve(w.descmap, w.cap_descmap, n4 * 128 * sizeof(float)))) return rc;
    if ((rc = reserve(w.sta, w.cap_sta, n4 * 3 * sizeof(float)))) return rc;
    if ((rc = reserve(w.heat, w.cap_heat, (size_t)H * W * sizeof(float)))) return rc;
    if ((rc = reserve(w.nmsdbg, w.cap_nmsdbg, (size_t)H * W * sizeof(float)))) return rc;
    // NMS survivors are >= 5 px apart except on exact plateaus: H*W/16 leaves 1.5x headroom.
    w.cap = (int)(((size_t)H * W) / 16) + 4096;
    if (c->cand_cap_override > 0 && c->cand_cap_override < w.cap) w.cap = c->cand_cap_override;
    int cap2 = 1;
    while (cap2 < w.cap) cap2 <<= 1;
    if ((rc = reserve(w.cand, w.cap_cand, (size_t)w.cap * sizeof(unsigned long long)))) return rc;
    bool grew = false;
    if ((rc = reserve(w.scratch, w.cap_scratch, (size_t)cap2 * sizeof(unsigned long long), &grew))) return rc;
    // select_kernel's rank / arrival counters: zero once, the kernel leaves them zero
    if (grew) SFD2_CUDA(cudaMemset(w.scratch, 0, w.cap_scratch));
    if (!w.counter) SFD2_CUDA(cudaMalloc(&w.counter, sizeof(int)));
    if (!w.status) {
      SFD2_CUDA(cudaMalloc(&w.status, sizeof(int)));
      SFD2_CUDA(cudaMemset(w.status, 0, sizeof(int)));
    }
  }
  if (!want_tc) {
    for (int i = 0; i < NUM_ACTS; ++i) {
      Act& a = w.acts[i];
      if ((rc = reserve(a.f32, w.cap_f32[i], a.elems() * sizeof(float)))) return rc;
    }
  } else {
    for (int i = 0; i < NUM_ACTS; ++i) {
      Act& a = w.acts[i];
      bool g0 = false, g1 = false;
      if ((rc = reserve(a.hi, w.cap_hi[i], a.elems() * sizeof(__half), &g0))) return rc;
      if ((rc = reserve(a.lo, w.cap_lo[i], a.elems() * sizeof(__half), &g1))) return rc;
      if (g0 || g1) w.maps_tc = false;
    }
  }
  if (want_tc && !w.maps_tc) {
    for (int i = 0; i < NUM_ACTS; ++i) {
      Act& a = w.acts[i];
      rc = tc_make_act_maps(a, a.hi, &w.maps[i][0], &w.maps[i][2], &w.maps[i][4], &w.maps[i][6]);
      if (rc) return rc;
      rc = tc_make_act_maps(a, a.lo, &w.maps[i][1], &w.maps[i][3], &w.maps[i][5], &w.maps[i][7]);
      if (rc) return rc;
      a.tm = w.maps[i];
      for (int b = 0; b < 2 && !rc; ++b) {
        rc = tc_make_store_map(&w.st_maps[i][2 * b], a.hi, a.C, a.W, a.H, a.Wp, 0, b ? 8 : 16);
        if (!rc) rc = tc_make_store_map(&w.st_maps[i][2 * b + 1], a.lo, a.C, a.W, a.H, a.Wp, 0, b ? 8 : 16);
      }
      if (rc) return rc;
      a.tm_st = w.st_maps[i];
    }
    rc = 0;
    for (int pl = 0; pl < 4 && !rc; ++pl) {
      const Act& a = w.acts[A1A];
      const uint64_t dims[3] = {64, (uint64_t)a.W, (uint64_t)a.H};
      const uint64_t str[2] = {128, (uint64_t)a.Wp * 128};
      const uint32_t box[3] = {64u, pl < 2 ? 256u : 128u, 1u};
      rc = make_tmap(&w.map_1a[pl], (pl & 1) ? (const void*)a.lo : (const void*)a.hi, 3, dims, str, box, 0, 128);
    }
    for (int b = 0; b < 2 && !rc; ++b) {
      rc = tc_make_store_map(&w.map_logits[b], w.logits, 80, w.W8, w.H8, w.W8, 1, b ? 8 : 16);
      if (!rc) rc = tc_make_store_map(&w.map_desc[b], w.descmap, 128, w.W4, w.H4, w.W4, 1, b ? 8 : 16);
      if (!rc) rc = tc_make_store_map(&w.map_semi[b], w.semi, 64, w.W8, w.H8, w.W8, 1, b ? 8 : 16);
    }
    if (rc) return rc;
    w.maps_tc = true;
  }
  return SFD2_OK;
}

// The even-padded row / column of an activation (Hp > H or Wp > W) is read by stride-2 consumers as the conv's zero
// padding and is never written by a producer (TMA stores clip at W x H): zero the padded planes once per shape, on
// the stream that is about to use the workspace.
static int zero_padding(Ws& w, bool tc, cudaStream_t st) {
  bool& flag = tc ? w.zero_tc : w.zero_f32;
  if (!flag) return SFD2_OK;
  for (int i = 0; i < NUM_ACTS; ++i) {
    Act& a = w.acts[i];
    if (a.Hp == a.H && a.Wp == a.W) continue;
    if (tc) {
      SFD2_CUDA(cudaMemsetAsync(a.hi, 0, a.elems() * sizeof(__half), st));
      SFD2_CUDA(cudaMemsetAsync(a.lo, 0, a.elems() * sizeof(__half), st));
    } else {
      SFD2_CUDA(cudaMemsetAsync(a.f32, 0, a.elems() * sizeof(float), st));
    }
  }
  flag = false;
  return SFD2_OK;
}

static cudaEvent_t prof_event(sfd2_ctx* c) {
  if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
static inline void prof_begin(sfd2_ctx* c, const char* label, cudaStream_t st) {
  if (!c->prof_on) return;
  sfd2_ctx::ProfRec r{label, prof_event(c), prof_event(c)};
  cudaEventRecord(r.a, st);
  c->prof.push_back(r);
}
static inline void prof_end(sfd2_ctx* c, cudaStream_t st) {
  if (!c->prof_on) return;
  cudaEventRecord(c->prof.back().b, st);
}

// A single host image arrives in row bands (sfd2_extract_host, n == 1): ev[b] fires on the copy stream when rows
// [.., row_end[b]) are in HBM, and conv1a runs band by band behind the copy instead of after all of it.
struct Bands {
  int n = 0;
  int row_end[8];
  cudaEvent_t ev[8];
};

// one image through network + post-processing, all on `st`
static int extract_one(sfd2_ctx* c, Ws& w, const void* img, int img_dtype, int H, int W, const sfd2_extract_params* p,
                       float* kpts, float* scores, float* desc, int32_t* count, cudaStream_t st, const Bands* bands = nullptr) {
  const int prec = p->precision;
  const bool tc = prec != SFD2_PREC_FP32;
  const int split = (prec == SFD2_PREC_TC_EXACT || prec == SFD2_PREC_TC_MIXED) ? 3 : 1;
  // MIXED: the descriptor head only has to meet the 1e-3 tolerance, so it runs single-pass (hi planes only)
  const int split_d = (prec == SFD2_PREC_TC_MIXED) ? 1 : split;
  Act* A = w.acts;
  int rc = zero_padding(w, tc, st);
  if (rc) return rc;
#define RUN(x) do { rc = (x); if (rc) return rc; } while (0)
#define RUNP(label, x) do { prof_begin(c, label, st); rc = (x); prof_end(c, st); if (rc) return rc; } while (0)
  // the first layers can run band by band behind a banded upload: {layer, input, output, stride, output rows done}
  struct BandLayer { const char* name; int in, out, stride, done; };
  BandLayer bl[3] = {{"conv1b", A1A, A1B, 2, 0}, {"conv2a", A1B, A2A, 1, 0}, {"conv2b", A2A, A2B, 2, 0}};
  int nbl = 0;               // layers of bl[] completed inside the band loop
  if (bands && bands->n > 0) {
    const bool by_band = conv1a_bands_ok(tc ? split : 0);
    if (by_band) nbl = std::max(0, std::min(g_band_layers, 3));
    int y0 = 0;
    for (int b = 0; b < bands->n; ++b) {
      if (cudaStreamWaitEvent(st, bands->ev[b], 0) != cudaSuccess) { set_error("cudaStreamWaitEvent(band %d) failed", b); return SFD2_ERR_CUDA; }
      // output row y reads input rows y-1 .. y+1: with rows < row_end[b] uploaded, rows < row_end[b] - 1 can be computed
      const int y1 = (b == bands->n - 1) ? H : bands->row_end[b] - 1;
      if (by_band && y1 > y0) {
        RUNP("conv1a", launch_conv1a(img, img_dtype, H, W, c->L("conv1a"), A[A1A], split, w.nimg, w.map_1a, c->num_sms, st, y0, y1));
        y0 = y1;
        // the layers behind it follow as far as their inputs reach: a 3x3 output row y reads input rows s*y-1 .. s*y+1
        int in_done = y1, in_h = H;
        for (int l = 0; l < nbl; ++l) {
          const int out_h = A[bl[l].out].H;
          const int avail = in_done >= in_h ? out_h : (bl[l].stride == 2 ? (in_done >= 2 ? (in_done - 2) / 2 + 1 : 0) : std::max(in_done - 1, 0));
          prof_begin(c, (std::string("tc_conv:") + bl[l].name).c_str(), st);
          rc = launch_conv_tc(A[bl[l].in], c->L(bl[l].name), A[bl[l].out], nullptr, nullptr, split, c->num_sms, st, 0, nullptr, nullptr,
                              &bl[l].done, avail);
          prof_end(c, st);
          if (rc) return rc;
          in_done = bl[l].done; in_h = out_h;
        }
      }
    }
    if (!by_band) RUNP("conv1a", launch_conv1a(img, img_dtype, H, W, c->L("conv1a"), A[A1A], tc ? split : 0, w.nimg, tc ? w.map_1a : nullptr, c->num_sms, st));
  } else {
    RUNP("conv1a", launch_conv1a(img, img_dtype, H, W, c->L("conv1a"), A[A1A], tc ? split : 0, w.nimg, tc ? w.map_1a : nullptr, c->num_sms, st));
  }
  // ConvSta rides in the epilogue of the layer that produces out4 (tcgen05 modes)
  const bool fuse_sta = tc && p->use_stability && g_fuse_sta;
  // single-pass descriptor head (`mixed`, `fast`): evaluate it after selection, only where descriptors are sampled
  const bool sparse_d = tc && split_d == 1 && g_sparse_desc && w.drows != nullptr;
  auto conv = [&](const char* name, int in, int out, int res, int sp = 0, bool with_sta = false) -> int {
    const Layer& L = c->L(name);
    // the fused ConvSta accumulates with atomicAdd (two epilogue warps per pixel): start from zero
    if (with_sta && cudaMemsetAsync(w.sta, 0, (size_t)w.H4 * w.W4 * 3 * sizeof(float), st) != cudaSuccess) {
      set_error("cudaMemsetAsync(sta) failed");
      return SFD2_ERR_CUDA;
    }
    prof_begin(c, (std::string(tc ? "tc_conv:" : "conv_f32:") + name).c_str(), st);
    const int r = tc ? launch_conv_tc(A[in], L, A[out], res >= 0 ? &A[res] : nullptr, nullptr, sp ? sp : split, c->num_sms, st,
                                      0, with_sta ? &c->L("sta") : nullptr, with_sta ? w.sta : nullptr)
                     : launch_conv_simt(A[in], L, A[out], res >= 0 ? &A[res] : nullptr, st);
    prof_end(c, st);
    return r;
  };
  if (nbl < 1) RUN(conv("conv1b", A1A, A1B, -1));
  if (nbl < 2) RUN(conv("conv2a", A1B, A2A, -1));
  if (nbl < 3) RUN(conv("conv2b", A2A, A2B, -1));
  RUN(conv("conv3a", A2B, A3A, -1));
  RUN(conv("conv3b", A3A, A3B, -1));
  RUN(conv("rb0c1", A3B, T1, -1)); RUN(conv("rb0c2", T1, T2, -1)); RUN(conv("rb0c3", T2, BA, A3B));
  RUN(conv("rb1c1", BA, T1, -1));  RUN(conv("rb1c2", T1, T2, -1)); RUN(conv("rb1c3", T2, BB, BA));
  RUN(conv("rb2c1", BB, T1, -1));  RUN(conv("rb2c2", T1, T2, -1)); RUN(conv("rb2c3", T2, BA, BB, 0, fuse_sta));
  RUN(conv("convPa0", BA, PA, -1));
  RUN(conv("convDa0", BA, DA, -1, split_d));
  // heads: fp32 outputs
  Act logit_act; logit_act.f32 = w.logits; logit_act.H = w.H8; logit_act.W = w.W8; logit_act.Wp = w.W8; logit_act.Hp = w.H8; logit_act.C = 80;
  Act desc_act;  desc_act.f32 = w.descmap; desc_act.H = w.H4; desc_act.W = w.W4; desc_act.Wp = w.W4; desc_act.Hp = w.H4; desc_act.C = 128;
  if (tc) {
    // head epilogues fused: the detector head writes the exp-normalised 64 cell scores straight into `semi`,
    // the descriptor head writes L2-normalised rows (no softmax65 / l2norm128 launches in the tcgen05 modes)
    RUNP("tc_conv:headP", launch_conv_tc(A[PA], c->L("headP"), logit_act, nullptr, w.map_semi, split, c->num_sms, st, 2));
    if (!sparse_d) RUNP("tc_conv:headD", launch_conv_tc(A[DA], c->L("headD"), desc_act, nullptr, w.map_desc, split_d, c->num_sms, st, 1));
  } else {
    RUNP("conv_f32:headP", launch_conv_simt(A[PA], c->L("headP"), logit_act, nullptr, st));
    RUNP("conv_f32:headD", launch_conv_simt(A[DA], c->L("headD"), desc_act, nullptr, st));
  }
  if (!tc) {
    RUNP("softmax65", launch_softmax65(w.logits, w.H8 * w.W8, w.semi, st));
    RUNP("l2norm128", launch_l2norm128(w.descmap, w.H4 * w.W4, st));
  }
  if (p->use_stability && !fuse_sta) RUNP("sta", launch_sta(A[BA], tc ? (split == 3 ? 1 : 2) : 0, c->L("sta"), w.sta, st));
  if (cudaMemsetAsync(w.counter, 0, sizeof(int), st) != cudaSuccess) { set_error("cudaMemsetAsync(counter) failed"); return SFD2_ERR_CUDA; }
  RUNP("heat", launch_heat(w.semi, w.H8, w.W8, w.sta, w.H4, w.W4, p->use_stability, w.heat, H, W, st));
  RUNP("nms", launch_nms(w.heat, H, W, p->conf_th, p->border, p->border_w > 0 ? p->border_w : W, p->border_h > 0 ? p->border_h : H,
                         (c->debug_flags & 1) ? w.nmsdbg : nullptr, w.cand, w.cap,
                 w.counter, st, /*zero_counter=*/false));
  RUNP("select", launch_select(w.cand, w.cap, w.counter, W, p->topk, kpts, scores, count, w.status, w.scratch, st));
  if (sparse_d)   // the descriptor head only at the pixels that are sampled (tc_desc_sparse.cu); bit-identical rows
    RUNP("tc_conv:headD", launch_desc_sparse(A[DA], c->L("headD"), H, W, kpts, count, p->topk, w.drows, desc, st));
  else
    RUNP("sample", launch_sample(w.descmap, w.H4, w.W4, H, W, kpts, count, p->topk, desc, st));
#undef RUN
#undef RUNP
  return SFD2_OK;
}

static int check_params(const sfd2_extract_params* p, int n, int h, int w) {
  SFD2_CHECK(p != nullptr, SFD2_ERR_ARG, "params is NULL");
  SFD2_CHECK(n >= 1 && h >= 16 && w >= 16, SFD2_ERR_ARG, "bad image batch %d x %d x %d (min 16x16)", n, h, w);
  SFD2_CHECK(p->nms_radius == 4, SFD2_ERR_ARG, "only nms_radius == 4 is implemented (got %d)", p->nms_radius);
  SFD2_CHECK(p->topk >= 1, SFD2_ERR_ARG, "topk must be >= 1 (capacity of the output buffers)");
  SFD2_CHECK(p->precision >= 0 && p->precision <= 3, SFD2_ERR_ARG, "bad precision %d", p->precision);
  SFD2_CHECK(p->border >= 0, SFD2_ERR_ARG, "bad border");
  return SFD2_OK;
}

extern "C" {

SFD2_API int sfd2_abi_version(void) { return SFD2_ABI_VERSION; }
SFD2_API const char* sfd2_last_error(void) { return g_err; }

SFD2_API int sfd2_create(const void* blob, size_t nbytes, int device, sfd2_ctx** out) {
  SFD2_CHECK(out, SFD2_ERR_ARG, "sfd2_create: NULL argument");
  *out = nullptr;
  // blob == NULL: a matcher-only context (no network weights; extract calls fail with SFD2_ERR_WEIGHTS)
  const uint8_t* base = static_cast<const uint8_t*>(blob);
  uint32_t nl = 0;
  if (blob) {
    SFD2_CHECK(nbytes >= 16 && memcmp(blob, "SFD2W001", 8) == 0, SFD2_ERR_WEIGHTS, "bad weight blob magic");
    memcpy(&nl, base + 8, 4);
    SFD2_CHECK(nl > 0 && nl < 64 && 16 + (size_t)nl * sizeof(BlobLayer) <= nbytes, SFD2_ERR_WEIGHTS, "bad layer count %u", nl);
  }
  SFD2_CUDA(cudaSetDevice(device));
  if (const char* e = getenv("SFD2_TC_MULTICAST")) g_tc_multicast = atoi(e) != 0;
  if (const char* e = getenv("SFD2_TC_PDL")) g_tc_pdl = atoi(e) != 0;
  if (const char* e = getenv("SFD2_TC_CG2")) g_tc_cg2 = atoi(e);
  if (const char* e = getenv("SFD2_TC_STILES")) g_tc_stiles = atoi(e);
  if (const char* e = getenv("SFD2_TC_SLIM")) g_tc_slim = atoi(e) != 0;
  if (const char* e = getenv("SFD2_POST_PDL")) g_post_pdl = atoi(e) != 0;
  if (const char* e = getenv("SFD2_SPARSE_DESC")) g_sparse_desc = atoi(e) != 0;
  if (const char* e = getenv("SFD2_HOST_BANDS")) g_host_bands = atoi(e);
  if (const char* e = getenv("SFD2_BAND_LAYERS")) g_band_layers = atoi(e);
  if (const char* e = getenv("SFD2_TC_HALO")) g_tc_halo = atoi(e) != 0;
  if (const char* e = getenv("SFD2_TC_NSPLIT")) g_tc_nsplit = atoi(e) != 0;
  if (const char* e = getenv("SFD2_CONV1A_MMA")) g_conv1a_mma = atoi(e) != 0;
  if (const char* e = getenv("SFD2_FUSE_STA")) g_fuse_sta = atoi(e) != 0;
  if (const char* e = getenv("SFD2_TC_DIAGCAT")) g_tc_diagcat = atoi(e) != 0;
  if (const char* e = getenv("SFD2_TC_SPLIT1X1")) g_tc_split1x1 = atoi(e) != 0;
  const char* env_streams = getenv("SFD2_STREAMS");
  sfd2_ctx* c = new sfd2_ctx();
  c->device = device;
  if (env_streams) c->nstreams = atoi(env_streams);
  if (const char* e = getenv("SFD2_CAND_CAP")) c->cand_cap_override = atoi(e);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; set_error("cudaGetDeviceProperties failed"); return SFD2_ERR_CUDA; }
  c->num_sms = prop.multiProcessorCount;
  if (prop.major != 10) { delete c; set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return SFD2_ERR_CUDA; }
  for (uint32_t i = 0; i < nl; ++i) {
    BlobLayer bl;
    memcpy(&bl, base + 16 + (size_t)i * sizeof(BlobLayer), sizeof(BlobLayer));
    Layer L;
    L.name = std::string(bl.name, strnlen(bl.name, 16));
    L.cin = bl.cin; L.cout = bl.cout; L.k = bl.k; L.stride = bl.stride; L.groups = bl.groups; L.relu = bl.relu;
    const size_t wn = (size_t)L.cout * (L.cin / (L.groups > 0 ? L.groups : 1)) * L.k * L.k;
    if (L.groups < 1 || bl.w_off + wn * 4 > nbytes || bl.b_off + (size_t)L.cout * 4 > nbytes) {
      delete c; set_error("layer %s out of blob bounds", L.name.c_str()); return SFD2_ERR_WEIGHTS;
    }
    L.w.resize(wn); L.b.resize(L.cout);
    memcpy(L.w.data(), base + bl.w_off, wn * 4);
    memcpy(L.b.data(), base + bl.b_off, (size_t)L.cout * 4);
    c->lidx[L.name] = (int)c->layers.size();
    c->layers.push_back(std::move(L));
  }
  for (const char* n : kLayerNames)
    if (blob && !c->lidx.count(n)) { delete c; set_error("weight blob lacks layer %s", n); return SFD2_ERR_WEIGHTS; }
  for (Layer& L : c->layers) {
    int rc = upload_simt(L);
    if (!rc && L.cin % 64 == 0 && L.cout > 3) rc = tc_encode_weights(L);
    if (!rc && L.name == "conv1a") rc = conv1a_mma_encode(L);
    if (rc) { sfd2_destroy(c); return rc; }
  }
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { sfd2_destroy(c); set_error("stream create failed"); return SFD2_ERR_CUDA; }
  *out = c;
  return SFD2_OK;
}

SFD2_API int sfd2_destroy(sfd2_ctx* c) {
  if (!c) return SFD2_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (Ws& w : c->ws) free_workspace(w);
  for (Layer& L : c->layers) free_layer(L);
  cudaFree(c->img_dev); cudaFree(c->kp_dev); cudaFree(c->sc_dev); cudaFree(c->de_dev); cudaFree(c->cnt_dev);
  cudaFree(c->row_key); cudaFree(c->col_key); cudaFree(c->row2); cudaFree(c->col2); cudaFree(c->m_d0); cudaFree(c->m_d1);
  {
    sfd2_ctx::MatchWs& m = c->mws;
    cudaFree(m.hi); cudaFree(m.lo); cudaFree(m.keys); cudaFree(m.sec); cudaFree(m.remap); cudaFree(m.efflen); cudaFree(m.done); cudaFree(m.tab);
    for (auto& sl : m.ring) { if (sl.host) cudaFreeHost(sl.host); if (sl.ev) cudaEventDestroy(sl.ev); }
  }
  cudaFree(c->m_out); cudaFree(c->m_sim);
  for (auto& r : c->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  for (int k = 0; k < sfd2_ctx::kMaxStreams; ++k) { if (c->aux[k]) cudaStreamDestroy(c->aux[k]); if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]); }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (auto e : c->img_ready) cudaEventDestroy(e);
  delete c;
  return SFD2_OK;
}

// `ready`: optional per-image events (recorded on another stream when that image's bytes are in HBM)
static int extract_batch(sfd2_ctx* c, const void* img, int img_dtype, int n, int h, int w, const sfd2_extract_params* p,
                         float* kpts, float* scores, float* desc, int32_t* counts, void* stream, const cudaEvent_t* ready,
                         const Bands* bands = nullptr) {
  SFD2_CHECK(c && img && kpts && scores && desc && counts, SFD2_ERR_ARG, "sfd2_extract_dev: NULL argument");
  SFD2_CHECK(!c->layers.empty(), SFD2_ERR_WEIGHTS, "this context was created without network weights (matcher only)");
  int rc = check_params(p, n, h, w);
  if (rc) return rc;
  SFD2_CHECK(img_dtype == SFD2_IMG_F32_NCHW || img_dtype == SFD2_IMG_U8_NHWC, SFD2_ERR_ARG, "bad image dtype %d", img_dtype);
  SFD2_CUDA(cudaSetDevice(c->device));
  // A batch alternates between two workspaces on two internal streams (forked from / joined to the caller's
  // stream with events): every conv kernel is a persistent 1-CTA/SM grid whose last wave leaves SMs idle
  // (950 tiles over 148 SMs = 6.4 waves), and the other image's next kernel fills them.  Per-launch
  // profiling needs un-overlapped kernels, so it forces a single stream.
  int ns = ((n > 1 || ready) && c->nstreams > 1 && !c->prof_on) ? std::min(c->nstreams, (int)sfd2_ctx::kMaxStreams) : 1;
  if (ns > n && n > 1) ns = n;
  for (int k = 0; k < ns; ++k) {
    rc = ensure_workspace(c, c->ws[k], h, w, p->precision, p->topk);
    if (rc) return rc;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ns > 1) {
    if (!c->ev_fork) {
      SFD2_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
      for (int k = 0; k < sfd2_ctx::kMaxStreams; ++k) {
        SFD2_CUDA(cudaStreamCreateWithFlags(&c->aux[k], cudaStreamNonBlocking));
        SFD2_CUDA(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
      }
    }
    SFD2_CUDA(cudaEventRecord(c->ev_fork, st));
    for (int k = 0; k < ns; ++k) SFD2_CUDA(cudaStreamWaitEvent(c->aux[k], c->ev_fork, 0));
  }
  const size_t img_stride = (size_t)h * w * 3 * (img_dtype == SFD2_IMG_F32_NCHW ? 4 : 1);
  const long long before = g_launches;
  for (int i = 0; i < n && !rc; ++i) {
    const int k = (ns > 1) ? (i % ns) : 0;
    cudaStream_t si = ns > 1 ? c->aux[k] : st;
    if (ready && cudaStreamWaitEvent(si, ready[i], 0) != cudaSuccess) { set_error("cudaStreamWaitEvent(image %d) failed", i); rc = SFD2_ERR_CUDA; break; }
    rc = extract_one(c, c->ws[k], static_cast<const uint8_t*>(img) + i * img_stride, img_dtype, h, w, p,
                     kpts + (size_t)i * p->topk * 2, scores + (size_t)i * p->topk,
                     desc + (size_t)i * p->topk * SFD2_DESC_DIM, counts + i, si, n == 1 ? bands : nullptr);
  }
  // join the internal streams to the caller's stream also when an image failed: whatever was launched must
  // be ordered before the caller's next work (and before the workspaces are reused)
  if (ns > 1)
    for (int k = 0; k < ns; ++k) {
      const bool ok = cudaEventRecord(c->ev_join[k], c->aux[k]) == cudaSuccess &&
                      cudaStreamWaitEvent(st, c->ev_join[k], 0) == cudaSuccess;
      if (!ok && !rc) { set_error("joining internal stream %d failed", k); rc = SFD2_ERR_CUDA; }
    }
  c->launches += g_launches - before;
  c->last_prec = p->precision;
  return rc;
}

SFD2_API int sfd2_extract_dev(sfd2_ctx* c, const void* img, int img_dtype, int n, int h, int w, const sfd2_extract_params* p,
                     float* kpts, float* scores, float* desc, int32_t* counts, void* stream) {
  return extract_batch(c, img, img_dtype, n, h, w, p, kpts, scores, desc, counts, stream, nullptr);
}

SFD2_API int sfd2_extract_host(sfd2_ctx* c, const void* img, int img_dtype, int n, int h, int w, const sfd2_extract_params* p,
                      float* kpts, float* scores, float* desc, int32_t* counts) {
  SFD2_CHECK(c && img && kpts && scores && desc && counts, SFD2_ERR_ARG, "sfd2_extract_host: NULL argument");
  int rc = check_params(p, n, h, w);
  if (rc) return rc;
  SFD2_CUDA(cudaSetDevice(c->device));
  const size_t img_bytes = (size_t)n * h * w * 3 * (img_dtype == SFD2_IMG_F32_NCHW ? 4 : 1);
  if (img_bytes > c->img_cap) {
    cudaFree(c->img_dev); c->img_dev = nullptr; c->img_cap = 0;
    SFD2_CUDA(cudaMalloc(&c->img_dev, img_bytes));
    c->img_cap = img_bytes;
  }
  const size_t rows = (size_t)n * p->topk;
  if (rows > c->out_cap) {
    cudaFree(c->kp_dev); cudaFree(c->sc_dev); cudaFree(c->de_dev); cudaFree(c->cnt_dev);
    c->kp_dev = c->sc_dev = c->de_dev = nullptr; c->cnt_dev = nullptr; c->out_cap = 0;
    SFD2_CUDA(cudaMalloc(&c->kp_dev, rows * 2 * sizeof(float)));
    SFD2_CUDA(cudaMalloc(&c->sc_dev, rows * sizeof(float)));
    SFD2_CUDA(cudaMalloc(&c->de_dev, rows * SFD2_DESC_DIM * sizeof(float)));
    SFD2_CUDA(cudaMalloc(&c->cnt_dev, (rows + 1) * sizeof(int32_t)));
    c->out_cap = rows;
  }
  cudaStream_t st = c->stream;
  if (n == 1 && img_bytes >= (4u << 20) && h >= 64 && g_host_bands > 1) {
    // one large image: the upload goes out in row bands on the copy stream and conv1a follows it band by band, so only
    // the last band's share of conv1a is left on the critical path behind the 23 MB copy (1600 x 1200 float32)
    if (!c->copy_stream) SFD2_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    Bands bands;
    bands.n = std::min(g_host_bands, 8);
    while ((int)c->img_ready.size() < bands.n) {
      cudaEvent_t e = nullptr;
      SFD2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      c->img_ready.push_back(e);
    }
    SFD2_CUDA(cudaEventRecord(c->img_ready[0], st));               // previous users of img_dev on st are done
    SFD2_CUDA(cudaStreamWaitEvent(c->copy_stream, c->img_ready[0], 0));
    const uint8_t* src = static_cast<const uint8_t*>(img);
    uint8_t* dst = static_cast<uint8_t*>(c->img_dev);
    int r0 = 0;
    for (int b = 0; b < bands.n; ++b) {
      const int r1 = (int)((long long)h * (b + 1) / bands.n);
      if (img_dtype == SFD2_IMG_F32_NCHW) {
        // the band's rows of the three channel planes: one strided copy (3 "rows" of band bytes, pitch = one plane)
        const size_t off = (size_t)r0 * w * sizeof(float), plane = (size_t)h * w * sizeof(float);
        SFD2_CUDA(cudaMemcpy2DAsync(dst + off, plane, src + off, plane, (size_t)(r1 - r0) * w * sizeof(float), 3,
                                    cudaMemcpyHostToDevice, c->copy_stream));
      } else {
        const size_t off = (size_t)r0 * w * 3;
        SFD2_CUDA(cudaMemcpyAsync(dst + off, src + off, (size_t)(r1 - r0) * w * 3, cudaMemcpyHostToDevice, c->copy_stream));
      }
      SFD2_CUDA(cudaEventRecord(c->img_ready[b], c->copy_stream));
      bands.row_end[b] = r1;
      bands.ev[b] = c->img_ready[b];
      r0 = r1;
    }
    rc = extract_batch(c, c->img_dev, img_dtype, n, h, w, p, c->kp_dev, c->sc_dev, c->de_dev, c->cnt_dev, st, nullptr, &bands);
  } else if (n == 1) {
    SFD2_CUDA(cudaMemcpyAsync(c->img_dev, img, img_bytes, cudaMemcpyHostToDevice, st));
    rc = extract_batch(c, c->img_dev, img_dtype, n, h, w, p, c->kp_dev, c->sc_dev, c->de_dev, c->cnt_dev, st, nullptr);
  } else {
    // batch: the images stream in on a copy stream, one event per image, so image i+1's H2D overlaps image i's kernels
    if (!c->copy_stream) SFD2_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    while ((int)c->img_ready.size() < n) {
      cudaEvent_t e = nullptr;
      SFD2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      c->img_ready.push_back(e);
    }
    const size_t one = img_bytes / n;
    SFD2_CUDA(cudaEventRecord(c->img_ready[0], st));               // previous users of img_dev on st are done
    SFD2_CUDA(cudaStreamWaitEvent(c->copy_stream, c->img_ready[0], 0));
    for (int i = 0; i < n; ++i) {
      SFD2_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(c->img_dev) + i * one, static_cast<const uint8_t*>(img) + i * one, one,
                                cudaMemcpyHostToDevice, c->copy_stream));
      SFD2_CUDA(cudaEventRecord(c->img_ready[i], c->copy_stream));
    }
    rc = extract_batch(c, c->img_dev, img_dtype, n, h, w, p, c->kp_dev, c->sc_dev, c->de_dev, c->cnt_dev, st, c->img_ready.data());
  }
  if (rc) return rc;
  SFD2_CUDA(cudaMemcpyAsync(kpts, c->kp_dev, rows * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  SFD2_CUDA(cudaMemcpyAsync(scores, c->sc_dev, rows * sizeof(float), cudaMemcpyDeviceToHost, st));
  SFD2_CUDA(cudaMemcpyAsync(desc, c->de_dev, rows * SFD2_DESC_DIM * sizeof(float), cudaMemcpyDeviceToHost, st));
  SFD2_CUDA(cudaMemcpyAsync(counts, c->cnt_dev, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  return sfd2_extract_status(c, st);
}

// Candidate-overflow flag of the extract calls issued so far (select_kernel raises it; it is sticky per workspace
// until queried).  Synchronises `stream`, clears the flags.
SFD2_API int sfd2_extract_status(sfd2_ctx* c, void* stream) {
  SFD2_CHECK(c != nullptr, SFD2_ERR_ARG, "sfd2_extract_status: NULL ctx");
  SFD2_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int status[sfd2_ctx::kMaxStreams] = {};
  for (int k = 0; k < sfd2_ctx::kMaxStreams; ++k)
    if (c->ws[k].status) {
      SFD2_CUDA(cudaMemcpyAsync(&status[k], c->ws[k].status, sizeof(int), cudaMemcpyDeviceToHost, st));
      SFD2_CUDA(cudaMemsetAsync(c->ws[k].status, 0, sizeof(int), st));
    }
  SFD2_CUDA(cudaStreamSynchronize(st));
  if (status[0] | status[1] | status[2] | status[3]) {
    set_error("NMS produced more candidates than the workspace holds (cap %d); results truncated", c->ws[0].cap);
    return SFD2_ERR_OVERFLOW;
  }
  return SFD2_OK;
}

SFD2_API int sfd2_preprocess_dev(sfd2_ctx* c, const uint8_t* img, int h, int w, int swap_rb, int hn, int wn, float* out, void* stream) {
  SFD2_CHECK(c && img && out, SFD2_ERR_ARG, "sfd2_preprocess_dev: NULL argument");
  SFD2_CHECK(h >= 1 && w >= 1 && hn >= 1 && wn >= 1 && hn <= 65535, SFD2_ERR_ARG, "sfd2_preprocess_dev: bad size %d x %d -> %d x %d", h, w, hn, wn);
  SFD2_CUDA(cudaSetDevice(c->device));
  const long long before = g_launches;
  const int rc = launch_preprocess(img, h, w, swap_rb ? 1 : 0, hn, wn, out, static_cast<cudaStream_t>(stream));
  c->launches += g_launches - before;
  return rc;
}

// ---- matcher -----------------------------------------------------------------------------------------------
static int ensure_match_ws_simt(sfd2_ctx* c, int n0, int n1) {
  const size_t need = (size_t)(n0 > n1 ? n0 : n1) + 128;
  if (need > c->key_cap) {
    cudaFree(c->row_key); cudaFree(c->col_key); cudaFree(c->row2); cudaFree(c->col2);
    c->row_key = c->col_key = nullptr; c->row2 = c->col2 = nullptr; c->key_cap = 0;
    SFD2_CUDA(cudaMalloc(&c->row_key, need * sizeof(unsigned long long)));
    SFD2_CUDA(cudaMalloc(&c->col_key, need * sizeof(unsigned long long)));
    SFD2_CUDA(cudaMalloc(&c->row2, need * sizeof(unsigned)));
    SFD2_CUDA(cudaMalloc(&c->col2, need * sizeof(unsigned)));
    c->key_cap = need;
  }
  return SFD2_OK;
}

// CUDA-core fp32 reference mode (SFD2_PREC_FP32): one pair, host-known sizes, [n, 128] rows
static int match_one_simt(sfd2_ctx* c, const float* d0, int n0, const float* d1, int n1, int d, const sfd2_match_params* p,
                          int32_t* matches0, float* sim0, cudaStream_t st) {
  int rc = ensure_match_ws_simt(c, n0, n1);
  if (rc) return rc;
  prof_begin(c, "match_simt", st);
  rc = launch_match_simt(d0, n0, d1, n1, d, c->row_key, c->col_key, st);
  prof_end(c, st);
  if (rc) return rc;
  if (p->ratio_threshold > 0.f) {   // second-best pass on CUDA cores (fp32 mode only; the tcgen05 modes keep top-2 in the epilogue)
    rc = launch_match_second(d0, n0, d1, n1, d, c->row_key, c->col_key, c->row2, c->col2, st);
    if (rc) return rc;
  }
  return launch_match_finish(c->row_key, c->col_key, n0, n1, p->do_mutual_check, p->distance_threshold,
                             p->ratio_threshold, ((p->ratio_mode & 0xFF) == 1 ? 2 : 1) | (p->ratio_mode & SFD2_MATCH_PLAIN_CODES),
                             c->row2, c->col2, matches0, sim0, st);
}

// the grouped tcgen05 matcher: tables -> device, prep (split / compaction / resets), one GEMM + arg-max + finish launch
static int match_pairs_tc(sfd2_ctx* c, const sfd2_desc_set* sets, int nsets, const int32_t* pa, const int32_t* pb, int npairs,
                          const sfd2_match_params* p, int32_t* matches0, float* sim0, cudaStream_t st) {
  sfd2_ctx::MatchWs& m = c->mws;
  std::vector<MOperD> opers(nsets);
  long long prow = 0;
  bool any_ids = false;
  for (int i = 0; i < nsets; ++i) {
    const sfd2_desc_set& s = sets[i];
    SFD2_CHECK(s.n >= 0 && (s.n == 0 || s.data), SFD2_ERR_ARG, "match: set %d has n = %d, data = %p", i, s.n, (const void*)s.data);
    SFD2_CHECK(s.layout == SFD2_DESC_ROWS || s.layout == SFD2_DESC_COLS, SFD2_ERR_ARG, "match: set %d: bad layout %d", i, s.layout);
    MOperD& o = opers[i];
    o.src = s.data;
    o.layout = s.layout;
    o.rs = s.layout == SFD2_DESC_ROWS ? SFD2_DESC_DIM : 1;
    o.cs = s.layout == SFD2_DESC_ROWS ? 1 : s.n;
    o.cap = s.n; o.prow0 = (int)prow; o.pad_ = 0;
    o.count = s.count; o.ids = s.ids;
    any_ids |= s.ids != nullptr;
    prow += round_up(s.n > 0 ? s.n : 1, 128);
    SFD2_CHECK(prow < (1ll << 30), SFD2_ERR_ARG, "match: too many descriptor rows in one call");
  }
  const int passes = p->ratio_threshold > 0.f ? 2 : 1;
  // mutual mode (one product + column reduction) is bound by its epilogue: one row-block per unit; the row-only modes are
  // bound by the B traffic: two row-blocks share every B tile
  const int sub = (passes == 1 && p->do_mutual_check) ? 1 : 2;
  std::vector<MProbD> probs(npairs);
  long long koff = 0, ooff = 0, tiles = 0;
  for (int k = 0; k < npairs; ++k) {
    SFD2_CHECK(pa[k] >= 0 && pa[k] < nsets && pb[k] >= 0 && pb[k] < nsets, SFD2_ERR_ARG, "match: pair %d refers to a set out of range", k);
    MProbD& q = probs[k];
    q.a = pa[k]; q.b = pb[k];
    const int na = sets[q.a].n, nb = sets[q.b].n;
    q.tm = cdiv(na, 128); q.tn = cdiv(nb, 128);
    q.tile0 = (int)tiles;
    q.ntiles = tm_units(q.tm, q.tn, passes, sub);
    tiles += q.ntiles;
    q.key_a = koff; koff += round_up(na > 0 ? na : 1, 128);
    q.key_b = koff; koff += round_up(nb > 0 ? nb : 1, 128);
    q.out_off = ooff; ooff += na;
    SFD2_CHECK(tiles < (1ll << 30), SFD2_ERR_ARG, "match: too many tiles in one call");
    if (q.ntiles == 0 && na > 0) {      // empty db: no CTA will ever finish this pair - every row is unmatched
      const size_t isz = (p->ratio_mode & SFD2_MATCH_I64) ? 8 : 4;
      SFD2_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(matches0) + q.out_off * isz, 0xFF, (size_t)na * isz, st));
      SFD2_CUDA(cudaMemsetAsync(sim0 + q.out_off, 0, (size_t)na * sizeof(float), st));
    }
  }
  if (tiles == 0) return SFD2_OK;
  // workspace (grow-only)
  int rc;
  bool grew_hi = false, grew_lo = false;
  if ((rc = reserve(m.hi, m.cap_hi, (size_t)prow * 128 * sizeof(__half), &grew_hi))) return rc;
  if ((rc = reserve(m.lo, m.cap_lo, (size_t)prow * 128 * sizeof(__half), &grew_lo))) return rc;
  if (grew_hi || grew_lo || !m.maps_ok) {
    // the maps cover the whole allocation, so they only change when the planes are re-allocated
    if ((rc = tm_make_plane_map(&m.tm_hi, m.hi, m.cap_hi / 256))) return rc;
    if ((rc = tm_make_plane_map(&m.tm_lo, m.lo, m.cap_lo / 256))) return rc;
    m.maps_ok = true;
  }
  if ((rc = reserve(m.keys, m.cap_keys, (size_t)koff * sizeof(unsigned long long)))) return rc;
  if (passes == 2 && (rc = reserve(m.sec, m.cap_sec, (size_t)koff * sizeof(unsigned)))) return rc;
  if (any_ids && (rc = reserve(m.remap, m.cap_remap, (size_t)prow * sizeof(int)))) return rc;
  if ((rc = reserve(m.efflen, m.cap_efflen, (size_t)nsets * sizeof(int)))) return rc;
  if ((rc = reserve(m.done, m.cap_done, (size_t)npairs * sizeof(int)))) return rc;
  // tables: a single pair rides in the kernel parameters; bigger calls go through a pinned staging ring -> device
  // (stream-ordered behind the previous call's kernels)
  MTabInline inl{};
  const MOperD* opers_dev = nullptr;
  const MProbD* probs_dev = nullptr;
  if (nsets <= 2 && npairs == 1) {
    for (int i = 0; i < nsets; ++i) inl.opers[i] = opers[i];
    inl.probs[0] = probs[0];
  } else {
    const size_t tab_bytes = opers.size() * sizeof(MOperD) + probs.size() * sizeof(MProbD);
    if ((rc = reserve(m.tab, m.cap_tab, tab_bytes))) return rc;
    sfd2_ctx::MatchWs::Slot& slot = m.ring[m.next++ % sfd2_ctx::MatchWs::kRing];
    if (!slot.ev) SFD2_CUDA(cudaEventCreateWithFlags(&slot.ev, cudaEventDisableTiming));
    else SFD2_CUDA(cudaEventSynchronize(slot.ev));      // the copy that last used this slot has completed
    if (tab_bytes > slot.cap) {
      if (slot.host) cudaFreeHost(slot.host);
      slot.host = nullptr; slot.cap = 0;
      SFD2_CUDA(cudaMallocHost(&slot.host, tab_bytes * 2));
      slot.cap = tab_bytes * 2;
    }
    memcpy(slot.host, opers.data(), opers.size() * sizeof(MOperD));
    memcpy(static_cast<uint8_t*>(slot.host) + opers.size() * sizeof(MOperD), probs.data(), probs.size() * sizeof(MProbD));
    SFD2_CUDA(cudaMemcpyAsync(m.tab, slot.host, tab_bytes, cudaMemcpyHostToDevice, st));
    SFD2_CUDA(cudaEventRecord(slot.ev, st));
    opers_dev = reinterpret_cast<const MOperD*>(m.tab);
    probs_dev = reinterpret_cast<const MProbD*>(m.tab + opers.size() * sizeof(MOperD));
  }
  prof_begin(c, "match_prep", st);
  rc = launch_match_prep(opers_dev, &inl, nsets, (int)prow, any_ids, m.hi, m.lo, m.remap, m.efflen, m.keys, passes == 2 ? m.sec : nullptr,
                         koff, m.done, npairs, c->num_sms, st);
  prof_end(c, st);
  if (rc) return rc;
  TcMatchArgs a{};
  a.opers = opers_dev; a.probs = probs_dev; a.inl = inl;
  a.nprob = npairs; a.total_tiles = (int)tiles;
  a.passes = passes;
  a.sub = sub;
  a.cols = (passes == 1 && p->do_mutual_check) ? 1 : 0;
  a.split = p->precision != SFD2_PREC_TC_FAST ? 3 : 1;
  a.mutual = p->do_mutual_check ? 1 : 0;
  a.ratio_mode = ((p->ratio_mode & 0xFF) == 1 ? 2 : 1) | (p->ratio_mode & (SFD2_MATCH_PLAIN_CODES | SFD2_MATCH_HLOC_SCORES | SFD2_MATCH_I64));
  a.dist_th = p->distance_threshold; a.ratio_th = p->ratio_threshold;
  a.keys = m.keys; a.sec = passes == 2 ? m.sec : nullptr;
  a.efflen = m.efflen; a.remap = m.remap; a.done = m.done;
  a.matches0 = matches0; a.sim0 = sim0;
  prof_begin(c, "match_tc", st);
  rc = launch_match_tc(m.tm_hi, m.tm_lo, a, c->num_sms, st);
  prof_end(c, st);
  return rc;
}

static int check_match_params(const sfd2_match_params* p) {
  SFD2_CHECK(p != nullptr, SFD2_ERR_ARG, "match params is NULL");
  SFD2_CHECK(p->precision >= 0 && p->precision <= 3, SFD2_ERR_ARG, "bad precision %d", p->precision);
  SFD2_CHECK(p->layout == SFD2_DESC_ROWS || p->layout == SFD2_DESC_COLS, SFD2_ERR_ARG, "bad descriptor layout %d", p->layout);
  return SFD2_OK;
}

SFD2_API int sfd2_match_pairs_dev(sfd2_ctx* c, const sfd2_desc_set* sets, int nsets, const int32_t* pa, const int32_t* pb,
                                  int npairs, const sfd2_match_params* p, int32_t* matches0, float* sim0, void* stream) {
  SFD2_CHECK(c && sets && pa && pb && nsets >= 1 && npairs >= 0, SFD2_ERR_ARG, "sfd2_match_pairs_dev: bad argument");
  int rc = check_match_params(p);
  if (rc) return rc;
  if (npairs == 0) return SFD2_OK;
  SFD2_CHECK(matches0 && sim0, SFD2_ERR_ARG, "sfd2_match_pairs_dev: NULL output");
  SFD2_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long before = g_launches;
  if (p->precision == SFD2_PREC_FP32) {   // CUDA-core reference mode: plain loop over host-sized row-major sets
    SFD2_CHECK(!(p->ratio_mode & (SFD2_MATCH_HLOC_SCORES | SFD2_MATCH_I64)), SFD2_ERR_ARG,
               "the fp32 CUDA-core mode writes int32 matches and raw similarities only");
    long long off = 0;
    for (int k = 0; k < npairs && !rc; ++k) {
      SFD2_CHECK(pa[k] >= 0 && pa[k] < nsets && pb[k] >= 0 && pb[k] < nsets, SFD2_ERR_ARG, "match: pair %d refers to a set out of range", k);
      const sfd2_desc_set &a = sets[pa[k]], &b = sets[pb[k]];
      SFD2_CHECK(!a.count && !b.count && !a.ids && !b.ids && a.layout == SFD2_DESC_ROWS && b.layout == SFD2_DESC_ROWS, SFD2_ERR_ARG,
                 "the fp32 CUDA-core mode takes host-sized [n,128] sets only (no device counts / ids / [128,n] layout)");
      if (a.n > 0) rc = match_one_simt(c, a.data, a.n, b.data, b.n, SFD2_DESC_DIM, p, matches0 + off, sim0 + off, st);
      off += a.n;
    }
  } else {
    rc = match_pairs_tc(c, sets, nsets, pa, pb, npairs, p, matches0, sim0, st);
  }
  c->launches += g_launches - before;
  return rc;
}

SFD2_API int sfd2_match_dev(sfd2_ctx* c, const float* d0, int n0, const float* d1, int n1, int d, const sfd2_match_params* p,
                   int32_t* matches0, float* sim0, void* stream) {
  SFD2_CHECK(c && p && (n0 == 0 || (d0 && matches0 && sim0)) && (n1 == 0 || d1), SFD2_ERR_ARG, "sfd2_match_dev: NULL argument");
  SFD2_CHECK(n0 >= 0 && n1 >= 0 && d >= 1, SFD2_ERR_ARG, "sfd2_match_dev: bad shape %d x %d x %d", n0, n1, d);
  int rc = check_match_params(p);
  if (rc) return rc;
  if (n0 == 0) return SFD2_OK;
  if (p->precision == SFD2_PREC_FP32) {
    SFD2_CHECK(p->layout == SFD2_DESC_ROWS && !(p->ratio_mode & (SFD2_MATCH_HLOC_SCORES | SFD2_MATCH_I64)), SFD2_ERR_ARG,
               "the fp32 CUDA-core mode takes [n,128] rows and writes int32 matches / raw similarities only");
    SFD2_CUDA(cudaSetDevice(c->device));
    const long long before = g_launches;
    rc = match_one_simt(c, d0, n0, d1, n1, d, p, matches0, sim0, static_cast<cudaStream_t>(stream));
    c->launches += g_launches - before;
    return rc;
  }
  SFD2_CHECK(d == SFD2_DESC_DIM, SFD2_ERR_ARG, "match: descriptor dim must be 128 (got %d)", d);
  const sfd2_desc_set sets[2] = {{d0, n0, p->layout, nullptr, nullptr}, {d1, n1, p->layout, nullptr, nullptr}};
  const int32_t a = 0, b = 1;
  return sfd2_match_pairs_dev(c, sets, 2, &a, &b, 1, p, matches0, sim0, stream);
}

SFD2_API int sfd2_match_batched_dev(sfd2_ctx* c, const float* d0, const int32_t* off0, const float* d1, const int32_t* off1,
                           int npairs, int d, const sfd2_match_params* p, int32_t* matches0, float* sim0, void* stream) {
  SFD2_CHECK(c && p && off0 && off1 && npairs >= 0, SFD2_ERR_ARG, "sfd2_match_batched_dev: bad argument");
  SFD2_CHECK(d == SFD2_DESC_DIM, SFD2_ERR_ARG, "match: descriptor dim must be 128 (got %d)", d);
  std::vector<sfd2_desc_set> sets(2 * (size_t)npairs);
  std::vector<int32_t> pa(npairs), pb(npairs);
  for (int i = 0; i < npairs; ++i) {
    SFD2_CHECK(off0[i + 1] >= off0[i] && off1[i + 1] >= off1[i], SFD2_ERR_ARG, "offsets must be non-decreasing");
    sets[2 * i] = sfd2_desc_set{d0 + (size_t)off0[i] * d, off0[i + 1] - off0[i], SFD2_DESC_ROWS, nullptr, nullptr};
    sets[2 * i + 1] = sfd2_desc_set{d1 + (size_t)off1[i] * d, off1[i + 1] - off1[i], SFD2_DESC_ROWS, nullptr, nullptr};
    pa[i] = 2 * i; pb[i] = 2 * i + 1;
  }
  if (npairs == 0) return SFD2_OK;
  // outputs are indexed like d0: pair i's rows start at off0[i] - off0[0] = the packed offset when off0[0] == 0
  return sfd2_match_pairs_dev(c, sets.data(), 2 * npairs, pa.data(), pb.data(), npairs, p, matches0 + off0[0], sim0 + off0[0], stream);
}

SFD2_API int sfd2_match_one_to_many_dev(sfd2_ctx* c, const float* q, int nq, const float* db, const int32_t* db_off,
                                        int ndb, int d, const sfd2_match_params* p, int32_t* matches0, float* sim0,
                                        void* stream) {
  SFD2_CHECK(c && p && q && db && db_off && matches0 && sim0, SFD2_ERR_ARG, "sfd2_match_one_to_many_dev: NULL argument");
  SFD2_CHECK(nq >= 1 && ndb >= 1 && d == SFD2_DESC_DIM, SFD2_ERR_ARG, "sfd2_match_one_to_many_dev: bad shape (d must be 128)");
  std::vector<sfd2_desc_set> sets(1 + (size_t)ndb);
  std::vector<int32_t> pa(ndb, 0), pb(ndb);
  sets[0] = sfd2_desc_set{q, nq, SFD2_DESC_ROWS, nullptr, nullptr};
  for (int i = 0; i < ndb; ++i) {
    SFD2_CHECK(db_off[i + 1] >= db_off[i], SFD2_ERR_ARG, "offsets must be non-decreasing");
    sets[1 + i] = sfd2_desc_set{db + (size_t)db_off[i] * d, db_off[i + 1] - db_off[i], SFD2_DESC_ROWS, nullptr, nullptr};
    pb[i] = 1 + i;
  }
  return sfd2_match_pairs_dev(c, sets.data(), 1 + ndb, pa.data(), pb.data(), ndb, p, matches0, sim0, stream);
}

SFD2_API int sfd2_match_host(sfd2_ctx* c, const float* d0, int n0, const float* d1, int n1, int d, const sfd2_match_params* p,
                    int32_t* matches0, float* sim0) {
  SFD2_CHECK(c && p, SFD2_ERR_ARG, "sfd2_match_host: NULL argument");
  SFD2_CHECK(n0 >= 0 && n1 >= 0 && d >= 1, SFD2_ERR_ARG, "sfd2_match_host: bad shape");
  if (n0 == 0) return SFD2_OK;
  SFD2_CUDA(cudaSetDevice(c->device));
  const size_t b0 = (size_t)n0 * d * sizeof(float), b1 = (size_t)(n1 > 0 ? n1 : 1) * d * sizeof(float);
  if (b0 > c->m_d0_cap) { cudaFree(c->m_d0); c->m_d0 = nullptr; c->m_d0_cap = 0; SFD2_CUDA(cudaMalloc(&c->m_d0, b0)); c->m_d0_cap = b0; }
  if (b1 > c->m_d1_cap) { cudaFree(c->m_d1); c->m_d1 = nullptr; c->m_d1_cap = 0; SFD2_CUDA(cudaMalloc(&c->m_d1, b1)); c->m_d1_cap = b1; }
  if ((size_t)n0 > c->m_out_cap) {
    cudaFree(c->m_out); cudaFree(c->m_sim); c->m_out = nullptr; c->m_sim = nullptr; c->m_out_cap = 0;
    SFD2_CUDA(cudaMalloc(&c->m_out, (size_t)n0 * sizeof(int32_t)));
    SFD2_CUDA(cudaMalloc(&c->m_sim, (size_t)n0 * sizeof(float)));
    c->m_out_cap = n0;
  }
  cudaStream_t st = c->stream;
  SFD2_CUDA(cudaMemcpyAsync(c->m_d0, d0, b0, cudaMemcpyHostToDevice, st));
  if (n1 > 0) SFD2_CUDA(cudaMemcpyAsync(c->m_d1, d1, (size_t)n1 * d * sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = sfd2_match_dev(c, c->m_d0, n0, c->m_d1, n1, d, p, c->m_out, c->m_sim, st);
  if (rc) return rc;
  SFD2_CUDA(cudaMemcpyAsync(matches0, c->m_out, (size_t)n0 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  SFD2_CUDA(cudaMemcpyAsync(sim0, c->m_sim, (size_t)n0 * sizeof(float), cudaMemcpyDeviceToHost, st));
  SFD2_CUDA(cudaStreamSynchronize(st));
  return SFD2_OK;
}

SFD2_API long long sfd2_launch_count(sfd2_ctx* c) { return c ? c->launches : -1; }

SFD2_API int sfd2_profile(sfd2_ctx* c, int enable) {
  SFD2_CHECK(c != nullptr, SFD2_ERR_ARG, "sfd2_profile: NULL ctx");
  c->prof_on = enable != 0;
  return SFD2_OK;
}

// "label\tlaunches\ttotal_ms\n" per kernel label since the last read; synchronises the device.
SFD2_API long long sfd2_profile_read(sfd2_ctx* c, char* buf, long long capacity) {
  if (!c || !buf || capacity < 1) { set_error("sfd2_profile_read: bad argument"); return SFD2_ERR_ARG; }
  cudaSetDevice(c->device);
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("sync failed"); return SFD2_ERR_CUDA; }
  std::map<std::string, std::pair<long long, double>> agg;
  std::vector<std::string> order;
  for (auto& r : c->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    if (!agg.count(r.label)) order.push_back(r.label);
    auto& e = agg[r.label];
    e.first += 1; e.second += ms;
    c->ev_pool.push_back(r.a); c->ev_pool.push_back(r.b);
  }
  c->prof.clear();
  std::string out;
  char line[256];
  for (auto& k : order) {
    snprintf(line, sizeof(line), "%s\t%lld\t%.6f\n", k.c_str(), agg[k].first, agg[k].second);
    out += line;
  }
  if ((long long)out.size() + 1 > capacity) { set_error("profile buffer too small"); return SFD2_ERR_ARG; }
  memcpy(buf, out.c_str(), out.size() + 1);
  return (long long)out.size();
}

SFD2_API int sfd2_nms_select_dev(sfd2_ctx* c, const float* heat, int h, int w, const sfd2_extract_params* p, float* kpts,
                        float* scores, int32_t* count, float* nms_out, void* stream) {
  SFD2_CHECK(c && heat && p && kpts && scores && count, SFD2_ERR_ARG, "sfd2_nms_select_dev: NULL argument");
  SFD2_CHECK(h >= 1 && w >= 1 && p->topk >= 1 && p->nms_radius == 4, SFD2_ERR_ARG, "sfd2_nms_select_dev: bad argument");
  SFD2_CUDA(cudaSetDevice(c->device));
  // private workspace sized for this map (the extract workspace may belong to another size)
  const int cap = (int)(((size_t)h * w) / 16) + 4096;
  int cap2 = 1;
  while (cap2 < cap) cap2 <<= 1;
  unsigned long long *cand = nullptr, *scratch = nullptr;
  int *counter = nullptr, *status = nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int hstatus = 0, rc = SFD2_OK;
  const long long before = g_launches;
  auto ck = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && !rc) { set_error("sfd2_nms_select_dev: %s -> %s", what, cudaGetErrorString(e)); rc = SFD2_ERR_CUDA; }
    return e == cudaSuccess;
  };
  if (ck(cudaMalloc(&cand, (size_t)cap * 8), "cudaMalloc(cand)") && ck(cudaMalloc(&scratch, (size_t)cap2 * 8), "cudaMalloc(scratch)") &&
      ck(cudaMalloc(&counter, 4), "cudaMalloc(counter)") && ck(cudaMalloc(&status, 4), "cudaMalloc(status)") &&
      ck(cudaMemsetAsync(status, 0, 4, st), "memset(status)") && ck(cudaMemsetAsync(scratch, 0, (size_t)cap2 * 8, st), "memset(scratch)")) {
    rc = launch_nms(heat, h, w, p->conf_th, p->border, p->border_w > 0 ? p->border_w : w, p->border_h > 0 ? p->border_h : h,
                    nms_out, cand, cap, counter, st);
    if (!rc) rc = launch_select(cand, cap, counter, w, p->topk, kpts, scores, count, status, scratch, st);
    ck(cudaMemcpyAsync(&hstatus, status, 4, cudaMemcpyDeviceToHost, st), "copy(status)");
  }
  // the private buffers must outlive the kernels: synchronise (also on the error paths) before freeing them
  ck(cudaStreamSynchronize(st), "cudaStreamSynchronize");
  c->launches += g_launches - before;
  cudaFree(cand); cudaFree(scratch); cudaFree(counter); cudaFree(status);
  if (!rc && hstatus) { set_error("candidate overflow (cap %d)", cap); rc = SFD2_ERR_OVERFLOW; }
  return rc;
}

// name -> intermediate of the last image.  Activations are returned as dense [H][W][C] fp32.
SFD2_API long long sfd2_debug_fetch(sfd2_ctx* c, const char* name, float* out, long long capacity) {
  if (!c || !name) { set_error("sfd2_debug_fetch: NULL argument"); return SFD2_ERR_ARG; }
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  const std::string n(name);
  if (n == "enable_nms_out") { c->debug_flags |= 1; return 0; }
  Ws& w = c->ws[0];   // single-image calls (and the even images of a batch) use workspace 0
  if (w.wsH == 0) { set_error("no image has been extracted yet"); return SFD2_ERR_ARG; }
  const float* src = nullptr;
  long long cnt = 0;
  if (n == "heat") { src = w.heat; cnt = (long long)w.wsH * w.wsW; }
  else if (n == "nms") { src = w.nmsdbg; cnt = (long long)w.wsH * w.wsW; }
  else if (n == "semi") { src = w.semi; cnt = (long long)w.H8 * w.W8 * 64; }
  else if (n == "logits") { src = w.logits; cnt = (long long)w.H8 * w.W8 * 80; }
  else if (n == "desc_map") { src = w.descmap; cnt = (long long)w.H4 * w.W4 * 128; }
  else if (n == "sta_logits") { src = w.sta; cnt = (long long)w.H4 * w.W4 * 3; }
  if (src) {
    if (cnt > capacity) { set_error("buffer too small: need %lld floats", cnt); return SFD2_ERR_ARG; }
    if (cudaMemcpy(out, src, (size_t)cnt * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("copy failed"); return SFD2_ERR_CUDA; }
    return cnt;
  }
  static const std::map<std::string, int> ids = {{"conv1a", A1A}, {"conv1b", A1B}, {"conv2a", A2A}, {"conv2b", A2B},
                                                 {"conv3a", A3A}, {"conv3b", A3B}, {"t1", T1}, {"t2", T2},
                                                 {"out4", BA}, {"rb1", BB}, {"convPa0", PA}, {"convDa0", DA}};
  auto it = ids.find(n);
  if (it == ids.end()) { set_error("unknown intermediate '%s'", name); return SFD2_ERR_ARG; }
  const Act& a = w.acts[it->second];
  cnt = (long long)a.H * a.W * a.C;
  if (cnt > capacity) { set_error("buffer too small: need %lld floats", cnt); return SFD2_ERR_ARG; }
  const bool tc = (c->last_prec != SFD2_PREC_FP32);
  const size_t ne = a.elems();
  if (!tc) {
    std::vector<float> tmp(ne);
    if (cudaMemcpy(tmp.data(), a.f32, ne * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("copy failed"); return SFD2_ERR_CUDA; }
    for (int y = 0; y < a.H; ++y)
      memcpy(out + (size_t)y * a.W * a.C, tmp.data() + (size_t)y * a.Wp * a.C, (size_t)a.W * a.C * 4);
  } else {
    std::vector<__half> hi(ne), lo(ne);
    if (cudaMemcpy(hi.data(), a.hi, ne * 2, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(lo.data(), a.lo, ne * 2, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("copy failed"); return SFD2_ERR_CUDA; }
    const bool use_lo = (c->last_prec == SFD2_PREC_TC_EXACT) || (c->last_prec == SFD2_PREC_TC_MIXED && it->second != DA);
    for (int y = 0; y < a.H; ++y)
      for (size_t i = 0; i < (size_t)a.W * a.C; ++i) {
        const size_t s = (size_t)y * a.Wp * a.C + i;
        out[(size_t)y * a.W * a.C + i] = __half2float(hi[s]) + (use_lo ? __half2float(lo[s]) : 0.f);
      }
  }
  return cnt;
}

// One convolution layer in isolation (unit tests): x NHWC fp32 [h][w][cin] on the host, weights OIHW,
// y NHWC fp32 [ho][wo][cout] on the host.  precision selects the CUDA-core or tcgen05 kernel.
SFD2_API int sfd2_debug_conv(sfd2_ctx* c, const float* x, int h, int w, int cin, const float* wt, const float* bs, int cout,
                    int ksize, int stride, int groups, int relu, int precision, float* y) {
  SFD2_CHECK(c && x && wt && bs && y, SFD2_ERR_ARG, "sfd2_debug_conv: NULL argument");
  SFD2_CHECK((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2) && (groups == 1 || groups == 32), SFD2_ERR_ARG, "sfd2_debug_conv: unsupported conv");
  SFD2_CUDA(cudaSetDevice(c->device));
  Layer L;
  L.name = "dbg"; L.cin = cin; L.cout = cout; L.k = ksize; L.stride = stride; L.groups = groups; L.relu = relu;
  const size_t wn = (size_t)cout * (cin / groups) * ksize * ksize;
  L.w.assign(wt, wt + wn); L.b.assign(bs, bs + cout);
  int rc = upload_simt(L);
  const bool tc = precision != SFD2_PREC_FP32;
  if (!rc && tc) rc = tc_encode_weights(L);
  Act in, out;
  in.H = h; in.W = w; in.C = cin; in.Hp = round_up(h, 2); in.Wp = round_up(w, 2);
  out.H = conv_out(h, stride); out.W = conv_out(w, stride); out.C = cout; out.Hp = round_up(out.H, 2); out.Wp = round_up(out.W, 2);
  const int outC_f32 = round_up(cout, 16);
  CUtensorMap maps[8];
  float* yf = nullptr;
  std::vector<float> xin(in.elems(), 0.f);
  for (int yy = 0; yy < h; ++yy) memcpy(xin.data() + (size_t)yy * in.Wp * cin, x + (size_t)yy * w * cin, (size_t)w * cin * 4);
  auto cleanup = [&]() { cudaFree(in.f32); cudaFree(in.hi); cudaFree(in.lo); cudaFree(out.f32); cudaFree(yf); free_layer(L); };
#define DBG_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_error("%s -> %s", #call, cudaGetErrorString(e_)); cleanup(); return SFD2_ERR_CUDA; } } while (0)
  if (rc) { cleanup(); return rc; }
  const long long before = g_launches;
  if (!tc) {
    DBG_CUDA(cudaMalloc(&in.f32, in.elems() * 4));
    DBG_CUDA(cudaMemcpy(in.f32, xin.data(), in.elems() * 4, cudaMemcpyHostToDevice));
    out.Wp = out.W; out.Hp = out.H;
    DBG_CUDA(cudaMalloc(&out.f32, out.elems() * 4));
    rc = launch_conv_simt(in, L, out, nullptr, nullptr);
    if (!rc) { DBG_CUDA(cudaDeviceSynchronize()); DBG_CUDA(cudaMemcpy(y, out.f32, out.elems() * 4, cudaMemcpyDeviceToHost)); }
  } else {
    std::vector<__half> hi(in.elems()), lo(in.elems());
    for (size_t i = 0; i < in.elems(); ++i) { hi[i] = __float2half_rn(xin[i]); lo[i] = __float2half_rn(xin[i] - __half2float(hi[i])); }
    DBG_CUDA(cudaMalloc(&in.hi, in.elems() * 2));
    DBG_CUDA(cudaMalloc(&in.lo, in.elems() * 2));
    DBG_CUDA(cudaMemcpy(in.hi, hi.data(), in.elems() * 2, cudaMemcpyHostToDevice));
    DBG_CUDA(cudaMemcpy(in.lo, lo.data(), in.elems() * 2, cudaMemcpyHostToDevice));
    rc = tc_make_act_maps(in, in.hi, &maps[0], &maps[2], &maps[4], &maps[6]);
    if (!rc) rc = tc_make_act_maps(in, in.lo, &maps[1], &maps[3], &maps[5], &maps[7]);
    in.tm = maps;
    Act o2 = out; o2.Wp = out.W; o2.Hp = out.H; o2.C = outC_f32;
    DBG_CUDA(cudaMalloc(&yf, o2.elems() * 4));
    CUtensorMap omap[2];
    if (!rc) rc = tc_make_store_map(&omap[0], yf, o2.C, o2.W, o2.H, o2.Wp, 1, 16);
    if (!rc) rc = tc_make_store_map(&omap[1], yf, o2.C, o2.W, o2.H, o2.Wp, 1, 8);
    if (!rc) rc = launch_conv_tc(in, L, o2, nullptr, omap, precision == SFD2_PREC_TC_EXACT ? 3 : 1, c->num_sms, nullptr);
    if (!rc) {
      DBG_CUDA(cudaDeviceSynchronize());
      std::vector<float> tmp(o2.elems());
      DBG_CUDA(cudaMemcpy(tmp.data(), yf, o2.elems() * 4, cudaMemcpyDeviceToHost));
      for (size_t pix = 0; pix < (size_t)out.H * out.W; ++pix)
        memcpy(y + pix * cout, tmp.data() + pix * outC_f32, (size_t)cout * 4);
    }
  }
#undef DBG_CUDA
  c->launches += g_launches - before;
  cleanup();
  return rc;
}

}  // extern "C"
