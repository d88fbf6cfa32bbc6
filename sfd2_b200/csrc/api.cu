// C ABI of libsfd2_b200.so (see include/sfd2_b200.h): context, weight blob parsing, workspace
// management and the per-image kernel schedule of the extract path, plus the matcher entry points.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "common.cuh"

namespace sfd2 {

static thread_local char g_err[1024] = "";
thread_local long long g_launches = 0;

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

// per-image extract workspace (activations, TMA views, head buffers, candidate list), sized for wsH x wsW
struct Ws {
  int wsH = 0, wsW = 0;
  bool have_f32 = false, have_tc = false;
  Act acts[NUM_ACTS];
  CUtensorMap maps[NUM_ACTS][8];
  CUtensorMap st_maps[NUM_ACTS][4];
  CUtensorMap map_1a[4];                    // conv1a output rows: [hi, lo] box {64 ch, 256 px, 1 row}, [hi, lo] box {64, 128, 1}
  CUtensorMap map_logits[2], map_desc[2], map_semi[2];   // fp32 head outputs (TMA store views: 16x2 and 8x4 boxes)
  int H2 = 0, W2 = 0, H4 = 0, W4 = 0, H8 = 0, W8 = 0;
  float4* nimg = nullptr;  // normalised image, NHWC4 fp32
  float *logits = nullptr, *semi = nullptr, *descmap = nullptr, *sta = nullptr, *heat = nullptr, *nmsdbg = nullptr;
  unsigned long long *cand = nullptr, *scratch = nullptr;
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
  int last_prec = -1;
  // host-API staging
  void* img_dev = nullptr; size_t img_cap = 0;
  float *kp_dev = nullptr, *sc_dev = nullptr, *de_dev = nullptr; int32_t* cnt_dev = nullptr; size_t out_cap = 0;
  // matcher workspace
  unsigned long long *row_key = nullptr, *col_key = nullptr; size_t key_cap = 0;
  unsigned *row2 = nullptr, *col2 = nullptr;   // second-best similarities (ratio tests)
  int* seg_dev = nullptr; size_t seg_cap = 0;  // segment table of the one-to-many call
  __half* mhalf = nullptr; size_t mhalf_cap = 0;
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
  }
  cudaFree(w.nimg); w.nimg = nullptr;
  cudaFree(w.logits); cudaFree(w.semi); cudaFree(w.descmap); cudaFree(w.sta); cudaFree(w.heat); cudaFree(w.nmsdbg);
  cudaFree(w.cand); cudaFree(w.scratch); cudaFree(w.counter); cudaFree(w.status);
  w.logits = w.semi = w.descmap = w.sta = w.heat = w.nmsdbg = nullptr;
  w.cand = w.scratch = nullptr; w.counter = w.status = nullptr;
  w.wsH = w.wsW = 0; w.have_f32 = w.have_tc = false;
}

static int ensure_workspace(const sfd2_ctx* c, Ws& w, int H, int W, int prec) {
  if (w.wsH != H || w.wsW != W) {
    free_workspace(w);
    w.wsH = H; w.wsW = W;
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
    const size_t n8 = (size_t)w.H8 * w.W8, n4 = (size_t)w.H4 * w.W4;
    SFD2_CUDA(cudaMalloc(&w.nimg, (size_t)H * W * sizeof(float4)));
    SFD2_CUDA(cudaMalloc(&w.logits, n8 * 80 * sizeof(float)));
    SFD2_CUDA(cudaMalloc(&w.semi, n8 * 64 * sizeof(float)));
    SFD2_CUDA(cudaMalloc(&w.descmap, n4 * 128 * sizeof(float)));
    SFD2_CUDA(cudaMalloc(&w.sta, n4 * 3 * sizeof(float)));
    SFD2_CUDA(cudaMalloc(&w.heat, (size_t)H * W * sizeof(float)));
    SFD2_CUDA(cudaMalloc(&w.nmsdbg, (size_t)H * W * sizeof(float)));
    // NMS survivors are >= 5 px apart except on exact plateaus: H*W/16 leaves 1.5x headroom.
    w.cap = (int)(((size_t)H * W) / 16) + 4096;
    int cap2 = 1;
    while (cap2 < w.cap) cap2 <<= 1;
    SFD2_CUDA(cudaMalloc(&w.cand, (size_t)w.cap * sizeof(unsigned long long)));
    SFD2_CUDA(cudaMalloc(&w.scratch, (size_t)cap2 * sizeof(unsigned long long)));
    SFD2_CUDA(cudaMemset(w.scratch, 0, (size_t)cap2 * sizeof(unsigned long long)));   // select_kernel's rank / arrival counters
    SFD2_CUDA(cudaMalloc(&w.counter, sizeof(int)));
    SFD2_CUDA(cudaMalloc(&w.status, sizeof(int)));
    SFD2_CUDA(cudaMemset(w.status, 0, sizeof(int)));
  }
  const bool want_tc = (prec != SFD2_PREC_FP32);
  if (!want_tc && !w.have_f32) {
    for (int i = 0; i < NUM_ACTS; ++i) {
      Act& a = w.acts[i];
      SFD2_CUDA(cudaMalloc(&a.f32, a.elems() * sizeof(float)));
      SFD2_CUDA(cudaMemset(a.f32, 0, a.elems() * sizeof(float)));
    }
    w.have_f32 = true;
  }
  if (want_tc && !w.have_tc) {
    for (int i = 0; i < NUM_ACTS; ++i) {
      Act& a = w.acts[i];
      SFD2_CUDA(cudaMalloc(&a.hi, a.elems() * sizeof(__half)));
      SFD2_CUDA(cudaMalloc(&a.lo, a.elems() * sizeof(__half)));
      SFD2_CUDA(cudaMemset(a.hi, 0, a.elems() * sizeof(__half)));
      SFD2_CUDA(cudaMemset(a.lo, 0, a.elems() * sizeof(__half)));
      int rc = tc_make_act_maps(a, a.hi, &w.maps[i][0], &w.maps[i][2], &w.maps[i][4], &w.maps[i][6]);
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
    int rc = 0;
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
    w.have_tc = true;
  }
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

// one image through network + post-processing, all on `st`
static int extract_one(sfd2_ctx* c, Ws& w, const void* img, int img_dtype, int H, int W, const sfd2_extract_params* p,
                       float* kpts, float* scores, float* desc, int32_t* count, cudaStream_t st) {
  const int prec = p->precision;
  const bool tc = prec != SFD2_PREC_FP32;
  const int split = (prec == SFD2_PREC_TC_EXACT || prec == SFD2_PREC_TC_MIXED) ? 3 : 1;
  // MIXED: the descriptor head only has to meet the 1e-3 tolerance, so it runs single-pass (hi planes only)
  const int split_d = (prec == SFD2_PREC_TC_MIXED) ? 1 : split;
  Act* A = w.acts;
  int rc;
#define RUN(x) do { rc = (x); if (rc) return rc; } while (0)
#define RUNP(label, x) do { prof_begin(c, label, st); rc = (x); prof_end(c, st); if (rc) return rc; } while (0)
  RUNP("conv1a", launch_conv1a(img, img_dtype, H, W, c->L("conv1a"), A[A1A], tc ? split : 0, w.nimg, tc ? w.map_1a : nullptr, c->num_sms, st));
  // ConvSta rides in the epilogue of the layer that produces out4 (tcgen05 modes)
  const bool fuse_sta = tc && p->use_stability && g_fuse_sta;
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
  RUN(conv("conv1b", A1A, A1B, -1));
  RUN(conv("conv2a", A1B, A2A, -1));
  RUN(conv("conv2b", A2A, A2B, -1));
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
    RUNP("tc_conv:headD", launch_conv_tc(A[DA], c->L("headD"), desc_act, nullptr, w.map_desc, split_d, c->num_sms, st, 1));
  } else {
    RUNP("conv_f32:headP", launch_conv_simt(A[PA], c->L("headP"), logit_act, nullptr, st));
    RUNP("conv_f32:headD", launch_conv_simt(A[DA], c->L("headD"), desc_act, nullptr, st));
  }
  if (!tc) {
    RUNP("softmax65", launch_softmax65(w.logits, w.H8 * w.W8, w.semi, st));
    RUNP("l2norm128", launch_l2norm128(w.descmap, w.H4 * w.W4, st));
  }
  if (p->use_stability && !fuse_sta) RUNP("sta", launch_sta(A[BA], tc ? (split == 3 ? 1 : 2) : 0, c->L("sta"), w.sta, st));
  RUNP("heat", launch_heat(w.semi, w.H8, w.W8, w.sta, w.H4, w.W4, p->use_stability, w.heat, H, W, st));
  RUNP("nms", launch_nms(w.heat, H, W, p->conf_th, p->border, p->border_w > 0 ? p->border_w : W, p->border_h > 0 ? p->border_h : H,
                         (c->debug_flags & 1) ? w.nmsdbg : nullptr, w.cand, w.cap,
                 w.counter, st));
  RUNP("select", launch_select(w.cand, w.cap, w.counter, W, p->topk, kpts, scores, count, w.status, w.scratch, st));
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
  cudaFree(c->row_key); cudaFree(c->col_key); cudaFree(c->row2); cudaFree(c->col2); cudaFree(c->seg_dev); cudaFree(c->mhalf); cudaFree(c->m_d0); cudaFree(c->m_d1);
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
                         float* kpts, float* scores, float* desc, int32_t* counts, void* stream, const cudaEvent_t* ready) {
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
    rc = ensure_workspace(c, c->ws[k], h, w, p->precision);
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
  for (int i = 0; i < n; ++i) {
    const int k = (ns > 1) ? (i % ns) : 0;
    if (ready) SFD2_CUDA(cudaStreamWaitEvent(ns > 1 ? c->aux[k] : st, ready[i], 0));
    rc = extract_one(c, c->ws[k], static_cast<const uint8_t*>(img) + i * img_stride, img_dtype, h, w, p,
                     kpts + (size_t)i * p->topk * 2, scores + (size_t)i * p->topk,
                     desc + (size_t)i * p->topk * SFD2_DESC_DIM, counts + i, ns > 1 ? c->aux[k] : st);
    if (rc) return rc;
  }
  if (ns > 1)
    for (int k = 0; k < ns; ++k) {
      SFD2_CUDA(cudaEventRecord(c->ev_join[k], c->aux[k]));
      SFD2_CUDA(cudaStreamWaitEvent(st, c->ev_join[k], 0));
    }
  c->launches += g_launches - before;
  c->last_prec = p->precision;
  return SFD2_OK;
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
  SFD2_CUDA(cudaMemsetAsync(c->kp_dev, 0, rows * 2 * sizeof(float), st));
  SFD2_CUDA(cudaMemsetAsync(c->sc_dev, 0, rows * sizeof(float), st));
  if (n == 1) {
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
  int status[sfd2_ctx::kMaxStreams] = {};
  for (int k = 0; k < sfd2_ctx::kMaxStreams; ++k)
    if (c->ws[k].status) SFD2_CUDA(cudaMemcpyAsync(&status[k], c->ws[k].status, sizeof(int), cudaMemcpyDeviceToHost, st));
  SFD2_CUDA(cudaStreamSynchronize(st));
  if (status[0] | status[1] | status[2] | status[3]) {
    for (int k = 0; k < sfd2_ctx::kMaxStreams; ++k)
      if (c->ws[k].status) cudaMemset(c->ws[k].status, 0, sizeof(int));
    set_error("NMS produced more candidates than the workspace holds (cap %d); results truncated", c->ws[0].cap);
    return SFD2_ERR_OVERFLOW;
  }
  return SFD2_OK;
}

static int ensure_match_ws(sfd2_ctx* c, int n0, int n1) {
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
  const size_t hneed = 2 * ((size_t)round_up(n0 > 0 ? n0 : 1, 128) + round_up(n1 > 0 ? n1 : 1, 128)) * 128;
  if (hneed > c->mhalf_cap) {
    cudaFree(c->mhalf); c->mhalf = nullptr; c->mhalf_cap = 0;
    SFD2_CUDA(cudaMalloc(&c->mhalf, hneed * sizeof(__half)));
    c->mhalf_cap = hneed;
  }
  return SFD2_OK;
}

static int match_one(sfd2_ctx* c, const float* d0, int n0, const float* d1, int n1, int d, const sfd2_match_params* p,
                     int32_t* matches0, float* sim0, cudaStream_t st) {
  int rc;
  prof_begin(c, p->precision == SFD2_PREC_FP32 ? "match_simt" : "match_tc", st);
  if (p->precision == SFD2_PREC_FP32)
    rc = launch_match_simt(d0, n0, d1, n1, d, c->row_key, c->col_key, st);
  else
    rc = launch_match_tc(d0, n0, d1, n1, d, p->precision != SFD2_PREC_TC_FAST ? 3 : 1, c->mhalf, c->row_key,
                         c->col_key, c->num_sms, st);
  prof_end(c, st);
  if (rc) return rc;
  if (p->ratio_threshold > 0.f) {   // second-best pass (CUDA-core fp32 in every precision mode)
    rc = launch_match_second(d0, n0, d1, n1, d, c->row_key, c->col_key, c->row2, c->col2, st);
    if (rc) return rc;
  }
  return launch_match_finish(c->row_key, c->col_key, n0, n1, p->do_mutual_check, p->distance_threshold,
                             p->ratio_threshold, ((p->ratio_mode & 0xFF) == 1 ? 2 : 1) | (p->ratio_mode & SFD2_MATCH_PLAIN_CODES),
                             c->row2, c->col2, matches0, sim0, st);
}

SFD2_API int sfd2_match_dev(sfd2_ctx* c, const float* d0, int n0, const float* d1, int n1, int d, const sfd2_match_params* p,
                   int32_t* matches0, float* sim0, void* stream) {
  SFD2_CHECK(c && p && (n0 == 0 || (d0 && matches0 && sim0)) && (n1 == 0 || d1), SFD2_ERR_ARG, "sfd2_match_dev: NULL argument");
  SFD2_CHECK(n0 >= 0 && n1 >= 0 && d >= 1, SFD2_ERR_ARG, "sfd2_match_dev: bad shape %d x %d x %d", n0, n1, d);
  SFD2_CHECK(p->precision >= 0 && p->precision <= 3, SFD2_ERR_ARG, "bad precision %d", p->precision);
  SFD2_CUDA(cudaSetDevice(c->device));
  int rc = ensure_match_ws(c, n0, n1);
  if (rc) return rc;
  const long long before = g_launches;
  rc = match_one(c, d0, n0, d1, n1, d, p, matches0, sim0, static_cast<cudaStream_t>(stream));
  c->launches += g_launches - before;
  return rc;
}

SFD2_API int sfd2_match_batched_dev(sfd2_ctx* c, const float* d0, const int32_t* off0, const float* d1, const int32_t* off1,
                           int npairs, int d, const sfd2_match_params* p, int32_t* matches0, float* sim0, void* stream) {
  SFD2_CHECK(c && p && off0 && off1 && npairs >= 0, SFD2_ERR_ARG, "sfd2_match_batched_dev: bad argument");
  SFD2_CUDA(cudaSetDevice(c->device));
  int mx0 = 0, mx1 = 0;
  for (int i = 0; i < npairs; ++i) {
    SFD2_CHECK(off0[i + 1] >= off0[i] && off1[i + 1] >= off1[i], SFD2_ERR_ARG, "offsets must be non-decreasing");
    mx0 = std::max(mx0, off0[i + 1] - off0[i]);
    mx1 = std::max(mx1, off1[i + 1] - off1[i]);
  }
  int rc = ensure_match_ws(c, mx0, mx1);
  if (rc) return rc;
  const long long before = g_launches;
  for (int i = 0; i < npairs && !rc; ++i)
    rc = match_one(c, d0 + (size_t)off0[i] * d, off0[i + 1] - off0[i], d1 + (size_t)off1[i] * d, off1[i + 1] - off1[i],
                   d, p, matches0 + off0[i], sim0 + off0[i], static_cast<cudaStream_t>(stream));
  c->launches += g_launches - before;
  return rc;
}

SFD2_API int sfd2_match_one_to_many_dev(sfd2_ctx* c, const float* q, int nq, const float* db, const int32_t* db_off,
                                        int ndb, int d, const sfd2_match_params* p, int32_t* matches0, float* sim0,
                                        void* stream) {
  SFD2_CHECK(c && p && q && db && db_off && matches0 && sim0, SFD2_ERR_ARG, "sfd2_match_one_to_many_dev: NULL argument");
  SFD2_CHECK(nq >= 1 && ndb >= 1 && d == 128, SFD2_ERR_ARG, "sfd2_match_one_to_many_dev: bad shape (d must be 128)");
  SFD2_CHECK(p->ratio_threshold <= 0.f, SFD2_ERR_ARG, "ratio tests are not available in the grouped call: use sfd2_match_batched_dev");
  SFD2_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->precision == SFD2_PREC_FP32) {   // CUDA-core mode: plain loop
    int rc = SFD2_OK;
    for (int i = 0; i < ndb && !rc; ++i)
      rc = sfd2_match_dev(c, q, nq, db + (size_t)db_off[i] * d, db_off[i + 1] - db_off[i], d, p, matches0 + (size_t)i * nq,
                          sim0 + (size_t)i * nq, stream);
    return rc;
  }
  std::vector<int> seg(3 * (ndb + 1));
  int P1 = 0;
  for (int i = 0; i < ndb; ++i) {
    SFD2_CHECK(db_off[i + 1] >= db_off[i], SFD2_ERR_ARG, "offsets must be non-decreasing");
    seg[i] = db_off[i];
    seg[(ndb + 1) + i] = P1;
    seg[2 * (ndb + 1) + i] = db_off[i + 1] - db_off[i];
    P1 += round_up(std::max(db_off[i + 1] - db_off[i], 1), 128);
  }
  seg[ndb] = db_off[ndb];
  seg[(ndb + 1) + ndb] = P1;
  // workspace: keys [max(nq*ndb, P1)], split planes [(nq_pad + P1) * 2 * 128] halves, segment table
  const size_t need_keys = std::max((size_t)nq * ndb, (size_t)P1) + 128;
  if (need_keys > c->key_cap) {
    cudaFree(c->row_key); cudaFree(c->col_key); cudaFree(c->row2); cudaFree(c->col2);
    c->row_key = c->col_key = nullptr; c->row2 = c->col2 = nullptr; c->key_cap = 0;
    SFD2_CUDA(cudaMalloc(&c->row_key, need_keys * sizeof(unsigned long long)));
    SFD2_CUDA(cudaMalloc(&c->col_key, need_keys * sizeof(unsigned long long)));
    SFD2_CUDA(cudaMalloc(&c->row2, need_keys * sizeof(unsigned)));
    SFD2_CUDA(cudaMalloc(&c->col2, need_keys * sizeof(unsigned)));
    c->key_cap = need_keys;
  }
  const size_t hneed = 2 * ((size_t)round_up(nq, 128) + P1) * 128;
  if (hneed > c->mhalf_cap) {
    cudaFree(c->mhalf); c->mhalf = nullptr; c->mhalf_cap = 0;
    SFD2_CUDA(cudaMalloc(&c->mhalf, hneed * sizeof(__half)));
    c->mhalf_cap = hneed;
  }
  if (seg.size() > c->seg_cap) {
    cudaFree(c->seg_dev); c->seg_dev = nullptr; c->seg_cap = 0;
    SFD2_CUDA(cudaMalloc(&c->seg_dev, seg.size() * sizeof(int)));
    c->seg_cap = seg.size();
  }
  SFD2_CUDA(cudaMemcpyAsync(c->seg_dev, seg.data(), seg.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  const long long before = g_launches;
  const int rc = launch_match_one_to_many(q, nq, db, c->seg_dev, ndb, P1, p->precision != SFD2_PREC_TC_FAST ? 3 : 1,
                                          p->do_mutual_check, p->distance_threshold, c->mhalf, c->row_key, c->col_key,
                                          matches0, sim0, c->num_sms, st);
  c->launches += g_launches - before;
  return rc;
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
  SFD2_CUDA(cudaMalloc(&cand, (size_t)cap * 8));
  SFD2_CUDA(cudaMalloc(&scratch, (size_t)cap2 * 8));
  SFD2_CUDA(cudaMalloc(&counter, 4));
  SFD2_CUDA(cudaMalloc(&status, 4));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SFD2_CUDA(cudaMemsetAsync(status, 0, 4, st));
  SFD2_CUDA(cudaMemsetAsync(scratch, 0, (size_t)cap2 * 8, st));
  const long long before = g_launches;
  int rc = launch_nms(heat, h, w, p->conf_th, p->border, p->border_w > 0 ? p->border_w : w, p->border_h > 0 ? p->border_h : h,
                      nms_out, cand, cap, counter, st);
  if (!rc) rc = launch_select(cand, cap, counter, w, p->topk, kpts, scores, count, status, scratch, st);
  c->launches += g_launches - before;
  int hstatus = 0;
  cudaMemcpyAsync(&hstatus, status, 4, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
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
