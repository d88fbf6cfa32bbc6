// Shared declarations for libsfd2_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/sfd2_b200.h"

namespace sfd2 {

void set_error(const char* fmt, ...);

#define SFD2_CUDA(call)                                                                 \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      sfd2::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return SFD2_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)

#define SFD2_CHECK(cond, code, ...)   \
  do {                                \
    if (!(cond)) {                    \
      sfd2::set_error(__VA_ARGS__);   \
      return (code);                  \
    }                                 \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return cdiv(a, b) * b; }
static inline int conv_out(int n, int stride) { return (n - 1) / stride + 1; }  // k=3,p=1 or k=1,p=0

// An activation map in HBM, channels-last.  Pixel (y, x) starts at (y*Wp + x)*C.
// Hp/Wp are the allocated (even) extents; the padding rows/cols stay zero so that a
// stride-2 consumer can view the buffer as [Hp/2][2][Wp/2][2][C] for TMA.
struct Act {
  float* f32 = nullptr;  // FP32 mode
  __half* hi = nullptr;  // TC modes: x ~ hi (+ lo)
  __half* lo = nullptr;
  int H = 0, W = 0, C = 0, Hp = 0, Wp = 0;
  const CUtensorMap* tm = nullptr;     // [s1_hi, s1_lo, s2_hi, s2_lo, halo_hi, halo_lo, t8x16_hi, t8x16_lo] TMA load views
  const CUtensorMap* tm_st = nullptr;  // [hi, lo] 16x2-pixel boxes, [hi, lo] 8x4-pixel boxes: TMA store views
  size_t elems() const { return (size_t)Hp * Wp * C; }
};

// One folded convolution (BatchNorm already multiplied in on the host).
struct Layer {
  std::string name;
  int cin = 0, cout = 0, k = 0, stride = 1, groups = 1, relu = 0;
  int cout_pad = 0;            // SIMT: multiple of 64; TC: multiple of 16
  std::vector<float> w;        // OIHW [cout][cin/groups][k][k]
  std::vector<float> b;        // [cout]
  float* w_simt = nullptr;     // [k*k][cin/groups][cout_pad64] fp32
  float* b_dev = nullptr;      // [cout_pad64] fp32
  // tcgen05 operands: K-major rows of 64 input channels.
  //  dense : row = tap*cout_tc + o           , col = ci            -> [k*k*cout_tc][cin]
  //  diag  : row = tap*256 + o (groups=32)   , col = ci - 64*(o/64)-> [9*256][64]
  __half* w_hi = nullptr;
  __half* w_lo = nullptr;
  __half* w_cat = nullptr;     // grouped layers: [tap][chunk][w_hi 64 rows | w_lo 64 rows][64] (TcConvArgs::cat)
  CUtensorMap tm_w_cat, tm_w_cat_half;
  int cout_tc = 0;             // rows per tap in the TC packing (multiple of 16)
  CUtensorMap tm_w_hi, tm_w_lo;
  CUtensorMap tm_w_hi_half, tm_w_lo_half;  // boxes of n_mma/2 rows (2-CTA multicast)
  CUtensorMap tm_w_hi_quarter, tm_w_lo_quarter;
};

struct TcConvLaunch;  // tc_conv.cu

// ---- kernels' host-side launchers (each returns a sfd2_status) -------------------
// simt_conv.cu
// [y_begin, y_end): output rows of this launch (y_end < 0 = H); row bands need the tcgen05 kernel (conv1a_bands_ok)
int launch_conv1a(const void* img, int img_dtype, int H, int W, const Layer& L, Act out, int tc_out, float4* nimg,
                  const CUtensorMap* tm1a, int num_sms, cudaStream_t st, int y_begin = 0, int y_end = -1);
bool conv1a_bands_ok(int tc_out);
// tc_conv1a.cu
int conv1a_mma_encode(Layer& L);
int launch_conv1a_mma(const void* img, int img_dtype, int H, int W, const Layer& L, const CUtensorMap* tm1a, int split,
                      int num_sms, cudaStream_t st, int y_begin = 0, int y_end = -1);
int launch_conv_simt(const Act& in, const Layer& L, Act out, const Act* res, cudaStream_t st);
// tc_in: 0 = fp32 input, 1 = fp16 hi+lo, 2 = fp16 hi only
int launch_sta(const Act& in, int tc_in, const Layer& L, float* logits, cudaStream_t st);
int launch_softmax65(const float* logits, int npix, float* semi, cudaStream_t st);
int launch_l2norm128(float* desc, int npix, cudaStream_t st);
// tc_conv.cu
int tc_encode_weights(Layer& L);
int tc_make_act_maps(const Act& t, const __half* base, CUtensorMap* s1, CUtensorMap* s2, CUtensorMap* halo,
                     CUtensorMap* t8x16 = nullptr);
int tc_make_store_map(CUtensorMap* tm, const void* base, int C, int W, int H, int Wp, int is_f32, int box_w);
// epi_fn (fp32 outputs only): 0 raw, 1 L2-normalised channels, 2 exp-normalised (softmax-with-eps) channels 0..63
// sta / sta_out: fuse ConvSta (1x1 256 -> 3) on this layer's output into the epilogue (fp16-plane outputs only)
// rows_done / rows_avail (row-band launches behind a banded upload): compute only the whole TILE rows inside output rows
// [*rows_done, rows_avail) - up to the last row when rows_avail >= out.H - and advance *rows_done; nothing to do = no launch
int launch_conv_tc(const Act& in, const Layer& L, Act out, const Act* res, const CUtensorMap* out_f32_map, int split,
                   int num_sms, cudaStream_t st, int epi_fn = 0, const Layer* sta = nullptr, float* sta_out = nullptr,
                   int* rows_done = nullptr, int rows_avail = 0);
// tc_desc_sparse.cu
int launch_desc_sparse(const Act& in, const Layer& L, int H, int W, const float* kpts, const int32_t* count, int topk,
                       float* rows, float* desc_out, cudaStream_t st);
// post.cu
int launch_heat(const float* semi, int H8, int W8, const float* sta, int H4, int W4, int use_sta, float* heat,
                int H, int W, cudaStream_t st);
int launch_nms(const float* heat, int H, int W, float conf_th, int border, int bw, int bh, float* nms_out,
               unsigned long long* cand, int cap, int* counter, cudaStream_t st, bool zero_counter = true);
int launch_select(unsigned long long* cand, int cap, const int* counter, int W, int topk, float* kpts,
                  float* scores, int32_t* count_out, int* status, unsigned long long* scratch,
                  cudaStream_t st);
int launch_sample(const float* desc_map, int H4, int W4, int H, int W, const float* kpts, const int32_t* count,
                  int topk, float* desc_out, cudaStream_t st);
// preprocess.cu
int launch_preprocess(const uint8_t* src, int h, int w, int swap_rb, int hn, int wn, float* out, cudaStream_t st);
// match.cu
int launch_match_simt(const float* d0, int n0, const float* d1, int n1, int d, unsigned long long* row_key,
                      unsigned long long* col_key, cudaStream_t st);
int launch_match_second(const float* d0, int n0, const float* d1, int n1, int d, const unsigned long long* row_key,
                        const unsigned long long* col_key, unsigned* row2, unsigned* col2, cudaStream_t st);
int launch_match_finish(const unsigned long long* row_key, const unsigned long long* col_key, int n0, int n1,
                        int mutual, float dist_th, float ratio_th, int ratio_mode, const unsigned* row2,
                        const unsigned* col2, int32_t* matches0, float* sim0, cudaStream_t st);

// driver entry point for TMA descriptors, resolved once through cudart (no -lcuda link)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();
// fp16, 128B-swizzled tiled map; dims/strides innermost first (strides[0] implied = 2 bytes)
int make_tmap_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);
int make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int is_f32, int swizzle);

// Launch with programmatic stream serialization (SFD2_POST_PDL=0 disables): the kernel may be scheduled while its
// predecessor in the stream drains; every kernel launched this way executes griddepcontrol.wait before it touches global memory.
extern int g_post_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_post_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
extern int g_sparse_desc, g_tc_slim, g_tc_stiles, g_tc_cg2, g_tc_pdl, g_tc_multicast, g_tc_halo, g_tc_nsplit, g_conv1a_mma, g_fuse_sta, g_tc_diagcat, g_tc_split1x1;
extern thread_local long long g_launches;  // kernels launched by this thread (for sfd2_launch_count)

}  // namespace sfd2
