/* sfd2_b200.h - C ABI of libsfd2_b200.so: the B200 (sm_100a) implementation of the
 * SFD2 feature-extraction + mutual-NN matching hot path.
 *
 * The reference (feixue94/sfd2) is pure Python/PyTorch and has no FFI of its own;
 * each entry point below names the reference Python interface it replaces
 * (paths relative to the reference root).  The Python drop-ins in sfd2_b200/
 * (extractor.py, matchers.py) bind these symbols with ctypes - see INTEGRATION.md
 * for the stub a reference maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, a negative sfd2_status on failure and
 *    never throws; sfd2_last_error() gives the message for the calling thread;
 *  - the CALLER owns every buffer; the library owns its packed weights and a
 *    workspace that is (re)sized when the image size grows;
 *  - *_dev entry points take DEVICE pointers and are asynchronous on `stream`
 *    (a cudaStream_t passed as void*; NULL = default stream): no hidden
 *    synchronisation, no allocation once the workspace exists;
 *  - *_host entry points take HOST pointers, do the H2D/D2H copies themselves and
 *    return when the results are in the host buffers;
 *  - outputs have fixed capacity (topk rows per image) plus a count per image;
 *  - a context is bound to one device and must not be used from two threads at once.
 */
#ifndef SFD2_B200_H
#define SFD2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SFD2_API __attribute__((visibility("default")))
#else
#define SFD2_API
#endif

#define SFD2_ABI_VERSION 2
#define SFD2_DESC_DIM 128
#define SFD2_MATCH_PLAIN_CODES 0x100
#define SFD2_MATCH_HLOC_SCORES 0x200
#define SFD2_MATCH_I64 0x400

typedef struct sfd2_ctx sfd2_ctx;

typedef enum {
  SFD2_OK = 0,
  SFD2_ERR_ARG = -1,       /* bad argument / unsupported shape            */
  SFD2_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed          */
  SFD2_ERR_WEIGHTS = -3,   /* malformed weight blob                        */
  SFD2_ERR_OVERFLOW = -4,  /* more NMS candidates than the workspace holds */
  SFD2_ERR_NOMEM = -5
} sfd2_status;

/* precision of the convolution stack */
typedef enum {
  SFD2_PREC_FP32 = 0,     /* CUDA-core fp32 FMA (reference-grade, slow)                          */
  SFD2_PREC_TC_EXACT = 1, /* tcgen05 fp16 x3 split (a_hi*w_hi + a_hi*w_lo + a_lo*w_hi), fp32 acc */
  SFD2_PREC_TC_FAST = 2,  /* tcgen05 fp16 x1, fp32 accumulate                                    */
  SFD2_PREC_TC_MIXED = 3  /* everything that feeds the heat-map (keypoint indices, scores) as TC_EXACT;
                             the descriptor head (convDa, convDb: 35 % of the MACs) as TC_FAST - its
                             outputs only have to meet the 1e-3 tolerance.  Matcher calls: = TC_EXACT */
} sfd2_precision;

/* image element type for sfd2_extract_* */
typedef enum {
  SFD2_IMG_F32_NCHW = 0, /* float32 [n,3,h,w] RGB in [0,1]  (what ImageDataset yields)  */
  SFD2_IMG_U8_NHWC = 1   /* uint8   [n,h,w,3] RGB; divided by 255 on the device         */
} sfd2_img_dtype;

typedef struct {
  float conf_th;      /* nets/extractor.py:143 conf_thresh (0.001 in every preset)      */
  int32_t nms_radius; /* nets/extractor.py:145 nms_dist = 4 (only 4 is implemented)     */
  int32_t border;     /* nets/extractor.py:146 border_remove = 4                        */
  int32_t topk;       /* extract_localization.py max_keypoints; rows of every output    */
  int32_t precision;  /* sfd2_precision                                                 */
  int32_t use_stability; /* extract_localization.py:30 use_stability                    */
  int32_t border_w;   /* extents the border test uses; 0 = the image's own w / h.  The  */
  int32_t border_h;   /* reference's multi-scale loop tests SCALED coordinates against   */
                      /* the ORIGINAL size (nets/extractor.py:181-182)                   */
} sfd2_extract_params;

typedef struct {
  int32_t do_mutual_check;  /* hloc/matchers/nearest_neighbor.py:31                      */
  float distance_threshold; /* :30 ; <= 0 means None                                     */
  float ratio_threshold;    /* :29 ; <= 0 means None                                     */
  int32_t precision;        /* sfd2_precision (FP32 = CUDA cores, TC_* = tcgen05)        */
  int32_t ratio_mode;       /* bits 0..7: 0 = hloc find_nn formula (nearest_neighbor.py:10-11),
                               1 = it_loc mutual_nn_ratio_matcher formula (it_loc/matcher.py:172-174);
                               bit 8 (SFD2_MATCH_PLAIN_CODES): report every unmatched row as -1
                               (no -2 for "rejected only by the mutual check");
                               bit 9 (SFD2_MATCH_HLOC_SCORES, tcgen05 modes): sim0 = hloc's matching_scores0, i.e.
                               (sim + 1) / 2 where the row passed its own ratio / distance test, else 0 (:14-15);
                               bit 10 (SFD2_MATCH_I64, tcgen05 modes): matches0 is an int64 buffer (hloc's dtype) */
  int32_t layout;           /* sfd2_desc_layout of d0 / d1 in sfd2_match_dev / sfd2_match_host */
} sfd2_match_params;

/* memory layout of a descriptor set */
typedef enum {
  SFD2_DESC_ROWS = 0, /* float32 [n, 128] row-major  (it_loc/matcher.py:94-95, the extractor's output)        */
  SFD2_DESC_COLS = 1  /* float32 [128, n]            (hloc: data['descriptors0'][0], nearest_neighbor.py:39)  */
} sfd2_desc_layout;

/* One descriptor set of a grouped match call (all pointers are DEVICE pointers). */
typedef struct {
  const float* data;    /* descriptors in `layout`                                                                 */
  int32_t n;            /* number of descriptors; the CAPACITY of the set when `count` is given                    */
  int32_t layout;       /* sfd2_desc_layout                                                                        */
  const int32_t* count; /* optional: the set's valid row count lives in device memory (e.g. sfd2_extract_dev's     */
                        /* counts[i]) - an extract -> match pipeline then never synchronises with the host         */
  const int32_t* ids;   /* optional int32 [n]: rows with ids[r] == -1 take no part (the localizer matches only db   */
                        /* keypoints that have a 3-D point: desc_db[db_3D_ids != -1], it_loc/localize_cv2.py:540-555); */
                        /* matches are reported as ORIGINAL row indices (the remap of :557-559 happens on the device) */
} sfd2_desc_set;

SFD2_API int sfd2_abi_version(void);
SFD2_API const char* sfd2_last_error(void);

/* Replaces get_model(...)[0] + model.cuda()  (extract_localization.py:208-218,227).
 * `blob` is the folded weight blob made by sfd2_b200/weights.py (host memory); blob == NULL creates a
 * matcher-only context (the matcher entry points need no network weights). */
SFD2_API int sfd2_create(const void* blob, size_t nbytes, int device, sfd2_ctx** out);
SFD2_API int sfd2_destroy(sfd2_ctx* ctx);

/* Replaces extract_resnet_return(model, img, conf_th, mask=None, topK, scales=[1.0])
 * (nets/extractor.py:97-337): normalise -> ResSegNetV2.det (nets/sfd2.py:313-354)
 * -> x stability -> simple_nms(4) -> >conf_th -> border -> sort -> bilinear
 * descriptor sampling + L2 -> top-K.
 *   img      [n] images of dtype `img_dtype`, h x w each
 *   kpts     float32 [n, topk, 2]  (x, y) pixel indices, score-descending
 *   scores   float32 [n, topk]
 *   desc     float32 [n, topk, 128] unit-norm rows
 *   counts   int32   [n] valid rows per image (<= topk); rows beyond are zero */
SFD2_API int sfd2_extract_dev(sfd2_ctx* ctx, const void* img_dev, int img_dtype, int n, int h, int w,
                     const sfd2_extract_params* p, float* kpts_dev, float* scores_dev,
                     float* desc_dev, int32_t* counts_dev, void* stream);
SFD2_API int sfd2_extract_host(sfd2_ctx* ctx, const void* img_host, int img_dtype, int n, int h, int w,
                      const sfd2_extract_params* p, float* kpts_host, float* scores_host,
                      float* desc_host, int32_t* counts_host);
/* The *_dev extract call is asynchronous, so it cannot report that NMS produced more candidates than the workspace
 * holds (the results are then a truncated, order-dependent subset).  This query synchronises `stream`, returns
 * SFD2_ERR_OVERFLOW if any extract since the last query overflowed, SFD2_OK otherwise, and clears the flag.
 * sfd2_extract_host calls it itself.  (The reference has no such limit: nets/extractor.py:158 uses nonzero().) */
SFD2_API int sfd2_extract_status(sfd2_ctx* ctx, void* stream);

/* Replaces the arithmetic of ImageDataset.__getitem__ (extract_localization.py:158-190): decoded uint8 image [h, w, 3]
 * (DEVICE pointer; swap_rb = 1 for cv2.imread's BGR order) -> float32 RGB [3, h_new, w_new] in the reference's value
 * range (cv2.resize INTER_CUBIC on the float image, then / 255; no clamping).  h_new == h && w_new == w: conversion only.
 * The output is what sfd2_extract_dev takes as SFD2_IMG_F32_NCHW.  Asynchronous on `stream`. */
SFD2_API int sfd2_preprocess_dev(sfd2_ctx* ctx, const uint8_t* img_u8_dev, int h, int w, int swap_rb, int h_new, int w_new,
                                 float* out_dev, void* stream);

/* Replaces NearestNeighbor._forward (hloc/matchers/nearest_neighbor.py:38-57) and
 * Matcher.mutual_nn_matcher (it_loc/matcher.py:122-130).
 *   d0 float32 [n0, d] row-major, d1 float32 [n1, d]   (d = 128)
 *   matches0 int32 [n0]  index into d1; -1 = no candidate / rejected by the row's own ratio or distance
 *                        test; -2 = rejected only by the mutual check (hloc keeps that row's score)
 *   sim0     float32 [n0] raw max cosine of every row (callers map to (s+1)/2)  */
SFD2_API int sfd2_match_dev(sfd2_ctx* ctx, const float* d0_dev, int n0, const float* d1_dev, int n1, int d,
                   const sfd2_match_params* p, int32_t* matches0_dev, float* sim0_dev, void* stream);
SFD2_API int sfd2_match_host(sfd2_ctx* ctx, const float* d0_host, int n0, const float* d1_host, int n1, int d,
                    const sfd2_match_params* p, int32_t* matches0_host, float* sim0_host);
/* Many pairs in ONE grouped launch: pair k matches sets[pair_a[k]] (rows of the result) against sets[pair_b[k]].
 * Replaces the per-pair loops hloc/match_features.py:90-121 and it_loc/localize_cv2.py:511-560,705-731.
 *   matches0 int32, sim0 float32: pair k's rows start at sum_{q<k} sets[pair_a[q]].n (capacity rows; rows beyond a
 *   device-side count read -1 / 0).  Codes as in sfd2_match_dev.  pair_a / pair_b / sets are HOST arrays. */
SFD2_API int sfd2_match_pairs_dev(sfd2_ctx* ctx, const sfd2_desc_set* sets, int nsets, const int32_t* pair_a,
                                  const int32_t* pair_b, int npairs, const sfd2_match_params* p,
                                  int32_t* matches0_dev, float* sim0_dev, void* stream);
/* Many independent pairs in one call (hloc/match_features.py:90 pair loop;
 * it_loc/localize_cv2.py:705 one query against <= 50 db images).  Pair i uses rows
 * [off0[i], off0[i+1]) of d0 and [off1[i], off1[i+1]) of d1; outputs are indexed
 * like d0 and matches are relative to off1[i].  Offsets are HOST arrays. */
SFD2_API int sfd2_match_batched_dev(sfd2_ctx* ctx, const float* d0_dev, const int32_t* off0_host,
                           const float* d1_dev, const int32_t* off1_host, int npairs, int d,
                           const sfd2_match_params* p, int32_t* matches0_dev, float* sim0_dev,
                           void* stream);

/* One query set against many db sets in ONE grouped launch (it_loc/localize_cv2.py:705-731: a query against
 * its <= 50 retrieved db images, each db set filtered by db_3D_ids != -1, :540-555).  db holds the db
 * descriptor sets back to back, db_off_host[ndb+1] their row offsets (HOST array).
 *   matches0 int32 [ndb, nq]  index into db set i (local), -1 / -2 as in sfd2_match_dev
 *   sim0     float32 [ndb, nq] */
SFD2_API int sfd2_match_one_to_many_dev(sfd2_ctx* ctx, const float* q_dev, int nq, const float* db_dev,
                                        const int32_t* db_off_host, int ndb, int d, const sfd2_match_params* p,
                                        int32_t* matches0_dev, float* sim0_dev, void* stream);

/* Test / profiling hooks (not part of the reference surface). */
/* Copy an intermediate of the LAST extracted image to host as float32:
 * "heat" [h,w], "nms" [h,w], "score" [h8*8? see DESIGN.md], "desc_map" [h4,w4,128],
 * "sta_logits" [h4,w4,3], "out4" [h4,w4,256], "conv1a".."conv3b" NHWC.  Returns the
 * number of floats written or a negative status. */
SFD2_API long long sfd2_debug_fetch(sfd2_ctx* ctx, const char* name, float* out_host, long long capacity);
/* Number of kernels this library launched on behalf of ctx since creation. */
SFD2_API long long sfd2_launch_count(sfd2_ctx* ctx);
/* Per-launch CUDA-event timing on the launching stream (bench.py's roofline numbers):
 * sfd2_profile(ctx, 1) starts recording; sfd2_profile_read synchronises and writes
 * "label\tlaunches\ttotal_ms\n" lines for everything recorded since the last read. */
SFD2_API int sfd2_profile(sfd2_ctx* ctx, int enable);
SFD2_API long long sfd2_profile_read(sfd2_ctx* ctx, char* buf, long long capacity);
/* Standalone NMS + selection on a caller-supplied heat-map (device fp32 [h,w]):
 * the part of the path that is compare-only and therefore bit-exact. */
SFD2_API int sfd2_nms_select_dev(sfd2_ctx* ctx, const float* heat_dev, int h, int w,
                        const sfd2_extract_params* p, float* kpts_dev, float* scores_dev,
                        int32_t* count_dev, float* nms_out_dev /* may be NULL */, void* stream);
/* Standalone conv layer on tcgen05 for unit tests: see sfd2_b200/csrc/api.cu. */
SFD2_API int sfd2_debug_conv(sfd2_ctx* ctx, const float* x_host, int h, int w, int cin, const float* w_host,
                    const float* b_host, int cout, int ksize, int stride, int groups, int relu,
                    int precision, float* y_host);

/* (The two hardware probes used while designing the conv kernel - tools/umma_probe.py, tools/mma_rate_probe.py -
 * are NOT part of this ABI: they are compiled in only by `SFD2_WITH_PROBES=1 python -m sfd2_b200.build --force` and
 * declared in sfd2_b200/csrc/probes.h.) */

#ifdef __cplusplus
}
#endif
#endif /* SFD2_B200_H */
