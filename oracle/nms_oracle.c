/* CPU ORACLE (plain C) for the compare-only stages of the SFD2 hot path.
 * TEST INFRASTRUCTURE ONLY - never linked into or called by the product (sfd2_b200/).
 *
 * Restates, without any library:
 *   simple_nms(scores, 4)          nets/extractor.py:20-35   (3 rounds, 9x9 windows, -inf padding)
 *   threshold / border / sort / K  nets/extractor.py:158-183, :322-326
 *   mutual nearest neighbour       hloc/matchers/nearest_neighbor.py:6-24, it_loc/matcher.py:122-130
 * Pinned by tests/test_oracle_c.py against the reference-generated fixtures in tests/golden/
 * (nms_cases.npz, match_cases.npz) and against the PyTorch restatement in sfd2_oracle.py.
 * Tie rule (the reference's is unspecified): higher score first, then lower pixel / column index.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void maxpool(const float* s, int H, int W, int r, float* out) {
  /* F.max_pool2d(kernel 2r+1, stride 1, padding r): out-of-image taps are skipped (-inf) */
  float* tmp = (float*)malloc(sizeof(float) * (size_t)H * W);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float m = -INFINITY;
      for (int k = x - r; k <= x + r; ++k)
        if (k >= 0 && k < W && s[(size_t)y * W + k] > m) m = s[(size_t)y * W + k];
      tmp[(size_t)y * W + x] = m;
    }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float m = -INFINITY;
      for (int k = y - r; k <= y + r; ++k)
        if (k >= 0 && k < H && tmp[(size_t)k * W + x] > m) m = tmp[(size_t)k * W + x];
      out[(size_t)y * W + x] = m;
    }
  free(tmp);
}

void sfd2o_nms(const float* s, int H, int W, int r, float* out) {
  const size_t n = (size_t)H * W;
  float* mp = (float*)malloc(sizeof(float) * n);
  float* mask = (float*)malloc(sizeof(float) * n);
  float* supp = (float*)malloc(sizeof(float) * n);
  float* ss = (float*)malloc(sizeof(float) * n);
  maxpool(s, H, W, r, mp);
  for (size_t i = 0; i < n; ++i) mask[i] = (s[i] == mp[i]) ? 1.f : 0.f;       /* max_mask */
  for (int round = 0; round < 2; ++round) {
    maxpool(mask, H, W, r, supp);                                             /* supp_mask = pool(mask) > 0 */
    for (size_t i = 0; i < n; ++i) ss[i] = (supp[i] > 0.f) ? 0.f : s[i];      /* supp_scores */
    maxpool(ss, H, W, r, mp);
    for (size_t i = 0; i < n; ++i)
      if (ss[i] == mp[i] && !(supp[i] > 0.f)) mask[i] = 1.f;                  /* |= new_max & ~supp */
  }
  for (size_t i = 0; i < n; ++i) out[i] = (mask[i] > 0.f) ? s[i] : 0.f;
  free(mp); free(mask); free(supp); free(ss);
}

typedef struct { float s; int lin; } cand_t;
static int cand_cmp(const void* a, const void* b) {
  const cand_t* p = (const cand_t*)a; const cand_t* q = (const cand_t*)b;
  if (p->s != q->s) return (p->s > q->s) ? -1 : 1;
  return (p->lin < q->lin) ? -1 : (p->lin > q->lin);
}

/* returns the number of keypoints written (<= topk when topk > 0) */
int sfd2o_select(const float* nms, int H, int W, float conf_th, int border, int topk, int* xy, float* sc) {
  cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)H * W);
  int n = 0;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const float v = nms[(size_t)y * W + x];
      if (v > conf_th && x >= border && x < W - border && y >= border && y < H - border) {
        c[n].s = v; c[n].lin = y * W + x; ++n;
      }
    }
  qsort(c, (size_t)n, sizeof(cand_t), cand_cmp);
  if (topk > 0 && n > topk) n = topk;
  for (int i = 0; i < n; ++i) { xy[2 * i] = c[i].lin % W; xy[2 * i + 1] = c[i].lin / W; sc[i] = c[i].s; }
  free(c);
  return n;
}

/* matches0[i] = argmax_j <d0_i, d1_j> if (no mutual check or argmax_i' <d0_i', d1_j> == i) else -1 */
void sfd2o_mutual_nn(const float* d0, int n, const float* d1, int m, int d, int mutual, int* matches0, float* sim0) {
  int* nn21 = (int*)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
  float* best21 = (float*)malloc(sizeof(float) * (size_t)(m > 0 ? m : 1));
  for (int j = 0; j < m; ++j) { nn21[j] = -1; best21[j] = -INFINITY; }
  for (int i = 0; i < n; ++i) {
    int bj = -1; float bs = -INFINITY;
    for (int j = 0; j < m; ++j) {
      float acc = 0.f;
      for (int k = 0; k < d; ++k) acc += d0[(size_t)i * d + k] * d1[(size_t)j * d + k];
      if (acc > bs) { bs = acc; bj = j; }
      if (acc > best21[j]) { best21[j] = acc; nn21[j] = i; }
    }
    matches0[i] = bj; sim0[i] = (bj >= 0) ? bs : 0.f;
  }
  if (mutual)
    for (int i = 0; i < n; ++i)
      if (matches0[i] >= 0 && nn21[matches0[i]] != i) matches0[i] = -1;
  free(nn21); free(best21);
}
