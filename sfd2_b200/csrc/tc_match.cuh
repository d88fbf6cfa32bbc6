// Device-side tables of the grouped tcgen05 matcher (tc_match.cu) and its host launchers.
#pragma once
#include "common.cuh"

namespace sfd2 {

// One descriptor set = one GEMM operand.  Its fp16 hi / lo planes occupy rows [prow0, prow0 + round_up(cap, 128)).
struct MOperD {
  const float* src;       // device fp32 descriptors
  long long rs, cs;       // element (row r, k) at src[r * rs + k * cs]
  int layout;             // 0: [n][128] rows (cs == 1), 1: [128][n] (hloc, rs == 1)
  int cap;                // rows (capacity when `count` is set)
  int prow0;              // first plane row (multiple of 128)
  int pad_;
  const int* count;       // optional DEVICE row count (<= cap)
  const int* ids;         // optional DEVICE int32 [cap]: rows with id == -1 are dropped (order-preserving compaction)
};

// One pair (set a, set b).  Work units [tile0, tile0 + ntiles) of the launch's list: ceil(tm/2) x tn units of a * b^T
// (a unit = two 128-row blocks of a against one 128-row tile of b), followed (two-product mode) by ceil(tn/2) x tm units
// of b * a^T.  tm / tn = 128-row tile counts of a / b.
struct MProbD {
  int a, b;
  int tile0, ntiles, tm, tn;
  long long key_a, key_b;   // offsets of the a-row / b-row keys in keys[] (and sec[])
  long long out_off;        // matches0 / sim0 rows of this pair start here (cap(a) rows)
};

// a single pair travels inside the kernel parameters (constant bank): no table upload on the latency path
struct MTabInline {
  MOperD opers[2];
  MProbD probs[1];
};

struct TcMatchArgs {
  const MOperD* opers;    // device tables; NULL = use `inl`
  const MProbD* probs;
  MTabInline inl;
  int nprob, total_tiles;
  int passes;             // 1: one product, rows thread-local + columns through the threshold filter; 2: both products, rows only (top-2)
  int cols;               // passes == 1: reduce the columns too (mutual check)
  int split, stages, aslots;
  int sub;                // row-blocks per work unit (1 or 2), see tc_match.cu
  int mutual, ratio_mode;
  int debug;              // SFD2_TM_DEBUG (experiments only): 1 = epilogue handshakes without reductions, 2 = no column path, 4 = no finish tail
  float dist_th, ratio_th;
  unsigned long long* keys;
  unsigned* sec;          // second-best similarities (ordered uint), passes == 2 only
  const int* efflen;      // effective rows per set (written by the prep kernels)
  const int* remap;       // compaction tables, indexed by plane row
  int* done;              // per-pair completed-tile counters
  int32_t* matches0;
  float* sim0;
};

size_t tm_smem_bytes(int split, int sub, int aslots, int stages);
int tm_stages(int split, int sub, int aslots);
int tm_units(int tm, int tn, int passes, int sub);
int tm_make_plane_map(CUtensorMap* tm, const __half* base, size_t rows);
int launch_match_prep(const MOperD* opers_dev, const MTabInline* inl, int noper, int total_prows, bool any_ids, __half* hi, __half* lo, int* remap,
                      int* efflen, unsigned long long* keys, unsigned* sec, long long nkeys, int* done, int nprob,
                      int num_sms, cudaStream_t st);
int launch_match_tc(const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, TcMatchArgs a, int num_sms, cudaStream_t st);

}  // namespace sfd2
