// Internal (non-ABI) declarations shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/tristage.h"

namespace ts {

void set_error(const char* fmt, ...);
const char* get_error();

#define TS_CUDA_OK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ts::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return TS_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

// experiment switches (TS_DBG_*, TS_SELECT_V1): on when set to anything but "" or "0"
inline bool env_on(const char* name) {
  const char* e = getenv(name);
  return e && e[0] && !(e[0] == '0' && !e[1]);
}
// kernel variants with a build-time default that the environment can override in both directions
// (NAME=1 / NAME=0).  Flip a default here once the variant has been validated and timed on hardware.
// Validated on a B200 (bit-equal to the two-launch single-CTA scan; profiles/README.md): TS_FUSE x1.02-1.03 on
// 1.25 M-row shards, TS_PAIR x1.12 at B = 256, tf32 for fp32 storage replaces B/4 CUDA-core passes by one.
#ifdef TS_CUDASIM
constexpr bool kDefaultFuse = false;   // the emulator would need all 148 x 192 fibers of a cooperative grid alive at once
#else
constexpr bool kDefaultFuse = true;    // TS_FUSE : threshold pre-pass + scan in one cooperative launch
#endif
constexpr bool kDefaultS2V2 = false;   // TS_S2_V2: second Stage-2 epilogue
constexpr bool kDefaultS2Flow = true;   // TS_S2_FLOW: Stage-2 tensor kernel with the resident query tile (s2_flow.cu); 0 = first kernel
constexpr bool kDefaultS2Epi2 = false; // TS_S2_EPI2: two Stage-2 epilogue warpgroups (320 threads), one per accumulator
constexpr bool kDefaultPdl = false;    // TS_PDL  : select / wait-merge / wait-take kernels use programmatic dependent launch.  Off: a clean,
                                       // interleaved A/B (no events between the kernels) shows the pre-launched kernel neutral or slower:
                                       // +4 us per 440 us step at B = 32, +4-8 us under CUDA-graph replay, +14 us on the k = 500 host call
constexpr bool kDefaultPair = true;    // TS_PAIR : cta_group::2 CTA pairs for B >= 129
constexpr bool kDefaultTf32 = true;    // TS_TF32 : fp32 storage takes the tensor path (kind::tf32) for B > 4 under TS_PATH_AUTO
inline bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  if (!e || !e[0]) return dflt;
  return !(e[0] == '0' && !e[1]);
}

inline int dtype_size(int dt) { return dt == TS_F32 ? 4 : 2; }
// row pitch in elements: rows are 16-byte aligned (TMA / 128-bit loads)
inline int row_pitch(int dim, int dt) { return dt == TS_F32 ? ((dim + 3) / 4) * 4 : ((dim + 7) / 8) * 8; }
inline int cap_for_k(int k) { return k <= 128 ? 256 : 1024; }

struct DeviceInfo {
  int sm_count;
  int cc_major, cc_minor;
  size_t smem_optin;
};
int get_device_info(int device, DeviceInfo* out);

// ---- convert / normalise rows (ingest K1/K3 and query prep) ---------------
enum NormMode { kNormNone = 0, kNormStage1 = 1 /* x/(|x|+1e-8) */, kNormStage2 = 2 /* x/max(|x|,1e-12) */ };
// src [n, dim] (pitch src_ld) of src_dtype -> dst [n, dst_ld] of dst_dtype, pad
// columns zeroed.  norm_mode applies to the stored values unless inv_norm_out
// is given, in which case values are stored un-normalised and 1/(|x|+eps) is
// written to inv_norm_out[n] (METRIC_COSINE).
int launch_convert_rows(const void* src, int src_dtype, int64_t src_ld, void* dst, int dst_dtype, int64_t dst_ld,
                        int64_t n, int dim, int norm_mode, float* inv_norm_out, cudaStream_t st,
                        unsigned int* zero_word = nullptr,    // 16 scheduling words of the scan, zeroed by the kernel
                        unsigned long long* tl = nullptr);    // TS_DBG_TIMELINE slots (see ts_index_debug_timeline)

// ---- Stage-1 scans -> partial keys [L][B][k] ------------------------------
struct ScanArgs {
  const void* rows;       // [n][ld] storage dtype
  int64_t n;
  int dim, ld, dtype;
  const float* inv_norm;  // nullptr unless METRIC_COSINE
  const void* q;          // [B][ld] storage dtype (prepared)
  int B, k;
  uint64_t* lists;        // scratch candidate lists
  size_t lists_keys;      // capacity in keys
  uint64_t* partial;      // stream path out: [L][B][k] keys
  size_t partial_keys;    // capacity in keys
  int* counts;            // umma path out: entries per candidate list
  float* pub;             // umma path scratch: published per-slice thresholds
  unsigned int* grid_bar; // umma path: 16 words zeroed by the query-prep kernel of the same call -- [0] arrival counter of the
                          // in-kernel grid barrier, [1 + mt] next-tile counter of query tile mt (dynamic tile schedule)
  int coop;               // the device supports cooperative launches (needed for the fused pre-pass)
  unsigned long long* tl; // TS_DBG_TIMELINE: time-line slots of the step (null = off): [2] first CTA entry, [3] last CTA exit
  int sm_count;
};
// where the umma scan leaves its candidates (consumed by launch_merge_lists)
struct UmmaLayout {
  int n_slices, n_mt, grid, cap, spread, bpad, jrank, rows_per_cta, fused, pair;
  int kth_rule;   // jrank == 1 and n_slices >= k: the slices' published bests are n_slices distinct rows, so the select kernel may
                  // filter with their k-th largest instead of their minimum
  size_t lists_keys, counts_n, pub_n;
  unsigned long long* tl;          // TS_DBG_TIMELINE: the step's time-line slots (null = off)
};
// number of partial lists L a scan will emit / scratch it needs
int s1_stream_plan(const ScanArgs& a, int* L, size_t* lists_keys);
int s1_umma_plan(const ScanArgs& a, UmmaLayout* lay);
int launch_s1_stream(const ScanArgs& a, cudaStream_t st, int* launches);
int launch_s1_umma(const ScanArgs& a, const UmmaLayout& lay, cudaStream_t st, int* launches);

// ---- top-k selection / merge ----------------------------------------------
// Fused exchange (multi-GPU, peer memory): the select kernel's final pass also stores query b's [k] scores and ids
// into slot `rank` of EVERY rank's receive buffer (16/8-byte stores over NVLink; the own buffer is one of them) and
// then publishes flag[parity][rank][b] = seq there with a system-scope release.  Layout of a receive buffer
// (ts_exchange_buffer_bytes): [parity][rank] slots of slot_bytes = scores [B_max*k_max] f32 | ids [B_max*k_max] i64,
// then flags [parity][rank][B_max] u32.
struct PushTarget {
  const long long* peer_bases;   // device array [n_ranks]: base address of every rank's buffer as seen from this GPU
  int n_ranks;
  long long scores_off;          // byte offset of this rank's score slot (parity applied) inside a buffer
  long long ids_off;             // ... of its id slot
  long long flags_off;           // ... of flags[parity][rank][0]
  unsigned int seq;              // sequence number of this step (never 0)
  // fused wait + merge (null merge_scores = off): after pushing query b's row the same CTA waits for the G rows of
  // query b in THIS rank's buffer and merges them -- the whole exchange is one kernel (see select_kernel)
  const float* merge_scores;     // this rank's buffer: score slot of rank 0, this parity
  const int64_t* merge_ids;      // ... id slot of rank 0
  long long merge_stride_f, merge_stride_i;   // floats / int64s between consecutive ranks' slots
  const unsigned int* merge_flags;            // flags[parity][0][0] of this rank's buffer; flag (l, b) at l * merge_flag_stride + b
  int merge_flag_stride;
  float* merge_out_scores;       // [B][k] merged result
  int64_t* merge_out_ids;
};
// keys [L][B][k] -> final (scores, ids) [B][k]; tmp0/tmp1 each hold
// ceil(L/2)*B*k keys (only used when L*k exceeds one selection pass).
int launch_merge_keys(const uint64_t* keys, int L, int B, int k, int64_t id_base, uint64_t* tmp0, uint64_t* tmp1,
                      float* out_scores, int64_t* out_ids, cudaStream_t st, int* launches, const PushTarget* push = nullptr);
size_t merge_tmp_keys(int L, int B, int k);
// unsorted per-(CTA, query) candidate lists of the umma scan -> final (scores, ids) [B][k]
int launch_merge_lists(const uint64_t* lists, const int* counts, const float* pub, const UmmaLayout& lay, int B, int k,
                       int64_t id_base, float* out_scores, int64_t* out_ids, cudaStream_t st, int* launches,
                       const PushTarget* push = nullptr);
// strides = elements between consecutive lists (floats for scores, int64s for ids)
int launch_merge_pairs(const float* scores, const int64_t* ids, long long stride_scores, long long stride_ids, int L,
                       int B, int k, float* out_scores, int64_t* out_ids, cudaStream_t st);
// peer-memory exchange (multi-GPU merge without a collective library call)
int launch_exchange_push(const void* blob, long long nbytes, const long long* peer_bases_dev, int n_ranks, long long slot_off,
                         long long flag_off, unsigned int seq, cudaStream_t st);
int launch_exchange_wait_take(void* matrix, const unsigned int* flags, int n_ranks, unsigned int seq, long long n, float* out,
                              cudaStream_t st);
int launch_exchange_wait_sum(const void* slots, long long slot_bytes, const unsigned int* flags, int n_ranks, unsigned int seq,
                             long long n, float* out, cudaStream_t st);
// wait_flag_stride = 0: one flag per list (whole [B,k] blob published at once); > 0: flag of (list l, query b) at
// wait_flags[l * stride + b] (the fused exchange publishes per query)
int launch_merge_pairs_wait(const float* scores, const int64_t* ids, long long stride_scores, long long stride_ids, int L, int B,
                            int k, const unsigned int* wait_flags, unsigned int wait_seq, float* out_scores, int64_t* out_ids,
                            cudaStream_t st, int wait_flag_stride = 0);
int launch_rank_desc(const float* scores, const int32_t* n_cand, int B, int C, int top_k, float* out_scores,
                     int32_t* out_pos, cudaStream_t st);

// ---- Stage-2 ---------------------------------------------------------------
// Layout of a token shard in HBM.  Both keep every doc on an 8-row boundary and use the same doc tables; they
// differ in the byte order INSIDE each 8-row group (8 * dim elements) and in what the pad rows hold:
//   kTokRowMajor : tok[row][dim], pad rows zero                      (first tensor kernel, fp32 / odd dims)
//   kTokTile     : tok[row / 8][dim / 8][row % 8][8], pad rows repeat the doc's last token
// kTokTile is the shared-memory image tcgen05.mma reads for a K-major operand without swizzle (core matrices of
// 8 rows x 16 bytes, LBO = 128 B along K, SBO = 16 * dim B between 8-row groups): a doc -- or any 8-row-aligned
// part of it -- moves from HBM into its columns of a tile with ONE contiguous cp.async.bulk, and the repeated
// last token makes the pad columns harmless for a row maximum, so the epilogue never masks.
enum TokLayout { kTokRowMajor = 0, kTokTile = 1 };
inline bool tok_tile_layout_ok(int dim, int dtype) { return dtype != TS_F32 && dim >= 16 && dim % 16 == 0 && dim <= 256; }
// scatter ragged docs into the 8-row padded store (tok_ingest.cu)
int launch_tok_ingest(const void* src, int src_dtype, const int64_t* src_off_dev, const int64_t* dst_off_dev,
                      const int32_t* len_dev, int n_docs, void* dst, int dst_dtype, int dim, int normalize, int layout,
                      cudaStream_t st);
// in-place conversion of docs [doc_lo, doc_lo + n_docs) between the two layouts (2-byte dtypes, dim % 8 == 0)
int launch_tok_relayout(void* tok, int dtype, const int64_t* doc_off_dev, const int32_t* doc_len_dev, int64_t doc_lo,
                        int64_t n_docs, int dim, int to_layout, cudaStream_t st);

struct MaxSimArgs {
  int layout;              // TokLayout of tok
  const void* tok;         // [ntok_pad][dim] storage dtype, docs padded to 8 rows
  const int64_t* doc_off;  // [ndocs] first row of each doc (multiple of 8)
  const int32_t* doc_len;  // [ndocs]
  int64_t ndocs, id_base;
  int64_t ntok_rows;       // rows allocated in tok (for the tensor map)
  int dim, dtype;
  const void* q;           // [B][lq_stride][dim] storage dtype, normalised
  const int32_t* q_len;    // [B] or nullptr
  int B, lq_stride;
  const int64_t* cand;     // [B][C]
  const int32_t* n_cand;   // [B] or nullptr
  int C, mode;
  float* out;              // [B][C]
  int sm_count;
  // multi-GPU scatter (flow kernel only; all null / 0 otherwise): every finalised score is ALSO stored at its flat
  // position of the [B][C] matrix at scatter_off inside every rank's receive buffer, and the last CTA publishes
  // `seq` in flag `rank` of every buffer (system-scope release) -- see ts_maxsim_scatter in include/tristage.h
  const long long* scatter_bases;   // device array [scatter_n] of the buffers' base addresses as seen from this GPU
  int scatter_n, scatter_rank;
  long long scatter_off, scatter_flags_off;
  unsigned int scatter_seq;
  unsigned int* scatter_done;       // device counter (zero between launches) for the last-CTA election
};
int launch_maxsim(const MaxSimArgs& a, cudaStream_t st, int* launches);
// s2_flow.cu: resident query tile, docs split across full tiles (dim <= 256); launch_maxsim dispatches to it
bool maxsim_flow_takes(const MaxSimArgs& a);
int launch_maxsim_flow(const MaxSimArgs& a, cudaStream_t st, int* launches);

}  // namespace ts
