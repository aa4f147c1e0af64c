// internal.cuh -- shared declarations of libb200match (not part of the C-ABI).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/b200match.h"

#define B200M_MAX_K 32
#define B200M_MAX_DIM 1024
#define B200M_TILE_N 256     // train rows per candidate-kernel tile; operand row counts are padded to this
#define B200M_TILE_M 128     // query rows per CTA
#define B200M_AUG_COLS 3     // extra K columns carrying |b|^2 as an FP16 triple
#define B200M_SENTINEL 60000.0f  // |b|^2 stand-in for invalid/padding train rows (FP16-representable)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes);   // grow-only
    void release();
    template <typename T> T *as() const { return (T *) p; }
};

struct Side {
    size_t n = 0;          // rows
    size_t n_pad = 0;      // rows rounded up to B200M_TILE_N
    int dim = 0;           // descriptor length
    int dp = 0;            // f32 row pitch in floats (dim rounded up to 4)
    int kp = 0;            // fp16 operand row pitch in halves (multiple of 64)
    int64_t index_offset = 0;
    uint64_t version = 0;  // bumped on every upload
    DevBuf staging;        // raw AoS as uploaded from the host
    DevBuf f32;            // [n][dp] dense FP32, zero padded columns
    DevBuf valid;          // [n] uint8
    DevBuf stats;          // per-CTA column statistics of the pack kernel (see pack.cu)
    DevBuf op_query;       // [n_pad][kp] fp16: -2*x16, 1,1,1, 0...
    DevBuf op_train;       // [n_pad][kp] fp16:  x16, nb_hi, nb_mid, nb_lo, 0...
    DevBuf norm16;         // [n_pad] float: |x16|^2 in scaled units
};

struct TcPrep {            // state shared by both sides' FP16 operands
    uint64_t ver[2] = {0, 0};
    bool ready = false;
    DevBuf mean;           // [dp] float centre
    DevBuf red;            // reduction scratch
    float scale = 1.f;     // power of two applied after centring
    float max_norm[2] = {0.f, 0.f};   // max |x16| per side (scaled units)
    bool usable = false;   // false when the data cannot be scaled into FP16 range (non-finite spread)
};

struct b200m_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    std::string err;
    Side side[2];
    TcPrep prep;
    DevBuf ws_cand_idx, ws_cand_cnt, ws_flag_rows, ws_counters, ws_scan, ws_out, ws_misc;
    DevBuf ws_part_i, ws_part_d, ws_done, ws_cand_val, ws_cand_thr;
    DevBuf ws_sweep_hint;   // [<= 64 train splits] int: where the candidate kernel's CTA pairs currently are in their sweep
    DevBuf ws_row_list, ws_row_flags, ws_sel_ops, ws_sel_norm;   // row selection (masked kNN)
    bool done_init = false;
    DevBuf ws_fidx, ws_fdist, ws_fcnt, ws_ridx, ws_rdist, ws_rcnt, ws_thr[2], ws_corr, ws_totals;
    bool totals_init = false;
    void *tmap_cache = nullptr;
    void *multiscale = nullptr;   // MultiscaleState (multiscale.cu)
    void *cluster = nullptr;      // ClusterState (cluster.cu)
    void *wide = nullptr;         // WideState (wide.cu)
    void *comm = nullptr;         // Comm (multi.cu): this context's rank in an NCCL communicator
    void *local = nullptr;        // LocalState (local.cu): the cell list of the gated kNN
    void *stage = nullptr;        // StagePool (api.cu): pinned bounce buffers for uploads from pageable host memory
    int local_min_rows = 4096;    // B200M_LOCAL_MIN_ROWS: train sets below this use the brute-force gate kernel
    int tc_cluster = 0;    // 0 = default; test/tuning override of the multicast cluster size (B200M_TC_CLUSTER)
    double masked_min_pairs = 1e9;   // B200M_MASKED_MIN_PAIRS: b200m_match skips unreferenced target rows in the reverse pass
                                     // from this many (source, target) pairs on (below, the row selection's host round trip
                                     // costs more than it saves)
    int tc_splits = 0;     // B200M_TC_SPLITS: 0 = chosen per launch (wave balance); > 0 forces the number of train splits
    int tc_splitn = 1;     // split-N candidate kernel for one-atom descriptors (B200M_TC_SPLITN=0: one N = 256 MMA per tile)
    int tc_alt = -1;       // B200M_TC_ALT: epilogue layout of the split-N kernel (4 / 5 chunk entries drained by 32- / 16-column
                           // TMEM loads; column entries: 1 alternating tiles, 2 quarter columns, 0 eight warps);
                           // -1 = by measurement: 5 up to k = 2, 4 beyond
    int tc_sweep_lag = -1; // B200M_TC_SWEEP_LAG: tiles between followers of the chunk-entry kernels' rotated sweep (0 = off, -1 = by size)
    int tc_lean = 1;       // B200M_TC_LEAN=0: general MMA issue loop also for one-atom descriptors (comparison)
    int tc_debug = 0;      // B200M_TC_DEBUG: timing experiments (results are NOT valid when set)
    int tc_pair = 1;       // 1 = CTA-pair (cta_group::2) candidate kernel; 0 = cta_group::1 + multicast (B200M_TC_MODE=mcast)
    bool profiling = false;
    b200m_stats stats{};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    struct EventPool *pool = nullptr;
};

int b200m_fail(b200m_ctx *ctx, const char *what, cudaError_t e, const char *file, int line);
int b200m_fail_msg(b200m_ctx *ctx, const std::string &msg);

#define CK(expr)                                                                   \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) return b200m_fail(ctx, #expr, _e, __FILE__, __LINE__); \
    } while (0)

// Per-region device timing without host synchronisation: when profiling is on, a region records a pair of
// events from the context's pool on the stream; pairs are resolved (cudaEventElapsedTime) when the stats are
// read or the pool fills up, so the timed program runs exactly as it does unprofiled.
struct EventPool {
    static const int kPairs = 128;
    cudaEvent_t ev[2 * kPairs] = {};
    double *slot[kPairs] = {};
    int used = 0;
    bool created = false;
};
void b200m_resolve_events(b200m_ctx *ctx);

struct StatTimer {
    b200m_ctx *ctx;
    int pair = -1;
    StatTimer(b200m_ctx *c, double *s);
    void stop();
};

// ---- kernel launchers (one translation unit each) ---------------------------
// pack.cu
// dense FP32 copy + validity + per-CTA column statistics (sum, min, max over valid rows) into `stats`
// (pack_stats_bytes() bytes), which launch_tc_prepare turns into the common centre and FP16 scale
cudaError_t launch_pack_f32(const float *aos, size_t n, size_t stride_bytes, int dim, int dp,
                            float *f32, uint8_t *valid, float *stats, int sm_count, cudaStream_t st);
size_t pack_stats_bytes(int sm_count, int dp);
cudaError_t launch_tc_prepare(b200m_ctx *ctx);   // centre/scale + FP16 operand tiles for both sides

// exact.cu
cudaError_t launch_exact_rows(const float *q_f32, const uint8_t *q_valid, int dp, int dim,
                              const float *t_f32, const uint8_t *t_valid, size_t nt, int64_t t_index_offset,
                              size_t row_begin, size_t n_rows, const int32_t *row_list, const int32_t *row_list_count,
                              int k, int32_t *idx, float *dist, int32_t *count, int max_blocks,
                              // few flagged rows (row_list given): split_blocks CTAs share every row; workspace of
                              // exact_split_ws_entries() int32 + float entries and exact_split_max_rows() zeroed counters
                              int split_blocks, int32_t *part_i, float *part_d, unsigned int *done,
                              // row_map (or null): the call's rows are row_map[0..n_rows) instead of row_begin + 0..n_rows;
                              // results are always stored at (row - row_begin)
                              const int32_t *row_map, cudaStream_t st);
cudaError_t launch_local_rows(const float *q_f32, const uint8_t *q_valid, int dp, const float *t_f32,
                              const uint8_t *t_valid, size_t nt, int64_t t_index_offset, size_t n_rows,
                              const float *q_xyz, const float *t_xyz, size_t xyz_stride_bytes, float radius, int k,
                              int32_t *idx, float *dist, int32_t *count, int max_blocks, cudaStream_t st);
size_t exact_split_ws_entries(int split_blocks, int k);
int exact_split_max_rows();
cudaError_t launch_rerank(const float *q_f32, const uint8_t *q_valid, int dp, int dim,
                          const float *t_f32, const uint8_t *t_valid, size_t nt, int64_t t_index_offset,
                          size_t row_begin, size_t n_rows, int k,
                          const int32_t *cand_idx, const int32_t *cand_cnt, int n_lists, int cap,
                          const float *cand_val, const float *cand_thr /* both null: no pruning */,
                          int32_t *idx, float *dist, int32_t *count,
                          int32_t *flag_rows, int32_t *counters /*[0]=flagged rows, [1..2]=candidate pairs (u64)*/,
                          int sm_count, const int32_t *row_map,
                          int chunk_entries /* 1: an entry is the first of 32 consecutive train rows + the chunk's minimum */,
                          cudaStream_t st);

// filter.cu
cudaError_t launch_filter(int mode, int k, float ratio_thr, float distance_thr, size_t row_begin, size_t n_rows,
                          const int32_t *fidx, const float *fdist, const int32_t *fcount,
                          const int32_t *ridx, const float *rdist, const int32_t *rcount, size_t n_rev_rows,
                          const float *thr_src, const float *thr_tgt, int64_t src_offset, int64_t tgt_offset,
                          b200m_corr *out, size_t cap, unsigned long long *n_out, float *avg,
                          void *scan_ws, size_t scan_ws_bytes, cudaStream_t st, int *n_launches,
                          const float *cdist = nullptr /* B200M_MODE_CLUSTER: [n_rows][k] from cluster.cu */);
size_t filter_scan_ws_bytes(size_t n_rows, int k);
cudaError_t launch_average(const float *fdist, const int32_t *fcount, size_t n_rows, int k, float *avg,
                           cudaStream_t st);
cudaError_t launch_merge(int k, int n_lists, size_t nq, const int32_t *idx_in, const float *dist_in,
                         const int32_t *count_in, int32_t *idx, float *dist, int32_t *count, cudaStream_t st);

// candidates_tc.cu
// Fills ws_cand_idx [n_lists][n_rows][cap] and ws_cand_cnt [n_lists][n_rows] (entries appended per list; a
// count above cap marks an overflowed list).  dump != nullptr: single tile, raw accumulators to dump[128][256].
// *has_values_out = 1: ws_cand_val [n_lists][n_rows][cap] and ws_cand_thr [n_lists][n_rows] are filled too;
// = 2: the same, and every entry is a 32-row CHUNK (first train row, smallest accumulator of the chunk).
// q_ops != null: the query rows are rows 0..n_rows of this compact [q_pad][kp] operand array (norms in q_norm) instead of
// rows row_begin.. of the query side's own array.
int tc_candidates(b200m_ctx *ctx, int direction, size_t row_begin, size_t n_rows, int k, int cap_request,
                  int *n_lists_out, int *cap_out, int *has_values_out, float *dump, size_t dump_t_tile,
                  const void *q_ops = nullptr, const float *q_norm = nullptr, size_t q_pad = 0);
bool tc_supported(const b200m_ctx *ctx, int dim, int k);
void tc_release(b200m_ctx *ctx);

// multiscale.cu: one table of per-scale k-lists filed under the query keypoints (match_multiscale's concatenation)
struct MultiscaleState {
    size_t n_query_kps = 0;
    int n_scales = 0, k = 0;
    DevBuf idx, dist, cnt, bad, qmap, tmap, xyz, oidx, odist, ocnt, kidx, kdist, kcnt;
};
void multiscale_release(b200m_ctx *ctx);
void ms_free(MultiscaleState *ms);
int ms_begin(b200m_ctx *ctx, MultiscaleState *ms, size_t n_query_kps, int n_scales, int k);
int ms_add_device(b200m_ctx *ctx, MultiscaleState *ms, int scale, size_t n_rows, const int32_t *d_idx, const float *d_dist,
                  const int32_t *d_count, const int32_t *d_query_map, const int32_t *d_train_map, size_t n_train_rows,
                  int64_t train_index_offset, size_t n_train_kps);
int ms_vote_device(b200m_ctx *ctx, MultiscaleState *ms, const float *d_train_xyz, size_t xyz_stride_bytes, float iss_radius,
                   int32_t *d_idx, float *d_dist, int32_t *d_count);

// wide.cu
void wide_release(b200m_ctx *ctx);

// multi.cu
void comm_release(b200m_ctx *ctx);

// local.cu: matchLocal with a finite radius over a cell list of the train keypoints; 0 = done, 1 = error, 2 = not worth a
// grid (the caller runs launch_local_rows)
void local_release(b200m_ctx *ctx);
int launch_local_cells(b200m_ctx *ctx, int direction, const float *d_query_xyz, const float *d_train_xyz, size_t xyz_stride_bytes,
                       float radius, int k, int32_t *d_idx, float *d_dist, int32_t *d_count);

// api.cu: kNN of query rows [row_begin, row_begin + n_rows) of `direction` on device buffers; with d_flags only the flagged
// (and valid) rows are answered, the others get empty lists.  Outputs are indexed by (row - row_begin).
extern "C" int b200m_knn_rows(b200m_ctx *ctx, const b200m_params *p, int direction, size_t row_begin, size_t n_rows, const uint8_t *d_flags,
                   int32_t *d_idx, float *d_dist, int32_t *d_count);

// cluster.cu
void cluster_release(b200m_ctx *ctx);
