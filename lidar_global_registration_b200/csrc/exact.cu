// exact.cu -- exact FP32 CUDA-core kernels.
//
// Arithmetic is the reference's own, bit for bit: pcl::L2_Norm as called by matchLocal
// (reference include/matching.h:663) == FLANN L2_Simple under matchFLANN (:581) --
//     s = 0; for d: diff = a[d]-b[d]; s = s + diff*diff;  dist = sqrt(s)
// sequential over d in FP32 with separate multiply and add (the reference builds for
// baseline x86-64: no FMA), so distances are BIT-IDENTICAL to the CPU oracle and the
// order (dist, index) -- KNNResult's "earlier stays first", include/matching.h:69-93 --
// is reproduced exactly, ties included.
//
//   exact_rows_kernel   one CTA per query row, all train rows: the universal exact path
//                       (B200M_PREC_F32_EXACT, and the fallback for rows whose candidate
//                       list overflowed).
//   rerank_kernel       one warp per query row over the candidate lists of the tensor-core
//                       pass (a certified superset of the exact top-k, see candidates_tc.cu):
//                       exact distances and the final (dist, idx) order.
#include <float.h>
#include <limits.h>
#include <math.h>

#include "internal.cuh"

namespace {

__device__ __forceinline__ bool lex_less(float d1, int i1, float d2, int i2) {
    return d1 < d2 || (d1 == d2 && i1 < i2);
}

// pcl::L2_Norm's squared sum over a dp-padded row pair (padding columns are 0 on both
// sides and add +0.0f, which leaves s unchanged).  q in shared memory, t in global.
__device__ __forceinline__ float seq_sqdist(const float *__restrict__ q, const float *__restrict__ t, int dp) {
    float s = 0.f;
    const float4 *t4 = reinterpret_cast<const float4 *>(t);
    for (int c = 0; c < dp / 4; ++c) {
        float4 v = __ldg(t4 + c);
        float d0 = __fsub_rn(q[4 * c + 0], v.x);
        s = __fadd_rn(s, __fmul_rn(d0, d0));
        float d1 = __fsub_rn(q[4 * c + 1], v.y);
        s = __fadd_rn(s, __fmul_rn(d1, d1));
        float d2 = __fsub_rn(q[4 * c + 2], v.z);
        s = __fadd_rn(s, __fmul_rn(d2, d2));
        float d3 = __fsub_rn(q[4 * c + 3], v.w);
        s = __fadd_rn(s, __fmul_rn(d3, d3));
    }
    return s;
}

constexpr int kExactThreads = 256;

template <int KMAX>
__global__ void __launch_bounds__(kExactThreads)
exact_rows_kernel(const float *__restrict__ q_f32, const uint8_t *__restrict__ q_valid, int dp, int dim,
                  const float *__restrict__ t_f32, const uint8_t *__restrict__ t_valid, size_t nt,
                  long long t_off, size_t row_begin, size_t n_rows, const int32_t *__restrict__ row_list,
                  const int32_t *__restrict__ row_list_count, int k, int32_t *__restrict__ idx,
                  float *__restrict__ dist, int32_t *__restrict__ count) {
    extern __shared__ float smem[];
    float *sq = smem;                                   // [dp] query row
    float *red_d = smem + dp;                           // [warps]
    int *red_i = reinterpret_cast<int *>(red_d + kExactThreads / 32);
    int *red_w = red_i + kExactThreads / 32;            // winner thread id per warp
    __shared__ float win_d;
    __shared__ int win_i, win_t;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t total = row_list ? (size_t) *row_list_count : n_rows;

    for (size_t r = blockIdx.x; r < total; r += gridDim.x) {
        const size_t local = row_list ? (size_t) row_list[r] : r;
        const size_t qi = row_begin + local;
        int32_t *oi = idx + local * k;
        float *od = dist + local * k;
        if (!q_valid[qi]) {   // non-finite query -> empty entry (reference include/matching.h:576)
            if (tid < k) { oi[tid] = -1; od[tid] = 0.f; }
            if (tid == 0) count[local] = 0;
            continue;
        }
        __syncthreads();
        for (int d = tid; d < dp; d += kExactThreads) sq[d] = q_f32[qi * (size_t) dp + d];
        __syncthreads();

        float ld[KMAX];
        int li[KMAX];
#pragma unroll
        for (int m = 0; m < KMAX; ++m) { ld[m] = INFINITY; li[m] = INT_MAX; }
        for (size_t j = tid; j < nt; j += kExactThreads) {
            if (!t_valid[j]) continue;   // invalid train rows are never candidates (:661)
            float d = __fsqrt_rn(seq_sqdist(sq, t_f32 + j * (size_t) dp, dp));
            if (lex_less(d, (int) j, ld[KMAX - 1], li[KMAX - 1])) {
                // sorted insertion, fully unrolled so the list stays in registers
                float cd = d;
                int ci = (int) j;
#pragma unroll
                for (int m = 0; m < KMAX; ++m) {
                    if (lex_less(cd, ci, ld[m], li[m])) {
                        float td = ld[m]; int ti = li[m];
                        ld[m] = cd; li[m] = ci;
                        cd = td; ci = ti;
                    }
                }
            }
        }
        // k rounds of block-wide arg-min over the per-thread list heads
        int head = 0, found = 0;
        for (int round = 0; round < k; ++round) {
            float hd = INFINITY;
            int hi = INT_MAX;
#pragma unroll
            for (int m = 0; m < KMAX; ++m)
                if (m == head) { hd = ld[m]; hi = li[m]; }
            float bd = hd;
            int bi = hi, bt = tid;
            for (int o = 16; o > 0; o >>= 1) {
                float od2 = __shfl_xor_sync(0xffffffffu, bd, o);
                int oi2 = __shfl_xor_sync(0xffffffffu, bi, o);
                int ot2 = __shfl_xor_sync(0xffffffffu, bt, o);
                if (lex_less(od2, oi2, bd, bi)) { bd = od2; bi = oi2; bt = ot2; }
            }
            if (lane == 0) { red_d[warp] = bd; red_i[warp] = bi; red_w[warp] = bt; }
            __syncthreads();
            if (tid == 0) {
                float wd = red_d[0];
                int wi = red_i[0], wt = red_w[0];
                for (int w = 1; w < kExactThreads / 32; ++w)
                    if (lex_less(red_d[w], red_i[w], wd, wi)) { wd = red_d[w]; wi = red_i[w]; wt = red_w[w]; }
                win_d = wd; win_i = wi; win_t = wt;
            }
            __syncthreads();
            if (win_i == INT_MAX) break;   // fewer than k valid train rows
            if (tid == win_t) head++;
            if (tid == 0) { oi[round] = (int32_t) (win_i + t_off); od[round] = win_d; }
            found = round + 1;
            __syncthreads();
        }
        if (tid >= found && tid < k) { oi[tid] = -1; od[tid] = 0.f; }
        if (tid == 0) count[local] = found;
    }
}

// ---- re-rank -----------------------------------------------------------------
// One warp per query row.  The candidate lists of the tensor-core pass are walked 32 at a
// time: candidate c belongs to lane c, which runs the reference's sequential FP32 chain for
// it.  Train rows are fetched 128 B at a time by the whole warp (coalesced, each line read
// once) into a 32x33 shared tile, then every lane consumes its own row of the tile.
constexpr int kRerankWarps = 8;
constexpr int kMaxLists = 16;

template <int KMAX>
__global__ void __launch_bounds__(kRerankWarps * 32)
rerank_kernel(const float *__restrict__ q_f32, const uint8_t *__restrict__ q_valid, int dp, int dim,
              const float *__restrict__ t_f32, const uint8_t *__restrict__ t_valid, size_t nt, long long t_off,
              size_t row_begin, size_t n_rows, int k,
              const int32_t *__restrict__ cand_idx, const int32_t *__restrict__ cand_cnt, int n_lists, int cap,
              int32_t *__restrict__ idx, float *__restrict__ dist, int32_t *__restrict__ count,
              int32_t *__restrict__ flag_rows, int32_t *__restrict__ counters) {
    extern __shared__ float smem[];
    __shared__ unsigned long long blk_cands;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *sq = smem + (size_t) warp * (dp + 32 * 33);
    float *tile = sq + dp;
    if (threadIdx.x == 0) blk_cands = 0ull;
    __syncthreads();
    const size_t local = (size_t) blockIdx.x * kRerankWarps + warp;
    unsigned long long my_cands = 0ull;
    if (local < n_rows) {
        const size_t qi = row_begin + local;
        int32_t *oi = idx + local * k;
        float *od = dist + local * k;
        if (!q_valid[qi]) {   // non-finite query -> empty entry (reference include/matching.h:576)
            for (int m = lane; m < k; m += 32) { oi[m] = -1; od[m] = 0.f; }
            if (lane == 0) count[local] = 0;
        } else {
            for (int d = lane; d < dp; d += 32) sq[d] = q_f32[qi * (size_t) dp + d];
            // list sizes (warp-uniform)
            int total = 0;
            bool overflow = false;
            for (int l = 0; l < n_lists; ++l) {
                int c = cand_cnt[(size_t) l * n_rows + local];
                if (c > cap) { overflow = true; c = cap; }
                total += c;
            }
            float ld[KMAX];
            int li[KMAX];
#pragma unroll
            for (int m = 0; m < KMAX; ++m) { ld[m] = INFINITY; li[m] = INT_MAX; }
            __syncwarp();
            if (!overflow) {
                my_cands = (unsigned long long) total;
                for (int base = 0; base < total; base += 32) {
                    int c = base + lane;
                    int j = -1;
                    if (c < total) {
                        int l = 0;
                        for (; l < n_lists; ++l) {
                            int cl = min(cand_cnt[(size_t) l * n_rows + local], cap);
                            if (c < cl) break;
                            c -= cl;
                        }
                        j = cand_idx[((size_t) l * n_rows + local) * cap + c];
                        if (j < 0 || (size_t) j >= nt || !t_valid[j]) j = -1;   // padding / invalid train rows (:661)
                    }
                    const unsigned live = __ballot_sync(0xffffffffu, j >= 0);
                    float s = 0.f;
                    for (int d0 = 0; d0 < dp; d0 += 32) {
                        const bool in = d0 + lane < dp;
#pragma unroll 8
                        for (int cc = 0; cc < 32; ++cc) {
                            int jj = __shfl_sync(0xffffffffu, j, cc);
                            if ((live >> cc) & 1u) {
                                float v = in ? __ldg(t_f32 + (size_t) jj * dp + d0 + lane) : 0.f;
                                tile[cc * 33 + lane] = v;
                            }
                        }
                        __syncwarp();
                        if (j >= 0) {
                            const int ne = min(32, dp - d0);
                            const float *tr = tile + lane * 33;
                            for (int e = 0; e < ne; ++e) {
                                float df = __fsub_rn(sq[d0 + e], tr[e]);
                                s = __fadd_rn(s, __fmul_rn(df, df));
                            }
                        }
                        __syncwarp();
                    }
                    if (j >= 0) {
                        float cd = __fsqrt_rn(s);
                        int ci = j;
                        if (lex_less(cd, ci, ld[KMAX - 1], li[KMAX - 1])) {
#pragma unroll
                            for (int m = 0; m < KMAX; ++m) {
                                if (lex_less(cd, ci, ld[m], li[m])) {
                                    float td = ld[m]; int ti = li[m];
                                    ld[m] = cd; li[m] = ci;
                                    cd = td; ci = ti;
                                }
                            }
                        }
                    }
                }
            }
            // k rounds of warp arg-min over the per-lane list heads
            int head = 0, found = 0;
            for (int round = 0; round < k; ++round) {
                float hd = INFINITY;
                int hi = INT_MAX;
#pragma unroll
                for (int m = 0; m < KMAX; ++m)
                    if (m == head) { hd = ld[m]; hi = li[m]; }
                float wd = hd;
                int wi = hi;
                for (int o = 16; o > 0; o >>= 1) {
                    float od2 = __shfl_xor_sync(0xffffffffu, wd, o);
                    int oi2 = __shfl_xor_sync(0xffffffffu, wi, o);
                    if (lex_less(od2, oi2, wd, wi)) { wd = od2; wi = oi2; }
                }
                if (wi == INT_MAX) break;
                if (hi == wi) head++;   // candidate indices are unique per row
                if (lane == 0) { oi[round] = (int32_t) (wi + t_off); od[round] = wd; }
                found = round + 1;
            }
            for (int m = found + lane; m < k; m += 32) { oi[m] = -1; od[m] = 0.f; }
            if (lane == 0) {
                count[local] = found;
                if (overflow) {
                    int pos = atomicAdd(counters, 1);
                    flag_rows[pos] = (int32_t) local;
                }
            }
        }
    }
    if (lane == 0 && my_cands) atomicAdd(&blk_cands, my_cands);
    __syncthreads();
    if (threadIdx.x == 0 && blk_cands)
        atomicAdd(reinterpret_cast<unsigned long long *>(counters + 2), blk_cands);
}

template <int KMAX>
cudaError_t launch_exact_t(const float *q_f32, const uint8_t *q_valid, int dp, int dim, const float *t_f32,
                           const uint8_t *t_valid, size_t nt, int64_t t_off, size_t row_begin, size_t n_rows,
                           const int32_t *row_list, const int32_t *row_list_count, int k, int32_t *idx, float *dist,
                           int32_t *count, int blocks, cudaStream_t st) {
    size_t smem = sizeof(float) * dp + (sizeof(float) + 2 * sizeof(int)) * (kExactThreads / 32);
    exact_rows_kernel<KMAX><<<blocks, kExactThreads, smem, st>>>(q_f32, q_valid, dp, dim, t_f32, t_valid, nt,
                                                                 (long long) t_off, row_begin, n_rows, row_list,
                                                                 row_list_count, k, idx, dist, count);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_exact_rows(const float *q_f32, const uint8_t *q_valid, int dp, int dim,
                              const float *t_f32, const uint8_t *t_valid, size_t nt, int64_t t_index_offset,
                              size_t row_begin, size_t n_rows, const int32_t *row_list, const int32_t *row_list_count,
                              int k, int32_t *idx, float *dist, int32_t *count, int max_blocks, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    int blocks = (int) (n_rows < (size_t) max_blocks ? n_rows : (size_t) max_blocks);
#define B200M_EXACT_CASE(K)                                                                                   \
    return launch_exact_t<K>(q_f32, q_valid, dp, dim, t_f32, t_valid, nt, t_index_offset, row_begin, n_rows,  \
                             row_list, row_list_count, k, idx, dist, count, blocks, st)
    if (k <= 1) B200M_EXACT_CASE(1);
    if (k <= 2) B200M_EXACT_CASE(2);
    if (k <= 4) B200M_EXACT_CASE(4);
    if (k <= 8) B200M_EXACT_CASE(8);
    if (k <= 16) B200M_EXACT_CASE(16);
    B200M_EXACT_CASE(32);
#undef B200M_EXACT_CASE
}

cudaError_t launch_rerank(const float *q_f32, const uint8_t *q_valid, int dp, int dim,
                          const float *t_f32, const uint8_t *t_valid, size_t nt, int64_t t_index_offset,
                          size_t row_begin, size_t n_rows, int k,
                          const int32_t *cand_idx, const int32_t *cand_cnt, int n_lists, int cap,
                          int32_t *idx, float *dist, int32_t *count,
                          int32_t *flag_rows, int32_t *counters, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    if (n_lists > kMaxLists) return cudaErrorInvalidValue;
    unsigned blocks = (unsigned) ((n_rows + kRerankWarps - 1) / kRerankWarps);
    size_t smem = sizeof(float) * (size_t) (dp + 32 * 33) * kRerankWarps;
#define B200M_RERANK_CASE(K)                                                                                    \
    do {                                                                                                        \
        if (smem > 48 * 1024) {                                                                                 \
            cudaError_t e = cudaFuncSetAttribute(rerank_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 (int) smem);                                                   \
            if (e != cudaSuccess) return e;                                                                     \
        }                                                                                                       \
        rerank_kernel<K><<<blocks, kRerankWarps * 32, smem, st>>>(                                              \
            q_f32, q_valid, dp, dim, t_f32, t_valid, nt, (long long) t_index_offset, row_begin, n_rows, k,      \
            cand_idx, cand_cnt, n_lists, cap, idx, dist, count, flag_rows, counters);                           \
        return cudaGetLastError();                                                                              \
    } while (0)
    if (k <= 1) B200M_RERANK_CASE(1);
    if (k <= 2) B200M_RERANK_CASE(2);
    if (k <= 4) B200M_RERANK_CASE(4);
    if (k <= 8) B200M_RERANK_CASE(8);
    if (k <= 16) B200M_RERANK_CASE(16);
    B200M_RERANK_CASE(32);
#undef B200M_RERANK_CASE
}
