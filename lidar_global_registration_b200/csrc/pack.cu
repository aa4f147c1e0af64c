// pack.cu -- descriptor upload kernels: the GPU replacement of pcl2cv<FeatureT>
// (reference include/matching.h:553-560), done ONCE per descriptor set instead of once
// per (query block, train block) pair.
//
//   pack_f32      AoS rows (stride sizeof(FeatureT)) -> dense FP32 [n][dp] + validity byte
//                 (pcl::PointRepresentation::isValid: every value finite).
//   tc_prepare    common centre + power-of-two scale for both sides, then the two FP16
//                 operand tiles per side for the tcgen05 candidate kernel:
//                     as query : [-2*x16 , 1, 1, 1, 0..]          (A operand, K-major)
//                     as train : [   x16 , nb_hi, nb_mid, nb_lo, 0..]  (B operand, K-major)
//                 so that  A.B = |b16|^2 - 2 a16.b16  comes out of the tensor core with
//                 the train-row norm already fused in.  Invalid / padding train rows carry
//                 the FP16 sentinel 60000 in nb_hi so they can never look near.
//
// All HBM-bound: algorithmic bytes per row = stride_in + 4*dp + 1 (pack_f32) and
// 4*dp + 2*2*kp + 4 (operand pack) -- see DESIGN.md.
#include <math.h>
#include <string.h>

#include "internal.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pack_f32_kernel(const float *__restrict__ aos, size_t n, size_t stride_floats, int dim, int dp,
                float *__restrict__ f32, uint8_t *__restrict__ valid) {
    const int lane = threadIdx.x & 31;
    size_t row = (size_t) blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= n) return;
    const float *src = aos + row * stride_floats;
    float *dst = f32 + row * (size_t) dp;
    bool ok = true;
    for (int d = lane; d < dp; d += 32) {
        float v = 0.f;
        if (d < dim) {
            v = __ldg(src + d);
            ok = ok && isfinite(v);
        }
        dst[d] = v;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) valid[row] = ok ? 1 : 0;
}

// ---- column sums over valid rows (for the common centre) --------------------
// grid.x blocks stride over rows; thread t owns columns t, t+256, ... (<= 4 of them).
__global__ void __launch_bounds__(256)
colsum_kernel(const float *__restrict__ f32, const uint8_t *__restrict__ valid, size_t n, int dp,
              double *__restrict__ partial /*[grid][dp+1]*/) {
    double acc[4] = {0, 0, 0, 0};
    double cnt = 0;
    for (size_t r = blockIdx.x; r < n; r += gridDim.x) {
        if (!valid[r]) continue;
        const float *row = f32 + r * (size_t) dp;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int d = threadIdx.x + 256 * c;
            if (d < dp) acc[c] += (double) row[d];
        }
        cnt += 1.0;
    }
    double *out = partial + (size_t) blockIdx.x * (dp + 1);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        int d = threadIdx.x + 256 * c;
        if (d < dp) out[d] = acc[c];
    }
    if (threadIdx.x == 0) out[dp] = cnt;
}

__global__ void mean_kernel(const double *__restrict__ pa, int ga, const double *__restrict__ pb, int gb,
                            int dp, float *__restrict__ mean) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dp) return;
    double s = 0, c = 0;
    for (int g = 0; g < ga; ++g) { s += pa[(size_t) g * (dp + 1) + d]; c += pa[(size_t) g * (dp + 1) + dp]; }
    for (int g = 0; g < gb; ++g) { s += pb[(size_t) g * (dp + 1) + d]; c += pb[(size_t) g * (dp + 1) + dp]; }
    mean[d] = c > 0 ? (float) (s / c) : 0.f;
}

// max |x - mean| over valid rows -> bits of a non-negative float, atomicMax as int
__global__ void __launch_bounds__(256)
maxabs_kernel(const float *__restrict__ f32, const uint8_t *__restrict__ valid, size_t n, int dp, int dim,
              const float *__restrict__ mean, int *__restrict__ out_bits) {
    float m = 0.f;
    for (size_t r = blockIdx.x; r < n; r += gridDim.x) {
        if (!valid[r]) continue;
        const float *row = f32 + r * (size_t) dp;
        for (int d = threadIdx.x; d < dim; d += 256) m = fmaxf(m, fabsf(row[d] - mean[d]));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_int(m));
}

// One warp per (padded) row: FP16 operand rows + |x16|^2.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pack_operands_kernel(const float *__restrict__ f32, const uint8_t *__restrict__ valid, size_t n, size_t n_pad,
                     int dim, int dp, int kp, const float *__restrict__ mean, float scale,
                     __half *__restrict__ op_query, __half *__restrict__ op_train, float *__restrict__ norm16,
                     int *__restrict__ max_norm_bits) {
    const int lane = threadIdx.x & 31;
    size_t row = (size_t) blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= n_pad) return;
    const bool ok = row < n && valid[row];
    __half *oq = op_query + row * (size_t) kp;
    __half *ot = op_train + row * (size_t) kp;
    double nrm = 0.0;
    for (int d = lane; d < kp; d += 32) {
        __half hq = __float2half_rn(0.f), ht = hq;
        if (d < dim) {
            if (ok) {
                float x = (f32[row * (size_t) dp + d] - mean[d]) * scale;
                ht = __float2half_rn(x);
                float xf = __half2float(ht);
                hq = __float2half_rn(-2.f * xf);   // exact: |x| <= 1, power-of-two factor
                nrm += (double) xf * (double) xf;
            }
        } else if (d < dim + B200M_AUG_COLS) {
            hq = __float2half_rn(1.f);             // picks up the three norm pieces of the train row
        }
        oq[d] = hq;
        if (d < dim) ot[d] = ht;
        else if (d >= dim + B200M_AUG_COLS) ot[d] = __float2half_rn(0.f);
    }
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0) {
        float nf = ok ? (float) nrm : 0.f;
        norm16[row] = nf;
        float hi_src = ok ? nf : B200M_SENTINEL;
        __half hi = __float2half_rn(hi_src);
        float r1 = hi_src - __half2float(hi);
        __half mid = __float2half_rn(r1);
        float r2 = r1 - __half2float(mid);
        __half lo = __float2half_rn(r2);
        ot[dim + 0] = hi;
        ot[dim + 1] = mid;
        ot[dim + 2] = lo;
        if (ok) atomicMax(max_norm_bits, __float_as_int(sqrtf(nf)));
    }
}

}  // namespace

cudaError_t launch_pack_f32(const float *aos, size_t n, size_t stride_bytes, int dim, int dp,
                            float *f32, uint8_t *valid, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    unsigned blocks = (unsigned) ((n + kWarpsPerBlock - 1) / kWarpsPerBlock);
    pack_f32_kernel<<<blocks, kWarpsPerBlock * 32, 0, st>>>(aos, n, stride_bytes / 4, dim, dp, f32, valid);
    return cudaGetLastError();
}

cudaError_t launch_tc_prepare(b200m_ctx *ctx) {
    Side &a = ctx->side[0], &b = ctx->side[1];
    TcPrep &pr = ctx->prep;
    cudaStream_t st = ctx->stream;
    const int dp = a.dp, dim = a.dim;
    const int grid = ctx->sm_count * 4;
    cudaError_t e;
    size_t part_bytes = sizeof(double) * (size_t) grid * (dp + 1);
    if ((e = pr.red.reserve(2 * part_bytes + 64)) != cudaSuccess) return e;
    if ((e = pr.mean.reserve(sizeof(float) * dp)) != cudaSuccess) return e;
    double *pa = pr.red.as<double>();
    double *pb = (double *) ((char *) pr.red.p + part_bytes);
    int *scal = (int *) ((char *) pr.red.p + 2 * part_bytes);   // [0]=maxabs bits, [1],[2]=max norm bits per side
    if ((e = cudaMemsetAsync(scal, 0, 64, st)) != cudaSuccess) return e;
    int ga = a.n ? grid : 0, gb = b.n ? grid : 0;
    if (ga) colsum_kernel<<<ga, 256, 0, st>>>(a.f32.as<float>(), a.valid.as<uint8_t>(), a.n, dp, pa);
    if (gb) colsum_kernel<<<gb, 256, 0, st>>>(b.f32.as<float>(), b.valid.as<uint8_t>(), b.n, dp, pb);
    mean_kernel<<<(dp + 127) / 128, 128, 0, st>>>(pa, ga, pb, gb, dp, pr.mean.as<float>());
    if (ga) maxabs_kernel<<<ga, 256, 0, st>>>(a.f32.as<float>(), a.valid.as<uint8_t>(), a.n, dp, dim, pr.mean.as<float>(), scal);
    if (gb) maxabs_kernel<<<gb, 256, 0, st>>>(b.f32.as<float>(), b.valid.as<uint8_t>(), b.n, dp, dim, pr.mean.as<float>(), scal);
    ctx->stats.launches += (ga ? 2 : 0) + (gb ? 2 : 0) + 1;
    int h[4] = {0, 0, 0, 0};
    if ((e = cudaMemcpyAsync(h, scal, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    float maxabs;
    memcpy(&maxabs, &h[0], 4);
    // power of two s with maxabs*s in (0.5, 1]; all-equal data (maxabs == 0) keeps s = 1
    float scale = 1.f;
    if (maxabs > 0.f && isfinite(maxabs)) {
        int ex;
        float fr = frexpf(maxabs, &ex);   // maxabs = fr * 2^ex, fr in [0.5, 1)
        if (fr == 0.5f) ex -= 1;          // exactly a power of two -> map to 1.0
        scale = ldexpf(1.f, -ex);
    }
    pr.scale = scale;
    // data whose spread cannot be brought into FP16 range (non-finite after centring, or a
    // spread so small/large that the power-of-two scale leaves float range) stays on the exact path
    pr.usable = isfinite(maxabs) && isfinite(scale) && scale > 0.f && maxabs < 1e30f && (maxabs == 0.f || maxabs > 1e-30f);
    for (int s = 0; s < 2; ++s) {
        Side &sd = ctx->side[s];
        if (sd.n_pad == 0) continue;
        size_t opb = sizeof(__half) * sd.n_pad * (size_t) sd.kp;
        if ((e = sd.op_query.reserve(opb)) != cudaSuccess) return e;
        if ((e = sd.op_train.reserve(opb)) != cudaSuccess) return e;
        if ((e = sd.norm16.reserve(sizeof(float) * sd.n_pad)) != cudaSuccess) return e;
        unsigned blocks = (unsigned) ((sd.n_pad + kWarpsPerBlock - 1) / kWarpsPerBlock);
        pack_operands_kernel<<<blocks, kWarpsPerBlock * 32, 0, st>>>(
            sd.f32.as<float>(), sd.valid.as<uint8_t>(), sd.n, sd.n_pad, dim, dp, sd.kp, pr.mean.as<float>(), scale,
            sd.op_query.as<__half>(), sd.op_train.as<__half>(), sd.norm16.as<float>(), scal + 1 + s);
        ctx->stats.launches += 1;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(h, scal, 12, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    memcpy(&pr.max_norm[0], &h[1], 4);
    memcpy(&pr.max_norm[1], &h[2], 4);
    pr.ver[0] = a.version;
    pr.ver[1] = b.version;
    pr.ready = true;
    return cudaSuccess;
}
