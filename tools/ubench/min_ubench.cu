// Micro-benchmark: cycles for one warp to decide "is any of 128 registers below thr" (the candidate kernel's fast
// path), for several instruction choices, with 1 or 2 such warps per scheduler.  The 128 values stay in registers; a
// loop-carried value feeds every reduction chain so that nothing can be hoisted out of the timing loop.
#include <cstdio>
#include <cuda_runtime.h>
#define N 128
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ int imin3(int a, int b, int c) { return min(min(a, b), c); }

template <int V>
__global__ void k(const float *in, float *out, long long *cyc, int iters, float thr) {
    float r[N];
    for (int i = 0; i < N; ++i) r[i] = in[(threadIdx.x * N + i) % 4096];
    int hits = 0;
    float carry = 1e30f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (V == 0) {   // four chains of FMNMX3
            float m[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                m[c] = min3(carry, r[32 * c], r[32 * c + 1]);
#pragma unroll
                for (int i = 2; i < 32; i += 2) m[c] = min3(m[c], r[32 * c + i], r[32 * c + i + 1]);
            }
            float mm = fminf(min3(m[0], m[1], m[2]), m[3]);
            if (mm < thr) hits++;
            carry = mm + 1e30f;
        } else if (V == 1) {   // four chains of two-input FMNMX
            float m[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                asm volatile("min.f32 %0, %1, %2;" : "=f"(m[c]) : "f"(carry), "f"(r[32 * c]));
#pragma unroll
                for (int i = 1; i < 32; ++i) asm volatile("min.f32 %0, %1, %2;" : "=f"(m[c]) : "f"(m[c]), "f"(r[32 * c + i]));
            }
            float mm = fminf(fminf(m[0], m[1]), fminf(m[2], m[3]));
            if (mm < thr) hits++;
            carry = mm + 1e30f;
        } else if (V == 2) {   // FADD (a - thr) + three-input OR of the sign words, four accumulators
            unsigned acc[4];
            const float th = thr + carry * 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] = 0;
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                float a = r[i] - th, b = r[i + 1] - th;
                acc[(i / 2) & 3] |= __float_as_uint(a) | __float_as_uint(b);
            }
            unsigned all = acc[0] | acc[1] | acc[2] | acc[3];
            if (all >> 31) hits++;
            carry = __uint_as_float(all & 0x3f800000u);
        } else if (V == 3) {   // predicate chain: setp.lt.or
            unsigned any;
            const float th = thr + carry * 0.f;
            asm volatile("{\n.reg .pred p;\nsetp.lt.f32 p, %1, %2;\n" : "=r"(any) : "f"(r[0]), "f"(th));
#pragma unroll
            for (int i = 1; i < N; ++i) asm volatile("setp.lt.or.f32 p, %0, %1, p;\n" ::"f"(r[i]), "f"(th));
            asm volatile("selp.u32 %0, 1, 0, p;\n}" : "=r"(any));
            if (any) hits++;
            carry = any ? 1.f : 2.f;
        } else if (V == 4) {   // four chains of integer VIMNMX3
            int m[4];
            const int ci = __float_as_int(carry);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                m[c] = imin3(ci, __float_as_int(r[32 * c]), __float_as_int(r[32 * c + 1]));
#pragma unroll
                for (int i = 2; i < 32; i += 2) m[c] = imin3(m[c], __float_as_int(r[32 * c + i]), __float_as_int(r[32 * c + i + 1]));
            }
            int mm = min(imin3(m[0], m[1], m[2]), m[3]);
            if (mm < __float_as_int(thr)) hits++;
            carry = __int_as_float(mm | 0x40000000);
        } else if (V == 5) {   // FADD + LOP3, two accumulators
            unsigned acc0 = 0, acc1 = 0;
            const float th = thr + carry * 0.f;
#pragma unroll
            for (int i = 0; i < N; i += 4) {
                float a = r[i] - th, b = r[i + 1] - th, c = r[i + 2] - th, d = r[i + 3] - th;
                acc0 |= __float_as_uint(a) | __float_as_uint(b);
                acc1 |= __float_as_uint(c) | __float_as_uint(d);
            }
            unsigned all = acc0 | acc1;
            if (all >> 31) hits++;
            carry = __uint_as_float(all & 0x3f800000u);
        } else if (V == 6) {   // half the elements FMNMX3, half FADD+LOP3 (both pipes busy)
            float m0 = min3(carry, r[0], r[1]), m1 = min3(carry, r[32], r[33]);
#pragma unroll
            for (int i = 2; i < 32; i += 2) { m0 = min3(m0, r[i], r[i + 1]); m1 = min3(m1, r[32 + i], r[32 + i + 1]); }
            unsigned acc0 = 0, acc1 = 0;
            const float th = thr + carry * 0.f;
#pragma unroll
            for (int i = 64; i < N; i += 4) {
                float a = r[i] - th, b = r[i + 1] - th, c = r[i + 2] - th, d = r[i + 3] - th;
                acc0 |= __float_as_uint(a) | __float_as_uint(b);
                acc1 |= __float_as_uint(c) | __float_as_uint(d);
            }
            float mm = fminf(m0, m1);
            if (mm < thr || ((acc0 | acc1) >> 31)) hits++;
            carry = mm + 1e30f;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x % 32 == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = hits + r[5] + carry;
}

template <int V>
void run(const char *name, const float *in, float *out, long long *cyc, int threads) {
    const int iters = 2000;
    k<V><<<148, threads>>>(in, out, cyc, iters, -1e30f);
    cudaDeviceSynchronize();
    k<V><<<148, threads>>>(in, out, cyc, iters, -1e30f);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[64];
    cudaMemcpy(h, cyc, sizeof(long long) * (threads / 32), cudaMemcpyDeviceToHost);
    printf("%-34s %d warp(s)/scheduler: %6.1f cycles per 128-value batch per warp  (%s)\n", name, threads / 128,
           (double) h[0] / iters, cudaGetErrorString(e));
}

int main() {
    float *in, *out;
    long long *cyc;
    cudaMalloc(&in, 4096 * 4);
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 32 * 8);
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = (float) (i % 97) * 0.37f + 1.f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int threads : {128, 256}) {
        run<0>("FMNMX3, 4 chains", in, out, cyc, threads);
        run<1>("FMNMX (2-input), 4 chains", in, out, cyc, threads);
        run<2>("FADD + LOP3 sign-OR, 4 acc", in, out, cyc, threads);
        run<5>("FADD + LOP3 sign-OR, 2 acc", in, out, cyc, threads);
        run<3>("FSETP.LT.OR predicate chain", in, out, cyc, threads);
        run<4>("VIMNMX3 (int), 4 chains", in, out, cyc, threads);
        run<6>("half FMNMX3 / half FADD+LOP3", in, out, cyc, threads);
    }
    return 0;
}
