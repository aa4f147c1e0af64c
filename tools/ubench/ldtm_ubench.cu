// Micro-benchmark: TMEM read bandwidth of one SM (tcgen05.ld), the bound on how fast an epilogue can drain accumulators.
// One CTA allocates all 512 TMEM columns; W warps (4, 8 or 16: 1, 2 or 4 per scheduler / TMEM lane quarter) each read their
// lane quarter's columns over and over with tcgen05.ld.32x32b of width x16 / x32 / x64 / x128, `depth` loads per wait.
// Prints bytes per SM cycle.  (The FPFH candidate kernel has to drain 128 rows x 256 FP32 columns = 128 KB per CTA and tile
// against 384 tensor-pipe cycles of MMA work per tile.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_ubench ldtm_ubench.cu && ./ldtm_ubench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define LD_ARGS16(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7]), \
                        "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15])

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : LD_ARGS16(r, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : LD_ARGS16(r, 0), LD_ARGS16(r, 16) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
                 "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
                 "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                 : LD_ARGS16(r, 0), LD_ARGS16(r, 16), LD_ARGS16(r, 32), LD_ARGS16(r, 48) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// WIDTH: columns per load (16, 32, 64); DEPTH: loads in flight per wait
template <int WIDTH, int DEPTH>
__global__ void __launch_bounds__(512, 1) k(long long *cyc, int iters, uint32_t *sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t) __cvta_generic_to_shared(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_slot + ((uint32_t) ((warp & 3) * 32) << 16);
    uint32_t r[WIDTH * DEPTH];
    uint32_t acc = 0;
    uint32_t col = (uint32_t) ((warp >> 2) * 128) & 511u;   // warps of one lane quarter start in different column ranges
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const uint32_t c = (col + (uint32_t) (d * WIDTH)) & 511u;
            if (WIDTH == 16) ld16(base + c, r + d * WIDTH);
            else if (WIDTH == 32) ld32(base + c, r + d * WIDTH);
            else ld64(base + c, r + d * WIDTH);
        }
        ld_wait();
        acc ^= r[0] ^ r[WIDTH * DEPTH - 1];
        col = (col + (uint32_t) (WIDTH * DEPTH)) & 511u;
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_slot) : "memory");
}

template <int WIDTH, int DEPTH>
void run(int warps, long long *d_cyc, uint32_t *d_sink) {
    const int iters = 2000;
    k<WIDTH, DEPTH><<<1, warps * 32>>>(d_cyc, iters, d_sink);
    k<WIDTH, DEPTH><<<1, warps * 32>>>(d_cyc, iters, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double bytes = (double) warps * iters * WIDTH * DEPTH * 32 * 4;
    printf("x%-3d depth %d warps %2d: %8lld cycles, %6.1f B/cycle/SM, %6.1f cycles per 128x256 FP32 tile (128 KB)%s\n", WIDTH, DEPTH, warps, c,
           bytes / (double) c, 131072.0 / (bytes / (double) c), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long *d_cyc;
    uint32_t *d_sink;
    cudaMalloc(&d_cyc, 1024);
    cudaMalloc(&d_sink, 64);
    for (int w : {4, 8, 16}) {
        run<16, 1>(w, d_cyc, d_sink);
        run<16, 4>(w, d_cyc, d_sink);
        run<32, 1>(w, d_cyc, d_sink);
        run<32, 2>(w, d_cyc, d_sink);
        run<64, 1>(w, d_cyc, d_sink);
    }
    return 0;
}
