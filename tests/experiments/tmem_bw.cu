// Microbenchmark (experiment, not product): TMEM read bandwidth per SM and its overlap with the MUFU pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/experiments/build/tmem_bw tests/experiments/tmem_bw.cu
// mode 0: every warp streams tcgen05.ld 32x32b.x32 over 128 columns; mode 1: 128 ex2 per thread per iteration;
// mode 2: both, software-pipelined (the loads of chunk c+1 under the exponentials of chunk c).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MODE>
__global__ void __launch_bounds__(512) k(int iters, float* sink, long long* clk) {
    extern __shared__ unsigned char dyn[];      // 200 KB requested: exactly one CTA per SM
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    float acc = 0.f;
    uint32_t a[32], b[32];
    const long long t0 = clock64();
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 4; c += 2) {
                ld32(tmem + c * 32, a);
                ld32(tmem + c * 32 + 32, b);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t xa = 0;
#pragma unroll
                for (int q = 0; q < 32; ++q) xa ^= a[q] ^ b[q];
                acc += __uint_as_float(xa);
            }
        }
    } else if (MODE == 1) {
        float x = (float)lane * 1e-3f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 128; ++i) acc += ex2(x + (float)i);
            x += 1e-6f;
        }
    } else {
        ld32(tmem, a);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t* cur = (c & 1) ? b : a;
                uint32_t* nxt = (c & 1) ? a : b;
                ld32(tmem + ((c + 1) & 3) * 32, nxt);
#pragma unroll
                for (int i = 0; i < 32; ++i) acc += ex2(__uint_as_float(cur[i]) * 1e-30f);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
        }
    }
    const long long t1 = clock64();
    if (acc == 123.456f) sink[0] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int MODE>
static void run(const char* name, int ctas_per_sm, int sms, int iters) {
    float* sink;
    long long* clk;
    cudaMalloc(&sink, 4);
    cudaMalloc(&clk, 8 * sms * ctas_per_sm);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<MODE><<<sms, 128 * ctas_per_sm, 200 * 1024>>>(iters, sink, clk);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<sms, 128 * ctas_per_sm, 200 * 1024>>>(iters, sink, clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c0 = 0;
    cudaMemcpy(&c0, clk, 8, cudaMemcpyDeviceToHost);
    const double per_iter = (double)c0 / iters;                     // clocks per iteration of one CTA (4 warps x 128 columns)
    const double bytes_sm = (double)ctas_per_sm * 4 * 32 * 128 * 4;   // TMEM bytes read per SM per iteration
    const double exps_sm = (double)ctas_per_sm * 128 * 128;
    printf("%-28s %d x 4 warps per SM: %8.1f clk/iter  -> %6.1f B/clk/SM TMEM, %5.2f ex2/clk/SM   (%.3f ms, err %s)\n", name, ctas_per_sm,
           per_iter, MODE != 1 ? bytes_sm / per_iter : 0.0, MODE != 0 ? exps_sm / per_iter : 0.0, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
    cudaFree(clk);
}

int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    printf("%s, %d SMs\n", pr.name, sms);
    for (int c = 1; c <= 4; ++c) run<0>("tcgen05.ld only", c, sms, 2000);
    for (int c = 1; c <= 4; ++c) run<1>("ex2 only", c, sms, 2000);
    for (int c = 1; c <= 4; ++c) run<2>("ld(c+1) under ex2(c)", c, sms, 2000);
    return 0;
}
