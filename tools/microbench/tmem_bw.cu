// Microbenchmark: tensor-memory load/store throughput of the .32x32b.x8 shape used by the replay kernel's
// operand stack (tcgen05.ld / tcgen05.st), against shared-memory LDS.128 x2 of the same 1 KB per warp.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_bw tmem_bw.cu ; run: ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0 = LDTM, 1 = STTM, 2 = LDS.128 x2, 3 = LDTM + dependent FADD chain (latency)
__global__ void __launch_bounds__(1024, 1) k(int iters, int cols, float *out, long long *cyc) {
    __shared__ uint32_t base_slot;
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&base_slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0 = base_slot + ((uint32_t)(warp & 3) * 32u << 16) + (uint32_t)(warp >> 2) * 16u;
    float a[8] = {1, 2, 3, 4, 5, 6, 7, 8}, s = 0.f;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(t0), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(t0 + 8), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long c0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const uint32_t t = t0 + ((i & 1) << 3);
        if (MODE == 0 || MODE == 3) {
            float v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\ntcgen05.wait::ld.sync.aligned;"
                         : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(t) : "memory");
            if (MODE == 3) { s += v[0]; asm volatile("" : "+f"(s)); a[0] = s; }
            else { s += v[0] + v[7]; }
        } else if (MODE == 1) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(t), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]) : "memory");
        } else {
            const float4 x = *reinterpret_cast<const float4 *>(sm + (i & 1) * 2048 + warp * 256 % 1024 + lane * 4);
            const float4 y = *reinterpret_cast<const float4 *>(sm + (i & 1) * 2048 + warp * 256 % 1024 + 128 + lane * 4);
            s += x.x + y.w;
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const long long c1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + a[0];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_slot), "r"(cols) : "memory");
}

template <int MODE>
void run(const char *name, int warps) {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    k<MODE><<<148, warps * 32, 32768>>>(100, 512, out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, warps * 32, 32768>>>(iters, 512, out, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double bytes_per_clk = (double)iters * warps * 1024.0 / (double)h;
    printf("%-10s warps/SM %2d: %8.1f cycles/iter/warp-set, %7.1f B/clk/SM, %.3f ms  (%s)\n", name, warps, (double)h / iters,
           bytes_per_clk, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {1, 4, 8, 16, 32}) run<0>("LDTM.x8", w);
    for (int w : {1, 4, 8, 16, 32}) run<1>("STTM.x8", w);
    for (int w : {1, 4, 8, 16, 32}) run<2>("LDS.128x2", w);
    for (int w : {1, 4}) run<3>("LDTM lat", w);
    return 0;
}
