// Microbenchmark: cost of one device-wide barrier among G co-resident CTAs of 1024 threads on B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=true barrier_bench.cu -o barrier_bench -lcudadevrt
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ unsigned ld_rlx(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rel(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_rlx(unsigned* p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// A: one atomic counter, everyone polls it
__device__ void barA(unsigned* f, unsigned& e) {
    __syncthreads(); e++;
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(f, 1u); while (ld_acq(f) < e * gridDim.x) {} __threadfence(); }
    __syncthreads();
}
// B: per-CTA flag in one array, everyone polls all
__device__ void barB(unsigned* f, unsigned& e) {
    __syncthreads(); e++;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) st_rel(f + blockIdx.x, e);
        bool ok; do { ok = true; for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) ok = ok && ld_rlx(f + i) >= e; } while (!__all_sync(~0u, ok));
    }
    __syncthreads();
}
// D: master gathers arrivals (one 128B line per CTA), then releases each CTA through its private line
__device__ void barD(unsigned* f, unsigned& e) {
    __syncthreads(); e++;
    const int G = gridDim.x, c = blockIdx.x;
    unsigned* arrive = f;             // [G][32]
    unsigned* release = f + 32 * 256; // [G][32]
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) st_rel(arrive + c * 32, e);
        if (c == 0) {
            bool ok; do { ok = true; for (int i = threadIdx.x; i < G; i += 32) ok = ok && ld_rlx(arrive + i * 32) >= e; } while (!__all_sync(~0u, ok));
            for (int i = threadIdx.x; i < G; i += 32) st_rlx(release + i * 32, e);
        }
        if (threadIdx.x == 0) while (ld_rlx(release + c * 32) < e) {}
    }
    __syncthreads();
}
// F: like B but the writer only uses a relaxed store after a block-level barrier (no MEMBAR): lower bound
__device__ void barF(unsigned* f, unsigned& e) {
    __syncthreads(); e++;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) st_rlx(f + blockIdx.x, e);
        bool ok; do { ok = true; for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) ok = ok && ld_rlx(f + i) >= e; } while (!__all_sync(~0u, ok));
    }
    __syncthreads();
}
template <int V>
__global__ void __launch_bounds__(1024) k(unsigned* f, float* data, int iters, int work) {
    unsigned e = 0;
    cg::grid_group g = cg::this_grid();
    for (int it = 0; it < iters; it++) {
        if (work) { size_t i = ((size_t)blockIdx.x * 1024 + threadIdx.x); float v; asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(data + ((i * 7 + it * 1024) & ((1 << 24) - 1))) : "memory"); data[i] = v + 1.f; }
        if (V == 0) barA(f, e);
        else if (V == 1) barB(f, e);
        else if (V == 3) barD(f, e);
        else if (V == 4) g.sync();
        else if (V == 5) barF(f, e);
    }
}
template <int V> void run(const char* name, int G, int work) {
    unsigned* f; float* d;
    cudaMalloc(&f, 1 << 20); cudaMalloc(&d, sizeof(float) << 24);
    cudaMemset(d, 0, sizeof(float) << 24);
    int iters = 2000;
    void* args[] = {&f, &d, &iters, &work};
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaMemset(f, 0, 1 << 20);
        cudaEventRecord(a);
        cudaError_t err = cudaLaunchCooperativeKernel((void*)k<V>, dim3(G), dim3(1024), args, 0, 0);
        cudaEventRecord(b); cudaEventSynchronize(b);
        if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("%s G=%d launch failed\n", name, G); return; }
        float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
    }
    printf("%-34s G=%3d work=%d  %.3f us per barrier\n", name, G, work, 1e3 * best / iters);
    cudaFree(f); cudaFree(d);
}
int main() {
    for (int work = 0; work < 2; work++)
        for (int G : {148, 32, 8}) {
            run<0>("A atomic counter + acquire poll", G, work);
            run<1>("B flag per CTA, all poll all", G, work);
            run<3>("D master gather / private release", G, work);
            run<4>("E cooperative_groups grid.sync", G, work);
            run<5>("F like B, relaxed store (no membar)", G, work);
        }
    return 0;
}
