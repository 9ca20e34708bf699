// Microbenchmark (B200): what one dependent "pass" costs under the three ways of separating passes
//   (a) a kernel launch boundary in one stream (plain launches and a CUDA graph),
//   (b) barrier.cluster (arrive.release / wait.acquire) among C CTAs x 1024 threads, with and without a DSMEM read per pass,
//   (c) a plain __syncthreads pass in one CTA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 cluster_bench.cu -o cluster_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); } } while (0)

__global__ void empty_kernel(float* p) { if (p && threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.f; }
__global__ void small_kernel(float* p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = p[i] * 1.0001f + 1.f; }

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) { unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ float ld_dsmem(unsigned addr) { float v; asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }

__device__ __forceinline__ void cluster_sync_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_dsmem(unsigned addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
// mode 0: barrier only (release/acquire); 1: scattered dsmem read + smem write + barrier; 2: relaxed barrier only;
// 3: local smem write + barrier; 4: remote store by 1/8 of the threads (coalesced) + barrier; 5: remote store by every thread (coalesced);
// 6: coalesced dsmem read + smem write + barrier; 7: like 4 with the relaxed barrier + fence.acq_rel.cluster
__global__ void __launch_bounds__(1024) cluster_kernel(int iters, int mode, float* out) {
    __shared__ float buf[2][1024];
    const unsigned rank = cluster_ctarank(), C = cluster_nctarank();
    buf[0][threadIdx.x] = (float)threadIdx.x; buf[1][threadIdx.x] = 0.f;
    cluster_sync_all();
    const unsigned local0 = (unsigned)__cvta_generic_to_shared(&buf[0][0]);
    const unsigned peer = mapa(local0, (rank + 1) % C);
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        const int cur = it & 1;
        if (mode == 1) {
            float v = ld_dsmem(peer + (cur * 1024 + ((threadIdx.x * 33) & 1023)) * 4);
            buf[cur ^ 1][threadIdx.x] = v + 1.f;
            acc += v;
        } else if (mode == 3) {
            float v = buf[cur][(threadIdx.x * 33) & 1023];
            buf[cur ^ 1][threadIdx.x] = v + 1.f;
            acc += v;
        } else if (mode == 4 || mode == 7) {
            float v = buf[cur][(threadIdx.x * 33) & 1023];
            buf[cur ^ 1][threadIdx.x] = v + 1.f;
            if ((threadIdx.x & 7) == 0) st_dsmem(peer + (cur * 1024 + (threadIdx.x >> 3)) * 4, v);
            acc += v;
        } else if (mode == 5) {
            float v = buf[cur][(threadIdx.x * 33) & 1023];
            st_dsmem(peer + ((cur ^ 1) * 1024 + threadIdx.x) * 4, v + 1.f);
            acc += v;
        } else if (mode == 6) {
            float v = ld_dsmem(peer + (cur * 1024 + threadIdx.x) * 4);
            buf[cur ^ 1][threadIdx.x] = v + 1.f;
            acc += v;
        }
        if (mode == 2) cluster_sync_relaxed();
        else if (mode == 7) { asm volatile("fence.acq_rel.cluster;" ::: "memory"); cluster_sync_relaxed(); }
        else cluster_sync_all();
    }
    if (out && threadIdx.x == 0) out[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(1024) cta_kernel(int iters, float* out) {
    __shared__ float buf[2][1024];
    buf[0][threadIdx.x] = (float)threadIdx.x; buf[1][threadIdx.x] = 0.f;
    __syncthreads();
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        const int cur = it & 1;
        float v = buf[cur][(threadIdx.x * 33) & 1023];
        buf[cur ^ 1][threadIdx.x] = v + 1.f;
        acc += v;
        __syncthreads();
    }
    if (out && threadIdx.x == 0) out[blockIdx.x] = acc;
}

static float time_ms(cudaStream_t s, void (*f)(cudaStream_t, void*), void* ctx, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(s, ctx); cudaStreamSynchronize(s);
    cudaEventRecord(a, s);
    for (int i = 0; i < reps; i++) f(s, ctx);
    cudaEventRecord(b, s); cudaStreamSynchronize(s);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

struct Ctx { float* d; int n; int iters; int mode; int C; int threads; cudaGraphExec_t ge; };

static void launch_chain_empty(cudaStream_t s, void* c) { Ctx* x = (Ctx*)c; for (int i = 0; i < 200; i++) empty_kernel<<<1, 32, 0, s>>>(x->d); }
static void launch_chain_small(cudaStream_t s, void* c) { Ctx* x = (Ctx*)c; for (int i = 0; i < 200; i++) small_kernel<<<(x->n + 255) / 256, 256, 0, s>>>(x->d, x->n); }
static void launch_graph(cudaStream_t s, void* c) { Ctx* x = (Ctx*)c; cudaGraphLaunch(x->ge, s); }
static void launch_cluster(cudaStream_t s, void* c) {
    Ctx* x = (Ctx*)c;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(x->C); cfg.blockDim = dim3(x->threads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = x->C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, cluster_kernel, x->iters, x->mode, x->d));
}
static void launch_cta(cudaStream_t s, void* c) { Ctx* x = (Ctx*)c; cta_kernel<<<1, 1024, 0, s>>>(x->iters, x->d); }

int main() {
    cudaStream_t s; cudaStreamCreate(&s);
    Ctx c; c.n = 1 << 16; c.iters = 2000; c.mode = 0; c.C = 8;
    cudaMalloc(&c.d, (1 << 20) * 4); cudaMemset(c.d, 0, (1 << 20) * 4);
    printf("launch gap, 200 dependent empty kernels: %.2f us each\n", 1e3f * time_ms(s, launch_chain_empty, &c, 5) / 200);
    for (int n : {1 << 12, 1 << 16, 1 << 20}) { c.n = n; printf("launch chain, 200 dependent small kernels over %d floats: %.2f us each\n", n, 1e3f * time_ms(s, launch_chain_small, &c, 5) / 200); }
    {   // the same 200-kernel chain as a graph
        c.n = 1 << 12;
        cudaGraph_t g; cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal); launch_chain_small(s, &c); cudaStreamEndCapture(s, &g);
        CK(cudaGraphInstantiate(&c.ge, g, 0));
        printf("graph of 200 dependent small kernels (4096 floats): %.2f us each\n", 1e3f * time_ms(s, launch_graph, &c, 5) / 200);
        c.n = 1 << 20;
        cudaGraph_t g2; cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal); launch_chain_small(s, &c); cudaStreamEndCapture(s, &g2);
        CK(cudaGraphInstantiate(&c.ge, g2, 0));
        printf("graph of 200 dependent small kernels (1M floats): %.2f us each\n", 1e3f * time_ms(s, launch_graph, &c, 5) / 200);
    }
    CK(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int C : {8, 16}) {
        c.C = C;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C); cfg.blockDim = dim3(1024);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, cluster_kernel, &cfg);
        printf("cluster size %d: max active clusters %d (%s)\n", C, nc, cudaGetErrorString(e));
        if (e != cudaSuccess || nc < 1) { cudaGetLastError(); continue; }
        for (int threads : {256, 1024}) {
            c.threads = threads;
            for (int mode = 0; mode < 8; mode++) {
                c.mode = mode;
                float ms = time_ms(s, launch_cluster, &c, 3);
                printf("  cluster %2d x %4d thr, mode %d: %.3f us per pass\n", C, threads, mode, 1e3f * ms / c.iters);
            }
        }
    }
    printf("one CTA x 1024 thr, smem read/write + __syncthreads: %.3f us per pass\n", 1e3f * time_ms(s, launch_cta, &c, 3) / c.iters);
    CK(cudaGetLastError());
    return 0;
}
