// Micro-benchmarks that inform the design of the persistent sweep kernel: FP64 dependent-issue latency,
// DFMA vs DMMA (mma.sync.m8n8k4.f64) throughput, cluster barrier cost, DSMEM push + remote mbarrier arrive, L2 flag
// round trip between CTAs, and whether a cooperative launch may carry a cluster dimension.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ttc_ubench ttc_ubench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); std::exit(1); } } while (0)

template <int OP>
__global__ void k_dep(double* out, long long* cyc, int iters, double a, double b) {
    double x = a + threadIdx.x * 1e-9, y = b;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) x = __dadd_rn(x, y);
            if (OP == 1) x = __dmul_rn(x, y);
            if (OP == 2) x = __fma_rn(x, y, y);
            if (OP == 3) { x = __dmul_rn(x, y); x = __dadd_rn(x, y); }
            if (OP == 4) x = y / x;
            if (OP == 5) x = __longlong_as_double(__shfl_sync(0xffffffffu, __double_as_longlong(x), (threadIdx.x + 1) & 31));
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: NACC independent accumulators per thread
template <int OP>
__global__ void k_tput(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = a + u + threadIdx.x * 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (OP == 0) x[u] = __fma_rn(x[u], b, b);
                if (OP == 1) { x[u] = __dmul_rn(x[u], b); }
                if (OP == 2) { x[u] = __dadd_rn(x[u], b); }
            }
    }
    double s = 0; for (int u = 0; u < 8; ++u) s += x[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// DMMA: mma.sync.aligned.m8n8k4.row.col.f64: D(8x8) += A(8x4) B(4x8): 512 flops per warp instruction
__global__ void k_dmma(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int u = 0; u < 8; ++u) { c[u][0] = 0; c[u][1] = 0; }
    double av = a + threadIdx.x * 1e-9, bv = b;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(av), "d"(bv));
    }
    double s = 0; for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_cluster_sync(long long* cyc, int iters) {
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) cl.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// push-style exchange: every CTA stores 32 B into every CTA's shared memory, then arrives on every CTA's mbarrier
// (remote arrive, release.cluster); each CTA waits on its own barrier (acquire.cluster) and block-syncs.
__global__ void k_cluster_mbar(long long* cyc, int iters, double* sink) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ double box[2][16][4];
    __shared__ unsigned long long bar[2];
    const int cs = cl.num_blocks(), me = cl.block_rank();
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            unsigned a = (unsigned)__cvta_generic_to_shared(&bar[b]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(cs));
        }
    }
    __syncthreads();
    cl.sync();
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const int b = i & 1, ph = (i >> 1) & 1;
        if (threadIdx.x < cs) {
            const int dst = threadIdx.x;
            double* rb = cl.map_shared_rank(&box[b][me][0], dst);
            rb[0] = i + me; rb[1] = 1; rb[2] = 2; rb[3] = 3;
            unsigned la = (unsigned)__cvta_generic_to_shared(&bar[b]);
            unsigned ra;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(dst));
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
        }
        if (threadIdx.x == 0) {
            unsigned la = (unsigned)__cvta_generic_to_shared(&bar[b]);
            unsigned ok = 0;
            while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(la), "r"(ph) : "memory");
        }
        __syncthreads();
        for (int c = 0; c < cs; ++c) acc += box[b][c][0];
    }
    long long t1 = clock64();
    cl.sync();
    if (threadIdx.x == 0) sink[blockIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// L2 flag ping-pong between CTA 0 and CTA 1 (different SMs)
__global__ void k_pingpong(unsigned long long* flags, long long* cyc, int iters) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    if (me > 1) return;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (me == 0) {
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flags), "l"((unsigned long long)i) : "memory");
            unsigned long long v = 0;
            while (v < (unsigned long long)i) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + 16) : "memory");
        } else {
            unsigned long long v = 0;
            while (v < (unsigned long long)i) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flags) : "memory");
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flags + 16), "l"((unsigned long long)i) : "memory");
        }
    }
    long long t1 = clock64();
    if (me == 0) cyc[0] = t1 - t0;
}
// all-to-all flag barrier among G CTAs: each CTA stores its flag, 32 lanes poll the G flags
__global__ void k_flagbarrier(unsigned long long* flags, long long* cyc, int iters) {
    const int G = gridDim.x, me = blockIdx.x;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flags + 16 * me), "l"((unsigned long long)i) : "memory");
        if (threadIdx.x < G) {
            unsigned long long v = 0;
            while (v < (unsigned long long)i) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + 16 * threadIdx.x) : "memory");
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && me == 0) cyc[0] = t1 - t0;
}
__global__ void k_coop_cluster(int* out) {
    cg::grid_group g = cg::this_grid();
    cg::cluster_group cl = cg::this_cluster();
    if (threadIdx.x == 0) atomicAdd(out, 1);
    g.sync();
    cl.sync();
    if (g.thread_rank() == 0) out[1] = out[0];
}

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    std::printf("device %s  SMs %d  clock %.0f MHz\n", pr.name, pr.multiProcessorCount, clk / 1e3);
    double* out; long long* cyc; CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&cyc, 64));
    long long hc = 0;
    const char* names[] = {"DADD", "DMUL", "DFMA", "DMUL+DADD", "DDIV", "SHFL64"};
    for (int warps : {1, 4, 8, 16}) {
        for (int op = 0; op < 6; ++op) {
            const int iters = 200;
            for (int rep = 0; rep < 2; ++rep) {
                switch (op) {
                    case 0: k_dep<0><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 1e-3); break;
                    case 1: k_dep<1><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 1.0000001); break;
                    case 2: k_dep<2><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 0.5); break;
                    case 3: k_dep<3><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 0.5); break;
                    case 4: k_dep<4><<<1, 32 * warps>>>(out, cyc, iters, 1.5, 1.25); break;
                    case 5: k_dep<5><<<1, 32 * warps>>>(out, cyc, iters, 1.5, 1.25); break;
                }
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
            std::printf("dependent chain  %-10s warps/SM %2d : %.1f cycles per step\n", names[op], warps, (double)hc / (iters * 16.0));
        }
    }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int op = 0; op < 4; ++op) {
        const int iters = 4000, blocks = pr.multiProcessorCount * 4, threads = 256;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            if (op == 0) k_tput<0><<<blocks, threads>>>(out, iters, 1.0, 0.999999);
            if (op == 1) k_tput<1><<<blocks, threads>>>(out, iters, 1.0, 0.999999);
            if (op == 2) k_tput<2><<<blocks, threads>>>(out, iters, 1.0, 1e-9);
            if (op == 3) k_dmma<<<blocks, threads>>>(out, iters, 1.0, 1e-3);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        double flops;
        if (op == 3) flops = (double)blocks * (threads / 32) * iters * 8.0 * 512.0;
        else flops = (double)blocks * threads * iters * 32.0 * (op == 0 ? 2.0 : 1.0);
        const char* nm[] = {"DFMA (2 flop)", "DMUL", "DADD", "DMMA m8n8k4 f64 (mma.sync)"};
        std::printf("throughput %-28s : %.2f TFLOP/s  (%.3f ms)\n", nm[op], flops / best / 1e9, best);
    }
    // cluster barriers
    for (int cs : {2, 4, 8, 16}) {
        for (int threads : {256, 512}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * 8); cfg.blockDim = dim3(threads);
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            if (cs > 8) { cudaFuncSetAttribute(k_cluster_sync, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); cudaFuncSetAttribute(k_cluster_mbar, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); }
            const int iters = 2000;
            cudaError_t e = cudaLaunchKernelEx(&cfg, k_cluster_sync, cyc, iters);
            if (e != cudaSuccess) { std::printf("cluster %d launch failed: %s\n", cs, cudaGetErrorString(e)); (void)cudaGetLastError(); continue; }
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
            std::printf("barrier.cluster   cluster %2d x %3d threads : %.0f cycles\n", cs, threads, (double)hc / iters);
            e = cudaLaunchKernelEx(&cfg, k_cluster_mbar, cyc, iters, out);
            if (e != cudaSuccess) { std::printf("cluster %d mbar launch failed: %s\n", cs, cudaGetErrorString(e)); (void)cudaGetLastError(); continue; }
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
            std::printf("DSMEM push+mbarrier cluster %2d x %3d threads : %.0f cycles (incl. 2 __syncthreads-equivalents)\n", cs, threads, (double)hc / iters);
        }
    }
    unsigned long long* flags; CK(cudaMalloc(&flags, 16 * 8 * 256)); CK(cudaMemset(flags, 0, 16 * 8 * 256));
    {
        const int iters = 2000;
        k_pingpong<<<2, 32>>>(flags, cyc, iters); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
        std::printf("L2 flag ping-pong (release/acquire.gpu) round trip : %.0f cycles\n", (double)hc / iters);
        for (int G : {8, 16, 32}) {
            CK(cudaMemset(flags, 0, 16 * 8 * 256));
            k_flagbarrier<<<G, 64>>>(flags, cyc, iters); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
            std::printf("all-to-all flag barrier among %2d CTAs : %.0f cycles\n", G, (double)hc / iters);
        }
    }
    {
        int* o; CK(cudaMalloc(&o, 8)); CK(cudaMemset(o, 0, 8));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(128); cfg.blockDim = dim3(256);
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 2;
        cudaFuncSetAttribute(k_coop_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_coop_cluster, o);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        int ho[2] = {0, 0};
        if (e == cudaSuccess) cudaMemcpy(ho, o, 8, cudaMemcpyDeviceToHost);
        std::printf("cooperative launch with cluster dimension 16 (128 CTAs): %s (counter %d)\n", e == cudaSuccess ? "OK" : cudaGetErrorString(e), ho[1]);
        (void)cudaGetLastError();
    }
    return 0;
}
