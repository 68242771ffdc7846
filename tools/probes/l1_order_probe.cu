// Micro-probe: does an L1 hit of one warp wait behind outstanding L2 / DRAM misses of ANOTHER warp of the same SM?
// Warp 0 walks 64 L1-resident 512-byte records (load, modify, store: the quality chain's access pattern) and is timed;
// warp 1 meanwhile issues one 256-byte fetch per step from a 4 GB area (misses to DRAM) or a 2 MB area (L2 hits) by
// LDG.cg, by cp.async.cg (16 lanes x 16 B, the decoder's base-context ring) or by one cp.async.bulk (TMA).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/l1_order_probe tools/probes/l1_order_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32;
typedef unsigned long long u64;

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void probe(uint4 *tab, u32 kmask, const uint4 *area, u64 area_mask16, int iters, long long *out, u32 *sink) {
    __shared__ volatile int done;
    __shared__ __align__(128) uint4 ring[8][16];
    __shared__ __align__(8) u64 mbar[8];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        done = 0;
        for (int i = 0; i < 8; i++) asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(smem_u32(&mbar[i])));
    }
    __syncthreads();
    if (wid == 0) {
        u32 ctx = 0, acc = 0;
        long long t0 = 0;
        for (int it = -4 * (int)(kmask + 1); it < iters; it++) {
            if (it == 0) t0 = clock64();
            uint4 *p = tab + (size_t)ctx * 32 + lane;
            uint4 M;
            asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(M.x), "=r"(M.y), "=r"(M.z), "=r"(M.w) : "l"(p) : "memory");
            acc += M.y;
            M.x += 1;
            asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(M.x), "r"(M.y), "r"(M.z), "r"(M.w) : "memory");
            u32 nx = __shfl_sync(0xffffffffu, M.w, 0);
            ctx = (ctx * 5 + 1 + (nx & 0)) & kmask;
        }
        long long t1 = clock64();
        if (lane == 0) { out[0] = t1 - t0; sink[0] = acc; done = 1; }
    } else {
        u64 x = 88172645463325252ull + lane * 0;          // same sequence on every lane
        u32 acc = 0, step = 0;
        long long n = 0;
        while (!done) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            const uint4 *src = area + ((x & area_mask16) & ~15ull);          // 256-byte aligned
            if (MODE == 1) {                              // LDG.cg, dependent (one outstanding request)
                uint4 v;
                if (lane < 16) { asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + lane) : "memory"); acc += v.x; }
                x += __shfl_sync(0xffffffffu, acc, 0) & 0;
            } else if (MODE == 2) {                       // cp.async ring, three steps in flight
                if (lane < 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&ring[(step + 3) & 7][lane])), "l"(src + lane) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 3;" ::: "memory");
                __syncwarp();
                acc += ring[step & 7][lane & 15].x;
            } else if (MODE == 3) {                       // one bulk copy (TMA) per step, three in flight
                if (lane == 0) {
                    u32 mb = smem_u32(&mbar[(step + 3) & 7]);
                    asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], 256;" ::"r"(mb) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];"
                                 ::"r"(smem_u32(&ring[(step + 3) & 7][0])), "l"(src), "r"(mb) : "memory");
                }
                if (step >= 3) {
                    u32 mb = smem_u32(&mbar[step & 7]), par = ((step - 3) >> 3) & 1, ok = 0;
                    while (!ok)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(mb), "r"(par) : "memory");
                    acc += ring[step & 7][lane & 15].x;
                }
            } else {
                __nanosleep(100);
            }
            step++; n++;
        }
        if (lane == 0) { sink[1] = acc; out[1] = n; }
    }
}

int main() {
    const int iters = 300000;
    uint4 *tab, *area; long long *out; u32 *sink;
    const u64 big = 4ull << 30;
    cudaMalloc(&tab, 4 << 20); cudaMemset(tab, 0, 4 << 20);
    cudaMalloc(&area, big); cudaMemset(area, 1, big);
    cudaMalloc(&out, 16); cudaMalloc(&sink, 8);
    const char *names[] = {"idle", "LDG.cg 256 B per step", "cp.async.cg 16 x 16 B, 3 in flight", "cp.async.bulk 256 B, 3 in flight"};
    for (u32 kmask : {63u, 2047u})
    for (int areasel = 0; areasel < 2; areasel++) {
        u64 bytes = areasel ? big : (2ull << 20);
        u64 mask16 = bytes / 16 - 1;
        for (int mode = 0; mode < 4; mode++) {
            long long h[2] = {0, 0};
            cudaMemset(out, 0, 16);
            switch (mode) {
                case 0: probe<0><<<1, 64>>>(tab, kmask, area, mask16, iters, out, sink); break;
                case 1: probe<1><<<1, 64>>>(tab, kmask, area, mask16, iters, out, sink); break;
                case 2: probe<2><<<1, 64>>>(tab, kmask, area, mask16, iters, out, sink); break;
                case 3: probe<3><<<1, 64>>>(tab, kmask, area, mask16, iters, out, sink); break;
            }
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("chain over %4u records; other warp: %-36s over %4llu MB : chain %6.1f cycles / step   (other warp: %.0f cycles / fetch)  %s\n", kmask + 1, names[mode], bytes >> 20,
                   (double)h[0] / iters, h[1] ? (double)h[0] / h[1] : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
