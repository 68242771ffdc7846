// Micro-probe: does a global store keep (update) or drop the L1 line a later load of the same address would hit?
// One warp walks K 512-byte records (lane l owns 16 bytes) in a data-dependent order: load, modify, store, next.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/l1_probe tools/probes/l1_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32;
typedef unsigned long long u64;

template <int MODE>
__global__ void probe(uint4 *tab, int K, int iters, long long *out, u32 *sink) {
    int lane = threadIdx.x & 31;
    u32 ctx = 0, acc = 0;
    long long t0 = 0;
    for (int it = -K * 4; it < iters; it++) {                 // warm the lines first
        if (it == 0) t0 = clock64();
        uint4 *p = tab + (size_t)ctx * 32 + lane;
        uint4 M;
        if (MODE == 5) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(M.x), "=r"(M.y), "=r"(M.z), "=r"(M.w) : "l"(p) : "memory");
        else asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(M.x), "=r"(M.y), "=r"(M.z), "=r"(M.w) : "l"(p) : "memory");
        acc += M.y;
        M.x += 1;
        if (MODE == 1) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(M.x), "r"(M.y), "r"(M.z), "r"(M.w) : "memory");
        if (MODE == 2) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(M.x), "r"(M.y), "r"(M.z), "r"(M.w) : "memory");
        if (MODE == 3) asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(M.x), "r"(M.y), "r"(M.z), "r"(M.w) : "memory");
        if (MODE == 4 || MODE == 5) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + (size_t)K * 32), "r"(M.x), "r"(M.y), "r"(M.z), "r"(M.w) : "memory");   // another line
        if (MODE == 6) { if (lane < 16) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(M.x), "r"(M.y), "r"(M.z), "r"(M.w) : "memory"); }
        u32 nx = __shfl_sync(0xffffffffu, M.w, 0);            // the next record depends on the loaded data
        ctx = (ctx * 5 + 1 + (nx & 0)) % (u32)K;
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; sink[0] = acc; }
}

int main() {
    const int iters = 200000;
    uint4 *tab; long long *out; u32 *sink;
    cudaMalloc(&tab, 64 << 20); cudaMemset(tab, 0, 64 << 20);
    cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
    const char *names[] = {"load only", "load + st (same line)", "load + st.wt", "load + st.cg", "load + st (other line)", "ld.cg + st other (L2 latency)", "load + st lanes<16"};
    for (int K : {16, 64, 256, 2048}) {
        for (int mode = 0; mode < 7; mode++) {
            long long h = 0;
            switch (mode) {
                case 0: probe<0><<<1, 32>>>(tab, K, iters, out, sink); break;
                case 1: probe<1><<<1, 32>>>(tab, K, iters, out, sink); break;
                case 2: probe<2><<<1, 32>>>(tab, K, iters, out, sink); break;
                case 3: probe<3><<<1, 32>>>(tab, K, iters, out, sink); break;
                case 4: probe<4><<<1, 32>>>(tab, K, iters, out, sink); break;
                case 5: probe<5><<<1, 32>>>(tab, K, iters, out, sink); break;
                case 6: probe<6><<<1, 32>>>(tab, K, iters, out, sink); break;
            }
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("K=%4d records (%5d KB)  %-32s %7.1f cycles / step  %s\n", K, K / 2, names[mode], (double)h / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
