// Single-warp latency microbenchmarks on B200 (dependent chains, one warp resident per SM sub-partition):
//   mma.sync.m16n8k16 bf16 (dependent accumulator chain vs independent chains), ld.shared.v4, shfl, FFMA, IMAD chain.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench_latency.bin tools/bench_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int MODE>
__global__ void k(float* out, unsigned long long* cyc, int reps, uint32_t seed) {
    __shared__ __align__(16) uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (i * 16) & 16383;   // pointer-chase table: sm[i] -> byte offset
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    float c[8][4] = {};
    uint32_t a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, b0 = seed + 4, b1 = seed + 5;
    float f = (float)seed; int x = (int)seed; uint32_t p = lane * 16;
    const unsigned long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) { mma(c[0], a0, a1, a2, a3, b0, b1); }                                   // dependent chain
        else if (MODE == 1) { for (int i = 0; i < 4; ++i) mma(c[i], a0, a1, a2, a3, b0, b1); }  // 4 independent
        else if (MODE == 2) { for (int i = 0; i < 8; ++i) mma(c[i], a0, a1, a2, a3, b0, b1); }  // 8 independent
        else if (MODE == 3) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(sm) + (p & 16383))); p = v.x + lane * 16; }
        else if (MODE == 4) { f = __shfl_xor_sync(0xffffffffu, f, 1) + 1.0f; }
        else if (MODE == 5) { f = fmaf(f, 1.0001f, 0.5f); }
        else if (MODE == 6) { x = x * 3 + 1; }
        else if (MODE == 7) { asm volatile("bar.sync 1, 32;"); }
    }
    const unsigned long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    float s = f + (float)x + (float)p;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[lane] = s;
}
template <int MODE> void run(const char* name, int per) {
    float* out; unsigned long long* cyc;
    CK(cudaMalloc(&out, 256)); CK(cudaMalloc(&cyc, 8));
    const int reps = 2000;
    k<MODE><<<1, 64>>>(out, cyc, reps, 0); CK(cudaDeviceSynchronize());
    k<MODE><<<1, 64>>>(out, cyc, reps, 0); CK(cudaDeviceSynchronize());
    unsigned long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("%-52s %7.1f cycles per iteration (%d op%s)\n", name, (double)h / reps, per, per > 1 ? "s" : "");
}
int main() {
    run<0>("mma.sync m16n8k16 bf16, dependent chain", 1);
    run<1>("mma.sync m16n8k16 bf16, 4 independent", 4);
    run<2>("mma.sync m16n8k16 bf16, 8 independent", 8);
    run<3>("ld.shared.v4 pointer chase", 1);
    run<4>("shfl.xor + fadd chain", 1);
    run<5>("ffma chain", 1);
    run<6>("imad chain", 1);
    run<7>("bar.sync (1 warp)", 1);
    return 0;
}
