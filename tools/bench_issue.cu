// Microbenchmark: what does SERIAL code cost when ONE warp of a CTA runs it (the other warps parked at a named barrier)?
// This is the shape of the sampler's finishing section and of every single-warp epilogue in the persistent frame kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/bench_issue tools/bench_issue.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void park(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }

// variant 0: dependent FADD chain from registers (64 adds)
// variant 1: serial sum of 64 floats read from shared memory 8 at a time (2 LDS.128 per 8 adds), loop not unrolled
// variant 2: rank loop as in the sampler: per 8 entries 2 LDS.128, 16 compares with a bounds branch per entry
// variant 3: the same rank loop without the per-entry branch (padded arrays)
// variant 4: serial sum with the values broadcast by shuffles from lane registers (no shared memory, fully unrolled)
// variant 5: rank by shuffles (64 entries, fully unrolled, no branches)
template <int V>
__global__ void __launch_bounds__(384, 1) k(float* out, unsigned long long* cyc, int ns, int reps) {
    __shared__ __align__(16) float a[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) a[tid] = (tid < ns) ? 1.0f / (float)(tid + 3) : 0.f;
    __syncthreads();
    if (warp != 0 && warp < 8) { park(256); return; }
    if (warp >= 8) return;
    const float p0 = a[lane], p1 = a[lane + 32];
    float acc = 0.f; int r0 = 0, r1 = 0;
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < reps; ++rep) {
        if (V == 0) {
            float s = acc;
#pragma unroll
            for (int i = 0; i < 64; ++i) s += p0;
            acc = s;
        } else if (V == 1) {
            float s = 0.f;
#pragma unroll 1
            for (int i = 0; i < ns; i += 8) {
                const float4 u = *reinterpret_cast<const float4*>(a + i), w = *reinterpret_cast<const float4*>(a + i + 4);
                s += u.x; s += u.y; s += u.z; s += u.w; s += w.x; s += w.y; s += w.z; s += w.w;
            }
            acc += s;
        } else if (V == 2 || V == 3) {
#pragma unroll 1
            for (int j = 0; j < ns; j += 8) {
                const float4 u = *reinterpret_cast<const float4*>(a + j), w = *reinterpret_cast<const float4*>(a + j + 4);
                const float pv[8] = {u.x, u.y, u.z, u.w, w.x, w.y, w.z, w.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int jj = j + q;
                    if (V == 3 || jj < ns) {
                        r0 += (pv[q] > p0 || (pv[q] == p0 && jj < lane)) ? 1 : 0;
                        r1 += (pv[q] > p1 || (pv[q] == p1 && jj < lane + 32)) ? 1 : 0;
                    }
                }
            }
        } else if (V == 4) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) s += __shfl_sync(0xffffffffu, p0, i);
#pragma unroll
            for (int i = 0; i < 32; ++i) s += __shfl_sync(0xffffffffu, p1, i);
            acc += s;
        } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const float v = __shfl_sync(0xffffffffu, i < 32 ? p0 : p1, i & 31);
                r0 += (v > p0 || (v == p0 && i < lane)) ? 1 : 0;
                r1 += (v > p1 || (v == p1 && i < lane + 32)) ? 1 : 0;
            }
        }
    }
    const unsigned long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    out[lane] = acc + (float)(r0 + r1);
    __syncwarp();
    // release the parked warps
    for (int i = 0; i < 7; ++i) { }
    park(256);
}

template <int V>
static void run(const char* name, int ns) {
    float* out; unsigned long long* cyc;
    CK(cudaMalloc(&out, 256)); CK(cudaMalloc(&cyc, 8));
    const int reps = 200;
    k<V><<<1, 384>>>(out, cyc, ns, reps);
    CK(cudaDeviceSynchronize());
    k<V><<<148, 384>>>(out, cyc, ns, reps);
    CK(cudaDeviceSynchronize());
    unsigned long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("%-64s ns=%2d  %8.1f cycles per pass\n", name, ns, (double)h / reps);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("64 dependent FADDs (registers)", 50);
    run<1>("serial sum, 8 per iteration from shared memory", 50);
    run<2>("rank loop, LDS + per-entry bounds branch (sampler today)", 50);
    run<3>("rank loop, LDS, no bounds branch", 50);
    run<4>("serial sum by 64 shuffles", 50);
    run<5>("rank by 64 shuffles, no branches", 50);
    return 0;
}
