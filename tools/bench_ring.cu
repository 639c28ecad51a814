// Microbenchmark: how fast can one producer lane per SM stream HBM -> shared memory through an mbarrier ring of
// 1-D TMA bulk copies (the weight path of the persistent frame kernel)?  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// consumers: 8 warps; each stage: wait full, (optionally touch the data), arrive empty
template <int TOUCH>
__global__ void __launch_bounds__(288, 1) ring_kernel(const unsigned char* w, size_t bytes_per_cta, int stage_bytes, int nstages, unsigned long long* out, float* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[16], empty[16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < nstages; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[i])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[i])), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n = (int)(bytes_per_cta / stage_bytes);
    const unsigned char* src = w + (size_t)blockIdx.x * bytes_per_cta;
    const unsigned long long t0 = clock64();
    if (warp == 8) {
        if (lane == 0) {
            for (int s = 0; s < n; ++s) {
                const int slot = s % nstages; const unsigned par = ((s / nstages) & 1) ^ 1;
                while (!try_wait(&empty[slot], par)) {}
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[slot])), "r"(stage_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem + (size_t)slot * stage_bytes)), "l"(src + (size_t)s * stage_bytes), "r"(stage_bytes), "r"(smem_u32(&full[slot])) : "memory");
            }
        }
    } else {
        float acc = 0.f;
        for (int s = 0; s < n; ++s) {
            const int slot = s % nstages; const unsigned par = (s / nstages) & 1;
            while (!try_wait(&full[slot], par)) {}
            if (TOUCH) {
                const uint2* p = reinterpret_cast<const uint2*>(smem + (size_t)slot * stage_bytes);
                for (int i = tid; i < stage_bytes / 8; i += 256) { const uint2 v = p[i]; acc += __uint_as_float(v.x << 16) + __uint_as_float(v.y & 0xffff0000u); }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
        }
        if (acc == 1.2345f) sink[0] = acc;
    }
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = clock64() - t0;
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double mhz = khz / 1000.0;
    const size_t per_cta = 64ull << 20;   // 64 MB per CTA
    unsigned char* w; unsigned long long* out; float* sink;
    for (int ncta : {148, 128}) {
        CK(cudaMalloc(&w, per_cta * ncta)); CK(cudaMemset(w, 0, per_cta * ncta));
        CK(cudaMalloc(&out, ncta * 8)); CK(cudaMalloc(&sink, 4));
        for (int touch = 0; touch < 2; ++touch)
            for (int sb : {8192, 16384, 32768})
                for (int ns : {2, 4, 8}) {
                    const size_t smem = (size_t)sb * ns;
                    if (smem > 200 * 1024) continue;
                    auto k = touch ? ring_kernel<1> : ring_kernel<0>;
                    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    k<<<ncta, 288, smem>>>(w, per_cta, sb, ns, out, sink);
                    CK(cudaDeviceSynchronize());
                    unsigned long long h[148], mx = 0;
                    CK(cudaMemcpy(h, out, ncta * 8, cudaMemcpyDeviceToHost));
                    for (int i = 0; i < ncta; ++i) mx = h[i] > mx ? h[i] : mx;
                    const double sec = mx / (mhz * 1e6);
                    printf("ctas %3d touch %d stage %5d B x %d stages: %7.1f GB/s  (%.2f us per stage)\n", ncta, touch, sb, ns, per_cta * ncta / sec / 1e9, sec * 1e6 / (per_cta / sb));
                }
        cudaFree(w); cudaFree(out); cudaFree(sink);
    }
    return 0;
}
