// Microbenchmark of the inter-CTA exchange primitives the persistent frame kernel depends on (B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/bench_exchange tools/bench_exchange.cu
// Every test: 148 CTAs x 256 threads (one per SM, cooperative), ITER hops; each hop = every CTA publishes its slice
// of a vector of W (value, sequence) words, a grid hand-over, then every CTA reads the WHOLE vector.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

enum { LD_VOLATILE = 0, LD_RELAXED = 1, LD_CG = 2, LD_BULK = 3 };
enum { SYNC_POLL_DATA = 0, SYNC_COUNTER = 1, SYNC_COUNTER_FENCE = 2 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int LD>
__device__ __forceinline__ uint4 load16(const uint2* p) {
    uint4 r;
    if (LD == LD_VOLATILE) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    else if (LD == LD_RELAXED) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    else asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

struct Params {
    uint2* buf[2];          // double-buffered vector of W words
    unsigned* ctr;          // grid counter
    int W, iters;
    unsigned long long* out;   // per-CTA total cycles; [ncta] = retries
    float* sink;
};

template <int LD, int SYNC>
__global__ void __launch_bounds__(256, 1) hop_kernel(Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x;
    const int per = (p.W + ncta - 1) / ncta;
    const int w0 = min(p.W, cta * per), w1 = min(p.W, w0 + per);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    unsigned long long retries = 0;
    unsigned nwait = 0;
    const unsigned long long t0 = clock64();
    for (int it = 1; it <= p.iters; ++it) {
        uint2* b = p.buf[it & 1];
        // publish
        for (int w = w0 + tid; w < w1; w += 256)
            asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(b + w), "r"(__float_as_uint((float)(w + it))), "r"((unsigned)it) : "memory");
        if (SYNC != SYNC_POLL_DATA) {
            __syncthreads();
            if (tid == 0) {
                if (SYNC == SYNC_COUNTER_FENCE) __threadfence();
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p.ctr) : "memory");
                const unsigned target = (unsigned)it * (unsigned)ncta;
                unsigned v;
                do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ctr) : "memory"); } while ((int)(v - target) < 0);
            }
            __syncthreads();
        }
        // read everything
        if (LD == LD_BULK) {
            const uint32_t bytes = (uint32_t)p.W * 8u;
            bool ok = false;
            while (!ok) {
                if (tid == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(smem)), "l"(b), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
                }
                uint32_t done = 0;
                while (!done) {
                    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(nwait & 1u) : "memory");
                }
                ++nwait;
                int bad = 0;
                for (int w = tid * 2; w < p.W; w += 512) {
                    const uint4 v = *reinterpret_cast<const uint4*>(smem + (size_t)w * 8);
                    if (v.y != (unsigned)it || v.w != (unsigned)it) bad = 1; else acc += __uint_as_float(v.x) + __uint_as_float(v.z);
                }
                ok = !__syncthreads_or(bad);
                if (!ok) ++retries;
            }
        } else {
            for (int w = tid * 2; w < p.W; w += 512) {
                uint4 v = load16<LD>(b + w);
                while (v.y != (unsigned)it || v.w != (unsigned)it) { ++retries; v = load16<LD>(b + w); }
                acc += __uint_as_float(v.x) + __uint_as_float(v.z);
            }
        }
        if (SYNC == SYNC_POLL_DATA) __syncthreads();
    }
    const unsigned long long t1 = clock64();
    if (tid == 0) p.out[cta] = t1 - t0;
    atomicAdd(&p.out[ncta], retries);
    if (acc == 12345.678f) p.sink[0] = acc;
}

// pure broadcast read (no writes): all CTAs read the same W words ITER times
template <int LD>
__global__ void __launch_bounds__(256, 1) read_kernel(Params p) {
    const int tid = threadIdx.x;
    float acc = 0.f;
    const unsigned long long t0 = clock64();
    for (int it = 1; it <= p.iters; ++it) {
        for (int w = tid * 2; w < p.W; w += 512) {
            const uint4 v = load16<LD>(p.buf[it & 1] + w);
            acc += __uint_as_float(v.x) + __uint_as_float(v.z);
        }
        __syncthreads();
    }
    const unsigned long long t1 = clock64();
    if (tid == 0) p.out[blockIdx.x] = t1 - t0;
    if (acc == 12345.678f) p.sink[0] = acc;
}

// grid barrier alone
__global__ void __launch_bounds__(256, 1) barrier_kernel(Params p) {
    const int tid = threadIdx.x, ncta = gridDim.x;
    const unsigned long long t0 = clock64();
    for (int it = 1; it <= p.iters; ++it) {
        __syncthreads();
        if (tid == 0) {
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p.ctr) : "memory");
            const unsigned target = (unsigned)it * (unsigned)ncta;
            unsigned v;
            do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ctr) : "memory"); } while ((int)(v - target) < 0);
        }
        __syncthreads();
    }
    const unsigned long long t1 = clock64();
    if (tid == 0) p.out[blockIdx.x] = t1 - t0;
}

template <typename K>
static void run(const char* name, K kern, Params p, int ncta, size_t smem, double mhz) {
    CK(cudaMemset(p.ctr, 0, 256));
    CK(cudaMemset(p.buf[0], 0, (size_t)p.W * 8));
    CK(cudaMemset(p.buf[1], 0, (size_t)p.W * 8));
    CK(cudaMemset(p.out, 0, (ncta + 1) * 8));
    void* args[] = {(void*)&p};
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaLaunchCooperativeKernel((const void*)kern, dim3(ncta), dim3(256), args, smem, 0));
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(ncta + 1);
    CK(cudaMemcpy(h.data(), p.out, (ncta + 1) * 8, cudaMemcpyDeviceToHost));
    unsigned long long mx = 0;
    for (int i = 0; i < ncta; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-44s W=%5d (%3d KB)  %8.0f cyc/hop  %6.2f us/hop  retries/hop %.1f\n", name, p.W, p.W * 8 / 1024, (double)mx / p.iters,
           (double)mx / p.iters / mhz, (double)h[ncta] / p.iters);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int ncta = prop.multiProcessorCount;
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double mhz = khz / 1000.0;
    printf("%s: %d SMs, %.0f MHz\n", prop.name, ncta, mhz);
    Params p{};
    const int WMAX = 16384;
    CK(cudaMalloc(&p.buf[0], WMAX * 8)); CK(cudaMalloc(&p.buf[1], WMAX * 8));
    CK(cudaMalloc(&p.ctr, 256)); CK(cudaMalloc(&p.out, (ncta + 1) * 8)); CK(cudaMalloc(&p.sink, 4));
    p.iters = 2000;
    run("grid barrier only (red.add + 1 poller/CTA)", barrier_kernel, p, ncta, 0, mhz);
    const int Ws[] = {1024, 3072, 8192};
    for (int W : Ws) {
        p.W = W;
        run("broadcast read only, ld.volatile", read_kernel<LD_VOLATILE>, p, ncta, 0, mhz);
        run("broadcast read only, ld.relaxed.gpu", read_kernel<LD_RELAXED>, p, ncta, 0, mhz);
        run("broadcast read only, ld.cg", read_kernel<LD_CG>, p, ncta, 0, mhz);
        run("hop: LL poll data, ld.volatile", hop_kernel<LD_VOLATILE, SYNC_POLL_DATA>, p, ncta, 0, mhz);
        run("hop: LL poll data, ld.relaxed.gpu", hop_kernel<LD_RELAXED, SYNC_POLL_DATA>, p, ncta, 0, mhz);
        run("hop: counter + LL, ld.volatile", hop_kernel<LD_VOLATILE, SYNC_COUNTER>, p, ncta, 0, mhz);
        run("hop: counter + LL, ld.relaxed.gpu", hop_kernel<LD_RELAXED, SYNC_COUNTER>, p, ncta, 0, mhz);
        run("hop: counter + LL, ld.cg", hop_kernel<LD_CG, SYNC_COUNTER>, p, ncta, 0, mhz);
        run("hop: fence + counter, ld.cg", hop_kernel<LD_CG, SYNC_COUNTER_FENCE>, p, ncta, 0, mhz);
        run("hop: counter + TMA bulk to smem + validate", hop_kernel<LD_BULK, SYNC_COUNTER>, p, ncta, (size_t)W * 8, mhz);
    }
    return 0;
}
