// tokenizer12hz_decode (src/tts_onnx.cpp:759-776) building blocks, channels-last activations.
//   conv_gemm_kernel : implicit GEMM  y[L][N] = epi( A[L][taps*Cin] . W[N][taps*Cin]^T )
//                      covers Linear, causal dilated conv (A row p = taps shifted input rows),
//                      and transposed conv k=2s/stride s as s phase-GEMMs (N = s*Cout, A row =
//                      [x[p], x[p-1]]), output [L][s*Cout] == channels-last [L*s][Cout].
//   plus the memory-bound pieces: RVQ gather-sum, SnakeBeta, RMSNorm rows, depthwise conv +
//   LayerNorm, SiLU*mul, final conv + clamp.
// v1: fp32 CUDA-core tiles with bf16 weights (exact activations -> parity with the fp32 oracle).
#pragma once
#include "common.cuh"

namespace lqt {

struct ConvGemmParams {
    const float* x;              // [L][Cin]
    const __nv_bfloat16* W;      // [N][taps*Cin]
    const float* bias;           // nullable, index n % bias_mod
    const float* scale;          // nullable, index n % bias_mod (LayerScale / ConvNeXt gamma)
    const float* residual;       // nullable [L][N]
    float* y;                    // [L][N]
    int L, Cin, N, taps, dil;
    int tap_rev;                 // 0: tap j reads x[p-(taps-1-j)*dil] (causal conv); 1: x[p-j*dil] (tconv)
    int shift;                   // added to the source position (0 = causal; (taps/2)*dil = 'same' padding)
    int bias_mod;
    int act;                     // 0 none, 1 SiLU, 3 GELU(erf)
    int hist;                    // streaming decode: rows [-hist, 0) in front of x hold the previous chunk's tail (0 = zero padding)
};

constexpr int CG_BM = 64, CG_BN = 64, CG_BK = 16, CG_THREADS = 256;

__global__ void __launch_bounds__(CG_THREADS)
conv_gemm_kernel(const ConvGemmParams p) {
    __shared__ float As[CG_BK][CG_BM + 4];
    __shared__ float Bs[CG_BK][CG_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * CG_BM, n0 = blockIdx.y * CG_BN;
    const int K = p.taps * p.Cin;
    const bool fast = (p.Cin % 16) == 0;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int a_row = tid >> 2, a_kq = (tid & 3) * 4;
    const int b_n = tid >> 1, b_kq = (tid & 1) * 8;

    float a_reg[4];
    float b_reg[8];

    auto load_slab = [&](int k0) {
        // ---- A: 64 rows x 16 k ---------------------------------------------------------------
        const int pidx = m0 + a_row;
        if (fast) {
            const int k = k0 + a_kq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pidx < p.L && k < K) {
                const int tap = k / p.Cin, c = k - tap * p.Cin;
                const int src = pidx + p.shift - (p.tap_rev ? tap : (p.taps - 1 - tap)) * p.dil;
                if (src >= -p.hist && src < p.L) v = *reinterpret_cast<const float4*>(p.x + (long long)src * p.Cin + c);
            }
            a_reg[0] = v.x; a_reg[1] = v.y; a_reg[2] = v.z; a_reg[3] = v.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = k0 + a_kq + e;
                float v = 0.f;
                if (pidx < p.L && k < K) {
                    const int tap = k / p.Cin, c = k - tap * p.Cin;
                    const int src = pidx + p.shift - (p.tap_rev ? tap : (p.taps - 1 - tap)) * p.dil;
                    if (src >= -p.hist && src < p.L) v = p.x[(long long)src * p.Cin + c];
                }
                a_reg[e] = v;
            }
        }
        // ---- B: 64 n x 16 k (threads 0..127) ----------------------------------------------------
        if (tid < 128) {
            const int n = n0 + b_n, k = k0 + b_kq;
            if (fast) {
                uint4 w = make_uint4(0u, 0u, 0u, 0u);
                if (n < p.N && k < K) w = *reinterpret_cast<const uint4*>(p.W + (size_t)n * K + k);
                b_reg[0] = bf16lo(w.x); b_reg[1] = bf16hi(w.x); b_reg[2] = bf16lo(w.y); b_reg[3] = bf16hi(w.y);
                b_reg[4] = bf16lo(w.z); b_reg[5] = bf16hi(w.z); b_reg[6] = bf16lo(w.w); b_reg[7] = bf16hi(w.w);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    b_reg[e] = (n < p.N && k + e < K) ? __bfloat162float(p.W[(size_t)n * K + k + e]) : 0.f;
            }
        }
    };
    auto store_slab = [&]() {
#pragma unroll
        for (int e = 0; e < 4; ++e) As[a_kq + e][a_row] = a_reg[e];
        if (tid < 128) {
#pragma unroll
            for (int e = 0; e < 8; ++e) Bs[b_kq + e][b_n] = b_reg[e];
        }
    };

    load_slab(0);
    for (int k0 = 0; k0 < K; k0 += CG_BK) {
        store_slab();
        __syncthreads();
        if (k0 + CG_BK < K) load_slab(k0 + CG_BK);
#pragma unroll
        for (int kk = 0; kk < CG_BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.L) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            const int bn = n % p.bias_mod;
            if (p.bias) v += p.bias[bn];
            if (p.act == 1) v = silu_f(v);
            else if (p.act == 3) v = gelu_erf_f(v);
            if (p.scale) v *= p.scale[bn];
            if (p.residual) v += p.residual[(size_t)m * p.N + n];
            p.y[(size_t)m * p.N + n] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core version of conv_gemm_kernel (same parameters, Cin % 16 == 0): mma.sync.m16n8k16, bf16 weights as the B
// operand, the fp32 activations as the A operand split into three bf16 planes (x = hi + mid + lo, remainders exact), so
// every product is exact in fp32 and the result differs from the CUDA-core kernel only by summation order.
// Block tile 128 positions x 64 outputs x 16 k, 8 warps x (16 x 64): 24 MMAs per warp and k-step, operands staged through
// padded shared memory (row stride 24 bf16: conflict-free fragment loads), next slab prefetched into registers.
// ------------------------------------------------------------------------------------------------
constexpr int CM_BM = 128, CM_BN = 64, CM_BK = 16, CM_LD = 24, CM_THREADS = 256, CM_FLUSH = 2;

// two fp32 values -> three packed bf16x2 planes (low half = first value); one packed conversion per plane, exact remainders
__device__ __forceinline__ void cm_split3x2(float x0, float x1, uint32_t& h, uint32_t& m, uint32_t& l) {
    const __nv_bfloat162 bh = __floats2bfloat162_rn(x0, x1);
    h = *reinterpret_cast<const uint32_t*>(&bh);
    const float r0 = x0 - __uint_as_float(h << 16), r1 = x1 - __uint_as_float(h & 0xffff0000u);
    const __nv_bfloat162 bm = __floats2bfloat162_rn(r0, r1);
    m = *reinterpret_cast<const uint32_t*>(&bm);
    const float q0 = r0 - __uint_as_float(m << 16), q1 = r1 - __uint_as_float(m & 0xffff0000u);
    const __nv_bfloat162 bl = __floats2bfloat162_rn(q0, q1);
    l = *reinterpret_cast<const uint32_t*>(&bl);
}

__global__ void __launch_bounds__(CM_THREADS)
conv_gemm_mma_kernel(const ConvGemmParams p) {
    __shared__ __align__(16) unsigned short As[3][CM_BM][CM_LD];
    __shared__ __align__(16) unsigned short Bs[CM_BN][CM_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const int m0 = blockIdx.x * CM_BM, n0 = blockIdx.y * CM_BN;
    const int K = p.taps * p.Cin;

    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }

    const int a_row = tid >> 1, a_k = (tid & 1) * 8;            // 128 rows x 16 k: 8 consecutive k per thread
    const int b_n = tid >> 1, b_k = (tid & 1) * 8;              // 64 n x 16 k (threads 0..127)
    float4 a0, a1;
    uint4 bw;
    auto load_slab = [&](int k0) {
        a0 = make_float4(0.f, 0.f, 0.f, 0.f); a1 = a0;
        const int pidx = m0 + a_row, k = k0 + a_k;
        if (pidx < p.L) {
            const int tap = k / p.Cin, c = k - tap * p.Cin;       // Cin % 16 == 0: the slab lies inside one tap
            const int src = pidx + p.shift - (p.tap_rev ? tap : (p.taps - 1 - tap)) * p.dil;
            if (src >= -p.hist && src < p.L) {
                const float4* q = reinterpret_cast<const float4*>(p.x + (long long)src * p.Cin + c);
                a0 = q[0]; a1 = q[1];
            }
        }
        bw = make_uint4(0u, 0u, 0u, 0u);
        if (tid < 128) {
            const int n = n0 + b_n;
            if (n < p.N) bw = *reinterpret_cast<const uint4*>(p.W + (size_t)n * K + k0 + b_k);
        }
    };
    auto store_slab = [&]() {
        uint32_t h[4], m[4], l[4];
        cm_split3x2(a0.x, a0.y, h[0], m[0], l[0]); cm_split3x2(a0.z, a0.w, h[1], m[1], l[1]);
        cm_split3x2(a1.x, a1.y, h[2], m[2], l[2]); cm_split3x2(a1.z, a1.w, h[3], m[3], l[3]);
        *reinterpret_cast<uint4*>(&As[0][a_row][a_k]) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(&As[1][a_row][a_k]) = make_uint4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<uint4*>(&As[2][a_row][a_k]) = make_uint4(l[0], l[1], l[2], l[3]);
        if (tid < 128) *reinterpret_cast<uint4*>(&Bs[b_n][b_k]) = bw;
    };

    // The tensor core adds into its accumulator with truncation, a systematic error that grows with the number of chained MMAs
    // (K up to 5376 here: 3e-4 relative). The MMAs therefore chain over CM_FLUSH k-steps only; the chunk sums are added to the
    // running result with ordinary round-to-nearest FADDs.
    float tacc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { tacc[j][0] = 0.f; tacc[j][1] = 0.f; tacc[j][2] = 0.f; tacc[j][3] = 0.f; }
    int since = 0;
    load_slab(0);
    for (int k0 = 0; k0 < K; k0 += CM_BK) {
        store_slab();
        __syncthreads();
        if (k0 + CM_BK < K) load_slab(k0 + CM_BK);
        uint32_t af[3][4];
        const int r = warp * 16 + g;
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            af[pl][0] = *reinterpret_cast<const uint32_t*>(&As[pl][r][tg * 2]);
            af[pl][1] = *reinterpret_cast<const uint32_t*>(&As[pl][r + 8][tg * 2]);
            af[pl][2] = *reinterpret_cast<const uint32_t*>(&As[pl][r][tg * 2 + 8]);
            af[pl][3] = *reinterpret_cast<const uint32_t*>(&As[pl][r + 8][tg * 2 + 8]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Bs[j * 8 + g][tg * 2]);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Bs[j * 8 + g][tg * 2 + 8]);
#pragma unroll
            for (int pl = 2; pl >= 0; --pl)                       // smallest plane first
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(tacc[j][0]), "+f"(tacc[j][1]), "+f"(tacc[j][2]), "+f"(tacc[j][3])
                             : "r"(af[pl][0]), "r"(af[pl][1]), "r"(af[pl][2]), "r"(af[pl][3]), "r"(b0), "r"(b1));
        }
        if (++since == CM_FLUSH || k0 + CM_BK >= K) {
            since = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) { acc[j][e] += tacc[j][e]; tacc[j][e] = 0.f; }
        }
        __syncthreads();
    }

#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int m = m0 + warp * 16 + g + hf * 8;
            if (m >= p.L) continue;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = n0 + j * 8 + tg * 2 + e;
                if (n >= p.N) continue;
                float v = acc[j][hf * 2 + e];
                const int bn = n % p.bias_mod;
                if (p.bias) v += p.bias[bn];
                if (p.act == 1) v = silu_f(v);
                else if (p.act == 3) v = gelu_erf_f(v);
                if (p.scale) v *= p.scale[bn];
                if (p.residual) v += p.residual[(size_t)m * p.N + n];
                p.y[(size_t)m * p.N + n] = v;
            }
        }
    }
}

// ---- RVQ dequantisation: gather-SUM per group (codebook order), output [T][Dc] each -------------
__global__ void rvq_gather_kernel(const long long* __restrict__ codes, int T, int n_q,
                                  const __nv_bfloat16* __restrict__ cb_sem,   // [1][size][Dc]
                                  const __nv_bfloat16* __restrict__ cb_aco,   // [n_q-1][size][Dc]
                                  int cb_size, int Dc, float* sem, float* aco) {
    const int t = blockIdx.x;
    if (t >= T) return;
    for (int d = threadIdx.x; d < Dc; d += blockDim.x) {
        const long long c0 = codes[(size_t)t * n_q];
        sem[(size_t)t * Dc + d] = __bfloat162float(cb_sem[(size_t)c0 * Dc + d]);
        float a = 0.f;
        for (int j = 1; j < n_q; ++j) {
            const long long c = codes[(size_t)t * n_q + j];
            a += __bfloat162float(cb_aco[((size_t)(j - 1) * cb_size + c) * Dc + d]);
        }
        aco[(size_t)t * Dc + d] = a;
    }
}

// ---- SnakeBeta: y = x + 1/(exp(beta)+1e-9) * sin(x*exp(alpha))^2, per channel --------------------
__global__ void snake_kernel(const float* __restrict__ x, float* __restrict__ y, long long n4, int C,
                             const float* __restrict__ alpha, const float* __restrict__ beta) {
    // n4 = number of float4 groups; C % 4 == 0
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i * 4) % C);
        float4 v = reinterpret_cast<const float4*>(x)[i];
        const float4 al = *reinterpret_cast<const float4*>(alpha + c);
        const float4 be = *reinterpret_cast<const float4*>(beta + c);
        float s;
        s = sinf(v.x * expf(al.x)); v.x = v.x + (1.0f / (expf(be.x) + 1e-9f)) * (s * s);
        s = sinf(v.y * expf(al.y)); v.y = v.y + (1.0f / (expf(be.y) + 1e-9f)) * (s * s);
        s = sinf(v.z * expf(al.z)); v.z = v.z + (1.0f / (expf(be.z) + 1e-9f)) * (s * s);
        s = sinf(v.w * expf(al.w)); v.w = v.w + (1.0f / (expf(be.w) + 1e-9f)) * (s * s);
        reinterpret_cast<float4*>(y)[i] = v;
    }
}

// ---- RMSNorm over rows [L][C]; one warp per row ---------------------------------------------------
__global__ void rmsnorm_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int L, int C,
                                    const float* __restrict__ w, float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= L) return;
    const float* xr = x + (size_t)row * C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) ss += xr[c] * xr[c];
    ss = warp_sum(ss);
    const float r = 1.0f / sqrtf(ss / (float)C + eps);
    for (int c = lane; c < C; c += 32) y[(size_t)row * C + c] = (xr[c] * r) * w[c];
}

__global__ void silu_mul_kernel(const float* g, const float* u, float* y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        y[i] = silu_f(g[i]) * u[i];
}

// ---- ConvNeXt front: depthwise causal conv k7 (+bias) then LayerNorm over channels ----------------
// one CTA per position; dynamic smem = C floats
// hist: rows [-hist, 0) in front of x hold the previous chunk's tail (streaming decode); 0 = causal zero padding
__global__ void dwconv_ln_kernel(const float* __restrict__ x, float* __restrict__ y, int L, int C,
                                 const float* __restrict__ dw_w /*[7][C]*/, const float* __restrict__ dw_b,
                                 const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, int hist) {
    extern __shared__ float hbuf[];
    __shared__ float red[32];
    __shared__ float stat;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float s = 0.f;
    for (int c = tid; c < C; c += blockDim.x) {
        float a = dw_b[c];
#pragma unroll
        for (int tap = 0; tap < 7; ++tap) {
            const int src = p - (6 - tap);
            if (src >= -hist) a = fmaf(dw_w[tap * C + c], x[(long long)src * C + c], a);
        }
        hbuf[c] = a;
        s += a;
    }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int w = 0; w < nw; ++w) t += red[w]; stat = t / (float)C; }
    __syncthreads();
    const float mean = stat;
    float v = 0.f;
    for (int c = tid; c < C; c += blockDim.x) { const float d = hbuf[c] - mean; v += d * d; }
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int w = 0; w < nw; ++w) t += red[w]; stat = 1.0f / sqrtf(t / (float)C + eps); }
    __syncthreads();
    const float rstd = stat;
    for (int c = tid; c < C; c += blockDim.x)
        y[(size_t)p * C + c] = (hbuf[c] - mean) * rstd * ln_w[c] + ln_b[c];
}

// ---- final causal conv k7 C->1 (+bias) + clamp[-1,1]; input already SnakeBeta-activated ----------
// one warp per output sample
__global__ void conv_out_kernel(const float* __restrict__ x, float* __restrict__ y, long long L, int C,
                                const float* __restrict__ w /*[7][C]*/, const float* __restrict__ b, int hist) {
    const long long pos = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (pos >= L) return;
    float a = 0.f;
    for (int tap = 0; tap < 7; ++tap) {
        const long long src = pos - (6 - tap);
        if (src < -hist) continue;
        for (int c = lane; c < C; c += 32) a = fmaf(w[tap * C + c], x[src * C + c], a);
    }
    a = warp_sum(a);
    if (lane == 0) y[pos] = fminf(1.0f, fmaxf(-1.0f, a + b[0]));
}

// ---- speaker encoder helpers ------------------------------------------------------------------------
// mean / std pooling over frames: x [F][C] -> out [2C] (mean | sqrt(var + 1e-5))
__global__ void stat_pool_kernel(const float* __restrict__ x, int F, int C, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int f = 0; f < F; ++f) s += x[(size_t)f * C + c];
    const float mean = s / (float)F;
    float v = 0.f;
    for (int f = 0; f < F; ++f) { const float d = x[(size_t)f * C + c] - mean; v += d * d; }
    out[c] = mean;
    out[C + c] = sqrtf(v / (float)F + 1e-5f);
}

__global__ void relu_add_kernel(const float* h, const float* res, float* y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        y[i] = (res ? res[i] : 0.f) + fmaxf(h[i], 0.f);
}

// ---- log-mel front end of the clone path on the device (src/io/mel.cpp:132-236; src/tts_onnx.cpp:331-359) -----------------
// One CTA per frame: window, 1024-point radix-2 decimation-in-time FFT in shared memory (f32, the twiddle TABLE is the host's:
// cos/sin of the f32 angle, tabulated per stage exactly as the reference evaluates them), power spectrum, HTK-mel triangles
// summed in bin order, log(e + 1e-10). Output transposed, [frames][num_mels]: the layout speaker_encoder wants (:374-380).
struct MelParams {
    const float* audio; long long n;          // 24 kHz mono
    const float* window;                      // [win]
    const float* tw_re; const float* tw_im;   // twiddles of the log2(n_fft) stages, concatenated (n_fft - 1 entries)
    const int* tri;                           // [num_mels][3]: left, centre, right bin
    float* out;                               // [frames][num_mels]
    int frames, hop, win, n_fft, log2n, num_mels;
};
__global__ void __launch_bounds__(256)
logmel_kernel(const MelParams p) {
    extern __shared__ float mel_smem[];       // re[n_fft] | im[n_fft] | pw[n_fft/2+1]
    float* re = mel_smem; float* im = re + p.n_fft; float* pw = im + p.n_fft;
    const int t = blockIdx.x, tid = threadIdx.x, n = p.n_fft;
    const long long start = (long long)t * p.hop;
    for (int i = tid; i < n; i += blockDim.x) {
        const long long idx = start + i;
        const float v = (i < p.win && idx < p.n) ? p.audio[idx] * p.window[i] : 0.0f;
        const int j = (int)(__brev((unsigned)i) >> (32 - p.log2n));          // bit-reversal permutation
        re[j] = v; im[j] = 0.0f;
    }
    __syncthreads();
    int tw0 = 0;
    for (int size = 2; size <= n; size <<= 1) {
        const int half = size >> 1;
        for (int b = tid; b < (n >> 1); b += blockDim.x) {
            const int k = b & (half - 1), base = (b - k) << 1;
            const float wr = p.tw_re[tw0 + k], wi = p.tw_im[tw0 + k];
            const int e = base + k, o = e + half;
            const float tr = __fsub_rn(__fmul_rn(wr, re[o]), __fmul_rn(wi, im[o]));   // no FMA contraction: the reference's f32 butterflies
            const float ti = __fadd_rn(__fmul_rn(wr, im[o]), __fmul_rn(wi, re[o]));
            const float er = re[e], ei = im[e];
            re[o] = er - tr; im[o] = ei - ti;
            re[e] = er + tr; im[e] = ei + ti;
        }
        tw0 += half;
        __syncthreads();
    }
    const int nb = n / 2 + 1;
    for (int k = tid; k < nb; k += blockDim.x) pw[k] = __fadd_rn(__fmul_rn(re[k], re[k]), __fmul_rn(im[k], im[k]));
    __syncthreads();
    for (int m = tid; m < p.num_mels; m += blockDim.x) {
        const int l = p.tri[m * 3], c = p.tri[m * 3 + 1], r = p.tri[m * 3 + 2];
        float e = 0.0f;
        for (int k = l; k < c && k < nb; ++k) e = __fadd_rn(e, __fmul_rn((float)(k - l) / (float)(c - l), pw[k]));
        for (int k = c; k < r && k < nb; ++k) e = __fadd_rn(e, __fmul_rn((float)(r - k) / (float)(r - c), pw[k]));
        p.out[(size_t)t * p.num_mels + m] = logf(e + 1e-10f);
    }
}

}  // namespace lqt
