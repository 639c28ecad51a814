// Vocoder decoder convolutions as TMA-fed tcgen05 implicit GEMMs (north_star (4): "causal conv and transposed-conv
// upsampling to 24 kHz as implicit-GEMM tensor-core kernels"; replaces conv_gemm_mma_kernel + snake_kernel for the decoder
// blocks of tokenizer12hz_decode, src/tts_onnx.cpp:759-776).
//
//   y[l][n] = epi( sum_tap sum_plane sum_c  Xp[l + off(tap)][plane][c] * W[n][tap][c] )
//
// * Time is the M side: 128 consecutive positions per CTA = 128 TMEM lanes; output channels are the N side (<= 256 per tile).
// * There is no im2col: tap `t` of a causal (dilated) conv is the SAME activation matrix shifted by off(t) rows, so the
//   producer just issues the 2-D TMA box load at row l0 + off(t); rows before the start of the sequence are out of bounds and
//   arrive as zeros -- exactly the causal left padding. A transposed conv (kernel 2s, stride s) is the GEMM with two taps
//   (x[p], x[p-1]) and N = s * Cout: [L][s*Cout] is channels-last [L*s][Cout].
// * Activations live in HBM as bf16 PLANES (hi | mid | lo per row, layout [L][planes][Cp], Cp = channels padded to 64):
//   x = hi + mid (+ lo) with exact remainders, every plane is one more K pass into the same fp32 accumulator, products are
//   exact. The planes are written by the PRODUCING kernel's epilogue (fp32 -> planes once per element), with the SnakeBeta
//   activation of the consuming layer already applied, so neither the split nor the activation is a separate pass.
// * Epilogue (one thread = one position, tcgen05.ld 32x32b): + bias, + residual, fp32 store (the residual stream) and/or
//   SnakeBeta + plane store for the next conv.
// Warp roles / pipeline as in tc_gemm.cuh.
#pragma once
#include "tc_gemm.cuh"

namespace lqt {

struct TcConvParams {
    int L;                     // positions (rows of x and y)
    int N;                     // output channels (for a transposed conv: s * Cout)
    int BN;                    // channel tile (multiple of 16, <= 256)
    int taps, dil, tap_rev;    // tap t reads x[l - (taps-1-t)*dil] (causal conv) or x[l - t*dil] (tap_rev: transposed conv)
    int planes, Cp;            // input planes, padded input channels (multiple of 64)
    int stages;
    const float* bias; int bias_mod;        // nullable; index n % bias_mod
    int act;                                // 0 none, 1 SiLU, 3 GELU(erf), after the bias
    const float* scale;                     // nullable, index n % bias_mod (LayerScale / ConvNeXt gamma), before the residual
    const float* residual;                  // nullable fp32 [L][N]
    float* y;                               // nullable fp32 [L][N]
    int y_snake;                            // 1: the fp32 output is stored AFTER SnakeBeta (input of conv_out_kernel)
    // output planes for the next conv: row = l * up + n / cout, channel = n % cout, layout [rows][oplanes][oCp]
    __nv_bfloat16* xo; int oplanes, oCp, cout, up;
    const float* sn_ea; const float* sn_ib; // nullable SnakeBeta of the consumer: exp(alpha)[cout], 1/(exp(beta)+1e-9)[cout]
};

template <int TMEM_COLS>
__global__ void __launch_bounds__(TG_THREADS, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcConvParams p) {
    extern __shared__ __align__(1024) unsigned char tg_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tg_smem_raw) + 1023) & ~(uintptr_t)1023);
    const size_t stage_bytes = (size_t)TG_BM * 128 + (size_t)p.BN * 128;
    TgShared* sh = reinterpret_cast<TgShared*>(smem + (size_t)p.stages * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l0 = blockIdx.x * TG_BM, n0 = blockIdx.y * p.BN;
    const int kbc = p.Cp / TG_BK;                                  // k-blocks per (tap, plane)
    const int niter = p.taps * p.planes * kbc;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) { tg_mbar_init(&sh->full[i], 1); tg_mbar_init(&sh->empty[i], 1); }
        tg_mbar_init(&sh->tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tg_smem_u32(&sh->tmem_base)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sh->tmem_base;

    if (warp == 0) {
        if (tg_elect_one()) {
            const uint32_t bytes = (uint32_t)stage_bytes;
            int i = 0;
            for (int t = 0; t < p.taps; ++t) {
                const int off = -(p.tap_rev ? t : (p.taps - 1 - t)) * p.dil;
                for (int pl = 0; pl < p.planes; ++pl)
                    for (int kb = 0; kb < kbc; ++kb, ++i) {
                        const int s = i % p.stages;
                        if (i >= p.stages) tg_mbar_wait(&sh->empty[s], (uint32_t)((i / p.stages) - 1) & 1u);
                        unsigned char* a = smem + (size_t)s * stage_bytes;
                        tg_mbar_expect_tx(&sh->full[s], bytes);
                        tg_tma_2d(a, &map_x, pl * p.Cp + kb * TG_BK, l0 + off, &sh->full[s]);        // rows < 0: zero fill = causal padding
                        tg_tma_2d(a + TG_BM * 128, &map_w, t * p.Cp + kb * TG_BK, n0, &sh->full[s]);
                    }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TG_BM >> 4) << 24);
        for (int i = 0; i < niter; ++i) {
            const int s = i % p.stages;
            tg_mbar_wait(&sh->full[s], (uint32_t)(i / p.stages) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tg_elect_one()) {
                const uint32_t a = tg_smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t ad = tg_desc_sw128(a), bd = tg_desc_sw128(a + TG_BM * 128);
#pragma unroll
                for (int k = 0; k < TG_BK / 16; ++k)
                    tg_umma(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (i | k) ? 1u : 0u);
                tg_commit(&sh->empty[s]);
                if (i == niter - 1) tg_commit(&sh->tmem_full);
            }
            __syncwarp();
        }
    } else {
        tg_mbar_wait(&sh->tmem_full, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        const int l = l0 + q * 32 + lane;
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16);
        const bool live = l < p.L;
        const int nmax = min(p.BN, p.N - n0);
        for (int c = 0; c < nmax; c += 16) {
            float v[16];
            tg_tmem_ld16(tbase + (uint32_t)c, v);
            if (!live) continue;
            const int n = n0 + c;
            if (p.bias) {
                const float* b = p.bias + (n % p.bias_mod);
#pragma unroll
                for (int j = 0; j < 16; j += 4) { const float4 w = __ldg(reinterpret_cast<const float4*>(b + j)); v[j] += w.x; v[j + 1] += w.y; v[j + 2] += w.z; v[j + 3] += w.w; }
            }
            if (p.act == 3) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = gelu_erf_f(v[j]);
            } else if (p.act == 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
            }
            if (p.scale) {
                const float* sc = p.scale + (n % p.bias_mod);
#pragma unroll
                for (int j = 0; j < 16; j += 4) { const float4 w = __ldg(reinterpret_cast<const float4*>(sc + j)); v[j] *= w.x; v[j + 1] *= w.y; v[j + 2] *= w.z; v[j + 3] *= w.w; }
            }
            if (p.residual) {
                const float* r = p.residual + (size_t)l * p.N + n;
#pragma unroll
                for (int j = 0; j < 16; j += 4) { const float4 w = *reinterpret_cast<const float4*>(r + j); v[j] += w.x; v[j + 1] += w.y; v[j + 2] += w.z; v[j + 3] += w.w; }
            }
            if (p.y && !p.y_snake) {
                float* y = p.y + (size_t)l * p.N + n;
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(y + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            const int ph = n / p.cout, ch = n - ph * p.cout;         // a 16-channel chunk never straddles a phase (cout % 16 == 0)
            if (p.sn_ea) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float s = sinf(v[j] * __ldg(p.sn_ea + ch + j));
                    v[j] = v[j] + __ldg(p.sn_ib + ch + j) * (s * s);
                }
            }
            if (p.y && p.y_snake) {
                float* y = p.y + (size_t)l * p.N + n;
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(y + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            if (p.xo) {
                __nv_bfloat16* xo = p.xo + ((size_t)((size_t)l * p.up + ph) * p.oplanes) * p.oCp + ch;
                for (int pl = 0; pl < p.oplanes; ++pl) {
                    uint32_t u[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        u[j] = *reinterpret_cast<const uint32_t*>(&h2);
                        v[2 * j] -= bf16lo(u[j]); v[2 * j + 1] -= bf16hi(u[j]);
                    }
                    *reinterpret_cast<uint4*>(xo + (size_t)pl * p.oCp) = make_uint4(u[0], u[1], u[2], u[3]);
                    *reinterpret_cast<uint4*>(xo + (size_t)pl * p.oCp + 8) = make_uint4(u[4], u[5], u[6], u[7]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
    }
}

// fp32 [L][C] -> planes [L][planes][Cp] (optionally through SnakeBeta): the hand-over from the fp32 stages into the tcgen05 decoder
__global__ void voc_planes_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xo, long long L, int C, int planes, int Cp,
                                  const float* __restrict__ sn_ea, const float* __restrict__ sn_ib) {
    const long long n4 = L * (C / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long l = i / (C / 4);
        const int c = (int)(i - l * (C / 4)) * 4;
        const float4 f = *reinterpret_cast<const float4*>(x + l * C + c);
        float v[4] = {f.x, f.y, f.z, f.w};
        if (sn_ea) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float s = sinf(v[j] * sn_ea[c + j]); v[j] = v[j] + sn_ib[c + j] * (s * s); }
        }
        for (int pl = 0; pl < planes; ++pl) {
            const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
            uint2 u;
            u.x = *reinterpret_cast<const uint32_t*>(&a); u.y = *reinterpret_cast<const uint32_t*>(&b);
            *reinterpret_cast<uint2*>(xo + ((size_t)l * planes + pl) * Cp + c) = u;
            v[0] -= bf16lo(u.x); v[1] -= bf16hi(u.x); v[2] -= bf16lo(u.y); v[3] -= bf16hi(u.y);
        }
    }
}

// W [N][taps*Cin] bf16 -> [N][taps][Cp] zero padded (once per model)
__global__ void voc_pad_weight_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ o, int N, int taps, int Cin, int Cp) {
    const long long total = (long long)N * taps * Cp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cp);
        const long long nt = i / Cp;
        o[i] = c < Cin ? w[nt * Cin + c] : __float2bfloat16_rn(0.f);
    }
}
__global__ void voc_snake_consts_kernel(const float* alpha, const float* beta, float* ea, float* ib, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) { ea[c] = expf(alpha[c]); ib[c] = 1.0f / (expf(beta[c]) + 1e-9f); }
}

}  // namespace lqt
