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
    int x_row0;                // row of the tensor map that is position 0 (streaming decode: the rows before it are the previous chunk's tail)
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

struct TcShared {
    uint64_t full[TG_MAX_STAGES], empty[TG_MAX_STAGES];
    uint64_t tmem_full[2], tmem_empty[2];      // two accumulator sets: the epilogue of tile i overlaps the MMAs of tile i + 1
    uint32_t tmem_base;
};

// PERSISTENT: grid = min(tiles, co-resident CTAs); CTA c takes tiles c, c + grid, ... (tile = position block x channel tile).
// TMEM_COLS = 2 * (power of two >= BN): the two accumulator sets.
constexpr int TC_EPI_GROUPS = 4;                                  // epilogue warps = 4 TMEM lane quarters x 4 column groups
constexpr int TC_THREADS = 64 + 128 * TC_EPI_GROUPS;              // + producer warp + MMA warp
constexpr int TC_STG_PITCH = 36;                                  // transpose tile row pitch (floats)

template <int TMEM_COLS>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcConvParams p) {
    extern __shared__ __align__(1024) unsigned char tg_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tg_smem_raw) + 1023) & ~(uintptr_t)1023);
    const size_t stage_bytes = (size_t)TG_BM * 128 + (size_t)p.BN * 128;
    TcShared* sh = reinterpret_cast<TcShared*>(smem + (size_t)p.stages * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kbc = p.Cp / TG_BK;                                  // k-blocks per (tap, plane)
    const int niter = p.taps * p.planes * kbc;
    const int n_ntiles = p.N / p.BN;
    const int n_tiles = ((p.L + TG_BM - 1) / TG_BM) * n_ntiles;
    constexpr uint32_t ACC_COLS = TMEM_COLS / 2;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) { tg_mbar_init(&sh->full[i], 1); tg_mbar_init(&sh->empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tg_mbar_init(&sh->tmem_full[i], 1); tg_mbar_init(&sh->tmem_empty[i], 4 * TC_EPI_GROUPS); }   // every epilogue warp releases a set
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
            int i = 0;                                               // ring position, runs on across tiles
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int l0 = (tile / n_ntiles) * TG_BM, n0 = (tile % n_ntiles) * p.BN;
                for (int t = 0; t < p.taps; ++t) {
                    const int off = -(p.tap_rev ? t : (p.taps - 1 - t)) * p.dil;
                    for (int pl = 0; pl < p.planes; ++pl)
                        for (int kb = 0; kb < kbc; ++kb, ++i) {
                            const int s = i % p.stages;
                            if (i >= p.stages) tg_mbar_wait(&sh->empty[s], (uint32_t)((i / p.stages) - 1) & 1u);
                            unsigned char* a = smem + (size_t)s * stage_bytes;
                            tg_mbar_expect_tx(&sh->full[s], bytes);
                            tg_tma_2d(a, &map_x, pl * p.Cp + kb * TG_BK, p.x_row0 + l0 + off, &sh->full[s]);    // rows < 0: zero fill = causal padding
                            tg_tma_2d(a + TG_BM * 128, &map_w, t * p.Cp + kb * TG_BK, n0, &sh->full[s]);
                        }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TG_BM >> 4) << 24);
        int i = 0, nt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++nt) {
            const int acc = nt & 1;
            if (nt >= 2) tg_mbar_wait(&sh->tmem_empty[acc], (uint32_t)((nt >> 1) - 1) & 1u);   // the epilogue has drained this set
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t dcol = tmem + (uint32_t)acc * ACC_COLS;
            for (int it = 0; it < niter; ++it, ++i) {
                const int s = i % p.stages;
                tg_mbar_wait(&sh->full[s], (uint32_t)(i / p.stages) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tg_elect_one()) {
                    const uint32_t a = tg_smem_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t ad = tg_desc_sw128(a), bd = tg_desc_sw128(a + TG_BM * 128);
#pragma unroll
                    for (int k = 0; k < TG_BK / 16; ++k)
                        tg_umma(dcol, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (it | k) ? 1u : 0u);
                    tg_commit(&sh->empty[s]);
                    if (it == niter - 1) tg_commit(&sh->tmem_full[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue =====
        // TMEM hands every thread one POSITION (lane = row) with the channels in registers; global memory wants the opposite
        // (consecutive lanes = consecutive channels of one row: [L][N] fp32 and [L][planes][Cp] bf16 are channels-last).
        // So each warp transposes 32 rows x 32 channels at a time through a padded shared-memory tile and does ALL the
        // epilogue arithmetic in the transposed domain: lane = channel (bias / scale / SnakeBeta constants are loaded once per
        // chunk), loop over the 32 rows with 128-byte coalesced residual loads and stores. (Thread-per-row stores of 16 bytes
        // at a row stride of N * 4 bytes half-fill every sector: the k = 1 convolutions took as long as the k = 7 ones.)
        // The elementwise tail (SnakeBeta = sinf per element, plane split) is what bounds these layers, not the MMAs: 16 warps
        // share it -- warp (2 + 4 g + i) owns TMEM lane quarter (warp & 3) and the 32-channel chunks g, g + 4, ...
        const int q = warp & 3, grp = (warp - 2) >> 2;
        constexpr int SP = TC_STG_PITCH;                             // 16-byte aligned rows, conflict-free both ways
        float* stg = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sh) + sizeof(TcShared) + 15) & ~(uintptr_t)15) + (warp - 2) * (32 * SP);
        int nt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++nt) {
        const int acc = nt & 1;
        const int l0 = (tile / n_ntiles) * TG_BM, n0 = (tile % n_ntiles) * p.BN;
        tg_mbar_wait(&sh->tmem_full[acc], (uint32_t)(nt >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lbase = l0 + q * 32;
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * ACC_COLS;
        const int nmax = min(p.BN, p.N - n0);
        const int rows = min(32, p.L - lbase);                       // live rows of this warp (<= 0: nothing to store)
        const int r4 = lane >> 3, c4 = (lane & 7) * 4;               // transposed domain: 8 lanes x 4 channels = one row's 32 channels
        for (int c = grp * 32; c < nmax; c += 32 * TC_EPI_GROUPS) {
            const int cw = min(32, nmax - c);                        // 16 or 32 channels in this chunk
            float v[16];
            tg_tmem_ld16(tbase + (uint32_t)c, v);
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(stg + lane * SP + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (cw > 16) {
                tg_tmem_ld16(tbase + (uint32_t)(c + 16), v);
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(stg + lane * SP + 16 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            __syncwarp();
            if (c4 < cw && rows > 0) {
                const int n = n0 + c + c4;
                const int bn = n % p.bias_mod;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f), one = make_float4(1.f, 1.f, 1.f, 1.f);
                const float4 bias = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + bn)) : z;
                const float4 scale = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + bn)) : one;
                const int ph = n / p.cout, ch = n - ph * p.cout;
                const float4 ea = p.sn_ea ? __ldg(reinterpret_cast<const float4*>(p.sn_ea + ch)) : z;
                const float4 ib = p.sn_ea ? __ldg(reinterpret_cast<const float4*>(p.sn_ib + ch)) : z;
#pragma unroll 2
                for (int r = r4; r < rows; r += 4) {
                    const size_t l = (size_t)(lbase + r);
                    const float4 a = *reinterpret_cast<const float4*>(stg + r * SP + c4);
                    float x[4] = {a.x + bias.x, a.y + bias.y, a.z + bias.z, a.w + bias.w};
                    if (p.act == 3) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) x[j] = gelu_erf_f(x[j]);
                    } else if (p.act == 1) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) x[j] = silu_f(x[j]);
                    }
                    x[0] *= scale.x; x[1] *= scale.y; x[2] *= scale.z; x[3] *= scale.w;
                    if (p.residual) {
                        const float4 w = *reinterpret_cast<const float4*>(p.residual + l * p.N + n);
                        x[0] += w.x; x[1] += w.y; x[2] += w.z; x[3] += w.w;
                    }
                    if (p.y && !p.y_snake) *reinterpret_cast<float4*>(p.y + l * p.N + n) = make_float4(x[0], x[1], x[2], x[3]);
                    if (p.sn_ea) {
                        float sn;
                        sn = sinf(x[0] * ea.x); x[0] = x[0] + ib.x * (sn * sn);
                        sn = sinf(x[1] * ea.y); x[1] = x[1] + ib.y * (sn * sn);
                        sn = sinf(x[2] * ea.z); x[2] = x[2] + ib.z * (sn * sn);
                        sn = sinf(x[3] * ea.w); x[3] = x[3] + ib.w * (sn * sn);
                    }
                    if (p.y && p.y_snake) *reinterpret_cast<float4*>(p.y + l * p.N + n) = make_float4(x[0], x[1], x[2], x[3]);
                    if (p.xo) {
                        __nv_bfloat16* xo = p.xo + ((l * p.up + ph) * p.oplanes) * p.oCp + ch;
                        for (int pl = 0; pl < p.oplanes; ++pl) {
                            const __nv_bfloat162 h0 = __floats2bfloat162_rn(x[0], x[1]), h1 = __floats2bfloat162_rn(x[2], x[3]);
                            uint2 u;
                            u.x = *reinterpret_cast<const uint32_t*>(&h0); u.y = *reinterpret_cast<const uint32_t*>(&h1);
                            *reinterpret_cast<uint2*>(xo + (size_t)pl * p.oCp) = u;
                            x[0] -= bf16lo(u.x); x[1] -= bf16hi(u.x); x[2] -= bf16lo(u.y); x[3] -= bf16hi(u.y);   // exact remainders
                        }
                    }
                }
            }
            __syncwarp();
        }
        // this warp has read its 32 lanes of the accumulator set: hand it back to the MMA warp
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tg_smem_u32(&sh->tmem_empty[acc])) : "memory");
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
