// TMA-fed tcgen05 GEMM for the BATCHED path (BASELINE configs[3]/[4]: many concurrent utterances, north_star (1):
// "TMA-fed tcgen05 tensor-core GEMMs once the batch makes them a dense contraction").
//
//   out[split][b][n] = sum_{k in split} W[n][k] * X[b][k]          W: bf16 [N, K] row-major (the .lqw layout, no regrouping)
//
// "Swap-AB" skinny GEMM: the WEIGHT rows are the M side of the MMA (128 rows per CTA tile = 128 TMEM lanes), the
// utterances are the N side. Activations stay fp32-exact when asked to: the fp32 row of utterance b is split into
// `planes` bf16 planes (hi / mid / lo: 8 + 8 + 8 mantissa bits, remainders exact -- the same trick as the batch-1 frame
// kernel) that occupy separate N columns, col = plane * Bt + b; every product weight x plane is exact in the fp32
// accumulator and the epilogue adds the planes. planes = 1 is the usual bf16-activation GEMM (north_star: logits within
// 2e-2), planes = 3 reproduces the oracle's fp32 activations (token-exact parity tests).
//
// Per CTA (192 threads): warp 0 = TMA producer (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes of 64 bf16 x {128 | Bt*planes}
// rows, mbarrier complete_tx), warp 1 = MMA issuer (one elected lane: tcgen05.mma.cta_group::1.kind::f16, M = 128,
// N = Bt*planes, K = 16, fp32 accumulators in TMEM; tcgen05.commit releases the ring stage), warps 2-5 = epilogue
// (tcgen05.ld 32x32b: one TMEM lane = one weight row per thread, plane sum, coalesced fp32 stores along n).
// Grid = (N / 128 weight tiles, K splits, utterance tiles). Split-K partials are written separately and summed in fixed
// order by the consumer kernel (batched.cuh), so results are bit-reproducible; the splits exist to put >= 128 CTAs on the
// machine when N is small (down projection: 8 tiles) -- these GEMMs are weight-streaming (HBM) bound until
// B * planes reaches the ridge (~213 flop/B).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace lqt {

constexpr int TG_BM = 128;            // weight rows per CTA tile (UMMA M)
constexpr int TG_BK = 64;             // bf16 per k-block = one 128-byte swizzle atom
constexpr int TG_THREADS = 192;
constexpr int TG_MAX_STAGES = 8;

struct TcGemmParams {
    int N, K;                 // weight rows, reduction length (K % 64 == 0)
    int n_split2;             // rows >= n_split2 come from the second weight map (gate | up); = N when there is one matrix
    int kb_per_split;         // k-blocks of 64 per split
    int BN;                   // MMA N = Bt * planes (multiple of 16, <= 256)
    int Bt, planes;           // utterances per tile (multiple of 16), bf16 planes per utterance
    int B;                    // valid utterances in total (rows b >= B are padding and are not stored)
    int stages;
    float* out;               // [splits][Bpad][N] fp32, Bpad = gridDim.z * Bt
    long long split_stride;   // elements between splits
};

LQT_DEVINL uint32_t tg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
LQT_DEVINL void tg_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tg_smem_u32(bar)), "r"(count));
}
LQT_DEVINL void tg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tg_smem_u32(bar)), "r"(bytes) : "memory");
}
LQT_DEVINL void tg_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}"
        ::"r"(tg_smem_u32(bar)), "r"(parity) : "memory");
}
LQT_DEVINL void tg_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(tg_smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(tg_smem_u32(bar)) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO 1 << 16 |
// SBO (8 rows x 128 B = 1024 B) >> 4 << 32 | version 1 << 46 | SWIZZLE_128B (2) << 61
LQT_DEVINL uint64_t tg_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
LQT_DEVINL void tg_umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
LQT_DEVINL void tg_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tg_smem_u32(bar)) : "memory");
}
LQT_DEVINL bool tg_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
LQT_DEVINL void tg_tmem_ld16(uint32_t taddr, float (&r)[16]) {      // this thread's TMEM lane, 16 consecutive columns
    uint32_t u[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}

struct TgShared {
    uint64_t full[TG_MAX_STAGES], empty[TG_MAX_STAGES], tmem_full;
    uint32_t tmem_base;
};

inline size_t tc_gemm_stage_bytes(int BN) { return (size_t)TG_BM * 128 + (size_t)BN * 128; }
inline size_t tc_gemm_smem_bytes(int BN, int stages) { return 1024 + (size_t)stages * tc_gemm_stage_bytes(BN) + sizeof(TgShared); }

template <int TMEM_COLS>
__global__ void __launch_bounds__(TG_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w2,
               const __grid_constant__ CUtensorMap map_x, const TcGemmParams p) {
    extern __shared__ __align__(1024) unsigned char tg_smem_raw[];
    // SWIZZLE_128B tiles must sit on 1024-byte boundaries of the shared window
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tg_smem_raw) + 1023) & ~(uintptr_t)1023);
    const size_t stage_bytes = (size_t)TG_BM * 128 + (size_t)p.BN * 128;
    TgShared* sh = reinterpret_cast<TgShared*>(smem + (size_t)p.stages * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * TG_BM;
    const int nkb = p.K / TG_BK;
    const int kb0 = blockIdx.y * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
    const int niter = kb1 - kb0;                                  // >= 1 by construction of the grid

    pdl_trigger();                                                // the next kernel's prologue may overlap this one
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
        // ===== TMA producer =====
        if (tg_elect_one()) {
            const bool second = row0 >= p.n_split2;
            const CUtensorMap* mw = second ? &map_w2 : &map_w;
            const int wrow = second ? row0 - p.n_split2 : row0;
            const int xrow = blockIdx.z * p.BN;
            const uint32_t bytes = (uint32_t)stage_bytes;
            // The WEIGHTS do not depend on the predecessor kernel: the first ring-full of weight boxes is requested before the
            // programmatic-dependency wait, so their HBM latency overlaps the predecessor's tail; only the X boxes (L2) follow it
            const int npre = min(niter, p.stages);
            for (int i = 0; i < npre; ++i) {
                tg_mbar_expect_tx(&sh->full[i], bytes);
                tg_tma_2d(smem + (size_t)i * stage_bytes, mw, (kb0 + i) * TG_BK, wrow, &sh->full[i]);
            }
            pdl_wait();                                           // the predecessor's outputs (X) are complete and visible
            for (int i = 0; i < npre; ++i)
                tg_tma_2d(smem + (size_t)i * stage_bytes + TG_BM * 128, &map_x, (kb0 + i) * TG_BK, xrow, &sh->full[i]);
            for (int i = npre; i < niter; ++i) {
                const int s = i % p.stages;
                tg_mbar_wait(&sh->empty[s], (uint32_t)((i / p.stages) - 1) & 1u);
                unsigned char* a = smem + (size_t)s * stage_bytes;
                tg_mbar_expect_tx(&sh->full[s], bytes);
                tg_tma_2d(a, mw, (kb0 + i) * TG_BK, wrow, &sh->full[s]);
                tg_tma_2d(a + TG_BM * 128, &map_x, (kb0 + i) * TG_BK, xrow, &sh->full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 | A bf16 | B bf16 | both K-major | N >> 3 | M >> 4
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TG_BM >> 4) << 24);
        for (int i = 0; i < niter; ++i) {
            const int s = i % p.stages;
            tg_mbar_wait(&sh->full[s], (uint32_t)(i / p.stages) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tg_elect_one()) {
                const uint32_t a = tg_smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t ad = tg_desc_sw128(a), bd = tg_desc_sw128(a + TG_BM * 128);
#pragma unroll
                for (int k = 0; k < TG_BK / 16; ++k)              // 32 bytes along K inside the swizzle atom = +2 in the address field
                    tg_umma(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (i | k) ? 1u : 0u);
                tg_commit(&sh->empty[s]);                         // the stage is free once these MMAs have read it
                if (i == niter - 1) tg_commit(&sh->tmem_full);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: TMEM lane (warp & 3) * 32 + lane = weight row; columns = plane * Bt + b =====
        tg_mbar_wait(&sh->tmem_full, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        pdl_wait();                                               // (returns at once: the X boxes were loaded behind the producer's wait)
        const int q = warp & 3;
        const int n = row0 + q * 32 + lane;
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16);
        const int b0 = blockIdx.z * p.Bt;
        float* out = p.out + (long long)blockIdx.y * p.split_stride;
        for (int c = 0; c < p.Bt; c += 16) {
            float v[16], w[16];
            tg_tmem_ld16(tbase + (uint32_t)c, v);
            if (p.planes > 1) {                                   // hi + (mid + lo): the small terms first
                tg_tmem_ld16(tbase + (uint32_t)(p.Bt + c), w);
                if (p.planes > 2) {
                    float u[16];
                    tg_tmem_ld16(tbase + (uint32_t)(2 * p.Bt + c), u);
#pragma unroll
                    for (int j = 0; j < 16; ++j) w[j] += u[j];
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += w[j];
            }
            if (n < p.N) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int b = b0 + c + j;
                    if (b < p.B) out[(long long)b * p.N + n] = v[j];
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

}  // namespace lqt
