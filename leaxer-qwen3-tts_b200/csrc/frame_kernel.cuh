// Persistent frame kernel: loops A and B of the reference (src/tts_onnx.cpp:782-872) as ONE
// cluster launch per utterance (or per chunk of frames).
//
// Why one kernel: a frame is 31 dependent network passes (1 talker step + 15 predictor passes, each
// followed by a draw) = ~475 dependent matrix-vector / attention phases of a few microseconds each. As
// separate launches the HBM pipe drains at every kernel boundary (round-1 v1: 577 launches, 4.2 ms per
// frame). Here every SM keeps one CTA resident (15 clusters of 8 CTAs on B200):
//   * warp 8 (producer) walks the static weight schedule of the whole launch and streams this CTA's
//     slice of every matrix into a shared-memory ring with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), running AHEAD of the math across phase boundaries.
//     The weights are regrouped once at load time into per-CTA images in mma A-fragment order, so the
//     ring is filled by contiguous copies and read with one conflict-free 16-byte load per MMA;
//   * warps 0-7 (consumers) run the phases. The kernel is bound by instruction issue and dependent
//     latency on these 8 warps, not by bandwidth, so the matrix-vector products run on the tensor
//     cores: mma.sync.m16n8k16 (bf16 x bf16 -> fp32) consumes a 16 x 16 weight block per instruction
//     where the fp32-FMA form needed ~20 instructions per 16-byte load. The fp32 input vector is split
//     into three bf16 planes (exact: 8 + 8 + 8 mantissa bits) that occupy three columns of the B
//     operand (gemv_mma). RMSNorm is folded: the product runs on x*w_norm, 1/rms is applied in the
//     epilogue. (Round-1 v10 ran the products on tcgen05: TMEM allocation, single-thread issue,
//     commit/mbarrier and tcgen05.ld add several dependent latencies per phase, which is what a batch-1
//     phase cannot afford; the synchronous warp-level MMA has none of that. tcgen05 returns with the
//     batched path);
//   * phases hand over through a grid counter (one arrival per CTA after it has ISSUED its stores, one
//     polling lane per CTA); activations travel between CTAs as 8-byte (value, sequence) words ("LL"
//     words, as in NCCL's low-latency protocol) that every reader validates, so no fence is needed; a
//     vector that every CTA needs is fetched ONCE per cluster with a TMA multicast copy (mc_fetch), and
//     the O-projection partials of the 8 kv groups are reduced through distributed shared memory;
//   * the sampler and the embedding glue (src/tts_onnx.cpp:803-842, 854-868, 878-950) run redundantly
//     in every CTA, so a draw costs no broadcast.
// Code size matters: the layer loop must stay warm in the instruction cache (one shared routine per kind
// of work; see tools/fk_codesize.py).
// Phases per layer: QKV | [talker: split-KV attention] | O-projection by kv-group (the code predictor
// computes its <=16-position attention inside this phase) | gate/up (SwiGLU) | down.
#pragma once
#include "attention.cuh"
#include "common.cuh"
#include "sampler.cuh"

namespace lqt {

typedef __nv_bfloat16 bf16_t;

extern __shared__ __align__(1024) unsigned char fk_smem[];     // the frame kernel's dynamic shared memory

constexpr int FK_CWARPS = 8;
constexpr int FK_CTHREADS = FK_CWARPS * 32;       // consumer threads
constexpr int FK_THREADS = FK_CTHREADS + 128;     // + one producer warpgroup (one lane works; a whole warpgroup so that setmaxnreg can hand its registers over)
constexpr int FK_STAGE_BYTES = 32 * 1024;         // one ring stage: a slice's fragment-ordered image is streamed as a flat byte stream
constexpr int FK_CLUSTER = 8;                     // CTAs per thread-block cluster (multicast + DSMEM domain)
constexpr int FK_LAND_WORDS = 3072;               // landing buffer of a multicast vector fetch (LL words), two of them
constexpr int FK_X1OWN = 16;                      // max rows of the down projection per CTA
constexpr int FK_RPP_MAX = 16;                    // O-projection rows reduced per CTA (DSMEM)
constexpr int FK_NS_MAX = 24;                     // max attention splits per kv group
constexpr int FK_ATT_MIN_CHUNK = 32;              // positions per split at least (8 warps x 4 positions in flight): short contexts use few
                                                  // splits, and the combine step reads one L2 round trip per four splits
constexpr int FK_NGRP_MAX = 8;                    // kv groups
constexpr int FK_PART_ROWS = 80;                  // max rows of a slice (64 for the flat phases, 72 for the O-projection of 15 clusters)
constexpr int FK_TILE_ROWS = 8;                   // rows per weight tile of the tensor-core matrix-vector phases
constexpr int FK_PT_MAX = 128;                    // KV pages per slot at most (8192 positions)
constexpr int FK_CP_POS = 32;                     // code-predictor KV capacity (positions)
constexpr int FK_RED_STRIDE = 96;                 // warp-partial row sums: [warp][FK_RED_STRIDE]; M = 2 -> second row at +48
constexpr int FK_XS_STRIDE = 512;                 // attention output rows (input of the grouped O-projection)
constexpr int FK_MAX_TLAYERS = 32, FK_MAX_CLAYERS = 8;
constexpr unsigned long long FK_SPIN_LIMIT = 6000000000ull;   // ~3 s of SM clocks: abort, never hang

// Weight "images": for every matrix and every CTA c, the rows this CTA owns, row-major [rmax][K] bf16
// (rmax = the largest row count of any CTA, unused rows zero); CTA c's image starts at c * rmax * K
// elements. Built once at engine init (fk_build_image_kernel) from the .lqw tensors.
struct FkLayer {
    const bf16_t* wqkv;     // image of [q_dim + 2 kv_dim][H]
    const bf16_t* wo_g;     // image of the O-projection sliced by kv group (K = rep*128)
    const bf16_t* wgu;      // image of gate/up interleaved (row 2n = gate n, 2n+1 = up n)
    const bf16_t* wdown;    // image of [H][inter]
    const float *ln1, *ln2, *qnorm, *knorm;
};

struct FkStack {
    int n_layers, H, heads, kv_heads, inter;
    const float *cos, *sin, *final_norm;
    uint2 *x, *qkv, *x1, *act;            // LL buffers: layer output [H], projections [qkv_dim], post-attention stream [H], SwiGLU [inter]
};

struct FkSmemOffsets { unsigned land, scratch, att, xs, red, nxt, res0, lh, shared; int maxV; };

struct FkParams {
    FkSmemOffsets so;          // carve-up of the dynamic shared memory (constant bank): shared-memory addresses are derived from the fk_smem
                               // symbol where they are used, so they are shared-space accesses (no generic->shared conversion, which under
                               // clusters costs an S2R SR_CgaCtaId per access) and occupy no registers
    FkStack talker, cp;
    FkLayer t_layers[FK_MAX_TLAYERS];     // in the kernel-parameter constant bank: no load latency
    FkLayer c_layers[FK_MAX_CLAYERS];
    const bf16_t* t_head; int vocab;          // images
    const bf16_t* c_heads; int cp_vocab, cp_steps; long long c_head_stride;   // image elements per predictor head
    const bf16_t* c_inproj_w; const float* c_inproj_b; uint2* cxin;     // 1.7B: talker width -> predictor width (LL [2][Hc])
    float eps;
    void* kv_pool; const int* page_table; int n_pages; int page_shift; long long page_stride; int kv_f32;   // n_pages <= FK_PT_MAX
    uint2* pa;                 // talker attention partials, LL [n_kv][FK_NS_MAX][2][ATT_PSTRIDE]
    float* cp_kv;              // PER-CTA copies [cta][layer][k|v][FK_CP_POS][128] fp32 of the predictor KV of the CTA's own kv group: every CTA
                               // computes the new k/v row of its group anyway, so it keeps them itself -- no writer fence, no shared lines
    uint2 *logits_ll, *clogits_ll;
    float *logits, *clogits, *last_hidden, *next_in;   // plain copies: API outputs, resume across launches, mode-1 input
    const bf16_t *codec_embed, *cp_embed;
    const float* prompt; int P;           // prefill rows (run when st->pos == 0)
    const float *trailing, *tts_pad;
    GenState* st; const SamplingDev* sp;
    long long* codes_out; const long long* forced; float* trace; int trace_stride;
    volatile int* progress;    // nullable, pinned HOST memory: frames completed so far (the host vocodes them while the kernel runs)
    unsigned* ctrl;            // [1] abort flag, [32] grid arrival counter (all zero at launch)
    int frame_end;             // run frames while frame < frame_end (<= max_frames)
    unsigned producer_sleep_ns; // back-off of the producer lane while the ring is full ($LQT_FK_SLEEP, default 800: 0-400 ns measured 0.5 % slower, 1600 ns 1.5 % slower)
    int mode;                  // 0 = prefill (if pos == 0) + frames; 1 = one talker token from next_in (head on), no frames
    unsigned long long* dbg;   // nullable: phase timeline of CTA dbg_cta, [0] = count
    int dbg_cap, dbg_cta;
};

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
LQT_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

LQT_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
LQT_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
LQT_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
LQT_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0)
LQT_DEVINL void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 1-D TMA bulk copy global -> the same shared offset of every CTA in `mask` of this cluster; each destination's
// mbarrier (same offset) receives complete_tx for the bytes written to it
LQT_DEVINL void bulk_g2s_mc(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
LQT_DEVINL uint32_t dsmem_addr(const void* local, unsigned cta_rank) {       // address of `local` in another CTA of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(cta_rank));
    return r;
}
LQT_DEVINL void st_ll_dsmem(uint32_t addr, float v, unsigned seq) {
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(__float_as_uint(v)), "r"(seq) : "memory");
}
LQT_DEVINL uint2 ld_ll_smem(const uint2* p) {
    uint2 r;
    asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(smem_u32(p)) : "memory");
    return r;
}
LQT_DEVINL void cluster_sync_all() {                // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
LQT_DEVINL void csync() { asm volatile("bar.sync 1, %0;" ::"n"(FK_CTHREADS) : "memory"); }
// plain loads (not asm volatile) so that ptxas can batch them ahead of the FMAs; visibility of the TMA
// writes is ordered by the mbarrier wait (asm volatile with a memory clobber) that precedes them
LQT_DEVINL uint2 lds64(const void* p) { return *reinterpret_cast<const uint2*>(p); }
LQT_DEVINL uint4 lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }
// shared-space load by 32-bit shared address: the generic->shared conversion (an S2R SR_CgaCtaId under clusters) is done once
// by the caller instead of once per access. volatile: must stay behind the mbarrier wait that made the data visible.
LQT_DEVINL uint4 lds128_s(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
// LL exchange: 8-byte (value, sequence) words. volatile accesses always go to L2 (the coherence point).
LQT_DEVINL void st_ll(uint2* p, float v, unsigned seq) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(seq) : "memory");
}
LQT_DEVINL uint4 ld_ll2(const uint2* p) {          // two consecutive words (16-byte aligned)
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
LQT_DEVINL uint2 ld_ll1(const uint2* p) {
    uint2 r;
    asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}

// ------------------------------------------------------------------------------------------------
// shared-memory layout
// ------------------------------------------------------------------------------------------------
struct FkDesc {                 // this CTA's weight slice of one phase kind
    int row0, nrows, K, RG;
    int nst;                    // ring stages of the slice (flat byte stream, FK_STAGE_BYTES each)
    unsigned img_off;           // element offset of this CTA's image inside the matrix image
    int lw;                     // log2(warps per tile pair) of the tensor-core product (gemv_mma)
    int rpp;                    // O-projection: rows reduced per cluster partner
};
// One warp's share of a matrix-vector phase, precomputed at kernel start (per stack, phase kind and consumer warp) so that the
// product starts with one 16-byte shared load instead of ~100 dependent integer instructions: the (tile pair, K slice) unit of gemv_mma.
struct alignas(16) FkUnit {
    uint32_t off;               // byte offset of the unit's first block inside the slice image (before the lane offset)
    uint16_t step;              // bytes between consecutive blocks of this warp (block size x warps per pair)
    uint16_t iters;             // blocks of this warp (even)
    uint8_t st0, st1;           // first / last ring stage of the slice that the unit reads
    uint8_t flags;              // 1 = this warp has a unit, 2 = the unit is the odd last tile (8 rows)
    uint8_t ks;                 // K slice index = row of part[][]
    uint16_t pidx;              // first row of the unit inside the slice
    uint16_t binc;              // B-fragment stride of this warp (96 bytes x warps per pair)
};
enum { FKU_ACTIVE = 1, FKU_SINGLE = 2 };
struct FkShared {
    uint64_t full[16];
    uint64_t empty[16];
    volatile int stop;            // consumers -> producer: stop issuing
    volatile int consumed;        // stages consumed when stop was raised
    volatile int aborted;
    int hist[256];
    int wtot[FK_CWARPS];
    float redf[FK_CWARPS][2];
    int redi[FK_CWARPS];
    uint32_t sel_prefix; int sel_k;
    int tok; float fsum;
    float ssred[2][FK_CWARPS];        // [phase parity][warp]: partial sums of squares (RMSNorm)
    alignas(8) uint32_t zero8[2];     // always zero: B-fragment source of the lanes that hold the zero columns (gemv_mma)
    float part[FK_CWARPS][FK_PART_ROWS];  // matrix-vector partial sums [warp][row of the slice] (K is split over the warps)
    float x1own[FK_X1OWN];            // post-attention stream at the rows of this CTA's down-projection slice (its residual)
    uint64_t land_bar[2];             // multicast landing buffers: complete_tx from all CTAs of the cluster
    uint2 redc[FK_NGRP_MAX][FK_RPP_MAX];  // O-projection partials of the partner CTAs (DSMEM, (value, sequence) words)         // post-attention residual stream at the rows of this CTA's down-projection slice
    FkDesc desc[2][10];           // [stack][phase kind]
    FkUnit unit[2][10][FK_CWARPS];   // [stack][phase kind][consumer warp]
    int pt[FK_PT_MAX];                // the slot's page table (static during a launch)
    uint16_t pushmap[2][FK_PART_ROWS];   // O-projection row r of this CTA's slice -> (cluster partner that reduces it << 8) | its slot there
};

struct FkCtx {
    const FkParams* p;
    unsigned land_n;          // fetches so far (buffer = land_n & 1, barrier parity = (land_n >> 1) & 1)
    unsigned land_a;          // landing buffer (0/1) that holds this layer's input row (residual of the O-projection)
    unsigned rank;            // CTA rank in the cluster
    int tid, lane, warp;
    int cta, ncta;
    unsigned seq;             // number of the current phase (1, 2, ...): tag of everything it publishes
    unsigned stage_ctr;       // ring stages consumed so far
    bool aborted;
    unsigned long long* dbg; int dbg_n, dbg_cap, dbg_tag;   // dbg_tag = (stack << 9) | (kind << 4) of the current phase
};

// shared-memory regions, derived from the carve-up in the constant bank (see fk_smem_layout)
#define FK_SH(c)   (reinterpret_cast<FkShared*>(fk_smem + (c).p->so.shared))
#define FK_RING(c) (fk_smem)
#define FK_ATT(c)  (reinterpret_cast<float*>(fk_smem + (c).p->so.att))
#define FK_XS(c)   (reinterpret_cast<float*>(fk_smem + (c).p->so.xs))
#define FK_XP(c)   (reinterpret_cast<float*>(fk_smem + (c).p->so.scratch))
#define FK_LAND(c) (reinterpret_cast<uint2*>(fk_smem + (c).p->so.land))
#define FK_NXT(c)  (reinterpret_cast<float*>(fk_smem + (c).p->so.nxt))
#define FK_RES0(c) (reinterpret_cast<float*>(fk_smem + (c).p->so.res0))
#define FK_LH(c)   (reinterpret_cast<float*>(fk_smem + (c).p->so.lh))

// timeline entries: (SM clock << 16) | (stack << 9) | (phase kind << 4) | point. Points:
//  0 phase begin   2 inputs in registers (polling done)   3 inputs complete (attention / staging done; before the GEMV)
//  4 glue done (sampler phases)   5 weights consumed (all FMAs done)   6 outputs published
enum { FKT_A = 1, FKT_B = 2, FKT_C = 3, FKT_D = 4, FKT_E = 5, FKT_HEAD = 6, FKT_SAMPLE = 7, FKT_INPROJ = 8 };
// The marks are compiled only into the profiling build (liblqt_b200_prof.so, -DFK_MARKS; tools/fk_timeline.py loads it through
// $LQT_B200_LIB): even disabled, ~10 mark sites per phase cost ~2 % of the frame in this issue-bound kernel.
#ifdef FK_MARKS
// thread 0 records into the first half of the buffer, lane 0 of warp FK_MARK_W2 into the second half (same SM clock: the two
// timelines show where the critical warp 0 lags behind an ordinary consumer warp)
#ifndef FK_MARK_W2
#define FK_MARK_W2 5
#endif
LQT_DEVINL void fk_mark(FkCtx& c, int point) {
    if (c.dbg && c.lane == 0 && (c.warp == 0 || c.warp == FK_MARK_W2) && c.dbg_n < c.dbg_cap)
        c.dbg[c.dbg_n++] = ((unsigned long long)clock64() << 16) | (unsigned)(c.dbg_tag | point);
}
LQT_DEVINL void fk_phase(FkCtx& c, int stack, int kind) { c.dbg_tag = (stack << 9) | (kind << 4); }
#else
LQT_DEVINL void fk_mark(FkCtx&, int) {}
LQT_DEVINL void fk_phase(FkCtx&, int, int) {}
#endif

// ------------------------------------------------------------------------------------------------
// weight slices: which rows of a matrix this CTA owns (identical arithmetic in producer and consumers)
// ------------------------------------------------------------------------------------------------
struct FkSlice { int row0, nrows; };

// N rows in groups of RG consecutive rows, dealt contiguously over all CTAs
LQT_DEVINL FkSlice flat_slice(int N, int RG, int cta, int ncta) {
    const int ng = N / RG;
    const int g0 = (int)(((unsigned)cta * (unsigned)ng) / (unsigned)ncta), g1 = (int)(((unsigned)(cta + 1) * (unsigned)ng) / (unsigned)ncta);
    return FkSlice{g0 * RG, (g1 - g0) * RG};
}
// kv-group decomposition: CTA c serves group c % n_kv as member c / n_kv of ns_g members
LQT_DEVINL int grp_members(int g, int n_kv, int ncta) { return (ncta - g + n_kv - 1) / n_kv; }
LQT_DEVINL FkSlice group_slice(int Nout, int cta, int ncta, int n_kv) {
    const int g = cta % n_kv, s = cta / n_kv, ns = grp_members(g, n_kv, ncta);
    const int T = Nout / FK_TILE_ROWS;           // rows are dealt in tiles of 8 (tensor-core fragment height)
    const int t0 = (int)(((unsigned)s * (unsigned)T) / (unsigned)ns), t1 = (int)(((unsigned)(s + 1) * (unsigned)T) / (unsigned)ns);
    return FkSlice{t0 * FK_TILE_ROWS, (t1 - t0) * FK_TILE_ROWS};
}
LQT_DEVINL FkDesc make_desc(const FkParams& p, bool is_cp, int kind, int cta, int ncta) {
    const FkStack& S = is_cp ? p.cp : p.talker;
    const int H = S.H, qkv_dim = (S.heads + 2 * S.kv_heads) * ATT_D, gK = (S.heads / S.kv_heads) * ATT_D;
    FkSlice s{0, 0}; int K = 256, RG = 1, rmax = 0;
    // the matrix-vector phases deal rows in tiles of FK_TILE_ROWS (the tensor-core fragment height); a gate/up pair never straddles a tile
    auto flat_max = [&](int N, int rg) { return ((N / rg + ncta - 1) / ncta) * rg; };
    switch (kind) {
        case FKT_INPROJ: s = flat_slice(H, FK_TILE_ROWS, cta, ncta); K = p.talker.H; rmax = flat_max(H, FK_TILE_ROWS); break;
        case FKT_A: s = flat_slice(qkv_dim, FK_TILE_ROWS, cta, ncta); K = H; rmax = flat_max(qkv_dim, FK_TILE_ROWS); break;
        case FKT_C: s = group_slice(H, cta, ncta, S.kv_heads); K = gK; rmax = ((H / FK_TILE_ROWS + ncta / S.kv_heads - 1) / (ncta / S.kv_heads)) * FK_TILE_ROWS; break;
        case FKT_D: s = flat_slice(2 * S.inter, FK_TILE_ROWS, cta, ncta); K = H; RG = 2; rmax = flat_max(2 * S.inter, FK_TILE_ROWS); break;
        case FKT_E: s = flat_slice(H, FK_TILE_ROWS, cta, ncta); K = S.inter; rmax = flat_max(H, FK_TILE_ROWS); break;
        case FKT_HEAD: { const int V = is_cp ? p.cp_vocab : p.vocab; s = flat_slice(V, FK_TILE_ROWS, cta, ncta); K = H; rmax = flat_max(V, FK_TILE_ROWS); break; }
        default: break;
    }
    FkDesc d;
    d.row0 = s.row0; d.nrows = s.nrows; d.K = K; d.RG = RG;
    d.nst = (int)(((unsigned)s.nrows * (unsigned)K * 2u + FK_STAGE_BYTES - 1) / FK_STAGE_BYTES);
    d.img_off = (unsigned)cta * (unsigned)rmax * (unsigned)K;
    { const int npu = (s.nrows + 15) >> 4; d.lw = npu <= 1 ? 3 : npu <= 2 ? 2 : npu <= 4 ? 1 : 0; }   // the fewer pairs, the more warps split K
    d.rpp = (s.nrows + S.kv_heads - 1) / S.kv_heads;
    return d;
}

LQT_DEVINL FkUnit make_unit(const FkDesc& d, int warp) {
    FkUnit u{};
    const int nkt = d.K >> 4, nt = d.nrows >> 3, npair = nt >> 1, npu = npair + (nt & 1);
    const int wpp = 1 << d.lw, ks = warp & (wpp - 1), p = warp >> d.lw;
    // npu <= FK_CWARPS >> lw: the host admits at most 64 rows per slice (80 for the O-projection) = 10 tiles = 5 units (fk_init)
    if (p < npu && nkt > 0) {
        const bool single = p == npair;
        const uint32_t bsz = single ? 256u : 512u;
        const uint32_t base = (uint32_t)p * (uint32_t)nkt * 512u, end = base + (uint32_t)nkt * bsz;
        u.flags |= FKU_ACTIVE | (single ? FKU_SINGLE : 0);
        u.off = base + (uint32_t)ks * bsz;
        u.step = (uint16_t)(bsz * (uint32_t)wpp);
        u.iters = (uint16_t)(nkt >> d.lw);
        u.st0 = (uint8_t)(base / FK_STAGE_BYTES); u.st1 = (uint8_t)((end - 1u) / FK_STAGE_BYTES);
        u.ks = (uint8_t)ks;
        u.pidx = (uint16_t)(p * 16);
        u.binc = (uint16_t)(96 * wpp);
    }
    return u;
}

// flat schedule of one token pass: [in_proj] + n_layers x (A qkv, B attention, C o-proj, D gate/up, E down) + [head]
struct FkOp { int kind, layer; };
LQT_DEVINL int pass_ops(int n_layers, bool inproj, bool head) { return (inproj ? 1 : 0) + n_layers * 5 + (head ? 1 : 0); }
LQT_DEVINL FkOp pass_op(int it, int n_layers, bool inproj) {
    const int n_pre = inproj ? 1 : 0;
    if (it < n_pre) return FkOp{FKT_INPROJ, 0};
    const int r = it - n_pre;
    if (r < n_layers * 5) return FkOp{1 + r % 5, r / 5};
    return FkOp{FKT_HEAD, 0};
}
// which pass comes q-th in this launch (identical in producer and consumers as long as no EOS)
struct FkPassId { bool is_cp; int cb; bool head; };    // cb: predictor pass index (0..cp_steps-1)
LQT_DEVINL FkPassId launch_pass(long long q, int mode, int n_prefill, int cp_steps) {
    if (mode == 1) return FkPassId{false, 0, true};
    if (q < n_prefill) return FkPassId{false, 0, q == n_prefill - 1};
    // per frame: predictor row 0 (talker last_hidden, no head), predictor passes 0..cp_steps-1 (head = pass), talker step
    const int r = (int)((q - n_prefill) % (cp_steps + 2));
    if (r == 0) return FkPassId{true, 0, false};
    return (r <= cp_steps) ? FkPassId{true, r - 1, true} : FkPassId{false, 0, true};
}

// ------------------------------------------------------------------------------------------------
// LL polling. The retry paths are cold and kept out of line (instruction-cache footprint of the layer
// loop matters); they take plain scalars so that the context struct can stay in registers.
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ bool ll_giveup_slow(volatile int* aborted, unsigned* ctrl, unsigned long long* t0) {
    if (*aborted || *reinterpret_cast<volatile unsigned*>(&ctrl[1]) != 0) { *aborted = 1; return true; }
    if (*t0 == 0) { *t0 = clock64(); return false; }
    if (clock64() - *t0 > FK_SPIN_LIMIT) { atomicExch(&ctrl[1], 1u); *aborted = 1; return true; }
    return false;
}
struct FkLL4 { uint4 a, b; };
__device__ __noinline__ FkLL4 ll_poll4_slow(const uint2* p, unsigned seq, volatile int* aborted, unsigned* ctrl) {
    FkLL4 r;
    int spins = 0; unsigned long long t0 = 0;
    do {
        if ((++spins & 255) == 0 && ll_giveup_slow(aborted, ctrl, &t0)) { r.a = ld_ll2(p); r.b = ld_ll2(p + 2); break; }
        __nanosleep(40);
        r.a = ld_ll2(p); r.b = ld_ll2(p + 2);
    } while (r.a.y != seq || r.a.w != seq || r.b.y != seq || r.b.w != seq);
    return r;
}
__device__ __noinline__ uint2 ll_poll1_slow(const uint2* p, unsigned seq, volatile int* aborted, unsigned* ctrl, unsigned sleep_ns) {
    uint2 a;
    int spins = 0; unsigned long long t0 = 0;
    do {
        if ((++spins & 255) == 0 && ll_giveup_slow(aborted, ctrl, &t0)) { a = ld_ll1(p); break; }
        __nanosleep(sleep_ns);
        a = ld_ll1(p);
    } while (a.y != seq);
    return a;
}
__device__ __noinline__ bool wait_full_slow(uint64_t* bar, unsigned par, unsigned* ctrl) {
    unsigned long long t0 = clock64();
    while (!mbar_try_wait(bar, par)) {
        if (clock64() - t0 > FK_SPIN_LIMIT) { atomicExch(&ctrl[1], 1u); return false; }
    }
    return true;
}

// four consecutive words (32-byte aligned group), all tagged `seq`
struct FkRaw4 { uint4 a, b; };
LQT_DEVINL FkRaw4 ll_issue4(const uint2* p) { FkRaw4 r; r.a = ld_ll2(p); r.b = ld_ll2(p + 2); return r; }
LQT_DEVINL float4 ll_finish4(FkCtx& c, FkRaw4 r, const uint2* p, unsigned seq) {
    if (r.a.y != seq || r.a.w != seq || r.b.y != seq || r.b.w != seq) {
        const FkLL4 q = ll_poll4_slow(p, seq, &FK_SH(c)->aborted, c.p->ctrl);
        r.a = q.a; r.b = q.b;
    }
    return make_float4(__uint_as_float(r.a.x), __uint_as_float(r.a.z), __uint_as_float(r.b.x), __uint_as_float(r.b.z));
}
LQT_DEVINL float4 ll_poll4(FkCtx& c, const uint2* p, unsigned seq) { return ll_finish4(c, ll_issue4(p), p, seq); }

// ------------------------------------------------------------------------------------------------
// Phase hand-over. Polling the data words themselves from 38k threads swamps the L2 slices that hold them
// (and delays the very stores being waited for), so the "when" travels separately: after a CTA has ISSUED
// the stores of phase n, one lane adds 1 to a grid counter (no fence: the stores may still be in flight);
// the next phase starts when one lane per CTA sees ncta * n, then every thread reads its inputs once and
// validates the (value, sequence) words -- a word whose store has not landed yet is simply re-read.
// ------------------------------------------------------------------------------------------------
LQT_DEVINL void grid_arrive(FkCtx& c) {          // one lane, after the CTA's stores of this phase were issued
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(c.p->ctrl + 32) : "memory");
}
__device__ __noinline__ void grid_wait_slow(const unsigned* ctr, unsigned target, volatile int* aborted, unsigned* ctrl) {
    int spins = 0; unsigned long long t0 = 0;
    // (Measured and rejected: four polls in flight a quarter of a round trip apart -- 2.15-2.19 instead of 2.10 ms per frame: more
    // polling delays the arrivals it waits for.)
    unsigned v;
    do {
        if ((++spins & 1023) == 0 && ll_giveup_slow(aborted, ctrl, &t0)) break;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
#ifdef FK_POLL_BACKOFF
        if ((int)(target - v) > FK_POLL_BACKOFF) __nanosleep(120);      // many CTAs still missing: do not crowd the counter's L2 slice
#endif
    } while ((int)(v - target) < 0);
}
// all consumer threads; returns when every CTA has issued the outputs of phase `n`.
// The poller is lane 0 of warp 1, not of warp 0: warp 0 runs the epilogue of the previous phase, the other warps get here early, so the
// poll is already in flight when the last arrival lands and neither the epilogue's tail nor the loop top of warp 0 sits between "every
// CTA has arrived" and the request for the next input vector. For the same reason the poller issues that request (fetch_src != nullptr:
// the cluster-multicast copy of mc_fetch) straight after the poll, before the CTA barrier. This is safe: a landing buffer is rewritten
// two fetches after it was filled, i.e. behind a grid hand-over that every reader of its previous content -- in every CTA of the
// cluster -- has passed; and the only reader of a landing buffer in an epilogue (the O-projection's residual) reads the OTHER buffer.
constexpr int FK_POLLER = 32;
#if defined(FK_NO_MC) && !defined(FK_NO_LEADER_FETCH)
#define FK_NO_LEADER_FETCH                        // the per-CTA copy of the bisecting build needs every CTA to wait for the grid itself
#endif
#ifndef FK_PRODUCER_WARP
#define FK_PRODUCER_WARP 3                        // which warp of the producer warpgroup streams the weights = the scheduler it shares: 3 (with consumer warps 3 and 7) measured 0.5 % faster than 0 (with warp 0, which runs every epilogue)
#endif
LQT_DEVINL void mc_issue(FkCtx& c, const uint2* src, int W);
LQT_DEVINL void mc_arm(FkCtx& c, int W);
LQT_DEVINL void mc_issue_all(FkCtx& c, const uint2* src, int W);
LQT_DEVINL void grid_wait(FkCtx& c, unsigned n, const uint2* fetch_src = nullptr, int fetch_w = 0) {
    if (c.tid == FK_POLLER) {
#ifndef FK_NO_LEADER_FETCH
        // One poller per cluster when the phase starts with a fetch (2.033 -> 2.011 ms per frame; several polls in flight by that one poller: 2.03;
        // the same for the phases WITHOUT a fetch, rank 0 releasing its partners through distributed shared memory: 2.005 -> 2.026): the landing barrier of a CTA completes when ALL copies into its buffer
        // have landed, i.e. with eight issuers when the LAST of eight pollers has seen the counter (each samples it once per L2 round trip).
        // Rank 0 polls and copies the whole vector into all eight CTAs; the others only arm their landing barrier.
        if (fetch_src) {
            mc_arm(c, fetch_w);
            if (c.rank == 0) {
                if (n != 0) {
                    const unsigned target = n * (unsigned)c.ncta;
                    unsigned v;
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c.p->ctrl + 32) : "memory");
                    if ((int)(v - target) < 0) grid_wait_slow(c.p->ctrl + 32, target, &FK_SH(c)->aborted, c.p->ctrl);
                }
                mc_issue_all(c, fetch_src, fetch_w);
            }
        } else
#endif
        {
            if (n != 0) {
                const unsigned target = n * (unsigned)c.ncta;
                unsigned v;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c.p->ctrl + 32) : "memory");
                if ((int)(v - target) < 0) grid_wait_slow(c.p->ctrl + 32, target, &FK_SH(c)->aborted, c.p->ctrl);
            }
            if (fetch_src) mc_issue(c, fetch_src, fetch_w);
        }
    }
    if (c.warp != 0) fk_mark(c, 14);                  // (profiling build with FK_MARK_W2=1, the poller's warp: poll done, fetch issued)
    csync();
}
LQT_DEVINL float ll_poll1(FkCtx& c, const uint2* p, unsigned seq) {
    uint2 a = ld_ll1(p);
    if (a.y != seq) a = ll_poll1_slow(p, seq, &FK_SH(c)->aborted, c.p->ctrl, 100);
    return __uint_as_float(a.x);
}

// ------------------------------------------------------------------------------------------------
// K-split matrix-vector product over this CTA's rows, weights from the ring
// ------------------------------------------------------------------------------------------------
LQT_DEVINL void wait_full(FkCtx& c, unsigned st, int nstages) {
    const unsigned slot = st % (unsigned)nstages, par = (st / (unsigned)nstages) & 1u;
    if (!mbar_try_wait(&FK_SH(c)->full[slot], par)) {
        if (!wait_full_slow(&FK_SH(c)->full[slot], par, c.p->ctrl)) c.aborted = true;
    }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core matrix-vector product (every matrix phase, the grouped O-projection included).
// The kernel is bound by instruction issue on its 8 consumer warps, not by bandwidth: with fp32 FMAs every 16-byte shared load
// of weights costs ~20 instructions (bf16 -> fp32 unpacking + FFMA2). One mma.sync.m16n8k16 (bf16 x bf16 -> fp32) consumes a
// 16 x 16 weight block per instruction instead. Exactness is kept by splitting the fp32 input vector into three bf16 planes
// (x = hi + mid + lo, 8 + 8 + 8 mantissa bits; every product w * plane is exact in fp32) that occupy columns 0..2 of the
// 16 x 8 B operand; the three result columns are added at the end.
//  * Weight image (built once by the host, engine.cu fk_build_image_kernel): this CTA's rows in tiles of 8; two tiles form the
//    16 rows of an A operand. For tile pair p and 16-column block kt the 32 lanes' A fragments (4 registers = 16 bytes each) are
//    stored contiguously, [p][kt][lane][a0 a1 a2 a3]; an odd last tile stores [kt][lane][a0 a2] (rows 8..15 of the operand are
//    zero registers). No padding: image bytes = rows * K * 2. One 16-byte shared load per lane per MMA, conflict-free.
//  * Input vector: B fragments in shared memory, [kt][plane][tg][b0 b1] (96 bytes per kt), written by the staging code.
//  * Every warp takes ONE (tile pair, K slice) unit: the fewer pairs a slice has, the more warps split the K dimension of each
//    (FkDesc.lw). Partial sums go to part[K slice][row]; one thread per row adds them in the epilogue.
// ------------------------------------------------------------------------------------------------
LQT_DEVINL void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
LQT_DEVINL uint2 lds64_s(uint32_t a) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
    return r;
}
// fp32 -> three bf16 planes (round-to-nearest each; the remainders are exact)
LQT_DEVINL void split3(float x, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);
    const __nv_bfloat16 l = __float2bfloat16_rn(r2);
    hi = (uint32_t)__bfloat16_as_ushort(h); mid = (uint32_t)__bfloat16_as_ushort(m); lo = (uint32_t)__bfloat16_as_ushort(l);
}
// two fp32 values -> three packed bf16x2 planes (low half = first value): one packed conversion per plane, exact remainders
LQT_DEVINL void split3x2(float x0, float x1, uint32_t& h, uint32_t& m, uint32_t& l) {
    const __nv_bfloat162 bh = __floats2bfloat162_rn(x0, x1);
    h = *reinterpret_cast<const uint32_t*>(&bh);
    const float r0 = x0 - __uint_as_float(h << 16), r1 = x1 - __uint_as_float(h & 0xffff0000u);
    const __nv_bfloat162 bm = __floats2bfloat162_rn(r0, r1);
    m = *reinterpret_cast<const uint32_t*>(&bm);
    const float q0 = r0 - __uint_as_float(m << 16), q1 = r1 - __uint_as_float(m & 0xffff0000u);
    const __nv_bfloat162 bl = __floats2bfloat162_rn(q0, q1);
    l = *reinterpret_cast<const uint32_t*>(&bl);
}
// this thread's four consecutive columns k..k+3 (k % 4 == 0) of the input vector -> B fragments
LQT_DEVINL void stage_bfrag(uint32_t xf_s, int k, float x0, float x1, float x2, float x3) {
    uint32_t h0, m0, l0, h1, m1, l1;
    split3x2(x0, x1, h0, m0, l0); split3x2(x2, x3, h1, m1, l1);
    const int kt = k >> 4, kk = k & 15;                          // kk in {0, 4, 8, 12}
    const uint32_t base = xf_s + (uint32_t)kt * 96u + (uint32_t)((kk & 7) >> 1) * 8u + (uint32_t)(kk >> 3) * 4u;   // (tg0, reg)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base), "r"(h0) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + 8u), "r"(h1) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + 32u), "r"(m0) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + 40u), "r"(m1) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + 64u), "r"(l0) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + 72u), "r"(l1) : "memory");
}
// one column k of an input vector -> its three bf16 planes in the B-fragment layout
LQT_DEVINL void stage_bfrag1(uint32_t xf_s, int k, float x) {
    uint32_t h, m, l;
    split3(x, h, m, l);
    const int kt = k >> 4, kk = k & 15;
    const uint32_t a = xf_s + (uint32_t)kt * 96u + (uint32_t)((kk & 7) >> 1) * 8u + (uint32_t)(kk >> 3) * 4u + (uint32_t)(kk & 1) * 2u;
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)h) : "memory");
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a + 32u), "h"((unsigned short)m) : "memory");
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a + 64u), "h"((unsigned short)l) : "memory");
}
template <int NST>
LQT_DEVINL uint32_t ring_wrap(uint32_t x) {
    constexpr uint32_t RING = (uint32_t)NST * FK_STAGE_BYTES;
    if constexpr ((NST & (NST - 1)) == 0) return x & (RING - 1u);
    else { while (x >= RING) x -= RING; return x; }
}
// blocks kt = ks, ks + wpp, ... of one tile pair (or of the odd last tile): `iters` (even) blocks, two accumulator sets.
// lin: ring offset of this lane's fragment of the first block, step: bytes between consecutive blocks of this warp;
// baddr / binc: this lane's B fragment and its stride (lanes that hold the zero columns of B read one fixed zero word
// pair, stride 0): no predicates, no bounds checks.
// (Measured and rejected: a software-pipelined form that requests the fragments of the next two blocks before the MMAs of the current
// two -- 12 more live registers, a few spills elsewhere in the kernel, 2.17 instead of 2.10 ms per frame.)
template <int NST, bool SINGLE>
LQT_DEVINL void mma_blocks(float (&acc)[2][4], uint32_t ring_s, uint32_t lin, uint32_t step, uint32_t baddr, uint32_t binc, int iters) {
#pragma unroll 1
    for (int it = 0; it < iters; it += 2) {
        const uint32_t l1 = ring_wrap<NST>(lin + step);
        uint4 a0, a1;
        if constexpr (SINGLE) {
            const uint2 t0 = lds64_s(ring_s + lin), t1 = lds64_s(ring_s + l1);
            a0 = make_uint4(t0.x, 0u, t0.y, 0u); a1 = make_uint4(t1.x, 0u, t1.y, 0u);
        } else {
            a0 = lds128_s(ring_s + lin); a1 = lds128_s(ring_s + l1);
        }
        const uint2 b0 = lds64_s(baddr), b1 = lds64_s(baddr + binc);
        mma_bf16_16816(acc[0], a0.x, a0.y, a0.z, a0.w, b0.x, b0.y);
        mma_bf16_16816(acc[1], a1.x, a1.y, a1.z, a1.w, b1.x, b1.y);
        lin = ring_wrap<NST>(l1 + step);
        baddr += 2u * binc;
    }
}

// Every warp takes ONE (tile pair, K slice) unit of the slice (FkUnit, precomputed): the per-unit overhead (stage waits,
// reduction, partial store) is paid once per warp and phase. gemv_wait runs early in the phase (while the input vector is still
// in flight: the weights were requested long before), gemv_mma after the inputs are staged, gemv_release after the CTA barrier
// that follows the product (one thread hands every stage of the slice back to the producer).
// (Kept inline with a rolled loop: the out-of-line variant had a better instruction-cache hit rate, 87.8 % vs 83.9 %, and was 1.3 % slower.)
LQT_DEVINL void gemv_wait(FkCtx& c, const FkUnit& u, int nstages) {
    if (u.flags & FKU_ACTIVE) {
#pragma unroll 1
        for (unsigned st = u.st0; st <= u.st1; ++st) wait_full(c, c.stage_ctr + st, nstages);
    }
}
template <int NST>
LQT_DEVINL void gemv_mma(FkCtx& c, const FkUnit& u, const uint32_t xf_s) {
    fk_mark(c, 9);
    if (u.flags & FKU_ACTIVE) {
        const uint32_t ring_s = smem_u32(FK_RING(c));
        const int g = c.lane >> 2, tg = c.lane & 3;
        // B fragments: lanes g < 3 hold the three planes; the others (zero columns of B) read a fixed pair of zero words
        const uint32_t b0addr = (g < 3) ? xf_s + (uint32_t)((g * 4 + tg) * 8) + (uint32_t)u.ks * 96u : smem_u32(&FK_SH(c)->zero8[0]);
        const uint32_t binc = (g < 3) ? (uint32_t)u.binc : 0u;
        const uint32_t ring0 = (c.stage_ctr % (unsigned)NST) * FK_STAGE_BYTES;       // ring offset of byte 0 of this slice
        const bool single = (u.flags & FKU_SINGLE) != 0;
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const uint32_t lin = ring_wrap<NST>(ring0 + u.off + c.lane * (single ? 8u : 16u));
        if (single) mma_blocks<NST, true>(acc, ring_s, lin, (uint32_t)u.step, b0addr, binc, (int)u.iters);
        else        mma_blocks<NST, false>(acc, ring_s, lin, (uint32_t)u.step, b0addr, binc, (int)u.iters);
        fk_mark(c, 10);
        // lane (g, tg): acc[.][0..1] = row g, columns 2tg, 2tg + 1; acc[.][2..3] = row g + 8. Columns 0..2 carry the planes.
        float v0 = (acc[0][0] + acc[1][0]) + (acc[0][1] + acc[1][1]);
        float v1 = (acc[0][2] + acc[1][2]) + (acc[0][3] + acc[1][3]);
        v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
        if (tg == 0) {
            float* part = &FK_SH(c)->part[u.ks][u.pidx];
            part[g] = v0;
            if (!single) part[8 + g] = v1;
        }
    }
    fk_mark(c, 11);
}
// after the CTA barrier that follows gemv_mma: every warp is done with every stage of the slice (those it never read included)
LQT_DEVINL void gemv_release(FkCtx& c, const FkDesc& d, int nstages) {
    if (c.tid == 32)
        for (int st = 0; st < d.nst; ++st) mbar_arrive(&FK_SH(c)->empty[(c.stage_ctr + st) % (unsigned)nstages]);
    c.stage_ctr += d.nst;
}
// sum of the K-slice partials of row r (after the CTA barrier that follows gemv_mma)
LQT_DEVINL float part_sum(FkCtx& c, int r, int wpp) {
    float v[FK_CWARPS];                            // all loads issued before the adds (this runs on the critical warp of every phase)
#pragma unroll
    for (int w = 0; w < FK_CWARPS; ++w) v[w] = (w < wpp) ? FK_SH(c)->part[w][r] : 0.f;
    return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
}

LQT_DEVINL float ss_rstd(FkCtx& c, int K, float eps) {
    const float* q = &FK_SH(c)->ssred[c.seq & 1u][0];
    float t = q[0];
#pragma unroll
    for (int w = 1; w < FK_CWARPS; ++w) t += q[w];
    return 1.0f / sqrtf(t / (float)K + eps);
}
// this thread's partial sum of squares -> shared (parity of c.seq); read after the phase's CTA barrier
LQT_DEVINL void ss_publish(FkCtx& c, float ss) {
    ss = warp_sum(ss);
    if (c.lane == 0) FK_SH(c)->ssred[c.seq & 1u][c.warp] = ss;
}

// ------------------------------------------------------------------------------------------------
// attention pieces
// ------------------------------------------------------------------------------------------------
// attention scratch layout (floats) inside FK_ATT(c)
constexpr int FA_Q = 0;                              // q_s   [2 m][2 r][128]
constexpr int FA_KN = FA_Q + 4 * ATT_D;              // knew  [2 m][128]
constexpr int FA_VN = FA_KN + 2 * ATT_D;             // vnew  [2 m][128]
constexpr int FA_SC = FA_VN + 2 * ATT_D;             // sc    [2 m][2 r][32]
constexpr int FA_WM = FA_SC + 4 * FK_CP_POS;         // wm    [8 w][2 r]  , wl [8][2]
constexpr int FA_WL = FA_WM + FK_CWARPS * 2;
constexpr int FA_WO = FA_WL + FK_CWARPS * 2;         // wo    [8 w][2 r][128]
constexpr int FA_FLOATS = FA_WO + FK_CWARPS * 2 * ATT_D;

// Four consecutive K/V values as loaded (bf16: 8 bytes, unpacked only where they are used -- an unpack right behind the load would make the
// warp wait for the load, and the loads are issued ahead of the grid hand-over precisely so that nobody waits for them there).
template <typename KVT> struct FkKvRaw;
template <> struct FkKvRaw<bf16_t> { uint2 u; };
template <> struct FkKvRaw<float> { float4 u; };
LQT_DEVINL void kv_load_raw(FkKvRaw<bf16_t>& r, const bf16_t* p) { r.u = __ldcg(reinterpret_cast<const uint2*>(p)); }
LQT_DEVINL void kv_load_raw(FkKvRaw<float>& r, const float* p) { r.u = __ldcg(reinterpret_cast<const float4*>(p)); }
LQT_DEVINL float4 kv_unpack(const FkKvRaw<bf16_t>& r) { return make_float4(bf16lo(r.u.x), bf16hi(r.u.x), bf16lo(r.u.y), bf16hi(r.u.y)); }
LQT_DEVINL float4 kv_unpack(const FkKvRaw<float>& r) { return r.u; }

// talker: split-KV partial attention of kv group g over this CTA's chunk of positions -> pa (LL).
// Everything that does not depend on the new row is requested BEFORE the grid hand-over (the function waits for the grid itself): the
// norm weights, cos / sin, and the first round of cached K/V rows (4 positions per warp = the whole chunk up to 32 positions per CTA),
// located through the copy of the slot's page table in shared memory. Returns true in the threads that stored the new K/V row: they
// owe a __threadfence() before the CTA's NEXT arrival (the caller issues it after this phase's arrival, off the critical path; the row
// is read one frame later at the earliest).
template <typename KVT>
LQT_DEVINL void talker_kv_round(FkCtx& c, const KVT* pool, long long goff, long long v_off, int jb, int j1, int t, FkKvRaw<KVT> (&kk)[4], FkKvRaw<KVT> (&vv)[4]) {
    const FkParams& p = *c.p;
    const int PS = 1 << p.page_shift;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int j = jb + u * FK_CWARPS;
        if (j < j1 && j != t) {
            const KVT* kp = pool + (long long)FK_SH(c)->pt[j >> p.page_shift] * p.page_stride + goff + (long long)(j & (PS - 1)) * ATT_D + c.lane * 4;
            kv_load_raw(kk[u], kp);
            kv_load_raw(vv[u], kp + v_off);
        }
    }
}
template <typename KVT>
LQT_DEVINL bool talker_attn_partial(FkCtx& c, const FkLayer& L, int layer, int t, unsigned want) {
    const FkParams& p = *c.p;
    const FkStack& S = p.talker;
    const int n_kv = S.kv_heads, g = c.cta % n_kv, s = c.cta / n_kv, ns = min(grp_members(g, n_kv, c.ncta), FK_NS_MAX);
    const int n_pos = t + 1, chunk = max(FK_ATT_MIN_CHUNK, (n_pos + ns - 1) / ns);
    const int j0 = s * chunk, j1 = min(n_pos, j0 + chunk);
    const bool active = !(s >= ns || j0 >= j1);                // else: idle split (short contexts / spare CTAs)
    const int PS = 1 << p.page_shift;
    const int q_dim = S.heads * ATT_D, kv_dim = n_kv * ATT_D;
    KVT* pool = reinterpret_cast<KVT*>(p.kv_pool);
    const long long layer_off = (long long)layer * 2 * n_kv * PS * ATT_D;
    const long long head_off = (long long)g * PS * ATT_D, v_off = (long long)n_kv * PS * ATT_D;
    float4 nw4 = make_float4(0.f, 0.f, 0.f, 0.f), c4 = nw4, s4 = nw4;
    FkKvRaw<KVT> kk[4], vv[4];
    if (active) {
        if (c.warp < 3) {
            nw4 = __ldg(reinterpret_cast<const float4*>(c.warp == 2 ? L.knorm : L.qnorm) + c.lane);
            c4 = __ldg(reinterpret_cast<const float4*>(S.cos + (size_t)t * (ATT_D / 2)) + (c.lane & 15));
            s4 = __ldg(reinterpret_cast<const float4*>(S.sin + (size_t)t * (ATT_D / 2)) + (c.lane & 15));
        }
        talker_kv_round<KVT>(c, pool, layer_off + head_off, v_off, j0 + c.warp, j1, t, kk, vv);
    }
    grid_wait(c, want);
    fk_mark(c, 1);
    if (!active) return false;
    float* q_s = FK_ATT(c) + FA_Q;
    float* kn = FK_ATT(c) + FA_KN;
    float* vn = FK_ATT(c) + FA_VN;
    const bool owns_new = (j1 == n_pos);
    bool fence = false;
    if (c.warp < 2) {
        float4 v = ll_poll4(c, S.qkv + (size_t)(g * 2 + c.warp) * ATT_D + c.lane * 4, want);
        v = head_norm_rope_regs(v, true, nw4, p.eps, c4, s4, c.lane);
        reinterpret_cast<float4*>(q_s + c.warp * ATT_D)[c.lane] = v;
    } else if (owns_new && c.warp < 4) {
        const long long base = (long long)FK_SH(c)->pt[t >> p.page_shift] * p.page_stride + layer_off + head_off +
                               (long long)(t & (PS - 1)) * ATT_D;
        if (c.warp == 2) {
            float4 v = ll_poll4(c, S.qkv + q_dim + (size_t)g * ATT_D + c.lane * 4, want);
            v = head_norm_rope_regs(v, true, nw4, p.eps, c4, s4, c.lane);
            KvIO<KVT>::store4(pool + base + c.lane * 4, v);
            v.x = KvIO<KVT>::round(v.x); v.y = KvIO<KVT>::round(v.y); v.z = KvIO<KVT>::round(v.z); v.w = KvIO<KVT>::round(v.w);
            reinterpret_cast<float4*>(kn)[c.lane] = v;
        } else {
            float4 v = ll_poll4(c, S.qkv + q_dim + kv_dim + (size_t)g * ATT_D + c.lane * 4, want);
            KvIO<KVT>::store4(pool + base + v_off + c.lane * 4, v);
            v.x = KvIO<KVT>::round(v.x); v.y = KvIO<KVT>::round(v.y); v.z = KvIO<KVT>::round(v.z); v.w = KvIO<KVT>::round(v.w);
            reinterpret_cast<float4*>(vn)[c.lane] = v;
        }
        fence = true;
    }
    csync();
    const float4 q0 = reinterpret_cast<const float4*>(q_s)[c.lane];
    const float4 q1 = reinterpret_cast<const float4*>(q_s + ATT_D)[c.lane];
    const float scale = 1.0f / sqrtf((float)ATT_D);
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    // The arithmetic of the loop is pinned with explicit intrinsics (which product is rounded and which is fused): the bf16 K/V rounding
    // amplifies a 1-ulp difference into other tokens within a few frames, so a refactoring must not change what the compiler contracts.
    for (int jb = j0 + c.warp;;) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = jb + u * FK_CWARPS;
            if (j < j1) {
                const float4 k4 = (j == t) ? reinterpret_cast<const float4*>(kn)[c.lane] : kv_unpack(kk[u]);
                const float4 v4 = (j == t) ? reinterpret_cast<const float4*>(vn)[c.lane] : kv_unpack(vv[u]);
                float d0 = __fmaf_rn(k4.w, q0.w, __fmaf_rn(k4.z, q0.z, __fmaf_rn(k4.x, q0.x, __fmul_rn(k4.y, q0.y))));
                float d1 = __fmaf_rn(k4.w, q1.w, __fmaf_rn(k4.z, q1.z, __fmaf_rn(k4.x, q1.x, __fmul_rn(k4.y, q1.y))));
                d0 = __fmul_rn(warp_sum(d0), scale); d1 = __fmul_rn(warp_sum(d1), scale);
                const float n0 = fmaxf(m0, d0), n1 = fmaxf(m1, d1);
                const float c0 = expf(m0 - n0), c1 = expf(m1 - n1), p0 = expf(d0 - n0), p1 = expf(d1 - n1);
                l0 = __fmaf_rn(l0, c0, p0); l1 = __fmaf_rn(l1, c1, p1);
                a0.x = __fmaf_rn(a0.x, c0, __fmul_rn(p0, v4.x)); a0.y = __fmaf_rn(a0.y, c0, __fmul_rn(p0, v4.y));
                a0.z = __fmaf_rn(a0.z, c0, __fmul_rn(p0, v4.z)); a0.w = __fmaf_rn(a0.w, c0, __fmul_rn(p0, v4.w));
                a1.x = __fmaf_rn(a1.x, c1, __fmul_rn(p1, v4.x)); a1.y = __fmaf_rn(a1.y, c1, __fmul_rn(p1, v4.y));
                a1.z = __fmaf_rn(a1.z, c1, __fmul_rn(p1, v4.z)); a1.w = __fmaf_rn(a1.w, c1, __fmul_rn(p1, v4.w));
                m0 = n0; m1 = n1;
            }
        }
        jb += FK_CWARPS * 4;
        if (jb >= j1) break;
        talker_kv_round<KVT>(c, pool, layer_off + head_off, v_off, jb, j1, t, kk, vv);
    }
    float* wm = FK_ATT(c) + FA_WM; float* wl = FK_ATT(c) + FA_WL; float* wo = FK_ATT(c) + FA_WO;
    reinterpret_cast<float4*>(wo + (c.warp * 2 + 0) * ATT_D)[c.lane] = a0;
    reinterpret_cast<float4*>(wo + (c.warp * 2 + 1) * ATT_D)[c.lane] = a1;
    if (c.lane == 0) { wm[c.warp * 2] = m0; wm[c.warp * 2 + 1] = m1; wl[c.warp * 2] = l0; wl[c.warp * 2 + 1] = l1; }
    csync();
    {
        const int r = c.tid >> 7, d = c.tid & 127;              // 256 threads = 2 heads x 128 dims
        float Mx = -INFINITY;
#pragma unroll
        for (int w2 = 0; w2 < FK_CWARPS; ++w2) Mx = fmaxf(Mx, wm[w2 * 2 + r]);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < FK_CWARPS; ++w2) {
            const float mw = wm[w2 * 2 + r];
            const float f = (mw == -INFINITY) ? 0.f : expf(mw - Mx);
            num = fmaf(f, wo[(w2 * 2 + r) * ATT_D + d], num);
            den = fmaf(f, wl[w2 * 2 + r], den);
        }
        uint2* part = p.pa + ((size_t)(g * FK_NS_MAX + s) * 2 + r) * ATT_PSTRIDE;
        st_ll(part + d, num, c.seq);
        if (d == 0) { st_ll(part + ATT_D, Mx, c.seq); st_ll(part + ATT_D + 1, den, c.seq); }
    }
    csync();                     // scratch is reused by the next phase
    return fence;
}

// talker: combine the splits of group g -> FK_XS(c)[0][0..rep*128)  (input of the grouped O-projection)
LQT_DEVINL void talker_attn_combine(FkCtx& c, int t, unsigned want, const FkUnit& u, int nstages) {
    const FkParams& p = *c.p;
    gemv_wait(c, u, nstages);
    const int n_kv = p.talker.kv_heads, g = c.cta % n_kv, ns = min(grp_members(g, n_kv, c.ncta), FK_NS_MAX);
    const int n_pos = t + 1, chunk = max(FK_ATT_MIN_CHUNK, (n_pos + ns - 1) / ns), active = (n_pos + chunk - 1) / chunk;
    const int r = c.tid >> 7, d = c.tid & 127;
    float Mx = -INFINITY, num = 0.f, den = 0.f;
    for (int s0 = 0; s0 < active; s0 += 4) {
        float ms[4], ls[4], os[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            ms[u] = -INFINITY; ls[u] = 0.f; os[u] = 0.f;
            if (s0 + u < active) {
                const uint2* ps = p.pa + ((size_t)(g * FK_NS_MAX + s0 + u) * 2 + r) * ATT_PSTRIDE;
                const uint4 ml = ld_ll2(ps + ATT_D);
                const uint2 ov = ld_ll1(ps + d);
                if (ml.y == want && ml.w == want && ov.y == want) {
                    ms[u] = __uint_as_float(ml.x); ls[u] = __uint_as_float(ml.z); os[u] = __uint_as_float(ov.x);
                } else {
                    ms[u] = ll_poll1(c, ps + ATT_D, want); ls[u] = ll_poll1(c, ps + ATT_D + 1, want); os[u] = ll_poll1(c, ps + d, want);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (s0 + u < active) {
                const float Mn = fmaxf(Mx, ms[u]);
                const float f0 = expf(Mx - Mn), f1 = expf(ms[u] - Mn);
                num = num * f0 + f1 * os[u];
                den = den * f0 + f1 * ls[u];
                Mx = Mn;
            }
        }
    }
    stage_bfrag1(smem_u32(FK_XS(c)), c.tid, num / den);        // column tid of the O-projection input, as B fragments
    csync();
}

// code predictor: full attention of kv group g for the ONE new position p0 (< FK_CP_POS / 2), result -> FK_XS(c)[0..256).
// Written for a small instruction footprint (it runs in every predictor layer): one polling / norm / rope path shared by
// q, k and v, two-value butterfly reductions.
LQT_DEVINL void cp_attn_local(FkCtx& c, const FkLayer& L, int layer, int p0, unsigned want, const FkUnit& u, int nstages) {
    const FkParams& p = *c.p;
    const FkStack& S = p.cp;
    const int n_kv = S.kv_heads, g = c.cta % n_kv;
    const int q_dim = S.heads * ATT_D, kv_dim = n_kv * ATT_D;
    float* q_s = FK_ATT(c) + FA_Q; float* kn = FK_ATT(c) + FA_KN; float* vn = FK_ATT(c) + FA_VN; float* sc = FK_ATT(c) + FA_SC;
    float* kc = p.cp_kv + (((size_t)c.cta * S.n_layers + layer) * 2 + 0) * FK_CP_POS * ATT_D;     // this CTA's private copy
    float* vc = kc + (size_t)FK_CP_POS * ATT_D;
    // prefetch the cached V column of this thread and the cached K rows of this warp (positions < p0)
    // while q/k/v of the new row are polled
    const int r_t = c.tid >> 7, d_t = c.tid & 127;
    float vcol[FK_CP_POS / 2];
#pragma unroll
    for (int j = 0; j < FK_CP_POS / 2; ++j) vcol[j] = (j < p0) ? __ldcg(vc + (size_t)j * ATT_D + d_t) : 0.f;
    float4 kpre[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int j = c.warp + u * FK_CWARPS;
        kpre[u] = (j < p0) ? __ldcg(reinterpret_cast<const float4*>(kc + (size_t)j * ATT_D) + c.lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // everything that does not depend on the new row is requested before the row -- and, like the cached K/V rows above, before the
    // grid hand-over, which this function waits for itself: norm weights, cos / sin (then the ring stages while the row is in flight)
    float4 nw4 = make_float4(0.f, 0.f, 0.f, 0.f), c4 = nw4, s4 = nw4;
    if (c.warp < 3) {
        nw4 = __ldg(reinterpret_cast<const float4*>(c.warp == 2 ? L.knorm : L.qnorm) + c.lane);
        c4 = __ldg(reinterpret_cast<const float4*>(S.cos + (size_t)p0 * (ATT_D / 2)) + (c.lane & 15));
        s4 = __ldg(reinterpret_cast<const float4*>(S.sin + (size_t)p0 * (ATT_D / 2)) + (c.lane & 15));
    }
#ifndef FK_DATAFLOW_QKV
    grid_wait(c, want);
#endif
    fk_mark(c, 1);
    if (c.warp < 4) {                                // warps 0, 1: the two q heads of the group; 2: k; 3: v
        const int job = c.warp;
        const int off = (job < 2) ? (g * 2 + job) * ATT_D : (job == 2 ? q_dim : q_dim + kv_dim) + g * ATT_D;
        const uint2* src = S.qkv + off + c.lane * 4;
        const FkRaw4 raw = ll_issue4(src);
        gemv_wait(c, u, nstages);
        float4 v = ll_finish4(c, raw, src, want);
        if (job < 3) v = head_norm_rope_regs(v, true, nw4, p.eps, c4, s4, c.lane);
        float* dst = (job < 2) ? q_s + job * ATT_D : (job == 2 ? kn : vn);
        reinterpret_cast<float4*>(dst)[c.lane] = v;
        if (job >= 2) reinterpret_cast<float4*>((job == 2 ? kc : vc) + (size_t)p0 * ATT_D)[c.lane] = v;   // read back only by this CTA, after CTA barriers
    } else {
        gemv_wait(c, u, nstages);
    }
    fk_mark(c, 13);
    csync();
    const float scale = 1.0f / sqrtf((float)ATT_D);
    // scores: key j against both heads (warp handles j = warp, warp + 8)
    {
        const float4 q0 = reinterpret_cast<const float4*>(q_s)[c.lane], q1 = reinterpret_cast<const float4*>(q_s + ATT_D)[c.lane];
        const float4 knew = reinterpret_cast<const float4*>(kn)[c.lane];
        const bool hi = (c.lane & 16) != 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = c.warp + u * FK_CWARPS;
            if (j <= p0) {
                const float4 k4 = (j == p0) ? knew : kpre[u];
                const float d0 = k4.x * q0.x + k4.y * q0.y + k4.z * q0.z + k4.w * q0.w;
                const float d1 = k4.x * q1.x + k4.y * q1.y + k4.z * q1.z + k4.w * q1.w;
                // both sums with five shuffles: the halves of the warp swap one value each, then reduce within the half
                float t = (hi ? d1 : d0) + __shfl_xor_sync(0xffffffffu, hi ? d0 : d1, 16);
                t += __shfl_xor_sync(0xffffffffu, t, 8);
                t += __shfl_xor_sync(0xffffffffu, t, 4);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                if ((c.lane & 15) == 0) sc[(c.lane >> 4) * FK_CP_POS + j] = t * scale;
            }
        }
    }
    csync();
    fk_mark(c, 14);
    if (c.warp < 2) {                                // softmax of head `warp` over j <= p0
        float* row = sc + c.warp * FK_CP_POS;
        const float v = (c.lane <= p0) ? row[c.lane] : -INFINITY;
        const float mx = warp_max(v);
        const float e = (c.lane <= p0) ? expf(v - mx) : 0.f;
        const float sum = warp_sum(e);
        if (c.lane <= p0) row[c.lane] = e / sum;
    }
    fk_mark(c, 5);
    csync();
    fk_mark(c, 15);
    {
        const float* row = sc + r_t * FK_CP_POS;
        float o = 0.f;
#pragma unroll
        for (int j = 0; j < FK_CP_POS / 2; ++j) if (j < p0) o = fmaf(row[j], vcol[j], o);
        o = fmaf(row[p0], vn[d_t], o);
        stage_bfrag1(smem_u32(FK_XS(c)), c.tid, o);            // column tid of the O-projection input, as B fragments
    }
    csync();
}
// ------------------------------------------------------------------------------------------------
// one token pass (one row) through a stack, as ONE loop over the flat op schedule so that every helper
// is instantiated exactly once. The layer-0 input row is in FK_RES0(c) (smem, plain layout). The residual
// stream of the current layer stays in registers (xin) from phase A to D.
// ------------------------------------------------------------------------------------------------
struct FkPass { bool is_cp; int pos0; bool head; };

LQT_DEVINL void f4_to(float (&d)[4], const float4& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }

// ------------------------------------------------------------------------------------------------
// Multicast fetch of a broadcast vector. Every CTA needs every activation vector, and 148 SMs reading the
// same freshly written lines is what saturates the L2 slices that hold them (tools/bench_exchange.cu:
// ~2 TB/s aggregate, which also delays the weight stream). So the eight CTAs of a cluster share the
// read: CTA r copies words [r*W/8, (r+1)*W/8) ONCE from L2 into the landing buffer of all eight
// (cp.async.bulk .multicast::cluster), every landing barrier collects complete_tx from all eight copies.
// Two landing buffers alternate; a buffer is rewritten two fetches later, i.e. behind at least one grid
// hand-over that every reader of its previous content has already passed.
// Readers validate the (value, sequence) words they use; a word whose store had not landed in L2 when the
// copy read it is re-polled from global memory and patched in place.
// ------------------------------------------------------------------------------------------------
LQT_DEVINL void mc_issue(FkCtx& c, const uint2* src, int W) {            // one thread (the poller)
#ifndef FK_NO_MC
    const unsigned b = c.land_n & 1u;
    uint2* dst = FK_LAND(c) + b * FK_LAND_WORDS;
    uint64_t* bar = &FK_SH(c)->land_bar[b];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic writes to the buffer (patches, sampler scratch)
    mbar_expect_tx(bar, (uint32_t)W * 8u);
    const int share = W / FK_CLUSTER;                                // W % 16 == 0: 16-byte multiples
    bulk_g2s_mc(dst + c.rank * share, src + c.rank * share, (uint32_t)share * 8u, bar, (uint16_t)((1u << FK_CLUSTER) - 1u));
#endif
}
LQT_DEVINL void mc_arm(FkCtx& c, int W) {                                  // one thread: expect W words in the current landing buffer
    const unsigned b = c.land_n & 1u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&FK_SH(c)->land_bar[b], (uint32_t)W * 8u);
}
LQT_DEVINL void mc_issue_all(FkCtx& c, const uint2* src, int W) {          // one thread of ONE CTA of the cluster: the whole vector into all eight
    const unsigned b = c.land_n & 1u;
    bulk_g2s_mc(FK_LAND(c) + b * FK_LAND_WORDS, src, (uint32_t)W * 8u, &FK_SH(c)->land_bar[b], (uint16_t)((1u << FK_CLUSTER) - 1u));
}
// issued = the request was already made by grid_wait
LQT_DEVINL uint2* mc_fetch(FkCtx& c, const uint2* src, int W, bool issued, const FkUnit* u = nullptr, int nstages = 1) {
    const unsigned b = c.land_n & 1u, par = (c.land_n >> 1) & 1u;
    uint2* dst = FK_LAND(c) + b * FK_LAND_WORDS;
    uint64_t* bar = &FK_SH(c)->land_bar[b];
#ifdef FK_NO_MC                                   // bisecting aid: plain per-CTA copy instead of the multicast
    for (int w = c.tid * 2; w < W; w += FK_CTHREADS * 2) *reinterpret_cast<uint4*>(dst + w) = ld_ll2(src + w);
    csync();
    ++c.land_n;
    (void)par; (void)bar; (void)issued;
    if (u) gemv_wait(c, *u, nstages);
    return dst;
#endif
    if (!issued && c.tid == FK_POLLER) mc_issue(c, src, W);
    if (u) gemv_wait(c, *u, nstages);                 // this warp's ring stages, while the vector is in flight
    if (!mbar_try_wait(bar, par)) {
        if (!wait_full_slow(bar, par, c.p->ctrl)) c.aborted = true;
    }
    fk_mark(c, 7);
    ++c.land_n;
    return dst;
}
// this thread's columns (4*tid.. of every 1024-chunk, J chunks) of a landed LL vector of K words tagged `want`
template <int J>
LQT_DEVINL void land_row(FkCtx& c, uint2* land, const uint2* src, int K, unsigned want, float (&out)[J][4]) {
    const int tid4 = c.tid * 4;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = j * 1024 + tid4;
        out[j][0] = 0.f; out[j][1] = 0.f; out[j][2] = 0.f; out[j][3] = 0.f;
        if (k < K) {
            uint4 a = *reinterpret_cast<const uint4*>(land + k), b = *reinterpret_cast<const uint4*>(land + k + 2);
            if (a.y != want || a.w != want || b.y != want || b.w != want) {
                const FkLL4 q = ll_poll4_slow(src + k, want, &FK_SH(c)->aborted, c.p->ctrl);
                a = q.a; b = q.b;
                *reinterpret_cast<uint4*>(land + k) = a; *reinterpret_cast<uint4*>(land + k + 2) = b;
            }
            out[j][0] = __uint_as_float(a.x); out[j][1] = __uint_as_float(a.z); out[j][2] = __uint_as_float(b.x); out[j][3] = __uint_as_float(b.z);
        }
    }
}
__device__ __noinline__ uint2 redc_poll_slow(const uint2* p, unsigned seq, volatile int* aborted, unsigned* ctrl) {
    uint2 a;
    int spins = 0; unsigned long long t0 = 0;
    do {
        if ((++spins & 1023) == 0 && ll_giveup_slow(aborted, ctrl, &t0)) { a = ld_ll_smem(p); break; }
        a = ld_ll_smem(p);
    } while (a.y != seq);
    return a;
}
// warp 0: sum the O-projection partials that the partner CTAs pushed into FK_SH(c)->redc (+ residual) for the rows this CTA
// reduces, publish the post-attention stream x1 (global LL), then signal the grid
LQT_DEVINL void reduce_partials(FkCtx& c, const FkDesc& d, int n_kv, int rpp, const float* res_plain, const uint2* res_land, uint2* x1) {
    if (c.warp != 0) return;
    const int g = (int)(c.rank % (unsigned)n_kv);
    const int r0 = g * rpp, nmine = max(0, min(d.nrows, r0 + rpp) - r0);
    if (c.lane < nmine) {
        const int row = d.row0 + r0 + c.lane;
        float acc = res_plain ? res_plain[row] : __uint_as_float(res_land[row].x);
        // (sequential on purpose: requesting all eight words before the first is examined samples them all at the earliest moment, and
        // every word that had not landed yet then takes the slow path -- measured slower for the CTAs that get here first)
#pragma unroll 1
        for (int gg = 0; gg < n_kv; ++gg) {
#ifdef FK_NO_DSMEM
            acc += ll_poll1(c, (c.p->pa + 8 * FK_NS_MAX * 2 * ATT_PSTRIDE) + (size_t)gg * 4096 + row, c.seq);
#else
            uint2 v = ld_ll_smem(&FK_SH(c)->redc[gg][c.lane]);
            if (v.y != c.seq) v = redc_poll_slow(&FK_SH(c)->redc[gg][c.lane], c.seq, &FK_SH(c)->aborted, c.p->ctrl);
            acc += __uint_as_float(v.x);
#endif
        }
        st_ll(x1 + row, acc, c.seq);
    }
    __syncwarp();
    if (c.lane == 0) grid_arrive(c);
}

template <int KJ, int NST>
LQT_DEVINL void consume_token(FkCtx& c, const FkParams& p, const FkPass& ps) {
    constexpr int HJ = (KJ > 3) ? 2 : 1;                         // 1024-column chunks of the hidden size
    const bool is_cp = ps.is_cp;
    const FkStack& S = is_cp ? p.cp : p.talker;
    const int tk = is_cp ? 1 : 0;
    const int H = S.H, n_kv = S.kv_heads;
    const bool inproj = is_cp && p.c_inproj_w != nullptr;
    const int tid4 = c.tid * 4;

    const int total = pass_ops(S.n_layers, inproj, ps.head);
    for (int it = 0; it < total && !c.aborted; ++it) {
        const FkOp op = pass_op(it, S.n_layers, inproj);
        const int kind = op.kind, l = op.layer;
        if (kind == FKT_B && is_cp) continue;                 // the predictor's attention lives inside phase C
        ++c.seq;
        const unsigned want = c.seq - 1;
        fk_phase(c, tk, kind);
        fk_mark(c, 0);
        const FkLayer& L = is_cp ? p.c_layers[l] : p.t_layers[l];
        const FkDesc d = FK_SH(c)->desc[tk][kind];
        const FkUnit u = FK_SH(c)->unit[tk][kind][c.warp];
        // RMSNorm weight of this phase: fetched BEFORE the grid hand-over (independent of the activations)
        const float* nw = (kind == FKT_A) ? L.ln1 : (kind == FKT_D) ? L.ln2 : (kind == FKT_HEAD) ? S.final_norm : nullptr;
        float4 nwv[HJ];
#pragma unroll
        for (int j = 0; j < HJ; ++j)
            nwv[j] = (nw && j * 1024 + tid4 < H) ? __ldg(reinterpret_cast<const float4*>(nw + j * 1024 + tid4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        // (Measured and rejected, -DFK_DATAFLOW_QKV: the predictor's phase C and the talker's phase B need only the q/k/v words of
        // their own kv group, each validated by its sequence tag, so they could skip the grid-wide wait. 2.375 instead of 2.275 ms
        // per frame: 4 warps x 120 CTAs re-polling the words delays the very stores they wait for.)
        // the layer input row: res0 (smem) for layer 0 without in_proj, else an LL buffer
        const bool in_res0 = (l == 0 && !inproj);
        const uint2* lin = (l == 0) ? p.cxin : S.x;           // (only read when !in_res0)
        // the vector this phase fetches first (requested by the poller as soon as the grid has arrived)
        const uint2* fsrc = (kind == FKT_A) ? (in_res0 ? nullptr : lin) : (kind == FKT_D) ? S.x1 : (kind == FKT_E) ? S.act : (kind == FKT_HEAD) ? S.x : nullptr;
        const int fw = (kind == FKT_E) ? min(FK_LAND_WORDS, d.K) : H;
        if (kind != FKT_B && !(kind == FKT_C && is_cp)) {      // (the attention phases request their cached K/V rows first and then wait themselves)
            grid_wait(c, want, fsrc, fw);
            fk_mark(c, 1);
        }
        if (kind == FKT_B) {
            const bool fence = p.kv_f32 ? talker_attn_partial<float>(c, L, l, ps.pos0, want) : talker_attn_partial<bf16_t>(c, L, l, ps.pos0, want);
            csync();
            if (c.tid == 0) grid_arrive(c);
            if (fence) __threadfence();                         // the new K/V row, before this CTA's next arrival (see talker_attn_partial)
            fk_mark(c, 3);
            if (FK_SH(c)->aborted) { c.aborted = true; break; }
            continue;
        }
        if (kind == FKT_C) {
            if (is_cp) cp_attn_local(c, L, l, ps.pos0, want, u, NST);
            else       talker_attn_combine(c, ps.pos0, want, u, NST);
            fk_mark(c, 3);
            if (FK_SH(c)->aborted) { c.aborted = true; break; }
            const int rpp = d.rpp;
            gemv_mma<NST>(c, u, smem_u32(FK_XS(c)));
            csync();                                           // every warp's partial sums are in shared memory
            fk_mark(c, 12);
            gemv_release(c, d, NST);
            // The n_kv CTAs that hold the partials of the same rows (one per kv group) sit in the same cluster: the partial of
            // row r goes straight into the shared memory of partner r / rpp as a (value, sequence) word (DSMEM store, no barrier)
            if (c.tid < d.nrows) {
                const unsigned g = c.rank % (unsigned)n_kv, base = c.rank - g;
                const unsigned pm = FK_SH(c)->pushmap[tk][c.tid];          // (partner << 8) | row slot at the partner: tid / rpp, tid % rpp
                st_ll_dsmem(dsmem_addr(&FK_SH(c)->redc[g][pm & 255u], base + (pm >> 8)), part_sum(c, c.tid, 1 << d.lw), c.seq);
            }
            fk_mark(c, 4);
            reduce_partials(c, d, n_kv, rpp, in_res0 ? FK_RES0(c) : nullptr, FK_LAND(c) + c.land_a * FK_LAND_WORDS, S.x1);
            fk_mark(c, 6);
            if (c.aborted || FK_SH(c)->aborted) { c.aborted = true; break; }
            continue;
        }
        // ---- K phases: stage the input vector as bf16x3 B fragments (x * w_norm for the normalised phases) ------
        float xin[HJ][4];                                      // the raw row (this thread's columns): sum of squares, last_hidden
#pragma unroll
        for (int j = 0; j < HJ; ++j) { xin[j][0] = 0.f; xin[j][1] = 0.f; xin[j][2] = 0.f; xin[j][3] = 0.f; }
        const uint32_t xf_s = smem_u32(FK_XP(c));
        switch (kind) {
            case FKT_INPROJ: {                                     // plain row in shared memory, no norm
                gemv_wait(c, u, NST);
#pragma unroll
                for (int j = 0; j < HJ; ++j)
                    if (j * 1024 + tid4 < d.K) {
                        const float4 v = *reinterpret_cast<const float4*>(FK_RES0(c) + j * 1024 + tid4);
                        stage_bfrag(xf_s, j * 1024 + tid4, v.x, v.y, v.z, v.w);
                    }
                break;
            }
            case FKT_A: {
                if (in_res0) {
                    gemv_wait(c, u, NST);
#pragma unroll
                    for (int j = 0; j < HJ; ++j)
                        if (j * 1024 + tid4 < H) f4_to(xin[j], *reinterpret_cast<const float4*>(FK_RES0(c) + j * 1024 + tid4));
                } else {
                    uint2* land = mc_fetch(c, lin, H, true, &u, NST);
                    land_row<HJ>(c, land, lin, H, want, xin);
                    c.land_a = (c.land_n - 1u) & 1u;
                }
                break;
            }
            case FKT_D: {
                uint2* land = mc_fetch(c, S.x1, H, true, &u, NST);
                land_row<HJ>(c, land, S.x1, H, want, xin);
                const FkDesc& de = FK_SH(c)->desc[tk][FKT_E];
#pragma unroll
                for (int j = 0; j < HJ; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {                          // residual for this CTA's down-projection rows
                        const unsigned rel = (unsigned)(j * 1024 + tid4 + i - de.row0);
                        if (rel < (unsigned)de.nrows) FK_SH(c)->x1own[rel] = xin[j][i];
                    }
                break;
            }
            case FKT_E: {
#pragma unroll 1
                for (int w0 = 0; w0 < d.K; w0 += FK_LAND_WORDS) {          // wide models: two landing-buffer loads
                    const int wn = min(FK_LAND_WORDS, d.K - w0);
                    float t3[3][4];
                    uint2* land = mc_fetch(c, S.act + w0, wn, w0 == 0, w0 == 0 ? &u : nullptr, NST);
                    land_row<3>(c, land, S.act + w0, wn, want, t3);
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (j * 1024 + tid4 < wn) stage_bfrag(xf_s, w0 + j * 1024 + tid4, t3[j][0], t3[j][1], t3[j][2], t3[j][3]);
                }
                break;
            }
            default: {  // FKT_HEAD: final norm + head
                uint2* land = mc_fetch(c, S.x, H, true, &u, NST);
                land_row<HJ>(c, land, S.x, H, want, xin);
                break;
            }
        }
        fk_mark(c, 2);
        if (nw) {                                              // RMSNorm, folded: product on x*w, 1/rms in the epilogue
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < HJ; ++j) {
                if (j * 1024 + tid4 < H) {
                    const float4 w = nwv[j];
                    ss = fmaf(xin[j][0], xin[j][0], ss); ss = fmaf(xin[j][1], xin[j][1], ss);
                    ss = fmaf(xin[j][2], xin[j][2], ss); ss = fmaf(xin[j][3], xin[j][3], ss);
                    stage_bfrag(xf_s, j * 1024 + tid4, xin[j][0] * w.x, xin[j][1] * w.y, xin[j][2] * w.z, xin[j][3] * w.w);
                }
            }
            ss_publish(c, ss);
        }
        if (c.aborted || FK_SH(c)->aborted) { c.aborted = true; }
        csync();                                               // the input fragments (and the sum-of-squares partials) are complete
        fk_mark(c, 8);
        if (FK_SH(c)->aborted) { c.aborted = true; break; }
        // ---- product (all warps, K split), then the epilogue on warp 0: one lane per row (pair) ----------------------
        {
            gemv_mma<NST>(c, u, xf_s);
            csync();                                           // every warp's partial sums are in shared memory
            fk_mark(c, 12);
            gemv_release(c, d, NST);
            const bool lh = (kind == FKT_HEAD && !is_cp);
            float rs = 1.f;
            if (nw && (c.warp == 0 || lh)) rs = ss_rstd(c, H, p.eps);
            if (c.warp == 0) {
                // (Measured and rejected: one straight-line epilogue for all kinds with two rows per lane and every load up front --
                // phase D 0.10 us faster, A and E 0.05-0.11 us slower, 1 % slower overall.)
                const int wpp = 1 << d.lw;
                if (kind == FKT_D) {                           // rows 2q (gate), 2q + 1 (up) -> act[q]
                    const int q = c.lane;
                    if (2 * q < d.nrows) {
                        const float gt = part_sum(c, 2 * q, wpp) * rs, up = part_sum(c, 2 * q + 1, wpp) * rs;
                        st_ll(S.act + (d.row0 >> 1) + q, silu_f(gt) * up, c.seq);
                    }
                } else {
#pragma unroll 1
                    for (int r = c.lane; r < d.nrows; r += 32) {
                        const float v = part_sum(c, r, wpp) * rs;
                        const int n = d.row0 + r;
                        if (kind == FKT_A) st_ll(S.qkv + n, v, c.seq);
                        else if (kind == FKT_E) st_ll(S.x + n, FK_SH(c)->x1own[r] + v, c.seq);
                        else if (kind == FKT_INPROJ) st_ll(p.cxin + n, v + __ldg(p.c_inproj_b + n), c.seq);
                        else {
                            st_ll((is_cp ? p.clogits_ll : p.logits_ll) + n, v, c.seq);
                            (is_cp ? p.clogits : p.logits)[n] = v;
                        }
                    }
                }
                __syncwarp();
                if (c.lane == 0) grid_arrive(c);               // the outputs of this CTA are issued
            }
            if (kind == FKT_HEAD && !is_cp) {                  // talker last_hidden = final-norm of the row (:859)
#pragma unroll
                for (int j = 0; j < HJ; ++j) {
                    if (j * 1024 + tid4 < H) {
                        const float4 w = nwv[j];
                        const float4 o = make_float4((xin[j][0] * rs) * w.x, (xin[j][1] * rs) * w.y, (xin[j][2] * rs) * w.z, (xin[j][3] * rs) * w.w);
                        *reinterpret_cast<float4*>(FK_LH(c) + j * 1024 + tid4) = o;
                        if (c.cta == 0) *reinterpret_cast<float4*>(p.last_hidden + j * 1024 + tid4) = o;
                    }
                }
            }
        }
        fk_mark(c, 6);
        if (c.aborted) break;
    }
}

// ------------------------------------------------------------------------------------------------
// sampler (same bit-exact semantics as sample_kernel in sampler.cuh), 256 consumer threads,
// executed redundantly by every CTA on the same logits
// ------------------------------------------------------------------------------------------------
struct FkSampScratch { float* x; float* pr; float* spr; unsigned short* idx; unsigned short* rank; };
// x: own scratch; the arrays of the general path reuse the landing buffers. Built where it is used so that the pointers stay
// shared-space addresses derived from the fk_smem symbol.
LQT_DEVINL FkSampScratch fk_scratch(const FkCtx& c) {
    FkSampScratch s;
    const FkSmemOffsets& so = c.p->so;
    s.x = reinterpret_cast<float*>(fk_smem + so.scratch);
    s.pr = reinterpret_cast<float*>(fk_smem + so.land);
    s.spr = s.pr + so.maxV;
    s.idx = reinterpret_cast<unsigned short*>(s.spr + so.maxV); s.rank = s.idx + so.maxV;
    return s;
}

LQT_DEVINL int block_excl_scan(FkCtx& c, int v, int* total) {          // 256-thread exclusive scan
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (c.lane >= o) inc += t; }
    csync();
    if (c.lane == 31) FK_SH(c)->wtot[c.warp] = inc;
    csync();
    int base = 0, tot = 0;
#pragma unroll
    for (int w2 = 0; w2 < FK_CWARPS; ++w2) { const int t = FK_SH(c)->wtot[w2]; if (w2 < c.warp) base += t; tot += t; }
    *total = tot;
    return base + inc - v;
}

// Fast path of the draw for 0 < top_k <= 64 (same results as the general path below, bit for bit):
// a 256-bin histogram of (max - x) * 16 locates the bin that holds the k-th largest value, the <= 64
// candidates up to that bin are compacted in index order, and ONE warp finishes: exact top-k by rank
// (ties survive, src/tts_onnx.cpp:917-927), softmax (:907-915), top-p (:929-950), renormalisation and the
// categorical draw in index order (:893-905), all sums serial in the reference's order.
// Returns -1 when the shape does not fit (flat logits): the caller falls back to the general path.
LQT_DEVINL int fk_sample_fast(FkCtx& c, int i0, int i1, float mx, const SamplingDev& sp,
                              uint32_t frame, int codebook) {
    FkShared* sh = FK_SH(c);
    const FkSampScratch s = fk_scratch(c);
    const int k = sp.top_k;
    sh->hist[c.tid] = 0;
    csync();
    for (int i = i0; i < i1; ++i) {
        const float dlt = (mx - s.x[i]) * 16.0f;               // >= 0; +inf for masked entries
        if (dlt < 255.0f) atomicAdd(&sh->hist[(int)dlt], 1);
    }
    csync();
    if (c.warp == 0) {
        int h[8], loc = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { h[q] = sh->hist[c.lane * 8 + q]; loc += h[q]; }
        int inc = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (c.lane >= o) inc += t; }
        int cum = inc - loc, myB = -1;
#pragma unroll
        for (int q = 0; q < 8; ++q) { cum += h[q]; if (myB < 0 && cum >= k) myB = c.lane * 8 + q; }
        const unsigned bal = __ballot_sync(0xffffffffu, myB >= 0);
        const int B = bal ? __shfl_sync(0xffffffffu, myB, __ffs(bal) - 1) : -1;
        if (c.lane == 0) sh->sel_k = B;
    }
    csync();
    fk_mark(c, 10);
    const int B = sh->sel_k;
    if (B < 0) { csync(); return -1; }
    const float lim = (float)(B + 1);
    int cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += ((mx - s.x[i]) * 16.0f < lim) ? 1 : 0;
    int n_c;
    int wpos = block_excl_scan(c, cnt, &n_c);
    if (n_c > 64) { csync(); return -1; }
    for (int i = i0; i < i1; ++i) {
        const float v = s.x[i];
        if ((mx - v) * 16.0f < lim) { s.idx[wpos] = (unsigned short)i; s.pr[wpos] = v; ++wpos; }
    }
    csync();
    fk_mark(c, 11);
    if (c.warp == 0) {
        // One warp, lane l owns candidates l and l + 32 (candidates are in index order). Everything stays in candidate order:
        // a candidate that drops out (below the k-th value, beyond the top-p cut) becomes an exact 0.0f, which is neutral in
        // every serial sum, so the sums see the reference's operands in the reference's order without any compaction.
        // Arrays are padded to 64 entries and read eight at a time (two 16-byte loads ahead of each dependent chain); no
        // per-entry bounds checks anywhere. The lane index goes through an opaque asm: ptxas would otherwise re-read
        // SR_TID.X (a long-latency S2R) in front of every comparison of this single-warp section.
        int lane;
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
        const int e0 = lane, e1 = lane + 32;
        const int n8 = (n_c + 7) & ~7;
        const float x0 = (e0 < n_c) ? s.pr[e0] : -INFINITY, x1 = (e1 < n_c) ? s.pr[e1] : -INFINITY;
        __syncwarp();
        if (e0 >= n_c) s.pr[e0] = -INFINITY;
        if (e1 >= n_c) s.pr[e1] = -INFINITY;
        __syncwarp();
        int gt0 = 0, gt1 = 0;
#pragma unroll 1
        for (int j = 0; j < n8; j += 8) {
            const float4 u = *reinterpret_cast<const float4*>(s.pr + j), w = *reinterpret_cast<const float4*>(s.pr + j + 4);
            const float xv[8] = {u.x, u.y, u.z, u.w, w.x, w.y, w.z, w.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) { gt0 += (xv[q] > x0) ? 1 : 0; gt1 += (xv[q] > x1) ? 1 : 0; }
        }
        fk_mark(c, 12);
        const bool sv0 = e0 < n_c && gt0 < k, sv1 = e1 < n_c && gt1 < k;      // x >= (k-th largest)  <=>  fewer than k values above it
        const unsigned b0 = __ballot_sync(0xffffffffu, sv0), b1 = __ballot_sync(0xffffffffu, sv1);
        const int ns = __popc(b0) + __popc(b1);
        float pr0 = sv0 ? (float)exp((double)(x0 - mx)) : 0.f, pr1 = sv1 ? (float)exp((double)(x1 - mx)) : 0.f;
        s.spr[e0] = pr0; s.spr[e1] = pr1;
        __syncwarp();
        float sum = 0.f;
#pragma unroll 1
        for (int i = 0; i < n8; i += 8) {                                        // serial, index order (every lane redundantly)
            const float4 u = *reinterpret_cast<const float4*>(s.spr + i), w = *reinterpret_cast<const float4*>(s.spr + i + 4);
            sum += u.x; sum += u.y; sum += u.z; sum += u.w; sum += w.x; sum += w.y; sum += w.z; sum += w.w;
        }
        fk_mark(c, 13);
        pr0 = pr0 / sum; pr1 = pr1 / sum;                                        // (dropped candidates stay 0)
        __syncwarp();
        s.spr[e0] = pr0; s.spr[e1] = pr1;
        __syncwarp();
        if (sp.top_p < 1.0f) {
            int r0 = 0, r1 = 0;                                                  // position in descending order (ties: index order)
#pragma unroll 1
            for (int j = 0; j < n8; j += 8) {
                const float4 u = *reinterpret_cast<const float4*>(s.spr + j), w = *reinterpret_cast<const float4*>(s.spr + j + 4);
                const float pv[8] = {u.x, u.y, u.z, u.w, w.x, w.y, w.z, w.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    r0 += (pv[q] > pr0 || (pv[q] == pr0 && j + q < e0)) ? 1 : 0;
                    r1 += (pv[q] > pr1 || (pv[q] == pr1 && j + q < e1)) ? 1 : 0;
                }
            }
            s.pr[e0] = 0.f; s.pr[e1] = 0.f;
            __syncwarp();
            if (sv0) s.pr[r0] = pr0;                                             // survivors only: their positions are 0..ns-1
            if (sv1) s.pr[r1] = pr1;
            __syncwarp();
            int cut = ns;
            float cs = 0.f;
#pragma unroll 1
            for (int r = 0; r < ns && cut == ns; r += 8) {
                const float4 u = *reinterpret_cast<const float4*>(s.pr + r), w = *reinterpret_cast<const float4*>(s.pr + r + 4);
                const float pv[8] = {u.x, u.y, u.z, u.w, w.x, w.y, w.z, w.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    cs += pv[q];                                                 // (zeros beyond ns never move cs across the threshold)
                    if (cut == ns && cs > sp.top_p) cut = min(ns, r + q + 1);
                }
            }
            if (r0 >= cut) pr0 = 0.f;
            if (r1 >= cut) pr1 = 0.f;
            s.spr[e0] = pr0; s.spr[e1] = pr1;
            __syncwarp();
            float s2 = 0.f;
#pragma unroll 1
            for (int i = 0; i < n8; i += 8) {
                const float4 u = *reinterpret_cast<const float4*>(s.spr + i), w = *reinterpret_cast<const float4*>(s.spr + i + 4);
                s2 += u.x; s2 += u.y; s2 += u.z; s2 += u.w; s2 += w.x; s2 += w.y; s2 += w.z; s2 += w.w;
            }
            if (s2 > 0.f) { pr0 = pr0 / s2; pr1 = pr1 / s2; }
            __syncwarp();
            s.spr[e0] = pr0; s.spr[e1] = pr1;
            __syncwarp();
        }
        fk_mark(c, 14);
        uint32_t r4[4];
        philox4x32_10(frame, (uint32_t)codebook, 0u, 0u, sp.seed, sp.utt, r4);
        const float u01 = (float)(r4[0] >> 8) * 5.9604644775390625e-08f;
        const int first = b0 ? (__ffs(b0) - 1) : (b1 ? 32 + __ffs(b1) - 1 : 0);
        float cdf = 0.f; int last = (int)s.idx[first]; bool hit = false;
#pragma unroll 1
        for (int i = 0; i < n8 && !hit; i += 8) {
            const float4 u = *reinterpret_cast<const float4*>(s.spr + i), w = *reinterpret_cast<const float4*>(s.spr + i + 4);
            const float pv[8] = {u.x, u.y, u.z, u.w, w.x, w.y, w.z, w.w};
            const uint4 ri = *reinterpret_cast<const uint4*>(s.idx + i);       // 8 ushort token indices
            const unsigned rw[4] = {ri.x, ri.y, ri.z, ri.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (!hit && pv[q] > 0.f) {
                    cdf += pv[q];
                    last = (int)((rw[q >> 1] >> ((q & 1) * 16)) & 0xffffu);
                    if (cdf > u01) hit = true;
                }
            }
        }
        fk_mark(c, 15);
        if (lane == 0) sh->tok = last;
    }
    csync();
    const int tok = sh->tok;
    csync();                     // everyone has read tok/scratch before the glue overwrites anything
    return tok;
}

// logits: LL words tagged `want` (ll != nullptr; already multicast into `land`) or a plain array (first draw after a resume)
LQT_DEVINL int fk_sample(FkCtx& c, const uint2* ll, const uint2* land, const float* plain, unsigned want, int V,
                         int mask_lo, int mask_hi, int mask_keep, const SamplingDev& sp, uint32_t frame, int codebook,
                         float* trace_row) {
    FkShared* sh = FK_SH(c);
    const FkSampScratch s = fk_scratch(c);
    const bool temper = !sp.greedy && sp.temperature > 0.0f && sp.temperature != 1.0f;
    // thread owns the contiguous range [i0, i1) (V % 4 == 0, ranges are multiples of 4)
    const int per = (((V + FK_CTHREADS - 1) / FK_CTHREADS) + 3) & ~3;
    const int i0 = min(V, c.tid * per), i1 = min(V, i0 + per);
    float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll 1
    for (int i = i0; i < i1; i += 4) {                         // rolled: this code runs once per draw, cold in the instruction cache
        {
            float4 q;
            if (ll) {                                          // landed (value, sequence) words, validated like every other input
                uint4 a = *reinterpret_cast<const uint4*>(land + i), b = *reinterpret_cast<const uint4*>(land + i + 2);
                if (a.y != want || a.w != want || b.y != want || b.w != want) {
                    const FkLL4 r = ll_poll4_slow(ll + i, want, &FK_SH(c)->aborted, c.p->ctrl);
                    a = r.a; b = r.b;
                }
                q = make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z));
            } else {
                q = *reinterpret_cast<const float4*>(plain + i);
            }
            const float vv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float v = vv[u];
                const int ii = i + u;
                if (ii >= mask_lo && ii < mask_hi && ii != mask_keep) v = -INFINITY;
                if (trace_row) trace_row[ii] = v;
                if (temper) v = v / sp.temperature;
                s.x[ii] = v;
                if (v > bv) { bv = v; bi = ii; }                   // ascending ii: first maximum wins
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    fk_mark(c, 8);
    if (c.lane == 0) { sh->redf[c.warp][0] = bv; sh->redi[c.warp] = bi; }
    csync();
    {
        float v = sh->redf[0][0]; int i = sh->redi[0];
#pragma unroll
        for (int w2 = 1; w2 < FK_CWARPS; ++w2)
            if (sh->redf[w2][0] > v || (sh->redf[w2][0] == v && sh->redi[w2] < i)) { v = sh->redf[w2][0]; i = sh->redi[w2]; }
        bv = v; bi = (i == 0x7fffffff) ? 0 : i;
    }
    if (sp.greedy) { csync(); return bi; }
    const float mx = bv;
    fk_mark(c, 9);
    if (sp.top_k > 0 && sp.top_k < V && sp.top_k <= 64) {      // common case: a handful of survivors, finished by one warp
        const int t = fk_sample_fast(c, i0, i1, mx, sp, frame, codebook);
        if (t >= 0) return t;
    }

    // top-k threshold: 4 x 8-bit radix select of the k-th largest key (warp-aggregated histogram)
    float thr = -INFINITY;
    if (sp.top_k > 0 && sp.top_k < V) {
        if (c.tid == 0) { sh->sel_prefix = 0u; sh->sel_k = sp.top_k; }
        for (int shift = 24; shift >= 0; shift -= 8) {
            sh->hist[c.tid] = 0;
            csync();
            const uint32_t prefix = sh->sel_prefix;
            const int kk = sh->sel_k;
            const uint32_t himask = (shift == 24) ? 0u : (0xffffffffu << (shift + 8));
            for (int i = i0; i < i0 + per; ++i) {               // uniform trip count: __match_any needs the full warp
                const bool in = i < i1;
                const uint32_t key = in ? float_key(s.x[i]) : 0u;
                const bool hit = in && ((key & himask) == prefix);
                const int bin = hit ? (int)((key >> shift) & 255) : (256 + c.lane);   // misses never match each other
                const unsigned peers = __match_any_sync(0xffffffffu, bin);
                if (hit && (__ffs(peers) - 1) == c.lane) atomicAdd(&sh->hist[bin], __popc(peers));
            }
            csync();
            // suffix sums S(b) = sum_{b' >= b} hist[b']; pick b with S(b) >= kk > S(b+1)
            const int hv = sh->hist[c.tid];
            int tot;
            const int excl = block_excl_scan(c, hv, &tot);         // sum of bins below tid
            const int s_ge = tot - excl;                           // S(tid)
            const int s_gt = s_ge - hv;                            // S(tid+1)
            if (s_ge >= kk && s_gt < kk) { sh->sel_prefix = prefix | ((uint32_t)c.tid << shift); sh->sel_k = kk - s_gt; }
            csync();
        }
        const uint32_t kkey = sh->sel_prefix;
        thr = __uint_as_float((kkey & 0x80000000u) ? (kkey & 0x7fffffffu) : ~kkey);
    }
    // compaction in index order (thread owns a contiguous range)
    int cnt = 0;
    for (int i = i0; i < i1; ++i) { const float v = s.x[i]; cnt += (!(v < thr) && v != -INFINITY) ? 1 : 0; }
    int n_surv;
    int wpos = block_excl_scan(c, cnt, &n_surv);
    for (int i = i0; i < i1; ++i) {
        const float v = s.x[i];
        if (!(v < thr) && v != -INFINITY) {
            s.idx[wpos] = (unsigned short)i;
            s.pr[wpos] = (float)exp((double)(v - mx));
            ++wpos;
        }
    }
    csync();
    if (c.tid == 0) {
        float sum = 0.f;
        for (int i = 0; i < n_surv; ++i) sum += s.pr[i];
        sh->fsum = sum;
    }
    csync();
    {
        const float sum = sh->fsum;
        for (int i = c.tid; i < n_surv; i += FK_CTHREADS) s.pr[i] = s.pr[i] / sum;
    }
    csync();
    const bool use_top_p = sp.top_p < 1.0f;
    if (use_top_p) {
        for (int i = c.tid; i < n_surv; i += FK_CTHREADS) {
            const float pi = s.pr[i];
            int r = 0;
            for (int j = 0; j < n_surv; ++j) { const float pj = s.pr[j]; r += (pj > pi || (pj == pi && j < i)) ? 1 : 0; }
            s.rank[i] = (unsigned short)r;
            s.spr[r] = pi;
        }
        csync();
    }
    if (c.tid == 0) {
        int cut = n_surv;
        float s2 = 1.0f;
        if (use_top_p) {
            float cs = 0.f;
            for (int r = 0; r < n_surv; ++r) { cs += s.spr[r]; if (cs > sp.top_p) { cut = r + 1; break; } }
            s2 = 0.f;
            for (int i = 0; i < n_surv; ++i) if ((int)s.rank[i] < cut && s.pr[i] > 0.f) s2 += s.pr[i];
        }
        uint32_t r4[4];
        philox4x32_10(frame, (uint32_t)codebook, 0u, 0u, sp.seed, sp.utt, r4);
        const float u = (float)(r4[0] >> 8) * 5.9604644775390625e-08f;
        float cdf = 0.f; int last = (n_surv > 0) ? (int)s.idx[0] : 0;
        for (int i = 0; i < n_surv; ++i) {
            float pi = s.pr[i];
            if (use_top_p) {
                if ((int)s.rank[i] >= cut) continue;
                if (s2 > 0.f) pi = pi / s2;
            }
            if (pi > 0.f) {
                cdf += pi; last = (int)s.idx[i];
                if (cdf > u) break;
            }
        }
        sh->tok = last;
    }
    csync();
    const int tok = sh->tok;
    csync();                     // everyone has read tok/scratch before the glue overwrites anything
    return tok;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
struct FkSmemLayout { size_t land, scratch, att, xs, red, nxt, res0, lh, shared, total; };
inline FkSmemLayout fk_smem_layout(int nstages, int maxV, int H, int maxK) {
    FkSmemLayout L{};
    auto up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
    size_t off = (size_t)nstages * FK_STAGE_BYTES;
    L.land = off; off += (size_t)2 * FK_LAND_WORDS * 8;        // multicast landing buffers | sampler arrays pr/spr/idx/rank (general path)
    L.scratch = off;                                           // sampler logits x[V] | attention scratch + attention output row
    const size_t attb = up((size_t)FA_FLOATS * 4), xsb = (size_t)2 * FK_XS_STRIDE * 4;
    L.att = off; L.xs = off + attb;
    size_t sc = attb + xsb;
    if ((size_t)maxV * 4 > sc) sc = (size_t)maxV * 4;
    if ((size_t)maxK * 6 > sc) sc = (size_t)maxK * 6;       // input vector of a matrix-vector phase as bf16x3 B fragments (96 bytes per 16 columns)
    off += up(sc);
    L.red = off;
    L.nxt = off; off += up((size_t)H * 4);
    L.res0 = off; off += up((size_t)H * 4);
    L.lh = off; off += up((size_t)H * 4);
    L.shared = off; off += up(sizeof(FkShared));
    L.total = off;
    return L;
}

// KJ = 1024-column chunks of the widest matrix (3: inter <= 3072, 6: <= 6144); NST = ring stages
template <int KJ, int NST>
__global__ void __launch_bounds__(FK_THREADS, 1)
frame_kernel(const __grid_constant__ FkParams p) {
    FkShared* sh = reinterpret_cast<FkShared*>(fk_smem + p.so.shared);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&sh->full[i], 1); mbar_init(&sh->empty[i], 1); }   // one consumer thread releases a stage (gemv_release)
        mbar_init(&sh->land_bar[0], 1); mbar_init(&sh->land_bar[1], 1);
        sh->stop = 0; sh->consumed = 0; sh->aborted = 0; sh->zero8[0] = 0u; sh->zero8[1] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 160) {
        const FkDesc dd = make_desc(p, tid >= 80, (tid % 80) >> 3, cta, ncta);
        if ((tid & 7) == 0) sh->desc[tid / 80][(tid % 80) >> 3] = dd;
        sh->unit[tid / 80][(tid % 80) >> 3][tid & 7] = make_unit(dd, tid & 7);
    }
    for (int i = tid; i < FK_NGRP_MAX * FK_RPP_MAX; i += FK_THREADS) (&sh->redc[0][0])[i] = make_uint2(0u, 0u);
    for (int i = tid; i < min(p.n_pages, FK_PT_MAX); i += FK_THREADS) sh->pt[i] = p.page_table[i];
    __syncthreads();
    if (tid < 2 * FK_PART_ROWS) {
        const int st = tid / FK_PART_ROWS, r = tid % FK_PART_ROWS, rpp = max(1, sh->desc[st][FKT_C].rpp);
        sh->pushmap[st][r] = (uint16_t)(((r / rpp) << 8) | (r % rpp));
    }
    __syncthreads();
    cluster_sync_all();                            // every landing barrier of the cluster is initialised before any multicast can arrive

    const GenState st0 = *p.st;                    // written by the host before launch
    const int n_prefill = (p.mode == 0 && st0.pos == 0) ? p.P : 0;
    const int frame_end = min(p.frame_end, st0.max_frames);

    if (warp >= FK_CWARPS) {
#ifndef FK_NO_SETMAXNREG
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
#endif
        // ============================ producer warp ============================================
        // Walks the same flat schedule as the consumers, one stage at a time, as far ahead as the
        // ring allows. All state in registers of one lane.
        if (warp == FK_CWARPS + FK_PRODUCER_WARP && lane == 0) {
            long long n_pass;
            if (p.mode == 1) n_pass = 1;
            else {
                const int nf = (!st0.done && frame_end > st0.frame) ? (frame_end - st0.frame) : 0;
                n_pass = (long long)n_prefill + (long long)nf * (p.cp_steps + 2);
            }
            unsigned issued = 0;
            bool stopped = false;
            for (long long q = 0; q < n_pass && !stopped; ++q) {
                const FkPassId id = launch_pass(q, p.mode, n_prefill, p.cp_steps);
                const FkStack& S = id.is_cp ? p.cp : p.talker;
                const bool inproj = id.is_cp && p.c_inproj_w != nullptr;
                const int total = pass_ops(S.n_layers, inproj, id.head);
                for (int it = 0; it < total && !stopped; ++it) {
                    const FkOp op = pass_op(it, S.n_layers, inproj);
                    if (op.kind == FKT_B) continue;
                    const FkLayer& L = id.is_cp ? p.c_layers[op.layer] : p.t_layers[op.layer];
                    const FkDesc d = sh->desc[id.is_cp ? 1 : 0][op.kind];
                    const bf16_t* W;
                    switch (op.kind) {
                        case FKT_INPROJ: W = p.c_inproj_w; break;
                        case FKT_A: W = L.wqkv; break;
                        case FKT_C: W = L.wo_g; break;
                        case FKT_D: W = L.wgu; break;
                        case FKT_E: W = L.wdown; break;
                        default: W = id.is_cp ? p.c_heads + (size_t)id.cb * p.c_head_stride : p.t_head; break;
                    }
                    if (d.nrows <= 0) continue;
                    const uint32_t row_bytes = (uint32_t)d.K * 2u;
                    const char* srcb = reinterpret_cast<const char*>(W + d.img_off);
                    // the fragment-ordered image of the slice is staged as a flat byte stream in full stages
                    const uint32_t total = (uint32_t)d.nrows * row_bytes;
                    const uint32_t sbytes = (uint32_t)FK_STAGE_BYTES;
                    for (uint32_t off = 0; off < total && !stopped; off += sbytes) {
                        const unsigned slot = issued % (unsigned)NST, par = ((issued / (unsigned)NST) & 1u) ^ 1u;
                        unsigned long long t0 = 0;
                        while (!mbar_try_wait(&sh->empty[slot], par)) {
                            // The ring is full: the producer is ~10 us ahead of the consumers (profiles/r1_v19_ring_latency.txt), so it
                            // sleeps between polls instead of spinning -- it shares its scheduler with consumer warps 0 and 4, and
                            // warp 0 (grid hand-over, reductions, sampler tail) is the critical path of every phase.
                            if (p.producer_sleep_ns) __nanosleep(p.producer_sleep_ns);
                            if (sh->stop) { stopped = true; break; }
                            if (t0 == 0) t0 = clock64();
                            else if (clock64() - t0 > FK_SPIN_LIMIT) { stopped = true; break; }
                        }
                        if (stopped) break;
                        const uint32_t bytes = min(sbytes, total - off);
#ifdef FK_FINE_MARKS
                        if (p.dbg && cta == p.dbg_cta) {
                            unsigned long long* pd = p.dbg + p.dbg_cap / 2;
                            if ((int)issued + 2 < p.dbg_cap / 2) { pd[1 + issued] = ((unsigned long long)clock64() << 16) | (issued & 0xffffu); pd[0] = issued + 1; }
                        }
#endif
                        mbar_expect_tx(&sh->full[slot], bytes);
                        bulk_g2s(fk_smem + (size_t)slot * FK_STAGE_BYTES, srcb + off, bytes, &sh->full[slot]);
                        ++issued;
                    }
                }
            }
            // drain: every issued copy must land before the CTA may exit
            unsigned long long t0 = clock64();
            while (!sh->stop) { if (clock64() - t0 > 4 * FK_SPIN_LIMIT) break; __nanosleep(200); }
            __threadfence_block();
            for (unsigned stg = (unsigned)sh->consumed; stg < issued; ++stg) {
                const unsigned slot = stg % (unsigned)NST, par = (stg / (unsigned)NST) & 1u;
                unsigned long long t1 = clock64();
                while (!mbar_try_wait(&sh->full[slot], par)) { if (clock64() - t1 > FK_SPIN_LIMIT) break; }
            }
        }
        __syncwarp();
        cluster_sync_all();                        // no CTA leaves while a partner may still write into its shared memory
        return;
    }

    // ================================ consumer warps ===============================================
#ifndef FK_NO_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
#endif
    FkCtx c;
    c.p = &p;
    c.land_n = 0; c.land_a = 0;
    c.rank = (unsigned)cta % FK_CLUSTER;
    c.tid = tid; c.lane = lane; c.warp = warp; c.cta = cta; c.ncta = ncta;
    c.seq = 0; c.stage_ctr = 0; c.aborted = false;
    c.dbg = (p.dbg && cta == p.dbg_cta) ? p.dbg + 1 : nullptr; c.dbg_n = 0; c.dbg_cap = p.dbg_cap / 2 - 1; c.dbg_tag = 0;
#ifdef FK_MARKS
    if (c.dbg && warp == FK_MARK_W2) c.dbg += p.dbg_cap / 2;
#endif
    const int H = p.talker.H, H4 = H >> 2;
    int pos = st0.pos, frame = st0.frame, done = st0.done, n_frames = st0.n_frames;
    const SamplingDev sp = *p.sp;
    int prefill_i = 0;
    int cb = 0;                      // next codebook to draw in the current frame (0 = talker code)
    bool mode1_done = false;
    bool row1_next = false;          // the next iteration runs predictor position 1 (no draw in between)
    bool resumed = (p.mode == 0 && n_prefill == 0);   // first draw reads the plain logits / last_hidden of the previous launch
    if (resumed) {
        for (int k4 = tid; k4 < H4; k4 += FK_CTHREADS)
            reinterpret_cast<float4*>(FK_LH(c))[k4] = __ldcg(reinterpret_cast<const float4*>(p.last_hidden) + k4);
        csync();
    }

    // One loop, one pass per iteration: [draw + glue ->] token pass. (src/tts_onnx.cpp:794, 801-846)
    while (!c.aborted) {
        FkPass ps;
        if (p.mode == 1 || prefill_i < n_prefill) {
            if (p.mode == 1 && mode1_done) break;
            const float* src = (p.mode == 1) ? p.next_in : p.prompt + (size_t)prefill_i * H;
            for (int k4 = tid; k4 < H4; k4 += FK_CTHREADS)
                reinterpret_cast<float4*>(FK_RES0(c))[k4] = __ldcg(reinterpret_cast<const float4*>(src) + k4);
            csync();
            const bool head = (p.mode == 1) || (prefill_i == n_prefill - 1);
            ps = FkPass{false, pos, head};
        } else if (row1_next) {
            for (int k4 = tid; k4 < H4; k4 += FK_CTHREADS)                      // predictor position 1: codec_embed(code0) (= the running sum so far)
                reinterpret_cast<float4*>(FK_RES0(c))[k4] = reinterpret_cast<const float4*>(FK_NXT(c))[k4];
            csync();
            ps = FkPass{true, 1, true};
            row1_next = false;
        } else {
            if (done || frame >= frame_end) break;
            // ---- draw codebook cb of this frame (:803-812 for cb 0, :863-864 otherwise) -----------------
            const int tk = cb ? 1 : 0;
            fk_phase(c, tk, FKT_SAMPLE);
            fk_mark(c, 0);
            const bool t0 = (cb == 0);                          // one call site: the sampler is instantiated once
            const uint2* lg = t0 ? (resumed ? nullptr : p.logits_ll) : p.clogits_ll;
            const int Vd = t0 ? p.vocab : p.cp_vocab;
            grid_wait(c, resumed ? 0u : c.seq, lg, Vd);         // the logits of the head phase have been issued everywhere
            fk_mark(c, 1);
            float* tr = (p.trace && cta == 0) ? p.trace + ((size_t)frame * 16 + cb) * p.trace_stride : nullptr;
            const uint2* land = lg ? mc_fetch(c, lg, Vd, true) : nullptr;
            int tok = fk_sample(c, lg, land, t0 ? p.logits : nullptr, c.seq, Vd, t0 ? 2048 : 0, t0 ? p.vocab : 0, t0 ? 2150 : -1,
                                sp, (uint32_t)frame, cb, tr);
            resumed = false;
            if (FK_SH(c)->aborted) { c.aborted = true; break; }
            if (p.forced && frame < st0.n_forced) tok = (int)p.forced[(size_t)frame * 16 + cb];
            fk_mark(c, 3);
            if (cb == 0 && tok == 2150) { done = 1; break; }                        // CODEC_EOS (:812)
            if (cta == 0 && tid == 0) {
                p.codes_out[(size_t)frame * 16 + cb] = tok;                         // :818-821
                if (cb == p.cp_steps && p.progress) {                               // the frame's 16 codes are stored: tell the host
                    __threadfence_system();
                    *p.progress = frame + 1;
                }
            }
            // ---- glue: embedding of the drawn code, running 16-way sum, next input rows (in res0) --------
            const bool last_cb = (cb == p.cp_steps);
            const bool use_tr = frame < st0.trailing_len;
            const bf16_t* row = (cb == 0) ? p.codec_embed + (size_t)tok * H
                                          : p.cp_embed + ((size_t)(cb - 1) * p.cp_vocab + tok) * H;
            for (int k4 = tid; k4 < H4; k4 += FK_CTHREADS) {
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(row) + k4);
                const float4 e = make_float4(bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y));
                float4 acc = e;                                                 // :824
                if (cb != 0) {                                                  // :825-830
                    acc = reinterpret_cast<float4*>(FK_NXT(c))[k4];
                    acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
                }
                if (last_cb) {                                                  // :833-842
                    const float4 tt = use_tr ? __ldg(reinterpret_cast<const float4*>(p.trailing + (size_t)frame * H) + k4)
                                             : __ldg(reinterpret_cast<const float4*>(p.tts_pad) + k4);
                    acc.x += tt.x; acc.y += tt.y; acc.z += tt.z; acc.w += tt.w;
                }
                reinterpret_cast<float4*>(FK_NXT(c))[k4] = acc;
                if (cb == 0) {                                                  // rows [last_hidden, codec_embed(code0)] (:854-860): two single-row passes
                    reinterpret_cast<float4*>(FK_RES0(c))[k4] = reinterpret_cast<const float4*>(FK_LH(c))[k4];
                } else {
                    reinterpret_cast<float4*>(FK_RES0(c))[k4] = last_cb ? acc : e;  // :867-868 / :845
                }
                if (last_cb && cta == 0) reinterpret_cast<float4*>(p.next_in)[k4] = acc;
            }
            csync();
            fk_mark(c, 4);
            if (last_cb) {
                n_frames = frame + 1;
                ps = FkPass{false, pos, true};                                  // :845
            } else {
                if (cb == 0) { ps = FkPass{true, 0, false}; row1_next = true; }  // predictor position 0: talker last_hidden, no head
                else ps = FkPass{true, cb + 1, true};
            }
        }
        consume_token<KJ, NST>(c, p, ps);
        if (c.aborted) break;
        if (p.mode == 1) { pos += 1; mode1_done = true; }
        else if (prefill_i < n_prefill) { pos += 1; ++prefill_i; }
        else if (row1_next) { }                                                 // position 0 done, position 1 follows without a draw
        else if (cb == p.cp_steps) { pos += 1; frame += 1; cb = 0; }
        else ++cb;
    }
    if (p.mode == 0 && !done && frame >= st0.max_frames) done = 1;
    // ---- exit: publish state, stop the producer -------------------------------------------------------
    csync();
#ifdef FK_MARKS
    if (c.dbg && warp == FK_MARK_W2 && lane == 0) c.dbg[-1] = (unsigned long long)c.dbg_n;
#endif
    if (tid == 0) {
        if (cta == 0) {
            p.st->pos = pos; p.st->frame = frame; p.st->done = done; p.st->n_frames = n_frames;
        }
        if (c.dbg) c.dbg[-1] = (unsigned long long)c.dbg_n;
        sh->consumed = (int)c.stage_ctr;
        __threadfence_block();
        sh->stop = 1;
    }
    cluster_sync_all();
}

}  // namespace lqt