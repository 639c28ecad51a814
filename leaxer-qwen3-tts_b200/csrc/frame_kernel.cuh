// Persistent frame kernel: loops A and B of the reference (src/tts_onnx.cpp:782-872) as ONE
// cooperative launch per utterance (or per chunk of frames).
//
// Why one kernel: a frame is 31 dependent network passes (1 talker step + 15 predictor passes, each
// followed by a draw) = ~460 dependent matrix-vector phases of ~1 us each. As separate launches the
// HBM pipe drains at every kernel boundary (round-1 v1: 577 launches, 4.2 ms per frame).
// Here every SM keeps one CTA resident:
//   * warp 8 (producer) walks the static weight schedule of the whole launch and streams this CTA's
//     rows of every matrix into a shared-memory ring with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), running AHEAD of the math across phase boundaries.
//     The weights are regrouped once at load time into per-CTA row-major images, so one stage of the
//     ring = a few complete rows = one contiguous copy;
//   * warps 0-7 (consumers) run the phases. A matrix-vector product is K-split over the 256 threads:
//     thread t owns columns 4t..4t+3 of every 1024-column chunk, so its slice of the input vector lives
//     in REGISTERS (polled straight from L2, no shared-memory staging, no barrier), the weights are read
//     with conflict-free 8-byte shared loads, eight rows are reduced at a time with a transposing
//     butterfly (9 shuffles for 8 rows) and the eight warp partials meet in shared memory: ONE CTA
//     barrier per phase. RMSNorm is folded: the product runs on x*w_norm while the sum of squares
//     travels with the partials, and 1/rms is applied in the epilogue (exact up to rounding order).
//     (Round-1 v10 ran this on tcgen05 with M64N8K16 MMAs: K/16 dependent MMAs of ~46 cycles per phase
//     plus bf16x3 operand staging made a phase 8-10 us; the FMA form is bounded by the shared-memory
//     read of the weights, ~0.25 us per phase.);
//   * there is NO grid barrier. Activations travel between CTAs as 8-byte (value, sequence) pairs
//     written with one store each ("LL" exchange, as in NCCL's low-latency protocol): a reader polls
//     the data itself until every word carries the sequence number of the phase that produces it.
//     One L2 write + one L2 read per hop instead of fence + atomic + poll + fence.
//     Every CTA executes the same numbered phases, so the expected number is always "previous phase";
//     a CTA can run at most one phase ahead of the slowest one, which makes one buffer per phase
//     kind race-free;
//   * the sampler and the embedding glue (src/tts_onnx.cpp:803-842, 854-868, 878-950) run redundantly
//     in every CTA, so a draw costs no broadcast.
// Phases per layer: QKV | [talker: split-KV attention] | O-projection by kv-group (the code predictor
// computes its <=17-position attention inside this phase) | gate/up (SwiGLU) | down.
#pragma once
#include "attention.cuh"
#include "common.cuh"
#include "sampler.cuh"

namespace lqt {

typedef __nv_bfloat16 bf16_t;

constexpr int FK_CWARPS = 8;
constexpr int FK_CTHREADS = FK_CWARPS * 32;       // consumer threads
constexpr int FK_THREADS = FK_CTHREADS + 32;      // + one producer warp
constexpr int FK_STAGE_BYTES = 24 * 1024;         // 8 rows of K=1024 (16 KB used), 4 rows of K=3072, 48 rows of K=256
constexpr int FK_NS_MAX = 24;                     // max attention splits per kv group
constexpr int FK_NGRP_MAX = 8;                    // kv groups
constexpr int FK_CP_POS = 32;                     // code-predictor KV capacity (positions)
constexpr int FK_RED_STRIDE = 96;                 // warp-partial row sums: [warp][FK_RED_STRIDE]; M = 2 -> second row at +48
constexpr int FK_X1OWN = 16;                      // max rows of the down projection per CTA
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
    uint2 *x, *qkv, *po, *act;            // LL buffers: [2][H] [2][qkv_dim] [2][n_kv][H] [2][inter]
};

struct FkParams {
    FkStack talker, cp;
    FkLayer t_layers[FK_MAX_TLAYERS];     // in the kernel-parameter constant bank: no load latency
    FkLayer c_layers[FK_MAX_CLAYERS];
    const bf16_t* t_head; int vocab;          // images
    const bf16_t* c_heads; int cp_vocab, cp_steps; long long c_head_stride;   // image elements per predictor head
    const bf16_t* c_inproj_w; const float* c_inproj_b; uint2* cxin;     // 1.7B: talker width -> predictor width (LL [2][Hc])
    float eps;
    void* kv_pool; const int* page_table; int page_shift; long long page_stride; int kv_f32;
    uint2* pa;                 // talker attention partials, LL [n_kv][FK_NS_MAX][2][ATT_PSTRIDE]
    float* cp_kv;              // [layer][k|v][n_kv][FK_CP_POS][128] fp32
    uint2 *logits_ll, *clogits_ll;
    float *logits, *clogits, *last_hidden, *next_in;   // plain copies: API outputs, resume across launches, mode-1 input
    const bf16_t *codec_embed, *cp_embed;
    const float* prompt; int P;           // prefill rows (run when st->pos == 0)
    const float *trailing, *tts_pad;
    GenState* st; const SamplingDev* sp;
    long long* codes_out; const long long* forced; float* trace; int trace_stride;
    unsigned* ctrl;            // [1] abort flag, [32] grid arrival counter (all zero at launch)
    int frame_end;             // run frames while frame < frame_end (<= max_frames)
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
LQT_DEVINL void csync() { asm volatile("bar.sync 1, %0;" ::"n"(FK_CTHREADS) : "memory"); }
LQT_DEVINL uint2 lds64(const void* p) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(smem_u32(p)));
    return r;
}
LQT_DEVINL uint4 lds128(const void* p) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)));
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
    int rps;                    // rows per ring stage
    unsigned img_off;           // element offset of this CTA's image inside the matrix image
    int pad0_, pad1_;
};
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
    float x1own[FK_X1OWN];         // post-attention residual stream at the rows of this CTA's down-projection slice
    FkDesc desc[2][10];           // [stack][phase kind]
};

struct FkCtx {
    const FkParams* p;
    FkShared* sh;
    unsigned char* ring;      // nstages * FK_STAGE_BYTES
    float* att;               // attention scratch
    float* xs;                // attention output rows [2][FK_XS_STRIDE] (input of the grouped O-projection)
    float* red;               // [2 parity][FK_CWARPS][FK_RED_STRIDE] warp-partial row sums
    float* nxt;               // running next talker input [H]
    float* res0;              // layer-0 input rows of the current pass [M][H0] (also its residual)
    float* lh;                // talker last_hidden [H] (code-predictor row 0, src/tts_onnx.cpp:859)
    int tid, lane, warp;
    int cta, ncta;
    unsigned seq;             // number of the current phase (1, 2, ...): tag of everything it publishes
    unsigned stage_ctr;       // ring stages consumed so far
    bool aborted;
    unsigned long long* dbg; int dbg_n, dbg_cap, dbg_tag;   // dbg_tag = (stack << 9) | (kind << 4) of the current phase
};

// timeline entries: (SM clock << 16) | (stack << 9) | (phase kind << 4) | point. Points:
//  0 phase begin   2 inputs in registers (polling done)   3 inputs complete (attention / staging done; before the GEMV)
//  4 glue done (sampler phases)   5 weights consumed (all FMAs done)   6 outputs published
enum { FKT_A = 1, FKT_B = 2, FKT_C = 3, FKT_D = 4, FKT_E = 5, FKT_HEAD = 6, FKT_SAMPLE = 7, FKT_INPROJ = 8 };
LQT_DEVINL void fk_mark(FkCtx& c, int point) {
    if (c.dbg && c.tid == 0 && c.dbg_n < c.dbg_cap)
        c.dbg[c.dbg_n++] = ((unsigned long long)clock64() << 16) | (unsigned)(c.dbg_tag | point);
}
LQT_DEVINL void fk_phase(FkCtx& c, int stack, int kind) { c.dbg_tag = (stack << 9) | (kind << 4); }

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
    const int r0 = (int)(((unsigned)s * (unsigned)Nout) / (unsigned)ns), r1 = (int)(((unsigned)(s + 1) * (unsigned)Nout) / (unsigned)ns);
    return FkSlice{r0, r1 - r0};
}
LQT_DEVINL FkDesc make_desc(const FkParams& p, bool is_cp, int kind, int cta, int ncta) {
    const FkStack& S = is_cp ? p.cp : p.talker;
    const int H = S.H, qkv_dim = (S.heads + 2 * S.kv_heads) * ATT_D, gK = (S.heads / S.kv_heads) * ATT_D;
    FkSlice s{0, 0}; int K = 256, RG = 1, rmax = 0;
    auto flat_max = [&](int N, int rg) { return ((N / rg + ncta - 1) / ncta) * rg; };
    switch (kind) {
        case FKT_INPROJ: s = flat_slice(H, 1, cta, ncta); K = p.talker.H; rmax = flat_max(H, 1); break;
        case FKT_A: s = flat_slice(qkv_dim, 1, cta, ncta); K = H; rmax = flat_max(qkv_dim, 1); break;
        case FKT_C: s = group_slice(H, cta, ncta, S.kv_heads); K = gK; rmax = (H + ncta / S.kv_heads - 1) / (ncta / S.kv_heads); break;
        case FKT_D: s = flat_slice(2 * S.inter, 2, cta, ncta); K = H; RG = 2; rmax = flat_max(2 * S.inter, 2); break;
        case FKT_E: s = flat_slice(H, 1, cta, ncta); K = S.inter; rmax = flat_max(H, 1); break;
        case FKT_HEAD: { const int V = is_cp ? p.cp_vocab : p.vocab; s = flat_slice(V, 1, cta, ncta); K = H; rmax = flat_max(V, 1); break; }
        default: break;
    }
    FkDesc d;
    d.row0 = s.row0; d.nrows = s.nrows; d.K = K; d.RG = RG;
    d.rps = max(1, FK_STAGE_BYTES / (K * 2));
    if (kind != FKT_C) { const int grp = (K <= 1024) ? 8 : 4; if (d.rps > grp) d.rps -= d.rps % grp; }
    d.img_off = (unsigned)cta * (unsigned)rmax * (unsigned)K;
    d.pad0_ = 0; d.pad1_ = 0;
    return d;
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
        const FkLL4 q = ll_poll4_slow(p, seq, &c.sh->aborted, c.p->ctrl);
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
    unsigned v;
    do {
        if ((++spins & 1023) == 0 && ll_giveup_slow(aborted, ctrl, &t0)) break;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while ((int)(v - target) < 0);
}
// all consumer threads; returns when every CTA has issued the outputs of phase `n`
LQT_DEVINL void grid_wait(FkCtx& c, unsigned n) {
    if (c.tid == 0 && n != 0) {
        const unsigned target = n * (unsigned)c.ncta;
        unsigned v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c.p->ctrl + 32) : "memory");
        if ((int)(v - target) < 0) grid_wait_slow(c.p->ctrl + 32, target, &c.sh->aborted, c.p->ctrl);
    }
    csync();
}
LQT_DEVINL float ll_poll1(FkCtx& c, const uint2* p, unsigned seq) {
    uint2 a = ld_ll1(p);
    if (a.y != seq) a = ll_poll1_slow(p, seq, &c.sh->aborted, c.p->ctrl, 100);
    return __uint_as_float(a.x);
}

// ------------------------------------------------------------------------------------------------
// K-split matrix-vector product over this CTA's rows, weights from the ring
// ------------------------------------------------------------------------------------------------
LQT_DEVINL void wait_full(FkCtx& c, unsigned st, int nstages) {
    const unsigned slot = st % (unsigned)nstages, par = (st / (unsigned)nstages) & 1u;
    if (!mbar_try_wait(&c.sh->full[slot], par)) {
        if (!wait_full_slow(&c.sh->full[slot], par, c.p->ctrl)) c.aborted = true;
    }
}

// sums of 8 values over the 32 lanes with a transposing butterfly: on return every lane holds the
// full-warp sum of a[lane >> 2] (9 shuffles instead of 40)
LQT_DEVINL float reduce8(const float (&a)[8], int lane) {
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
    float q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b4 ? a[i] : a[i + 4], keep = b4 ? a[i + 4] : a[i];
        q[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float d[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b3 ? q[i] : q[i + 2], keep = b3 ? q[i + 2] : q[i];
        d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float send = b2 ? d[0] : d[1], keep = b2 ? d[1] : d[0];
    float s = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    return s;
}

// sums of 4 values over the 32 lanes: on return every lane holds the full-warp sum of a[lane >> 3] (6 shuffles)
LQT_DEVINL float reduce4(const float (&a)[4], int lane) {
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0;
    float d[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b4 ? a[i] : a[i + 2], keep = b4 ? a[i + 2] : a[i];
        d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    const float send = b3 ? d[0] : d[1], keep = b3 ? d[1] : d[0];
    float s = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    return s;
}

// One group of NR rows x NJ 1024-column chunks, branch-free: all weight loads are issued before the FMAs.
// Rows beyond nr re-read the last row (discarded), chunks a thread does not own read offset 0 with x = 0.
template <int KJ, int NJ, int NR>
LQT_DEVINL void ks_group(const unsigned char* sb, int rowbytes, int nr, const int (&coff)[KJ], const float (&xr)[KJ][4],
                         float* red, int lane) {
    uint2 w[NR][NJ];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const unsigned char* rb = sb + min(r, nr - 1) * rowbytes;
#pragma unroll
        for (int j = 0; j < NJ; ++j) w[r][j] = lds64(rb + coff[j]);
    }
    float a[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        a[r] = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            a[r] = fmaf(bf16lo(w[r][j].x), xr[j][0], a[r]); a[r] = fmaf(bf16hi(w[r][j].x), xr[j][1], a[r]);
            a[r] = fmaf(bf16lo(w[r][j].y), xr[j][2], a[r]); a[r] = fmaf(bf16hi(w[r][j].y), xr[j][3], a[r]);
        }
    }
    constexpr int SH = (NR == 8) ? 2 : 3;                       // log2(lanes per row) after the butterfly
    const int rl = lane >> SH;
    float s;
    if (NR == 8) s = reduce8(reinterpret_cast<const float(&)[8]>(a), lane);
    else         s = reduce4(reinterpret_cast<const float(&)[4]>(a), lane);
    if ((lane & ((1 << SH) - 1)) == 0 && rl < nr) red[rl] = s;
}

// xr[j][i] = input element j*1024 + 4*tid + i (zero beyond K). Leaves the eight warp partials of every row in
// c.red (parity of c.seq); the caller synchronises and runs the epilogue.
template <int KJ, int NST>
LQT_DEVINL void gemv_ks(FkCtx& c, const FkDesc& d, const float (&xr)[KJ][4]) {
    float* red = c.red + (c.seq & 1u) * (FK_CWARPS * FK_RED_STRIDE) + c.warp * FK_RED_STRIDE;
    const int rowbytes = d.K * 2, tid4 = c.tid * 4;
    const int kj = (d.K + 1023) >> 10;
    int coff[KJ];
#pragma unroll
    for (int j = 0; j < KJ; ++j) coff[j] = (j * 1024 + tid4 < d.K) ? j * 2048 + c.tid * 8 : 0;
    const int nst = (d.nrows + d.rps - 1) / d.rps;
    int row = 0;
#pragma unroll 1
    for (int st = 0; st < nst; ++st) {
        const unsigned ast = c.stage_ctr + st, slot = ast % (unsigned)NST;
        wait_full(c, ast, NST);
        const int nrs = min(d.rps, d.nrows - row);
        const unsigned char* sb = c.ring + (size_t)slot * FK_STAGE_BYTES;
        if (kj == 1) {
#pragma unroll 1
            for (int r0 = 0; r0 < nrs; r0 += 8) ks_group<KJ, 1, 8>(sb + (size_t)r0 * rowbytes, rowbytes, nrs - r0, coff, xr, red + row + r0, c.lane);
        } else if (kj <= 3 || KJ <= 3) {
#pragma unroll 1
            for (int r0 = 0; r0 < nrs; r0 += 4) ks_group<KJ, 3, 4>(sb + (size_t)r0 * rowbytes, rowbytes, nrs - r0, coff, xr, red + row + r0, c.lane);
        } else {
#pragma unroll 1
            for (int r0 = 0; r0 < nrs; r0 += 4) ks_group<KJ, KJ, 4>(sb + (size_t)r0 * rowbytes, rowbytes, nrs - r0, coff, xr, red + row + r0, c.lane);
        }
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.sh->empty[slot]);       // this warp is done with the stage
        row += nrs;
    }
    c.stage_ctr += nst;
}

// Row-per-warp product for the grouped O-projection (K = rep*128 <= 512, input row in c.xs): warp w owns
// rows w, w + 8, ... of every stage; results are published straight from the reduction (no CTA barrier).
template <int NST, int NC>
LQT_DEVINL void gemv_rw_n(FkCtx& c, const FkDesc& d, uint2* out) {
    const int K = d.K, rowbytes = K * 2;
    float x0[NC][8];
    int coff[NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc) {
        const int k = cc * 256 + c.lane * 8;
        const bool act = k < K;
        coff[cc] = act ? k * 2 : 0;
        float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0;
        if (act) { u0 = *reinterpret_cast<const float4*>(c.xs + k); u1 = *reinterpret_cast<const float4*>(c.xs + k + 4); }
        x0[cc][0] = u0.x; x0[cc][1] = u0.y; x0[cc][2] = u0.z; x0[cc][3] = u0.w; x0[cc][4] = u1.x; x0[cc][5] = u1.y; x0[cc][6] = u1.z; x0[cc][7] = u1.w;
    }
    const int nst = (d.nrows + d.rps - 1) / d.rps;
    int row = 0;
#pragma unroll 1
    for (int st = 0; st < nst; ++st) {
        const unsigned ast = c.stage_ctr + st, slot = ast % (unsigned)NST;
        wait_full(c, ast, NST);
        const int nrs = min(d.rps, d.nrows - row);
        const unsigned char* sb = c.ring + (size_t)slot * FK_STAGE_BYTES;
#pragma unroll 1
        for (int r0 = 0; r0 < nrs; r0 += 64) {
            uint4 w[8][NC];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned char* rb = sb + min(r0 + c.warp + 8 * i, nrs - 1) * rowbytes;
#pragma unroll
                for (int cc = 0; cc < NC; ++cc) w[i][cc] = lds128(rb + coff[cc]);
            }
            float a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = 0.f;
#pragma unroll
                for (int cc = 0; cc < NC; ++cc) {
                    a[i] = fmaf(bf16lo(w[i][cc].x), x0[cc][0], a[i]); a[i] = fmaf(bf16hi(w[i][cc].x), x0[cc][1], a[i]);
                    a[i] = fmaf(bf16lo(w[i][cc].y), x0[cc][2], a[i]); a[i] = fmaf(bf16hi(w[i][cc].y), x0[cc][3], a[i]);
                    a[i] = fmaf(bf16lo(w[i][cc].z), x0[cc][4], a[i]); a[i] = fmaf(bf16hi(w[i][cc].z), x0[cc][5], a[i]);
                    a[i] = fmaf(bf16lo(w[i][cc].w), x0[cc][6], a[i]); a[i] = fmaf(bf16hi(w[i][cc].w), x0[cc][7], a[i]);
                }
            }
            const float s = reduce8(a, c.lane);
            const int r = r0 + c.warp + 8 * (c.lane >> 2);
            if ((c.lane & 3) == 0 && r < nrs) st_ll(out + d.row0 + row + r, s, c.seq);
        }
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.sh->empty[slot]);
        row += nrs;
    }
    c.stage_ctr += nst;
}
template <int NST>
LQT_DEVINL void gemv_rw(FkCtx& c, const FkDesc& d, uint2* out) {
    if (d.K <= 256) gemv_rw_n<NST, 1>(c, d, out); else gemv_rw_n<NST, 2>(c, d, out);
}

LQT_DEVINL float red_total(const float* red, int idx) {        // fixed order: bit-reproducible
    float s = red[idx];
#pragma unroll
    for (int w = 1; w < FK_CWARPS; ++w) s += red[w * FK_RED_STRIDE + idx];
    return s;
}
LQT_DEVINL float ss_rstd(FkCtx& c, int K, float eps) {
    const float* q = &c.sh->ssred[c.seq & 1u][0];
    float t = q[0];
#pragma unroll
    for (int w = 1; w < FK_CWARPS; ++w) t += q[w];
    return 1.0f / sqrtf(t / (float)K + eps);
}
// this thread's partial sum of squares -> shared (parity of c.seq); read after the phase's CTA barrier
LQT_DEVINL void ss_publish(FkCtx& c, float ss) {
    ss = warp_sum(ss);
    if (c.lane == 0) c.sh->ssred[c.seq & 1u][c.warp] = ss;
}

// ------------------------------------------------------------------------------------------------
// attention pieces
// ------------------------------------------------------------------------------------------------
// attention scratch layout (floats) inside c.att
constexpr int FA_Q = 0;                              // q_s   [2 m][2 r][128]
constexpr int FA_KN = FA_Q + 4 * ATT_D;              // knew  [2 m][128]
constexpr int FA_VN = FA_KN + 2 * ATT_D;             // vnew  [2 m][128]
constexpr int FA_SC = FA_VN + 2 * ATT_D;             // sc    [2 m][2 r][32]
constexpr int FA_WM = FA_SC + 4 * FK_CP_POS;         // wm    [8 w][2 r]  , wl [8][2]
constexpr int FA_WL = FA_WM + FK_CWARPS * 2;
constexpr int FA_WO = FA_WL + FK_CWARPS * 2;         // wo    [8 w][2 r][128]
constexpr int FA_FLOATS = FA_WO + FK_CWARPS * 2 * ATT_D;

template <typename KVT>
LQT_DEVINL float4 kv_load4_cg(const KVT* p);
template <> LQT_DEVINL float4 kv_load4_cg<bf16_t>(const bf16_t* p) {
    const uint2 u = __ldcg(reinterpret_cast<const uint2*>(p));
    return make_float4(bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y));
}
template <> LQT_DEVINL float4 kv_load4_cg<float>(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// talker: split-KV partial attention of kv group g over this CTA's chunk of positions -> pa (LL)
template <typename KVT>
LQT_DEVINL void talker_attn_partial(FkCtx& c, const FkLayer& L, int layer, int t, unsigned want) {
    const FkParams& p = *c.p;
    const FkStack& S = p.talker;
    const int n_kv = S.kv_heads, g = c.cta % n_kv, s = c.cta / n_kv, ns = min(grp_members(g, n_kv, c.ncta), FK_NS_MAX);
    const int n_pos = t + 1, chunk = (n_pos + ns - 1) / ns;
    const int j0 = s * chunk, j1 = min(n_pos, j0 + chunk);
    if (s >= ns || j0 >= j1) return;                           // idle split (short contexts / spare CTAs)
    const int PS = 1 << p.page_shift;
    const int q_dim = S.heads * ATT_D, kv_dim = n_kv * ATT_D;
    const float* cosr = S.cos + (size_t)t * (ATT_D / 2);
    const float* sinr = S.sin + (size_t)t * (ATT_D / 2);
    KVT* pool = reinterpret_cast<KVT*>(p.kv_pool);
    const long long layer_off = (long long)layer * 2 * n_kv * PS * ATT_D;
    const long long head_off = (long long)g * PS * ATT_D, v_off = (long long)n_kv * PS * ATT_D;
    float* q_s = c.att + FA_Q;
    float* kn = c.att + FA_KN;
    float* vn = c.att + FA_VN;
    const bool owns_new = (j1 == n_pos);
    if (c.warp < 2) {
        float4 v = ll_poll4(c, S.qkv + (size_t)(g * 2 + c.warp) * ATT_D + c.lane * 4, want);
        v = head_norm_rope(v, L.qnorm, p.eps, cosr, sinr, c.lane);
        reinterpret_cast<float4*>(q_s + c.warp * ATT_D)[c.lane] = v;
    } else if (owns_new && c.warp < 4) {
        const long long base = (long long)p.page_table[t >> p.page_shift] * p.page_stride + layer_off + head_off +
                               (long long)(t & (PS - 1)) * ATT_D;
        if (c.warp == 2) {
            float4 v = ll_poll4(c, S.qkv + q_dim + (size_t)g * ATT_D + c.lane * 4, want);
            v = head_norm_rope(v, L.knorm, p.eps, cosr, sinr, c.lane);
            KvIO<KVT>::store4(pool + base + c.lane * 4, v);
            v.x = KvIO<KVT>::round(v.x); v.y = KvIO<KVT>::round(v.y); v.z = KvIO<KVT>::round(v.z); v.w = KvIO<KVT>::round(v.w);
            reinterpret_cast<float4*>(kn)[c.lane] = v;
        } else {
            float4 v = ll_poll4(c, S.qkv + q_dim + kv_dim + (size_t)g * ATT_D + c.lane * 4, want);
            KvIO<KVT>::store4(pool + base + v_off + c.lane * 4, v);
            v.x = KvIO<KVT>::round(v.x); v.y = KvIO<KVT>::round(v.y); v.z = KvIO<KVT>::round(v.z); v.w = KvIO<KVT>::round(v.w);
            reinterpret_cast<float4*>(vn)[c.lane] = v;
        }
        __threadfence();        // the cache line must be visible before any later word of this CTA says "done"
    }
    csync();
    const float4 q0 = reinterpret_cast<const float4*>(q_s)[c.lane];
    const float4 q1 = reinterpret_cast<const float4*>(q_s + ATT_D)[c.lane];
    const float scale = 1.0f / sqrtf((float)ATT_D);
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    for (int jb = j0 + c.warp; jb < j1; jb += FK_CWARPS * 4) {
        float4 kk[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = jb + u * FK_CWARPS;
            if (j < j1) {
                if (j == t) {
                    kk[u] = reinterpret_cast<const float4*>(kn)[c.lane];
                    vv[u] = reinterpret_cast<const float4*>(vn)[c.lane];
                } else {
                    const KVT* kp = pool + (long long)p.page_table[j >> p.page_shift] * p.page_stride + layer_off + head_off +
                                    (long long)(j & (PS - 1)) * ATT_D + c.lane * 4;
                    kk[u] = kv_load4_cg<KVT>(kp);
                    vv[u] = kv_load4_cg<KVT>(kp + v_off);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = jb + u * FK_CWARPS;
            if (j < j1) {
                float d0 = kk[u].x * q0.x + kk[u].y * q0.y + kk[u].z * q0.z + kk[u].w * q0.w;
                float d1 = kk[u].x * q1.x + kk[u].y * q1.y + kk[u].z * q1.z + kk[u].w * q1.w;
                d0 = warp_sum(d0) * scale; d1 = warp_sum(d1) * scale;
                const float n0 = fmaxf(m0, d0), n1 = fmaxf(m1, d1);
                const float c0 = expf(m0 - n0), c1 = expf(m1 - n1), p0 = expf(d0 - n0), p1 = expf(d1 - n1);
                l0 = l0 * c0 + p0; l1 = l1 * c1 + p1;
                a0.x = a0.x * c0 + p0 * vv[u].x; a0.y = a0.y * c0 + p0 * vv[u].y; a0.z = a0.z * c0 + p0 * vv[u].z; a0.w = a0.w * c0 + p0 * vv[u].w;
                a1.x = a1.x * c1 + p1 * vv[u].x; a1.y = a1.y * c1 + p1 * vv[u].y; a1.z = a1.z * c1 + p1 * vv[u].z; a1.w = a1.w * c1 + p1 * vv[u].w;
                m0 = n0; m1 = n1;
            }
        }
    }
    float* wm = c.att + FA_WM; float* wl = c.att + FA_WL; float* wo = c.att + FA_WO;
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
}

// talker: combine the splits of group g -> c.xs[0][0..rep*128)  (input of the grouped O-projection)
LQT_DEVINL void talker_attn_combine(FkCtx& c, int t, unsigned want) {
    const FkParams& p = *c.p;
    const int n_kv = p.talker.kv_heads, g = c.cta % n_kv, ns = min(grp_members(g, n_kv, c.ncta), FK_NS_MAX);
    const int n_pos = t + 1, chunk = (n_pos + ns - 1) / ns, active = (n_pos + chunk - 1) / chunk;
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
    c.xs[c.tid] = num / den;
    csync();
}

// code predictor: full attention of kv group g for the M new positions p0.., result -> c.xs[m][0..256)
LQT_DEVINL void cp_attn_local(FkCtx& c, const FkLayer& L, int layer, int M, int p0, unsigned want) {
    const FkParams& p = *c.p;
    const FkStack& S = p.cp;
    const int n_kv = S.kv_heads, g = c.cta % n_kv, s = c.cta / n_kv;
    const int q_dim = S.heads * ATT_D, kv_dim = n_kv * ATT_D, qkv_dim = q_dim + 2 * kv_dim;
    float* q_s = c.att + FA_Q; float* kn = c.att + FA_KN; float* vn = c.att + FA_VN; float* sc = c.att + FA_SC;
    float* kc = p.cp_kv + ((size_t)(layer * 2 + 0) * n_kv + g) * FK_CP_POS * ATT_D;
    float* vc = p.cp_kv + ((size_t)(layer * 2 + 1) * n_kv + g) * FK_CP_POS * ATT_D;
    // prefetch the cached V column of this thread and the cached K rows of this warp (positions < p0)
    // while q/k/v of the new rows are polled
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
    // warps 0..2M-1: q heads ; 2M..3M-1: k ; 3M..4M-1: v
    for (int job = c.warp; job < 4 * M; job += FK_CWARPS) {
        if (job < 2 * M) {
            const int m = job >> 1, r = job & 1, pos = p0 + m;
            float4 v = ll_poll4(c, S.qkv + (size_t)m * qkv_dim + (size_t)(g * 2 + r) * ATT_D + c.lane * 4, want);
            v = head_norm_rope(v, L.qnorm, p.eps, S.cos + (size_t)pos * (ATT_D / 2), S.sin + (size_t)pos * (ATT_D / 2), c.lane);
            reinterpret_cast<float4*>(q_s + (m * 2 + r) * ATT_D)[c.lane] = v;
        } else if (job < 3 * M) {
            const int m = job - 2 * M, pos = p0 + m;
            float4 v = ll_poll4(c, S.qkv + (size_t)m * qkv_dim + q_dim + (size_t)g * ATT_D + c.lane * 4, want);
            v = head_norm_rope(v, L.knorm, p.eps, S.cos + (size_t)pos * (ATT_D / 2), S.sin + (size_t)pos * (ATT_D / 2), c.lane);
            reinterpret_cast<float4*>(kn + m * ATT_D)[c.lane] = v;
            if (s == 0) { reinterpret_cast<float4*>(kc + (size_t)pos * ATT_D)[c.lane] = v; __threadfence(); }
        } else {
            const int m = job - 3 * M, pos = p0 + m;
            const float4 v = ll_poll4(c, S.qkv + (size_t)m * qkv_dim + q_dim + kv_dim + (size_t)g * ATT_D + c.lane * 4, want);
            reinterpret_cast<float4*>(vn + m * ATT_D)[c.lane] = v;
            if (s == 0) { reinterpret_cast<float4*>(vc + (size_t)pos * ATT_D)[c.lane] = v; __threadfence(); }
        }
    }
    csync();
    const float scale = 1.0f / sqrtf((float)ATT_D);
    // scores: key j against both heads of every row m that may see it (warp handles j = warp, warp + 8, warp + 16)
    const int n_last = p0 + M;                       // positions visible to the last row
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const int j = c.warp + u * FK_CWARPS;
        if (j < n_last) {
            float4 k4;
            if (j >= p0) k4 = reinterpret_cast<const float4*>(kn + (j - p0) * ATT_D)[c.lane];
            else if (u < 2) k4 = kpre[u < 2 ? u : 0];
            else k4 = __ldcg(reinterpret_cast<const float4*>(kc + (size_t)j * ATT_D) + c.lane);
            float d[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {            // i = m*2 + r
                const float4 q = reinterpret_cast<const float4*>(q_s + i * ATT_D)[c.lane];
                d[i] = k4.x * q.x + k4.y * q.y + k4.z * q.z + k4.w * q.w;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int i = 0; i < 4; ++i) d[i] += __shfl_xor_sync(0xffffffffu, d[i], o);
            if (c.lane < 2 * M) {
                const int m = c.lane >> 1;
                float v = d[0];
                if (c.lane == 1) v = d[1]; else if (c.lane == 2) v = d[2]; else if (c.lane == 3) v = d[3];
                if (j <= p0 + m) sc[c.lane * FK_CP_POS + j] = v * scale;
            }
        }
    }
    csync();
    if (c.warp < 2 * M) {                            // softmax of row (m, r) over j <= p0 + m
        const int m = c.warp >> 1, np = p0 + m + 1;
        float* row = sc + c.warp * FK_CP_POS;
        const float v = (c.lane < np) ? row[c.lane] : -INFINITY;
        const float mx = warp_max(v);
        const float e = (c.lane < np) ? expf(v - mx) : 0.f;
        const float sum = warp_sum(e);
        if (c.lane < np) row[c.lane] = e / sum;
    }
    csync();
    for (int m = 0; m < M; ++m) {
        const int np = p0 + m + 1;
        const float* row = sc + (m * 2 + r_t) * FK_CP_POS;
        float o = 0.f;
#pragma unroll
        for (int j = 0; j < FK_CP_POS / 2; ++j) if (j < p0) o = fmaf(row[j], vcol[j], o);
        for (int j = p0; j < np; ++j) o = fmaf(row[j], vn[(j - p0) * ATT_D + d_t], o);
        c.xs[m * FK_XS_STRIDE + c.tid] = o;
    }
    csync();
}
// ------------------------------------------------------------------------------------------------
// one token pass (one row) through a stack, as ONE loop over the flat op schedule so that every helper
// is instantiated exactly once. The layer-0 input row is in c.res0 (smem, plain layout). The residual
// stream of the current layer stays in registers (xin) from phase A to D.
// ------------------------------------------------------------------------------------------------
struct FkPass { bool is_cp; int pos0; bool head; };

LQT_DEVINL void f4_to(float (&d)[4], const float4& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }

// this thread's columns (4*tid.. of every 1024-chunk, J chunks) of an LL vector of K words tagged `want`
template <int J>
LQT_DEVINL void ll_row(FkCtx& c, const uint2* src, int K, unsigned want, float (&out)[J][4]) {
    const int tid4 = c.tid * 4;
    FkRaw4 raw[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { const int k = j * 1024 + tid4; raw[j] = ll_issue4(src + (k < K ? k : 0)); }
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = j * 1024 + tid4;
        const float4 v = ll_finish4(c, raw[j], src + (k < K ? k : 0), want);
        if (k < K) f4_to(out[j], v);
        else { out[j][0] = 0.f; out[j][1] = 0.f; out[j][2] = 0.f; out[j][3] = 0.f; }
    }
}
// x += sum over the kv groups of the O-projection partials, in group order
template <int J>
LQT_DEVINL void ll_add_partials(FkCtx& c, const uint2* po, int n_kv, int H, unsigned want, float (&x)[J][4]) {
    const int tid4 = c.tid * 4;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = j * 1024 + tid4;
        const bool act = k < H;
        FkRaw4 raw[FK_NGRP_MAX];
#pragma unroll
        for (int g = 0; g < FK_NGRP_MAX; ++g) raw[g] = ll_issue4(po + (size_t)(g < n_kv ? g : 0) * H + (act ? k : 0));
#pragma unroll
        for (int g = 0; g < FK_NGRP_MAX; ++g) {
            const float4 b = ll_finish4(c, raw[g], po + (size_t)(g < n_kv ? g : 0) * H + (act ? k : 0), want);
            if (g < n_kv && act) { x[j][0] += b.x; x[j][1] += b.y; x[j][2] += b.z; x[j][3] += b.w; }
        }
    }
}

template <int KJ, int NST>
LQT_DEVINL void consume_token(FkCtx& c, const FkParams& p, const FkPass& ps) {
    constexpr int HJ = (KJ > 3) ? 2 : 1;                         // 1024-column chunks of the hidden size
    const bool is_cp = ps.is_cp;
    const FkStack& S = is_cp ? p.cp : p.talker;
    const int tk = is_cp ? 1 : 0;
    const int H = S.H, n_kv = S.kv_heads;
    const int H0 = p.talker.H;                                   // width of the row in res0
    const bool inproj = is_cp && p.c_inproj_w != nullptr;
    const int total = pass_ops(S.n_layers, inproj, ps.head);
    const int tid4 = c.tid * 4;
    float xin[HJ][4];                                            // layer input (residual stream), this thread's columns
#pragma unroll
    for (int j = 0; j < HJ; ++j) { xin[j][0] = 0.f; xin[j][1] = 0.f; xin[j][2] = 0.f; xin[j][3] = 0.f; }

    for (int it = 0; it < total && !c.aborted; ++it) {
        const FkOp op = pass_op(it, S.n_layers, inproj);
        const int kind = op.kind, l = op.layer;
        if (kind == FKT_B && is_cp) continue;                 // the predictor's attention lives inside phase C
        ++c.seq;
        const unsigned want = c.seq - 1;
        fk_phase(c, tk, kind);
        fk_mark(c, 0);
        grid_wait(c, want);
        fk_mark(c, 1);
        const FkLayer& L = is_cp ? p.c_layers[l] : p.t_layers[l];
        const FkDesc d = c.sh->desc[tk][kind];
        // the layer input row: res0 (smem) for layer 0 without in_proj, else an LL buffer
        const bool in_res0 = (l == 0 && !inproj);
        const uint2* lin = (l == 0) ? p.cxin : S.x;           // (only read when !in_res0)
        if (kind == FKT_B) {
            if (p.kv_f32) talker_attn_partial<float>(c, L, l, ps.pos0, want);
            else          talker_attn_partial<bf16_t>(c, L, l, ps.pos0, want);
            csync();
            if (c.tid == 0) grid_arrive(c);
            fk_mark(c, 3);
            if (c.sh->aborted) { c.aborted = true; break; }
            continue;
        }
        if (kind == FKT_C) {
            if (is_cp) cp_attn_local(c, L, l, 1, ps.pos0, want);
            else       talker_attn_combine(c, ps.pos0, want);
            fk_mark(c, 3);
            if (c.sh->aborted) { c.aborted = true; break; }
            gemv_rw<NST>(c, d, S.po + (size_t)(c.cta % n_kv) * H);
            csync();
            if (c.tid == 0) grid_arrive(c);
            fk_mark(c, 6);
            if (c.aborted) break;
            continue;
        }
        // ---- K-split phases: inputs -> registers ------------------------------------------------------
        float xr[KJ][4];
#pragma unroll
        for (int j = 0; j < KJ; ++j) { xr[j][0] = 0.f; xr[j][1] = 0.f; xr[j][2] = 0.f; xr[j][3] = 0.f; }
        const float* nw = nullptr;
        switch (kind) {
            case FKT_INPROJ: {
#pragma unroll
                for (int j = 0; j < KJ; ++j)
                    if (j * 1024 + tid4 < H0) f4_to(xr[j], *reinterpret_cast<const float4*>(c.res0 + j * 1024 + tid4));
                break;
            }
            case FKT_A: {
                if (in_res0) {
#pragma unroll
                    for (int j = 0; j < HJ; ++j)
                        if (j * 1024 + tid4 < H) f4_to(xin[j], *reinterpret_cast<const float4*>(c.res0 + j * 1024 + tid4));
                } else {
                    ll_row<HJ>(c, lin, H, want, xin);
                }
                nw = L.ln1;
                break;
            }
            case FKT_D: {
                ll_add_partials<HJ>(c, S.po, n_kv, H, want, xin);
                const FkDesc& de = c.sh->desc[tk][FKT_E];
#pragma unroll
                for (int j = 0; j < HJ; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {                          // residual for this CTA's down-projection rows
                        const unsigned rel = (unsigned)(j * 1024 + tid4 + i - de.row0);
                        if (rel < (unsigned)de.nrows) c.sh->x1own[rel] = xin[j][i];
                    }
                nw = L.ln2;
                break;
            }
            case FKT_E: ll_row<KJ>(c, S.act, d.K, want, xr); break;
            default:   // FKT_HEAD: final norm + head
                ll_row<HJ>(c, S.x, H, want, xin);
                nw = S.final_norm;
                break;
        }
        fk_mark(c, 2);
        if (nw) {                                              // RMSNorm, folded: product on x*w, 1/rms in the epilogue
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < HJ; ++j) {
                if (j * 1024 + tid4 < H) {
                    const float4 w = __ldg(reinterpret_cast<const float4*>(nw + j * 1024 + tid4));
                    ss = fmaf(xin[j][0], xin[j][0], ss); ss = fmaf(xin[j][1], xin[j][1], ss);
                    ss = fmaf(xin[j][2], xin[j][2], ss); ss = fmaf(xin[j][3], xin[j][3], ss);
                    xr[j][0] = xin[j][0] * w.x; xr[j][1] = xin[j][1] * w.y; xr[j][2] = xin[j][2] * w.z; xr[j][3] = xin[j][3] * w.w;
                }
            }
            ss_publish(c, ss);
        }
        fk_mark(c, 3);
        if (c.sh->aborted) { c.aborted = true; break; }
        gemv_ks<KJ, NST>(c, d, xr);
        fk_mark(c, 5);
        csync();
        // ---- epilogue: warp partials -> outputs (warp 0 publishes everything, then signals the grid) ----------
        {
            const float* red = c.red + (c.seq & 1u) * (FK_CWARPS * FK_RED_STRIDE);
            const float rs = nw ? ss_rstd(c, H, p.eps) : 1.f;
            if (c.warp == 0) {
                if (kind == FKT_D) {
                    const int nq = d.nrows >> 1;
                    for (int q = c.lane; q < nq; q += 32) {
                        const float g = red_total(red, 2 * q) * rs, u = red_total(red, 2 * q + 1) * rs;
                        st_ll(S.act + (d.row0 >> 1) + q, silu_f(g) * u, c.seq);
                    }
                } else {
                    for (int r = c.lane; r < d.nrows; r += 32) {
                        const int n = d.row0 + r;
                        const float v = red_total(red, r) * rs;
                        if (kind == FKT_A) st_ll(S.qkv + n, v, c.seq);
                        else if (kind == FKT_E) st_ll(S.x + n, c.sh->x1own[r] + v, c.seq);
                        else if (kind == FKT_INPROJ) st_ll(p.cxin + n, v + __ldg(p.c_inproj_b + n), c.seq);
                        else {
                            st_ll((is_cp ? p.clogits_ll : p.logits_ll) + n, v, c.seq);
                            (is_cp ? p.clogits : p.logits)[n] = v;
                        }
                    }
                }
                __syncwarp();
                if (c.lane == 0) grid_arrive(c);
            }
            if (kind == FKT_HEAD && !is_cp) {                  // talker last_hidden = final-norm of the row (:859)
#pragma unroll
                for (int j = 0; j < HJ; ++j) {
                    if (j * 1024 + tid4 < H) {
                        const float4 w = __ldg(reinterpret_cast<const float4*>(nw + j * 1024 + tid4));
                        const float4 o = make_float4((xin[j][0] * rs) * w.x, (xin[j][1] * rs) * w.y, (xin[j][2] * rs) * w.z, (xin[j][3] * rs) * w.w);
                        *reinterpret_cast<float4*>(c.lh + j * 1024 + tid4) = o;
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

LQT_DEVINL int block_excl_scan(FkCtx& c, int v, int* total) {          // 256-thread exclusive scan
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (c.lane >= o) inc += t; }
    csync();
    if (c.lane == 31) c.sh->wtot[c.warp] = inc;
    csync();
    int base = 0, tot = 0;
#pragma unroll
    for (int w2 = 0; w2 < FK_CWARPS; ++w2) { const int t = c.sh->wtot[w2]; if (w2 < c.warp) base += t; tot += t; }
    *total = tot;
    return base + inc - v;
}

// logits: LL words tagged `want` (ll != nullptr) or a plain array (first draw after a resume)
LQT_DEVINL int fk_sample(FkCtx& c, const FkSampScratch& s, const uint2* ll, const float* plain, unsigned want, int V,
                         int mask_lo, int mask_hi, int mask_keep, const SamplingDev& sp, uint32_t frame, int codebook,
                         float* trace_row) {
    FkShared* sh = c.sh;
    const bool temper = !sp.greedy && sp.temperature > 0.0f && sp.temperature != 1.0f;
    // thread owns the contiguous range [i0, i1) (V % 4 == 0, ranges are multiples of 4)
    const int per = (((V + FK_CTHREADS - 1) / FK_CTHREADS) + 3) & ~3;
    const int i0 = min(V, c.tid * per), i1 = min(V, i0 + per);
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = i0; i < i1; i += 4) {
        const float4 q = ll ? ll_poll4(c, ll + i, want) : *reinterpret_cast<const float4*>(plain + i);
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
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
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
struct FkSmemLayout { size_t scratch, att, xs, red, nxt, res0, lh, shared, total; };
inline FkSmemLayout fk_smem_layout(int nstages, int maxV, int H, int res0_floats) {
    FkSmemLayout L{};
    auto up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
    size_t off = (size_t)nstages * FK_STAGE_BYTES;
    L.scratch = off;                                           // sampler scratch | attention scratch + attention output rows
    const size_t samp = (size_t)maxV * (4 + 4 + 4 + 2 + 2);
    const size_t attb = up((size_t)FA_FLOATS * 4), xsb = (size_t)2 * FK_XS_STRIDE * 4;
    L.att = off; L.xs = off + attb;
    off += up(samp > attb + xsb ? samp : attb + xsb);
    L.red = off; off += up((size_t)2 * FK_CWARPS * FK_RED_STRIDE * 4);
    L.nxt = off; off += up((size_t)H * 4);
    L.res0 = off; off += up((size_t)res0_floats * 4);
    L.lh = off; off += up((size_t)H * 4);
    L.shared = off; off += up(sizeof(FkShared));
    L.total = off;
    return L;
}

struct FkSmemOffsets { unsigned scratch, att, xs, red, nxt, res0, lh, shared; int maxV; };

// KJ = 1024-column chunks of the widest matrix (3: inter <= 3072, 6: <= 6144); NST = ring stages
template <int KJ, int NST>
__global__ void __launch_bounds__(FK_THREADS, 1)
frame_kernel(const __grid_constant__ FkParams p, const FkSmemOffsets so) {
    extern __shared__ __align__(1024) unsigned char fk_smem[];
    FkShared* sh = reinterpret_cast<FkShared*>(fk_smem + so.shared);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&sh->full[i], 1); mbar_init(&sh->empty[i], FK_CWARPS); }   // every consumer warp releases a stage
        sh->stop = 0; sh->consumed = 0; sh->aborted = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 20) sh->desc[tid / 10][tid % 10] = make_desc(p, tid >= 10, tid % 10, cta, ncta);
    __syncthreads();

    const GenState st0 = *p.st;                    // written by the host before launch
    const int n_prefill = (p.mode == 0 && st0.pos == 0) ? p.P : 0;
    const int frame_end = min(p.frame_end, st0.max_frames);

    if (warp == FK_CWARPS) {
        // ============================ producer warp ============================================
        // Walks the same flat schedule as the consumers, one stage at a time, as far ahead as the
        // ring allows. All state in registers of one lane.
        if (lane == 0) {
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
                    for (int r0 = 0; r0 < d.nrows && !stopped; r0 += d.rps) {
                        const int nr = min(d.rps, d.nrows - r0);
                        const unsigned slot = issued % (unsigned)NST, par = ((issued / (unsigned)NST) & 1u) ^ 1u;
                        unsigned long long t0 = 0;
                        while (!mbar_try_wait(&sh->empty[slot], par)) {
                            if (sh->stop) { stopped = true; break; }
                            if (t0 == 0) t0 = clock64();
                            else if (clock64() - t0 > FK_SPIN_LIMIT) { stopped = true; break; }
                        }
                        if (stopped) break;
                        const uint32_t bytes = (uint32_t)nr * row_bytes;
                        mbar_expect_tx(&sh->full[slot], bytes);
                        bulk_g2s(fk_smem + (size_t)slot * FK_STAGE_BYTES, srcb + (size_t)r0 * row_bytes, bytes, &sh->full[slot]);
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
        return;
    }

    // ================================ consumer warps ===============================================
    FkCtx c;
    c.p = &p; c.sh = sh; c.ring = fk_smem;
    c.att = reinterpret_cast<float*>(fk_smem + so.att);
    c.xs = reinterpret_cast<float*>(fk_smem + so.xs);
    c.red = reinterpret_cast<float*>(fk_smem + so.red);
    c.nxt = reinterpret_cast<float*>(fk_smem + so.nxt);
    c.res0 = reinterpret_cast<float*>(fk_smem + so.res0);
    c.lh = reinterpret_cast<float*>(fk_smem + so.lh);
    c.tid = tid; c.lane = lane; c.warp = warp; c.cta = cta; c.ncta = ncta;
    c.seq = 0; c.stage_ctr = 0; c.aborted = false;
    c.dbg = (p.dbg && cta == p.dbg_cta) ? p.dbg + 1 : nullptr; c.dbg_n = 0; c.dbg_cap = p.dbg_cap - 1; c.dbg_tag = 0;
    FkSampScratch ss;
    ss.x = reinterpret_cast<float*>(fk_smem + so.scratch);
    ss.pr = ss.x + so.maxV; ss.spr = ss.pr + so.maxV;
    ss.idx = reinterpret_cast<unsigned short*>(ss.spr + so.maxV); ss.rank = ss.idx + so.maxV;

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
            reinterpret_cast<float4*>(c.lh)[k4] = __ldcg(reinterpret_cast<const float4*>(p.last_hidden) + k4);
        csync();
    }

    // One loop, one pass per iteration: [draw + glue ->] token pass. (src/tts_onnx.cpp:794, 801-846)
    while (!c.aborted) {
        FkPass ps;
        if (p.mode == 1 || prefill_i < n_prefill) {
            if (p.mode == 1 && mode1_done) break;
            const float* src = (p.mode == 1) ? p.next_in : p.prompt + (size_t)prefill_i * H;
            for (int k4 = tid; k4 < H4; k4 += FK_CTHREADS)
                reinterpret_cast<float4*>(c.res0)[k4] = __ldcg(reinterpret_cast<const float4*>(src) + k4);
            csync();
            const bool head = (p.mode == 1) || (prefill_i == n_prefill - 1);
            ps = FkPass{false, pos, head};
        } else if (row1_next) {
            for (int k4 = tid; k4 < H4; k4 += FK_CTHREADS)                      // predictor position 1: codec_embed(code0) (= the running sum so far)
                reinterpret_cast<float4*>(c.res0)[k4] = reinterpret_cast<const float4*>(c.nxt)[k4];
            csync();
            ps = FkPass{true, 1, true};
            row1_next = false;
        } else {
            if (done || frame >= frame_end) break;
            // ---- draw codebook cb of this frame (:803-812 for cb 0, :863-864 otherwise) -----------------
            const int tk = cb ? 1 : 0;
            fk_phase(c, tk, FKT_SAMPLE);
            fk_mark(c, 0);
            grid_wait(c, resumed ? 0u : c.seq);                 // the logits of the head phase have been issued everywhere
            fk_mark(c, 1);
            float* tr = (p.trace && cta == 0) ? p.trace + ((size_t)frame * 16 + cb) * p.trace_stride : nullptr;
            int tok;
            if (cb == 0) tok = fk_sample(c, ss, resumed ? nullptr : p.logits_ll, p.logits, c.seq, p.vocab, 2048, p.vocab, 2150,
                                         sp, (uint32_t)frame, 0, tr);
            else         tok = fk_sample(c, ss, p.clogits_ll, nullptr, c.seq, p.cp_vocab, 0, 0, -1, sp, (uint32_t)frame, cb, tr);
            resumed = false;
            if (c.sh->aborted) { c.aborted = true; break; }
            if (p.forced && frame < st0.n_forced) tok = (int)p.forced[(size_t)frame * 16 + cb];
            fk_mark(c, 3);
            if (cb == 0 && tok == 2150) { done = 1; break; }                        // CODEC_EOS (:812)
            if (cta == 0 && tid == 0) p.codes_out[(size_t)frame * 16 + cb] = tok;   // :818-821
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
                    acc = reinterpret_cast<float4*>(c.nxt)[k4];
                    acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
                }
                if (last_cb) {                                                  // :833-842
                    const float4 tt = use_tr ? __ldg(reinterpret_cast<const float4*>(p.trailing + (size_t)frame * H) + k4)
                                             : __ldg(reinterpret_cast<const float4*>(p.tts_pad) + k4);
                    acc.x += tt.x; acc.y += tt.y; acc.z += tt.z; acc.w += tt.w;
                }
                reinterpret_cast<float4*>(c.nxt)[k4] = acc;
                if (cb == 0) {                                                  // rows [last_hidden, codec_embed(code0)] (:854-860): two single-row passes
                    reinterpret_cast<float4*>(c.res0)[k4] = reinterpret_cast<const float4*>(c.lh)[k4];
                } else {
                    reinterpret_cast<float4*>(c.res0)[k4] = last_cb ? acc : e;  // :867-868 / :845
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
    if (tid == 0) {
        if (cta == 0) {
            p.st->pos = pos; p.st->frame = frame; p.st->done = done; p.st->n_frames = n_frames;
        }
        if (c.dbg) c.dbg[-1] = (unsigned long long)c.dbg_n;
        sh->consumed = (int)c.stage_ctr;
        __threadfence_block();
        sh->stop = 1;
    }
}

}  // namespace lqt