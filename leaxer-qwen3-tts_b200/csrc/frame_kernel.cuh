// Persistent frame kernel: loops A and B of the reference (src/tts_onnx.cpp:782-872) as ONE
// cooperative launch per utterance (or per chunk of frames).
//
// Why one kernel: a frame is 31 dependent network passes (1 talker step + 15 predictor passes, each
// followed by a draw) = ~460 dependent matrix-vector phases of ~1 us each. As separate launches the
// HBM pipe drains at every kernel boundary (round-1 v1: 577 launches, 4.2 ms per frame).
// Here every SM keeps one CTA resident:
//   * warp 8 (producer) walks the static weight schedule of the whole launch and streams this CTA's
//     slice of every matrix into a shared-memory ring with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), running AHEAD of the math across phase boundaries;
//   * warps 0-7 (consumers) run the phases: assemble the input vector (RMSNorm / attention / partial
//     sums) and publish the outputs. The matrix-vector product itself runs on the 5th-gen tensor
//     cores: the weights are stored in HBM per CTA as ready-made K-major SWIZZLE_128B tiles, so the
//     bulk copies land operand A in canonical UMMA layout; the activation vector is split exactly
//     into three bf16 terms (x = hi + mid + lo, 24 mantissa bits) that form operand B (N = 8
//     columns); ONE thread issues K/16 tcgen05.mma (M=64, N=8, fp32 accumulate in TMEM) per phase,
//     tcgen05.commit frees the ring slots, and 64-128 threads read their row back with tcgen05.ld and
//     publish it in parallel. (The fp32 FMA version of this loop was issue/latency bound: 2-4 us
//     per phase with 8 warps.);
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
constexpr int FK_STAGE_BYTES = 16 * 1024;
constexpr int FK_STAGES = 8;                      // 128 KB weight ring per SM
constexpr int FK_NS_MAX = 24;                     // max attention splits per kv group
constexpr int FK_NGRP_MAX = 8;                    // kv groups
constexpr int FK_MAXV = 4096;
constexpr int FK_CP_POS = 32;                     // code-predictor KV capacity (positions)
constexpr int FK_NACC = 8;                        // TMEM accumulator sets (16 columns each), one per issuing warp
constexpr int FK_TMEM_COLS = FK_NACC * 16;        // 256
constexpr int FK_MAX_TLAYERS = 32, FK_MAX_CLAYERS = 8;
constexpr unsigned long long FK_SPIN_LIMIT = 6000000000ull;   // ~3 s of SM clocks: abort, never hang

// Weight "images": for every matrix and every CTA c, the rows this CTA owns (padded to a multiple of 8
// with zero rows) as K/64 tiles of [R8 rows][128 bytes], each tile in the canonical K-major
// SWIZZLE_128B layout (16-byte chunk index XOR (row & 7)); CTA c's image starts at c * rmax8 * K
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
    unsigned* ctrl;            // [1] abort flag (zero at launch)
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
LQT_DEVINL uint4 lds128(const void* p) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)));
    return r;
}
// ---- tcgen05 / TMEM ----
LQT_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
LQT_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
LQT_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100): start>>4 | LBO(1)<<16 | SBO(1024>>4)<<32 |
// version 1 <<46 | layout SWIZZLE_128B (2) << 61   (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp)
LQT_DEVINL uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A/B bf16, both K-major, N = 8, M = 64 (cute::UMMA::InstrDescriptor)
constexpr uint32_t FK_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 17) | (4u << 24);
LQT_DEVINL void umma_bf16_m64n8k16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(FK_IDESC), "r"(accumulate) : "memory");
}
LQT_DEVINL bool elect_one() {                     // one lane of a converged warp (warp-uniform control flow around it)
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
LQT_DEVINL void umma_commit(uint64_t* bar) {      // arrives on the mbarrier when all prior tcgen05.mma of this thread are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
LQT_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {     // this thread's TMEM lane, 16 consecutive columns
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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
    int r8;                     // rows padded to a multiple of 8 (tile = r8 rows x 128 bytes)
    int tps;                    // tiles (64-element K chunks) per ring stage
    unsigned img_off;           // element offset of this CTA's image inside the matrix image
    int pad_;
};
struct FkShared {
    uint64_t full[FK_STAGES];
    uint64_t empty[FK_STAGES];
    volatile int stop;            // consumers -> producer: stop issuing
    volatile int consumed;        // stages consumed when stop was raised
    volatile int aborted;
    int hist[256];
    int wtot[FK_CWARPS];
    float redf[FK_CWARPS][2];
    int redi[FK_CWARPS];
    uint32_t sel_prefix; int sel_k;
    int tok; float fsum;
    uint64_t mma_done;            // tcgen05.commit of the last MMA of a phase arrives here
    uint32_t tmem_base;           // written by tcgen05.alloc
    FkDesc desc[2][10];           // [stack][phase kind]
};

// Operand B of the tensor-core GEMV. Row n = 3*m + j of the 8-row tile holds term j (hi, mid, lo) of
// activation row m; element k lives in tile k/64 at byte n*128 + (((k%64)/8) ^ n)*16 + (k%8)*2.
LQT_DEVINL void bf16_split3(float x, unsigned short& h, unsigned short& m, unsigned short& l) {
    const __nv_bfloat16 bh = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(bh);
    const __nv_bfloat16 bm = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(bm);
    const __nv_bfloat16 bl = __float2bfloat16_rn(r2);
    h = __bfloat16_as_ushort(bh); m = __bfloat16_as_ushort(bm); l = __bfloat16_as_ushort(bl);
}
LQT_DEVINL void xb_store4(unsigned char* bt, int m, int k4, const float4& v) {     // elements 4*k4 .. 4*k4+3 of row m
    const int k = k4 << 2, tile = k >> 6, c16 = (k & 63) >> 3, e0 = k & 7;
    unsigned short h[4], md[4], l[4];
    bf16_split3(v.x, h[0], md[0], l[0]); bf16_split3(v.y, h[1], md[1], l[1]);
    bf16_split3(v.z, h[2], md[2], l[2]); bf16_split3(v.w, h[3], md[3], l[3]);
    unsigned char* base = bt + (size_t)tile * 1024 + e0 * 2;
    const int n0 = 3 * m;
    *reinterpret_cast<uint2*>(base + (n0 + 0) * 128 + ((c16 ^ (n0 + 0)) << 4)) =
        make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    *reinterpret_cast<uint2*>(base + (n0 + 1) * 128 + ((c16 ^ (n0 + 1)) << 4)) =
        make_uint2((uint32_t)md[0] | ((uint32_t)md[1] << 16), (uint32_t)md[2] | ((uint32_t)md[3] << 16));
    *reinterpret_cast<uint2*>(base + (n0 + 2) * 128 + ((c16 ^ (n0 + 2)) << 4)) =
        make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}
LQT_DEVINL void xb_store1(unsigned char* bt, int m, int k, float v) {
    const int tile = k >> 6, c16 = (k & 63) >> 3, e = k & 7;
    unsigned short h, md, l;
    bf16_split3(v, h, md, l);
    unsigned char* base = bt + (size_t)tile * 1024 + e * 2;
    const int n0 = 3 * m;
    *reinterpret_cast<unsigned short*>(base + (n0 + 0) * 128 + ((c16 ^ (n0 + 0)) << 4)) = h;
    *reinterpret_cast<unsigned short*>(base + (n0 + 1) * 128 + ((c16 ^ (n0 + 1)) << 4)) = md;
    *reinterpret_cast<unsigned short*>(base + (n0 + 2) * 128 + ((c16 ^ (n0 + 2)) << 4)) = l;
}

struct FkCtx {
    const FkParams* p;
    FkShared* sh;
    unsigned char* ring;      // FK_STAGES * FK_STAGE_BYTES
    unsigned char* bt;        // operand B: x as bf16 (hi, mid, lo) per row, K-major SWIZZLE_128B, K/64 tiles of 8 x 128 B | aliases the sampler scratch
    float* att;               // attention scratch
    float* nxt;               // running next talker input [H]
    float* res0;              // layer-0 input rows of the current pass [M][H0] (also its residual)
    float* lh;                // talker last_hidden [H] (code-predictor row 0, src/tts_onnx.cpp:859)
    int tid, lane, warp;
    int cta, ncta;
    unsigned seq;             // number of the current phase (1, 2, ...): tag of everything it publishes
    unsigned stage_ctr;       // ring stages consumed so far
    unsigned mma_phase;       // parity of the next wait on sh->mma_done
    uint32_t tmem;            // TMEM base address (lane 0, column 0) of the 32 allocated columns
    bool aborted;
    unsigned long long* dbg; int dbg_n, dbg_cap, dbg_tag;   // dbg_tag = (stack << 9) | (kind << 4) of the current phase
};

// timeline entries: (SM clock << 16) | (stack << 9) | (phase kind << 4) | point. Points:
//  0 phase begin   1 input probe passed   2 inputs fetched   3 inputs complete (polling, attention, RMSNorm done; before the GEMV)
//  4 glue done (sampler phases)   5 all MMAs of the phase done   7 accumulators read back   6 outputs published
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
    d.r8 = (s.nrows + 7) & ~7;
    const int tile_bytes = d.r8 * 128;
    d.tps = tile_bytes > 0 ? max(1, FK_STAGE_BYTES / tile_bytes) : 1;
    d.img_off = (unsigned)cta * (unsigned)((rmax + 7) & ~7) * (unsigned)K;
    d.pad_ = 0;
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
    const int r = (int)((q - n_prefill) % (cp_steps + 1));
    return (r < cp_steps) ? FkPassId{true, r, true} : FkPassId{false, 0, true};
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
        __nanosleep(100);
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

// Readiness probe before a CTA-wide read: warp 0 polls 32 words spread over the region (one per lane,
// with back-off) and everybody else waits at the CTA barrier, so that a not-yet-complete vector costs
// 32 L2 requests per CTA and round instead of one per thread (148 CTAs polling the same lines would
// otherwise starve the very stores they wait for).
LQT_DEVINL void ll_probe(FkCtx& c, const uint2* base, int nwords, unsigned seq) {
    if (c.warp == 0) {
        const uint2* p = base + (((c.lane + 1) * nwords) >> 5) - 1;
        if (ld_ll1(p).y != seq) ll_poll1_slow(p, seq, &c.sh->aborted, c.p->ctrl, 40);
    }
    csync();
}
// four consecutive words (32-byte aligned group), all tagged `seq`
LQT_DEVINL float4 ll_poll4(FkCtx& c, const uint2* p, unsigned seq) {
    uint4 a = ld_ll2(p), b = ld_ll2(p + 2);
    if (a.y != seq || a.w != seq || b.y != seq || b.w != seq) {
        const FkLL4 r = ll_poll4_slow(p, seq, &c.sh->aborted, c.p->ctrl);
        a = r.a; b = r.b;
    }
    return make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z));
}
LQT_DEVINL float ll_poll1(FkCtx& c, const uint2* p, unsigned seq) {
    uint2 a = ld_ll1(p);
    if (a.y != seq) a = ll_poll1_slow(p, seq, &c.sh->aborted, c.p->ctrl, 100);
    return __uint_as_float(a.x);
}
// values already verified by this CTA in an earlier phase
LQT_DEVINL float4 ll_val4(const uint2* p) {
    const uint4 a = ld_ll2(p), b = ld_ll2(p + 2);
    return make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z));
}
LQT_DEVINL float ll_val1(const uint2* p) { return __uint_as_float(ld_ll1(p).x); }

// ------------------------------------------------------------------------------------------------
// input staging: rows -> xs (permuted float4), optional partial sums, optional RMSNorm (fused:
// values stay in registers until rstd is known, so xs is written exactly once)
// ------------------------------------------------------------------------------------------------
struct FkStage {
    const uint2* ll;        // polled source rows [M][ll_stride] (nullptr: rows come from `sm`)
    int ll_stride;
    bool verified;          // source words were already verified by this CTA (no polling)
    const float* sm;        // smem source rows [M][sm_stride] (plain layout)
    int sm_stride;
    const uint2* part;      // nullable: n_part polled partial vectors per row at part[(m*n_part + g)*K + k]
    int n_part;
    const float* nw;        // nullable RMSNorm weight [K]
    float* copy_sm;         // nullable: normalised row 0 also to smem plain [K] (last_hidden)
    float* copy_gl;         // nullable: and to global [K] by CTA 0
};

LQT_DEVINL float4 stage_fetch(FkCtx& c, const FkStage& s, int m, int k4, int K, unsigned want) {
    float4 a;
    if (s.ll) a = s.verified ? ll_val4(s.ll + (size_t)m * s.ll_stride + k4 * 4)
                             : ll_poll4(c, s.ll + (size_t)m * s.ll_stride + k4 * 4, want);
    else a = reinterpret_cast<const float4*>(s.sm + (size_t)m * s.sm_stride)[k4];
    if (s.part) {
        // all partial vectors in flight at once, validated afterwards (stragglers: slow path)
        uint4 q[FK_NGRP_MAX][2];
#pragma unroll
        for (int g = 0; g < FK_NGRP_MAX; ++g) {
            if (g < s.n_part) {
                const uint2* pp = s.part + ((size_t)m * s.n_part + g) * K + k4 * 4;
                q[g][0] = ld_ll2(pp); q[g][1] = ld_ll2(pp + 2);
            }
        }
#pragma unroll
        for (int g = 0; g < FK_NGRP_MAX; ++g) {
            if (g < s.n_part) {
                float4 b;
                if (q[g][0].y == want && q[g][0].w == want && q[g][1].y == want && q[g][1].w == want)
                    b = make_float4(__uint_as_float(q[g][0].x), __uint_as_float(q[g][0].z), __uint_as_float(q[g][1].x), __uint_as_float(q[g][1].z));
                else
                    b = ll_poll4(c, s.part + ((size_t)m * s.n_part + g) * K + k4 * 4, want);
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
        }
    }
    return a;
}

LQT_DEVINL void stage_rows(FkCtx& c, const FkStage& s, int M, int K, unsigned want, float eps) {
    const int K4 = K >> 2;
    if (s.part) ll_probe(c, s.part, M * s.n_part * K, want);
    else if (s.ll && !s.verified) ll_probe(c, s.ll, (M - 1) * s.ll_stride + K, want);
    fk_mark(c, 1);
    const bool norm = s.nw != nullptr;         // norm path: K = hidden <= 2048 -> at most 2 float4 per thread per row,
    float4 v00, v01, v10, v11;                 // kept in registers until rstd is known (operand B is written once)
    v00 = v01 = v10 = v11 = make_float4(0.f, 0.f, 0.f, 0.f);
    float ss0 = 0.f, ss1 = 0.f;
#pragma unroll 1
    for (int m = 0; m < M; ++m) {
#pragma unroll 1
        for (int k4 = c.tid, i = 0; k4 < K4; k4 += FK_CTHREADS, ++i) {
            const float4 a = stage_fetch(c, s, m, k4, K, want);
            if (norm) {
                const float q = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
                if (m == 0) { ss0 += q; if (i == 0) v00 = a; else v01 = a; }
                else        { ss1 += q; if (i == 0) v10 = a; else v11 = a; }
            } else {
                xb_store4(c.bt, m, k4, a);
            }
        }
    }
    fk_mark(c, 2);
    if (norm) {
        ss0 = warp_sum(ss0); ss1 = warp_sum(ss1);
        if (c.lane == 0) { c.sh->redf[c.warp][0] = ss0; c.sh->redf[c.warp][1] = ss1; }
        csync();
        float t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < FK_CWARPS; ++w2) { t0 += c.sh->redf[w2][0]; t1 += c.sh->redf[w2][1]; }
        const float r0 = 1.0f / sqrtf(t0 / (float)K + eps), r1 = 1.0f / sqrtf(t1 / (float)K + eps);
#pragma unroll 1
        for (int m = 0; m < M; ++m) {
#pragma unroll 1
            for (int k4 = c.tid, i = 0; k4 < K4; k4 += FK_CTHREADS, ++i) {
                float4 a = (m == 0) ? (i == 0 ? v00 : v01) : (i == 0 ? v10 : v11);
                const float r = (m == 0) ? r0 : r1;
                const float4 w = __ldg(reinterpret_cast<const float4*>(s.nw) + k4);
                a.x = (a.x * r) * w.x; a.y = (a.y * r) * w.y; a.z = (a.z * r) * w.z; a.w = (a.w * r) * w.w;
                if (m == 0) {
                    if (s.copy_sm) reinterpret_cast<float4*>(s.copy_sm)[k4] = a;
                    if (s.copy_gl && c.cta == 0) reinterpret_cast<float4*>(s.copy_gl)[k4] = a;
                }
                xb_store4(c.bt, m, k4, a);
            }
        }
    }
    fence_proxy_async_smem();                  // operand B was written through the generic proxy; the MMA reads it through the async proxy
    csync();
}

// ------------------------------------------------------------------------------------------------
// GEMV over this CTA's slice, weights from the ring, outputs published as LL words
// ------------------------------------------------------------------------------------------------
enum { EPI_STORE = 0, EPI_GLU = 1, EPI_RESID = 2, EPI_LOGITS = 3 };
struct FkEpi {
    int kind;
    uint2* out; int out_stride;          // out[m*out_stride + n]   (GLU: n/2)
    const float* bias;                   // STORE only, nullable
    float* plain;                        // LOGITS: second, plain copy [n]
    // RESID: out = (x_in[n] + sum_g po[g][n]) + v, recomputed exactly as phase D staged it
    const float* res_sm; int res_sm_stride;     // layer-0 residual rows in smem, or
    const uint2* res_ll; int res_ll_stride;     // verified LL rows
    const uint2* po; int n_part, H;
};

LQT_DEVINL float epi_resid(const FkEpi& e, int m, int n) {
    float r = e.res_sm ? e.res_sm[(size_t)m * e.res_sm_stride + n] : ll_val1(e.res_ll + (size_t)m * e.res_ll_stride + n);
    float pv[FK_NGRP_MAX];
#pragma unroll
    for (int g = 0; g < FK_NGRP_MAX; ++g)
        pv[g] = (g < e.n_part) ? ll_val1(e.po + ((size_t)m * e.n_part + g) * e.H + n) : 0.f;
#pragma unroll
    for (int g = 0; g < FK_NGRP_MAX; ++g) if (g < e.n_part) r += pv[g];
    return r;
}

LQT_DEVINL void wait_full(FkCtx& c, unsigned st) {
    const unsigned slot = st % FK_STAGES, par = (st / FK_STAGES) & 1u;
    if (!mbar_try_wait(&c.sh->full[slot], par)) {
        if (!wait_full_slow(&c.sh->full[slot], par, c.p->ctrl)) c.aborted = true;
    }
}

// Tensor-core GEMV of one phase. Operand A = this CTA's weight tiles as they land in the ring
// (K/64 tiles of r8 x 128 B, SWIZZLE_128B), operand B = c.bt, D = TMEM columns 0-7 (rows 0-63) and
// 8-15 (rows 64-127): row r of a 64-row block sits in TMEM lane (r % 16) + 32 * (r / 16).
// One thread issues everything; tcgen05.commit releases each ring stage and finally signals mma_done.
// Then thread (warp q < 4, lane l < 16) owns rows 16q + l and 64 + 16q + l and publishes them.
LQT_DEVINL void gemv_tc(FkCtx& c, const FkDesc& d, int M, const FkEpi& e) {
    const int ntile = d.K >> 6;
    const int nst = d.nrows > 0 ? (ntile + d.tps - 1) / d.tps : 0;
    const bool two = d.r8 > 64;
    if (nst > 0) {
        // All eight consumer warps issue: warp w takes K-step (w & 3) of the tiles with parity (w >> 2). A
        // tcgen05.mma costs its issuing warp ~140 cycles (measured), whatever surrounds it, while MMAs of
        // different warps overlap, so the issue is spread as wide as possible. Control flow is warp-uniform,
        // one elected lane issues. Accumulator set = warp (its own TMEM columns): 8 independent chains.
        tc_fence_after();
        const int ks = c.warp & 3, par = c.warp >> 2;
        const uint32_t tile16 = ((uint32_t)d.r8 * 128u) >> 4;                 // tile size in 16-byte units
        const uint32_t dcol = c.tmem + (uint32_t)c.warp * 16u;
        const uint64_t bd0 = umma_desc_sw128(smem_u32(c.bt) + ks * 32);
        uint32_t acc = 0u;                                                     // first MMA of this warp overwrites
        int tile = 0;                                                          // first tile of the current stage
        const uint64_t astep = (uint64_t)(2u * tile16);
#pragma unroll 1
        for (int st = 0; st < nst; ++st) {
            const unsigned ast = c.stage_ctr + st, slot = ast % FK_STAGES;
            wait_full(c, ast);
            tc_fence_after();
            const int nt = min(d.tps, ntile - tile);
            const int t0 = ((tile & 1) == par) ? 0 : 1;                         // this warp's first tile inside the stage
            uint64_t ad = umma_desc_sw128(smem_u32(c.ring + (size_t)slot * FK_STAGE_BYTES) + ks * 32) + (uint64_t)((uint32_t)t0 * tile16);
            uint64_t bd = bd0 + (uint64_t)((uint32_t)(tile + t0) * 64u);
#pragma unroll 1
            for (int t = t0; t < nt; t += 2) {
                if (elect_one()) {
                    umma_bf16_m64n8k16(dcol, ad, bd, acc);
                    if (two) umma_bf16_m64n8k16(dcol + 8, ad + 512, bd, acc);   // rows 64..127: + 8192 B
                }
                acc = 1u;
                ad += astep; bd += 128;
            }
            tile += nt;
            if (elect_one()) umma_commit(&c.sh->empty[slot]);   // ring slot reusable once these MMAs have read it
            __syncwarp();
        }
        if (elect_one()) umma_commit(&c.sh->mma_done);
        __syncwarp();
    }
    c.stage_ctr += nst;
    if (nst == 0) return;
    // ---- epilogue ------------------------------------------------------------------------------
    const bool epi_thread = c.warp < 4;
    const int rA = 16 * c.warp + c.lane, rB = 64 + rA;       // rows of this thread (lanes < 16 only)
    const bool vA = epi_thread && c.lane < 16 && rA < d.nrows, vB = epi_thread && c.lane < 16 && two && rB < d.nrows;
    float resA[2] = {0.f, 0.f}, resB[2] = {0.f, 0.f};
    if (e.kind == EPI_RESID) {                               // overlaps the MMAs
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (m < M) {
                if (vA) resA[m] = epi_resid(e, m, d.row0 + rA);
                if (vB) resB[m] = epi_resid(e, m, d.row0 + rB);
            }
        }
    }
    {   // everybody waits: operand B and the TMEM accumulators are reused by the next phase
        if (!mbar_try_wait(&c.sh->mma_done, c.mma_phase)) {
            if (!wait_full_slow(&c.sh->mma_done, c.mma_phase, c.p->ctrl)) c.aborted = true;
        }
        c.mma_phase ^= 1u;
    }
    tc_fence_after();
    fk_mark(c, 5);
    if (epi_thread) {
        const int nset = (ntile >= 2) ? FK_NACC : 4;          // warps with tile parity 1 issue nothing when there is one tile
        float yA[2] = {0.f, 0.f}, yB[2] = {0.f, 0.f};
#pragma unroll 1
        for (int s0 = 0; s0 < nset; s0 += 2) {
            uint32_t r0[16], r1[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]),
                           "=r"(r0[8]), "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15])
                         : "r"(c.tmem + ((uint32_t)(32 * c.warp) << 16) + (uint32_t)s0 * 16u) : "memory");
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]),
                           "=r"(r1[8]), "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15])
                         : "r"(c.tmem + ((uint32_t)(32 * c.warp) << 16) + (uint32_t)(s0 + 1) * 16u) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                yA[m] += (__uint_as_float(r0[3 * m]) + __uint_as_float(r0[3 * m + 1])) + __uint_as_float(r0[3 * m + 2]);
                yB[m] += (__uint_as_float(r0[8 + 3 * m]) + __uint_as_float(r0[8 + 3 * m + 1])) + __uint_as_float(r0[8 + 3 * m + 2]);
                if (s0 + 1 < nset) {
                    yA[m] += (__uint_as_float(r1[3 * m]) + __uint_as_float(r1[3 * m + 1])) + __uint_as_float(r1[3 * m + 2]);
                    yB[m] += (__uint_as_float(r1[8 + 3 * m]) + __uint_as_float(r1[8 + 3 * m + 1])) + __uint_as_float(r1[8 + 3 * m + 2]);
                }
            }
        }
        tc_fence_before();
        fk_mark(c, 7);
        // SwiGLU: the gate row (even) needs the up row (odd) of the next lane
        float uA[2], uB[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            uA[m] = __shfl_down_sync(0xffffffffu, yA[m], 1);
            uB[m] = __shfl_down_sync(0xffffffffu, yB[m], 1);
        }
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (m < M) {
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    const bool valid = blk ? vB : vA;
                    if (!valid) continue;
                    const int n = d.row0 + (blk ? rB : rA);
                    float v = blk ? yB[m] : yA[m];
                    if (d.RG == 2) {
                        if ((n & 1) == 0) st_ll(e.out + (size_t)m * e.out_stride + (n >> 1), silu_f(v) * (blk ? uB[m] : uA[m]), c.seq);
                    } else if (e.kind == EPI_RESID) {
                        st_ll(e.out + (size_t)m * e.out_stride + n, (blk ? resB[m] : resA[m]) + v, c.seq);
                    } else {
                        if (e.bias) v += __ldg(e.bias + n);
                        st_ll(e.out + (size_t)m * e.out_stride + n, v, c.seq);
                        if (e.kind == EPI_LOGITS) e.plain[n] = v;
                    }
                }
            }
        }
    }
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

// talker: combine the splits of group g -> xs[0][0..rep*128)  (input of the grouped O-projection)
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
    xb_store1(c.bt, 0, c.tid, num / den);
    fence_proxy_async_smem();
    csync();
}

// code predictor: full attention of kv group g for the M new positions p0.., result -> xs[m][0..256)
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
        xb_store1(c.bt, m, c.tid, o);
    }
    fence_proxy_async_smem();
    csync();
}

// ------------------------------------------------------------------------------------------------
// one token pass (M rows) through a stack, as ONE loop over the flat op schedule so that every helper
// is instantiated exactly once (the context stays in registers; no local memory on the hot path).
// The layer-0 input rows are in c.res0 (smem, plain layout, stride = talker hidden).
// ------------------------------------------------------------------------------------------------
struct FkPass {
    bool is_cp; int M, pos0;
    const bf16_t* head_w; int head_n;
};

LQT_DEVINL void consume_token(FkCtx& c, const FkParams& p, const FkPass& ps) {
    const bool is_cp = ps.is_cp;
    const FkStack& S = is_cp ? p.cp : p.talker;
    const int tk = is_cp ? 1 : 0, M = ps.M;
    const int H = S.H, qd = S.heads * ATT_D, kvd = S.kv_heads * ATT_D, qkv_dim = qd + 2 * kvd;
    const int n_kv = S.kv_heads;
    const int H0 = p.talker.H;                                   // width of the rows in res0
    const bool inproj = is_cp && p.c_inproj_w != nullptr;
    const int total = pass_ops(S.n_layers, inproj, ps.head_w != nullptr);
    for (int it = 0; it < total && !c.aborted; ++it) {
        const FkOp op = pass_op(it, S.n_layers, inproj);
        const int kind = op.kind, l = op.layer;
        if (kind == FKT_B && is_cp) continue;                 // the predictor's attention lives inside phase C
        ++c.seq;
        const unsigned want = c.seq - 1;
        fk_phase(c, tk, kind);
        fk_mark(c, 0);
        const FkLayer& L = is_cp ? p.c_layers[l] : p.t_layers[l];
        const FkDesc d = c.sh->desc[tk][kind];
        // the layer input rows: res0 (smem) for layer 0 without in_proj, else an LL buffer
        const bool in_res0 = (l == 0 && !inproj);
        const uint2* lin = (l == 0) ? p.cxin : S.x;           // (only read when !in_res0)
        if (kind == FKT_B) {
            if (p.kv_f32) talker_attn_partial<float>(c, L, l, ps.pos0, want);
            else          talker_attn_partial<bf16_t>(c, L, l, ps.pos0, want);
            fk_mark(c, 3);
            if (c.sh->aborted) { c.aborted = true; break; }
            continue;
        }
        FkStage sg{nullptr, 0, false, nullptr, 0, nullptr, 0, nullptr, nullptr, nullptr};
        FkEpi e{EPI_STORE, nullptr, 0, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, 0, 0};
        int sM = M, gM = M;
        bool do_stage = true;
        switch (kind) {
            case FKT_INPROJ:
                sg.sm = c.res0; sg.sm_stride = H0;
                e.out = p.cxin; e.out_stride = H; e.bias = p.c_inproj_b;
                break;
            case FKT_A:
                if (in_res0) { sg.sm = c.res0; sg.sm_stride = H0; }
                else { sg.ll = lin; sg.ll_stride = H; }
                sg.nw = L.ln1;
                e.out = S.qkv; e.out_stride = qkv_dim;
                break;
            case FKT_C:
                do_stage = false;
                e.out = S.po + (size_t)(c.cta % n_kv) * H; e.out_stride = n_kv * H;
                break;
            case FKT_D:
                if (in_res0) { sg.sm = c.res0; sg.sm_stride = H0; }
                else { sg.ll = lin; sg.ll_stride = H; sg.verified = true; }
                sg.part = S.po; sg.n_part = n_kv; sg.nw = L.ln2;
                e.kind = EPI_GLU; e.out = S.act; e.out_stride = S.inter;
                break;
            case FKT_E:
                sg.ll = S.act; sg.ll_stride = S.inter;
                e.kind = EPI_RESID; e.out = S.x; e.out_stride = H;
                if (in_res0) { e.res_sm = c.res0; e.res_sm_stride = H0; }
                else { e.res_ll = lin; e.res_ll_stride = H; }
                e.po = S.po; e.n_part = n_kv; e.H = H;
                break;
            default:   // FKT_HEAD: final norm of the LAST row + head
                sg.ll = S.x + (size_t)(M - 1) * H; sg.ll_stride = H; sM = 1; gM = 1;
                sg.nw = S.final_norm;
                if (!is_cp) { sg.copy_sm = c.lh; sg.copy_gl = p.last_hidden; }
                e.kind = EPI_LOGITS; e.out = is_cp ? p.clogits_ll : p.logits_ll; e.out_stride = 0;
                e.plain = is_cp ? p.clogits : p.logits;
                break;
        }
        if (kind == FKT_C) {
            if (is_cp) cp_attn_local(c, L, l, M, ps.pos0, want);
            else       talker_attn_combine(c, ps.pos0, want);
        }
        if (do_stage) stage_rows(c, sg, sM, d.K, want, p.eps);
        fk_mark(c, 3);
        if (c.sh->aborted) { c.aborted = true; break; }
        gemv_tc(c, d, gM, e);
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
struct FkSmemLayout { size_t scratch, att, nxt, res0, lh, shared, total; };
inline FkSmemLayout fk_smem_layout(int maxV, int maxK, int H, int res0_floats) {
    FkSmemLayout L{};
    auto up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
    size_t off = (size_t)FK_STAGES * FK_STAGE_BYTES;
    L.scratch = off;                                           // operand B (maxK * 16 bytes) | sampler scratch
    const size_t samp = (size_t)maxV * (4 + 4 + 4 + 2 + 2), bt = (size_t)maxK * 16;
    off += up(samp > bt ? samp : bt);
    L.att = off; off += up((size_t)FA_FLOATS * 4);
    L.nxt = off; off += up((size_t)H * 4);
    L.res0 = off; off += up((size_t)res0_floats * 4);
    L.lh = off; off += up((size_t)H * 4);
    L.shared = off; off += up(sizeof(FkShared));
    L.total = off;
    return L;
}

struct FkSmemOffsets { unsigned scratch, att, nxt, res0, lh, shared; int maxV; };

__global__ void __launch_bounds__(FK_THREADS, 1)
frame_kernel(const __grid_constant__ FkParams p, const FkSmemOffsets so) {
    extern __shared__ __align__(1024) unsigned char fk_smem[];
    FkShared* sh = reinterpret_cast<FkShared*>(fk_smem + so.shared);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    if (tid == 0) {
        for (int i = 0; i < FK_STAGES; ++i) { mbar_init(&sh->full[i], 1); mbar_init(&sh->empty[i], FK_CWARPS); }   // every consumer warp issues MMAs and commits
        mbar_init(&sh->mma_done, FK_CWARPS);
        sh->stop = 0; sh->consumed = 0; sh->aborted = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 20) sh->desc[tid / 10][tid % 10] = make_desc(p, tid >= 10, tid % 10, cta, ncta);
    if (warp == 0) {                               // TMEM: FK_NACC accumulator sets of the tensor-core GEMV
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "n"(FK_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

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
                n_pass = (long long)n_prefill + (long long)nf * (p.cp_steps + 1);
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
                    const int ntile = d.K >> 6;
                    const uint32_t tile_bytes = (uint32_t)d.r8 * 128u;
                    const char* srcb = reinterpret_cast<const char*>(W + d.img_off);
                    for (int t0i = 0; t0i < ntile && !stopped; t0i += d.tps) {
                        const int nt = min(d.tps, ntile - t0i);
                        const unsigned slot = issued % FK_STAGES, par = ((issued / FK_STAGES) & 1u) ^ 1u;
                        unsigned long long t0 = 0;
                        while (!mbar_try_wait(&sh->empty[slot], par)) {
                            if (sh->stop) { stopped = true; break; }
                            if (t0 == 0) t0 = clock64();
                            else if (clock64() - t0 > FK_SPIN_LIMIT) { stopped = true; break; }
                        }
                        if (stopped) break;
                        const uint32_t bytes = (uint32_t)nt * tile_bytes;
                        mbar_expect_tx(&sh->full[slot], bytes);
                        bulk_g2s(fk_smem + (size_t)slot * FK_STAGE_BYTES, srcb + (size_t)t0i * tile_bytes, bytes, &sh->full[slot]);
                        ++issued;
                    }
                }
            }
            // drain: every issued copy must land before the CTA may exit
            unsigned long long t0 = clock64();
            while (!sh->stop) { if (clock64() - t0 > 4 * FK_SPIN_LIMIT) break; __nanosleep(200); }
            __threadfence_block();
            for (unsigned stg = (unsigned)sh->consumed; stg < issued; ++stg) {
                const unsigned slot = stg % FK_STAGES, par = (stg / FK_STAGES) & 1u;
                unsigned long long t1 = clock64();
                while (!mbar_try_wait(&sh->full[slot], par)) { if (clock64() - t1 > FK_SPIN_LIMIT) break; }
            }
        }
        return;
    }

    // ================================ consumer warps ===============================================
    FkCtx c;
    c.p = &p; c.sh = sh; c.ring = fk_smem;
    c.bt = fk_smem + so.scratch;
    c.att = reinterpret_cast<float*>(fk_smem + so.att);
    c.nxt = reinterpret_cast<float*>(fk_smem + so.nxt);
    c.res0 = reinterpret_cast<float*>(fk_smem + so.res0);
    c.lh = reinterpret_cast<float*>(fk_smem + so.lh);
    c.tid = tid; c.lane = lane; c.warp = warp; c.cta = cta; c.ncta = ncta;
    c.seq = 0; c.stage_ctr = 0; c.aborted = false; c.mma_phase = 0; c.tmem = sh->tmem_base;
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
            ps = FkPass{false, 1, pos, head ? p.t_head : nullptr, p.vocab};
        } else {
            if (done || frame >= frame_end) break;
            // ---- draw codebook cb of this frame (:803-812 for cb 0, :863-864 otherwise) -----------------
            const int tk = cb ? 1 : 0;
            fk_phase(c, tk, FKT_SAMPLE);
            fk_mark(c, 0);
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
                if (cb == 0) {                                                  // rows [last_hidden, codec_embed(code0)] (:854-860)
                    reinterpret_cast<float4*>(c.res0)[k4] = reinterpret_cast<const float4*>(c.lh)[k4];
                    reinterpret_cast<float4*>(c.res0 + H)[k4] = e;
                } else {
                    reinterpret_cast<float4*>(c.res0)[k4] = last_cb ? acc : e;  // :867-868 / :845
                }
                if (last_cb && cta == 0) reinterpret_cast<float4*>(p.next_in)[k4] = acc;
            }
            csync();
            fk_mark(c, 4);
            if (last_cb) {
                n_frames = frame + 1;
                ps = FkPass{false, 1, pos, p.t_head, p.vocab};                  // :845
            } else {
                ps = FkPass{true, cb == 0 ? 2 : 1, cb == 0 ? 0 : cb + 1, p.c_heads + (size_t)cb * p.cp_vocab * p.cp.H, p.cp_vocab};
            }
        }
        consume_token(c, p, ps);
        if (c.aborted) break;
        if (p.mode == 1) { pos += 1; mode1_done = true; }
        else if (prefill_i < n_prefill) { pos += 1; ++prefill_i; }
        else if (cb == p.cp_steps) { pos += 1; frame += 1; cb = 0; }
        else ++cb;
    }
    if (p.mode == 0 && !done && frame >= st0.max_frames) done = 1;
    // ---- exit: publish state, stop the producer, release TMEM -----------------------------------------
    tc_fence_before();
    csync();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "n"(FK_TMEM_COLS) : "memory");
    }
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
