// Persistent frame kernel: loops A and B of the reference (src/tts_onnx.cpp:782-872) as ONE
// cooperative launch per utterance (or per chunk of frames).
//
// Why one kernel: a frame is 31 dependent network passes (1 talker step + 15 predictor passes, each
// followed by a draw) = ~460 dependent matrix-vector phases of a few microseconds each. As separate
// launches the HBM pipe drains at every kernel boundary. Here every SM keeps one CTA resident:
//   * warp 8 (producer) walks the static weight schedule of the whole frame and streams this CTA's
//     slice of every matrix into a shared-memory ring with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), running AHEAD of the math across phase boundaries, so
//     HBM stays busy while the consumers sit in a grid barrier;
//   * warps 0-7 (consumers) run the phases: assemble the input vector (RMSNorm / attention / partial
//     sums), dot it with the rows in the ring (x in registers, weights read once from smem), write
//     their few outputs, and meet the other CTAs at a release/acquire grid barrier;
//   * the sampler and the embedding glue (src/tts_onnx.cpp:803-842, 854-868, 878-950) run redundantly
//     in every CTA, so a draw costs no extra barrier and no broadcast.
// Phases per layer: QKV | [talker: split-KV attention] | O-projection by kv-group (the code predictor
// computes its <=17-position attention inside this phase) | gate/up (SwiGLU) | down.
// All cross-CTA activations are read with ld.global.cg (L2), never through L1.
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
constexpr int FK_STAGES = 9;                      // 144 KB weight ring per SM
constexpr int FK_NS_MAX = 24;                     // max CTAs per kv group (attention splits)
constexpr int FK_NGRP_MAX = 8;                    // kv groups
constexpr int FK_MAXV = 4096;
constexpr int FK_CP_POS = 32;                     // code-predictor KV capacity (positions)
constexpr unsigned long long FK_SPIN_LIMIT = 6000000000ull;   // ~3 s of SM clocks: abort, never hang

struct FkLayer {
    const bf16_t* wqkv;     // [q_dim + 2 kv_dim][H]
    const bf16_t* wo_g;     // [n_kv][H][rep*128]   (O-projection regrouped by kv group)
    const bf16_t* wgu;      // [2*inter][H]         (gate row n, up row n interleaved)
    const bf16_t* wdown;    // [H][inter]
    const float *ln1, *ln2, *qnorm, *knorm;
};

constexpr int FK_MAX_TLAYERS = 32, FK_MAX_CLAYERS = 8;

struct FkStack {
    int n_layers, H, heads, kv_heads, inter;
    const float *cos, *sin, *final_norm;
    float *x, *xmid, *qkv, *po, *act;     // [2][H] [2][H] [2][qkv_dim] [2][n_kv][H] [2][inter]
};

struct FkParams {
    FkStack talker, cp;
    FkLayer t_layers[FK_MAX_TLAYERS];     // in the kernel-parameter constant bank: no load latency
    FkLayer c_layers[FK_MAX_CLAYERS];
    const bf16_t* t_head; int vocab;
    const bf16_t* c_heads; int cp_vocab, cp_steps;
    const bf16_t* c_inproj_w; const float* c_inproj_b; float* cxin;     // 1.7B: talker width -> predictor width
    float eps;
    void* kv_pool; const int* page_table; int page_shift; long long page_stride; int kv_f32;
    float* pa;                 // talker attention partials [n_kv][FK_NS_MAX][2][ATT_PSTRIDE]
    float* cp_kv;              // [layer][k|v][n_kv][FK_CP_POS][128] fp32
    float *logits, *clogits, *last_hidden;
    float *cp_in, *next_in;    // layer-0 inputs of the predictor pass [2][H] / talker step [H] (written by CTA 0)
    const bf16_t *codec_embed, *cp_embed;
    const float* prompt; int P;           // prefill rows (run when st->pos == 0)
    const float *trailing, *tts_pad;
    GenState* st; const SamplingDev* sp;
    long long* codes_out; const long long* forced; float* trace; int trace_stride;
    unsigned* ctrl;            // [0] grid-barrier counter, [1] abort flag (both zero at launch)
    int frame_end;             // run frames while frame < frame_end (<= max_frames)
    int mode;                  // 0 = prefill (if pos == 0) + frames; 1 = one talker token from next_in (head on), no frames
    unsigned long long* dbg;   // nullable: phase timeline of CTA dbg_cta, entries (clock64 << 8 | tag), [0] = count
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
LQT_DEVINL unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
LQT_DEVINL void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
LQT_DEVINL unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
LQT_DEVINL void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
LQT_DEVINL void csync() { asm volatile("bar.sync 1, %0;" ::"n"(FK_CTHREADS) : "memory"); }
LQT_DEVINL uint4 lds128(const void* p) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)));
    return r;
}

// ------------------------------------------------------------------------------------------------
// shared-memory layout
// ------------------------------------------------------------------------------------------------
struct FkShared {
    uint64_t full[FK_STAGES];
    uint64_t empty[FK_STAGES];
    volatile int stop;            // consumers -> producer: stop issuing
    volatile int consumed;        // stages consumed when stop was raised
    int aborted;
    int hist[256];
    int wtot[FK_CWARPS];
    float redf[FK_CWARPS][2];
    int redi[FK_CWARPS];
    uint32_t sel_prefix; int sel_k;
    int tok; float fsum;
    float rstd[2];
};

// x vectors live in smem as float4, permuted inside every 256-float chunk so that a lane's two
// float4 reads (elements lane*8 .. lane*8+7 of the chunk) are bank-conflict free.
LQT_DEVINL int xs_perm4(int k4) { return (k4 & ~63) + ((k4 & 1) << 5) + ((k4 & 63) >> 1); }
LQT_DEVINL int xs_idx(int k) { return xs_perm4(k >> 2) * 4 + (k & 3); }

struct FkCtx {
    const FkParams* p;
    FkShared* sh;
    unsigned char* ring;      // FK_STAGES * FK_STAGE_BYTES
    float* xs;                // x staging: [M][Kpad] (permuted)        | aliases the sampler scratch
    float* att;               // attention scratch                        |
    float* nxt;               // running next talker input [H]
    int tid, lane, warp;
    int cta, ncta;
    unsigned gen;             // grid-barrier generation (arrivals so far)
    unsigned stage_ctr;       // ring stages consumed so far
    bool aborted;
    unsigned long long* dbg; int dbg_n, dbg_cap, dbg_tag;   // dbg_tag = (stack << 9) | (kind << 4) of the current phase
};

// timeline entries: (SM clock << 16) | (stack << 9) | (phase kind << 4) | point. Points:
//  0 phase begin   1 descriptor ready (before the grid wait)   2 grid wait done   3 attention / rows staged
//  4 RMSNorm done  5 first weight stage landed (inside the GEMV)   6 this warp's GEMV rows done
//  7 all warps done (CTA barrier inside the arrive)   8 release-add issued
enum { FKT_A = 1, FKT_B = 2, FKT_C = 3, FKT_D = 4, FKT_E = 5, FKT_HEAD = 6, FKT_SAMPLE = 7, FKT_INPROJ = 8, FKT_GLUE = 9 };
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
LQT_DEVINL int rows_per_stage(int K, int RG) {
    int r = FK_STAGE_BYTES / (K * 2);
    r = (r / RG) * RG;
    return r < RG ? RG : r;       // host guarantees RG * K * 2 <= FK_STAGE_BYTES
}

// ------------------------------------------------------------------------------------------------
// producer: stream this CTA's slices in program order
// ------------------------------------------------------------------------------------------------
// flat schedule of one token pass: [in_proj] + n_layers x (A qkv, B attention, C o-proj, D gate/up, E down) + [head]
struct FkOp { int kind, layer; };
LQT_DEVINL int pass_ops(int n_layers, bool inproj, bool head) { return (inproj ? 1 : 0) + n_layers * 5 + (head ? 1 : 0); }
LQT_DEVINL FkOp pass_op(int it, int n_layers, bool inproj) {
    const int n_pre = inproj ? 1 : 0;
    if (it < n_pre) return FkOp{8 /*FKT_INPROJ*/, 0};
    const int r = it - n_pre;
    if (r < n_layers * 5) return FkOp{1 + r % 5, r / 5};
    return FkOp{6 /*FKT_HEAD*/, 0};
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
// grid barrier (all CTAs co-resident: cooperative launch)
// ------------------------------------------------------------------------------------------------
LQT_DEVINL void gbar_arrive(FkCtx& c) {
    csync();                                   // this CTA's global writes are ordered before the release
    fk_mark(c, 7);
    ++c.gen;
    if (c.tid == 0) red_release_add(&c.p->ctrl[0], 1u);
    fk_mark(c, 8);
}
LQT_DEVINL void gbar_wait(FkCtx& c) {
    if (c.tid == 0) {
        const unsigned target = c.gen * (unsigned)c.ncta;
        unsigned long long t0 = 0;
        int it = 0;
        while (ld_relaxed_u32(&c.p->ctrl[0]) < target) {
            if ((++it & 255) == 0) {
                if (ld_relaxed_u32(&c.p->ctrl[1]) != 0) { c.sh->aborted = 1; break; }
                if (t0 == 0) t0 = clock64();
                else if (clock64() - t0 > FK_SPIN_LIMIT) { atomicExch(&c.p->ctrl[1], 1u); c.sh->aborted = 1; break; }
            }
        }
        fence_acq_rel_gpu();
    }
    csync();
    if (c.sh->aborted) c.aborted = true;
}

// ------------------------------------------------------------------------------------------------
// input staging
// ------------------------------------------------------------------------------------------------
// xs[m][:] = src[m*stride + :] (+ sum of n_part partial vectors at part[(m*n_part + g)*K + :])
LQT_DEVINL void stage_rows(FkCtx& c, const float* src, int stride, int M, int K, const float* part, int n_part,
                           float* copy_out /* nullable: CTA 0 writes the assembled rows [M][K] */) {
    const int K4 = K >> 2, Kpad4 = ((K + 255) & ~255) >> 2;
    float4* xs4 = reinterpret_cast<float4*>(c.xs);
    for (int m = 0; m < M; ++m) {
        for (int k4 = c.tid; k4 < Kpad4; k4 += FK_CTHREADS) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k4 < K4) {
                v = __ldcg(reinterpret_cast<const float4*>(src + (size_t)m * stride) + k4);
                for (int g = 0; g < n_part; ++g) {
                    const float4 a = __ldcg(reinterpret_cast<const float4*>(part + ((size_t)m * n_part + g) * K) + k4);
                    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
                }
                if (copy_out && c.cta == 0) reinterpret_cast<float4*>(copy_out + (size_t)m * K)[k4] = v;
            }
            xs4[m * Kpad4 + xs_perm4(k4)] = v;
        }
    }
    csync();
}

// in-place RMSNorm of the staged rows: xs = (xs * rstd) * w ; optional copy of the result (CTA 0)
LQT_DEVINL void norm_rows(FkCtx& c, const float* w, int M, int K, float eps, float* copy_out) {
    const int K4 = K >> 2, Kpad4 = ((K + 255) & ~255) >> 2;
    float4* xs4 = reinterpret_cast<float4*>(c.xs);
    float ss[2] = {0.f, 0.f};
    for (int k4 = c.tid; k4 < K4; k4 += FK_CTHREADS) {
        const int d = xs_perm4(k4);
        for (int m = 0; m < M; ++m) {
            const float4 v = xs4[m * Kpad4 + d];
            ss[m] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
    }
    for (int m = 0; m < M; ++m) {
        const float s = warp_sum(ss[m]);
        if (c.lane == 0) c.sh->redf[c.warp][m] = s;
    }
    csync();
    if (c.tid < M) {
        float s = 0.f;
        for (int w2 = 0; w2 < FK_CWARPS; ++w2) s += c.sh->redf[w2][c.tid];
        c.sh->rstd[c.tid] = 1.0f / sqrtf(s / (float)K + eps);
    }
    csync();
    for (int k4 = c.tid; k4 < K4; k4 += FK_CTHREADS) {
        const int d = xs_perm4(k4);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + k4);
        for (int m = 0; m < M; ++m) {
            float4 v = xs4[m * Kpad4 + d];
            const float r = c.sh->rstd[m];
            v.x = (v.x * r) * wv.x; v.y = (v.y * r) * wv.y; v.z = (v.z * r) * wv.z; v.w = (v.w * r) * wv.w;
            xs4[m * Kpad4 + d] = v;
            if (copy_out && c.cta == 0) reinterpret_cast<float4*>(copy_out + (size_t)m * K)[k4] = v;
        }
    }
    csync();
}

// ------------------------------------------------------------------------------------------------
// GEMV over this CTA's slice, weights from the ring
// ------------------------------------------------------------------------------------------------
enum { EPI_STORE = 0, EPI_GLU = 1, EPI_RESID = 2 };
struct FkEpi {
    int kind;
    float* out; int out_stride;          // out[m*out_stride + n]   (GLU: n/2)
    const float* bias; int act;          // STORE only
    const float* resid; int resid_stride;
};

LQT_DEVINL void epi_store(const FkEpi& e, int m, int n, float v, float v2) {
    if (e.kind == EPI_GLU) {
        e.out[(size_t)m * e.out_stride + (n >> 1)] = silu_f(v) * v2;
    } else if (e.kind == EPI_RESID) {
        e.out[(size_t)m * e.out_stride + n] = __ldcg(e.resid + (size_t)m * e.resid_stride + n) + v;
    } else {
        if (e.bias) v += __ldg(e.bias + n);
        if (e.act == 1) v = silu_f(v);
        e.out[(size_t)m * e.out_stride + n] = v;
    }
}

LQT_DEVINL void wait_full(FkCtx& c, unsigned st) {
    const unsigned slot = st % FK_STAGES, par = (st / FK_STAGES) & 1u;
    unsigned long long t0 = 0;
    while (!mbar_try_wait(&c.sh->full[slot], par)) {
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > FK_SPIN_LIMIT) { c.aborted = true; atomicExch(&c.p->ctrl[1], 1u); break; }
    }
}

// x in registers: KC chunks of 256 elements, M rows. RG rows per group (GLU: gate, up).
template <int KC, int M, int RG>
LQT_DEVINL void gemv_reg(FkCtx& c, int row0, int nrows, const FkEpi& e) {
    constexpr int K = KC * 256;
    const int Kpad4 = K >> 2;
    const float4* xs4 = reinterpret_cast<const float4*>(c.xs);
    float4 xa[M][KC], xb[M][KC];
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int j = 0; j < KC; ++j) {
            xa[m][j] = xs4[m * Kpad4 + j * 64 + c.lane];
            xb[m][j] = xs4[m * Kpad4 + j * 64 + 32 + c.lane];
        }
    const int rps = rows_per_stage(K, RG);
    const int nst = (nrows + rps - 1) / rps;
    for (int st = 0; st < nst; ++st) {
        const unsigned ast = c.stage_ctr + st;
        wait_full(c, ast);
        if (st == 0) fk_mark(c, 5);
        const unsigned char* base = c.ring + (size_t)(ast % FK_STAGES) * FK_STAGE_BYTES;
        const int rs = min(rps, nrows - st * rps);          // rows in this stage
        for (int g = c.warp; g * RG < rs; g += FK_CWARPS) {
            float acc[RG][M];
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int m = 0; m < M; ++m) acc[r][m] = 0.f;
#pragma unroll
            for (int r = 0; r < RG; ++r) {
                const unsigned char* rowp = base + (size_t)(g * RG + r) * (K * 2) + c.lane * 16;
                constexpr int JB = KC < 4 ? KC : 4;          // weight loads in flight per lane
#pragma unroll
                for (int j0 = 0; j0 < KC; j0 += JB) {
                    uint4 w[JB];
#pragma unroll
                    for (int j = 0; j < JB; ++j) w[j] = lds128(rowp + (j0 + j) * 512);
#pragma unroll
                    for (int j = 0; j < JB; ++j)
#pragma unroll
                        for (int m = 0; m < M; ++m) acc[r][m] = dot8(w[j], xa[m][j0 + j], xb[m][j0 + j], acc[r][m]);
                }
            }
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int m = 0; m < M; ++m) acc[r][m] = warp_sum(acc[r][m]);
            if (c.lane == 0) {
                const int n = row0 + st * rps + g * RG;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    if (RG == 2) epi_store(e, m, n, acc[0][m], acc[RG - 1][m]);
                    else epi_store(e, m, n, acc[0][m], 0.f);
                }
            }
        }
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.sh->empty[ast % FK_STAGES]);
    }
    c.stage_ctr += nst;
}

// generic fallback: x read from smem inside the loop (any K % 256 == 0, M <= 2)
template <int RG>
LQT_DEVINL void gemv_smem(FkCtx& c, int K, int M, int row0, int nrows, const FkEpi& e) {
    const int KC = K >> 8, Kpad4 = K >> 2;
    const float4* xs4 = reinterpret_cast<const float4*>(c.xs);
    const int rps = rows_per_stage(K, RG);
    const int nst = (nrows + rps - 1) / rps;
    for (int st = 0; st < nst; ++st) {
        const unsigned ast = c.stage_ctr + st;
        wait_full(c, ast);
        const unsigned char* base = c.ring + (size_t)(ast % FK_STAGES) * FK_STAGE_BYTES;
        const int rs = min(rps, nrows - st * rps);
        for (int g = c.warp; g * RG < rs; g += FK_CWARPS) {
            float acc[RG][2];
#pragma unroll
            for (int r = 0; r < RG; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
            for (int j = 0; j < KC; ++j) {
                const float4 a0 = xs4[j * 64 + c.lane], b0 = xs4[j * 64 + 32 + c.lane];
                float4 a1 = a0, b1 = b0;
                if (M > 1) { a1 = xs4[Kpad4 + j * 64 + c.lane]; b1 = xs4[Kpad4 + j * 64 + 32 + c.lane]; }
#pragma unroll
                for (int r = 0; r < RG; ++r) {
                    const uint4 w = lds128(base + (size_t)(g * RG + r) * ((size_t)K * 2) + c.lane * 16 + j * 512);
                    acc[r][0] = dot8(w, a0, b0, acc[r][0]);
                    if (M > 1) acc[r][1] = dot8(w, a1, b1, acc[r][1]);
                }
            }
#pragma unroll
            for (int r = 0; r < RG; ++r) { acc[r][0] = warp_sum(acc[r][0]); if (M > 1) acc[r][1] = warp_sum(acc[r][1]); }
            if (c.lane == 0) {
                const int n = row0 + st * rps + g * RG;
                for (int m = 0; m < M; ++m) {
                    if (RG == 2) epi_store(e, m, n, acc[0][m], acc[RG - 1][m]);
                    else epi_store(e, m, n, acc[0][m], 0.f);
                }
            }
        }
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.sh->empty[ast % FK_STAGES]);
    }
    c.stage_ctr += nst;
}

LQT_DEVINL void gemv_phase(FkCtx& c, int K, int M, int RG, int row0, int nrows, const FkEpi& e) {
    if (nrows <= 0) return;
    const int KC = K >> 8;
    if (RG == 2) {                      // SwiGLU pairs: K = hidden
        if (M == 1) {
            switch (KC) {
                case 1: gemv_reg<1, 1, 2>(c, row0, nrows, e); return;
                case 4: gemv_reg<4, 1, 2>(c, row0, nrows, e); return;
                case 8: gemv_reg<8, 1, 2>(c, row0, nrows, e); return;
                default: break;
            }
        } else {
            switch (KC) {
                case 1: gemv_reg<1, 2, 2>(c, row0, nrows, e); return;
                case 4: gemv_reg<4, 2, 2>(c, row0, nrows, e); return;
                default: break;
            }
        }
        gemv_smem<2>(c, K, M, row0, nrows, e);
        return;
    }
    if (M == 1) {
        switch (KC) {
            case 1:  gemv_reg<1, 1, 1>(c, row0, nrows, e); return;
            case 2:  gemv_reg<2, 1, 1>(c, row0, nrows, e); return;
            case 4:  gemv_reg<4, 1, 1>(c, row0, nrows, e); return;
            case 8:  gemv_reg<8, 1, 1>(c, row0, nrows, e); return;
            case 12: gemv_reg<12, 1, 1>(c, row0, nrows, e); return;
            default: break;
        }
    } else {
        switch (KC) {
            case 1: gemv_reg<1, 2, 1>(c, row0, nrows, e); return;
            case 2: gemv_reg<2, 2, 1>(c, row0, nrows, e); return;
            case 4: gemv_reg<4, 2, 1>(c, row0, nrows, e); return;
            default: break;
        }
    }
    gemv_smem<1>(c, K, M, row0, nrows, e);
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

// talker: split-KV partial attention of kv group g over this CTA's chunk of positions
template <typename KVT>
LQT_DEVINL void talker_attn_partial(FkCtx& c, const FkLayer& L, int layer, int t) {
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
        float4 v = __ldcg(reinterpret_cast<const float4*>(S.qkv + (size_t)(g * 2 + c.warp) * ATT_D) + c.lane);
        v = head_norm_rope(v, L.qnorm, p.eps, cosr, sinr, c.lane);
        reinterpret_cast<float4*>(q_s + c.warp * ATT_D)[c.lane] = v;
    } else if (owns_new && c.warp < 4) {
        const long long base = (long long)p.page_table[t >> p.page_shift] * p.page_stride + layer_off + head_off +
                               (long long)(t & (PS - 1)) * ATT_D;
        if (c.warp == 2) {
            float4 v = __ldcg(reinterpret_cast<const float4*>(S.qkv + q_dim + (size_t)g * ATT_D) + c.lane);
            v = head_norm_rope(v, L.knorm, p.eps, cosr, sinr, c.lane);
            KvIO<KVT>::store4(pool + base + c.lane * 4, v);
            v.x = KvIO<KVT>::round(v.x); v.y = KvIO<KVT>::round(v.y); v.z = KvIO<KVT>::round(v.z); v.w = KvIO<KVT>::round(v.w);
            reinterpret_cast<float4*>(kn)[c.lane] = v;
        } else {
            float4 v = __ldcg(reinterpret_cast<const float4*>(S.qkv + q_dim + kv_dim + (size_t)g * ATT_D) + c.lane);
            KvIO<KVT>::store4(pool + base + v_off + c.lane * 4, v);
            v.x = KvIO<KVT>::round(v.x); v.y = KvIO<KVT>::round(v.y); v.z = KvIO<KVT>::round(v.z); v.w = KvIO<KVT>::round(v.w);
            reinterpret_cast<float4*>(vn)[c.lane] = v;
        }
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
        for (int w2 = 0; w2 < FK_CWARPS; ++w2) Mx = fmaxf(Mx, wm[w2 * 2 + r]);
        float num = 0.f, den = 0.f;
        for (int w2 = 0; w2 < FK_CWARPS; ++w2) {
            const float mw = wm[w2 * 2 + r];
            const float f = (mw == -INFINITY) ? 0.f : expf(mw - Mx);
            num = fmaf(f, wo[(w2 * 2 + r) * ATT_D + d], num);
            den = fmaf(f, wl[w2 * 2 + r], den);
        }
        float* part = p.pa + ((size_t)(g * FK_NS_MAX + s) * 2 + r) * ATT_PSTRIDE;
        part[d] = num;
        if (d == 0) { part[ATT_D] = Mx; part[ATT_D + 1] = den; }
    }
}

// talker: combine the splits of group g -> xs[0][0..rep*128)  (input of the grouped O-projection)
LQT_DEVINL void talker_attn_combine(FkCtx& c, int t) {
    const FkParams& p = *c.p;
    const int n_kv = p.talker.kv_heads, g = c.cta % n_kv, ns = min(grp_members(g, n_kv, c.ncta), FK_NS_MAX);
    const int n_pos = t + 1, chunk = (n_pos + ns - 1) / ns, active = (n_pos + chunk - 1) / chunk;
    const int r = c.tid >> 7, d = c.tid & 127;
    float Mx = -INFINITY;
    for (int s = 0; s < active; ++s)
        Mx = fmaxf(Mx, __ldcg(p.pa + ((size_t)(g * FK_NS_MAX + s) * 2 + r) * ATT_PSTRIDE + ATT_D));
    float num = 0.f, den = 0.f;
    for (int s = 0; s < active; ++s) {
        const float* ps = p.pa + ((size_t)(g * FK_NS_MAX + s) * 2 + r) * ATT_PSTRIDE;
        const float f = expf(__ldcg(ps + ATT_D) - Mx);
        num = fmaf(f, __ldcg(ps + d), num);
        den = fmaf(f, __ldcg(ps + ATT_D + 1), den);
    }
    c.xs[xs_idx(c.tid)] = num / den;
    csync();
}

// code predictor: full attention of kv group g for the M new positions p0.., result -> xs[m][0..256)
LQT_DEVINL void cp_attn_local(FkCtx& c, const FkLayer& L, int layer, int M, int p0) {
    const FkParams& p = *c.p;
    const FkStack& S = p.cp;
    const int n_kv = S.kv_heads, g = c.cta % n_kv, s = c.cta / n_kv;
    const int q_dim = S.heads * ATT_D, kv_dim = n_kv * ATT_D, qkv_dim = q_dim + 2 * kv_dim;
    float* q_s = c.att + FA_Q; float* kn = c.att + FA_KN; float* vn = c.att + FA_VN; float* sc = c.att + FA_SC;
    float* kc = p.cp_kv + ((size_t)(layer * 2 + 0) * n_kv + g) * FK_CP_POS * ATT_D;
    float* vc = p.cp_kv + ((size_t)(layer * 2 + 1) * n_kv + g) * FK_CP_POS * ATT_D;
    // warps 0..2M-1: q heads ; 2M..3M-1: k ; 3M..4M-1: v
    for (int job = c.warp; job < 4 * M; job += FK_CWARPS) {
        if (job < 2 * M) {
            const int m = job >> 1, r = job & 1, pos = p0 + m;
            float4 v = __ldcg(reinterpret_cast<const float4*>(S.qkv + (size_t)m * qkv_dim + (size_t)(g * 2 + r) * ATT_D) + c.lane);
            v = head_norm_rope(v, L.qnorm, p.eps, S.cos + (size_t)pos * (ATT_D / 2), S.sin + (size_t)pos * (ATT_D / 2), c.lane);
            reinterpret_cast<float4*>(q_s + (m * 2 + r) * ATT_D)[c.lane] = v;
        } else if (job < 3 * M) {
            const int m = job - 2 * M, pos = p0 + m;
            float4 v = __ldcg(reinterpret_cast<const float4*>(S.qkv + (size_t)m * qkv_dim + q_dim + (size_t)g * ATT_D) + c.lane);
            v = head_norm_rope(v, L.knorm, p.eps, S.cos + (size_t)pos * (ATT_D / 2), S.sin + (size_t)pos * (ATT_D / 2), c.lane);
            reinterpret_cast<float4*>(kn + m * ATT_D)[c.lane] = v;
            if (s == 0) reinterpret_cast<float4*>(kc + (size_t)pos * ATT_D)[c.lane] = v;
        } else {
            const int m = job - 3 * M, pos = p0 + m;
            const float4 v = __ldcg(reinterpret_cast<const float4*>(S.qkv + (size_t)m * qkv_dim + q_dim + kv_dim + (size_t)g * ATT_D) + c.lane);
            reinterpret_cast<float4*>(vn + m * ATT_D)[c.lane] = v;
            if (s == 0) reinterpret_cast<float4*>(vc + (size_t)pos * ATT_D)[c.lane] = v;
        }
    }
    csync();
    const float scale = 1.0f / sqrtf((float)ATT_D);
    // scores: combos (m, j) -> both heads
    const int n_last = p0 + M;                       // positions visible to the last row
    for (int cb = c.warp; cb < M * n_last; cb += FK_CWARPS) {
        const int m = cb / n_last, j = cb - m * n_last;
        if (j > p0 + m) continue;                    // causal
        const float4 k4 = (j < p0) ? __ldcg(reinterpret_cast<const float4*>(kc + (size_t)j * ATT_D) + c.lane)
                                   : reinterpret_cast<const float4*>(kn + (j - p0) * ATT_D)[c.lane];
        const float4 qa = reinterpret_cast<const float4*>(q_s + (m * 2 + 0) * ATT_D)[c.lane];
        const float4 qb = reinterpret_cast<const float4*>(q_s + (m * 2 + 1) * ATT_D)[c.lane];
        float d0 = k4.x * qa.x + k4.y * qa.y + k4.z * qa.z + k4.w * qa.w;
        float d1 = k4.x * qb.x + k4.y * qb.y + k4.z * qb.z + k4.w * qb.w;
        d0 = warp_sum(d0) * scale; d1 = warp_sum(d1) * scale;
        if (c.lane == 0) { sc[(m * 2 + 0) * FK_CP_POS + j] = d0; sc[(m * 2 + 1) * FK_CP_POS + j] = d1; }
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
    {
        const int r = c.tid >> 7, d = c.tid & 127;
        for (int m = 0; m < M; ++m) {
            const int np = p0 + m + 1;
            const float* row = sc + (m * 2 + r) * FK_CP_POS;
            float o = 0.f;
            for (int j = 0; j < np; ++j) {
                const float vv = (j < p0) ? __ldcg(vc + (size_t)j * ATT_D + d) : vn[(j - p0) * ATT_D + d];
                o = fmaf(row[j], vv, o);
            }
            c.xs[m * 256 + xs_idx(c.tid)] = o;       // K = 256 -> Kpad = 256
        }
    }
    csync();
}

// ------------------------------------------------------------------------------------------------
// one token pass (M rows) through a stack, as ONE loop over the flat op schedule so that every helper
// is instantiated exactly once (the context stays in registers; no local memory on the hot path).
// x0: layer-0 input rows in global memory [M][x0_stride] (also the layer-0 residual);
// x0_in_smem: the rows are already staged in xs (sampler glue).
// ------------------------------------------------------------------------------------------------
struct FkPass {
    bool is_cp; int M, pos0;
    const float* x0; int x0_stride; bool x0_in_smem;
    const bf16_t* head_w; int head_n; float* head_out; float* hidden_out;
};

LQT_DEVINL void consume_token(FkCtx& c, const FkParams& p, const FkPass& ps) {
    const bool is_cp = ps.is_cp;
    const FkStack& S = is_cp ? p.cp : p.talker;
    const int tk = is_cp ? 1 : 0, M = ps.M;
    const int H = S.H, qd = S.heads * ATT_D, kvd = S.kv_heads * ATT_D, qkv_dim = qd + 2 * kvd;
    const int rep = S.heads / S.kv_heads, gK = rep * ATT_D, n_kv = S.kv_heads;
    const bool inproj = is_cp && p.c_inproj_w != nullptr;
    const int total = pass_ops(S.n_layers, inproj, ps.head_w != nullptr);
    const float* lin = ps.x0; int lin_stride = ps.x0_stride;   // layer input rows (global)
    bool in_smem = ps.x0_in_smem;
    for (int it = 0; it < total && !c.aborted; ++it) {
        const FkOp op = pass_op(it, S.n_layers, inproj);
        const int kind = op.kind, l = op.layer;
        if (kind == FKT_B && is_cp) continue;                 // the predictor's attention lives inside phase C
        fk_phase(c, tk, kind);
        fk_mark(c, 0);
        const FkLayer& L = is_cp ? p.c_layers[l] : p.t_layers[l];
        // ---- what this phase stages and multiplies ---------------------------------------------
        const float* src = nullptr; int sstride = 0, sK = 0, sM = M, npart = 0;
        const float* part = nullptr; float* scopy = nullptr; bool do_stage = false;
        const float* nw = nullptr; float* ncopy = nullptr;
        int gK_ = 0, gM = M, rg = 1; FkSlice sl{0, 0};
        FkEpi e{EPI_STORE, nullptr, 0, nullptr, 0, nullptr, 0};
        bool need_wait = true;
        switch (kind) {
            case FKT_INPROJ:
                need_wait = false;
                do_stage = !in_smem; src = lin; sstride = lin_stride; sK = p.talker.H;
                gK_ = p.talker.H; sl = flat_slice(H, 1, c.cta, c.ncta);
                e = FkEpi{EPI_STORE, p.cxin, H, p.c_inproj_b, 0, nullptr, 0};
                break;
            case FKT_A:
                need_wait = !(l == 0 && !inproj);
                do_stage = !(l == 0 && in_smem); src = lin; sstride = lin_stride; sK = H;
                nw = L.ln1;
                gK_ = H; sl = flat_slice(qkv_dim, 1, c.cta, c.ncta);
                e = FkEpi{EPI_STORE, S.qkv, qkv_dim, nullptr, 0, nullptr, 0};
                break;
            case FKT_B:
                break;
            case FKT_C:
                gK_ = gK; sl = group_slice(H, c.cta, c.ncta, n_kv);
                e = FkEpi{EPI_STORE, S.po + (size_t)(c.cta % n_kv) * H, n_kv * H, nullptr, 0, nullptr, 0};
                break;
            case FKT_D:
                do_stage = true; src = lin; sstride = lin_stride; sK = H; part = S.po; npart = n_kv; scopy = S.xmid;
                nw = L.ln2;
                gK_ = H; rg = 2; sl = flat_slice(2 * S.inter, 2, c.cta, c.ncta);
                e = FkEpi{EPI_GLU, S.act, S.inter, nullptr, 0, nullptr, 0};
                break;
            case FKT_E:
                do_stage = true; src = S.act; sstride = S.inter; sK = S.inter;
                gK_ = S.inter; sl = flat_slice(H, 1, c.cta, c.ncta);
                e = FkEpi{EPI_RESID, S.x, H, nullptr, 0, S.xmid, H};
                break;
            default:   // FKT_HEAD: final norm of the LAST row + head
                do_stage = true; src = S.x + (size_t)(M - 1) * H; sstride = H; sK = H; sM = 1;
                nw = S.final_norm; ncopy = ps.hidden_out;
                gK_ = H; gM = 1; sl = flat_slice(ps.head_n, 1, c.cta, c.ncta);
                e = FkEpi{EPI_STORE, ps.head_out, ps.head_n, nullptr, 0, nullptr, 0};
                break;
        }
        fk_mark(c, 1);
        if (need_wait) { gbar_wait(c); if (c.aborted) break; }
        fk_mark(c, 2);
        if (kind == FKT_B) {
            if (p.kv_f32) talker_attn_partial<float>(c, L, l, ps.pos0);
            else          talker_attn_partial<bf16_t>(c, L, l, ps.pos0);
            fk_mark(c, 3);
            gbar_arrive(c);
            continue;
        }
        if (kind == FKT_C) {
            if (is_cp) cp_attn_local(c, L, l, M, ps.pos0);
            else       talker_attn_combine(c, ps.pos0);
        }
        if (do_stage) stage_rows(c, src, sstride, sM, sK, part, npart, scopy);
        fk_mark(c, 3);
        if (nw) norm_rows(c, nw, sM, sK, p.eps, ncopy);
        fk_mark(c, 4);
        gemv_phase(c, gK_, gM, rg, sl.row0, sl.nrows, e);
        fk_mark(c, 6);
        gbar_arrive(c);
        if (kind == FKT_E) { lin = S.x; lin_stride = H; in_smem = false; }
        if (kind == FKT_INPROJ) { lin = p.cxin; lin_stride = H; in_smem = false; }
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

LQT_DEVINL int fk_sample(FkCtx& c, const FkSampScratch& s, const float* logits, int V, int mask_lo, int mask_hi,
                         int mask_keep, const SamplingDev& sp, uint32_t frame, int codebook, float* trace_row) {
    FkShared* sh = c.sh;
    const bool temper = !sp.greedy && sp.temperature > 0.0f && sp.temperature != 1.0f;
    for (int i = c.tid; i < V; i += FK_CTHREADS) {
        float v = __ldcg(logits + i);
        if (i >= mask_lo && i < mask_hi && i != mask_keep) v = -INFINITY;
        if (trace_row) trace_row[i] = v;
        if (temper) v = v / sp.temperature;
        s.x[i] = v;
    }
    csync();
    // block max + argmax (lowest index on ties)
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = c.tid; i < V; i += FK_CTHREADS) {
        const float v = s.x[i];
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
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
    if (sp.greedy) return bi;
    const float mx = bv;

    // top-k threshold: 4 x 8-bit radix select of the k-th largest key
    float thr = -INFINITY;
    if (sp.top_k > 0 && sp.top_k < V) {
        if (c.tid == 0) { sh->sel_prefix = 0u; sh->sel_k = sp.top_k; }
        for (int shift = 24; shift >= 0; shift -= 8) {
            sh->hist[c.tid] = 0;
            csync();
            const uint32_t prefix = sh->sel_prefix;
            const int kk = sh->sel_k;
            const uint32_t himask = (shift == 24) ? 0u : (0xffffffffu << (shift + 8));
            for (int i = c.tid; i < V; i += FK_CTHREADS) {
                const uint32_t key = float_key(s.x[i]);
                if ((key & himask) == prefix) atomicAdd(&sh->hist[(key >> shift) & 255], 1);
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
    const int per = (V + FK_CTHREADS - 1) / FK_CTHREADS;
    const int i0 = c.tid * per, i1 = min(V, i0 + per);
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
    return sh->tok;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
struct FkSmemLayout { size_t ring, scratch, nxt, shared, total; };
inline FkSmemLayout fk_smem_layout(int maxV, int max_xs_floats, int H) {
    FkSmemLayout L{};
    auto up = [](size_t v) { return (v + 127) & ~(size_t)127; };
    L.ring = 0;
    size_t off = (size_t)FK_STAGES * FK_STAGE_BYTES;
    L.scratch = off;
    const size_t samp = (size_t)maxV * (4 + 4 + 4 + 2 + 2);
    const size_t xsatt = up((size_t)max_xs_floats * 4) + (size_t)FA_FLOATS * 4;
    off += up(samp > xsatt ? samp : xsatt);
    L.nxt = off; off += up((size_t)H * 4);
    L.shared = off; off += up(sizeof(FkShared));
    L.total = off;
    return L;
}

struct FkSmemOffsets { unsigned scratch, xs_bytes, nxt, shared; int maxV; };

__global__ void __launch_bounds__(FK_THREADS, 1)
frame_kernel(const __grid_constant__ FkParams p, const FkSmemOffsets so) {
    extern __shared__ __align__(1024) unsigned char fk_smem[];
    FkShared* sh = reinterpret_cast<FkShared*>(fk_smem + so.shared);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < FK_STAGES; ++i) { mbar_init(&sh->full[i], 1); mbar_init(&sh->empty[i], FK_CWARPS); }
        sh->stop = 0; sh->consumed = 0; sh->aborted = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const GenState st0 = *p.st;                    // written by the host before launch
    const int cta = blockIdx.x, ncta = gridDim.x;
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
                const int H = S.H, qd = S.heads * ATT_D, kvd = S.kv_heads * ATT_D, rep = S.heads / S.kv_heads;
                const bool inproj = id.is_cp && p.c_inproj_w != nullptr;
                const int total = pass_ops(S.n_layers, inproj, id.head);
                for (int it = 0; it < total && !stopped; ++it) {
                    const FkOp op = pass_op(it, S.n_layers, inproj);
                    if (op.kind == FKT_B) continue;
                    const FkLayer& L = id.is_cp ? p.c_layers[op.layer] : p.t_layers[op.layer];
                    const bf16_t* W; int K, RG = 1; FkSlice sl;
                    switch (op.kind) {
                        case FKT_INPROJ: W = p.c_inproj_w; K = p.talker.H; sl = flat_slice(H, 1, cta, ncta); break;
                        case FKT_A: W = L.wqkv; K = H; sl = flat_slice(qd + 2 * kvd, 1, cta, ncta); break;
                        case FKT_C: W = L.wo_g + (size_t)(cta % S.kv_heads) * H * (rep * ATT_D); K = rep * ATT_D;
                                    sl = group_slice(H, cta, ncta, S.kv_heads); break;
                        case FKT_D: W = L.wgu; K = H; RG = 2; sl = flat_slice(2 * S.inter, 2, cta, ncta); break;
                        case FKT_E: W = L.wdown; K = S.inter; sl = flat_slice(H, 1, cta, ncta); break;
                        default:
                            if (id.is_cp) { W = p.c_heads + (size_t)id.cb * p.cp_vocab * H; sl = flat_slice(p.cp_vocab, 1, cta, ncta); }
                            else          { W = p.t_head; sl = flat_slice(p.vocab, 1, cta, ncta); }
                            K = H; break;
                    }
                    const int rps = rows_per_stage(K, RG);
                    const char* srcb = reinterpret_cast<const char*>(W + (size_t)sl.row0 * K);
                    for (int r = 0; r < sl.nrows && !stopped; r += rps) {
                        const int n = min(rps, sl.nrows - r);
                        const unsigned slot = issued % FK_STAGES, par = ((issued / FK_STAGES) & 1u) ^ 1u;
                        unsigned long long t0 = 0;
                        while (!mbar_try_wait(&sh->empty[slot], par)) {
                            if (sh->stop) { stopped = true; break; }
                            if (t0 == 0) t0 = clock64();
                            else if (clock64() - t0 > FK_SPIN_LIMIT) { stopped = true; break; }
                        }
                        if (stopped) break;
                        const uint32_t bytes = (uint32_t)n * K * 2;
                        mbar_expect_tx(&sh->full[slot], bytes);
                        bulk_g2s(fk_smem + (size_t)slot * FK_STAGE_BYTES, srcb + (size_t)r * K * 2, bytes, &sh->full[slot]);
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
    c.xs = reinterpret_cast<float*>(fk_smem + so.scratch);
    c.att = reinterpret_cast<float*>(fk_smem + so.scratch + so.xs_bytes);
    c.nxt = reinterpret_cast<float*>(fk_smem + so.nxt);
    c.tid = tid; c.lane = lane; c.warp = warp; c.cta = cta; c.ncta = ncta;
    c.gen = 0; c.stage_ctr = 0; c.aborted = false;
    c.dbg = (p.dbg && cta == p.dbg_cta) ? p.dbg + 1 : nullptr; c.dbg_n = 0; c.dbg_cap = p.dbg_cap - 1; c.dbg_tag = 0;
    FkSampScratch ss;
    ss.x = reinterpret_cast<float*>(fk_smem + so.scratch);
    ss.pr = ss.x + so.maxV; ss.spr = ss.pr + so.maxV;
    ss.idx = reinterpret_cast<unsigned short*>(ss.spr + so.maxV); ss.rank = ss.idx + so.maxV;

    const int H = p.talker.H;
    const int Kpad4 = ((H + 255) & ~255) >> 2;
    float4* xs4 = reinterpret_cast<float4*>(c.xs);
    int pos = st0.pos, frame = st0.frame, done = st0.done, n_frames = st0.n_frames;
    const SamplingDev sp = *p.sp;
    int prefill_i = 0;
    int cb = 0;                      // next codebook to draw in the current frame (0 = talker code)
    bool mode1_done = false;

    // One loop, one pass per iteration: [draw + glue ->] token pass. (src/tts_onnx.cpp:794, 801-846)
    while (!c.aborted) {
        FkPass ps;
        if (p.mode == 1) {
            if (mode1_done) break;
            ps = FkPass{false, 1, pos, p.next_in, H, false, p.t_head, p.vocab, p.logits, p.last_hidden};
        } else if (prefill_i < n_prefill) {
            const bool last = (prefill_i == n_prefill - 1);
            ps = FkPass{false, 1, pos, p.prompt + (size_t)prefill_i * H, H, false, last ? p.t_head : nullptr, p.vocab,
                        p.logits, p.last_hidden};
        } else {
            if (done || frame >= frame_end) break;
            // ---- draw codebook cb of this frame (:803-812 for cb 0, :863-864 otherwise) -----------------
            const int tk = cb ? 1 : 0;
            fk_phase(c, tk, FKT_SAMPLE);
            fk_mark(c, 1);
            gbar_wait(c); if (c.aborted) break;
            fk_mark(c, 2);
            float* tr = (p.trace && cta == 0) ? p.trace + ((size_t)frame * 16 + cb) * p.trace_stride : nullptr;
            int tok = (cb == 0) ? fk_sample(c, ss, p.logits, p.vocab, 2048, p.vocab, 2150, sp, (uint32_t)frame, 0, tr)
                                : fk_sample(c, ss, p.clogits, p.cp_vocab, 0, 0, -1, sp, (uint32_t)frame, cb, tr);
            if (p.forced && frame < st0.n_forced) tok = (int)p.forced[(size_t)frame * 16 + cb];
            fk_mark(c, 3);
            if (cb == 0 && tok == 2150) { done = 1; break; }                        // CODEC_EOS (:812)
            if (cta == 0 && tid == 0) p.codes_out[(size_t)frame * 16 + cb] = tok;   // :818-821
            // ---- glue: embedding of the drawn code, running 16-way sum, next input rows ------------------
            const bool last_cb = (cb == p.cp_steps);
            const bool use_tr = frame < st0.trailing_len;
            const bf16_t* row = (cb == 0) ? p.codec_embed + (size_t)tok * H
                                          : p.cp_embed + ((size_t)(cb - 1) * p.cp_vocab + tok) * H;
            for (int k4 = tid; k4 < Kpad4; k4 += FK_CTHREADS) {
                float4 e = make_float4(0.f, 0.f, 0.f, 0.f), acc = e, lh = e;
                if (k4 < (H >> 2)) {
                    const uint2 u = __ldg(reinterpret_cast<const uint2*>(row) + k4);
                    e = make_float4(bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y));
                    if (cb == 0) {
                        acc = e;                                                    // :824
                        lh = __ldcg(reinterpret_cast<const float4*>(p.last_hidden) + k4);   // :859
                    } else {
                        acc = reinterpret_cast<float4*>(c.nxt)[k4];                 // :825-830
                        acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
                    }
                    if (last_cb) {                                                  // :833-842
                        const float4 tt = use_tr ? __ldg(reinterpret_cast<const float4*>(p.trailing + (size_t)frame * H) + k4)
                                                 : __ldg(reinterpret_cast<const float4*>(p.tts_pad) + k4);
                        acc.x += tt.x; acc.y += tt.y; acc.z += tt.z; acc.w += tt.w;
                    }
                    reinterpret_cast<float4*>(c.nxt)[k4] = acc;
                    if (cta == 0) {                                                 // layer-0 residual of the next pass
                        if (last_cb) reinterpret_cast<float4*>(p.next_in)[k4] = acc;
                        else if (cb == 0) { reinterpret_cast<float4*>(p.cp_in)[k4] = lh; reinterpret_cast<float4*>(p.cp_in + H)[k4] = e; }
                        else reinterpret_cast<float4*>(p.cp_in)[k4] = e;
                    }
                }
                if (cb == 0) { xs4[xs_perm4(k4)] = lh; xs4[Kpad4 + xs_perm4(k4)] = e; }
                else xs4[xs_perm4(k4)] = last_cb ? acc : e;
            }
            csync();
            fk_mark(c, 4);
            if (last_cb) {
                n_frames = frame + 1;
                ps = FkPass{false, 1, pos, p.next_in, H, true, p.t_head, p.vocab, p.logits, p.last_hidden};   // :845
            } else {
                ps = FkPass{true, cb == 0 ? 2 : 1, cb == 0 ? 0 : cb + 1, p.cp_in, H, true,
                            p.c_heads + (size_t)cb * p.cp_vocab * p.cp.H, p.cp_vocab, p.clogits, nullptr};
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
    // ---- exit: publish state, stop the producer -----------------------------------------------------
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
