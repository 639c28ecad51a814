// Kernels of the BATCHED frame loop (BASELINE configs[3]/[4]: many concurrent utterances per GPU): everything around the
// tcgen05 GEMMs of tc_gemm.cuh. B utterance slots run loops A and B of the reference (src/tts_onnx.cpp:782-872) in lockstep;
// every slot has its own position, Philox key, KV pages and state (BatchState), so a slot can be (re)filled at any frame
// boundary (continuous batching) and a finished slot simply idles.
//
//   bprep_kernel    residual + split-K partials (+bias) -> residual stream, RMSNorm, bf16 planes of the GEMM's X operand
//   battn_kernel    q/k RMSNorm + RoPE + paged KV append + GQA split-KV attention, output as bf16 planes (O-projection X)
//   bswiglu_kernel  silu(gate) * up from the gate|up partials -> bf16 planes (down-projection X)
//   bsample_kernel  the reference's sampler + frame glue per slot (sampler.cuh sample_block)
//   badvance_kernel per-slot position / prefill bookkeeping
// X operand layout (tc_gemm.cuh): row = (tile * planes + plane) * Bt + (b % Bt), tile = b / Bt; K contiguous.
#pragma once
#include "attention.cuh"
#include "common.cuh"
#include "sampler.cuh"

namespace lqt {

struct BatchState {
    GenState g;               // first member: sampler.cuh works on it unchanged
    int P;                    // prompt rows of the utterance in this slot
    int prefill_pos;          // prompt rows already fed; < P: the slot is prefilling (no draws, talker input = prompt row)
    int active;               // 0: empty slot
    int pad_;
};

LQT_DEVINL bool bslot_idle(const BatchState& s) { return !s.active || s.g.done; }
LQT_DEVINL bool bslot_prefilling(const BatchState& s) { return s.prefill_pos < s.P; }

// fp32 -> bf16 planes at row (tile * planes + p) * Bt + bl of the X operand, 4 consecutive columns
LQT_DEVINL void bstore_planes4(__nv_bfloat16* X, int K, int planes, int Bt, int b, int k, float4 v) {
    const int tile = b / Bt, bl = b - tile * Bt;
    float r[4] = {v.x, v.y, v.z, v.w};
    for (int p = 0; p < planes; ++p) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(r[0], r[1]), c = __floats2bfloat162_rn(r[2], r[3]);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&a); u.y = *reinterpret_cast<const uint32_t*>(&c);
        *reinterpret_cast<uint2*>(X + ((size_t)(tile * planes + p) * Bt + bl) * K + k) = u;
        r[0] -= bf16lo(u.x); r[1] -= bf16hi(u.x); r[2] -= bf16lo(u.y); r[3] -= bf16hi(u.y);      // exact remainders
    }
}

// sum of the split-K partials of 4 consecutive outputs, in the fixed order 0..S-1 (bit-reproducible). The loads of up to 8
// splits are issued before the first add: the partials sit in L2 and a dependent chain of S round trips is what this costs
// otherwise (profiles/r2_batched_launches_*.md: 7.5 us per bprep launch with a rolled loop).
LQT_DEVINL float4 bsum_splits4(const float* base, int n_splits, long long split_stride, float4 v) {
    for (int q0 = 0; q0 < n_splits; q0 += 8) {
        float4 w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            w[u] = (q0 + u < n_splits) ? __ldcg(reinterpret_cast<const float4*>(base + (size_t)(q0 + u) * split_stride)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (q0 + u < n_splits) { v.x += w[u].x; v.y += w[u].y; v.z += w[u].z; v.w += w[u].w; }
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
struct BPrepParams {
    const BatchState* st;
    // input row of slot b: `resid` [B][H]; or, when `prompt` is set (talker layer 0), the slot's next prompt row while it
    // is prefilling and resid[b] (= next_in) otherwise
    const float* resid;
    const float* prompt; int prompt_rows;        // [B][prompt_rows][H]
    const float* part; int n_splits; long long split_stride;   // nullable split-K partials [S][Bpad][H]
    const float* bias;                           // nullable [H]
    const float* norm_w; float eps;              // nullable: no normalisation (in_proj input)
    float* x_out;                                // nullable [B][H]: the updated residual stream
    float* hid_out;                              // nullable [B][H]: the normalised row in fp32 (talker last_hidden)
    __nv_bfloat16* X; int planes, Bt;            // nullable
    int H;
};

__global__ void __launch_bounds__(256)
bprep_kernel(const BPrepParams p) {
    extern __shared__ float bp_row[];            // [H]
    __shared__ float red[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    pdl_trigger();
    pdl_wait();
    const BatchState s = p.st[b];
    if (bslot_idle(s)) return;
    const int H = p.H;
    const float* in = p.resid ? p.resid + (size_t)b * H : nullptr;
    if (p.prompt && bslot_prefilling(s)) in = p.prompt + ((size_t)b * p.prompt_rows + s.prefill_pos) * H;
    float ss = 0.f;
    for (int k = tid * 4; k < H; k += 1024) {
        float4 v = in ? *reinterpret_cast<const float4*>(in + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.n_splits) v = bsum_splits4(p.part + (size_t)b * H + k, p.n_splits, p.split_stride, v);
        if (p.bias) { const float4 w = *reinterpret_cast<const float4*>(p.bias + k); v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
        *reinterpret_cast<float4*>(bp_row + k) = v;
        if (p.x_out) *reinterpret_cast<float4*>(p.x_out + (size_t)b * H + k) = v;
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    float rstd = 1.f;
    if (p.norm_w) {
        ss = warp_sum(ss);
        if ((tid & 31) == 0) red[tid >> 5] = ss;
        __syncthreads();
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        rstd = 1.0f / sqrtf(t / (float)H + p.eps);
    }
    for (int k = tid * 4; k < H; k += 1024) {
        float4 v = *reinterpret_cast<const float4*>(bp_row + k);
        if (p.norm_w) {
            const float4 w = *reinterpret_cast<const float4*>(p.norm_w + k);
            v.x = (v.x * rstd) * w.x; v.y = (v.y * rstd) * w.y; v.z = (v.z * rstd) * w.z; v.w = (v.w * rstd) * w.w;
        }
        if (p.hid_out) *reinterpret_cast<float4*>(p.hid_out + (size_t)b * H + k) = v;
        if (p.X) bstore_planes4(p.X, H, p.planes, p.Bt, b, k, v);
    }
}

// ------------------------------------------------------------------------------------------------
struct BSwigluParams {
    const BatchState* st;
    const float* part; int n_splits; long long split_stride;   // [S][Bpad][2I]: gate rows [0,I), up rows [I,2I)
    __nv_bfloat16* X; int planes, Bt;
    int I;
};
__global__ void __launch_bounds__(256)
bswiglu_kernel(const BSwigluParams p) {
    const int b = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    if (bslot_idle(p.st[b])) return;
    const int I = p.I;
    for (int k = (blockIdx.y * 256 + threadIdx.x) * 4; k < I; k += gridDim.y * 1024) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 g = bsum_splits4(p.part + (size_t)b * 2 * I + k, p.n_splits, p.split_stride, z);
        const float4 u = bsum_splits4(p.part + (size_t)b * 2 * I + I + k, p.n_splits, p.split_stride, z);
        bstore_planes4(p.X, I, p.planes, p.Bt, b, k, make_float4(silu_f(g.x) * u.x, silu_f(g.y) * u.y, silu_f(g.z) * u.z, silu_f(g.w) * u.w));
    }
}

// ------------------------------------------------------------------------------------------------
// Attention of ONE new position per slot (decode shape). grid = (n_kv, nsplit, B). Same arithmetic as attn_decode_kernel
// (attention.cuh); differences: q/k/v come from the QKV GEMM's split-K partials, position / page table / KV pages are per
// slot, and the output goes straight into the O-projection's X operand as bf16 planes.
struct BAttnParams {
    const BatchState* st;
    const float* qkv_part; int n_splits; long long split_stride;   // [S][Bpad][qkv_dim]
    const float* qnorm; const float* knorm;
    const float* rope_cos; const float* rope_sin;
    int fixed_pos;             // >= 0: position of the new token (code predictor); < 0: st[b].g.pos (talker)
    void* kv_pool;
    const int* page_table; int pt_stride;       // logical page -> physical page, [B][pt_stride]
    float* partial;            // [B][n_kv][nsplit][2][ATT_PSTRIDE]
    int* counters;             // [B][n_kv]
    __nv_bfloat16* X; int planes, Bt;           // O-projection operand, K = q_dim
    long long page_stride, layer_off;
    int page_shift, n_kv;
    float eps, scale;
};

template <typename KVT>
__global__ void __launch_bounds__(ATT_THREADS)
battn_kernel(const BAttnParams p) {
    constexpr int REP = 2;
    extern __shared__ float sc[];
    __shared__ __align__(16) float q_s[REP][ATT_D];
    __shared__ float red_m[REP][ATT_WARPS], red_l[REP][ATT_WARPS];
    __shared__ __align__(16) float o_s[ATT_WARPS][REP][ATT_D];
    __shared__ int ticket_s;
    const int b = blockIdx.z;
    pdl_trigger();
    pdl_wait();
    const BatchState s = p.st[b];
    if (bslot_idle(s)) return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x, split = blockIdx.y, nsplit = gridDim.y;
    const int t = p.fixed_pos >= 0 ? p.fixed_pos : s.g.pos;
    const int PS = 1 << p.page_shift;
    const int n_pos = t + 1;
    const int n_pages = (n_pos + PS - 1) >> p.page_shift;
    const int active = min(nsplit, n_pages);
    if (split >= active) return;
    const int n_heads = p.n_kv * REP;
    const int q_dim = n_heads * ATT_D, kv_dim = p.n_kv * ATT_D, qkv_dim = q_dim + 2 * kv_dim;
    const float* cosr = p.rope_cos + (size_t)t * (ATT_D / 2);
    const float* sinr = p.rope_sin + (size_t)t * (ATT_D / 2);
    KVT* pool = reinterpret_cast<KVT*>(p.kv_pool);
    const int* pt = p.page_table + (size_t)b * p.pt_stride;
    const long long head_off = (long long)g * PS * ATT_D;
    const long long v_off = (long long)p.n_kv * PS * ATT_D;
    auto load_qkv = [&](int off) {                                   // sum of the split-K partials, fixed order
        return bsum_splits4(p.qkv_part + (size_t)b * qkv_dim + off + lane * 4, p.n_splits, p.split_stride, make_float4(0.f, 0.f, 0.f, 0.f));
    };
    if (warp < REP) {
        float4 v = load_qkv((g * REP + warp) * ATT_D);
        v = head_norm_rope(v, p.qnorm, p.eps, cosr, sinr, lane);
        reinterpret_cast<float4*>(q_s[warp])[lane] = v;
    } else if (warp == REP || warp == REP + 1) {
        const int tpage = t >> p.page_shift;
        if (tpage % nsplit == split) {
            const long long base = (long long)pt[tpage] * p.page_stride + p.layer_off + head_off + (long long)(t & (PS - 1)) * ATT_D;
            if (warp == REP) {
                float4 v = load_qkv(q_dim + g * ATT_D);
                v = head_norm_rope(v, p.knorm, p.eps, cosr, sinr, lane);
                KvIO<KVT>::store4(pool + base + lane * 4, v);
            } else {
                const float4 v = load_qkv(q_dim + kv_dim + g * ATT_D);
                KvIO<KVT>::store4(pool + base + v_off + lane * 4, v);
            }
        }
    }
    __syncthreads();

    float4 q[REP];
#pragma unroll
    for (int r = 0; r < REP; ++r) q[r] = reinterpret_cast<const float4*>(q_s[r])[lane];
    const int cap = ((n_pages + nsplit - 1) / nsplit) * PS;
    float mloc[REP];
#pragma unroll
    for (int r = 0; r < REP; ++r) mloc[r] = -INFINITY;
    int li = 0;
    for (int pg = split; pg < n_pages; pg += nsplit, ++li) {
        const KVT* kb = pool + (long long)pt[pg] * p.page_stride + p.layer_off + head_off;
        const int p0 = pg << p.page_shift;
        const int cnt = min(PS, n_pos - p0);
        for (int i0 = warp * 4; i0 < cnt; i0 += ATT_WARPS * 4) {
            float4 kv4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u < cnt) kv4[u] = KvIO<KVT>::load4(kb + (long long)(i0 + u) * ATT_D + lane * 4);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u < cnt) {
#pragma unroll
                    for (int r = 0; r < REP; ++r) {
                        float d = kv4[u].x * q[r].x + kv4[u].y * q[r].y + kv4[u].z * q[r].z + kv4[u].w * q[r].w;
                        d = warp_sum(d) * p.scale;
                        mloc[r] = fmaxf(mloc[r], d);
                        if (lane == 0) sc[r * cap + li * PS + i0 + u] = d;
                    }
                }
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < REP; ++r) red_m[r][warp] = mloc[r];
    }
    __syncthreads();
    float mcta[REP];
#pragma unroll
    for (int r = 0; r < REP; ++r) {
        float m = red_m[r][0];
#pragma unroll
        for (int w = 1; w < ATT_WARPS; ++w) m = fmaxf(m, red_m[r][w]);
        mcta[r] = m;
    }
    float4 acc[REP];
    float lsum[REP];
#pragma unroll
    for (int r = 0; r < REP; ++r) { acc[r] = make_float4(0.f, 0.f, 0.f, 0.f); lsum[r] = 0.f; }
    li = 0;
    for (int pg = split; pg < n_pages; pg += nsplit, ++li) {
        const KVT* vb = pool + (long long)pt[pg] * p.page_stride + p.layer_off + head_off + v_off;
        const int p0 = pg << p.page_shift;
        const int cnt = min(PS, n_pos - p0);
        for (int i0 = warp * 4; i0 < cnt; i0 += ATT_WARPS * 4) {
            float4 vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u < cnt) vv[u] = KvIO<KVT>::load4(vb + (long long)(i0 + u) * ATT_D + lane * 4);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u < cnt) {
#pragma unroll
                    for (int r = 0; r < REP; ++r) {
                        const float pr = expf(sc[r * cap + li * PS + i0 + u] - mcta[r]);
                        lsum[r] += pr;
                        acc[r].x = fmaf(pr, vv[u].x, acc[r].x); acc[r].y = fmaf(pr, vv[u].y, acc[r].y);
                        acc[r].z = fmaf(pr, vv[u].z, acc[r].z); acc[r].w = fmaf(pr, vv[u].w, acc[r].w);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < REP; ++r) {
        reinterpret_cast<float4*>(o_s[warp][r])[lane] = acc[r];
        if (lane == 0) red_l[r][warp] = lsum[r];
    }
    __syncthreads();
    if (active == 1) {
        // one split (always the case for the code predictor's <= 17 positions and for short talker contexts): no partial record,
        // no ticket, no fence -- the CTA holds the whole softmax
        if (tid < REP * ATT_D / 4) {
            const int r = tid / (ATT_D / 4), d = (tid % (ATT_D / 4)) * 4;
            float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
            float den = 0.f;
#pragma unroll
            for (int w = 0; w < ATT_WARPS; ++w) {
                const float4 o = *reinterpret_cast<const float4*>(&o_s[w][r][d]);
                num.x += o.x; num.y += o.y; num.z += o.z; num.w += o.w;
                den += red_l[r][w];
            }
            bstore_planes4(p.X, q_dim, p.planes, p.Bt, b, (g * REP + r) * ATT_D + d, make_float4(num.x / den, num.y / den, num.z / den, num.w / den));
        }
        return;
    }
    float* partial = p.partial + (size_t)b * p.n_kv * nsplit * REP * ATT_PSTRIDE;
    float* part = partial + ((size_t)(g * nsplit + split) * REP) * ATT_PSTRIDE;
    for (int e = tid; e < REP * ATT_D; e += ATT_THREADS) {
        const int r = e / ATT_D, d = e % ATT_D;
        float sm = 0.f;
#pragma unroll
        for (int w = 0; w < ATT_WARPS; ++w) sm += o_s[w][r][d];
        part[r * ATT_PSTRIDE + d] = sm;
    }
    if (tid < REP) {
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < ATT_WARPS; ++w) l += red_l[tid][w];
        part[tid * ATT_PSTRIDE + ATT_D] = mcta[tid];
        part[tid * ATT_PSTRIDE + ATT_D + 1] = l;
    }
    __threadfence();
    __syncthreads();
    int* counter = p.counters + (size_t)b * p.n_kv + g;
    if (tid == 0) ticket_s = atomicAdd(counter, 1);
    __syncthreads();
    if (ticket_s != active - 1) return;
    __threadfence();
    // last CTA of (slot, kv head): combine the splits; thread -> 4 consecutive dims of one head (64 threads busy)
    if (tid < REP * ATT_D / 4) {
        const int r = tid / (ATT_D / 4), d = (tid % (ATT_D / 4)) * 4;
        float M = -INFINITY;
        for (int sp = 0; sp < active; ++sp)
            M = fmaxf(M, __ldcg(partial + ((size_t)(g * nsplit + sp) * REP + r) * ATT_PSTRIDE + ATT_D));
        float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
        float den = 0.f;
        for (int sp = 0; sp < active; ++sp) {
            const float* ps = partial + ((size_t)(g * nsplit + sp) * REP + r) * ATT_PSTRIDE;
            const float w = expf(__ldcg(ps + ATT_D) - M);
            const float4 o = __ldcg(reinterpret_cast<const float4*>(ps + d));
            num.x = fmaf(w, o.x, num.x); num.y = fmaf(w, o.y, num.y); num.z = fmaf(w, o.z, num.z); num.w = fmaf(w, o.w, num.w);
            den = fmaf(w, __ldcg(ps + ATT_D + 1), den);
        }
        bstore_planes4(p.X, q_dim, p.planes, p.Bt, b, (g * REP + r) * ATT_D + d, make_float4(num.x / den, num.y / den, num.z / den, num.w / den));
    }
    if (tid == 0) *counter = 0;
}

// ------------------------------------------------------------------------------------------------
// Code-predictor attention: <= 17 positions, fp32 KV of 32 positions per slot. One WARP per (slot, kv group): q/k RMSNorm +
// RoPE, KV append, both query heads' softmax and P.V in registers -- no shared memory, no block barrier (battn_kernel spent
// 25 us per launch at 256 slots on block-level machinery for a 17-position softmax; 80 of these run per frame).
struct BCpAttnParams {
    const BatchState* st;
    const float* qkv_part; int n_splits; long long split_stride;
    const float* qnorm; const float* knorm; const float* rope_cos; const float* rope_sin;
    int pos;                   // position of the new token (0..16)
    float* kv;                 // [B][layers][k|v][n_kv][PS][128] fp32
    long long slot_stride, layer_off;
    int PS, n_kv, B;
    __nv_bfloat16* X; int planes, Bt;
    float eps, scale;
};
// (256, 2): <= 128 registers so that two CTAs (16 warps) share an SM -- 256 slots x 8 groups = 2048 warps then fit in one wave;
// the K rows and the V rows take turns in the same registers.
__global__ void __launch_bounds__(256, 2)
bcp_attn_kernel(const BCpAttnParams p) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    pdl_trigger();
    pdl_wait();
    if (w >= p.B * p.n_kv) return;
    const int b = w / p.n_kv, g = w - b * p.n_kv;
    if (bslot_idle(p.st[b])) return;
    const int q_dim = p.n_kv * 2 * ATT_D, kv_dim = p.n_kv * ATT_D, qkv_dim = q_dim + 2 * kv_dim;
    const float* cosr = p.rope_cos + (size_t)p.pos * (ATT_D / 2);
    const float* sinr = p.rope_sin + (size_t)p.pos * (ATT_D / 2);
    const float* base = p.qkv_part + (size_t)b * qkv_dim + lane * 4;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float* kc = p.kv + (size_t)b * p.slot_stride + p.layer_off + (size_t)g * p.PS * ATT_D;
    float* vc = kc + (size_t)p.n_kv * p.PS * ATT_D;
    constexpr int MAXP = 17;
    // the cached K rows are requested first (predicated on j < pos, independent of everything else): they travel while the
    // new token's q/k/v are summed, normalised and rotated
    float4 rows[MAXP];
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
        rows[j] = z;
        if (j < p.pos) rows[j] = __ldcg(reinterpret_cast<const float4*>(kc + (size_t)j * ATT_D) + lane);
    }
    float4 q0 = bsum_splits4(base + (g * 2) * ATT_D, p.n_splits, p.split_stride, z);
    float4 q1 = bsum_splits4(base + (g * 2 + 1) * ATT_D, p.n_splits, p.split_stride, z);
    float4 kn = bsum_splits4(base + q_dim + g * ATT_D, p.n_splits, p.split_stride, z);
    const float4 vn = bsum_splits4(base + q_dim + kv_dim + g * ATT_D, p.n_splits, p.split_stride, z);
    q0 = head_norm_rope(q0, p.qnorm, p.eps, cosr, sinr, lane);
    q1 = head_norm_rope(q1, p.qnorm, p.eps, cosr, sinr, lane);
    kn = head_norm_rope(kn, p.knorm, p.eps, cosr, sinr, lane);
    reinterpret_cast<float4*>(kc + (size_t)p.pos * ATT_D)[lane] = kn;
    reinterpret_cast<float4*>(vc + (size_t)p.pos * ATT_D)[lane] = vn;
    // scores of both heads with ONE butterfly per position: the halves of the warp swap one partial sum, then reduce within the
    // half (lanes 0-15 end up with head 0, lanes 16-31 with head 1)
    const bool hi = (lane & 16) != 0;
    float sc[MAXP];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
        sc[j] = -INFINITY;
        if (j <= p.pos) {
            const float4 k4 = (j == p.pos) ? kn : rows[j];
            const float d0 = k4.x * q0.x + k4.y * q0.y + k4.z * q0.z + k4.w * q0.w;
            const float d1 = k4.x * q1.x + k4.y * q1.y + k4.z * q1.z + k4.w * q1.w;
            float t = (hi ? d1 : d0) + __shfl_xor_sync(0xffffffffu, hi ? d0 : d1, 16);
            t += __shfl_xor_sync(0xffffffffu, t, 8);
            t += __shfl_xor_sync(0xffffffffu, t, 4);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            sc[j] = t * p.scale;
            mx = fmaxf(mx, sc[j]);
        }
    }
    // V rows into the same registers
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
        rows[j] = z;
        if (j < p.pos) rows[j] = __ldcg(reinterpret_cast<const float4*>(vc + (size_t)j * ATT_D) + lane);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
        if (j <= p.pos) { sc[j] = expf(sc[j] - mx); den += sc[j]; }          // this half's head
    }
    float4 a0 = z, a1 = z;
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
        if (j <= p.pos) {
            const float4 v4 = (j == p.pos) ? vn : rows[j];
            const float mine = sc[j] / den;                                   // probability of position j for this half's head
            const float other = __shfl_xor_sync(0xffffffffu, mine, 16);
            const float e0 = hi ? other : mine, e1 = hi ? mine : other;
            a0.x = fmaf(e0, v4.x, a0.x); a0.y = fmaf(e0, v4.y, a0.y); a0.z = fmaf(e0, v4.z, a0.z); a0.w = fmaf(e0, v4.w, a0.w);
            a1.x = fmaf(e1, v4.x, a1.x); a1.y = fmaf(e1, v4.y, a1.y); a1.z = fmaf(e1, v4.z, a1.z); a1.w = fmaf(e1, v4.w, a1.w);
        }
    }
    bstore_planes4(p.X, q_dim, p.planes, p.Bt, b, (g * 2) * ATT_D + lane * 4, a0);
    bstore_planes4(p.X, q_dim, p.planes, p.Bt, b, (g * 2 + 1) * ATT_D + lane * 4, a1);
}

// ------------------------------------------------------------------------------------------------
// per-slot sampler + glue: the batch-1 sampler block (sampler.cuh) with per-slot pointers
struct BSampleParams {
    BatchState* st; const SamplingDev* sp;        // [B] each (per-slot Philox utterance id)
    const float* logits_part; int n_splits; long long split_stride; int V;     // [S][Bpad][V]
    int mask_lo, mask_hi, mask_keep, codebook;
    const __nv_bfloat16* embed_table; int H;
    float* cp_in; float* next_in;                 // [B][H]
    const float* trailing; long long trailing_stride;   // [B][trailing_stride]
    const float* tts_pad;                         // [B][H]
    long long* codes_out; long long codes_stride; // [B][codes_stride]
    const long long* forced;                      // nullable, same layout as codes_out
    float* trace; long long trace_bstride; int trace_stride;   // nullable [B][frames][16][trace_stride]
    int eos_id, n_codebooks;
};
constexpr int BSMP_THREADS = 256;
__global__ void __launch_bounds__(BSMP_THREADS)
bsample_kernel(const BSampleParams q) {
    const int b = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    BatchState* st = q.st + b;
    if (!st->active || bslot_prefilling(*st)) return;              // (done slots leave inside sample_block)
    SampleParams p{};
    p.logits = q.logits_part + (size_t)b * q.V; p.V = q.V; p.n_splits = q.n_splits; p.split_stride = q.split_stride;
    p.mask_lo = q.mask_lo; p.mask_hi = q.mask_hi; p.mask_keep = q.mask_keep;
    p.sp = q.sp + b; p.codebook = q.codebook; p.st = &st->g; p.token_out = nullptr;
    p.embed_table = q.embed_table; p.H = q.H;
    p.cp_in = q.cp_in + (size_t)b * q.H; p.next_in = q.next_in + (size_t)b * q.H;
    p.trailing = q.trailing + (size_t)b * q.trailing_stride; p.tts_pad = q.tts_pad + (size_t)b * q.H;
    p.codes_out = q.codes_out + (size_t)b * q.codes_stride;
    p.forced = q.forced ? q.forced + (size_t)b * q.codes_stride : nullptr;
    p.trace = q.trace ? q.trace + (size_t)b * q.trace_bstride : nullptr; p.trace_stride = q.trace_stride;
    p.eos_id = q.eos_id; p.n_codebooks = q.n_codebooks;
    sample_block<BSMP_THREADS>(p);
}

__global__ void badvance_kernel(BatchState* st, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (b >= B) return;
    BatchState& s = st[b];
    if (bslot_idle(s)) return;
    s.g.pos += 1;
    if (bslot_prefilling(s)) s.prefill_pos += 1;
}

}  // namespace lqt
