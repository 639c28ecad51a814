// Decode-time attention for one new token: fused per-head q/k RMSNorm + rotate-half RoPE +
// KV append into the paged cache + GQA attention (split over KV pages, flash-decoding style)
// + split combine by the last-arriving CTA.
//
// Replaces the attention inside talker_decode.onnx / code_predictor.onnx and the host-side
// KVCache round trip of the reference (src/tts_onnx.cpp:684-691, 721-729): K/V never leave HBM.
//
// KV pool layout: [phys_page][layer][k|v][kv_head][page_size][D], D = 128.
// Talker: bf16 pages of 64 positions. Code predictor: one fp32 "page" of 32 positions.
#pragma once
#include "common.cuh"

namespace lqt {

constexpr int ATT_D = 128;
constexpr int ATT_THREADS = 256;
constexpr int ATT_WARPS = ATT_THREADS / 32;
constexpr int ATT_PSTRIDE = ATT_D + 4;        // partial record: o[128], m, l, pad, pad

struct AttnParams {
    const float* qkv;          // [q_dim + 2*kv_dim] raw projections of the new token
    const float* qnorm;        // nullable [D]
    const float* knorm;        // nullable [D]
    const float* rope_cos;     // [max_pos][D/2]
    const float* rope_sin;
    const int* pos_ptr;        // device: position t of the new token; it attends to [0..t]
    void* kv_pool;             // pool base (element type = template KVT)
    const int* page_table;     // logical page -> physical page (this slot)
    float* partial;            // [n_kv][nsplit][REP][ATT_PSTRIDE]
    int* counters;             // [n_kv] tickets (zero on entry; reset by the last CTA)
    float* out;                // [n_heads * D]
    const int* done;           // nullable early-exit flag
    long long page_stride;     // elements between physical pages  (= layers*2*n_kv*PS*D)
    long long layer_off;       // element offset of this layer's K block inside a page
    int page_shift;            // log2(page_size)
    int n_kv;
    float eps, scale;
};

// one warp: RMSNorm (optional) + RoPE on a 128-wide head vector; lane holds dims [4*lane, 4*lane+4)
// per-head RMSNorm (weights w; skipped when !normed) + rotate-half RoPE with this lane's cos / sin values already in registers
// (callers on a latency-critical path load them BEFORE they wait for v)
LQT_DEVINL float4 head_norm_rope_regs(float4 v, bool normed, const float4 w, float eps, const float4 c, const float4 s, int lane) {
    if (normed) {
        float ss = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
        const float r = 1.0f / sqrtf(ss / (float)ATT_D + eps);
        v.x = (v.x * r) * w.x; v.y = (v.y * r) * w.y; v.z = (v.z * r) * w.z; v.w = (v.w * r) * w.w;
    }
    float4 o;                                            // rotate-half: partner dims are +-64 -> lane ^ 16
    o.x = __shfl_xor_sync(0xffffffffu, v.x, 16); o.y = __shfl_xor_sync(0xffffffffu, v.y, 16);
    o.z = __shfl_xor_sync(0xffffffffu, v.z, 16); o.w = __shfl_xor_sync(0xffffffffu, v.w, 16);
    float4 r;
    if (lane < 16) {       // first half:  x1*c - x2*s
        r.x = v.x * c.x - o.x * s.x; r.y = v.y * c.y - o.y * s.y;
        r.z = v.z * c.z - o.z * s.z; r.w = v.w * c.w - o.w * s.w;
    } else {               // second half: x2*c + x1*s
        r.x = v.x * c.x + o.x * s.x; r.y = v.y * c.y + o.y * s.y;
        r.z = v.z * c.z + o.z * s.z; r.w = v.w * c.w + o.w * s.w;
    }
    return r;
}
LQT_DEVINL float4 head_norm_rope(float4 v, const float* norm_w, float eps,
                                 const float* cosr, const float* sinr, int lane) {
    if (norm_w) {
        float ss = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
        const float r = 1.0f / sqrtf(ss / (float)ATT_D + eps);
        const float4 w = reinterpret_cast<const float4*>(norm_w)[lane];
        v.x = (v.x * r) * w.x; v.y = (v.y * r) * w.y; v.z = (v.z * r) * w.z; v.w = (v.w * r) * w.w;
    }
    // rotate-half: partner dims are +-64 -> lane ^ 16
    float4 o;
    o.x = __shfl_xor_sync(0xffffffffu, v.x, 16); o.y = __shfl_xor_sync(0xffffffffu, v.y, 16);
    o.z = __shfl_xor_sync(0xffffffffu, v.z, 16); o.w = __shfl_xor_sync(0xffffffffu, v.w, 16);
    const int f = (lane & 15) * 4;                       // frequency index of dim (mod 64)
    const float4 c = *reinterpret_cast<const float4*>(cosr + f);
    const float4 s = *reinterpret_cast<const float4*>(sinr + f);
    float4 r;
    if (lane < 16) {       // first half:  x1*c - x2*s
        r.x = v.x * c.x - o.x * s.x; r.y = v.y * c.y - o.y * s.y;
        r.z = v.z * c.z - o.z * s.z; r.w = v.w * c.w - o.w * s.w;
    } else {               // second half: x2*c + x1*s
        r.x = v.x * c.x + o.x * s.x; r.y = v.y * c.y + o.y * s.y;
        r.z = v.z * c.z + o.z * s.z; r.w = v.w * c.w + o.w * s.w;
    }
    return r;
}

// grid = (n_kv, nsplit); dynamic smem = REP * pages_per_cta_max * page_size floats (scores)
template <typename KVT, int REP>
__global__ void __launch_bounds__(ATT_THREADS)
attn_decode_kernel(const AttnParams p) {
    extern __shared__ float sc[];                          // [REP][cap] scores -> probabilities
    __shared__ __align__(16) float q_s[REP][ATT_D];
    __shared__ float red_m[REP][ATT_WARPS], red_l[REP][ATT_WARPS];
    __shared__ __align__(16) float o_s[ATT_WARPS][REP][ATT_D];
    __shared__ int ticket_s;
    if (p.done && *p.done) return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x, split = blockIdx.y, nsplit = gridDim.y;
    const int t = *p.pos_ptr;
    const int PS = 1 << p.page_shift;
    const int n_pos = t + 1;
    const int n_pages = (n_pos + PS - 1) >> p.page_shift;
    const int active = min(nsplit, n_pages);
    if (split >= active) return;
    const int n_heads = p.n_kv * REP;
    const int q_dim = n_heads * ATT_D, kv_dim = p.n_kv * ATT_D;
    const float* cosr = p.rope_cos + (size_t)t * (ATT_D / 2);
    const float* sinr = p.rope_sin + (size_t)t * (ATT_D / 2);
    KVT* pool = reinterpret_cast<KVT*>(p.kv_pool);
    const long long head_off = (long long)g * PS * ATT_D;
    const long long v_off = (long long)p.n_kv * PS * ATT_D;    // K block -> V block

    // ---- q heads of this group; new k/v into the cache (owner split only) ---------------------
    if (warp < REP) {
        const int h = g * REP + warp;
        float4 v = reinterpret_cast<const float4*>(p.qkv + (size_t)h * ATT_D)[lane];
        v = head_norm_rope(v, p.qnorm, p.eps, cosr, sinr, lane);
        reinterpret_cast<float4*>(q_s[warp])[lane] = v;
    } else if (warp == REP || warp == REP + 1) {
        const int tpage = t >> p.page_shift;
        if (tpage % nsplit == split) {
            const long long base = (long long)p.page_table[tpage] * p.page_stride + p.layer_off +
                                   head_off + (long long)(t & (PS - 1)) * ATT_D;
            if (warp == REP) {
                float4 v = reinterpret_cast<const float4*>(p.qkv + q_dim + (size_t)g * ATT_D)[lane];
                v = head_norm_rope(v, p.knorm, p.eps, cosr, sinr, lane);
                KvIO<KVT>::store4(pool + base + lane * 4, v);
            } else {
                float4 v = reinterpret_cast<const float4*>(p.qkv + q_dim + kv_dim + (size_t)g * ATT_D)[lane];
                KvIO<KVT>::store4(pool + base + v_off + lane * 4, v);
            }
        }
    }
    __syncthreads();

    // ---- pass 1: scores for the positions of my pages ----------------------------------------
    float4 q[REP];
#pragma unroll
    for (int r = 0; r < REP; ++r) q[r] = reinterpret_cast<const float4*>(q_s[r])[lane];
    const int cap = ((n_pages + nsplit - 1) / nsplit) * PS;     // score slots per head
    float mloc[REP];
#pragma unroll
    for (int r = 0; r < REP; ++r) mloc[r] = -INFINITY;
    int li = 0;                                                 // local page counter
    for (int pg = split; pg < n_pages; pg += nsplit, ++li) {
        const KVT* kb = pool + (long long)p.page_table[pg] * p.page_stride + p.layer_off + head_off;
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

    // ---- pass 2: p = exp(s - m), accumulate P.V per warp --------------------------------------
    float4 acc[REP];
    float lsum[REP];
#pragma unroll
    for (int r = 0; r < REP; ++r) { acc[r] = make_float4(0.f, 0.f, 0.f, 0.f); lsum[r] = 0.f; }
    li = 0;
    for (int pg = split; pg < n_pages; pg += nsplit, ++li) {
        const KVT* vb = pool + (long long)p.page_table[pg] * p.page_stride + p.layer_off + head_off + v_off;
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
    // CTA partial: thread -> (r, d)
    float* part = p.partial + ((size_t)(g * nsplit + split) * REP) * ATT_PSTRIDE;
    for (int e = tid; e < REP * ATT_D; e += ATT_THREADS) {
        const int r = e / ATT_D, d = e % ATT_D;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < ATT_WARPS; ++w) s += o_s[w][r][d];
        part[r * ATT_PSTRIDE + d] = s;
    }
    if (tid < REP) {
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < ATT_WARPS; ++w) l += red_l[tid][w];
        part[tid * ATT_PSTRIDE + ATT_D] = mcta[tid];
        part[tid * ATT_PSTRIDE + ATT_D + 1] = l;
    }

    // ---- last-arriving CTA of this kv head combines the splits --------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) ticket_s = atomicAdd(&p.counters[g], 1);
    __syncthreads();
    if (ticket_s != active - 1) return;
    __threadfence();
    for (int e = tid; e < REP * ATT_D; e += ATT_THREADS) {
        const int r = e / ATT_D, d = e % ATT_D;
        float M = -INFINITY;
        for (int s = 0; s < active; ++s)
            M = fmaxf(M, __ldcg(p.partial + ((size_t)(g * nsplit + s) * REP + r) * ATT_PSTRIDE + ATT_D));
        float num = 0.f, den = 0.f;
        for (int s = 0; s < active; ++s) {
            const float* ps = p.partial + ((size_t)(g * nsplit + s) * REP + r) * ATT_PSTRIDE;
            const float w = expf(__ldcg(ps + ATT_D) - M);
            num = fmaf(w, __ldcg(ps + d), num);
            den = fmaf(w, __ldcg(ps + ATT_D + 1), den);
        }
        p.out[(size_t)(g * REP + r) * ATT_D + d] = num / den;
    }
    if (tid == 0) p.counters[g] = 0;
}

// ------------------------------------------------------------------------------------------------
// Vocoder pre-transformer attention: T positions, MHA, causal sliding window, RoPE, head_dim 64.
// qkv rows [T][3*n_heads*64] (q | k | v), output [T][n_heads*64]. One warp per (position, head).
// ------------------------------------------------------------------------------------------------
struct WinAttnParams {
    const float* qkv; float* out;
    const float* rope_cos; const float* rope_sin;   // [max_pos][32]
    int T, n_heads, window;
    float scale;
    int pos0;                  // absolute position of row 0 (streaming decode: rows are a chunk of a longer sequence)
    int hist;                  // rows [-hist, 0) in front of qkv hold the q|k|v of the previous positions (<= window - 1)
};

__global__ void __launch_bounds__(256)
window_attn_kernel(const WinAttnParams p) {
    constexpr int D = 64;
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (wid >= (long long)p.T * p.n_heads) return;
    const int i = (int)(wid / p.n_heads), h = (int)(wid % p.n_heads);
    const int qd = p.n_heads * D, row = 3 * qd;
    // lane holds dims (lane, lane+32): rotate-half partners live in the same lane
    auto rope = [&](const float* v, int pos, float& a, float& b) {
        const float x1 = v[lane], x2 = v[lane + 32];
        const float c = p.rope_cos[(size_t)pos * 32 + lane], s = p.rope_sin[(size_t)pos * 32 + lane];
        a = x1 * c - x2 * s;
        b = x2 * c + x1 * s;
    };
    float q1, q2;
    rope(p.qkv + (size_t)i * row + h * D, p.pos0 + i, q1, q2);
    const int j0 = max(-p.hist, i - p.window + 1);
    float m = -INFINITY, l = 0.f, o1 = 0.f, o2 = 0.f;
    for (int j = j0; j <= i; ++j) {
        float k1, k2;
        rope(p.qkv + (long long)j * row + qd + h * D, p.pos0 + j, k1, k2);
        const float s = warp_sum(q1 * k1 + q2 * k2) * p.scale;
        const float mn = fmaxf(m, s);
        const float corr = expf(m - mn), pr = expf(s - mn);
        const float* v = p.qkv + (long long)j * row + 2 * qd + h * D;
        l = l * corr + pr;
        o1 = o1 * corr + pr * v[lane];
        o2 = o2 * corr + pr * v[lane + 32];
        m = mn;
    }
    p.out[(size_t)i * qd + h * D + lane] = o1 / l;
    p.out[(size_t)i * qd + h * D + lane + 32] = o2 / l;
}

}  // namespace lqt
