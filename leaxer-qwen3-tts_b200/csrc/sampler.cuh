// Fused sampler: special-token mask -> temperature -> top-k -> softmax -> top-p -> renormalise ->
// seeded Philox4x32-10 categorical draw, bit-for-bit the semantics of the reference's host sampler
// (src/tts_onnx.cpp:803-807, 878-950) with the RNG replaced by Philox
// (key = (seed, utterance id), counter = (frame, codebook, 0, 0)).
//
// Bit-exactness rules shared with oracle/qwen3_tts_oracle.py:
//   * x/T is an IEEE f32 division; exp is evaluated in f64 and rounded to f32;
//   * every sum the reference accumulates left-to-right in f32 (softmax denominator, top-p
//     running sum, renormalisation, cdf) is accumulated serially in index order here too;
//   * top-k keeps ties at the threshold; top-p order is (prob desc, index asc).
// The kernel also fuses the frame-loop glue that follows a draw in the reference
// (:812 EOS check, :818-830 frame store + embedding sum, :833-842 trailing text / pad add,
// :854-860 / :867-868 predictor input rows).
#pragma once
#include "common.cuh"

namespace lqt {

constexpr int SMP_THREADS = 1024;
constexpr int SMP_MAXV = 4096;

// sampling knobs live in device memory so a captured frame graph can be replayed with new values
struct SamplingDev {
    float temperature, top_p;
    int top_k, greedy;
    uint32_t seed, utt;
};

struct SampleParams {
    const float* logits; int V;
    int n_splits; long long split_stride;   // batched path: the logits are the sum of n_splits split-K partials (0/1 = plain array)
    int mask_lo, mask_hi, mask_keep;     // logits[i] = -inf for mask_lo <= i < mask_hi, i != mask_keep
    const SamplingDev* sp;
    uint32_t frame_imm;
    int codebook;                        // 0 = talker code, 1..15 = sub-codes
    GenState* st;                        // nullable: standalone call uses frame_imm
    int* token_out;                      // nullable
    // fused glue (nullable table => no glue)
    const __nv_bfloat16* embed_table;    // [rows][H]: codec_embed (cb 0) or cp_embed[cb-1]
    int H;
    float* cp_in;                        // next code-predictor input row [H]
    float* next_in;                      // next talker input accumulator [H]
    const float* trailing;               // [trailing_len][H]
    const float* tts_pad;                // [H]
    long long* codes_out;                // [max_frames][16]
    const long long* forced;             // nullable [n_forced][16]
    float* trace; int trace_stride;      // nullable [max_frames][16][trace_stride]
    int eos_id; int n_codebooks;         // 2150, 16
};

LQT_DEVINL void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

LQT_DEVINL uint32_t float_key(float x) {          // order-preserving float -> uint
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// block-wide exclusive prefix of a 0/1 flag; returns the prefix, *total = block total
template <int NT>
LQT_DEVINL int block_excl_scan_flag(int flag, int* warp_tot, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    const int in_warp = __popc(b & ((1u << lane) - 1u));
    __syncthreads();                                  // protect warp_tot reuse
    if (lane == 0) warp_tot[warp] = __popc(b);
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
        const int c = warp_tot[w];
        if (w < warp) base += c;
        tot += c;
    }
    *total = tot;
    return base + in_warp;
}

// Fast path of the draw for 0 < top_k <= 64 (bit for bit the results of the general path in sample_block): a 256-bin
// histogram of (max - x) * 16 finds the bin that holds the k-th largest value, the <= 64 candidates up to that bin are
// compacted in index order and ONE warp finishes -- exact top-k by rank (ties at the threshold survive, src/tts_onnx.cpp
// :917-927), softmax (:907-915), top-p (:929-950), renormalisation and the categorical draw (:893-905), every sum serial in
// the reference's order (a candidate that drops out is an exact 0.0f, neutral in a serial sum). Same scheme as the
// persistent frame kernel's sampler (frame_kernel.cuh fk_sample_fast). Returns -1 if the shape does not fit (flat logits:
// more than 64 candidates): the caller then runs the general path.
template <int NT>
LQT_DEVINL int sample_fast(const SamplingDev& sp, const float* s_x, float* pr, float* spr, int* idx, int V,
                           uint32_t frame, int codebook, int* hist, int* warp_tot, float* redf, int* sel_s, int* tok_s) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = sp.top_k;
    float mx = -INFINITY;
    for (int i = tid; i < V; i += NT) mx = fmaxf(mx, s_x[i]);
    mx = warp_max(mx);
    if (tid < 256) hist[tid] = 0;
    if (lane == 0) redf[warp] = mx;
    __syncthreads();
    mx = redf[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) mx = fmaxf(mx, redf[w]);
    for (int i = tid; i < V; i += NT) {
        const float dlt = (mx - s_x[i]) * 16.0f;                    // >= 0; +inf for masked entries
        if (dlt < 255.0f) atomicAdd(&hist[(int)dlt], 1);
    }
    __syncthreads();
    if (warp == 0) {
        int h[8], loc = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { h[q] = hist[lane * 8 + q]; loc += h[q]; }
        int inc = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        int cum = inc - loc, myB = -1;
#pragma unroll
        for (int q = 0; q < 8; ++q) { cum += h[q]; if (myB < 0 && cum >= k) myB = lane * 8 + q; }
        const unsigned bal = __ballot_sync(0xffffffffu, myB >= 0);
        const int B = bal ? __shfl_sync(0xffffffffu, myB, __ffs(bal) - 1) : -1;
        if (lane == 0) *sel_s = B;
    }
    __syncthreads();
    const int B = *sel_s;
    if (B < 0) return -1;
    const float lim = (float)(B + 1);
    // compaction in INDEX order: thread t owns the contiguous range [t * per, (t + 1) * per)
    const int per = (V + NT - 1) / NT;
    const int i0 = min(V, tid * per), i1 = min(V, i0 + per);
    int cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += ((mx - s_x[i]) * 16.0f < lim) ? 1 : 0;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int base = 0, n_c = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) { const int t = warp_tot[w]; if (w < warp) base += t; n_c += t; }
    if (n_c > 64) return -1;
    int wpos = base + inc - cnt;
    for (int i = i0; i < i1; ++i) {
        const float v = s_x[i];
        if ((mx - v) * 16.0f < lim) { idx[wpos] = i; pr[wpos] = v; ++wpos; }
    }
    __syncthreads();
    if (warp == 0) {
        // lane l owns candidates l and l + 32; arrays padded to 64 entries and read eight at a time
        const int e0 = lane, e1 = lane + 32;
        const int n8 = (n_c + 7) & ~7;
        const float x0 = (e0 < n_c) ? pr[e0] : -INFINITY, x1 = (e1 < n_c) ? pr[e1] : -INFINITY;
        __syncwarp();
        if (e0 >= n_c) pr[e0] = -INFINITY;
        if (e1 >= n_c) pr[e1] = -INFINITY;
        __syncwarp();
        int gt0 = 0, gt1 = 0;
        for (int j = 0; j < n8; ++j) { const float xv = pr[j]; gt0 += (xv > x0) ? 1 : 0; gt1 += (xv > x1) ? 1 : 0; }
        const bool sv0 = e0 < n_c && gt0 < k, sv1 = e1 < n_c && gt1 < k;      // x >= (k-th largest)  <=>  fewer than k values above it
        const unsigned b0 = __ballot_sync(0xffffffffu, sv0), b1 = __ballot_sync(0xffffffffu, sv1);
        const int ns = __popc(b0) + __popc(b1);
        float pr0 = sv0 ? (float)exp((double)(x0 - mx)) : 0.f, pr1 = sv1 ? (float)exp((double)(x1 - mx)) : 0.f;
        spr[e0] = pr0; spr[e1] = pr1;
        __syncwarp();
        float sum = 0.f;
        for (int i = 0; i < n8; ++i) sum += spr[i];                              // serial, index order (every lane redundantly)
        pr0 = pr0 / sum; pr1 = pr1 / sum;                                        // (dropped candidates stay 0)
        __syncwarp();
        spr[e0] = pr0; spr[e1] = pr1;
        __syncwarp();
        if (sp.top_p < 1.0f) {
            int r0 = 0, r1 = 0;                                                  // position in descending order (ties: index order)
            for (int j = 0; j < n8; ++j) {
                const float pv = spr[j];
                r0 += (pv > pr0 || (pv == pr0 && j < e0)) ? 1 : 0;
                r1 += (pv > pr1 || (pv == pr1 && j < e1)) ? 1 : 0;
            }
            pr[e0] = 0.f; pr[e1] = 0.f;
            __syncwarp();
            if (sv0) pr[r0] = pr0;                                               // survivors only: their positions are 0..ns-1
            if (sv1) pr[r1] = pr1;
            __syncwarp();
            int cut = ns;
            float cs = 0.f;
            for (int r = 0; r < ns; ++r) { cs += pr[r]; if (cs > sp.top_p) { cut = r + 1; break; } }
            if (r0 >= cut) pr0 = 0.f;
            if (r1 >= cut) pr1 = 0.f;
            __syncwarp();
            spr[e0] = pr0; spr[e1] = pr1;
            __syncwarp();
            float s2 = 0.f;
            for (int i = 0; i < n8; ++i) s2 += spr[i];
            if (s2 > 0.f) { pr0 = pr0 / s2; pr1 = pr1 / s2; }
            __syncwarp();
            spr[e0] = pr0; spr[e1] = pr1;
            __syncwarp();
        }
        uint32_t r4[4];
        philox4x32_10(frame, (uint32_t)codebook, 0u, 0u, sp.seed, sp.utt, r4);
        const float u01 = (float)(r4[0] >> 8) * 5.9604644775390625e-08f;
        const int first = b0 ? (__ffs(b0) - 1) : (b1 ? 32 + __ffs(b1) - 1 : 0);
        float cdf = 0.f; int last = idx[first];
        for (int i = 0; i < n8; ++i) {
            const float pv = spr[i];
            if (pv > 0.f) {
                cdf += pv; last = idx[i];
                if (cdf > u01) break;
            }
        }
        if (lane == 0) *tok_s = last;
    }
    __syncthreads();
    return *tok_s;
}

// dynamic smem: V * (float x + int idx + float p + float sorted_p + int rank) = 20 V bytes
// One draw + glue by one CTA of NT threads (called by sample_kernel and by the batched path's bsample_kernel).
// NT = threads of the CTA: 1024 for the single-utterance kernel, 256 in the batched path (one CTA per slot, many slots per SM).
template <int NT>
LQT_DEVINL void sample_block(const SampleParams& p) {
    extern __shared__ unsigned char smp_raw[];
    float* s_x    = reinterpret_cast<float*>(smp_raw);          // (masked, tempered) logits
    int*   s_idx  = reinterpret_cast<int*>(s_x + p.V);          // survivor -> original index
    float* s_p    = reinterpret_cast<float*>(s_idx + p.V);      // survivor e / prob
    float* s_sp   = s_p + p.V;                                  // probs in top-p order
    int*   s_rank = reinterpret_cast<int*>(s_sp + p.V);         // survivor -> rank
    __shared__ int hist[256];
    __shared__ int warp_tot[NT / 32];
    __shared__ float redf[NT / 32];
    __shared__ int redi[NT / 32];
    __shared__ uint32_t sel_prefix; __shared__ int sel_k;
    __shared__ int tok_s; __shared__ int cutoff_s; __shared__ float sum_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    GenState* st = p.st;
    uint32_t frame = p.frame_imm;
    if (st) {
        const int done = st->done, f = st->frame, mx = st->max_frames;
        __syncthreads();
        if (done) return;
        if (p.codebook == 0 && f >= mx) { if (tid == 0) st->done = 1; return; }
        frame = (uint32_t)f;
    }
    const int V = p.V;
    const SamplingDev sp = *p.sp;
    const bool temper = !sp.greedy && sp.temperature > 0.0f && sp.temperature != 1.0f;

    // ---- load + mask (:803-807) + temperature (:882-884) ---------------------------------------
    for (int i = tid; i < V; i += NT) {
        float v = p.logits[i];
        if (p.n_splits > 1) {                                   // fixed order, loads in flight
            float w[8];
            for (int q0 = 1; q0 < p.n_splits; q0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) w[u] = (q0 + u < p.n_splits) ? __ldcg(p.logits + (size_t)(q0 + u) * p.split_stride + i) : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) if (q0 + u < p.n_splits) v += w[u];
            }
        }
        if (i >= p.mask_lo && i < p.mask_hi && i != p.mask_keep) v = -INFINITY;
        if (p.trace) p.trace[((size_t)frame * p.n_codebooks + p.codebook) * p.trace_stride + i] = v;
        if (temper) v = v / sp.temperature;
        s_x[i] = v;
    }
    __syncthreads();

    int token;
    if (sp.greedy) {
        // argmax, lowest index on ties
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int i = tid; i < V; i += NT) {
            const float v = s_x[i];
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { redf[warp] = bv; redi[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            float v = redf[0]; int i = redi[0];
            for (int w = 1; w < NT / 32; ++w)
                if (redf[w] > v || (redf[w] == v && redi[w] < i)) { v = redf[w]; i = redi[w]; }
            tok_s = (i == 0x7fffffff) ? 0 : i;
        }
        __syncthreads();
        token = tok_s;
    } else if (sp.top_k > 0 && sp.top_k <= 64 && sp.top_k < V &&
               (token = sample_fast<NT>(sp, s_x, s_p, s_sp, s_idx, V, frame, p.codebook, hist, warp_tot, redf, &sel_k, &tok_s)) >= 0) {
        // common case (top-k 50): done by the histogram / one-warp path above
    } else {
        __syncthreads();
        // ---- top-k threshold = k-th largest value (:917-927), 4 x 8-bit radix select ----------
        float thr = -INFINITY;
        if (sp.top_k > 0 && sp.top_k < V) {
            if (tid == 0) { sel_prefix = 0u; sel_k = sp.top_k; }
            for (int shift = 24; shift >= 0; shift -= 8) {
                if (tid < 256) hist[tid] = 0;
                __syncthreads();
                const uint32_t prefix = sel_prefix;
                const uint32_t himask = (shift == 24) ? 0u : (0xffffffffu << (shift + 8));
                for (int i = tid; i < V; i += NT) {
                    const uint32_t key = float_key(s_x[i]);
                    if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
                }
                __syncthreads();
                if (tid == 0) {
                    int k = sel_k, b = 255;
                    for (; b > 0; --b) { if (hist[b] >= k) break; k -= hist[b]; }
                    sel_k = k;
                    sel_prefix = prefix | ((uint32_t)b << shift);
                }
                __syncthreads();
            }
            const uint32_t kk = sel_prefix;            // key of the k-th largest value
            const uint32_t u = (kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk;
            thr = __uint_as_float(u);
        }
        // ---- max over survivors ----------------------------------------------------------------
        float mx = -INFINITY;
        for (int i = tid; i < V; i += NT) {
            const float v = s_x[i];
            if (!(v < thr) && v != -INFINITY) mx = fmaxf(mx, v);
        }
        mx = warp_max(mx);
        if (lane == 0) redf[warp] = mx;
        __syncthreads();
        mx = redf[0];
#pragma unroll
        for (int w = 1; w < NT / 32; ++w) mx = fmaxf(mx, redf[w]);
        // ---- compaction in index order; e = (float)exp((double)(x - m)) (:907-915) ------------
        int n_surv = 0;
        for (int base = 0; base < V; base += NT) {
            const int i = base + tid;
            const float v = (i < V) ? s_x[i] : -INFINITY;
            const bool sv = (i < V) && !(v < thr) && (v != -INFINITY);
            int tot;
            const int pre = block_excl_scan_flag<NT>(sv ? 1 : 0, warp_tot, &tot);
            if (sv) {
                s_idx[n_surv + pre] = i;
                s_p[n_surv + pre] = (float)exp((double)(v - mx));
            }
            n_surv += tot;
        }
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int i = 0; i < n_surv; ++i) s += s_p[i];
            sum_s = s;
        }
        __syncthreads();
        {
            const float s = sum_s;
            for (int i = tid; i < n_surv; i += NT) s_p[i] = s_p[i] / s;
        }
        __syncthreads();
        // ---- top-p (:929-950) + renormalise (:893-898) -----------------------------------------
        if (sp.top_p < 1.0f) {
            for (int i = tid; i < n_surv; i += NT) {
                const float pi = s_p[i];
                int r = 0;
                for (int j = 0; j < n_surv; ++j) {
                    const float pj = s_p[j];
                    r += (pj > pi || (pj == pi && j < i)) ? 1 : 0;
                }
                s_rank[i] = r;
                s_sp[r] = pi;
            }
            __syncthreads();
            if (tid == 0) {
                float c = 0.f; int cut = n_surv;
                for (int r = 0; r < n_surv; ++r) {
                    c += s_sp[r];
                    if (c > sp.top_p) { cut = r + 1; break; }
                }
                cutoff_s = cut;
            }
            __syncthreads();
            const int cut = cutoff_s;
            for (int i = tid; i < n_surv; i += NT)
                if (s_rank[i] >= cut) s_p[i] = 0.f;
            __syncthreads();
            if (tid == 0) {
                float s = 0.f;
                for (int i = 0; i < n_surv; ++i) if (s_p[i] > 0.f) s += s_p[i];
                sum_s = s;
            }
            __syncthreads();
            const float s2 = sum_s;
            if (s2 > 0.f)
                for (int i = tid; i < n_surv; i += NT) s_p[i] = s_p[i] / s2;
            __syncthreads();
        }
        // ---- categorical draw in index order: smallest i with cdf[i] > u ------------------------
        if (tid == 0) {
            uint32_t r4[4];
            philox4x32_10(frame, (uint32_t)p.codebook, 0u, 0u, sp.seed, sp.utt, r4);
            const float u = (float)(r4[0] >> 8) * 5.9604644775390625e-08f;     // * 2^-24
            float c = 0.f; int last = (n_surv > 0) ? s_idx[0] : 0;
            for (int i = 0; i < n_surv; ++i) {
                const float pi = s_p[i];
                if (pi > 0.f) {
                    c += pi; last = s_idx[i];
                    if (c > u) break;
                }
            }
            tok_s = last;
        }
        __syncthreads();
        token = tok_s;
    }

    // ---- glue -----------------------------------------------------------------------------------
    if (p.forced && st && (int)frame < st->n_forced)
        token = (int)p.forced[(size_t)frame * p.n_codebooks + p.codebook];
    if (p.token_out && tid == 0) *p.token_out = token;
    if (!p.embed_table) return;
    if (p.codebook == 0 && token == p.eos_id) {          // :812
        if (tid == 0) st->done = 1;
        return;
    }
    if (tid == 0) {
        p.codes_out[(size_t)frame * p.n_codebooks + p.codebook] = token;
        if (p.codebook == p.n_codebooks - 1) { st->n_frames = (int)frame + 1; st->frame = (int)frame + 1; }
    }
    const bool last_cb = (p.codebook == p.n_codebooks - 1);
    const bool use_trailing = st && ((int)frame < st->trailing_len);
    const __nv_bfloat16* row = p.embed_table + (size_t)token * p.H;
    for (int h = tid; h < p.H; h += NT) {
        const float e = __bfloat162float(row[h]);
        p.cp_in[h] = e;
        float acc = (p.codebook == 0) ? e : p.next_in[h] + e;               // :824-830
        if (last_cb) acc += use_trailing ? p.trailing[(size_t)frame * p.H + h] : p.tts_pad[h];   // :833-842
        p.next_in[h] = acc;
    }
}

__global__ void __launch_bounds__(SMP_THREADS, 1)
sample_kernel(const SampleParams p) { sample_block<SMP_THREADS>(p); }

}  // namespace lqt
