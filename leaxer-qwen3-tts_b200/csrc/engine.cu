// C-ABI implementation (include/lqt_b200.h): weight loading, device buffers, the captured frame
// graph and the per-graph entry points that stand where the reference's Ort::Session::Run calls
// were (src/tts_onnx.cpp:545-776), plus the device frame loop (:782-872).
#include "../../include/lqt_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include "attention.cuh"
#include "common.cuh"
#include "frame_kernel.cuh"
#include "gemv.cuh"
#include "lqw_loader.h"
#include "batched.cuh"
#include "sampler.cuh"
#include "tc_conv.cuh"
#include "tc_gemm.cuh"
#include "vocoder.cuh"

using namespace lqt;
typedef __nv_bfloat16 bf16;

namespace {

// reference constants (src/tts_onnx.h:39-62)
constexpr long long TTS_BOS = 151672, TTS_EOS = 151673, TTS_PAD = 151671;
constexpr int CODEC_BOS = 2149, CODEC_EOS = 2150, CODEC_PAD = 2148;
constexpr int CODEC_THINK = 2154, CODEC_NOTHINK = 2155, CODEC_THINK_BOS = 2156, CODEC_THINK_EOS = 2157;
constexpr int KV_PAGE_SHIFT = 6;                 // talker KV pages of 64 positions
constexpr int KV_PAGE = 1 << KV_PAGE_SHIFT;
constexpr int ATT_NSPLIT = 16;
constexpr int CP_PAGE_SHIFT = 5;                 // predictor: one fp32 page of 32 positions
constexpr int N_CODEBOOKS = 16;

thread_local std::string g_create_error;    // why the last lqt_create on THIS thread failed (no handle exists to hold it)

struct LayerW {
    const float *ln1 = nullptr, *ln2 = nullptr, *qnorm = nullptr, *knorm = nullptr, *ls1 = nullptr, *ls2 = nullptr;
    const bf16 *wqkv = nullptr, *wo = nullptr, *wgate = nullptr, *wup = nullptr, *wdown = nullptr;
};

struct Spec {
    int hidden, layers, heads, kv_heads, head_dim, inter, vocab, max_pos;
    int cp_hidden, cp_layers, cp_heads, cp_kv_heads, cp_inter, cp_vocab, cp_steps, cp_max_pos;
    int text_vocab, text_dim;
    int voc_codebook_size, voc_codebook_dim, voc_rvq_out, voc_hidden, voc_layers, voc_heads, voc_head_dim,
        voc_inter, voc_window, voc_max_pos, voc_decoder_dim;
    std::vector<int> voc_up_ratios, voc_up_rates;
    int spk_mels, spk_channels, spk_layers;
    float rms_eps, voc_rms_eps;
    int samples_per_frame;
};

struct VocBlockW {
    const float *snake_a, *snake_b, *tconv_b;
    const bf16* tconv_w;
    struct Res { const float *s1a, *s1b, *c1b, *s2a, *s2b, *c2b; const bf16 *c1w, *c2w; } res[3];
    int cin, cout, stride;
};
struct VocUpW {
    const bf16 *tconv_w, *pw1_w, *pw2_w;
    const float *tconv_b, *dw_w, *dw_b, *ln_w, *ln_b, *pw1_b, *pw2_b, *gamma;
    int factor;
};

}  // namespace

struct lqt_batch;
namespace { struct VocTcModel; }

struct lqt_engine {
    int device = 0;
    lqt_batch* batch = nullptr;               // batched path context (batch_engine.inl), created on first use
    // streaming vocoder (lqt_vocoder_stream_*): frames decoded so far + the tail every layer with left context keeps between chunks
    struct VocStream { int t0 = 0; std::map<std::string, std::pair<void*, size_t>> halo; } voc_stream;
    float *mel_window = nullptr, *mel_tw_re = nullptr, *mel_tw_im = nullptr; int* mel_tri = nullptr;   // log-mel tables (lqt_log_mel)
    VocTcModel* voc_tc = nullptr;             // tcgen05 vocoder decoder (voc_tc.inl): padded weights, SnakeBeta constants, plane buffers
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    Spec sp{};
    LqwFile f_text, f_codec, f_cpe, f_talker, f_cp, f_voc, f_spk;
    bool has_spk = false;
    std::string err;
    lqt_stats stats{};

    // talker / predictor weights
    std::vector<LayerW> tl, cl;
    const float *t_norm = nullptr, *t_cos = nullptr, *t_sin = nullptr;
    const bf16* t_head = nullptr;
    const float *c_norm = nullptr, *c_cos = nullptr, *c_sin = nullptr, *c_inproj_b = nullptr;
    const bf16 *c_heads = nullptr, *c_inproj_w = nullptr;
    const bf16 *text_embed = nullptr, *fc1w = nullptr, *fc2w = nullptr, *codec_embed = nullptr, *cp_embed = nullptr;
    const float *fc1b = nullptr, *fc2b = nullptr;
    // vocoder weights
    std::vector<LayerW> vl;
    const float *v_norm = nullptr, *v_cos = nullptr, *v_sin = nullptr;
    const bf16 *rvq_sem_cb = nullptr, *rvq_aco_cb = nullptr, *rvq_sem_proj = nullptr, *rvq_aco_proj = nullptr;
    const bf16 *pre_conv_w = nullptr, *dec_in_w = nullptr;
    const float *pre_conv_b = nullptr, *dec_in_b = nullptr, *out_sa = nullptr, *out_sb = nullptr, *out_w = nullptr, *out_b = nullptr;
    std::vector<VocUpW> vup;
    std::vector<VocBlockW> vblk;

    // decode buffers
    float *x = nullptr, *qkv = nullptr, *attn = nullptr, *act = nullptr, *logits = nullptr, *last_hidden = nullptr;
    float *cx = nullptr, *cxin = nullptr, *cqkv = nullptr, *cattn = nullptr, *cact = nullptr, *clogits = nullptr;
    float *cp_in = nullptr, *next_in = nullptr;
    float *partial = nullptr, *cpartial = nullptr;
    int *counters = nullptr, *ccounters = nullptr;
    void* kv_pool = nullptr;                  // bf16 (default) or fp32 pages
    bool kv_f32 = false;
    float* cp_kv = nullptr;
    int *page_tables = nullptr, *cp_page_table = nullptr, *cp_pos_consts = nullptr;
    int n_slots = 2, max_pages = 0;
    std::vector<int> slot_len;
    GenState* st = nullptr;
    GenState* st_host = nullptr;              // pinned
    SamplingDev* sampling_dev = nullptr;
    long long *codes_dev = nullptr, *forced_dev = nullptr;
    float *trailing_dev = nullptr, *tts_pad_dev = nullptr, *prompt_dev = nullptr;
    float* trace_dev = nullptr; int trace_stride = 0;
    int max_frames_cap = 0;
    int* token_dev = nullptr;
    // prompt building scratch
    long long* ids_dev = nullptr; float *tp_emb = nullptr, *tp_h = nullptr, *tp_out = nullptr; int tp_cap = 0;
    float* spk_dev = nullptr;
    // frame graphs keyed by slot*2 + trace
    std::map<int, cudaGraphExec_t> graphs;
    int kernels_per_frame = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_t0 = nullptr;
    // streaming path: finished frames are vocoded on a second stream while the frame kernel keeps generating -- a short first
    // chunk (first-audio latency), then `stream_chunk` frames at a time
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_chunk = nullptr, ev_first = nullptr, ev_end = nullptr;
    int first_chunk = 4;                      // frames of the first chunk (320 ms of audio); 0 = streaming off; $LQT_FIRST_CHUNK
    int stream_chunk = 25;                    // frames of every later chunk (2 s); $LQT_STREAM_CHUNK
    float* chunk_audio_dev = nullptr; size_t chunk_audio_cap = 0;
    float* chunk_audio_out = nullptr; int64_t chunk_audio_out_cap = 0;    // caller's buffer of the running lqt_synthesize_tokens call
    bool chunk_pending = false;               // stream2 holds work of the current call (synchronised on every exit path)
    // persistent frame kernel (frame_kernel.cuh)
    int frame_impl = 0;                       // 0 = persistent kernel, 1 = v1 graph of kernels
    std::vector<void*> fk_allocs;             // regrouped weights, layer tables, activation buffers
    FkStack fk_talker{}, fk_cp{};
    std::vector<FkLayer> fk_tl, fk_cl;
    const bf16 *fk_t_head = nullptr, *fk_c_heads = nullptr, *fk_c_inproj = nullptr; long long fk_c_head_stride = 0;
    float* fk_cp_kv = nullptr;                // per-CTA predictor KV copies of the frame kernel
    uint2* fk_arena = nullptr; size_t fk_arena_words = 0;     // all LL exchange buffers (zeroed at every launch)
    uint2 *fk_pa = nullptr, *fk_cxin = nullptr, *fk_logits_ll = nullptr, *fk_clogits_ll = nullptr;
    unsigned* fk_ctrl = nullptr;
    unsigned* fk_ctrl_host = nullptr;         // pinned
    int* fk_progress_host = nullptr;          // pinned: frames completed by the running frame kernel (streaming vocoder)
    bool fk_progress_on = false;
    int voc_reserved_frames = 0;              // largest streaming chunk the vocoder workspaces are sized for
    lqt_audio_callback on_audio = nullptr; void* on_audio_user = nullptr;     // lqt_synthesize_stream: called per finished PCM chunk
    std::vector<cudaEvent_t> chunk_events;    // one per PCM chunk in flight (created on demand, reused)
    bool streamed_last = false;               // the last lqt_synthesize_tokens call went through synthesize_streaming
    unsigned long long* fk_dbg = nullptr; int fk_dbg_cap = 0, fk_dbg_cta = 0;
    FkSmemOffsets fk_so{};
    bool fk_wide = false;
    int fk_ncta = 0;                          // CTAs of the frame kernel = 8 x co-resident clusters
    bool fk_coop = true;
    size_t fk_smem = 0;
    // vocoder workspace
    std::map<std::string, std::pair<float*, size_t>> ws;
    long long* voc_codes_dev = nullptr; size_t voc_codes_cap = 0;
    float* audio_dev = nullptr; size_t audio_cap = 0;
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

// A tensor the kernels will index with dimensions derived from the model spec: it must exist with exactly that dtype and
// shape, otherwise a smaller vocab / hidden size or a wrong dtype in the file would become an out-of-bounds device read.
template <typename T>
const T* need(lqt_engine* h, const LqwFile& f, const std::string& name, bool& ok, std::initializer_list<long long> dims) {
    const DevTensor* t = f.find(name);
    if (!t) { if (ok) h->err = "missing tensor " + name; ok = false; return nullptr; }
    const int want_dt = std::is_same<T, float>::value ? 1 : 0;
    bool same = t->dtype == want_dt && t->dims.size() == dims.size();
    if (same) { size_t i = 0; for (long long d : dims) same = same && t->dims[i++] == d; }
    if (!same) {
        if (ok) {
            std::string got, want;
            for (auto d : t->dims) got += (got.empty() ? "" : "x") + std::to_string(d);
            for (auto d : dims) want += (want.empty() ? "" : "x") + std::to_string(d);
            h->err = "tensor " + name + ": expected " + (want_dt ? "f32 [" : "bf16 [") + want + "], file has " + (t->dtype ? "f32 [" : "bf16 [") + got + "]";
        }
        ok = false;
        return nullptr;
    }
    return reinterpret_cast<const T*>(t->ptr);
}

// one decoder layer: hidden H, q width qd, kv width kvd, head dim D, MLP width I
bool load_layers(lqt_engine* h, const LqwFile& f, const std::string& pre, int n, int H, int qd, int kvd, int D, int I,
                 bool qk_norm, bool ls, std::vector<LayerW>& out) {
    bool ok = true;
    out.resize(n);
    for (int i = 0; i < n; ++i) {
        const std::string p = pre + "l" + std::to_string(i) + ".";
        LayerW& L = out[i];
        L.ln1 = need<float>(h, f, p + "ln1", ok, {H});   L.ln2 = need<float>(h, f, p + "ln2", ok, {H});
        L.wqkv = need<bf16>(h, f, p + "wqkv", ok, {qd + 2 * kvd, H});  L.wo = need<bf16>(h, f, p + "wo", ok, {H, qd});
        L.wgate = need<bf16>(h, f, p + "wgate", ok, {I, H}); L.wup = need<bf16>(h, f, p + "wup", ok, {I, H});
        L.wdown = need<bf16>(h, f, p + "wdown", ok, {H, I});
        if (qk_norm) { L.qnorm = need<float>(h, f, p + "qnorm", ok, {D}); L.knorm = need<float>(h, f, p + "knorm", ok, {D}); }
        if (ls) { L.ls1 = need<float>(h, f, p + "ls1", ok, {H}); L.ls2 = need<float>(h, f, p + "ls2", ok, {H}); }
    }
    return ok;
}

template <typename T>
int dalloc(lqt_engine* h, T** p, size_t n) {
    CK(cudaMalloc((void**)p, n * sizeof(T)));
    CK(cudaMemset(*p, 0, n * sizeof(T)));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
void launch_gemv(lqt_engine* h, GemvParams p, bool glu) {
    const int opi = glu ? 2 : 4;
    const int warps = (p.N + opi - 1) / opi;
    int ctas = (warps + GEMV_WARPS - 1) / GEMV_WARPS;
    ctas = std::max(1, std::min(ctas, h->num_sms * 4));
    const int kpad = (p.K + 255) & ~255;
    const int MT = p.M <= 1 ? 1 : (p.M == 2 ? 2 : 4);
    dim3 grid(ctas, (p.M + MT - 1) / MT);
    const size_t smem = (size_t)MT * kpad * sizeof(float);
#define LQT_GEMV_LAUNCH(MTV, GLUV) gemv_kernel<MTV, GLUV><<<grid, GEMV_THREADS, smem, h->stream>>>(p)
    if (glu) {
        if (MT == 1) LQT_GEMV_LAUNCH(1, true); else if (MT == 2) LQT_GEMV_LAUNCH(2, true); else LQT_GEMV_LAUNCH(4, true);
    } else {
        if (MT == 1) LQT_GEMV_LAUNCH(1, false); else if (MT == 2) LQT_GEMV_LAUNCH(2, false); else LQT_GEMV_LAUNCH(4, false);
    }
#undef LQT_GEMV_LAUNCH
    h->stats.kernel_launches++;
}

GemvParams gemv_params(const bf16* W, int N, int K, const float* x, float* y, int M = 1) {
    GemvParams p{};
    p.W = W; p.N = N; p.K = K; p.x = x; p.y = y; p.M = M;
    p.x_stride = K; p.y_stride = N; p.res_stride = N;
    return p;
}

struct XfmrCtx {               // one decoder stack (talker or predictor) in decode mode
    const std::vector<LayerW>* layers;
    int H, heads, kv_heads, inter;
    float eps;
    const float *cos, *sin;
    float *x, *qkv, *attn, *act, *partial;
    int* counters;
    void* kv_pool; const int* page_table; int page_shift; long long page_stride; bool kv_bf16; bool paged;
    int nsplit;
    const int* done;
};

// one token through all layers: x_in [H] -> ctx.x [H] (residual stream). pos_ptr = device position.
void run_stack(lqt_engine* h, const XfmrCtx& c, const float* x_in, const int* pos_ptr) {
    const int D = ATT_D, qd = c.heads * D, kvd = c.kv_heads * D;
    const int PS = 1 << c.page_shift;
    const int nl = (int)c.layers->size();
    for (int l = 0; l < nl; ++l) {
        const LayerW& L = (*c.layers)[l];
        const float* xin = (l == 0) ? x_in : c.x;
        {   // RMSNorm + QKV
            GemvParams p = gemv_params(L.wqkv, qd + 2 * kvd, c.H, xin, c.qkv);
            p.norm_w = L.ln1; p.eps = c.eps; p.done = c.done;
            launch_gemv(h, p, false);
        }
        {   // q/k norm + RoPE + KV append + attention
            AttnParams a{};
            a.qkv = c.qkv; a.qnorm = L.qnorm; a.knorm = L.knorm; a.rope_cos = c.cos; a.rope_sin = c.sin;
            a.pos_ptr = pos_ptr; a.kv_pool = c.kv_pool; a.page_table = c.page_table;
            a.partial = c.partial; a.counters = c.counters; a.out = c.attn; a.done = c.done;
            a.page_stride = c.page_stride;
            a.layer_off = (long long)l * 2 * c.kv_heads * PS * D;
            a.page_shift = c.page_shift; a.n_kv = c.kv_heads; a.eps = c.eps;
            a.scale = 1.0f / sqrtf((float)D);
            dim3 grid(c.kv_heads, c.nsplit);
            const int max_pages = c.paged ? h->max_pages : 1;
            const size_t smem = (size_t)2 * ((max_pages + c.nsplit - 1) / c.nsplit) * PS * sizeof(float);
            if (c.kv_bf16) attn_decode_kernel<bf16, 2><<<grid, ATT_THREADS, smem, h->stream>>>(a);
            else           attn_decode_kernel<float, 2><<<grid, ATT_THREADS, smem, h->stream>>>(a);
            h->stats.kernel_launches++;
        }
        {   // O projection + residual
            GemvParams p = gemv_params(L.wo, c.H, qd, c.attn, c.x);
            p.residual = xin; p.scale = L.ls1; p.done = c.done;
            launch_gemv(h, p, false);
        }
        {   // RMSNorm + SwiGLU
            GemvParams p = gemv_params(L.wgate, c.inter, c.H, c.x, c.act);
            p.W2 = L.wup; p.norm_w = L.ln2; p.eps = c.eps; p.done = c.done;
            launch_gemv(h, p, true);
        }
        {   // down projection + residual
            GemvParams p = gemv_params(L.wdown, c.H, c.inter, c.act, c.x);
            p.residual = c.x; p.scale = L.ls2; p.done = c.done;
            launch_gemv(h, p, false);
        }
    }
}

XfmrCtx talker_ctx(lqt_engine* h, int slot, const int* done) {
    XfmrCtx c{};
    c.layers = &h->tl; c.H = h->sp.hidden; c.heads = h->sp.heads; c.kv_heads = h->sp.kv_heads; c.inter = h->sp.inter;
    c.eps = h->sp.rms_eps; c.cos = h->t_cos; c.sin = h->t_sin;
    c.x = h->x; c.qkv = h->qkv; c.attn = h->attn; c.act = h->act; c.partial = h->partial; c.counters = h->counters;
    c.kv_pool = h->kv_pool; c.page_table = h->page_tables + (size_t)slot * h->max_pages;
    c.page_shift = KV_PAGE_SHIFT;
    c.page_stride = (long long)h->sp.layers * 2 * h->sp.kv_heads * KV_PAGE * ATT_D;
    c.kv_bf16 = !h->kv_f32; c.paged = true; c.nsplit = ATT_NSPLIT; c.done = done;
    return c;
}

XfmrCtx cp_ctx(lqt_engine* h, const int* done) {
    XfmrCtx c{};
    c.layers = &h->cl; c.H = h->sp.cp_hidden; c.heads = h->sp.cp_heads; c.kv_heads = h->sp.cp_kv_heads; c.inter = h->sp.cp_inter;
    c.eps = h->sp.rms_eps; c.cos = h->c_cos; c.sin = h->c_sin;
    c.x = h->cx; c.qkv = h->cqkv; c.attn = h->cattn; c.act = h->cact; c.partial = h->cpartial; c.counters = h->ccounters;
    c.kv_pool = h->cp_kv; c.page_table = h->cp_page_table; c.page_shift = CP_PAGE_SHIFT;
    c.page_stride = (long long)h->sp.cp_layers * 2 * h->sp.cp_kv_heads * (1 << CP_PAGE_SHIFT) * ATT_D;
    c.kv_bf16 = false; c.paged = false; c.nsplit = 1; c.done = done;
    return c;
}

// talker token: x_in -> (optionally) logits + last_hidden
void run_talker_token(lqt_engine* h, int slot, const float* x_in, bool with_head, const int* done) {
    XfmrCtx c = talker_ctx(h, slot, done);
    run_stack(h, c, x_in, &h->st->pos);
    if (with_head) {
        GemvParams p = gemv_params(h->t_head, h->sp.vocab, h->sp.hidden, h->x, h->logits);
        p.norm_w = h->t_norm; p.eps = h->sp.rms_eps; p.xnorm_out = h->last_hidden; p.done = done;
        launch_gemv(h, p, false);
    }
}

// predictor token at position s (0..16). x_row is talker-width [hidden]; head_idx >= 0 -> logits.
void run_cp_token(lqt_engine* h, const float* x_row, int s, int head_idx, const int* done) {
    XfmrCtx c = cp_ctx(h, done);
    const float* xin = x_row;
    if (h->c_inproj_w) {          // 1.7B: talker width -> predictor width
        GemvParams p = gemv_params(h->c_inproj_w, h->sp.cp_hidden, h->sp.hidden, x_row, h->cxin);
        p.bias = h->c_inproj_b; p.done = done;
        launch_gemv(h, p, false);
        xin = h->cxin;
    }
    run_stack(h, c, xin, h->cp_pos_consts + s);
    if (head_idx >= 0) {
        GemvParams p = gemv_params(h->c_heads + (size_t)head_idx * h->sp.cp_vocab * h->sp.cp_hidden,
                                   h->sp.cp_vocab, h->sp.cp_hidden, h->cx, h->clogits);
        p.norm_w = h->c_norm; p.eps = h->sp.rms_eps; p.done = done;
        launch_gemv(h, p, false);
    }
}

__global__ void advance_kernel(GenState* st) {
    if (st->done) return;
    st->pos += 1;
}

void launch_sample(lqt_engine* h, SampleParams sp) {
    const size_t smem = (size_t)sp.V * 20;
    sample_kernel<<<1, SMP_THREADS, smem, h->stream>>>(sp);
    h->stats.kernel_launches++;
}

SampleParams frame_sample_params(lqt_engine* h, int codebook, bool trace) {
    SampleParams s{};
    s.sp = h->sampling_dev; s.codebook = codebook; s.st = h->st; s.token_out = nullptr;
    s.H = h->sp.hidden; s.cp_in = h->cp_in; s.next_in = h->next_in;
    s.trailing = h->trailing_dev; s.tts_pad = h->tts_pad_dev; s.codes_out = h->codes_dev;
    s.forced = h->forced_dev; s.eos_id = CODEC_EOS; s.n_codebooks = N_CODEBOOKS;
    if (trace) { s.trace = h->trace_dev; s.trace_stride = h->trace_stride; }
    if (codebook == 0) {
        s.logits = h->logits; s.V = h->sp.vocab;
        s.mask_lo = 2048; s.mask_hi = h->sp.vocab; s.mask_keep = CODEC_EOS;      // src/tts_onnx.cpp:803-807
        s.embed_table = h->codec_embed;
    } else {
        s.logits = h->clogits; s.V = h->sp.cp_vocab;
        s.embed_table = h->cp_embed + (size_t)(codebook - 1) * h->sp.cp_vocab * h->sp.hidden;
    }
    return s;
}

// one frame of loops A+B (src/tts_onnx.cpp:801-846): sample code0 -> 15 sub-codes -> talker step
void enqueue_frame(lqt_engine* h, int slot, bool trace) {
    const int* done = &h->st->done;
    launch_sample(h, frame_sample_params(h, 0, trace));
    run_cp_token(h, h->last_hidden, 0, -1, done);                 // row 0 = talker last_hidden (:859)
    for (int s = 1; s <= h->sp.cp_steps; ++s) {
        run_cp_token(h, h->cp_in, s, s - 1, done);                // row s = embedding of the previous code
        launch_sample(h, frame_sample_params(h, s, trace));
    }
    run_talker_token(h, slot, h->next_in, true, done);            // run_decode (:845)
    advance_kernel<<<1, 1, 0, h->stream>>>(h->st);
    h->stats.kernel_launches++;
}

int get_frame_graph(lqt_engine* h, int slot, bool trace, cudaGraphExec_t* out) {
    const int key = slot * 2 + (trace ? 1 : 0);
    auto it = h->graphs.find(key);
    if (it != h->graphs.end()) { *out = it->second; return 0; }
    cudaGraph_t g;
    const uint64_t before = h->stats.kernel_launches;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    enqueue_frame(h, slot, trace);
    CK(cudaStreamEndCapture(h->stream, &g));
    h->kernels_per_frame = (int)(h->stats.kernel_launches - before);
    h->stats.kernel_launches = before;                 // capture does not execute
    cudaGraphExec_t ge;
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaGraphDestroy(g);
    h->graphs[key] = ge;
    *out = ge;
    return 0;
}

}  // namespace

// ================================================================================================
// part 2: embeddings, prompt assembly, vocoder pipeline, generation loop, C-ABI
// ================================================================================================
namespace {

__global__ void gather_rows_kernel(const bf16* __restrict__ table, const long long* __restrict__ ids,
                                   int width, float* __restrict__ out) {
    const long long id = ids[blockIdx.x];
    const bf16* row = table + (size_t)id * width;
    for (int i = threadIdx.x; i < width; i += blockDim.x)
        out[(size_t)blockIdx.x * width + i] = __bfloat162float(row[i]);
}

// out[r][:] = tp[desc[r].x] (if >=0) + codec_embed[desc[r].y] (if >=0) + spk (if desc[r].z)
__global__ void assemble_rows_kernel(const int* __restrict__ desc, const float* __restrict__ tp,
                                     const bf16* __restrict__ codec, const float* __restrict__ spk,
                                     int H, float* __restrict__ out) {
    const int r = blockIdx.x;
    const int ti = desc[r * 3], ci = desc[r * 3 + 1], sf = desc[r * 3 + 2];
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        float v = 0.f;
        if (ti >= 0) v = tp[(size_t)ti * H + i];
        if (ci >= 0) v = v + __bfloat162float(codec[(size_t)ci * H + i]);
        if (sf) v = v + spk[i];
        out[(size_t)r * H + i] = v;
    }
}

float* wsbuf(lqt_engine* h, const std::string& name, size_t n) {
    auto& e = h->ws[name];
    if (e.second < n) {
        if (e.first) cudaFree(e.first);
        e.first = nullptr; e.second = 0;
        if (cudaMalloc((void**)&e.first, n * sizeof(float)) != cudaSuccess) { h->err = "workspace cudaMalloc failed: " + name; return nullptr; }
        e.second = n;
    }
    return e.first;
}

bool voc_tc_try(lqt_engine* h, const ConvGemmParams& c);

void launch_conv_gemm(lqt_engine* h, ConvGemmParams p) {
    if (p.bias_mod <= 0) p.bias_mod = p.N;
    if (p.dil <= 0) p.dil = 1;
    if (p.taps <= 0) p.taps = 1;
    if (voc_tc_try(h, p)) return;                 // TMA-fed tcgen05 implicit GEMM (tc_conv.cuh) where the shape allows
    static const bool no_mma = getenv("LQT_CONV_FP32") != nullptr;             // A/B aid: the CUDA-core kernel
    if (!no_mma && p.Cin % 16 == 0) {                                       // tensor-core path (bf16x3 split activations, exact products)
        dim3 grid((p.L + CM_BM - 1) / CM_BM, (p.N + CM_BN - 1) / CM_BN);
        conv_gemm_mma_kernel<<<grid, CM_THREADS, 0, h->stream>>>(p);
    } else {
        dim3 grid((p.L + CG_BM - 1) / CG_BM, (p.N + CG_BN - 1) / CG_BN);
        conv_gemm_kernel<<<grid, CG_THREADS, 0, h->stream>>>(p);
    }
    h->stats.kernel_launches++;
}

ConvGemmParams cg(const float* x, int L, int Cin, const bf16* W, int N, float* y) {
    ConvGemmParams p{};
    p.x = x; p.L = L; p.Cin = Cin; p.W = W; p.N = N; p.y = y; p.taps = 1; p.dil = 1; p.bias_mod = N;
    return p;
}

void launch_snake(lqt_engine* h, const float* x, float* y, long long n, int C, const float* a, const float* b) {
    const long long n4 = n / 4;
    const int blocks = (int)std::min<long long>((n4 + 255) / 256, (long long)h->num_sms * 16);
    snake_kernel<<<std::max(blocks, 1), 256, 0, h->stream>>>(x, y, n4, C, a, b);
    h->stats.kernel_launches++;
}

// text_project.onnx on the device: ids_dev [S] -> tp_out [S][H]
int run_text_project(lqt_engine* h, const long long* ids_dev, int S, float* out) {
    const int Dt = h->sp.text_dim, H = h->sp.hidden;
    float* emb = wsbuf(h, "tp_emb", (size_t)S * Dt);
    float* hid = wsbuf(h, "tp_hid", (size_t)S * Dt);
    if (!emb || !hid) return 1;
    gather_rows_kernel<<<S, 256, 0, h->stream>>>(h->text_embed, ids_dev, Dt, emb);
    h->stats.kernel_launches++;
    GemvParams p1 = gemv_params(h->fc1w, Dt, Dt, emb, hid, S);
    p1.bias = h->fc1b; p1.act = 1;
    launch_gemv(h, p1, false);
    GemvParams p2 = gemv_params(h->fc2w, H, Dt, hid, out, S);
    p2.bias = h->fc2b;
    launch_gemv(h, p2, false);
    CK(cudaGetLastError());
    return 0;
}

// build_prompt_embeddings (src/tts_onnx.cpp:442-539) -> h->prompt_dev [P][H], h->trailing_dev, h->tts_pad_dev
int build_prompt_device(lqt_engine* h, const int64_t* ids, int n, int lang_id, const float* spk_host,
                        int* P_out, int* trailing_len_out) {
    if (n < 5) { h->err = "token_ids must hold at least 5 ids (role x3, >=1 text, 2 trailer)"; return 1; }
    // lang_id indexes codec_embed directly (:470-473): 0 = auto, else one of the language rows of the codec vocabulary
    if (lang_id != 0 && (lang_id < 2048 || lang_id >= h->sp.vocab)) { h->err = "lang_codec_id out of range (0 = auto, else a codec id in [2048, vocab))"; return 1; }
    const int H = h->sp.hidden;
    const int n_text_rest = std::max(0, n - 6);              // ids[4 .. n-3]
    if (n_text_rest + 1 > h->sp.max_pos) { h->err = "text too long"; return 1; }
    std::vector<long long> all = {TTS_BOS, TTS_EOS, TTS_PAD, ids[0], ids[1], ids[2], ids[3]};
    for (int i = 4; i < n - 2; ++i) all.push_back(ids[i]);
    for (long long v : all)
        if (v < 0 || v >= h->sp.text_vocab) { h->err = "text token id out of range"; return 1; }
    const int S = (int)all.size();
    long long* ids_dev = (long long*)wsbuf(h, "tp_ids", (size_t)S * 2);
    float* tp = wsbuf(h, "tp_out", (size_t)S * H);
    if (!ids_dev || !tp) return 1;
    CK(cudaMemcpyAsync(ids_dev, all.data(), S * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    if (run_text_project(h, ids_dev, S, tp)) return 1;
    if (spk_host) CK(cudaMemcpyAsync(h->spk_dev, spk_host, H * sizeof(float), cudaMemcpyHostToDevice, h->stream));

    std::vector<int> prefill;                                                   // :466-476
    if (lang_id == 0) prefill = {CODEC_NOTHINK, CODEC_THINK_BOS, CODEC_THINK_EOS};
    else prefill = {CODEC_THINK, CODEC_THINK_BOS, lang_id, CODEC_THINK_EOS};
    prefill.push_back(CODEC_PAD); prefill.push_back(CODEC_BOS);
    std::vector<std::pair<int, int>> cod;                                       // (codec id, is_speaker)
    for (size_t i = 0; i + 1 < prefill.size(); ++i) cod.push_back({prefill[i], 0});
    if (spk_host) cod.push_back({-1, 1});                                       // :481-490
    cod.push_back({prefill.back(), 0});
    const int pad_count = (int)prefill.size() - 2 + (spk_host ? 1 : 0);         // :497-498
    std::vector<int> desc;
    auto row = [&](int tpi, int ci, int sf) { desc.push_back(tpi); desc.push_back(ci); desc.push_back(sf); };
    row(3, -1, 0); row(4, -1, 0); row(5, -1, 0);                                // role (:493-494)
    for (int i = 0; i < pad_count; ++i) row(2, cod[i].first, cod[i].second);    // tts_pad + codec (:499-512)
    row(0, cod[pad_count].first, cod[pad_count].second);                        // tts_bos + codec/speaker
    row(6, cod[pad_count + 1].first, cod[pad_count + 1].second);                // first text + codec_bos (:515-520)
    const int P = (int)desc.size() / 3;
    for (int i = 0; i < n_text_rest; ++i) row(7 + i, -1, 0);                    // trailing (:530-534)
    row(1, -1, 0);                                                              // tts_eos (:535)
    const int TL = n_text_rest + 1;
    row(2, -1, 0);                                                              // tts_pad_embed_ (:463)
    const int rows = (int)desc.size() / 3;
    int* desc_dev = (int*)wsbuf(h, "tp_desc", desc.size());
    float* asm_out = wsbuf(h, "tp_asm", (size_t)rows * H);
    if (!desc_dev || !asm_out) return 1;
    CK(cudaMemcpyAsync(desc_dev, desc.data(), desc.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    assemble_rows_kernel<<<rows, 256, 0, h->stream>>>(desc_dev, tp, h->codec_embed, h->spk_dev, H, asm_out);
    h->stats.kernel_launches++;
    CK(cudaMemcpyAsync(h->prompt_dev, asm_out, (size_t)P * H * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(h->trailing_dev, asm_out + (size_t)P * H, (size_t)TL * H * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(h->tts_pad_dev, asm_out + (size_t)(P + TL) * H, (size_t)H * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));      // desc/all vectors go out of scope
    *P_out = P; *trailing_len_out = TL;
    return 0;
}

#include "tma_host.inl"
#include "voc_tc.inl"

// tokenizer12hz_decode (src/tts_onnx.cpp:759-776) on device codes [T][16] -> audio_dev [T*spf]
// STREAMING (vs != nullptr): codes are frames [vs->t0, vs->t0 + T) of a longer utterance. Every op with left context (pre_conv, the
// window attention's K/V, the depthwise convs, every k > 1 convolution of the decoder, the final conv) finds the previous chunk's
// tail in the rows in FRONT of its input (copied in from the per-site state before it runs, saved again afterwards), so a chunk
// computes exactly what the one-shot decode computes for those frames -- bit for bit (same per-element arithmetic, same K order).
int voc_halo(lqt_engine* h, lqt_engine::VocStream* vs, const std::string& key, void* row0, size_t row_bytes, int halo_rows, long long n_rows) {
    if (!vs || halo_rows <= 0) return 0;
    auto& st = vs->halo[key];
    const size_t bytes = (size_t)halo_rows * row_bytes;
    if (st.second != bytes) {
        if (st.first) cudaFree(st.first);
        st.first = nullptr; st.second = 0;
        CK(cudaMalloc(&st.first, bytes));
        CK(cudaMemsetAsync(st.first, 0, bytes, h->stream));          // before the first frame: zeros = the causal padding
        st.second = bytes;
    }
    char* r0 = reinterpret_cast<char*>(row0);
    CK(cudaMemcpyAsync(r0 - bytes, st.first, bytes, cudaMemcpyDeviceToDevice, h->stream));                       // previous tail in front of row 0
    CK(cudaMemcpyAsync(st.first, r0 + (n_rows - halo_rows) * (long long)row_bytes, bytes, cudaMemcpyDeviceToDevice, h->stream));   // new tail
    return 0;
}
void voc_stream_clear(lqt_engine* h) {
    for (auto& e : h->voc_stream.halo) if (e.second.first) cudaFree(e.second.first);
    h->voc_stream.halo.clear();
    h->voc_stream.t0 = 0;
}
// new utterance: zero the carried tails (buffers stay allocated: no cudaMalloc/cudaFree while a frame kernel may be running)
int voc_stream_reset(lqt_engine* h, cudaStream_t st) {
    for (auto& e : h->voc_stream.halo) if (e.second.first) CK(cudaMemsetAsync(e.second.first, 0, e.second.second, st));
    h->voc_stream.t0 = 0;
    return 0;
}

int run_vocoder(lqt_engine* h, const long long* codes_dev, int T, float* audio, cudaStream_t vstream = nullptr, lqt_engine::VocStream* vs = nullptr) {
    // every launch below goes to h->stream: a caller that wants another stream (the first-audio chunk) passes it here and the
    // handle's stream is swapped for the duration of the enqueue only (one host thread per handle, include/lqt_b200.h)
    struct StreamSwap { lqt_engine* h; cudaStream_t keep; StreamSwap(lqt_engine* e, cudaStream_t v) : h(e), keep(e->stream) { if (v) e->stream = v; } ~StreamSwap() { h->stream = keep; } } swap_(h, vstream);
    const Spec& s = h->sp;
    const int Dc = s.voc_codebook_dim, R = s.voc_rvq_out, Cv = s.voc_hidden, I = s.voc_inter;
    const int vqd = s.voc_heads * s.voc_head_dim;
    const int t0 = vs ? vs->t0 : 0;
    if (t0 + T > s.voc_max_pos) { h->err = "vocoder: too many frames"; return 1; }
    // front margin of every workspace buffer: room for the largest left context (the attention window) of the widest row
    const size_t margin = (size_t)std::max(s.voc_window, 64) * std::max(std::max(3 * vqd, I), std::max(4 * Cv, s.voc_decoder_dim));
    // largest activation (elements) over all stages
    size_t maxel = (size_t)T * std::max(std::max(3 * vqd, I), std::max(Cv, R));
    {
        size_t L = T;
        for (int f : s.voc_up_ratios) { L *= f; maxel = std::max(maxel, L * (size_t)(4 * Cv)); }
        maxel = std::max(maxel, L * (size_t)s.voc_decoder_dim);
        int c = s.voc_decoder_dim;
        for (int r : s.voc_up_rates) { L *= r; c /= 2; maxel = std::max(maxel, L * (size_t)c); }
    }
    float* b0 = wsbuf(h, "v0", maxel + margin); float* b1 = wsbuf(h, "v1", maxel + margin);
    float* b2 = wsbuf(h, "v2", maxel + margin); float* b3 = wsbuf(h, "v3", maxel + margin);
    if (!b0 || !b1 || !b2 || !b3) return 1;
    b0 += margin; b1 += margin; b2 += margin; b3 += margin;

    // RVQ gather-sum + 1x1 output projections
    float *sem = b1, *aco = b2;
    rvq_gather_kernel<<<T, 256, 0, h->stream>>>(codes_dev, T, N_CODEBOOKS, h->rvq_sem_cb, h->rvq_aco_cb,
                                                s.voc_codebook_size, Dc, sem, aco);
    h->stats.kernel_launches++;
    launch_conv_gemm(h, cg(sem, T, Dc, h->rvq_sem_proj, R, b0));
    { ConvGemmParams p = cg(aco, T, Dc, h->rvq_aco_proj, R, b0); p.residual = b0; launch_conv_gemm(h, p); }
    // pre_conv k3
    float* xa = b3;
    if (voc_halo(h, vs, "pre_conv", b0, (size_t)R * 4, 2, T)) return 1;
    { ConvGemmParams p = cg(b0, T, R, h->pre_conv_w, Cv, xa); p.taps = 3; p.bias = h->pre_conv_b; p.hist = vs ? 2 : 0; launch_conv_gemm(h, p); }
    // pre-transformer (sliding-window attention, LayerScale)
    for (int l = 0; l < s.voc_layers; ++l) {
        const LayerW& L = h->vl[l];
        rmsnorm_rows_kernel<<<(T + 7) / 8, 256, 0, h->stream>>>(xa, b0, T, Cv, L.ln1, s.voc_rms_eps);
        h->stats.kernel_launches++;
        launch_conv_gemm(h, cg(b0, T, Cv, L.wqkv, 3 * vqd, b1));
        WinAttnParams w{};
        w.qkv = b1; w.out = b2; w.rope_cos = h->v_cos; w.rope_sin = h->v_sin; w.T = T; w.n_heads = s.voc_heads;
        w.window = s.voc_window; w.scale = 1.0f / sqrtf((float)s.voc_head_dim);
        w.pos0 = t0; w.hist = std::min(t0, s.voc_window - 1);
        if (voc_halo(h, vs, "att" + std::to_string(l), b1, (size_t)3 * vqd * 4, s.voc_window - 1, T)) return 1;
        window_attn_kernel<<<(int)(((long long)T * s.voc_heads + 7) / 8), 256, 0, h->stream>>>(w);
        h->stats.kernel_launches++;
        { ConvGemmParams p = cg(b2, T, vqd, L.wo, Cv, xa); p.scale = L.ls1; p.residual = xa; launch_conv_gemm(h, p); }
        rmsnorm_rows_kernel<<<(T + 7) / 8, 256, 0, h->stream>>>(xa, b0, T, Cv, L.ln2, s.voc_rms_eps);
        h->stats.kernel_launches++;
        launch_conv_gemm(h, cg(b0, T, Cv, L.wgate, I, b1));
        launch_conv_gemm(h, cg(b0, T, Cv, L.wup, I, b2));
        silu_mul_kernel<<<std::min((int)(((size_t)T * I + 255) / 256), h->num_sms * 16), 256, 0, h->stream>>>(b1, b2, b1, (long long)T * I);
        h->stats.kernel_launches++;
        { ConvGemmParams p = cg(b1, T, I, L.wdown, Cv, xa); p.scale = L.ls2; p.residual = xa; launch_conv_gemm(h, p); }
    }
    rmsnorm_rows_kernel<<<(T + 7) / 8, 256, 0, h->stream>>>(xa, b0, T, Cv, h->v_norm, s.voc_rms_eps);
    h->stats.kernel_launches++;
    // upsample stages: transposed conv (kernel = stride) + ConvNeXt
    float* cur = b0; float* o1 = b1; float* o2 = b2; float* o3 = b3;
    int L = T;
    for (size_t u = 0; u < h->vup.size(); ++u) {
        const VocUpW& U = h->vup[u];
        { ConvGemmParams p = cg(cur, L, Cv, U.tconv_w, U.factor * Cv, o1); p.bias = U.tconv_b; p.bias_mod = Cv; launch_conv_gemm(h, p); }
        L *= U.factor;
        if (voc_halo(h, vs, "dw" + std::to_string(u), o1, (size_t)Cv * 4, 6, L)) return 1;
        dwconv_ln_kernel<<<L, 256, Cv * sizeof(float), h->stream>>>(o1, o2, L, Cv, U.dw_w, U.dw_b, U.ln_w, U.ln_b, 1e-6f, vs ? 6 : 0);
        h->stats.kernel_launches++;
        { ConvGemmParams p = cg(o2, L, Cv, U.pw1_w, 4 * Cv, o3); p.bias = U.pw1_b; p.act = 3; launch_conv_gemm(h, p); }
        { ConvGemmParams p = cg(o3, L, 4 * Cv, U.pw2_w, Cv, cur); p.bias = U.pw2_b; p.scale = U.gamma; p.residual = o1; launch_conv_gemm(h, p); }
    }
    // decoder: TMA-fed tcgen05 implicit-GEMM convolutions with fused SnakeBeta (tc_conv.cuh) ...
    if (h->voc_tc && h->voc_tc->ready) return voc_tc_decoder(h, cur, L, audio, vs);
    // ... or the round-1 kernels (mma.sync implicit GEMM on fp32 activations + separate SnakeBeta passes; LQT_VOC_TC=0)
    if (voc_halo(h, vs, "dec_in", cur, (size_t)Cv * 4, 6, L)) return 1;
    { ConvGemmParams p = cg(cur, L, Cv, h->dec_in_w, s.voc_decoder_dim, o1); p.taps = 7; p.bias = h->dec_in_b; p.hist = vs ? 6 : 0; launch_conv_gemm(h, p); }
    float* t = o1;                       // running activation
    float* fa = cur; float* fb = o2;                      // free buffers (o3 unused from here)
    for (size_t b = 0; b < h->vblk.size(); ++b) {
        const VocBlockW& B = h->vblk[b];
        launch_snake(h, t, fa, (long long)L * B.cin, B.cin, B.snake_a, B.snake_b);
        if (voc_halo(h, vs, "tconv" + std::to_string(b), fa, (size_t)B.cin * 4, 1, L)) return 1;
        { ConvGemmParams p = cg(fa, L, B.cin, B.tconv_w, B.stride * B.cout, fb); p.taps = 2; p.tap_rev = 1; p.bias = B.tconv_b; p.bias_mod = B.cout; p.hist = vs ? 1 : 0; launch_conv_gemm(h, p); }
        L *= B.stride;
        std::swap(t, fb);                // t = tconv output; fb = old t (free)
        const int dil[3] = {1, 3, 9};
        for (int r = 0; r < 3; ++r) {
            const auto& Rr = B.res[r];
            launch_snake(h, t, fa, (long long)L * B.cout, B.cout, Rr.s1a, Rr.s1b);
            if (voc_halo(h, vs, "c1_" + std::to_string(b) + "_" + std::to_string(r), fa, (size_t)B.cout * 4, 6 * dil[r], L)) return 1;
            { ConvGemmParams p = cg(fa, L, B.cout, Rr.c1w, B.cout, fb); p.taps = 7; p.dil = dil[r]; p.bias = Rr.c1b; p.hist = vs ? 6 * dil[r] : 0; launch_conv_gemm(h, p); }
            launch_snake(h, fb, fa, (long long)L * B.cout, B.cout, Rr.s2a, Rr.s2b);
            { ConvGemmParams p = cg(fa, L, B.cout, Rr.c2w, B.cout, t); p.bias = Rr.c2b; p.residual = t; launch_conv_gemm(h, p); }
        }
    }
    const int Cl = h->vblk.empty() ? s.voc_decoder_dim : h->vblk.back().cout;
    launch_snake(h, t, fa, (long long)L * Cl, Cl, h->out_sa, h->out_sb);
    if (voc_halo(h, vs, "conv_out", fa, (size_t)Cl * 4, 6, L)) return 1;
    conv_out_kernel<<<(L + 7) / 8, 256, 0, h->stream>>>(fa, audio, L, Cl, h->out_w, h->out_b, vs ? 6 : 0);
    h->stats.kernel_launches++;
    CK(cudaGetLastError());
    return 0;
}

int ensure_audio(lqt_engine* h, int T) {
    const size_t need_codes = (size_t)T * N_CODEBOOKS, need_audio = (size_t)T * h->sp.samples_per_frame;
    if (h->voc_codes_cap < need_codes) {
        if (h->voc_codes_dev) cudaFree(h->voc_codes_dev);
        CK(cudaMalloc((void**)&h->voc_codes_dev, need_codes * sizeof(long long)));
        h->voc_codes_cap = need_codes;
    }
    if (h->audio_cap < need_audio) {
        if (h->audio_dev) cudaFree(h->audio_dev);
        CK(cudaMalloc((void**)&h->audio_dev, need_audio * sizeof(float)));
        h->audio_cap = need_audio;
    }
    return 0;
}

int upload_sampling(lqt_engine* h, const lqt_sampling* sp) {
    SamplingDev d{sp->temperature, sp->top_p, sp->top_k, sp->greedy, sp->seed, sp->utterance_id};
    CK(cudaMemcpyAsync(h->sampling_dev, &d, sizeof(d), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int fk_launch(lqt_engine* h, int slot, int mode, const float* prompt, int P, int frame_end, bool trace);
int fk_check_abort(lqt_engine* h);

// v1 frame loop (kept for A/B measurements): prefill as P decode-shaped steps + one CUDA graph of
// per-op kernels per frame. GenState already uploaded.
int generate_core_graph(lqt_engine* h, int slot, int P, const lqt_sampling* sp, bool trace) {
    cudaGraphExec_t graph;
    if (get_frame_graph(h, slot, trace, &graph)) return 1;
    CK(cudaEventRecord(h->ev0, h->stream));
    // prefill (src/tts_onnx.cpp:794): P decode-shaped steps, head only on the last row
    const int H = h->sp.hidden;
    for (int i = 0; i < P; ++i) {
        run_talker_token(h, slot, h->prompt_dev + (size_t)i * H, i == P - 1, nullptr);
        advance_kernel<<<1, 1, 0, h->stream>>>(h->st);
        h->stats.kernel_launches++;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    cudaEvent_t evg0, evg1, evpoll;
    CK(cudaEventCreate(&evg0)); CK(cudaEventCreate(&evg1)); CK(cudaEventCreateWithFlags(&evpoll, cudaEventDisableTiming));
    CK(cudaEventRecord(evg0, h->stream));
    bool poll_pending = false;
    int launched = 0;
    for (int f = 0; f < sp->max_new_tokens; ++f) {
        CK(cudaGraphLaunch(graph, h->stream));
        h->stats.graph_launches++;
        h->stats.kernel_launches += h->kernels_per_frame;
        ++launched;
        if ((f & 7) == 7) {
            if (poll_pending && cudaEventQuery(evpoll) == cudaSuccess) {
                poll_pending = false;
                if (h->st_host->done) break;
            }
            if (!poll_pending) {
                CK(cudaMemcpyAsync(h->st_host, h->st, sizeof(GenState), cudaMemcpyDeviceToHost, h->stream));
                CK(cudaEventRecord(evpoll, h->stream));
                poll_pending = true;
            }
        }
    }
    CK(cudaEventRecord(evg1, h->stream));
    CK(cudaMemcpyAsync(h->st_host, h->st, sizeof(GenState), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    cudaEventElapsedTime(&h->stats.last_prefill_ms, h->ev0, h->ev1);
    cudaEventElapsedTime(&h->stats.last_generate_ms, evg0, evg1);
    cudaEventDestroy(evg0); cudaEventDestroy(evg1); cudaEventDestroy(evpoll);
    return 0;
}

// loops A+B on the device. prompt_dev / trailing_dev / tts_pad_dev / forced_dev already filled.
int generate_core(lqt_engine* h, int slot, int P, int trailing_len, const lqt_sampling* sp, int n_forced,
                  bool trace, int* n_frames_out) {
    if (slot < 0 || slot >= h->n_slots) { h->err = "bad slot"; return 1; }
    if (P < 1 || sp->max_new_tokens < 0 || sp->max_new_tokens > h->max_frames_cap ||
        P + sp->max_new_tokens > h->sp.max_pos) { h->err = "P + max_new_tokens exceeds max_pos"; return 1; }
    if (upload_sampling(h, sp)) return 1;
    GenState g{};
    g.pos = 0; g.frame = 0; g.done = 0; g.n_frames = 0; g.trailing_len = trailing_len;
    g.max_frames = sp->max_new_tokens; g.cp_pos = 0; g.n_forced = n_forced;
    *h->st_host = g;
    CK(cudaMemcpyAsync(h->st, h->st_host, sizeof(GenState), cudaMemcpyHostToDevice, h->stream));
    if (h->frame_impl == 0) {
        // persistent frame kernel: prefill + all frames in one cluster launch (frame_kernel.cuh). The kernel can also RESUME an
        // utterance (from GenState and the plain logits / last_hidden copies of the previous launch): $LQT_FK_SPLIT=n splits the
        // launch after n frames, which is how the tests exercise that path.
        CK(cudaEventRecord(h->ev0, h->stream));
        static const int split_env = getenv("LQT_FK_SPLIT") ? atoi(getenv("LQT_FK_SPLIT")) : 0;
        const int split = (split_env > 0 && split_env < sp->max_new_tokens) ? split_env : 0;
        if (split) {
            if (fk_launch(h, slot, 0, h->prompt_dev, P, split, trace)) return 1;
            CK(cudaMemcpyAsync(h->st_host, h->st, sizeof(GenState), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            if (fk_check_abort(h)) return 1;
            if (!h->st_host->done && h->st_host->n_frames == split)
                if (fk_launch(h, slot, 0, h->prompt_dev, P, sp->max_new_tokens, trace)) return 1;       // resumes: pos != 0
        } else {
            if (fk_launch(h, slot, 0, h->prompt_dev, P, sp->max_new_tokens, trace)) return 1;
        }
        CK(cudaEventRecord(h->ev1, h->stream));
        CK(cudaMemcpyAsync(h->st_host, h->st, sizeof(GenState), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
        if (fk_check_abort(h)) return 1;
        cudaEventElapsedTime(&h->stats.last_generate_ms, h->ev0, h->ev1);
        h->stats.last_prefill_ms = 0.f;             // included in last_generate_ms
    } else {
        if (generate_core_graph(h, slot, P, sp, trace)) return 1;
    }
    h->slot_len[slot] = h->st_host->pos;
    *n_frames_out = h->st_host->n_frames;
    h->stats.last_frames = h->st_host->n_frames;
    return 0;
}

int ensure_trace(lqt_engine* h, int frames, int stride) {
    const size_t need_n = (size_t)std::max(frames, 1) * N_CODEBOOKS * stride;
    auto& e = h->ws["trace"];
    if (e.second < need_n || h->trace_stride != stride || h->trace_dev != e.first) {
        for (auto it = h->graphs.begin(); it != h->graphs.end();) {
            if (it->first & 1) { cudaGraphExecDestroy(it->second); it = h->graphs.erase(it); } else ++it;
        }
        float* p = wsbuf(h, "trace", need_n);
        if (!p) return 1;
        h->trace_dev = p; h->trace_stride = stride;
    }
    CK(cudaMemsetAsync(h->trace_dev, 0, need_n * sizeof(float), h->stream));
    return 0;
}

bool load_vocoder_weights(lqt_engine* h) {
    bool ok = true;
    const LqwFile& f = h->f_voc;
    const Spec& s = h->sp;
    const int Cv = s.voc_hidden, Dc = s.voc_codebook_dim, R = s.voc_rvq_out, vqd = s.voc_heads * s.voc_head_dim;
    h->rvq_sem_cb = need<bf16>(h, f, "rvq.sem.codebook", ok, {1, s.voc_codebook_size, Dc});  h->rvq_sem_proj = need<bf16>(h, f, "rvq.sem.out_proj", ok, {R, Dc});
    h->rvq_aco_cb = need<bf16>(h, f, "rvq.aco.codebook", ok, {s.cp_steps, s.voc_codebook_size, Dc});  h->rvq_aco_proj = need<bf16>(h, f, "rvq.aco.out_proj", ok, {R, Dc});
    h->pre_conv_w = need<bf16>(h, f, "pre_conv.weight", ok, {Cv, 3, R});   h->pre_conv_b = need<float>(h, f, "pre_conv.bias", ok, {Cv});
    ok = load_layers(h, f, "pt.", s.voc_layers, Cv, vqd, vqd, s.voc_head_dim, s.voc_inter, false, true, h->vl) && ok;
    h->v_norm = need<float>(h, f, "pt.norm", ok, {Cv});
    h->v_cos = need<float>(h, f, "pt.rope_cos", ok, {s.voc_max_pos, s.voc_head_dim / 2}); h->v_sin = need<float>(h, f, "pt.rope_sin", ok, {s.voc_max_pos, s.voc_head_dim / 2});
    for (size_t u = 0; u < s.voc_up_ratios.size(); ++u) {
        const std::string p = "up" + std::to_string(u) + ".";
        VocUpW U{};
        U.factor = s.voc_up_ratios[u];
        U.tconv_w = need<bf16>(h, f, p + "tconv.weight", ok, {U.factor, Cv, 1, Cv}); U.tconv_b = need<float>(h, f, p + "tconv.bias", ok, {Cv});
        U.dw_w = need<float>(h, f, p + "dw.weight", ok, {7, Cv}); U.dw_b = need<float>(h, f, p + "dw.bias", ok, {Cv});
        U.ln_w = need<float>(h, f, p + "ln.weight", ok, {Cv}); U.ln_b = need<float>(h, f, p + "ln.bias", ok, {Cv});
        U.pw1_w = need<bf16>(h, f, p + "pw1.weight", ok, {4 * Cv, Cv}); U.pw1_b = need<float>(h, f, p + "pw1.bias", ok, {4 * Cv});
        U.pw2_w = need<bf16>(h, f, p + "pw2.weight", ok, {Cv, 4 * Cv}); U.pw2_b = need<float>(h, f, p + "pw2.bias", ok, {Cv});
        U.gamma = need<float>(h, f, p + "gamma", ok, {Cv});
        h->vup.push_back(U);
    }
    h->dec_in_w = need<bf16>(h, f, "dec.conv_in.weight", ok, {s.voc_decoder_dim, 7, Cv}); h->dec_in_b = need<float>(h, f, "dec.conv_in.bias", ok, {s.voc_decoder_dim});
    int cin = s.voc_decoder_dim;
    for (size_t b = 0; b < s.voc_up_rates.size(); ++b) {
        const std::string p = "dec.b" + std::to_string(b) + ".";
        VocBlockW B{};
        B.cin = cin; B.cout = cin / 2; B.stride = s.voc_up_rates[b];
        B.snake_a = need<float>(h, f, p + "snake.alpha", ok, {B.cin}); B.snake_b = need<float>(h, f, p + "snake.beta", ok, {B.cin});
        B.tconv_w = need<bf16>(h, f, p + "tconv.weight", ok, {B.stride, B.cout, 2, B.cin}); B.tconv_b = need<float>(h, f, p + "tconv.bias", ok, {B.cout});
        for (int r = 0; r < 3; ++r) {
            const std::string q = p + "r" + std::to_string(r) + ".";
            auto& R = B.res[r];
            R.s1a = need<float>(h, f, q + "snake1.alpha", ok, {B.cout}); R.s1b = need<float>(h, f, q + "snake1.beta", ok, {B.cout});
            R.c1w = need<bf16>(h, f, q + "conv1.weight", ok, {B.cout, 7, B.cout});  R.c1b = need<float>(h, f, q + "conv1.bias", ok, {B.cout});
            R.s2a = need<float>(h, f, q + "snake2.alpha", ok, {B.cout}); R.s2b = need<float>(h, f, q + "snake2.beta", ok, {B.cout});
            R.c2w = need<bf16>(h, f, q + "conv2.weight", ok, {B.cout, 1, B.cout});  R.c2b = need<float>(h, f, q + "conv2.bias", ok, {B.cout});
        }
        h->vblk.push_back(B);
        cin /= 2;
    }
    h->out_sa = need<float>(h, f, "dec.snake_out.alpha", ok, {cin}); h->out_sb = need<float>(h, f, "dec.snake_out.beta", ok, {cin});
    h->out_w = need<float>(h, f, "dec.conv_out.weight", ok, {1, 7, cin});  h->out_b = need<float>(h, f, "dec.conv_out.bias", ok, {1});
    return ok;
}

// ------------------------------------------------------------------------------------------------
// persistent frame kernel: one-time weight regrouping, tables, launch
// ------------------------------------------------------------------------------------------------
// Per-CTA weight images (see frame_kernel.cuh): CTA c's rows of the matrix.
//  mode 0 (flat rows) / 1 (gate/up interleaved rows) / 2 (O-projection: rows of a cluster's member index, columns of the CTA's kv group): mma.m16n8k16 A-fragment order (gemv_mma): rows in tiles of 8, two tiles
//  per 16-row operand; for tile pair p and 16-column block kt the 32 lanes' fragments are contiguous, [p][kt][lane][a0 a1 a2 a3]
//  (a0/a2: row g of the first tile, columns 2tg.. and 8 + 2tg..; a1/a3: row g of the second tile); an odd last tile stores
//  [kt][lane][a0 a2]. No padding: image bytes = rows * K * 2.
struct ImgJob {
    const bf16* src0; const bf16* src1;   // mode 1: gate / up
    bf16* dst;
    int N, K, RG, mode;
    int src_stride, n_kv, rmax;
};
__global__ void fk_build_image_kernel(const ImgJob j) {
    const int c = blockIdx.x, ncta = gridDim.x;
    bf16* dst = j.dst + (size_t)c * j.rmax * j.K;
    const FkSlice sl = (j.mode == 2) ? group_slice(j.N, c, ncta, j.n_kv) : flat_slice(j.N, FK_TILE_ROWS, c, ncta);
    const int nkt = j.K >> 4, nt = sl.nrows >> 3, npair = nt >> 1;
    const long long nword = (long long)sl.nrows * j.K / 2;          // 32-bit words (bf16 pairs)
    const long long pair_words = (long long)npair * nkt * 128;
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst);
    for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < nword; i += (long long)gridDim.y * blockDim.x) {
        int row, col;
        if (i < pair_words) {
            const int p = (int)(i / (nkt * 128)), rem = (int)(i % (nkt * 128));
            const int kt = rem >> 7, w = rem & 127, lane = w >> 2, reg = w & 3, g = lane >> 2, tg = lane & 3;
            row = (2 * p + (reg & 1)) * 8 + g; col = 16 * kt + 2 * tg + (reg >> 1) * 8;
        } else {
            const int rem = (int)(i - pair_words);
            const int kt = rem >> 6, w = rem & 63, lane = w >> 1, reg = w & 1, g = lane >> 2, tg = lane & 3;
            row = (nt - 1) * 8 + g; col = 16 * kt + 2 * tg + reg * 8;
        }
        const int n = sl.row0 + row;
        const bf16* src = (j.mode == 1) ? ((n & 1) ? j.src1 : j.src0) + (size_t)(n >> 1) * j.src_stride + col
                        : (j.mode == 2) ? j.src0 + (size_t)n * j.src_stride + (size_t)(c % j.n_kv) * j.K + col
                                        : j.src0 + (size_t)n * j.src_stride + col;
        dw[i] = *reinterpret_cast<const uint32_t*>(src);
    }
}

template <typename T>
int fk_alloc(lqt_engine* h, T** p, size_t n) {
    CK(cudaMalloc((void**)p, n * sizeof(T)));
    CK(cudaMemset(*p, 0, n * sizeof(T)));
    h->fk_allocs.push_back((void*)*p);
    return 0;
}

int fk_rmax(int N, int RG, int mode, int n_kv, int ncta) {        // must match make_desc() in frame_kernel.cuh
    if (mode == 2) { const int ns = ncta / n_kv; return ((N / FK_TILE_ROWS + ns - 1) / ns) * FK_TILE_ROWS; }
    (void)RG;
    return ((N / FK_TILE_ROWS + ncta - 1) / ncta) * FK_TILE_ROWS;           // rows are dealt in tiles of 8
}

// builds one image; returns its device pointer (nullptr on failure)
bf16* fk_image(lqt_engine* h, const bf16* src0, const bf16* src1, int N, int K, int RG, int mode, int src_stride, int n_kv,
               size_t* elems_out = nullptr) {
    ImgJob j{};
    j.src0 = src0; j.src1 = src1; j.N = N; j.K = K; j.RG = RG; j.mode = mode; j.src_stride = src_stride; j.n_kv = n_kv;
    j.rmax = fk_rmax(N, RG, mode, n_kv, h->fk_ncta);
    const size_t elems = (size_t)h->fk_ncta * j.rmax * K + 64;
    bf16* dst = nullptr;
    if (fk_alloc(h, &dst, elems)) return nullptr;
    j.dst = dst;
    fk_build_image_kernel<<<dim3(h->fk_ncta, 4), 256, 0, h->stream>>>(j);
    if (elems_out) *elems_out = elems;
    return dst;
}

int fk_build_stack(lqt_engine* h, const std::vector<LayerW>& layers, int H, int heads, int kv_heads, int inter,
                   const float* cosr, const float* sinr, const float* final_norm, FkStack* out, std::vector<FkLayer>* tab_out) {
    const int gK = (heads / kv_heads) * ATT_D, qkv_dim = (heads + 2 * kv_heads) * ATT_D;
    std::vector<FkLayer> tab(layers.size());
    for (size_t l = 0; l < layers.size(); ++l) {
        const bf16* wqkv = fk_image(h, layers[l].wqkv, nullptr, qkv_dim, H, 1, 0, H, kv_heads);
        const bf16* wo = fk_image(h, layers[l].wo, nullptr, H, gK, 1, 2, kv_heads * gK, kv_heads);
        const bf16* wgu = fk_image(h, layers[l].wgate, layers[l].wup, 2 * inter, H, 2, 1, H, kv_heads);
        const bf16* wdown = fk_image(h, layers[l].wdown, nullptr, H, inter, 1, 0, inter, kv_heads);
        if (!wqkv || !wo || !wgu || !wdown) return 1;
        tab[l] = FkLayer{wqkv, wo, wgu, wdown, layers[l].ln1, layers[l].ln2, layers[l].qnorm, layers[l].knorm};
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    *tab_out = tab;
    FkStack S{};
    S.n_layers = (int)layers.size(); S.H = H; S.heads = heads; S.kv_heads = kv_heads; S.inter = inter;
    S.cos = cosr; S.sin = sinr; S.final_norm = final_norm;
    *out = S;
    return 0;
}

int fk_init(lqt_engine* h) {
    const Spec& s = h->sp;
    const int maxK = std::max(std::max(s.hidden, s.inter), std::max(s.cp_hidden, s.cp_inter));
    auto chk = [&](int K, const char* what) -> bool {
        if (K % 256 != 0 || K > 6144) {                 // gemv_mma: 16-column blocks, an even number per warp
            h->err = std::string("frame kernel: unsupported dimension for ") + what; return false; }
        return true;
    };
    if (!chk(s.hidden, "hidden") || !chk(s.inter, "inter") || !chk(s.cp_hidden, "cp_hidden") || !chk(s.cp_inter, "cp_inter"))
        return 1;
    h->fk_wide = maxK > 3072;                       // frame_kernel<6, 3> instead of <3, 8>
    const int maxV = std::max(s.vocab, s.cp_vocab);
    if (s.hidden > 2048 || s.cp_hidden > 2048 || (maxV % 16) || (s.vocab % 16) || (s.cp_vocab % 16) || maxV > FK_LAND_WORDS || (s.hidden % 16) || (s.cp_hidden % 16) || (s.inter % 16) || (s.cp_inter % 16)) {
        h->err = "frame kernel: hidden > 2048, vocab > 3072 or a dimension that is not a multiple of 16"; return 1;
    }
    {   // shared-memory carve-up and the number of co-resident clusters (the grid)
        const FkSmemLayout L = fk_smem_layout(h->fk_wide ? 3 : 4, maxV, s.hidden, maxK);
        h->fk_so.land = (unsigned)L.land; h->fk_so.scratch = (unsigned)L.scratch; h->fk_so.att = (unsigned)L.att; h->fk_so.xs = (unsigned)L.xs;
        h->fk_so.red = (unsigned)L.red; h->fk_so.nxt = (unsigned)L.nxt; h->fk_so.res0 = (unsigned)L.res0; h->fk_so.lh = (unsigned)L.lh;
        h->fk_so.shared = (unsigned)L.shared; h->fk_so.maxV = maxV;
        h->fk_smem = L.total;
        const void* fn = h->fk_wide ? (const void*)frame_kernel<6, 3> : (const void*)frame_kernel<3, 4>;
        // The attribute belongs to the FUNCTION (process-wide), not to this handle: opt in to the device maximum so that another
        // engine with a smaller model (smaller carve-up) can never lower it under this one's launches.
        int optin = 0;
        CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
        if ((size_t)optin < h->fk_smem) { h->err = "frame kernel: shared memory carve-up exceeds the device limit"; return 1; }
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(FK_CLUSTER * 64); cfg.blockDim = dim3(FK_THREADS); cfg.dynamicSmemBytes = h->fk_smem; cfg.stream = h->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = FK_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = 0;
        CK(cudaOccupancyMaxActiveClusters(&ncl, fn, &cfg));
        if (ncl < 1) { h->err = "frame kernel: no cluster of 8 CTAs fits on this device"; return 1; }
        if (getenv("LQT_FK_NOCOOP")) h->fk_coop = false;        // profilers that cannot replay cooperative launches
        h->fk_ncta = FK_CLUSTER * ncl;
        if (const char* e = getenv("LQT_FK_CLUSTERS")) { const int v = atoi(e); if (v >= 1 && v <= ncl) h->fk_ncta = FK_CLUSTER * v; }
        if (getenv("LQT_DEBUG")) fprintf(stderr, "[lqt] frame kernel: %d clusters of %d CTAs co-resident, using %d CTAs, %zu B shared memory\n", ncl, FK_CLUSTER, h->fk_ncta, h->fk_smem);
    }
    {   // row counts per CTA: warp partials [FK_RED_STRIDE] (two rows at once: 48 each), x1own
        const int nc = h->fk_ncta;
        const int worst = std::max(std::max(fk_rmax(2 * s.inter, 2, 1, s.kv_heads, nc), fk_rmax((s.heads + 2 * s.kv_heads) * ATT_D, 1, 0, s.kv_heads, nc)),
                                   fk_rmax(std::max(s.vocab, s.cp_vocab), 1, 0, s.kv_heads, nc));
        const int worst_c = std::max(fk_rmax(2 * s.cp_inter, 2, 1, s.cp_kv_heads, nc), fk_rmax((s.cp_heads + 2 * s.cp_kv_heads) * ATT_D, 1, 0, s.cp_kv_heads, nc));
        if (worst > 64 || worst_c > 64) { h->err = "frame kernel: too many rows per SM"; return 1; }
        if (h->max_pages > FK_PT_MAX) { h->err = "frame kernel: max_pos exceeds the page-table copy in shared memory"; return 1; }
        {   // a slice must fit the ring: at most NST stages in flight per phase
            auto nst_of = [&](int rows, int K) { return (int)(((size_t)rows * K * 2 + FK_STAGE_BYTES - 1) / FK_STAGE_BYTES); };
            const int a = nst_of(fk_rmax((s.heads + 2 * s.kv_heads) * ATT_D, 1, 0, s.kv_heads, nc), s.hidden), dd = nst_of(fk_rmax(2 * s.inter, 2, 1, s.kv_heads, nc), s.hidden);
            const int ee = nst_of(fk_rmax(s.hidden, 1, 0, s.kv_heads, nc), s.inter), hh = nst_of(fk_rmax(std::max(s.vocab, s.cp_vocab), 1, 0, s.kv_heads, nc), s.hidden);
            const int ce = nst_of(fk_rmax(s.cp_hidden, 1, 0, s.cp_kv_heads, nc), s.cp_inter), cd = nst_of(fk_rmax(2 * s.cp_inter, 2, 1, s.cp_kv_heads, nc), s.cp_hidden);
            const int ca = nst_of(fk_rmax((s.cp_heads + 2 * s.cp_kv_heads) * ATT_D, 1, 0, s.cp_kv_heads, nc), s.cp_hidden);
            if (std::max(std::max(std::max(a, dd), std::max(ee, hh)), std::max(std::max(ce, cd), ca)) > (h->fk_wide ? 3 : 4)) { h->err = "frame kernel: a weight slice exceeds the ring"; return 1; }
        }
        if (std::max(fk_rmax(s.hidden, 1, 0, s.kv_heads, nc), fk_rmax(s.cp_hidden, 1, 0, s.cp_kv_heads, nc)) > FK_X1OWN) { h->err = "frame kernel: too many down-projection rows per SM"; return 1; }
        if (FK_CLUSTER % s.kv_heads || FK_CLUSTER % s.cp_kv_heads) { h->err = "frame kernel: kv heads must divide the cluster size (8)"; return 1; }
        const int rpp_t = (fk_rmax(s.hidden, 1, 2, s.kv_heads, nc) + s.kv_heads - 1) / s.kv_heads, rpp_c = (fk_rmax(s.cp_hidden, 1, 2, s.cp_kv_heads, nc) + s.cp_kv_heads - 1) / s.cp_kv_heads;
        if (std::max(rpp_t, rpp_c) > FK_RPP_MAX) { h->err = "frame kernel: too many O-projection rows per SM"; return 1; }
        if (std::max(fk_rmax(s.hidden, 1, 2, s.kv_heads, nc), fk_rmax(s.cp_hidden, 1, 2, s.cp_kv_heads, nc)) > FK_PART_ROWS) { h->err = "frame kernel: too many O-projection rows per SM"; return 1; }
        if ((s.heads / s.kv_heads) * ATT_D > FK_XS_STRIDE || (s.cp_heads / s.cp_kv_heads) * ATT_D > FK_XS_STRIDE || s.heads != 2 * s.kv_heads || s.cp_heads != 2 * s.cp_kv_heads) {
            h->err = "frame kernel: needs 2 query heads per kv head"; return 1;
        }
    }
    if (s.kv_heads > FK_NGRP_MAX || s.cp_kv_heads > FK_NGRP_MAX || s.kv_heads != s.cp_kv_heads) { h->err = "frame kernel: kv head count"; return 1; }
    if (s.cp_steps + 1 > FK_CP_POS / 2) { h->err = "frame kernel: cp_steps"; return 1; }   // predictor positions 0..cp_steps
    if (h->fk_ncta < s.kv_heads) { h->err = "frame kernel: too few SMs"; return 1; }
    if (s.layers > FK_MAX_TLAYERS || s.cp_layers > FK_MAX_CLAYERS) { h->err = "frame kernel: too many layers"; return 1; }
    if (fk_build_stack(h, h->tl, s.hidden, s.heads, s.kv_heads, s.inter, h->t_cos, h->t_sin, h->t_norm, &h->fk_talker, &h->fk_tl)) return 1;
    if (fk_build_stack(h, h->cl, s.cp_hidden, s.cp_heads, s.cp_kv_heads, s.cp_inter, h->c_cos, h->c_sin, h->c_norm, &h->fk_cp, &h->fk_cl)) return 1;
    {   // head / in_proj images
        h->fk_t_head = fk_image(h, h->t_head, nullptr, s.vocab, s.hidden, 1, 0, s.hidden, s.kv_heads);
        if (!h->fk_t_head) return 1;
        const int r8 = fk_rmax(s.cp_vocab, 1, 0, s.cp_kv_heads, h->fk_ncta);
        h->fk_c_head_stride = (long long)h->fk_ncta * r8 * s.cp_hidden;
        bf16* all = nullptr;
        if (fk_alloc(h, &all, (size_t)h->fk_c_head_stride * s.cp_steps + 64)) return 1;
        for (int j = 0; j < s.cp_steps; ++j) {
            ImgJob job{};
            job.src0 = h->c_heads + (size_t)j * s.cp_vocab * s.cp_hidden; job.N = s.cp_vocab; job.K = s.cp_hidden; job.RG = 1; job.mode = 0;
            job.src_stride = s.cp_hidden; job.n_kv = s.cp_kv_heads; job.rmax = r8; job.dst = all + (size_t)j * h->fk_c_head_stride;
            fk_build_image_kernel<<<dim3(h->fk_ncta, 4), 256, 0, h->stream>>>(job);
        }
        h->fk_c_heads = all;
        if (h->c_inproj_w) {
            h->fk_c_inproj = fk_image(h, h->c_inproj_w, nullptr, s.cp_hidden, s.hidden, 1, 0, s.hidden, s.cp_kv_heads);
            if (!h->fk_c_inproj) return 1;
        }
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(h->stream));
    }
    {   // one arena for every LL exchange buffer
        auto al = [](size_t n) { return (n + 31) & ~(size_t)31; };
        const size_t qkv_t = (size_t)(s.heads + 2 * s.kv_heads) * ATT_D, qkv_c = (size_t)(s.cp_heads + 2 * s.cp_kv_heads) * ATT_D;
        size_t off = 0;
        auto take = [&](size_t n) { const size_t o = off; off += al(n); return o; };
        const size_t o_tx = take(2 * (size_t)s.hidden), o_tq = take(2 * qkv_t), o_tp = take(2 * (size_t)s.kv_heads * s.hidden), o_ta = take(2 * (size_t)s.inter);
        const size_t o_cx = take(2 * (size_t)s.cp_hidden), o_cq = take(2 * qkv_c), o_cp = take(2 * (size_t)s.cp_kv_heads * s.cp_hidden), o_ca = take(2 * (size_t)s.cp_inter);
        const size_t o_pa = take((size_t)8 * FK_NS_MAX * 2 * ATT_PSTRIDE + 8 * 4096), o_ci = take(2 * (size_t)s.cp_hidden);
        const size_t o_lg = take((size_t)s.vocab), o_cl = take((size_t)s.cp_vocab);
        if (fk_alloc(h, &h->fk_arena, off)) return 1;
        h->fk_arena_words = off;
        uint2* a = h->fk_arena;
        h->fk_talker.x = a + o_tx; h->fk_talker.qkv = a + o_tq; h->fk_talker.x1 = a + o_tp; h->fk_talker.act = a + o_ta;
        h->fk_cp.x = a + o_cx; h->fk_cp.qkv = a + o_cq; h->fk_cp.x1 = a + o_cp; h->fk_cp.act = a + o_ca;
        h->fk_pa = a + o_pa; h->fk_cxin = a + o_ci; h->fk_logits_ll = a + o_lg; h->fk_clogits_ll = a + o_cl;
    }
    if (fk_alloc(h, &h->fk_cp_kv, (size_t)h->fk_ncta * s.cp_layers * 2 * FK_CP_POS * ATT_D)) return 1;
    if (fk_alloc(h, &h->fk_ctrl, 64)) return 1;      // [1] abort flag, [32] grid arrival counter (own cache line)
    CK(cudaMallocHost((void**)&h->fk_ctrl_host, 64 * sizeof(unsigned)));
    CK(cudaMallocHost((void**)&h->fk_progress_host, sizeof(int)));
    *h->fk_progress_host = 0;
    return 0;
}

// one cooperative launch. mode 0: prefill (if st->pos == 0) + frames < frame_end ; mode 1: one talker token from next_in
int fk_launch(lqt_engine* h, int slot, int mode, const float* prompt, int P, int frame_end, bool trace) {
    const Spec& s = h->sp;
    FkParams p{};
    p.talker = h->fk_talker; p.cp = h->fk_cp;
    { static const int sl = getenv("LQT_FK_SLEEP") ? atoi(getenv("LQT_FK_SLEEP")) : 800; p.producer_sleep_ns = (unsigned)(sl < 0 ? 0 : sl); }
    for (size_t l = 0; l < h->fk_tl.size(); ++l) p.t_layers[l] = h->fk_tl[l];
    for (size_t l = 0; l < h->fk_cl.size(); ++l) p.c_layers[l] = h->fk_cl[l];
    p.t_head = h->fk_t_head; p.vocab = s.vocab;
    p.c_heads = h->fk_c_heads; p.cp_vocab = s.cp_vocab; p.cp_steps = s.cp_steps; p.c_head_stride = h->fk_c_head_stride;
    p.c_inproj_w = h->fk_c_inproj; p.c_inproj_b = h->c_inproj_b; p.cxin = h->fk_cxin;
    p.eps = s.rms_eps;
    p.kv_pool = h->kv_pool; p.page_table = h->page_tables + (size_t)slot * h->max_pages; p.n_pages = h->max_pages; p.page_shift = KV_PAGE_SHIFT;
    p.page_stride = (long long)s.layers * 2 * s.kv_heads * KV_PAGE * ATT_D; p.kv_f32 = h->kv_f32 ? 1 : 0;
    p.pa = h->fk_pa; p.cp_kv = h->fk_cp_kv;
    p.logits_ll = h->fk_logits_ll; p.clogits_ll = h->fk_clogits_ll;
    p.logits = h->logits; p.clogits = h->clogits; p.last_hidden = h->last_hidden;
    p.next_in = h->next_in;
    p.codec_embed = h->codec_embed; p.cp_embed = h->cp_embed;
    p.prompt = prompt; p.P = P;
    p.trailing = h->trailing_dev; p.tts_pad = h->tts_pad_dev;
    p.st = h->st; p.sp = h->sampling_dev;
    p.codes_out = h->codes_dev; p.forced = h->forced_dev;
    p.trace = trace ? h->trace_dev : nullptr; p.trace_stride = h->trace_stride;
    p.ctrl = h->fk_ctrl; p.frame_end = frame_end; p.mode = mode;
    p.progress = h->fk_progress_on ? h->fk_progress_host : nullptr;
    p.dbg = h->fk_dbg; p.dbg_cap = h->fk_dbg_cap; p.dbg_cta = h->fk_dbg_cta;
    CK(cudaMemsetAsync(h->fk_ctrl, 0, 64 * sizeof(unsigned), h->stream));
    CK(cudaMemsetAsync(h->fk_arena, 0, h->fk_arena_words * sizeof(uint2), h->stream));   // sequence numbers restart at 1
    p.so = h->fk_so;
    void* args[] = {(void*)&p};
    {
        const void* fn = h->fk_wide ? (const void*)frame_kernel<6, 3> : (const void*)frame_kernel<3, 4>;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(h->fk_ncta); cfg.blockDim = dim3(FK_THREADS); cfg.dynamicSmemBytes = h->fk_smem; cfg.stream = h->stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = FK_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = h->fk_coop ? 2 : 1;
        cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
        if (e != cudaSuccess && h->fk_coop) {
            // cluster + cooperative rejected. The grid equals the occupancy limit, so all CTAs are co-resident as long as nothing
            // else holds SMs: from here on the chunk vocoder is NOT overlapped on the second stream (fk_coop gates it in
            // generate_core), and the device-side spin limit stays as the last guard.
            (void)cudaGetLastError();
            fprintf(stderr, "[lqt] cooperative cluster launch refused (%s): launching without the attribute, first-audio overlap off\n", cudaGetErrorString(e));
            h->fk_coop = false; cfg.numAttrs = 1;
            e = cudaLaunchKernelExC(&cfg, fn, args);
        }
        CK(e);
    }
    h->stats.kernel_launches++;
    CK(cudaMemcpyAsync(h->fk_ctrl_host, h->fk_ctrl, 64 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

int fk_check_abort(lqt_engine* h) {       // after a stream synchronize
    if (getenv("LQT_DEBUG") && h->fk_ctrl_host[41]) fprintf(stderr, "[lqt] frame kernel: %u multicast fetches (all CTAs), %u stale 4-word groups re-polled\n", h->fk_ctrl_host[41], h->fk_ctrl_host[40]);
    if (h->fk_ctrl_host[1] != 0) { h->err = "frame kernel aborted: a device-side wait timed out"; return 1; }
    return 0;
}

// opt in to large dynamic shared memory once, outside any stream capture
int prime_kernel_attributes(lqt_engine* h) {
    const int big = 160 * 1024;
    CK(cudaFuncSetAttribute(gemv_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CK(cudaFuncSetAttribute(gemv_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CK(cudaFuncSetAttribute(gemv_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CK(cudaFuncSetAttribute(gemv_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CK(cudaFuncSetAttribute(gemv_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CK(cudaFuncSetAttribute(gemv_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CK(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    return 0;
}

int init_engine(lqt_engine* h, const std::string& dir) {
    if (prime_kernel_attributes(h)) return 1;
    struct { const char* name; LqwFile* f; bool required; } files[] = {
        {"text_project", &h->f_text, true}, {"codec_embed", &h->f_codec, true},
        {"code_predictor_embed", &h->f_cpe, true}, {"talker_prefill", &h->f_talker, true},
        {"code_predictor", &h->f_cp, true}, {"tokenizer12hz_decode", &h->f_voc, true},
        {"speaker_encoder", &h->f_spk, false}};
    // talker_decode.lqw must exist too (7-file layout, src/tts_onnx.cpp:91-104) but shares talker_prefill's tensors
    {
        FILE* t = std::fopen((dir + "/talker_decode.lqw").c_str(), "rb");
        if (!t) { h->err = "Failed to load required model files: talker_decode.lqw missing"; return 1; }
        std::fclose(t);
    }
    for (auto& e : files) {
        const std::string path = dir + "/" + e.name + ".lqw";
        FILE* t = std::fopen(path.c_str(), "rb");
        if (!t) {
            if (e.required) { h->err = std::string("Failed to load required model files: ") + e.name + ".lqw missing"; return 1; }
            continue;
        }
        std::fclose(t);
        const std::string er = load_lqw(path, *e.f);
        if (!er.empty()) { h->err = er; return 1; }
        if (e.f == &h->f_spk) h->has_spk = true;
    }
    const LqwFile& m = h->f_talker;
    Spec& s = h->sp;
    s.hidden = m.meta_int("hidden", 1024); s.layers = m.meta_int("layers", 28); s.heads = m.meta_int("heads", 16);
    s.kv_heads = m.meta_int("kv_heads", 8); s.head_dim = m.meta_int("head_dim", 128); s.inter = m.meta_int("inter", 3072);
    s.vocab = m.meta_int("vocab", 3072); s.max_pos = m.meta_int("max_pos", 2304);
    s.cp_hidden = m.meta_int("cp_hidden", 1024); s.cp_layers = m.meta_int("cp_layers", 5); s.cp_heads = m.meta_int("cp_heads", 16);
    s.cp_kv_heads = m.meta_int("cp_kv_heads", 8); s.cp_inter = m.meta_int("cp_inter", 3072); s.cp_vocab = m.meta_int("cp_vocab", 2048);
    s.cp_steps = m.meta_int("cp_steps", 15); s.cp_max_pos = m.meta_int("cp_max_pos", 32);
    s.text_vocab = m.meta_int("text_vocab", 151936); s.text_dim = m.meta_int("text_dim", 2048);
    s.voc_codebook_size = m.meta_int("voc_codebook_size", 2048); s.voc_codebook_dim = m.meta_int("voc_codebook_dim", 256);
    s.voc_rvq_out = m.meta_int("voc_rvq_out", 512); s.voc_hidden = m.meta_int("voc_hidden", 1024);
    s.voc_layers = m.meta_int("voc_layers", 8); s.voc_heads = m.meta_int("voc_heads", 16); s.voc_head_dim = m.meta_int("voc_head_dim", 64);
    s.voc_inter = m.meta_int("voc_inter", 3072); s.voc_window = m.meta_int("voc_window", 72); s.voc_max_pos = m.meta_int("voc_max_pos", 2304);
    s.voc_decoder_dim = m.meta_int("voc_decoder_dim", 1536);
    s.voc_up_ratios = m.meta_ints("voc_upsampling_ratios"); s.voc_up_rates = m.meta_ints("voc_upsample_rates");
    s.spk_mels = m.meta_int("spk_mels", 128); s.spk_channels = m.meta_int("spk_channels", 512); s.spk_layers = m.meta_int("spk_layers", 3);
    s.rms_eps = (float)m.meta_f("rms_eps", 1e-6); s.voc_rms_eps = (float)m.meta_f("voc_rms_eps", 1e-5);
    s.samples_per_frame = 1;
    for (int v : s.voc_up_ratios) s.samples_per_frame *= v;
    for (int v : s.voc_up_rates) s.samples_per_frame *= v;
    if (s.head_dim != ATT_D || s.voc_head_dim != 64) { h->err = "unsupported head_dim"; return 1; }
    if (s.heads != 2 * s.kv_heads || s.cp_heads != 2 * s.cp_kv_heads) { h->err = "unsupported GQA ratio (need 2)"; return 1; }
    if (s.cp_steps != N_CODEBOOKS - 1 || s.cp_steps + 2 > (1 << CP_PAGE_SHIFT)) { h->err = "unsupported cp_steps"; return 1; }
    if (s.vocab > SMP_MAXV || s.cp_vocab > SMP_MAXV) { h->err = "vocab too large for the sampler"; return 1; }
    if ((s.hidden % 8) || (s.inter % 8) || (s.text_dim % 8)) { h->err = "dims must be multiples of 8"; return 1; }

    bool ok = true;
    {
        const int H0 = s.hidden, Hc0 = s.cp_hidden, D0 = s.head_dim;
        h->text_embed = need<bf16>(h, h->f_text, "embed", ok, {s.text_vocab, s.text_dim});
        h->fc1w = need<bf16>(h, h->f_text, "fc1.weight", ok, {s.text_dim, s.text_dim}); h->fc1b = need<float>(h, h->f_text, "fc1.bias", ok, {s.text_dim});
        h->fc2w = need<bf16>(h, h->f_text, "fc2.weight", ok, {H0, s.text_dim}); h->fc2b = need<float>(h, h->f_text, "fc2.bias", ok, {H0});
        h->codec_embed = need<bf16>(h, h->f_codec, "embed", ok, {s.vocab, H0});
        h->cp_embed = need<bf16>(h, h->f_cpe, "embed", ok, {s.cp_steps, s.cp_vocab, H0});
        ok = load_layers(h, h->f_talker, "", s.layers, H0, s.heads * D0, s.kv_heads * D0, D0, s.inter, true, false, h->tl) && ok;
        h->t_norm = need<float>(h, h->f_talker, "norm", ok, {H0}); h->t_head = need<bf16>(h, h->f_talker, "head", ok, {s.vocab, H0});
        h->t_cos = need<float>(h, h->f_talker, "rope_cos", ok, {s.max_pos, D0 / 2}); h->t_sin = need<float>(h, h->f_talker, "rope_sin", ok, {s.max_pos, D0 / 2});
        ok = load_layers(h, h->f_cp, "", s.cp_layers, Hc0, s.cp_heads * D0, s.cp_kv_heads * D0, D0, s.cp_inter, true, false, h->cl) && ok;
        h->c_norm = need<float>(h, h->f_cp, "norm", ok, {Hc0}); h->c_heads = need<bf16>(h, h->f_cp, "heads", ok, {s.cp_steps, s.cp_vocab, Hc0});
        h->c_cos = need<float>(h, h->f_cp, "rope_cos", ok, {s.cp_max_pos, D0 / 2}); h->c_sin = need<float>(h, h->f_cp, "rope_sin", ok, {s.cp_max_pos, D0 / 2});
        if (s.hidden != s.cp_hidden) {
            h->c_inproj_w = need<bf16>(h, h->f_cp, "in_proj.weight", ok, {Hc0, H0}); h->c_inproj_b = need<float>(h, h->f_cp, "in_proj.bias", ok, {Hc0});
        }
    }
    ok = load_vocoder_weights(h) && ok;
    if (!ok) return 1;

    const int H = s.hidden, Hc = s.cp_hidden, D = ATT_D;
    if (dalloc(h, &h->x, H) || dalloc(h, &h->qkv, (s.heads + 2 * s.kv_heads) * D) || dalloc(h, &h->attn, s.heads * D) ||
        dalloc(h, &h->act, s.inter) || dalloc(h, &h->logits, s.vocab) || dalloc(h, &h->last_hidden, H) ||
        dalloc(h, &h->cx, Hc) || dalloc(h, &h->cxin, Hc) || dalloc(h, &h->cqkv, (s.cp_heads + 2 * s.cp_kv_heads) * D) ||
        dalloc(h, &h->cattn, s.cp_heads * D) || dalloc(h, &h->cact, s.cp_inter) || dalloc(h, &h->clogits, s.cp_vocab) ||
        dalloc(h, &h->cp_in, (size_t)2 * H) || dalloc(h, &h->next_in, H) || dalloc(h, &h->spk_dev, H) ||
        dalloc(h, &h->partial, (size_t)s.kv_heads * ATT_NSPLIT * 2 * ATT_PSTRIDE) ||
        dalloc(h, &h->cpartial, (size_t)s.cp_kv_heads * 2 * ATT_PSTRIDE) ||
        dalloc(h, &h->counters, s.kv_heads) || dalloc(h, &h->ccounters, s.cp_kv_heads))
        return 1;
    h->max_pages = (s.max_pos + KV_PAGE - 1) / KV_PAGE;
    const size_t page_elems = (size_t)s.layers * 2 * s.kv_heads * KV_PAGE * D;
    {
        const size_t bytes = page_elems * h->max_pages * h->n_slots * (h->kv_f32 ? sizeof(float) : sizeof(bf16));
        CK(cudaMalloc(&h->kv_pool, bytes));
        CK(cudaMemset(h->kv_pool, 0, bytes));
    }
    if (dalloc(h, &h->cp_kv, (size_t)s.cp_layers * 2 * s.cp_kv_heads * (1 << CP_PAGE_SHIFT) * D)) return 1;
    {
        std::vector<int> pt((size_t)h->n_slots * h->max_pages);
        for (size_t i = 0; i < pt.size(); ++i) pt[i] = (int)i;           // slot s owns pages [s*max_pages, ...)
        if (dalloc(h, &h->page_tables, pt.size())) return 1;
        CK(cudaMemcpy(h->page_tables, pt.data(), pt.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (dalloc(h, &h->cp_page_table, 1)) return 1;
        std::vector<int> pc(32);
        for (int i = 0; i < 32; ++i) pc[i] = i;
        if (dalloc(h, &h->cp_pos_consts, 32)) return 1;
        CK(cudaMemcpy(h->cp_pos_consts, pc.data(), 32 * sizeof(int), cudaMemcpyHostToDevice));
    }
    h->slot_len.assign(h->n_slots, 0);
    h->max_frames_cap = 4096;
    if (dalloc(h, &h->st, 1) || dalloc(h, &h->sampling_dev, 1) || dalloc(h, &h->token_dev, 1) ||
        dalloc(h, &h->codes_dev, (size_t)h->max_frames_cap * N_CODEBOOKS) ||
        dalloc(h, &h->forced_dev, (size_t)h->max_frames_cap * N_CODEBOOKS) ||
        dalloc(h, &h->trailing_dev, (size_t)s.max_pos * H) || dalloc(h, &h->tts_pad_dev, H) ||
        dalloc(h, &h->prompt_dev, (size_t)16 * H))
        return 1;
    CK(cudaMallocHost((void**)&h->st_host, sizeof(GenState)));
    CK(cudaEventCreate(&h->ev0)); CK(cudaEventCreate(&h->ev1)); CK(cudaEventCreate(&h->ev_t0));
    CK(cudaEventCreate(&h->ev_chunk)); CK(cudaEventCreate(&h->ev_first)); CK(cudaEventCreate(&h->ev_end));
    CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
    if (const char* e = getenv("LQT_FIRST_CHUNK")) h->first_chunk = std::max(0, atoi(e));
    if (const char* e = getenv("LQT_STREAM_CHUNK")) h->stream_chunk = std::max(1, atoi(e));
    h->stream_chunk = std::max(h->stream_chunk, h->first_chunk);
    if (voc_tc_init(h)) return 1;
    if (h->frame_impl != LQT_FRAME_GRAPH && h->frame_impl != LQT_FRAME_BATCHED && fk_init(h)) {
        // shapes the persistent kernel does not cover (e.g. the 1.7B talker: > 64 rows per CTA). An explicit request for the
        // persistent kernel fails here, loudly; LQT_FRAME_AUTO runs loops A+B as the CUDA graph of per-op sm_100a kernels instead
        // (still device-only; there is no CPU path) and says so in lqt_stats.frame_impl_active.
        if (h->frame_impl == LQT_FRAME_PERSISTENT) { h->err = "persistent frame kernel unavailable for this model: " + h->err; return 1; }
        fprintf(stderr, "[lqt] persistent frame kernel unavailable (%s): lqt_synthesize_tokens runs the batched tcgen05 path with one slot\n", h->err.c_str());
        h->err.clear();
        h->frame_impl = LQT_FRAME_BATCHED;
    }
    if (h->frame_impl == LQT_FRAME_AUTO) h->frame_impl = LQT_FRAME_PERSISTENT;
    return 0;
}

}  // namespace

#include "batch_engine.inl"

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

int lqt_synthesize_batch(lqt_engine* h, const lqt_batch_request* reqs, int32_t n_reqs, const lqt_sampling* sp, const lqt_batch_options* opt) {
    if (!h || !reqs || n_reqs < 1 || !sp) return 1;
    cudaSetDevice(h->device);
    const int rc = synthesize_batch_impl(h, reqs, n_reqs, sp, opt);
    if (rc) cudaStreamSynchronize(h->stream);
    return rc;
}

int lqt_debug_tc_gemm(lqt_engine* h, const float* W, const float* x, int32_t N, int32_t K, int32_t B, int32_t planes, int32_t splits, float* out) {
    if (!h || !W || !x || !out) return 1;
    cudaSetDevice(h->device);
    return debug_tc_gemm_impl(h, W, x, N, K, B, planes, splits, out);
}

const char* lqt_create_error(void) { return g_create_error.c_str(); }

int lqt_create(const char* model_dir, int device_id, lqt_engine** out) {
    lqt_options o{};
    o.kv_dtype = LQT_KV_BF16; o.n_slots = 0; o.frame_impl = LQT_FRAME_AUTO;
    return lqt_create_ex(model_dir, device_id, &o, out);
}

int lqt_check_model_file(const char* path, char* err, int32_t err_cap) {
    const std::string e = path ? check_lqw(path) : std::string("null path");
    if (err && err_cap > 0) { std::strncpy(err, e.c_str(), (size_t)err_cap - 1); err[err_cap - 1] = 0; }
    return e.empty() ? 0 : 1;
}

int lqt_create_ex(const char* model_dir, int device_id, const lqt_options* opt, lqt_engine** out) {
    if (out) *out = nullptr;
    if (!model_dir || !out) { g_create_error = "null argument"; return 1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_create_error = "no CUDA device (this library has no CPU fallback)";
        return 1;
    }
    if (device_id < 0 || device_id >= ndev) { g_create_error = "bad device id"; return 1; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device_id);
    if (prop.major < 10) { g_create_error = "device is not sm_100 (Blackwell) class"; return 1; }
    if (cudaSetDevice(device_id) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return 1; }
    lqt_engine* h = new lqt_engine();
    h->device = device_id;
    h->num_sms = prop.multiProcessorCount;
    h->n_slots = 2;
    if (opt) {
        if (opt->kv_dtype != LQT_KV_BF16 && opt->kv_dtype != LQT_KV_F32) { g_create_error = "bad kv_dtype"; delete h; return 1; }
        h->kv_f32 = opt->kv_dtype == LQT_KV_F32;
        if (opt->n_slots > 0) h->n_slots = opt->n_slots;
        if (opt->frame_impl < LQT_FRAME_PERSISTENT || opt->frame_impl > LQT_FRAME_BATCHED) { g_create_error = "bad frame_impl"; delete h; return 1; }
        h->frame_impl = opt->frame_impl;
    }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_create_error = "cudaStreamCreate failed"; delete h; return 1;
    }
    if (init_engine(h, model_dir)) {
        g_create_error = h->err.empty() ? "engine initialisation failed" : h->err;
        lqt_destroy(h);
        return 1;
    }
    *out = h;
    return 0;
}

void lqt_destroy(lqt_engine* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    batch_destroy(h->batch); h->batch = nullptr;
    voc_tc_destroy(h);
    voc_stream_clear(h);
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.second);
    for (auto& w : h->ws) if (w.second.first) cudaFree(w.second.first);
    void* bufs[] = {h->x, h->qkv, h->attn, h->act, h->logits, h->last_hidden, h->cx, h->cxin, h->cqkv, h->cattn, h->cact,
                    h->clogits, h->cp_in, h->next_in, h->partial, h->cpartial, h->counters, h->ccounters, h->kv_pool,
                    h->cp_kv, h->page_tables, h->cp_page_table, h->cp_pos_consts, h->st, h->sampling_dev, h->codes_dev,
                    h->forced_dev, h->trailing_dev, h->tts_pad_dev, h->prompt_dev, h->token_dev, h->spk_dev,
                    h->voc_codes_dev, h->audio_dev};
    for (void* b : bufs) if (b) cudaFree(b);
    for (void* b : h->fk_allocs) if (b) cudaFree(b);
    if (h->fk_ctrl_host) cudaFreeHost(h->fk_ctrl_host);
    if (h->fk_progress_host) cudaFreeHost(h->fk_progress_host);
    if (h->st_host) cudaFreeHost(h->st_host);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_chunk) cudaEventDestroy(h->ev_chunk);
    if (h->ev_first) cudaEventDestroy(h->ev_first);
    if (h->ev_end) cudaEventDestroy(h->ev_end);
    for (cudaEvent_t e : h->chunk_events) cudaEventDestroy(e);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->chunk_audio_dev) cudaFree(h->chunk_audio_dev);
    if (h->mel_window) { cudaFree(h->mel_window); cudaFree(h->mel_tw_re); cudaFree(h->mel_tw_im); cudaFree(h->mel_tri); }
    h->f_text.release(); h->f_codec.release(); h->f_cpe.release(); h->f_talker.release();
    h->f_cp.release(); h->f_voc.release(); h->f_spk.release();
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char* lqt_last_error(lqt_engine* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int lqt_get_info(lqt_engine* h, lqt_info* o) {
    if (!h || !o) return 1;
    const Spec& s = h->sp;
    o->hidden = s.hidden; o->layers = s.layers; o->heads = s.heads; o->kv_heads = s.kv_heads; o->head_dim = s.head_dim;
    o->vocab = s.vocab; o->cp_vocab = s.cp_vocab; o->cp_steps = s.cp_steps; o->samples_per_frame = s.samples_per_frame;
    o->sample_rate = 24000; o->has_speaker_encoder = h->has_spk ? 1 : 0; o->max_pos = s.max_pos; o->num_sms = h->num_sms;
    return 0;
}

int lqt_get_stats(lqt_engine* h, lqt_stats* o) {
    if (!h || !o) return 1;
    *o = h->stats;
    o->frame_impl_active = h->frame_impl;
    o->cooperative_launch = (h->frame_impl == LQT_FRAME_PERSISTENT && h->fk_coop) ? 1 : 0;
    return 0;
}
int lqt_reset_stats(lqt_engine* h) { if (!h) return 1; h->stats = lqt_stats{}; return 0; }

int lqt_text_project(lqt_engine* h, const int64_t* ids, int32_t S, float* out) {
    if (!h || !ids || !out || S <= 0) return 1;
    cudaSetDevice(h->device);
    for (int i = 0; i < S; ++i)
        if (ids[i] < 0 || ids[i] >= h->sp.text_vocab) { h->err = "text token id out of range"; return 1; }
    long long* d = (long long*)wsbuf(h, "tp_ids", (size_t)S * 2);
    float* o = wsbuf(h, "tp_out", (size_t)S * h->sp.hidden);
    if (!d || !o) return 1;
    CK(cudaMemcpyAsync(d, ids, S * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    if (run_text_project(h, d, S, o)) return 1;
    CK(cudaMemcpyAsync(out, o, (size_t)S * h->sp.hidden * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

static int gather_api(lqt_engine* h, const bf16* table, int rows, const int64_t* ids, int N, float* out) {
    cudaSetDevice(h->device);
    for (int i = 0; i < N; ++i)
        if (ids[i] < 0 || ids[i] >= rows) { h->err = "embedding id out of range"; return 1; }
    const int H = h->sp.hidden;
    long long* d = (long long*)wsbuf(h, "tp_ids", (size_t)N * 2);
    float* o = wsbuf(h, "tp_out", (size_t)N * H);
    if (!d || !o) return 1;
    CK(cudaMemcpyAsync(d, ids, N * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    gather_rows_kernel<<<N, 256, 0, h->stream>>>(table, d, H, o);
    h->stats.kernel_launches++;
    CK(cudaMemcpyAsync(out, o, (size_t)N * H * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int lqt_codec_embed(lqt_engine* h, const int64_t* ids, int32_t N, float* out) {
    if (!h || !ids || !out || N <= 0) return 1;
    return gather_api(h, h->codec_embed, h->sp.vocab, ids, N, out);
}

int lqt_code_predictor_embed(lqt_engine* h, int64_t id, int64_t step, float* out) {
    if (!h || !out) return 1;
    if (step < 0 || step >= h->sp.cp_steps) { h->err = "generation_step out of range"; return 1; }
    return gather_api(h, h->cp_embed + (size_t)step * h->sp.cp_vocab * h->sp.hidden, h->sp.cp_vocab, &id, 1, out);
}

int lqt_kv_reset(lqt_engine* h, int32_t slot) {
    if (!h || slot < 0 || slot >= h->n_slots) return 1;
    h->slot_len[slot] = 0;
    return 0;
}
int lqt_kv_len(lqt_engine* h, int32_t slot) {
    if (!h || slot < 0 || slot >= h->n_slots) return -1;
    return h->slot_len[slot];
}

int lqt_talker_prefill(lqt_engine* h, int32_t slot, const float* embeds, int32_t P, float* logits_last, float* last_hidden) {
    if (!h || !embeds || P <= 0 || slot < 0 || slot >= h->n_slots) return 1;
    cudaSetDevice(h->device);
    const int H = h->sp.hidden;
    if (P > h->sp.max_pos) { h->err = "prefill longer than max_pos"; return 1; }
    float* e = wsbuf(h, "prefill_in", (size_t)P * H);
    if (!e) return 1;
    CK(cudaMemcpyAsync(e, embeds, (size_t)P * H * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    GenState g{}; *h->st_host = g;
    CK(cudaMemcpyAsync(h->st, h->st_host, sizeof(GenState), cudaMemcpyHostToDevice, h->stream));
    if (h->frame_impl == 0) {
        if (fk_launch(h, slot, 0, e, P, 0, false)) return 1;        // prefill only (frame_end = 0)
    } else {
        for (int i = 0; i < P; ++i) {
            run_talker_token(h, slot, e + (size_t)i * H, i == P - 1, nullptr);
            advance_kernel<<<1, 1, 0, h->stream>>>(h->st);
            h->stats.kernel_launches++;
        }
    }
    CK(cudaGetLastError());
    if (logits_last) CK(cudaMemcpyAsync(logits_last, h->logits, h->sp.vocab * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (last_hidden) CK(cudaMemcpyAsync(last_hidden, h->last_hidden, H * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->frame_impl == 0 && fk_check_abort(h)) return 1;
    h->slot_len[slot] = P;
    return 0;
}

int lqt_talker_decode(lqt_engine* h, int32_t slot, const float* embed, float* logits, float* last_hidden) {
    if (!h || !embed || slot < 0 || slot >= h->n_slots) return 1;
    cudaSetDevice(h->device);
    const int H = h->sp.hidden;
    if (h->slot_len[slot] >= h->sp.max_pos) { h->err = "KV cache full"; return 1; }
    GenState g{}; g.pos = h->slot_len[slot]; *h->st_host = g;
    CK(cudaMemcpyAsync(h->st, h->st_host, sizeof(GenState), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->next_in, embed, H * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (h->frame_impl == 0) { if (fk_launch(h, slot, 1, nullptr, 0, 0, false)) return 1; }
    else run_talker_token(h, slot, h->next_in, true, nullptr);
    CK(cudaGetLastError());
    if (logits) CK(cudaMemcpyAsync(logits, h->logits, h->sp.vocab * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (last_hidden) CK(cudaMemcpyAsync(last_hidden, h->last_hidden, H * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->frame_impl == 0 && fk_check_abort(h)) return 1;
    h->slot_len[slot] += 1;
    return 0;
}

int lqt_code_predictor(lqt_engine* h, const float* embeds, int32_t L, int64_t step, float* logits) {
    if (!h || !embeds || !logits) return 1;
    cudaSetDevice(h->device);
    if (L < 1 || L > h->sp.cp_steps + 2 || step < 0 || step >= h->sp.cp_steps) { h->err = "code_predictor: bad L or generation_step"; return 1; }
    const int H = h->sp.hidden;
    float* e = wsbuf(h, "cp_rows", (size_t)L * H);
    if (!e) return 1;
    CK(cudaMemcpyAsync(e, embeds, (size_t)L * H * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    for (int i = 0; i < L; ++i) run_cp_token(h, e + (size_t)i * H, i, i == L - 1 ? (int)step : -1, nullptr);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(logits, h->clogits, h->sp.cp_vocab * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int lqt_vocoder_decode(lqt_engine* h, const int64_t* codes, int32_t T, float* audio, int64_t* length) {
    if (!h || !codes || !audio || T <= 0) return 1;
    cudaSetDevice(h->device);
    for (long long i = 0; i < (long long)T * N_CODEBOOKS; ++i)
        if (codes[i] < 0 || codes[i] >= h->sp.voc_codebook_size) { h->err = "audio code out of range"; return 1; }
    if (ensure_audio(h, T)) return 1;
    CK(cudaMemcpyAsync(h->voc_codes_dev, codes, (size_t)T * N_CODEBOOKS * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->ev0, h->stream));
    if (run_vocoder(h, h->voc_codes_dev, T, h->audio_dev)) return 1;
    CK(cudaEventRecord(h->ev1, h->stream));
    const size_t n = (size_t)T * h->sp.samples_per_frame;
    CK(cudaMemcpyAsync(audio, h->audio_dev, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->stats.last_vocoder_ms, h->ev0, h->ev1);
    if (length) *length = (int64_t)n;
    return 0;
}

// log-mel tables, built like the host extractor's (host/io/mel.cpp == reference src/io/mel.cpp:15-18, 32-80, 132-158):
// symmetric Hann with the N-1 denominator evaluated in double, HTK mel points on floor((n_fft+1) f / sr) bins, twiddles
// cos/sin of the f32 angle per stage
static int mel_tables_init(lqt_engine* h) {
    if (h->mel_window) return 0;
    const int win = 1024, nfft = 1024, mels = 128, sr = 24000;
    const float fmin = 0.0f, fmax = 12000.0f;
    std::vector<float> w(win), tr, ti;
    for (int i = 0; i < win; ++i) w[i] = static_cast<float>(0.5f * (1.0f - std::cos(2.0f * M_PI * i / (win - 1))));
    auto hz_to_mel = [](float hz) { return 2595.0f * std::log10(1.0f + hz / 700.0f); };
    auto mel_to_hz = [](float mel) { return 700.0f * (std::pow(10.0f, mel / 2595.0f) - 1.0f); };
    const int bins = nfft / 2 + 1;
    const float lo = hz_to_mel(fmin), hi = hz_to_mel(fmax);
    std::vector<int> edge(mels + 2), tri(mels * 3);
    for (int i = 0; i < mels + 2; ++i) {
        const float mel = lo + (hi - lo) * i / (mels + 1);
        edge[i] = std::min((int)std::floor((nfft + 1) * mel_to_hz(mel) / sr), bins - 1);
    }
    for (int m = 0; m < mels; ++m) { tri[m * 3] = edge[m]; tri[m * 3 + 1] = edge[m + 1]; tri[m * 3 + 2] = edge[m + 2]; }
    for (int size = 2; size <= nfft; size *= 2) {
        const float step = static_cast<float>(-2.0f * M_PI / size);
        for (int k = 0; k < size / 2; ++k) { const float ang = step * k; tr.push_back(std::cos(ang)); ti.push_back(std::sin(ang)); }
    }
    CK(cudaMalloc((void**)&h->mel_window, win * 4)); CK(cudaMalloc((void**)&h->mel_tw_re, tr.size() * 4));
    CK(cudaMalloc((void**)&h->mel_tw_im, ti.size() * 4)); CK(cudaMalloc((void**)&h->mel_tri, tri.size() * 4));
    CK(cudaMemcpy(h->mel_window, w.data(), win * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->mel_tw_re, tr.data(), tr.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->mel_tw_im, ti.data(), ti.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->mel_tri, tri.data(), tri.size() * 4, cudaMemcpyHostToDevice));
    return 0;
}

// audio (host, 24 kHz mono) -> device log-mel [frames][128] in the "spk_in" workspace; *frames_out = frame count
static int logmel_device(lqt_engine* h, const float* audio, int64_t n, int* frames_out) {
    if (mel_tables_init(h)) return 1;
    const int win = 1024, hop = 256, nfft = 1024, mels = 128;
    const int frames = (n < win) ? 1 : (int)((n - win) / hop + 1);
    float* a = wsbuf(h, "mel_audio", (size_t)n);
    float* in = wsbuf(h, "spk_in", (size_t)frames * mels);
    if (!a || !in) return 1;
    CK(cudaMemcpyAsync(a, audio, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    MelParams p{};
    p.audio = a; p.n = n; p.window = h->mel_window; p.tw_re = h->mel_tw_re; p.tw_im = h->mel_tw_im; p.tri = h->mel_tri; p.out = in;
    p.frames = frames; p.hop = hop; p.win = win; p.n_fft = nfft; p.log2n = 10; p.num_mels = mels;
    logmel_kernel<<<frames, 256, (size_t)(2 * nfft + nfft / 2 + 1) * sizeof(float), h->stream>>>(p);
    h->stats.kernel_launches++;
    CK(cudaGetLastError());
    *frames_out = frames;
    return 0;
}

static int speaker_encoder_device(lqt_engine* h, int frames, float* out);

int lqt_log_mel(lqt_engine* h, const float* audio, int64_t n_samples, float* mel_out, int32_t* frames) {
    if (!h || !audio || n_samples <= 0 || !frames) return 1;
    cudaSetDevice(h->device);
    int nf = 0;
    if (logmel_device(h, audio, n_samples, &nf)) return 1;
    *frames = nf;
    if (mel_out) {                                   // reference layout [num_mels][frames] (MelExtractor::extract)
        std::vector<float> t((size_t)nf * 128);
        CK(cudaMemcpyAsync(t.data(), h->ws["spk_in"].first, t.size() * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int f = 0; f < nf; ++f) for (int m = 0; m < 128; ++m) mel_out[(size_t)m * nf + f] = t[(size_t)f * 128 + m];
    }
    return 0;
}

int lqt_speaker_embed_audio(lqt_engine* h, const float* audio, int64_t n_samples, float* out) {
    if (!h || !audio || !out || n_samples <= 0) return 1;
    if (!h->has_spk) { h->err = "speaker encoder not available"; return 1; }
    if (h->sp.spk_mels != 128) { h->err = "speaker encoder expects 128 mel bands"; return 1; }
    cudaSetDevice(h->device);
    int nf = 0;
    if (logmel_device(h, audio, n_samples, &nf)) return 1;
    return speaker_encoder_device(h, nf, out);
}

int lqt_vocoder_stream_reset(lqt_engine* h) {
    if (!h) return 1;
    cudaSetDevice(h->device);
    CK(cudaStreamSynchronize(h->stream));
    return voc_stream_reset(h, h->stream);
}

int lqt_vocoder_stream_chunk(lqt_engine* h, const int64_t* codes, int32_t T, float* audio, int64_t* length) {
    if (!h || !codes || !audio || T <= 0) return 1;
    cudaSetDevice(h->device);
    for (long long i = 0; i < (long long)T * N_CODEBOOKS; ++i)
        if (codes[i] < 0 || codes[i] >= h->sp.voc_codebook_size) { h->err = "audio code out of range"; return 1; }
    if (ensure_audio(h, T)) return 1;
    CK(cudaMemcpyAsync(h->voc_codes_dev, codes, (size_t)T * N_CODEBOOKS * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    if (run_vocoder(h, h->voc_codes_dev, T, h->audio_dev, nullptr, &h->voc_stream)) return 1;
    h->voc_stream.t0 += T;
    const size_t n = (size_t)T * h->sp.samples_per_frame;
    CK(cudaMemcpyAsync(audio, h->audio_dev, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (length) *length = (int64_t)n;
    return 0;
}

int lqt_speaker_encoder(lqt_engine* h, const float* mel_t, int32_t frames, float* out) {
    if (!h || !mel_t || !out || frames <= 0) return 1;
    if (!h->has_spk) { h->err = "speaker encoder not available"; return 1; }
    cudaSetDevice(h->device);
    float* in = wsbuf(h, "spk_in", (size_t)frames * h->sp.spk_mels);
    if (!in) return 1;
    CK(cudaMemcpyAsync(in, mel_t, (size_t)frames * h->sp.spk_mels * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    return speaker_encoder_device(h, frames, out);
}

// the speaker_encoder graph on the device mel [frames][mels] in the "spk_in" workspace
static int speaker_encoder_device(lqt_engine* h, int frames, float* out) {
    const Spec& s = h->sp;
    const int Cs = s.spk_channels, M = s.spk_mels;
    bool ok = true;
    const LqwFile& f = h->f_spk;
    float* in = h->ws["spk_in"].first;
    float* a = wsbuf(h, "spk_a", (size_t)frames * Cs);
    float* b = wsbuf(h, "spk_b", (size_t)frames * Cs);
    float* pool = wsbuf(h, "spk_pool", (size_t)2 * Cs);
    float* o = wsbuf(h, "spk_out", s.hidden);
    if (!in || !a || !b || !pool || !o) return 1;
    const long long n = (long long)frames * Cs;
    const int eb = (int)std::min<long long>((n + 255) / 256, 4096);
    { ConvGemmParams p = cg(in, frames, M, need<bf16>(h, f, "in_conv.weight", ok, {Cs, 5, M}), Cs, b); p.taps = 5; p.shift = 2; p.bias = need<float>(h, f, "in_conv.bias", ok, {Cs});
      if (!ok) return 1; launch_conv_gemm(h, p); }
    relu_add_kernel<<<eb, 256, 0, h->stream>>>(b, nullptr, a, n); h->stats.kernel_launches++;
    for (int i = 0; i < s.spk_layers; ++i) {
        const std::string q = "l" + std::to_string(i) + ".conv.";
        ConvGemmParams p = cg(a, frames, Cs, need<bf16>(h, f, q + "weight", ok, {Cs, 3, Cs}), Cs, b); p.taps = 3; p.shift = 1; p.bias = need<float>(h, f, q + "bias", ok, {Cs});
        if (!ok) return 1;
        launch_conv_gemm(h, p);
        relu_add_kernel<<<eb, 256, 0, h->stream>>>(b, a, a, n); h->stats.kernel_launches++;
    }
    stat_pool_kernel<<<(Cs + 127) / 128, 128, 0, h->stream>>>(a, frames, Cs, pool); h->stats.kernel_launches++;
    GemvParams g = gemv_params(need<bf16>(h, f, "fc.weight", ok, {s.hidden, 2 * Cs}), s.hidden, 2 * Cs, pool, o);
    g.bias = need<float>(h, f, "fc.bias", ok, {s.hidden});
    if (!ok) return 1;
    launch_gemv(h, g, false);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, o, s.hidden * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int lqt_sample(lqt_engine* h, const float* logits, int32_t V, const lqt_sampling* sp, uint32_t frame, uint32_t codebook,
               int32_t mask_codec_specials, int64_t* id) {
    if (!h || !logits || !sp || !id || V <= 0 || V > SMP_MAXV) return 1;
    cudaSetDevice(h->device);
    float* d = wsbuf(h, "sample_logits", V);
    if (!d) return 1;
    CK(cudaMemcpyAsync(d, logits, V * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (upload_sampling(h, sp)) return 1;
    SampleParams s{};
    s.logits = d; s.V = V; s.sp = h->sampling_dev; s.frame_imm = frame; s.codebook = (int)codebook; s.token_out = h->token_dev;
    s.n_codebooks = N_CODEBOOKS; s.eos_id = CODEC_EOS;
    if (mask_codec_specials) { s.mask_lo = 2048; s.mask_hi = V; s.mask_keep = CODEC_EOS; }
    launch_sample(h, s);
    CK(cudaGetLastError());
    int tok = 0;
    CK(cudaMemcpyAsync(&tok, h->token_dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *id = tok;
    return 0;
}

int lqt_generate(lqt_engine* h, int32_t slot, const float* prompt, int32_t P, const float* trailing, int32_t trailing_len,
                 const float* tts_pad, const lqt_sampling* sp, const int64_t* forced_codes, int32_t n_forced,
                 int64_t* codes_out, int32_t* n_frames, float* logits_trace, int32_t trace_stride) {
    if (!h || !prompt || !tts_pad || !sp || !codes_out || !n_frames) return 1;
    cudaSetDevice(h->device);
    const int H = h->sp.hidden;
    if (P < 1 || P > 16) { h->err = "P must be in [1,16]"; return 1; }
    if (trailing_len < 0 || trailing_len > h->sp.max_pos) { h->err = "bad trailing_len"; return 1; }
    if (n_forced < 0 || n_forced > h->max_frames_cap) { h->err = "bad n_forced"; return 1; }
    const bool trace = logits_trace != nullptr;
    if (trace) {
        if (trace_stride < std::max(h->sp.vocab, h->sp.cp_vocab)) { h->err = "trace_stride too small"; return 1; }
        if (ensure_trace(h, sp->max_new_tokens, trace_stride)) return 1;
    }
    CK(cudaMemcpyAsync(h->prompt_dev, prompt, (size_t)P * H * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (trailing_len > 0) CK(cudaMemcpyAsync(h->trailing_dev, trailing, (size_t)trailing_len * H * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->tts_pad_dev, tts_pad, H * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (n_forced > 0) CK(cudaMemcpyAsync(h->forced_dev, forced_codes, (size_t)n_forced * N_CODEBOOKS * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    int nf = 0;
    if (generate_core(h, slot, P, trailing_len, sp, n_forced, trace, &nf)) return 1;
    if (nf > 0) CK(cudaMemcpyAsync(codes_out, h->codes_dev, (size_t)nf * N_CODEBOOKS * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    if (trace && sp->max_new_tokens > 0)
        CK(cudaMemcpyAsync(logits_trace, h->trace_dev, (size_t)sp->max_new_tokens * N_CODEBOOKS * trace_stride * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *n_frames = nf;
    return 0;
}

int lqt_debug_timeline(lqt_engine* h, int32_t enable_entries, int32_t cta, uint64_t* out, int32_t out_cap) {
    if (!h) return -1;
    cudaSetDevice(h->device);
    if (enable_entries > 0) {                      // arm: the next frame-kernel launches record CTA `cta`
        if (h->fk_dbg_cap < enable_entries + 1) {
            unsigned long long* d = nullptr;
            if (cudaMalloc((void**)&d, (size_t)(enable_entries + 1) * 8) != cudaSuccess) { h->err = "timeline cudaMalloc failed"; return -1; }
            h->fk_allocs.push_back(d);
            h->fk_dbg = d; h->fk_dbg_cap = enable_entries + 1;
        }
        cudaMemsetAsync(h->fk_dbg, 0, (size_t)h->fk_dbg_cap * 8, h->stream);
        h->fk_dbg_cta = cta;
        return 0;
    }
    if (!h->fk_dbg || !out) return 0;
    unsigned long long n = 0;
    const unsigned long long* base = h->fk_dbg + (cta == 1 ? h->fk_dbg_cap / 2 : 0);      // read mode: cta 1 = the producer's half (FK_FINE_MARKS builds)
    cudaMemcpyAsync(&n, base, 8, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    const int m = (int)std::min<unsigned long long>(n, (unsigned long long)std::max(out_cap, 0));
    if (m > 0) cudaMemcpy(out, base + 1, (size_t)m * 8, cudaMemcpyDeviceToHost);
    return m;
}

// Debug aid (no reference counterpart): values of one exchange buffer of the persistent frame kernel as left by the
// last launch. which: 0 talker x, 1 talker qkv, 2 talker x1, 3 talker act, 4..7 the same for the code predictor.
int lqt_debug_exchange(lqt_engine* h, int32_t which, float* out, int32_t n) {
    if (!h || !out || n <= 0 || h->frame_impl != 0) return 1;
    cudaSetDevice(h->device);
    const FkStack& S = (which & 4) ? h->fk_cp : h->fk_talker;
    const uint2* src = (which & 3) == 0 ? S.x : (which & 3) == 1 ? S.qkv : (which & 3) == 2 ? S.x1 : S.act;
    std::vector<uint2> tmp((size_t)n);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(tmp.data(), src, (size_t)n * sizeof(uint2), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) { float f; memcpy(&f, &tmp[i].x, 4); out[i] = f; }
    return 0;
}

int lqt_build_prompt(lqt_engine* h, const int64_t* token_ids, int32_t n_ids, int32_t lang_codec_id, const float* speaker_embed,
                     float* prompt_out, int32_t* P, float* trailing_out, int32_t* trailing_len, float* tts_pad_out) {
    if (!h || !token_ids || !P || !trailing_len) return 1;
    cudaSetDevice(h->device);
    int p = 0, tl = 0;
    if (build_prompt_device(h, token_ids, n_ids, lang_codec_id, speaker_embed, &p, &tl)) return 1;
    const int H = h->sp.hidden;
    if (prompt_out) CK(cudaMemcpyAsync(prompt_out, h->prompt_dev, (size_t)p * H * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (trailing_out) CK(cudaMemcpyAsync(trailing_out, h->trailing_dev, (size_t)tl * H * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (tts_pad_out) CK(cudaMemcpyAsync(tts_pad_out, h->tts_pad_dev, H * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *P = p; *trailing_len = tl;
    return 0;
}

static int synthesize_tokens_impl(lqt_engine* h, const int64_t* token_ids, int32_t n_ids, int32_t lang_codec_id, const float* speaker_embed,
                                  const lqt_sampling* sp, float* audio_out, int64_t audio_capacity, int64_t* n_samples,
                                  int64_t* codes_out, int32_t* n_frames);

int lqt_synthesize_stream(lqt_engine* h, const int64_t* token_ids, int32_t n_ids, int32_t lang_codec_id, const float* speaker_embed,
                          const lqt_sampling* sp, float* audio_out, int64_t audio_capacity, int64_t* n_samples,
                          int64_t* codes_out, int32_t* n_frames, lqt_audio_callback on_audio, void* user) {
    if (!h) return 1;
    h->on_audio = on_audio; h->on_audio_user = user;
    int64_t ns = 0;
    const int rc = lqt_synthesize_tokens(h, token_ids, n_ids, lang_codec_id, speaker_embed, sp, audio_out, audio_capacity, &ns, codes_out, n_frames);
    // paths that do not stream (graph / batched frame loops, chunking off, short utterances): one callback with everything
    if (rc == 0 && on_audio && ns > 0 && !h->streamed_last) on_audio(user, audio_out, 0, ns);
    h->on_audio = nullptr; h->on_audio_user = nullptr;
    if (n_samples) *n_samples = ns;
    return rc;
}

int lqt_synthesize_tokens(lqt_engine* h, const int64_t* token_ids, int32_t n_ids, int32_t lang_codec_id, const float* speaker_embed,
                          const lqt_sampling* sp, float* audio_out, int64_t audio_capacity, int64_t* n_samples,
                          int64_t* codes_out, int32_t* n_frames) {
    if (!h || !token_ids || !sp || !n_samples) return 1;
    h->streamed_last = false;
    cudaSetDevice(h->device);
    h->chunk_audio_out = audio_out; h->chunk_audio_out_cap = audio_capacity; h->chunk_pending = false;
    const int rc = synthesize_tokens_impl(h, token_ids, n_ids, lang_codec_id, speaker_embed, sp, audio_out, audio_capacity, n_samples, codes_out, n_frames);
    // whatever happened above, no copy into the caller's buffer may still be in flight when this call returns
    if (h->chunk_pending) { cudaStreamSynchronize(h->stream2); h->chunk_pending = false; }
    if (rc) cudaStreamSynchronize(h->stream);
    h->chunk_audio_out = nullptr; h->chunk_audio_out_cap = 0;
    return rc;
}

// Streaming synthesis on the persistent kernel: ONE frame-kernel launch generates the whole utterance and
// reports every finished frame through a pinned host word; the host vocodes the finished frames chunk by chunk on the second
// stream (streaming vocoder with carried state: each chunk is bit-identical to the one-shot decode, nothing is decoded twice)
// and copies the PCM into the caller's buffer while generation continues on the other SMs. After the last frame only the last
// chunk is left to decode. The first chunk is h->first_chunk frames (the first-audio latency), the others h->stream_chunk.
static int synthesize_streaming(lqt_engine* h, int P, int TL, const lqt_sampling* sp, float* audio_out, int64_t audio_capacity,
                                int64_t* n_samples, int64_t* codes_out, int32_t* n_frames) {
    const int chunk = h->stream_chunk, first_n = h->first_chunk, spf = h->sp.samples_per_frame, max_new = sp->max_new_tokens;
    if (P < 1 || max_new < 0 || max_new > h->max_frames_cap || P + max_new > h->sp.max_pos) { h->err = "P + max_new_tokens exceeds max_pos"; return 1; }
    if (audio_capacity < (int64_t)max_new * spf) { h->err = "audio_out too small"; return 1; }
    if (upload_sampling(h, sp)) return 1;
    if (ensure_audio(h, std::max(max_new, chunk))) return 1;      // the throw-away chunk below writes `chunk` frames of PCM
    if (h->voc_reserved_frames < chunk) {
        // size every vocoder workspace (and the per-layer state) for a chunk BEFORE the frame kernel starts: allocations while it
        // runs would serialise behind it. One throw-away chunk of zeros does exactly the allocations the real chunks need.
        CK(cudaMemsetAsync(h->codes_dev, 0, (size_t)chunk * N_CODEBOOKS * sizeof(long long), h->stream));
        if (run_vocoder(h, h->codes_dev, chunk, h->audio_dev, nullptr, &h->voc_stream)) return 1;
        CK(cudaStreamSynchronize(h->stream));
        h->voc_reserved_frames = chunk;
    }
    if (voc_stream_reset(h, h->stream)) return 1;
    GenState g{};
    g.trailing_len = TL; g.max_frames = max_new;
    *h->st_host = g;
    CK(cudaMemcpyAsync(h->st, h->st_host, sizeof(GenState), cudaMemcpyHostToDevice, h->stream));
    *h->fk_progress_host = 0;
    h->fk_progress_on = true;
    CK(cudaEventRecord(h->ev0, h->stream));
    const int rc = fk_launch(h, 0, 0, h->prompt_dev, P, max_new, false);
    h->fk_progress_on = false;
    if (rc) return 1;
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaMemcpyAsync(h->st_host, h->st, sizeof(GenState), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaEventRecord(h->ev_chunk, h->stream));                  // = "the kernel and the state copy are done"
    CK(cudaStreamWaitEvent(h->stream2, h->ev0, 0));               // stream2 starts after the state reset above
    h->chunk_pending = true;
    int done = 0;                                                 // frames vocoded so far
    bool first = true, finished = false;
    volatile int* prog = h->fk_progress_host;
    struct Pending { int frame0, frames; };
    std::vector<Pending> pend;                                    // chunks enqueued, in order; [delivered, pend.size()) not yet reported
    size_t delivered = 0;
    auto deliver = [&](bool wait) {                               // report the chunks whose PCM has landed in the caller's buffer
        while (delivered < pend.size()) {
            if (wait) cudaEventSynchronize(h->chunk_events[delivered]);
            else if (cudaEventQuery(h->chunk_events[delivered]) != cudaSuccess) break;
            if (h->on_audio) h->on_audio(h->on_audio_user, audio_out + (size_t)pend[delivered].frame0 * spf,
                                         (int64_t)pend[delivered].frame0 * spf, (int64_t)pend[delivered].frames * spf);
            ++delivered;
        }
    };
    auto next_size = [&]() { return done == 0 ? first_n : chunk; };
    auto vocode = [&](int upto) -> int {                          // frames [done, upto): the first chunk, then chunks of at most `chunk`
        while (done < upto) {
            const int n = std::min(next_size(), upto - done);
            if (run_vocoder(h, h->codes_dev + (size_t)done * N_CODEBOOKS, n, h->audio_dev + (size_t)done * spf, h->stream2, &h->voc_stream)) return 1;
            h->voc_stream.t0 += n;
            CK(cudaMemcpyAsync(audio_out + (size_t)done * spf, h->audio_dev + (size_t)done * spf, (size_t)n * spf * sizeof(float), cudaMemcpyDeviceToHost, h->stream2));
            if (first) { CK(cudaEventRecord(h->ev_first, h->stream2)); first = false; }
            if (h->chunk_events.size() <= pend.size()) {
                cudaEvent_t e;
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                h->chunk_events.push_back(e);
            }
            CK(cudaEventRecord(h->chunk_events[pend.size()], h->stream2));
            pend.push_back(Pending{done, n});
            done += n;
        }
        return 0;
    };
    while (!finished) {
        finished = cudaEventQuery(h->ev_chunk) == cudaSuccess;
        const int avail = finished ? h->st_host->n_frames : *prog;
        const int upto = finished ? avail : (avail - done >= next_size() ? done + next_size() : done);   // whole chunks while the kernel runs, the rest at the end
        if (upto > done) { if (vocode(upto)) return 1; }
        else if (!finished) { struct timespec ts = {0, 50000}; nanosleep(&ts, nullptr); }
        deliver(false);
    }
    CK(cudaEventRecord(h->ev_end, h->stream2));
    CK(cudaStreamSynchronize(h->stream2));
    deliver(true);
    h->chunk_pending = false;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    if (fk_check_abort(h)) return 1;
    const int nf = h->st_host->n_frames;
    cudaEventElapsedTime(&h->stats.last_generate_ms, h->ev0, h->ev1);
    h->stats.last_prefill_ms = 0.f;
    if (nf > 0) {
        cudaEventElapsedTime(&h->stats.last_total_ms, h->ev_t0, h->ev_end);
        cudaEventElapsedTime(&h->stats.first_audio_ms, h->ev_t0, h->ev_first);
        float tail = 0.f;
        cudaEventElapsedTime(&tail, h->ev1, h->ev_end);           // what is left of the vocoder after the last frame
        h->stats.last_vocoder_ms = std::max(tail, 0.f);
    }
    h->slot_len[0] = h->st_host->pos;
    h->stats.last_frames = nf;
    if (n_frames) *n_frames = nf;
    if (nf > 0 && codes_out) {
        CK(cudaMemcpyAsync(codes_out, h->codes_dev, (size_t)nf * N_CODEBOOKS * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    *n_samples = (int64_t)nf * spf;
    h->streamed_last = true;
    return 0;
}

static int synthesize_tokens_impl(lqt_engine* h, const int64_t* token_ids, int32_t n_ids, int32_t lang_codec_id, const float* speaker_embed,
                                  const lqt_sampling* sp, float* audio_out, int64_t audio_capacity, int64_t* n_samples,
                                  int64_t* codes_out, int32_t* n_frames) {
    *n_samples = 0;
    if (h->frame_impl == LQT_FRAME_BATCHED) {            // shapes the persistent kernel does not take (1.7B): one slot of the batched path
        lqt_batch_request rq{};
        rq.token_ids = token_ids; rq.n_ids = n_ids; rq.lang_codec_id = lang_codec_id; rq.speaker_embed = speaker_embed;
        rq.utterance_id = sp->utterance_id; rq.max_new_tokens = sp->max_new_tokens;
        rq.audio_out = audio_out; rq.audio_capacity = audio_capacity; rq.n_samples = n_samples; rq.codes_out = codes_out;
        int32_t nf = 0; rq.n_frames = &nf;
        std::vector<int64_t> codes_tmp;
        if (!codes_out) { codes_tmp.resize((size_t)std::max(sp->max_new_tokens, 1) * N_CODEBOOKS); rq.codes_out = codes_tmp.data(); }
        lqt_batch_options bo{}; bo.max_concurrent = 1; bo.planes = 3; bo.poll_frames = 8;
        if (synthesize_batch_impl(h, &rq, 1, sp, &bo)) return 1;
        if (n_frames) *n_frames = nf;
        h->stats.first_audio_ms = h->stats.last_total_ms;
        return 0;
    }
    if (n_frames) *n_frames = 0;
    int P = 0, TL = 0;
    CK(cudaEventRecord(h->ev_t0, h->stream));
    if (build_prompt_device(h, token_ids, n_ids, lang_codec_id, speaker_embed, &P, &TL)) return 1;
    h->stats.first_audio_ms = 0.f; h->stats.last_total_ms = 0.f;
    if (h->frame_impl == LQT_FRAME_PERSISTENT && audio_out && h->fk_coop && h->first_chunk > 0 && sp->max_new_tokens > h->first_chunk)
        return synthesize_streaming(h, P, TL, sp, audio_out, audio_capacity, n_samples, codes_out, n_frames);
    int nf = 0, chunk = 0;
    if (generate_core(h, 0, P, TL, sp, 0, false, &nf)) return 1;
    if (n_frames) *n_frames = nf;
    h->stats.last_total_ms = 0.f;
    if (nf == 0) return 0;                                   // empty result, like src/tts_onnx.cpp:418
    if (codes_out) CK(cudaMemcpyAsync(codes_out, h->codes_dev, (size_t)nf * N_CODEBOOKS * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    const int64_t n = (int64_t)nf * h->sp.samples_per_frame;
    if (!audio_out || audio_capacity < n) { h->err = "audio_out too small"; return 1; }
    if (ensure_audio(h, nf)) return 1;
    if (chunk) CK(cudaStreamWaitEvent(h->stream, h->ev_first, 0));      // the chunk pass shares the vocoder workspace
    CK(cudaEventRecord(h->ev0, h->stream));
    if (run_vocoder(h, h->codes_dev, nf, h->audio_dev)) return 1;
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaMemcpyAsync(audio_out, h->audio_dev, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->stats.last_vocoder_ms, h->ev0, h->ev1);
    cudaEventElapsedTime(&h->stats.last_total_ms, h->ev_t0, h->ev1);
    if (chunk) cudaEventElapsedTime(&h->stats.first_audio_ms, h->ev_t0, h->ev_first);
    else h->stats.first_audio_ms = h->stats.last_total_ms;
    *n_samples = n;
    return 0;
}

}  // extern "C"
