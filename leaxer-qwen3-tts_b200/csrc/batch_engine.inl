// Host side of the BATCHED path (included by engine.cu): lqt_synthesize_batch runs up to `max_concurrent` utterances in
// lockstep slots on one GPU -- continuous batching over KV slots with a page allocator -- through a CUDA graph of
// tcgen05 GEMMs (tc_gemm.cuh) and the per-slot kernels of batched.cuh. BASELINE configs[3] (256 concurrent utterances) and
// [4] (1.7B talker): any hidden / MLP width that is a multiple of 64 works here; there is no shape the graph cannot take.
// The reference itself is strictly one utterance per call (src/tts_onnx.cpp:405-436, batch dim 1 at :547, 618, 672-674).

struct BGemm {
    CUtensorMap w, w2;
    const CUtensorMap* x = nullptr;
    int N = 0, K = 0, n_split2 = 0, S = 1, kb_per_split = 1;
    float* part = nullptr; long long split_stride = 0;
};
struct BLayer { BGemm qkv, o, gu, down; const float *ln1, *ln2, *qnorm, *knorm; };

struct lqt_batch {
    int B = 0, Bt = 0, n_tiles = 0, planes = 3, Bpad = 0, BN = 0, frames_cap = 0, tl_cap = 0, stages = 2;
    int max_pages = 0, n_pages = 0, nsplit_attn = 1;
    bool with_trace = false; int trace_stride = 0;
    std::vector<void*> allocs;
    BatchState* st = nullptr; BatchState* st_host = nullptr;
    SamplingDev* sp = nullptr;
    bf16 *xn_t = nullptr, *attn_t = nullptr, *act_t = nullptr, *xn_c = nullptr, *attn_c = nullptr, *act_c = nullptr, *xin_c = nullptr;
    CUtensorMap m_xn_t, m_attn_t, m_act_t, m_xn_c, m_attn_c, m_act_c, m_xin_c;
    float *xa_t = nullptr, *xb_t = nullptr, *xa_c = nullptr, *xb_c = nullptr, *last_hidden = nullptr, *cp_in = nullptr, *next_in = nullptr;
    float *prompt = nullptr, *trailing = nullptr, *tts_pad = nullptr, *trace = nullptr;
    long long *codes = nullptr, *forced = nullptr;
    std::vector<BLayer> tl, cl;
    BGemm t_head, c_inproj; std::vector<BGemm> c_heads;
    float *p_qkv = nullptr, *p_o = nullptr, *p_gu = nullptr, *p_down = nullptr, *p_logits = nullptr;      // split-K partials (shared by both stacks)
    void* kv_pool = nullptr; int* page_table = nullptr; std::vector<int> page_table_host; std::vector<int> free_pages;
    std::vector<std::vector<int>> slot_pages;
    float* cp_kv = nullptr; int* cp_page_table = nullptr;
    float* attn_partial = nullptr; int* attn_counters = nullptr;
    cudaGraphExec_t graph = nullptr; int kernels_per_frame = 0;
};

namespace {

template <typename T>
int balloc(lqt_engine* h, lqt_batch* bt, T** p, size_t n) {
    CK(cudaMalloc((void**)p, n * sizeof(T)));
    CK(cudaMemsetAsync(*p, 0, n * sizeof(T), h->stream));
    bt->allocs.push_back((void*)*p);
    return 0;
}

// K splits of a GEMM with N_total weight rows, from a small cost model (microseconds): waves x (k-blocks per CTA x time per
// k-block + fixed CTA cost) + the consumers' reduction of the split partials. Few utterances: the machine is empty, split
// until ~every SM has a CTA. Many utterances (BN = 256): a CTA's ring fills the SM (one CTA per SM), partials are megabytes --
// 192 CTAs on 148 SMs is two waves for the price of two, so fewer splits win (profiles/r2_batched_launches_b256.md).
int bgemm_splits(const lqt_engine* h, const lqt_batch* bt, int N_total, int K) {
    const int tiles = (N_total + TG_BM - 1) / TG_BM * bt->n_tiles, nkb = K / TG_BK;
    const double stage_kb = (double)tc_gemm_stage_bytes(bt->BN) / 1024.0;
    const double t_k = 0.10 + stage_kb / 140.0;                      // one k-block: ~140 KB/us of TMA ingest per SM + issue
    const double t_fix = 3.0;                                        // prologue (barriers, TMEM) + epilogue + launch slack
    int best = 1; double best_cost = 1e30;
    for (int S = 1; S <= nkb; ++S) {
        const int per = (nkb + S - 1) / S, Se = (nkb + per - 1) / per;
        if (Se != S) continue;
        const int st = std::min(bt->stages, per);
        const size_t smem = tc_gemm_smem_bytes(bt->BN, st);
        const int occ = std::max(1, std::min(4, (int)((220 * 1024) / smem)));
        const int waves = (tiles * S + h->num_sms * occ - 1) / (h->num_sms * occ);
        const double reduce = (double)S * bt->Bpad * N_total * 4.0 / 3.0e6 + 0.15 * S;      // partial bytes at ~3 TB/s (L2) + latency per split
        const double cost = waves * (per * t_k * (occ > 1 ? 0.75 : 1.0) + t_fix) + reduce;
        if (cost < best_cost) { best_cost = cost; best = S; }
    }
    return best;
}

// one weight matrix [N][K] (optionally a second one stacked behind it: gate | up) against X operand `x`
int bgemm_init(lqt_engine* h, lqt_batch* bt, BGemm* g, const bf16* W, const bf16* W2, int N, int K, const CUtensorMap* x, float* part) {
    g->N = W2 ? 2 * N : N; g->K = K; g->n_split2 = W2 ? N : g->N; g->x = x; g->part = part;
    if (N % TG_BM && W2) { h->err = "batched path: the MLP width must be a multiple of 128"; return 1; }
    if (make_map(h, &g->w, W, N, K, TG_BM)) return 1;
    if (make_map(h, &g->w2, W2 ? W2 : W, N, K, TG_BM)) return 1;
    const int nkb = K / TG_BK;
    g->S = bgemm_splits(h, bt, g->N, K);
    g->kb_per_split = (nkb + g->S - 1) / g->S;
    g->split_stride = (long long)bt->Bpad * g->N;
    return 0;
}

// every kernel of the frame graph is launched with the programmatic-serialization attribute (PDL, common.cuh): inside the
// captured graph the edges become programmatic dependencies, so kernel k+1's prologue overlaps kernel k's tail ($LQT_PDL=0: off)
template <typename... KArgs, typename... Args>
void pdl_launch(lqt_engine* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    static const bool on = !(getenv("LQT_PDL") && atoi(getenv("LQT_PDL")) == 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = on ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
    h->stats.kernel_launches++;
}

template <int COLS>
void tc_launch(lqt_engine* h, const lqt_batch* bt, const BGemm& g, const TcGemmParams& p, dim3 grid, size_t smem) {
    pdl_launch(h, tc_gemm_kernel<COLS>, grid, dim3(TG_THREADS), smem, g.w, g.w2, *g.x, p);
}
void bgemm_launch(lqt_engine* h, const lqt_batch* bt, const BGemm& g) {
    TcGemmParams p{};
    p.N = g.N; p.K = g.K; p.n_split2 = g.n_split2; p.kb_per_split = g.kb_per_split; p.BN = bt->BN; p.Bt = bt->Bt; p.planes = bt->planes;
    p.B = bt->B; p.stages = std::min(bt->stages, g.kb_per_split); p.out = g.part; p.split_stride = g.split_stride;
    const dim3 grid((g.N + TG_BM - 1) / TG_BM, g.S, bt->n_tiles);
    const size_t smem = tc_gemm_smem_bytes(bt->BN, p.stages);
    if (bt->BN <= 32) tc_launch<32>(h, bt, g, p, grid, smem);
    else if (bt->BN <= 64) tc_launch<64>(h, bt, g, p, grid, smem);
    else if (bt->BN <= 128) tc_launch<128>(h, bt, g, p, grid, smem);
    else tc_launch<256>(h, bt, g, p, grid, smem);
}

void bprep_launch(lqt_engine* h, const lqt_batch* bt, const float* resid, bool select_prompt, const BGemm* from, const float* bias,
                  const float* norm_w, float* x_out, float* hid_out, bf16* X, int H) {
    BPrepParams p{};
    p.st = bt->st; p.resid = resid;
    if (select_prompt) { p.prompt = bt->prompt; p.prompt_rows = 16; }
    if (from) { p.part = from->part; p.n_splits = from->S; p.split_stride = from->split_stride; }
    p.bias = bias; p.norm_w = norm_w; p.eps = h->sp.rms_eps; p.x_out = x_out; p.hid_out = hid_out;
    p.X = X; p.planes = bt->planes; p.Bt = bt->Bt; p.H = H;
    pdl_launch(h, bprep_kernel, dim3(bt->B), dim3(256), (size_t)H * sizeof(float), p);
}

struct BStackCtx {
    const std::vector<BLayer>* layers;
    int H, heads, kv_heads, inter;
    const float *cos, *sin;
    float *xa, *xb;
    bf16 *xn, *attn, *act;
    bool talker;
};

// one new position per slot through all layers. `in`: [B][H] input rows (talker: prompt row or next_in per slot state).
// Leaves the last layer's down-projection partials in layers.back().down; the caller applies the final norm.
// kv_only_last: the pass only has to leave its K/V rows behind (predictor position 0: its hidden state is never read, src/tts_onnx.cpp:854-868
// feeds the head with the LAST row only) -- the last layer stops after the attention kernel, which appends them.
void bstack_run(lqt_engine* h, lqt_batch* bt, const BStackCtx& c, const float* in, const BGemm* in_part, const float* in_bias, int fixed_pos, bool kv_only_last = false) {
    const int D = ATT_D, qd = c.heads * D;
    const int nl = (int)c.layers->size();
    for (int l = 0; l < nl; ++l) {
        const BLayer& L = (*c.layers)[l];
        if (l == 0) bprep_launch(h, bt, in, c.talker, in_part, in_bias, L.ln1, c.xa, nullptr, c.xn, c.H);
        else bprep_launch(h, bt, c.xb, false, &(*c.layers)[l - 1].down, nullptr, L.ln1, c.xa, nullptr, c.xn, c.H);
        bgemm_launch(h, bt, L.qkv);
        {
            BAttnParams a{};
            a.st = bt->st; a.qkv_part = L.qkv.part; a.n_splits = L.qkv.S; a.split_stride = L.qkv.split_stride;
            a.qnorm = L.qnorm; a.knorm = L.knorm; a.rope_cos = c.cos; a.rope_sin = c.sin; a.fixed_pos = fixed_pos;
            a.partial = bt->attn_partial; a.counters = bt->attn_counters;
            a.X = c.attn; a.planes = bt->planes; a.Bt = bt->Bt; a.n_kv = c.kv_heads; a.eps = h->sp.rms_eps; a.scale = 1.0f / sqrtf((float)D);
            if (c.talker) {
                a.kv_pool = bt->kv_pool; a.page_table = bt->page_table; a.pt_stride = bt->max_pages; a.page_shift = KV_PAGE_SHIFT;
                a.page_stride = (long long)h->sp.layers * 2 * c.kv_heads * KV_PAGE * D;
                a.layer_off = (long long)l * 2 * c.kv_heads * KV_PAGE * D;
                const dim3 grid(c.kv_heads, bt->nsplit_attn, bt->B);
                const size_t smem = (size_t)2 * ((bt->max_pages + bt->nsplit_attn - 1) / bt->nsplit_attn) * KV_PAGE * sizeof(float);
                if (h->kv_f32) pdl_launch(h, battn_kernel<float>, grid, dim3(ATT_THREADS), smem, a);
                else pdl_launch(h, battn_kernel<bf16>, grid, dim3(ATT_THREADS), smem, a);
            } else {
                const int PSc = 1 << CP_PAGE_SHIFT;
                BCpAttnParams c2{};
                c2.st = bt->st; c2.qkv_part = L.qkv.part; c2.n_splits = L.qkv.S; c2.split_stride = L.qkv.split_stride;
                c2.qnorm = L.qnorm; c2.knorm = L.knorm; c2.rope_cos = c.cos; c2.rope_sin = c.sin; c2.pos = fixed_pos;
                c2.kv = bt->cp_kv; c2.slot_stride = (long long)h->sp.cp_layers * 2 * c.kv_heads * PSc * D;
                c2.layer_off = (long long)l * 2 * c.kv_heads * PSc * D; c2.PS = PSc; c2.n_kv = c.kv_heads; c2.B = bt->B;
                c2.X = c.attn; c2.planes = bt->planes; c2.Bt = bt->Bt; c2.eps = h->sp.rms_eps; c2.scale = a.scale;
                pdl_launch(h, bcp_attn_kernel, dim3((bt->B * c.kv_heads + 7) / 8), dim3(256), 0, c2);
            }
        }
        if (kv_only_last && l == nl - 1) break;
        bgemm_launch(h, bt, L.o);
        bprep_launch(h, bt, c.xa, false, &L.o, nullptr, L.ln2, c.xb, nullptr, c.xn, c.H);
        bgemm_launch(h, bt, L.gu);
        {
            BSwigluParams s{};
            s.st = bt->st; s.part = L.gu.part; s.n_splits = L.gu.S; s.split_stride = L.gu.split_stride;
            s.X = c.act; s.planes = bt->planes; s.Bt = bt->Bt; s.I = c.inter;
            pdl_launch(h, bswiglu_kernel, dim3(bt->B, std::max(1, std::min(4, c.inter / 1024))), dim3(256), 0, s);
        }
        bgemm_launch(h, bt, L.down);
        (void)qd;
    }
}

void bsample_launch(lqt_engine* h, lqt_batch* bt, int codebook, const BGemm& logits) {
    BSampleParams q{};
    q.st = bt->st; q.sp = bt->sp; q.logits_part = logits.part; q.n_splits = logits.S; q.split_stride = logits.split_stride; q.V = logits.N;
    q.codebook = codebook; q.H = h->sp.hidden;
    q.cp_in = bt->cp_in; q.next_in = bt->next_in; q.trailing = bt->trailing; q.trailing_stride = (long long)bt->tl_cap * h->sp.hidden;
    q.tts_pad = bt->tts_pad; q.codes_out = bt->codes; q.codes_stride = (long long)bt->frames_cap * N_CODEBOOKS; q.forced = bt->forced;
    if (bt->with_trace) { q.trace = bt->trace; q.trace_stride = bt->trace_stride; q.trace_bstride = (long long)bt->frames_cap * N_CODEBOOKS * bt->trace_stride; }
    q.eos_id = CODEC_EOS; q.n_codebooks = N_CODEBOOKS;
    if (codebook == 0) {
        q.mask_lo = 2048; q.mask_hi = h->sp.vocab; q.mask_keep = CODEC_EOS;                    // src/tts_onnx.cpp:803-807
        q.embed_table = h->codec_embed;
    } else {
        q.embed_table = h->cp_embed + (size_t)(codebook - 1) * h->sp.cp_vocab * h->sp.hidden;
    }
    pdl_launch(h, bsample_kernel, dim3(bt->B), dim3(BSMP_THREADS), (size_t)logits.N * 20, q);
}

// one lockstep frame (src/tts_onnx.cpp:801-846 for every slot): draw code 0 -> 15 x (predictor pass, draw) -> talker step
void benqueue_frame(lqt_engine* h, lqt_batch* bt) {
    const Spec& s = h->sp;
    bsample_launch(h, bt, 0, bt->t_head);
    BStackCtx cc{&bt->cl, s.cp_hidden, s.cp_heads, s.cp_kv_heads, s.cp_inter, h->c_cos, h->c_sin, bt->xa_c, bt->xb_c, bt->xn_c, bt->attn_c, bt->act_c, false};
    for (int pos = 0; pos <= s.cp_steps; ++pos) {                  // position 0 = talker last_hidden, position j = embedding of the previous code
        const float* row = pos == 0 ? bt->last_hidden : bt->cp_in;
        if (h->c_inproj_w) {                                       // 1.7B: talker width -> predictor width
            bprep_launch(h, bt, row, false, nullptr, nullptr, nullptr, nullptr, nullptr, bt->xin_c, s.hidden);
            bgemm_launch(h, bt, bt->c_inproj);
            bstack_run(h, bt, cc, nullptr, &bt->c_inproj, h->c_inproj_b, pos, pos == 0);
        } else {
            bstack_run(h, bt, cc, row, nullptr, nullptr, pos, pos == 0);
        }
        if (pos >= 1) {
            bprep_launch(h, bt, bt->xb_c, false, &bt->cl.back().down, nullptr, h->c_norm, nullptr, nullptr, bt->xn_c, s.cp_hidden);
            bgemm_launch(h, bt, bt->c_heads[pos - 1]);
            bsample_launch(h, bt, pos, bt->c_heads[pos - 1]);
        }
    }
    BStackCtx tc{&bt->tl, s.hidden, s.heads, s.kv_heads, s.inter, h->t_cos, h->t_sin, bt->xa_t, bt->xb_t, bt->xn_t, bt->attn_t, bt->act_t, true};
    bstack_run(h, bt, tc, bt->next_in, nullptr, nullptr, -1);
    bprep_launch(h, bt, bt->xb_t, false, &bt->tl.back().down, nullptr, h->t_norm, nullptr, bt->last_hidden, bt->xn_t, s.hidden);
    bgemm_launch(h, bt, bt->t_head);
    pdl_launch(h, badvance_kernel, dim3((bt->B + 127) / 128), dim3(128), 0, bt->st, bt->B);
}

void batch_destroy(lqt_batch* bt) {
    if (!bt) return;
    if (bt->graph) cudaGraphExecDestroy(bt->graph);
    for (void* p : bt->allocs) if (p) cudaFree(p);
    if (bt->st_host) cudaFreeHost(bt->st_host);
    delete bt;
}

int batch_stack_init(lqt_engine* h, lqt_batch* bt, const std::vector<LayerW>& w, std::vector<BLayer>* out, int H, int heads, int kv_heads, int inter,
                     const CUtensorMap* xn, const CUtensorMap* attn, const CUtensorMap* act) {
    const int qd = heads * ATT_D, qkvd = (heads + 2 * kv_heads) * ATT_D;
    out->resize(w.size());
    for (size_t l = 0; l < w.size(); ++l) {
        BLayer& L = (*out)[l];
        if (bgemm_init(h, bt, &L.qkv, w[l].wqkv, nullptr, qkvd, H, xn, bt->p_qkv)) return 1;
        if (bgemm_init(h, bt, &L.o, w[l].wo, nullptr, H, qd, attn, bt->p_o)) return 1;
        if (bgemm_init(h, bt, &L.gu, w[l].wgate, w[l].wup, inter, H, xn, bt->p_gu)) return 1;
        if (bgemm_init(h, bt, &L.down, w[l].wdown, nullptr, H, inter, act, bt->p_down)) return 1;
        L.ln1 = w[l].ln1; L.ln2 = w[l].ln2; L.qnorm = w[l].qnorm; L.knorm = w[l].knorm;
    }
    return 0;
}

int batch_create(lqt_engine* h, int B, int planes, int frames_cap, int tl_cap, bool with_trace, lqt_batch** out) {
    const Spec& s = h->sp;
    if (planes < 1 || planes > 3) { h->err = "batch: planes must be 1, 2 or 3"; return 1; }
    if (B < 1 || B > 1024) { h->err = "batch: 1 <= max_concurrent <= 1024"; return 1; }
    if ((s.hidden % 64) || (s.inter % 128) || (s.cp_hidden % 64) || (s.cp_inter % 128)) { h->err = "batch: hidden sizes must be multiples of 64, MLP widths of 128"; return 1; }
    lqt_batch* bt = new lqt_batch();
    *out = bt;
    bt->B = B; bt->planes = planes; bt->frames_cap = frames_cap; bt->tl_cap = tl_cap; bt->with_trace = with_trace;
    const int bt_max = (256 / planes) / 16 * 16;                   // utterances per N tile: planes * Bt <= 256, multiple of 16
    bt->Bt = std::min(bt_max, (B + 15) / 16 * 16);
    bt->n_tiles = (B + bt->Bt - 1) / bt->Bt;
    bt->Bpad = bt->n_tiles * bt->Bt;
    bt->BN = bt->Bt * planes;
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    CK(cudaFuncSetAttribute(tc_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    CK(cudaFuncSetAttribute(tc_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    CK(cudaFuncSetAttribute(tc_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    CK(cudaFuncSetAttribute(tc_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    CK(cudaFuncSetAttribute(bsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    bt->stages = std::max(2, std::min(TG_MAX_STAGES, (int)((optin - 2048 - (int)sizeof(TgShared)) / (int)tc_gemm_stage_bytes(bt->BN))));
    const int H = s.hidden, Hc = s.cp_hidden, D = ATT_D;
    const int qd = s.heads * D, qkvd = (s.heads + 2 * s.kv_heads) * D, cqd = s.cp_heads * D, cqkvd = (s.cp_heads + 2 * s.cp_kv_heads) * D;
    const size_t rows = (size_t)bt->n_tiles * planes * bt->Bt;
    CK(cudaMallocHost((void**)&bt->st_host, (size_t)B * sizeof(BatchState)));
    memset(bt->st_host, 0, (size_t)B * sizeof(BatchState));
    if (balloc(h, bt, &bt->st, B) || balloc(h, bt, &bt->sp, B)) return 1;
    if (balloc(h, bt, &bt->xn_t, rows * H) || balloc(h, bt, &bt->attn_t, rows * qd) || balloc(h, bt, &bt->act_t, rows * s.inter) ||
        balloc(h, bt, &bt->xn_c, rows * Hc) || balloc(h, bt, &bt->attn_c, rows * cqd) || balloc(h, bt, &bt->act_c, rows * s.cp_inter) ||
        balloc(h, bt, &bt->xin_c, rows * H)) return 1;
    if (make_map(h, &bt->m_xn_t, bt->xn_t, rows, H, bt->BN) || make_map(h, &bt->m_attn_t, bt->attn_t, rows, qd, bt->BN) ||
        make_map(h, &bt->m_act_t, bt->act_t, rows, s.inter, bt->BN) || make_map(h, &bt->m_xn_c, bt->xn_c, rows, Hc, bt->BN) ||
        make_map(h, &bt->m_attn_c, bt->attn_c, rows, cqd, bt->BN) || make_map(h, &bt->m_act_c, bt->act_c, rows, s.cp_inter, bt->BN) ||
        make_map(h, &bt->m_xin_c, bt->xin_c, rows, H, bt->BN)) return 1;
    if (balloc(h, bt, &bt->xa_t, (size_t)B * H) || balloc(h, bt, &bt->xb_t, (size_t)B * H) || balloc(h, bt, &bt->xa_c, (size_t)B * Hc) ||
        balloc(h, bt, &bt->xb_c, (size_t)B * Hc) || balloc(h, bt, &bt->last_hidden, (size_t)B * H) || balloc(h, bt, &bt->cp_in, (size_t)B * H) ||
        balloc(h, bt, &bt->next_in, (size_t)B * H) || balloc(h, bt, &bt->prompt, (size_t)B * 16 * H) ||
        balloc(h, bt, &bt->trailing, (size_t)B * tl_cap * H) || balloc(h, bt, &bt->tts_pad, (size_t)B * H) ||
        balloc(h, bt, &bt->codes, (size_t)B * frames_cap * N_CODEBOOKS) || balloc(h, bt, &bt->forced, (size_t)B * frames_cap * N_CODEBOOKS)) return 1;
    if (with_trace) {
        bt->trace_stride = std::max(s.vocab, s.cp_vocab);
        if (balloc(h, bt, &bt->trace, (size_t)B * frames_cap * N_CODEBOOKS * bt->trace_stride)) return 1;
    }
    // split-K partial buffers, one per kind of GEMM and shared by both stacks: [splits][Bpad][N]
    auto psize = [&](int N, int K) { return (size_t)bt->Bpad * N * bgemm_splits(h, bt, N, K); };
    if (balloc(h, bt, &bt->p_qkv, std::max(psize(qkvd, H), psize(cqkvd, Hc))) ||
        balloc(h, bt, &bt->p_o, std::max(std::max(psize(H, qd), psize(Hc, cqd)), psize(Hc, H))) ||
        balloc(h, bt, &bt->p_gu, std::max(psize(2 * s.inter, H), psize(2 * s.cp_inter, Hc))) ||
        balloc(h, bt, &bt->p_down, std::max(psize(H, s.inter), psize(Hc, s.cp_inter))) ||
        balloc(h, bt, &bt->p_logits, std::max(psize(s.vocab, H), psize(s.cp_vocab, Hc)))) return 1;
    if (batch_stack_init(h, bt, h->tl, &bt->tl, H, s.heads, s.kv_heads, s.inter, &bt->m_xn_t, &bt->m_attn_t, &bt->m_act_t)) return 1;
    if (batch_stack_init(h, bt, h->cl, &bt->cl, Hc, s.cp_heads, s.cp_kv_heads, s.cp_inter, &bt->m_xn_c, &bt->m_attn_c, &bt->m_act_c)) return 1;
    if (bgemm_init(h, bt, &bt->t_head, h->t_head, nullptr, s.vocab, H, &bt->m_xn_t, bt->p_logits)) return 1;
    bt->c_heads.resize(s.cp_steps);
    for (int j = 0; j < s.cp_steps; ++j)
        if (bgemm_init(h, bt, &bt->c_heads[j], h->c_heads + (size_t)j * s.cp_vocab * Hc, nullptr, s.cp_vocab, Hc, &bt->m_xn_c, bt->p_logits)) return 1;
    if (h->c_inproj_w && bgemm_init(h, bt, &bt->c_inproj, h->c_inproj_w, nullptr, Hc, H, &bt->m_xin_c, bt->p_o)) return 1;
    // KV: page pool + allocator (talker), one fp32 page per slot (predictor)
    bt->max_pages = (std::min(s.max_pos, 16 + frames_cap) + KV_PAGE - 1) / KV_PAGE;
    bt->n_pages = bt->max_pages * B;
    const size_t page_elems = (size_t)s.layers * 2 * s.kv_heads * KV_PAGE * D;
    {
        const size_t bytes = page_elems * bt->n_pages * (h->kv_f32 ? sizeof(float) : sizeof(bf16));
        CK(cudaMalloc(&bt->kv_pool, bytes));
        bt->allocs.push_back(bt->kv_pool);
        CK(cudaMemsetAsync(bt->kv_pool, 0, bytes, h->stream));
    }
    bt->free_pages.resize(bt->n_pages);
    for (int i = 0; i < bt->n_pages; ++i) bt->free_pages[i] = bt->n_pages - 1 - i;
    bt->slot_pages.assign(B, {});
    bt->page_table_host.assign((size_t)B * bt->max_pages, 0);
    if (balloc(h, bt, &bt->page_table, (size_t)B * bt->max_pages)) return 1;
    if (balloc(h, bt, &bt->cp_kv, (size_t)B * s.cp_layers * 2 * s.cp_kv_heads * (1 << CP_PAGE_SHIFT) * D)) return 1;
    {
        std::vector<int> ident(B);
        for (int i = 0; i < B; ++i) ident[i] = i;
        if (balloc(h, bt, &bt->cp_page_table, B)) return 1;
        CK(cudaMemcpyAsync(bt->cp_page_table, ident.data(), B * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    bt->nsplit_attn = std::max(1, std::min(std::min(ATT_NSPLIT, bt->max_pages), (2 * h->num_sms) / std::max(1, B * s.kv_heads)));
    if (balloc(h, bt, &bt->attn_partial, (size_t)B * std::max(s.kv_heads, s.cp_kv_heads) * bt->nsplit_attn * 2 * ATT_PSTRIDE) ||
        balloc(h, bt, &bt->attn_counters, (size_t)B * std::max(s.kv_heads, s.cp_kv_heads))) return 1;
    {
        const size_t att_smem = (size_t)2 * ((bt->max_pages + bt->nsplit_attn - 1) / bt->nsplit_attn) * KV_PAGE * sizeof(float);
        if (att_smem > 48 * 1024) {
            CK(cudaFuncSetAttribute(battn_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem));
            CK(cudaFuncSetAttribute(battn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem));
        }
    }
    CK(cudaStreamSynchronize(h->stream));
    // the frame graph
    cudaGraph_t g;
    const uint64_t before = h->stats.kernel_launches;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    benqueue_frame(h, bt);
    CK(cudaStreamEndCapture(h->stream, &g));
    bt->kernels_per_frame = (int)(h->stats.kernel_launches - before);
    h->stats.kernel_launches = before;
    CK(cudaGraphInstantiate(&bt->graph, g, 0));
    cudaGraphDestroy(g);
    return 0;
}

}  // namespace

namespace {

struct BReqState { int slot = -1; int P = 0, TL = 0; bool done = false; int n_frames = 0; };

// put request r into slot `slot`: prompt rows, trailing text, pad row, sampler key, KV pages, state
int batch_admit(lqt_engine* h, lqt_batch* bt, const lqt_batch_request& rq, const lqt_sampling* sp, int slot, BReqState* rs) {
    const int H = h->sp.hidden;
    int P = 0, TL = 0;
    if (build_prompt_device(h, rq.token_ids, rq.n_ids, rq.lang_codec_id, rq.speaker_embed, &P, &TL)) return 1;
    if (TL > bt->tl_cap) { h->err = "batch: text longer than the batch context was sized for"; return 1; }
    const int max_new = std::min(rq.max_new_tokens, bt->frames_cap);
    if (P + max_new > h->sp.max_pos) { h->err = "P + max_new_tokens exceeds max_pos"; return 1; }
    const int need = (P + max_new + KV_PAGE - 1) / KV_PAGE;
    if ((int)bt->free_pages.size() < need || need > bt->max_pages) { h->err = "batch: out of KV pages"; return 1; }
    std::vector<int>& mine = bt->slot_pages[slot];
    for (int i = 0; i < need; ++i) {
        mine.push_back(bt->free_pages.back()); bt->free_pages.pop_back();
        bt->page_table_host[(size_t)slot * bt->max_pages + i] = mine.back();
    }
    CK(cudaMemcpyAsync(bt->page_table + (size_t)slot * bt->max_pages, &bt->page_table_host[(size_t)slot * bt->max_pages], need * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(bt->prompt + (size_t)slot * 16 * H, h->prompt_dev, (size_t)P * H * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(bt->trailing + (size_t)slot * bt->tl_cap * H, h->trailing_dev, (size_t)TL * H * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(bt->tts_pad + (size_t)slot * H, h->tts_pad_dev, (size_t)H * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    const int n_forced = rq.forced_codes ? std::min(rq.n_forced, max_new) : 0;
    if (n_forced > 0)
        CK(cudaMemcpyAsync(bt->forced + (size_t)slot * bt->frames_cap * N_CODEBOOKS, rq.forced_codes, (size_t)n_forced * N_CODEBOOKS * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    SamplingDev d{sp->temperature, sp->top_p, sp->top_k, sp->greedy, sp->seed, rq.utterance_id};
    CK(cudaMemcpyAsync(bt->sp + slot, &d, sizeof(d), cudaMemcpyHostToDevice, h->stream));
    BatchState& st = bt->st_host[slot];
    st = BatchState{};
    st.g.trailing_len = TL; st.g.max_frames = max_new; st.g.n_forced = n_forced;
    st.P = P; st.prefill_pos = 0; st.active = 1;
    CK(cudaMemcpyAsync(bt->st + slot, &st, sizeof(BatchState), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));              // `d` and the host page-table row go out of scope / may be rewritten
    rs->slot = slot; rs->P = P; rs->TL = TL;
    return 0;
}

int synthesize_batch_impl(lqt_engine* h, const lqt_batch_request* reqs, int n_reqs, const lqt_sampling* sp, const lqt_batch_options* opt) {
    const Spec& s = h->sp;
    const int planes = opt && opt->planes ? opt->planes : 3;
    const int B = std::max(1, std::min(opt && opt->max_concurrent > 0 ? opt->max_concurrent : n_reqs, n_reqs));
    int frames_cap = 1, tl_cap = 1; bool with_trace = false;
    for (int i = 0; i < n_reqs; ++i) {
        if (!reqs[i].token_ids || reqs[i].n_ids < 5 || reqs[i].max_new_tokens < 0) { h->err = "batch: bad request"; return 1; }
        frames_cap = std::max(frames_cap, std::min(reqs[i].max_new_tokens, s.max_pos - 16));
        tl_cap = std::max(tl_cap, reqs[i].n_ids);
        with_trace = with_trace || reqs[i].logits_trace != nullptr;
        if (reqs[i].n_samples) *reqs[i].n_samples = 0;
        if (reqs[i].n_frames) *reqs[i].n_frames = 0;
    }
    lqt_batch* bt = h->batch;
    if (!bt || bt->B != B || bt->planes != planes || bt->frames_cap < frames_cap || bt->tl_cap < tl_cap || bt->with_trace != with_trace) {
        batch_destroy(h->batch); h->batch = nullptr;
        if (batch_create(h, B, planes, frames_cap, tl_cap, with_trace, &bt)) { batch_destroy(bt); return 1; }
        h->batch = bt;
    }
    // fresh run: every slot empty, every page free
    memset(bt->st_host, 0, (size_t)B * sizeof(BatchState));
    CK(cudaMemcpyAsync(bt->st, bt->st_host, (size_t)B * sizeof(BatchState), cudaMemcpyHostToDevice, h->stream));
    bt->free_pages.resize(bt->n_pages);
    for (int i = 0; i < bt->n_pages; ++i) bt->free_pages[i] = bt->n_pages - 1 - i;
    for (auto& v : bt->slot_pages) v.clear();
    if (with_trace) CK(cudaMemsetAsync(bt->trace, 0, (size_t)B * bt->frames_cap * N_CODEBOOKS * bt->trace_stride * sizeof(float), h->stream));

    std::vector<BReqState> rs(n_reqs);
    std::vector<int> slot_req(B, -1);
    int next_req = 0, finished = 0;
    CK(cudaEventRecord(h->ev_t0, h->stream));
    CK(cudaEventRecord(h->ev0, h->stream));
    const int poll_every = opt && opt->poll_frames > 0 ? opt->poll_frames : 8;
    long long frames_launched = 0;
    while (finished < n_reqs) {
        for (int slot = 0; slot < B && next_req < n_reqs; ++slot) {                     // admit into free slots
            if (slot_req[slot] >= 0) continue;
            if (batch_admit(h, bt, reqs[next_req], sp, slot, &rs[next_req])) return 1;
            slot_req[slot] = next_req++;
        }
        for (int f = 0; f < poll_every; ++f) {
            CK(cudaGraphLaunch(bt->graph, h->stream));
            h->stats.graph_launches++;
            h->stats.kernel_launches += bt->kernels_per_frame;
            ++frames_launched;
        }
        CK(cudaMemcpyAsync(bt->st_host, bt->st, (size_t)B * sizeof(BatchState), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
        for (int slot = 0; slot < B; ++slot) {
            const int r = slot_req[slot];
            if (r < 0 || !bt->st_host[slot].g.done) continue;
            const int nf = bt->st_host[slot].g.n_frames;                                   // utterance finished: harvest, free the slot
            rs[r].done = true; rs[r].n_frames = nf;
            const lqt_batch_request& rq = reqs[r];
            if (rq.n_frames) *rq.n_frames = nf;
            if (nf > 0 && rq.codes_out)
                CK(cudaMemcpyAsync(rq.codes_out, bt->codes + (size_t)slot * bt->frames_cap * N_CODEBOOKS, (size_t)nf * N_CODEBOOKS * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
            if (rq.logits_trace && rq.max_new_tokens > 0)
                CK(cudaMemcpyAsync(rq.logits_trace, bt->trace + (size_t)slot * bt->frames_cap * N_CODEBOOKS * bt->trace_stride,
                                   (size_t)std::min(rq.max_new_tokens, bt->frames_cap) * N_CODEBOOKS * bt->trace_stride * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
            if (nf > 0 && rq.audio_out) {                                                  // vocoder (src/tts_onnx.cpp:430) for this utterance
                const int64_t n = (int64_t)nf * s.samples_per_frame;
                if (rq.audio_capacity < n) { h->err = "audio_out too small"; return 1; }
                if (ensure_audio(h, nf)) return 1;
                if (run_vocoder(h, bt->codes + (size_t)slot * bt->frames_cap * N_CODEBOOKS, nf, h->audio_dev)) return 1;
                CK(cudaMemcpyAsync(rq.audio_out, h->audio_dev, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
                if (rq.n_samples) *rq.n_samples = n;
            }
            for (int pg : bt->slot_pages[slot]) bt->free_pages.push_back(pg);
            bt->slot_pages[slot].clear();
            bt->st_host[slot] = BatchState{};
            CK(cudaMemcpyAsync(bt->st + slot, &bt->st_host[slot], sizeof(BatchState), cudaMemcpyHostToDevice, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            slot_req[slot] = -1;
            ++finished;
        }
    }
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->stats.last_total_ms, h->ev_t0, h->ev1);
    h->stats.last_generate_ms = h->stats.last_total_ms;
    h->stats.last_frames = (int)std::min<long long>(frames_launched, 0x7fffffff);
    return 0;
}

__global__ void f32_to_bf16_kernel(const float* in, bf16* out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void bsum_splits_kernel(const float* part, int S, long long stride, float* out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = 0.f;
        for (int q = 0; q < S; ++q) v += part[q * stride + i];
        out[i] = v;
    }
}
__global__ void bplanes_kernel(const float* x, bf16* X, int B, int K, int planes, int Bt) {
    const int b = blockIdx.x;
    for (int k = threadIdx.x * 4; k < K; k += blockDim.x * 4)
        bstore_planes4(X, K, planes, Bt, b, k, *reinterpret_cast<const float4*>(x + (size_t)b * K + k));
}

// parity surface of the tcgen05 GEMM alone: out[b][n] = sum_k bf16(W[n][k]) * x[b][k] (x fp32, split into `planes` bf16 planes)
int debug_tc_gemm_impl(lqt_engine* h, const float* W, const float* x, int N, int K, int B, int planes, int splits, float* out) {
    if (N < 1 || K % 64 || B < 1 || planes < 1 || planes > 3) { h->err = "debug gemm: bad shape"; return 1; }
    lqt_batch tmp;
    lqt_batch* bt = &tmp;
    const int bt_max = (256 / planes) / 16 * 16;
    bt->B = B; bt->planes = planes; bt->Bt = std::min(bt_max, (B + 15) / 16 * 16); bt->n_tiles = (B + bt->Bt - 1) / bt->Bt;
    bt->Bpad = bt->n_tiles * bt->Bt; bt->BN = bt->Bt * planes;
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    CK(cudaFuncSetAttribute(tc_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    CK(cudaFuncSetAttribute(tc_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    CK(cudaFuncSetAttribute(tc_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    CK(cudaFuncSetAttribute(tc_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    bt->stages = std::max(2, std::min(TG_MAX_STAGES, (int)((optin - 2048 - (int)sizeof(TgShared)) / (int)tc_gemm_stage_bytes(bt->BN))));
    const size_t rows = (size_t)bt->n_tiles * planes * bt->Bt;
    float *Wf = nullptr, *xf = nullptr, *part = nullptr, *of = nullptr; bf16 *Wb = nullptr, *X = nullptr;
    auto cleanup = [&]() { for (void* p : bt->allocs) cudaFree(p); bt->allocs.clear(); };
    BGemm g;
    int rc = 1;
    do {
        if (balloc(h, bt, &Wf, (size_t)N * K) || balloc(h, bt, &xf, (size_t)B * K) || balloc(h, bt, &Wb, (size_t)N * K) || balloc(h, bt, &X, rows * K) ||
            balloc(h, bt, &of, (size_t)B * N)) break;
        CK(cudaMemcpyAsync(Wf, W, (size_t)N * K * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(xf, x, (size_t)B * K * 4, cudaMemcpyHostToDevice, h->stream));
        f32_to_bf16_kernel<<<256, 256, 0, h->stream>>>(Wf, Wb, (long long)N * K);
        bplanes_kernel<<<B, 256, 0, h->stream>>>(xf, X, B, K, planes, bt->Bt);
        CUtensorMap mx;
        if (make_map(h, &mx, X, rows, K, bt->BN)) break;
        if (bgemm_init(h, bt, &g, Wb, nullptr, N, K, &mx, nullptr)) break;
        if (splits > 0) { const int nkb = K / 64; g.S = std::min(splits, nkb); g.kb_per_split = (nkb + g.S - 1) / g.S; g.S = (nkb + g.kb_per_split - 1) / g.kb_per_split; }
        if (balloc(h, bt, &part, (size_t)g.S * bt->Bpad * N)) break;
        g.part = part;
        bgemm_launch(h, bt, g);
        bsum_splits_kernel<<<256, 256, 0, h->stream>>>(part, g.S, g.split_stride, of, (long long)B * N);
        if (cudaGetLastError() != cudaSuccess) { h->err = "debug gemm: launch failed"; break; }
        if (cudaMemcpyAsync(out, of, (size_t)B * N * 4, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) break;
        const cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { h->err = std::string("debug gemm: ") + cudaGetErrorString(e); break; }
        rc = 0;
    } while (false);
    cleanup();
    return rc;
}

}  // namespace
