// Host side of the tcgen05 vocoder decoder (tc_conv.cuh), included by engine.cu. Replaces, for dec.conv_in and the four
// decoder blocks (95 % of the vocoder's flops), the chain snake_kernel -> conv_gemm_mma_kernel of run_vocoder with one
// implicit-GEMM launch per convolution whose epilogue already applies the next SnakeBeta and writes the bf16 planes.
// (included inside engine.cu's anonymous namespace)

int voc_halo(lqt_engine* h, lqt_engine::VocStream* vs, const std::string& key, void* row0, size_t row_bytes, int halo_rows, long long n_rows);

struct VocTcW {                 // per convolution: zero-padded weights [N][taps][Cp] + bias
    bf16* w = nullptr; const float* bias = nullptr;
    int N = 0, taps = 1, Cin = 0, Cp = 0, bias_mod = 0;
    CUtensorMap mw; int BN = 0;
};
struct VocTcSnake { float *ea = nullptr, *ib = nullptr; };

struct VocTcModel {
    bool ready = false;
    int planes = 2;
    VocTcW conv_in;
    struct Blk { VocTcW tconv; VocTcW c1[3], c2[3]; VocTcSnake s_in, s1[3], s2[3]; } blk[8];
    VocTcSnake s_out;
    std::vector<void*> allocs;
    bf16 *xa = nullptr, *xb = nullptr; size_t x_cap = 0;        // plane ping-pong buffers
    bf16* xs = nullptr; size_t xs_cap = 0;                      // planes of an fp32 operand of the front stages (voc_tc_try)
    std::map<const bf16*, VocTcW> lin;                          // front-stage Linear / conv weights used in place (Cin % 64 == 0)
    float *t0 = nullptr; size_t t_cap = 0;                      // fp32 residual stream / final activation
};

int voc_tc_pick_bn(int N) {                                     // largest multiple of 16 that divides N, <= 256
    for (int bn = 256; bn >= 16; bn -= 16) if (N % bn == 0) return bn;
    return 0;
}

int voc_tc_weight(lqt_engine* h, VocTcModel* m, VocTcW* o, const bf16* w, const float* bias, int N, int taps, int Cin, int bias_mod) {
    o->N = N; o->taps = taps; o->Cin = Cin; o->Cp = (Cin + 63) / 64 * 64; o->bias = bias; o->bias_mod = bias_mod;
    o->BN = voc_tc_pick_bn(N);
    if (!o->BN || (Cin % 16)) { h->err = "tcgen05 vocoder: channel counts must be multiples of 16"; return 1; }
    const size_t n = (size_t)N * taps * o->Cp;
    CK(cudaMalloc((void**)&o->w, n * sizeof(bf16)));
    m->allocs.push_back(o->w);
    voc_pad_weight_kernel<<<(int)std::min<size_t>((n + 255) / 256, 4096), 256, 0, h->stream>>>(w, o->w, N, taps, Cin, o->Cp);
    return make_map(h, &o->mw, o->w, N, taps * o->Cp, o->BN);
}
int voc_tc_snake(lqt_engine* h, VocTcModel* m, VocTcSnake* o, const float* alpha, const float* beta, int C) {
    CK(cudaMalloc((void**)&o->ea, (size_t)C * 4)); m->allocs.push_back(o->ea);
    CK(cudaMalloc((void**)&o->ib, (size_t)C * 4)); m->allocs.push_back(o->ib);
    voc_snake_consts_kernel<<<(C + 127) / 128, 128, 0, h->stream>>>(alpha, beta, o->ea, o->ib, C);
    return 0;
}

int voc_tc_init(lqt_engine* h) {
    VocTcModel* m = new VocTcModel();
    h->voc_tc = m;
    const Spec& s = h->sp;
    if (const char* e = getenv("LQT_VOC_PLANES")) m->planes = std::max(1, std::min(3, atoi(e)));
    if (h->vblk.size() > 8 || (s.voc_hidden % 64)) return 0;                     // shapes this path does not take: the fp32-staged kernels stay
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    CK(cudaFuncSetAttribute(tc_conv_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    CK(cudaFuncSetAttribute(tc_conv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    CK(cudaFuncSetAttribute(tc_conv_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    CK(cudaFuncSetAttribute(tc_conv_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    if (voc_tc_weight(h, m, &m->conv_in, h->dec_in_w, h->dec_in_b, s.voc_decoder_dim, 7, s.voc_hidden, s.voc_decoder_dim)) return 1;
    for (size_t b = 0; b < h->vblk.size(); ++b) {
        const VocBlockW& B = h->vblk[b];
        if ((B.cout % 16) || (B.cin % 16)) { h->err.clear(); return 0; }
        auto& K = m->blk[b];
        if (voc_tc_snake(h, m, &K.s_in, B.snake_a, B.snake_b, B.cin)) return 1;
        if (voc_tc_weight(h, m, &K.tconv, B.tconv_w, B.tconv_b, B.stride * B.cout, 2, B.cin, B.cout)) return 1;
        for (int r = 0; r < 3; ++r) {
            if (voc_tc_snake(h, m, &K.s1[r], B.res[r].s1a, B.res[r].s1b, B.cout) || voc_tc_snake(h, m, &K.s2[r], B.res[r].s2a, B.res[r].s2b, B.cout)) return 1;
            if (voc_tc_weight(h, m, &K.c1[r], B.res[r].c1w, B.res[r].c1b, B.cout, 7, B.cout, B.cout)) return 1;
            if (voc_tc_weight(h, m, &K.c2[r], B.res[r].c2w, B.res[r].c2b, B.cout, 1, B.cout, B.cout)) return 1;
        }
    }
    const int Cl = h->vblk.empty() ? s.voc_decoder_dim : h->vblk.back().cout;
    if (voc_tc_snake(h, m, &m->s_out, h->out_sa, h->out_sb, Cl)) return 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    m->ready = getenv("LQT_VOC_TC") ? atoi(getenv("LQT_VOC_TC")) != 0 : true;
    return 0;
}

void voc_tc_destroy(lqt_engine* h) {
    VocTcModel* m = h->voc_tc;
    if (!m) return;
    for (void* p : m->allocs) if (p) cudaFree(p);
    if (m->xa) cudaFree(m->xa);
    if (m->xb) cudaFree(m->xb);
    if (m->xs) cudaFree(m->xs);
    if (m->t0) cudaFree(m->t0);
    delete m;
    h->voc_tc = nullptr;
}

struct VocTcOut {               // what one convolution's epilogue produces
    float* y = nullptr; bool y_snake = false; const float* residual = nullptr; const float* scale = nullptr; int act = 0;
    bf16* xo = nullptr; int cout = 0, up = 1; const VocTcSnake* sn = nullptr;
};

// x: planes of position 0; `halo` rows in front of it hold the previous chunk's tail (streaming decode), else 0
int voc_tc_conv(lqt_engine* h, VocTcModel* m, const VocTcW& W, const bf16* x, long long L, int dil, bool tap_rev, const VocTcOut& o, int halo = 0) {
    CUtensorMap mx;
    if (make_map(h, &mx, x - (size_t)halo * m->planes * W.Cp, L + halo, m->planes * W.Cp, TG_BM)) return 1;
    TcConvParams p{};
    p.x_row0 = halo;
    p.L = (int)L; p.N = W.N; p.BN = W.BN; p.taps = W.taps; p.dil = dil; p.tap_rev = tap_rev ? 1 : 0;
    p.planes = m->planes; p.Cp = W.Cp;
    p.bias = W.bias; p.bias_mod = W.bias_mod; p.residual = o.residual; p.scale = o.scale; p.act = o.act; p.y = o.y; p.y_snake = o.y_snake ? 1 : 0;
    p.xo = o.xo; p.oplanes = m->planes; p.cout = o.cout ? o.cout : W.N; p.oCp = (p.cout + 63) / 64 * 64; p.up = o.up;
    if (o.sn) { p.sn_ea = o.sn->ea; p.sn_ib = o.sn->ib; }
    const size_t stage = tc_gemm_stage_bytes(W.BN);
    const size_t epi_smem = 16 + (size_t)4 * TC_EPI_GROUPS * 32 * TC_STG_PITCH * sizeof(float);      // the epilogue warps' transpose tiles
    p.stages = std::max(2, std::min(TG_MAX_STAGES, (int)((224 * 1024 - 2048 - epi_smem) / stage)));   // persistent CTA: the ring runs on across tiles
    const size_t smem = 1024 + (size_t)p.stages * stage + sizeof(TcShared) + epi_smem;
    const long long tiles = ((L + TG_BM - 1) / TG_BM) * (long long)(W.N / W.BN);
    const dim3 grid((unsigned)std::min<long long>(tiles, h->num_sms));                       // persistent: one CTA per SM
    if (W.BN <= 32) tc_conv_kernel<64><<<grid, TC_THREADS, smem, h->stream>>>(mx, W.mw, p);              // template = 2 accumulator sets
    else if (W.BN <= 64) tc_conv_kernel<128><<<grid, TC_THREADS, smem, h->stream>>>(mx, W.mw, p);
    else if (W.BN <= 128) tc_conv_kernel<256><<<grid, TC_THREADS, smem, h->stream>>>(mx, W.mw, p);
    else tc_conv_kernel<512><<<grid, TC_THREADS, smem, h->stream>>>(mx, W.mw, p);
    h->stats.kernel_launches++;
    return 0;
}

// The GEMM-shaped ops of the stages in FRONT of the decoder (RVQ output projections, pre_conv, the pre-transformer's Linear
// layers, the upsampling stages' transposed / point-wise convs): same arguments as the round-1 conv_gemm kernels, run on the
// tcgen05 kernel instead when the shape allows -- the fp32 operand is split into planes by voc_planes_kernel first (these
// tensors are small: L <= 4 T rows). Returns false if the op has to stay on the old kernel.
bool voc_tc_try(lqt_engine* h, const ConvGemmParams& c) {
    VocTcModel* m = h->voc_tc;
    if (!m || !m->ready || c.shift != 0 || (c.Cin % 64) || (c.N % 16) || c.L < 1 || c.act == 1) return false;
    auto it = m->lin.find(c.W);
    if (it == m->lin.end()) {
        VocTcW w;
        w.w = const_cast<bf16*>(c.W); w.N = c.N; w.taps = c.taps; w.Cin = c.Cin; w.Cp = c.Cin; w.BN = voc_tc_pick_bn(c.N);
        if (w.BN > 128 && c.L <= 2048) w.BN = (c.N % 128 == 0) ? 128 : w.BN;        // few rows: more channel tiles = more CTAs
        if (!w.BN || make_map(h, &w.mw, w.w, c.N, c.taps * c.Cin, w.BN)) { h->err.clear(); return false; }
        it = m->lin.emplace(c.W, w).first;
    }
    VocTcW w = it->second;
    w.bias = c.bias; w.bias_mod = c.bias_mod > 0 ? c.bias_mod : c.N;
    const int hist = c.hist;                                       // streaming decode: rows in front of x carry the previous chunk's tail
    const size_t need = (size_t)(c.L + hist) * m->planes * c.Cin;
    if (m->xs_cap < need) {
        if (m->xs) cudaFree(m->xs);
        m->xs = nullptr; m->xs_cap = 0;
        if (cudaMalloc((void**)&m->xs, need * sizeof(bf16)) != cudaSuccess) { (void)cudaGetLastError(); return false; }
        m->xs_cap = need;
    }
    const long long n4 = (long long)(c.L + hist) * (c.Cin / 4);
    voc_planes_kernel<<<(int)std::min<long long>((n4 + 255) / 256, (long long)h->num_sms * 16), 256, 0, h->stream>>>(c.x - (size_t)hist * c.Cin, m->xs, c.L + hist, c.Cin, m->planes, c.Cin, nullptr, nullptr);
    h->stats.kernel_launches++;
    VocTcOut o; o.y = c.y; o.residual = c.residual; o.scale = c.scale; o.act = c.act;
    return voc_tc_conv(h, m, w, m->xs + (size_t)hist * m->planes * c.Cin, c.L, c.dil > 0 ? c.dil : 1, c.tap_rev != 0, o, hist) == 0;
}

// decoder of tokenizer12hz_decode from the output of the upsampling stages: cur fp32 [L0][Cv] -> audio [L0 * prod(rates)]
int voc_tc_decoder(lqt_engine* h, const float* cur, long long L0, float* audio, lqt_engine::VocStream* vs) {
    VocTcModel* m = h->voc_tc;
    const Spec& s = h->sp;
    const int P = m->planes;
    // front margin of the plane / fp32 buffers: 64 rows (>= the longest left context: k7 at dilation 9 = 54) of the widest row
    const size_t xmargin = (size_t)64 * P * std::max(s.voc_hidden, (s.voc_decoder_dim + 63) / 64 * 64);
    const size_t tmargin = (size_t)64 * s.voc_decoder_dim;
    // capacity: largest plane tensor and largest fp32 tensor over the chain
    size_t xmax = (size_t)L0 * P * std::max(s.voc_hidden, (s.voc_decoder_dim + 63) / 64 * 64), tmax = 0;
    {
        long long L = L0;
        for (const VocBlockW& B : h->vblk) {
            L *= B.stride;
            xmax = std::max(xmax, (size_t)L * P * ((B.cout + 63) / 64 * 64));
            tmax = std::max(tmax, (size_t)L * B.cout);
        }
    }
    xmax += xmargin; tmax += tmargin;
    if (m->x_cap < xmax) {
        if (m->xa) cudaFree(m->xa);
        if (m->xb) cudaFree(m->xb);
        m->xa = m->xb = nullptr; m->x_cap = 0;
        CK(cudaMalloc((void**)&m->xa, xmax * sizeof(bf16)));
        CK(cudaMalloc((void**)&m->xb, xmax * sizeof(bf16)));
        m->x_cap = xmax;
    }
    if (m->t_cap < tmax) {
        if (m->t0) cudaFree(m->t0);
        m->t0 = nullptr; m->t_cap = 0;
        CK(cudaMalloc((void**)&m->t0, tmax * sizeof(float)));
        m->t_cap = tmax;
    }
    // channel-padding columns (Cin = 96 -> Cp = 128) must read as zeros: the epilogues never write them
    CK(cudaMemsetAsync(m->xa, 0, xmax * sizeof(bf16), h->stream));
    CK(cudaMemsetAsync(m->xb, 0, xmax * sizeof(bf16), h->stream));
    bf16 *pin = m->xa + xmargin, *pout = m->xb + xmargin;
    float* t0buf = m->t0 + tmargin;
    auto row_bytes = [&](int C) { return (size_t)P * ((C + 63) / 64 * 64) * sizeof(bf16); };
    long long L = L0;
    {
        const long long n4 = L * (s.voc_hidden / 4);
        voc_planes_kernel<<<(int)std::min<long long>((n4 + 255) / 256, (long long)h->num_sms * 16), 256, 0, h->stream>>>(cur, pin, L, s.voc_hidden, P, s.voc_hidden, nullptr, nullptr);
        h->stats.kernel_launches++;
    }
    {   // dec.conv_in (k7) -> planes of snake_b0(.)
        VocTcOut o; o.xo = pout; o.cout = s.voc_decoder_dim; o.sn = h->vblk.empty() ? &m->s_out : &m->blk[0].s_in;
        if (voc_halo(h, vs, "tc_in", pin, row_bytes(s.voc_hidden), 6, L)) return 1;
        if (voc_tc_conv(h, m, m->conv_in, pin, L, 1, false, o, vs ? 6 : 0)) return 1;
        std::swap(pin, pout);
    }
    float* fa = t0buf;
    for (size_t b = 0; b < h->vblk.size(); ++b) {
        const VocBlockW& B = h->vblk[b];
        auto& K = m->blk[b];
        {   // transposed conv: [L][s*cout] == [L*s][cout]; fp32 residual stream t0 + planes of snake1_r0
            VocTcOut o; o.y = t0buf; o.xo = pout; o.cout = B.cout; o.up = B.stride; o.sn = &K.s1[0];
            if (voc_halo(h, vs, "tc_t" + std::to_string(b), pin, row_bytes(B.cin), 1, L)) return 1;
            if (voc_tc_conv(h, m, K.tconv, pin, L, 1, true, o, vs ? 1 : 0)) return 1;
            std::swap(pin, pout);
        }
        L *= B.stride;
        const int dil[3] = {1, 3, 9};
        for (int r = 0; r < 3; ++r) {
            {   // conv1 k7 dilated -> planes of snake2(.)
                VocTcOut o; o.xo = pout; o.cout = B.cout; o.sn = &K.s2[r];
                if (voc_halo(h, vs, "tc_c" + std::to_string(b) + "_" + std::to_string(r), pin, row_bytes(B.cout), 6 * dil[r], L)) return 1;
                if (voc_tc_conv(h, m, K.c1[r], pin, L, dil[r], false, o, vs ? 6 * dil[r] : 0)) return 1;
                std::swap(pin, pout);
            }
            {   // conv2 k1 + residual -> t0 (in place) and the planes of the next consumer's SnakeBeta
                VocTcOut o; o.residual = t0buf; o.cout = B.cout;
                const bool last = (r == 2 && b + 1 == h->vblk.size());
                if (last) { o.y = t0buf; o.y_snake = true; o.sn = &m->s_out; }           // fp32 snake_out(x) for conv_out_kernel
                else { o.y = t0buf; o.xo = pout; o.sn = (r < 2) ? &K.s1[r + 1] : &m->blk[b + 1].s_in; }
                if (voc_tc_conv(h, m, K.c2[r], pin, L, 1, false, o)) return 1;
                std::swap(pin, pout);
            }
        }
    }
    const int Cl = h->vblk.empty() ? s.voc_decoder_dim : h->vblk.back().cout;
    if (voc_halo(h, vs, "tc_out", fa, (size_t)Cl * 4, 6, L)) return 1;
    conv_out_kernel<<<(unsigned)((L + 7) / 8), 256, 0, h->stream>>>(fa, audio, L, Cl, h->out_w, h->out_b, vs ? 6 : 0);
    h->stats.kernel_launches++;
    CK(cudaGetLastError());
    return 0;
}
