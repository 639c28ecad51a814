// TEST INFRASTRUCTURE -- deterministic stub graphs behind the ORT shim (see onnxruntime_cxx_api.h).
// Every stub honours the I/O contract of the graph it stands for (tensor names, dtypes, shapes: SURVEY.md Appendix A,
// src/tts_onnx.cpp:545-776) and derives its outputs from a hash of its inputs, in exact uint32/float32 arithmetic that
// oracle/stub_graphs.py repeats in numpy. One trace line per Run call: "<graph> <name>:<shape>:<sum>:<xor> ...".
// Include once with ORT_SHIM_IMPLEMENTATION defined.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <sstream>

#include "onnxruntime_cxx_api.h"

namespace ort_shim {

struct Digest { uint32_t s = 0, x = 0; };
inline uint32_t mix32(uint32_t x) {          // lowbias32 (same as modelspec._mix32)
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16; return x;
}
inline float u2f(uint32_t h) { return (float)(h >> 8) * 1.1920928955078125e-07f - 1.0f; }   // [-1, 1), exact
inline Digest digest_words(const uint32_t* w, size_t n, uint32_t salt) {
    Digest d;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t h = mix32(w[i] ^ mix32((uint32_t)i * 0x9E3779B9u + salt));
        d.s += h; d.x ^= h;
    }
    return d;
}
inline uint32_t key_of(Digest d) { return mix32(d.s ^ mix32(d.x + 0x85EBCA6Bu)); }
inline Digest digest_value(const Ort::Value& v, uint32_t salt = 0) {
    const size_t words = v.count * (v.kind == Ort::Value::I64 ? 2 : 1);
    return digest_words(reinterpret_cast<const uint32_t*>(v.ptr), words, salt);
}

struct State {
    std::string trace;
    long calls = 0;
    long eos_at = -1;              // talker_decode: attention_mask length at which CODEC_EOS gets the top logit
    int hidden = 1024, vocab = 3072, cp_vocab = 2048, layers = 28, kv_heads = 8, head_dim = 128, spf = 1920;
};
State& state();
std::vector<Ort::Value> run_stub(const std::string& stem, const char* const* in_names, const Ort::Value* in, size_t n_in,
                                 const char* const* out_names, size_t n_out);

#ifdef ORT_SHIM_IMPLEMENTATION
State& state() { static State s; return s; }

static void fill_hash(float* dst, size_t n, uint32_t base, float scale) {
    for (size_t j = 0; j < n; ++j) dst[j] = scale * u2f(mix32(base + (uint32_t)j));
}
// causal chain of row keys: ck[p] depends on rows 0..p
static std::vector<uint32_t> row_chain(const float* rows, size_t n_rows, size_t width, uint32_t salt) {
    std::vector<uint32_t> ck(n_rows);
    uint32_t prev = salt;
    for (size_t p = 0; p < n_rows; ++p) {
        const uint32_t rk = key_of(digest_words(reinterpret_cast<const uint32_t*>(rows + p * width), width, (uint32_t)p));
        prev = mix32(prev ^ rk);
        ck[p] = prev;
    }
    return ck;
}

std::vector<Ort::Value> run_stub(const std::string& stem, const char* const* in_names, const Ort::Value* in, size_t n_in,
                                 const char* const* out_names, size_t n_out) {
    State& S = state();
    const int H = S.hidden;
    {   // trace line
        std::ostringstream os;
        os << stem;
        for (size_t i = 0; i < n_in; ++i) {
            const Digest d = digest_value(in[i]);
            os << ' ' << in_names[i] << ':';
            for (size_t k = 0; k < in[i].shape.size(); ++k) os << (k ? "x" : "") << in[i].shape[k];
            char buf[32]; std::snprintf(buf, sizeof buf, ":%08x:%08x", d.s, d.x);
            os << buf;
        }
        os << " ->";
        for (size_t i = 0; i < n_out; ++i) os << ' ' << out_names[i];
        S.trace += os.str(); S.trace += '\n';
        ++S.calls;
    }
    std::vector<Ort::Value> out;
    using V = Ort::Value;
    if (stem == "text_project" || stem == "codec_embed") {
        const int64_t* ids = reinterpret_cast<const int64_t*>(in[0].ptr);
        const int64_t n = (int64_t)in[0].count;
        V o = V::Owned(V::F32, {1, n, H});
        const uint32_t salt = stem == "text_project" ? 0x1111u : 0x2222u;
        for (int64_t s = 0; s < n; ++s) fill_hash(o.GetTensorMutableData<float>() + s * H, H, mix32((uint32_t)ids[s] ^ salt), 1.0f);
        out.push_back(std::move(o));
    } else if (stem == "code_predictor_embed") {
        const uint32_t id = (uint32_t) * reinterpret_cast<const int64_t*>(in[0].ptr), step = (uint32_t) * reinterpret_cast<const int64_t*>(in[1].ptr);
        V o = V::Owned(V::F32, {1, 1, H});
        fill_hash(o.GetTensorMutableData<float>(), H, mix32(mix32(id ^ 0x3333u) ^ (step * 0x9E3779B9u)), 1.0f);
        out.push_back(std::move(o));
    } else if (stem == "talker_prefill" || stem == "talker_decode") {
        const bool dec = stem == "talker_decode";
        const int64_t n_new = dec ? 1 : in[0].shape[1];
        const int64_t T = (int64_t)in[1].count;                       // attention_mask length = total positions
        const int64_t past = T - n_new;
        uint32_t salt = 0x4444u;
        if (dec) {                                                     // the past KV tensors feed the chain
            uint32_t pk = 0x9999u;
            for (size_t i = 2; i < n_in; ++i) pk = mix32(pk ^ key_of(digest_value(in[i], (uint32_t)i)));
            salt = pk;
        }
        std::vector<uint32_t> ck;
        {
            const float* rows = reinterpret_cast<const float*>(in[0].ptr);
            ck.resize((size_t)n_new);
            uint32_t prev = salt;
            for (int64_t p = 0; p < n_new; ++p) {
                const uint32_t rk = key_of(digest_words(reinterpret_cast<const uint32_t*>(rows + p * H), (size_t)H, (uint32_t)(past + p)));
                prev = mix32(prev ^ rk);
                ck[(size_t)p] = prev;
            }
        }
        V lg = V::Owned(V::F32, {1, n_new, S.vocab});
        for (int64_t p = 0; p < n_new; ++p) fill_hash(lg.GetTensorMutableData<float>() + p * S.vocab, (size_t)S.vocab, ck[(size_t)p], 4.0f);
        if (dec && T == S.eos_at) lg.GetTensorMutableData<float>()[2150] = 100.0f;
        V lh = V::Owned(V::F32, {1, 1, H});
        fill_hash(lh.GetTensorMutableData<float>(), (size_t)H, ck.back() ^ 0x55555555u, 1.0f);
        out.push_back(std::move(lg)); out.push_back(std::move(lh));
        const int KH = S.kv_heads, D = S.head_dim;
        for (int i = 0; i < 2 * S.layers; ++i) {                       // present_key_0, present_value_0, present_key_1, ...
            V kv = V::Owned(V::F32, {1, KH, T, D});
            float* dst = kv.GetTensorMutableData<float>();
            const float* src = dec ? reinterpret_cast<const float*>(in[2 + i].ptr) : nullptr;
            for (int h = 0; h < KH; ++h) {
                if (past) std::memcpy(dst + (size_t)h * T * D, src + (size_t)h * past * D, (size_t)past * D * sizeof(float));
                for (int64_t p = 0; p < n_new; ++p)
                    fill_hash(dst + ((size_t)h * T + past + p) * D, (size_t)D, ck[(size_t)p] + 0x10000u * (uint32_t)i + 128u * (uint32_t)h + 0x777u, 1.0f);
            }
            out.push_back(std::move(kv));
        }
    } else if (stem == "code_predictor") {
        const int64_t L = in[0].shape[1];
        const uint32_t step = (uint32_t) * reinterpret_cast<const int64_t*>(in[1].ptr);
        const std::vector<uint32_t> ck = row_chain(reinterpret_cast<const float*>(in[0].ptr), (size_t)L, (size_t)H, 0x6666u);
        V o = V::Owned(V::F32, {1, 1, S.cp_vocab});
        fill_hash(o.GetTensorMutableData<float>(), (size_t)S.cp_vocab, ck.back() ^ (step * 0x9E3779B9u), 4.0f);
        out.push_back(std::move(o));
    } else if (stem == "tokenizer12hz_decode") {
        const int64_t T = in[0].shape[1];
        V a = V::Owned(V::F32, {1, T * S.spf});
        fill_hash(a.GetTensorMutableData<float>(), (size_t)(T * S.spf), key_of(digest_value(in[0])) ^ 0x8888u, 0.5f);
        V len = V::Owned(V::I64, {1});
        len.GetTensorMutableData<int64_t>()[0] = T * S.spf;
        out.push_back(std::move(a)); out.push_back(std::move(len));
    } else if (stem == "speaker_encoder") {
        V o = V::Owned(V::F32, {1, H});
        fill_hash(o.GetTensorMutableData<float>(), (size_t)H, key_of(digest_value(in[0])) ^ 0xAAAAu, 1.0f);
        out.push_back(std::move(o));
    } else {
        throw Ort::Exception("ort_shim: unknown graph " + stem);
    }
    return out;
}
#endif  // ORT_SHIM_IMPLEMENTATION

}  // namespace ort_shim
