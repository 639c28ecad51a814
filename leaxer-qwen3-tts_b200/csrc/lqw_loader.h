// .lqw weight-file loader (format: leaxer-qwen3-tts_b200/modelspec.py). Replaces
// TTSEngine::load_model (src/tts_onnx.cpp:134-232): one file per former .onnx graph, uploaded to
// HBM once; tensors are addressed by name.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace lqt {

struct DevTensor {
    void* ptr = nullptr;
    int dtype = 0;                 // 0 = bf16, 1 = f32
    std::vector<int64_t> dims;
    size_t nbytes = 0;
    int64_t numel() const { int64_t n = 1; for (auto d : dims) n *= d; return n; }
};

struct LqwFile {
    std::map<std::string, std::string> meta;
    std::map<std::string, DevTensor> tensors;
    void* slab = nullptr;          // one device allocation per file
    size_t slab_bytes = 0;

    void release() { if (slab) cudaFree(slab); slab = nullptr; tensors.clear(); }

    const DevTensor* find(const std::string& name) const {
        auto it = tensors.find(name);
        return it == tensors.end() ? nullptr : &it->second;
    }
    int meta_int(const char* k, int dflt) const {
        auto it = meta.find(k);
        return it == meta.end() ? dflt : std::atoi(it->second.c_str());
    }
    double meta_f(const char* k, double dflt) const {
        auto it = meta.find(k);
        return it == meta.end() ? dflt : std::atof(it->second.c_str());
    }
    std::vector<int> meta_ints(const char* k) const {
        std::vector<int> out;
        auto it = meta.find(k);
        if (it == meta.end()) return out;
        const std::string& s = it->second;
        size_t p = 0;
        while (p < s.size()) {
            size_t q = s.find(',', p);
            if (q == std::string::npos) q = s.size();
            if (q > p) out.push_back(std::atoi(s.substr(p, q - p).c_str()));
            p = q + 1;
        }
        return out;
    }
};

// returns empty string on success, else an error message
inline std::string load_lqw(const std::string& path, LqwFile& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return "cannot open " + path;
    unsigned char head[24];
    if (std::fread(head, 1, 24, f) != 24 || std::memcmp(head, "LQTW0001", 8) != 0) {
        std::fclose(f);
        return "bad magic in " + path;
    }
    uint32_t nt, nm; uint64_t data_start;
    std::memcpy(&nt, head + 8, 4); std::memcpy(&nm, head + 12, 4); std::memcpy(&data_start, head + 16, 8);
    if (data_start < 24 || data_start > (1u << 26)) { std::fclose(f); return "bad header in " + path; }
    std::vector<unsigned char> hdr(data_start - 24);
    if (!hdr.empty() && std::fread(hdr.data(), 1, hdr.size(), f) != hdr.size()) {
        std::fclose(f);
        return "short header in " + path;
    }
    size_t p = 0;
    auto rd16 = [&](uint16_t& v) { std::memcpy(&v, hdr.data() + p, 2); p += 2; };
    auto rdstr = [&](std::string& s) { uint16_t n; rd16(n); s.assign((const char*)hdr.data() + p, n); p += n; };
    for (uint32_t i = 0; i < nm; ++i) { std::string k, v; rdstr(k); rdstr(v); out.meta[k] = v; }
    struct Ent { std::string name; DevTensor t; uint64_t off; };
    std::vector<Ent> ents(nt);
    uint64_t total = 0;
    for (uint32_t i = 0; i < nt; ++i) {
        Ent& e = ents[i];
        rdstr(e.name);
        uint8_t dt = hdr[p++], nd = hdr[p++];
        e.t.dtype = dt;
        for (int d = 0; d < nd; ++d) { uint32_t v; std::memcpy(&v, hdr.data() + p, 4); p += 4; e.t.dims.push_back(v); }
        uint64_t nb; std::memcpy(&e.off, hdr.data() + p, 8); p += 8; std::memcpy(&nb, hdr.data() + p, 8); p += 8;
        e.t.nbytes = nb;
        if (e.off + nb > total) total = e.off + nb;
    }
    if (total > 0) {
        total = (total + 255) & ~(uint64_t)255;
        if (cudaMalloc(&out.slab, total) != cudaSuccess) { std::fclose(f); return "cudaMalloc failed for " + path; }
        out.slab_bytes = total;
        // stream the data section through a pinned bounce buffer
        const size_t CH = 64u << 20;
        void* bounce = nullptr;
        if (cudaMallocHost(&bounce, CH) != cudaSuccess) { std::fclose(f); return "cudaMallocHost failed"; }
        std::fseek(f, (long)data_start, SEEK_SET);
        uint64_t done = 0;
        while (done < total) {
            size_t want = (size_t)std::min<uint64_t>(CH, total - done);
            size_t got = std::fread(bounce, 1, want, f);
            if (got == 0) break;        // trailing pad may be absent
            if (cudaMemcpy((char*)out.slab + done, bounce, got, cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaFreeHost(bounce); std::fclose(f);
                return "cudaMemcpy failed for " + path;
            }
            done += got;
        }
        cudaFreeHost(bounce);
        for (auto& e : ents) {
            if (e.off + e.t.nbytes > done) { std::fclose(f); return "truncated file " + path; }
            e.t.ptr = (char*)out.slab + e.off;
            out.tensors[e.name] = e.t;
        }
    }
    std::fclose(f);
    return "";
}

}  // namespace lqt
