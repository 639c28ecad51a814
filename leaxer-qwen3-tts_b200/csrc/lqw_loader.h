// .lqw weight-file loader (format: leaxer-qwen3-tts_b200/modelspec.py). Replaces
// TTSEngine::load_model (src/tts_onnx.cpp:134-232): one file per former .onnx graph, uploaded to
// HBM once; tensors are addressed by name.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
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

// Header of one tensor as stored in the file (no device state): parsed and validated without touching CUDA so that corrupt or
// hostile files are rejected before any allocation (tests: lqt_check_model_file, CPU only).
struct LqwEntry { std::string name; int dtype = 0; std::vector<int64_t> dims; uint64_t off = 0, nbytes = 0; };
struct LqwHeader {
    std::map<std::string, std::string> meta;
    std::vector<LqwEntry> entries;
    uint64_t data_start = 0, data_bytes = 0;       // data section: [data_start, data_start + data_bytes) must lie inside the file
};

// Parses and validates the header of an open .lqw file. Every read is bounds-checked against the header block, every tensor's
// byte count must equal numel x element size, offsets must be 256-byte aligned and the data section must fit in `file_size`.
// returns empty string on success, else an error message
inline std::string parse_lqw_header(FILE* f, uint64_t file_size, const std::string& path, LqwHeader& out) {
    unsigned char head[24];
    if (std::fread(head, 1, 24, f) != 24 || std::memcmp(head, "LQTW0001", 8) != 0) return "bad magic in " + path;
    uint32_t nt, nm; uint64_t data_start;
    std::memcpy(&nt, head + 8, 4); std::memcpy(&nm, head + 12, 4); std::memcpy(&data_start, head + 16, 8);
    if (data_start < 24 || data_start > (1u << 26) || data_start > file_size) return "bad header in " + path;
    if (nt > (1u << 20) || nm > (1u << 16)) return "bad header in " + path + " (tensor/meta count)";
    std::vector<unsigned char> hdr(data_start - 24);
    if (!hdr.empty() && std::fread(hdr.data(), 1, hdr.size(), f) != hdr.size()) return "short header in " + path;
    size_t p = 0;
    bool ok = true;
    auto have = [&](size_t n) { if (!ok || n > hdr.size() - p) { ok = false; return false; } return true; };
    auto rd = [&](void* dst, size_t n) { if (!have(n)) { std::memset(dst, 0, n); return; } std::memcpy(dst, hdr.data() + p, n); p += n; };
    auto rdstr = [&](std::string& s) { uint16_t n = 0; rd(&n, 2); if (!have(n)) { s.clear(); return; } s.assign((const char*)hdr.data() + p, n); p += n; };
    for (uint32_t i = 0; i < nm && ok; ++i) { std::string k, v; rdstr(k); rdstr(v); if (ok) out.meta[k] = v; }
    if (!ok) return "corrupt header in " + path + " (meta section runs past the header)";
    out.entries.resize(nt);
    uint64_t total = 0;
    for (uint32_t i = 0; i < nt; ++i) {
        LqwEntry& e = out.entries[i];
        rdstr(e.name);
        uint8_t dt = 0, nd = 0; rd(&dt, 1); rd(&nd, 1);
        if (!ok) return "corrupt header in " + path + " (tensor table runs past the header)";
        if (dt > 1) return "corrupt header in " + path + ": unknown dtype for tensor " + e.name;
        if (nd > 8) return "corrupt header in " + path + ": rank > 8 for tensor " + e.name;
        e.dtype = dt;
        uint64_t numel = 1;
        for (int d = 0; d < nd; ++d) {
            uint32_t v = 0; rd(&v, 4);
            e.dims.push_back(v);
            if (v != 0 && numel > (UINT64_MAX / 8) / v) return "corrupt header in " + path + ": dims overflow for tensor " + e.name;
            numel *= v;
        }
        rd(&e.off, 8); rd(&e.nbytes, 8);
        if (!ok) return "corrupt header in " + path + " (tensor table runs past the header)";
        if (e.nbytes != numel * (dt == 0 ? 2u : 4u)) return "corrupt header in " + path + ": byte count of tensor " + e.name + " does not match its dims";
        if (e.off % 256 != 0) return "corrupt header in " + path + ": misaligned tensor " + e.name;
        if (e.off > UINT64_MAX - e.nbytes || e.off + e.nbytes > file_size - data_start)
            return "truncated file " + path + " (tensor " + e.name + " lies outside the data section)";
        if (e.off + e.nbytes > total) total = e.off + e.nbytes;
    }
    out.data_start = data_start; out.data_bytes = total;
    return "";
}

inline uint64_t file_size_of(FILE* f) {
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    return n < 0 ? 0 : (uint64_t)n;
}

// header-only validation of a file (no CUDA): empty string = well formed
inline std::string check_lqw(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return "cannot open " + path;
    LqwHeader h;
    const std::string er = parse_lqw_header(f, file_size_of(f), path, h);
    std::fclose(f);
    return er;
}

// returns empty string on success, else an error message
inline std::string load_lqw(const std::string& path, LqwFile& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return "cannot open " + path;
    LqwHeader hd;
    const std::string er = parse_lqw_header(f, file_size_of(f), path, hd);
    if (!er.empty()) { std::fclose(f); return er; }
    out.meta = hd.meta;
    if (hd.data_bytes > 0) {
        const uint64_t total = hd.data_bytes;
        if (cudaMalloc(&out.slab, (total + 255) & ~(uint64_t)255) != cudaSuccess) { std::fclose(f); return "cudaMalloc failed for " + path; }
        out.slab_bytes = total;
        // stream the data section through a pinned bounce buffer
        const size_t CH = 64u << 20;
        void* bounce = nullptr;
        if (cudaMallocHost(&bounce, CH) != cudaSuccess) { std::fclose(f); return "cudaMallocHost failed"; }
        std::fseek(f, (long)hd.data_start, SEEK_SET);
        uint64_t done = 0;
        while (done < total) {
            size_t want = (size_t)std::min<uint64_t>(CH, total - done);
            size_t got = std::fread(bounce, 1, want, f);
            if (got == 0) break;
            if (cudaMemcpy((char*)out.slab + done, bounce, got, cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaFreeHost(bounce); std::fclose(f);
                return "cudaMemcpy failed for " + path;
            }
            done += got;
        }
        cudaFreeHost(bounce);
        if (done < total) { std::fclose(f); return "truncated file " + path; }
        for (const LqwEntry& e : hd.entries) {
            DevTensor t;
            t.ptr = (char*)out.slab + e.off; t.dtype = e.dtype; t.dims = e.dims; t.nbytes = (size_t)e.nbytes;
            out.tensors[e.name] = t;
        }
    }
    std::fclose(f);
    return "";
}

}  // namespace lqt
