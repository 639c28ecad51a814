// Shared device helpers for the sm_100a kernels. No fast-math: IEEE div/sqrt so that the oracle
// (fp32 torch) and the device differ only by summation order.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define LQT_DEVINL __device__ __forceinline__

namespace lqt {

// Device-side utterance state: everything that changes between graph replays lives here so the
// frame graph can be captured once.
struct GenState {
    int pos;            // talker KV length == position of the next talker token
    int frame;          // current frame index (Philox counter, trailing-text schedule)
    int done;           // set when CODEC_EOS was sampled (src/tts_onnx.cpp:812) or max frames hit
    int n_frames;       // frames emitted so far
    int trailing_len;   // src/tts_onnx.cpp:536
    int max_frames;     // params.max_new_tokens
    int cp_pos;         // code predictor position inside the current frame (0..16)
    int n_forced;       // teacher forcing: number of forced frames (0 = off)
};

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its predecessor in the stream / graph is still running. pdl_trigger() lets the successor's CTAs be scheduled as
// soon as every CTA of this grid has issued it; pdl_wait() blocks until the predecessor grid has completed and its memory
// is visible -- everything before it (barrier init, TMEM allocation, descriptor prefetch, index arithmetic) overlaps the
// predecessor's tail. Both are no-ops in a launch without the attribute.
LQT_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
LQT_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

LQT_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
LQT_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 16-byte load (weights are read once per pass: do not pollute L1)
LQT_DEVINL uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
LQT_DEVINL uint2 ldg_stream8(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// bf16 pair (packed in a u32, low half = element 0) -> two floats (exact: bf16 is truncated f32)
LQT_DEVINL float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
LQT_DEVINL float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

LQT_DEVINL float dot8(const uint4& w, const float4& a, const float4& b, float acc) {
    acc = fmaf(bf16lo(w.x), a.x, acc); acc = fmaf(bf16hi(w.x), a.y, acc);
    acc = fmaf(bf16lo(w.y), a.z, acc); acc = fmaf(bf16hi(w.y), a.w, acc);
    acc = fmaf(bf16lo(w.z), b.x, acc); acc = fmaf(bf16hi(w.z), b.y, acc);
    acc = fmaf(bf16lo(w.w), b.z, acc); acc = fmaf(bf16hi(w.w), b.w, acc);
    return acc;
}

LQT_DEVINL float silu_f(float x) { return x / (1.0f + expf(-x)); }
LQT_DEVINL float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// round-to-nearest-even f32 -> bf16 -> f32 (the KV-cache rounding point, mirrored by the oracle)
LQT_DEVINL float bf16_round_f(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <typename T> struct KvIO;
template <> struct KvIO<__nv_bfloat16> {
    static LQT_DEVINL float4 load4(const __nv_bfloat16* p) {
        uint2 u = *reinterpret_cast<const uint2*>(p);
        return make_float4(bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y));
    }
    static LQT_DEVINL void store4(__nv_bfloat16* p, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = u;
    }
    static LQT_DEVINL float round(float x) { return bf16_round_f(x); }
};
template <> struct KvIO<float> {
    static LQT_DEVINL float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
    static LQT_DEVINL void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
    static LQT_DEVINL float round(float x) { return x; }
};

}  // namespace lqt
