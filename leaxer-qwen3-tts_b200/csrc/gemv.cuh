// Weight-streaming GEMV for M <= 4 rows in flight:  y[m][n] = epilogue( sum_k W[n][k] * xin[m][k] )
// W is bf16 [N][K] row-major (nn.Linear layout), activations fp32, fp32 accumulate.
// Prologue (optional): RMSNorm of the input rows (talker/predictor layers, final norms).
// Epilogues: bias, SiLU/GELU, SwiGLU (two matrices), per-output scale, residual add.
//
// HBM-bound by construction (1 flop/byte): one warp owns 4 weight rows at a time and keeps
// 16 x 16-byte streaming loads per lane in flight; x lives in shared memory in a lane-permuted
// order so the two float4 reads per lane are bank-conflict free.
#pragma once
#include "common.cuh"

namespace lqt {

struct GemvParams {
    const __nv_bfloat16* W;     // [N][K]
    const __nv_bfloat16* W2;    // GLU only: second matrix [N][K]; y = silu(W x) * (W2 x)
    const float* x;             // [M][x_stride]
    const float* norm_w;        // nullable: RMSNorm weight [K]
    const float* bias;          // nullable [N]
    const float* scale;         // nullable [N]: multiplies the result before the residual add
    const float* residual;      // nullable [M][res_stride]
    float* y;                   // [M][y_stride]
    float* xnorm_out;           // nullable: block (0,*) writes the normalised input rows [M][K]
    const int* done;            // nullable early-exit flag (device frame loop)
    int x_stride, res_stride, y_stride;
    int M, N, K;
    int act;                    // 0 none, 1 SiLU, 3 GELU(erf)  (applied after bias)
    float eps;
};

constexpr int GEMV_THREADS = 256;
constexpr int GEMV_WARPS = GEMV_THREADS / 32;

LQT_DEVINL int gemv_kpad(int K) { return (K + 255) & ~255; }

template <int MT, bool GLU>
__global__ void __launch_bounds__(GEMV_THREADS)
gemv_kernel(const GemvParams p) {
    extern __shared__ float4 xs4[];                 // [MT][Kpad/4], lane-permuted
    __shared__ float red[GEMV_WARPS][MT];
    __shared__ float rstd_s[MT];
    if (p.done && *p.done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = p.K, K4 = K >> 2, Kpad4 = gemv_kpad(K) >> 2;
    const int m_base = blockIdx.y * MT;

    // ---- stage the input rows (and their sum of squares) -------------------------------------
    float ss[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) ss[m] = 0.f;
    for (int k4 = tid; k4 < Kpad4; k4 += GEMV_THREADS) {
        const int dst = (k4 & ~63) + ((k4 & 1) << 5) + ((k4 & 63) >> 1);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k4 < K4 && m_base + m < p.M)
                v = reinterpret_cast<const float4*>(p.x + (size_t)(m_base + m) * p.x_stride)[k4];
            ss[m] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            xs4[m * Kpad4 + dst] = v;
        }
    }
    if (p.norm_w) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            float s = warp_sum(ss[m]);
            if (lane == 0) red[warp][m] = s;
        }
        __syncthreads();
        if (tid < MT) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < GEMV_WARPS; ++w) s += red[w][tid];
            rstd_s[tid] = 1.0f / sqrtf(s / (float)K + p.eps);
        }
        __syncthreads();
        for (int k4 = tid; k4 < K4; k4 += GEMV_THREADS) {
            const int dst = (k4 & ~63) + ((k4 & 1) << 5) + ((k4 & 63) >> 1);
            const float4 w = reinterpret_cast<const float4*>(p.norm_w)[k4];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                float4 v = xs4[m * Kpad4 + dst];
                const float r = rstd_s[m];
                v.x = (v.x * r) * w.x; v.y = (v.y * r) * w.y;
                v.z = (v.z * r) * w.z; v.w = (v.w * r) * w.w;
                xs4[m * Kpad4 + dst] = v;
                if (p.xnorm_out && blockIdx.x == 0 && m_base + m < p.M)
                    reinterpret_cast<float4*>(p.xnorm_out + (size_t)(m_base + m) * K)[k4] = v;
            }
        }
    }
    __syncthreads();

    // ---- stream the weight rows --------------------------------------------------------------
    constexpr int OPI = GLU ? 2 : 4;                 // outputs per warp iteration (4 weight rows)
    const int total_warps = gridDim.x * GEMV_WARPS;
    const int N = p.N;
    for (int n0 = (blockIdx.x * GEMV_WARPS + warp) * OPI; n0 < N; n0 += total_warps * OPI) {
        const __nv_bfloat16* rowp[4];
        if (GLU) {
            const int n1 = min(n0 + 1, N - 1);
            rowp[0] = p.W + (size_t)n0 * K;  rowp[1] = p.W2 + (size_t)n0 * K;
            rowp[2] = p.W + (size_t)n1 * K;  rowp[3] = p.W2 + (size_t)n1 * K;
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) rowp[r] = p.W + (size_t)min(n0 + r, N - 1) * K;
        }
        float acc[4][MT];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int m = 0; m < MT; ++m) acc[r][m] = 0.f;

        for (int kb = lane * 8; kb < K; kb += 1024) {
            uint4 w[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = kb + j * 256;
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    w[r][j] = (k < K) ? ldg_stream(rowp[r] + k) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = kb + j * 256;
                if (k < K) {
                    const int c4 = (k >> 8) * 64 + lane;          // chunk base + lane
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        const float4 a = xs4[m * Kpad4 + c4];
                        const float4 b = xs4[m * Kpad4 + c4 + 32];
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[r][m] = dot8(w[r][j], a, b, acc[r][m]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int m = 0; m < MT; ++m) acc[r][m] = warp_sum(acc[r][m]);

        // ---- epilogue: lane (o*MT+m) finishes output (n0+o, m) -------------------------------
#pragma unroll
        for (int o = 0; o < OPI; ++o) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                if (lane == o * MT + m) {
                    const int n = n0 + o, mm = m_base + m;
                    if (n < N && mm < p.M) {
                        float v;
                        if (GLU) {
                            v = silu_f(acc[2 * o][m]) * acc[2 * o + 1][m];
                        } else {
                            v = acc[o][m];
                            if (p.bias) v += p.bias[n];
                            if (p.act == 1) v = silu_f(v);
                            else if (p.act == 3) v = gelu_erf_f(v);
                        }
                        if (p.scale) v *= p.scale[n];
                        if (p.residual) v += p.residual[(size_t)mm * p.res_stride + n];
                        p.y[(size_t)mm * p.y_stride + n] = v;
                    }
                }
            }
        }
    }
}

}  // namespace lqt
