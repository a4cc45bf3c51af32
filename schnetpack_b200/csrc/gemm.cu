// Dense layer kernel (fp32 FFMA path):  Y = act( (A .* act'(a_pre)) * B + bias ) + addend
// Reference: nn/base.py:52-55 (Dense.forward) and its autograd input-gradient.  True-fp32 accumulation like the
// reference's default torch matmul precision ("highest", cli.py:95-97), so 1e-5 parity holds by construction.
//
// Tiling: BM=128 x BN=64 output tile per 256-thread CTA, BK=16, 8x4 register micro-tile per thread, A tile stored
// k-major in shared memory so both operand fetches are conflict-free 128-bit LDS; next k-tile is prefetched into
// registers while the current one is multiplied.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int AS_LD = BM + 4;

struct GemmArgs {
    const float* A;
    const float* a_pre;
    const float* B;
    const float* bias;
    const float* addend;
    float* Y;
    float* y_pre;
    int64_t M, lda, ld_add, ldy;
    int K, N, a_act, act;
};

__global__ void __launch_bounds__(NT) k_dense(GemmArgs g) {
    __shared__ __align__(16) float As[BK][AS_LD];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // global->register staging
    float4 ra[2];
    float4 rb;
    const int a_row = tid >> 2, a_kq = (tid & 3) * 4;  // rows a_row and a_row+64, k offset a_kq
    const int b_k = tid >> 4, b_n = (tid & 15) * 4;
    const bool a_vec = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) &&
                       (!g.a_pre || (reinterpret_cast<uintptr_t>(g.a_pre) & 15) == 0);
    const bool b_vec = ((g.N & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            int64_t m = m0 + a_row + it * 64;
            int k = k0 + a_kq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < g.M) {
                const float* p = g.A + m * g.lda + k;
                if (a_vec && k + 3 < g.K) {
                    v = *reinterpret_cast<const float4*>(p);
                    if (g.a_pre) {
                        float4 q = *reinterpret_cast<const float4*>(g.a_pre + m * g.lda + k);
                        v.x *= spk_act_grad(q.x, g.a_act);
                        v.y *= spk_act_grad(q.y, g.a_act);
                        v.z *= spk_act_grad(q.z, g.a_act);
                        v.w *= spk_act_grad(q.w, g.a_act);
                    }
                } else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (k + i < g.K) {
                            t[i] = p[i];
                            if (g.a_pre) t[i] *= spk_act_grad(g.a_pre[m * g.lda + k + i], g.a_act);
                        }
                    v = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
            ra[it] = v;
        }
        {
            int k = k0 + b_k, n = n0 + b_n;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < g.K) {
                const float* p = g.B + (int64_t)k * g.N + n;
                if (b_vec && n + 3 < g.N) {
                    v = *reinterpret_cast<const float4*>(p);
                } else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (n + i < g.N) t[i] = p[i];
                    v = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
            rb = v;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            int r = a_row + it * 64;
            As[a_kq + 0][r] = ra[it].x;
            As[a_kq + 1][r] = ra[it].y;
            As[a_kq + 2][r] = ra[it].z;
            As[a_kq + 3][r] = ra[it].w;
        }
        *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = rb;
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nk = (g.K + BK - 1) / BK;
    load_tiles(0);
    for (int kt = 0; kt < nk; ++kt) {
        __syncthreads();
        store_tiles();
        __syncthreads();
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
    }

    // epilogue
    const int n = n0 + tx * 4;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (g.bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n + j < g.N) bv[j] = g.bias[n + j];
    }
    const bool y_vec = ((g.ldy & 3) == 0) && (n + 3 < g.N) && ((reinterpret_cast<uintptr_t>(g.Y) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int64_t m = m0 + ty * 8 + i;
        if (m >= g.M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bv[j];
        if (g.y_pre) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < g.N) g.y_pre[m * g.ldy + n + j] = v[j];
        }
        if (g.act != SPK_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = spk_act(v[j], g.act);
        }
        if (g.addend) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < g.N) v[j] += g.addend[m * g.ld_add + n + j];
        }
        if (y_vec) {
            *reinterpret_cast<float4*>(g.Y + m * g.ldy + n) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < g.N) g.Y[m * g.ldy + n + j] = v[j];
        }
    }
}

}  // namespace

extern "C" int spk_dense(const float* A, int64_t M, int K, int64_t lda, const float* a_pre, int a_act, const float* B,
                         int N, const float* bias, int act, const float* addend, int64_t ld_add, float* Y, int64_t ldy,
                         float* y_pre, spk_stream_t stream) {
    if (M < 0 || K <= 0 || N <= 0 || lda < K || ldy < N) return SPK_ERR_ARG;
    if (act < 0 || act > 2 || a_act < 0 || a_act > 2) return SPK_ERR_ARG;
    if (M == 0) return SPK_OK;
    if (!A || !B || !Y) return SPK_ERR_ARG;
    if (addend && ld_add < N) return SPK_ERR_ARG;
    GemmArgs g;
    g.A = A; g.a_pre = a_pre; g.B = B; g.bias = bias; g.addend = addend; g.Y = Y; g.y_pre = y_pre;
    g.M = M; g.lda = lda; g.ld_add = ld_add; g.ldy = ldy; g.K = K; g.N = N; g.a_act = a_act; g.act = act;
    dim3 grid((unsigned)spk_cdiv(M, BM), (unsigned)spk_cdiv(N, BN));
    k_dense<<<grid, NT, 0, spk_st(stream)>>>(g);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
