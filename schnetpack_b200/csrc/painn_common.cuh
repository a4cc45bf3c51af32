// Shared device helpers of the PaiNN edge kernels (painn.cu: LDG/LDS variant, painn_tma.cu: TMA + mbarrier pipeline).
#pragma once
#include "common.cuh"

// Filter weights of one channel as k-PAIRS (w[2m], w[2m+1]): with the radial basis read from shared memory as natural
// (phi[2m], phi[2m+1]) pairs, one packed FFMA2 (fma.rn.f32x2 -- Blackwell's full-rate fp32 path; scalar FFMA issues at
// half rate) advances an even-k and an odd-k partial sum at once; the two partials are added at the end.
template <int NRB>
struct FilterRegs {
    float2 a[NRB / 2], b[NRB / 2], c[NRB / 2];
    float ba, bb, bc;
};

template <int NRB>
__device__ __forceinline__ void load_filter(FilterRegs<NRB>& w, const float* __restrict__ wf,
                                            const float* __restrict__ bf, int F, int n_rbf, int c) {
#pragma unroll
    for (int m = 0; m < NRB / 2; ++m) {
        const int k0 = 2 * m, k1 = 2 * m + 1;
        const bool ok0 = k0 < n_rbf, ok1 = k1 < n_rbf;
        w.a[m] = make_float2(ok0 ? wf[(int64_t)c * n_rbf + k0] : 0.f, ok1 ? wf[(int64_t)c * n_rbf + k1] : 0.f);
        w.b[m] = make_float2(ok0 ? wf[(int64_t)(F + c) * n_rbf + k0] : 0.f, ok1 ? wf[(int64_t)(F + c) * n_rbf + k1] : 0.f);
        w.c[m] = make_float2(ok0 ? wf[(int64_t)(2 * F + c) * n_rbf + k0] : 0.f,
                             ok1 ? wf[(int64_t)(2 * F + c) * n_rbf + k1] : 0.f);
    }
    w.ba = bf[c];
    w.bb = bf[F + c];
    w.bc = bf[2 * F + c];
}

