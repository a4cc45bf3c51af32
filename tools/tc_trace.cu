// Stand-alone micro-benchmark + phase trace of the tcgen05 dense kernel (not part of the product library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DSPK_TC_TRACE [-DTC_NST=4 -DTC_MINB=2] \
//        -I include -I schnetpack_b200/csrc tools/tc_trace.cu -o gpurun_out/tc_trace
// Prints per-shape launch times (CUDA events, 50 launches) and the clock64 phase stamps of a few CTAs.
#include "../schnetpack_b200/csrc/gemm_tc.cu"

#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

static void run(int64_t M, int K, int N, int act, bool prologue) {
    float *A, *Apre, *W, *Wp, *bias, *Y, *Ypre;
    long long* dbg;
    CK(cudaMalloc(&A, M * K * 4));
    CK(cudaMalloc(&Apre, M * K * 4));
    CK(cudaMalloc(&W, (size_t)N * K * 4));
    CK(cudaMalloc(&Wp, spk_tc_packed_floats(N, K) * 4));
    CK(cudaMalloc(&bias, N * 4));
    CK(cudaMalloc(&Y, M * N * 4));
    CK(cudaMalloc(&Ypre, M * N * 4));
    const int TN = tile_n(M, N);
    const int ncta = (int)(((M + TM - 1) / TM) * ((N + TN - 1) / TN));
    CK(cudaMalloc(&dbg, (size_t)ncta * 128 * 8));
    CK(cudaMemset(dbg, 0, (size_t)ncta * 128 * 8));
    std::vector<float> h((size_t)M * K);
    for (auto& v : h) v = (float)rand() / RAND_MAX - 0.5f;
    CK(cudaMemcpy(A, h.data(), M * K * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(Apre, h.data(), M * K * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(W, h.data(), (size_t)N * K * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(bias, h.data(), N * 4, cudaMemcpyHostToDevice));
    if (spk_tc_pack_weight(W, N, K, Wp, nullptr)) { printf("pack failed\n"); exit(1); }
    TcArgs g;
    g.A = A; g.a_pre = prologue ? Apre : nullptr; g.Wp = TN == 128 ? Wp + packed_floats_tn(N, K, 64) : Wp; g.bias = bias; g.addend = nullptr; g.Y = Y;
    g.y_pre = act ? Ypre : nullptr;
    g.M = M; g.lda = K; g.ld_add = N; g.ldy = N; g.K = K; g.N = N; g.a_act = prologue ? 3 : 0; g.act = act;
    g.save_deriv = act ? 1 : 0;
    g.dbg = dbg;
    auto launch = [&]() {
        int rc = TN == 128 ? dispatch_tc<128>(g, 0) : dispatch_tc<64>(g, 0);
        if (rc) { printf("launch rc %d\n", rc); exit(1); }
    };
    for (int i = 0; i < 5; ++i) launch();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 50;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int dev_clk = 0; cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
    printf("M=%lld K=%d N=%d act=%d prologue=%d ctas=%d : %.2f us/launch (TN=%d)\n", (long long)M, K, N, act,
           (int)prologue, ncta, ms * 1000.f / reps, TN);
    if (!act && !prologue) {   // sampled correctness check against fp64 on the host: Y = A W^T + bias
        std::vector<float> y((size_t)M * N);
        CK(cudaMemcpy(y.data(), Y, y.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxref = 0;
        for (int t = 0; t < 4096; ++t) {
            const int64_t m = (t * 7919LL) % M;
            const int n = (t * 104729) % N;
            double acc = h[n];
            for (int k = 0; k < K; ++k) acc += (double)h[m * K + k] * (double)h[(size_t)n * K + k];
            maxerr = fmax(maxerr, fabs(acc - y[m * N + n]));
            maxref = fmax(maxref, fabs(acc));
        }
        printf("  check: max abs err %.3e (max |ref| %.3f) -> rel %.2e\n", maxerr, maxref, maxerr / maxref);
    }
    std::vector<long long> d((size_t)ncta * 128);
    CK(cudaMemcpy(d.data(), dbg, d.size() * 8, cudaMemcpyDeviceToHost));
    const int nk = (K + TK - 1) / TK;
    int show[3] = {0, ncta / 2, ncta - 1};
    for (int si = 0; si < 3; ++si) {
        const long long* t = &d[(size_t)show[si] * 128];
        const long long t0 = t[0];
        printf("  cta %d [cycles from entry]: setup %lld | epilogue_done %lld | exit %lld\n", show[si], t[1] - t0, t[2] - t0,
               t[3] - t0);
        printf("    kt: loads_issued stage_published mma_issued acc_ready drained\n");
        for (int kt = 0; kt < nk && kt < 16; ++kt)
            printf("    %2d: %6lld %6lld %6lld %6lld %6lld\n", kt, t[16 + kt] - t0, t[32 + kt] - t0, t[48 + kt] - t0,
                   t[80 + kt] - t0, t[64 + kt] - t0);
    }
    // distribution of CTA lifetimes
    double sum = 0; long long mx = 0;
    for (int c = 0; c < ncta; ++c) { long long l = d[(size_t)c * 128 + 3] - d[(size_t)c * 128]; sum += l; if (l > mx) mx = l; }
    printf("  CTA lifetime: mean %.0f cycles, max %lld cycles\n", sum / ncta, mx);
    cudaFree(A); cudaFree(Apre); cudaFree(W); cudaFree(Wp); cudaFree(bias); cudaFree(Y); cudaFree(Ypre); cudaFree(dbg);
}

int main() {
    run(5376, 128, 384, 0, false);
    run(5376, 128, 128, 1, false);
    run(16128, 128, 256, 0, false);
    run(5376, 256, 128, 1, false);
    run(5376, 384, 128, 0, true);
    run(16128, 256, 128, 0, false);
    run(5376, 128, 256, 0, false);
    return 0;
}
