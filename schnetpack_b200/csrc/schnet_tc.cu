// SchNet interaction block, forward, as ONE fused edge kernel with the filter network on the tensor cores (tcgen05, TMEM).
// Reference: representation/schnet.py:56-70
//     Wij = filter_network(f_ij) * rcut_ij        filter_network = Dense(n_rbf -> F, ssp) -> Dense(F -> F)      :61-62
//     x   = scatter_add( x[idx_j] * Wij, idx_i )                                                                :65-67
// The reference (and round 1 of this repo) materialises the [E,F] hidden layer and the [E,F] filter in HBM per block (2 x
// 512 B per edge and layer, round trips included ~2 KB) -- 38 kFLOP per edge and layer of dense work that is tensor-pipe
// bound.  Here a chunk of 32 edges goes through both layers without leaving the SM:
//
//   producers   Phi' = [phi | 1 | 0] rows of the chunk (hi/lo TF32 split) -> shared memory (K-major, 64 B swizzle);
//   MMA 1       H = [W0 | b0] Phi'^T        128 hidden channels (TMEM lanes) x 32 edges (columns), 3xTF32, K = n_rbf + 1;
//               "N-folded": the hi and lo tiles of an edge operand are adjacent, so [X_hi ; X_lo] is ONE operand of 64 rows
//               and W_hi [X_hi ; X_lo]^T yields the main product (columns 0..31) and the W_hi X_lo terms (32..63) in one
//               instruction, W_lo X_hi accumulates onto the latter: 2 MMAs per k-step instead of 3 (an MMA costs ~130 cycles
//               for any N <= 256, and the chunk's serial MMA chain is what bounds this kernel);
//   activation  the 16 consumer warps (all 512 threads: lane = hidden channel, a group takes its own 8 edges of the NEXT chunk)
//               read H with tcgen05.ld, apply shifted softplus (nn/activations.py:9-22), split hi/lo and write it to shared
//               memory as the K-major B operand of the second layer -- 4 dedicated activation warps needed 7.2 k cycles per
//               chunk for the 32 log1p(exp()) per thread and were the chunk period (phase trace, profiles/);
//   MMA 2       D = W1 ssp(H)            128 output channels (lanes) x 32 edges, 3xTF32, K = 128: the main products in one
//               accumulator, the small hi*lo terms in a second one (the tensor core truncates when it accumulates);
//   consumers   16 warps = 4 groups x 128 channels, thread = channel as in the PaiNN kernels: filter value
//               W = (D_main + D_corr + b1) * fc for the group's 8 edges of the chunk with two tcgen05.ld.x8, gather
//               h[j, c], accumulate the receiver's row in registers (CSR order, deterministic, no atomics), m[i] on flush.
//
// Both weight operands ([W0|b0]: 32 KB, W1: 128 KB as hi|lo tiles) stay resident in shared memory for the life of the
// persistent CTA (one per SM) and arrive by two TMA bulk copies.  The receiver ranges come from a CSR that may have been
// built from the ACTIVE edges only (spk_graph_build_active: d < cutoff), so the padding slots of a padded neighbour list
// (cfg3: 62 % of 831 488 slots) cost nothing -- their filter is exactly 0 in the reference (nn/cutoff.py:30-32).
// The reverse pass (forces) keeps the round-1 pipeline (filter tensors materialised): the named SchNet workload with forces is
// the 9-atom cfg1; cfg3 is energy-only inference.
#include "tcgen05.cuh"

// development aid (build with -DSPK_SCHNET_TRACE, tools/build_variant.sh): clock64 stamps of CTA 0, one row of 16 per chunk
#ifdef SPK_SCHNET_TRACE
__device__ long long g_schnet_trace[256 * 16];
#define STRACE(k, slot)                                                                              \
    do {                                                                                             \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (k) < 256) g_schnet_trace[(k) * 16 + (slot)] = clock64(); \
    } while (0)
extern "C" int spk_debug_schnet_trace(long long* host) {
    return (int)cudaMemcpyFromSymbol(host, g_schnet_trace, sizeof(long long) * 256 * 16);
}
#else
#define STRACE(k, slot) do { } while (0)
#endif

namespace {

constexpr int F_TC = 128;
constexpr int EG = 8, NG = 4, NE = EG * NG;              // 32 edge rows per chunk == UMMA N
constexpr int NCW = NG * 4;                              // 16 consumer warps
constexpr int W_MMA = NCW;
constexpr int W_PROD0 = NCW + 1;
constexpr int NPROD = 2, NST = 2;                        // Phi' stages (chunk k -> stage k % 2 -> producer k % 2)
constexpr int NMETA = 4;                                 // (sender, fc) ring
constexpr int NTHREADS = (W_PROD0 + NPROD) * 32;         // 608
constexpr int KT = 16;
constexpr int A_TILE = F_TC * KT * 4;                    // 8192 B
constexpr int W0_BYTES = 2 * 2 * A_TILE;                 // [hi,lo][2 k-tiles]
constexpr int W1_BYTES = 2 * 8 * A_TILE;                 // [hi,lo][8 k-tiles]
constexpr int B_TILE = NE * KT * 4;                      // 2048 B
constexpr int PHI_STAGE = 2 * 2 * B_TILE;                // [hi,lo][2 k-tiles]
constexpr int B2_BYTES = 2 * 8 * B_TILE;                 // [hi,lo][8 k-tiles]
constexpr int META_STAGE = NE * 8;                       // int sender + float fc per row
constexpr int SMEM_BYTES = W0_BYTES + W1_BYTES + NST * PHI_STAGE + B2_BYTES + NMETA * META_STAGE + 1024;
constexpr int TMEM_COLS = 256;                           // H[2] x (main 32 | corr 32) | D[2] x (main 32 | corr 32)
constexpr int COL_H = 0, COL_D = 128;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

// packed operands: [W0|b0|0] as [hi,lo][2 k-tiles][128 x 16], then W1 as [hi,lo][8 k-tiles][128 x 16]
__global__ void k_pack_schnet_filter(const float* __restrict__ w0, const float* __restrict__ b0,
                                     const float* __restrict__ w1, int n_rbf, float* __restrict__ out) {
    SPK_PDL_WAIT_ONLY();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int N0 = F_TC * 32, N1 = F_TC * F_TC;
    if (t >= N0 + N1) return;
    float w;
    int row, k, nkt;
    float* base;
    if (t < N0) {
        k = t & 31;
        row = t >> 5;
        w = k < n_rbf ? w0[(int64_t)row * n_rbf + k] : (k == n_rbf ? b0[row] : 0.f);
        nkt = 2;
        base = out;
    } else {
        const int u = t - N0;
        k = u & (F_TC - 1);
        row = u >> 7;
        w = w1[(int64_t)row * F_TC + k];
        nkt = 8;
        base = out + W0_BYTES / 4;
    }
    const float hi = tf32_rn(w), lo = w - hi;
    const int kt = k >> 4, kk = k & 15;
    const int off = tile_off(row, kk >> 2) / 4 + (kk & 3);
    base[(0 * nkt + kt) * (A_TILE / 4) + off] = hi;
    base[(1 * nkt + kt) * (A_TILE / 4) + off] = lo;
}

__global__ void __launch_bounds__(NTHREADS, 1) k_schnet_cfconv_fwd_tc(
    const float* __restrict__ h, const float* __restrict__ phi, const float* __restrict__ geo,
    const int* __restrict__ rowptr, const int* __restrict__ slot_j, const float* __restrict__ wpk,
    const float* __restrict__ b1, int act, int n_atoms, int n_rbf, float* __restrict__ m) {
    constexpr int F = F_TC;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW0 = smem;
    uint8_t* sW1 = sW0 + W0_BYTES;
    uint8_t* sPhi = sW1 + W1_BYTES;
    uint8_t* sB2 = sPhi + NST * PHI_STAGE;
    uint8_t* sMeta = sB2 + B2_BYTES;
    __shared__ __align__(8) uint64_t w_full, phi_full[NST], phi_empty[NST], h_full[2], h_empty[2], b2_full, b2_empty,
        d_full[2], d_empty[2], meta_full[NMETA], meta_empty[NMETA];
    __shared__ uint32_t s_tmem;
    __shared__ int s_rlo[NG], s_rhi[NG], s_sb[NG], s_se[NG];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // broadcast: the compiler may treat the role index as warp-uniform
    SPK_PDL_LAUNCH_DEPENDENTS();
    if (tid == 0) {
        mbar_init(&w_full, 1);
#pragma unroll
        for (int s = 0; s < NST; ++s) {
            mbar_init(&phi_full[s], 1);
            mbar_init(&phi_empty[s], 1);
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            mbar_init(&h_full[b], 1);
            mbar_init(&h_empty[b], NCW);
            mbar_init(&d_full[b], 1);
            mbar_init(&d_empty[b], NCW);
        }
        mbar_init(&b2_full, NCW);
        mbar_init(&b2_empty, 1);
#pragma unroll
        for (int s = 0; s < NMETA; ++s) {
            mbar_init(&meta_full[s], 1);
            mbar_init(&meta_empty[s], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&w_full, W0_BYTES + W1_BYTES);      // static operand: fetched BEFORE griddepcontrol.wait (pack kernels never trigger their dependents early, common.cuh)
        tma_load(sW0, wpk, W0_BYTES, &w_full);
        tma_load(sW1, wpk + W0_BYTES / 4, W1_BYTES, &w_full);
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    SPK_PDL_WAIT();
    if (warp <= NG) {                              // warp w finds boundary w of this CTA's NG group ranges (warp-wide search)
        const int n_act = rowptr[n_atoms];                  // edges in the (possibly active-only) CSR
        const int bnd = spk_block_row_begin_warp(rowptr, n_atoms, n_act, gridDim.x * NG, blockIdx.x * NG + warp);
        if (lane == 0) {
            const int e = rowptr[bnd];
            if (warp < NG) {
                s_rlo[warp] = bnd;
                s_sb[warp] = e;
            }
            if (warp > 0) {
                s_rhi[warp - 1] = bnd;
                s_se[warp - 1] = e;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    int n_chunks = 0;
#pragma unroll
    for (int g = 0; g < NG; ++g) n_chunks = max(n_chunks, (s_se[g] - s_sb[g] + EG - 1) / EG);
    const int KP = spk_kp(n_rbf);
    const int nks1 = (n_rbf + 1 + 7) >> 3;                  // k-steps of the first layer (K = n_rbf + 1 padded to 8)

    if (warp >= W_PROD0) {
        // =========================================== producers ===========================================
        const int p = warp - W_PROD0;
        const int qc = lane & 7, rsub = lane >> 3;          // 8 x 16 B chunks per row (32 floats), 4 rows per pass, 8 passes
        const bool q_in = qc * 4 < KP;
        const int kb_chunk = n_rbf >> 2, kb = n_rbf & 3;    // the bias column k = n_rbf
        for (int k = p; k < n_chunks; k += NPROD) {
            const int st = k % NST, use = k / NST;
            const int ms = k % NMETA, muse = k / NMETA;
            uint8_t* stP = sPhi + st * PHI_STAGE;
            int* st_j = reinterpret_cast<int*>(sMeta + ms * META_STAGE);
            float* st_fc = reinterpret_cast<float*>(sMeta + ms * META_STAGE + NE * 4);
            // unconditional clamped loads, issued before the waits (rows past a group's end read slot 0, zeroed below)
            const int g0 = lane / EG;
            const int s0 = s_sb[g0] + k * EG + (lane % EG);
            const bool ok0 = s0 < s_se[g0];
            const int sl0 = ok0 ? s0 : 0;
            const int mj = slot_j[sl0];
            const float mfc = geo[(int64_t)sl0 * SPK_GEO_STRIDE + 4];
            float4 pv[8];
            bool okv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + rsub, g = r / EG;
                const int s = s_sb[g] + k * EG + (r % EG);
                okv[i] = s < s_se[g];
                const int sl = okv[i] ? s : 0;
                pv[i] = *reinterpret_cast<const float4*>(phi + (int64_t)sl * KP + (q_in ? qc * 4 : 0));
            }
            if (use >= 1) mbar_wait(&phi_empty[st], (use - 1) & 1);
            if (muse >= 1) mbar_wait(&meta_empty[ms], (muse - 1) & 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + rsub;
                float4 v = q_in ? pv[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                if (qc == kb_chunk) {
                    if (kb == 0) v.x = 1.f; else if (kb == 1) v.y = 1.f; else if (kb == 2) v.z = 1.f; else v.w = 1.f;
                }
                if (!okv[i]) v = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 hi, lo;
                hi.x = tf32_rn(v.x); hi.y = tf32_rn(v.y); hi.z = tf32_rn(v.z); hi.w = tf32_rn(v.w);
                lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
                const int off = tile_off(r, qc & 3);
                *reinterpret_cast<float4*>(stP + ((qc >> 2) * 2 + 0) * B_TILE + off) = hi;     // [k-tile][hi | lo]
                *reinterpret_cast<float4*>(stP + ((qc >> 2) * 2 + 1) * B_TILE + off) = lo;
            }
            st_j[lane] = mj;
            st_fc[lane] = ok0 ? mfc : 0.f;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&phi_full[st]);
                mbar_arrive(&meta_full[ms]);
            }
            STRACE(k, 0);                                   // producer: chunk published
        }
    } else if (warp == W_MMA) {
        // =========================================== MMA issuer ===========================================
        if (lane == 0) {
            const uint32_t idesc =
                (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NE >> 3) << 17) | ((uint32_t)(F_TC >> 4) << 24);
            const uint32_t idesc_w =                                       // N = 2 NE: [X_hi ; X_lo] as one operand
                (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * NE) >> 3) << 17) | ((uint32_t)(F_TC >> 4) << 24);
            mbar_wait(&w_full, 0);
            const uint32_t a0 = smem_u32(sW0), a1 = smem_u32(sW1), bb2 = smem_u32(sB2);
            auto mma1 = [&](int k) {                                      // H(k) = [W0|b0] Phi'(k)^T
                const int st = k % NST, hb = k & 1;
                mbar_wait(&phi_full[st], (k / NST) & 1);
                if (k >= 2) mbar_wait(&h_empty[hb], ((k >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t bp = smem_u32(sPhi + st * PHI_STAGE);
                const uint32_t d = tmem_base + (uint32_t)(COL_H + hb * 2 * NE);
                for (int s3 = 0; s3 < nks1; ++s3) {
                    const int kt = s3 >> 1, ko = (s3 & 1) * 32;
                    const uint64_t ah = make_desc(a0 + (0 * 2 + kt) * A_TILE + ko);
                    const uint64_t al = make_desc(a0 + (1 * 2 + kt) * A_TILE + ko);
                    const uint64_t bw = make_desc(bp + (kt * 2) * B_TILE + ko);       // hi tile, lo tile follows
                    umma_tf32(d, ah, bw, idesc_w, s3 ? 1u : 0u);                      // [Wh Xh | Wh Xl]
                    umma_tf32(d + (uint32_t)NE, al, bw, idesc, 1u);                   // Wl Xh onto the small-term columns
                }
                umma_commit(&phi_empty[st]);
                umma_commit(&h_full[hb]);
                STRACE(k, 1);                               // MMA 1 issued
            };
            if (n_chunks > 0) mma1(0);
            for (int k = 0; k < n_chunks; ++k) {
                if (k + 1 < n_chunks) mma1(k + 1);                        // keeps the activation warps one chunk ahead
                const int db = k & 1;
                mbar_wait(&b2_full, k & 1);
                STRACE(k, 2);                               // B2 of this chunk available
                if (k >= 2) mbar_wait(&d_empty[db], ((k >> 1) - 1) & 1);
                STRACE(k, 3);                               // D buffer free
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_main = tmem_base + (uint32_t)(COL_D + db * 2 * NE);
#pragma unroll
                for (int s16 = 0; s16 < 16; ++s16) {                      // D(k) = W1 ssp(H(k)), K = 128
                    const int kt = s16 >> 1, ko = (s16 & 1) * 32;
                    const uint64_t ah = make_desc(a1 + (0 * 8 + kt) * A_TILE + ko);
                    const uint64_t al = make_desc(a1 + (1 * 8 + kt) * A_TILE + ko);
                    const uint64_t bw = make_desc(bb2 + (kt * 2) * B_TILE + ko);      // [act_hi ; act_lo]
                    umma_tf32(d_main, ah, bw, idesc_w, s16 ? 1u : 0u);
                    umma_tf32(d_main + (uint32_t)NE, al, bw, idesc, 1u);
                }
                umma_commit(&b2_empty);
                umma_commit(&d_full[db]);
                STRACE(k, 4);                               // MMA 2 issued
            }
        }
    } else {
        // =========================================== consumers ===========================================
        const int g = warp >> 2, qd = warp & 3;
        const int c = qd * 32 + lane;
        const float b1c = b1 ? b1[c] : 0.f;
        const int row_hi = s_rhi[g], s_begin = s_sb[g], s_end = s_se[g];
        int i = s_rlo[g];
        int next_boundary = i < row_hi ? rowptr[i + 1] : 0x7fffffff;
        float acc = 0.f;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(COL_D + g * EG);
        // activation of chunk ka for this group's 8 edges: H[lane = hidden channel c, columns g*8..] -> ssp -> B2 rows g*8..
        const uint32_t lane_h = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(COL_H + g * EG);
        const int akt = c >> 4, akk = c & 15;
        auto activate = [&](int ka) {
            const int hb = ka & 1;
            mbar_wait(&h_full[hb], (ka >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp == 0) STRACE(ka, 5);                  // H ready
            float hm[EG], hc[EG];
            tmem_ld8_nowait(lane_h + (uint32_t)(hb * 2 * NE), hm);
            tmem_ld8_nowait(lane_h + (uint32_t)(hb * 2 * NE + NE), hc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&h_empty[hb]);
            float a[EG];
            if (act == SPK_ACT_SSP) {                      // the block's own activation (schnet.py:45), SFU version
#pragma unroll
                for (int e = 0; e < EG; ++e) a[e] = spk_ssp_fast(hm[e] + hc[e]);
            } else {
#pragma unroll
                for (int e = 0; e < EG; ++e) a[e] = spk_act(hm[e] + hc[e], act);
            }
            if (warp == 0) STRACE(ka, 6);                  // values computed
            if (ka >= 1) mbar_wait(&b2_empty, (ka - 1) & 1);              // MMA 2 of the previous chunk has read B2
            if (warp == 0) STRACE(ka, 7);                  // B2 free
            uint8_t* hi_t = sB2 + (akt * 2 + 0) * B_TILE;                 // [k-tile][hi | lo]
            uint8_t* lo_t = sB2 + (akt * 2 + 1) * B_TILE;
#pragma unroll
            for (int e = 0; e < EG; ++e) {
                const float hi = tf32_rn(a[e]);
                const int off = tile_off(g * EG + e, akk >> 2) + (akk & 3) * 4;
                *reinterpret_cast<float*>(hi_t + off) = hi;
                *reinterpret_cast<float*>(lo_t + off) = a[e] - hi;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&b2_full);
            if (warp == 0) STRACE(ka, 8);                  // B2 stored
        };
        if (n_chunks > 0) activate(0);
        for (int k = 0; k < n_chunks; ++k) {
            const int db = k & 1, ms = k % NMETA;
            if (k + 1 < n_chunks) activate(k + 1);          // the next chunk's hidden layer, while MMA 2 of this one runs
            mbar_wait(&d_full[db], (k >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp == 0) STRACE(k, 9);                    // consumer: D ready
            float wm[EG], wc[EG];
            tmem_ld8_nowait(lane_addr + (uint32_t)(db * 2 * NE), wm);
            tmem_ld8_nowait(lane_addr + (uint32_t)(db * 2 * NE + NE), wc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[db]);
            mbar_wait(&meta_full[ms], (k / NMETA) & 1);
            const int* st_j = reinterpret_cast<const int*>(sMeta + ms * META_STAGE) + g * EG;
            const float* st_fc = reinterpret_cast<const float*>(sMeta + ms * META_STAGE + NE * 4) + g * EG;
            const int base = s_begin + k * EG;
            if (base < s_end) {
                float hv[EG], fc[EG];
#pragma unroll
                for (int u = 0; u < EG; ++u) {                           // all gathers first (rows past the end read sender 0)
                    hv[u] = h[(size_t)st_j[u] * F + c];
                    fc[u] = st_fc[u];
                }
#pragma unroll
                for (int u = 0; u < EG; ++u) {
                    const int s = base + u;
                    if (s < s_end) {
                        while (s >= next_boundary) {
                            m[(size_t)i * F + c] = acc;
                            acc = 0.f;
                            ++i;
                            next_boundary = rowptr[i + 1];
                        }
                        acc = fmaf((wm[u] + wc[u] + b1c) * fc[u], hv[u], acc);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&meta_empty[ms]);
            if (warp == 0) STRACE(k, 10);                   // consumer: chunk done
        }
        for (; i < row_hi; ++i) {
            m[(size_t)i * F + c] = acc;
            acc = 0.f;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

}  // namespace

extern "C" size_t spk_schnet_filter_packed_floats(void) { return (W0_BYTES + W1_BYTES) / 4; }

extern "C" int spk_schnet_pack_filter(const float* w0, const float* b0, const float* w1, int F, int n_rbf, float* packed,
                                      spk_stream_t stream) {
    if (!w0 || !b0 || !w1 || !packed) return SPK_ERR_ARG;
    if (F != F_TC || n_rbf <= 0 || n_rbf > 31) return SPK_ERR_UNSUPPORTED;
    const int total = F_TC * 32 + F_TC * F_TC;
    spk_launch(k_pack_schnet_filter, (unsigned)spk_cdiv(total, 256), 256, 0, spk_st(stream), w0, b0, w1, n_rbf, packed);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_schnet_cfconv_fwd_tc(const float* h, const float* phi, const float* geo, const int32_t* rowptr,
                                        const int32_t* slot_j, const float* filter_packed, const float* b1, int act,
                                        int64_t n_atoms, int64_t n_edges, int F, int n_rbf, float* m, spk_stream_t stream) {
    if (n_atoms < 0 || n_edges < 0 || F <= 0 || n_rbf <= 0 || act < 0 || act > 2) return SPK_ERR_ARG;
    if (F != F_TC || n_rbf > 31) return SPK_ERR_UNSUPPORTED;
    if (n_atoms == 0) return SPK_OK;
    if (!h || !rowptr || !filter_packed || !m) return SPK_ERR_ARG;
    if (n_edges > 0 && (!phi || !geo || !slot_j)) return SPK_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(phi) | reinterpret_cast<uintptr_t>(filter_packed)) & 15) return SPK_ERR_UNSUPPORTED;
    static SpkSmemOnce once;
    if (cudaError_t e = once.set(k_schnet_cfconv_fwd_tc, SMEM_BYTES); e != cudaSuccess) return SPK_CUDA_ERR(e);
    int64_t nb = spk_num_sms();
    if (nb > spk_cdiv(n_edges, NE) + 1) nb = spk_cdiv(n_edges, NE) + 1;
    if (nb > n_atoms) nb = n_atoms;
    if (nb < 1) nb = 1;
    spk_launch(k_schnet_cfconv_fwd_tc, (unsigned)nb, NTHREADS, SMEM_BYTES, spk_st(stream), h, phi, geo, rowptr, slot_j,
               filter_packed, b1, act, (int)n_atoms, n_rbf, m);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
