// Halo exchange of ghost-atom rows over NVLink PEER MEMORY: the pack (gather of the rows a peer
// needs), the transfer and the unpack are ONE kernel on the receiving side -- every rank reads the rows of its ghost atoms
// straight out of their owners' buffers (peer-mapped symmetric memory, CUDA IPC over NVLink / NVSwitch) with coalesced
// 16-byte loads, and writes them into the ghost block of its own table.  No send-side pack kernel, no staging copy, no
// NCCL call; the only synchronisation is the symmetric-memory barrier the caller enqueues before the pull (owners have
// published) -- buffers alternate between two generations so that no second barrier is needed.
//
//   forward   ghost[k, :]   = peer[g_rank[k]][ g_row[k], : ]                          k in [0, n_ghost)
//   reverse   g_rows[r, :] += sum over the peers p (ascending) that hold r as a ghost of  peer[p][ pos, : ]
//             (per-row entry lists in CSR form; fixed order -> deterministic, no atomics)
// Rows have C floats (C % 4 == 0 uses 16-byte accesses; the positions exchange has C = 3).
#include "common.cuh"

namespace {

__global__ void k_halo_pull(float* __restrict__ dst, const unsigned long long* __restrict__ peer_base,
                            const int* __restrict__ g_rank, const int* __restrict__ g_row, int64_t n_ghost, int C) {
    SPK_PDL_ENTER();
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if ((C & 3) == 0) {
        const int C4 = C >> 2;
        if (t >= n_ghost * C4) return;
        const int64_t k = t / C4;
        const int c = (int)(t - k * C4);
        const float4* src = reinterpret_cast<const float4*>(peer_base[g_rank[k]]) + (int64_t)g_row[k] * C4 + c;
        reinterpret_cast<float4*>(dst)[t] = __ldcv(src);          // peer memory: never served from a stale local cache line
    } else {
        if (t >= n_ghost * C) return;
        const int64_t k = t / C;
        const int c = (int)(t - k * C);
        const float* src = reinterpret_cast<const float*>(peer_base[g_rank[k]]) + (int64_t)g_row[k] * C + c;
        dst[t] = __ldcv(src);
    }
}

__global__ void k_halo_pull_add(float* __restrict__ g_rows, const unsigned long long* __restrict__ peer_base,
                                const int* __restrict__ rows, const int* __restrict__ entry_ptr,
                                const int* __restrict__ e_rank, const int* __restrict__ e_pos, int64_t n_listed, int C) {
    SPK_PDL_ENTER();
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if ((C & 3) == 0) {
        const int C4 = C >> 2;
        if (t >= n_listed * C4) return;
        const int64_t k = t / C4;
        const int c = (int)(t - k * C4);
        float4* out = reinterpret_cast<float4*>(g_rows) + (int64_t)rows[k] * C4 + c;
        float4 acc = *out;
        for (int e = entry_ptr[k]; e < entry_ptr[k + 1]; ++e) {
            const float4 v = __ldcv(reinterpret_cast<const float4*>(peer_base[e_rank[e]]) + (int64_t)e_pos[e] * C4 + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *out = acc;
    } else {
        if (t >= n_listed * C) return;
        const int64_t k = t / C;
        const int c = (int)(t - k * C);
        float* out = g_rows + (int64_t)rows[k] * C + c;
        float acc = *out;
        for (int e = entry_ptr[k]; e < entry_ptr[k + 1]; ++e)
            acc += __ldcv(reinterpret_cast<const float*>(peer_base[e_rank[e]]) + (int64_t)e_pos[e] * C + c);
        *out = acc;
    }
}

}  // namespace

#define GRID1D(n, T) (unsigned)spk_cdiv((n), (T)), (T), 0, spk_st(stream)

extern "C" int spk_halo_pull(float* ghost, const uint64_t* peer_base, const int32_t* ghost_rank, const int32_t* ghost_row,
                             int64_t n_ghost, int row_floats, spk_stream_t stream) {
    if (n_ghost < 0 || row_floats <= 0) return SPK_ERR_ARG;
    if (n_ghost == 0) return SPK_OK;
    if (!ghost || !peer_base || !ghost_rank || !ghost_row) return SPK_ERR_ARG;
    if (!(row_floats & 3) && (reinterpret_cast<uintptr_t>(ghost) & 15)) return SPK_ERR_ARG;
    const int64_t n = n_ghost * ((row_floats & 3) ? row_floats : row_floats / 4);
    spk_launch(k_halo_pull, GRID1D(n, 256), ghost, reinterpret_cast<const unsigned long long*>(peer_base), ghost_rank,
               ghost_row, n_ghost, row_floats);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_halo_pull_add(float* g_rows, const uint64_t* peer_base, const int32_t* rows, const int32_t* entry_ptr,
                                 const int32_t* entry_rank, const int32_t* entry_pos, int64_t n_listed, int row_floats,
                                 spk_stream_t stream) {
    if (n_listed < 0 || row_floats <= 0) return SPK_ERR_ARG;
    if (n_listed == 0) return SPK_OK;
    if (!g_rows || !peer_base || !rows || !entry_ptr || !entry_rank || !entry_pos) return SPK_ERR_ARG;
    if (!(row_floats & 3) && (reinterpret_cast<uintptr_t>(g_rows) & 15)) return SPK_ERR_ARG;
    const int64_t n = n_listed * ((row_floats & 3) ? row_floats : row_floats / 4);
    spk_launch(k_halo_pull_add, GRID1D(n, 256), g_rows, reinterpret_cast<const unsigned long long*>(peer_base), rows, entry_ptr,
               entry_rank, entry_pos, n_listed, row_floats);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
