// tcgen05 / TMA / mbarrier primitives shared by the tensor-core kernels (gemm_tc.cu, painn_tc.cu): thin inline-PTX wrappers
// and the one operand layout both kernels use -- K-major tiles of 16 floats (64 B rows) with the 64-byte swizzle.
#pragma once
#include "common.cuh"

constexpr int SPK_SW64_ROW = 64;               // bytes per operand row (16 floats)
constexpr int SPK_SW64_SBO = 8 * SPK_SW64_ROW; // 512 B between consecutive 8-row groups

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// arrives on `bar` when every tcgen05.mma issued so far by this thread has completed (implies before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_64B:
// start>>4 [0,14) | LBO (unused, 1) [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout 4 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(SPK_SW64_SBO >> 4) << 32) | (1ull << 46) |
           (4ull << 61);
}
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// byte offset of (row r, 16 B K-chunk c in 0..3) inside an operand tile: rows are 64 B, groups of 8 rows are 512 B atoms and
// the chunk index is XOR-swizzled with bits [1,3) of the row (Swizzle<2,4,3>, the pattern the tensor core applies to the
// byte address when the descriptor says SWIZZLE_64B).  Tiles are 512 B aligned.
__host__ __device__ __forceinline__ int tile_off(int r, int c) {
    return (r >> 3) * SPK_SW64_SBO + (r & 7) * SPK_SW64_ROW + ((c ^ ((r >> 1) & 3)) << 4);
}
