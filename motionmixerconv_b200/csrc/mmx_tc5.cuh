// Blackwell (sm_100a) primitives used by the tcgen05 kernels: tensor memory (TMEM) allocation and access, tcgen05.mma with
// shared-memory / TMEM operands, mbarriers, the bulk-copy engine (TMA, cp.async.bulk) and the proxy fences between them.
// Device-only; everything is inline PTX (no CUTLASS).
//
// Operand layout ("panel layout", used for every shared-memory operand of these kernels):
//
//     element (row r, col c) of an [R x C] fp32 matrix lives at byte   (c / 4) * (R * 16)  +  r * 16  +  (c % 4) * 4
//
// i.e. one panel per 4-column chunk, a panel being R rows of 16 bytes.  R is a multiple of 8, C a multiple of 4.
// The same bytes are a valid tcgen05 shared-memory operand in BOTH orientations (SWIZZLE_NONE canonical layouts):
//   * "K-major",  rows = M (or N) index, cols = K index:  core matrix = 8 rows x 16 B contiguous (128 B),
//                 stride between 8-row groups (SBO) = 128 B, stride between the two 16-byte K chunks of one K=8 MMA (LBO) = R*16
//   * "MN-major", rows = K index, cols = M (or N) index:  core matrix = 8 K-rows x 16 B (4 MN elements),
//                 stride between 4-element MN chunks (SBO) = R*16, stride between 8-row K groups (LBO) = 128 B
// so a tile written once (one thread per row, 16-byte stores: lanes hit consecutive 16-byte slots, conflict-free) serves as the
// A operand of  D = A * W^T  and as an operand of the weight-gradient product  dW = dY^T * A  without a transposed copy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmx {
namespace tc5 {

#define MMX_TC5_D __device__ __forceinline__

MMX_TC5_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------ mbarrier
MMX_TC5_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make mbarrier.init visible to the async proxy (TMA / tcgen05.commit arrive on it)
MMX_TC5_D void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
MMX_TC5_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
MMX_TC5_D void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
MMX_TC5_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a mis-programmed pipeline must end the kernel (and fail the test) instead of hanging the GPU.  `*abort_flag`
// (shared) is set on timeout; callers check it after the kernel's join points.
MMX_TC5_D bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        if (mbar_try_wait(bar, parity)) return true;
        if (it > 64 && *abort_flag) return false;
    }
    *abort_flag = 1;
    return false;
}

// ------------------------------------------------------------------------------------------ proxy / tcgen05 fences
// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy (tcgen05.mma operand reads, bulk stores)
MMX_TC5_D void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
MMX_TC5_D void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MMX_TC5_D void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
MMX_TC5_D void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
MMX_TC5_D void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ TMEM allocation (one full warp)
template <int COLS>
MMX_TC5_D void tmem_alloc(uint32_t* slot_in_smem) {
    static_assert(COLS == 32 || COLS == 64 || COLS == 128 || COLS == 256 || COLS == 512, "TMEM columns: power of two >= 32");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
MMX_TC5_D void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// ------------------------------------------------------------------------------------------ descriptors
// shared-memory matrix descriptor, SWIZZLE_NONE, Blackwell version field = 1
MMX_TC5_D uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// panel-layout operand, rows = M/N index, cols = K index; k0 = first column of this K=8 step (multiple of 8)
MMX_TC5_D uint64_t desc_kmajor(uint32_t base, uint32_t panel_bytes, int k0) {
    return smem_desc(base + (uint32_t)(k0 >> 2) * panel_bytes, panel_bytes, 128u);
}
// panel-layout operand, rows = K index, cols = M/N index; r0 = first row of this K=8 step (multiple of 8), c0 = first M/N column
MMX_TC5_D uint64_t desc_mnmajor(uint32_t base, uint32_t panel_bytes, int r0, int c0) {
    return smem_desc(base + (uint32_t)(c0 >> 2) * panel_bytes + (uint32_t)r0 * 16u, 128u, panel_bytes);
}
// instruction descriptor: kind::tf32, fp32 accumulate, dense
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ MMA issue (ONE thread)
// D[tmem] (+)= A[smem] * B[smem]
MMX_TC5_D void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]   (A: one row per lane, one K element per 32-bit column; K-major only)
MMX_TC5_D void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they have completed (implies fence::before_thread_sync)
MMX_TC5_D void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------ TMEM <-> registers (whole warp)
// Lane l of warp w (w % 4 = lane quarter) touches TMEM lane 32*(w%4) + l, columns [col, col + N).
MMX_TC5_D uint32_t tmem_addr(uint32_t base, int lane_quarter, int col) { return base + ((uint32_t)(lane_quarter * 32) << 16) + (uint32_t)col; }

MMX_TC5_D void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}
MMX_TC5_D void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
MMX_TC5_D void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(a)),
                 "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d))
                 : "memory");
}
MMX_TC5_D void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

// ------------------------------------------------------------------------------------------ bulk copies (TMA engine, 1-D)
// global -> shared, completes `bytes` on the mbarrier; addresses and size multiples of 16
MMX_TC5_D void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global (bulk async-group)
MMX_TC5_D void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
MMX_TC5_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the source shared memory of all committed bulk stores has been read (buffer reusable)
MMX_TC5_D void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
MMX_TC5_D void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ programmatic dependent launch
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the stream is
// still running: everything before pdl_wait() (barrier / TMEM set-up, staging of the weights -- nothing the predecessor
// writes) overlaps the predecessor's tail; pdl_wait() returns once the predecessor has completed and its writes are visible.
// pdl_launch_dependents() lets the NEXT kernel's CTAs be scheduled as soon as this kernel's CTAs have all started.
MMX_TC5_D void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
MMX_TC5_D void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ TF32 operand split
// hi = x with the 13 low mantissa bits cleared (exactly what the tensor core keeps of an fp32 word), lo = x - hi (exact in
// fp32; the tensor core again keeps its top bits).  hi*w_hi + lo*w_hi + hi*w_lo recovers the fp32 product to ~2^-21.
MMX_TC5_D float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

}  // namespace tc5
}  // namespace mmx
