// tc5.cuh -- thin inline-PTX layer over the sm_100a tensor-core path: tcgen05.mma (UMMA), TMEM alloc/ld/st,
// tcgen05.commit + mbarrier, proxy fences.  Only what the b200sr kernels use; bf16 x bf16 -> fp32, cta_group::1.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace b200sr {
namespace tc5 {

// ---- shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleaved" core matrices) ------------------------
// A core matrix is 8 rows x 16 bytes stored contiguously (128 B).  Element (row r, 16-byte k-chunk c in {0,1}) of an
// operand lives at   start + (r / 8) * SBO + c * LBO + (r % 8) * 16.
// Because LBO is a free stride, ONE K=16 instruction may pair any two 8-channel chunks whose distance is the same in
// every 8-row group -- the kernels use that to contract ragged K (24 = 16 + 8, 9 taps x 24) without zero padding.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);  // version = 1 (Blackwell), layout_type = 0
}

// ---- instruction descriptor: kind::f16, A = B = bf16 (K-major), D = fp32, M in {64,128}, N % 16 == 0 ------------
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T      (issued by ONE thread)
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T      (A: 128 lanes x K/2 columns of packed bf16 pairs)
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// make the completion of all previously issued MMAs of this thread arrive (count 1) on an mbarrier
__device__ __forceinline__ void commit(uint32_t bar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_saddr) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes (st.shared / cp.async) -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One leader lane of a CONVERGED warp.  Issuing tcgen05.mma from `if (leader)` of a converged warp (instead of a
// lane whose warp-mates run ahead) keeps the per-instruction scalar overhead at ~35-40 clk (measured, tools/umma_bench2.cu).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- programmatic dependent launch (no-ops when the grid was launched without the attribute) ------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- register re-balancing between warpgroups (all 4 warps of an aligned warpgroup must execute it) ----------------------
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ---- TMEM ---------------------------------------------------------------------------------------------------------
// address = (lane << 16) | column.  One warp allocates / frees; a warp can only touch lanes 32*(warp%4) .. +31.
__device__ __forceinline__ void tmem_alloc(uint32_t dst_saddr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_saddr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// relu + round-to-nearest-even + pack two fp32 into bf16x2 in ONE instruction (lo in bits 0..15)
__device__ __forceinline__ uint32_t relu_pack_bf16x2(uint32_t lo_f32, uint32_t hi_f32) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi_f32)), "f"(__uint_as_float(lo_f32)));
    return d;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// 16x128b shape: a warp reads 16 TMEM lanes (taddr.lane = 32 * (warp % 4) + {0, 16}) x 4 columns per repeat: thread (g = lane / 4,
// j = lane % 4) gets, per repeat i, r[2i] = (lane g, column 4i + j) and r[2i+1] = (lane g + 8, column 4i + j)  (verified on B200:
// tools/tmem_frag_probe.cu).  Four repeats = 16 columns = one K = 16 step of mma.sync.m16n8k16 with the K order permuted: A fragment
// a0 = pack(r0, r2), a1 = pack(r1, r3), a2 = pack(r4, r6), a3 = pack(r5, r7)  <=>  the B fragment holds columns {j, 4 + j} and
// {8 + j, 12 + j}.  Measured tcgen05.ld rates per SM (tools/tmem_ld_bench.cu): 32x32b 814 B/clk, 16x128b 485, 16x256b (the shape whose
// registers are the A fragment without a K permutation) only 251; one warp alone: 56 / 39 / 31 B/clk.
__device__ __forceinline__ void tmem_ld_16x128b_x4(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_16x128b_x8(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x128b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_16x128b_x16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x128b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// legacy warp-level MMA (HMMA.16816): D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col).  B200: 8.2 clk per instruction and
// SM sub-partition = ~1000 MAC/clk/SM, a quarter of tcgen05's rate, and it shares the datapath with it (tools/hmma_bench.cu).
__device__ __forceinline__ void hmma_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- TMA (cp.async.bulk.tensor): 5-D tiled load global -> shared, completion counted in bytes on an mbarrier --------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst_saddr, const void *tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// one 8-channel plane box of the planar-8 trunk (tma_map.h): coordinates in pixels / rows / plane / image
__device__ __forceinline__ void tma_load_plane(uint32_t dst_saddr, const void *tmap, uint32_t bar, int x, int y, int plane, int n) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(4 * x), "r"(y), "r"(plane), "r"(n)
        : "memory");
}
// plain (non-tensor) bulk copy global -> shared, completion counted in bytes on an mbarrier; addresses and size multiples of 16
__device__ __forceinline__ void bulk_load_g2s(uint32_t dst_saddr, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// ---- mbarrier -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
// arrive WITHOUT release semantics: a release arrive also waits for the thread's outstanding GLOBAL stores (~800 clk after
// the output stores of an epilogue).  Use only where everything the barrier publishes has already completed
// (tcgen05.wait::ld / ::st retired, shared-memory values consumed).
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.relaxed.cta.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
// the executing thread's prior cp.async copies, once complete, arrive on the mbarrier (the arrival is pre-counted in its init count)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// Developer switch: -DB200SR_MBAR_SUSPEND_HINT_NS=<ns> passes a suspend-time hint.  Without it try_wait comes back after a short
// system-defined time and a waiting warp's retry loop keeps re-issuing -- a third of ALL executed instructions in the row-streaming block
// kernel (ncu source page, profiles/r02_block_rs_ncu.md).  Measured: neither the hint (10 ms) nor letting only one warp per warpgroup
// poll (mbar_wait_wg) changed the kernel times -- the retries only take issue slots nobody else wanted.
#ifndef B200SR_MBAR_SUSPEND_HINT_NS
#define B200SR_MBAR_SUSPEND_HINT_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
#if B200SR_MBAR_SUSPEND_HINT_NS > 0
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"((uint32_t)B200SR_MBAR_SUSPEND_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
#endif
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {  // non-blocking probe
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
#ifdef B200SR_TC5_PROF
__device__ volatile unsigned *g_tc5_dbg;  // mapped host memory: who timed out on which barrier (developer probes only)
#endif
// bounded wait: a protocol bug must fault the launch (caught by the host as a CUDA error), never hang the GPU box
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it)
        if (it > (1u << 22)) {
#ifdef B200SR_TC5_PROF
            if (g_tc5_dbg && (threadIdx.x & 31) == 0) {
                volatile unsigned *d = g_tc5_dbg + 4 * (blockIdx.x * 32 + (threadIdx.x >> 5));
                d[0] = 0xDEAD0000u | (threadIdx.x >> 5), d[1] = bar, d[2] = parity, d[3] = blockIdx.x;
                __threadfence_system();
            }
            for (volatile int spin = 0; spin < 2000000; ++spin) {}
#endif
            __trap();
        }
}

// A whole warpgroup (4 consecutive warps) waits for one mbarrier phase: only its first warp polls, the other three block in a named
// hardware barrier (bar.sync costs no issue slots while blocked; a polling warp re-issues its retry loop and competes with the warps that
// have work -- a third of all executed instructions in the row-streaming block kernel before this).  id: 1..15, one per warpgroup.
__device__ __forceinline__ void mbar_wait_wg(uint32_t bar, uint32_t parity, int id) {
    if (((threadIdx.x >> 5) & 3) == 0) mbar_wait(bar, parity);
    asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");
}
// spin on the non-blocking probe: lower wake-up latency than the suspending try_wait, at the price of issue slots and shared-memory
// probes while waiting.  For the one or two barriers that sit on a kernel's critical loop only.
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; !mbar_test(bar, parity); ++it)
        if (it > (1u << 28)) __trap();
}

}  // namespace tc5
}  // namespace b200sr
