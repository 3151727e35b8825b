// ptx.cuh -- thin inline-PTX wrappers for the sm_100a features the Pixie kernels use:
// mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (alloc / mma.kind::tf32 / commit / ld) and
// named barriers.  Written for sm_100a only; nothing here compiles for another target.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pixie {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    // no suspend-time hint: with a hint ptxas emits TRYWAIT + NANOSLEEP and the wake-up after the
    // phase completes is slow enough to cost 40 % of the kernel (measured, profiles/r01_notes.md)
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a wedged pipeline traps (-> cudaErrorLaunchFailure) after ~2 s instead of hanging
// the GPU.  The clock is only read every 4096 failed probes, i.e. never on the normal path.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 4095u) == 0u) {
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > 2000000000ull) __trap();
        }
    }
}

// One thread of a CONVERGED warp: true in exactly one lane (the same one every time).  The
// single-thread instructions (TMA, tcgen05.mma / commit) are issued under this predicate from
// warp-uniform control flow; under a plain `if (lane == 0)` region the compiler serialises every
// uniform-datapath instruction with an ELECT / BRA.U.ANY loop, which made the MMA issuer spend
// ~900 cycles per tile issuing six instructions (profiles/r01_notes.md).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tensormap(const void *tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void *tmap, uint32_t bar,
                                            int32_t c0, int32_t c1, uint64_t cache_hint)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        :
        : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1),
          "l"(cache_hint)
        : "memory");
}
// Same box, but only as far as L2: no shared memory, no barrier.  The producer issues it one stage
// cycle ahead of the real load, which then finds its lines in L2 (shorter and steadier latency
// than HBM under load: profiles/r02_notes.md section 8).
__device__ __forceinline__ void tma_prefetch_2d(const void *tmap, int32_t c0, int32_t c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1)
                 : "memory");
}
// 1-D bulk copy global -> shared, completion on an mbarrier.  bytes % 16 == 0.
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes,
                                          uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :
        : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
        : "memory");
}

// ---------------------------------------------------------------- named barriers
__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before()
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after()
{
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_wait_ld()
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 operands (fp32 bit patterns, low 13 mantissa bits
// ignored by the tensor core), fp32 accumulate.  Issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :
        : "r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All tcgen05.mma issued so far by this thread arrive (once) on the mbarrier when they retire.
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     bar)
                 : "memory");
}

// tcgen05.ld 32x32b: thread i of the warp receives N consecutive fp32 columns of TMEM lane
// (32 * (warp % 4) + i).
#define PIXIE_R4(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3])
#define PIXIE_R8(v, o) PIXIE_R4(v, o), PIXIE_R4(v, o + 4)
#define PIXIE_R16(v, o) PIXIE_R8(v, o), PIXIE_R8(v, o + 8)
#define PIXIE_R32(v, o) PIXIE_R16(v, o), PIXIE_R16(v, o + 16)

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];"
                 : "=r"(v[0]), "=r"(v[1])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : PIXIE_R4(v, 0)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t *v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : PIXIE_R8(v, 0)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : PIXIE_R16(v, 0)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : PIXIE_R32(v, 0)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t *v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : PIXIE_R32(v, 0), PIXIE_R32(v, 32)
        : "r"(taddr)
        : "memory");
}

// Loads N (even, <= 128) consecutive columns with the fewest instructions.
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t *v)
{
    static_assert(N % 2 == 0 && N >= 2 && N <= 128, "slice width");
    int o = 0;
    if constexpr (N >= 64) {
        tmem_ld64(taddr, v);
        o = 64;
    }
    if constexpr (N == 128) {
        tmem_ld64(taddr + 64, v + 64);
        o = 128;
    }
    constexpr int R = (N == 128) ? 0 : (N % 64);
    if constexpr (R >= 32) {
        tmem_ld32(taddr + o, v + o);
        o += 32;
    }
    if constexpr ((R % 32) >= 16) {
        tmem_ld16(taddr + o, v + o);
        o += 16;
    }
    if constexpr ((R % 16) >= 8) {
        tmem_ld8(taddr + o, v + o);
        o += 8;
    }
    if constexpr ((R % 8) >= 4) {
        tmem_ld4(taddr + o, v + o);
        o += 4;
    }
    if constexpr ((R % 4) >= 2) {
        tmem_ld2(taddr + o, v + o);
        o += 2;
    }
}

// 128-bit shared-memory load from a 32-bit shared address
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "r"(addr));
    return r;
}

// 128-bit shared-memory store to a 32-bit shared address
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
                 "r"(d)
                 : "memory");
}

// ---------------------------------------------------------------- packed fp32 (FADD2 / FFMA2)
__device__ __forceinline__ uint64_t pack2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t pack2u(uint32_t lo, uint32_t hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void unpack2u(uint64_t v, uint32_t &lo, uint32_t &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// ---------------------------------------------------------------- descriptors
// K-major operand tile, rows at a 128-byte pitch, SWIZZLE_128B (Swizzle<3,4,3>): 8-row groups are
// 1024 bytes apart (SBO); LBO is unused for swizzled K-major layouts (encoded as 1).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) |
           (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major operand tile with 32-byte rows (ONE tf32 K-step per row), SWIZZLE_32B (Swizzle<1,4,3>):
// 8-row groups are 256 bytes apart (SBO).
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr)
{
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) |
           (static_cast<uint64_t>(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// K-major, no swizzle ("interleave"): 8x16-byte core matrices; lbo = byte distance between the two
// K-adjacent core matrices of one MMA, sbo = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo)
{
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) |
           (static_cast<uint64_t>(lbo >> 4) << 16) | (static_cast<uint64_t>(sbo >> 4) << 32) |
           (1ull << 46);
}
// kind::tf32 instruction descriptor: fp32 accumulator, tf32 A and B, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

}  // namespace ptx
}  // namespace pixie
