// common.cuh — shared host/device utilities for liblocov_b200 (sm_100a only).
// Hand-written PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit /
// ld) and the host-side error plumbing of the C ABI (include/locov_b200.h).
#pragma once

#include <cuda.h>          // CUtensorMap & enums only — the driver entry point is resolved at run time
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

#include "../../include/locov_b200.h"

namespace loco {

// ------------------------------------------------------------------------------------------------
// host side: thread-local error string + checks
// ------------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);   // records message, returns (int)e

#define LOCO_CUDA(call)                                             \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return ::loco::cuda_fail(_e, #call); \
    } while (0)

#define LOCO_REQUIRE(cond, code, ...)      \
    do {                                   \
        if (!(cond)) {                     \
            ::loco::set_error(__VA_ARGS__); \
            return (code);                 \
        }                                  \
    } while (0)

// 2-D bf16 row-major tensor map: dims {cols (inner), rows}, box {64 elements = 128 B, box_rows},
// 128-byte swizzle, zero fill out of bounds.  Returns LOCO_OK or an error code.
int make_tmap_bf16_2d(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows);

// Same for an fp32 (TF32 operand) matrix: box {32 elements = 128 B, box_rows}.
int make_tmap_f32_2d(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                     uint32_t box_rows);

// 4-D store map of a pooled [R, C, PH*PW] fp32 output seen as [R][C/4][4][PH*PW]: box = {PH*PW, 1, 32, 1}, i.e. the 32 dense
// tile rows that hold channels 4*l + j (l = 0..31) of one 128-channel slab (roi_align.cu); stores past C/4 are clipped.
int make_tmap_roi_out(CUtensorMap *map, const float *out, uint64_t R, uint64_t C, uint64_t PHW);
int make_tmap_bf16_out(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems);
int current_device_sm_count();
void count_launch(int n = 1);   // bumps the library-wide kernel-launch counter (loco_launch_count)

// ------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------
// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// Every kernel of the library starts with pdl_trigger(); pdl_wait(); and is launched through launch_kernel() with the
// programmatic-stream-serialization attribute: the next kernel of the stream is scheduled while this one still runs, does
// its set-up (barrier init, TMEM allocation, tensor-map prefetch) and blocks in griddepcontrol.wait until this grid has
// completed and flushed — the 1-2 us launch gap between the short dependent kernels of the path disappears.  All global
// reads of produced data and all global writes come after pdl_wait().  LOCOV_B200_PDL=0 launches without the attribute
// (the device-side instructions are then no-ops).
bool pdl_enabled();      // api.cu

#if defined(__CUDACC__)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    unsigned n = 0;
    if (cluster > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = (unsigned)cluster;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
    uint32_t l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
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
// Bounded wait: a pipeline bug must surface as a trapped kernel (a CUDA error on the host), never as
// a hung GPU.  2 s is ~1e4 x the longest legitimate wait in any kernel of this library.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = global_timer_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (global_timer_ns() - t0 > 2000000000ull) {
            printf("locov_b200: mbarrier wait timed out (block %d thread %d parity %u)\n", (int)blockIdx.x,
                   (int)threadIdx.x, parity);
            __trap();
        }
    }
}

// ---- TMA ------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tile load global -> shared, completion signalled on `bar` (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int32_t c_inner,
                                            int32_t c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}

// L2 prefetch of a 2-D tile (no shared-memory destination, no completion signal).
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *m, int32_t c_inner, int32_t c_outer) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c_inner),
                 "r"(c_outer)
                 : "memory");
}

// 2-D tile load multicast to every CTA of the cluster whose bit is set in `cta_mask`: the tile lands at the same
// shared-memory offset in each destination CTA and signals complete_tx on the mbarrier at the same offset there.
__device__ __forceinline__ void tma_load_2d_mc(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int32_t c_inner,
                                               int32_t c_outer, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer), "h"(cta_mask)
        : "memory");
}
// One lane of the (converged) warp, chosen by the hardware: code under this predicate is known to the compiler to run in a
// single thread, so tcgen05 / TMA instructions (which take warp-uniform operands) need no per-lane serialisation loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {      // every thread of every CTA in the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Execution-only rendezvous of the cluster (no memory ordering of its own): the set-up sync is ordered by fence.mbarrier_init.release.cluster,
// the exit sync only keeps a CTA alive while peers may still signal its barriers.  The release / acquire form above costs a GPU-scope
// memory barrier per warp (~0.5 us of the ~1 us set-up the phase timeline shows).
__device__ __forceinline__ void cluster_sync_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}

// ---- tcgen05 / TMEM ---------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {   // whole warp, ncols pow2 >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major bf16 operand tile stored as [rows][64] (128 B per row)
// with the 128-byte swizzle TMA wrote it with: SBO = 8 rows * 128 B = 1024 B, LBO unused (1),
// version 1 (sm_100), layout type 2 = SWIZZLE_128B.  Bit layout: cute/arch/mma_sm100_desc.hpp.
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);        // [0,14)  start address >> 4
    d |= static_cast<uint64_t>(1) << 16;                           // [16,30) leading byte offset >> 4 (ignored)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // [32,46) stride byte offset >> 4
    d |= static_cast<uint64_t>(1) << 46;                           // [46,48) descriptor version = 1
    d |= static_cast<uint64_t>(2) << 61;                           // [61,64) SWIZZLE_128B
    return d;
}
// Instruction descriptor, kind::f16: D fp32, A/B bf16, both K-major, M x N tile.
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
    uint32_t d = 0;
    d |= 1u << 4;              // [4,6)   D format  = F32
    d |= 1u << 7;              // [7,10)  A format  = BF16
    d |= 1u << 10;             // [10,13) B format  = BF16
    d |= (N >> 3) << 17;       // [17,23) N >> 3
    d |= (M >> 4) << 24;       // [24,29) M >> 4
    return d;                  // a_major = b_major = 0 (K-major), no negate, dense
}
// Instruction descriptor, kind::tf32: D fp32, A/B tf32 (fp32 bit patterns in shared memory), both K-major.
__device__ __forceinline__ uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
    uint32_t d = 0;
    d |= 1u << 4;              // D format = F32
    d |= 2u << 7;              // A format = TF32
    d |= 2u << 10;             // B format = TF32
    d |= (N >> 3) << 17;
    d |= (M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {      // M x N x 8 per instruction
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on `bar` when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// Same, arriving on the barrier at this offset in EVERY CTA of the cluster selected by `cta_mask`.
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}
// ---- CTA pair (cta_group::2): two SMs of a TPC execute one M = 256 MMA; each CTA holds its own 128 rows of A and HALF of
// the B tile in shared memory, the accumulator rows land in each CTA's own TMEM.  Per SM the operand bytes per MAC drop by
// (128 + N) / (128 + N/2): the 128 x N single-CTA tiles of this library are bound by the SM's shared-memory fill rate.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {   // shared::cta address -> shared::cluster address in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {      // arrive on a (possibly remote) mbarrier
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Same with CTA-scope release: enough when the arrival publishes no shared / global data to the other CTA — the accumulator-empty
// signal orders tensor-memory reads, which tcgen05.fence::before_thread_sync covers — and avoids the cluster-scope memory barrier.
__device__ __forceinline__ void mbar_arrive_cluster_cta(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cta.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-D tile load into THIS CTA's shared memory whose completion is signalled on an mbarrier of either CTA of the pair
// (`bar_cluster_addr` = shared::cluster address, normally the leader's barrier).
__device__ __forceinline__ void tma_load_2d_2sm(void *smem_dst, const CUtensorMap *m, uint32_t bar_cluster_addr, int32_t c_inner,
                                                int32_t c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c_inner), "r"(c_outer)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *smem_dst, uint32_t ncols) {  // one warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in both CTAs of the pair once all previously issued pair MMAs have completed
__device__ __forceinline__ void umma_commit_2cta(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

// One K step (32 bytes of K per operand row) of the tile MMA, kind and CTA-group fixed at compile time.
template <bool TF32, bool PAIR>
__device__ __forceinline__ void umma_step(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (PAIR) {
        if constexpr (TF32) umma_tf32_2cta(tmem_d, desc_a, desc_b, idesc, accumulate);
        else umma_bf16_2cta(tmem_d, desc_a, desc_b, idesc, accumulate);
    } else {
        if constexpr (TF32) umma_tf32(tmem_d, desc_a, desc_b, idesc, accumulate);
        else umma_bf16(tmem_d, desc_a, desc_b, idesc, accumulate);
    }
}

// TMEM -> registers: lane i of the warp reads TMEM lane (lane_base + i), 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Asynchronous TMEM loads for software-pipelined epilogues: tmem_ld_async<W> issues tcgen05.ld (lane i <- TMEM lane base + i,
// W consecutive fp32 columns) WITHOUT waiting; tmem_ld_fence<W> is tcgen05.wait::ld with the destination registers as in/out
// operands, so the compiler cannot move a use of them above the wait.
template <int W>
__device__ __forceinline__ void tmem_ld_async(uint32_t taddr, uint32_t (&r)[32]);
template <>
__device__ __forceinline__ void tmem_ld_async<32>(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_async<16>(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_async<8>(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_async<4>(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
template <int W>
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&r)[32]);
template <>
__device__ __forceinline__ void tmem_ld_fence<32>(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                   "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),
                   "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),
                   "+r"(r[31])
                 :
                 : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_fence<16>(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                   "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_fence<8>(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) : : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_fence<4>(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) : : "memory");
}

// 2^x, single MUFU.EX2 (flush-to-zero: no denormal fix-up code around it)
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- small helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint16_t f32_to_bf16_rn(float x) {   // round-to-nearest-even: one F2FP instruction
    uint16_t h;
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(x));
    return h;
}
// two floats -> packed bf16x2 (`a` in the low half, `b` in the high half), one instruction
__device__ __forceinline__ uint32_t pack_bf16x2_rn(float a, float b) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
__device__ __forceinline__ float bf16_to_f32(uint16_t h) { return __uint_as_float(static_cast<uint32_t>(h) << 16); }
__device__ __forceinline__ void split_bf16(float x, uint16_t &hi, uint16_t &lo) {
    hi = f32_to_bf16_rn(x);
    lo = f32_to_bf16_rn(x - bf16_to_f32(hi));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

#endif  // __CUDACC__

}  // namespace loco
