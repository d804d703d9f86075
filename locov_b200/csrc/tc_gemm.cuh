// tc_gemm.cuh — the tcgen05 / TMEM / TMA tensor-core core shared by every dense contraction of the
// region-text path (projection, RoI x class scoring, LSM word<->region pair scoring).
//
// One CTA computes one or more 128 x BLOCK_N fp32 accumulator tiles ("chunks") in tensor memory:
//   warp 0   : TMA producer  — cp.async.bulk.tensor 2-D tiles (128B swizzle) of A [128 x 64] and
//              B [BLOCK_N x 64] bf16 into a `stages`-deep shared-memory ring, mbarrier complete_tx.
//   warp 1   : MMA issuer    — one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BLOCK_N,
//              K=16) x4 per stage, tcgen05.commit releases the stage / publishes the accumulator.
//              Also owns the TMEM allocation.
//   warps 2-5: epilogue      — tcgen05.ld 32x32b (lane = accumulator row) and an operation-specific
//              fused reduction (bias / softmax / log-sum-exp / argmax / attention pooling) supplied as the
//              `Epi` policy.  With two accumulator stages the epilogue of chunk i overlaps the MMAs
//              of chunk i+1.
// fp32-accurate mode = three bf16 passes (hi*hi, hi*lo, lo*hi) accumulated into the same TMEM tile.
// BLOCK_N, the pipeline depth and the pass count are run-time values (the UMMA instruction descriptor
// and the TMA boxes are built from them), so a single instantiation per epilogue serves all shapes.
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace loco {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;   // bf16 elements: 128 bytes = one swizzle span
constexpr int TC_THREADS = 192;          // 2 control warps + 4 epilogue warps (policies may ask for 8: Epi::kEpiWarps)
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;   // 16 KB

struct TcMaps {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    CUtensorMap o_hi, o_lo;      // output maps of policies that leave through TMA stores (Epi::kWantsMaps), unused otherwise
};

struct TcCore {
    int block_n;        // UMMA N: multiple of 16 in [16, 256]
    int num_k_blocks;   // ceil(K / 64); the K tail is zero-filled by TMA
    int passes;         // 1 = bf16, 3 = hi*hi + hi*lo + lo*hi
    int stages;         // shared-memory ring depth
    int chunks;         // accumulator tiles per CTA
    int acc_stages;     // TMEM accumulator buffers (1 or 2)
    int tmem_cols;      // allocation: power of two >= 32
    int epi_smem;       // bytes of epilogue scratch
    int epi_tail;       // bytes of epilogue scratch that never overlay the ring (they follow the barriers even with epi_overlay): buffers the
                        // policy fills while the tile's operands are still in flight
    // thread-block cluster with TMA multicast (cm x cn CTAs; 1 x 1 = no cluster): CTA (rm, rn) computes tile
    // (tile_m0 + rm, tile_n0 + rn); the A tile of a row is shared by its cn CTAs (each loads 1/cn of it and
    // multicasts), the B tile of a column by its cm CTAs — L2->SM operand traffic drops by the same factors.
    int cm, cn;
    int clusters_n;     // clusters along the virtual tile grid's n axis (virtual tiles_n = clusters_n * cn)
    int tf32;           // 1: operands are fp32 in memory and multiplied as TF32 (kind::tf32, 32 elements per 128-byte k block)
    int total_tiles;    // > 0: persistent mode — CTA b processes tiles b, b + grid, ... (chunks = its share); 0: `chunks` each
    int debug_mode;     // developer experiments only (LOCOV_B200_DEBUG): 1 = skip the MMAs (pure TMA ingest rate),
                        // 2 = skip the TMA loads (pure tensor-pipe + operand-read rate); results are garbage
    int ring_bytes;     // bytes reserved for the operand ring (>= stages * stage_bytes; barriers follow it)
    int two_cta;        // 1: CTA pair along M (cm = 2, cn = 1): cta_group::2 MMAs of M = 256, each CTA stages its own 128 rows of
                        // A and HALF of the B tile (no multicast); the leader CTA (rank 0) issues the MMAs
    int epi_overlay;    // 1: the epilogue scratch overlays the operand ring (legal only with chunks == 1: the ring is
                        // dead once the accumulator barrier fired — every load of this tile has landed and been consumed)
    unsigned long long *timeline;   // developer probe (LOCOV_B200_TIMELINE=1): per CTA 16 globaltimer stamps (0-6 set by the core, 8-15 free for the epilogue policy), NULL otherwise
    int prefetch;       // > 0: the producer prefetches the A panel into L2 in bursts of `prefetch` consecutive k blocks, one burst
                        // ahead of the loads: a 128-byte-wide k block touches every row's DRAM page for 128 bytes only; a burst
                        // turns that into `prefetch` x 128 contiguous bytes per row while the page is open
    int fuse3;          // set by tc_finalize for passes == 3 when the ring allows it: ONE stage holds the k block of A_hi, A_lo, B_hi and B_lo and
                        // feeds the three products hi*hi + hi*lo + lo*hi — each operand slice is loaded once per k block instead of once
                        // per pass (78 -> 52 KB of TMA fill per k block at N = 160: the 3-pass main loop is fill bound, not tensor bound)
    int single_wave;    // 1: the grid is at most one CTA per SM, so a second resident CTA would never exist — give the whole
                        // shared memory to the operand ring (bytes in flight per SM bound the TMA ingest rate: Little's law)
};

__host__ __device__ inline int tc_round_up(int x, int m) { return (x + m - 1) / m * m; }

inline int tc_tmem_cols(int block_n, int acc_stages) {
    const int need = (acc_stages - 1) * block_n + tc_round_up(block_n, 32);
    int c = 32;
    while (c < need) c <<= 1;
    return c;
}

// Fill stages / tmem / smem for a given block_n; returns dynamic shared memory bytes.
inline size_t tc_finalize(TcCore &core, int K, int passes, int chunks, int epi_smem) {
    const int k_elems = core.tf32 ? TC_BLOCK_K / 2 : TC_BLOCK_K;      // one 128-byte swizzle span per k block
    core.num_k_blocks = (K + k_elems - 1) / k_elems;
    core.passes = passes;
    core.chunks = chunks;
    core.acc_stages = chunks > 1 ? 2 : 1;
    if (core.acc_stages * core.block_n > 512) core.acc_stages = 1;
    core.tmem_cols = tc_tmem_cols(core.block_n, core.acc_stages);
    core.epi_smem = epi_smem;
    if (core.cm < 1) core.cm = 1;
    if (core.cn < 1) core.cn = 1;
    size_t stage_bytes = TC_A_BYTES + (size_t)(core.two_cta ? core.block_n / 2 : core.block_n) * 128;
    if (chunks != 1) core.epi_overlay = 0;
    const long long fixed = 1024 /*alignment slack*/ + 512 /*barriers*/ + core.epi_tail + (core.epi_overlay ? 0 : (long long)epi_smem);
    long long budget = core.single_wave ? 225 : 110;                                // two CTAs per SM unless the grid is one wave
    if (const char *e = getenv("LOCOV_B200_SMEMKB")) { const int v = atoi(e); if (v >= 32 && v <= 225) budget = v; }   // developer sweep knob
    // fused three-pass stages (see TcCore::fuse3): only without multicast clusters, and only when at least 3 double-width stages fit
    core.fuse3 = 0;
    static const bool fuse3_allowed = []() { const char *e = getenv("LOCOV_B200_FUSE3"); return !(e != nullptr && e[0] == '0'); }();
    if (passes == 3 && fuse3_allowed && !core.tf32 && (core.two_cta || core.cm * core.cn == 1) && !core.epi_overlay &&
        (budget * 1024 - fixed) / (long long)(2 * stage_bytes) >= 3) {
        core.fuse3 = 1;
        stage_bytes *= 2;
    }
    int stages = (int)((budget * 1024 - fixed) / (long long)stage_bytes);
    if (core.epi_overlay && (long long)epi_smem + fixed > 110 * 1024) stages = 0;    // the overlay itself needs a whole SM
    if (stages < 3) stages = (int)(((long long)225 * 1024 - fixed) / (long long)stage_bytes);
    const int total_iters = core.num_k_blocks * (core.fuse3 ? 1 : passes) * chunks;
    if (stages > (core.single_wave ? TC_MAX_STAGES : 6)) stages = core.single_wave ? TC_MAX_STAGES : 6;
    if (const char *e = getenv("LOCOV_B200_STAGES")) { const int v = atoi(e); if (v >= 1 && v < stages) stages = v; }   // developer sweep knob
    if (stages > total_iters) stages = total_iters;
    if (stages < 1) stages = 1;
    core.stages = stages;
    core.debug_mode = 0;
    if (const char *e = getenv("LOCOV_B200_DEBUG")) core.debug_mode = atoi(e);
    if (const char *e = getenv("LOCOV_B200_PF")) { const int v = atoi(e); if (v >= 0 && v <= 64) core.prefetch = v; }   // developer sweep knob
    size_t ring = (size_t)stages * stage_bytes;
    if (core.epi_overlay && ring < (size_t)epi_smem) ring = ((size_t)epi_smem + 1023) / 1024 * 1024;
    core.ring_bytes = (int)ring;
    return (size_t)fixed + ring;
}

#if defined(__CUDACC__)

// ---- epilogue store helpers (per-warp shared scratch of 32 x 33 words) ----------------------------------
constexpr int TC_WARP_SCRATCH_WORDS = 32 * 33;

// v = this lane's row (tile row = 32*q + lane), columns [0,32) of a 32x32 block.  Writes rows < rows_valid
// and columns < cols_valid to dst (row-major, ld) as fully coalesced 128-byte rows.
__device__ __forceinline__ void warp_store_f32(float *scratch, const float (&v)[32], float *dst, int64_t ld,
                                               int rows_valid, int cols_valid, int lane) {
#pragma unroll
    for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = v[j];
    __syncwarp();
    if (lane < cols_valid)
        for (int rr = 0; rr < rows_valid; ++rr) dst[(int64_t)rr * ld + lane] = scratch[rr * 33 + lane];
    __syncwarp();
}

// Same block written as bf16 hi (and optionally lo): two rows per instruction, 64 contiguous bytes each.
__device__ __forceinline__ void warp_store_bf16(uint32_t *scratch, const float (&v)[32], uint16_t *dst_hi,
                                                uint16_t *dst_lo, int64_t ld, int rows_valid, int cols_valid,
                                                int lane) {
    uint32_t lo_pack[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        uint16_t h0, l0, h1, l1;
        split_bf16(v[2 * j], h0, l0);
        split_bf16(v[2 * j + 1], h1, l1);
        scratch[lane * 17 + j] = (uint32_t)h0 | ((uint32_t)h1 << 16);
        lo_pack[j] = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
    __syncwarp();
    const int sub = lane >> 4, cc = lane & 15;
    for (int it = 0; it < 16; ++it) {
        const int rr = 2 * it + sub;
        if (rr < rows_valid && 2 * cc < cols_valid)
            *reinterpret_cast<uint32_t *>(dst_hi + (int64_t)rr * ld + 2 * cc) = scratch[rr * 17 + cc];
    }
    __syncwarp();
    if (dst_lo != nullptr) {
#pragma unroll
        for (int j = 0; j < 16; ++j) scratch[lane * 17 + j] = lo_pack[j];
        __syncwarp();
        for (int it = 0; it < 16; ++it) {
            const int rr = 2 * it + sub;
            if (rr < rows_valid && 2 * cc < cols_valid)
                *reinterpret_cast<uint32_t *>(dst_lo + (int64_t)rr * ld + 2 * cc) = scratch[rr * 17 + cc];
        }
        __syncwarp();
    }
}

// ---- MMA issue loop (one elected thread) ---------------------------------------------------------------
// The issuing thread is a single warp with no instruction-level parallelism: every instruction between two tcgen05.mma is
// latency the tensor pipe may have to wait for (measured: ~170-200 cycles per MMA with run-time kind / pair / debug branches
// and per-lane serialisation loops around each MMA, against an 80-128 cycle MMA).  Kind and CTA-group are therefore
// template parameters, the four K steps of a stage are straight-line code and the descriptors advance by immediates.
template <bool TF32, bool PAIR>
__device__ __forceinline__ void mma_issue_loop(const TcCore &core, unsigned char *smem, uint32_t stage_bytes, uint64_t *full,
                                               uint64_t *empty, uint64_t *tfull, uint64_t *tempty, uint32_t tmem_base, int nchunks,
                                               int iters_per_chunk, unsigned long long *tl) {
    const uint32_t mma_m = PAIR ? 2u * TC_BLOCK_M : (uint32_t)TC_BLOCK_M;
    const uint32_t idesc = TF32 ? umma_idesc_tf32(mma_m, (uint32_t)core.block_n) : umma_idesc_bf16(mma_m, (uint32_t)core.block_n);
    const bool skip_mma = core.debug_mode == 1;
    const bool multicast_commit = !PAIR && core.cm * core.cn > 1;
    uint16_t commit_mask = 1;
    if (multicast_commit) {
        const int rank = (int)cluster_ctarank();
        const int rm = rank / core.cn, rn = rank - rm * core.cn;
        uint32_t mb = ((1u << core.cn) - 1u) << (rm * core.cn);
        for (int j = 0; j < core.cm; ++j) mb |= 1u << (j * core.cn + rn);
        commit_mask = (uint16_t)mb;
    }
    const uint32_t smem_base = smem_u32(smem);
    uint32_t stage = 0, phase = 0;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int acc = ch % core.acc_stages;
        const uint32_t acc_phase = (uint32_t)(ch / core.acc_stages) & 1u;
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * core.block_n);
        uint32_t accumulate = 0;
        for (int it = 0; it < iters_per_chunk; ++it) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            if (tl != nullptr && it == 0 && ch == 0) tl[2] = global_timer_ns();     // first operand stage landed
            const uint32_t a_addr = smem_base + stage * stage_bytes;
            if (core.fuse3) {
                // stage = [A_hi][A_lo][B_hi][B_lo] of one k block: hi*hi + hi*lo + lo*hi per 16-wide K step
                const uint32_t bb = (stage_bytes - 2u * TC_A_BYTES) / 2u;
                const uint64_t da_h = umma_desc_k128(a_addr), da_l = umma_desc_k128(a_addr + TC_A_BYTES);
                const uint64_t db_h = umma_desc_k128(a_addr + 2u * TC_A_BYTES), db_l = umma_desc_k128(a_addr + 2u * TC_A_BYTES + bb);
                if (!skip_mma) {
#pragma unroll
                    for (uint32_t ks = 0; ks < 8u; ks += 2u) {
                        umma_step<TF32, PAIR>(tmem_d, da_h + ks, db_h + ks, idesc, ks == 0u ? accumulate : 1u);
                        umma_step<TF32, PAIR>(tmem_d, da_h + ks, db_l + ks, idesc, 1u);
                        umma_step<TF32, PAIR>(tmem_d, da_l + ks, db_h + ks, idesc, 1u);
                    }
                }
            } else {
            const uint64_t da = umma_desc_k128(a_addr);
            const uint64_t db = umma_desc_k128(a_addr + TC_A_BYTES);
            if (!skip_mma) {
                umma_step<TF32, PAIR>(tmem_d, da, db, idesc, accumulate);            // +32 B of K per step (descriptor units of 16 B)
                umma_step<TF32, PAIR>(tmem_d, da + 2u, db + 2u, idesc, 1u);
                umma_step<TF32, PAIR>(tmem_d, da + 4u, db + 4u, idesc, 1u);
                umma_step<TF32, PAIR>(tmem_d, da + 6u, db + 6u, idesc, 1u);
            }
            }
            accumulate = 1;
            if constexpr (PAIR) {
                umma_commit_2cta(&empty[stage]);
            } else {
                if (multicast_commit) umma_commit_mc(&empty[stage], commit_mask); else umma_commit(&empty[stage]);
            }
            if (++stage == (uint32_t)core.stages) { stage = 0; phase ^= 1u; }
        }
        if constexpr (PAIR) umma_commit_2cta(&tfull[acc]); else umma_commit(&tfull[acc]);
        if (tl != nullptr && ch == nchunks - 1) tl[3] = global_timer_ns();           // last MMA issued
    }
}

// ---- the kernel ------------------------------------------------------------------------------------
// Epi interface:
//   struct Params;                                       (trivially copyable, passed by value)
//   static __device__ void coords(const Params&, const TcCore&, int cta, int chunk, int &row_a, int &row_b);
//   __device__ void begin(const Params&, const TcCore&, int cta, int row, int lane, int q, unsigned char *smem);
//   __device__ void chunk(const Params&, const TcCore&, int cta, int chunk, uint32_t taddr, ...same...);
//   __device__ void finish(const Params&, const TcCore&, int cta, ...same...);
//   static constexpr bool kHasPrefetch;  true: prefetch(const Params&, const TcCore&, int cta, int chunk, row, lane, q, smem) runs before the
//                                        wait for the chunk's accumulator (masks / per-tile inputs load under the MMAs)
//   static constexpr bool kWantsMaps;    true: the policy stores through TMA: members map_hi / map_lo (const CUtensorMap *) are set to TcMaps::o_hi / o_lo
//   static constexpr bool kSelfRelease;  true: chunk() itself arrives on the accumulator-empty barrier (members release_bar / release_local)
// PAIR = true is the CTA-pair (cta_group::2) instantiation: a kernel containing cta_group::2 instructions can only be
// launched as a cluster of two, so it is a separate instantiation from the single-CTA / multicast-cluster one.
template <class Epi, bool PAIR = false>
__global__ void __launch_bounds__(64 + 32 * Epi::kEpiWarps, Epi::kMinBlocks)
    tc_gemm_kernel(const __grid_constant__ TcMaps maps, const TcCore core, const typename Epi::Params ep) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);

    constexpr bool pair = PAIR;                               // CTA pair: this CTA stages half of the B tile
    const uint32_t b_bytes = (uint32_t)(pair ? core.block_n / 2 : core.block_n) * 128u;
    const uint32_t stage_bytes = (TC_A_BYTES + b_bytes) * (core.fuse3 ? 2u : 1u);
    unsigned char *bar_base = smem + (size_t)core.ring_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(bar_base);
    uint64_t *empty = full + TC_MAX_STAGES;
    uint64_t *tfull = empty + TC_MAX_STAGES;
    uint64_t *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    unsigned char *epi_smem = core.epi_overlay ? smem : bar_base + 512;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int k_elems = core.tf32 ? TC_BLOCK_K / 2 : TC_BLOCK_K;
    unsigned long long *tl = core.timeline ? core.timeline + (size_t)blockIdx.x * 16 : nullptr;
    if (tl && threadIdx.x == 0) tl[0] = global_timer_ns();                       // CTA entry

    // cluster geometry -> virtual CTA index handed to the epilogue policy
    const int csize = core.cm * core.cn;
    int rm = 0, rn = 0, cta = blockIdx.x;
    uint16_t mask_a = 1, mask_b = 1;
    if (csize > 1) {
        const int rank = (int)cluster_ctarank();
        rm = rank / core.cn;
        rn = rank - rm * core.cn;
        const int cluster_id = blockIdx.x / csize;
        const int tile_m = (cluster_id / core.clusters_n) * core.cm + rm;
        const int tile_n = (cluster_id % core.clusters_n) * core.cn + rn;
        cta = tile_m * (core.clusters_n * core.cn) + tile_n;
        mask_a = (uint16_t)(((1u << core.cn) - 1u) << (rm * core.cn));          // CTAs (rm, *)
        uint32_t mb = 0;
        for (int j = 0; j < core.cm; ++j) mb |= 1u << (j * core.cn + rn);        // CTAs (*, rn)
        mask_b = (uint16_t)mb;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < core.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], pair ? 1u : (uint32_t)(core.cm + core.cn - 1));   // one release per CTA that reads what this CTA loads
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], (pair ? 64u : 32u) * Epi::kEpiWarps);          // pair: the epilogue threads of both CTAs
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a_hi);
        prefetch_tmap(&maps.b_hi);
        if (core.passes > 1) {
            prefetch_tmap(&maps.a_lo);
            prefetch_tmap(&maps.b_lo);
        }
    }
    if (warp == 1) {
        if constexpr (pair) tmem_alloc2(tmem_slot, (uint32_t)core.tmem_cols);
        else tmem_alloc(tmem_slot, (uint32_t)core.tmem_cols);
    }
    tc_fence_before();
    __syncthreads();                                               // TMEM address / barrier words visible inside the CTA
    if (csize > 1) cluster_sync_relaxed();                         // peers' barriers must be initialised (fence.mbarrier_init above) before any remote signal
    tc_fence_after();
    pdl_trigger();                                                                // the next kernel may start its own set-up
    pdl_wait();                                                                   // operands / masks of the previous kernel are complete
    if (tl && threadIdx.x == 0) tl[1] = global_timer_ns();                       // setup done (barriers, TMEM, cluster sync)
    const uint32_t tmem_base = *tmem_slot;
    const int iters_per_chunk = core.num_k_blocks * (core.fuse3 ? 1 : core.passes);
    // persistent mode: unit u (a CTA, or a CTA pair) of U launched units processes tiles u, u + U, ...
    const int unit = pair ? (int)blockIdx.x / 2 : (int)blockIdx.x, units = pair ? (int)gridDim.x / 2 : (int)gridDim.x;
    const int nchunks = core.total_tiles > 0 ? max(0, (core.total_tiles - unit + units - 1) / units) : core.chunks;

    if (warp == 0) {
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int ch = 0; ch < nchunks; ++ch) {
                int row_a, row_b;
                Epi::coords(ep, core, cta, ch, row_a, row_b);
                for (int pass = 0; pass < (core.fuse3 ? 1 : core.passes); ++pass) {
                    const CUtensorMap *ma = (pass == 2) ? &maps.a_lo : &maps.a_hi;
                    const CUtensorMap *mb = (pass == 1) ? &maps.b_lo : &maps.b_hi;
                    for (int kb = 0; kb < core.num_k_blocks; ++kb) {
                        if (core.prefetch > 0 && kb % core.prefetch == 0 && core.debug_mode != 2) {
                            const int a_rows_pf = pair ? TC_BLOCK_M : TC_BLOCK_M / core.cn;      // rows this CTA itself loads
                            const int row_pf = pair ? row_a : row_a + rn * a_rows_pf;
                            const int kb0 = kb + core.prefetch;
                            for (int j = kb0; j < kb0 + core.prefetch && j < core.num_k_blocks; ++j) tma_prefetch_2d(ma, j * k_elems, row_pf);
                        }
                        if (core.fuse3) {          // one stage = the k block of all four operand matrices (csize == 1 or CTA pair)
                            mbar_wait(&empty[stage], phase ^ 1u);
                            unsigned char *s4 = smem + (size_t)stage * stage_bytes;
                            if (core.debug_mode == 2) {
                                mbar_arrive(&full[stage]);
                            } else if constexpr (pair) {
                                if (rm == 0) mbar_arrive_expect_tx(&full[stage], 2u * stage_bytes);
                                const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
                                const int rb = row_b + rm * (core.block_n / 2);
                                tma_load_2d_2sm(s4, &maps.a_hi, lead_full, kb * k_elems, row_a);
                                tma_load_2d_2sm(s4 + TC_A_BYTES, &maps.a_lo, lead_full, kb * k_elems, row_a);
                                tma_load_2d_2sm(s4 + 2 * TC_A_BYTES, &maps.b_hi, lead_full, kb * k_elems, rb);
                                tma_load_2d_2sm(s4 + 2 * TC_A_BYTES + b_bytes, &maps.b_lo, lead_full, kb * k_elems, rb);
                            } else {
                                mbar_arrive_expect_tx(&full[stage], stage_bytes);
                                tma_load_2d(s4, &maps.a_hi, &full[stage], kb * k_elems, row_a);
                                tma_load_2d(s4 + TC_A_BYTES, &maps.a_lo, &full[stage], kb * k_elems, row_a);
                                tma_load_2d(s4 + 2 * TC_A_BYTES, &maps.b_hi, &full[stage], kb * k_elems, row_b);
                                tma_load_2d(s4 + 2 * TC_A_BYTES + b_bytes, &maps.b_lo, &full[stage], kb * k_elems, row_b);
                            }
                            if (++stage == (uint32_t)core.stages) { stage = 0; phase ^= 1u; }
                            continue;
                        }
                        mbar_wait(&empty[stage], phase ^ 1u);
                        unsigned char *sa = smem + (size_t)stage * stage_bytes;
                        if (core.debug_mode == 2) {
                            mbar_arrive(&full[stage]);
                        } else {
                        if constexpr (pair) {
                            // both CTAs signal the LEADER's full barrier; the leader announces the bytes of both
                            if (rm == 0) mbar_arrive_expect_tx(&full[stage], 2u * stage_bytes);
                            const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
                            tma_load_2d_2sm(sa, ma, lead_full, kb * k_elems, row_a);
                            tma_load_2d_2sm(sa + TC_A_BYTES, mb, lead_full, kb * k_elems, row_b + rm * (core.block_n / 2));
                        } else {
                        mbar_arrive_expect_tx(&full[stage], stage_bytes);
                        if (csize == 1) {
                            tma_load_2d(sa, ma, &full[stage], kb * k_elems, row_a);
                            tma_load_2d(sa + TC_A_BYTES, mb, &full[stage], kb * k_elems, row_b);
                        } else {
                            // this CTA's slice of the shared tiles, delivered to every CTA of the row / column
                            const int a_rows = TC_BLOCK_M / core.cn, b_rows = core.block_n / core.cm;
                            tma_load_2d_mc(sa + (size_t)rn * a_rows * 128, ma, &full[stage], kb * k_elems, row_a + rn * a_rows, mask_a);
                            tma_load_2d_mc(sa + TC_A_BYTES + (size_t)rm * b_rows * 128, mb, &full[stage], kb * k_elems,
                                           row_b + rm * b_rows, mask_b);
                        }
                        }
                        }
                        if (++stage == (uint32_t)core.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (!(pair && rm != 0) && elect_one()) {              // CTA pair: only the leader issues
            if (core.tf32) mma_issue_loop<true, pair>(core, smem, stage_bytes, full, empty, tfull, tempty, tmem_base, nchunks, iters_per_chunk, tl);
            else mma_issue_loop<false, pair>(core, smem, stage_bytes, full, empty, tfull, tempty, tmem_base, nchunks, iters_per_chunk, tl);
        }
    } else {
        const int q = warp & 3;            // TMEM lane quarter accessible to this warp
        const int row = q * 32 + lane;     // accumulator row owned by this thread
        Epi epi;
        if constexpr (Epi::kWantsMaps) { epi.map_hi = &maps.o_hi; epi.map_lo = &maps.o_lo; }
        epi.begin(ep, core, cta, row, lane, q, epi_smem);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int acc = ch % core.acc_stages;
            const uint32_t acc_phase = (uint32_t)(ch / core.acc_stages) & 1u;
            if constexpr (Epi::kHasPrefetch) epi.prefetch(ep, core, cta, ch, row, lane, q, epi_smem);   // per-tile inputs of the epilogue, loaded
                                                                                                          // while the tile's MMAs still run
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            if (tl && ch == nchunks - 1 && threadIdx.x == 64) tl[4] = global_timer_ns();   // last accumulator complete
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * core.block_n);
            if constexpr (Epi::kSelfRelease) {
                // the policy frees the accumulator itself as soon as its last tcgen05.ld has completed (every epilogue thread
                // arrives exactly once per chunk), so the MMAs of the tile after next can start under the rest of the epilogue
                epi.release_bar = pair ? mapa_u32(smem_u32(&tempty[acc]), 0) : 0u;
                epi.release_local = &tempty[acc];
                epi.chunk(ep, core, cta, ch, taddr, row, lane, q, epi_smem);
            } else {
                epi.chunk(ep, core, cta, ch, taddr, row, lane, q, epi_smem);
                tc_fence_before();
                if constexpr (pair) mbar_arrive_cluster_cta(mapa_u32(smem_u32(&tempty[acc]), 0));   // the leader's MMA thread waits for both CTAs
                else mbar_arrive(&tempty[acc]);
            }
        }
        epi.finish(ep, core, cta, row, lane, q, epi_smem);
        if (tl && threadIdx.x == 64) tl[5] = global_timer_ns();                  // epilogue done
    }
    tc_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_relaxed();                         // no CTA may exit while peers still signal its barriers
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        if constexpr (pair) tmem_dealloc2(tmem_base, (uint32_t)core.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)core.tmem_cols);
    }
    if (tl && threadIdx.x == 0) tl[6] = global_timer_ns();                       // CTA exit
}

unsigned long long *debug_timeline_buffer(int ctas);   // api.cu; NULL unless LOCOV_B200_TIMELINE=1

template <class Epi, bool PAIR = false>
int tc_launch(const TcMaps &maps, const TcCore &core_in, const typename Epi::Params &ep, int grid, size_t smem_bytes,
              cudaStream_t st) {
    TcCore core = core_in;
    core.timeline = debug_timeline_buffer(grid);
    LOCO_REQUIRE(smem_bytes <= 227 * 1024, LOCO_E_UNSUPPORTED, "tensor-core kernel needs %zu B of shared memory", smem_bytes);
    LOCO_REQUIRE(core.block_n >= 16 && core.block_n <= 256 && core.block_n % 16 == 0, LOCO_E_BADARG, "bad block_n %d", core.block_n);
    LOCO_REQUIRE(core.tmem_cols <= 512, LOCO_E_UNSUPPORTED, "tensor memory request %d columns", core.tmem_cols);
    LOCO_REQUIRE(PAIR == (core.two_cta != 0) && (!PAIR || (core.cm == 2 && core.cn == 1)), LOCO_E_BADARG, "CTA-pair launch mismatch");
    static thread_local size_t configured = 0;   // per (thread, Epi, PAIR) high-water mark
    if (smem_bytes > configured) {
        LOCO_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<Epi, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        configured = 227 * 1024;
    }
    const int csize = core.cm * core.cn;
    if (csize > 1) {
        LOCO_REQUIRE(csize <= 8 && grid % csize == 0, LOCO_E_BADARG, "bad cluster %dx%d for grid %d", core.cm, core.cn, grid);
        LOCO_REQUIRE(TC_BLOCK_M % core.cn == 0 && (TC_BLOCK_M / core.cn) % 8 == 0 && core.block_n % core.cm == 0 && (core.block_n / core.cm) % 8 == 0,
                     LOCO_E_BADARG, "cluster %dx%d does not slice a 128x%d tile on 8-row boundaries", core.cm, core.cn, core.block_n);
    }
    LOCO_CUDA(launch_kernel(tc_gemm_kernel<Epi, PAIR>, dim3((unsigned)grid), dim3(64 + 32 * Epi::kEpiWarps), smem_bytes, st, csize, maps, core, ep));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

#endif  // __CUDACC__

}  // namespace loco
