// tc_ops.cu — the dense contractions of the region-text path on the tcgen05 core (tc_gemm.cuh), each
// with its reductions fused into the TMEM epilogue:
//   EpiLinear : out = A·Wᵀ + bias                  (emb_pred / bbox_pred / v2l_projection)
//   EpiScore  : logits = E·Cᵀ (+bias), online softmax statistics, argmax, probabilities
//   EpiLsm    : caption×image similarity tile with BOTH attention poolings (word→region, region→word)
//               from one GEMM; the backward variant emits dS for the dEmb / dCap GEMMs
#include <cfloat>
#include <cstdlib>

#include "tc_gemm.cuh"

namespace loco {

constexpr float LSM_FILL = -1.0e30f;
constexpr float LSM_PAD = -3.0e38f;    // padding columns of a tile: exp2(LSM_PAD - anything finite) == 0   // finite "masked" logit: an all-masked row stays uniform, never NaN

// ------------------------------------------------------------------------------------------------
// EpiLinear
// ------------------------------------------------------------------------------------------------
struct EpiLinear {
    static constexpr int kEpiWarps = 4;
    static constexpr int kMinBlocks = 1;
    static constexpr bool kHasPrefetch = false, kSelfRelease = false, kWantsMaps = true;
    static constexpr int kStageBytes = 4 * 2 * 2048 + 512;       // tma_out: per epilogue warp one 32 x 32 bf16 block for hi and one for lo (+ alignment)
    const CUtensorMap *map_hi, *map_lo;
    struct Params {
        const float *bias;
        float *out_f32;
        int64_t ld_f32;
        uint16_t *out_hi, *out_lo;
        int n_bf16;
        int64_t ld_bf16;
        int M, N, tiles_n;
        int grid;        // CTAs launched: persistent tile striding (tile = cta + chunk * grid)
        int vec_f32;     // out_f32 rows are 16-byte aligned: 128-bit stores
        int tma_out;     // bf16 outputs leave through shared memory + TMA tensor stores (TcMaps::o_hi / o_lo)
        int pair_units;  // > 0: persistent CTA pairs — pair u of `pair_units` owns the 256-row tiles u, u + pair_units, ... of the
                         // (row pair, column) grid; `cta` is the virtual index of this CTA's 128-row tile of the first one
    };
    // virtual 128-row tile (row tile * tiles_n + column tile) of this CTA's chunk `ch`
    static __device__ __forceinline__ int tile_of(const Params &p, int cta, int ch) {
        if (p.pair_units == 0) return cta + ch * p.grid;
        const int tm0 = cta / p.tiles_n, tn0 = cta - tm0 * p.tiles_n;
        const int u = (tm0 >> 1) * p.tiles_n + tn0 + ch * p.pair_units;
        const int um = u / p.tiles_n;
        return (2 * um + (tm0 & 1)) * p.tiles_n + (u - um * p.tiles_n);
    }
    static __device__ __forceinline__ void coords(const Params &p, const TcCore &core, int cta, int ch, int &row_a, int &row_b) {
        const int tile = tile_of(p, cta, ch);
        row_a = (tile / p.tiles_n) * TC_BLOCK_M;
        row_b = (tile % p.tiles_n) * core.block_n;
    }
    __device__ __forceinline__ void begin(const Params &, const TcCore &, int, int, int, int, unsigned char *) {}
    // Each thread owns one accumulator row: 32 consecutive columns per tcgen05.ld = 128 contiguous bytes of fp32 (or
    // 64 of bf16) in the row-major output, written with 128-bit stores straight from registers — whole 32-byte
    // sectors per thread, no shared-memory transpose, ~12 instructions per 32 outputs.
    __device__ __forceinline__ void chunk(const Params &p, const TcCore &core, int cta, int ch, uint32_t taddr, int row, int lane,
                                          int q, unsigned char *smem) {
        const int tile = tile_of(p, cta, ch);
        const int grow = (tile / p.tiles_n) * TC_BLOCK_M + row;
        const int n0 = (tile % p.tiles_n) * core.block_n;
        const bool row_ok = grow < p.M;
        for (int c0 = 0; c0 < core.block_n; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            const int gc0 = n0 + c0;
            const int cols_valid = max(0, min(min(32, core.block_n - c0), p.N - gc0));
            if (cols_valid <= 0) continue;     // warp-uniform
            if (p.bias != nullptr) {
                const float bl = (lane < cols_valid) ? __ldg(p.bias + gc0 + lane) : 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bl, j);
            }
            if (p.out_f32 != nullptr && row_ok) {
                float *dst = p.out_f32 + (int64_t)grow * p.ld_f32 + gc0;
                if (p.vec_f32 && cols_valid == 32) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < cols_valid) dst[j] = v[j];
                }
            }
            if (p.out_hi != nullptr && gc0 < p.n_bf16) {
                const int cb = min(cols_valid, p.n_bf16 - gc0);
                uint32_t hp[16], lp[16];
                const bool want_lo = p.out_lo != nullptr;          // warp-uniform
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    hp[j] = pack_bf16x2_rn(v[2 * j], v[2 * j + 1]);
                    lp[j] = 0u;
                    if (want_lo)
                        lp[j] = pack_bf16x2_rn(v[2 * j] - __uint_as_float(hp[j] << 16), v[2 * j + 1] - __uint_as_float(hp[j] & 0xFFFF0000u));
                }
                if (p.tma_out && core.block_n - c0 >= 32) {     // (a narrower last block of the tile would spill into the neighbouring tile's columns: direct stores)
                    // Through the copy engine: the direct stores below touch 32 different 128-byte lines per instruction (lane = row),
                    // ~32 LSU cycles each; here the warp's 32 x 32 block is staged with conflict-free 16-byte shared stores (64-byte
                    // swizzle of the output map) and one thread issues a tensor store per matrix.  Rows past M and columns past
                    // n_bf16 are clipped by the map.  One buffer per warp: the wait for the previous block's read sits behind
                    // this block's TMEM load and conversion.
                    unsigned char *stage = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem) + 511) & ~(uintptr_t)511) + q * 4096;
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                    const int sw = (lane >> 1) & 3;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int pos = (c ^ sw) * 16 + lane * 64;
                        *reinterpret_cast<uint4 *>(stage + pos) = make_uint4(hp[4 * c], hp[4 * c + 1], hp[4 * c + 2], hp[4 * c + 3]);
                        if (want_lo) *reinterpret_cast<uint4 *>(stage + 2048 + pos) = make_uint4(lp[4 * c], lp[4 * c + 1], lp[4 * c + 2], lp[4 * c + 3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const int r0 = (tile / p.tiles_n) * TC_BLOCK_M + q * 32;
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(reinterpret_cast<uint64_t>(map_hi)), "r"(smem_u32(stage)), "r"(gc0), "r"(r0) : "memory");
                        if (want_lo)
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(reinterpret_cast<uint64_t>(map_lo)), "r"(smem_u32(stage + 2048)), "r"(gc0), "r"(r0) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    continue;
                }
                if (!row_ok) continue;
                uint16_t *dh = p.out_hi + (int64_t)grow * p.ld_bf16 + gc0;     // 64-byte aligned: base 16 B, ld % 8, gc0 % 32
                uint16_t *dl = p.out_lo ? p.out_lo + (int64_t)grow * p.ld_bf16 + gc0 : nullptr;
                if (cb == 32) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        *reinterpret_cast<uint4 *>(dh + 2 * j) = make_uint4(hp[j], hp[j + 1], hp[j + 2], hp[j + 3]);
                        if (dl) *reinterpret_cast<uint4 *>(dl + 2 * j) = make_uint4(lp[j], lp[j + 1], lp[j + 2], lp[j + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < cb) {
                            dh[j] = (uint16_t)((j & 1) ? (hp[j >> 1] >> 16) : (hp[j >> 1] & 0xFFFFu));
                            if (dl) dl[j] = (uint16_t)((j & 1) ? (lp[j >> 1] >> 16) : (lp[j >> 1] & 0xFFFFu));
                        }
                }
            }
        }
    }
    __device__ __forceinline__ void finish(const Params &p, const TcCore &, int, int, int lane, int, unsigned char *) {
        if (p.tma_out && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the staging buffers die with the CTA
    }
};

// ------------------------------------------------------------------------------------------------
// EpiScore — RoI x class logits with online softmax statistics across N-chunks
// ------------------------------------------------------------------------------------------------
struct EpiScore {
    static constexpr int kEpiWarps = 4;
    static constexpr int kMinBlocks = 1;
    static constexpr bool kHasPrefetch = false, kSelfRelease = false, kWantsMaps = false;
    struct Params {
        const float *bias;
        float *logits;
        int64_t ld;
        float *lse;
        int64_t *argmax_fg;
        float *part;    // groups > 1: per (row, group) partial statistics {max, sum, best value, best index} for box_score_finalize_kernel
        int R, K1;
        int groups;     // CTAs per 128-row tile: CTA (tile, g) scores the column chunks g, g + groups, ...
        int vec;        // logits rows are 16-byte aligned (base and ld): 128-bit stores straight from registers
    };
    float run_max, run_sum, best_val;
    int best_idx;

    static __device__ __forceinline__ void coords(const Params &p, const TcCore &core, int cta, int chunk, int &row_a, int &row_b) {
        row_a = (cta / p.groups) * TC_BLOCK_M;
        row_b = (cta % p.groups + chunk * p.groups) * core.block_n;      // past K1: TMA zero fill, skipped by the epilogue
    }
    static constexpr int kBiasCap = 4096;   // class-bias values staged in shared memory (wider lists read it from global)
    __device__ __forceinline__ void begin(const Params &p, const TcCore &, int, int row, int, int, unsigned char *smem) {
        run_max = -FLT_MAX;
        run_sum = 0.f;
        best_val = -FLT_MAX;
        best_idx = 0;
        if (p.bias != nullptr) {
            // the class bias is staged once per CTA: a global load per 32-column block in front of the dependent
            // shuffles (L2 latency x 8 blocks x chunks) dominated this epilogue
            float *sb = reinterpret_cast<float *>(smem) + 4 * TC_WARP_SCRATCH_WORDS;
            for (int c = row; c < min(p.K1, kBiasCap); c += 128) sb[c] = __ldg(p.bias + c);
            asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps only
        }
    }
    __device__ __forceinline__ void chunk(const Params &p, const TcCore &core, int cta, int ch, uint32_t taddr, int row, int lane,
                                          int q, unsigned char *smem) {
        float *scratch = reinterpret_cast<float *>(smem) + (size_t)q * TC_WARP_SCRATCH_WORDS;
        const float *sb = reinterpret_cast<const float *>(smem) + 4 * TC_WARP_SCRATCH_WORDS;
        const int tile = cta / p.groups;
        const int m0 = tile * TC_BLOCK_M + q * 32;
        const int rows_valid = max(0, min(32, p.R - m0));
        const int n0 = (cta % p.groups + ch * p.groups) * core.block_n;
        if (n0 >= p.K1) return;
        for (int c0 = 0; c0 < core.block_n; c0 += 32) {
            const int gc0 = n0 + c0;
            const int cols_valid = max(0, min(min(32, core.block_n - c0), p.K1 - gc0));
            if (cols_valid <= 0) break;
            float v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            if (p.bias != nullptr) {           // one coalesced load per block, broadcast by shuffles (32 dependent global loads per
                                               // block made this epilogue latency bound)
                const int bc = gc0 + lane;
                const float bl = (lane < cols_valid) ? (bc < kBiasCap ? sb[bc] : __ldg(p.bias + bc)) : 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bl, j);
            }
            float blk_max = -FLT_MAX;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < cols_valid) {
                    blk_max = fmaxf(blk_max, v[j]);
                    if (gc0 + j < p.K1 - 1 && v[j] > best_val) {   // strict >: first maximum wins (torch.argmax)
                        best_val = v[j];
                        best_idx = gc0 + j;
                    }
                }
            }
            const float new_max = fmaxf(run_max, blk_max);
            float s = run_sum * __expf(run_max - new_max);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < cols_valid) s += __expf(v[j] - new_max);
            run_max = new_max;
            run_sum = s;
            if (p.vec && cols_valid == 32) {
                if (row < rows_valid + q * 32) {
                    float *dst = p.logits + (int64_t)(tile * TC_BLOCK_M + row) * p.ld + gc0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            } else {
                warp_store_f32(scratch, v, p.logits + (int64_t)m0 * p.ld + gc0, p.ld, rows_valid, cols_valid, lane);
            }
        }
    }
    __device__ __forceinline__ void finish(const Params &p, const TcCore &, int cta, int row, int, int, unsigned char *) {
        const int grow = (cta / p.groups) * TC_BLOCK_M + row;
        if (grow >= p.R) return;
        if (p.groups == 1) {
            if (p.lse != nullptr) p.lse[grow] = run_max + logf(run_sum);
            if (p.argmax_fg != nullptr) p.argmax_fg[grow] = (int64_t)best_idx;
        } else {
            float4 *dst = reinterpret_cast<float4 *>(p.part) + (int64_t)grow * p.groups + (cta % p.groups);
            *dst = make_float4(run_max, run_sum, best_val, __int_as_float(best_idx));
        }
    }
};

// Second (bandwidth) pass of the scoring: one warp per RoI row combines the per-group statistics (groups > 1) into the
// log-sum-exp and the foreground argmax, and streams out the softmax probabilities exp(logit - lse) when requested —
// coalesced 128-bit rows, every SM busy (the GEMM epilogue that did this row by row per thread was latency bound).
__global__ void __launch_bounds__(256) box_score_finalize_kernel(const float *__restrict__ logits, float *__restrict__ probs, int64_t ld,
                                                                 int R, int K1, const float4 *__restrict__ part, int groups,
                                                                 float *__restrict__ lse, int64_t *__restrict__ argmax_fg) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < R; row += gridDim.x * wpb) {
        float l;
        if (part != nullptr) {
            float4 st = make_float4(-FLT_MAX, 0.f, -FLT_MAX, 0.f);
            if (lane < groups) st = part[(int64_t)row * groups + lane];
            float mx = st.x;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sm = st.y * __expf(st.x - mx);
            float bv = st.z;
            int bi = __float_as_int(st.w);
            if (lane >= groups) bi = 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sm += __shfl_xor_sync(0xffffffffu, sm, o);
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }    // first maximum (smallest column) wins
            }
            l = mx + logf(sm);
            if (lane == 0) {
                if (lse != nullptr) lse[row] = l;
                if (argmax_fg != nullptr) argmax_fg[row] = (int64_t)(bv == -FLT_MAX ? 0 : bi);
            }
        } else {
            l = lse[row];
        }
        if (probs != nullptr) {
            const float *src = logits + (int64_t)row * ld;
            float *dst = probs + (int64_t)row * ld;
            if ((ld & 3) == 0 && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(probs)) & 15) == 0) {
                const int k4 = K1 >> 2;
                for (int c = lane; c < k4; c += 32) {
                    const float4 v = *reinterpret_cast<const float4 *>(src + 4 * c);
                    *reinterpret_cast<float4 *>(dst + 4 * c) = make_float4(__expf(v.x - l), __expf(v.y - l), __expf(v.z - l), __expf(v.w - l));
                }
                for (int c = 4 * k4 + lane; c < K1; c += 32) dst[c] = __expf(src[c] - l);
            } else {
                for (int c = lane; c < K1; c += 32) dst[c] = __expf(src[c] - l);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LSM pair epilogue (forward and backward share it).
//
// One CTA owns the similarity tile of `per_tile` captions x ONE image: rows = caption words
// (row = cl*T + t), columns = the image's regions.  After the MMAs the tile is scaled by 1/temperature
// and parked in shared memory (row stride odd => conflict-free row- and column-wise walks), so BOTH
// softmax directions of grounding_head.py:162-166 are evaluated from a single GEMM:
//   row phase    (thread = word row)          : softmax over regions  -> attention-pooled f_t   (w2r)
//   column phase (thread = (caption, region)) : softmax over the caption's T words -> h_r       (r2w)
// followed by fixed-order warp reductions (deterministic, no atomics) to the two scalars of each pair.
// BWD additionally turns the tile into dS = d(loss)/d(raw similarity) in place and streams it out
// transposed (dS^T [Bi*Rg, Bc*T], the A operand of dEmb = dS^T . cap) and optionally as dS (for dCap).
// ------------------------------------------------------------------------------------------------
constexpr int LSM_MAX_PER_TILE = 16;

struct LsmParams {
    const float *cap_mask;   // [Bc, T]
    const float *reg_mask;   // [Bi, Rg]
    float *out_w2r, *out_r2w;   // [Bc, Bi] (ld_out); forward outputs, each may be null
    int64_t ld_out;
    const float *g_w2r, *g_r2w; // [Bc, Bi] (ld_g); upstream gradients (BWD), each may be null
    int64_t ld_g;
    uint16_t *dst_hi, *dst_lo;  // dS^T [Bi*Rg, Bc*T] (ld_dst) bf16 hi / lo (lo may be null)
    int64_t ld_dst;
    uint16_t *ds_hi, *ds_lo;    // dS [Bc*T, Bi*Rg] (ld_ds), optional
    int64_t ld_ds;
    int Bc, T, Bi, Rg;
    int Bi_pad;                 // images padded to the cluster width: virtual CTA index = g * Bi_pad + i
    int per_tile;               // captions per 128-row tile
    int lds;                    // shared-memory row stride of the tile (floats, odd)
    float inv_temp;
    int hardmax;
};

// shared-memory carve-up (floats) after the [128][lds] tile
struct LsmSmem {
    float *S, *rbias, *cbias, *rowval, *colval, *pmax, *pden, *pnum, *rowmx, *rowden, *rowf, *colmx, *colden, *colh, *capnw;
    __device__ __forceinline__ LsmSmem(unsigned char *base, int lds, int block_n, int per_tile, int Rg, bool bwd) {
        float *p = reinterpret_cast<float *>(base);
        S = p; p += 128 * lds;
        rbias = p; p += block_n;          // 0 for a valid region, LSM_FILL for a masked one (additive mask)
        cbias = p; p += 128;              // same per caption-word row
        rowval = p; p += 128;
        colval = p; p += per_tile * Rg;
        pmax = p; p += 256;               // [2][128] partial row statistics of the two column halves
        pden = p; p += 256;
        pnum = p; p += 256;
        capnw = p; p += LSM_MAX_PER_TILE;
        rowmx = rowden = rowf = colmx = colden = colh = nullptr;
        if (bwd) {
            rowmx = p; p += 128;
            rowden = p; p += 128;
            rowf = p; p += 128;
            colmx = p; p += per_tile * Rg;
            colden = p; p += per_tile * Rg;
            colh = p; p += per_tile * Rg;
        }
    }
};
static size_t lsm_epi_smem_bytes(int lds, int block_n, int per_tile, int Rg, bool bwd) {
    size_t f = (size_t)128 * lds + block_n + 128 + 128 + (size_t)per_tile * Rg + 3 * 256 + LSM_MAX_PER_TILE;
    if (bwd) f += 3 * 128 + 3 * (size_t)per_tile * Rg;
    return f * sizeof(float) + 16;
}

constexpr float LSM_LOG2E = 1.4426950408889634f;
constexpr float LSM_LN2 = 0.6931471805599453f;

// Eight epilogue warps: warps w and w + 4 own the same 32 accumulator rows (TMEM lane quarter w % 4) and split the
// columns in interleaved 32-column blocks; the column / reduction phases run from shared memory on all 256 threads.
// The tile is parked in log2 units (raw * inv_temp * log2 e) so that every exponential is one FADD + one MUFU.EX2.
template <bool BWD>
struct EpiLsm {
    static constexpr int kEpiWarps = 8;
    static constexpr int kMinBlocks = 2;      // two CTAs per SM: one tile's epilogue overlaps the other's loads / MMAs
    static constexpr bool kHasPrefetch = false, kSelfRelease = false, kWantsMaps = false;
    static constexpr int kThreads = 32 * kEpiWarps;
    typedef LsmParams Params;
    static __device__ __forceinline__ void coords(const Params &p, const TcCore &, int cta, int, int &row_a, int &row_b) {
        const int g = cta / p.Bi_pad, i = cta - g * p.Bi_pad;
        row_a = g * p.per_tile * p.T;    // caption word rows (A operand)
        row_b = i * p.Rg;                // region rows of image i (B operand)
    }
    __device__ __forceinline__ void begin(const Params &, const TcCore &, int, int, int, int, unsigned char *) {}
    __device__ __forceinline__ void finish(const Params &, const TcCore &, int, int, int, int, unsigned char *) {}

    __device__ __forceinline__ void chunk(const Params &p, const TcCore &core, int cta, int, uint32_t taddr, int row, int lane,
                                          int q, unsigned char *smem) {
        const LsmSmem sm(smem, p.lds, core.block_n, p.per_tile, p.Rg, BWD);
        const int g = cta / p.Bi_pad, i = cta - g * p.Bi_pad;
        if (i >= p.Bi || g * p.per_tile >= p.Bc) return;            // padding tile of the cluster grid (CTA-uniform)
        if (core.debug_mode == 3) return;                           // developer timing experiments (LOCOV_B200_DEBUG)
        const int et = threadIdx.x - 64;                            // 0..255 among the epilogue threads
        const int ew = et >> 5;                                     // epilogue warp 0..7
        const int half = ew >> 2;                                   // which interleaved half of the column blocks
        (void)q;
        const int c_first = g * p.per_tile;
        const int ncap = max(0, min(p.per_tile, p.Bc - c_first));    // valid captions of this tile
        const int nrows = ncap * p.T;                                // valid word rows
        const int T = p.T, Rg = p.Rg, lds = p.lds;
        const bool want_w = BWD ? (p.g_w2r != nullptr) : (p.out_w2r != nullptr);
        const bool want_r = BWD ? (p.g_r2w != nullptr) : (p.out_r2w != nullptr);
        const float scale2 = p.inv_temp * LSM_LOG2E;

        // ---- phase 0: additive masks -----------------------------------------------------------------------
        // (columns >= Rg are padding of the 32-column blocks: a fill far below LSM_FILL keeps them out of an all-masked
        //  row's uniform softmax; the accumulator holds finite values there — block_n is a multiple of 32)
        for (int r = et; r < core.block_n; r += kThreads)
            sm.rbias[r] = (r < Rg) ? ((__ldg(p.reg_mask + (int64_t)i * Rg + r) > 0.f) ? 0.f : LSM_FILL) : LSM_PAD;
        const float mc = (row < nrows) ? __ldg(p.cap_mask + (int64_t)c_first * T + row) : 0.f;
        if (half == 0) sm.cbias[row] = (mc > 0.f) ? 0.f : LSM_FILL;
        named_bar_sync(1, kThreads);

        // ---- phase 1: TMEM -> tile (log2 units) in smem; row softmax statistics, columns split between the halves ----
        const int nblk = (Rg + 31) / 32;
        float *srow = sm.S + (size_t)row * lds;
        const bool row_on = mc > 0.f;
        float mx = -FLT_MAX, best_s = 0.f;
        int best_r = 0;
        float f_t = 0.f, den = 1.f;
        if (!p.hardmax) {
            // fast path (softmax): no per-element predicates.  Rows without a valid word (mc == 0) and padding columns
            // need no special case: the additive region bias masks columns, and a masked row's result is discarded below.
            for (int b = half; b < nblk; b += 2) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)(b * 32), v);
                const float4 *rb4 = reinterpret_cast<const float4 *>(sm.rbias + b * 32);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 rb = rb4[j >> 2];
                    const float s0 = v[j] * scale2, s1 = v[j + 1] * scale2, s2 = v[j + 2] * scale2, s3 = v[j + 3] * scale2;
                    srow[b * 32 + j] = s0; srow[b * 32 + j + 1] = s1; srow[b * 32 + j + 2] = s2; srow[b * 32 + j + 3] = s3;
                    mx = fmaxf(mx, fmaxf(fmaxf(s0 + rb.x, s1 + rb.y), fmaxf(s2 + rb.z, s3 + rb.w)));
                }
            }
            sm.pmax[half * 128 + row] = mx;
            named_bar_sync(1, kThreads);
            mx = fmaxf(sm.pmax[row], sm.pmax[128 + row]);
            if (want_w) {
                // second sweep straight from TMEM: FFMA, FADD, MUFU.EX2, FADD, FFMA per element, four accumulator chains
                float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
                const float nmx = -mx;
                for (int b = half; b < nblk; b += 2) {
                    float v[32];
                    tmem_ld32(taddr + (uint32_t)(b * 32), v);
                    const float4 *rb4 = reinterpret_cast<const float4 *>(sm.rbias + b * 32);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 rb = rb4[j >> 2];
                        const float e0 = ex2_ftz(fmaf(v[j], scale2, nmx) + rb.x);
                        const float e1 = ex2_ftz(fmaf(v[j + 1], scale2, nmx) + rb.y);
                        const float e2 = ex2_ftz(fmaf(v[j + 2], scale2, nmx) + rb.z);
                        const float e3 = ex2_ftz(fmaf(v[j + 3], scale2, nmx) + rb.w);
                        d0 += e0; n0 = fmaf(e0, v[j], n0);
                        d1 += e1; n1 = fmaf(e1, v[j + 1], n1);
                        d2 += e2; n2 = fmaf(e2, v[j + 2], n2);
                        d3 += e3; n3 = fmaf(e3, v[j + 3], n3);
                    }
                }
                sm.pden[half * 128 + row] = (d0 + d1) + (d2 + d3);
                sm.pnum[half * 128 + row] = ((n0 + n1) + (n2 + n3)) * p.inv_temp;      // sum e * s in natural units
            }
            named_bar_sync(1, kThreads);
            if (want_w) {
                den = sm.pden[row] + sm.pden[128 + row];
                f_t = (sm.pnum[row] + sm.pnum[128 + row]) / den;
            }
        } else {
            for (int b = half; b < nblk; b += 2) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)(b * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int r = b * 32 + j;
                    const float s2 = v[j] * scale2;
                    srow[r] = s2;
                    const float sv = (r < Rg) ? (row_on ? s2 + sm.rbias[r] : LSM_FILL) : LSM_PAD;
                    if (sv > mx) { mx = sv; best_s = s2; best_r = r; }     // strict >: first maximum (torch.argmax)
                }
            }
            sm.pmax[half * 128 + row] = mx;
            sm.pden[half * 128 + row] = (float)best_r;
            sm.pnum[half * 128 + row] = best_s;
            named_bar_sync(1, kThreads);
            // first maximum over the whole row: prefer the lower column index on ties (blocks are interleaved)
            const float m0 = sm.pmax[row], m1 = sm.pmax[128 + row];
            const float r0 = sm.pden[row], r1 = sm.pden[128 + row];
            const bool take1 = (m1 > m0) || (m1 == m0 && r1 < r0);
            mx = fmaxf(m0, m1);
            den = take1 ? r1 : r0;                                  // hardmax: `den` carries the argmax index
            f_t = (take1 ? sm.pnum[128 + row] : sm.pnum[row]) * LSM_LN2;
            named_bar_sync(1, kThreads);
        }
        if (half == 0) {
            sm.rowval[row] = row_on ? f_t : 0.f;
            if (BWD) {
                sm.rowmx[row] = mx;
                sm.rowden[row] = den;
                sm.rowf[row] = f_t;
            }
        }

        // ---- phase 2: column softmax over the T words of each caption (tile complete after the barrier above) ----
        if (core.debug_mode == 4) return;
        if (want_r) {
            const int items = ncap * Rg;
            for (int it = et; it < items; it += kThreads) {
                const int cl = it / Rg, r = it - cl * Rg;
                const bool rm_on = sm.rbias[r] == 0.f;
                const float *col = sm.S + (size_t)cl * T * lds + r;
                const float *cb = sm.cbias + cl * T;
                float cmx = -FLT_MAX, h, cden;
                if (!p.hardmax) {
                    // a masked region's column needs no special case (its result is discarded below)
#pragma unroll 4
                    for (int t = 0; t < T; ++t) cmx = fmaxf(cmx, col[(size_t)t * lds] + cb[t]);
                    float d0 = 0.f, d1 = 0.f, n0 = 0.f, n1 = 0.f;
                    const float ncmx = -cmx;
                    int t = 0;
#pragma unroll 2
                    for (; t + 1 < T; t += 2) {
                        const float s0 = col[(size_t)t * lds], s1 = col[(size_t)(t + 1) * lds];
                        const float e0 = ex2_ftz((s0 + ncmx) + cb[t]);
                        const float e1 = ex2_ftz((s1 + ncmx) + cb[t + 1]);
                        d0 += e0; n0 = fmaf(e0, s0, n0);
                        d1 += e1; n1 = fmaf(e1, s1, n1);
                    }
                    if (t < T) {
                        const float s0 = col[(size_t)t * lds];
                        const float e0 = ex2_ftz((s0 + ncmx) + cb[t]);
                        d0 += e0; n0 = fmaf(e0, s0, n0);
                    }
                    cden = d0 + d1;
                    h = (n0 + n1) / cden * LSM_LN2;
                } else {
                    float cbest = 0.f;
                    int best_t = 0;
                    for (int t = 0; t < T; ++t) {
                        const float s2 = col[(size_t)t * lds];
                        const float sv = rm_on ? s2 + cb[t] : LSM_FILL;
                        if (sv > cmx) { cmx = sv; cbest = s2; best_t = t; }
                    }
                    cden = (float)best_t;                              // hardmax: argmax index
                    h = cbest * LSM_LN2;
                }
                sm.colval[it] = rm_on ? h : 0.f;
                if (BWD) {
                    sm.colmx[it] = cmx;
                    sm.colden[it] = cden;
                    sm.colh[it] = h;
                }
            }
        }
        named_bar_sync(1, kThreads);

        // ---- phase 3: per-caption reductions in a fixed order (epilogue warp w handles captions w, w+8, ...) ----------
        if (core.debug_mode == 5) return;
        float nr = 0.f;
        for (int r = lane; r < Rg; r += 32) nr += (sm.rbias[r] == 0.f) ? 1.f : 0.f;
        nr = warp_sum(nr);
        for (int cl = ew; cl < ncap; cl += kEpiWarps) {
            float a = 0.f, nw = 0.f, bsum = 0.f;
            for (int t = lane; t < T; t += 32) {
                a += sm.rowval[cl * T + t];
                nw += (sm.cbias[cl * T + t] == 0.f) ? 1.f : 0.f;
            }
            a = warp_sum(a);
            nw = warp_sum(nw);
            if (want_r) {
                for (int r = lane; r < Rg; r += 32) bsum += sm.colval[cl * Rg + r];
                bsum = warp_sum(bsum);
            }
            if (lane == 0) {
                const int c = c_first + cl;
                if (!BWD) {
                    if (p.out_w2r) p.out_w2r[(int64_t)c * p.ld_out + i] = -a / fmaxf(nw, 1.f);
                    if (p.out_r2w) p.out_r2w[(int64_t)c * p.ld_out + i] = -bsum / fmaxf(nr, 1.f);
                } else {
                    sm.capnw[cl] = nw;
                }
            }
        }
        if (!BWD) return;

        // ---- phase 4 (BWD): dS in place (row-owning threads, columns split between the halves) --------------------
        named_bar_sync(1, kThreads);
        if (row < nrows) {
            const int cl = row / T, t = row - cl * T;
            const int c = c_first + cl;
            // d(d_w2r)/ds = -(m_c/nw) P (1 + valid (s - f_t));  d(d_r2w)/ds = -(m_r/nr) Q (1 + valid (s - h_r));
            // raw similarity = s * temperature  =>  one more factor 1/temperature
            const float gw = want_w ? -__ldg(p.g_w2r + (int64_t)c * p.ld_g + i) * mc / fmaxf(sm.capnw[cl], 1.f) * p.inv_temp : 0.f;
            const float gr = want_r ? -__ldg(p.g_r2w + (int64_t)c * p.ld_g + i) / fmaxf(nr, 1.f) * p.inv_temp : 0.f;
            const float rmx = sm.rowmx[row], rinv = 1.f / sm.rowden[row], rf = sm.rowf[row];
            for (int b = half; b < nblk; b += 2) {
                const int r_end = min(Rg, b * 32 + 32);
#pragma unroll 4
                for (int r = b * 32; r < r_end; ++r) {
                    const float s2 = srow[r];
                    const float s = s2 * LSM_LN2;
                    const bool rm_on = sm.rbias[r] == 0.f;
                    const bool valid = row_on && rm_on;
                    const float sv = valid ? s2 : LSM_FILL;
                    float d = 0.f;
                    if (want_w) {
                        if (!p.hardmax) d += gw * ex2_ftz(sv - rmx) * rinv * (1.f + (valid ? s - rf : 0.f));
                        else d += (r == (int)sm.rowden[row]) ? gw : 0.f;
                    }
                    if (want_r) {
                        const int it = cl * Rg + r;
                        if (!p.hardmax) d += rm_on ? gr * ex2_ftz(sv - sm.colmx[it]) / sm.colden[it] * (1.f + (valid ? s - sm.colh[it] : 0.f)) : 0.f;
                        else d += (rm_on && t == (int)sm.colden[it]) ? gr : 0.f;
                    }
                    srow[r] = d;
                }
            }
        }
        named_bar_sync(1, kThreads);

        // ---- phase 5 (BWD): coalesced write-out ------------------------------------------------------------------
        const int64_t col0 = (int64_t)c_first * T;                  // first caption-word column of this tile in dS^T
        for (int r = ew; r < Rg; r += kEpiWarps) {                  // warp per region row of dS^T
            uint16_t *dh = p.dst_hi + ((int64_t)i * Rg + r) * p.ld_dst + col0;
            uint16_t *dl = p.dst_lo ? p.dst_lo + ((int64_t)i * Rg + r) * p.ld_dst + col0 : nullptr;
            for (int w = lane; w < nrows; w += 32) {
                uint16_t h, l;
                split_bf16(sm.S[(size_t)w * lds + r], h, l);
                dh[w] = h;
                if (dl) dl[w] = l;
            }
        }
        if (p.ds_hi != nullptr) {
            for (int w = ew; w < nrows; w += kEpiWarps) {           // warp per caption-word row of dS
                uint16_t *dh = p.ds_hi + (col0 + w) * p.ld_ds + (int64_t)i * Rg;
                uint16_t *dl = p.ds_lo ? p.ds_lo + (col0 + w) * p.ld_ds + (int64_t)i * Rg : nullptr;
                for (int r = lane; r < Rg; r += 32) {
                    uint16_t h, l;
                    split_bf16(sm.S[(size_t)w * lds + r], h, l);
                    dh[r] = h;
                    if (dl) dl[r] = l;
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// LSM pair FORWARD epilogue, second generation (softmax alignment): persistent CTA pairs.
//
// A CTA pair (tcgen05 cta_group::2, M = 256) owns the similarity tile of 2 x `per_tile` captions (CTA rm holds the word rows of
// caption group 2*gp + rm) x `ipt` consecutive images (image j of the tile = accumulator columns [j*Rg, j*Rg + Rg)): the region
// operand of the images is staged ONCE for both caption groups (each CTA loads half of it), so the bytes an SM pulls from L2 per
// MAC drop to 0.6x of the single-CTA 128 x 128 tile the first generation used — that kernel was bound by TMA ingest (30 KB per
// 224-cycle k block), not by the tensor pipe.  Pairs are persistent (pair p takes tiles p, p + P, ...) with two accumulator stages
// in tensor memory, so the epilogue of tile k runs under the MMAs of tile k + 1, and it frees its accumulator right after its last
// tcgen05.ld so that tile k + 2 can start under the rest of it.
//
// Epilogue = two independent halves of four warps (TMEM lane quarters 0-3 each); half h takes the images j = h, h + 2, ... of
// the tile.  Per image:
//   row pass    (thread = word row)  : ONE sweep over the image's Rg accumulator columns with an online softmax (running max,
//                                      rescaled sums) -> attention-pooled f_t (w2r); the scaled similarities (log2 units) are
//                                      parked in the half's private shared-memory sub-tile with conflict-free 128-bit stores
//   column pass (thread = region)    : per caption, softmax over its T words down the parked column -> h_r (r2w), reduced over the
//                                      regions with a fixed-order shuffle tree per warp
//   finish      (thread = caption)   : fixed-order sums of the T row values / the four warp partials -> the two scalars of the pair
// Masks of the NEXT tile are fetched by prefetch() while its MMAs still run.  All reductions have a fixed order that depends
// only on (caption, image): any block of the pair matrix is bit-identical to the same block computed in another tiling / shard.
// ------------------------------------------------------------------------------------------------
struct LsmFwdParams {
    const float *cap_mask;   // [Bc, T]
    const float *reg_mask;   // [Bi, Rg]
    float *out_w2r, *out_r2w;   // [Bc, Bi] (ld_out), each may be null
    int64_t ld_out;
    int Bc, T, Bi, Rg;
    int per_tile;            // captions per CTA (one 128-row half of the pair tile)
    int ipt;                 // images per tile
    int slots;               // images per epilogue half = ceil(ipt / 2)
    int halves;              // epilogue halves with a shared-memory allocation (1 when ipt == 1)
    int Tp;                  // words per caption in the parked layout = round_up(T, 4): a caption's column segment is whole float4s
    int ldt;                 // parked sub-tile row stride (floats): >= per_tile * Tp, multiple of 4, ldt / 4 odd (conflict-free 128-bit loads)
    int rb;                  // region-bias stride per image slot = round_up(Rg, 32)
    int npairs;              // CTA pairs launched
    int tiles_i;             // image tiles = ceil(Bi / ipt)
    float inv_temp;
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float exp_col(float x) { return ex2_ftz(x); }
__device__ __forceinline__ float exp_row(float x) { return ex2_ftz(x); }

// LDT > 0: the parked sub-tile's row stride is this compile-time constant (every transposed store is one STS with an immediate
// offset); LDT == 0: run-time stride p.ldt (caption groups whose padded word count exceeds the constant).
//
// SIXTEEN epilogue warps = four warpgroups of 128 threads, each covering the 128 accumulator rows (TMEM lane quarter = warp & 3).
// Warpgroup g works on image (g & 1) of the current pair of images ("slot") and, in the row pass, on column half (g >> 1) of that
// image's accumulator columns; the two warpgroups of an image ("image group", 256 threads, one named barrier) merge their partial
// row statistics through shared memory, then split the CAPTIONS of the column pass between them.  Four warps per scheduler instead
// of two: the epilogue is a chain of dependent packed-fp32 / MUFU / shared-memory operations, and with two warps per scheduler every
// one of those latencies was exposed (issue slots 14 % busy in the r2 profile).
template <int LDT>
struct EpiLsmFwd {
    static constexpr int kEpiWarps = 16;
    static constexpr int kMinBlocks = 1;
    static constexpr bool kHasPrefetch = true, kSelfRelease = true, kWantsMaps = false;
    static constexpr int kCbt = 192;       // >= per_tile * Tp (per_tile <= 16, per_tile * T <= 128, Tp <= T + 3)
    typedef LsmFwdParams Params;
    uint32_t release_bar;        // set by the core: shared::cluster address of the leader's accumulator-empty barrier
    uint64_t *release_local;
    // per-tile state computed by prefetch() (before the accumulator is awaited) and used by chunk(): no integer divisions,
    // mask reads or tile decoding between the arrival of the accumulator and the first tcgen05.ld
    struct Tile { int rm, gp, it, c_first, ncap, nrows; };
    Tile tile_;
    int pcol_, nslots_;
    bool row_in_, row_on_;

    static __device__ __forceinline__ Tile decode(const Params &p, int cta, int ch) {
        Tile t;
        t.rm = cta / p.npairs;
        const int tile = (cta - t.rm * p.npairs) + ch * p.npairs;
        t.gp = tile / p.tiles_i;
        t.it = tile - t.gp * p.tiles_i;
        t.c_first = (2 * t.gp + t.rm) * p.per_tile;
        t.ncap = max(0, min(p.per_tile, p.Bc - t.c_first));
        t.nrows = t.ncap * p.T;
        return t;
    }
    static __device__ __forceinline__ void coords(const Params &p, const TcCore &, int cta, int ch, int &row_a, int &row_b) {
        const Tile t = decode(p, cta, ch);
        row_a = t.c_first * p.T;                 // word rows of this CTA's caption group (A operand; rows past Bc*T are zero-filled)
        row_b = t.it * p.ipt * p.Rg;             // region rows of the tile's images (B operand; the core adds this CTA's half)
    }
    // shared memory of one image group.  All additive masks are in LOG2 units (pre-multiplied by c = inv_temp * log2 e), so one packed
    // FFMA turns two raw accumulator values into two masked softmax arguments.
    struct Half {
        float *park;        // [Rg][ldt]   RAW similarities of the image being processed, TRANSPOSED: region-major, a caption's words
                            //             contiguous at [cl * Tp, cl * Tp + T)
        float *rbias;       // [slots][rb] additive region mask: 0 valid, LSM_FILL * c masked, LSM_PAD * c past the image
        float *cbt;         // [kCbt]      additive word mask in the parked layout: 0 valid, LSM_FILL * c masked, LSM_PAD * c for the pad words
        float *rowval;      // [128]       f_t of the accumulator row (0 for a masked word)
        float *part;        // [2][3][128] partial row statistics (max, sum, weighted sum) of the two column halves
        float *capnw;       // [16]        valid words per caption
        float *nreg;        // [16]        valid regions per image slot
        int *ncols;         // [16]        accumulator columns of the image that can matter (last valid region + 1; Rg when none is valid)
        int *capq;          // [16]        float4 quads of a caption's parked column segment that can matter (up to its last valid word)
        float *hbuf;        // [16][128]   h_r of (caption, region thread), summed per caption in the finish phase
    };
    // "big" part of an image group's scratch (parked sub-tile + per-region results): may overlay the operand ring when every pair
    // has a single tile; "small" part (masks, counts, row statistics): filled by prefetch() while operands are in flight, never overlaid
    static __host__ __device__ __forceinline__ size_t big_floats(int Rg, int ldt, int per_tile) {
        return ((size_t)Rg * ldt + (size_t)per_tile * 128 + 3) & ~(size_t)3;
    }
    static __host__ __device__ __forceinline__ size_t small_floats(int slots, int rb) {
        return ((size_t)slots * rb + kCbt + 128 + 6 * 128 + LSM_MAX_PER_TILE + 16 + 16 + 16 + 3) & ~(size_t)3;
    }
    static __device__ __forceinline__ int stride(const Params &p) { return LDT > 0 ? LDT : p.ldt; }
    static __device__ __forceinline__ Half carve(const Params &p, const TcCore &core, unsigned char *smem, int img) {
        // core hands over the ring base when the scratch overlays it, else the first byte after the barriers
        float *small = reinterpret_cast<float *>(core.epi_overlay ? smem + core.ring_bytes + 512 : smem);
        float *big = core.epi_overlay ? reinterpret_cast<float *>(smem) : small + 2 * small_floats(p.slots, p.rb);
        float *b = small + (size_t)img * small_floats(p.slots, p.rb);
        Half h;
        h.rbias = b; b += (size_t)p.slots * p.rb;
        h.cbt = b; b += kCbt;
        h.rowval = b; b += 128;
        h.part = b; b += 6 * 128;
        h.capnw = b; b += LSM_MAX_PER_TILE;
        h.nreg = b; b += 16;
        h.ncols = reinterpret_cast<int *>(b); b += 16;
        h.capq = reinterpret_cast<int *>(b);
        h.park = big + (size_t)img * big_floats(p.Rg, p.ldt, p.per_tile);
        h.hbuf = h.park + (size_t)p.Rg * p.ldt;
        return h;
    }
    __device__ __forceinline__ void begin(const Params &, const TcCore &, int, int, int, int, unsigned char *) {}
    __device__ __forceinline__ void finish(const Params &, const TcCore &, int, int, int, int, unsigned char *) {}

    // masks of tile `ch` -> additive biases and counts in the image group's shared memory (all before the accumulator is awaited)
    __device__ __forceinline__ void prefetch(const Params &p, const TcCore &core, int cta, int ch, int row, int lane, int, unsigned char *smem) {
        const int e = (threadIdx.x >> 5) - 2, g = e >> 2, img = g & 1, hcol = g >> 1;
        const int gt = hcol * 128 + (e & 3) * 32 + lane, gw = hcol * 4 + (e & 3);      // thread / warp index inside the image group
        const Tile t = decode(p, cta, ch);
        tile_ = t;
        nslots_ = 0;                                            // images this image group really has in this tile
        if (img < p.halves)
            for (int s = 0; s < p.slots; ++s)
                if (img + 2 * s < p.ipt && t.it * p.ipt + img + 2 * s < p.Bi) nslots_ = s + 1;
        row_in_ = row < t.nrows;
        row_on_ = false;
        const int rcl = row / p.T;
        pcol_ = rcl * p.Tp + (row - rcl * p.T);                 // this row's position in the parked (padded) word layout
        if (img >= p.halves) return;
        const Half h = carve(p, core, smem, img);
        const float c2 = p.inv_temp * LSM_LOG2E;
        const float fill2 = LSM_FILL * c2, pad2 = LSM_PAD * c2;
        named_bar_sync(2 + img, 256);                          // the image group has finished reading the previous tile's buffers
        if (row_in_) {
            row_on_ = __ldg(p.cap_mask + (int64_t)t.c_first * p.T + row) > 0.f;
            if (hcol == 0) h.cbt[pcol_] = row_on_ ? 0.f : fill2;
        }
        if (gt < t.ncap)
            for (int w = p.T; w < p.Tp; ++w) h.cbt[gt * p.Tp + w] = pad2;
        for (int s = 0; s < p.slots; ++s) {
            const int i = t.it * p.ipt + img + 2 * s;
            const bool img_on = (img + 2 * s < p.ipt) && i < p.Bi;
            for (int r = gt; r < p.rb; r += 256)
                h.rbias[s * p.rb + r] = (r < p.Rg && img_on) ? ((__ldg(p.reg_mask + (int64_t)i * p.Rg + r) > 0.f) ? 0.f : fill2) : pad2;
        }
        named_bar_sync(2 + img, 256);
        if (gt < t.ncap) {                                     // valid words of caption gt
            float nw = 0.f;
            int last = 0;
            for (int w = 0; w < p.T; ++w)
                if (h.cbt[gt * p.Tp + w] == 0.f) { nw += 1.f; last = w + 1; }
            h.capnw[gt] = nw;
            h.capq[gt] = last > 0 ? (last + 3) >> 2 : (p.Tp >> 2);       // no valid word: uniform over all T words
        }
        for (int s = gw; s < p.slots; s += 8) {                // the group's images: one warp each
            float nr = 0.f;
            int last = 0;
            for (int r = lane; r < p.Rg; r += 32)
                if (h.rbias[s * p.rb + r] == 0.f) { nr += 1.f; last = r + 1; }
            nr = warp_sum(nr);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
            if (lane == 0) {
                h.nreg[s] = nr;
                h.ncols[s] = last > 0 ? last : p.Rg;           // none valid: uniform over all Rg regions, every column matters
            }
        }
        named_bar_sync(2 + img, 256);                          // counts are read by every thread of the group in chunk()
    }

    // W accumulator columns of this thread's row (RAW units): park them transposed and fold them into the running softmax statistics
    // (m: running max of the masked log2-domain values, d: sum of 2^(x - m), n: sum of those weights times the RAW value), two
    // columns per instruction with Blackwell's packed fp32x2 arithmetic.
    // TAIL: only the first `valid` columns belong to the image (the others may hold anything, NaN patterns included: select, never
    // multiply).  MASKED: add the region bias (an image whose regions are all valid needs none in its full blocks).
    template <int W, bool TAIL, bool MASKED>
    __device__ __forceinline__ void fold(uint32_t (&v)[32], int valid, float c, const float *rb, float *pp, int ldt, bool store, float &m,
                                         float2 &d, float2 &n) {
        float2 x[W / 2];
        const float2 cc = f2(c, c);
        float bm = -FLT_MAX;
#pragma unroll
        for (int j = 0; j < W; j += 4) {
            float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MASKED || TAIL) r4 = *reinterpret_cast<const float4 *>(rb + j);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int k = j + 2 * u;
                if (TAIL) {
                    if (k >= valid) v[k] = 0u;
                    if (k + 1 >= valid) v[k + 1] = 0u;
                }
                const float2 raw = f2(__uint_as_float(v[k]), __uint_as_float(v[k + 1]));
                if (store) {
                    if (!TAIL || k < valid) pp[(size_t)k * ldt] = raw.x;
                    if (!TAIL || k + 1 < valid) pp[(size_t)(k + 1) * ldt] = raw.y;
                }
                x[k >> 1] = (MASKED || TAIL) ? __ffma2_rn(raw, cc, u == 0 ? f2(r4.x, r4.y) : f2(r4.z, r4.w)) : __fmul2_rn(raw, cc);
                bm = fmaxf(bm, fmaxf(x[k >> 1].x, x[k >> 1].y));
            }
        }
        const float m_new = fmaxf(m, bm);
        const float corr = ex2_ftz(m - m_new);
        const float2 cr = f2(corr, corr), nm = f2(-m_new, -m_new);
        float2 d0 = __fmul2_rn(d, cr), n0 = __fmul2_rn(n, cr), d1 = f2(0.f, 0.f), n1 = f2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < W / 2; k += 2) {
            // x - m: exactly 0 where every region is masked (x == m == the fill)
            const float2 t0 = __fadd2_rn(x[k], nm), t1 = __fadd2_rn(x[k + 1], nm);
            const float2 e0 = f2(exp_row(t0.x), exp_row(t0.y)), e1 = f2(exp_row(t1.x), exp_row(t1.y));
            d0 = __fadd2_rn(d0, e0); n0 = __ffma2_rn(e0, f2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1])), n0);
            d1 = __fadd2_rn(d1, e1); n1 = __ffma2_rn(e1, f2(__uint_as_float(v[2 * k + 2]), __uint_as_float(v[2 * k + 3])), n1);
        }
        m = m_new;
        d = __fadd2_rn(d0, d1);
        n = __fadd2_rn(n0, n1);
    }

    // softmax over one caption's words down a parked column (NQ float4 quads; `tlast` real words in the last one):
    // returns sum_t softmax_t * RAW similarity
    template <int NQ>
    static __device__ __forceinline__ float col_caption(const float *base, const float *cb, int tlast, float c) {
        float2 a[2 * NQ], x[2 * NQ];
        const float2 cc = f2(c, c);
        float mx = -FLT_MAX;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float4 v = *reinterpret_cast<const float4 *>(base + 4 * q);
            const float4 b = *reinterpret_cast<const float4 *>(cb + 4 * q);
            if (q == NQ - 1) {                                  // pad words were never parked: select them away
                if (tlast < 2) v.y = 0.f;
                if (tlast < 3) v.z = 0.f;
                if (tlast < 4) v.w = 0.f;
            }
            a[2 * q] = f2(v.x, v.y);
            a[2 * q + 1] = f2(v.z, v.w);
            x[2 * q] = __ffma2_rn(a[2 * q], cc, f2(b.x, b.y));
            x[2 * q + 1] = __ffma2_rn(a[2 * q + 1], cc, f2(b.z, b.w));
            mx = fmaxf(mx, fmaxf(fmaxf(x[2 * q].x, x[2 * q].y), fmaxf(x[2 * q + 1].x, x[2 * q + 1].y)));
        }
        const float2 nm = f2(-mx, -mx);
        float2 d0 = f2(0.f, 0.f), d1 = f2(0.f, 0.f), n0 = f2(0.f, 0.f), n1 = f2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 2 * NQ; k += 2) {
            const float2 t0 = __fadd2_rn(x[k], nm), t1 = __fadd2_rn(x[k + 1], nm);       // (exactly 0 for an all-masked caption)
            const float2 e0 = f2(exp_col(t0.x), exp_col(t0.y)), e1 = f2(exp_col(t1.x), exp_col(t1.y));
            d0 = __fadd2_rn(d0, e0); n0 = __ffma2_rn(e0, a[k], n0);
            d1 = __fadd2_rn(d1, e1); n1 = __ffma2_rn(e1, a[k + 1], n1);
        }
        const float2 dd = __fadd2_rn(d0, d1), nn = __fadd2_rn(n0, n1);
        return __fdividef(nn.x + nn.y, dd.x + dd.y);            // 2 ulp: far inside the 1e-4 bar
    }
    // any number of quads (more than 8: T > 32): two sweeps over the column segment
    static __device__ __forceinline__ float col_caption_long(const float *base, const float *cb, int nq, int tlast, float c) {
        float mx = -FLT_MAX;
        for (int q = 0; q < nq; ++q) {
            float4 v = *reinterpret_cast<const float4 *>(base + 4 * q);
            const float4 b = *reinterpret_cast<const float4 *>(cb + 4 * q);
            if (q == nq - 1) { if (tlast < 2) v.y = 0.f; if (tlast < 3) v.z = 0.f; if (tlast < 4) v.w = 0.f; }
            mx = fmaxf(mx, fmaxf(fmaxf(fmaf(v.x, c, b.x), fmaf(v.y, c, b.y)), fmaxf(fmaf(v.z, c, b.z), fmaf(v.w, c, b.w))));
        }
        float d0 = 0.f, d1 = 0.f, n0 = 0.f, n1 = 0.f;
        for (int q = 0; q < nq; ++q) {
            float4 v = *reinterpret_cast<const float4 *>(base + 4 * q);
            const float4 b = *reinterpret_cast<const float4 *>(cb + 4 * q);
            if (q == nq - 1) { if (tlast < 2) v.y = 0.f; if (tlast < 3) v.z = 0.f; if (tlast < 4) v.w = 0.f; }
            const float e0 = ex2_ftz(fmaf(v.x, c, b.x) - mx), e1 = ex2_ftz(fmaf(v.y, c, b.y) - mx);
            const float e2 = ex2_ftz(fmaf(v.z, c, b.z) - mx), e3 = ex2_ftz(fmaf(v.w, c, b.w) - mx);
            d0 += e0; n0 = fmaf(e0, v.x, n0);
            d1 += e1; n1 = fmaf(e1, v.y, n1);
            d0 += e2; n0 = fmaf(e2, v.z, n0);
            d1 += e3; n1 = fmaf(e3, v.w, n1);
        }
        return __fdividef(n0 + n1, d0 + d1);
    }
    // one caption's column segment, `q` quads of it (up to the caption's last valid word; the last quad of a whole segment holds `tlast`
    // real words)
    static __device__ __forceinline__ float col_any(const float *base, const float *cb, int q, int nq, int tlast, float c2) {
        const int tl = q == nq ? tlast : 4;
        switch (q) {
            case 1: return col_caption<1>(base, cb, tl, c2);
            case 2: return col_caption<2>(base, cb, tl, c2);
            case 3: return col_caption<3>(base, cb, tl, c2);
            case 4: return col_caption<4>(base, cb, tl, c2);
            case 5: return col_caption<5>(base, cb, tl, c2);
            case 6: return col_caption<6>(base, cb, tl, c2);
            case 7: return col_caption<7>(base, cb, tl, c2);
            case 8: return col_caption<8>(base, cb, tl, c2);
            default: return col_caption_long(base, cb, q, tl, c2);
        }
    }

    __device__ __forceinline__ void release_acc() {
        tc_fence_before();
        mbar_arrive_cluster_cta(release_bar);
    }

    __device__ __forceinline__ void chunk(const Params &p, const TcCore &core, int, int, uint32_t taddr, int row, int lane, int,
                                          unsigned char *smem) {
        const int e = (threadIdx.x >> 5) - 2, g = e >> 2, img = g & 1, hcol = g >> 1;
        const int ht = (e & 3) * 32 + lane, gw = hcol * 4 + (e & 3);
        const Tile t = tile_;
        const int T = p.T, Rg = p.Rg, ldt = stride(p), Tp = p.Tp;
        const int nslots = nslots_;
        if (t.ncap == 0 || nslots == 0 || core.debug_mode == 3) {   // padding half of the last pair / nothing for this image group
            release_acc();
            return;
        }
        const Half h = carve(p, core, smem, img);
        unsigned long long *tl = (core.timeline != nullptr && threadIdx.x == 64) ? core.timeline + (size_t)blockIdx.x * 16 : nullptr;
        const float c2 = p.inv_temp * LSM_LOG2E;
        const bool row_in = row_in_, row_on = row_on_;
        const int pcol = pcol_;
        const int nq = Tp >> 2, tlast = T - 4 * (nq - 1);
        for (int s = 0; s < nslots; ++s) {
            const int j = img + 2 * s, i = t.it * p.ipt + j;
            const float *rb = h.rbias + s * p.rb;
            const int ncols = h.ncols[s];
            // ---- row pass: this warpgroup's column half of the image's accumulator columns ----------------------------------------
            {
                const uint32_t tcol = taddr + (uint32_t)(j * Rg);
                const int nb = ncols >> 5, rem = ncols & 31;
                const int tw = rem == 0 ? 0 : (rem <= 16 ? 16 : 32);
                const int nblk = nb + (tw > 0 ? 1 : 0), nsplit = (nblk + 1) >> 1;
                const int k0 = hcol == 0 ? 0 : nsplit, k1 = hcol == 0 ? nsplit : nblk;
                const bool masked = h.nreg[s] != (float)Rg;
                float m = -FLT_MAX;
                float2 d = f2(0.f, 0.f), n = f2(0.f, 0.f);
                if (tl) tl[8] = global_timer_ns();
                for (int k = k0; k < k1; ++k) {
                    uint32_t cur[32];
                    const uint32_t adr = tcol + (uint32_t)(k * 32);
                    float *pp = h.park + (size_t)(k * 32) * ldt + pcol;
                    const float *rbk = rb + k * 32;
                    if (k < nb || tw == 32) {
                        tmem_ld_async<32>(adr, cur);
                        tmem_ld_fence<32>(cur);
                        if (k < nb) {
                            if (masked) fold<32, false, true>(cur, 32, c2, rbk, pp, ldt, row_in, m, d, n);
                            else fold<32, false, false>(cur, 32, c2, rbk, pp, ldt, row_in, m, d, n);
                        } else fold<32, true, true>(cur, rem, c2, rbk, pp, ldt, row_in, m, d, n);
                    } else {
                        tmem_ld_async<16>(adr, cur);
                        tmem_ld_fence<16>(cur);
                        fold<16, true, true>(cur, rem, c2, rbk, pp, ldt, row_in, m, d, n);
                    }
                }
                if (tl) tl[10] = global_timer_ns();
                if (s == nslots - 1) release_acc();                                   // last tcgen05.ld of this thread for this tile
                float *pt = h.part + hcol * 384 + row;
                pt[0] = m; pt[128] = d.x + d.y; pt[256] = n.x + n.y;
            }
            named_bar_sync(2 + img, 256);
            if (tl) tl[11] = global_timer_ns();
            if (hcol == 0) {                                                          // merge the two column halves of the row
                const float m0 = h.part[row], m1 = h.part[384 + row];
                const float mm = fmaxf(m0, m1), w0 = ex2_ftz(m0 - mm), w1 = ex2_ftz(m1 - mm);
                const float dd = h.part[128 + row] * w0 + h.part[512 + row] * w1, nn = h.part[256 + row] * w0 + h.part[640 + row] * w1;
                h.rowval[row] = row_on ? __fdividef(nn, dd) * p.inv_temp : 0.f;
            }
            // ---- column pass: thread = region, softmax over the T words of each caption down the parked column; the captions are split
            //      between the two warpgroups; per-(caption, region) results go to shared memory and are summed in the finish phase ----
            const bool do_col = core.debug_mode != 4 && p.out_r2w != nullptr;
            if (do_col) {
                const int csplit = (t.ncap + 1) >> 1;
                const int c_lo = hcol == 0 ? 0 : csplit, c_hi = hcol == 0 ? csplit : t.ncap;
                for (int r0 = 0; r0 < ncols; r0 += 128) {
                    const int r = r0 + ht;
                    const bool r_on = r < ncols && rb[min(r, ncols - 1)] == 0.f;
                    const float *colbase = h.park + (size_t)min(r, ncols - 1) * ldt;
                    const bool warp_on = __any_sync(0xffffffffu, r_on);               // a warp whose regions are all masked has nothing to do
                    for (int cl = c_lo; cl < c_hi; ++cl) {
                        float v = 0.f;
                        if (warp_on) v = col_any(colbase + cl * Tp, h.cbt + cl * Tp, h.capq[cl], nq, tlast, c2);
                        float *hb = h.hbuf + cl * 128 + ht;
                        hb[0] = (r0 == 0 ? 0.f : hb[0]) + (r_on ? v * p.inv_temp : 0.f);
                    }
                }
            }
            if (tl) tl[12] = global_timer_ns();
            named_bar_sync(2 + img, 256);
            if (tl) tl[13] = global_timer_ns();
            // ---- finish: warp = caption; both sums in a fixed order (lane-strided partials, then the shuffle tree) ------------------------
            for (int cl = gw; cl < t.ncap; cl += 8) {
                float a = 0.f, b = 0.f;
                for (int w = lane; w < T; w += 32) a += h.rowval[cl * T + w];
                if (do_col) b = (h.hbuf[cl * 128 + lane] + h.hbuf[cl * 128 + 32 + lane]) + (h.hbuf[cl * 128 + 64 + lane] + h.hbuf[cl * 128 + 96 + lane]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                }
                if (lane == 0) {
                    const int c = t.c_first + cl;
                    if (p.out_w2r != nullptr) p.out_w2r[(int64_t)c * p.ld_out + i] = -a / fmaxf(h.capnw[cl], 1.f);
                    if (p.out_r2w != nullptr) p.out_r2w[(int64_t)c * p.ld_out + i] = -b / fmaxf(h.nreg[s], 1.f);
                }
            }
            if (s + 1 < nslots) named_bar_sync(2 + img, 256);       // park / rowval / hbuf are rewritten by the next image
        }
    }
};
constexpr int LSM_LDT = 132;      // compile-time parked stride: covers per_tile * round_up(T, 4) <= 132 (T = 20: 120; T = 70: 72; T = 7, 8, 16: 128)

// ------------------------------------------------------------------------------------------------
// host-side helpers
// ------------------------------------------------------------------------------------------------
static int fill_maps(TcMaps &maps, const uint16_t *a_hi, const uint16_t *a_lo, uint64_t a_rows, uint64_t a_ld,
                     const uint16_t *b_hi, const uint16_t *b_lo, uint64_t b_rows, uint64_t b_ld, uint64_t K, const TcCore &core) {
    // TMA boxes are the per-CTA SLICES of the tiles (the whole tile without a cluster)
    const uint32_t a_box = (uint32_t)(TC_BLOCK_M / core.cn), b_box = (uint32_t)(core.block_n / core.cm);
    int rc;
    if (core.tf32) {       // operands are fp32 matrices read in place (no hi/lo split)
        if ((rc = make_tmap_f32_2d(&maps.a_hi, a_hi, a_rows, K, a_ld, a_box)) != LOCO_OK) return rc;
        if ((rc = make_tmap_f32_2d(&maps.b_hi, b_hi, b_rows, K, b_ld, b_box)) != LOCO_OK) return rc;
        maps.a_lo = maps.a_hi;
        maps.b_lo = maps.b_hi;
        return LOCO_OK;
    }
    if ((rc = make_tmap_bf16_2d(&maps.a_hi, a_hi, a_rows, K, a_ld, a_box)) != LOCO_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&maps.a_lo, a_lo ? a_lo : a_hi, a_rows, K, a_ld, a_box)) != LOCO_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&maps.b_hi, b_hi, b_rows, K, b_ld, b_box)) != LOCO_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&maps.b_lo, b_lo ? b_lo : b_hi, b_rows, K, b_ld, b_box)) != LOCO_OK) return rc;
    return LOCO_OK;
}

// Thread-block clusters with TMA multicast are implemented (tc_gemm.cuh) but OFF by default: re-swept after every change of
// the core (profiles/README.md §1), they have not beaten CTA pairs / single-CTA tiles at any shape of the path.
// LOCOV_B200_CLUSTER=1 enables them for A/B measurements.
static bool clusters_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("LOCOV_B200_CLUSTER");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v != 0;
}

// Tile / cluster shape for a plain GEMM [M,K]x[N,K]^T: minimise  waves * k_blocks * max(MMA, L2)  per tile with
//   MMA  = 4 * BLOCK_N/2 cycles per 64-wide k block (128 x BLOCK_N x 16 MACs at 4096 MAC/cycle/SM),
//   L2   = operand bytes this CTA pulls per k block * concurrently running CTAs / ~3400 B/cycle  — the L2 -> SM
//          fabric of a B200 moves about as much as HBM (profiles/README.md), so operand re-reads, not the tensor
//          pipe, bound these GEMMs unless tiles are shared through TMA multicast.
static void pick_gemm_shape(int M, int N, int sms, TcCore &core) {
    const int mt = (M + TC_BLOCK_M - 1) / TC_BLOCK_M;
    double best_cost = 1e300;
    const int cms[2] = {1, 2}, cns[3] = {1, 2, 4};
    core.block_n = 128; core.cm = 1; core.cn = 1;
    for (int ci = 0; ci < 2; ++ci)
        for (int cj = 0; cj < 3; ++cj) {
            const int cm = cms[ci], cn = cns[cj];
            if (cm * cn > 1 && !clusters_enabled()) continue;
            for (int bn = 256; bn >= 32; bn -= 16) {
                if ((bn / cm) % 8 != 0 || bn % cm != 0) continue;
                const int tm = tc_round_up(mt, cm), tn = tc_round_up((N + bn - 1) / bn, cn);
                if (cm > 1 && mt < cm) continue;
                if (cn > 1 && (N + bn - 1) / bn < cn) continue;
                const int ctas = tm * tn, csize = cm * cn;
                const int slots = (sms / csize) * csize;
                const int waves = (ctas + slots - 1) / slots;
                // per 64-wide k block: tensor pipe 4 * bn/2 cycles vs. shared-memory port (TMA fill + SS operand reads)
                const double mma = 4.0 * bn / 2.0;
                const double port = ((128.0 + bn) * 128.0 * 2.0) / 128.0;
                const double per_k = (mma > port ? mma : port);
                const double cost = waves * (per_k + 60.0 + bn * 0.5) + 1e-3 * csize;   // + amortised per-tile fixed cost / epilogue
                if (cost < best_cost) { best_cost = cost; core.block_n = bn; core.cm = cm; core.cn = cn; }
            }
        }
    // developer overrides for shape sweeps (scripts/gemm_sweep.py): LOCOV_B200_BN / _CM / _CN
    if (const char *e = getenv("LOCOV_B200_BN")) { const int v = atoi(e); if (v >= 16 && v <= 256 && v % 16 == 0) core.block_n = v; }
    if (const char *e = getenv("LOCOV_B200_CM")) { const int v = atoi(e); if (v == 1 || v == 2) core.cm = v; }
    if (const char *e = getenv("LOCOV_B200_CN")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) core.cn = v; }
    core.clusters_n = tc_round_up((N + core.block_n - 1) / core.block_n, core.cn) / core.cn;
}

// CTA pairs (tcgen05 cta_group::2, tc_gemm.cuh) for the plain GEMMs: on by default, LOCOV_B200_2CTA=0 falls back to
// single-CTA tiles (A/B measurements).
static bool pair_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("LOCOV_B200_2CTA");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

// Tile width for a CTA-pair GEMM: a pair owns a 256 x bn tile, per CTA and 128-byte k block the tensor pipe needs 2 * bn
// cycles and the shared-memory fill (128 rows of A + bn/2 rows of B at ~52 B/cycle/SM, measured) (16384 + 64 * bn) / 52;
// minimise waves * max(pipe, fill) + per-tile fixed cost.  Returns false when a single row tile makes pairs pointless.
static bool pick_pair_shape(int M, int N, int sms, TcCore &core) {
    const int mt = (M + TC_BLOCK_M - 1) / TC_BLOCK_M;
    if (!pair_enabled() || mt < 2 || sms < 2) return false;
    double best = 1e300;
    int best_bn = 0;
    for (int bn = 256; bn >= 32; bn -= 16) {
        const int pairs = ((mt + 1) / 2) * ((N + bn - 1) / bn);
        const int waves = (pairs + sms / 2 - 1) / (sms / 2);
        const double pipe = 2.0 * bn, fill = (16384.0 + 64.0 * bn) / 52.0;
        const double cost = waves * ((pipe > fill ? pipe : fill) + 60.0 + bn * 0.5);
        if (cost < best) { best = cost; best_bn = bn; }
    }
    if (const char *e = getenv("LOCOV_B200_BN")) { const int v = atoi(e); if (v >= 32 && v <= 256 && v % 16 == 0) best_bn = v; }
    core.block_n = best_bn;
    core.cm = 2; core.cn = 1; core.two_cta = 1;
    core.clusters_n = (N + best_bn - 1) / best_bn;
    return true;
}

}  // namespace loco

using namespace loco;

extern "C" {

static int linear_launch(const void *A_hi, const void *A_lo, int64_t lda, const void *W_hi, const void *W_lo, int64_t ldw, bool tf32,
                         const float *bias, int M, int N, int K, float *out_f32, int64_t ld_f32, uint16_t *out_hi, uint16_t *out_lo,
                         int n_bf16, int64_t ld_bf16, void *stream) {
    LOCO_REQUIRE(M >= 0 && N > 0 && K > 0, LOCO_E_BADARG, "linear_fwd: bad shape M=%d N=%d K=%d", M, N, K);
    if (M == 0) return LOCO_OK;
    LOCO_REQUIRE(A_hi && W_hi, LOCO_E_BADARG, "linear_fwd: null operand");
    LOCO_REQUIRE((A_lo == nullptr) == (W_lo == nullptr), LOCO_E_BADARG, "linear_fwd: A_lo and W_lo must both be given (fp32-accurate mode) or both be NULL");
    LOCO_REQUIRE(out_f32 || out_hi, LOCO_E_BADARG, "linear_fwd: no output requested");
    LOCO_REQUIRE(!out_f32 || ld_f32 >= N, LOCO_E_BADARG, "linear_fwd: ld_f32 < N");
    if (out_hi) {
        LOCO_REQUIRE(n_bf16 > 0 && n_bf16 <= N && ld_bf16 >= n_bf16 && ld_bf16 % 8 == 0, LOCO_E_ALIGN, "linear_fwd: bf16 output needs 0 < n_bf16 <= N and ld_bf16 %% 8 == 0");
        LOCO_REQUIRE((reinterpret_cast<uintptr_t>(out_hi) & 15) == 0 && (!out_lo || (reinterpret_cast<uintptr_t>(out_lo) & 15) == 0), LOCO_E_ALIGN, "linear_fwd: bf16 outputs must be 16-byte aligned");
        LOCO_REQUIRE(n_bf16 % 2 == 0 || ld_bf16 > n_bf16, LOCO_E_ALIGN, "linear_fwd: odd n_bf16 needs a padded ld_bf16");
    }
    TcCore core = {};
    core.tf32 = tf32 ? 1 : 0;
    const int sms = current_device_sm_count();
    if (!pick_pair_shape(M, N, sms, core)) pick_gemm_shape(M, N, sms, core);
    EpiLinear::Params p;
    p.bias = bias; p.out_f32 = out_f32; p.ld_f32 = ld_f32; p.out_hi = out_hi; p.out_lo = out_lo;
    p.n_bf16 = out_hi ? n_bf16 : 0; p.ld_bf16 = ld_bf16; p.M = M; p.N = N;
    p.vec_f32 = (out_f32 != nullptr && (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0 && ld_f32 % 4 == 0) ? 1 : 0;
    p.tiles_n = core.clusters_n * core.cn;                                      // virtual (padded) tile grid
    const int tiles = tc_round_up((M + TC_BLOCK_M - 1) / TC_BLOCK_M, core.cm) * p.tiles_n;
    int grid = tiles, chunks = 1;
    p.pair_units = 0;
    static const bool pair_persist = []() { const char *e = getenv("LOCOV_B200_PAIR_PERSIST"); return !(e != nullptr && e[0] == '0'); }();
    if (core.two_cta && tiles / 2 > sms / 2 && pair_persist) {
        // persistent CTA pairs: one pair per two SMs strides over the 256-row tiles; two TMEM accumulator stages let the epilogue of
        // tile i overlap the MMAs of tile i + 1 (multi-wave pair grids paid set-up + epilogue per tile before)
        const int pairs_total = tiles / 2, slots = sms / 2;
        grid = 2 * slots;
        chunks = (pairs_total + slots - 1) / slots;
        core.total_tiles = pairs_total;
        p.pair_units = slots;
    } else if (core.cm * core.cn == 1 && tiles > sms) {
        // persistent: one CTA per SM strides over the tiles; two TMEM accumulator stages let the epilogue of tile i
        // overlap the MMAs of tile i + 1
        grid = sms;
        chunks = (tiles + grid - 1) / grid;
        core.total_tiles = tiles;
    }
    p.grid = grid;
    core.single_wave = grid <= sms ? 1 : 0;
    // bf16 outputs through TMA stores when the grid is a single wave (nothing else on the SM hides the epilogue, and the whole shared
    // memory is this CTA's); LOCOV_B200_EPI_TMA=0 keeps the direct stores for A/B measurements
    static const bool tma_allowed = []() { const char *e = getenv("LOCOV_B200_EPI_TMA"); return !(e != nullptr && e[0] == '0'); }();
    p.tma_out = (out_hi != nullptr && core.single_wave && tma_allowed) ? 1 : 0;
    const size_t smem = tc_finalize(core, K, A_lo ? 3 : 1, chunks, p.tma_out ? EpiLinear::kStageBytes : 0);
    TcMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc = fill_maps(maps, static_cast<const uint16_t *>(A_hi), static_cast<const uint16_t *>(A_lo), M, lda, static_cast<const uint16_t *>(W_hi),
                       static_cast<const uint16_t *>(W_lo), N, ldw, K, core);
    if (rc != LOCO_OK) return rc;
    if (p.tma_out) {
        if ((rc = make_tmap_bf16_out(&maps.o_hi, out_hi, (uint64_t)M, (uint64_t)p.n_bf16, (uint64_t)ld_bf16)) != LOCO_OK) return rc;
        if ((rc = make_tmap_bf16_out(&maps.o_lo, out_lo ? out_lo : out_hi, (uint64_t)M, (uint64_t)p.n_bf16, (uint64_t)ld_bf16)) != LOCO_OK) return rc;
    }
    if (core.two_cta) return tc_launch<EpiLinear, true>(maps, core, p, grid, smem, static_cast<cudaStream_t>(stream));
    return tc_launch<EpiLinear>(maps, core, p, grid, smem, static_cast<cudaStream_t>(stream));
}

int loco_linear_fwd(const uint16_t *A_hi, const uint16_t *A_lo, int64_t lda, const uint16_t *W_hi, const uint16_t *W_lo,
                    int64_t ldw, const float *bias, int M, int N, int K, float *out_f32, int64_t ld_f32,
                    uint16_t *out_hi, uint16_t *out_lo, int n_bf16, int64_t ld_bf16, void *stream) {
    return linear_launch(A_hi, A_lo, lda, W_hi, W_lo, ldw, false, bias, M, N, K, out_f32, ld_f32, out_hi, out_lo, n_bf16, ld_bf16, stream);
}

int loco_linear_tf32_fwd(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, int M, int N, int K,
                         float *out_f32, int64_t ld_f32, uint16_t *out_hi, uint16_t *out_lo, int n_bf16, int64_t ld_bf16,
                         void *stream) {
    return linear_launch(A, nullptr, lda, W, nullptr, ldw, true, bias, M, N, K, out_f32, ld_f32, out_hi, out_lo, n_bf16, ld_bf16, stream);
}

// column-chunk groups per 128-row tile: spread a wide class list over the SMs that the row tiles alone leave idle
static int box_score_groups(int R, int K1) {
    if (K1 <= 256) return 1;
    const int chunks = (K1 + 255) / 256;
    const int tiles = (R + TC_BLOCK_M - 1) / TC_BLOCK_M;
    int g = current_device_sm_count() / (tiles > 0 ? tiles : 1);
    if (g < 1) g = 1;
    if (g > chunks) g = chunks;
    if (g > 32) g = 32;
    return g;
}

int64_t loco_box_score_workspace_bytes(int R, int K1) {
    const int g = box_score_groups(R, K1);
    return g > 1 ? (int64_t)R * g * 16 : 16;
}

int loco_box_score_fwd(const uint16_t *E_hi, const uint16_t *E_lo, int64_t lde, const uint16_t *C_hi, const uint16_t *C_lo,
                       int64_t ldc, const float *cls_bias, int R, int K1, int D, float *logits, float *probs,
                       int64_t ld_logits, float *lse, int64_t *argmax_fg, void *workspace, void *stream) {
    LOCO_REQUIRE(R >= 0 && K1 >= 1 && D > 0, LOCO_E_BADARG, "box_score_fwd: bad shape R=%d K1=%d D=%d", R, K1, D);
    if (R == 0) return LOCO_OK;
    LOCO_REQUIRE(E_hi && C_hi && logits, LOCO_E_BADARG, "box_score_fwd: null pointer");
    LOCO_REQUIRE((E_lo == nullptr) == (C_lo == nullptr), LOCO_E_BADARG, "box_score_fwd: E_lo and C_lo must both be given or both be NULL");
    LOCO_REQUIRE(ld_logits >= K1, LOCO_E_BADARG, "box_score_fwd: ld_logits < K1");
    LOCO_REQUIRE(probs == nullptr || lse != nullptr, LOCO_E_BADARG, "box_score_fwd: probs need the lse output");
    TcCore core = {};
    // one N tile when the class list fits (<= 256), otherwise 256-wide chunks with online softmax, dealt round-robin
    // to `groups` CTAs per row tile (partial statistics combined by box_score_finalize_kernel)
    core.block_n = K1 <= 256 ? tc_round_up(K1, 32) : 256;
    const int total_chunks = (K1 + core.block_n - 1) / core.block_n;
    const int groups = box_score_groups(R, K1);
    LOCO_REQUIRE(groups == 1 || workspace != nullptr, LOCO_E_BADARG, "box_score_fwd: K1 > 256 needs loco_box_score_workspace_bytes() of workspace");
    const int chunks = (total_chunks + groups - 1) / groups;
    const size_t smem = tc_finalize(core, D, E_lo ? 3 : 1, chunks,
                                    4 * TC_WARP_SCRATCH_WORDS * 4 + (cls_bias ? (K1 < EpiScore::kBiasCap ? K1 : EpiScore::kBiasCap) * 4 : 0));
    TcMaps maps;
    int rc = fill_maps(maps, E_hi, E_lo, R, lde, C_hi, C_lo, K1, ldc, D, core);
    if (rc != LOCO_OK) return rc;
    EpiScore::Params p;
    p.bias = cls_bias; p.logits = logits; p.ld = ld_logits; p.lse = lse; p.argmax_fg = argmax_fg;
    p.part = groups > 1 ? static_cast<float *>(workspace) : nullptr;
    p.R = R; p.K1 = K1; p.groups = groups;
    p.vec = ((reinterpret_cast<uintptr_t>(logits) & 15) == 0 && ld_logits % 4 == 0) ? 1 : 0;
    const int grid = ((R + TC_BLOCK_M - 1) / TC_BLOCK_M) * groups;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = tc_launch<EpiScore>(maps, core, p, grid, smem, st);
    if (rc != LOCO_OK) return rc;
    if (groups > 1 || probs != nullptr) {
        int blocks = (R + 7) / 8;
        const int cap = current_device_sm_count() * 8;
        if (blocks > cap) blocks = cap;
        LOCO_CUDA(launch_kernel(box_score_finalize_kernel, dim3(blocks), dim3(256), 0, st, 1, logits, probs, ld_logits, R, K1,
                                groups > 1 ? static_cast<const float4 *>(workspace) : static_cast<const float4 *>(nullptr), groups, lse, argmax_fg));
        count_launch();
        LOCO_CUDA(cudaGetLastError());
    }
    return LOCO_OK;
}

int64_t loco_lsm_pair_workspace_bytes(int Bc, int T, int Bi, int Rg) {
    (void)Bc; (void)T; (void)Bi; (void)Rg;
    return 16;   // reserved (all reductions happen on-chip); kept so callers always pass a valid pointer
}

static int lsm_launch(bool bwd, const uint16_t *cap_hi, const uint16_t *cap_lo, int64_t ldcap, const uint16_t *emb_hi,
                      const uint16_t *emb_lo, int64_t ldemb, int D, LsmParams &p, cudaStream_t st) {
    LOCO_REQUIRE(p.Bc >= 0 && p.Bi >= 0 && p.T > 0 && p.Rg > 0 && D > 0, LOCO_E_BADARG, "lsm_pair: bad shape Bc=%d T=%d Bi=%d Rg=%d D=%d", p.Bc, p.T, p.Bi, p.Rg, D);
    LOCO_REQUIRE(p.T <= 128 && p.Rg <= 256, LOCO_E_UNSUPPORTED, "lsm_pair: supports T <= 128 words and Rg <= 256 regions (got T=%d Rg=%d)", p.T, p.Rg);
    LOCO_REQUIRE(cap_hi && emb_hi && p.cap_mask && p.reg_mask, LOCO_E_BADARG, "lsm_pair: null pointer");
    LOCO_REQUIRE((cap_lo == nullptr) == (emb_lo == nullptr), LOCO_E_BADARG, "lsm_pair: cap_lo and emb_lo must both be given or both be NULL");
    TcCore core = {};
    core.block_n = tc_round_up(p.Rg, 32);      // whole 32-column epilogue blocks: every accumulator column read is written by the MMA
    p.per_tile = TC_BLOCK_M / p.T;
    if (p.per_tile > LSM_MAX_PER_TILE) p.per_tile = LSM_MAX_PER_TILE;
    if (p.per_tile > p.Bc) p.per_tile = p.Bc;
    p.lds = core.block_n | 1;
    int groups = (p.Bc + p.per_tile - 1) / p.per_tile;
    // cluster = cm caption groups x cn images: the caption tile is multicast to the cn image CTAs of its row, the
    // region tile of an image to the cm caption-group CTAs of its column
    core.cm = (clusters_enabled() && groups >= 2) ? 2 : 1;
    core.cn = (clusters_enabled() && p.Bi >= 4) ? 4 : ((clusters_enabled() && p.Bi >= 2) ? 2 : 1);
    groups = tc_round_up(groups, core.cm);
    p.Bi_pad = tc_round_up(p.Bi, core.cn);
    core.clusters_n = p.Bi_pad / core.cn;
    const size_t epi = lsm_epi_smem_bytes(p.lds, core.block_n, p.per_tile, p.Rg, bwd);
    core.epi_overlay = 1;      // the parked tile re-uses the operand ring once the MMAs are done: two CTAs fit an SM
    const size_t smem = tc_finalize(core, D, cap_lo ? 3 : 1, 1, (int)epi);
    TcMaps maps;
    int rc = fill_maps(maps, cap_hi, cap_lo, (uint64_t)p.Bc * p.T, ldcap, emb_hi, emb_lo, (uint64_t)p.Bi * p.Rg, ldemb, D, core);
    if (rc != LOCO_OK) return rc;
    LOCO_REQUIRE((long long)groups * p.Bi_pad < (1ll << 31), LOCO_E_UNSUPPORTED, "lsm_pair: too many tiles");
    const int grid = groups * p.Bi_pad;
    return bwd ? tc_launch<EpiLsm<true>>(maps, core, p, grid, smem, st) : tc_launch<EpiLsm<false>>(maps, core, p, grid, smem, st);
}

// LOCOV_B200_LSM_V2=0 keeps the first-generation forward kernel (one CTA per (caption group, image) tile) for A/B runs.
static bool lsm_v2_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("LOCOV_B200_LSM_V2");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

// second-generation forward (softmax alignment): persistent CTA pairs, see EpiLsmFwd.  Returns LOCO_E_UNSUPPORTED (without an
// error message of its own) when no tile shape fits shared / tensor memory: the caller then uses the first-generation kernel.
static int lsm_fwd2_launch(const uint16_t *cap_hi, const uint16_t *cap_lo, int64_t ldcap, const uint16_t *emb_hi, const uint16_t *emb_lo,
                           int64_t ldemb, int D, LsmFwdParams &p, cudaStream_t st, bool *declined) {
    *declined = false;
    const int sms = current_device_sm_count();
    int per_max = TC_BLOCK_M / p.T;
    if (per_max > LSM_MAX_PER_TILE) per_max = LSM_MAX_PER_TILE;
    if (per_max > p.Bc) per_max = p.Bc;
    if (per_max < 1) per_max = 1;
    p.Tp = tc_round_up(p.T, 4);
    p.rb = tc_round_up(p.Rg, 32);
    // images per tile: as many as fit 256 accumulator columns (two accumulator stages = 512 TMEM columns), at most 16; fewer when the
    // epilogue's reads would leave the allocation or the parked sub-tiles leave too little shared memory for the operand ring
    int ipt = 256 / p.Rg;
    if (ipt < 1) ipt = 1;
    if (ipt > p.Bi) ipt = p.Bi;
    if (ipt > 16) ipt = 16;
    if (const char *e = getenv("LOCOV_B200_LSM_IPT")) { const int v = atoi(e); if (v >= 1 && v <= ipt) ipt = v; }   // developer sweep knob
    TcCore core;
    size_t smem = 0;
    bool fixed_ldt = false;
    for (;; --ipt) {
        if (ipt < 1) { *declined = true; return LOCO_E_UNSUPPORTED; }
        p.ipt = ipt;
        p.slots = (ipt + 1) / 2;
        p.halves = ipt > 1 ? 2 : 1;
        p.tiles_i = (p.Bi + ipt - 1) / ipt;
        // captions per CTA: the most that fit 128 accumulator rows — unless fewer still give every pair a single tile: then the epilogue
        // (the longer half of a lone tile's life) is spread over more SMs
        p.per_tile = per_max;
        for (int pt = 1; pt < per_max; ++pt) {
            const long long tot = (long long)(((p.Bc + pt - 1) / pt + 1) / 2) * p.tiles_i;
            if (tot <= sms / 2) { p.per_tile = pt; break; }
        }
        if (const char *e = getenv("LOCOV_B200_LSM_PT")) { const int v = atoi(e); if (v >= 1 && v <= per_max) p.per_tile = v; }   // developer sweep knob
        p.ldt = tc_round_up(p.per_tile * p.Tp, 4);
        if (((p.ldt / 4) & 1) == 0) p.ldt += 4;
        fixed_ldt = p.ldt <= LSM_LDT;
        if (fixed_ldt) p.ldt = LSM_LDT;
        const int groups = (p.Bc + p.per_tile - 1) / p.per_tile;
        const long long total = (long long)((groups + 1) / 2) * p.tiles_i;
        LOCO_REQUIRE(total < (1ll << 30), LOCO_E_UNSUPPORTED, "lsm_pair: too many tiles");
        core = TcCore{};
        core.block_n = tc_round_up(ipt * p.Rg, 16);
        core.cm = 2; core.cn = 1; core.two_cta = 1;
        int npairs = sms / 2;
        if (npairs > total) npairs = (int)total;
        if (npairs < 1) npairs = 1;
        p.npairs = npairs;
        core.clusters_n = npairs;                       // virtual CTA index = rm * npairs + pair (see tc_gemm_kernel)
        core.total_tiles = (int)total;
        core.single_wave = 1;
        const int chunks = (int)((total + npairs - 1) / npairs);
        const size_t big = 2 * EpiLsmFwd<0>::big_floats(p.Rg, p.ldt, p.per_tile) * sizeof(float);
        core.epi_tail = (int)(2 * EpiLsmFwd<0>::small_floats(p.slots, p.rb) * sizeof(float) + 16);
        if (big + core.epi_tail > 200 * 1024) continue;
        // a lone tile per pair: its parked sub-tiles re-use the operand ring (dead once the accumulator is complete), the ring gets all
        // of shared memory; several tiles: separate scratch, two accumulator stages
        core.epi_overlay = chunks == 1 ? 1 : 0;
        if (const char *e = getenv("LOCOV_B200_LSM_OVERLAY")) { if (e[0] == '0') core.epi_overlay = 0; }
        smem = tc_finalize(core, D, cap_lo ? 3 : 1, core.epi_overlay ? 1 : (chunks > 1 ? chunks : 2), (int)big);
        // every accumulator column the epilogue reads must lie inside the allocation (worst case: the last image read as 32-wide blocks)
        const int reach = (core.acc_stages - 1) * core.block_n + (ipt - 1) * p.Rg + tc_round_up(p.Rg, 32);
        while (core.tmem_cols < reach && core.tmem_cols < 512) core.tmem_cols <<= 1;
        if (reach > core.tmem_cols || core.stages < 2 || smem > 227 * 1024) continue;
        break;
    }
    TcMaps maps;
    int rc = fill_maps(maps, cap_hi, cap_lo, (uint64_t)p.Bc * p.T, ldcap, emb_hi, emb_lo, (uint64_t)p.Bi * p.Rg, ldemb, D, core);
    if (rc != LOCO_OK) return rc;
    if (fixed_ldt) return tc_launch<EpiLsmFwd<LSM_LDT>, true>(maps, core, p, 2 * p.npairs, smem, st);
    return tc_launch<EpiLsmFwd<0>, true>(maps, core, p, 2 * p.npairs, smem, st);
}

int loco_lsm_pair_fwd(const uint16_t *cap_hi, const uint16_t *cap_lo, int64_t ldcap, const float *cap_mask,
                      const uint16_t *emb_hi, const uint16_t *emb_lo, int64_t ldemb, const float *reg_mask, int Bc, int T,
                      int Bi, int Rg, int D, float inv_temperature, int alignment, float *d_w2r, float *d_r2w,
                      int64_t ld_out, void *workspace, void *stream) {
    (void)workspace;
    if (Bc == 0 || Bi == 0) return LOCO_OK;
    LOCO_REQUIRE(T <= 128 && Rg <= 256, LOCO_E_UNSUPPORTED, "lsm_pair_fwd: supports T <= 128 words and Rg <= 256 regions (got T=%d Rg=%d)", T, Rg);
    LOCO_REQUIRE(alignment == LOCO_ALIGN_SOFTMAX || alignment == LOCO_ALIGN_HARDMAX, LOCO_E_UNSUPPORTED, "lsm_pair_fwd: alignment %d not implemented", alignment);
    LOCO_REQUIRE(d_w2r || d_r2w, LOCO_E_BADARG, "lsm_pair_fwd: no output requested");
    LOCO_REQUIRE(ld_out >= Bi, LOCO_E_BADARG, "lsm_pair_fwd: ld_out < Bi");
    if (alignment == LOCO_ALIGN_SOFTMAX && lsm_v2_enabled() && current_device_sm_count() >= 2) {
        LOCO_REQUIRE(cap_hi && emb_hi && cap_mask && reg_mask, LOCO_E_BADARG, "lsm_pair_fwd: null pointer");
        LOCO_REQUIRE((cap_lo == nullptr) == (emb_lo == nullptr), LOCO_E_BADARG, "lsm_pair_fwd: cap_lo and emb_lo must both be given or both be NULL");
        LOCO_REQUIRE(T > 0 && Rg > 0 && D > 0, LOCO_E_BADARG, "lsm_pair_fwd: bad shape T=%d Rg=%d D=%d", T, Rg, D);
        LsmFwdParams q = {};
        q.cap_mask = cap_mask; q.reg_mask = reg_mask; q.out_w2r = d_w2r; q.out_r2w = d_r2w; q.ld_out = ld_out;
        q.Bc = Bc; q.T = T; q.Bi = Bi; q.Rg = Rg; q.inv_temp = inv_temperature;
        bool declined = false;
        const int rc2 = lsm_fwd2_launch(cap_hi, cap_lo, ldcap, emb_hi, emb_lo, ldemb, D, q, static_cast<cudaStream_t>(stream), &declined);
        if (!declined) return rc2;        // (declined: no tile shape fits — the first-generation kernel below takes it)
    }
    LsmParams p = {};
    p.cap_mask = cap_mask; p.reg_mask = reg_mask; p.out_w2r = d_w2r; p.out_r2w = d_r2w; p.ld_out = ld_out;
    p.Bc = Bc; p.T = T; p.Bi = Bi; p.Rg = Rg; p.inv_temp = inv_temperature; p.hardmax = (alignment == LOCO_ALIGN_HARDMAX);
    return lsm_launch(false, cap_hi, cap_lo, ldcap, emb_hi, emb_lo, ldemb, D, p, static_cast<cudaStream_t>(stream));
}

int loco_lsm_pair_bwd(const uint16_t *cap_hi, const uint16_t *cap_lo, int64_t ldcap, const float *cap_mask,
                      const uint16_t *emb_hi, const uint16_t *emb_lo, int64_t ldemb, const float *reg_mask, int Bc, int T,
                      int Bi, int Rg, int D, float inv_temperature, int alignment, const float *g_w2r, const float *g_r2w,
                      int64_t ld_g, uint16_t *dst_hi, uint16_t *dst_lo, int64_t ld_dst, uint16_t *ds_hi, uint16_t *ds_lo,
                      int64_t ld_ds, void *stream) {
    if (Bc == 0 || Bi == 0) return LOCO_OK;
    LOCO_REQUIRE(alignment == LOCO_ALIGN_SOFTMAX || alignment == LOCO_ALIGN_HARDMAX, LOCO_E_UNSUPPORTED, "lsm_pair_bwd: alignment %d not implemented", alignment);
    LOCO_REQUIRE(g_w2r || g_r2w, LOCO_E_BADARG, "lsm_pair_bwd: no upstream gradient given");
    LOCO_REQUIRE(ld_g >= Bi, LOCO_E_BADARG, "lsm_pair_bwd: ld_g < Bi");
    LOCO_REQUIRE(dst_hi != nullptr && ld_dst >= (int64_t)Bc * T, LOCO_E_BADARG, "lsm_pair_bwd: dS^T buffer missing or ld_dst < Bc*T");
    LOCO_REQUIRE(ds_hi == nullptr || ld_ds >= (int64_t)Bi * Rg, LOCO_E_BADARG, "lsm_pair_bwd: ld_ds < Bi*Rg");
    LsmParams p = {};
    p.cap_mask = cap_mask; p.reg_mask = reg_mask; p.g_w2r = g_w2r; p.g_r2w = g_r2w; p.ld_g = ld_g;
    p.dst_hi = dst_hi; p.dst_lo = dst_lo; p.ld_dst = ld_dst; p.ds_hi = ds_hi; p.ds_lo = ds_lo; p.ld_ds = ld_ds;
    p.Bc = Bc; p.T = T; p.Bi = Bi; p.Rg = Rg; p.inv_temp = inv_temperature; p.hardmax = (alignment == LOCO_ALIGN_HARDMAX);
    return lsm_launch(true, cap_hi, cap_lo, ldcap, emb_hi, emb_lo, ldemb, D, p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
