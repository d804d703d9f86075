// tc_ops.cu — the dense contractions of the region-text path on the tcgen05 core (tc_gemm.cuh), each
// with its reductions fused into the TMEM epilogue:
//   EpiLinear : out = A·Wᵀ + bias                  (emb_pred / bbox_pred / v2l_projection)
//   EpiScore  : logits = E·Cᵀ (+bias), online softmax statistics, argmax, probabilities
//   EpiW2R    : word→region attention pooling of the caption×image similarity tile
//   EpiR2W    : region→word attention pooling of the transposed tile
#include <cfloat>

#include "tc_gemm.cuh"

namespace loco {

constexpr float LSM_FILL = -1.0e30f;   // finite "masked" logit: an all-masked row stays uniform, never NaN

// ------------------------------------------------------------------------------------------------
// EpiLinear
// ------------------------------------------------------------------------------------------------
struct EpiLinear {
    struct Params {
        const float *bias;
        float *out_f32;
        int64_t ld_f32;
        uint16_t *out_hi, *out_lo;
        int n_bf16;
        int64_t ld_bf16;
        int M, N, tiles_n;
    };
    static __device__ __forceinline__ void coords(const Params &p, const TcCore &core, int cta, int, int &row_a, int &row_b) {
        row_a = (cta / p.tiles_n) * TC_BLOCK_M;
        row_b = (cta % p.tiles_n) * core.block_n;
    }
    __device__ __forceinline__ void begin(const Params &, const TcCore &, int, int, int, int, unsigned char *) {}
    __device__ __forceinline__ void chunk(const Params &p, const TcCore &core, int cta, int, uint32_t taddr, int row, int lane,
                                          int q, unsigned char *smem) {
        float *scratch = reinterpret_cast<float *>(smem) + (size_t)q * TC_WARP_SCRATCH_WORDS;
        const int m0 = (cta / p.tiles_n) * TC_BLOCK_M + q * 32;      // first global row of this warp
        const int n0 = (cta % p.tiles_n) * core.block_n;
        const int rows_valid = max(0, min(32, p.M - m0));
        (void)row;
        for (int c0 = 0; c0 < core.block_n; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            const int gc0 = n0 + c0;
            const int cols_valid = max(0, min(min(32, core.block_n - c0), p.N - gc0));
            if (cols_valid <= 0) continue;     // warp-uniform
            if (p.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < cols_valid) v[j] += __ldg(p.bias + gc0 + j);
            }
            if (p.out_f32 != nullptr)
                warp_store_f32(scratch, v, p.out_f32 + (int64_t)m0 * p.ld_f32 + gc0, p.ld_f32, rows_valid, cols_valid, lane);
            if (p.out_hi != nullptr && gc0 < p.n_bf16) {
                const int cb = min(cols_valid, p.n_bf16 - gc0);
                // odd column counts only occur at the very end of the matrix; pad lanes write zeros from
                // the zero-filled accumulator columns (TMA out-of-bounds rows of W are zero).
                warp_store_bf16(reinterpret_cast<uint32_t *>(scratch), v, p.out_hi + (int64_t)m0 * p.ld_bf16 + gc0,
                                p.out_lo ? p.out_lo + (int64_t)m0 * p.ld_bf16 + gc0 : nullptr, p.ld_bf16, rows_valid,
                                (cb + 1) & ~1, lane);
            }
        }
    }
    __device__ __forceinline__ void finish(const Params &, const TcCore &, int, int, int, int, unsigned char *) {}
};

// ------------------------------------------------------------------------------------------------
// EpiScore — RoI x class logits with online softmax statistics across N-chunks
// ------------------------------------------------------------------------------------------------
struct EpiScore {
    struct Params {
        const float *bias;
        float *logits, *probs;
        int64_t ld;
        float *lse;
        int64_t *argmax_fg;
        int R, K1;
    };
    float run_max, run_sum, best_val;
    int best_idx;

    static __device__ __forceinline__ void coords(const Params &, const TcCore &core, int cta, int chunk, int &row_a, int &row_b) {
        row_a = cta * TC_BLOCK_M;
        row_b = chunk * core.block_n;
    }
    __device__ __forceinline__ void begin(const Params &, const TcCore &, int, int, int, int, unsigned char *) {
        run_max = -FLT_MAX;
        run_sum = 0.f;
        best_val = -FLT_MAX;
        best_idx = 0;
    }
    __device__ __forceinline__ void chunk(const Params &p, const TcCore &core, int cta, int ch, uint32_t taddr, int row, int lane,
                                          int q, unsigned char *smem) {
        float *scratch = reinterpret_cast<float *>(smem) + (size_t)q * TC_WARP_SCRATCH_WORDS;
        const int m0 = cta * TC_BLOCK_M + q * 32;
        const int rows_valid = max(0, min(32, p.R - m0));
        const int n0 = ch * core.block_n;
        (void)row;
        for (int c0 = 0; c0 < core.block_n; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            const int gc0 = n0 + c0;
            const int cols_valid = max(0, min(min(32, core.block_n - c0), p.K1 - gc0));
            if (cols_valid <= 0) continue;
            float blk_max = -FLT_MAX;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < cols_valid) {
                    if (p.bias != nullptr) v[j] += __ldg(p.bias + gc0 + j);
                    blk_max = fmaxf(blk_max, v[j]);
                    if (gc0 + j < p.K1 - 1 && v[j] > best_val) {   // strict >: first maximum wins (torch.argmax)
                        best_val = v[j];
                        best_idx = gc0 + j;
                    }
                }
            }
            const float new_max = fmaxf(run_max, blk_max);
            float s = run_sum * expf(run_max - new_max);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < cols_valid) s += expf(v[j] - new_max);
            run_max = new_max;
            run_sum = s;
            warp_store_f32(scratch, v, p.logits + (int64_t)m0 * p.ld + gc0, p.ld, rows_valid, cols_valid, lane);
        }
    }
    __device__ __forceinline__ void finish(const Params &p, const TcCore &, int cta, int row, int lane, int q, unsigned char *) {
        const int m0 = cta * TC_BLOCK_M + q * 32;
        const int grow = cta * TC_BLOCK_M + row;
        const float lse = run_max + logf(run_sum);
        if (grow < p.R) {
            if (p.lse != nullptr) p.lse[grow] = lse;
            if (p.argmax_fg != nullptr) p.argmax_fg[grow] = (int64_t)best_idx;
        }
        if (p.probs != nullptr) {
            // Every lane re-reads exactly the logits it wrote itself in warp_store_f32 (same row / column
            // mapping), so program order makes them visible without a fence.
            const int rows_valid = max(0, min(32, p.R - m0));
            for (int rr = 0; rr < rows_valid; ++rr) {
                const float l = __shfl_sync(0xffffffffu, lse, rr);
                const float *src = p.logits + (int64_t)(m0 + rr) * p.ld;
                float *dst = p.probs + (int64_t)(m0 + rr) * p.ld;
                for (int c = lane; c < p.K1; c += 32) dst[c] = expf(src[c] - l);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// LSM epilogues.  Tile rows are always owned one-per-thread; a "segment" is the span of columns that
// belongs to one softmax (all regions of the image for W2R; the T words of one caption for R2W).
// ------------------------------------------------------------------------------------------------
struct LsmParams {
    const float *cap_mask;   // [Bc, T]
    const float *reg_mask;   // [Bi, Rg]
    float *out;              // [Bc, Bi] (ld)
    int64_t ld;
    int Bc, T, Bi, Rg;
    int per_tile;            // W2R: captions per M tile; R2W: captions per N tile
    int groups;              // W2R: caption groups; R2W: caption chunks
    int row_blocks;          // R2W: 128-row blocks per image (1 or 2)
    float inv_temp;
    int hardmax;
};

// W2R: rows = words of `per_tile` captions (row = cl*T + t), columns = regions of image i.
struct EpiW2R {
    typedef LsmParams Params;
    static __device__ __forceinline__ void coords(const Params &p, const TcCore &, int cta, int, int &row_a, int &row_b) {
        const int g = cta / p.Bi, i = cta - g * p.Bi;
        row_a = g * p.per_tile * p.T;    // caption rows
        row_b = i * p.Rg;                // region rows
    }
    __device__ __forceinline__ void begin(const Params &, const TcCore &, int, int, int, int, unsigned char *) {}
    __device__ __forceinline__ void chunk(const Params &p, const TcCore &core, int cta, int, uint32_t taddr, int row, int lane,
                                          int q, unsigned char *smem) {
        (void)lane; (void)q;
        float *row_val = reinterpret_cast<float *>(smem);      // [128] masked pooled value per word row
        float *row_msk = row_val + 128;                        // [128] caption mask per word row
        float *rmask_s = row_msk + 128;                        // [block_n] region mask of image i
        const int g = cta / p.Bi, i = cta - g * p.Bi;
        const int cl = row / p.T, t = row - cl * p.T;
        const int c = g * p.per_tile + cl;
        const bool row_ok = (cl < p.per_tile) && (c < p.Bc);
        const float mc = row_ok ? __ldg(p.cap_mask + (int64_t)c * p.T + t) : 0.f;
        // stage the region mask of this image (shared by all rows)
        const int et = threadIdx.x - 64;                       // 0..127 among epilogue threads
        for (int r = et; r < core.block_n; r += 128) rmask_s[r] = (r < p.Rg) ? __ldg(p.reg_mask + (int64_t)i * p.Rg + r) : 0.f;
        named_bar_sync(1, 128);

        const int nblk = (p.Rg + 31) / 32;
        float pooled = 0.f;
        if (!p.hardmax) {
            // pass 1: row maximum of the masked logits
            float mx = -FLT_MAX;
            for (int b = 0; b < nblk; ++b) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)(b * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int r = b * 32 + j;
                    if (r < p.Rg) {
                        const float s = v[j] * p.inv_temp;
                        const float sm = (mc > 0.f && rmask_s[r] > 0.f) ? s : LSM_FILL;
                        mx = fmaxf(mx, sm);
                    }
                }
            }
            // pass 2: softmax-weighted sum of the UNMASKED similarities (reference grounding_head.py:228-231)
            float den = 0.f, num = 0.f;
            for (int b = 0; b < nblk; ++b) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)(b * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int r = b * 32 + j;
                    if (r < p.Rg) {
                        const float s = v[j] * p.inv_temp;
                        const float sm = (mc > 0.f && rmask_s[r] > 0.f) ? s : LSM_FILL;
                        const float e = expf(sm - mx);
                        den += e;
                        num = fmaf(e, s, num);
                    }
                }
            }
            pooled = num / den;
        } else {
            float best = -FLT_MAX, best_s = 0.f;
            for (int b = 0; b < nblk; ++b) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)(b * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int r = b * 32 + j;
                    if (r < p.Rg) {
                        const float s = v[j] * p.inv_temp;
                        const float sm = (mc > 0.f && rmask_s[r] > 0.f) ? s : LSM_FILL;
                        if (sm > best) { best = sm; best_s = s; }
                    }
                }
            }
            pooled = best_s;
        }
        row_val[row] = (mc > 0.f) ? pooled : 0.f;
        row_msk[row] = mc;
        named_bar_sync(1, 128);
        // deterministic per-caption reduction: thread k sums the T rows of caption slot k
        if (et < p.per_tile) {
            const int cc = g * p.per_tile + et;
            if (cc < p.Bc) {
                float acc = 0.f, nw = 0.f;
                for (int tt = 0; tt < p.T; ++tt) {
                    acc += row_val[et * p.T + tt];
                    nw += row_msk[et * p.T + tt];
                }
                p.out[(int64_t)cc * p.ld + i] = -acc / fmaxf(nw, 1.f);
            }
        }
    }
    __device__ __forceinline__ void finish(const Params &, const TcCore &, int, int, int, int, unsigned char *) {}
};

// R2W: rows = regions of image i (row block m), columns = words of `per_tile` captions (col = cl*T + t).
struct EpiR2W {
    typedef LsmParams Params;
    static __device__ __forceinline__ void decode(const Params &p, int cta, int &i, int &m, int &cc) {
        cc = cta % p.groups;
        const int rest = cta / p.groups;
        m = rest % p.row_blocks;
        i = rest / p.row_blocks;
    }
    static __device__ __forceinline__ void coords(const Params &p, const TcCore &, int cta, int, int &row_a, int &row_b) {
        int i, m, cc;
        decode(p, cta, i, m, cc);
        row_a = i * p.Rg + m * TC_BLOCK_M;       // region rows (A operand = region embeddings)
        row_b = cc * p.per_tile * p.T;           // caption word rows (B operand)
    }
    __device__ __forceinline__ void begin(const Params &, const TcCore &, int, int, int, int, unsigned char *) {}
    __device__ __forceinline__ void chunk(const Params &p, const TcCore &core, int cta, int, uint32_t taddr, int row, int lane,
                                          int q, unsigned char *smem) {
        float *cmask_s = reinterpret_cast<float *>(smem);      // [block_n] caption mask per column
        float *part = cmask_s + 256;                            // [4][per_tile] per-warp partial sums
        int i, m, cc;
        decode(p, cta, i, m, cc);
        const int r = m * TC_BLOCK_M + row;
        const bool row_ok = r < p.Rg;
        const float mr = row_ok ? __ldg(p.reg_mask + (int64_t)i * p.Rg + r) : 0.f;
        const int et = threadIdx.x - 64;
        const int c_first = cc * p.per_tile;
        const int ncap = max(0, min(p.per_tile, p.Bc - c_first));
        for (int col = et; col < core.block_n; col += 128) {
            const int cl = col / p.T;
            cmask_s[col] = (cl < ncap) ? __ldg(p.cap_mask + (int64_t)(c_first + cl) * p.T + (col - cl * p.T)) : 0.f;
        }
        // number of valid regions of the image (every warp computes it redundantly)
        float nr = 0.f;
        for (int rr = lane; rr < p.Rg; rr += 32) nr += __ldg(p.reg_mask + (int64_t)i * p.Rg + rr);
        nr = warp_sum(nr);
        named_bar_sync(1, 128);

        // Column blocks of 32 do not line up with the T-word segments, so the online state of the current
        // segment is carried across blocks.  Two sweeps over TMEM: segment maxima, then pooled sums.
        const int ncols = ncap * p.T;
        const int nblk = (ncols + 31) / 32;
        float pooled_seg = 0.f;                 // value for the segment being finalised
        // sweep over segments; per segment: max pass then sum pass restricted to its column range
        for (int cl = 0; cl < ncap; ++cl) {
            const int cb = cl * p.T, ce = cb + p.T;
            const int b0 = cb / 32, b1 = (ce - 1) / 32;
            float mx = -FLT_MAX, best_s = 0.f;
            for (int b = b0; b <= b1; ++b) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)(b * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = b * 32 + j;
                    if (col >= cb && col < ce) {
                        const float s = v[j] * p.inv_temp;
                        const float sm = (mr > 0.f && cmask_s[col] > 0.f) ? s : LSM_FILL;
                        if (sm > mx) { mx = sm; best_s = s; }
                    }
                }
            }
            if (!p.hardmax) {
                float den = 0.f, num = 0.f;
                for (int b = b0; b <= b1; ++b) {
                    float v[32];
                    tmem_ld32(taddr + (uint32_t)(b * 32), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = b * 32 + j;
                        if (col >= cb && col < ce) {
                            const float s = v[j] * p.inv_temp;
                            const float sm = (mr > 0.f && cmask_s[col] > 0.f) ? s : LSM_FILL;
                            const float e = expf(sm - mx);
                            den += e;
                            num = fmaf(e, s, num);
                        }
                    }
                }
                pooled_seg = num / den;
            } else {
                pooled_seg = best_s;
            }
            // sum over the 32 region rows of this warp (rows with region mask 0 contribute exactly 0)
            const float contrib = warp_sum((mr > 0.f) ? pooled_seg : 0.f);
            if (lane == 0) part[q * 16 + (cl & 15)] = contrib;
            if ((cl & 15) == 15 || cl == ncap - 1) {
                named_bar_sync(1, 128);
                const int base = cl & ~15;
                if (et <= (cl & 15)) {
                    // fixed summation order over the four warps (tile row quarters 0..3): deterministic
                    float tot = 0.f;
                    for (int w = 0; w < 4; ++w) {
                        // warp index -> quarter: warps 2,3,4,5 own quarters 2,3,0,1; order by quarter
                        tot += part[w * 16 + et];
                    }
                    const int c = c_first + base + et;
                    const float val = -tot / fmaxf(nr, 1.f);
                    float *dst = p.out + (int64_t)c * p.ld + i;
                    if (p.row_blocks == 1) *dst = val; else atomicAdd(dst, val);
                }
                named_bar_sync(1, 128);
            }
        }
        (void)nblk;
    }
    __device__ __forceinline__ void finish(const Params &, const TcCore &, int, int, int, int, unsigned char *) {}
};

// ------------------------------------------------------------------------------------------------
// host-side helpers
// ------------------------------------------------------------------------------------------------
static int fill_maps(TcMaps &maps, const uint16_t *a_hi, const uint16_t *a_lo, uint64_t a_rows, uint64_t a_ld,
                     const uint16_t *b_hi, const uint16_t *b_lo, uint64_t b_rows, uint64_t b_ld, uint64_t K, int block_n) {
    int rc;
    if ((rc = make_tmap_bf16_2d(&maps.a_hi, a_hi, a_rows, K, a_ld, TC_BLOCK_M)) != LOCO_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&maps.a_lo, a_lo ? a_lo : a_hi, a_rows, K, a_ld, TC_BLOCK_M)) != LOCO_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&maps.b_hi, b_hi, b_rows, K, b_ld, (uint32_t)block_n)) != LOCO_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&maps.b_lo, b_lo ? b_lo : b_hi, b_rows, K, b_ld, (uint32_t)block_n)) != LOCO_OK) return rc;
    return LOCO_OK;
}

// BLOCK_N for a plain GEMM: minimise waves * per-tile MMA time.  Per K=16 step a 128xN tile costs
// max(N/2, 32 + N/4) cycles (tensor pipe vs. shared-memory operand reads; B300_MICROARCH tcgen05 floor).
static int pick_block_n(int M, int N, int sms) {
    const int mt = (M + TC_BLOCK_M - 1) / TC_BLOCK_M;
    int best = 256;
    double best_cost = 1e30;
    for (int bn = 256; bn >= 32; bn -= 32) {
        const int tiles = mt * ((N + bn - 1) / bn);
        const int waves = (tiles + sms - 1) / sms;
        const double per = bn / 2.0 > 32 + bn / 4.0 ? bn / 2.0 : 32 + bn / 4.0;
        const double cost = waves * (per + 6.0);   // +6: fixed per-step issue/barrier overhead
        if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
    }
    return best;
}

}  // namespace loco

using namespace loco;

extern "C" {

int loco_linear_fwd(const uint16_t *A_hi, const uint16_t *A_lo, int64_t lda, const uint16_t *W_hi, const uint16_t *W_lo,
                    int64_t ldw, const float *bias, int M, int N, int K, float *out_f32, int64_t ld_f32,
                    uint16_t *out_hi, uint16_t *out_lo, int n_bf16, int64_t ld_bf16, void *stream) {
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
    TcCore core;
    core.block_n = pick_block_n(M, N, current_device_sm_count());
    const size_t smem = tc_finalize(core, K, A_lo ? 3 : 1, 1, 4 * TC_WARP_SCRATCH_WORDS * 4);
    TcMaps maps;
    int rc = fill_maps(maps, A_hi, A_lo, M, lda, W_hi, W_lo, N, ldw, K, core.block_n);
    if (rc != LOCO_OK) return rc;
    EpiLinear::Params p;
    p.bias = bias; p.out_f32 = out_f32; p.ld_f32 = ld_f32; p.out_hi = out_hi; p.out_lo = out_lo;
    p.n_bf16 = out_hi ? n_bf16 : 0; p.ld_bf16 = ld_bf16; p.M = M; p.N = N;
    p.tiles_n = (N + core.block_n - 1) / core.block_n;
    const int grid = ((M + TC_BLOCK_M - 1) / TC_BLOCK_M) * p.tiles_n;
    return tc_launch<EpiLinear>(maps, core, p, grid, smem, static_cast<cudaStream_t>(stream));
}

int loco_box_score_fwd(const uint16_t *E_hi, const uint16_t *E_lo, int64_t lde, const uint16_t *C_hi, const uint16_t *C_lo,
                       int64_t ldc, const float *cls_bias, int R, int K1, int D, float *logits, float *probs,
                       int64_t ld_logits, float *lse, int64_t *argmax_fg, void *stream) {
    LOCO_REQUIRE(R >= 0 && K1 >= 1 && D > 0, LOCO_E_BADARG, "box_score_fwd: bad shape R=%d K1=%d D=%d", R, K1, D);
    if (R == 0) return LOCO_OK;
    LOCO_REQUIRE(E_hi && C_hi && logits, LOCO_E_BADARG, "box_score_fwd: null pointer");
    LOCO_REQUIRE((E_lo == nullptr) == (C_lo == nullptr), LOCO_E_BADARG, "box_score_fwd: E_lo and C_lo must both be given or both be NULL");
    LOCO_REQUIRE(ld_logits >= K1, LOCO_E_BADARG, "box_score_fwd: ld_logits < K1");
    TcCore core;
    // one N tile when the class list fits (<= 256), otherwise 256-wide chunks with online softmax
    core.block_n = K1 <= 256 ? tc_round_up(K1, 32) : 256;
    const int chunks = (K1 + core.block_n - 1) / core.block_n;
    const size_t smem = tc_finalize(core, D, E_lo ? 3 : 1, chunks, 4 * TC_WARP_SCRATCH_WORDS * 4);
    TcMaps maps;
    int rc = fill_maps(maps, E_hi, E_lo, R, lde, C_hi, C_lo, K1, ldc, D, core.block_n);
    if (rc != LOCO_OK) return rc;
    EpiScore::Params p;
    p.bias = cls_bias; p.logits = logits; p.probs = probs; p.ld = ld_logits; p.lse = lse; p.argmax_fg = argmax_fg;
    p.R = R; p.K1 = K1;
    const int grid = (R + TC_BLOCK_M - 1) / TC_BLOCK_M;
    return tc_launch<EpiScore>(maps, core, p, grid, smem, static_cast<cudaStream_t>(stream));
}

int64_t loco_lsm_pair_workspace_bytes(int Bc, int T, int Bi, int Rg) {
    (void)Bc; (void)T; (void)Bi; (void)Rg;
    return 16;   // reserved (partial sums are reduced on-chip); kept so callers always pass a valid pointer
}

int loco_lsm_pair_fwd(const uint16_t *cap_hi, const uint16_t *cap_lo, int64_t ldcap, const float *cap_mask,
                      const uint16_t *emb_hi, const uint16_t *emb_lo, int64_t ldemb, const float *reg_mask, int Bc, int T,
                      int Bi, int Rg, int D, float inv_temperature, int alignment, float *d_w2r, float *d_r2w,
                      int64_t ld_out, void *workspace, void *stream) {
    (void)workspace;
    LOCO_REQUIRE(Bc >= 0 && Bi >= 0 && T > 0 && Rg > 0 && D > 0, LOCO_E_BADARG, "lsm_pair_fwd: bad shape Bc=%d T=%d Bi=%d Rg=%d D=%d", Bc, T, Bi, Rg, D);
    if (Bc == 0 || Bi == 0) return LOCO_OK;
    LOCO_REQUIRE(T <= 128 && Rg <= 256, LOCO_E_UNSUPPORTED, "lsm_pair_fwd: supports T <= 128 words and Rg <= 256 regions (got T=%d Rg=%d)", T, Rg);
    LOCO_REQUIRE(alignment == LOCO_ALIGN_SOFTMAX || alignment == LOCO_ALIGN_HARDMAX, LOCO_E_UNSUPPORTED, "lsm_pair_fwd: alignment %d not implemented", alignment);
    LOCO_REQUIRE(cap_hi && emb_hi && cap_mask && reg_mask, LOCO_E_BADARG, "lsm_pair_fwd: null pointer");
    LOCO_REQUIRE((cap_lo == nullptr) == (emb_lo == nullptr), LOCO_E_BADARG, "lsm_pair_fwd: cap_lo and emb_lo must both be given or both be NULL");
    LOCO_REQUIRE(d_w2r || d_r2w, LOCO_E_BADARG, "lsm_pair_fwd: no output requested");
    LOCO_REQUIRE(ld_out >= Bi, LOCO_E_BADARG, "lsm_pair_fwd: ld_out < Bi");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int passes = cap_lo ? 3 : 1;
    LsmParams p;
    p.cap_mask = cap_mask; p.reg_mask = reg_mask; p.ld = ld_out; p.Bc = Bc; p.T = T; p.Bi = Bi; p.Rg = Rg;
    p.inv_temp = inv_temperature; p.hardmax = (alignment == LOCO_ALIGN_HARDMAX);
    int rc;
    if (d_w2r) {
        TcCore core;
        core.block_n = tc_round_up(Rg, 16);
        p.per_tile = TC_BLOCK_M / T;
        p.groups = (Bc + p.per_tile - 1) / p.per_tile;
        p.row_blocks = 1;
        p.out = d_w2r;
        const size_t smem = tc_finalize(core, D, passes, 1, (128 + 128 + 256) * 4);
        TcMaps maps;
        rc = fill_maps(maps, cap_hi, cap_lo, (uint64_t)Bc * T, ldcap, emb_hi, emb_lo, (uint64_t)Bi * Rg, ldemb, D, core.block_n);
        if (rc != LOCO_OK) return rc;
        rc = tc_launch<EpiW2R>(maps, core, p, p.groups * Bi, smem, st);
        if (rc != LOCO_OK) return rc;
    }
    if (d_r2w) {
        TcCore core;
        p.per_tile = 256 / T < Bc ? 256 / T : Bc;
        core.block_n = tc_round_up(p.per_tile * T, 16);
        p.groups = (Bc + p.per_tile - 1) / p.per_tile;
        p.row_blocks = (Rg + TC_BLOCK_M - 1) / TC_BLOCK_M;
        p.out = d_r2w;
        if (p.row_blocks > 1)
            for (int c = 0; c < Bc; ++c)   // rows are ld_out apart: zero each row's Bi entries
                LOCO_CUDA(cudaMemsetAsync(d_r2w + (int64_t)c * ld_out, 0, sizeof(float) * Bi, st));
        const size_t smem = tc_finalize(core, D, passes, 1, (256 + 64) * 4);
        TcMaps maps;
        rc = fill_maps(maps, emb_hi, emb_lo, (uint64_t)Bi * Rg, ldemb, cap_hi, cap_lo, (uint64_t)Bc * T, ldcap, D, core.block_n);
        if (rc != LOCO_OK) return rc;
        rc = tc_launch<EpiR2W>(maps, core, p, Bi * p.row_blocks * p.groups, smem, st);
        if (rc != LOCO_OK) return rc;
    }
    return LOCO_OK;
}

}  // extern "C"
