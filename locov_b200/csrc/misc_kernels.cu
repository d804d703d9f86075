// misc_kernels.cu — bandwidth-bound helpers around the tensor-core kernels:
//   * loco_split_bf16   : fp32 -> bf16 hi (+lo) operand preparation (optionally transposed)
//   * loco_box_ce_fwd_bwd: cross-entropy loss + dlogits from the fused-softmax statistics
//   * loco_pair_ce      : empty-pair guard, 4 CE losses and 4 batch accuracies on the [Bc,Bi] pair matrix
#include <cfloat>

#include "common.cuh"

namespace loco {

// ---- fp32 -> bf16 hi/lo -------------------------------------------------------------------------------
// one destination quad (4 columns) of the fp32 -> bf16 hi/lo split
__device__ __forceinline__ void split_quad(const float *__restrict__ src, int64_t cols, int64_t src_ld, uint16_t *__restrict__ hi,
                                           uint16_t *__restrict__ lo, int64_t dst_ld, int64_t quads_per_row, bool vec_ok, int64_t idx) {
    const int64_t r = idx / quads_per_row;
    const int64_t c = (idx - r * quads_per_row) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const float *s = src + r * src_ld + c;
    if (c + 3 < cols && vec_ok) {
        const float4 f = __ldg(reinterpret_cast<const float4 *>(s));
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c + j < cols) v[j] = __ldg(s + j);
    }
    uint16_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_bf16(v[j], h[j], l[j]);
    uint2 ph, pl;
    ph.x = (uint32_t)h[0] | ((uint32_t)h[1] << 16); ph.y = (uint32_t)h[2] | ((uint32_t)h[3] << 16);
    pl.x = (uint32_t)l[0] | ((uint32_t)l[1] << 16); pl.y = (uint32_t)l[2] | ((uint32_t)l[3] << 16);
    *reinterpret_cast<uint2 *>(hi + r * dst_ld + c) = ph;
    if (lo != nullptr) *reinterpret_cast<uint2 *>(lo + r * dst_ld + c) = pl;
}

// one element of the LSM masks: caption_mask = attention * (1 - special) (grounding_head.py:94-101), region mask -> fp32 (:105-106)
__device__ __forceinline__ void mask_item(const int64_t *__restrict__ att, const int64_t *__restrict__ spe, int64_t n_cap,
                                          const void *__restrict__ reg, int reg_kind, float *__restrict__ cap_mask,
                                          float *__restrict__ reg_mask, int64_t i) {
    if (i < n_cap) {
        cap_mask[i] = (float)(att[i] * (1 - spe[i]));
    } else {
        const int64_t j = i - n_cap;
        float v;
        if (reg_kind == 0) v = (float)static_cast<const uint8_t *>(reg)[j];
        else if (reg_kind == 1) v = static_cast<const float *>(reg)[j];
        else v = (float)static_cast<const int64_t *>(reg)[j];
        reg_mask[j] = v;
    }
}

// One thread per 4 destination columns; pad columns [cols, dst_ld) are written as zeros.
__global__ void __launch_bounds__(256) split_bf16_kernel(const float *__restrict__ src, int64_t rows, int64_t cols,
                                                         int64_t src_ld, uint16_t *__restrict__ hi,
                                                         uint16_t *__restrict__ lo, int64_t dst_ld) {
    pdl_trigger();
    pdl_wait();
    const int64_t quads_per_row = dst_ld / 4;      // dst_ld % 8 == 0
    const int64_t total = rows * quads_per_row;
    const bool vec_ok = (src_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        split_quad(src, cols, src_ld, hi, lo, dst_ld, quads_per_row, vec_ok, idx);
    }
}

// ---- LSM input preparation in ONE launch: caption embeddings -> bf16 operand (hi / lo) and both masks -> fp32.
// (the two jobs are independent; separate launches cost more in launch gaps than in work at the BASELINE shapes)
__global__ void __launch_bounds__(256) lsm_prep_kernel(const float *__restrict__ cap, int64_t rows, int64_t cols, int64_t cap_ld,
                                                       uint16_t *__restrict__ hi, uint16_t *__restrict__ lo, int64_t dst_ld,
                                                       const int64_t *__restrict__ att, const int64_t *__restrict__ spe, int64_t n_cap,
                                                       const void *__restrict__ reg, int reg_kind, int64_t n_reg,
                                                       float *__restrict__ cap_mask, float *__restrict__ reg_mask) {
    pdl_trigger();
    pdl_wait();
    const int64_t quads_per_row = dst_ld / 4;
    const int64_t nq = rows * quads_per_row, total = nq + n_cap + n_reg;
    const bool vec_ok = (cap_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(cap) & 15) == 0);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        if (idx < nq) split_quad(cap, cols, cap_ld, hi, lo, dst_ld, quads_per_row, vec_ok, idx);
        else mask_item(att, spe, n_cap, reg, reg_kind, cap_mask, reg_mask, idx - nq);
    }
}

// ---- LSM input preparation, all of it in ONE launch: up to three fp32 -> bf16 (hi / lo) jobs (region features, projection weight,
// caption embeddings) and both masks.  The caption / weight / mask jobs are latency-bound on their own (3.5-3.8 us each for a few MB);
// inside the launch of the HBM-bound feature split (9 us for 26 MB) they cost nothing.  Jobs are laid out largest first.
struct SplitJobs {
    const float *src[3];
    uint16_t *hi[3], *lo[3];
    long long rows[3], cols[3], src_ld[3], dst_ld[3], nq[3];     // nq = rows * dst_ld / 4 destination quads
    int n;
};
__global__ void __launch_bounds__(256) lsm_prep_multi_kernel(const SplitJobs jobs, const int64_t *__restrict__ att, const int64_t *__restrict__ spe,
                                                             int64_t n_cap, const void *__restrict__ reg, int reg_kind, int64_t n_reg,
                                                             float *__restrict__ cap_mask, float *__restrict__ reg_mask) {
    pdl_trigger();
    pdl_wait();
    int64_t total = n_cap + n_reg;
    for (int j = 0; j < jobs.n; ++j) total += jobs.nq[j];
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = idx;
        bool done = false;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (done || j >= jobs.n) continue;
            if (k < jobs.nq[j]) {
                const bool vec_ok = (jobs.src_ld[j] % 4 == 0) && ((reinterpret_cast<uintptr_t>(jobs.src[j]) & 15) == 0);
                split_quad(jobs.src[j], jobs.cols[j], jobs.src_ld[j], jobs.hi[j], jobs.lo[j], jobs.dst_ld[j], jobs.dst_ld[j] / 4, vec_ok, k);
                done = true;
            } else {
                k -= jobs.nq[j];
            }
        }
        if (!done) mask_item(att, spe, n_cap, reg, reg_kind, cap_mask, reg_mask, k);
    }
}

// transposed variant: dst[c, r] = src[r, c]; dst is [cols, dst_ld] with pad [rows, dst_ld) zeroed.
__global__ void __launch_bounds__(256) split_bf16_t_kernel(const float *__restrict__ src, int rows, int cols, int64_t src_ld,
                                                           uint16_t *__restrict__ hi, uint16_t *__restrict__ lo,
                                                           int64_t dst_ld) {
    pdl_trigger();
    pdl_wait();
    __shared__ float tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        tile[ty + 8 * k][tx] = (r < rows && c < cols) ? __ldg(src + (int64_t)r * src_ld + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;     // dst row = c, dst col = r
        if (c < cols && r < dst_ld) {
            uint16_t h, l;
            split_bf16(tile[tx][ty + 8 * k], h, l);
            hi[(int64_t)c * dst_ld + r] = h;
            if (lo != nullptr) lo[(int64_t)c * dst_ld + r] = l;
        }
    }
}

// ---- LSM mask preparation: caption_mask = attention_mask * (1 - special_tokens_mask) as fp32 (grounding_head.py:94-101),
// region_mask (uint8 / fp32 / int64) as fp32 (:105-106) — one launch instead of four ATen elementwise kernels.
__global__ void __launch_bounds__(256) lsm_masks_kernel(const int64_t *__restrict__ att, const int64_t *__restrict__ spe, int64_t n_cap,
                                                        const void *__restrict__ reg, int reg_kind, int64_t n_reg,
                                                        float *__restrict__ cap_mask, float *__restrict__ reg_mask) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cap + n_reg; i += (int64_t)gridDim.x * blockDim.x)
        mask_item(att, spe, n_cap, reg, reg_kind, cap_mask, reg_mask, i);
}

// ---- 16-bit transpose: dst[c, r] = src[r, c] (bf16 operands for the backward GEMMs) ---------------------------
__global__ void __launch_bounds__(256) transpose16_kernel(const uint16_t *__restrict__ src, int rows, int cols, int64_t src_ld,
                                                          uint16_t *__restrict__ dst, int64_t dst_ld) {
    pdl_trigger();
    pdl_wait();
    __shared__ uint16_t tile[32][34];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        tile[ty + 8 * k][tx] = (r < rows && c < cols) ? src[(int64_t)r * src_ld + c] : (uint16_t)0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;     // dst row = c, dst col = r; pad columns [rows, dst_ld) get zeros
        if (c < cols && r < dst_ld) dst[(int64_t)c * dst_ld + r] = tile[tx][ty + 8 * k];
    }
}

// ---- cross entropy over scored logits -------------------------------------------------------------------
// one warp per RoI row
__global__ void __launch_bounds__(256) box_ce_kernel(const float *__restrict__ logits, int64_t ld, const float *__restrict__ lse,
                                                     const int64_t *__restrict__ labels, int R, int K1, float scale,
                                                     float *__restrict__ loss_sum, float grad_scale,
                                                     float *__restrict__ dl_f32, uint16_t *__restrict__ dl_bf16,
                                                     int64_t ld_bf16) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    float local = 0.f;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += gridDim.x * wpb) {
        const int64_t y = labels[r];
        const float l = lse[r];
        const float *row = logits + (int64_t)r * ld;
        if (lane == 0 && y >= 0 && y < K1) local += (l - row[y]);
        if (dl_f32 != nullptr || dl_bf16 != nullptr) {
            const int64_t lim = dl_bf16 ? ld_bf16 : K1;
            for (int64_t c = lane; c < lim; c += 32) {
                float g = 0.f;
                if (c < K1) g = (expf(row[c] - l) - (c == y ? 1.f : 0.f)) * grad_scale;
                if (dl_f32 != nullptr && c < K1) dl_f32[(int64_t)r * ld + c] = g;
                if (dl_bf16 != nullptr) dl_bf16[(int64_t)r * ld_bf16 + c] = f32_to_bf16_rn(g);
            }
        }
    }
    // block reduction of lane-0 partials -> one atomic per block
    __shared__ float red[32];
    local = warp_sum(local);
    if (lane == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < wpb ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0 && v != 0.f) atomicAdd(loss_sum, v * scale);
    }
}

// ---- row softmax statistics over given logits (no GEMM in front) ----------------------------------------------------
// One warp per RoI row: log-sum-exp over the K1 columns, first-maximum argmax over the K1 - 1 foreground columns and, when
// requested, the probabilities — the same outputs as the scoring epilogue, for logits that did not come out of it
// (NORMALIZE_EMB_PRED / STANDARDIZE_EMB_PRED, or a caller-modified score matrix handed to losses / inference).
__global__ void __launch_bounds__(256) box_softmax_kernel(const float *__restrict__ logits, int64_t ld, int R, int K1, float *__restrict__ lse,
                                                          int64_t *__restrict__ argmax_fg, float *__restrict__ probs, int64_t ld_probs) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += gridDim.x * wpb) {
        const float *row = logits + (int64_t)r * ld;
        float m = -FLT_MAX, best = -FLT_MAX;
        int arg = 0x7fffffff;
        for (int c = lane; c < K1; c += 32) {
            const float v = row[c];
            m = fmaxf(m, v);
            if (c < K1 - 1 && v > best) { best = v; arg = c; }       // columns visited in increasing order per lane
        }
        m = warp_max(m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        float s = 0.f;
        for (int c = lane; c < K1; c += 32) s += expf(row[c] - m);
        s = warp_sum(s);
        const float l = m + logf(s);
        if (lane == 0) {
            if (lse != nullptr) lse[r] = l;
            if (argmax_fg != nullptr) argmax_fg[r] = (int64_t)(arg == 0x7fffffff ? 0 : arg);
        }
        if (probs != nullptr)
            for (int c = lane; c < K1; c += 32) probs[(int64_t)r * ld_probs + c] = expf(row[c] - l);
    }
}

// ---- row-wise L2 normalisation / standardisation of the projected embeddings (logged_module.py:55-72) ---------------
// mode 0: y = x / max(||x||_2, 1e-12)        (F.normalize)
// mode 1: y = (x - mean) / (std + 1e-12), std with Bessel's correction (torch.std default)
// One warp per row.  Backward (dy -> dx) in the same kernel when `dy` is given: x is the forward INPUT.
__global__ void __launch_bounds__(256) row_normalize_kernel(const float *__restrict__ x, int64_t ldx, int rows, int cols, int mode,
                                                            const float *__restrict__ dy, int64_t lddy, float *__restrict__ out, int64_t ldo) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
        const float *xr = x + (int64_t)r * ldx;
        float *o = out + (int64_t)r * ldo;
        float s1 = 0.f;
        for (int c = lane; c < cols; c += 32) s1 += xr[c];
        s1 = warp_sum(s1);
        const float mean = mode == 1 ? s1 / (float)cols : 0.f;
        float s2 = 0.f;
        for (int c = lane; c < cols; c += 32) { const float d = xr[c] - mean; s2 = fmaf(d, d, s2); }
        s2 = warp_sum(s2);
        if (mode == 0) {
            const float nrm = sqrtf(s2), den = fmaxf(nrm, 1e-12f);
            if (dy == nullptr) {
                for (int c = lane; c < cols; c += 32) o[c] = xr[c] / den;
            } else {
                const float *g = dy + (int64_t)r * lddy;
                float dot = 0.f;
                for (int c = lane; c < cols; c += 32) dot = fmaf(g[c], xr[c], dot);
                dot = warp_sum(dot);
                // y = x / den; den = ||x|| unless clamped (then it is a constant)
                const float k = nrm > 1e-12f ? dot / (den * den * den) : 0.f;
                for (int c = lane; c < cols; c += 32) o[c] = g[c] / den - xr[c] * k;
            }
        } else {
            const float sd = cols > 1 ? sqrtf(s2 / (float)(cols - 1)) : nanf("");
            const float den = sd + 1e-12f;
            if (dy == nullptr) {
                for (int c = lane; c < cols; c += 32) o[c] = (xr[c] - mean) / den;
            } else {
                const float *g = dy + (int64_t)r * lddy;
                float gs = 0.f, gc = 0.f;
                for (int c = lane; c < cols; c += 32) { gs += g[c]; gc = fmaf(g[c], xr[c] - mean, gc); }
                gs = warp_sum(gs);
                gc = warp_sum(gc);
                // dL/dc = g/den - (sum g c) / den^2 * c / (sd (n-1));  dx = dL/dc - mean(dL/dc), and sum(c) = 0
                const float k = sd > 0.f ? gc / (den * den * sd * (float)(cols - 1)) : 0.f;
                const float gm = gs / ((float)cols * den);
                for (int c = lane; c < cols; c += 32) o[c] = g[c] / den - (xr[c] - mean) * k - gm;
            }
        }
    }
}

// ---- pair-matrix losses (single CTA; the matrix is at most a few hundred KB) ----------------------------------
// Both reductions return the block-wide result to EVERY thread (broadcast through red[0]).
__device__ __forceinline__ float block_reduce_max(float v, float *red) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        float r = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -FLT_MAX;
        r = warp_max(r);
        if (threadIdx.x == 0) red[0] = r;
    }
    __syncthreads();
    const float out = red[0];
    __syncthreads();
    return out;
}
__device__ __forceinline__ float block_reduce_sum(float v, float *red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        float r = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
        r = warp_sum(r);
        if (threadIdx.x == 0) red[0] = r;
    }
    __syncthreads();
    const float out = red[0];
    __syncthreads();
    return out;
}

// Pair-matrix losses.  grid = (nmat, nblk): CTA (m, j) owns 32 image columns (choose caption: a warp per column) and
// 32 caption rows (choose image: a warp per row) of matrix m; lanes stride over the other index and every reduction is a
// fixed-order shuffle tree.  nblk == 1 (B <= 32): the matrix is staged in shared memory once.  nblk > 1: each CTA scans the
// whole (L2-resident) matrix for the guard's global maximum, applies the guard on the fly, and the last CTA to finish
// (ticket counter in `scratch`) adds the per-CTA partial sums in CTA order — deterministic, no float atomics.
// monotone float <-> unsigned mapping (larger float = larger unsigned; 0 is below every finite float): atomicMax on floats
__device__ __forceinline__ unsigned int float_to_ordered(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// scratch layout of the multi-CTA pair-CE (floats): [16] tickets, [16] ordered maxima, [Bc] empty-caption flags, [Bi] empty-image flags,
// then [nmat][nblk][4] partial sums
__device__ __forceinline__ float *pce_flags(float *scratch) { return scratch + 32; }

// First launch of the multi-CTA pair-CE (B > 32): what EVERY strip CTA needs and none should recompute — the empty-caption / empty-image
// flags (a warp per row of the masks) and the maximum of each ORIGINAL matrix for the empty-pair guard (grid-stride scan, one atomicMax per
// block on the ordered encoding).  With these replicated in each of the 64 strip CTAs of a 256 x 256 matrix they were most of its 28 us.
__global__ void __launch_bounds__(256) pair_ce_prep_kernel(const float *__restrict__ pw_all, int nmat, int64_t mat_stride, int64_t ld, int Bc, int Bi,
                                                           const float *__restrict__ cap_mask, int T, const float *__restrict__ reg_mask, int Rg,
                                                           float *__restrict__ scratch) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32];
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    float *cap_empty = pce_flags(scratch), *img_empty = cap_empty + Bc;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < Bc + Bi; row += gridDim.x * wpb) {
        float sum = 0.f;
        if (row < Bc) {
            for (int t = lane; t < T; t += 32) sum += cap_mask[(int64_t)row * T + t];
        } else {
            for (int r = lane; r < Rg; r += 32) sum += reg_mask[(int64_t)(row - Bc) * Rg + r];
        }
        sum = warp_sum(sum);
        if (lane == 0) (row < Bc ? cap_empty[row] : img_empty[row - Bc]) = sum > 0.f ? 0.f : 1.f;
    }
    unsigned int *maxenc = reinterpret_cast<unsigned int *>(scratch) + 16;
    for (int m = 0; m < nmat; ++m) {
        const float *pw = pw_all + (int64_t)m * mat_stride;
        float mx = -FLT_MAX;
        for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < Bc; c += gridDim.x * wpb) {      // warp = row, lane = column: coalesced
            const float *rowp = pw + (int64_t)c * ld;
            for (int i = lane; i < Bi; i += 32) mx = fmaxf(mx, rowp[i]);
        }
        mx = block_reduce_max(mx, red);
        if (threadIdx.x == 0) atomicMax(maxenc + m, float_to_ordered(mx));
    }
}

// STRIP (B > 32 only): image columns / caption rows per CTA — 32, or 8 for the large matrices of the sharded head (256 x 256: 64 CTAs
// instead of 16; the kernel is a chain of short dependent passes, so its time is per-CTA latency, not throughput).
template <bool STAGED, int STRIP = 32>
__global__ void __launch_bounds__(1024) pair_ce_kernel(float *__restrict__ pw_all, int64_t mat_stride, int64_t ld, int Bc, int Bi,
                                                       int diag_off, const float *__restrict__ cap_mask, int T,
                                                       const float *__restrict__ reg_mask, int Rg, float *__restrict__ out4_all,
                                                       float *__restrict__ dcap_all, float *__restrict__ dimg_all,
                                                       float *__restrict__ scratch) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];
    float *red = sm;                 // [32]
    float *cap_empty = sm + 32;      // [Bc] 1 if the caption has no valid word
    float *img_empty = cap_empty + Bc;   // [Bi]
    float *acc = img_empty + Bi;     // [4][32] per-warp partial sums: ce_cap, ce_img, acc_cap, acc_img (= out4 order)
    float *tile = acc + 128;         // [Bc][Bi + 1] when STAGED
    __shared__ int s_last;
    const int mat_id = blockIdx.x, blk = blockIdx.y, nblk = gridDim.y;
    float *pw = pw_all + (int64_t)mat_id * mat_stride;
    float *out4 = out4_all + 4 * mat_id;
    float *dcap = dcap_all ? dcap_all + (int64_t)mat_id * Bc * Bi : nullptr;
    float *dimg = dimg_all ? dimg_all + (int64_t)mat_id * Bc * Bi : nullptr;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int64_t mld = STAGED ? (Bi + 1) : ld;      // row stride of the matrix the passes read
    const float *mat = STAGED ? tile : pw;

    float mx = -FLT_MAX;
    if (STAGED) {
        for (int idx = tid; idx < Bc * Bi; idx += nt) {
            const int c = idx / Bi, i = idx - c * Bi;
            const float v = pw[(int64_t)c * ld + i];
            tile[c * (Bi + 1) + i] = v;
            mx = fmaxf(mx, v);
        }
    } else {
        // flags and the matrix maximum come from pair_ce_prep_kernel (previous launch)
        const float *flags = pce_flags(scratch);
        for (int k = tid; k < Bc + Bi; k += nt) cap_empty[k] = flags[k];             // (img_empty follows cap_empty in both places)
        mx = ordered_to_float(reinterpret_cast<const unsigned int *>(scratch)[16 + mat_id]);
    }
    if (STAGED) {
        for (int c = warp; c < Bc; c += nwarp) {
            float s = 0.f;
            for (int t = lane; t < T; t += 32) s += cap_mask[(int64_t)c * T + t];
            s = warp_sum(s);
            if (lane == 0) cap_empty[c] = s > 0.f ? 0.f : 1.f;
        }
        for (int i = warp; i < Bi; i += nwarp) {
            float s = 0.f;
            for (int r = lane; r < Rg; r += 32) s += reg_mask[(int64_t)i * Rg + r];
            s = warp_sum(s);
            if (lane == 0) img_empty[i] = s > 0.f ? 0.f : 1.f;
        }
    }
    if (tid < 128) acc[tid] = 0.f;
    // empty-pair guard (grounding_head.py:240-251): max over the whole ORIGINAL matrix; with several CTAs the entries
    // are rewritten by the last CTA only, after every CTA has finished scanning
    mx = block_reduce_max(mx, red);            // (contains the __syncthreads that publish tile / cap_empty / img_empty)
    const float guard = mx + 100.0f;
    // value of entry (c, i) with the guard applied on the fly
    auto val = [&](int c, int i) -> float {
        const float v = mat[(int64_t)c * mld + i];
        if (STAGED) return v;
        return (cap_empty[c] > 0.f && img_empty[i] > 0.f) ? guard : v;
    };
    if (STAGED) {
        for (int idx = tid; idx < Bc * Bi; idx += nt) {
            const int c = idx / Bi, i = idx - c * Bi;
            if (cap_empty[c] > 0.f && img_empty[i] > 0.f) {
                pw[(int64_t)c * ld + i] = guard;
                tile[c * (Bi + 1) + i] = guard;
            }
        }
        __syncthreads();
    }
    const float inv = 1.0f / (float)Bi;
    const int i_beg = blk * STRIP, i_end = min(Bi, i_beg + STRIP);
    const int c_beg = blk * STRIP, c_end = min(Bc, c_beg + STRIP);
    // with several CTAs the last block also takes the rows beyond nblk * STRIP (Bc > Bi never happens for square
    // matrices; kept general)
    const int c_end2 = (blk == nblk - 1) ? Bc : c_end;
    const int i_end2 = (blk == nblk - 1) ? Bi : i_end;

    // choose caption: per image column i, log-softmax over the rows of -pw; target row = i + diag_off
    if (!STAGED) {
        // matrix in global memory (B > 32): lane = column of this CTA's strip, warps stride over the rows, so that every
        // access is a coalesced 128-byte row segment (a warp walking one column touches 32 sectors per load: 20 of the
        // 45 us of this kernel at B = 256).  Per-warp partial statistics are combined over the warps in a fixed order.
        float *px = tile;                  // [nwarp][33] partial max / partial sums
        float *py = px + 32 * 33;          // [nwarp][33] partial minima
        int *pz = reinterpret_cast<int *>(py + 32 * 33);   // [nwarp][33] partial argmin
        float *col_m = reinterpret_cast<float *>(pz + 32 * 33), *col_s = col_m + 32;
        // lane = (row sub-lane, column of the strip): a warp covers SUB consecutive rows x STRIP columns per load
        constexpr int SUB = 32 / STRIP;
        const int colj = lane % STRIP, sub = lane / STRIP;
        const int i = i_beg + colj;
        const bool col_on = i < i_end2;
        const int tgt = i + diag_off;
        float m = -FLT_MAX, best = FLT_MAX;
        int arg = 0x7fffffff;
        if (col_on)
            for (int c = warp * SUB + sub; c < Bc; c += nwarp * SUB) {
                const float v = val(c, i);
                m = fmaxf(m, -v);
                if (v < best) { best = v; arg = c; }        // rows visited in increasing order: first minimum
            }
        px[warp * 33 + lane] = m; py[warp * 33 + lane] = best; pz[warp * 33 + lane] = arg;
        __syncthreads();
        if (warp == 0 && lane < STRIP) {
            float mm = -FLT_MAX, bb = FLT_MAX;
            int aa = 0x7fffffff;
            for (int w = 0; w < nwarp; ++w)
#pragma unroll
                for (int sl = 0; sl < SUB; ++sl) {
                    mm = fmaxf(mm, px[w * 33 + sl * STRIP + lane]);
                    const float ob = py[w * 33 + sl * STRIP + lane];
                    const int oa = pz[w * 33 + sl * STRIP + lane];
                    if (ob < bb || (ob == bb && oa < aa)) { bb = ob; aa = oa; }
                }
            col_m[lane] = mm;
            reinterpret_cast<int *>(col_s)[32 + lane] = aa;                 // column argmin
        }
        __syncthreads();
        const float cm = col_m[colj];
        float s = 0.f;
        if (col_on)
            for (int c = warp * SUB + sub; c < Bc; c += nwarp * SUB) s += expf(-val(c, i) - cm);
        px[warp * 33 + lane] = s;          // (the partial maxima in px were consumed before the barrier above)
        __syncthreads();
        if (warp == 0) {
            float l = 0.f, a = 0.f;
            if (lane < STRIP) {
                float ss = 0.f;
                for (int w = 0; w < nwarp; ++w)
#pragma unroll
                    for (int sl = 0; sl < SUB; ++sl) ss += px[w * 33 + sl * STRIP + lane];
                col_s[lane] = ss;
                const int i0 = i_beg + lane, tg = i0 + diag_off;
                if (i0 < i_end2 && tg < Bc) {
                    l = (col_m[lane] + logf(ss)) + val(tg, i0);
                    a = (reinterpret_cast<int *>(col_s)[32 + lane] == tg) ? 1.f : 0.f;
                }
            }
            l = warp_sum(l);
            a = warp_sum(a);
            if (lane == 0) { acc[0 * 32 + 0] += l; acc[2 * 32 + 0] += a; }
        }
        __syncthreads();
        if (dcap != nullptr && col_on) {
            const float cs = col_s[colj];
            for (int c = warp * SUB + sub; c < Bc; c += nwarp * SUB) {
                float g = 0.f;
                if (tgt < Bc && !(cap_empty[c] > 0.f && img_empty[i] > 0.f))
                    g = (((c == tgt) ? 1.f : 0.f) - expf(-val(c, i) - cm) / cs) * inv;
                dcap[(int64_t)c * Bi + i] = g;
            }
        }
    } else
    for (int i = i_beg + warp; i < i_end2; i += nwarp) {
        float m = -FLT_MAX, best = FLT_MAX;
        int arg = 0x7fffffff;
        for (int c = lane; c < Bc; c += 32) {
            const float v = val(c, i);
            m = fmaxf(m, -v);
            if (v < best) { best = v; arg = c; }
        }
        m = warp_max(m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {     // argmin with first-index tie-break (torch.argmin)
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        float s = 0.f;
        for (int c = lane; c < Bc; c += 32) s += expf(-val(c, i) - m);
        s = warp_sum(s);
        const int tgt = i + diag_off;
        if (lane == 0 && tgt < Bc) {
            acc[0 * 32 + warp] += (m + logf(s)) + val(tgt, i);
            acc[2 * 32 + warp] += (arg == tgt) ? 1.f : 0.f;
        }
        if (dcap != nullptr) {
            // d/dpw[c,i] of mean_i( lse_c(-pw[:,i]) + pw[tgt,i] ); guard-filled entries are constants
            for (int c = lane; c < Bc; c += 32) {
                float g = 0.f;
                if (tgt < Bc && !(cap_empty[c] > 0.f && img_empty[i] > 0.f))
                    g = (((c == tgt) ? 1.f : 0.f) - expf(-val(c, i) - m) / s) * inv;
                dcap[(int64_t)c * Bi + i] = g;
            }
        }
    }
    // choose image: per caption row c, log-softmax over the columns of -pw; rows outside
    // [diag_off, diag_off + Bi) carry no choose-image loss
    for (int c = c_beg + warp; c < c_end2; c += nwarp) {
        const int k = c - diag_off;
        const bool has = (k >= 0 && k < Bi);
        float m = -FLT_MAX, best = FLT_MAX, s = 0.f;
        int arg = 0x7fffffff;
        if (has) {
            for (int i = lane; i < Bi; i += 32) {
                const float v = val(c, i);
                m = fmaxf(m, -v);
                if (v < best) { best = v; arg = i; }
            }
            m = warp_max(m);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
            }
            for (int i = lane; i < Bi; i += 32) s += expf(-val(c, i) - m);
            s = warp_sum(s);
            if (lane == 0) {
                acc[1 * 32 + warp] += (m + logf(s)) + val(c, k);
                acc[3 * 32 + warp] += (arg == k) ? 1.f : 0.f;
            }
        }
        if (dimg != nullptr) {
            for (int i = lane; i < Bi; i += 32) {
                float g = 0.f;
                if (has && !(cap_empty[c] > 0.f && img_empty[i] > 0.f))
                    g = (((i == k) ? 1.f : 0.f) - expf(-val(c, i) - m) / s) * inv;
                dimg[(int64_t)c * Bi + i] = g;
            }
        }
    }
    __syncthreads();
    float part = 0.f;
    if (warp < 4) {                  // warp w sums partial w over the (fixed-order) per-warp slots
        float v = (lane < nwarp) ? acc[warp * 32 + lane] : 0.f;
        part = warp_sum(v);
    }
    if (nblk == 1) {
        if (warp < 4 && lane == 0) out4[warp] = part * inv;
        return;
    }
    // several CTAs: publish the partials, the last CTA adds them in CTA order
    float *parts = scratch + 32 + Bc + Bi + ((int64_t)mat_id * nblk) * 4;
    unsigned int *ticket = reinterpret_cast<unsigned int *>(scratch) + mat_id;
    if (warp < 4 && lane == 0) parts[blk * 4 + warp] = part;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == (unsigned)(nblk - 1)) ? 1 : 0;
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (tid < 4) {
            float v = 0.f;
            for (int j = 0; j < nblk; ++j) v += __ldcg(parts + j * 4 + tid);
            out4[tid] = v * inv;
        }
        if (tid == 0) {              // ready for the next launch
            *ticket = 0;
            reinterpret_cast<unsigned int *>(scratch)[16 + mat_id] = 0u;
        }
        for (int idx = tid; idx < Bc * Bi; idx += nt) {
            const int c = idx / Bi, i = idx - c * Bi;
            if (cap_empty[c] > 0.f && img_empty[i] > 0.f) pw[(int64_t)c * ld + i] = guard;
        }
    }
}

// ---- class-agnostic box regression loss (Detectron2 FastRCNNOutputLayers.box_reg_loss, smooth-L1 / L1) -----------------------
// loss = sum over foreground rows (0 <= gt < K) of smooth_l1(deltas - get_deltas(proposal, gt_box)) * scale, and its gradient
// with respect to the predicted deltas (zero rows for background), in ONE launch instead of ~30 ATen launches (masks, box
// transforms, abs / where / sum and their autograd twins): with the GEMMs at tens of microseconds those launches WERE the
// training step of the box head.  One thread per RoI; per-block partial sums in a fixed order, last block adds the partials
// in block order (ticket in `scratch`): deterministic.
__global__ void __launch_bounds__(256) box_reg_loss_kernel(const float *__restrict__ deltas, int64_t ld_d, const float *__restrict__ prop,
                                                           const float *__restrict__ gtb, const int64_t *__restrict__ labels, int R, int K,
                                                           float wx, float wy, float ww, float wh, float beta, float scale,
                                                           float *__restrict__ loss, float *__restrict__ ddeltas, int64_t ld_g,
                                                           float *__restrict__ scratch) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32];
    __shared__ int s_last;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (r < R) {
        const int64_t y = labels[r];
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        if (y >= 0 && y < K) {
            const float4 p = *reinterpret_cast<const float4 *>(prop + 4 * (int64_t)r), q = *reinterpret_cast<const float4 *>(gtb + 4 * (int64_t)r);
            const float sw = p.z - p.x, sh = p.w - p.y, sx = p.x + 0.5f * sw, sy = p.y + 0.5f * sh;
            const float tw = q.z - q.x, th = q.w - q.y, tx = q.x + 0.5f * tw, ty = q.y + 0.5f * th;
            const float tgt[4] = {wx * (tx - sx) / sw, wy * (ty - sy) / sh, ww * logf(tw / sw), wh * logf(th / sh)};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float df = deltas[(int64_t)r * ld_d + k] - tgt[k], n = fabsf(df);
                if (beta < 1e-5f) {
                    l += n;
                    g[k] = (df > 0.f) ? scale : ((df < 0.f) ? -scale : 0.f);         // sign(0) = 0, as torch.abs' gradient
                } else if (n < beta) {
                    l += 0.5f * n * n / beta;
                    g[k] = df / beta * scale;
                } else {
                    l += n - 0.5f * beta;
                    g[k] = (df > 0.f) ? scale : -scale;
                }
            }
        }
        if (ddeltas != nullptr) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ddeltas[(int64_t)r * ld_g + k] = g[k];
        }
    }
    l = block_reduce_sum(l, red);
    float *parts = scratch + 16;
    unsigned int *ticket = reinterpret_cast<unsigned int *>(scratch);
    if (threadIdx.x == 0) parts[blockIdx.x] = l;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
        __threadfence();
        float v = 0.f;
        for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) v += __ldcg(parts + j);     // (fixed order per thread, then the fixed tree)
        v = block_reduce_sum(v, red);
        if (threadIdx.x == 0) {
            *loss = v * scale;
            *ticket = 0;            // ready for the next launch
        }
    }
}

// ---- skinny weight gradient: dW[j, v] = sum_r dy[r, j] * x[r, v], db[j] = sum_r dy[r, j]  (J <= 8 output rows) ---------------------
// The gradient of the class-agnostic box regressor (bbox_pred: 4 x 2048) is a [4 x R] x [R x 2048] product: as a tensor-core GEMM it
// needs x TRANSPOSED (a 67 MB pass at 8192 RoIs) and wastes 124 of 128 tile rows.  Here x is streamed once (HBM bound, 128-bit
// loads): block (vc, rc) owns 512 columns x a chunk of rows; partials [row chunks][J][V] are added in chunk order by the second kernel.
template <int J>
__global__ void __launch_bounds__(128) skinny_grad_partial_kernel(const float *__restrict__ dy, int64_t ld_dy, const float *__restrict__ x,
                                                                  int64_t ld_x, int R, int V, int rows_per_chunk, float *__restrict__ part) {
    pdl_trigger();
    pdl_wait();
    const int v0 = (blockIdx.x * 128 + threadIdx.x) * 4;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(R, r0 + rows_per_chunk);
    float4 acc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v0 < V) {
        auto load_x = [&](int r) -> float4 {
            if (v0 + 3 < V) return __ldcs(reinterpret_cast<const float4 *>(x + (int64_t)r * ld_x + v0));
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
            const float *xr = x + (int64_t)r * ld_x + v0;
            xv.x = xr[0];
            if (v0 + 1 < V) xv.y = xr[1];
            if (v0 + 2 < V) xv.z = xr[2];
            return xv;
        };
        auto accumulate = [&](int r, const float4 &xv) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float g = __ldg(dy + (int64_t)r * ld_dy + j);           // warp-uniform address: one broadcast load
                acc[j].x = fmaf(g, xv.x, acc[j].x); acc[j].y = fmaf(g, xv.y, acc[j].y);
                acc[j].z = fmaf(g, xv.z, acc[j].z); acc[j].w = fmaf(g, xv.w, acc[j].w);
            }
        };
        int r = r0;
        for (; r + 8 <= r1; r += 8) {                      // 8 independent 128-bit loads in flight per thread, rows added in order
            float4 xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[u] = load_x(r + u);
#pragma unroll
            for (int u = 0; u < 8; ++u) accumulate(r + u, xv[u]);
        }
        for (; r < r1; ++r) accumulate(r, load_x(r));
#pragma unroll
        for (int j = 0; j < J; ++j) {
            float *dst = part + ((int64_t)blockIdx.y * J + j) * V + v0;
            dst[0] = acc[j].x;
            if (v0 + 1 < V) dst[1] = acc[j].y;
            if (v0 + 2 < V) dst[2] = acc[j].z;
            if (v0 + 3 < V) dst[3] = acc[j].w;
        }
    }
}
// partials [nchunk][J*V] -> dw: thread = (element, quarter of the chunks); the four quarter sums are combined in a fixed order.
// Block 0 also forms the bias gradient (column sums of dy): warp = (column j, quarter of the rows), same fixed-order combination.
__global__ void __launch_bounds__(256) skinny_grad_reduce_kernel(const float *__restrict__ part, int nchunk, int JV, float *__restrict__ dw,
                                                                 const float *__restrict__ dy, int64_t ld_dy, int R, int J, float *__restrict__ db) {
    __shared__ float sm[4][64];
    __shared__ float sb[32];
    pdl_trigger();
    pdl_wait();
    const int e = threadIdx.x & 63, q = threadIdx.x >> 6;
    const int idx = blockIdx.x * 64 + e;
    float v = 0.f;
    if (idx < JV) {
        int c = q;
        for (; c + 12 < nchunk; c += 16) {
            const float a0 = part[(int64_t)c * JV + idx], a1 = part[(int64_t)(c + 4) * JV + idx], a2 = part[(int64_t)(c + 8) * JV + idx],
                        a3 = part[(int64_t)(c + 12) * JV + idx];
            v += a0; v += a1; v += a2; v += a3;
        }
        for (; c < nchunk; c += 4) v += part[(int64_t)c * JV + idx];
    }
    sm[q][e] = v;
    __syncthreads();
    if (q == 0 && idx < JV) dw[idx] = (sm[0][e] + sm[1][e]) + (sm[2][e] + sm[3][e]);
    if (db != nullptr && blockIdx.x == 0) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;           // 8 warps
        for (int j0 = 0; j0 < J; j0 += 2) {                                // two columns per round, four row quarters each
            const int j = j0 + (w >> 2), rq = w & 3;
            float t = 0.f;
            if (j < J)
                for (int r = rq * 32 + lane; r < R; r += 128) t += dy[(int64_t)r * ld_dy + j];
            t = warp_sum(t);
            if (lane == 0) sb[w] = t;
            __syncthreads();
            if (threadIdx.x < 2 && j0 + (int)threadIdx.x < J) {
                const int b = threadIdx.x * 4;
                db[j0 + threadIdx.x] = (sb[b] + sb[b + 1]) + (sb[b + 2] + sb[b + 3]);
            }
            __syncthreads();
        }
    }
}

// ---- peer scatter: one local 2-D buffer -> the same place in EVERY rank's symmetric buffer (all-gather by peer stores) --------------------
// src [rows][row_bytes] (pitch src_pitch) is written to dst_p + dst_offset (pitch dst_pitch) for every peer p of `peers` (device array of
// `n_peers` base pointers into the ranks' symmetric allocations, this rank's own included): 16-byte stores over NVLink / NVSwitch
// straight from the producing rank — no NCCL launch, no staging copy.  Ordering against the consumers is the caller's job (a
// symmetric-memory barrier after this kernel).  One block per (chunk of 16-byte words, peer).
struct PeerSegs {
    const unsigned char *src[4];
    long long src_pitch[4], dst_pitch[4], dst_offset[4];
    int rows[4], row_words[4];
    int nseg;
};

__device__ __forceinline__ void st_release_sys_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// wait for the flag words [channel][peer] of THIS rank (set by the peers' signals), then clear them for the next round
__device__ __forceinline__ void peer_wait_flags(unsigned int *const *flag_peers, int n_peers, int rank, int channel) {
    if ((int)threadIdx.x < n_peers) {
        unsigned int *mine = flag_peers[rank] + channel * n_peers + threadIdx.x;
        const uint64_t t0 = global_timer_ns();
        while (*reinterpret_cast<volatile unsigned int *>(mine) == 0u)      // (relaxed polling; the system-scope fence below is the acquire)
            if (global_timer_ns() - t0 > 2000000000ull) { printf("locov_b200: rank %d timed out waiting for peer %d (channel %d)\n", rank, (int)threadIdx.x, channel); __trap(); }
        *mine = 0u;
    }
    __threadfence_system();
}

// Up to four local 2-D buffers -> the same places of every rank's symmetric buffer, then (mode & 2) a SIGNAL to every peer and (mode & 4)
// a WAIT for every peer's signal in the same launch: every block fences its peer stores at system scope and takes a ticket; the last
// block stores 1 (release, system scope) into flag word [channel][this rank] of every peer and, when asked to, spins (acquire) on its own
// flag words [channel][peer] and clears them.  A flag word has one writer (its peer) and is cleared by its reader before that peer can
// signal the same channel again (the two exchanges of a step alternate channels and each needs the other side's previous signal), so
// plain stores suffice and CUDA-graph replays need no sequence numbers.  Waits are bounded (2 s, then trap): a missing peer surfaces
// as a CUDA error, never as a hung GPU.
template <typename W>
__global__ void __launch_bounds__(256) peer_scatter_kernel(const PeerSegs segs, unsigned char *const *__restrict__ peers, int n_peers,
                                                           unsigned int *const *__restrict__ flag_peers, int rank, int channel, int mode,
                                                           unsigned int *__restrict__ ticket) {
    pdl_trigger();
    pdl_wait();
    if (segs.nseg > 0) {
        const int peer = blockIdx.y, sg = blockIdx.z;
        unsigned char *dst = peers[peer] + segs.dst_offset[sg];
        const unsigned char *src = segs.src[sg];
        const int row_words = segs.row_words[sg];
        const long long sp = segs.src_pitch[sg], dp = segs.dst_pitch[sg];
        const int64_t total = (int64_t)segs.rows[sg] * row_words;
        if (dst + 0 != src || sp != dp)       // (a slice produced in place in this rank's own buffer needs no copy to itself)
            for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
                const int64_t r = idx / row_words, w = idx - r * row_words;
                const W v = *reinterpret_cast<const W *>(src + r * sp + w * (int64_t)sizeof(W));
                *reinterpret_cast<W *>(dst + r * dp + w * (int64_t)sizeof(W)) = v;
            }
    }
    if ((mode & 6) == 0) return;
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y * gridDim.z - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x == 0) *ticket = 0;                       // ready for the next launch
    if ((mode & 2) && (int)threadIdx.x < n_peers) st_release_sys_u32(flag_peers[threadIdx.x] + channel * n_peers + rank, 1u);
    if (mode & 4) peer_wait_flags(flag_peers, n_peers, rank, channel);
}

// ---- distillation losses on the pair matrices (distill_mmss_gcnn.py:211-289 KD, :381-433 MSE) -------------------------------------------
// Three [B, B] matrices: a teacher (`trans`) and the two student matrices of the LSM head.  KD: for both softmax directions (dim 0 =
// "choose caption", dim 1 = "choose image") and both students, KL(target || input) * temperature^2, batch-mean; the target is the
// teacher (kind 0: DISTILLATION_TEACHER_TRANSFORMER) or the student (kind 1: the shipped coco_lsm.yaml).  MSE (kind 2): every pair
// compared as is and transposed (the same value twice).
// Launch 1 (lines): a warp per (direction, line) computes the three log-sum-exps and the two KL sums of that line — a line of one
// direction only needs the same line of the other matrices — and its loss share; the last block adds the shares in line order.
// Launch 2 (gradients, optional): a thread per element combines the statistics of its column and of its row.
// workspace (floats): [16] ticket, [2][B][5] line statistics {lse_t, lse_w, lse_r, kl_w, kl_r}, [2][B] loss shares.
__global__ void __launch_bounds__(256) pair_distill_lines_kernel(const float *__restrict__ tm, int64_t ld_t, const float *__restrict__ wm, int64_t ld_w,
                                                                 const float *__restrict__ rm, int64_t ld_r, int B, float inv_temp, int kind,
                                                                 float loss_scale, float *__restrict__ loss, float *__restrict__ ws) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    float *stats = ws + 16, *share = stats + 10 * (int64_t)B;
    for (int line = blockIdx.x * wpb + (threadIdx.x >> 5); line < 2 * B; line += gridDim.x * wpb) {
        const int dir = line / B, idx = line - dir * B;
        // element k of the line: dir 0 (softmax over dim 0) walks down column idx, dir 1 along row idx
        const int64_t st_t = dir == 0 ? ld_t : 1, st_w = dir == 0 ? ld_w : 1, st_r = dir == 0 ? ld_r : 1;
        const float *pt = tm + (dir == 0 ? idx : idx * ld_t), *pw = wm + (dir == 0 ? idx : idx * ld_w), *pr = rm + (dir == 0 ? idx : idx * ld_r);
        float out = 0.f;
        if (kind == 2) {
            if (dir == 1)                                   // (every element once: the row lines)
                for (int k = lane; k < B; k += 32) {
                    const float t = pt[k * st_t], dw = t - pw[k * st_w], dr = t - pr[k * st_r];
                    out += dw * dw + dr * dr;
                }
            out = warp_sum(out);
        } else {
            float mt = -FLT_MAX, mw = -FLT_MAX, mr = -FLT_MAX;
            for (int k = lane; k < B; k += 32) {
                mt = fmaxf(mt, -pt[k * st_t] * inv_temp); mw = fmaxf(mw, -pw[k * st_w] * inv_temp); mr = fmaxf(mr, -pr[k * st_r] * inv_temp);
            }
            mt = warp_max(mt); mw = warp_max(mw); mr = warp_max(mr);
            float s_t = 0.f, s_w = 0.f, s_r = 0.f;
            for (int k = lane; k < B; k += 32) {
                s_t += expf(-pt[k * st_t] * inv_temp - mt); s_w += expf(-pw[k * st_w] * inv_temp - mw); s_r += expf(-pr[k * st_r] * inv_temp - mr);
            }
            const float lse_t = mt + logf(warp_sum(s_t)), lse_w = mw + logf(warp_sum(s_w)), lse_r = mr + logf(warp_sum(s_r));
            float kw = 0.f, kr = 0.f;
            for (int k = lane; k < B; k += 32) {
                const float lp = -pt[k * st_t] * inv_temp - lse_t, lw = -pw[k * st_w] * inv_temp - lse_w, lr = -pr[k * st_r] * inv_temp - lse_r;
                if (kind == 0) { const float p = expf(lp); kw += p * (lp - lw); kr += p * (lp - lr); }
                else { kw += expf(lw) * (lw - lp); kr += expf(lr) * (lr - lp); }
            }
            kw = warp_sum(kw); kr = warp_sum(kr);
            if (lane == 0) {
                float *q = stats + 5 * (int64_t)line;
                q[0] = lse_t; q[1] = lse_w; q[2] = lse_r; q[3] = kw; q[4] = kr;
            }
            out = kw + kr;
        }
        if (lane == 0) share[line] = out;
    }
    unsigned int *ticket = reinterpret_cast<unsigned int *>(ws);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
        __threadfence();
        float v = 0.f;
        for (int j = threadIdx.x; j < 2 * B; j += blockDim.x) v += __ldcg(share + j);         // fixed order per thread, then the fixed tree
        v = block_reduce_sum(v, red);
        if (threadIdx.x == 0) {
            *loss = v * loss_scale;
            *ticket = 0;
        }
    }
}

__global__ void __launch_bounds__(256) pair_distill_grad_kernel(const float *__restrict__ tm, int64_t ld_t, const float *__restrict__ wm, int64_t ld_w,
                                                                const float *__restrict__ rm, int64_t ld_r, int B, float inv_temp, int kind, float gscale,
                                                                float *__restrict__ g_t, float *__restrict__ g_w, float *__restrict__ g_r,
                                                                const float *__restrict__ ws) {
    pdl_trigger();
    pdl_wait();
    const float *stats = ws + 16;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)B * B; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e / B), i = (int)(e - (int64_t)c * B);
        const float t = tm[(int64_t)c * ld_t + i], w = wm[(int64_t)c * ld_w + i], r = rm[(int64_t)c * ld_r + i];
        float gt = 0.f, gw = 0.f, gr = 0.f;
        if (kind == 2) {
            gw = (w - t) * gscale; gr = (r - t) * gscale; gt = -(gw + gr);
        } else {
            const float at = -t * inv_temp, aw = -w * inv_temp, ar = -r * inv_temp;
#pragma unroll
            for (int dir = 0; dir < 2; ++dir) {
                const float *q = stats + 5 * (int64_t)(dir * B + (dir == 0 ? i : c));
                const float lp = at - q[0], lw = aw - q[1], lr = ar - q[2];
                const float p = expf(lp), qw = expf(lw), qr = expf(lr);
                if (kind == 0) { gw += qw - p; gr += qr - p; gt += p * ((lp - lw) - q[3]) + p * ((lp - lr) - q[4]); }
                else { gt += (p - qw) + (p - qr); gw += qw * ((lw - lp) - q[3]); gr += qr * ((lr - lp) - q[4]); }
            }
            gt *= gscale; gw *= gscale; gr *= gscale;
        }
        if (g_t != nullptr) g_t[e] = gt;
        if (g_w != nullptr) g_w[e] = gw;
        if (g_r != nullptr) g_r[e] = gr;
    }
}

// ---- multi-token class scoring: GroundingModule of box_emb_grounding_head.py:60-237 -------------------------------------
// Every class k owns the token columns [seg[k], seg[k+1]) of the token-logit matrix raw = e . tokens^T (one GEMM on the tensor-core
// core).  score[r, k] = sum_t a_t * s_t with s = raw / temperature and a = softmax over the class's tokens (hardmax: the first
// maximum).  Thread = (RoI, class); a class has a handful of tokens, neighbouring threads read neighbouring segments.
// The reference pads every class to max_tok and fills the padding with min(s) - 100 before the softmax: weight exp(-100 - ...) = 0.
template <bool BWD>
__global__ void __launch_bounds__(256) token_pool_kernel(const float *__restrict__ raw, int64_t ld_raw, const int32_t *__restrict__ seg, int R, int K1,
                                                         float inv_temp, int hardmax, float *__restrict__ scores, int64_t ld_s,
                                                         float *__restrict__ att, const float *__restrict__ dscores, int64_t ld_ds,
                                                         float *__restrict__ draw, int64_t ld_draw) {
    pdl_trigger();
    pdl_wait();
    const int64_t total = (int64_t)R * K1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / K1), k = (int)(i - (int64_t)r * K1);
        const int t0 = seg[k], t1 = seg[k + 1];
        const float *p = raw + (int64_t)r * ld_raw;
        float m = -FLT_MAX;
        int am = t0;
        for (int t = t0; t < t1; ++t) {
            const float v = p[t] * inv_temp;
            if (v > m) { m = v; am = t; }
        }
        float sum = 0.f, acc = 0.f;
        if (!hardmax)
            for (int t = t0; t < t1; ++t) {
                const float v = p[t] * inv_temp, e = expf(v - m);
                sum += e;
                acc = fmaf(e, v, acc);
            }
        const float score = t1 > t0 ? (hardmax ? m : acc / sum) : 0.f;
        if (!BWD) {
            scores[(int64_t)r * ld_s + k] = score;
            if (att != nullptr)
                for (int t = t0; t < t1; ++t) att[(int64_t)r * ld_raw + t] = hardmax ? (t == am ? 1.f : 0.f) : expf(p[t] * inv_temp - m) / sum;
        } else {
            // d score / d raw_t = a_t (1 + s_t - score) / temperature  (softmax);  1 / temperature at the maximum (hardmax)
            const float g = dscores[(int64_t)r * ld_ds + k] * inv_temp;
            for (int t = t0; t < t1; ++t) {
                float d;
                if (hardmax) d = t == am ? g : 0.f;
                else {
                    const float v = p[t] * inv_temp;
                    d = g * (expf(v - m) / sum) * (1.f + v - score);
                }
                draw[(int64_t)r * ld_draw + t] = d;
            }
        }
    }
}

// ---- device-side tensor statistics for LoggedModule.log (logged_module.py:8-18: min / max / mean / std of every logged tensor) ----------
// The reference copies each logged tensor to the host and issues four scalar reductions per call (seven calls per LSM forward); here the
// four numbers of one tensor come from ONE launch and stay on the device until somebody reads log_info.  Partial (min, max, sum, sum of
// squares) per block in double precision, combined in block order by the last block; std is the unbiased one of torch.Tensor.std().
constexpr int TS_MAX_BLOCKS = 1024;
__global__ void __launch_bounds__(256) tensor_stats_kernel(const float *__restrict__ x, int64_t n, float *__restrict__ out4, double *__restrict__ ws) {
    pdl_trigger();
    pdl_wait();
    __shared__ double sd[4][8];
    __shared__ int s_last;
    float mn = FLT_MAX, mx = -FLT_MAX;
    double s1 = 0.0, s2 = 0.0;
    auto take = [&](float v) { mn = fminf(mn, v); mx = fmaxf(mx, v); s1 += (double)v; s2 += (double)v * (double)v; };
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {                  // 128-bit loads, two in flight per thread
        const int64_t n4 = n >> 2;
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        int64_t i = tid;
        for (; i + nthr < n4; i += 2 * nthr) {
            const float4 a = __ldcs(x4 + i), b = __ldcs(x4 + i + nthr);
            take(a.x); take(a.y); take(a.z); take(a.w); take(b.x); take(b.y); take(b.z); take(b.w);
        }
        for (; i < n4; i += nthr) {
            const float4 a = __ldcs(x4 + i);
            take(a.x); take(a.y); take(a.z); take(a.w);
        }
        for (int64_t j = (n4 << 2) + tid; j < n; j += nthr) take(x[j]);
    } else {
        for (int64_t i = tid; i < n; i += nthr) take(x[i]);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    auto warp_combine = [&]() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
    };
    auto block_combine = [&](double &a, double &b, double &c, double &d) {      // fixed order: shuffle tree, then warps 0..7 in order
        warp_combine();
        if (lane == 0) { sd[0][w] = mn; sd[1][w] = mx; sd[2][w] = s1; sd[3][w] = s2; }
        __syncthreads();
        a = sd[0][0]; b = sd[1][0]; c = 0.0; d = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { a = fmin(a, sd[0][k]); b = fmax(b, sd[1][k]); c += sd[2][k]; d += sd[3][k]; }
    };
    double a, b, c, d;
    block_combine(a, b, c, d);
    double *parts = ws + 2;                    // [grid][4]
    if (threadIdx.x == 0) {
        parts[4 * blockIdx.x] = a; parts[4 * blockIdx.x + 1] = b; parts[4 * blockIdx.x + 2] = c; parts[4 * blockIdx.x + 3] = d;
        __threadfence();
        unsigned int *ticket = reinterpret_cast<unsigned int *>(ws);
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // the last block combines the per-block partials: thread t takes blocks t, t + 256, ... in order, then the same fixed-order tree
    mn = FLT_MAX; mx = -FLT_MAX; s1 = 0.0; s2 = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) {
        mn = fminf(mn, (float)parts[4 * k]); mx = fmaxf(mx, (float)parts[4 * k + 1]); s1 += parts[4 * k + 2]; s2 += parts[4 * k + 3];
    }
    __syncthreads();                           // (sd is reused)
    block_combine(a, b, c, d);
    if (threadIdx.x == 0) {
        const double mean = c / (double)n;
        const double var = n > 1 ? fmax(d - (double)n * mean * mean, 0.0) / (double)(n - 1) : 0.0;
        out4[0] = (float)a; out4[1] = (float)b; out4[2] = (float)mean; out4[3] = (float)sqrt(var);
        *reinterpret_cast<unsigned int *>(ws) = 0;
    }
}

}  // namespace loco

using namespace loco;

extern "C" {

int loco_split_bf16(const float *src, int64_t rows, int64_t cols, int64_t src_ld, uint16_t *hi, uint16_t *lo,
                    int64_t dst_ld, int transpose, void *stream) {
    LOCO_REQUIRE(rows >= 0 && cols >= 0 && src_ld >= cols, LOCO_E_BADARG, "split_bf16: bad shape rows=%lld cols=%lld src_ld=%lld",
                 (long long)rows, (long long)cols, (long long)src_ld);
    if (rows == 0 || cols == 0) return LOCO_OK;
    LOCO_REQUIRE(src && hi, LOCO_E_BADARG, "split_bf16: null pointer");
    LOCO_REQUIRE(dst_ld % 8 == 0, LOCO_E_ALIGN, "split_bf16: dst_ld %lld is not a multiple of 8", (long long)dst_ld);
    LOCO_REQUIRE((reinterpret_cast<uintptr_t>(hi) & 15) == 0 && (!lo || (reinterpret_cast<uintptr_t>(lo) & 15) == 0), LOCO_E_ALIGN,
                 "split_bf16: destination must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!transpose) {
        LOCO_REQUIRE(dst_ld >= cols, LOCO_E_BADARG, "split_bf16: dst_ld < cols");
        const int64_t total = rows * (dst_ld / 4);
        const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        LOCO_CUDA(launch_kernel(split_bf16_kernel, dim3(blocks), dim3(256), 0, st, 1, src, rows, cols, src_ld, hi, lo, dst_ld));
    count_launch();
    } else {
        LOCO_REQUIRE(dst_ld >= rows, LOCO_E_BADARG, "split_bf16: transposed dst_ld < rows");
        LOCO_REQUIRE(rows < (1ll << 31) && cols < (1ll << 31), LOCO_E_UNSUPPORTED, "split_bf16: matrix too large to transpose");
        dim3 grid((unsigned)((dst_ld + 31) / 32), (unsigned)((cols + 31) / 32));
        LOCO_REQUIRE(grid.y <= 65535, LOCO_E_UNSUPPORTED, "split_bf16: too many columns to transpose");
        LOCO_CUDA(launch_kernel(split_bf16_t_kernel, grid, dim3(256), 0, st, 1, src, (int)rows, (int)cols, src_ld, hi, lo, dst_ld));
    count_launch();
    }
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_lsm_masks(const int64_t *attention_mask, const int64_t *special_tokens_mask, int64_t n_cap, const void *region_mask,
                   int region_kind, int64_t n_reg, float *cap_mask, float *reg_mask, void *stream) {
    LOCO_REQUIRE(n_cap >= 0 && n_reg >= 0 && region_kind >= 0 && region_kind <= 2, LOCO_E_BADARG, "lsm_masks: bad arguments");
    if (n_cap + n_reg == 0) return LOCO_OK;
    LOCO_REQUIRE((n_cap == 0 || (attention_mask && special_tokens_mask && cap_mask)) && (n_reg == 0 || (region_mask && reg_mask)), LOCO_E_BADARG,
                 "lsm_masks: null pointer");
    const int64_t total = n_cap + n_reg;
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    LOCO_CUDA(launch_kernel(lsm_masks_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, attention_mask, special_tokens_mask,
                            n_cap, region_mask, region_kind, n_reg, cap_mask, reg_mask));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_lsm_prep(const float *cap, int64_t rows, int64_t cols, int64_t cap_ld, uint16_t *cap_hi, uint16_t *cap_lo, int64_t dst_ld,
                  const int64_t *attention_mask, const int64_t *special_tokens_mask, int64_t n_cap, const void *region_mask,
                  int region_kind, int64_t n_reg, float *cap_mask, float *reg_mask, void *stream) {
    LOCO_REQUIRE(rows > 0 && cols > 0 && cap_ld >= cols && n_cap > 0 && n_reg > 0 && region_kind >= 0 && region_kind <= 2, LOCO_E_BADARG,
                 "lsm_prep: bad shape rows=%lld cols=%lld n_cap=%lld n_reg=%lld", (long long)rows, (long long)cols, (long long)n_cap, (long long)n_reg);
    LOCO_REQUIRE(cap && cap_hi && attention_mask && special_tokens_mask && region_mask && cap_mask && reg_mask, LOCO_E_BADARG, "lsm_prep: null pointer");
    LOCO_REQUIRE(dst_ld % 8 == 0 && dst_ld >= cols, LOCO_E_ALIGN, "lsm_prep: dst_ld %lld must be a multiple of 8 and >= cols", (long long)dst_ld);
    LOCO_REQUIRE((reinterpret_cast<uintptr_t>(cap_hi) & 15) == 0 && (!cap_lo || (reinterpret_cast<uintptr_t>(cap_lo) & 15) == 0), LOCO_E_ALIGN,
                 "lsm_prep: destination must be 16-byte aligned");
    const int64_t total = rows * (dst_ld / 4) + n_cap + n_reg;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    LOCO_CUDA(launch_kernel(lsm_prep_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, cap, rows, cols, cap_ld, cap_hi, cap_lo,
                            dst_ld, attention_mask, special_tokens_mask, n_cap, region_mask, region_kind, n_reg, cap_mask, reg_mask));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_lsm_prep_multi(int nsplit, const float *const *src, const int64_t *rows, const int64_t *cols, const int64_t *src_ld, uint16_t *const *hi,
                        uint16_t *const *lo, const int64_t *dst_ld, const int64_t *attention_mask, const int64_t *special_tokens_mask, int64_t n_cap,
                        const void *region_mask, int region_kind, int64_t n_reg, float *cap_mask, float *reg_mask, void *stream) {
    LOCO_REQUIRE(nsplit >= 1 && nsplit <= 3 && n_cap > 0 && n_reg > 0 && region_kind >= 0 && region_kind <= 2, LOCO_E_BADARG,
                 "lsm_prep_multi: bad arguments nsplit=%d n_cap=%lld n_reg=%lld", nsplit, (long long)n_cap, (long long)n_reg);
    LOCO_REQUIRE(src && rows && cols && src_ld && hi && lo && dst_ld && attention_mask && special_tokens_mask && region_mask && cap_mask && reg_mask,
                 LOCO_E_BADARG, "lsm_prep_multi: null pointer");
    SplitJobs jobs = {};
    jobs.n = nsplit;
    int64_t total = n_cap + n_reg;
    for (int j = 0; j < nsplit; ++j) {
        LOCO_REQUIRE(rows[j] > 0 && cols[j] > 0 && src_ld[j] >= cols[j] && src[j] && hi[j], LOCO_E_BADARG, "lsm_prep_multi: bad job %d", j);
        LOCO_REQUIRE(dst_ld[j] % 8 == 0 && dst_ld[j] >= cols[j], LOCO_E_ALIGN, "lsm_prep_multi: dst_ld of job %d must be a multiple of 8 and >= cols", j);
        LOCO_REQUIRE((reinterpret_cast<uintptr_t>(hi[j]) & 15) == 0 && (!lo[j] || (reinterpret_cast<uintptr_t>(lo[j]) & 15) == 0), LOCO_E_ALIGN,
                     "lsm_prep_multi: destination of job %d must be 16-byte aligned", j);
        jobs.src[j] = src[j]; jobs.hi[j] = hi[j]; jobs.lo[j] = lo[j];
        jobs.rows[j] = rows[j]; jobs.cols[j] = cols[j]; jobs.src_ld[j] = src_ld[j]; jobs.dst_ld[j] = dst_ld[j];
        jobs.nq[j] = rows[j] * (dst_ld[j] / 4);
        total += jobs.nq[j];
    }
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    LOCO_CUDA(launch_kernel(lsm_prep_multi_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, jobs, attention_mask, special_tokens_mask,
                            n_cap, region_mask, region_kind, n_reg, cap_mask, reg_mask));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_transpose_bf16(const uint16_t *src, int64_t rows, int64_t cols, int64_t src_ld, uint16_t *dst, int64_t dst_ld,
                        void *stream) {
    LOCO_REQUIRE(rows >= 0 && cols >= 0 && src_ld >= cols && dst_ld >= rows, LOCO_E_BADARG, "transpose_bf16: bad shape rows=%lld cols=%lld",
                 (long long)rows, (long long)cols);
    if (rows == 0 || cols == 0) return LOCO_OK;
    LOCO_REQUIRE(src && dst, LOCO_E_BADARG, "transpose_bf16: null pointer");
    LOCO_REQUIRE(rows < (1ll << 31) && cols < (1ll << 31), LOCO_E_UNSUPPORTED, "transpose_bf16: matrix too large");
    dim3 grid((unsigned)((dst_ld + 31) / 32), (unsigned)((cols + 31) / 32));
    LOCO_REQUIRE(grid.y <= 65535, LOCO_E_UNSUPPORTED, "transpose_bf16: too many columns");
    LOCO_CUDA(launch_kernel(transpose16_kernel, grid, dim3(256), 0, static_cast<cudaStream_t>(stream), 1, src, (int)rows, (int)cols, src_ld, dst, dst_ld));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_box_ce_fwd_bwd(const float *logits, int64_t ld_logits, const float *lse, const int64_t *labels, int R, int K1,
                        float scale, float *loss_sum, float grad_scale, float *dlogits_f32, uint16_t *dlogits_bf16,
                        int64_t ld_bf16, void *stream) {
    LOCO_REQUIRE(R >= 0 && K1 >= 1 && ld_logits >= K1, LOCO_E_BADARG, "box_ce: bad shape R=%d K1=%d ld=%lld", R, K1, (long long)ld_logits);
    if (R == 0) return LOCO_OK;
    LOCO_REQUIRE(logits && lse && labels && loss_sum, LOCO_E_BADARG, "box_ce: null pointer");
    if (dlogits_bf16) LOCO_REQUIRE(ld_bf16 >= K1 && ld_bf16 % 8 == 0, LOCO_E_ALIGN, "box_ce: ld_bf16 must be >= K1 and a multiple of 8");
    const int wpb = 8;
    int blocks = (R + wpb - 1) / wpb;
    if (blocks > 148 * 8) blocks = 148 * 8;
    LOCO_CUDA(launch_kernel(box_ce_kernel, dim3(blocks), dim3(wpb * 32), 0, static_cast<cudaStream_t>(stream), 1, logits, ld_logits, lse, labels, R, K1,
                            scale, loss_sum, grad_scale, dlogits_f32, dlogits_bf16, ld_bf16));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_box_softmax(const float *logits, int64_t ld_logits, int R, int K1, float *lse, int64_t *argmax_fg, float *probs,
                     int64_t ld_probs, void *stream) {
    LOCO_REQUIRE(R >= 0 && K1 >= 1 && ld_logits >= K1, LOCO_E_BADARG, "box_softmax: bad shape R=%d K1=%d ld=%lld", R, K1, (long long)ld_logits);
    if (R == 0) return LOCO_OK;
    LOCO_REQUIRE(logits && (lse || argmax_fg || probs), LOCO_E_BADARG, "box_softmax: null pointer / no output requested");
    LOCO_REQUIRE(probs == nullptr || ld_probs >= K1, LOCO_E_BADARG, "box_softmax: ld_probs < K1");
    int blocks = (R + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    LOCO_CUDA(launch_kernel(box_softmax_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, logits, ld_logits, R, K1, lse,
                            argmax_fg, probs, ld_probs));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_row_normalize(const float *x, int64_t ldx, int rows, int cols, int mode, const float *dy, int64_t lddy, float *out, int64_t ldo,
                       void *stream) {
    LOCO_REQUIRE(rows >= 0 && cols >= 1 && ldx >= cols && ldo >= cols && (mode == 0 || mode == 1), LOCO_E_BADARG,
                 "row_normalize: bad arguments rows=%d cols=%d mode=%d", rows, cols, mode);
    if (rows == 0) return LOCO_OK;
    LOCO_REQUIRE(x && out && (dy == nullptr || lddy >= cols), LOCO_E_BADARG, "row_normalize: null pointer or lddy < cols");
    int blocks = (rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    LOCO_CUDA(launch_kernel(row_normalize_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, x, ldx, rows, cols, mode, dy, lddy,
                            out, ldo));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int64_t loco_box_reg_loss_workspace_bytes(int R) { return (int64_t)(16 + (R + 255) / 256) * (int64_t)sizeof(float); }

int loco_box_reg_loss(const float *deltas, int64_t ld_deltas, const float *proposal_boxes, const float *gt_boxes, const int64_t *labels, int R,
                      int K, const float *reg_weights4_host, float smooth_l1_beta, float scale, float *loss, float *ddeltas, int64_t ld_ddeltas,
                      void *workspace, void *stream) {
    LOCO_REQUIRE(R >= 0 && K >= 0 && ld_deltas >= 4 && (ddeltas == nullptr || ld_ddeltas >= 4), LOCO_E_BADARG, "box_reg_loss: bad shape R=%d", R);
    LOCO_REQUIRE(loss != nullptr && reg_weights4_host != nullptr, LOCO_E_BADARG, "box_reg_loss: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (R == 0) {
        LOCO_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
        return LOCO_OK;
    }
    LOCO_REQUIRE(deltas && proposal_boxes && gt_boxes && labels && workspace, LOCO_E_BADARG, "box_reg_loss: null pointer");
    LOCO_REQUIRE(((reinterpret_cast<uintptr_t>(proposal_boxes) | reinterpret_cast<uintptr_t>(gt_boxes)) & 15) == 0, LOCO_E_ALIGN,
                 "box_reg_loss: box arrays must be 16-byte aligned (contiguous [R,4] fp32)");
    const int blocks = (R + 255) / 256;
    LOCO_CUDA(launch_kernel(box_reg_loss_kernel, dim3(blocks), dim3(256), 0, st, 1, deltas, ld_deltas, proposal_boxes, gt_boxes, labels, R, K,
                            reg_weights4_host[0], reg_weights4_host[1], reg_weights4_host[2], reg_weights4_host[3], smooth_l1_beta, scale, loss, ddeltas,
                            ld_ddeltas, static_cast<float *>(workspace)));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

// row chunks: enough blocks for ~4 per SM (the kernel streams x once and has to keep the HBM pipe full), at least 16 rows each
static int skinny_chunks(int R, int V) {
    const int bx = (V + 511) / 512;
    int c = (4 * 148 + bx - 1) / bx;
    const int cap = (R + 15) / 16;
    if (c > cap) c = cap;
    if (c > 256) c = 256;
    return c < 1 ? 1 : c;
}
int64_t loco_skinny_grad_workspace_bytes(int R, int J, int V) { return (int64_t)skinny_chunks(R, V) * J * V * (int64_t)sizeof(float); }

int loco_skinny_grad(const float *dy, int64_t ld_dy, const float *x, int64_t ld_x, int R, int J, int V, float *dw, float *db, void *workspace,
                     void *stream) {
    LOCO_REQUIRE(R >= 0 && J >= 1 && J <= 8 && V >= 1 && ld_dy >= J && ld_x >= V, LOCO_E_BADARG, "skinny_grad: bad shape R=%d J=%d V=%d", R, J, V);
    LOCO_REQUIRE(dw != nullptr, LOCO_E_BADARG, "skinny_grad: null output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (R == 0) {
        LOCO_CUDA(cudaMemsetAsync(dw, 0, (size_t)J * V * sizeof(float), st));
        if (db) LOCO_CUDA(cudaMemsetAsync(db, 0, (size_t)J * sizeof(float), st));
        return LOCO_OK;
    }
    LOCO_REQUIRE(dy && x && workspace, LOCO_E_BADARG, "skinny_grad: null pointer");
    LOCO_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && ld_x % 4 == 0, LOCO_E_ALIGN, "skinny_grad: x must be 16-byte aligned with ld_x %% 4 == 0");
    const int nchunk = skinny_chunks(R, V), rpc = (R + nchunk - 1) / nchunk;
    dim3 grid((unsigned)((V + 511) / 512), (unsigned)nchunk);
    float *part = static_cast<float *>(workspace);
#define LOCO_SKINNY(JJ) LOCO_CUDA(launch_kernel(skinny_grad_partial_kernel<JJ>, grid, dim3(128), 0, st, 1, dy, ld_dy, x, ld_x, R, V, rpc, part))
    switch (J) {
        case 1: LOCO_SKINNY(1); break;
        case 2: LOCO_SKINNY(2); break;
        case 3: LOCO_SKINNY(3); break;
        case 4: LOCO_SKINNY(4); break;
        case 5: LOCO_SKINNY(5); break;
        case 6: LOCO_SKINNY(6); break;
        case 7: LOCO_SKINNY(7); break;
        default: LOCO_SKINNY(8); break;
    }
#undef LOCO_SKINNY
    count_launch();
    LOCO_CUDA(launch_kernel(skinny_grad_reduce_kernel, dim3((unsigned)((J * V + 63) / 64)), dim3(256), 0, st, 1, static_cast<const float *>(part), nchunk,
                            J * V, dw, dy, ld_dy, R, J, db));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_peer_exchange(int nseg, const void *const *src, const int64_t *src_pitch, const int *rows, const int64_t *row_bytes, const int64_t *dst_pitch,
                       const int64_t *dst_offset, const void *const *peers_dev, int n_peers, const void *const *flag_peers_dev, int rank, int channel,
                       int mode, void *ticket_dev, void *stream) {
    LOCO_REQUIRE(nseg >= 0 && nseg <= 4 && n_peers >= 1 && n_peers <= 64 && rank >= 0 && rank < n_peers, LOCO_E_BADARG, "peer_exchange: bad arguments nseg=%d peers=%d rank=%d",
                 nseg, n_peers, rank);
    LOCO_REQUIRE(nseg == 0 || (src && src_pitch && rows && row_bytes && dst_pitch && dst_offset && peers_dev), LOCO_E_BADARG, "peer_exchange: null pointer");
    LOCO_REQUIRE((mode & ~7) == 0 && ((mode & 1) != 0) == (nseg > 0) && mode != 0, LOCO_E_BADARG, "peer_exchange: mode %d does not match %d segments", mode, nseg);
    LOCO_REQUIRE((mode & 6) == 0 || (flag_peers_dev != nullptr && ticket_dev != nullptr && channel >= 0 && channel < 16), LOCO_E_BADARG,
                 "peer_exchange: signal / wait need flag buffers, a ticket word and 0 <= channel < 16");
    PeerSegs segs = {};
    segs.nseg = nseg;
    int64_t all = 0, max_words = 1;
    for (int k = 0; k < nseg; ++k) {
        LOCO_REQUIRE(rows[k] >= 0 && row_bytes[k] >= 0 && (rows[k] == 0 || row_bytes[k] == 0 || src[k] != nullptr), LOCO_E_BADARG, "peer_exchange: bad segment %d", k);
        LOCO_REQUIRE(src_pitch[k] >= row_bytes[k] && dst_pitch[k] >= row_bytes[k], LOCO_E_BADARG, "peer_exchange: pitch < row_bytes in segment %d", k);
        all |= row_bytes[k] | src_pitch[k] | dst_pitch[k] | dst_offset[k] | (int64_t)(reinterpret_cast<uintptr_t>(src[k]) & 15);
    }
    LOCO_REQUIRE(all % 4 == 0, LOCO_E_ALIGN, "peer_exchange: rows, pitches and offsets must be multiples of 4 bytes");
    const bool wide = all % 16 == 0;           // 16-byte words when everything allows it, 4-byte words otherwise
    const int wbytes = wide ? 16 : 4;
    for (int k = 0; k < nseg; ++k) {
        segs.src[k] = static_cast<const unsigned char *>(src[k]);
        segs.src_pitch[k] = src_pitch[k]; segs.dst_pitch[k] = dst_pitch[k]; segs.dst_offset[k] = dst_offset[k];
        segs.rows[k] = rows[k]; segs.row_words[k] = (int)(row_bytes[k] / wbytes);
        const int64_t words = (int64_t)rows[k] * segs.row_words[k];
        if (words > max_words) max_words = words;
    }
    // enough blocks to keep NVLink busy (a few hundred KB per peer), few enough that the grid is one wave beside the GEMMs
    int cap = 512 / (n_peers * (nseg > 0 ? nseg : 1));
    if (cap < 4) cap = 4;
    if (cap > 64) cap = 64;
    int blocks = (int)((max_words + 511) / 512 < cap ? (max_words + 511) / 512 : cap);
    if (blocks < 1 || nseg == 0) blocks = 1;
    const dim3 grid((unsigned)blocks, nseg > 0 ? (unsigned)n_peers : 1u, nseg > 0 ? (unsigned)nseg : 1u);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (wide)
        LOCO_CUDA(launch_kernel(peer_scatter_kernel<uint4>, grid, dim3(256), 0, st, 1, segs, (unsigned char *const *)(peers_dev), n_peers,
                                (unsigned int *const *)(flag_peers_dev), rank, channel, mode, static_cast<unsigned int *>(ticket_dev)));
    else
        LOCO_CUDA(launch_kernel(peer_scatter_kernel<uint32_t>, grid, dim3(256), 0, st, 1, segs, (unsigned char *const *)(peers_dev), n_peers,
                                (unsigned int *const *)(flag_peers_dev), rank, channel, mode, static_cast<unsigned int *>(ticket_dev)));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

static int pair_ce_strip(int Bc, int Bi) { return (Bc > Bi ? Bc : Bi) >= 128 ? 8 : 32; }

int64_t loco_pair_distill_workspace_bytes(int B) { return (int64_t)(16 + 12 * (int64_t)B) * (int64_t)sizeof(float); }

int loco_pair_distill(const float *trans, int64_t ld_trans, const float *w2r, int64_t ld_w2r, const float *r2w, int64_t ld_r2w, int B, float temperature,
                      int kind, float loss_weight, float *loss, float *g_trans, float *g_w2r, float *g_r2w, void *workspace, void *stream) {
    LOCO_REQUIRE(B >= 1 && ld_trans >= B && ld_w2r >= B && ld_r2w >= B && kind >= 0 && kind <= 2 && temperature > 0.f, LOCO_E_BADARG,
                 "pair_distill: bad arguments B=%d kind=%d temperature=%g", B, kind, (double)temperature);
    LOCO_REQUIRE(trans && w2r && r2w && loss && workspace, LOCO_E_BADARG, "pair_distill: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float inv_temp = 1.0f / temperature;
    // KD: sum of KL * T^2 / B (kldiv 'batchmean' divides by the first dimension);  MSE: 2 * (mean + mean)
    const float loss_scale = kind == 2 ? 2.0f * loss_weight / ((float)B * (float)B) : loss_weight * temperature * temperature / (float)B;
    int blocks = (2 * B + 7) / 8;
    if (blocks > 148) blocks = 148;
    LOCO_CUDA(launch_kernel(pair_distill_lines_kernel, dim3(blocks), dim3(256), 0, st, 1, trans, ld_trans, w2r, ld_w2r, r2w, ld_r2w, B, inv_temp, kind,
                            loss_scale, loss, static_cast<float *>(workspace)));
    count_launch();
    if (g_trans || g_w2r || g_r2w) {
        // d/dx of a = -x / T brings -1 / T:  KD: -weight * T / B;  MSE: 2 * (2 / B^2) * weight
        const float gscale = kind == 2 ? 4.0f * loss_weight / ((float)B * (float)B) : -loss_weight * temperature / (float)B;
        int gb = (int)(((int64_t)B * B + 255) / 256);
        if (gb > 148 * 4) gb = 148 * 4;
        LOCO_CUDA(launch_kernel(pair_distill_grad_kernel, dim3(gb), dim3(256), 0, st, 1, trans, ld_trans, w2r, ld_w2r, r2w, ld_r2w, B, inv_temp, kind, gscale,
                                g_trans, g_w2r, g_r2w, static_cast<const float *>(workspace)));
        count_launch();
    }
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_token_pool_fwd(const float *raw, int64_t ld_raw, const int32_t *seg_off, int R, int K1, float inv_temp, int alignment, float *scores,
                        int64_t ld_scores, float *att, void *stream) {
    LOCO_REQUIRE(R >= 0 && K1 >= 1 && ld_scores >= K1 && inv_temp > 0.f && (alignment == 0 || alignment == 1), LOCO_E_BADARG,
                 "token_pool_fwd: bad arguments R=%d K1=%d alignment=%d", R, K1, alignment);
    if (R == 0) return LOCO_OK;
    LOCO_REQUIRE(raw && seg_off && scores, LOCO_E_BADARG, "token_pool_fwd: null pointer");
    const int blocks = (int)std::min<int64_t>(((int64_t)R * K1 + 255) / 256, (int64_t)current_device_sm_count() * 16);
    LOCO_CUDA(launch_kernel(token_pool_kernel<false>, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, raw, ld_raw, seg_off, R, K1, inv_temp,
                            alignment, scores, ld_scores, att, (const float *)nullptr, (int64_t)0, (float *)nullptr, (int64_t)0));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_token_pool_bwd(const float *raw, int64_t ld_raw, const int32_t *seg_off, int R, int K1, float inv_temp, int alignment, const float *dscores,
                        int64_t ld_dscores, float *draw, int64_t ld_draw, void *stream) {
    LOCO_REQUIRE(R >= 0 && K1 >= 1 && ld_dscores >= K1 && inv_temp > 0.f && (alignment == 0 || alignment == 1), LOCO_E_BADARG,
                 "token_pool_bwd: bad arguments R=%d K1=%d alignment=%d", R, K1, alignment);
    if (R == 0) return LOCO_OK;
    LOCO_REQUIRE(raw && seg_off && dscores && draw, LOCO_E_BADARG, "token_pool_bwd: null pointer");
    const int blocks = (int)std::min<int64_t>(((int64_t)R * K1 + 255) / 256, (int64_t)current_device_sm_count() * 16);
    LOCO_CUDA(launch_kernel(token_pool_kernel<true>, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, raw, ld_raw, seg_off, R, K1, inv_temp,
                            alignment, (float *)nullptr, (int64_t)0, (float *)nullptr, dscores, ld_dscores, draw, ld_draw));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int64_t loco_tensor_stats_workspace_bytes(void) { return (int64_t)(2 + 4 * TS_MAX_BLOCKS) * (int64_t)sizeof(double); }

int loco_tensor_stats(const float *x, int64_t n, float *out4, void *workspace, void *stream) {
    LOCO_REQUIRE(n >= 1 && x && out4 && workspace, LOCO_E_BADARG, "tensor_stats: bad arguments n=%lld", (long long)n);
    LOCO_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, LOCO_E_ALIGN, "tensor_stats: workspace must be 8-byte aligned");
    const int64_t cap = std::min<int64_t>(TS_MAX_BLOCKS, (int64_t)current_device_sm_count() * 4);
    int blocks = (int)std::min<int64_t>((n + 2047) / 2048, cap);
    if (blocks < 1) blocks = 1;
    LOCO_CUDA(launch_kernel(tensor_stats_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, x, n, out4, static_cast<double *>(workspace)));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int64_t loco_pair_ce_workspace_bytes(int nmat, int Bc, int Bi) {
    const int strip = pair_ce_strip(Bc, Bi);
    const int nblk = ((Bc > Bi ? Bc : Bi) + strip - 1) / strip;
    return (int64_t)(32 + Bc + Bi + (int64_t)nmat * nblk * 4) * (int64_t)sizeof(float);
}

int loco_pair_ce(float *pw, int nmat, int64_t mat_stride, int64_t ld, int Bc, int Bi, int diag_offset, const float *cap_mask,
                 int T, const float *reg_mask, int Rg, float *out4, float *dpw_caption, float *dpw_image, void *workspace,
                 void *stream) {
    LOCO_REQUIRE(Bc > 0 && Bi > 0 && ld >= Bi && T > 0 && Rg > 0 && diag_offset >= 0 && nmat >= 1 && nmat <= 16, LOCO_E_BADARG,
                 "pair_ce: bad shape nmat=%d Bc=%d Bi=%d ld=%lld", nmat, Bc, Bi, (long long)ld);
    LOCO_REQUIRE(pw && cap_mask && reg_mask && out4, LOCO_E_BADARG, "pair_ce: null pointer");
    const size_t base = (32 + (size_t)Bc + Bi + 128) * sizeof(float);
    LOCO_REQUIRE(base <= 48 * 1024, LOCO_E_UNSUPPORTED, "pair_ce: batch too large (Bc=%d Bi=%d)", Bc, Bi);
    // one CTA with the matrix staged in shared memory up to 64 x 64 (N = 2 of the sharded head: one launch instead of prep + strips)
    static const int staged_max = []() { const char *e = getenv("LOCOV_B200_PCE_STAGED_MAX"); const int v = e ? atoi(e) : 64; return (v >= 32 && v <= 96) ? v : 64; }();
    const int strip_w = (Bc <= staged_max && Bi <= staged_max) ? staged_max : pair_ce_strip(Bc, Bi);
    const int nblk = ((Bc > Bi ? Bc : Bi) + strip_w - 1) / strip_w;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (nblk == 1) {
        const size_t staged = base + (size_t)Bc * (Bi + 1) * sizeof(float);
        LOCO_CUDA(launch_kernel(pair_ce_kernel<true>, dim3(nmat, 1), dim3(1024), staged, st, 1, pw, mat_stride, ld, Bc, Bi, diag_offset, cap_mask, T,
                                reg_mask, Rg, out4, dpw_caption, dpw_image, static_cast<float *>(nullptr)));
    } else {
        LOCO_REQUIRE(workspace != nullptr, LOCO_E_BADARG, "pair_ce: B > 32 needs loco_pair_ce_workspace_bytes() of ZERO-INITIALISED workspace "
                     "(the kernel leaves it zeroed again)");
        const size_t strip = base + (size_t)(3 * 32 * 33 + 96) * sizeof(float);      // per-warp partial statistics of the column pass
        int pblocks = (Bc + Bi + 7) / 8;
        if (pblocks > 148) pblocks = 148;
        LOCO_CUDA(launch_kernel(pair_ce_prep_kernel, dim3(pblocks), dim3(256), 0, st, 1, static_cast<const float *>(pw), nmat, mat_stride, ld, Bc, Bi, cap_mask, T,
                                reg_mask, Rg, static_cast<float *>(workspace)));
        count_launch();
        if (strip_w == 8)
            LOCO_CUDA(launch_kernel(pair_ce_kernel<false, 8>, dim3(nmat, nblk), dim3(1024), strip, st, 1, pw, mat_stride, ld, Bc, Bi, diag_offset, cap_mask, T,
                                    reg_mask, Rg, out4, dpw_caption, dpw_image, static_cast<float *>(workspace)));
        else
            LOCO_CUDA(launch_kernel(pair_ce_kernel<false, 32>, dim3(nmat, nblk), dim3(1024), strip, st, 1, pw, mat_stride, ld, Bc, Bi, diag_offset, cap_mask, T,
                                    reg_mask, Rg, out4, dpw_caption, dpw_image, static_cast<float *>(workspace)));
    }
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

}  // extern "C"
