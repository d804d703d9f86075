// box_infer.cu — the inference tail of the box predictor: Detectron2's `fast_rcnn_inference` as the reference
// reaches it from `EmbeddingFastRCNNOutputLayers.inference` (ovr/modeling/roi_heads/roi_emb_heads.py:280 and :357 ->
// FastRCNNOutputLayers.inference -> fast_rcnn_inference_single_image): box decoding + clipping, score threshold, per-class
// NMS, top-k per image.  With K = 1203 classes the reference's tail is a [R, K] `nonzero`, an index gather and a Python loop over the
// classes around torchvision's NMS; here it is four launches for all images of the batch, no host synchronisation:
//
//   1. decode   : warp per RoI — Box2BoxTransform.apply_deltas + Boxes.clip (lane 0) and the finiteness scan of the RoI's score row.
//   2. iou mask : the embedding head regresses ONE class-agnostic box per RoI (box_emb_head.py:137), so the "IoU > nms_thresh"
//                 relation between the RoIs of an image is the same for every class: one bit matrix [R_img, R_img] per image.
//   3. class nms: CTA = (image, 8 consecutive classes), warp = class: gather the class's candidates (score > thresh) from a
//                 sector-wide staged tile of the probability matrix, bitonic-sort them by (score desc, RoI asc) in shared memory,
//                 greedy NMS as a walk over the sorted list with a "removed" bit set held across the lanes, OR-ing in the bit-matrix
//                 row of every kept RoI.  At most `topk` survivors per class are written (more can never reach the image's top-k).
//   4. top-k    : CTA per image: radix select of the k-th largest 64-bit key (score bits | inverted candidate index — unique, so
//                 ties resolve to the lower (RoI, class) index exactly as the reference's stable sort does), bitonic sort of the
//                 selected keys, outputs.
// Candidate order, tie rules and the kept-row numbering (rows are numbered among the VALID rows of the image, as the reference's
// boolean filtering leaves them) follow fast_rcnn_inference_single_image; the IoU expression is torchvision's
// (inter / (area_a + area_b - inter) > thresh, individually rounded fp32 operations).
#include <cstdint>
#include <cstdlib>

#include "common.cuh"

namespace loco {

constexpr int BI_THREADS = 256;
constexpr int BI_WARPS = BI_THREADS / 32;
constexpr int BI_CW = 8;            // classes per CTA of the class-NMS kernel (= warps; one 32-byte sector of a score row)
constexpr int BI_CHUNK = 256;       // rows staged per round

__device__ __forceinline__ uint32_t ordered_bits(float f) {       // monotone float -> uint (scores are > thresh >= 0 here, but be general)
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// ---- 1. decode + clip + validity -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BI_THREADS) box_decode_kernel(const float *__restrict__ probs, int64_t ldp, const float *__restrict__ deltas, int64_t ldd,
                                                                const float *__restrict__ proposals, const int32_t *__restrict__ img_off,
                                                                const float *__restrict__ img_hw, int n_img, int R, int K1, float wx, float wy, float ww,
                                                                float wh, float clampv, float4 *__restrict__ boxes, int32_t *__restrict__ valid) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * BI_WARPS + (threadIdx.x >> 5);
    if (r >= R) return;
    bool fin = true;
    for (int k = lane; k < K1; k += 32) fin = fin && isfinite(probs[(int64_t)r * ldp + k]);
    fin = __all_sync(0xffffffffu, fin);
    if (lane == 0) {
        int im = 0;
        while (im + 1 < n_img && r >= img_off[im + 1]) ++im;              // (a handful of images)
        const float ih = img_hw[2 * im], iw = img_hw[2 * im + 1];
        const float4 p = *reinterpret_cast<const float4 *>(proposals + (int64_t)r * 4);
        const float *d = deltas + (int64_t)r * ldd;
        // Box2BoxTransform.apply_deltas, operation by operation
        const float w = __fsub_rn(p.z, p.x), h = __fsub_rn(p.w, p.y);
        const float cx = __fadd_rn(p.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(p.y, __fmul_rn(0.5f, h));
        const float dx = __fdiv_rn(d[0], wx), dy = __fdiv_rn(d[1], wy);
        float dw = __fdiv_rn(d[2], ww), dh = __fdiv_rn(d[3], wh);
        dw = dw > clampv ? clampv : dw;                                   // torch.clamp(max=): a NaN stays a NaN (fminf would drop it)
        dh = dh > clampv ? clampv : dh;
        const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
        const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
        float4 b = make_float4(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), __fsub_rn(pcy, __fmul_rn(0.5f, ph)), __fadd_rn(pcx, __fmul_rn(0.5f, pw)),
                               __fadd_rn(pcy, __fmul_rn(0.5f, ph)));
        const bool bfin = isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w);
        // torch.clamp(min=0, max=size): NaN stays NaN; such rows are invalid and never used
        b.x = fminf(fmaxf(b.x, 0.f), iw); b.y = fminf(fmaxf(b.y, 0.f), ih); b.z = fminf(fmaxf(b.z, 0.f), iw); b.w = fminf(fmaxf(b.w, 0.f), ih);
        boxes[r] = b;
        valid[r] = (fin && bfin) ? 1 : 0;
    }
}

// ---- 2. IoU bit matrix per image ----------------------------------------------------------------------------------------------------
// mask[r][w] bit b = IoU(box r, box (row0 + 32 w + b)) > thresh, for the rows of r's image.  Warp per row r, lane = word.
__global__ void __launch_bounds__(BI_THREADS) box_iou_mask_kernel(const float4 *__restrict__ boxes, const int32_t *__restrict__ img_off, int n_img, int R, int W32,
                                                                  float thresh, uint32_t *__restrict__ mask) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * BI_WARPS + (threadIdx.x >> 5);
    if (r >= R) return;
    int im = 0;
    while (im + 1 < n_img && r >= img_off[im + 1]) ++im;
    const int row0 = img_off[im], nrow = img_off[im + 1] - row0;
    const float4 a = boxes[r];
    const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    for (int w = lane; w < W32; w += 32) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int j = 32 * w + b;
            if (j >= nrow) break;
            const float4 c = boxes[row0 + j];
            const float sb = __fmul_rn(__fsub_rn(c.z, c.x), __fsub_rn(c.w, c.y));
            const float iw = fmaxf(__fsub_rn(fminf(a.z, c.z), fmaxf(a.x, c.x)), 0.f);
            const float ih = fmaxf(__fsub_rn(fminf(a.w, c.w), fmaxf(a.y, c.y)), 0.f);
            const float inter = __fmul_rn(iw, ih);
            const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
            if (iou > thresh) bits |= 1u << b;
        }
        mask[(int64_t)r * W32 + w] = bits;
    }
}

// ---- 3. per-class NMS ----------------------------------------------------------------------------------------------------------------
// key = ordered score bits << 32 | (0xffffffff - local RoI index): descending key order = (score desc, RoI asc).
__device__ __forceinline__ void warp_bitonic_desc(uint64_t *key, int npad, int lane) {
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < npad; i += 32) {
                const int x = i ^ j;
                if (x > i) {
                    const uint64_t a = key[i], b = key[x];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { key[i] = b; key[x] = a; }
                }
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(BI_THREADS) box_class_nms_kernel(const float *__restrict__ probs, int64_t ldp, const int32_t *__restrict__ valid,
                                                                   const int32_t *__restrict__ img_off, const uint32_t *__restrict__ mask, int W32, int K,
                                                                   int ngroups, float score_thresh, int topk, int rmax, uint64_t *__restrict__ cls_keys,
                                                                   int32_t *__restrict__ cls_count) {
    extern __shared__ __align__(16) unsigned char bi_smem[];
    float *tile = reinterpret_cast<float *>(bi_smem);                                   // [BI_CHUNK][BI_CW]
    uint64_t *keys = reinterpret_cast<uint64_t *>(bi_smem + BI_CHUNK * BI_CW * sizeof(float));   // [BI_WARPS][rmax]
    pdl_trigger();
    pdl_wait();
    const int im = blockIdx.x / ngroups, grp = blockIdx.x - im * ngroups;
    const int k0 = grp * BI_CW;
    const int row0 = img_off[im], nrow = img_off[im + 1] - row0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = k0 + warp;
    uint64_t *mykeys = keys + (size_t)warp * rmax;
    int n = 0;                                                                          // candidates of this warp's class (warp-uniform)
    for (int c0 = 0; c0 < nrow; c0 += BI_CHUNK) {
        __syncthreads();
        for (int i = threadIdx.x; i < BI_CHUNK * BI_CW; i += BI_THREADS) {              // 8 consecutive threads = one 32-byte sector of a row
            const int rr = c0 + i / BI_CW, kk = k0 + (i % BI_CW);
            float v = -1.0f;
            if (rr < nrow && kk < K && valid[row0 + rr]) v = probs[(int64_t)(row0 + rr) * ldp + kk];
            tile[i] = v;
        }
        __syncthreads();
        if (k < K) {
            for (int j = lane; j < BI_CHUNK; j += 32) {
                const float v = tile[j * BI_CW + warp];
                const bool take = (c0 + j < nrow) && (v > score_thresh);
                const uint32_t bal = __ballot_sync(0xffffffffu, take);
                if (take) mykeys[n + __popc(bal & ((1u << lane) - 1))] = ((uint64_t)ordered_bits(v) << 32) | (uint32_t)(0xffffffffu - (uint32_t)(c0 + j));
                n += __popc(bal);
            }
        }
    }
    if (k >= K) return;
    int32_t *cnt = cls_count + (int64_t)im * K + k;
    if (n == 0) {
        if (lane == 0) *cnt = 0;
        return;
    }
    int npad = 1;
    while (npad < n) npad <<= 1;
    for (int i = n + lane; i < npad; i += 32) mykeys[i] = 0;
    __syncwarp();
    warp_bitonic_desc(mykeys, npad, lane);
    // greedy walk: `removed` bits of RoIs suppressed by a kept RoI of this class; lane holds words lane, lane + 32 (W32 <= 64)
    uint32_t rem0 = 0, rem1 = 0;
    uint64_t *outk = cls_keys + ((int64_t)im * K + k) * topk;
    int kept = 0;
    const uint32_t *mbase = mask + (int64_t)row0 * W32;
    // the bit-matrix row of the next candidate is always in flight (most candidates of a dense class are already removed, but when one is
    // kept its row is needed at once)
    auto row_of = [&](int i) -> int { return (int)(0xffffffffu - (uint32_t)(mykeys[i] & 0xffffffffu)); };
    int rn = row_of(0);
    uint32_t m0 = lane < W32 ? mbase[(int64_t)rn * W32 + lane] : 0u;
    uint32_t m1 = lane + 32 < W32 ? mbase[(int64_t)rn * W32 + lane + 32] : 0u;
    for (int i = 0; i < n && kept < topk; ++i) {
        const int r = rn;
        const uint32_t c0m = m0, c1m = m1;
        if (i + 1 < n) {
            rn = row_of(i + 1);
            m0 = lane < W32 ? mbase[(int64_t)rn * W32 + lane] : 0u;
            m1 = lane + 32 < W32 ? mbase[(int64_t)rn * W32 + lane + 32] : 0u;
        }
        const int word = r >> 5;
        const uint32_t wv = __shfl_sync(0xffffffffu, (word >> 5) ? rem1 : rem0, word & 31);
        if ((wv >> (r & 31)) & 1u) continue;                       // suppressed by a higher-scoring RoI of this class
        if (lane == 0) outk[kept] = mykeys[i];
        ++kept;
        rem0 |= c0m;
        rem1 |= c1m;
    }
    if (lane == 0) *cnt = kept;
}

// ---- 4. top-k per image -------------------------------------------------------------------------------------------------------------
// Final key = score bits << 32 | (0xffffffff - (valid-rank of the RoI * K + class)): the candidate order of `nonzero()` on the filtered
// [R_valid, K] score matrix.  Radix select (8 x 8 bits, from the top) of the k-th largest, then a bitonic sort of the selected keys.
constexpr int TK_THREADS = 1024;
constexpr int TK_MAX = 1024;        // topk <= TK_MAX

__global__ void __launch_bounds__(TK_THREADS) box_topk_kernel(const uint64_t *__restrict__ cls_keys, const int32_t *__restrict__ cls_count,
                                                              const int32_t *__restrict__ valid, const int32_t *__restrict__ img_off,
                                                              const float4 *__restrict__ boxes, int K, int topk, float4 *__restrict__ out_boxes,
                                                              float *__restrict__ out_scores, int64_t *__restrict__ out_classes, int64_t *__restrict__ out_rows,
                                                              int32_t *__restrict__ out_count, int32_t *__restrict__ vrank) {
    __shared__ uint32_t hist[256];
    __shared__ uint64_t sel[TK_MAX];
    __shared__ int s_total, s_nsel;
    __shared__ uint64_t s_prefix;
    __shared__ int s_need;
    __shared__ int s_digit;
    pdl_trigger();
    pdl_wait();
    const int im = blockIdx.x;
    const int row0 = img_off[im], nrow = img_off[im + 1] - row0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = TK_THREADS / 32;
    // rank of every row among the valid rows of the image (exclusive count) — the reference filters invalid rows out before indexing
    if (warp == 0) {
        int run = 0;
        for (int c0 = 0; c0 < nrow; c0 += 32) {
            const int v = (c0 + lane < nrow) ? valid[row0 + c0 + lane] : 0;
            const uint32_t bal = __ballot_sync(0xffffffffu, v != 0);
            if (c0 + lane < nrow) vrank[row0 + c0 + lane] = run + __popc(bal & ((1u << lane) - 1));
            run += __popc(bal);
        }
    }
    if (threadIdx.x == 0) { s_total = 0; s_prefix = 0; s_nsel = 0; }
    __syncthreads();
    const uint64_t *kb = cls_keys + (int64_t)im * K * topk;
    const int32_t *cb = cls_count + (int64_t)im * K;
    auto final_key = [&](uint64_t key, int k) -> uint64_t {
        const uint32_t r = 0xffffffffu - (uint32_t)(key & 0xffffffffu);
        const uint32_t flat = (uint32_t)vrank[row0 + r] * (uint32_t)K + (uint32_t)k;
        return (key & 0xffffffff00000000ull) | (uint32_t)(0xffffffffu - flat);
    };
    {   // total survivors
        int t = 0;
        for (int k = threadIdx.x; k < K; k += TK_THREADS) t += cb[k];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0 && t) atomicAdd(&s_total, t);
    }
    __syncthreads();
    const int total = s_total;
    const int want = min(topk, total);
    if (threadIdx.x == 0) { out_count[im] = want; s_need = want; }
    if (want == 0) return;
    __syncthreads();
    // radix select: after the loop s_prefix is the want-th largest final key
    for (int pass = 0; pass < 8 && total > want; ++pass) {
        const int shift = 56 - 8 * pass;
        if (threadIdx.x < 256) hist[threadIdx.x] = 0;
        __syncthreads();
        const uint64_t prefix = s_prefix;
        const uint64_t pmask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int k = warp; k < K; k += NW) {
            const int c = cb[k];
            for (int e = lane; e < c; e += 32) {
                const uint64_t fk = final_key(kb[(int64_t)k * topk + e], k);
                if ((fk & pmask) == prefix) atomicAdd(&hist[(fk >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int need = s_need, d = 255;
            for (; d > 0; --d) {
                if ((int)hist[d] >= need) break;
                need -= (int)hist[d];
            }
            s_need = need;
            s_digit = d;
            s_prefix = prefix | ((uint64_t)d << shift);
        }
        __syncthreads();
    }
    const uint64_t kth = (total > want) ? s_prefix : 0ull;
    for (int k = warp; k < K; k += NW) {
        const int c = cb[k];
        for (int e = lane; e < c; e += 32) {
            const uint64_t fk = final_key(kb[(int64_t)k * topk + e], k);
            if (fk >= kth) {
                const int slot = atomicAdd(&s_nsel, 1);
                if (slot < TK_MAX) sel[slot] = fk;
            }
        }
    }
    __syncthreads();
    int npad = 1;
    while (npad < want) npad <<= 1;
    for (int i = want + threadIdx.x; i < npad; i += TK_THREADS) sel[i] = 0;
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npad; i += TK_THREADS) {
                const int x = i ^ j;
                if (x > i) {
                    const uint64_t a = sel[i], b = sel[x];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { sel[i] = b; sel[x] = a; }
                }
            }
            __syncthreads();
        }
    // outputs: the valid-rank -> row map is monotone; invert it by search over the image's rows
    for (int i = threadIdx.x; i < want; i += TK_THREADS) {
        const uint64_t fk = sel[i];
        const uint32_t flat = 0xffffffffu - (uint32_t)(fk & 0xffffffffu);
        const int vr = (int)(flat / (uint32_t)K), cls = (int)(flat - (uint32_t)vr * (uint32_t)K);
        int lo = 0, hi = nrow - 1;                      // smallest row with vrank == vr and valid
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const int vm = vrank[row0 + mid] + (valid[row0 + mid] ? 1 : 0);     // number of valid rows in [0, mid]
            if (vm >= vr + 1) hi = mid; else lo = mid + 1;
        }
        const int64_t o = (int64_t)im * topk + i;
        out_boxes[o] = boxes[row0 + lo];
        out_scores[o] = from_ordered_bits((uint32_t)(fk >> 32));
        out_classes[o] = cls;
        out_rows[o] = vr;
    }
}

}  // namespace loco

using namespace loco;

extern "C" {

static int box_inference_w32(int max_rows) { return (max_rows + 31) / 32; }

int64_t loco_box_inference_workspace_bytes(int R, int K, int n_img, int max_rows_per_image, int topk) {
    if (R <= 0 || K <= 0 || n_img <= 0 || topk <= 0) return 16;
    const int64_t W32 = box_inference_w32(max_rows_per_image);
    int64_t b = 0;
    b += ((int64_t)R * 16 + 255) / 256 * 256;                               // decoded boxes
    b += ((int64_t)R * 4 + 255) / 256 * 256;                                // valid flags
    b += ((int64_t)R * 4 + 255) / 256 * 256;                                // valid ranks
    b += ((int64_t)R * W32 * 4 + 255) / 256 * 256;                          // IoU bit matrix
    b += ((int64_t)n_img * K * 4 + 255) / 256 * 256;                        // survivors per (image, class)
    b += (int64_t)n_img * K * topk * 8;                                     // their keys
    return b;
}

int loco_box_inference(const float *probs, int64_t ld_probs, const float *deltas, int64_t ld_deltas, const float *proposals, const int32_t *img_offsets,
                       const float *img_hw, int n_img, int max_rows_per_image, int R, int K, const float *reg_weights4_host, float scale_clamp,
                       float score_thresh, float nms_thresh, int topk, float *out_boxes, float *out_scores, int64_t *out_classes, int64_t *out_rows,
                       int32_t *out_count, void *workspace, void *stream) {
    LOCO_REQUIRE(R >= 0 && K >= 1 && n_img >= 1 && topk >= 1 && ld_probs >= K + 1 && ld_deltas >= 4 && max_rows_per_image >= 0, LOCO_E_BADARG,
                 "box_inference: bad shape R=%d K=%d n_img=%d topk=%d", R, K, n_img, topk);
    LOCO_REQUIRE(topk <= TK_MAX, LOCO_E_UNSUPPORTED, "box_inference: topk %d (at most %d)", topk, TK_MAX);
    LOCO_REQUIRE(max_rows_per_image <= 2048, LOCO_E_UNSUPPORTED, "box_inference: %d RoIs in one image (at most 2048)", max_rows_per_image);
    LOCO_REQUIRE((int64_t)max_rows_per_image * K < (1ll << 32), LOCO_E_UNSUPPORTED, "box_inference: RoIs x classes per image exceeds 32 bits");
    LOCO_REQUIRE(out_count && img_offsets && img_hw && reg_weights4_host, LOCO_E_BADARG, "box_inference: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (R == 0) {
        LOCO_CUDA(cudaMemsetAsync(out_count, 0, (size_t)n_img * sizeof(int32_t), st));
        return LOCO_OK;
    }
    LOCO_REQUIRE(probs && deltas && proposals && out_boxes && out_scores && out_classes && out_rows && workspace, LOCO_E_BADARG, "box_inference: null pointer");
    LOCO_REQUIRE(((reinterpret_cast<uintptr_t>(proposals) | reinterpret_cast<uintptr_t>(out_boxes) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0, LOCO_E_ALIGN,
                 "box_inference: proposals, out_boxes and workspace must be 16-byte aligned");
    const int W32 = box_inference_w32(max_rows_per_image);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    auto carve = [&](int64_t bytes) { unsigned char *p = ws; ws += (bytes + 255) / 256 * 256; return p; };
    float4 *boxes = reinterpret_cast<float4 *>(carve((int64_t)R * 16));
    int32_t *valid = reinterpret_cast<int32_t *>(carve((int64_t)R * 4));
    int32_t *vrank = reinterpret_cast<int32_t *>(carve((int64_t)R * 4));
    uint32_t *mask = reinterpret_cast<uint32_t *>(carve((int64_t)R * W32 * 4));
    int32_t *cls_count = reinterpret_cast<int32_t *>(carve((int64_t)n_img * K * 4));
    uint64_t *cls_keys = reinterpret_cast<uint64_t *>(ws);
    const int rblocks = (R + BI_WARPS - 1) / BI_WARPS;
    LOCO_CUDA(launch_kernel(box_decode_kernel, dim3(rblocks), dim3(BI_THREADS), 0, st, 1, probs, ld_probs, deltas, ld_deltas, proposals, img_offsets, img_hw, n_img,
                            R, K + 1, reg_weights4_host[0], reg_weights4_host[1], reg_weights4_host[2], reg_weights4_host[3], scale_clamp, boxes, valid));
    count_launch();
    LOCO_CUDA(launch_kernel(box_iou_mask_kernel, dim3(rblocks), dim3(BI_THREADS), 0, st, 1, (const float4 *)boxes, img_offsets, n_img, R, W32, nms_thresh, mask));
    count_launch();
    int rmax = 32;
    while (rmax < max_rows_per_image) rmax <<= 1;
    const size_t smem = (size_t)BI_CHUNK * BI_CW * sizeof(float) + (size_t)BI_WARPS * rmax * sizeof(uint64_t);
    static thread_local size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        LOCO_CUDA(cudaFuncSetAttribute(box_class_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const int ngroups = (K + BI_CW - 1) / BI_CW;
    LOCO_CUDA(launch_kernel(box_class_nms_kernel, dim3(n_img * ngroups), dim3(BI_THREADS), smem, st, 1, probs, ld_probs, (const int32_t *)valid, img_offsets,
                            (const uint32_t *)mask, W32, K, ngroups, score_thresh, topk, rmax, cls_keys, cls_count));
    count_launch();
    LOCO_CUDA(launch_kernel(box_topk_kernel, dim3(n_img), dim3(TK_THREADS), 0, st, 1, (const uint64_t *)cls_keys, (const int32_t *)cls_count, (const int32_t *)valid,
                            img_offsets, (const float4 *)boxes, K, topk, reinterpret_cast<float4 *>(out_boxes), out_scores, out_classes, out_rows, out_count, vrank));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

}  // extern "C"
