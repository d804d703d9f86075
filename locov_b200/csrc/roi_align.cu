// roi_align.cu — RoIAlign (ROIAlignV2, aligned=True) forward / backward / sampling-grid dump.
//
// Replaces the torchvision::roi_align CUDA kernel the reference reaches at
//   /root/reference/ovr/modeling/roi_heads/roi_emb_heads.py:182-187 (ROIPooler construction)
//   /root/reference/ovr/modeling/roi_heads/roi_emb_heads.py:243-245 (self.pooler(features, boxes))
//
// Design (B200, HBM-write bound: out is R*C*PH*PW*4 B, the feature map is L2 resident) — DESIGN.md 4.1:
//   * features are consumed channels-last ([N,H,W,C]); an NCHW input is transposed once per call by
//     nchw_to_nhwc_kernel into caller-provided workspace (<4 % of the traffic at the configs).
//   * vectorised kernel (roi_align_fwd_v4_kernel): one CTA = (roi, 1-4 slabs of 128 channels), warp = bin row,
//     lane = 4 consecutive channels (128-bit taps), so all control flow (adaptive sample counts, skipped
//     taps) is warp-uniform.  Generic kernel (any C / alignment): one channel per lane, 32-channel slabs.
//   * the per-roi sampling tables (y taps, x taps) are computed ONCE per CTA with explicitly rounded
//     fp32 intrinsics (__fmul_rn/__fadd_rn/__fdiv_rn: no FMA contraction) so that coordinates and
//     integer tap indices are bit-identical to the CPU reference arithmetic.
//   * bilinear pooling is evaluated separably with MERGED tap lists: per bin and axis the weights of the
//     samples are summed per distinct feature row / column first (samples are < 1 px apart by
//     construction of the adaptive grid), so each pixel of a bin's footprint is loaded once per bin row.
//     The common case walks consecutive feature columns with a sliding register window (walk_bin_row).
//   * results are staged in a [128][PH*PW] shared-memory tile and leave through the copy engine
//     (four TMA tensor stores per slab, evict-first so the feature map stays in the 126 MB L2) or, for pooled
//     sizes whose rows are not 16-byte multiples, as coalesced st.global.cs rows.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace loco {

struct RoiGeom {
    float sh, sw, bh, bw;
    int gh, gw;
    int batch;
};

// Exact restatement of the reference arithmetic; every operation individually rounded.
__device__ __forceinline__ RoiGeom roi_geometry(const float *__restrict__ roi, float scale, int aligned, int PH, int PW,
                                                int sampling_ratio) {
    RoiGeom g;
    const float off = aligned ? 0.5f : 0.0f;
    g.batch = (int)roi[0];
    const float start_w = __fsub_rn(__fmul_rn(roi[1], scale), off);
    const float start_h = __fsub_rn(__fmul_rn(roi[2], scale), off);
    const float end_w = __fsub_rn(__fmul_rn(roi[3], scale), off);
    const float end_h = __fsub_rn(__fmul_rn(roi[4], scale), off);
    float rw = __fsub_rn(end_w, start_w);
    float rh = __fsub_rn(end_h, start_h);
    if (!aligned) {
        rw = fmaxf(rw, 1.0f);
        rh = fmaxf(rh, 1.0f);
    }
    g.bh = __fdiv_rn(rh, (float)PH);
    g.bw = __fdiv_rn(rw, (float)PW);
    g.gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(g.bh);
    g.gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(g.bw);
    g.sh = start_h;
    g.sw = start_w;
    return g;
}

// coordinate of sample `i` (of `g` per bin) in bin `p`:  start + p*bin + ((i + .5)*bin)/g
__device__ __forceinline__ float sample_coord(float start, float bin, int p, int i, int g) {
    const float a = __fadd_rn(start, __fmul_rn((float)p, bin));
    const float b = __fdiv_rn(__fmul_rn(__fadd_rn((float)i, 0.5f), bin), (float)g);
    return __fadd_rn(a, b);
}

// One-axis tap: low/high index and the two interpolation weights; low = -1 marks a skipped sample
// (coordinate outside [-1, size]).
struct Tap {
    int lo, hi;
    float wl, wh;   // weight of value[lo] (= 1 - frac) and value[hi] (= frac)
};
__device__ __forceinline__ Tap make_tap(float v, int size) {
    Tap t;
    if (v < -1.0f || v > (float)size) {
        t.lo = -1; t.hi = -1; t.wl = 0.f; t.wh = 0.f;
        return t;
    }
    if (v <= 0.0f) v = 0.0f;
    int lo = __float2int_rz(v), hi;
    if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else { hi = lo + 1; }
    const float frac = __fsub_rn(v, (float)lo);
    t.lo = lo; t.hi = hi; t.wh = frac; t.wl = __fsub_rn(1.0f, frac);
    return t;
}

// ------------------------------------------------------------------------------------------------
// NCHW -> NHWC transpose (per image: [C, HW] -> [HW, C]) through a padded 32x32 tile.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float *__restrict__ src, float *__restrict__ dst, int C,
                                                           int HW) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const float *s = src + (size_t)n * C * HW;
    float *d = dst + (size_t)n * C * HW;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, p = p0 + tx;
        if (c < C && p < HW) tile[ty + 8 * k][tx] = __ldg(s + (size_t)c * HW + p);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int p = p0 + ty + 8 * k, c = c0 + tx;
        if (c < C && p < HW) d[(size_t)p * C + c] = tile[tx][ty + 8 * k];
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
constexpr int RA_WARPS = 7;
constexpr int RA_THREADS = RA_WARPS * 32;
constexpr int RA_CC = 32;          // channels per CTA (lane = channel)
constexpr int RA_TAB = 512;        // merged (index, weight) entries held per axis
constexpr int RA_UCOLS = 48;       // feature columns whose vertically pooled values a warp keeps in shared memory

// Merged tap list of one bin along one axis: the gh (gw) samples of a bin are < 1 px apart, so their 2*g bilinear
// taps hit at most g + 1 distinct rows (columns); adding up the weights per distinct index first means every
// feature pixel of the bin's footprint is loaded ONCE (L2 -> SM traffic ~ footprint, not 4 loads per sample).
struct MergedEntry {
    int idx;
    float w;
};

// entries of bin `p` are written at tab[p * stride ...]; returns their number
__device__ __forceinline__ int build_merged(MergedEntry *tab, float start, float bin, int p, int g, int size, int mul = 1,
                                            bool tag_parity = false) {
    int n = 0;
    // taps are monotonic in the sample index, so an index can only repeat one of the last two entries
    auto add = [&](int idx, float w) {
        if (n > 0 && tab[n - 1].idx == idx) tab[n - 1].w += w;
        else if (n > 1 && tab[n - 2].idx == idx) tab[n - 2].w += w;
        else { tab[n].idx = idx; tab[n].w = w; ++n; }
    };
    for (int i = 0; i < g; ++i) {
        const Tap t = make_tap(sample_coord(start, bin, p, i, g), size);
        if (t.lo < 0) continue;                       // sample outside [-1, size]: contributes 0
        add(t.lo, t.wl);
        add(t.hi, t.wh);
    }
    if (mul != 1)
        for (int i = 0; i < n; ++i)                      // index -> byte offset (v4 kernel; mul is a multiple of 16)
            tab[i].idx = tab[i].idx * mul | (tag_parity ? (tab[i].idx & 1) : 0);
    return n;
}

__global__ void __launch_bounds__(RA_THREADS, 3) roi_align_fwd_generic_kernel(const float *__restrict__ feat,   // [N,H,W,C]
                                                                   const float *__restrict__ rois, int C, int H, int W,
                                                                   int PH, int PW, float scale, int sampling_ratio,
                                                                   int aligned, int nchunks, float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MergedEntry *ytab = reinterpret_cast<MergedEntry *>(smem_raw);
    MergedEntry *xtab = ytab + RA_TAB;
    int *ycnt = reinterpret_cast<int *>(xtab + RA_TAB);      // [PH]
    int *xcnt = ycnt + PH;                                    // [PW]
    float *tile = reinterpret_cast<float *>(xcnt + PW);
    float *ubuf_all = tile + RA_CC * ((PH * PW) | 1);        // [RA_WARPS][RA_UCOLS][32]

    const int r = blockIdx.x / nchunks;
    const int chunk = blockIdx.x - r * nchunks;
    const int c0 = chunk * RA_CC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PHW = PH * PW;
    const int tstride = PHW | 1;

    const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, aligned, PH, PW, sampling_ratio);
    // max merged entries per bin: adaptive grids space the samples <= 1 px apart (g + 1 distinct indices, +1 slack);
    // a fixed sampling_ratio may spread them further (2 per sample)
    const int ystride = sampling_ratio > 0 ? 2 * g.gh : g.gh + 2, xstride = sampling_ratio > 0 ? 2 * g.gw : g.gw + 2;
    const bool empty = (g.gh <= 0 || g.gw <= 0);
    const bool tables_fit = !empty && (long long)PH * ystride <= RA_TAB && (long long)PW * xstride <= RA_TAB;
    const float cnt = (float)max(g.gh * g.gw, 1);

    if (tables_fit) {
        const int t = threadIdx.x;
        if (t < PH) ycnt[t] = build_merged(ytab + t * ystride, g.sh, g.bh, t, g.gh, H);
        else if (t >= 32 && t < 32 + PW) xcnt[t - 32] = build_merged(xtab + (t - 32) * xstride, g.sw, g.bw, t - 32, g.gw, W);
    }
    __syncthreads();

    const int c = c0 + lane;
    const bool active = c < C;
    const float *fb = feat + (size_t)g.batch * H * W * C + (active ? c : 0);
    float *trow = tile + lane * tstride;
    const size_t rowpitch = (size_t)W * C;

    if (empty) {
        for (int ph = warp; ph < PH; ph += RA_WARPS)
            for (int pw = 0; pw < PW; ++pw) trow[ph * PW + pw] = 0.f;
    } else if (tables_fit) {
        // column range touched by this roi (min / max over every merged entry: a fixed sampling_ratio on a box with
        // x2 < x1 walks the columns backwards, so the first / last list entries are not the extremes)
        int xmin = 0x7fffffff, xmax = -1;
        for (int e = lane; e < PW * xstride; e += 32) {
            const int pw = e / xstride, k = e - pw * xstride;
            if (k < xcnt[pw]) {
                xmin = min(xmin, xtab[e].idx);
                xmax = max(xmax, xtab[e].idx);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
            xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        }
        if (xmax < 0) xmin = 0;
        const int ncols = xmax - xmin + 1;
        float *ubuf = ubuf_all + (size_t)warp * RA_UCOLS * 32 + lane;
        for (int ph = warp; ph < PH; ph += RA_WARPS) {
            const MergedEntry *yt = ytab + ph * ystride;
            const int ny = ycnt[ph];
            if (ncols <= RA_UCOLS) {
                // pass 1: vertically pooled value u(x) of every column of the footprint, four columns per step so
                // that 4 * ny independent coalesced loads are in flight per warp (the kernel is latency-bound otherwise)
                if (active) {
                    for (int x0 = xmin; x0 <= xmax; x0 += 4) {
                        float u0 = 0.f, u1 = 0.f, u2 = 0.f, u3 = 0.f;
                        const float *f0 = fb + (size_t)x0 * C;
                        const float *f1 = fb + (size_t)min(x0 + 1, xmax) * C;
                        const float *f2 = fb + (size_t)min(x0 + 2, xmax) * C;
                        const float *f3 = fb + (size_t)min(x0 + 3, xmax) * C;
#pragma unroll 2
                        for (int k = 0; k < ny; ++k) {
                            const size_t ro = (size_t)yt[k].idx * rowpitch;
                            const float wy = yt[k].w;
                            const float a = __ldg(f0 + ro), b = __ldg(f1 + ro), c2 = __ldg(f2 + ro), d = __ldg(f3 + ro);
                            u0 = fmaf(wy, a, u0);
                            u1 = fmaf(wy, b, u1);
                            u2 = fmaf(wy, c2, u2);
                            u3 = fmaf(wy, d, u3);
                        }
                        float *ub = ubuf + (size_t)(x0 - xmin) * 32;
                        ub[0] = u0;
                        if (x0 + 1 <= xmax) ub[32] = u1;
                        if (x0 + 2 <= xmax) ub[64] = u2;
                        if (x0 + 3 <= xmax) ub[96] = u3;
                    }
                }
                __syncwarp();
                // pass 2: horizontal pooling of each bin from the buffered column values
                for (int pw = 0; pw < PW; ++pw) {
                    float acc = 0.f;
                    const MergedEntry *xt = xtab + pw * xstride;
                    const int nx = xcnt[pw];
                    if (active && ny > 0)
                        for (int k = 0; k < nx; ++k) acc = fmaf(xt[k].w, ubuf[(size_t)(xt[k].idx - xmin) * 32], acc);
                    trow[ph * PW + pw] = __fdiv_rn(acc, cnt);
                }
                __syncwarp();
            } else {
                // very wide roi: bin-driven sweep with the last column cached
                auto column = [&](int x) -> float {
                    const float *fx = fb + (size_t)x * C;
                    float u0 = 0.f, u1 = 0.f;
                    int k = 0;
                    for (; k + 1 < ny; k += 2) {
                        const float a = __ldg(fx + (size_t)yt[k].idx * rowpitch);
                        const float b = __ldg(fx + (size_t)yt[k + 1].idx * rowpitch);
                        u0 = fmaf(yt[k].w, a, u0);
                        u1 = fmaf(yt[k + 1].w, b, u1);
                    }
                    if (k < ny) u0 = fmaf(yt[k].w, __ldg(fx + (size_t)yt[k].idx * rowpitch), u0);
                    return u0 + u1;
                };
                int kx = -1;
                float uk = 0.f;
                for (int pw = 0; pw < PW; ++pw) {
                    float acc = 0.f;
                    const MergedEntry *xt = xtab + pw * xstride;
                    const int nx = xcnt[pw];
                    if (active && ny > 0) {
                        for (int k = 0; k < nx; ++k) {
                            const int x = xt[k].idx;
                            if (x != kx) { uk = column(x); kx = x; }
                            acc = fmaf(xt[k].w, uk, acc);
                        }
                    }
                    trow[ph * PW + pw] = __fdiv_rn(acc, cnt);
                }
            }
        }
    } else {
        // generic path for rois whose sampling grid exceeds the shared tables: direct taps
        for (int ph = warp; ph < PH; ph += RA_WARPS)
            for (int pw = 0; pw < PW; ++pw) {
                float acc = 0.f;
                for (int iy = 0; iy < g.gh; ++iy) {
                    const Tap ty = make_tap(sample_coord(g.sh, g.bh, ph, iy, g.gh), H);
                    if (ty.lo < 0) continue;
                    for (int ix = 0; ix < g.gw; ++ix) {
                        const Tap tx = make_tap(sample_coord(g.sw, g.bw, pw, ix, g.gw), W);
                        if (tx.lo < 0 || !active) continue;
                        const float v1 = __ldg(fb + ((size_t)ty.lo * W + tx.lo) * C);
                        const float v2 = __ldg(fb + ((size_t)ty.lo * W + tx.hi) * C);
                        const float v3 = __ldg(fb + ((size_t)ty.hi * W + tx.lo) * C);
                        const float v4 = __ldg(fb + ((size_t)ty.hi * W + tx.hi) * C);
                        acc += ty.wl * tx.wl * v1 + ty.wl * tx.wh * v2 + ty.wh * tx.wl * v3 + ty.wh * tx.wh * v4;
                    }
                }
                trow[ph * PW + pw] = __fdiv_rn(acc, cnt);
            }
    }
    __syncthreads();

    // coalesced streaming write-out of the [cc, PH*PW] slab (contiguous in the NCHW output)
    const int cc = min(RA_CC, C - c0);
    const int total = cc * PHW;
    float *ob = out + ((size_t)r * C + c0) * PHW;
    int ch = threadIdx.x / PHW, j = threadIdx.x - ch * PHW;
    for (int i = threadIdx.x; i < total; i += RA_THREADS) {
        __stcs(ob + i, tile[ch * tstride + j]);
        j += RA_THREADS;
        while (j >= PHW) { j -= PHW; ++ch; }
    }
}


// ------------------------------------------------------------------------------------------------
// Launch order of the rois: heaviest first.  The kernels below run one CTA per (roi, channel chunk) in block-index order; a roi's cost
// grows with its area (more feature columns per bin row, more row taps, the bin-driven fall-backs for the widest sampling grids), the
// 9 % of the benchmark's boxes above 448 px cost twice the average — and when such a roi is scheduled late, its CTAs are the tail
// the whole grid waits for (measured: the same boxes sorted largest-first run 10-12 % faster).  One small block buckets the rois
// by log2(area in feature pixels) and emits the roi indices bucket by bucket, heaviest bucket first (order inside a bucket is
// whatever the atomics give: it only affects scheduling, never results).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) roi_order_kernel(const float *__restrict__ rois, int R, float scale, int *__restrict__ order) {
    pdl_trigger();
    pdl_wait();
    __shared__ int cnt[16], base[16];
    if (threadIdx.x < 16) cnt[threadIdx.x] = 0;
    __syncthreads();
    auto bucket = [&](int r) -> int {
        const float *q = rois + (size_t)r * 5;
        const float a = fmaxf((q[3] - q[1]) * scale, 1.f) * fmaxf((q[4] - q[2]) * scale, 1.f);
        const int b = a == a ? (int)floorf(log2f(a) * 1.5f) : 0;       // (NaN boxes: lightest bucket)
        return min(max(b, 0), 15);
    };
    for (int r = threadIdx.x; r < R; r += blockDim.x) atomicAdd(&cnt[bucket(r)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 15; b >= 0; --b) { base[b] = run; run += cnt[b]; }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) order[atomicAdd(&base[bucket(r)], 1)] = r;
}

// ------------------------------------------------------------------------------------------------
// forward, vectorised: one CTA = (roi, 128-channel slab), lane = 4 consecutive channels (one LDG.128 per tap
// serves 4 channels, so the per-tap table lookups / address arithmetic are amortised 4x and a warp request is a
// full 512-byte run of the NHWC pixel), warp = one bin row.
//   * merged tap tables hold BYTE offsets (row offset = y * W * C * 4, column offset = x * C * 4), built once
//     per CTA with the exactly rounded coordinate arithmetic above.
//   * a bin row walks its bins left to right; the merged column lists of neighbouring bins are monotonic and
//     share at most their boundary column, so caching the last vertically pooled column in registers makes every
//     column of the row's footprint get loaded exactly once per bin row (no shared-memory column buffer).
//     The rows shared by neighbouring bin rows are re-read through L1 (ld.global.nc), not L2.
//   * up to four row taps per column are issued as independent predicated 128-bit loads (adaptive grids with
//     gh <= 3 have at most 4 merged rows); longer lists continue in a loop.
//   * results go to a [128][S] shared tile with channel 4*l + j stored at row j*32 + l.  For even PW two
//     neighbouring bins are stored as one 8-byte word and S = 2 (mod 4): both the column-wise STS.64 and the
//     row-wise LDS.64 of the write-out are bank-conflict free; odd PW uses scalar accesses with odd S.
//   * write-out: the slab is contiguous in the NCHW output; fully coalesced st.global.cs (evict-first).
// ------------------------------------------------------------------------------------------------
constexpr int RV_WARPS = 14;
constexpr int RV_THREADS = RV_WARPS * 32;
constexpr int RV_CC = 128;
// shared-memory layout of the vectorised kernel: tap tables (2 x RA_TAB entries), 4 x 32 ints + 4 header ints, 32 walk
// weight vectors, then the tile on the next 128-byte boundary
constexpr int RV_TILE_OFFSET = ((2 * RA_TAB * 8 + (128 + 4) * 4 + 32 * 16) + 127) / 128 * 128;

__device__ __forceinline__ float4 ldg128(const char *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void fma4(float4 &a, float w, const float4 &v) {
    a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
}

// Four channels of one lane as two packed fp32 pairs: the arithmetic below uses Blackwell's packed FFMA2 / FMUL2
// (two fp32 FMAs per issue slot) — the kernel is issue-bound, not FLOP-bound.
struct Quad {
    float2 lo, hi;
};
__device__ __forceinline__ Quad ldg_quad(const char *p) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    Quad q;
    q.lo = make_float2(v.x, v.y);
    q.hi = make_float2(v.z, v.w);
    return q;
}
__device__ __forceinline__ void quad_fma(Quad &a, float w, const Quad &v) {
    const float2 ww = make_float2(w, w);
    a.lo = __ffma2_rn(ww, v.lo, a.lo);
    a.hi = __ffma2_rn(ww, v.hi, a.hi);
}
__device__ __forceinline__ Quad quad_mul(float w, const Quad &v) {
    const float2 ww = make_float2(w, w);
    Quad r;
    r.lo = __fmul2_rn(ww, v.lo);
    r.hi = __fmul2_rn(ww, v.hi);
    return r;
}

// Vertically pooled value of one footprint column for 4 channels: NY independent 128-bit loads, then the FMAs.
template <int NY>
__device__ __forceinline__ Quad pooled_column(const char *p, const int (&yo)[4], const float (&yw)[4]) {
    Quad v[NY];
#pragma unroll
    for (int k = 0; k < NY; ++k) v[k] = ldg_quad(p + yo[k]);
    Quad u = quad_mul(yw[0], v[0]);
#pragma unroll
    for (int k = 1; k < NY; ++k) quad_fma(u, yw[k], v[k]);
    return u;
}

// One bin row (PW bins) of one roi for the 4 channels of this lane.  NY = number of merged row taps (1..4 static,
// 0 = run-time count with the taps read from shared memory).  Two pooled columns are cached in registers, slot =
// parity of the column index (bit 0 of the table offset): the merged column lists of neighbouring bins are monotonic
// and overlap in at most two (consecutive) columns, so every column of the row's footprint is loaded exactly once.
// The slot tags are compared at run time, so a non-monotonic list (fixed sampling ratio on a flipped box) only costs
// extra loads.  All branches are warp-uniform.
// Where the pooled 4 channels of bin `pw` go.  OUT = 0: the shared-memory tile of the NCHW write-out (row = channel; trow = this lane's
// tile row of the bin row, rstep = 32 tile rows).  OUT = 1 / 2: straight to a CHANNELS-LAST output [R, PH, PW, C] in fp32 / bf16 —
// the lane's 4 channels are 16 / 8 contiguous bytes, a warp writes 512 / 256 contiguous bytes per bin, no staging, no write-out phase
// (trow = address of the lane's channels in bin 0 of the bin row, NULL for lanes past C; rstep = bytes per bin = C * element size).
template <int OUT>
__device__ __forceinline__ void put_bin(float *trow, int rstep, int pw, const Quad &a) {
    if (OUT == 0) {
        trow[pw] = a.lo.x; trow[pw + rstep] = a.lo.y; trow[pw + 2 * rstep] = a.hi.x; trow[pw + 3 * rstep] = a.hi.y;
    } else if (trow != nullptr) {
        char *p = reinterpret_cast<char *>(trow) + (size_t)pw * rstep;
        if (OUT == 1) __stcs(reinterpret_cast<float4 *>(p), make_float4(a.lo.x, a.lo.y, a.hi.x, a.hi.y));
        else __stcs(reinterpret_cast<uint2 *>(p), make_uint2(pack_bf16x2_rn(a.lo.x, a.lo.y), pack_bf16x2_rn(a.hi.x, a.hi.y)));
    }
}

template <int NY, bool PAIR, int OUT>
__device__ __noinline__ void pool_bin_row(const char *fb, const MergedEntry *yt, int ny, const MergedEntry *xtab,
                                             const int *xcnt, int xstride, int PW, float inv_cnt, float *trow, int rstep) {
    int yo[4] = {0, 0, 0, 0};
    float yw[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < (NY > 0 ? NY : 1); ++k) {
        if (NY > 0) { yo[k] = yt[k].idx; yw[k] = yt[k].w; }
    }
    Quad zero;
    zero.lo = make_float2(0.f, 0.f);
    zero.hi = zero.lo;
    int off0 = -1, off1 = -1;
    Quad u0 = zero, u1 = zero;
    auto column = [&](int off) -> Quad {
        const char *p = fb + off;
        if (NY > 0) return pooled_column<(NY > 0 ? NY : 1)>(p, yo, yw);
        Quad t = zero;
        for (int k = 0; k < ny; ++k) quad_fma(t, yt[k].w, ldg_quad(p + yt[k].idx));
        return t;
    };
    auto bin = [&](const MergedEntry *xt, int nx) -> Quad {
        Quad acc = zero;
#pragma unroll 1
        for (int j = 0; j < nx; ++j) {
            const MergedEntry e = xt[j];
            if (e.idx & 1) {
                if (e.idx != off1) { u1 = column(e.idx - 1); off1 = e.idx; }
                quad_fma(acc, e.w, u1);
            } else {
                if (e.idx != off0) { u0 = column(e.idx); off0 = e.idx; }
                quad_fma(acc, e.w, u0);
            }
        }
        return quad_mul(inv_cnt, acc);
    };
    if (PAIR) {
#pragma unroll 1
        for (int pw = 0; pw < PW; pw += 2) {
            const Quad a = bin(xtab, xcnt[pw]);
            const Quad b = bin(xtab + xstride, xcnt[pw + 1]);
            xtab += 2 * xstride;
            *reinterpret_cast<float2 *>(trow + pw) = make_float2(a.lo.x, b.lo.x);
            *reinterpret_cast<float2 *>(trow + pw + rstep) = make_float2(a.lo.y, b.lo.y);
            *reinterpret_cast<float2 *>(trow + pw + 2 * rstep) = make_float2(a.hi.x, b.hi.x);
            *reinterpret_cast<float2 *>(trow + pw + 3 * rstep) = make_float2(a.hi.y, b.hi.y);
        }
    } else {
#pragma unroll 1
        for (int pw = 0; pw < PW; ++pw) {
            const Quad a = bin(xtab, xcnt[pw]);
            xtab += xstride;
            put_bin<OUT>(trow, rstep, pw, a);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Column-walk form of a bin row (the common case: adaptive sampling grid with GW = ceil(bin_w) <= 3).
// The gw samples of a bin are < 1 px apart and span < bin_w <= GW px, so the merged taps of bin pw lie in the GW + 1
// consecutive columns [base(pw), base(pw) + GW], and base(pw) is non-decreasing.  The row is therefore a single
// left-to-right walk over CONSECUTIVE feature columns with a sliding window of GW + 1 vertically pooled columns in
// registers: per bin a 4-float dense weight vector and an advance count (both built once per roi — the x structure is the
// same for all bin rows and channel slabs).  No per-tap table lookups, index compares or cache-slot branches, and the loads
// of the next column are always issued one advance ahead of their use (they are unconditional), which is what hides the
// L2 latency the bin-driven walk above exposes (profiles/README.md: long-scoreboard stalls, 45 % issue utilisation).
// ------------------------------------------------------------------------------------------------
struct WalkBin {
    float w[4];      // dense weights of columns base .. base + 3 (zero beyond GW)
};

template <int GW, int NY, bool PAIR, int OUT>
__device__ __noinline__ void walk_bin_row(const char *fb, const MergedEntry *yt, const WalkBin *xw, const int *xadv, int x0_bytes,
                                           int colstride, int xlast_bytes, int PW, float inv_cnt, float *trow, int rstep) {
    // xlast_bytes = byte offset of the LAST column the walk consumes: the look-ahead load past it is redirected to it
    // (an L1 hit instead of a new 512-byte line per bin row: rows of small rois are only 3-15 columns long)
    int yo[NY];
    float yw[NY];
#pragma unroll
    for (int k = 0; k < NY; ++k) { yo[k] = yt[k].idx; yw[k] = yt[k].w; }
    Quad u[GW + 1];
    Quad raw[NY];
#pragma unroll
    for (int k = 0; k <= GW; ++k) { u[k].lo = make_float2(0.f, 0.f); u[k].hi = u[k].lo; }
    int sb = x0_bytes;                                  // byte offset of the next column of the stream
    {
        const char *p = fb + min(sb, xlast_bytes);
#pragma unroll
        for (int k = 0; k < NY; ++k) raw[k] = ldg_quad(p + yo[k]);
    }
    auto bin = [&](int pw) -> Quad {
        int a = xadv[pw];
#pragma unroll 1
        for (; a > 0; --a) {
#pragma unroll
            for (int k = 0; k < GW; ++k) u[k] = u[k + 1];
            Quad t = quad_mul(yw[0], raw[0]);
#pragma unroll
            for (int k = 1; k < NY; ++k) quad_fma(t, yw[k], raw[k]);
            u[GW] = t;
            sb += colstride;
            const char *p = fb + min(sb, xlast_bytes);  // columns past the last one carry zero weight: clamp the address
#pragma unroll
            for (int k = 0; k < NY; ++k) raw[k] = ldg_quad(p + yo[k]);
        }
        const float4 w = *reinterpret_cast<const float4 *>(xw[pw].w);
        Quad acc = quad_mul(w.x, u[0]);
        quad_fma(acc, w.y, u[1]);
        if (GW >= 2) quad_fma(acc, w.z, u[GW >= 2 ? 2 : 0]);
        if (GW >= 3) quad_fma(acc, w.w, u[GW >= 3 ? 3 : 0]);
        return quad_mul(inv_cnt, acc);
    };
    if (PAIR) {
#pragma unroll 1
        for (int pw = 0; pw < PW; pw += 2) {           // even PW: paired 8-byte tile stores
            const Quad a = bin(pw);
            const Quad b = bin(pw + 1);
            *reinterpret_cast<float2 *>(trow + pw) = make_float2(a.lo.x, b.lo.x);
            *reinterpret_cast<float2 *>(trow + pw + rstep) = make_float2(a.lo.y, b.lo.y);
            *reinterpret_cast<float2 *>(trow + pw + 2 * rstep) = make_float2(a.hi.x, b.hi.x);
            *reinterpret_cast<float2 *>(trow + pw + 3 * rstep) = make_float2(a.hi.y, b.hi.y);
        }
    } else {
#pragma unroll 1
        for (int pw = 0; pw < PW; ++pw) {
            const Quad a = bin(pw);
            put_bin<OUT>(trow, rstep, pw, a);
        }
    }
}

template <int GW, bool PAIR, int OUT>
__device__ __forceinline__ bool walk_dispatch(int ny, const char *fb, const MergedEntry *yt, const WalkBin *xw, const int *xadv,
                                              int x0_bytes, int colstride, int xlast_bytes, int PW, float inv_cnt, float *trow, int rstep) {
    switch (ny) {                                      // warp-uniform
        case 1: walk_bin_row<GW, 1, PAIR, OUT>(fb, yt, xw, xadv, x0_bytes, colstride, xlast_bytes, PW, inv_cnt, trow, rstep); return true;
        case 2: walk_bin_row<GW, 2, PAIR, OUT>(fb, yt, xw, xadv, x0_bytes, colstride, xlast_bytes, PW, inv_cnt, trow, rstep); return true;
        case 3: walk_bin_row<GW, 3, PAIR, OUT>(fb, yt, xw, xadv, x0_bytes, colstride, xlast_bytes, PW, inv_cnt, trow, rstep); return true;
        case 4: walk_bin_row<GW, 4, PAIR, OUT>(fb, yt, xw, xadv, x0_bytes, colstride, xlast_bytes, PW, inv_cnt, trow, rstep); return true;
        default: return false;
    }
}

template <bool PAIR, bool MULTI, int OUT>
__global__ void __launch_bounds__(RV_THREADS, 2) roi_align_fwd_v4_kernel(const float *__restrict__ feat,   // [N,H,W,C]
                                                                        const float *__restrict__ rois, int C, int H, int W,
                                                                        int PH, int PW, float scale, int sampling_ratio,
                                                                        int aligned, int nchunks, int slabs, int tstride,
                                                                        int bulk_out, float *__restrict__ out,
                                                                        const __grid_constant__ CUtensorMap omap, const int *__restrict__ order) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MergedEntry *ytab = reinterpret_cast<MergedEntry *>(smem_raw);
    MergedEntry *xtab = ytab + RA_TAB;
    int *ycnt = reinterpret_cast<int *>(xtab + RA_TAB);      // [32]
    int *xcnt = ycnt + 32;                                    // [32]
    int *xadv = xcnt + 32;                                    // [32] column-walk advance counts
    int *xbase = xadv + 32;                                   // [32] first column of each bin (-1: no sample inside the map)
    int *walk_hdr = xbase + 32;                               // [4]  {-, first column of the walk, last column, -}
    WalkBin *xw = reinterpret_cast<WalkBin *>(walk_hdr + 4);  // [32] dense per-bin column weights (16-byte aligned)
    float *tile = reinterpret_cast<float *>(smem_raw + RV_TILE_OFFSET);   // [RV_CC][tstride], 128-byte aligned (TMA store source)

    const int slot = blockIdx.x / nchunks;
    const int chunk = blockIdx.x - slot * nchunks;            // this CTA pools `slabs` consecutive 128-channel slabs
    const int r = order != nullptr ? order[slot] : slot;      // launch order: heaviest rois first (roi_order_kernel)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PHW = PH * PW;

    const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, aligned, PH, PW, sampling_ratio);
    const int ystride = sampling_ratio > 0 ? 2 * g.gh : g.gh + 2, xstride = sampling_ratio > 0 ? 2 * g.gw : g.gw + 2;
    const bool empty = (g.gh <= 0 || g.gw <= 0);
    const bool tables_fit = !empty && (long long)PH * ystride <= RA_TAB && (long long)PW * xstride <= RA_TAB;
    const float inv_cnt = __fdiv_rn(1.0f, (float)max(g.gh * g.gw, 1));

    // column-walk program, part 1 (by the thread that built the bin's merged list): first column and dense weights
    const bool walk_try = tables_fit && sampling_ratio <= 0 && g.gw >= 1 && g.gw <= 3;
    bool walk_bad_a = false;
    if (tables_fit) {
        const int t = threadIdx.x;
        if (t < PH) ycnt[t] = build_merged(ytab + t * ystride, g.sh, g.bh, t, g.gh, H, W * C * 4);
        else if (t >= 32 && t < 32 + PW) {
            const int pw = t - 32;
            const MergedEntry *xt = xtab + pw * xstride;
            const int nx = build_merged(xtab + pw * xstride, g.sw, g.bw, pw, g.gw, W, C * 4, true);
            xcnt[pw] = nx;
            if (walk_try) {
                const int colbytes = C * 4;
                float w[4] = {0.f, 0.f, 0.f, 0.f};
                int base = -1;
                if (nx > 0) {
                    base = (xt[0].idx & ~1) / colbytes;                              // (bit 0 of the byte offset is the parity tag)
                    for (int j = 0; j < nx; ++j) {
                        const int d = (xt[j].idx & ~1) / colbytes - base;
                        if (d < 0 || d > g.gw) walk_bad_a = true;
                        else w[d] += xt[j].w;
                    }
                }
                xbase[pw] = base;
                xw[pw].w[0] = w[0]; xw[pw].w[1] = w[1]; xw[pw].w[2] = w[2]; xw[pw].w[3] = w[3];
            }
        }
    }
    if (threadIdx.x == 0) { walk_hdr[1] = 0; walk_hdr[2] = 0; }
    __syncthreads();
    // column-walk program (see walk_bin_row), part 2: advance counts from the bases of the preceding bins
    bool walk_bad = !walk_try;
    if (walk_try && threadIdx.x >= 32 && threadIdx.x < 32 + PW) {
        const int pw = threadIdx.x - 32;
        walk_bad = walk_bad_a;
        const int base = xbase[pw];
        int adv = 0;
        if (base >= 0) {
            int prev = -1;
            for (int q = pw - 1; q >= 0 && prev < 0; --q) prev = xbase[q];
            if (prev < 0) { adv = g.gw + 1; walk_hdr[1] = base; }      // first non-empty bin primes the whole window
            else adv = base - prev;
            if (adv < 0 || adv > g.gw + 1) walk_bad = true;
            atomicMax(&walk_hdr[2], base + g.gw);                     // last column the walk consumes
        }
        xadv[pw] = adv;
    }
    const bool walk_ok = __syncthreads_and(walk_bad ? 0 : 1) != 0;
#ifdef LOCOV_ROI_PROLOGUE_ONLY          // developer timing probe: what the per-CTA prologue costs on its own (results are garbage)
    if (walk_ok || !walk_ok) return;
#endif
    const int walk_x0 = walk_hdr[1] * C * 4, walk_colstride = C * 4, walk_xlast = min(W - 1, walk_hdr[2]) * C * 4;

    // Warp roles: a group of PH warps pools one 128-channel slab (warp = bin row); with PH <= RV_WARPS / 2 (7 x 7 pooling)
    // several groups work on consecutive slabs at the same time, each into its own tile, so no warp idles.
    // (MULTI = false is the single-group instantiation: one tile, every warp index is a bin row)
    const int nconc = MULTI ? max(1, RV_WARPS / PH) : 1;
    const int wgroup = MULTI ? warp / PH : 0, wrow = warp - wgroup * PH;
    const int row_step = nconc == 1 ? RV_WARPS : PH;
    const size_t tile_floats = (size_t)RV_CC * tstride;
    const int cl = 4 * lane;
    float *t0 = tile + (wgroup < nconc ? wgroup : 0) * tile_floats + (size_t)lane * tstride;   // rows lane, 32 + lane, 64 + lane, 96 + lane
    // OUT = 0: rstep = 32 tile rows; channels-last output: bytes per bin (C * element size), t0 is re-pointed per slab below
    const int esize = OUT == 2 ? 2 : 4;
    const int rstep = OUT == 0 ? 32 * tstride : C * esize;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto store_bin = [&](int o, const float4 &a) {
        if (OUT == 0) {
            t0[o] = a.x; t0[o + rstep] = a.y; t0[o + 2 * rstep] = a.z; t0[o + 3 * rstep] = a.w;
        } else {
            Quad q;
            q.lo = make_float2(a.x, a.y);
            q.hi = make_float2(a.z, a.w);
            put_bin<OUT>(t0, rstep, o, q);
        }
    };

    for (int sl0 = 0; sl0 < slabs; sl0 += nconc) {
        const int sl = sl0 + wgroup;
        const int c0 = (chunk * slabs + sl) * RV_CC;
        const bool work = wgroup < nconc && sl < slabs && c0 < C;   // warp-uniform
        const bool active = work && (c0 + cl) < C;             // C % 4 == 0: a lane's 4 channels are all in or all out
        const char *fb = reinterpret_cast<const char *>(feat + (size_t)g.batch * H * W * C + (active ? c0 + cl : 0));
        if (OUT == 0) {
            if (sl0 > 0) __syncthreads();                      // the previous slabs' write-out has drained the tiles
        } else {                                               // channels-last output: this lane's 4 channels of bin 0 of the roi
            t0 = active ? reinterpret_cast<float *>(reinterpret_cast<char *>(out) + ((size_t)r * PHW * C + c0 + cl) * esize) : nullptr;
        }

        if (!work) {
        } else if (empty) {
            for (int ph = wrow; ph < PH; ph += row_step)
                for (int pw = 0; pw < PW; ++pw) store_bin(ph * PW + pw, zero4);
        } else if (tables_fit) {
            for (int ph = wrow; ph < PH; ph += row_step) {
                const MergedEntry *yt = ytab + ph * ystride;
                const int ny = ycnt[ph];
                float *trow = OUT == 0 ? t0 + ph * PW
                                       : (t0 != nullptr ? reinterpret_cast<float *>(reinterpret_cast<char *>(t0) + (size_t)ph * PW * rstep) : nullptr);
                if (walk_ok) {                                 // CTA-uniform: column walk (GW = sampling grid width of this roi)
                    bool done;
                    if (g.gw == 1) done = walk_dispatch<1, PAIR, OUT>(ny, fb, yt, xw, xadv, walk_x0, walk_colstride, walk_xlast, PW, inv_cnt, trow, rstep);
                    else if (g.gw == 2) done = walk_dispatch<2, PAIR, OUT>(ny, fb, yt, xw, xadv, walk_x0, walk_colstride, walk_xlast, PW, inv_cnt, trow, rstep);
                    else done = walk_dispatch<3, PAIR, OUT>(ny, fb, yt, xw, xadv, walk_x0, walk_colstride, walk_xlast, PW, inv_cnt, trow, rstep);
                    if (done) continue;
                }
                switch (ny) {                                  // warp-uniform
                    case 1: pool_bin_row<1, PAIR, OUT>(fb, yt, ny, xtab, xcnt, xstride, PW, inv_cnt, trow, rstep); break;
                    case 2: pool_bin_row<2, PAIR, OUT>(fb, yt, ny, xtab, xcnt, xstride, PW, inv_cnt, trow, rstep); break;
                    case 3: pool_bin_row<3, PAIR, OUT>(fb, yt, ny, xtab, xcnt, xstride, PW, inv_cnt, trow, rstep); break;
                    case 4: pool_bin_row<4, PAIR, OUT>(fb, yt, ny, xtab, xcnt, xstride, PW, inv_cnt, trow, rstep); break;
                    default: pool_bin_row<0, PAIR, OUT>(fb, yt, ny, xtab, xcnt, xstride, PW, inv_cnt, trow, rstep); break;
                }
            }
        } else {
            // rois whose sampling grid exceeds the shared tables: direct taps (rare: > 36 samples per bin and axis)
            for (int ph = wrow; ph < PH; ph += row_step)
                for (int pw = 0; pw < PW; ++pw) {
                    float4 acc = zero4;
                    for (int iy = 0; iy < g.gh; ++iy) {
                        const Tap ty = make_tap(sample_coord(g.sh, g.bh, ph, iy, g.gh), H);
                        if (ty.lo < 0) continue;
                        for (int ix = 0; ix < g.gw; ++ix) {
                            const Tap tx = make_tap(sample_coord(g.sw, g.bw, pw, ix, g.gw), W);
                            if (tx.lo < 0 || !active) continue;
                            const size_t pc = (size_t)C * 4;
                            fma4(acc, ty.wl * tx.wl, ldg128(fb + ((size_t)ty.lo * W + tx.lo) * pc));
                            fma4(acc, ty.wl * tx.wh, ldg128(fb + ((size_t)ty.lo * W + tx.hi) * pc));
                            fma4(acc, ty.wh * tx.wl, ldg128(fb + ((size_t)ty.hi * W + tx.lo) * pc));
                            fma4(acc, ty.wh * tx.wh, ldg128(fb + ((size_t)ty.hi * W + tx.hi) * pc));
                        }
                    }
                    acc.x *= inv_cnt; acc.y *= inv_cnt; acc.z *= inv_cnt; acc.w *= inv_cnt;
                    store_bin(ph * PW + pw, acc);
                }
        }
        if (OUT != 0) continue;                                // channels-last output went straight to global memory
        if (bulk_out) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // tile writes -> visible to the copy engine
        __syncthreads();

        for (int gi = 0; gi < nconc; ++gi) {                   // write out every tile pooled in this pass (all threads)
            const int sl_g = sl0 + gi;
            const int c0_g = (chunk * slabs + sl_g) * RV_CC;
            if (sl_g >= slabs || c0_g >= C) break;
            const float *tile_g = tile + gi * tile_floats;
            const int cc = min(RV_CC, C - c0_g);
            if (bulk_out) {
                // write-out by the copy engine: the tile rows j*32 .. j*32+31 (dense, tstride == PH*PW) hold channels 4*l + j of
                // the slab, which is exactly one box of the 4-D output map — four TMA tensor stores per slab (evict-first in
                // L2, like st.global.cs), issued by one thread: no LDS / STG instructions, no LSU wavefronts.  (128 per-row
                // cp.async.bulk copies were measured first: their per-lane issue loop alone drew 19 % of the stall samples.)
                // The generic-proxy tile writes were fenced before the barrier above; channels past C are clipped by the map.
                if (threadIdx.x == 0) {
                    uint64_t pol;
                    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t src = smem_u32(tile_g + (size_t)j * 32 * tstride);
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
                                     ::"l"(reinterpret_cast<uint64_t>(&omap)), "r"(src), "r"(0), "r"(j), "r"(c0_g >> 2), "r"(r), "l"(pol)
                                     : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile may be overwritten / the CTA may exit
                }
                continue;
            }
            // coalesced streaming write-out of the [cc, PH*PW] slab (contiguous in the NCHW output).  Thread = (channel
            // within a group of `cpi`, position); cpi is a multiple of 4 whenever possible so that stepping to the next
            // channel group is a constant stride in the (row-permuted) tile as well as in global memory.
            float *ob = out + ((size_t)r * C + c0_g) * PHW;
            const int words = PAIR ? (PHW >> 1) : PHW;           // 8-byte (PAIR) or 4-byte words per channel
            int cpi = RV_THREADS / words;
            if (cpi >= 4) cpi &= ~3;
            const int chs = threadIdx.x / words, wi = threadIdx.x - chs * words;
            if (chs < cpi) {
                if ((cpi & 3) == 0) {
                    const int rinc = (cpi >> 2) * tstride;
                    const float *sp = tile_g + (size_t)(((chs & 3) << 5) + (chs >> 2)) * tstride + (PAIR ? 2 * wi : wi);
                    float *gp = ob + (size_t)chs * PHW + (PAIR ? 2 * wi : wi);
                    const size_t ginc = (size_t)cpi * PHW;
#pragma unroll 4
                    for (int ch = chs; ch < cc; ch += cpi) {
                        if (PAIR) __stcs(reinterpret_cast<float2 *>(gp), *reinterpret_cast<const float2 *>(sp));
                        else __stcs(gp, *sp);
                        sp += rinc;
                        gp += ginc;
                    }
                } else {
                    for (int ch = chs; ch < cc; ch += cpi) {
                        const int row = ((ch & 3) << 5) + (ch >> 2);
                        if (PAIR) __stcs(reinterpret_cast<float2 *>(ob + (size_t)ch * PHW) + wi,
                                         *reinterpret_cast<const float2 *>(tile_g + (size_t)row * tstride + 2 * wi));
                        else __stcs(ob + (size_t)ch * PHW + wi, tile_g[(size_t)row * tstride + wi]);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward, vectorised: the adjoint of the column walk.  One CTA = (roi, slabs of 128 channels), warp = bin row, lane = 4
// consecutive channels of a CHANNELS-LAST gradient map (workspace, transposed to NCHW afterwards): the upstream
// gradient slab [128][PH*PW] is staged in shared memory (coalesced reads), a bin row scatters it along x into a sliding window
// of GW + 1 column accumulators in registers (dense per-bin weights, the same tables as the forward walk), and every
// finished column is added to the map with one 128-bit vector atomic per row tap (red.global.add.v4.f32: 512 contiguous
// bytes per warp) — against four scalar atomics per sample, bin and channel in the per-element kernel below
// (torchvision's scheme).  Rois the walk cannot take (fixed sampling ratio, GW > 3, oversized tables) are left to that
// kernel through the `handled` flags.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_quad(char *p, float w, const Quad &t) {
    atomicAdd(reinterpret_cast<float4 *>(p), make_float4(w * t.lo.x, w * t.lo.y, w * t.hi.x, w * t.hi.y));
}

struct WalkBin8 {
    float w[8];      // dense weights of columns base .. base + 7 (zero beyond GW): sampling grids up to 7 wide
};

template <int GW>
__device__ __noinline__ void walk_bin_row_bwd(char *gb, const MergedEntry *yt, int ny, const WalkBin8 *xw, const int *xadv, int x0_bytes,
                                               int colstride, int xlast_bytes, int PW, float inv_cnt, const float *trow, int rstep) {
    Quad acc[GW + 1];
#pragma unroll
    for (int k = 0; k <= GW; ++k) { acc[k].lo = make_float2(0.f, 0.f); acc[k].hi = acc[k].lo; }
    int cb = x0_bytes - (GW + 1) * colstride;           // byte offset of window column 0 (columns before x0 are phantoms)
    auto emit = [&](int col_bytes, const Quad &t) {
        if (col_bytes < x0_bytes || col_bytes > xlast_bytes) return;     // warp-uniform
        for (int k = 0; k < ny; ++k) red_add_quad(gb + yt[k].idx + col_bytes, yt[k].w, t);
    };
#pragma unroll 1
    for (int pw = 0; pw < PW; ++pw) {
        int a = xadv[pw];
#pragma unroll 1
        for (; a > 0; --a) {
            emit(cb, acc[0]);
#pragma unroll
            for (int k = 0; k < GW; ++k) acc[k] = acc[k + 1];
            acc[GW].lo = make_float2(0.f, 0.f);
            acc[GW].hi = acc[GW].lo;
            cb += colstride;
        }
        Quad g;
        g.lo = make_float2(trow[pw], trow[pw + rstep]);
        g.hi = make_float2(trow[pw + 2 * rstep], trow[pw + 3 * rstep]);
        g = quad_mul(inv_cnt, g);
        const float4 wa = *reinterpret_cast<const float4 *>(xw[pw].w);
        const float4 wb = *reinterpret_cast<const float4 *>(xw[pw].w + 4);
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int k = 0; k <= GW; ++k) quad_fma(acc[k], wv[k], g);
    }
#pragma unroll
    for (int k = 0; k <= GW; ++k) emit(cb + k * colstride, acc[k]);
}

__global__ void __launch_bounds__(RV_THREADS, 2) roi_align_bwd_v4_kernel(const float *__restrict__ dout, const float *__restrict__ rois, int C,
                                                                        int H, int W, int PH, int PW, float scale, int sampling_ratio,
                                                                        int aligned, int nchunks, int slabs, float *__restrict__ dfeat_nhwc,
                                                                        unsigned char *__restrict__ handled, const int *__restrict__ order) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MergedEntry *ytab = reinterpret_cast<MergedEntry *>(smem_raw);
    MergedEntry *xtab = ytab + RA_TAB;
    int *ycnt = reinterpret_cast<int *>(xtab + RA_TAB);
    int *xcnt = ycnt + 32;
    int *xadv = xcnt + 32;
    int *xbase = xadv + 32;
    int *walk_hdr = xbase + 32;
    WalkBin8 *xw = reinterpret_cast<WalkBin8 *>(walk_hdr + 4);
    float *tile = reinterpret_cast<float *>(xw + 32);         // [RV_CC][PH*PW | 1]: odd pitch, conflict-free lane-per-row reads

    const int slot = blockIdx.x / nchunks;
    const int chunk = blockIdx.x - slot * nchunks;
    const int r = order != nullptr ? order[slot] : slot;      // launch order: heaviest rois first (roi_order_kernel)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PHW = PH * PW, pitch = PHW | 1;

    const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, aligned, PH, PW, sampling_ratio);
    const bool empty = (g.gh <= 0 || g.gw <= 0);
    if (empty) {                                              // no sample: zero gradient
        if (chunk == 0 && threadIdx.x == 0) handled[r] = 1;
        return;
    }
    const int ystride = g.gh + 2, xstride = g.gw + 2;
    const bool walk_try = sampling_ratio <= 0 && g.gw >= 1 && g.gw <= 7 && PH <= RV_WARPS && (long long)PH * ystride <= RA_TAB &&
                          (long long)PW * xstride <= RA_TAB;
    if (!walk_try) return;                                    // CTA-uniform: left to the per-element kernel
    const float inv_cnt = __fdiv_rn(1.0f, (float)max(g.gh * g.gw, 1));

    bool walk_bad_a = false;
    {
        const int t = threadIdx.x;
        if (t < PH) ycnt[t] = build_merged(ytab + t * ystride, g.sh, g.bh, t, g.gh, H, W * C * 4);
        else if (t >= 32 && t < 32 + PW) {
            const int pw = t - 32;
            const MergedEntry *xt = xtab + pw * xstride;
            const int nx = build_merged(xtab + pw * xstride, g.sw, g.bw, pw, g.gw, W, C * 4, true);
            xcnt[pw] = nx;
            const int colbytes = C * 4;
            float w[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            int base = -1;
            if (nx > 0) {
                base = (xt[0].idx & ~1) / colbytes;
                for (int j = 0; j < nx; ++j) {
                    const int d = (xt[j].idx & ~1) / colbytes - base;
                    if (d < 0 || d > g.gw) walk_bad_a = true;
                    else w[d] += xt[j].w;
                }
            }
            xbase[pw] = base;
#pragma unroll
            for (int d = 0; d < 8; ++d) xw[pw].w[d] = w[d];
        }
    }
    if (threadIdx.x == 0) { walk_hdr[1] = 0; walk_hdr[2] = 0; }
    __syncthreads();
    bool walk_bad = false;
    if (threadIdx.x >= 32 && threadIdx.x < 32 + PW) {
        const int pw = threadIdx.x - 32;
        walk_bad = walk_bad_a;
        const int base = xbase[pw];
        int adv = 0;
        if (base >= 0) {
            int prev = -1;
            for (int q = pw - 1; q >= 0 && prev < 0; --q) prev = xbase[q];
            if (prev < 0) { adv = g.gw + 1; walk_hdr[1] = base; }
            else adv = base - prev;
            if (adv < 0 || adv > g.gw + 1) walk_bad = true;
            atomicMax(&walk_hdr[2], base + g.gw);
        }
        xadv[pw] = adv;
    }
    if (__syncthreads_and(walk_bad ? 0 : 1) == 0) return;     // CTA-uniform
    if (chunk == 0 && threadIdx.x == 0) handled[r] = 1;
    const int x0_bytes = walk_hdr[1] * C * 4, colstride = C * 4, xlast_bytes = min(W - 1, walk_hdr[2]) * C * 4;
    const int rstep = 32 * pitch;

    for (int sl = 0; sl < slabs; ++sl) {
        const int c0 = (chunk * slabs + sl) * RV_CC;
        if (c0 >= C) break;
        const int cc = min(RV_CC, C - c0);
        if (sl > 0) __syncthreads();                          // every warp is done with the previous slab's tile
        // stage dout[r, c0 : c0 + cc, :, :] (one contiguous run) as tile[(ch & 3) * 32 + (ch >> 2)][bin]
        const float *src = dout + ((size_t)r * C + c0) * PHW;
        if (PHW % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            // 128-bit loads, seven in flight per thread before the first is stored (the loop is latency bound otherwise:
            // 43 % of the kernel's stall samples sat on the store that waits for its own load)
            const int nvec = PHW >> 2, total = cc * nvec;
            const int dch = RV_THREADS / nvec, dq = RV_THREADS - dch * nvec;
            int ch = threadIdx.x / nvec, q = threadIdx.x - ch * nvec;
            for (int e0 = threadIdx.x; e0 < total; e0 += 7 * RV_THREADS) {
                float4 v[7];
#pragma unroll
                for (int u = 0; u < 7; ++u) {
                    const int e = e0 + u * RV_THREADS;
                    v[u] = (e < total) ? __ldg(reinterpret_cast<const float4 *>(src) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 7; ++u) {
                    if (e0 + u * RV_THREADS < total) {
                        float *d = tile + (size_t)(((ch & 3) << 5) + (ch >> 2)) * pitch + 4 * q;
                        d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
                    }
                    ch += dch; q += dq;
                    if (q >= nvec) { q -= nvec; ++ch; }
                }
            }
        } else {
            int ch = threadIdx.x / PHW, b = threadIdx.x - ch * PHW;
            const int dch = RV_THREADS / PHW, db = RV_THREADS - dch * PHW;
            for (int e = threadIdx.x; e < cc * PHW; e += RV_THREADS) {
                tile[(size_t)(((ch & 3) << 5) + (ch >> 2)) * pitch + b] = __ldg(src + e);
                ch += dch; b += db;
                if (b >= PHW) { b -= PHW; ++ch; }
            }
        }
        __syncthreads();
        const bool active = (c0 + 4 * lane) < C;
        if (warp < PH && active) {
            char *gb = reinterpret_cast<char *>(dfeat_nhwc + (size_t)g.batch * H * W * C + c0 + 4 * lane);
            const MergedEntry *yt = ytab + warp * ystride;
            const float *trow = tile + (size_t)lane * pitch + warp * PW;
            const int ny = ycnt[warp];
            if (ny > 0) {
                switch (g.gw) {                               // CTA-uniform
                    case 1: walk_bin_row_bwd<1>(gb, yt, ny, xw, xadv, x0_bytes, colstride, xlast_bytes, PW, inv_cnt, trow, rstep); break;
                    case 2: walk_bin_row_bwd<2>(gb, yt, ny, xw, xadv, x0_bytes, colstride, xlast_bytes, PW, inv_cnt, trow, rstep); break;
                    case 3: walk_bin_row_bwd<3>(gb, yt, ny, xw, xadv, x0_bytes, colstride, xlast_bytes, PW, inv_cnt, trow, rstep); break;
                    case 4: walk_bin_row_bwd<4>(gb, yt, ny, xw, xadv, x0_bytes, colstride, xlast_bytes, PW, inv_cnt, trow, rstep); break;
                    case 5: walk_bin_row_bwd<5>(gb, yt, ny, xw, xadv, x0_bytes, colstride, xlast_bytes, PW, inv_cnt, trow, rstep); break;
                    case 6: walk_bin_row_bwd<6>(gb, yt, ny, xw, xadv, x0_bytes, colstride, xlast_bytes, PW, inv_cnt, trow, rstep); break;
                    default: walk_bin_row_bwd<7>(gb, yt, ny, xw, xadv, x0_bytes, colstride, xlast_bytes, PW, inv_cnt, trow, rstep); break;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward (atomic scatter; one thread per (r, c, ph, pw))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) roi_align_bwd_kernel(const float *__restrict__ dout, const float *__restrict__ rois,
                                                            int C, int H, int W, int PH, int PW, float scale,
                                                            int sampling_ratio, int aligned, size_t total,
                                                            float *__restrict__ dfeat) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int pw = (int)(idx % PW);
        const int ph = (int)((idx / PW) % PH);
        const int c = (int)((idx / ((size_t)PW * PH)) % C);
        const int r = (int)(idx / ((size_t)PW * PH * C));
        const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, aligned, PH, PW, sampling_ratio);
        if (g.gh <= 0 || g.gw <= 0) continue;
        const float gv = dout[idx] / (float)max(g.gh * g.gw, 1);
        float *plane = dfeat + ((size_t)g.batch * C + c) * H * W;
        for (int iy = 0; iy < g.gh; ++iy) {
            const Tap ty = make_tap(sample_coord(g.sh, g.bh, ph, iy, g.gh), H);
            if (ty.lo < 0) continue;
            for (int ix = 0; ix < g.gw; ++ix) {
                const Tap tx = make_tap(sample_coord(g.sw, g.bw, pw, ix, g.gw), W);
                if (tx.lo < 0) continue;
                atomicAdd(plane + ty.lo * W + tx.lo, gv * ty.wl * tx.wl);
                atomicAdd(plane + ty.lo * W + tx.hi, gv * ty.wl * tx.wh);
                atomicAdd(plane + ty.hi * W + tx.lo, gv * ty.wh * tx.wl);
                atomicAdd(plane + ty.hi * W + tx.hi, gv * ty.wh * tx.wh);
            }
        }
    }
}

// the rois the vectorised backward declined (one CTA per roi; a handled roi costs one flag load)
__global__ void __launch_bounds__(256) roi_align_bwd_rest_kernel(const float *__restrict__ dout, const float *__restrict__ rois, int C, int H,
                                                                 int W, int PH, int PW, float scale, int sampling_ratio, int aligned,
                                                                 float *__restrict__ dfeat, const unsigned char *__restrict__ handled) {
    const int r = blockIdx.x;
    if (handled[r]) return;
    const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, aligned, PH, PW, sampling_ratio);
    if (g.gh <= 0 || g.gw <= 0) return;
    const int PHW = PH * PW;
    const float inv = 1.0f / (float)max(g.gh * g.gw, 1);
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < C * PHW; e += gridDim.y * blockDim.x) {
        const int c = e / PHW, b = e - c * PHW;
        const int ph = b / PW, pw = b - ph * PW;
        const float gv = dout[(size_t)r * C * PHW + e] * inv;
        float *plane = dfeat + ((size_t)g.batch * C + c) * H * W;
        for (int iy = 0; iy < g.gh; ++iy) {
            const Tap ty = make_tap(sample_coord(g.sh, g.bh, ph, iy, g.gh), H);
            if (ty.lo < 0) continue;
            for (int ix = 0; ix < g.gw; ++ix) {
                const Tap tx = make_tap(sample_coord(g.sw, g.bw, pw, ix, g.gw), W);
                if (tx.lo < 0) continue;
                atomicAdd(plane + ty.lo * W + tx.lo, gv * ty.wl * tx.wl);
                atomicAdd(plane + ty.lo * W + tx.hi, gv * ty.wl * tx.wh);
                atomicAdd(plane + ty.hi * W + tx.lo, gv * ty.wh * tx.wl);
                atomicAdd(plane + ty.hi * W + tx.hi, gv * ty.wh * tx.wh);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// sampling-grid dump (parity tests: coordinates + integer indices must be bit-exact)
// ------------------------------------------------------------------------------------------------
__global__ void roi_align_grid_kernel(const float *__restrict__ rois, int R, int H, int W, int PH, int PW, float scale,
                                      int sampling_ratio, int aligned, int max_grid, int32_t *__restrict__ grid_hw,
                                      float *__restrict__ yx, int32_t *__restrict__ idx) {
    const int r = blockIdx.x;
    if (r >= R) return;
    const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, aligned, PH, PW, sampling_ratio);
    if (threadIdx.x == 0) {
        grid_hw[2 * r] = g.gh;
        grid_hw[2 * r + 1] = g.gw;
    }
    const int per_roi = PH * PW * max_grid * max_grid;
    for (int s = threadIdx.x; s < per_roi; s += blockDim.x) {
        const int ix = s % max_grid;
        const int iy = (s / max_grid) % max_grid;
        const int pw = (s / (max_grid * max_grid)) % PW;
        const int ph = s / (max_grid * max_grid * PW);
        if (iy >= g.gh || ix >= g.gw) continue;
        const float y = sample_coord(g.sh, g.bh, ph, iy, g.gh);
        const float x = sample_coord(g.sw, g.bw, pw, ix, g.gw);
        const size_t o = (size_t)r * per_roi + s;
        yx[2 * o] = y;
        yx[2 * o + 1] = x;
        const bool skip = (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
        const Tap ty = make_tap(y, H), tx = make_tap(x, W);
        idx[4 * o + 0] = skip ? -1 : ty.lo;
        idx[4 * o + 1] = skip ? -1 : tx.lo;
        idx[4 * o + 2] = skip ? -1 : ty.hi;
        idx[4 * o + 3] = skip ? -1 : tx.hi;
    }
}


// ------------------------------------------------------------------------------------------------------------------------------
// Spatial mean of the res5 output — the reference's ``box_features.mean(dim=[2, 3])`` (roi_emb_heads.py:262, :329, :351) — which is the
// largest read between RoIAlign and the predictor: [R, 2048, 7, 7] is 3.2 GB at R = 8000.  One pass, fp32 accumulation in a fixed
// order, and the bf16 (hi, lo) operand of the projection GEMM written by the same pass (what would otherwise be a separate split
// kernel re-reading the [R, C] means).
//   CL = true  : x is channels-last [R, HW, C] (fp32 or bf16): a thread owns 4 channels, walks HW; a warp reads 512 / 256 contiguous
//                bytes per position.
//   CL = false : x is [R, C, HW] fp32: a warp owns 32 consecutive (r, c) rows, 4 rows in flight at a time (8 loads per lane), so
//                the 32 means leave as one 128-byte store.
template <typename T>
__device__ __forceinline__ float4 load_ch4(const T *p);
template <>
__device__ __forceinline__ float4 load_ch4<float>(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }
template <>
__device__ __forceinline__ float4 load_ch4<uint16_t>(const uint16_t *p) {
    const uint2 v = __ldcs(reinterpret_cast<const uint2 *>(p));
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
}

__device__ __forceinline__ void mean_store(float m, int64_t r, int c, float *out, int64_t ldo, uint16_t *hi, uint16_t *lo, int64_t ldh) {
    out[r * ldo + c] = m;
    if (hi != nullptr) {
        uint16_t h, l;
        split_bf16(m, h, l);
        hi[r * ldh + c] = h;
        if (lo != nullptr) lo[r * ldh + c] = l;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) spatial_mean_cl_kernel(const T *__restrict__ x, int64_t R, int C, int HW, float inv, float *__restrict__ out,
                                                              int64_t ldo, uint16_t *__restrict__ hi, uint16_t *__restrict__ lo, int64_t ldh) {
    pdl_trigger();
    pdl_wait();
    const int c4n = C / 4;
    const int64_t total = R * c4n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / c4n;
        const int c = (int)(i - r * c4n) * 4;
        const T *p = x + (r * HW) * C + c;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        int s = 0;
        for (; s + 7 <= HW; s += 7) {                    // 7 independent loads in flight (HW = 49 -> 7 rounds)
            float4 v[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) v[k] = load_ch4<T>(p + (int64_t)(s + k) * C);
#pragma unroll
            for (int k = 0; k < 7; ++k) { a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w; }
        }
        for (; s < HW; ++s) {
            const float4 v = load_ch4<T>(p + (int64_t)s * C);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        const float4 m = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
        *reinterpret_cast<float4 *>(out + r * ldo + c) = m;
        if (hi != nullptr) {
            uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
            split_bf16(m.x, h0, l0); split_bf16(m.y, h1, l1); split_bf16(m.z, h2, l2); split_bf16(m.w, h3, l3);
            *reinterpret_cast<uint2 *>(hi + r * ldh + c) = make_uint2(h0 | ((uint32_t)h1 << 16), h2 | ((uint32_t)h3 << 16));
            if (lo != nullptr) *reinterpret_cast<uint2 *>(lo + r * ldh + c) = make_uint2(l0 | ((uint32_t)l1 << 16), l2 | ((uint32_t)l3 << 16));
        }
    }
}

// One warp owns 32 consecutive (r, c) rows = one contiguous span of 32 * HW floats: the span is copied to shared memory with 16-byte
// cp.async (no registers, everything in flight at once — 6.3 KB per warp at HW = 49), then lane j adds up row j from shared memory
// in a fixed order (row stride HW | 1 floats: odd, so the 32 lanes hit 32 different banks) and the 32 means leave as one 128-byte store.
constexpr int SM_WARPS = 4;
__global__ void __launch_bounds__(SM_WARPS * 32) spatial_mean_rows_kernel(const float *__restrict__ x, int64_t rows, int C, int HW, float inv,
                                                                          float *__restrict__ out, int64_t ldo, uint16_t *__restrict__ hi,
                                                                          uint16_t *__restrict__ lo, int64_t ldh) {
    extern __shared__ __align__(16) float sm_rows[];
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int stride = HW | 1;
    float *buf = sm_rows + (size_t)wib * 32 * stride;
    const uint32_t sbuf = smem_u32(buf);
    const int64_t warp = (int64_t)blockIdx.x * SM_WARPS + wib;
    const int64_t nwarps = (int64_t)gridDim.x * SM_WARPS;
    const bool linear = (HW & 1) != 0;                  // stride == HW: the span is copied as it lies
    for (int64_t base = warp * 32; base < rows; base += nwarps * 32) {
        const int nrow = (int)min((int64_t)32, rows - base);
        const int n = nrow * HW;
        const float *src = x + base * HW;
        if (linear && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const int n4 = n >> 2;
            for (int i = lane; i < n4; i += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbuf + i * 16), "l"(src + i * 4) : "memory");
            for (int i = (n4 << 2) + lane; i < n; i += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sbuf + i * 4), "l"(src + i) : "memory");
        } else {
            for (int i = lane; i < n; i += 32) {
                const int row = i / HW;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sbuf + (row * stride + (i - row * HW)) * 4), "l"(src + i) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (lane < nrow) {
            const float *p = buf + lane * stride;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;           // four chains, combined in a fixed order
            int t = 0;
            for (; t + 4 <= HW; t += 4) { a0 += p[t]; a1 += p[t + 1]; a2 += p[t + 2]; a3 += p[t + 3]; }
            for (; t < HW; ++t) a0 += p[t];
            const int64_t row = base + lane;
            mean_store(((a0 + a1) + (a2 + a3)) * inv, row / C, (int)(row % C), out, ldo, hi, lo, ldh);
        }
        __syncwarp();                                        // the buffer is refilled by the next round
    }
}

// fall-back for very large maps (the span of 32 rows does not fit shared memory): lanes stride over the row, warp reduction
__global__ void __launch_bounds__(256) spatial_mean_rows_big_kernel(const float *__restrict__ x, int64_t rows, int C, int HW, float inv, float *__restrict__ out,
                                                                    int64_t ldo, uint16_t *__restrict__ hi, uint16_t *__restrict__ lo, int64_t ldh) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < rows; row += nwarps) {
        float a = 0.f;
        for (int s = lane; s < HW; s += 32) a += __ldcs(x + row * HW + s);
        a = warp_sum(a);
        if (lane == 0) mean_store(a * inv, row / C, (int)(row % C), out, ldo, hi, lo, ldh);
    }
}

// gradient of the mean: dx[r, c, s] = dy[r, c] / HW in the layout / dtype of x (each thread writes 4 channels x one position, or 4 positions)
template <typename T, bool CL>
__global__ void __launch_bounds__(256) spatial_mean_bwd_kernel(const float *__restrict__ dy, int64_t lddy, int64_t R, int C, int HW, float inv, T *__restrict__ dx) {
    pdl_trigger();
    pdl_wait();
    const int64_t total = R * (int64_t)C * HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r;
        int c;
        if (CL) { c = (int)(i % C); r = i / ((int64_t)C * HW); }
        else    { const int64_t rc = i / HW; r = rc / C; c = (int)(rc - r * C); }
        const float g = dy[r * lddy + c] * inv;
        if (sizeof(T) == 4) reinterpret_cast<float *>(dx)[i] = g;
        else reinterpret_cast<uint16_t *>(dx)[i] = f32_to_bf16_rn(g);
    }
}

}  // namespace loco

using namespace loco;

extern "C" {

static int64_t roi_order_bytes(int R) { return R > 0 ? (((int64_t)R * 4 + 255) / 256) * 256 : 0; }

// [transposed map (NCHW input only)] [launch order: R ints]
int64_t loco_roi_align_workspace_bytes(int N, int C, int H, int W, int feat_layout, int R) {
    const int64_t map = feat_layout == LOCO_NHWC ? 0 : (((int64_t)N * C * H * W * (int64_t)sizeof(float) + 255) / 256) * 256;
    return map + roi_order_bytes(R);
}

int loco_roi_align_fwd(const float *feat, int N, int C, int H, int W, int feat_layout, const float *rois, int R,
                       int PH, int PW, float spatial_scale, int sampling_ratio, int aligned, void *out_v, int out_layout, int out_dtype,
                       void *workspace, void *stream) {
    float *out = static_cast<float *>(out_v);
    LOCO_REQUIRE((out_layout == LOCO_NCHW && out_dtype == LOCO_F32) || (out_layout == LOCO_NHWC && (out_dtype == LOCO_F32 || out_dtype == LOCO_BF16)),
                 LOCO_E_UNSUPPORTED, "roi_align_fwd: output layout %d / dtype %d (NCHW fp32, or channels-last fp32 / bf16)", out_layout, out_dtype);
    LOCO_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && PH > 0 && PW > 0 && R >= 0, LOCO_E_BADARG,
                 "roi_align_fwd: bad shape N=%d C=%d H=%d W=%d PH=%d PW=%d R=%d", N, C, H, W, PH, PW, R);
    LOCO_REQUIRE(feat_layout == LOCO_NCHW || feat_layout == LOCO_NHWC, LOCO_E_BADARG, "roi_align_fwd: bad layout %d", feat_layout);
    if (R == 0) return LOCO_OK;
    LOCO_REQUIRE(feat && rois && out, LOCO_E_BADARG, "roi_align_fwd: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float *nhwc = feat;
    if (feat_layout == LOCO_NCHW) {
        LOCO_REQUIRE(workspace != nullptr, LOCO_E_BADARG, "roi_align_fwd: NCHW features need loco_roi_align_workspace_bytes() of workspace");
        const int HW = H * W;
        dim3 grid((HW + 31) / 32, (C + 31) / 32, N);
        LOCO_REQUIRE(grid.y <= 65535 && grid.z <= 65535, LOCO_E_UNSUPPORTED, "roi_align_fwd: feature map too large for the transpose grid");
        nchw_to_nhwc_kernel<<<grid, 256, 0, st>>>(feat, static_cast<float *>(workspace), C, HW);
    count_launch();
        LOCO_CUDA(cudaGetLastError());
        nhwc = static_cast<const float *>(workspace);
    }
    LOCO_REQUIRE(PH <= 32 && PW <= 32, LOCO_E_UNSUPPORTED, "roi_align_fwd: output size %dx%d (at most 32x32)", PH, PW);
    const int PHW = PH * PW;
    // vectorised kernel: 4 channels per lane (128-bit taps), byte offsets inside one image kept in 32 bits
    const bool v4 = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(nhwc) & 15) == 0) && ((long long)H * W * C * 4 < (1ll << 31)) &&
                    PHW <= RV_THREADS;
    const int out_mode = out_layout == LOCO_NCHW ? 0 : (out_dtype == LOCO_F32 ? 1 : 2);
    LOCO_REQUIRE(out_mode == 0 || (v4 && (reinterpret_cast<uintptr_t>(out) & 15) == 0), LOCO_E_UNSUPPORTED,
                 "roi_align_fwd: channels-last output needs C %% 4 == 0 and 16-byte aligned feature / output pointers");
    if (v4) {
        const bool pair = out_mode == 0 && (PW % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
        // copy-engine write-out needs dense 16-byte aligned channel rows (PH*PW % 4 == 0: 14x14 yes, 7x7 no); its tile
        // stride PH*PW = 0 (mod 4) makes the paired 8-byte tile stores 2-way bank conflicted, which costs less than the
        // LDS + STG write-out loop it removes (LOCOV_B200_ROI_BULK=0 restores the loop for A/B measurements)
        static const bool bulk_allowed = []() { const char *e = getenv("LOCOV_B200_ROI_BULK"); return !(e != nullptr && e[0] == '0'); }();
        const int bulk_out = (pair && bulk_allowed && PHW % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
        int tstride;
        if (bulk_out) tstride = PHW;
        else if (pair) { tstride = PHW; while (tstride % 4 != 2) ++tstride; }     // S = 2 (mod 4): conflict-free STS.64 columns
        else tstride = PHW | 1;
        const int nslab = (C + RV_CC - 1) / RV_CC;
        // each CTA pools `slabs` consecutive slabs with one set of tap tables (the per-CTA prologue — geometry, tap tables, walk program —
        // is ~30 % of the samples at 2 slabs), as long as the grid keeps >= 8 waves; with the launch order below (no tail of late heavy
        // rois) the channels-last modes afford 6 waves: 4 slabs per CTA at the benchmark shape (225 -> 217 us)
        int slabs = 1;
        static const bool order_allowed = []() { const char *e = getenv("LOCOV_B200_ROI_ORDER"); return !(e != nullptr && e[0] == '0'); }();
        const bool will_order = order_allowed && workspace != nullptr;
        const long long min_ctas = (will_order && out_mode != 0 ? 6ll : 8ll) * 2 * 148;
        while (slabs < 4 && nslab % (slabs * 2) == 0 && (long long)R * (nslab / (slabs * 2)) >= min_ctas) slabs *= 2;
        static const int slabs_env = []() { const char *e = getenv("LOCOV_B200_ROI_SLABS"); return e != nullptr ? atoi(e) : 0; }();   // developer sweep knob
        if (slabs_env > 0 && nslab % slabs_env == 0) slabs = slabs_env;
        const int nconc = RV_WARPS / PH > 1 ? RV_WARPS / PH : 1;            // slabs pooled at the same time (see the kernel)
        while (slabs < nconc && nslab % (slabs * 2) == 0) slabs *= 2;       // keep every warp group busy
        const int nchunks = nslab / slabs;
        const size_t smem = (size_t)RV_TILE_OFFSET + (out_mode == 0 ? (size_t)nconc * RV_CC * tstride * sizeof(float) : 0);   // no tile for channels-last
        CUtensorMap omap;
        memset(&omap, 0, sizeof(omap));
        if (bulk_out) {
            const int rc = make_tmap_roi_out(&omap, out, (uint64_t)R, (uint64_t)C, (uint64_t)PHW);
            if (rc != LOCO_OK) return rc;
        }
        LOCO_REQUIRE(smem <= 200 * 1024, LOCO_E_UNSUPPORTED, "roi_align_fwd: output size %dx%d needs %zu B of shared memory", PH, PW, smem);
        LOCO_REQUIRE((long long)R * nchunks < (1ll << 31), LOCO_E_UNSUPPORTED, "roi_align_fwd: too many (roi, channel-slab) tiles");
        typedef void (*v4_fn)(const float *, const float *, int, int, int, int, int, float, int, int, int, int, int, int, float *, const CUtensorMap, const int *);
        const bool multi = nconc > 1;
        v4_fn fn;
        if (out_mode == 1) fn = multi ? roi_align_fwd_v4_kernel<false, true, 1> : roi_align_fwd_v4_kernel<false, false, 1>;
        else if (out_mode == 2) fn = multi ? roi_align_fwd_v4_kernel<false, true, 2> : roi_align_fwd_v4_kernel<false, false, 2>;
        else fn = pair ? (multi ? roi_align_fwd_v4_kernel<true, true, 0> : roi_align_fwd_v4_kernel<true, false, 0>)
                       : (multi ? roi_align_fwd_v4_kernel<false, true, 0> : roi_align_fwd_v4_kernel<false, false, 0>);
        static thread_local size_t smem_set[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int vi = out_mode == 0 ? (pair ? 2 : 0) + (multi ? 1 : 0) : 4 + 2 * (out_mode - 1) + (multi ? 1 : 0);
        if (smem > 48 * 1024 && smem > smem_set[vi]) {
            LOCO_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_set[vi] = smem;
        }
        // launch order (heaviest rois first) when the grid is more than a few waves and the caller gave workspace
        const int *order = nullptr;
        if (will_order && (long long)R * nchunks > 4ll * current_device_sm_count()) {
            const int64_t map_bytes = feat_layout == LOCO_NHWC ? 0 : (((int64_t)N * C * H * W * (int64_t)sizeof(float) + 255) / 256) * 256;
            int *ord = reinterpret_cast<int *>(static_cast<unsigned char *>(workspace) + map_bytes);
            LOCO_CUDA(launch_kernel(roi_order_kernel, dim3(1), dim3(1024), 0, st, 1, rois, R, spatial_scale, ord));
            count_launch();
            order = ord;
        }
        fn<<<R * nchunks, RV_THREADS, smem, st>>>(nhwc, rois, C, H, W, PH, PW, spatial_scale, sampling_ratio, aligned, nchunks, slabs, tstride,
                                                  bulk_out, out, omap, order);
        count_launch();
        LOCO_CUDA(cudaGetLastError());
        return LOCO_OK;
    }
    // generic kernel: any channel count / alignment, one channel per lane
    const int nchunks = (C + RA_CC - 1) / RA_CC;
    const size_t smem = 2 * RA_TAB * sizeof(MergedEntry) + (size_t)(PH + PW) * sizeof(int) + (size_t)RA_CC * ((PH * PW) | 1) * sizeof(float) +
                        (size_t)RA_WARPS * RA_UCOLS * 32 * sizeof(float);
    LOCO_REQUIRE(smem <= 200 * 1024, LOCO_E_UNSUPPORTED, "roi_align_fwd: output size %dx%d needs %zu B of shared memory", PH, PW, smem);
    LOCO_REQUIRE((long long)R * nchunks < (1ll << 31), LOCO_E_UNSUPPORTED, "roi_align_fwd: too many (roi, channel-slab) tiles");
    static thread_local size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        LOCO_CUDA(cudaFuncSetAttribute(roi_align_fwd_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    roi_align_fwd_generic_kernel<<<R * nchunks, RA_THREADS, smem, st>>>(nhwc, rois, C, H, W, PH, PW, spatial_scale, sampling_ratio,
                                                                aligned, nchunks, out);
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int64_t loco_roi_align_bwd_workspace_bytes(int N, int C, int H, int W, int R) {
    if (C % 4 != 0) return 0;                                 // the vectorised path needs 4-channel lanes
    // [channels-last accumulation map] [handled flags: R bytes] [launch order: R ints]
    return (((int64_t)N * C * H * W * (int64_t)sizeof(float) + 255) / 256) * 256 + (((int64_t)R + 255) / 256) * 256 + roi_order_bytes(R);
}

int loco_roi_align_bwd(const float *dout, int N, int C, int H, int W, const float *rois, int R, int PH, int PW,
                       float spatial_scale, int sampling_ratio, int aligned, float *dfeat, void *workspace, void *stream) {
    LOCO_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && PH > 0 && PW > 0 && R >= 0, LOCO_E_BADARG, "roi_align_bwd: bad shape");
    if (R == 0) return LOCO_OK;
    LOCO_REQUIRE(dout && rois && dfeat, LOCO_E_BADARG, "roi_align_bwd: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t total = (size_t)R * C * PH * PW;
    const int blocks = (int)min((size_t)148 * 32, (total + 255) / 256);
    const int PHW = PH * PW;
    const size_t smem = 2 * RA_TAB * sizeof(MergedEntry) + (128 + 4) * sizeof(int) + 32 * sizeof(WalkBin8) + (size_t)RV_CC * (PHW | 1) * sizeof(float);
    static const bool v4_allowed = []() { const char *e = getenv("LOCOV_B200_ROI_BWD_V4"); return !(e != nullptr && e[0] == '0'); }();
    const bool v4 = v4_allowed && workspace != nullptr && C % 4 == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && PH <= 32 && PW <= 32 &&
                    (long long)H * W * C * 4 < (1ll << 31) && smem <= 110 * 1024 && (long long)R * ((C + RV_CC - 1) / RV_CC) < (1ll << 31) &&
                    N <= 65535;
    if (!v4) {       // per-element scatter straight into the NCHW gradient (zero-filled by the caller)
        roi_align_bwd_kernel<<<blocks, 256, 0, st>>>(dout, rois, C, H, W, PH, PW, spatial_scale, sampling_ratio, aligned, total, dfeat);
        count_launch();
        LOCO_CUDA(cudaGetLastError());
        return LOCO_OK;
    }
    // channels-last accumulation map + per-roi "handled" flags live in the workspace
    float *ws = static_cast<float *>(workspace);
    const size_t map_bytes = (((size_t)N * C * H * W * sizeof(float) + 255) / 256) * 256;
    unsigned char *handled = static_cast<unsigned char *>(workspace) + map_bytes;
    LOCO_CUDA(cudaMemsetAsync(workspace, 0, map_bytes + (size_t)R, st));
    const int nslab = (C + RV_CC - 1) / RV_CC;
    int slabs = 1;
    while (slabs < 4 && nslab % (slabs * 2) == 0 && (long long)R * (nslab / (slabs * 2)) >= 8ll * 2 * 148) slabs *= 2;
    const int nchunks = nslab / slabs;
    static thread_local size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        LOCO_CUDA(cudaFuncSetAttribute(roi_align_bwd_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const int *order = nullptr;
    static const bool order_allowed = []() { const char *e = getenv("LOCOV_B200_ROI_ORDER"); return !(e != nullptr && e[0] == '0'); }();
    if (order_allowed && (long long)R * nchunks > 4ll * current_device_sm_count()) {
        int *ord = reinterpret_cast<int *>(handled + (((size_t)R + 255) / 256) * 256);
        LOCO_CUDA(launch_kernel(roi_order_kernel, dim3(1), dim3(1024), 0, st, 1, rois, R, spatial_scale, ord));
        count_launch();
        order = ord;
    }
    roi_align_bwd_v4_kernel<<<R * nchunks, RV_THREADS, smem, st>>>(dout, rois, C, H, W, PH, PW, spatial_scale, sampling_ratio, aligned, nchunks, slabs,
                                                                   ws, handled, order);
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    // [N][HW][C] -> [N][C][HW]: the forward's transpose with the roles of the two axes swapped (this ASSIGNS dfeat)
    {
        const int HW = H * W;
        dim3 grid((C + 31) / 32, (HW + 31) / 32, N);
        LOCO_REQUIRE(grid.y <= 65535, LOCO_E_UNSUPPORTED, "roi_align_bwd: feature map too large for the transpose grid");
        nchw_to_nhwc_kernel<<<grid, 256, 0, st>>>(ws, dfeat, HW, C);
        count_launch();
        LOCO_CUDA(cudaGetLastError());
    }
    // whatever the vectorised kernel declined
    roi_align_bwd_rest_kernel<<<dim3((unsigned)R, 32), 256, 0, st>>>(dout, rois, C, H, W, PH, PW, spatial_scale, sampling_ratio, aligned, dfeat, handled);
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_roi_align_grid_dump(const float *rois, int R, int H, int W, int PH, int PW, float spatial_scale,
                             int sampling_ratio, int aligned, int max_grid, int32_t *grid_hw, float *yx, int32_t *idx,
                             void *stream) {
    LOCO_REQUIRE(R >= 0 && H > 0 && W > 0 && PH > 0 && PW > 0 && max_grid > 0, LOCO_E_BADARG, "roi_align_grid_dump: bad shape");
    if (R == 0) return LOCO_OK;
    LOCO_REQUIRE(rois && grid_hw && yx && idx, LOCO_E_BADARG, "roi_align_grid_dump: null pointer");
    roi_align_grid_kernel<<<R, 256, 0, static_cast<cudaStream_t>(stream)>>>(rois, R, H, W, PH, PW, spatial_scale,
                                                                           sampling_ratio, aligned, max_grid, grid_hw, yx, idx);
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_spatial_mean(const void *x, int64_t R, int C, int HW, int layout, int dtype, float *out, int64_t ldo, uint16_t *hi, uint16_t *lo, int64_t ldh,
                      void *stream) {
    LOCO_REQUIRE(R >= 0 && C >= 1 && HW >= 1 && ldo >= C, LOCO_E_BADARG, "spatial_mean: bad shape R=%lld C=%d HW=%d", (long long)R, C, HW);
    LOCO_REQUIRE((layout == LOCO_NCHW && dtype == LOCO_F32) || (layout == LOCO_NHWC && (dtype == LOCO_F32 || dtype == LOCO_BF16)), LOCO_E_UNSUPPORTED,
                 "spatial_mean: layout %d / dtype %d (NCHW fp32, or channels-last fp32 / bf16)", layout, dtype);
    LOCO_REQUIRE(lo == nullptr || hi != nullptr, LOCO_E_BADARG, "spatial_mean: lo without hi");
    LOCO_REQUIRE(hi == nullptr || ldh >= C, LOCO_E_BADARG, "spatial_mean: operand row stride %lld < C", (long long)ldh);
    if (R == 0) return LOCO_OK;
    LOCO_REQUIRE(x && out, LOCO_E_BADARG, "spatial_mean: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float inv = 1.0f / (float)HW;
    const int sms = current_device_sm_count();
    if (layout == LOCO_NHWC) {
        const int esz = dtype == LOCO_F32 ? 4 : 2;
        LOCO_REQUIRE(C % 4 == 0 && (reinterpret_cast<uintptr_t>(x) % (4 * esz)) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && ldo % 4 == 0 &&
                         (hi == nullptr || (((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 7) == 0 && ldh % 4 == 0)),
                     LOCO_E_ALIGN, "spatial_mean: channels-last input needs C %% 4 == 0 and 16-byte aligned rows");
        const int64_t total = R * (C / 4);
        const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sms * 16);
        if (dtype == LOCO_F32)
            LOCO_CUDA(launch_kernel(spatial_mean_cl_kernel<float>, dim3(blocks), dim3(256), 0, st, 1, static_cast<const float *>(x), R, C, HW, inv, out, ldo,
                                    hi, lo, ldh));
        else
            LOCO_CUDA(launch_kernel(spatial_mean_cl_kernel<uint16_t>, dim3(blocks), dim3(256), 0, st, 1, static_cast<const uint16_t *>(x), R, C, HW, inv, out,
                                    ldo, hi, lo, ldh));
    } else {
        const int64_t rows = R * C;
        const size_t smem = (size_t)SM_WARPS * 32 * (HW | 1) * sizeof(float);
        if (smem <= 96 * 1024) {
            static thread_local size_t smem_set = 0;
            if (smem > 48 * 1024 && smem > smem_set) {
                LOCO_CUDA(cudaFuncSetAttribute(spatial_mean_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                smem_set = smem;
            }
            const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (200 * 1024) / smem));
            const int blocks = (int)std::min<int64_t>((rows + 32 * SM_WARPS - 1) / (32 * SM_WARPS), (int64_t)sms * per_sm);
            LOCO_CUDA(launch_kernel(spatial_mean_rows_kernel, dim3(blocks), dim3(SM_WARPS * 32), smem, st, 1, static_cast<const float *>(x), rows, C, HW, inv,
                                    out, ldo, hi, lo, ldh));
        } else {
            const int blocks = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)sms * 8);
            LOCO_CUDA(launch_kernel(spatial_mean_rows_big_kernel, dim3(blocks), dim3(256), 0, st, 1, static_cast<const float *>(x), rows, C, HW, inv, out, ldo,
                                    hi, lo, ldh));
        }
    }
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

int loco_spatial_mean_bwd(const float *dy, int64_t lddy, int64_t R, int C, int HW, int layout, int dtype, void *dx, void *stream) {
    LOCO_REQUIRE(R >= 0 && C >= 1 && HW >= 1 && lddy >= C, LOCO_E_BADARG, "spatial_mean_bwd: bad shape R=%lld C=%d HW=%d", (long long)R, C, HW);
    LOCO_REQUIRE((layout == LOCO_NCHW && dtype == LOCO_F32) || (layout == LOCO_NHWC && (dtype == LOCO_F32 || dtype == LOCO_BF16)), LOCO_E_UNSUPPORTED,
                 "spatial_mean_bwd: layout %d / dtype %d", layout, dtype);
    if (R == 0) return LOCO_OK;
    LOCO_REQUIRE(dy && dx, LOCO_E_BADARG, "spatial_mean_bwd: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float inv = 1.0f / (float)HW;
    const int64_t total = R * (int64_t)C * HW;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)current_device_sm_count() * 32);
    if (layout == LOCO_NCHW)
        LOCO_CUDA(launch_kernel(spatial_mean_bwd_kernel<float, false>, dim3(blocks), dim3(256), 0, st, 1, dy, lddy, R, C, HW, inv, static_cast<float *>(dx)));
    else if (dtype == LOCO_F32)
        LOCO_CUDA(launch_kernel(spatial_mean_bwd_kernel<float, true>, dim3(blocks), dim3(256), 0, st, 1, dy, lddy, R, C, HW, inv, static_cast<float *>(dx)));
    else
        LOCO_CUDA(launch_kernel(spatial_mean_bwd_kernel<uint16_t, true>, dim3(blocks), dim3(256), 0, st, 1, dy, lddy, R, C, HW, inv, static_cast<uint16_t *>(dx)));
    count_launch();
    LOCO_CUDA(cudaGetLastError());
    return LOCO_OK;
}

}  // extern "C"

