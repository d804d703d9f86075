/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called from the product
 * path (locov_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object.
 *
 * Scalar float32 restatement of RoIAlign as the reference reaches it:
 *   /root/reference/ovr/modeling/roi_heads/roi_emb_heads.py:182-187  (ROIPooler(14, (1/16,), 0, "ROIAlignV2"))
 *   /root/reference/ovr/modeling/roi_heads/roi_emb_heads.py:243-245  (_shared_roi_transform -> self.pooler)
 * The arithmetic itself lives in a third-party dependency that is NOT under /root/reference:
 * Detectron2 (un-vendored, version unpinned, README.md:28) -> detectron2.layers.ROIAlign ->
 * torchvision.ops.roi_align(input, rois[R,5], output_size, spatial_scale, sampling_ratio, aligned=True)
 * (torchvision unpinned; 0.26.0 is the build in this image).  This file restates torchvision's
 * published CPU algorithm (pre-computed bilinear taps, one separately-rounded float32 op per step,
 * SURVEY.md Appendix A) and is PINNED in tests/test_oracle_roi_align.py bit-for-bit against the
 * compiled CPU op torch.ops.torchvision.roi_align on in-range, out-of-range, clamped and
 * degenerate boxes.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off so that no multiply-add is fused).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

typedef struct {
    float y, x;            /* sample coordinate BEFORE clamping (what "sampling-grid coordinates" means) */
    int32_t y_low, x_low, y_high, x_high; /* indices after clamping; -1 when the sample is skipped */
    float w1, w2, w3, w4;  /* bilinear weights (0 when skipped) */
} tap_t;

static inline void roi_geometry(const float *roi, float scale, int aligned, int PH, int PW,
                                int sampling_ratio, float *sh, float *sw, float *bh, float *bw,
                                int *gh, int *gw)
{
    const float off = aligned ? 0.5f : 0.0f;
    const float start_w = roi[1] * scale - off;
    const float start_h = roi[2] * scale - off;
    const float end_w = roi[3] * scale - off;
    const float end_h = roi[4] * scale - off;
    float rw = end_w - start_w;
    float rh = end_h - start_h;
    if (!aligned) {
        rw = rw > 1.0f ? rw : 1.0f;
        rh = rh > 1.0f ? rh : 1.0f;
    }
    *bh = rh / (float)PH;
    *bw = rw / (float)PW;
    *gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)PH);
    *gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)PW);
    *sh = start_h;
    *sw = start_w;
}

static inline void make_tap(float y, float x, int H, int W, tap_t *t)
{
    t->y = y;
    t->x = x;
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) {
        t->y_low = t->x_low = t->y_high = t->x_high = -1;
        t->w1 = t->w2 = t->w3 = t->w4 = 0.0f;
        return;
    }
    if (y <= 0.0f) y = 0.0f;
    if (x <= 0.0f) x = 0.0f;
    int y_low = (int)y, x_low = (int)x, y_high, x_high;
    if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else { y_high = y_low + 1; }
    if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else { x_high = x_low + 1; }
    const float ly = y - (float)y_low, lx = x - (float)x_low;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    t->y_low = y_low; t->x_low = x_low; t->y_high = y_high; t->x_high = x_high;
    t->w1 = hy * hx; t->w2 = hy * lx; t->w3 = ly * hx; t->w4 = ly * lx;
}

/* Forward.  feat [N,C,H,W] fp32 NCHW, rois [R,5] = (batch_idx,x1,y1,x2,y2), out [R,C,PH,PW]. */
int oracle_roi_align_fwd(const float *feat, int N, int C, int H, int W, const float *rois, int R,
                         int PH, int PW, float scale, int sampling_ratio, int aligned, float *out)
{
    (void)N;
    for (int r = 0; r < R; ++r) {
        const float *roi = rois + (size_t)r * 5;
        const int b = (int)roi[0];
        float sh, sw, bh, bw; int gh, gw;
        roi_geometry(roi, scale, aligned, PH, PW, sampling_ratio, &sh, &sw, &bh, &bw, &gh, &gw);
        const int cnt_i = gh * gw;
        const float count = (float)(cnt_i > 1 ? cnt_i : 1);
        for (int c = 0; c < C; ++c) {
            const float *plane = feat + ((size_t)b * C + c) * H * W;
            float *o = out + ((size_t)r * C + c) * PH * PW;
            for (int ph = 0; ph < PH; ++ph)
                for (int pw = 0; pw < PW; ++pw) {
                    float acc = 0.0f;
                    for (int iy = 0; iy < gh; ++iy) {
                        const float yy = sh + (float)ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
                        for (int ix = 0; ix < gw; ++ix) {
                            const float xx = sw + (float)pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
                            tap_t t;
                            make_tap(yy, xx, H, W, &t);
                            if (t.y_low < 0) continue;   /* weights are 0: adds +0.0f in torchvision */
                            acc += t.w1 * plane[t.y_low * W + t.x_low] + t.w2 * plane[t.y_low * W + t.x_high] +
                                   t.w3 * plane[t.y_high * W + t.x_low] + t.w4 * plane[t.y_high * W + t.x_high];
                        }
                    }
                    o[ph * PW + pw] = acc / count;
                }
        }
    }
    return 0;
}

/* Grid dump for ONE roi: grid_hw[2] = (gh, gw); for every (ph, pw, iy, ix) in that nesting order
 * yx[2] fp32 (unclamped sample coordinate) and idx[4] int32 (y_low, x_low, y_high, x_high; -1 = skipped).
 * max_samples bounds the arrays; returns the number of samples written or -1 if it would overflow. */
int oracle_roi_align_grid(const float *roi, int H, int W, int PH, int PW, float scale, int sampling_ratio,
                          int aligned, int32_t *grid_hw, float *yx, int32_t *idx, int max_samples)
{
    float sh, sw, bh, bw; int gh, gw;
    roi_geometry(roi, scale, aligned, PH, PW, sampling_ratio, &sh, &sw, &bh, &bw, &gh, &gw);
    grid_hw[0] = gh; grid_hw[1] = gw;
    if (gh <= 0 || gw <= 0) return 0;
    const long total = (long)PH * PW * gh * gw;
    if (total > max_samples) return -1;
    int n = 0;
    for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw)
            for (int iy = 0; iy < gh; ++iy) {
                const float yy = sh + (float)ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
                for (int ix = 0; ix < gw; ++ix) {
                    const float xx = sw + (float)pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
                    tap_t t;
                    make_tap(yy, xx, H, W, &t);
                    yx[2 * n] = t.y; yx[2 * n + 1] = t.x;
                    idx[4 * n] = t.y_low; idx[4 * n + 1] = t.x_low; idx[4 * n + 2] = t.y_high; idx[4 * n + 3] = t.x_high;
                    ++n;
                }
            }
    return n;
}

/* Backward (torchvision roi_align_backward semantics): dfeat must be zero-initialised by the caller. */
int oracle_roi_align_bwd(const float *dout, int N, int C, int H, int W, const float *rois, int R,
                         int PH, int PW, float scale, int sampling_ratio, int aligned, float *dfeat)
{
    (void)N;
    for (int r = 0; r < R; ++r) {
        const float *roi = rois + (size_t)r * 5;
        const int b = (int)roi[0];
        float sh, sw, bh, bw; int gh, gw;
        roi_geometry(roi, scale, aligned, PH, PW, sampling_ratio, &sh, &sw, &bh, &bw, &gh, &gw);
        const int cnt_i = gh * gw;
        const float count = (float)(cnt_i > 1 ? cnt_i : 1);
        for (int c = 0; c < C; ++c) {
            float *plane = dfeat + ((size_t)b * C + c) * H * W;
            const float *g = dout + ((size_t)r * C + c) * PH * PW;
            for (int ph = 0; ph < PH; ++ph)
                for (int pw = 0; pw < PW; ++pw) {
                    const float gv = g[ph * PW + pw];
                    for (int iy = 0; iy < gh; ++iy) {
                        const float yy = sh + (float)ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
                        for (int ix = 0; ix < gw; ++ix) {
                            const float xx = sw + (float)pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
                            tap_t t;
                            make_tap(yy, xx, H, W, &t);
                            if (t.y_low < 0) continue;
                            plane[t.y_low * W + t.x_low] += gv * t.w1 / count;
                            plane[t.y_low * W + t.x_high] += gv * t.w2 / count;
                            plane[t.y_high * W + t.x_low] += gv * t.w3 / count;
                            plane[t.y_high * W + t.x_high] += gv * t.w4 / count;
                        }
                    }
                }
        }
    }
    return 0;
}
