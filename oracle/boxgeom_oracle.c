/*
 * boxgeom_oracle.c -- CPU restatement of the box-geometry hot path of
 * ches-001/vision-conglomerate.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke test
 * in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (vision_conglomerate_b200/csrc) never links or calls it.
 *
 * Plain scalar C, fp32 arithmetic without FMA contraction (build with
 * -ffp-contract=off), one function per reference routine.  Each function cites
 * the reference file:line it follows (paths relative to the reference root).
 *
 * Parity pinning: the reference has no tests or golden vectors of its own
 * (SURVEY.md section 4).  This restatement is pinned against OUTPUTS OF THE
 * REFERENCE ITSELF, produced by importing /root/reference in the build
 * container (oracle/make_golden.py) and committed under tests/golden/.
 * The NMS arithmetic lives in torchvision (unpinned third-party dependency of
 * the reference, installed version 0.26.0): bgo_nms restates the published
 * greedy CPU algorithm of torchvision/csrc/ops/cpu/nms_kernel.cpp and is pinned
 * against torchvision-CPU outputs in tests/golden and live in the tests.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BGO_API __attribute__((visibility("default")))

static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

/* ------------------------------------------------------------------ decode
 * modules/detection.py:98-173 (_get_scale_pred), detection branch only
 * (no proto_seg_module, no keypoints).
 * raw, out: [B, ny, nx, na, 5+C] fp32, channels [obj, cls*C, tx, ty, tw, th].
 * inference == 0: xy = 2*sig - 0.5, wh = (2*sig)^2                (:122,:125)
 * inference != 0: xy = (xy + grid) * stride, wh = wh*anchor*[nx,ny]*stride
 *                 with stride = float32([H/ny, W/nx]) applied to (x, y) in
 *                 that order, exactly as the reference does       (:145-155)
 */
BGO_API void bgo_decode_scale(const float *raw, int B, int ny, int nx, int na, int C,
                              const float *anchors, int H, int W, int inference, float *out)
{
    const int D = 5 + C;
    const float s0 = (float)H / (float)ny; /* multiplies x (reference quirk) */
    const float s1 = (float)W / (float)nx; /* multiplies y */
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
                for (int a = 0; a < na; ++a) {
                    size_t r = ((((size_t)b * ny + y) * nx + x) * na + a) * (size_t)D;
                    const float *in = raw + r;
                    float *o = out + r;
                    for (int c = 0; c <= C; ++c) o[c] = in[c]; /* logits pass through :116,:119 */
                    float bx = sigmoidf_(in[C + 1]) * 2.0f - 0.5f;
                    float by = sigmoidf_(in[C + 2]) * 2.0f - 0.5f;
                    float bw = sigmoidf_(in[C + 3]) * 2.0f; bw = bw * bw;
                    float bh = sigmoidf_(in[C + 4]) * 2.0f; bh = bh * bh;
                    if (inference) {
                        bx = (bx + (float)x) * s0;
                        by = (by + (float)y) * s1;
                        bw = ((bw * anchors[2 * a + 0]) * (float)nx) * s0;
                        bh = ((bh * anchors[2 * a + 1]) * (float)ny) * s1;
                    }
                    o[C + 1] = bx; o[C + 2] = by; o[C + 3] = bw; o[C + 4] = bh;
                }
}

/* modules/detection.py:175-190 (_bbox_to_size): box = (box / from) * to, in place,
 * from = [W,H,W,H], to = [W0,H0,W0,H0].  The caller applies the guard at :76. */
BGO_API void bgo_bbox_to_size(float *pred, size_t rows, int C, int H, int W, int H0, int W0)
{
    const int D = 5 + C;
    for (size_t r = 0; r < rows; ++r) {
        float *p = pred + r * D + C + 1;
        p[0] = (p[0] / (float)W) * (float)W0;
        p[1] = (p[1] / (float)H) * (float)H0;
        p[2] = (p[2] / (float)W) * (float)W0;
        p[3] = (p[3] / (float)H) * (float)H0;
    }
}

/* inference_det.py:57-76 + utils/utils.py:215-226.
 * preds: [n, 5+C] decoded rows.  Outputs per row: score = max_c sig(cls_c) * sig(obj),
 * cls = first argmax_c sig(cls_c), xyxy with (w,h) += box_allowance first and
 * x2 = x1 + w.  The reference mutates preds' w/h in place (:73-74); we do not. */
BGO_API void bgo_score_xyxy(const float *preds, size_t n, int C, float box_allowance,
                            float *score, int32_t *cls, float *xyxy)
{
    const int D = 5 + C;
    for (size_t r = 0; r < n; ++r) {
        const float *p = preds + r * D;
        float conf = sigmoidf_(p[0]);
        float best = -1.0f; int bi = 0;
        for (int c = 0; c < C; ++c) {
            float s = sigmoidf_(p[1 + c]);
            if (s > best) { best = s; bi = c; }
        }
        score[r] = best * conf;
        cls[r] = bi;
        float w = p[C + 3] + box_allowance, h = p[C + 4] + box_allowance;
        float x1 = p[C + 1] - w / 2.0f, y1 = p[C + 2] - h / 2.0f;
        xyxy[4 * r + 0] = x1; xyxy[4 * r + 1] = y1;
        xyxy[4 * r + 2] = x1 + w; xyxy[4 * r + 3] = y1 + h;
    }
}

/* -------------------------------------------------------------------- NMS
 * torchvision 0.26 csrc/ops/cpu/nms_kernel.cpp (nms_kernel_impl<float>), the
 * kernel behind torchvision.ops.nms that inference_det.py:77-82 reaches through
 * batched_nms.  Stable descending score order (ties: lower index first), areas
 * pre-rounded in fp32, IoU in fp32, compared as double against the un-rounded
 * threshold with strict '>'.  `sel` (length m) lists the candidate indices in
 * ascending order; keep receives original indices in greedy (score) order.
 */
typedef struct { float s; int64_t i; } bgo_key_t;
static int bgo_key_cmp(const void *a, const void *b)
{
    const bgo_key_t *x = (const bgo_key_t *)a, *y = (const bgo_key_t *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

static int64_t bgo_nms_subset(const float *boxes, const float *scores, const int64_t *sel,
                              int64_t m, double thr, int64_t *keep)
{
    if (m <= 0) return 0;
    bgo_key_t *ord = (bgo_key_t *)malloc(sizeof(bgo_key_t) * (size_t)m);
    float *bx = (float *)malloc(sizeof(float) * 5 * (size_t)m);
    unsigned char *sup = (unsigned char *)calloc((size_t)m, 1);
    for (int64_t k = 0; k < m; ++k) { ord[k].i = sel ? sel[k] : k; ord[k].s = scores[ord[k].i]; }
    qsort(ord, (size_t)m, sizeof(bgo_key_t), bgo_key_cmp);
    for (int64_t k = 0; k < m; ++k) { /* gather in sorted order: x1,y1,x2,y2,area */
        const float *b = boxes + 4 * ord[k].i;
        bx[5 * k + 0] = b[0]; bx[5 * k + 1] = b[1]; bx[5 * k + 2] = b[2]; bx[5 * k + 3] = b[3];
        bx[5 * k + 4] = (b[2] - b[0]) * (b[3] - b[1]);
    }
    int64_t nk = 0;
    for (int64_t i = 0; i < m; ++i) {
        if (sup[i]) continue;
        keep[nk++] = ord[i].i;
        const float ix1 = bx[5 * i], iy1 = bx[5 * i + 1], ix2 = bx[5 * i + 2], iy2 = bx[5 * i + 3];
        const float ia = bx[5 * i + 4];
        for (int64_t j = i + 1; j < m; ++j) {
            if (sup[j]) continue;
            const float *q = bx + 5 * j;
            float xx1 = ix1 < q[0] ? q[0] : ix1;
            float yy1 = iy1 < q[1] ? q[1] : iy1;
            float xx2 = q[2] < ix2 ? q[2] : ix2;
            float yy2 = q[3] < iy2 ? q[3] : iy2;
            float w = xx2 - xx1; w = (0.0f < w) ? w : 0.0f;
            float h = yy2 - yy1; h = (0.0f < h) ? h : 0.0f;
            float inter = w * h;
            float ovr = inter / (ia + q[4] - inter);
            if ((double)ovr > thr) sup[j] = 1;
        }
    }
    free(ord); free(bx); free(sup);
    return nk;
}

BGO_API int64_t bgo_nms(const float *boxes, const float *scores, int64_t n, double thr, int64_t *keep)
{
    return bgo_nms_subset(boxes, scores, NULL, n, thr, keep);
}

/* torchvision/ops/boxes.py _batched_nms_vanilla (the branch every BASELINE
 * config takes): greedy NMS independently per distinct idxs value,
 * result = all kept indices, score-descending.  torchvision's final sort is not
 * stable; we emit the canonical order (score desc, index asc). */
typedef struct { int64_t g; int64_t i; } bgo_gi_t;
static int bgo_gi_cmp(const void *a, const void *b)
{
    const bgo_gi_t *x = (const bgo_gi_t *)a, *y = (const bgo_gi_t *)b;
    if (x->g != y->g) return (x->g > y->g) - (x->g < y->g);
    return (x->i > y->i) - (x->i < y->i);
}

BGO_API int64_t bgo_batched_nms(const float *boxes, const float *scores, const int64_t *idxs,
                                int64_t n, double thr, int64_t *keep)
{
    if (n <= 0) return 0;
    bgo_gi_t *gi = (bgo_gi_t *)malloc(sizeof(bgo_gi_t) * (size_t)n);
    int64_t *sel = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int64_t *gstart = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int64_t *gkept = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t k = 0; k < n; ++k) { gi[k].g = idxs[k]; gi[k].i = k; }
    qsort(gi, (size_t)n, sizeof(bgo_gi_t), bgo_gi_cmp);
    int64_t ng = 0;
    for (int64_t k = 0; k < n; ++k) {
        sel[k] = gi[k].i;
        if (k == 0 || gi[k].g != gi[k - 1].g) gstart[ng++] = k;
    }
    gstart[ng] = n;
    /* groups are independent: one host thread per group when built with -fopenmp */
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t g = 0; g < ng; ++g)
        gkept[g] = bgo_nms_subset(boxes, scores, sel + gstart[g], gstart[g + 1] - gstart[g], thr,
                                  tmp + gstart[g]);
    int64_t nk = 0;
    for (int64_t g = 0; g < ng; ++g)
        for (int64_t k = 0; k < gkept[g]; ++k) keep[nk++] = tmp[gstart[g] + k];
    free(tmp); free(gstart); free(gkept);
    bgo_key_t *ord = (bgo_key_t *)malloc(sizeof(bgo_key_t) * (size_t)(nk ? nk : 1));
    for (int64_t k = 0; k < nk; ++k) { ord[k].i = keep[k]; ord[k].s = scores[keep[k]]; }
    qsort(ord, (size_t)nk, sizeof(bgo_key_t), bgo_key_cmp);
    for (int64_t k = 0; k < nk; ++k) keep[k] = ord[k].i;
    free(ord); free(gi); free(sel);
    return nk;
}

/* ------------------------------------------------------- target assignment
 * dataset/detection_dataset.py:90-246 (build_target_by_scale), detection
 * branch (overlap_masks=None, no keypoint columns).  SURVEY A.2.
 * targets [nt,6] = (img, cls, x, y, w, h) normalised; anchors [na,2] normalised.
 * Output order: offset k in 0..4, anchor a, target t (lexicographic).
 * out_idx4: [4, cap] rows = batch_idx, grid_j, grid_i, anchor_idx (int64);
 * returns M.  cap must be >= 5*na*nt.
 */
static int64_t bgo_assign_impl(const float *targets, int64_t nt, int stride, int ny, int nx, const float *anchors,
                               int na, float anchor_t, float edge_t, int64_t cap, int64_t *out_idx4,
                               int64_t *out_cls, float *out_anchor, float *out_box, int64_t *out_t)
{
    static const float offx[5] = {0.f, 1.f, 0.f, -1.f, 0.f};
    static const float offy[5] = {0.f, 0.f, 1.f, 0.f, -1.f};
    int64_t M = 0;
    const float fnx = (float)nx, fny = (float)ny;
    for (int k = 0; k < 5; ++k) {
        const float ox = offx[k] * edge_t, oy = offy[k] * edge_t; /* :220 */
        for (int a = 0; a < na; ++a) {
            const float aw = anchors[2 * a] * fnx, ah = anchors[2 * a + 1] * fny; /* :183 */
            for (int64_t t = 0; t < nt; ++t) {
                const float *tg = targets + (int64_t)stride * t;
                const float gx = tg[2] * fnx, gy = tg[3] * fny, gw = tg[4] * fnx, gh = tg[5] * fny; /* :184 */
                const float rw = gw / aw, rh = gh / ah;                                              /* :190 */
                const float irw = 1.0f / rw, irh = 1.0f / rh;
                float m = rw > irw ? rw : irw; /* torch.max propagates NaN; inputs assumed finite */
                float m2 = rh > irh ? rh : irh;
                if (m2 > m) m = m2;
                if (!(m < anchor_t)) continue;                                                       /* :191 */
                int sel;
                const float ix = fnx - gx, iy = fny - gy;                                           /* :195 */
                switch (k) {
                case 0: sel = 1; break;
                case 1: sel = (fmodf(gx, 1.0f) < edge_t) && (gx > 1.0f); break;                     /* :200 */
                case 2: sel = (fmodf(gy, 1.0f) < edge_t) && (gy > 1.0f); break;
                case 3: sel = (fmodf(ix, 1.0f) < edge_t) && (ix > 1.0f); break;                     /* :201 */
                default: sel = (fmodf(iy, 1.0f) < edge_t) && (iy > 1.0f); break;
                }
                if (!sel) continue;
                if (M >= cap) return -1;
                int64_t gi = (int64_t)(gx - ox), gj = (int64_t)(gy - oy); /* .long() truncates :231 */
                if (gi < 0) gi = 0; if (gi > nx - 1) gi = nx - 1;          /* :232-233 */
                if (gj < 0) gj = 0; if (gj > ny - 1) gj = ny - 1;
                out_idx4[0 * cap + M] = (int64_t)tg[0];
                out_idx4[1 * cap + M] = gj;
                out_idx4[2 * cap + M] = gi;
                out_idx4[3 * cap + M] = a;
                out_cls[M] = (int64_t)tg[1];
                out_anchor[2 * M] = aw; out_anchor[2 * M + 1] = ah;
                out_box[4 * M + 0] = gx - (float)gi; /* clamped gi,gj (aliasing) :237 */
                out_box[4 * M + 1] = gy - (float)gj;
                out_box[4 * M + 2] = gw; out_box[4 * M + 3] = gh;
                if (out_t) out_t[M] = t;
                ++M;
            }
        }
    }
    return M;
}

BGO_API int64_t bgo_assign(const float *targets, int64_t nt, int ny, int nx, const float *anchors,
                           int na, float anchor_t, float edge_t, int64_t cap, int64_t *out_idx4,
                           int64_t *out_cls, float *out_anchor, float *out_box)
{
    return bgo_assign_impl(targets, nt, 6, ny, nx, anchors, na, anchor_t, edge_t, cap, out_idx4, out_cls, out_anchor, out_box, NULL);
}

/* Segmentation / keypoint variants (detection_dataset.py:127-172,239-245): rows of `stride` floats, the extra
 * columns ride along unchanged (gain 1, :173-176); out_t [cap] = source target of every match, from which the
 * caller derives tmask_idx (:132-170) and the keypoint rows (:244). */
BGO_API int64_t bgo_assign_ex(const float *targets, int64_t nt, int stride, int ny, int nx, const float *anchors,
                              int na, float anchor_t, float edge_t, int64_t cap, int64_t *out_idx4,
                              int64_t *out_cls, float *out_anchor, float *out_box, int64_t *out_t)
{
    return bgo_assign_impl(targets, nt, stride, ny, nx, anchors, na, anchor_t, edge_t, cap, out_idx4, out_cls, out_anchor, out_box, out_t);
}

/* ------------------------------------------------------------------- CIoU
 * modules/detection_loss.py:229-264 (compute_ciou), element-wise form.
 * If grad_p != NULL also writes d ciou / d preds_xywh [M,4] with `a` held
 * constant (the reference computes it under no_grad, :261-262).
 * Forward is fp32 in the reference's operation order; the gradient is
 * evaluated in double from the same formulas (autograd in the reference is
 * fp32; parity tolerance is rtol 1e-5 against golden autograd outputs).
 */
static float bgo_ciou_one(const float p[4], const float t[4], float e, double *g /*4 or NULL*/)
{
    const float pw = p[2], ph = p[3], tw = t[2], th = t[3];
    const float px1 = p[0] - pw / 2.0f, py1 = p[1] - ph / 2.0f;
    const float px2 = px1 + pw, py2 = py1 + ph;
    const float tx1 = t[0] - tw / 2.0f, ty1 = t[1] - th / 2.0f;
    const float tx2 = tx1 + tw, ty2 = ty1 + th;
    float iw = fminf(px2, tx2) - fmaxf(px1, tx1); if (iw < 0.f) iw = 0.f;
    float ih = fminf(py2, ty2) - fmaxf(py1, ty1); if (ih < 0.f) ih = 0.f;
    const float inter = iw * ih;
    const float uni = (pw * ph) + (tw * th) - inter;
    const float iou = inter / (uni + e);
    const float cw = fmaxf(px2, tx2) - fminf(px1, tx1);
    const float ch = fmaxf(py2, ty2) - fminf(py1, ty1);
    const float c2 = cw * cw + ch * ch + e;
    const float k4pi2 = (float)(4.0 / (M_PI * M_PI));
    const float dat = atanf(tw / th) - atanf(pw / ph);
    const float v = k4pi2 * (dat * dat);
    const float dx = p[0] - t[0], dy = p[1] - t[1];
    const float rho2 = dx * dx + dy * dy;
    const float a = v / (v - iou + (1.0f + e));
    const float ciou = iou - ((rho2 / c2) + (a * v));
    if (g) {
        /* variables: x, y, w, h of the prediction; px1 = x - w/2, px2 = px1 + w = x + w/2 */
        const double dpx1[4] = {1, 0, -0.5, 0}, dpx2[4] = {1, 0, 0.5, 0};
        const double dpy1[4] = {0, 1, 0, -0.5}, dpy2[4] = {0, 1, 0, 0.5};
        const int iw_pos = (fminf(px2, tx2) - fmaxf(px1, tx1)) >= 0.f; /* clamp backward passes at equality */
        const int ih_pos = (fminf(py2, ty2) - fmaxf(py1, ty1)) >= 0.f;
        /* torch.min/max backward: gradient goes to the selected operand; on exact ties it is split
         * evenly between both, so the prediction operand receives 0.5. */
        const double s_minx2 = px2 < tx2 ? 1.0 : (px2 == tx2 ? 0.5 : 0.0);
        const double s_maxx1 = px1 > tx1 ? 1.0 : (px1 == tx1 ? 0.5 : 0.0);
        const double s_miny2 = py2 < ty2 ? 1.0 : (py2 == ty2 ? 0.5 : 0.0);
        const double s_maxy1 = py1 > ty1 ? 1.0 : (py1 == ty1 ? 0.5 : 0.0);
        const double s_maxx2 = px2 > tx2 ? 1.0 : (px2 == tx2 ? 0.5 : 0.0);
        const double s_minx1 = px1 < tx1 ? 1.0 : (px1 == tx1 ? 0.5 : 0.0);
        const double s_maxy2 = py2 > ty2 ? 1.0 : (py2 == ty2 ? 0.5 : 0.0);
        const double s_miny1 = py1 < ty1 ? 1.0 : (py1 == ty1 ? 0.5 : 0.0);
        const double den = (double)uni + e;
        const double r = (double)pw / ph;
        const double datd = (double)dat;
        for (int q = 0; q < 4; ++q) {
            double diw = iw_pos ? (s_minx2 * dpx2[q] - s_maxx1 * dpx1[q]) : 0.0;
            double dih = ih_pos ? (s_miny2 * dpy2[q] - s_maxy1 * dpy1[q]) : 0.0;
            double dinter = diw * ih + iw * dih;
            double dpwph = (q == 2 ? ph : 0.0) + (q == 3 ? pw : 0.0);
            double duni = dpwph - dinter;
            double diou = dinter / den - (double)inter * duni / (den * den);
            double dcw = s_maxx2 * dpx2[q] - s_minx1 * dpx1[q];
            double dch = s_maxy2 * dpy2[q] - s_miny1 * dpy1[q];
            double dc2 = 2.0 * cw * dcw + 2.0 * ch * dch;
            double drho2 = (q == 0 ? 2.0 * dx : 0.0) + (q == 1 ? 2.0 * dy : 0.0);
            /* d atan(pw/ph) */
            double dr = (q == 2 ? 1.0 / ph : 0.0) + (q == 3 ? -(double)pw / ((double)ph * ph) : 0.0);
            double datan_p = dr / (1.0 + r * r);
            double dv = (double)k4pi2 * 2.0 * datd * (-datan_p);
            /* v only depends on (w,h): autograd never multiplies `a` into the x,y gradients, which matters
             * when a is NaN (identical boxes can give iou > 1 in fp32 and a = 0/0). */
            double av = (q >= 2) ? (double)a * dv : 0.0;
            g[q] = diou - (drho2 / c2 - (double)rho2 * dc2 / ((double)c2 * c2) + av);
        }
    }
    return ciou;
}

BGO_API void bgo_ciou(const float *p, const float *t, int64_t M, float e, float *out, float *grad_p)
{
    for (int64_t m = 0; m < M; ++m) {
        double g[4];
        out[m] = bgo_ciou_one(p + 4 * m, t + 4 * m, e, grad_p ? g : NULL);
        if (grad_p) for (int q = 0; q < 4; ++q) grad_p[4 * m + q] = (float)g[q];
    }
}

/* ------------------------------------------------------------------- loss
 * modules/detection_loss.py:125-226 (loss_fn) for one scale, default config
 * (BCEWithLogits, no focal, no keypoints), SURVEY A.3.
 * preds [B,ny,nx,na,5+C] (training-decoded), assignment outputs as produced by
 * bgo_assign (idx4 with row stride cap).
 * scalars[0..7]  = ciou_loss(lbox, NaN->0), conf_loss, class_loss (NaN->0), mean_ciou,
 *                  avg_pos_conf, avg_neg_conf, M, n_neg
 * hist [3*C] int64 = tp_c, n_true_c, n_pred_c          (for sklearn macro metrics, :198-206)
 * If grad != NULL: grad [same shape as preds] = d(w_box*lbox + w_conf*lconf + w_cls*lcls)/d preds.
 * t_conf scatter: last match in (k,a,t) order wins (single-thread index_put_ semantics, :182).
 */
static inline double bce_logits(double x, double t)
{
    /* ATen: (1 - t) * x - log_sigmoid(x), log_sigmoid(x) = min(x,0) - log1p(exp(-|x|)) */
    double ls = (x < 0 ? x : 0.0) - log1p(exp(-fabs(x)));
    return (1.0 - t) * x - ls;
}

BGO_API void bgo_loss_scale(const float *preds, int B, int ny, int nx, int na, int C,
                            const int64_t *idx4, int64_t cap, const int64_t *cls,
                            const float *anchor, const float *box, int64_t M, float label_smoothing,
                            float w_box, float w_conf, float w_cls, double *scalars, int64_t *hist,
                            float *grad)
{
    const int D = 5 + C;
    const size_t cells = (size_t)B * ny * nx * na;
    float *tconf = (float *)calloc(cells, sizeof(float));
    float *ciou = (float *)malloc(sizeof(float) * (size_t)(M ? M : 1));
    double sum_1mciou = 0, sum_ciou = 0, sum_pos = 0, sum_cls = 0;
    const double cn = 0.5 * (double)label_smoothing, cp = 1.0 - cn;
    if (grad) memset(grad, 0, cells * D * sizeof(float));
    memset(hist, 0, sizeof(int64_t) * 3 * (size_t)C);
    for (int64_t m = 0; m < M; ++m) {
        const size_t cell = (((size_t)idx4[m] * ny + idx4[cap + m]) * nx + idx4[2 * cap + m]) * na + idx4[3 * cap + m];
        const float *row = preds + cell * D;
        float p[4] = {row[C + 1], row[C + 2], row[C + 3] * anchor[2 * m], row[C + 4] * anchor[2 * m + 1]};
        double g[4];
        ciou[m] = bgo_ciou_one(p, box + 4 * m, 1e-7f, grad ? g : NULL);
        sum_1mciou += (double)(1.0f - ciou[m]);
        sum_ciou += ciou[m];
        sum_pos += (double)sigmoidf_(row[0]);
        tconf[cell] = ciou[m]; /* later m overwrites earlier: last wins */
        int am = 0; float best = row[1];
        for (int c = 0; c < C; ++c) {
            double t = (c == cls[m]) ? cp : cn;
            sum_cls += bce_logits(row[1 + c], t);
            if (row[1 + c] > best) { best = row[1 + c]; am = c; }
            if (grad) grad[cell * D + 1 + c] += (float)(w_cls * ((double)sigmoidf_(row[1 + c]) - t) / ((double)M * C));
        }
        hist[0 * C + cls[m]] += (am == cls[m]);
        hist[1 * C + cls[m]] += 1;
        hist[2 * C + am] += 1;
        if (grad) {
            const double s = -(double)w_box / (double)M; /* d mean(1-ciou) */
            grad[cell * D + C + 1] += (float)(s * g[0]);
            grad[cell * D + C + 2] += (float)(s * g[1]);
            grad[cell * D + C + 3] += (float)(s * g[2] * anchor[2 * m]);
            grad[cell * D + C + 4] += (float)(s * g[3] * anchor[2 * m + 1]);
        }
    }
    double sum_conf = 0, sum_neg = 0; int64_t n_neg = 0;
    for (size_t c = 0; c < cells; ++c) {
        const float x = preds[c * D];
        sum_conf += bce_logits(x, tconf[c]);
        if (tconf[c] == 0.0f) { sum_neg += (double)sigmoidf_(x); ++n_neg; }
        if (grad) grad[c * D] += (float)(w_conf * ((double)sigmoidf_(x) - (double)tconf[c]) / (double)cells);
    }
    scalars[0] = M ? sum_1mciou / (double)M : 0.0;
    scalars[1] = sum_conf / (double)cells;
    scalars[2] = M ? sum_cls / ((double)M * C) : 0.0;
    scalars[3] = M ? sum_ciou / (double)M : NAN;
    scalars[4] = M ? sum_pos / (double)M : NAN;
    scalars[5] = n_neg ? sum_neg / (double)n_neg : NAN;
    scalars[6] = (double)M;
    scalars[7] = (double)n_neg;
    free(tconf); free(ciou);
}

/* --------------------------------------------------------- anchor metrics
 * utils/make_anchors.py:14-39 (ratio_metrics / ratio_metrics_w_extras).
 * wh [n,2], anchors [k,2]; out = {score, bpr, aat}. */
BGO_API void bgo_ratio_metrics(const float *wh, int64_t n, const float *anchors, int k, float threshold,
                               double *out3)
{
    double s_score = 0, s_m = 0;
    const float inv_t = 1.0f / threshold; /* python: 1 / threshold in double, compared against fp32 v */
    const double inv_td = 1.0 / (double)threshold;
    (void)inv_t;
    for (int64_t i = 0; i < n; ++i) {
        float best = -INFINITY;
        for (int j = 0; j < k; ++j) {
            float r0 = wh[2 * i] / anchors[2 * j], r1 = wh[2 * i + 1] / anchors[2 * j + 1];
            float i0 = 1.0f / r0, i1 = 1.0f / r1;
            float m0 = r0 < i0 ? r0 : i0, m1 = r1 < i1 ? r1 : i1;
            float v = m0 < m1 ? m0 : m1;
            if (v > best) best = v;
        }
        /* `v > 1/threshold`: torch compares the fp32 tensor with a python double scalar; the scalar is
         * cast to the tensor dtype (fp32) for the comparison. */
        const int m = best > (float)inv_td;
        s_m += m;
        s_score += m ? (double)best : 0.0;
    }
    out3[0] = n ? s_score / (double)n : NAN;
    out3[1] = n ? s_m / (double)n : NAN;
    out3[2] = s_m;
}
