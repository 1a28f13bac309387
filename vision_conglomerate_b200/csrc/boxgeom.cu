// libboxgeom.so -- host side of the C ABI declared in include/boxgeom.h.
// Single translation unit (the kernels live in the *_kernels.cuh headers), built for sm_100a only:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "nms_kernels.cuh"
#include "detect_kernels.cuh"
#include "train_kernels.cuh"

namespace bg {

unsigned long long g_launches = 0;
static cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;

IouThr make_iou_thr(double thr)
{
    IouThr t;
    if (thr != thr) {  // NaN threshold: nothing is ever suppressed
        t.tdn = INFINITY; t.lo = t.hi = INFINITY; t.fast_ok = 0; t.zero_suppresses = 0;
        return t;
    }
    float f = (float)thr;
    if ((double)f > thr) f = nextafterf(f, -INFINITY);  // largest float <= thr
    t.tdn = f;
    t.zero_suppresses = (0.0 > thr) ? 1 : 0;
    t.fast_ok = (f >= 1e-30f && f <= 1e30f) ? 1 : 0;
    t.lo = f * (1.0f - 4.76837158203125e-07f);  // 2^-21 guard band, see iou_suppresses
    t.hi = f * (1.0f + 4.76837158203125e-07f);
    return t;
}

struct Bump {  // workspace carving; with base == nullptr it only measures
    unsigned char *base;
    size_t off;
    template <typename T>
    T *take(size_t n)
    {
        off = align_up(off, 256);
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

static int num_sms()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

static void carve_seg(Bump &b, SegNms &p, long long S_max, long long elems, size_t mask_bytes)
{
    p.hdr = b.take<SegHdr>(1);
    p.seg_count = b.take<int>(S_max);
    p.seg_off = b.take<long long>(S_max + 1);
    p.tile_prefix = b.take<int>(S_max + 1);
    p.mask_off = b.take<long long>(S_max + 1);
    p.item_prefix = b.take<int>(S_max + 1);
    p.emit_count = b.take<int>(S_max);
    p.out_prefix = b.take<long long>(S_max + 1);
    p.keys = b.take<u64>(elems);
    p.sorted_box = b.take<float4>(elems);
    p.sorted_area = b.take<float>(elems);
    p.bkeys = b.take<u64>(elems);
    p.bbox = b.take<float4>(elems);
    p.barea = b.take<float>(elems);
    p.keepbits = b.take<u64>(elems / 64 + S_max + 1);
    p.ew32 = b.take<u32>(2 * (elems / 64 + S_max + 1));
    p.rank32 = b.take<u32>(2 * (elems / 64 + S_max + 1));
    p.emit_pos = b.take<u32>(elems);
    p.emit_key = b.take<u64>(elems);
    p.mask = b.take<u64>(mask_bytes / 8);
    p.mask_words = (long long)(mask_bytes / 8);
}

// ---------------------------------------------------------------------------------------------
struct GnmsWs { SegNms p; u32 *slot; };
static size_t gnms_carve(unsigned char *base, long long n, long long max_groups, size_t mask_bytes, GnmsWs &w)
{
    Bump b{base, 0};
    memset(&w.p, 0, sizeof(w.p));
    carve_seg(b, w.p, max_groups, 2 * n + 64, mask_bytes);
    w.slot = b.take<u32>(n);
    return align_up(b.off, 256);
}

struct DetWs { SegNms p; float4 *box_dense; int *cls_dense; long long stride; };
static size_t det_carve(unsigned char *base, int B, long long N, size_t mask_bytes, DetWs &w)
{
    Bump b{base, 0};
    memset(&w.p, 0, sizeof(w.p));
    w.stride = (long long)next_pow2((u32)N);
    carve_seg(b, w.p, B, (long long)B * w.stride, mask_bytes);
    w.box_dense = b.take<float4>((size_t)B * N);
    w.cls_dense = b.take<int>((size_t)B * N);
    return align_up(b.off, 256);
}

static long long det_candidates(const bg_detect_params *p)
{
    long long N = 0;
    for (int s = 0; s < 3; ++s) N += (long long)p->ny[s] * p->nx[s] * p->na;
    return N;
}

static bool det_valid(const bg_detect_params *p)
{
    if (!p || p->B <= 0 || p->C <= 0 || p->na <= 0 || p->na > BG_MAX_ANCHORS || p->H <= 0 || p->W <= 0) return false;
    if (p->n_tracked < 0 || p->n_tracked > BG_MAX_TRACKED) return false;
    for (int s = 0; s < 3; ++s)
        if (p->ny[s] <= 0 || p->nx[s] <= 0) return false;
    const long long N = det_candidates(p);
    return N > 0 && N < (1ll << 31) && (long long)p->B * N < (1ll << 31);
}

}  // namespace bg

using namespace bg;

extern "C" {

const char *bg_strerror(int code)
{
    switch (code) {
    case BG_OK: return "ok";
    case BG_ERR_INVALID: return "invalid argument";
    case BG_ERR_WORKSPACE: return "workspace too small";
    case BG_ERR_LAUNCH: return "CUDA error while enqueueing (is this an sm_100a device?)";
    default: return "unknown error";
    }
}

int bg_version(void) { return 100; }

uint64_t bg_launch_count(void) { return g_launches; }
void bg_profile_events(void *start, void *stop) { g_prof_start = (cudaEvent_t)start; g_prof_stop = (cudaEvent_t)stop; }
size_t bg_sizeof_detect_params(void) { return sizeof(bg_detect_params); }
size_t bg_sizeof_loss_params(void) { return sizeof(bg_loss_params); }

// ------------------------------------------------------------------------------------------ B4
size_t bg_batched_nms_workspace_bytes(int64_t n, int64_t max_groups, size_t mask_bytes)
{
    if (n < 0 || max_groups <= 0) return 0;
    GnmsWs w;
    return gnms_carve(nullptr, n, max_groups, mask_bytes, w);
}

int bg_batched_nms(const float *boxes, const float *scores, const int64_t *idxs, int64_t n, double iou_threshold,
                   int64_t max_groups, int64_t *out_keep, int32_t *out_counts, void *workspace,
                   size_t workspace_bytes, size_t mask_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 0 || n >= (1ll << 31) || max_groups <= 0 || max_groups > (1ll << 24) || !out_counts) return BG_ERR_INVALID;
    if (n == 0) {
        if (cudaMemsetAsync(out_counts, 0, 2 * sizeof(int32_t), st) != cudaSuccess) return BG_ERR_LAUNCH;
        return BG_OK;
    }
    if (!boxes || !scores || !idxs || !out_keep || !workspace) return BG_ERR_INVALID;
    if (((uintptr_t)boxes & 15) != 0) return BG_ERR_INVALID;  // float4 loads
    GnmsWs w;
    if (gnms_carve((unsigned char *)workspace, n, max_groups, mask_bytes, w) > workspace_bytes) return BG_ERR_WORKSPACE;
    SegNms &p = w.p;
    p.boxes = reinterpret_cast<const float4 *>(boxes);
    p.box_seg_stride = 0;
    p.cls = nullptr;
    p.n_tracked = 0;
    p.thr = make_iou_thr(iou_threshold);
    segnms_configure(p);
    const int sms = num_sms();
    const int gs = sms * 8;
    const long long *gidx = reinterpret_cast<const long long *>(idxs);
    gnms_init_kernel<<<gs, 256, 0, st>>>(p, max_groups);
    BG_LAUNCH_CHECK();
    gnms_minmax_kernel<<<gs, 256, 0, st>>>(p, gidx, n);
    BG_LAUNCH_CHECK();
    gnms_count_kernel<<<gs, 256, 0, st>>>(p, gidx, n, max_groups, w.slot);
    BG_LAUNCH_CHECK();
    gnms_offsets_kernel<<<1, 1024, 0, st>>>(p);
    BG_LAUNCH_CHECK();
    gnms_scatter_kernel<<<gs, 256, 0, st>>>(p, gidx, scores, n, w.slot);
    BG_LAUNCH_CHECK();
    const int S_launch = (int)(max_groups < n ? max_groups : n);
    int rc = segnms_run(p, S_launch, out_counts, 0, sms, st);
    if (rc != BG_OK) return rc;
    const int go = S_launch < 1 ? 1 : (S_launch > 4096 ? 4096 : S_launch);
    gnms_output_kernel<<<dim3(S_launch > 256 ? 1 : 8, go), 256, 0, st>>>(p, reinterpret_cast<long long *>(out_keep));
    BG_LAUNCH_CHECK();
    return BG_OK;
}

// ------------------------------------------------------------------------------------------ B5
size_t bg_detect_workspace_bytes(const bg_detect_params *p, size_t mask_bytes)
{
    if (!det_valid(p)) return 0;
    DetWs w;
    return det_carve(nullptr, p->B, det_candidates(p), mask_bytes, w);
}

int bg_detect(const float *raw_sm, const float *raw_md, const float *raw_lg, const bg_detect_params *pp,
              float *out_boxes, int64_t *out_img, int64_t *out_keep, int32_t *out_counts, void *workspace,
              size_t workspace_bytes, size_t mask_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!det_valid(pp) || !raw_sm || !raw_md || !raw_lg || !out_boxes || !out_img || !out_keep || !out_counts || !workspace)
        return BG_ERR_INVALID;
    const long long N = det_candidates(pp);
    DetWs w;
    if (det_carve((unsigned char *)workspace, pp->B, N, mask_bytes, w) > workspace_bytes) return BG_ERR_WORKSPACE;
    SegNms &p = w.p;
    p.boxes = w.box_dense;
    p.box_seg_stride = N;
    p.cls = w.cls_dense;
    p.n_tracked = pp->n_tracked;
    for (int i = 0; i < pp->n_tracked; ++i) p.tracked[i] = pp->tracked[i];
    p.thr = make_iou_thr(pp->iou_threshold);
    segnms_configure(p);

    DetectK k;
    memset(&k, 0, sizeof(k));
    const float *raws[3] = {raw_sm, raw_md, raw_lg};
    int off = 0;
    bool aligned = true;
    for (int s = 0; s < 3; ++s) {
        ScaleDesc &d = k.sc[s];
        d.raw = raws[s];
        d.ny = pp->ny[s]; d.nx = pp->nx[s];
        d.cells_na = pp->ny[s] * pp->nx[s] * pp->na;
        d.img_off = off;
        off += d.cells_na;
        d.rows = (long long)pp->B * d.cells_na;
        d.s0 = (float)pp->H / (float)pp->ny[s];
        d.s1 = (float)pp->W / (float)pp->nx[s];
        d.fnx = (float)pp->nx[s]; d.fny = (float)pp->ny[s];
        for (int a = 0; a < pp->na; ++a) { d.aw[a] = pp->anchors[s][a][0]; d.ah[a] = pp->anchors[s][a][1]; }
        aligned = aligned && (((uintptr_t)raws[s] & 15) == 0);
    }
    k.B = pp->B; k.C = pp->C; k.D = pp->C + 5; k.na = pp->na; k.N = (int)N;
    // guard of modules/detection.py:76: rescale only if BOTH dimensions differ
    k.rescale = (pp->og_H > 0 && pp->og_W > 0 && pp->og_H != pp->H && pp->og_W != pp->W) ? 1 : 0;
    k.fW = (float)pp->W; k.fH = (float)pp->H; k.fW0 = (float)pp->og_W; k.fH0 = (float)pp->og_H;
    k.use_allowance = pp->box_allowance != 0.0f;
    k.allowance = pp->box_allowance;
    k.score_thr = pp->score_threshold;
    k.seg_count = p.seg_count; k.seg_off = p.seg_off; k.keys = p.keys;
    k.box_dense = w.box_dense; k.cls_dense = w.cls_dense;

    const int sms = num_sms();
    detect_init_kernel<<<(pp->B + 1 + 255) / 256, 256, 0, st>>>(p, pp->B, w.stride, out_counts);
    BG_LAUNCH_CHECK();

    const bool prof = g_prof_start && g_prof_stop;
    if (prof) cudaEventRecord(g_prof_start, st);
    int variant = pp->variant;
    if (variant == 0) variant = aligned ? 2 : 1;
    if (variant == 2 && !aligned) return BG_ERR_INVALID;
    if (variant == 2) {
        TileMap tm;
        int TR = (int)(TMA_TILE_BYTES / ((size_t)k.D * 4));
        TR = TR > TMA_THREADS ? TMA_THREADS : (TR / 4) * 4;  // multiple of 4 rows keeps every full tile 16-byte sized
        if (TR < 4) variant = 1;
        else {
            tm.TR = TR;
            tm.total = 0;
            for (int s = 0; s < 3; ++s) { tm.tiles[s] = (int)((k.sc[s].rows + TR - 1) / TR); tm.total += tm.tiles[s]; }
            const size_t smem = (size_t)TR * k.D * 4;
            static bool attr_set = false;
            if (!attr_set) {
                if (cudaFuncSetAttribute(decode_filter_tma_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_TILE_BYTES) != cudaSuccess ||
                    cudaFuncSetAttribute(decode_filter_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_TILE_BYTES) != cudaSuccess) {
                    (void)cudaGetLastError();
                    return BG_ERR_LAUNCH;
                }
                attr_set = true;
            }
            const int cap = sms * TMA_CTAS_PER_SM;
            const int grid = tm.total < cap ? tm.total : cap;
            if (k.C == 80) decode_filter_tma_kernel<80><<<grid, TMA_THREADS, smem, st>>>(k, tm);
            else decode_filter_tma_kernel<0><<<grid, TMA_THREADS, smem, st>>>(k, tm);
            BG_LAUNCH_CHECK();
        }
    }
    if (variant == 1) {
        decode_filter_warp_kernel<<<sms * 8, 256, 0, st>>>(k);
        BG_LAUNCH_CHECK();
    }
    if (prof) { cudaEventRecord(g_prof_stop, st); g_prof_start = g_prof_stop = nullptr; }
    int rc = segnms_run(p, pp->B, out_counts, 1, sms, st);
    if (rc != BG_OK) return rc;
    const int go = pp->B < 4096 ? pp->B : 4096;
    detect_output_kernel<<<dim3(pp->B > 256 ? 1 : 8, go), 256, 0, st>>>(p, k, pp->order, out_boxes, reinterpret_cast<long long *>(out_img),
                                             reinterpret_cast<long long *>(out_keep), out_counts);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

int bg_decode_scale(const float *raw, float *out, int32_t B, int32_t ny, int32_t nx, int32_t na, int32_t C,
                    const float *anchors, int32_t H, int32_t W, int32_t inference, int32_t og_H, int32_t og_W,
                    void *stream)
{
    if (!raw || !out || B <= 0 || ny <= 0 || nx <= 0 || na <= 0 || na > BG_MAX_ANCHORS || C <= 0) return BG_ERR_INVALID;
    if (inference && !anchors) return BG_ERR_INVALID;
    DecodeK k;
    memset(&k, 0, sizeof(k));
    k.raw = raw; k.out = out;
    k.rows = (long long)B * ny * nx * na;
    k.ny = ny; k.nx = nx; k.na = na; k.C = C; k.D = C + 5;
    k.inference = inference;
    k.rescale = (og_H > 0 && og_W > 0 && og_H != H && og_W != W) ? 1 : 0;
    k.s0 = (float)H / (float)ny; k.s1 = (float)W / (float)nx;
    k.fnx = (float)nx; k.fny = (float)ny;
    k.fW = (float)W; k.fH = (float)H; k.fW0 = (float)og_W; k.fH0 = (float)og_H;
    if (anchors) for (int a = 0; a < na; ++a) { k.aw[a] = anchors[2 * a]; k.ah[a] = anchors[2 * a + 1]; }
    decode_scale_kernel<<<num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(k);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

// ------------------------------------------------------------------------------------------ B1
static void assign_fill(AssignK &k, const float *targets, long long nt, int ny, int nx, const float *anchors, int na,
                        float anchor_t, float edge_t)
{
    memset(&k, 0, sizeof(k));
    k.targets = targets; k.nt = nt; k.ny = ny; k.nx = nx; k.na = na;
    k.fnx = (float)nx; k.fny = (float)ny;
    for (int a = 0; a < na; ++a) { k.aw[a] = anchors[2 * a] * (float)nx; k.ah[a] = anchors[2 * a + 1] * (float)ny; }
    k.anchor_t = anchor_t; k.edge_t = edge_t;
    k.ncand = 5ll * na * nt;
}

static int assign_launch(AssignK &k, cudaStream_t st)
{
    if (k.nt == 0) {
        if (cudaMemsetAsync(k.count, 0, sizeof(int), st) != cudaSuccess) return BG_ERR_LAUNCH;
        return BG_OK;
    }
    const int nblk = (int)((k.ncand + ASSIGN_THREADS - 1) / ASSIGN_THREADS);
    assign_count_kernel<<<nblk, ASSIGN_THREADS, 0, st>>>(k);
    BG_LAUNCH_CHECK();
    assign_emit_kernel<<<nblk, ASSIGN_THREADS, 0, st>>>(k);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

size_t bg_assign_workspace_bytes(int64_t nt, int32_t na)
{
    if (nt < 0 || na <= 0) return 0;
    const long long ncand = 5ll * na * nt;
    return align_up((size_t)((ncand + ASSIGN_THREADS - 1) / ASSIGN_THREADS + 1) * sizeof(int), 256);
}

int bg_assign_targets(const float *targets, int64_t nt, int32_t ny, int32_t nx, const float *anchors, int32_t na,
                      float anchor_t, float edge_t, int64_t *out_idx4, int64_t *out_cls, float *out_anchor,
                      float *out_box, int64_t cap, int32_t *out_count, void *workspace, size_t workspace_bytes,
                      void *stream)
{
    if (nt < 0 || ny <= 0 || nx <= 0 || na <= 0 || na > BG_MAX_ANCHORS || !anchors || !out_count) return BG_ERR_INVALID;
    if (5ll * na * nt >= (1ll << 31)) return BG_ERR_INVALID;
    if (nt > 0 && (!targets || !out_idx4 || !out_cls || !out_anchor || !out_box || !workspace || cap < 5ll * na * nt))
        return BG_ERR_INVALID;
    if (out_box && (((uintptr_t)out_box & 15) != 0)) return BG_ERR_INVALID;
    if (workspace_bytes < bg_assign_workspace_bytes(nt, na)) return BG_ERR_WORKSPACE;
    AssignK k;
    assign_fill(k, targets, nt, ny, nx, anchors, na, anchor_t, edge_t);
    k.block_counts = (int *)workspace;
    k.idx4 = reinterpret_cast<long long *>(out_idx4);
    k.cls64 = reinterpret_cast<long long *>(out_cls);
    k.anchor = out_anchor; k.box = out_box; k.cap = cap; k.count = out_count;
    return assign_launch(k, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------ B2
int bg_ciou_fwd(const float *p, const float *t, int64_t M, float eps, float *out, void *stream)
{
    if (M < 0) return BG_ERR_INVALID;
    if (M == 0) return BG_OK;
    if (!p || !t || !out || (((uintptr_t)p | (uintptr_t)t) & 15)) return BG_ERR_INVALID;
    ciou_fwd_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, t, M, eps, out);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

int bg_ciou_bwd(const float *p, const float *t, const float *go, int64_t M, float eps, float *gp, void *stream)
{
    if (M < 0) return BG_ERR_INVALID;
    if (M == 0) return BG_OK;
    if (!p || !t || !go || !gp || (((uintptr_t)p | (uintptr_t)t | (uintptr_t)gp) & 15)) return BG_ERR_INVALID;
    ciou_bwd_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, t, go, M, eps, gp);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

// ------------------------------------------------------------------------------------------ B3
namespace {
struct LossWs {
    int *block_counts[3];
    int *M[3];
    int *cell[3];
    int *cls[3];
    float *anchor[3];
    float *box[3];
    float *ciou[3];
    int *winner[3];
    double *part_match[3];
    double *part_dense[3];
    long long cap;
    long long cells[3];
    int nblk_match, nblk_dense;
};

bool loss_valid(const bg_loss_params *p)
{
    if (!p || p->B <= 0 || p->C <= 0 || p->na <= 0 || p->na > BG_MAX_ANCHORS || p->nt < 0) return false;
    if (5ll * p->na * p->nt >= (1ll << 31)) return false;
    for (int s = 0; s < 3; ++s) {
        if (p->ny[s] <= 0 || p->nx[s] <= 0) return false;
        if ((long long)p->B * p->ny[s] * p->nx[s] * p->na * (p->C + 5) >= (1ll << 40)) return false;
        if ((long long)p->B * p->ny[s] * p->nx[s] * p->na >= (1ll << 31)) return false;
    }
    return true;
}

size_t loss_carve(unsigned char *base, const bg_loss_params *p, LossWs &w)
{
    Bump b{base, 0};
    const int sms = num_sms();
    w.cap = 5ll * p->na * p->nt;
    if (w.cap < 1) w.cap = 1;
    w.nblk_match = sms * 8;
    w.nblk_dense = sms * 8;
    const size_t nblk_assign = (size_t)((w.cap + ASSIGN_THREADS - 1) / ASSIGN_THREADS + 1);
    for (int s = 0; s < 3; ++s) {
        w.cells[s] = (long long)p->B * p->ny[s] * p->nx[s] * p->na;
        w.block_counts[s] = b.take<int>(nblk_assign);
        w.M[s] = b.take<int>(1);
        w.cell[s] = b.take<int>(w.cap);
        w.cls[s] = b.take<int>(w.cap);
        w.anchor[s] = b.take<float>(2 * w.cap);
        w.box[s] = b.take<float>(4 * w.cap);
        w.ciou[s] = b.take<float>(w.cap);
        w.winner[s] = b.take<int>(w.cells[s]);
        w.part_match[s] = b.take<double>((size_t)w.nblk_match * 4);
        w.part_dense[s] = b.take<double>((size_t)w.nblk_dense * 3);
    }
    return align_up(b.off, 256);
}

void loss_fill(LossScaleK &k, const bg_loss_params *p, const LossWs &w, int s, const float *preds, float *grad)
{
    memset(&k, 0, sizeof(k));
    k.preds = preds; k.grad = grad; k.cells = w.cells[s]; k.C = p->C; k.D = p->C + 5;
    k.M = w.M[s]; k.cell = w.cell[s]; k.cls = w.cls[s]; k.anchor = w.anchor[s]; k.box = w.box[s];
    k.ciou = w.ciou[s]; k.winner = w.winner[s]; k.part_match = w.part_match[s]; k.part_dense = w.part_dense[s];
    k.cn = 0.5f * p->label_smoothing;  // python: cn = 0.5*ls (double), written into an fp32 tensor
    k.cn = (float)(0.5 * (double)p->label_smoothing);
    k.cp = (float)(1.0 - 0.5 * (double)p->label_smoothing);
    k.nblk_match = w.nblk_match; k.nblk_dense = w.nblk_dense;
}
}  // namespace

size_t bg_loss_workspace_bytes(const bg_loss_params *p)
{
    if (!loss_valid(p)) return 0;
    LossWs w;
    return loss_carve(nullptr, p, w);
}

int bg_loss_fwd(const float *preds_sm, const float *preds_md, const float *preds_lg, const float *targets,
                const bg_loss_params *p, double *out_scalars, int64_t *out_hist, void *workspace,
                size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!loss_valid(p) || !preds_sm || !preds_md || !preds_lg || !out_scalars || !out_hist || !workspace) return BG_ERR_INVALID;
    if (p->nt > 0 && !targets) return BG_ERR_INVALID;
    LossWs w;
    if (loss_carve((unsigned char *)workspace, p, w) > workspace_bytes) return BG_ERR_WORKSPACE;
    const float *preds[3] = {preds_sm, preds_md, preds_lg};
    if (cudaMemsetAsync(out_hist, 0, sizeof(int64_t) * 9 * (size_t)p->C, st) != cudaSuccess) return BG_ERR_LAUNCH;
    for (int s = 0; s < 3; ++s) {
        AssignK a;
        float anc[2 * BG_MAX_ANCHORS];
        for (int q = 0; q < p->na; ++q) { anc[2 * q] = p->anchors[s][q][0]; anc[2 * q + 1] = p->anchors[s][q][1]; }
        assign_fill(a, targets, p->nt, p->ny[s], p->nx[s], anc, p->na, p->anchor_t, p->edge_t);
        a.block_counts = w.block_counts[s];
        a.anchor = w.anchor[s]; a.box = w.box[s]; a.cell = w.cell[s]; a.cls32 = w.cls[s];
        a.cap = w.cap; a.count = w.M[s];
        int rc = assign_launch(a, st);
        if (rc != BG_OK) return rc;
        if (cudaMemsetAsync(w.winner[s], 0xff, sizeof(int) * (size_t)w.cells[s], st) != cudaSuccess) return BG_ERR_LAUNCH;
        LossScaleK k;
        loss_fill(k, p, w, s, preds[s], nullptr);
        k.hist = reinterpret_cast<long long *>(out_hist) + (size_t)s * 3 * p->C;
        k.scalars = out_scalars + 8 * s;
        loss_match_kernel<<<w.nblk_match, LOSS_THREADS, 0, st>>>(k);
        BG_LAUNCH_CHECK();
        loss_dense_kernel<<<w.nblk_dense, LOSS_THREADS, 0, st>>>(k);
        BG_LAUNCH_CHECK();
        loss_finalize_kernel<<<1, 256, 0, st>>>(k);
        BG_LAUNCH_CHECK();
    }
    return BG_OK;
}

int bg_loss_bwd(const float *preds_sm, const float *preds_md, const float *preds_lg, const bg_loss_params *p,
                float grad_out, float *grad_sm, float *grad_md, float *grad_lg, void *workspace,
                size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!loss_valid(p) || !preds_sm || !preds_md || !preds_lg || !grad_sm || !grad_md || !grad_lg || !workspace)
        return BG_ERR_INVALID;
    LossWs w;
    if (loss_carve((unsigned char *)workspace, p, w) > workspace_bytes) return BG_ERR_WORKSPACE;
    const float *preds[3] = {preds_sm, preds_md, preds_lg};
    float *grads[3] = {grad_sm, grad_md, grad_lg};
    const int sms = num_sms();
    for (int s = 0; s < 3; ++s) {
        LossScaleK k;
        loss_fill(k, p, w, s, preds[s], grads[s]);
        const double sw = (double)p->scale_w[s] * (double)grad_out;
        k.w_box = (double)p->box_w * sw;
        k.w_conf = (double)p->conf_w * sw;
        k.w_cls = (double)p->class_w * sw;
        loss_bwd_dense_kernel<<<sms * 16, 256, 0, st>>>(k);
        BG_LAUNCH_CHECK();
        loss_bwd_match_kernel<<<sms * 8, LOSS_THREADS, 0, st>>>(k);
        BG_LAUNCH_CHECK();
    }
    return BG_OK;
}

// ------------------------------------------------------------------------------------------ a13
int bg_ratio_metrics(const float *wh, int64_t n, const float *anchors, int32_t kk, float threshold, double *out3,
                     void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 0 || kk <= 0 || kk > 32 || !anchors || !out3 || (n > 0 && !wh)) return BG_ERR_INVALID;
    if (cudaMemsetAsync(out3, 0, 3 * sizeof(double), st) != cudaSuccess) return BG_ERR_LAUNCH;
    if (n == 0) return BG_OK;
    RatioK k;
    k.wh = wh; k.n = n; k.k = kk; k.out = out3;
    for (int j = 0; j < kk; ++j) { k.aw[j] = anchors[2 * j]; k.ah[j] = anchors[2 * j + 1]; }
    k.inv_thr = (float)(1.0 / (double)threshold);  // python double 1/threshold, cast to fp32 by the comparison
    int grid = (int)((n + 255) / 256);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    ratio_metrics_kernel<<<grid, 256, 0, st>>>(k);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

}  // extern "C"
