// libboxgeom.so -- host side of the C ABI declared in include/boxgeom.h.
// Single translation unit (the kernels live in the *_kernels.cuh headers), built for sm_100a only:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <math.h>
#include <stdlib.h>
#include <atomic>
#include <string.h>

#include "common.cuh"
#include "nms_kernels.cuh"
#include "detect_kernels.cuh"
#include "imgnms_kernels.cuh"
#include "train_kernels.cuh"
#include "loss_kernels.cuh"
#include "seg_kernels.cuh"

namespace bg {

std::atomic<unsigned long long> g_launches{0};
// profiling hooks are armed per host thread (the thread that arms one is the thread whose next call sees it), so
// concurrent callers on other threads are unaffected and the entry points stay re-entrant
static thread_local cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;
// (armed from the caller's thread, consumed by bg_loss_bwd on whichever thread runs it -- autograd's worker thread under torch)
static std::atomic<cudaEvent_t> g_prof_loss_start{nullptr}, g_prof_loss_stop{nullptr};
static thread_local unsigned long long *g_prof_stamps = nullptr, *g_prof_cycles = nullptr;

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

// launch behind the previous kernel of the stream with programmatic stream serialization: the grid may become
// resident while its predecessor drains; the kernel itself calls griddepcontrol.wait before touching upstream data
template <typename K>
int launch_after(void (*kern)(K), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const K &arg, bool after_kernel = true)
{
    static const bool pdl_env = []() { const char *e = getenv("BG_PDL"); return !(e && e[0] == '0'); }();
    const bool pdl = pdl_env && after_kernel;  // only behind a kernel of ours (not behind a memset)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, kern, arg) != cudaSuccess) { (void)cudaGetLastError(); return BG_ERR_LAUNCH; }
    ++g_launches;
    return BG_OK;
}

static int cur_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); dev = 0; }
    return (dev >= 0 && dev < 64) ? dev : 0;
}

static int num_sms()
{
    static int n[64] = {0};
    const int dev = cur_device();
    if (!n[dev]) {
        if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0) {
            (void)cudaGetLastError();
            n[dev] = 148;
        }
    }
    return n[dev];
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
    p.bwh = b.take<u32>(elems);
    p.edge_off = b.take<long long>(S_max + 1);
    p.cell_off = b.take<long long>(S_max + 1);
    p.cell_start = b.take<int>(elems / 2 + 2 * S_max + 2);  // sum of G*G + 1 with G*G <= max(1, count / 2)
    p.grid = b.take<SegGrid>(S_max);
    p.edge_count = b.take<unsigned long long>(S_max);
    p.gstate = b.take<unsigned char>(2 * elems);
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

constexpr int DET_SPILL_EDGES = 1 << 17;  // overlap edges per image that may spill past the shared-memory list

struct DetWs {
    // decode stage (both NMS paths): per-tile survivor slots
    int *tile_count;
    u64 *slot_keys;
    float4 *box_slots;
    int *cls_slots;
    // per-image NMS path
    FusedHdr *hdr;
    u64 *chain;
    long long *f_seg_off;
    int *f_emit_count;
    int *cand_count;
    u32 *gflag;           // split mode hand-over (helper CTA -> main CTA)
    u32 *gedges;
    u32 *gspill;
    u64 *gsorted;
    float4 *gboxp;
    u64 *gkeyp;           // lean per-image kernel: keys / classes by survivor number
    u32 *gclsp;
    u64 *f_emit_key;      // globally ordered output only
    float4 *f_emit_box;
    int *f_emit_cls;
    // general path: boxes / classes by candidate index + the segmented engine
    float4 *box_dense;
    int *cls_dense;
    SegNms p;
    long long stride;
};
static size_t det_carve(unsigned char *base, int B, long long N, int tiles_per_image, bool general, bool global_order,
                        size_t mask_bytes, DetWs &w)
{
    Bump b{base, 0};
    memset(&w, 0, sizeof(w));
    w.tile_count = b.take<int>((size_t)B * tiles_per_image);
    w.slot_keys = b.take<u64>((size_t)B * N);
    w.box_slots = b.take<float4>((size_t)B * N);
    w.cls_slots = b.take<int>((size_t)B * N);
    w.hdr = b.take<FusedHdr>(1);
    w.chain = b.take<u64>((size_t)B + 1);
    w.f_seg_off = b.take<long long>((size_t)B + 1);
    w.f_emit_count = b.take<int>(B);
    w.cand_count = b.take<int>(B);
    w.gflag = b.take<u32>(4 * (size_t)B);
    if (!general) {
        w.gedges = b.take<u32>((size_t)B * INMS_HCAP);
        w.gspill = b.take<u32>((size_t)B * DET_SPILL_EDGES);
        w.gsorted = b.take<u64>((size_t)B * INMS_CAP_MAX);
        w.gboxp = b.take<float4>((size_t)B * INMS_CAP_MAX);
        w.gkeyp = b.take<u64>((size_t)B * InmsLean::CAP);
        w.gclsp = b.take<u32>((size_t)B * InmsLean::CAP);
    }
    w.stride = (long long)next_pow2((u32)N);
    if (general) {
        w.box_dense = b.take<float4>((size_t)B * N);
        w.cls_dense = b.take<int>((size_t)B * N);
        carve_seg(b, w.p, B, (long long)B * w.stride, mask_bytes);
    } else if (global_order) {
        w.f_emit_key = b.take<u64>((size_t)B * N);
        w.f_emit_box = b.take<float4>((size_t)B * N);
        w.f_emit_cls = b.take<int>((size_t)B * N);
    }
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
    if (p->extra_cols < 0 || p->extra_cols > 4096) return false;
    for (int s = 0; s < 3; ++s) {
        if (p->ny[s] <= 0 || p->nx[s] <= 0) return false;
        // the decode kernel divides by na and nx with 32-bit magic numbers: exact while n * d < 2^32
        if ((long long)p->ny[s] * p->nx[s] * p->nx[s] >= (1ll << 32)) return false;
        if ((long long)p->ny[s] * p->nx[s] * p->na * p->na >= (1ll << 32)) return false;
    }
    const long long N = det_candidates(p);
    return N > 0 && N < (1ll << 31) && (long long)p->B * N < (1ll << 31);
}

static TilePlan det_tile_plan(const bg_detect_params *p)
{
    TilePlan tp;
    const int D = p->C + 5 + p->extra_cols;
    int TR = (int)(DEC_TILE_BYTES / ((size_t)D * 4));
    TR = TR > DEC_THREADS ? DEC_THREADS : (TR / 4) * 4;  // multiple of 4 rows keeps every full tile 16-byte sized
    if (TR < 4) TR = 4;
    tp.TR = TR;
    tp.tpi_total = 0;
    for (int s = 0; s < 3; ++s) {
        const long long cells_na = (long long)p->ny[s] * p->nx[s] * p->na;
        tp.tpi[s] = (int)((cells_na + TR - 1) / TR);
        tp.tpi_total += tp.tpi[s];
    }
    tp.total = p->B * tp.tpi_total;
    return tp;
}

// 1 = general segmented engine, 0 = one CTA per image (needs a positive threshold for the reach bound and
// a tile list that fits the kernel's shared-memory prefix table)
static int det_nms_path(const bg_detect_params *p, const TilePlan &tp)
{
    if (p->nms_path == 1) return 1;
    const IouThr t = make_iou_thr(p->iou_threshold);
    const bool ok = t.fast_ok && !t.zero_suppresses && t.tdn >= 0.05f && t.tdn < 1.0f && tp.tpi_total < INMS_MAXT &&
                    det_candidates(p) <= InmsLarge::MAX_N && p->C <= 65535;
    if (p->nms_path >= 2 && p->nms_path <= 5) return ok ? 0 : -1;
    return ok ? 0 : 1;
}

static bool det_plan_valid(const bg_detect_params *p, const TilePlan &tp)
{
    return (size_t)tp.TR * (p->C + 5 + p->extra_cols) * 4 <= (size_t)DEC_TILE_BYTES && (long long)p->B * tp.tpi_total < (1ll << 31);
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

int bg_version(void) { return 201; }  // 200: bg_loss_* take the input form (decoded / raw / raw split) and per-scale pointer triples

uint64_t bg_launch_count(void) { return g_launches; }
void bg_profile_events(void *start, void *stop) { g_prof_start = (cudaEvent_t)start; g_prof_stop = (cudaEvent_t)stop; }
void bg_profile_events_loss(void *start, void *stop) { g_prof_loss_start = (cudaEvent_t)start; g_prof_loss_stop = (cudaEvent_t)stop; }
void bg_profile_stamps(void *dev_buf) { g_prof_stamps = (unsigned long long *)dev_buf; }
void bg_profile_decode_cycles(void *dev_buf) { g_prof_cycles = (unsigned long long *)dev_buf; }
int bg_profile_stamps_per_image(void) { return INMS_STAMPS; }
size_t bg_sizeof_detect_params(void) { return sizeof(bg_detect_params); }
void *bg_host_mapped_ptr(void *host_ptr)
{
    void *d = nullptr;
    if (!host_ptr || cudaHostGetDevicePointer(&d, host_ptr, 0) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return d;
}
size_t bg_sizeof_loss_params(void) { return sizeof(bg_loss_params); }
size_t bg_sizeof_seg_params(void) { return sizeof(bg_seg_params); }

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
    // global order: gather the per-segment lists, merge them pairwise level by level (ping-pong between two key
    // arrays the engine no longer needs), write the candidate indices
    const int go = S_launch < 1 ? 1 : (S_launch > 4096 ? 4096 : S_launch);
    u64 *ping = p.bkeys, *pong = p.keys;
    gnms_gather_kernel<<<dim3(S_launch > 256 ? 1 : 8, go), 256, 0, st>>>(p, ping);
    BG_LAUNCH_CHECK();
    for (int L = 0; (1ll << L) < S_launch; ++L) {
        gnms_merge_level_kernel<<<gs, 256, 0, st>>>(p, (L & 1) ? pong : ping, (L & 1) ? ping : pong, L);
        BG_LAUNCH_CHECK();
    }
    gnms_output_kernel<<<gs, 256, 0, st>>>(p, ping, pong, reinterpret_cast<long long *>(out_keep));
    BG_LAUNCH_CHECK();
    return BG_OK;
}

// ------------------------------------------------------------------------------------------ B5
size_t bg_detect_workspace_bytes(const bg_detect_params *p, size_t mask_bytes)
{
    if (!det_valid(p)) return 0;
    const TilePlan tp = det_tile_plan(p);
    const int path = det_nms_path(p, tp);
    if (path < 0 || !det_plan_valid(p, tp)) return 0;
    DetWs w;
    return det_carve(nullptr, p->B, det_candidates(p), tp.tpi_total, path == 1, p->order != 0, mask_bytes, w);
}

}  // extern "C"

// geometry / decode parameters of the kernel argument block; returns whether the three inputs are 16-byte aligned
static bool det_fill_geometry(DetectK &k, const float *const raws[3], int predecoded, const bg_detect_params *pp, long long N)
{
    memset(&k, 0, sizeof(k));
    int off = 0;
    bool aligned = true;
    for (int s = 0; s < 3; ++s) {
        ScaleDesc &d = k.sc[s];
        d.raw = raws[s];
        d.ny = pp->ny[s]; d.nx = pp->nx[s];
        d.cells_na = pp->ny[s] * pp->nx[s] * pp->na;
        d.img_stride = predecoded ? N : d.cells_na;
        d.img_off = off;
        off += d.cells_na;
        d.rows = (long long)pp->B * d.cells_na;
        d.s0 = (float)pp->H / (float)pp->ny[s];
        d.s1 = (float)pp->W / (float)pp->nx[s];
        d.fnx = (float)pp->nx[s]; d.fny = (float)pp->ny[s];
        d.magic_nx = (u32)(((1ull << 32) + (u64)pp->nx[s] - 1) / (u64)pp->nx[s]);
        for (int a = 0; a < pp->na; ++a) { d.aw[a] = pp->anchors[s][a][0]; d.ah[a] = pp->anchors[s][a][1]; }
        aligned = aligned && (((uintptr_t)raws[s] & 15) == 0);
    }
    k.B = pp->B; k.C = pp->C; k.D = pp->C + 5 + pp->extra_cols; k.na = pp->na; k.N = (int)N;
    k.magic_na = (u32)(((1ull << 32) + (u64)pp->na - 1) / (u64)pp->na);
    k.predecoded = predecoded;
    // guard of modules/detection.py:76: rescale only if BOTH dimensions differ (already applied to decoded rows)
    k.rescale = (!predecoded && pp->og_H > 0 && pp->og_W > 0 && pp->og_H != pp->H && pp->og_W != pp->W) ? 1 : 0;
    k.fW = (float)pp->W; k.fH = (float)pp->H; k.fW0 = (float)pp->og_W; k.fH0 = (float)pp->og_H;
    k.use_allowance = pp->box_allowance != 0.0f;
    k.allowance = pp->box_allowance;
    k.score_thr = pp->score_threshold;
    return aligned;
}

// shared body of bg_detect (three raw head tensors) and bg_post_process (one decoded [B,N,D] tensor)
static int detect_impl(const float *raw_sm, const float *raw_md, const float *raw_lg, int predecoded, const bg_detect_params *pp,
                       float *out_boxes, int64_t *out_img, int64_t *out_keep, int32_t *out_counts, void *workspace,
                       size_t workspace_bytes, size_t mask_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!det_valid(pp) || !raw_sm || !raw_md || !raw_lg || !out_boxes || !out_img || !out_keep || !out_counts || !workspace)
        return BG_ERR_INVALID;
    const long long N = det_candidates(pp);
    const TilePlan tp = det_tile_plan(pp);
    const int path = det_nms_path(pp, tp);
    if (path < 0 || !det_plan_valid(pp, tp)) return BG_ERR_INVALID;
    DetWs w;
    if (det_carve((unsigned char *)workspace, pp->B, N, tp.tpi_total, path == 1, pp->order != 0, mask_bytes, w) > workspace_bytes)
        return BG_ERR_WORKSPACE;

    DetectK k;
    const float *raws[3] = {raw_sm, raw_md, raw_lg};
    const bool aligned = det_fill_geometry(k, raws, predecoded, pp, N);
    k.keys = w.slot_keys; k.box_slots = w.box_slots; k.cls_slots = w.cls_slots;
    k.box_dense = w.box_dense; k.cls_dense = w.cls_dense;

    const int sms = num_sms();
    static bool attr_set[64] = {false};  // function attributes are per device
    const int dev = cur_device();
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(decode_filter_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_STAGES * DEC_TILE_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(decode_filter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_STAGES * DEC_TILE_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(image_nms_kernel<InmsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ImgNmsSmem<InmsSmall>)) != cudaSuccess ||
            cudaFuncSetAttribute(image_nms_kernel<InmsLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ImgNmsSmem<InmsLarge>)) != cudaSuccess ||
            // the lean NMS CTA shares an SM with two decode CTAs only if the SM keeps its full 228 KB of shared memory:
            // a smaller carve-out chosen for the decode kernel alone (196 KB) could not be changed while its CTAs run
            cudaFuncSetAttribute(decode_filter_kernel<80>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess ||
            cudaFuncSetAttribute(decode_filter_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess ||
            cudaFuncSetAttribute(image_nms_kernel<InmsLean>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) {
            (void)cudaGetLastError();
            return BG_ERR_LAUNCH;
        }
        attr_set[dev] = true;
    }

    // ---- decode + score filter: variant 1 forces plain loads, variant 2 insists on the TMA pipeline ----
    if (pp->variant == 2 && !aligned) return BG_ERR_INVALID;
    DecodeOut o;
    o.tile_count = w.tile_count; o.hdr = w.hdr; o.chain = w.chain; o.seg_off = w.f_seg_off;
    o.force_plain = (pp->variant == 1 || !aligned) ? 1 : 0;
    o.gflag = w.gflag;
    o.cycles = g_prof_cycles;
    const bool prof = g_prof_start && g_prof_stop;
    if (prof) cudaEventRecord(g_prof_start, st);
    {
        const int cap = sms * DEC_CTAS_PER_SM;
        const int grid = tp.total < cap ? tp.total : cap;
        const size_t smem = (size_t)DEC_STAGES * tp.TR * k.D * 4;
        if (k.C == 80 && k.D == 85) decode_filter_kernel<80><<<grid, DEC_THREADS, smem, st>>>(k, tp, o);
        else decode_filter_kernel<0><<<grid, DEC_THREADS, smem, st>>>(k, tp, o);
        BG_LAUNCH_CHECK();
    }
    if (prof) { cudaEventRecord(g_prof_stop, st); g_prof_start = g_prof_stop = nullptr; }

    // the NMS kernels on the caller's second (higher-priority) stream, ordered behind the decode by the caller's event
    cudaStream_t st_dec = st;
    const bool two_streams = pp->nms_stream != nullptr && pp->nms_event != nullptr;
    if (two_streams) {
        if (cudaEventRecord((cudaEvent_t)pp->nms_event, st) != cudaSuccess ||
            cudaStreamWaitEvent((cudaStream_t)pp->nms_stream, (cudaEvent_t)pp->nms_event, 0) != cudaSuccess) {
            (void)cudaGetLastError();
            return BG_ERR_LAUNCH;
        }
        st = (cudaStream_t)pp->nms_stream;
    }
    auto rejoin = [&](bool flag_written = false) -> int {   // `stream` continues behind the NMS kernels
        if (pp->host_flag && !flag_written) {   // (the per-image kernel in image order stores the flag itself)
            host_flag_kernel<<<1, 1, 0, st>>>((int32_t *)pp->host_flag, pp->host_flag_value);
            BG_LAUNCH_CHECK();
        }
        if (!two_streams) return BG_OK;
        if (cudaEventRecord((cudaEvent_t)pp->nms_event, st) != cudaSuccess ||
            cudaStreamWaitEvent(st_dec, (cudaEvent_t)pp->nms_event, 0) != cudaSuccess) {
            (void)cudaGetLastError();
            return BG_ERR_LAUNCH;
        }
        return BG_OK;
    };

    if (path == 0) {
        // ---- one CTA per image: sort, grid-pruned pair tests, greedy resolution, rows ----
        ImgNmsK q;
        memset(&q, 0, sizeof(q));
        q.B = pp->B; q.N = (int)N; q.TR = tp.TR; q.tpi_total = tp.tpi_total;
        for (int s = 0; s < 3; ++s) { q.tpi[s] = tp.tpi[s]; q.img_off[s] = k.sc[s].img_off; }
        q.tile_count = w.tile_count; q.keys = w.slot_keys; q.box_slots = w.box_slots; q.cls_slots = w.cls_slots;
        q.thr = make_iou_thr(pp->iou_threshold);
        q.reach = (1.0f - q.thr.tdn) / q.thr.tdn * 1.01f + 0.01f;
        q.n_tracked = pp->n_tracked;
        for (int i = 0; i < pp->n_tracked; ++i) q.tracked[i] = pp->tracked[i];
        q.hdr = w.hdr; q.chain = w.chain; q.order = pp->order;
        q.emit_key = w.f_emit_key; q.emit_box = w.f_emit_box; q.emit_cls = w.f_emit_cls;
        q.emit_count = w.f_emit_count; q.cand_count = w.cand_count;
        q.out_boxes = out_boxes; q.out_img = reinterpret_cast<long long *>(out_img);
        q.out_keep = reinterpret_cast<long long *>(out_keep); q.out_counts = out_counts;
        q.stamps = g_prof_stamps;
        q.host_flag = (int32_t *)pp->host_flag; q.host_flag_value = pp->host_flag_value;
        {   // two CTAs per image (helper + main) while every CTA of the grid can be resident at once
            static const int split_env = []() { const char *e = getenv("BG_NMS_SPLIT"); return e ? atoi(e) : -1; }();
            // (throughput mode, 1024-thread kernel: less total SM time matters more than latency; the lean kernel runs next
            // to the decode CTAs either way, and a shorter NMS shortens what is left when the last batch's decode ends)
            const bool can_split = pp->nms_path != 3 && (!pp->throughput || pp->nms_path == 5);
            q.split = !can_split ? 0 : (split_env >= 0 ? (split_env != 0) : (2 * pp->B <= sms ? 1 : 0));
        }
        q.gflag = w.gflag; q.gedges = w.gedges; q.gsorted = w.gsorted; q.gspill = w.gspill; q.gcap = DET_SPILL_EDGES;
        q.gboxp = w.gboxp; q.gkeyp = w.gkeyp; q.gclsp = w.gclsp;
        const bool large = pp->nms_path == 4;  // up to 8,192 survivors per image, boxes in L2 instead of shared memory
        // up to 2,048 survivors; a CTA small enough to run next to the decode CTAs of other streams (one tile count per thread)
        const bool lean = pp->nms_path == 5 && tp.tpi_total < InmsLean::MAXT;
        {   // programmatic dependent launch: the CTAs become resident while the decode kernel drains
            static const bool pdl = []() { const char *e = getenv("BG_PDL"); return !(e && e[0] == '0'); }();
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(pp->B * (q.split ? 2 : 1)); cfg.blockDim = dim3(lean ? InmsLean::THREADS : INMS_THREADS);
            cfg.dynamicSmemBytes = lean ? sizeof(ImgNmsSmem<InmsLean>) : large ? sizeof(ImgNmsSmem<InmsLarge>) : sizeof(ImgNmsSmem<InmsSmall>);
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            // (throughput mode: CTAs resident early would hold SMs the other streams' decode kernels can use)
            static const bool lean_pdl = []() { const char *e = getenv("BG_LEAN_PDL"); return e && e[0] == '1'; }();
            cfg.attrs = at; cfg.numAttrs = (pdl && !two_streams && (!pp->throughput || (lean && lean_pdl))) ? 1 : 0;
            const cudaError_t le = lean    ? cudaLaunchKernelEx(&cfg, image_nms_kernel<InmsLean>, q)
                                   : large ? cudaLaunchKernelEx(&cfg, image_nms_kernel<InmsLarge>, q)
                                           : cudaLaunchKernelEx(&cfg, image_nms_kernel<InmsSmall>, q);
            if (le != cudaSuccess) { (void)cudaGetLastError(); return BG_ERR_LAUNCH; }
            ++g_launches;
        }
        if (pp->order) {
            SegNms v;
            memset(&v, 0, sizeof(v));
            v.emit_count = w.f_emit_count; v.emit_key = w.f_emit_key; v.seg_off = w.f_seg_off; v.box_seg_stride = N;
            const int go = pp->B < 4096 ? pp->B : 4096;
            detect_output_kernel<<<dim3(pp->B > 256 ? 1 : 8, go), 256, 0, st>>>(v, k, 1, nullptr, w.f_emit_box, w.f_emit_cls, out_boxes, reinterpret_cast<long long *>(out_img),
                                                                           reinterpret_cast<long long *>(out_keep), out_counts);
            BG_LAUNCH_CHECK();
        }
        return rejoin(pp->order == 0);
    }

    // ---- general path: segmented engine over the compacted survivor lists ----
    SegNms &p = w.p;
    p.boxes = w.box_dense;
    p.box_seg_stride = N;
    p.cls = w.cls_dense;
    p.n_tracked = pp->n_tracked;
    for (int i = 0; i < pp->n_tracked; ++i) p.tracked[i] = pp->tracked[i];
    p.thr = make_iou_thr(pp->iou_threshold);
    segnms_configure(p);
    detect_compact_kernel<<<pp->B < 2 * sms ? pp->B : 2 * sms, 1024, 0, st>>>(p, k, tp, w.tile_count, w.stride, out_counts);
    BG_LAUNCH_CHECK();
    int rc = segnms_run(p, pp->B, out_counts, 1, sms, st);
    if (rc != BG_OK) return rc;
    const int go = pp->B < 4096 ? pp->B : 4096;
    detect_output_kernel<<<dim3(pp->B > 256 ? 1 : 8, go), 256, 0, st>>>(p, k, pp->order, p.seg_count, nullptr, nullptr, out_boxes, reinterpret_cast<long long *>(out_img),
                                                                   reinterpret_cast<long long *>(out_keep), out_counts);
    BG_LAUNCH_CHECK();
    return rejoin();
}

extern "C" {

int bg_detect(const float *raw_sm, const float *raw_md, const float *raw_lg, const bg_detect_params *pp,
              float *out_boxes, int64_t *out_img, int64_t *out_keep, int32_t *out_counts, void *workspace,
              size_t workspace_bytes, size_t mask_bytes, void *stream)
{
    return detect_impl(raw_sm, raw_md, raw_lg, 0, pp, out_boxes, out_img, out_keep, out_counts, workspace, workspace_bytes,
                       mask_bytes, stream);
}

int bg_post_process(const float *preds, const bg_detect_params *pp, float *out_boxes, int64_t *out_img, int64_t *out_keep,
                    int32_t *out_counts, void *workspace, size_t workspace_bytes, size_t mask_bytes, void *stream)
{
    if (!det_valid(pp) || !preds) return BG_ERR_INVALID;
    const long long D = pp->C + 5 + pp->extra_cols;
    const long long n0 = (long long)pp->ny[0] * pp->nx[0] * pp->na, n1 = (long long)pp->ny[1] * pp->nx[1] * pp->na;
    return detect_impl(preds, preds + n0 * D, preds + (n0 + n1) * D, 1, pp, out_boxes, out_img, out_keep, out_counts, workspace,
                       workspace_bytes, mask_bytes, stream);
}

int bg_decode_scale_ex(const float *raw, float *out, int32_t B, int32_t ny, int32_t nx, int32_t na, int32_t C,
                       int32_t extra_cols, int32_t tanh_cols, const float *anchors, int32_t H, int32_t W, int32_t inference,
                       int32_t og_H, int32_t og_W, void *stream)
{
    if (!raw || !out || B <= 0 || ny <= 0 || nx <= 0 || na <= 0 || na > BG_MAX_ANCHORS || C <= 0) return BG_ERR_INVALID;
    if (extra_cols < 0 || tanh_cols < 0 || tanh_cols > extra_cols) return BG_ERR_INVALID;
    if (inference && !anchors) return BG_ERR_INVALID;
    DecodeK k;
    memset(&k, 0, sizeof(k));
    k.raw = raw; k.out = out;
    k.rows = (long long)B * ny * nx * na;
    k.ny = ny; k.nx = nx; k.na = na; k.C = C; k.D = C + 5 + extra_cols;
    k.tanh_cols = tanh_cols;
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

int bg_decode_scale(const float *raw, float *out, int32_t B, int32_t ny, int32_t nx, int32_t na, int32_t C,
                    const float *anchors, int32_t H, int32_t W, int32_t inference, int32_t og_H, int32_t og_W,
                    void *stream)
{
    return bg_decode_scale_ex(raw, out, B, ny, nx, na, C, 0, 0, anchors, H, W, inference, og_H, og_W, stream);
}

int bg_decode_rows(const float *raw_sm, const float *raw_md, const float *raw_lg, const bg_detect_params *pp, const int64_t *idx,
                   int64_t n, float *out, void *stream)
{
    if (n < 0 || !det_valid(pp)) return BG_ERR_INVALID;
    if (n == 0) return BG_OK;
    if (!raw_sm || !raw_md || !raw_lg || !idx || !out) return BG_ERR_INVALID;
    DetectK k;
    const float *raws[3] = {raw_sm, raw_md, raw_lg};
    det_fill_geometry(k, raws, 0, pp, det_candidates(pp));
    long long blocks = (n * 32 + 255) / 256;  // one warp per row
    decode_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(k, reinterpret_cast<const long long *>(idx), n, out);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

int bg_bbox_to_size(float *pred, int64_t rows, int32_t C, int32_t D, const int64_t *from4, const int64_t *to4, void *stream)
{
    if (rows < 0 || C <= 0 || D < C + 5) return BG_ERR_INVALID;
    if (rows == 0) return BG_OK;
    if (!pred || !from4 || !to4) return BG_ERR_INVALID;
    long long blocks = (rows * 4 + 255) / 256;
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    bbox_to_size_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred, rows, C, D, reinterpret_cast<const long long *>(from4),
                                                                      reinterpret_cast<const long long *>(to4));
    BG_LAUNCH_CHECK();
    return BG_OK;
}

int bg_decode_train_bwd_ex(const float *raw, const float *grad_out, float *grad_raw, int64_t rows, int32_t C, int32_t extra_cols,
                           int32_t tanh_cols, void *stream)
{
    if (rows < 0 || C <= 0 || extra_cols < 0 || tanh_cols < 0 || tanh_cols > extra_cols) return BG_ERR_INVALID;
    if (rows == 0) return BG_OK;
    if (!raw || !grad_out || !grad_raw) return BG_ERR_INVALID;
    decode_train_bwd_kernel<<<num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(raw, grad_out, grad_raw, rows, C, extra_cols, tanh_cols);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

int bg_decode_train_bwd(const float *raw, const float *grad_out, float *grad_raw, int64_t rows, int32_t C, void *stream)
{
    return bg_decode_train_bwd_ex(raw, grad_out, grad_raw, rows, C, 0, 0, stream);
}

// ------------------------------------------------------------------------------------------ B1
static void assign_fill(AssignK &k, const float *targets, long long nt, int ny, int nx, const float *anchors, int na,
                        float anchor_t, float edge_t)
{
    memset(&k, 0, sizeof(k));
    k.targets = targets; k.row_stride = 6; k.nt = nt; k.ny = ny; k.nx = nx; k.na = na;
    k.fnx = (float)nx; k.fny = (float)ny;
    for (int a = 0; a < na; ++a) { k.aw[a] = anchors[2 * a] * (float)nx; k.ah[a] = anchors[2 * a + 1] * (float)ny; }
    k.anchor_t = anchor_t; k.edge_t = edge_t;
    k.ncand = 5ll * na * nt;
}

// `n` scales (all with the same target list, hence the same candidate count) in one launch; the caller has
// zeroed the chain words
static int assign_launch(Assign3K &kk, int n, cudaStream_t st)
{
    const AssignK &k = kk.a[0];
    if (k.nt == 0) {
        for (int s = 0; s < n; ++s)
            if (cudaMemsetAsync(kk.a[s].count, 0, sizeof(int), st) != cudaSuccess) return BG_ERR_LAUNCH;
        return BG_OK;
    }
    const int nblk = (int)((k.ncand + ASSIGN_BLOCK - 1) / ASSIGN_BLOCK);
    assign_onepass_kernel<<<dim3(nblk, n), ASSIGN_THREADS, 0, st>>>(kk);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

static size_t assign_chain_words(long long nt, int na)
{
    const long long ncand = 5ll * na * nt;
    return (size_t)((ncand + ASSIGN_BLOCK - 1) / ASSIGN_BLOCK + 2);
}

size_t bg_assign_workspace_bytes(int64_t nt, int32_t na)
{
    if (nt < 0 || na <= 0) return 0;
    return align_up(assign_chain_words(nt, na) * sizeof(u64), 256);
}

size_t bg_assign_ex_workspace_bytes(int64_t nt, int32_t na, int32_t batch_size)
{
    if (nt < 0 || na <= 0 || batch_size < 0) return 0;
    return bg_assign_workspace_bytes(nt, na) + align_up((size_t)(nt + 1) * sizeof(int), 256) +
           align_up((size_t)(batch_size + 2) * sizeof(int), 256) + 256;
}

int bg_assign_targets_ex(const float *targets, int64_t nt, int32_t row_stride, int32_t ny, int32_t nx, const float *anchors,
                         int32_t na, float anchor_t, float edge_t, int32_t tmask_mode, int32_t batch_size,
                         int64_t *out_idx4, int64_t *out_cls, float *out_anchor, float *out_box, int64_t *out_tmask,
                         float *out_kpts, int64_t cap, int32_t *out_count, void *workspace, size_t workspace_bytes,
                         void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (nt < 0 || ny <= 0 || nx <= 0 || na <= 0 || na > BG_MAX_ANCHORS || !anchors || !out_count || row_stride < 6) return BG_ERR_INVALID;
    if (tmask_mode < 0 || tmask_mode > 2 || (tmask_mode == 2 && batch_size <= 0) || batch_size < 0) return BG_ERR_INVALID;
    if (5ll * na * nt >= (1ll << 31)) return BG_ERR_INVALID;
    if (nt > 0 && (!targets || !out_idx4 || !out_cls || !out_anchor || !out_box || !workspace || cap < 5ll * na * nt))
        return BG_ERR_INVALID;
    if (nt > 0 && ((tmask_mode && !out_tmask) || (row_stride > 6 && !out_kpts))) return BG_ERR_INVALID;
    if (out_box && (((uintptr_t)out_box & 15) != 0)) return BG_ERR_INVALID;
    if (workspace_bytes < bg_assign_ex_workspace_bytes(nt, na, batch_size)) return BG_ERR_WORKSPACE;
    if (cudaMemsetAsync(out_count, 0, 2 * sizeof(int32_t), st) != cudaSuccess) return BG_ERR_LAUNCH;
    if (nt == 0) return BG_OK;
    unsigned char *wsp = (unsigned char *)workspace;
    u64 *chain = (u64 *)wsp; wsp += bg_assign_workspace_bytes(nt, na);
    if (cudaMemsetAsync(chain, 0, assign_chain_words(nt, na) * sizeof(u64), st) != cudaSuccess) return BG_ERR_LAUNCH;
    int *tmask_of_target = (int *)wsp; wsp += align_up((size_t)(nt + 1) * sizeof(int), 256);
    int *block_start = (int *)wsp;
    Assign3K kk;
    AssignK &k = kk.a[0];
    assign_fill(k, targets, nt, ny, nx, anchors, na, anchor_t, edge_t);
    k.row_stride = row_stride;
    k.chain = chain;
    k.idx4 = reinterpret_cast<long long *>(out_idx4);
    k.cls64 = reinterpret_cast<long long *>(out_cls);
    k.anchor = out_anchor; k.box = out_box; k.cap = cap; k.count = out_count;
    if (tmask_mode) {
        assign_tmask_kernel<<<1, 1024, 0, st>>>(targets, row_stride, nt, tmask_mode == 2, batch_size, tmask_of_target, block_start, out_count + 1);
        BG_LAUNCH_CHECK();
        k.tmask_of_target = tmask_of_target;
        k.tmask64 = reinterpret_cast<long long *>(out_tmask);
    }
    if (row_stride > 6) k.kpts = out_kpts;
    kk.a[1] = kk.a[2] = k;
    return assign_launch(kk, 1, st);
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
    Assign3K kk;
    AssignK &k = kk.a[0];
    assign_fill(k, targets, nt, ny, nx, anchors, na, anchor_t, edge_t);
    k.chain = (u64 *)workspace;
    if (nt > 0 && cudaMemsetAsync(workspace, 0, assign_chain_words(nt, na) * sizeof(u64), (cudaStream_t)stream) != cudaSuccess) return BG_ERR_LAUNCH;
    k.idx4 = reinterpret_cast<long long *>(out_idx4);
    k.cls64 = reinterpret_cast<long long *>(out_cls);
    k.anchor = out_anchor; k.box = out_box; k.cap = cap; k.count = out_count;
    kk.a[1] = kk.a[2] = k;
    return assign_launch(kk, 1, (cudaStream_t)stream);
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
    int *M;               // [3] + status [1]   } one zero-filled block: counters, status, confusion counters,
    long long *hist;      // [3,3,C]            } succ flags and the cells' list heads
    unsigned char *succ;  // all scales
    int *head;            // all scales
    size_t zero_begin, zero_bytes;
    float *gobj;          // all scales, contiguous
    int *cell[3], *cls[3], *key[3], *next[3];
    float *ciou[3];
    float4 *gbox[3];
    double *part_match[3];
    double *part_dense[3];
    long long cap;
    long long cells[3], cell_off[3], cells_total;
    int nblk_match, nblk_dense;
    int match_small;      // 256 instead of 1,024 candidates per block of the match kernel
};

bool loss_valid(const bg_loss_params *p)
{
    if (!p || p->B <= 0 || p->C <= 0 || p->C > 4096 || p->na <= 0 || p->na > BG_MAX_ANCHORS || p->nt < 0) return false;
    if (p->input_form < BG_LOSS_DECODED || p->input_form > BG_LOSS_RAW_SPLIT || p->extra_cols < 0 || p->extra_cols > 4096) return false;
    if (p->input_form == BG_LOSS_RAW_SPLIT && p->extra_cols != 0) return false;
    if (5ll * p->na * p->nt >= (1ll << 31)) return false;
    for (int s = 0; s < 3; ++s) {
        if (p->ny[s] <= 0 || p->nx[s] <= 0) return false;
        if ((long long)p->B * p->ny[s] * p->nx[s] * p->na * (p->C + 5 + p->extra_cols) >= (1ll << 40)) return false;
        if ((long long)p->B * p->ny[s] * p->nx[s] * p->na >= (1ll << 31)) return false;
    }
    return true;
}

size_t loss_carve(unsigned char *base, const bg_loss_params *p, LossWs &w)
{
    Bump b{base, 0};
    const int sms = num_sms();
    w.cap = 5ll * p->na * p->nt;
    w.nblk_match = (int)((w.cap + MATCH_CHUNK - 1) / MATCH_CHUNK);
    // small shards: 1,024-candidate blocks would leave most SMs idle (three scales x nblk blocks, six resident per SM)
    w.match_small = 3ll * w.nblk_match < 4ll * sms ? 1 : 0;
    if (w.match_small) w.nblk_match = (int)((w.cap + MATCH_CHUNK_SMALL - 1) / MATCH_CHUNK_SMALL);
    if (w.cap < 1) w.cap = 1;
    w.cells_total = 0;
    long long most = 0;
    for (int s = 0; s < 3; ++s) {
        w.cells[s] = (long long)p->B * p->ny[s] * p->nx[s] * p->na;
        w.cell_off[s] = w.cells_total;
        w.cells_total += w.cells[s];
        most = w.cells[s] > most ? w.cells[s] : most;
    }
    {   // blocks of the dense pass per scale: at least two cells per thread on the largest scale, between one and eight
        // CTAs per SM (the strided objectness loads need every warp the machine holds to hide their latency)
        long long nb = (most + LOSS_THREADS * 2 - 1) / (LOSS_THREADS * 2);
        nb = nb < sms ? sms : (nb > sms * 8 ? sms * 8 : nb);
        w.nblk_dense = (int)nb;
    }
    w.M = b.take<int>(4);
    w.zero_begin = b.off - 4 * sizeof(int);
    w.hist = b.take<long long>(9 * (size_t)p->C);
    w.succ = b.take<unsigned char>(3 * (size_t)w.cap);
    w.head = b.take<int>(w.cells_total);
    w.zero_bytes = b.off - w.zero_begin;
    w.gobj = b.take<float>(w.cells_total);
    for (int s = 0; s < 3; ++s) {
        w.cell[s] = b.take<int>(w.cap);
        w.cls[s] = b.take<int>(w.cap);
        w.key[s] = b.take<int>(w.cap);
        w.ciou[s] = b.take<float>(w.cap);
        w.gbox[s] = b.take<float4>(w.cap);
        w.next[s] = b.take<int>(w.cap);
        w.part_match[s] = b.take<double>((size_t)(w.nblk_match > 0 ? w.nblk_match : 1) * 4);
        w.part_dense[s] = b.take<double>((size_t)w.nblk_dense * 3);
    }
    return align_up(b.off, 256);
}

// in/grads: per scale {obj, cls, box}.  Interleaved forms pass the row base as `obj` (cls/box are derived).
bool loss_fill(Loss3K &k, const bg_loss_params *p, const LossWs &w, const bg_head_ptrs in[3], const bg_head_grads *grads,
               const float *targets)
{
    memset(&k, 0, sizeof(k));
    k.B = p->B; k.C = p->C;
    k.raw = p->input_form != BG_LOSS_DECODED;
    // python: cn = 0.5 * label_smoothing, cp = 1 - cn in double, written into fp32 tensors (detection_loss.py:191-195)
    k.cn = (float)(0.5 * (double)p->label_smoothing);
    k.cp = (float)(1.0 - 0.5 * (double)p->label_smoothing);
    k.nblk_match = w.nblk_match; k.nblk_dense = w.nblk_dense;
    k.box_w = p->box_w; k.conf_w = p->conf_w; k.class_w = p->class_w;
    k.status = w.M + 3;
    const int D = p->C + 5 + p->extra_cols;
    const bool split = p->input_form == BG_LOSS_RAW_SPLIT;
    for (int s = 0; s < 3; ++s) {
        LossScale &S = k.s[s];
        HeadView &v = S.v;
        if (split) {
            if (!in[s].obj || !in[s].cls || !in[s].box) return false;
            if (((uintptr_t)in[s].obj | (uintptr_t)in[s].cls | (uintptr_t)in[s].box) & 15) return false;
            v.obj = in[s].obj; v.cls = in[s].cls; v.box = in[s].box;
            v.so = 1; v.sc = p->C; v.sb = 4;
            if (grads) {
                if (!grads[s].obj || !grads[s].cls || !grads[s].box) return false;
                if (((uintptr_t)grads[s].obj | (uintptr_t)grads[s].cls | (uintptr_t)grads[s].box) & 15) return false;
                v.g_obj = grads[s].obj; v.g_cls = grads[s].cls; v.g_box = grads[s].box;
            }
        } else {
            if (!in[s].obj || ((uintptr_t)in[s].obj & 15)) return false;
            v.obj = in[s].obj; v.cls = in[s].obj + 1; v.box = in[s].obj + 1 + p->C;
            v.so = v.sc = v.sb = D;
            if (grads) {
                if (!grads[s].obj || ((uintptr_t)grads[s].obj & 15)) return false;
                v.g_obj = grads[s].obj; v.g_cls = grads[s].obj + 1; v.g_box = grads[s].obj + 1 + p->C;
            }
        }
        S.cells = w.cells[s];
        float anc[2 * BG_MAX_ANCHORS];
        for (int q = 0; q < p->na; ++q) { anc[2 * q] = p->anchors[s][q][0]; anc[2 * q + 1] = p->anchors[s][q][1]; }
        assign_fill(S.a, targets, p->nt, p->ny[s], p->nx[s], anc, p->na, p->anchor_t, p->edge_t);
        S.M = w.M + s; S.cell = w.cell[s]; S.cls = w.cls[s]; S.key = w.key[s];
        S.ciou = w.ciou[s]; S.gbox = w.gbox[s];
        S.head = w.head + w.cell_off[s]; S.next = w.next[s]; S.succ = w.succ + (size_t)s * w.cap; S.gobj = w.gobj + w.cell_off[s];
        S.part_match = w.part_match[s]; S.part_dense = w.part_dense[s];
        S.hist = w.hist + (size_t)s * 3 * p->C;
        S.scale_w = p->scale_w[s];
    }
    return true;
}

}  // namespace

size_t bg_loss_workspace_bytes(const bg_loss_params *p)
{
    if (!loss_valid(p)) return 0;
    LossWs w;
    return loss_carve(nullptr, p, w);
}

int bg_loss_fwd(const bg_head_ptrs in[3], const float *targets, const bg_loss_params *p, double *out_scalars,
                int64_t *out_hist, float *out_loss, int32_t *out_status, void *workspace, size_t workspace_bytes,
                void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!loss_valid(p) || !in || !out_scalars || !out_hist || !out_loss || !workspace) return BG_ERR_INVALID;
    if (p->nt > 0 && !targets) return BG_ERR_INVALID;
    LossWs w;
    if (loss_carve((unsigned char *)workspace, p, w) > workspace_bytes) return BG_ERR_WORKSPACE;
    Loss3K k;
    if (!loss_fill(k, p, w, in, nullptr, targets)) return BG_ERR_INVALID;
    for (int s = 0; s < 3; ++s) k.s[s].hist_out = reinterpret_cast<long long *>(out_hist) + (size_t)s * 3 * p->C;
    k.scalars = out_scalars;
    k.loss_out = out_loss;
    // one clear: match counters, status word, confusion counters, succ flags, list heads
    if (cudaMemsetAsync((unsigned char *)workspace + w.zero_begin, 0, w.zero_bytes, st) != cudaSuccess) return BG_ERR_LAUNCH;
    if (p->nt > 0) {
        const dim3 grid(w.nblk_match, 3);
        const size_t smem = sizeof(MatchRec) * (w.match_small ? MATCH_CHUNK_SMALL : MATCH_CHUNK) + 3 * (size_t)p->C * sizeof(int);
        // CTAs per SM of the match kernel: 6 (40 registers, some spills) hides the scattered-row latency better than 4
        static const int occ = []() { const char *e = getenv("BG_MATCH_OCC"); return (e && e[0] == '4') ? 4 : 6; }();
        // split form with C % 4 == 0: the class rows are 16-byte aligned -- vector loads, four lanes per match
        static const bool match_vec_env = []() { const char *e = getenv("BG_MATCH_VEC"); return !(e && e[0] == '0'); }();
        const bool vec = match_vec_env && p->input_form == BG_LOSS_RAW_SPLIT && (p->C & 3) == 0;
#define BG_MATCH_LAUNCH(CT, RAW)                                                                                       \
        do {                                                                                                           \
            if (w.match_small) loss_match_kernel<CT, RAW, 6, MATCH_PER_SMALL, 0><<<grid, LOSS_THREADS, smem, st>>>(k);  \
            else if (occ == 4) loss_match_kernel<CT, RAW, 4, MATCH_PER, 0><<<grid, LOSS_THREADS, smem, st>>>(k);       \
            else loss_match_kernel<CT, RAW, 6, MATCH_PER, 0><<<grid, LOSS_THREADS, smem, st>>>(k);                     \
        } while (0)
#define BG_MATCH_LAUNCH_VEC(CT)                                                                                        \
        do {                                                                                                           \
            if (w.match_small) loss_match_kernel<CT, 1, 6, MATCH_PER_SMALL, 1><<<grid, LOSS_THREADS, smem, st>>>(k);    \
            else loss_match_kernel<CT, 1, 6, MATCH_PER, 1><<<grid, LOSS_THREADS, smem, st>>>(k);                       \
        } while (0)
        if (vec) {
            if (p->C == 80) BG_MATCH_LAUNCH_VEC(80); else BG_MATCH_LAUNCH_VEC(0);
        } else if (p->C == 80) {
            if (k.raw) BG_MATCH_LAUNCH(80, 1); else BG_MATCH_LAUNCH(80, 0);
        } else {
            if (k.raw) BG_MATCH_LAUNCH(0, 1); else BG_MATCH_LAUNCH(0, 0);
        }
#undef BG_MATCH_LAUNCH
#undef BG_MATCH_LAUNCH_VEC
        BG_LAUNCH_CHECK();
    }
    static const int dense_occ = []() { const char *e = getenv("BG_DENSE_OCC"); return (e && e[0] == '8') ? 8 : 5; }();
    int rc = dense_occ == 5 ? launch_after(loss_dense_kernel<5>, dim3(w.nblk_dense, 3), dim3(LOSS_THREADS), 0, st, k, p->nt > 0)
                            : launch_after(loss_dense_kernel<8>, dim3(w.nblk_dense, 3), dim3(LOSS_THREADS), 0, st, k, p->nt > 0);
    if (rc != BG_OK) return rc;
    rc = launch_after(loss_finalize_kernel, dim3(1), dim3(768), 0, st, k);
    if (rc != BG_OK) return rc;
    if (out_status && cudaMemcpyAsync(out_status, k.status, sizeof(int), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return BG_ERR_LAUNCH;
    return BG_OK;
}

// split form: the class / box planes of the gradient are zeros except for the matched rows: cleared by memset
// (adjacent planes by one call)
static int loss_clear_planes(const bg_loss_params *p, const bg_head_grads grads[3], cudaStream_t st)
{
    struct Run { unsigned char *p; size_t n; } runs[6];
    int nr = 0;
    for (int s = 0; s < 3; ++s) {
        const size_t cells = (size_t)p->B * p->ny[s] * p->nx[s] * p->na;
        if (!grads[s].cls || !grads[s].box) return BG_ERR_INVALID;
        runs[nr++] = Run{(unsigned char *)grads[s].cls, cells * p->C * sizeof(float)};
        runs[nr++] = Run{(unsigned char *)grads[s].box, cells * 4 * sizeof(float)};
    }
    for (int i = 1; i < nr; ++i)  // insertion sort by address
        for (int j = i; j > 0 && runs[j].p < runs[j - 1].p; --j) { const Run t = runs[j]; runs[j] = runs[j - 1]; runs[j - 1] = t; }
    for (int i = 0; i < nr;) {
        unsigned char *b0 = runs[i].p;
        size_t n = runs[i].n;
        int j = i + 1;
        while (j < nr && runs[j].p == b0 + n) { n += runs[j].n; ++j; }
        if (cudaMemsetAsync(b0, 0, n, st) != cudaSuccess) return BG_ERR_LAUNCH;
        i = j;
    }
    return BG_OK;
}

int bg_loss_clear_grads(const bg_loss_params *p, const bg_head_grads grads[3], void *stream)
{
    if (!loss_valid(p) || !grads || p->input_form != BG_LOSS_RAW_SPLIT) return BG_ERR_INVALID;
    return loss_clear_planes(p, grads, (cudaStream_t)stream);
}

int bg_loss_bwd(const bg_head_ptrs in[3], const bg_loss_params *p, const float *grad_out_dev, float grad_out_host,
                const bg_head_grads grads[3], int32_t flags, void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!loss_valid(p) || !in || !grads || !workspace) return BG_ERR_INVALID;
    LossWs w;
    if (loss_carve((unsigned char *)workspace, p, w) > workspace_bytes) return BG_ERR_WORKSPACE;
    Loss3K k;
    if (!loss_fill(k, p, w, in, grads, nullptr)) return BG_ERR_INVALID;
    k.go_dev = grad_out_dev;
    k.go_host = grad_out_host;
    const int sms = num_sms();
    const cudaEvent_t prof_a = g_prof_loss_start.exchange(nullptr), prof_b = g_prof_loss_stop.exchange(nullptr);
    const bool prof = prof_a && prof_b;
    if (prof) cudaEventRecord(prof_a, st);
    if (p->input_form == BG_LOSS_RAW_SPLIT) {
        // class / box planes: cleared by memset unless the caller did so already (BG_LOSS_BWD_PRECLEARED: next to the
        // forward, on another stream); objectness plane by a kernel
        if (!(flags & BG_LOSS_BWD_PRECLEARED)) {
            const int rc = loss_clear_planes(p, grads, st);
            if (rc != BG_OK) return rc;
        }
        loss_bwd_conf_kernel<<<dim3(sms * 2, 3), 256, 0, st>>>(k);
        BG_LAUNCH_CHECK();
    } else {
        const int D = p->C + 5 + p->extra_cols;
        const size_t smem = (size_t)BWD_WARPS * 32 * D * sizeof(float);
        if (smem > 200 * 1024) return BG_ERR_INVALID;  // rows longer than ~780 floats do not fit the chunk images
        static size_t attr_smem_dev[64] = {0};
        size_t &attr_smem = attr_smem_dev[cur_device()];
        if (smem > attr_smem) {
            if (cudaFuncSetAttribute(loss_bwd_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                (void)cudaGetLastError();
                return BG_ERR_LAUNCH;
            }
            attr_smem = smem;
        }
        static const bool pin = []() { const char *e = getenv("BG_L2_PIN"); return !(e && e[0] == '0'); }();
        if (pin) {
            l2_pin_kernel<<<sms * 4, 256, 0, st>>>(reinterpret_cast<const float4 *>(w.gobj), w.cells_total / 4);
            BG_LAUNCH_CHECK();
        }
        const int per_sm = smem <= 110 * 1024 ? 2 : 1;
        const int rc = launch_after(loss_bwd_stream_kernel, dim3(sms * per_sm), dim3(BWD_WARPS * 32), smem, st, k, pin);
        if (rc != BG_OK) return rc;
    }
    if (prof) cudaEventRecord(prof_b, st);
    if (p->nt == 0) return BG_OK;
    static const bool rows_vec = []() { const char *e = getenv("BG_ROWS_VEC"); return !(e && e[0] == '0'); }();
    if (rows_vec && p->input_form == BG_LOSS_RAW_SPLIT && (k.C & 3) == 0 && k.C <= 128) {  // 16-byte rows: four lanes per match
        if (k.C == 80) return launch_after(loss_bwd_rows_vec_kernel<80>, dim3(sms * 8, 3), dim3(LOSS_THREADS), 0, st, k);
        return launch_after(loss_bwd_rows_vec_kernel<0>, dim3(sms * 8, 3), dim3(LOSS_THREADS), 0, st, k);
    }
    if (k.C == 80) return launch_after(loss_bwd_rows_kernel<80>, dim3(sms * 8, 3), dim3(LOSS_THREADS), 0, st, k);
    return launch_after(loss_bwd_rows_kernel<0>, dim3(sms * 8, 3), dim3(LOSS_THREADS), 0, st, k);
}

// ------------------------------------------------------------------------------------------ f2: mask term of SegmentationLoss
}  // extern "C" (templates below)
namespace {

struct SegWs {
    u64 *chain[3];
    int *tmask_of_target, *block_start;
    int *count;          // [3] matches per scale, [3] = status of the mask numbering
    int *cell[3];
    float4 *box[3];
    long long *tmask[3];
    int *cnt, *off;
    SegEntry *list;
    float4 *mstat;
    double *img_s;
    float *part;
    size_t zero_begin, zero_bytes;
    int G;
    long long cap_s, cap;
};

bool seg_valid(const bg_seg_params *p)
{
    if (!p || p->B <= 0 || p->C <= 0 || p->na <= 0 || p->na > BG_MAX_ANCHORS || p->nt < 0) return false;
    if (p->K != 8 && p->K != 16 && p->K != 32) return false;
    if (p->extra_cols < p->K || p->extra_cols > 4096) return false;
    if (p->Hp <= 0 || p->Wp <= 0 || p->Hm <= 0 || p->Wm <= 0 || (long long)p->Hp * p->Wp >= (1ll << 30)) return false;
    if (15ll * p->na * p->nt >= (1ll << 31)) return false;
    for (int s = 0; s < 3; ++s) {
        if (p->ny[s] <= 0 || p->nx[s] <= 0) return false;
        if ((long long)p->B * p->ny[s] * p->nx[s] * p->na >= (1ll << 31)) return false;
    }
    return true;
}

size_t seg_carve(unsigned char *base, const bg_seg_params *p, SegWs &w)
{
    Bump b{base, 0};
    memset(&w, 0, sizeof(w));
    w.cap_s = 5ll * p->na * p->nt;
    w.cap = 3 * w.cap_s;
    const size_t words = assign_chain_words(p->nt, p->na);
    w.zero_begin = align_up(b.off, 256);
    for (int s = 0; s < 3; ++s) w.chain[s] = b.take<u64>(words);
    w.count = b.take<int>(4);
    w.zero_bytes = b.off - w.zero_begin;
    w.tmask_of_target = b.take<int>((size_t)p->nt + 1);
    w.block_start = b.take<int>((size_t)p->B + 2);
    for (int s = 0; s < 3; ++s) {
        w.cell[s] = b.take<int>((size_t)w.cap_s + 1);
        w.box[s] = b.take<float4>((size_t)w.cap_s + 1);
        w.tmask[s] = b.take<long long>((size_t)w.cap_s + 1);
    }
    w.cnt = b.take<int>(4 * (size_t)p->B);
    w.off = b.take<int>((size_t)p->B + 1);
    w.list = b.take<SegEntry>((size_t)w.cap + 1);
    w.mstat = b.take<float4>((size_t)w.cap + 1);
    w.img_s = b.take<double>(6 * (size_t)p->B);
    // pixel groups per image: enough CTAs for two waves of the GPU, bounded by the tiles there are and by 64 MB of partials
    const int tiles = (int)(((long long)p->Hp * p->Wp + SEG_THREADS - 1) / SEG_THREADS);
    int G = (2 * num_sms() + p->B - 1) / p->B;
    G = G < 1 ? 1 : (G > tiles ? tiles : G);
    while (G > 1 && (size_t)G * w.cap * p->K * 4 > ((size_t)64 << 20)) --G;
    w.G = G;
    w.part = b.take<float>((size_t)G * (w.cap + 1) * (p->K > SEG_FWD_Q ? p->K : SEG_FWD_Q));
    return align_up(b.off, 256);
}

void seg_fill_k(SegK &k, const bg_seg_params *p, const SegWs &w, const float *const preds[3], const float *protos, const float *masks)
{
    memset(&k, 0, sizeof(k));
    k.B = p->B; k.K = p->K; k.D = 5 + p->C + p->extra_cols; k.coef_off = 5 + p->C;
    k.Hp = p->Hp; k.Wp = p->Wp; k.HW = p->Hp * p->Wp; k.Hm = p->Hm; k.Wm = p->Wm;
    k.sy = (float)p->Hm / (float)p->Hp; k.sx = (float)p->Wm / (float)p->Wp;
    for (int s = 0; s < 3; ++s) {
        k.preds[s] = preds[s];
        k.cells_per_img[s] = p->ny[s] * p->nx[s] * p->na;
        k.cell[s] = w.cell[s]; k.box[s] = w.box[s]; k.tmask[s] = w.tmask[s]; k.count[s] = w.count + s;
        k.scale_w[s] = p->scale_w[s];
    }
    k.seg_w = p->seg_w;
    k.protos = protos; k.masks = masks;
    k.cap_s = (int)w.cap_s; k.cap = (int)w.cap; k.G = w.G;
    k.cnt = w.cnt; k.off = w.off; k.list = w.list; k.mstat = w.mstat; k.img_s = w.img_s; k.part = w.part;
}

template <int K>
int seg_launch_fwd(const SegK &k, cudaStream_t st)
{
    seg_fwd_kernel<K><<<dim3(k.G, k.B), SEG_THREADS, SegTile<K>::BYTES, st>>>(k);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

template <int K>
int seg_launch_bwd(const SegK &k, cudaStream_t st)
{
    seg_bwd_coef_kernel<K><<<dim3(k.G, k.B), SEG_THREADS, SegTile<K>::BYTES, st>>>(k);
    BG_LAUNCH_CHECK();
    seg_bwd_protos_kernel<K><<<dim3((k.HW + SEG_THREADS - 1) / SEG_THREADS, k.B), SEG_THREADS, 0, st>>>(k);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

}  // namespace
extern "C" {

size_t bg_seg_loss_workspace_bytes(const bg_seg_params *p)
{
    if (!seg_valid(p)) return 0;
    SegWs w;
    return seg_carve(nullptr, p, w);
}

int bg_seg_loss_fwd(const float *const preds[3], const float *targets, const float *protos, const float *target_masks,
                    const bg_seg_params *p, float *inout_loss, double *out_scalars, int32_t *out_status, void *workspace,
                    size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!seg_valid(p) || !preds || !preds[0] || !preds[1] || !preds[2] || !protos || !target_masks || !inout_loss ||
        !out_scalars || !out_status || !workspace)
        return BG_ERR_INVALID;
    if (p->nt > 0 && !targets) return BG_ERR_INVALID;
    SegWs w;
    if (seg_carve((unsigned char *)workspace, p, w) > workspace_bytes) return BG_ERR_WORKSPACE;
    if (cudaMemsetAsync((unsigned char *)workspace + w.zero_begin, 0, w.zero_bytes, st) != cudaSuccess) return BG_ERR_LAUNCH;
    if (cudaMemsetAsync(out_status, 0, sizeof(int32_t), st) != cudaSuccess) return BG_ERR_LAUNCH;
    if (p->nt > 0) {
        // mask numbering of the targets (detection_dataset.py:146-157), then the three scales' matches in one launch
        assign_tmask_kernel<<<1, 1024, 0, st>>>(targets, 6, p->nt, 1, p->B, w.tmask_of_target, w.block_start, out_status);
        BG_LAUNCH_CHECK();
        Assign3K kk;
        for (int s = 0; s < 3; ++s) {
            AssignK &a = kk.a[s];
            assign_fill(a, targets, p->nt, p->ny[s], p->nx[s], &p->anchors[s][0][0], p->na, p->anchor_t, p->edge_t);
            a.chain = w.chain[s];
            a.cap = w.cap_s; a.count = w.count + s;
            a.cell = w.cell[s]; a.box = reinterpret_cast<float *>(w.box[s]);
            a.tmask_of_target = w.tmask_of_target; a.tmask64 = w.tmask[s];
        }
        const int rc = assign_launch(kk, 3, st);
        if (rc != BG_OK) return rc;
    }
    SegK k;
    seg_fill_k(k, p, w, preds, protos, target_masks);
    k.scalars = out_scalars; k.loss = inout_loss;
    seg_count_kernel<<<p->B, SEG_THREADS, 0, st>>>(k);
    BG_LAUNCH_CHECK();
    seg_fill_kernel<<<p->B, SEG_THREADS, 0, st>>>(k);
    BG_LAUNCH_CHECK();
    const int rc = p->K == 32 ? seg_launch_fwd<32>(k, st) : p->K == 16 ? seg_launch_fwd<16>(k, st) : seg_launch_fwd<8>(k, st);
    if (rc != BG_OK) return rc;
    seg_reduce_kernel<<<p->B, SEG_THREADS, 0, st>>>(k);
    BG_LAUNCH_CHECK();
    seg_final_kernel<<<1, SEG_THREADS, 0, st>>>(k);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

int bg_seg_loss_bwd(const float *const preds[3], const float *protos, const float *target_masks, const bg_seg_params *p,
                    const float *grad_out_dev, float *const grad_preds[3], float *grad_protos, void *workspace,
                    size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!seg_valid(p) || !preds || !preds[0] || !preds[1] || !preds[2] || !protos || !target_masks || !grad_preds ||
        !grad_preds[0] || !grad_preds[1] || !grad_preds[2] || !grad_protos || !workspace)
        return BG_ERR_INVALID;
    SegWs w;
    if (seg_carve((unsigned char *)workspace, p, w) > workspace_bytes) return BG_ERR_WORKSPACE;
    SegK k;
    seg_fill_k(k, p, w, preds, protos, target_masks);
    for (int s = 0; s < 3; ++s) k.gpreds[s] = grad_preds[s];
    k.gprotos = grad_protos;
    k.go_dev = grad_out_dev;
    const int rc = p->K == 32 ? seg_launch_bwd<32>(k, st) : p->K == 16 ? seg_launch_bwd<16>(k, st) : seg_launch_bwd<8>(k, st);
    if (rc != BG_OK) return rc;
    const long long work = (long long)w.cap * p->K;
    const int blocks = (int)(work / SEG_THREADS + 1 < 4ll * num_sms() ? work / SEG_THREADS + 1 : 4ll * num_sms());
    seg_bwd_scatter_kernel<<<blocks, SEG_THREADS, 0, st>>>(k);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

int bg_seg_masks(const float *coefs, const int32_t *row_offsets, const float *protos, int32_t B, int32_t K, int32_t Hp,
                 int32_t Wp, int64_t n, int32_t H, int32_t W, float *scratch, uint8_t *out_masks, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || K <= 0 || K > SEGM_KMAX || Hp <= 0 || Wp <= 0 || H <= 0 || W <= 0 || n < 0 || !row_offsets || !protos) return BG_ERR_INVALID;
    if ((long long)Hp * Wp >= (1ll << 30) || n * (long long)H * W >= (1ll << 40)) return BG_ERR_INVALID;
    if (n == 0) return BG_OK;
    if (!coefs || !scratch || !out_masks || ((uintptr_t)out_masks & 3)) return BG_ERR_INVALID;
    SegMaskK k;
    memset(&k, 0, sizeof(k));
    k.B = B; k.K = K; k.Hp = Hp; k.Wp = Wp; k.HW = Hp * Wp; k.H = H; k.W = W; k.n = n;
    k.coefs = coefs; k.row_off = row_offsets; k.protos = protos; k.low = scratch; k.out = out_masks;
    k.ry = (float)Hp / (float)H; k.rx = (float)Wp / (float)W;
    seg_lowres_kernel<<<dim3((k.HW + SEG_THREADS - 1) / SEG_THREADS, B), SEG_THREADS, 0, st>>>(k);
    BG_LAUNCH_CHECK();
    const bool vec = (W & 3) == 0;
    const long long blocks = n * ((H + SEGM_YCHUNK - 1) / SEGM_YCHUNK);   // one block per (row, 16 output lines)
    if (blocks >= (1ll << 31)) return BG_ERR_INVALID;
    const int wv = vec ? W / 4 : W;
    const int threads = wv >= SEG_THREADS ? SEG_THREADS : ((wv + 31) / 32) * 32;   // no idle warps when a line is short
    if (vec) seg_upsample_kernel<4><<<(unsigned)blocks, threads, 0, st>>>(k);
    else seg_upsample_kernel<1><<<(unsigned)blocks, threads, 0, st>>>(k);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

int bg_loss_pack(const double *scalars, const int64_t *cells3, int32_t C, double *pack15, void *stream)
{
    if (!scalars || !cells3 || !pack15 || C <= 0) return BG_ERR_INVALID;
    loss_pack_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scalars, (double)cells3[0], (double)cells3[1], (double)cells3[2], (double)C, pack15);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

int bg_loss_combine(const double *pack15, const bg_loss_params *p, double *out_loss, void *stream)
{
    if (!pack15 || !p || !out_loss || p->C <= 0) return BG_ERR_INVALID;
    loss_combine_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pack15, p->box_w, p->conf_w, p->class_w, p->scale_w[0], p->scale_w[1],
                                                     p->scale_w[2], (double)p->C, out_loss);
    BG_LAUNCH_CHECK();
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
