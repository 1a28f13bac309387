// Training side: YOLOv5-style target assignment (ordered compaction), element-wise CIoU with its
// analytic gradient, the fused detection loss (forward + backward) and the anchor-fit metrics.
//
// Reference semantics: dataset/detection_dataset.py:90-246,
// modules/detection_loss.py:125-264, utils/make_anchors.py:14-39.
#pragma once
#include "common.cuh"

namespace bg {

// ------------------------------------------------------------------------------------------------
// target assignment
// ------------------------------------------------------------------------------------------------
struct AssignK {
    const float *targets;  // [nt, row_stride]: (img, cls, x, y, w, h, keypoint columns...)
    int row_stride;        // 6 + number of keypoint columns
    long long nt;
    int ny, nx, na;
    float fnx, fny;
    float aw[BG_MAX_ANCHORS], ah[BG_MAX_ANCHORS];  // anchors in grid units: anchor * (nx, ny)   (:183)
    float anchor_t, edge_t;
    long long ncand;       // 5*na*nt
    u64 *chain;            // [nblocks + 2] ticket counter + look-back words, zeroed by the host
    // outputs (any may be null)
    long long *idx4;       // [4,cap]
    long long *cls64;      // [cap]
    float *anchor;         // [cap,2]
    float *box;            // [cap,4]
    long long cap;
    int *count;            // [1] M
    // packed outputs for the fused loss (any may be null)
    int *cell;             // [cap] ((b*ny+gj)*nx+gi)*na+a
    int *cls32;            // [cap]
    // segmentation / keypoint variants (detection_dataset.py:132-172,239-245); null = detection branch
    const int *tmask_of_target;  // [nt] mask index of every target (see assign_tmask_kernel)
    long long *tmask64;          // [cap]
    float *kpts;                 // [cap, row_stride-6]
};

constexpr int ASSIGN_THREADS = 1024;
struct Assign3K { AssignK a[3]; };  // one launch covers up to three scales: blockIdx.y selects the entry

struct AssignOut { int b, gj, gi, a, cls; long long t; float aw, ah, bx, by, bw, bh; };

// torch.remainder(x, 1): fmod(x, 1) == x - trunc(x) exactly in fp32 (the subtraction cannot round), NaN for +-inf
__device__ __forceinline__ float torch_remainder1(float x)
{
    float r = __fsub_rn(x, truncf(x));
    if (r != 0.0f && r < 0.0f) r += 1.0f;
    return r;
}

// candidate c = (k*na + a)*nt + t; returns whether it is emitted and, if so, its outputs
__device__ __forceinline__ bool assign_eval(const AssignK &k, long long c, AssignOut &o)
{
    const u32 ka_u = (u32)c / (u32)k.nt;  // 5*na*nt < 2^31 is checked on the host
    const long long t = (long long)((u32)c - ka_u * (u32)k.nt);
    const int ka = (int)ka_u;
    const int a = ka % k.na, kk = ka / k.na;
    const float *tg = k.targets + (long long)k.row_stride * t;
    const float gx = __fmul_rn(tg[2], k.fnx), gy = __fmul_rn(tg[3], k.fny);
    const float gw = __fmul_rn(tg[4], k.fnx), gh = __fmul_rn(tg[5], k.fny);
    const float rw = __fdiv_rn(gw, k.aw[a]), rh = __fdiv_rn(gh, k.ah[a]);
    const float irw = __fdiv_rn(1.0f, rw), irh = __fdiv_rn(1.0f, rh);
    const float m = fmaxf(fmaxf(rw, irw), fmaxf(rh, irh));
    if (!(m < k.anchor_t)) return false;
    float ox = 0.f, oy = 0.f;
    bool sel = true;
    if (kk == 1) { sel = (torch_remainder1(gx) < k.edge_t) && (gx > 1.0f); ox = k.edge_t; }
    else if (kk == 2) { sel = (torch_remainder1(gy) < k.edge_t) && (gy > 1.0f); oy = k.edge_t; }
    else if (kk == 3) { const float ix = __fsub_rn(k.fnx, gx); sel = (torch_remainder1(ix) < k.edge_t) && (ix > 1.0f); ox = -k.edge_t; }
    else if (kk == 4) { const float iy = __fsub_rn(k.fny, gy); sel = (torch_remainder1(iy) < k.edge_t) && (iy > 1.0f); oy = -k.edge_t; }
    if (!sel) return false;
    long long gi = (long long)__fsub_rn(gx, ox), gj = (long long)__fsub_rn(gy, oy);  // .long() truncates (:231)
    gi = gi < 0 ? 0 : (gi > k.nx - 1 ? k.nx - 1 : gi);
    gj = gj < 0 ? 0 : (gj > k.ny - 1 ? k.ny - 1 : gj);
    o.t = t;
    o.b = (int)(long long)tg[0];
    o.cls = (int)(long long)tg[1];
    o.gi = (int)gi; o.gj = (int)gj; o.a = a;
    o.aw = k.aw[a]; o.ah = k.ah[a];
    o.bx = __fsub_rn(gx, (float)gi); o.by = __fsub_rn(gy, (float)gj);  // clamped cell (aliasing at :232-237)
    o.bw = gw; o.bh = gh;
    return true;
}

// Mask index of every target (detection_dataset.py:132-170).  overlap = 0: the target's own position.
// overlap = 1: masks of one image are merged into one plane, the index is 1 + the position inside the image's
// block, the blocks being laid out by the per-image counts for image ids 0..batch_size-1 in order (which is the
// targets' own order when they are sorted by image, as collate_fn produces them).  status[0] is set to 1 when the
// counts do not add up to nt (the reference's torch.cat would raise).
__global__ void __launch_bounds__(1024) assign_tmask_kernel(const float *targets, int row_stride, long long nt, int overlap,
                                                             int batch_size, int *tmask_of_target, int *block_start /*[batch_size+1]*/,
                                                             int *status)
{
    if (!overlap) {
        for (long long t = threadIdx.x; t < nt; t += 1024) tmask_of_target[t] = (int)t;
        if (threadIdx.x == 0) status[0] = 0;
        return;
    }
    for (int i = threadIdx.x; i <= batch_size; i += 1024) block_start[i] = 0;
    __syncthreads();
    for (long long t = threadIdx.x; t < nt; t += 1024) {
        const float fi = targets[(long long)row_stride * t];
        const int i = (int)fi;
        if (i >= 0 && i < batch_size && (float)i == fi) atomicAdd(&block_start[i + 1], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // batch_size is small: a serial prefix is fine
        int run = 0;
        for (int i = 1; i <= batch_size; ++i) { run += block_start[i]; block_start[i] = run; }
        status[0] = (run == nt) ? 0 : 1;
    }
    __syncthreads();
    for (long long p = threadIdx.x; p < nt; p += 1024) {
        int lo = 0, hi = batch_size;  // largest block with block_start[block] <= p
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (block_start[mid] <= p) lo = mid; else hi = mid;
        }
        tmask_of_target[p] = (int)(p - block_start[lo]) + 1;
    }
}

// One pass: a block evaluates its 4096 candidates, counts the emitted ones, obtains its output offset by decoupled
// look-back over the earlier blocks (blocks are numbered by an atomic ticket, so every block a block waits for is
// already running) and writes its matches -- the reference's (k, a, t) boolean-mask order, deterministically.
// k.chain: [nblk + 2] u64, zeroed by the host before the launch: word 0 = the ticket counter, word 1 + i = block i
// (ASSIGN_AGG | own count, later ASSIGN_PREFIX | inclusive count).
constexpr u64 ASSIGN_PREFIX = 1ull << 63, ASSIGN_AGG = 1ull << 62, ASSIGN_VALUE = (1ull << 62) - 1;

constexpr int ASSIGN_SLICES = 4;  // candidates per thread: a block covers ASSIGN_SLICES * ASSIGN_THREADS consecutive candidates
constexpr int ASSIGN_BLOCK = ASSIGN_SLICES * ASSIGN_THREADS;

__global__ void __launch_bounds__(ASSIGN_THREADS) assign_onepass_kernel(Assign3K kk)
{
    const AssignK &k = kk.a[blockIdx.y];
    __shared__ int s_w[ASSIGN_SLICES][32];
    __shared__ long long s_base;
    __shared__ unsigned s_ticket;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_ticket = (unsigned)atomicAdd((unsigned long long *)k.chain, 1ull);
    __syncthreads();
    const int bid = (int)s_ticket;
    AssignOut o[ASSIGN_SLICES];
    bool f[ASSIGN_SLICES];
    u32 bal[ASSIGN_SLICES];
#pragma unroll
    for (int i = 0; i < ASSIGN_SLICES; ++i) {  // slice i = candidates [bid*BLOCK + i*1024, +1024): ballot order = candidate order
        const long long c = (long long)bid * ASSIGN_BLOCK + i * ASSIGN_THREADS + threadIdx.x;
        f[i] = (c < k.ncand) && assign_eval(k, c, o[i]);
        bal[i] = __ballot_sync(0xffffffffu, f[i]);
        if (lane == 0) s_w[i][wid] = __popc(bal[i]);
    }
    __syncthreads();
    int off[ASSIGN_SLICES], tot = 0;
#pragma unroll
    for (int i = 0; i < ASSIGN_SLICES; ++i) {
        // exclusive prefix of the 32 warp counts of slice i at this warp's index, and their total: one shuffle scan
        static_assert(ASSIGN_THREADS == 1024, "one warp count per lane");
        const int v = s_w[i][lane];
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        const int woff = __shfl_sync(0xffffffffu, inc - v, wid);
        const int st = __shfl_sync(0xffffffffu, inc, 31);
        off[i] = tot + woff;
        tot += st;
    }
    if (wid == 0) {
        volatile u64 *words = (volatile u64 *)k.chain + 1;
        if (lane == 0) words[bid] = ASSIGN_AGG | (u64)tot;
        long long base = 0;
        int hi = bid - 1;
        while (hi >= 0) {
            const int j = hi - lane;  // newest first
            u64 v = ASSIGN_PREFIX;    // lanes past block 0 read as "prefix 0"
            if (j >= 0) { do { v = words[j]; } while (!(v & (ASSIGN_AGG | ASSIGN_PREFIX))); }
            const u32 pm = __ballot_sync(0xffffffffu, (v & ASSIGN_PREFIX) != 0);
            const int stop = pm ? (__ffs(pm) - 1) : 32;
            long long part = (lane <= stop) ? (long long)(v & ASSIGN_VALUE) : 0;
            part = warp_sum(part);
            base += part;
            if (pm) break;
            hi -= 32;
        }
        if (lane == 0) {
            s_base = base;
            words[bid] = ASSIGN_PREFIX | (u64)(base + tot);
            if (bid == (int)gridDim.x - 1) *k.count = (int)(base + tot);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ASSIGN_SLICES; ++i) {
        if (!f[i]) continue;
        const long long m = s_base + off[i] + __popc(bal[i] & lanemask_lt());
        if (m >= k.cap) continue;
        const AssignOut &q = o[i];
        if (k.idx4) {
            k.idx4[0 * k.cap + m] = q.b; k.idx4[1 * k.cap + m] = q.gj;
            k.idx4[2 * k.cap + m] = q.gi; k.idx4[3 * k.cap + m] = q.a;
        }
        if (k.cls64) k.cls64[m] = q.cls;
        if (k.anchor) { k.anchor[2 * m] = q.aw; k.anchor[2 * m + 1] = q.ah; }
        if (k.box) { float4 *bp = reinterpret_cast<float4 *>(k.box) + m; *bp = make_float4(q.bx, q.by, q.bw, q.bh); }
        if (k.cell) k.cell[m] = ((q.b * k.ny + q.gj) * k.nx + q.gi) * k.na + q.a;
        if (k.cls32) k.cls32[m] = q.cls;
        if (k.tmask64) k.tmask64[m] = k.tmask_of_target[q.t];
        if (k.kpts) {
            const int nk = k.row_stride - 6;
            const float *src = k.targets + (long long)k.row_stride * q.t + 6;
            for (int r = 0; r < nk; ++r) k.kpts[m * nk + r] = src[r];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// CIoU (modules/detection_loss.py:229-264): fp32 forward in the reference's operation order; the
// gradient w.r.t. the prediction (alpha constant) is evaluated in double from the same quantities.
// ------------------------------------------------------------------------------------------------
template <typename G>  // G = double (stand-alone bg_ciou_bwd) or float (fused loss: rtol 1e-4 against fp32 autograd)
__device__ __forceinline__ float ciou_eval(const float p[4], const float t[4], float e, G *g /*4 or null*/)
{
    const float pw = p[2], ph = p[3], tw = t[2], th = t[3];
    const float px1 = __fsub_rn(p[0], __fdiv_rn(pw, 2.0f)), py1 = __fsub_rn(p[1], __fdiv_rn(ph, 2.0f));
    const float px2 = __fadd_rn(px1, pw), py2 = __fadd_rn(py1, ph);
    const float tx1 = __fsub_rn(t[0], __fdiv_rn(tw, 2.0f)), ty1 = __fsub_rn(t[1], __fdiv_rn(th, 2.0f));
    const float tx2 = __fadd_rn(tx1, tw), ty2 = __fadd_rn(ty1, th);
    const float iw_raw = __fsub_rn(fminf(px2, tx2), fmaxf(px1, tx1));
    const float ih_raw = __fsub_rn(fminf(py2, ty2), fmaxf(py1, ty1));
    const float iw = iw_raw < 0.f ? 0.f : iw_raw, ih = ih_raw < 0.f ? 0.f : ih_raw;
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(pw, ph), __fmul_rn(tw, th)), inter);
    const float iou = __fdiv_rn(inter, __fadd_rn(uni, e));
    const float cw = __fsub_rn(fmaxf(px2, tx2), fminf(px1, tx1));
    const float ch = __fsub_rn(fmaxf(py2, ty2), fminf(py1, ty1));
    const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(cw, cw), __fmul_rn(ch, ch)), e);
    const float k4pi2 = 0.40528473456935105f;  // float32(4 / pi^2)
    const float dat = __fsub_rn(atanf(__fdiv_rn(tw, th)), atanf(__fdiv_rn(pw, ph)));
    const float v = __fmul_rn(k4pi2, __fmul_rn(dat, dat));
    const float dx = __fsub_rn(p[0], t[0]), dy = __fsub_rn(p[1], t[1]);
    const float rho2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const float a = __fdiv_rn(v, __fadd_rn(__fsub_rn(v, iou), __fadd_rn(1.0f, e)));
    const float ciou = __fsub_rn(iou, __fadd_rn(__fdiv_rn(rho2, c2), __fmul_rn(a, v)));
    if (g) {
        const G dpx1[4] = {1, 0, -0.5, 0}, dpx2[4] = {1, 0, 0.5, 0};
        const G dpy1[4] = {0, 1, 0, -0.5}, dpy2[4] = {0, 1, 0, 0.5};
        const bool iw_pos = iw_raw >= 0.f, ih_pos = ih_raw >= 0.f;  // clamp passes the gradient at equality
        // torch.min / torch.max split the gradient evenly on exact ties
        const G s_minx2 = px2 < tx2 ? (G)1 : (px2 == tx2 ? (G)0.5 : (G)0), s_maxx1 = px1 > tx1 ? (G)1 : (px1 == tx1 ? (G)0.5 : (G)0);
        const G s_miny2 = py2 < ty2 ? (G)1 : (py2 == ty2 ? (G)0.5 : (G)0), s_maxy1 = py1 > ty1 ? (G)1 : (py1 == ty1 ? (G)0.5 : (G)0);
        const G s_maxx2 = px2 > tx2 ? (G)1 : (px2 == tx2 ? (G)0.5 : (G)0), s_minx1 = px1 < tx1 ? (G)1 : (px1 == tx1 ? (G)0.5 : (G)0);
        const G s_maxy2 = py2 > ty2 ? (G)1 : (py2 == ty2 ? (G)0.5 : (G)0), s_miny1 = py1 < ty1 ? (G)1 : (py1 == ty1 ? (G)0.5 : (G)0);
        const G den = (G)uni + (G)e;
        const G r = (G)pw / (G)ph;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const G diw = iw_pos ? (s_minx2 * dpx2[q] - s_maxx1 * dpx1[q]) : (G)0;
            const G dih = ih_pos ? (s_miny2 * dpy2[q] - s_maxy1 * dpy1[q]) : (G)0;
            const G dinter = diw * ih + iw * dih;
            const G dpwph = (q == 2 ? (G)ph : (G)0) + (q == 3 ? (G)pw : (G)0);
            const G duni = dpwph - dinter;
            const G diou = dinter / den - (G)inter * duni / (den * den);
            const G dcw = s_maxx2 * dpx2[q] - s_minx1 * dpx1[q];
            const G dch = s_maxy2 * dpy2[q] - s_miny1 * dpy1[q];
            const G dc2 = (G)2 * cw * dcw + (G)2 * ch * dch;
            const G drho2 = (q == 0 ? (G)2 * dx : (G)0) + (q == 1 ? (G)2 * dy : (G)0);
            const G dr = (q == 2 ? (G)1 / (G)ph : (G)0) + (q == 3 ? -(G)pw / ((G)ph * ph) : (G)0);
            const G dv = (G)k4pi2 * (G)2 * (G)dat * (-(dr / ((G)1 + r * r)));
            const G av = (q >= 2) ? (G)a * dv : (G)0;  // v depends on (w,h) only: a NaN alpha never reaches x,y
            g[q] = diou - (drho2 / c2 - (G)rho2 * dc2 / ((G)c2 * c2) + av);
        }
    }
    return ciou;
}

__global__ void ciou_fwd_kernel(const float *p, const float *t, long long M, float e, float *out)
{
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float4 pp = reinterpret_cast<const float4 *>(p)[m], tt = reinterpret_cast<const float4 *>(t)[m];
    const float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ta[4] = {tt.x, tt.y, tt.z, tt.w};
    out[m] = ciou_eval<double>(pa, ta, e, nullptr);
}

__global__ void ciou_bwd_kernel(const float *p, const float *t, const float *go, long long M, float e, float *gp)
{
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float4 pp = reinterpret_cast<const float4 *>(p)[m], tt = reinterpret_cast<const float4 *>(t)[m];
    const float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ta[4] = {tt.x, tt.y, tt.z, tt.w};
    double g[4];
    ciou_eval<double>(pa, ta, e, g);
    const double s = go[m];
    reinterpret_cast<float4 *>(gp)[m] = make_float4((float)(s * g[0]), (float)(s * g[1]), (float)(s * g[2]), (float)(s * g[3]));
}

// ------------------------------------------------------------------------------------------------
// fused detection loss: every kernel covers the three scales (blockIdx.y)
//   forward : loss_match_kernel  pass A, one thread per match   : gather, CIoU and its gradient, link the match
//                                                                  into its cell's list (atomicExch on the head)
//                                pass B, eight lanes per match  : class BCE, argmax, confusion counters
//             loss_dense_kernel  one thread per cell            : objectness BCE against the CIoU of the cell's last match;
//                                                                  keeps sigmoid(x) - t for the backward
//             loss_finalize_kernel                              : fixed-order reduction, scalars, total loss
//   backward: loss_bwd_stream_kernel  zeros + the objectness column, one 16-byte store per four elements
//             loss_bwd_rows_kernel    class / box columns of the matched rows, summed over the cell's match list
//                                     (gather backward = index_put(accumulate=True)); no atomics
// ------------------------------------------------------------------------------------------------
struct LossScale {
    const float *preds;  // [B,ny,nx,na,D]
    float *grad;         // same shape (backward only)
    long long cells;     // B*ny*nx*na
    const int *M;        // device count from the assignment
    const int *cell;     // [cap]
    const int *cls;      // [cap]
    const float *anchor; // [cap,2]
    const float *box;    // [cap,4]
    float *ciou;         // [cap]
    float4 *gbox;        // [cap] d ciou / d (x, y, w_raw, h_raw) of the matched prediction
    int *head;           // [cells] most recently linked match of the cell (-1: none); list through next[]
    int *next;           // [cap] previous match of the same cell, -1 at the end of the list (= the cell's first match)
    unsigned char *succ; // [cap] 1 if a later match was linked in front of this one (zeroed by the host)
    float *gobj;         // [cells] sigmoid(obj) - t_conf
    double *part_match;  // [nblk_match,4]: sum(1-ciou), sum(ciou), sum(sig(obj)), sum(bce_cls)
    double *part_dense;  // [nblk_dense,3]: sum(bce_obj), sum(sig(obj) | t==0), n_neg
    long long *hist;     // [3,C]
    double scale_w;
};

struct Loss3K {
    LossScale s[3];
    int C, D;
    float cn, cp;        // class targets: 0.5*label_smoothing and 1-cn
    int nblk_match, nblk_dense;
    double box_w, conf_w, class_w;
    double *scalars;     // [3,8]
    float *loss_out;     // [1] total loss (modules/detection_loss.py:107-110)
    const float *go_dev; // backward: upstream gradient on the device (or null -> go_host)
    float go_host;
};

__device__ __forceinline__ float bce_logits(float x, float t)
{
    // ATen binary_cross_entropy_with_logits: (1 - t) * x - log_sigmoid(x)
    const float ls = __fsub_rn(fminf(x, 0.0f), log1pf(expf(-fabsf(x))));
    return __fsub_rn(__fmul_rn(__fsub_rn(1.0f, t), x), ls);
}

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

constexpr int LOSS_THREADS = 256;
constexpr int ROWS_UNROLL = 10;  // 8 lanes x 10 = one 80-class row per batch of loads

template <int CT>  // compile-time class count (80: no bounds predicates in the unrolled class loop); 0 = runtime
__global__ void __launch_bounds__(LOSS_THREADS, 6) loss_match_kernel(Loss3K k)
{
    extern __shared__ int s_hist[];  // [3,C] block-local confusion counters
    __shared__ double s_red[LOSS_THREADS / 32][4];
    const LossScale &S = k.s[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int M = *S.M;
    const int C = CT ? CT : k.C, D = C + 5;
    constexpr bool kFull = CT != 0 && CT % (8 * ROWS_UNROLL) == 0;  // every lane's batch lies inside the row
    for (int i = tid; i < 3 * C; i += LOSS_THREADS) s_hist[i] = 0;
    __syncthreads();

    // pass A: one thread per match -- gather, CIoU and its gradient, link the match into its cell's list
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (long long m = (long long)blockIdx.x * LOSS_THREADS + tid; m < M; m += (long long)gridDim.x * LOSS_THREADS) {
        const int cell = S.cell[m];
        const float *row = S.preds + (long long)cell * D;
        const float aw = S.anchor[2 * m], ah = S.anchor[2 * m + 1];
        const float p[4] = {__ldg(row + C + 1), __ldg(row + C + 2), __fmul_rn(__ldg(row + C + 3), aw),
                            __fmul_rn(__ldg(row + C + 4), ah)};
        const float obj = __ldg(row);
        const float4 tb = reinterpret_cast<const float4 *>(S.box)[m];
        const float t[4] = {tb.x, tb.y, tb.z, tb.w};
        float g[4];
        const float ci = ciou_eval<float>(p, t, 1e-7f, g);
        S.ciou[m] = ci;
        S.gbox[m] = make_float4(g[0], g[1], g[2] * aw, g[3] * ah);
        const int prev = atomicExch(&S.head[cell], (int)m);
        S.next[m] = prev;
        if (prev >= 0) S.succ[prev] = 1;
        a0 += (double)__fsub_rn(1.0f, ci);
        a1 += (double)ci;
        a2 += (double)sigmoid_acc(obj);
    }

    // pass B: eight lanes per match (four matches per warp in flight)
    const int gl = lane & 7;
    for (long long mb = ((long long)blockIdx.x * (LOSS_THREADS / 32) + wid) * 4; mb < M;
         mb += (long long)gridDim.x * (LOSS_THREADS / 32) * 4) {
        const long long m = mb + (lane >> 3);
        const bool valid = m < M;
        // sum_c bce(x_c, t_c) = sum_c softplus(x_c) - cn * sum_c x_c - (cp - cn) * x_target, with
        // softplus(x) = max(x, 0) + log(1 + exp(-|x|)); the logs of a lane's classes are taken as ONE log of the
        // product (each factor lies in (1, 2], ten of them stay far from overflow) -- fast exp/log units, |error| of
        // the row sum < 1e-6 relative, far inside the rtol 1e-5 bar of the mean over M*C terms
        float bsum = 0.f, best = -INFINITY;
        int bi = 0x7fffffff, tc = -1;
        if (valid) {
            tc = S.cls[m];
            const float *row = S.preds + (long long)S.cell[m] * D + 1;
            float spos = 0.f, sx = 0.f, lsum = 0.f;
            for (int cb = 0; cb < C; cb += 8 * ROWS_UNROLL) {
                float x[ROWS_UNROLL];
#pragma unroll
                for (int u = 0; u < ROWS_UNROLL; ++u) { const int c = cb + 8 * u + gl; x[u] = (kFull || c < C) ? __ldg(row + c) : -INFINITY; }
                float prod = 1.f;
#pragma unroll
                for (int u = 0; u < ROWS_UNROLL; ++u) {
                    const int c = cb + 8 * u + gl;
                    if (kFull || c < C) {
                        prod *= 1.0f + __expf(-fabsf(x[u]));
                        spos += fmaxf(x[u], 0.0f);
                        sx += x[u];
                        if (x[u] > best) { best = x[u]; bi = c; }
                    }
                }
                lsum += __logf(prod);
            }
            bsum = spos + lsum - k.cn * sx;
            if (gl == 0 && tc >= 0 && tc < C) bsum -= (k.cp - k.cn) * __ldg(row + tc);  // (labels outside 0..C-1 have no target column)
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (valid && gl == 0) {
            a3 += (double)bsum;
            if (tc >= 0 && tc < C) {
                if (bi == tc) atomicAdd(&s_hist[tc], 1);
                atomicAdd(&s_hist[C + tc], 1);
            }
            if (bi >= 0 && bi < C) atomicAdd(&s_hist[2 * C + bi], 1);
        }
    }

    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
    if (lane == 0) { s_red[wid][0] = a0; s_red[wid][1] = a1; s_red[wid][2] = a2; s_red[wid][3] = a3; }
    __syncthreads();
    if (tid < 4) {
        double s = 0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) s += s_red[w][tid];
        S.part_match[(long long)blockIdx.x * 4 + tid] = s;
    }
    for (int i = tid; i < 3 * C; i += LOSS_THREADS) {
        const int v = s_hist[i];
        if (v) atomicAdd((unsigned long long *)&S.hist[i], (unsigned long long)v);
    }
}

// One float every D*4 bytes: ask L2 for 64-byte fills instead of the default (measured on B200 for a 340-byte
// stride: 89 instead of 122 bytes of DRAM time per element, scripts/micro/write_stride.cu)
__device__ __forceinline__ float ld_stride_f32(const float *p)
{
    float v;
    asm volatile("ld.global.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// dense objectness BCE over every cell
__global__ void __launch_bounds__(LOSS_THREADS) loss_dense_kernel(Loss3K k)
{
    __shared__ double s_red[LOSS_THREADS / 32][3];
    const LossScale &S = k.s[blockIdx.y];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double a0 = 0, a1 = 0, a2 = 0;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    // four cells per thread and pass, their strided loads in flight together (same per-thread order of the sums)
    constexpr int DENSE_PER = 4;
    const long long stride = (long long)gridDim.x * LOSS_THREADS;
    for (long long c0 = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x; c0 < S.cells; c0 += DENSE_PER * stride) {
        float xs[DENSE_PER];
        int ws[DENSE_PER];
#pragma unroll
        for (int u = 0; u < DENSE_PER; ++u) {
            const long long c = c0 + u * stride;
            xs[u] = 0.f; ws[u] = -1;
            if (c < S.cells) { xs[u] = ld_stride_f32(S.preds + c * k.D); ws[u] = S.head[c]; }
        }
#pragma unroll
        for (int u = 0; u < DENSE_PER; ++u) {
            const long long c = c0 + u * stride;
            if (c >= S.cells) break;
            const float x = xs[u];
            int w = ws[u];  // "last match wins": the highest match index of the cell's list
            for (int j = w; j >= 0; j = S.next[j]) w = j > w ? j : w;
            const float t = w >= 0 ? S.ciou[w] : 0.0f;
            const float sg = sigmoid_acc(x);
            a0 += (double)bce_logits(x, t);
            if (t == 0.0f) { a1 += (double)sg; a2 += 1.0; }
            // the backward's streaming kernel reads this residual in the middle of a 2 GB write stream: keep it in L2
            // (evict-last), so those reads do not turn into DRAM read/write turnarounds
            asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(S.gobj + c), "f"(__fsub_rn(sg, t)), "l"(pol) : "memory");
        }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) { s_red[wid][0] = a0; s_red[wid][1] = a1; s_red[wid][2] = a2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) s += s_red[w][threadIdx.x];
        S.part_dense[(long long)blockIdx.x * 3 + threadIdx.x] = s;
    }
}

// fixed-order final reduction -> scalars[3,8] and the combined loss; 768 threads, 256 per scale
__global__ void __launch_bounds__(768) loss_finalize_kernel(Loss3K k)
{
    __shared__ double s_red[3][8][7];
    __shared__ double s_terms[3][3];
    const int sc = threadIdx.x >> 8, t = threadIdx.x & 255;
    const int lane = t & 31, wid = t >> 5;
    const LossScale &S = k.s[sc];
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int b = t; b < k.nblk_match; b += 256)
        for (int q = 0; q < 4; ++q) v[q] += S.part_match[(long long)b * 4 + q];
    for (int b = t; b < k.nblk_dense; b += 256)
        for (int q = 0; q < 3; ++q) v[4 + q] += S.part_dense[(long long)b * 3 + q];
    for (int q = 0; q < 7; ++q) v[q] = warp_sum(v[q]);
    if (lane == 0) for (int q = 0; q < 7; ++q) s_red[sc][wid][q] = v[q];
    __syncthreads();
    if (t == 0) {
        double s[7];
        for (int q = 0; q < 7; ++q) { s[q] = 0; for (int w = 0; w < 8; ++w) s[q] += s_red[sc][w][q]; }
        const double M = (double)*S.M;
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        double *o = k.scalars + 8 * sc;
        o[0] = M > 0 ? s[0] / M : 0.0;                      // NaN -> 0 (:209-210)
        o[1] = s[4] / (double)S.cells;
        o[2] = M > 0 ? s[3] / (M * (double)k.C) : 0.0;
        o[3] = M > 0 ? s[1] / M : nan;
        o[4] = M > 0 ? s[2] / M : nan;
        o[5] = s[6] > 0 ? s[5] / s[6] : nan;
        o[6] = M;
        o[7] = s[6];
        s_terms[sc][0] = S.scale_w * o[0]; s_terms[sc][1] = S.scale_w * o[1]; s_terms[sc][2] = S.scale_w * o[2];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double lbox = s_terms[0][0] + s_terms[1][0] + s_terms[2][0];
        const double lconf = s_terms[0][1] + s_terms[1][1] + s_terms[2][1];
        const double lcls = s_terms[0][2] + s_terms[1][2] + s_terms[2][2];
        *k.loss_out = (float)(k.box_w * lbox + k.conf_w * lconf + k.class_w * lcls);
    }
}

struct BwdScales { float conf, cls, box; };
__device__ __forceinline__ BwdScales bwd_scales(const Loss3K &k, const LossScale &S)
{
    const double go = (double)(k.go_dev ? *k.go_dev : k.go_host) * S.scale_w;
    const int M = *S.M;
    BwdScales r;
    r.conf = (float)(k.conf_w * go / (double)S.cells);
    r.cls = M > 0 ? (float)(k.class_w * go / ((double)M * (double)k.C)) : 0.0f;
    r.box = M > 0 ? (float)(-k.box_w * go / (double)M) : 0.0f;
    return r;
}

// grad_preds is zeros, one objectness value per row, and the class / box columns of the matched rows (6 % of
// the rows).  Two kernels, no atomics, both with every warp of the machine busy:
//
// loss_bwd_stream_kernel: written the way a fill would be.  Every warp owns a shared-memory image of a 32-row
//   chunk (32*D floats, zero-filled once); per chunk it drops the 32 objectness values into column 0 of the rows
//   (lane = row, residual fetched with one coalesced load, the next chunk's prefetched) and copies the image out
//   with 16-byte loads from shared memory and 512-byte coalesced stores -- no per-element index arithmetic.
//   (A TMA bulk store of the image, cp.async.bulk.global.shared::cta, measured 5.5 TB/s; this loop is faster.)
// loss_bwd_rows_kernel: eight lanes per match; the first match linked into a cell owns the row (known from the
//   forward: next == -1; and succ == 0 says it is the cell's only match, so nothing has to be looked up by cell),
//   otherwise it walks the cell's list (gather backward = index_put(accumulate=True): every match of the cell contributes)
//   and rewrites the row's class / box columns:
//   class c: cls*(n*(sigmoid(x)-cn) - (cp-cn)*#{matches of class c}),  box j: box * sum of the CIoU gradients.

constexpr int BWD_WARPS = 8;  // warps per CTA of the streaming kernel (each owns one chunk image)

// Mixing the 26 MB of residual reads into the 2.1 GB write stream costs ~45 us of DRAM read/write turnarounds
// (measured: the same kernel without the loads runs 352 instead of 396 us).  So the residuals are pulled into L2
// first, marked evict-last, and the write stream below uses evict-first stores: the streaming kernel's loads hit L2.
__global__ void __launch_bounds__(256) l2_pin_kernel(const float4 *p, long long n4)
{
    float acc = 0.f;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float4 v;
        asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i), "l"(pol));
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 1.2345e-30f) asm volatile("trap;");  // keeps the loads alive
}

__global__ void __launch_bounds__(BWD_WARPS * 32) loss_bwd_stream_kernel(Loss3K k)
{
    extern __shared__ __align__(128) float bwd_smem[];  // [BWD_WARPS][32*D]
    const int D = k.D;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int chunk_floats = 32 * D, chunk_f4 = 8 * D;
    float *im = bwd_smem + (size_t)wid * chunk_floats;
    for (int i = lane; i < chunk_floats; i += 32) im[i] = 0.f;  // only column 0 of a row ever changes
    __syncwarp();
    const float4 *im4 = reinterpret_cast<const float4 *>(im);

    long long nch[3], tot = 0;
    for (int s = 0; s < 3; ++s) { nch[s] = k.s[s].cells >> 5; tot += nch[s]; }  // full chunks; remainders below
    const float cf0 = bwd_scales(k, k.s[0]).conf, cf1 = bwd_scales(k, k.s[1]).conf, cf2 = bwd_scales(k, k.s[2]).conf;
    const long long gw = (long long)blockIdx.x * BWD_WARPS + wid, nw = (long long)gridDim.x * BWD_WARPS;

    auto locate = [&](long long g, int &si, long long &row0) {
        si = 0;
        long long ch = g;
        if (ch >= nch[0]) { ch -= nch[0]; si = 1; if (ch >= nch[1]) { ch -= nch[1]; si = 2; } }
        row0 = ch << 5;
    };
    // the residuals of the next BWD_AHEAD chunks are kept in flight: a chunk is only ~22 store instructions long,
    // far shorter than the latency of the load that feeds the one after it
    constexpr int BWD_AHEAD = 4;
    float gq[BWD_AHEAD];
#pragma unroll
    for (int a = 0; a < BWD_AHEAD; ++a) {
        gq[a] = 0.f;
        const long long ga = gw + a * nw;
        if (ga < tot) { int s2; long long r2; locate(ga, s2, r2); gq[a] = k.s[s2].gobj[r2 + lane]; }
    }
    for (long long g = gw; g < tot; g += nw) {
        int si; long long row0;
        locate(g, si, row0);
        float4 *dst = reinterpret_cast<float4 *>(k.s[si].grad + row0 * D);
        im[lane * D] = (si == 0 ? cf0 : (si == 1 ? cf1 : cf2)) * gq[0];
#pragma unroll
        for (int a = 0; a + 1 < BWD_AHEAD; ++a) gq[a] = gq[a + 1];
        gq[BWD_AHEAD - 1] = 0.f;
        const long long ga = g + BWD_AHEAD * nw;
        if (ga < tot) { int s2; long long r2; locate(ga, s2, r2); gq[BWD_AHEAD - 1] = k.s[s2].gobj[r2 + lane]; }
        __syncwarp();
        for (int f = lane; f < chunk_f4; f += 32) __stcs(dst + f, im4[f]);  // shared-memory image -> 512-byte coalesced, evict-first stores
        __syncwarp();
    }

    // rows beyond the last full chunk of a scale (cells not a multiple of 32): plain stores by one CTA
    if (blockIdx.x == 0) {
        for (int s = 0; s < 3; ++s) {
            const LossScale &S = k.s[s];
            const float cf = s == 0 ? cf0 : (s == 1 ? cf1 : cf2);
            for (long long row = (S.cells & ~31LL) + wid; row < S.cells; row += BWD_WARPS)
                for (int col = lane; col < D; col += 32) S.grad[row * D + col] = col == 0 ? cf * S.gobj[row] : 0.f;
        }
    }
}

template <int CT>  // compile-time class count (80: no bounds predicates in the unrolled row loop); 0 = runtime
__global__ void __launch_bounds__(LOSS_THREADS) loss_bwd_rows_kernel(Loss3K k)
{
    const LossScale &S = k.s[blockIdx.y];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, gl = lane & 7;
    const int M = *S.M;
    if (M <= 0) return;
    const BwdScales sc = bwd_scales(k, S);
    const int C = CT ? CT : k.C, D = C + 5;
    for (long long mb = ((long long)blockIdx.x * (LOSS_THREADS / 32) + wid) * 4; mb < M;
         mb += (long long)gridDim.x * (LOSS_THREADS / 32) * 4) {
        const long long m = mb + (lane >> 3);
        if (m >= M) continue;
        if (S.next[m] != -1) continue;  // the cell's first match owns the row
        const int cell = S.cell[m];
        int n = 0, c1 = -1, c2 = -1;
        float gb[4] = {0.f, 0.f, 0.f, 0.f};
        int lhead = (int)m;
        if (!S.succ[m]) {  // the only match of its cell (the usual case): everything is addressed by m, no list walk
            const float4 gq = S.gbox[m];
            gb[0] = gq.x; gb[1] = gq.y; gb[2] = gq.z; gb[3] = gq.w;
            c1 = S.cls[m];
            n = 1;
        } else {
            lhead = S.head[cell];
            for (int j = lhead; j >= 0; j = S.next[j]) {
                const float4 gq = S.gbox[j];
                gb[0] += gq.x; gb[1] += gq.y; gb[2] += gq.z; gb[3] += gq.w;
                if (n == 0) c1 = S.cls[j]; else if (n == 1) c2 = S.cls[j];
                ++n;
            }
        }
        const float *xrow = S.preds + (long long)cell * D;
        float *grow = S.grad + (long long)cell * D;
        // class column c: cls*(n*(sg - cn) - (cp - cn)*hits(c)) = ka*sg - kb - kc*hits(c); the (at most two) columns
        // with hits are fixed up after the row loop by the lane that wrote them
        const float ka = sc.cls * (float)n, kb = ka * k.cn, kc = sc.cls * (k.cp - k.cn);
        for (int cb = 1; cb <= C; cb += 8 * ROWS_UNROLL) {
            float x[ROWS_UNROLL];
#pragma unroll
            for (int u = 0; u < ROWS_UNROLL; ++u) {  // all loads of the batch in flight before the first store
                const int col = cb + 8 * u + gl;
                x[u] = (CT != 0 && CT % (8 * ROWS_UNROLL) == 0) || col <= C ? __ldg(xrow + col) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < ROWS_UNROLL; ++u) {
                const int col = cb + 8 * u + gl;
                if (!(CT != 0 && CT % (8 * ROWS_UNROLL) == 0) && col > C) continue;
                grow[col] = ka * sigmoid_fast(x[u]) - kb;
            }
        }
        if (n <= 2) {
            if (c1 >= 0 && gl == (c1 & 7)) grow[1 + c1] -= kc;
            if (c2 >= 0 && gl == (c2 & 7)) grow[1 + c2] -= kc;
        } else if (gl == 0) {  // three or more matches on one cell: one subtraction per match
            for (int j = lhead; j >= 0; j = S.next[j]) grow[1 + S.cls[j]] -= kc;
        }
        if (gl < 4) grow[C + 1 + gl] = sc.box * (gl == 0 ? gb[0] : gl == 1 ? gb[1] : gl == 2 ? gb[2] : gb[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// anchor-fit metrics (utils/make_anchors.py:14-39)
// ------------------------------------------------------------------------------------------------
struct RatioK { const float *wh; long long n; int k; float aw[32], ah[32]; float inv_thr; double *out; };

__global__ void __launch_bounds__(256) ratio_metrics_kernel(RatioK k)
{
    double s0 = 0, s1 = 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < k.n; i += (long long)gridDim.x * 256) {
        const float w = k.wh[2 * i], h = k.wh[2 * i + 1];
        float best = -INFINITY;
        for (int j = 0; j < k.k; ++j) {
            const float r0 = __fdiv_rn(w, k.aw[j]), r1 = __fdiv_rn(h, k.ah[j]);
            const float v = fminf(fminf(r0, __fdiv_rn(1.0f, r0)), fminf(r1, __fdiv_rn(1.0f, r1)));
            best = fmaxf(best, v);
        }
        if (best > k.inv_thr) { s0 += (double)best; s1 += 1.0; }
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    if ((threadIdx.x & 31) == 0) { atomicAdd(k.out, s0); atomicAdd(k.out + 1, s1); }
    if (blockIdx.x == 0 && threadIdx.x == 0) k.out[2] = (double)k.n;
}

}  // namespace bg
