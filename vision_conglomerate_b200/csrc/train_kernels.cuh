// Training side: YOLOv5-style target assignment (ordered compaction), element-wise CIoU with its
// analytic gradient and the anchor-fit metrics.  The fused detection loss is in loss_kernels.cuh.
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
// division of the gradient arithmetic: exact in double, the fast reciprocal-multiply in float (2 ulp; the fused loss's
// gradients are held to rtol 1e-4 and its match kernel is instruction-bound)
__device__ __forceinline__ double gdiv(double a, double b) { return a / b; }
__device__ __forceinline__ float gdiv(float a, float b) { return __fdividef(a, b); }

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
        const G r = gdiv((G)pw, (G)ph);
        const G inv_den = gdiv((G)1, den), inv_c2 = gdiv((G)1, (G)c2), inv_ph = gdiv((G)1, (G)ph), inv_1r2 = gdiv((G)1, (G)1 + r * r);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const G diw = iw_pos ? (s_minx2 * dpx2[q] - s_maxx1 * dpx1[q]) : (G)0;
            const G dih = ih_pos ? (s_miny2 * dpy2[q] - s_maxy1 * dpy1[q]) : (G)0;
            const G dinter = diw * ih + iw * dih;
            const G dpwph = (q == 2 ? (G)ph : (G)0) + (q == 3 ? (G)pw : (G)0);
            const G duni = dpwph - dinter;
            const G diou = dinter * inv_den - (G)inter * duni * inv_den * inv_den;
            const G dcw = s_maxx2 * dpx2[q] - s_minx1 * dpx1[q];
            const G dch = s_maxy2 * dpy2[q] - s_miny1 * dpy1[q];
            const G dc2 = (G)2 * cw * dcw + (G)2 * ch * dch;
            const G drho2 = (q == 0 ? (G)2 * dx : (G)0) + (q == 1 ? (G)2 * dy : (G)0);
            const G dr = (q == 2 ? inv_ph : (G)0) + (q == 3 ? -(G)pw * inv_ph * inv_ph : (G)0);
            const G dv = (G)k4pi2 * (G)2 * (G)dat * (-(dr * inv_1r2));
            const G av = (q >= 2) ? (G)a * dv : (G)0;  // v depends on (w,h) only: a NaN alpha never reaches x,y
            g[q] = diou - (drho2 * inv_c2 - (G)rho2 * dc2 * inv_c2 * inv_c2 + av);
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
// anchor-fit metrics (utils/make_anchors.py:14-39)
// ------------------------------------------------------------------------------------------------
struct RatioK { const float *wh; long long n; int k; float aw[32], ah[32]; float inv_thr; double *out; };

// The two sums are accumulated as 64-bit integers (the values lie in (0, 1]: fixed point with 40 fractional bits is
// exact for every value >= 2^-17 and holds 2^23 boxes), so the result does not depend on the order in which the
// blocks finish.  The accumulators live in the output words themselves (out[0..2] viewed as u64: sum, count, blocks
// done; zeroed by the host) and the last block to finish converts them to doubles in place.
constexpr double RATIO_FIX = 1099511627776.0;  // 2^40

__global__ void __launch_bounds__(256) ratio_metrics_kernel(RatioK k)
{
    unsigned long long s0 = 0, s1 = 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < k.n; i += (long long)gridDim.x * 256) {
        const float w = k.wh[2 * i], h = k.wh[2 * i + 1];
        float best = -INFINITY;
        for (int j = 0; j < k.k; ++j) {
            const float r0 = __fdiv_rn(w, k.aw[j]), r1 = __fdiv_rn(h, k.ah[j]);
            const float v = fminf(fminf(r0, __fdiv_rn(1.0f, r0)), fminf(r1, __fdiv_rn(1.0f, r1)));
            best = fmaxf(best, v);
        }
        if (best > k.inv_thr) { s0 += (unsigned long long)((double)best * RATIO_FIX); s1 += 1; }
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(k.out);
    if ((threadIdx.x & 31) == 0 && s1) { atomicAdd(acc, s0); atomicAdd(acc + 1, s1); }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(acc + 2, 1ull) == (unsigned long long)gridDim.x - 1) {
            __threadfence();
            const unsigned long long a0 = atomicAdd(acc, 0ull), a1 = atomicAdd(acc + 1, 0ull);
            k.out[0] = (double)a0 / RATIO_FIX;
            k.out[1] = (double)a1;
            k.out[2] = (double)k.n;
        }
    }
}

}  // namespace bg
