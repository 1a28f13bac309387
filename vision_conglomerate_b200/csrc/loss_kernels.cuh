// The fused detection loss (modules/detection_loss.py:84-226) for the default BCE configuration: target
// assignment, matched-row gather, CIoU, "last match wins" objectness targets, dense objectness BCE, class BCE,
// confusion counters, and the backward that writes the dense gradient.  Every kernel covers the three scales.
//
// Input forms (HeadView / Loss3K::raw):
//   decoded rows   [B,ny,nx,na,5+C]: what DetectionNet._get_scale_pred(inference=False) returns (detection.py:98-173);
//   raw rows       the head's own output (logits): the training-mode decode xy = 2s-0.5, wh = (2s)^2 (:122,125) is
//                  applied in registers in the forward and its derivative in the backward;
//   raw, split     the head's three conv outputs before EffiDecHead.forward concatenates them (common.py:908-919):
//                  conf [B,ny,nx,na], cls [B,ny,nx,na,C], bbox [B,ny,nx,na,4]: the objectness plane is
//                  contiguous and the gradient comes back in the same three pieces.
//
//   forward : loss_match_kernel   one block per 512 assignment candidates (k, a, t): evaluates the YOLOv5 rule
//                                 (dataset/detection_dataset.py:90-246), allots the block's matches a range of the
//                                 match arrays with one atomic, then per match: gather, CIoU and its gradient, link
//                                 into the cell's list; eight lanes per match: class BCE, argmax, confusion counters.
//                                 The reference's (k, a, t) order survives as the candidate number ("key") of a match.
//             loss_dense_kernel   one thread per cell: objectness BCE against the CIoU of the cell's LAST match in
//                                 the reference's order (= highest key in the cell's list); keeps sigmoid(x) - t
//             loss_finalize_kernel  fixed-order reduction of the per-block partial sums, scalars, total loss
//   backward: loss_bwd_stream_kernel  (interleaved rows) zeros + the objectness column, written like a fill
//             loss_bwd_conf_kernel    (split) the objectness plane; the class / box planes are cleared by memset
//             loss_bwd_rows_kernel    class / box columns of the matched rows, summed over the cell's match list
//                                     (gather backward = index_put(accumulate=True)); no atomics
// Results do not depend on the order in which blocks run: sums are per block in a fixed order, the match arrays'
// slot order (the only thing the atomics decide) never reaches a result, integer counters are exact.
#pragma once
#include "train_kernels.cuh"

namespace bg {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct HeadView {
    const float *obj, *cls, *box;  // element j of a cell: obj[cell*so], cls[cell*sc + j], box[cell*sb + j]
    float *g_obj, *g_cls, *g_box;  // backward outputs, same strides
    int so, sc, sb;
};

struct LossScale {
    HeadView v;
    long long cells;     // B*ny*nx*na
    AssignK a;           // assignment parameters of this scale (targets, anchors in grid units, thresholds)
    int *M;              // [1] number of matches (the blocks' allocation counter, zeroed by the host)
    int *cell;           // [cap] ((b*ny+gj)*nx+gi)*na+a
    int *cls;            // [cap]
    int *key;            // [cap] candidate number (k*na + a)*nt + t: the reference's match order
    float *ciou;         // [cap]
    float4 *gbox;        // [cap] d ciou / d (the four box values as stored in the input tensor)
    int *head;           // [cells] 1 + the most recently linked match of the cell (0: none; zeroed by the host); list through next[]
    int *next;           // [cap] previously linked match of the same cell, -1 for the first linked
    unsigned char *succ; // [cap] 1 if another match was linked in front of this one (zeroed by the host)
    float *gobj;         // [cells] sigmoid(obj) - t_conf
    double *part_match;  // [nblk_match,4]: sum(1-ciou), sum(ciou), sum(sig(obj)), sum(bce_cls)
    double *part_dense;  // [nblk_dense,3]: sum(bce_obj), sum(sig(obj) | t==0), n_neg
    long long *hist;     // [3,C] in the workspace (zeroed by the host)
    long long *hist_out; // [3,C] caller's copy, written by the finalize kernel
    double scale_w;
};

struct Loss3K {
    LossScale s[3];
    int B, C;
    int raw;             // 1: the box values are logits, decode in registers
    float cn, cp;        // class targets: 0.5*label_smoothing and 1-cn
    int nblk_match, nblk_dense;
    double box_w, conf_w, class_w;
    double *scalars;     // [3,8]
    float *loss_out;     // [1] total loss (modules/detection_loss.py:107-110)
    int *status;         // [1] bit 0: a target's image id or class id is outside the batch / class range (dropped)
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
constexpr int MATCH_PER = 4;        // assignment candidates per thread (large problems)
constexpr int MATCH_PER_SMALL = 1;  // ... when the grid would otherwise not fill the GPU (small shards)
constexpr int MATCH_CHUNK = LOSS_THREADS * MATCH_PER;
constexpr int MATCH_CHUNK_SMALL = LOSS_THREADS * MATCH_PER_SMALL;

// per-block staging record of a match (shared memory): what the per-match phases need, so that they run on dense
// lanes and the evaluation's registers are dead by then
struct MatchRec { int cell; int cls_a; float bx, by, bw, bh; };  // cls_a = cls | anchor << 16 | local candidate number << 20

template <int CT, int RAW, int OCC, int PER, int VEC>  // CT: compile-time class count (80: no bounds predicates in the class loop); 0 = runtime;
                                                       // OCC: CTAs per SM; PER: candidates per thread (a block covers 256*PER);
                                                       // VEC: split form with C % 4 == 0 -- class rows read with 16-byte loads
__global__ void __launch_bounds__(LOSS_THREADS, OCC) loss_match_kernel(Loss3K k)
{
    constexpr int MATCH_CHUNK = LOSS_THREADS * PER, MATCH_PER = PER;  // (shadow the namespace-level defaults)
    extern __shared__ __align__(16) unsigned char s_dyn[];   // [MATCH_CHUNK] MatchRec, then [3*C] block-local confusion counters
    __shared__ double s_red[LOSS_THREADS / 32][4];
    __shared__ int s_wcnt[LOSS_THREADS / 32], s_wpre[LOSS_THREADS / 32 + 1];
    __shared__ int s_base;
    MatchRec *s_rec = reinterpret_cast<MatchRec *>(s_dyn);
    int *s_hist = reinterpret_cast<int *>(s_dyn + sizeof(MatchRec) * MATCH_CHUNK);
    const LossScale &S = k.s[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int C = CT ? CT : k.C;
    constexpr bool kFull = CT != 0 && CT % (8 * ROWS_UNROLL) == 0;  // every lane's batch lies inside the row
    for (int i = tid; i < 3 * C; i += LOSS_THREADS) s_hist[i] = 0;

    // ---- 1. the block's candidates.  Warp w owns candidates [chunk*CHUNK + w*128, +128) (slice i = its lanes' i-th
    //         candidates) and writes the records of the emitted ones, in candidate order, to its own quarter-kilobyte
    //         region of the staging array: one evaluation per candidate, no block-wide prefix needed for the position.
    const long long c0 = (long long)blockIdx.x * MATCH_CHUNK;
    bool bad = false;
    int wcount = 0;  // records this warp has written so far (warp-uniform)
#pragma unroll
    for (int i = 0; i < MATCH_PER; ++i) {
        const int lc = wid * (32 * MATCH_PER) + i * 32 + lane;  // local candidate number
        const long long c = c0 + lc;
        AssignOut o;
        bool f = (c < S.a.ncand) && assign_eval(S.a, c, o);
        if (f && ((unsigned)o.b >= (unsigned)k.B || (unsigned)o.cls >= (unsigned)C)) {
            f = false;  // the reference raises IndexError on these (preds[batch_idx...], t_cls[range, cls]); here: dropped + flagged
            bad = true;
        }
        const u32 bal = __ballot_sync(0xffffffffu, f);
        if (f) {
            MatchRec r;
            r.cell = ((o.b * S.a.ny + o.gj) * S.a.nx + o.gi) * S.a.na + o.a;
            r.cls_a = o.cls | (o.a << 16) | (lc << 20);
            r.bx = o.bx; r.by = o.by; r.bw = o.bw; r.bh = o.bh;
            s_rec[wid * (32 * MATCH_PER) + wcount + __popc(bal & lanemask_lt())] = r;
        }
        wcount += __popc(bal);
    }
    if (lane == 0) s_wcnt[wid] = wcount;
    __syncthreads();
    pdl_wait();  // everything above reads only the targets; the counters / match arrays / head words are cleared upstream
    if (bad) atomicOr(k.status, 1);
    if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) { const int v = s_wcnt[w]; s_wpre[w] = tot; tot += v; }
        s_wpre[LOSS_THREADS / 32] = tot;
        s_base = tot ? atomicAdd(S.M, tot) : 0;
    }
    __syncthreads();
    const int nloc = s_wpre[LOSS_THREADS / 32], base = s_base;
    // dense match number j of the block -> its record: the warp whose range holds j, then the position inside its region
    auto rec_of = [&](int j) -> const MatchRec & {
        int w = 0;
#pragma unroll
        for (int q = 1; q < LOSS_THREADS / 32; ++q) w += (j >= s_wpre[q]) ? 1 : 0;
        return s_rec[w * (32 * MATCH_PER) + (j - s_wpre[w])];
    };

    // ---- 2. one thread per match: gather, CIoU and its gradient, link into the cell's list; records to global memory
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int j = tid; j < nloc; j += LOSS_THREADS) {
        const MatchRec r = rec_of(j);
        const int m = base + j, cell = r.cell, an = (r.cls_a >> 16) & 15;
        const float aw = S.a.aw[an], ah = S.a.ah[an];
        const float *bp = S.v.box + (long long)cell * S.v.sb;
        float b0 = __ldg(bp), b1 = __ldg(bp + 1), b2 = __ldg(bp + 2), b3 = __ldg(bp + 3);
        const float obj = __ldg(S.v.obj + (long long)cell * S.v.so);
        float d0 = 1.f, d1 = 1.f, d2 = 1.f, d3 = 1.f;  // d(decoded)/d(stored)
        if (RAW) {  // detection.py:122,125: xy = sigmoid*2 - 0.5, wh = (sigmoid*2)^2
            const float s0 = sigmoid_acc(b0), s1 = sigmoid_acc(b1), s2 = sigmoid_acc(b2), s3 = sigmoid_acc(b3);
            b0 = __fsub_rn(__fmul_rn(s0, 2.0f), 0.5f); b1 = __fsub_rn(__fmul_rn(s1, 2.0f), 0.5f);
            const float w2 = __fmul_rn(s2, 2.0f), h2 = __fmul_rn(s3, 2.0f);
            b2 = __fmul_rn(w2, w2); b3 = __fmul_rn(h2, h2);
            d0 = 2.0f * s0 * (1.0f - s0); d1 = 2.0f * s1 * (1.0f - s1);
            d2 = 8.0f * s2 * s2 * (1.0f - s2); d3 = 8.0f * s3 * s3 * (1.0f - s3);
        }
        const float p[4] = {b0, b1, __fmul_rn(b2, aw), __fmul_rn(b3, ah)};
        const float t[4] = {r.bx, r.by, r.bw, r.bh};
        float g[4];
        const float ci = ciou_eval<float>(p, t, 1e-7f, g);
        S.cell[m] = cell; S.cls[m] = r.cls_a & 0xffff;
        S.key[m] = (int)c0 + (int)((u32)r.cls_a >> 20);
        S.ciou[m] = ci;
        S.gbox[m] = make_float4(g[0] * d0, g[1] * d1, g[2] * aw * d2, g[3] * ah * d3);
        const int prev = atomicExch(&S.head[cell], m + 1) - 1;
        S.next[m] = prev;
        if (prev >= 0) S.succ[prev] = 1;
        a0 += (double)__fsub_rn(1.0f, ci);
        a1 += (double)ci;
        a2 += (double)sigmoid_acc(obj);
    }

    // ---- 3. class BCE, argmax, confusion counters
    // sum_c bce(x_c, t_c) = sum_c softplus(x_c) - cn * sum_c x_c - (cp - cn) * x_target, with
    // softplus(x) = max(x, 0) + log(1 + exp(-|x|)); the logs of a lane's classes are taken as ONE log of the
    // product (each factor lies in (1, 2], twenty of them stay far from overflow) -- fast exp/log units, |error| of
    // the row sum < 1e-6 relative, far inside the rtol 1e-5 bar of the mean over M*C terms
    if constexpr (VEC) {
        // split form: four lanes per match, 16-byte loads (eight matches per warp in flight); lane g of a group owns
        // the float4 slots g, g+4, g+8, ... of the class row
        const int gl = lane & 3, C4 = C >> 2;
        constexpr int NV = CT ? (CT / 4 + 3) / 4 : 8;
        for (int jb = wid * 8; jb < nloc; jb += (LOSS_THREADS / 32) * 8) {
            const int j = jb + (lane >> 2);
            const bool valid = j < nloc;
            float bsum = 0.f, best = -INFINITY;
            int bi = 0x7fffffff, tc = -1;
            if (valid) {
                const MatchRec &rr = rec_of(j);
                tc = rr.cls_a & 0xffff;
                const float *row1 = S.v.cls + (long long)rr.cell * C;
                const float4 *row = reinterpret_cast<const float4 *>(row1);
                float spos = 0.f, sx = 0.f, lsum = 0.f;
                for (int vb = 0; vb < C4; vb += 4 * NV) {
                    float4 x[NV];
#pragma unroll
                    for (int u = 0; u < NV; ++u) { const int v = vb + 4 * u + gl; x[u] = v < C4 ? __ldg(row + v) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); }
                    float prod = 1.f, mx = -INFINITY;
#pragma unroll
                    for (int u = 0; u < NV; ++u) {
                        if (vb + 4 * u + gl < C4) {
                            const float xs[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                prod *= 1.0f + __expf(-fabsf(xs[e]));
                                spos += fmaxf(xs[e], 0.0f);
                                sx += xs[e];
                                mx = fmaxf(mx, xs[e]);
                            }
                        }
                    }
                    lsum += __logf(prod);
                    if (mx > best) {  // first index holding the batch maximum (slots ascend in class index)
                        best = mx;
#pragma unroll
                        for (int u = NV - 1; u >= 0; --u) {
                            const int c = 4 * (vb + 4 * u + gl);
                            if (x[u].w == mx) bi = c + 3;
                            if (x[u].z == mx) bi = c + 2;
                            if (x[u].y == mx) bi = c + 1;
                            if (x[u].x == mx) bi = c;
                        }
                    }
                }
                bsum = spos + lsum - k.cn * sx;
                if (gl == 0) bsum -= (k.cp - k.cn) * __ldg(row1 + tc);
            }
#pragma unroll
            for (int o2 = 2; o2 > 0; o2 >>= 1) {
                bsum += __shfl_xor_sync(0xffffffffu, bsum, o2);
                const float ob = __shfl_xor_sync(0xffffffffu, best, o2);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o2);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (valid && gl == 0) {
                a3 += (double)bsum;
                if (bi == tc) atomicAdd(&s_hist[tc], 1);
                atomicAdd(&s_hist[C + tc], 1);
                if (bi >= 0 && bi < C) atomicAdd(&s_hist[2 * C + bi], 1);
            }
        }
    } else {
    // eight lanes per match (four matches per warp in flight)
    const int gl = lane & 7;
    for (int jb = wid * 4; jb < nloc; jb += (LOSS_THREADS / 32) * 4) {
        const int j = jb + (lane >> 3);
        const bool valid = j < nloc;
        // sum_c bce(x_c, t_c) = sum_c softplus(x_c) - cn * sum_c x_c - (cp - cn) * x_target, with
        // softplus(x) = max(x, 0) + log(1 + exp(-|x|)); the logs of a lane's classes are taken as ONE log of the
        // product (each factor lies in (1, 2], ten of them stay far from overflow) -- fast exp/log units, |error| of
        // the row sum < 1e-6 relative, far inside the rtol 1e-5 bar of the mean over M*C terms
        float bsum = 0.f, best = -INFINITY;
        int bi = 0x7fffffff, tc = -1;
        if (valid) {
            const MatchRec &rr = rec_of(j);
            tc = rr.cls_a & 0xffff;
            const float *row = S.v.cls + (long long)rr.cell * S.v.sc;
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
            if (gl == 0) bsum -= (k.cp - k.cn) * __ldg(row + tc);
        }
#pragma unroll
        for (int o2 = 4; o2 > 0; o2 >>= 1) {
            bsum += __shfl_xor_sync(0xffffffffu, bsum, o2);
            const float ob = __shfl_xor_sync(0xffffffffu, best, o2);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o2);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (valid && gl == 0) {
            a3 += (double)bsum;
            if (bi == tc) atomicAdd(&s_hist[tc], 1);
            atomicAdd(&s_hist[C + tc], 1);
            if (bi >= 0 && bi < C) atomicAdd(&s_hist[2 * C + bi], 1);
        }
    }
    }
    pdl_launch_dependents();

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
// One cell of the dense pass: the objectness BCE against the CIoU of the cell's last match, the residual for the backward.
__device__ __forceinline__ float dense_cell(const LossScale &S, float x, int head_m, double &a0, double &a1, double &a2)
{
    int w = -1, wk = -1;  // "last match wins": the highest candidate number in the cell's list
    for (int j = head_m; j >= 0; j = S.next[j]) { const int kj = S.key[j]; if (kj > wk) { wk = kj; w = j; } }
    const float t = w >= 0 ? S.ciou[w] : 0.0f;
    // bce(x, t) = (1 - t) x - log_sigmoid(x) = max(x, 0) - t x + log(1 + e), e = exp(-|x|), and
    // sigmoid(x) = (x >= 0 ? 1 : e) / (1 + e): one fast exponential, one fast logarithm and one fast division
    // per cell (|error| < 3e-7 absolute on terms of order one -- the mean over millions of cells and the
    // gradient residual stay far inside rtol 1e-5 / 1e-4); the accurate forms made this kernel issue-bound
    // on the contiguous (split) objectness plane
    const float e = __expf(-fabsf(x));
    const float l = e < 1e-4f ? e * (1.0f - 0.5f * e) : __logf(1.0f + e);
    const float sg = __fdividef(x >= 0.0f ? 1.0f : e, 1.0f + e);
    a0 += (double)(fmaxf(x, 0.0f) - t * x + l);
    if (t == 0.0f) { a1 += (double)sg; a2 += 1.0; }
    return __fsub_rn(sg, t);
}

template <int OCC>  // CTAs per SM
__global__ void __launch_bounds__(LOSS_THREADS, OCC) loss_dense_kernel(Loss3K k)
{
    __shared__ double s_red[LOSS_THREADS / 32][3];
    const LossScale &S = k.s[blockIdx.y];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double a0 = 0, a1 = 0, a2 = 0;
    const long long stride = (long long)gridDim.x * LOSS_THREADS;
    const int so = S.v.so;
    const long long first = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x;
    if (so == 1 && (((unsigned long long)S.head | (unsigned long long)S.gobj | (unsigned long long)S.v.obj) & 15) == 0) {
        // contiguous objectness plane (split form): four consecutive cells per thread, 16-byte loads and stores
        const long long n4 = S.cells >> 2;
        const float4 *x4p = reinterpret_cast<const float4 *>(S.v.obj);
        const int4 *h4p = reinterpret_cast<const int4 *>(S.head);
        float4 *g4p = reinterpret_cast<float4 *>(S.gobj);
        float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xb = xa;
        if (first < n4) xa = __ldg(x4p + first);                     // the logits do not depend on the upstream kernels
        if (first + stride < n4) xb = __ldg(x4p + first + stride);
        pdl_wait();
        for (long long g = first; g < n4; g += 2 * stride) {
            const bool two = g + stride < n4;
            if (g != first) { xa = __ldg(x4p + g); if (two) xb = __ldg(x4p + g + stride); }
            const int4 ha = h4p[g];
            int4 hb = make_int4(0, 0, 0, 0);
            if (two) hb = h4p[g + stride];
            float4 r;
            r.x = dense_cell(S, xa.x, ha.x - 1, a0, a1, a2); r.y = dense_cell(S, xa.y, ha.y - 1, a0, a1, a2);
            r.z = dense_cell(S, xa.z, ha.z - 1, a0, a1, a2); r.w = dense_cell(S, xa.w, ha.w - 1, a0, a1, a2);
            g4p[g] = r;
            if (two) {
                r.x = dense_cell(S, xb.x, hb.x - 1, a0, a1, a2); r.y = dense_cell(S, xb.y, hb.y - 1, a0, a1, a2);
                r.z = dense_cell(S, xb.z, hb.z - 1, a0, a1, a2); r.w = dense_cell(S, xb.w, hb.w - 1, a0, a1, a2);
                g4p[g + stride] = r;
            }
        }
        if (blockIdx.x == 0 && threadIdx.x < (int)(S.cells & 3)) {   // cells beyond the last group of four
            const long long c = (n4 << 2) + threadIdx.x;
            S.gobj[c] = dense_cell(S, __ldg(S.v.obj + c), S.head[c] - 1, a0, a1, a2);
        }
    } else {
        // four cells per thread and pass, their strided loads in flight together (same per-thread order of the sums)
        constexpr int DENSE_PER = 4;
        // the logits do not depend on the upstream kernels: the first batch of loads is issued before the wait
        float xs[DENSE_PER];
#pragma unroll
        for (int u = 0; u < DENSE_PER; ++u) {
            const long long c = first + u * stride;
            xs[u] = 0.f;
            if (c < S.cells) xs[u] = so == 1 ? __ldg(S.v.obj + c) : ld_stride_f32(S.v.obj + c * so);
        }
        pdl_wait();
        for (long long c0 = first; c0 < S.cells; c0 += DENSE_PER * stride) {
            int ws[DENSE_PER];
#pragma unroll
            for (int u = 0; u < DENSE_PER; ++u) {
                const long long c = c0 + u * stride;
                ws[u] = -1;
                if (c < S.cells) {
                    if (c0 != first) xs[u] = so == 1 ? __ldg(S.v.obj + c) : ld_stride_f32(S.v.obj + c * so);
                    ws[u] = S.head[c] - 1;
                }
            }
#pragma unroll
            for (int u = 0; u < DENSE_PER; ++u) {
                const long long c = c0 + u * stride;
                if (c >= S.cells) break;
                S.gobj[c] = dense_cell(S, xs[u], ws[u], a0, a1, a2);
            }
        }
    }
    pdl_launch_dependents();
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
    pdl_wait();
    for (int i = t; i < 3 * k.C; i += 256) S.hist_out[i] = S.hist[i];
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

// ---- multi-GPU: the per-shard terms of the big-batch loss and their combination (shard.allreduce_loss_terms) ----
// pack[s] = {lbox*M, lconf*cells, lcls*M*C, M, cells}: sums that add up over image shards (one all-reduce of 15 doubles)
__global__ void loss_pack_kernel(const double *scalars /*[3,8]*/, double c0, double c1, double c2, double C, double *pack /*[3,5]*/)
{
    const int s = threadIdx.x;
    if (s >= 3) return;
    const double *o = scalars + 8 * s;
    const double M = o[6], cells = s == 0 ? c0 : (s == 1 ? c1 : c2);
    pack[5 * s + 0] = o[0] * M; pack[5 * s + 1] = o[1] * cells; pack[5 * s + 2] = o[2] * M * C;
    pack[5 * s + 3] = M; pack[5 * s + 4] = cells;
}
// the loss of the concatenated batch from the summed terms (modules/detection_loss.py:107-110 with global means)
__global__ void loss_combine_kernel(const double *pack, double box_w, double conf_w, double class_w, double w0, double w1, double w2,
                                    double C, double *out)
{
    if (threadIdx.x != 0) return;
    double lbox = 0, lconf = 0, lcls = 0;
    for (int s = 0; s < 3; ++s) {
        const double *q = pack + 5 * s;
        const double w = s == 0 ? w0 : (s == 1 ? w1 : w2), M = q[3];
        lbox += w * (M > 0 ? q[0] / M : 0.0);
        lconf += w * (q[1] / q[4]);
        lcls += w * (M > 0 ? q[2] / (M * C) : 0.0);
    }
    *out = box_w * lbox + conf_w * lconf + class_w * lcls;
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
// (measured: the same kernel without the loads runs 352 instead of 396 us).  So the backward first pulls the
// residuals into L2 marked evict-last (this kernel, ~5 us), the write stream uses evict-first stores, and the streaming
// kernel's loads -- which hit L2 -- demote each line to evict-first as they go: nothing stays pinned after the call.
// (Round 1 marked the lines evict-last already when the forward stored them; a forward that is never followed by its
// backward then leaves 26 MB of high-priority lines behind per call, and several of those starve every later kernel
// of L2.)
__global__ void __launch_bounds__(256) l2_pin_kernel(const float4 *p, long long n4)
{
    float acc = 0.f;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    pdl_launch_dependents();  // the fill kernel may become resident and prepare its chunk images meanwhile
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float4 v;
        asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i), "l"(pol));
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 1.2345e-30f) asm volatile("trap;");  // keeps the loads alive
}

// The backward reads each residual exactly once: the load demotes its L2 line to evict-first, so the evict-last
// lines of one step do not pile up in L2 when the next forward's workspace lies at another address.
__device__ __forceinline__ unsigned long long l2_evict_first_policy()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float ld_last_use_f32(const float *p, unsigned long long pol)
{
    float v;
    asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ld_last_use_f32x4(const float4 *p, unsigned long long pol)
{
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}

// interleaved rows only (v.g_obj is the gradient tensor, v.so its row length D)
__global__ void __launch_bounds__(BWD_WARPS * 32) loss_bwd_stream_kernel(Loss3K k)
{
    extern __shared__ __align__(128) float bwd_smem[];  // [BWD_WARPS][32*D]
    const int D = k.s[0].v.so;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int chunk_floats = 32 * D, chunk_f4 = 8 * D;
    float *im = bwd_smem + (size_t)wid * chunk_floats;
    for (int i = lane; i < chunk_floats; i += 32) im[i] = 0.f;  // only column 0 of a row ever changes
    __syncwarp();
    const float4 *im4 = reinterpret_cast<const float4 *>(im);
    pdl_wait();  // (launched behind the pinning pass)

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
    const unsigned long long lu_pol = l2_evict_first_policy();
    float gq[BWD_AHEAD];
#pragma unroll
    for (int a = 0; a < BWD_AHEAD; ++a) {
        gq[a] = 0.f;
        const long long ga = gw + a * nw;
        if (ga < tot) { int s2; long long r2; locate(ga, s2, r2); gq[a] = ld_last_use_f32(k.s[s2].gobj + r2 + lane, lu_pol); }
    }
    for (long long g = gw; g < tot; g += nw) {
        int si; long long row0;
        locate(g, si, row0);
        float4 *dst = reinterpret_cast<float4 *>(k.s[si].v.g_obj + row0 * D);
        im[lane * D] = (si == 0 ? cf0 : (si == 1 ? cf1 : cf2)) * gq[0];
#pragma unroll
        for (int a = 0; a + 1 < BWD_AHEAD; ++a) gq[a] = gq[a + 1];
        gq[BWD_AHEAD - 1] = 0.f;
        const long long ga = g + BWD_AHEAD * nw;
        if (ga < tot) { int s2; long long r2; locate(ga, s2, r2); gq[BWD_AHEAD - 1] = ld_last_use_f32(k.s[s2].gobj + r2 + lane, lu_pol); }
        __syncwarp();
        for (int f = lane; f < chunk_f4; f += 32) __stcs(dst + f, im4[f]);  // shared-memory image -> 512-byte coalesced, evict-first stores
        __syncwarp();
    }
    pdl_launch_dependents();

    // rows beyond the last full chunk of a scale (cells not a multiple of 32): plain stores by one CTA
    if (blockIdx.x == 0) {
        for (int s = 0; s < 3; ++s) {
            const LossScale &S = k.s[s];
            const float cf = s == 0 ? cf0 : (s == 1 ? cf1 : cf2);
            for (long long row = (S.cells & ~31LL) + wid; row < S.cells; row += BWD_WARPS)
                for (int col = lane; col < D; col += 32) S.v.g_obj[row * D + col] = col == 0 ? cf * S.gobj[row] : 0.f;
        }
    }
}

// split layout: the objectness plane of the gradient (the class / box planes are cleared by cudaMemsetAsync)
__global__ void __launch_bounds__(256) loss_bwd_conf_kernel(Loss3K k)
{
    const LossScale &S = k.s[blockIdx.y];
    const float cf = bwd_scales(k, S).conf;
    const long long n4 = S.cells >> 2;
    const float4 *src = reinterpret_cast<const float4 *>(S.gobj);
    float4 *dst = reinterpret_cast<float4 *>(S.v.g_obj);
    const unsigned long long lu_pol = l2_evict_first_policy();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float4 v = ld_last_use_f32x4(src + i, lu_pol);
        v.x *= cf; v.y *= cf; v.z *= cf; v.w *= cf;
        dst[i] = v;
    }
    if (blockIdx.x == 0)
        for (long long c = (n4 << 2) + threadIdx.x; c < S.cells; c += 256) S.v.g_obj[c] = cf * S.gobj[c];
}

template <int CT>  // compile-time class count (80: no bounds predicates in the unrolled row loop); 0 = runtime
__global__ void __launch_bounds__(LOSS_THREADS) loss_bwd_rows_kernel(Loss3K k)
{
    const LossScale &S = k.s[blockIdx.y];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, gl = lane & 7;
    const int M = *S.M;
    if (M <= 0) return;
    const BwdScales sc = bwd_scales(k, S);
    const int C = CT ? CT : k.C;
    constexpr bool kFull = CT != 0 && CT % (8 * ROWS_UNROLL) == 0;
    pdl_wait();  // the rows below were cleared by the kernel / memsets launched before this one
    for (long long mb = ((long long)blockIdx.x * (LOSS_THREADS / 32) + wid) * 4; mb < M;
         mb += (long long)gridDim.x * (LOSS_THREADS / 32) * 4) {
        const long long m = mb + (lane >> 3);
        if (m >= M) continue;
        // everything addressed by m is fetched at once (one memory round trip), then the ownership test
        const int nx = S.next[m], cell = S.cell[m], cls_m = S.cls[m];
        const unsigned char su = S.succ[m];
        const float4 gq_m = S.gbox[m];
        if (nx != -1) continue;  // the first match linked into the cell owns the row
        int n = 0, c1 = -1, c2 = -1;
        float gb[4] = {0.f, 0.f, 0.f, 0.f};
        int lhead = (int)m;
        if (!su) {  // the only match of its cell (the usual case): everything is addressed by m, no list walk
            gb[0] = gq_m.x; gb[1] = gq_m.y; gb[2] = gq_m.z; gb[3] = gq_m.w;
            c1 = cls_m;
            n = 1;
        } else {           // sums in double: the list order depends on the block schedule, the rounded sum must not
            lhead = S.head[cell] - 1;
            double gd[4] = {0, 0, 0, 0};
            for (int j = lhead; j >= 0; j = S.next[j]) {
                const float4 gq = S.gbox[j];
                gd[0] += gq.x; gd[1] += gq.y; gd[2] += gq.z; gd[3] += gq.w;
                if (n == 0) c1 = S.cls[j]; else if (n == 1) c2 = S.cls[j];
                ++n;
            }
            gb[0] = (float)gd[0]; gb[1] = (float)gd[1]; gb[2] = (float)gd[2]; gb[3] = (float)gd[3];
        }
        const float *xrow = S.v.cls + (long long)cell * S.v.sc;
        float *grow = S.v.g_cls + (long long)cell * S.v.sc;
        // class column c: cls*(n*(sg - cn) - (cp - cn)*hits(c)) = ka*sg - kb - kc*hits(c); the (at most two) columns
        // with hits are fixed up after the row loop by the lane that wrote them
        const float ka = sc.cls * (float)n, kb = ka * k.cn, kc = sc.cls * (k.cp - k.cn);
        for (int cb = 0; cb < C; cb += 8 * ROWS_UNROLL) {
            float x[ROWS_UNROLL];
#pragma unroll
            for (int u = 0; u < ROWS_UNROLL; ++u) {  // all loads of the batch in flight before the first store
                const int col = cb + 8 * u + gl;
                x[u] = (kFull || col < C) ? __ldg(xrow + col) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < ROWS_UNROLL; ++u) {
                const int col = cb + 8 * u + gl;
                if (!kFull && col >= C) continue;
                grow[col] = ka * sigmoid_fast(x[u]) - kb;
            }
        }
        if (n <= 2) {  // the lane that stored a column fixes it up: no cross-lane ordering needed
            if (c1 >= 0 && gl == (c1 & 7)) grow[c1] -= kc;
            if (c2 >= 0 && gl == (c2 & 7)) grow[c2] -= kc;
        } else {  // three or more matches on one cell: one subtraction per match, by one lane, after the group's
                  // stores are visible to it (n is uniform over the eight lanes of the group)
            __syncwarp(0xffu << (lane & 24));
            if (gl == 0)
                for (int j = lhead; j >= 0; j = S.next[j]) grow[S.cls[j]] -= kc;
        }
        if (gl < 4) S.v.g_box[(long long)cell * S.v.sb + gl] = sc.box * (gl == 0 ? gb[0] : gl == 1 ? gb[1] : gl == 2 ? gb[2] : gb[3]);
    }
}

// The same for the split form (class rows of C floats and box rows of 4 floats, 16-byte aligned, C % 4 == 0): four
// lanes per match with 16-byte loads and stores -- eight matches of a warp in flight instead of four, a quarter of
// the memory instructions.  Lane g of a group owns the float4 slots g, g+4, g+8, ... of the class row.
template <int CT>  // compile-time class count (80) or 0 = runtime
__global__ void __launch_bounds__(LOSS_THREADS) loss_bwd_rows_vec_kernel(Loss3K k)
{
    const LossScale &S = k.s[blockIdx.y];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, gl = lane & 3;
    const int M = *S.M;
    if (M <= 0) return;
    const BwdScales sc = bwd_scales(k, S);
    const int C = CT ? CT : k.C, C4 = C >> 2;
    constexpr int NV = CT ? (CT / 4 + 3) / 4 : 8;   // float4 slots per lane (runtime C: up to 128 classes per pass)
    pdl_wait();  // the rows below were cleared by the memsets launched before this kernel
    for (long long mb = ((long long)blockIdx.x * (LOSS_THREADS / 32) + wid) * 8; mb < M;
         mb += (long long)gridDim.x * (LOSS_THREADS / 32) * 8) {
        const long long m = mb + (lane >> 2);
        if (m >= M) continue;
        const int nx = S.next[m], cell = S.cell[m], cls_m = S.cls[m];
        const unsigned char su = S.succ[m];
        const float4 gq_m = S.gbox[m];
        if (nx != -1) continue;  // the first match linked into the cell owns the row
        int n = 0, c1 = -1, c2 = -1;
        float4 gb = gq_m;
        int lhead = (int)m;
        if (!su) {
            c1 = cls_m;
            n = 1;
        } else {           // sums in double: the list order depends on the block schedule, the rounded sum must not
            lhead = S.head[cell] - 1;
            double gd[4] = {0, 0, 0, 0};
            for (int j = lhead; j >= 0; j = S.next[j]) {
                const float4 gq = S.gbox[j];
                gd[0] += gq.x; gd[1] += gq.y; gd[2] += gq.z; gd[3] += gq.w;
                if (n == 0) c1 = S.cls[j]; else if (n == 1) c2 = S.cls[j];
                ++n;
            }
            gb = make_float4((float)gd[0], (float)gd[1], (float)gd[2], (float)gd[3]);
        }
        const float4 *xrow = reinterpret_cast<const float4 *>(S.v.cls + (long long)cell * C);
        float4 *grow = reinterpret_cast<float4 *>(S.v.g_cls + (long long)cell * C);
        float *grow1 = S.v.g_cls + (long long)cell * C;
        const float ka = sc.cls * (float)n, kb = ka * k.cn, kc = sc.cls * (k.cp - k.cn);
        for (int vb = 0; vb < C4; vb += 4 * NV) {
            float4 x[NV];
#pragma unroll
            for (int u = 0; u < NV; ++u) {  // all loads of the batch in flight before the first store
                const int v = vb + 4 * u + gl;
                x[u] = v < C4 ? __ldg(xrow + v) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const int v = vb + 4 * u + gl;
                if (v >= C4) continue;
                float4 g;
                g.x = ka * sigmoid_fast(x[u].x) - kb; g.y = ka * sigmoid_fast(x[u].y) - kb;
                g.z = ka * sigmoid_fast(x[u].z) - kb; g.w = ka * sigmoid_fast(x[u].w) - kb;
                grow[v] = g;
            }
        }
        if (n <= 2) {  // the lane that stored a column fixes it up: no cross-lane ordering needed
            if (c1 >= 0 && gl == ((c1 >> 2) & 3)) grow1[c1] -= kc;
            if (c2 >= 0 && gl == ((c2 >> 2) & 3)) grow1[c2] -= kc;
        } else {
            __syncwarp(0xfu << (lane & 28));
            if (gl == 0)
                for (int j = lhead; j >= 0; j = S.next[j]) grow1[S.cls[j]] -= kc;
        }
        if (gl == 0)
            *reinterpret_cast<float4 *>(S.v.g_box + (long long)cell * 4) = make_float4(sc.box * gb.x, sc.box * gb.y, sc.box * gb.z, sc.box * gb.w);
    }
}

}  // namespace bg
