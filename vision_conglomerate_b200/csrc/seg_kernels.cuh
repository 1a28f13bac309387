// The mask term of the reference's SegmentationLoss (modules/segmentation_loss.py:147-171, 208-231) for
// overlap_masks=True and BCEWithLogits, on top of the fused detection loss (loss_kernels.cuh), which supplies the box /
// objectness / class terms from the same prediction tensors (their mask-coefficient columns ride along as extra
// columns).  Per scale and image i with matches m:
//     pred_m = coefs_m @ protos_i                           coefs = the K columns after the box columns of the matched row
//     t_m    = (target_masks_i == tmask_idx_m)              masks nearest-resized to the protos' size (:152-153)
//     dice_i = mean_m (2 sum(sig t) + e) / (sum(sig) + sum(t) + e),  e = 1e-5            (utils/utils.py:152-172)
//     L_m    = sum(bce(pred_m, t_m) over the crop of m) / (Hp Wp) / (w_m h_m)            (:225, utils.py:130-149; the crop
//              box is the match's own (x, y, w, h) in GRID units, as the reference has it)
//     sl_i   = mean_m (1 - L_m) * (1 - dice_i);   seg_loss = sum_i sl_i / B;   dice_score likewise from round(sig)
// Nothing of size matches x pixels is ever stored: the forward keeps per-match sums, the backward recomputes the
// logits.  This IS a contraction (n_i x K x Hp*Wp per image), done on the FP32 pipes with explicit FMAs: the rtol 1e-5
// parity bar rules out single-pass TF32 / bf16 tensor-core products, and at the reference's sizes (a few dozen matches
// per image, K = 32) the kernels are bound by the transcendental per (match, pixel), not by the 32 FMAs in front of it.
//   seg_count / seg_fill    per-image match lists over the three scales, in the reference's order (scale, then (k, a, t))
//   seg_fwd_kernel          match-major: lane = match, warp = pixel slice; proto tile in shared memory, coefficients in
//                           registers; per-match sums (sig, sig*t, t, rounded ones, cropped bce) -> partials per pixel group
//   seg_reduce / seg_final  fixed-order sums in double -> per-match constants of the backward, per-image and per-scale terms
//   seg_bwd_coef_kernel     match-major again: d loss / d coefs (partials per pixel group) -> seg_bwd_scatter adds them to
//                           the coefficient columns of the dense gradient the detection backward has written
//   seg_bwd_protos_kernel   pixel-major: thread = pixel, its K proto values and K gradient sums in registers, the image's
//                           matches streamed through shared memory: d loss / d protos written once, coalesced, no atomics
#pragma once
#include "loss_kernels.cuh"

namespace bg {

struct SegEntry { int row; int s; float tmask; float area; float x1, y1, x2, y2; };  // 32 bytes
static_assert(sizeof(SegEntry) == 32, "SegEntry");

struct SegK {
    int B, K, D, coef_off, Hp, Wp, HW, Hm, Wm;
    float sy, sx;                 // nearest-neighbour source scale (in / out, fp32) of F.interpolate
    const float *preds[3];        // [cells_s, D]
    float *gpreds[3];             // backward: the dense gradients (already written by the detection backward)
    int cells_per_img[3];         // ny*nx*na
    const float *protos;          // [B, K, HW]
    float *gprotos;
    const float *masks;           // [B, Hm, Wm]
    const int *cell[3];           // assignment outputs per scale (assign_onepass_kernel): row index of the matched cell,
    const float4 *box[3];         //   (x, y, w, h) of the target in grid units relative to the cell,
    const long long *tmask[3];    //   mask index, and the number of matches
    const int *count[3];
    int cap_s;                    // capacity of a scale's match arrays (5*na*nt)
    int *cnt;                     // [B,4] matches of image i on each scale, [3] = all scales
    int *off;                     // [B+1] start of image i's entries in list
    SegEntry *list;               // [cap]
    int cap;
    int G;                        // pixel groups per image = CTAs per image of the match-major kernels
    float *part;                  // forward [G, cap, 6], backward [G, cap, K]
    float4 *mstat;                // [cap] (cL, cD / Den^2, Num, Den): constants of d sl / d pred of the match
    double *img_s;                // [B,3,2] (sl_i, ds_i) per scale
    double *scalars;              // [3,2] caller's: (seg_loss, dice_score) per scale
    float *loss;                  // [1] in: the detection loss; out: + seg_w * sum_s scale_w[s] * seg_loss_s
    double scale_w[3], seg_w;
    const float *go_dev;          // backward: upstream gradient
};

constexpr int SEG_THREADS = 256;
constexpr int SEG_FWD_Q = 6;      // per-match sums: sig, sig*t, t, round(sig), round(sig)*t, cropped bce
constexpr float SEG_DICE_E = 1e-5f;

__global__ void __launch_bounds__(SEG_THREADS) seg_count_kernel(SegK k)
{
    const int i = blockIdx.x;
    int tot = 0;
    for (int s = 0; s < 3; ++s) {
        const int M = min(*k.count[s], k.cap_s);
        int c = 0;
        for (int m0 = 0; m0 < M; m0 += SEG_THREADS) {
            const int m = m0 + threadIdx.x;
            c += __syncthreads_count(m < M && k.cell[s][m] / k.cells_per_img[s] == i && k.cell[s][m] >= 0);
        }
        if (threadIdx.x == 0) k.cnt[4 * i + s] = c;
        tot += c;
    }
    if (threadIdx.x == 0) k.cnt[4 * i + 3] = tot;
}

__global__ void __launch_bounds__(SEG_THREADS) seg_fill_kernel(SegK k)
{
    __shared__ int s_w[SEG_THREADS / 32];
    __shared__ int s_red[SEG_THREADS / 32];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int before = 0;
    for (int j = tid; j < i; j += SEG_THREADS) before += k.cnt[4 * j + 3];
    before = warp_sum(before);
    if (lane == 0) s_red[wid] = before;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < SEG_THREADS / 32; ++w) base += s_red[w];
    if (tid == 0) {
        k.off[i] = base;
        if (i == k.B - 1) k.off[k.B] = base + k.cnt[4 * i + 3];
    }
    for (int s = 0; s < 3; ++s) {
        const int M = min(*k.count[s], k.cap_s);
        for (int m0 = 0; m0 < M; m0 += SEG_THREADS) {
            const int m = m0 + tid;
            int row = -1;
            if (m < M) row = k.cell[s][m];
            const bool f = row >= 0 && row / k.cells_per_img[s] == i;
            const u32 bal = __ballot_sync(0xffffffffu, f);
            if (lane == 0) s_w[wid] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
            for (int w = 0; w < SEG_THREADS / 32; ++w) { if (w < wid) woff += s_w[w]; tot += s_w[w]; }
            const int pos = base + woff + __popc(bal & lanemask_lt());
            if (f && pos < k.cap) {
                const float4 b = k.box[s][m];
                SegEntry e;
                e.row = row; e.s = s; e.tmask = (float)k.tmask[s][m];
                e.area = __fmul_rn(b.z, b.w);
                // crop_section (utils/utils.py:142): x1 = x - w/2, x2 = x + w/2
                e.x1 = __fsub_rn(b.x, __fmul_rn(b.z, 0.5f)); e.y1 = __fsub_rn(b.y, __fmul_rn(b.w, 0.5f));
                e.x2 = __fadd_rn(b.x, __fmul_rn(b.z, 0.5f)); e.y2 = __fadd_rn(b.y, __fmul_rn(b.w, 0.5f));
                k.list[pos] = e;
            }
            base += tot;
            __syncthreads();
        }
    }
}

// pixel tile in shared memory: tile[p][K + 4] (rows 16-byte aligned: a thread reads a pixel's K values with K/4
// broadcast LDS.128), the resized target-mask value and the pixel coordinates of the tile's 256 pixels
template <int K>
struct SegTile {
    static constexpr int P = SEG_THREADS;
    static constexpr int LD = K + 4;
    static constexpr size_t BYTES = (size_t)(P * LD + 3 * P) * sizeof(float);
};

template <int K>
__device__ __forceinline__ void seg_load_tile(const SegK &k, int i, int tl, float *tile, float *mt, float *fx, float *fy)
{
    constexpr int LD = SegTile<K>::LD;
    const int tid = threadIdx.x;
    const int px = tl * SEG_THREADS + tid;
    const float *P = k.protos + (long long)i * K * k.HW;
    if (px < k.HW) {
#pragma unroll
        for (int kk = 0; kk < K; ++kk) tile[tid * LD + kk] = P[(long long)kk * k.HW + px];
        const int y = px / k.Wp, x = px - y * k.Wp;
        const int ys = min((int)floorf(__fmul_rn((float)y, k.sy)), k.Hm - 1), xs = min((int)floorf(__fmul_rn((float)x, k.sx)), k.Wm - 1);
        mt[tid] = k.masks[((long long)i * k.Hm + ys) * k.Wm + xs];
        fx[tid] = (float)x; fy[tid] = (float)y;
    } else {
#pragma unroll
        for (int kk = 0; kk < K; ++kk) tile[tid * LD + kk] = 0.0f;
        mt[tid] = -1.0f; fx[tid] = -1e30f; fy[tid] = -1e30f;
    }
}

template <int K>
__device__ __forceinline__ float seg_dot(const float (&c)[K], const float *row, float (&v)[K])
{
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
        const float4 u = reinterpret_cast<const float4 *>(row)[q];
        v[4 * q] = u.x; v[4 * q + 1] = u.y; v[4 * q + 2] = u.z; v[4 * q + 3] = u.w;
    }
    // four independent chains (a single one waits the FMA latency 32 times in a row: half of the first version's stalls)
    float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
#pragma unroll
    for (int kk = 0; kk < K; kk += 4) {
        d0 = __fmaf_rn(c[kk], v[kk], d0); d1 = __fmaf_rn(c[kk + 1], v[kk + 1], d1);
        d2 = __fmaf_rn(c[kk + 2], v[kk + 2], d2); d3 = __fmaf_rn(c[kk + 3], v[kk + 3], d3);
    }
    return (d0 + d1) + (d2 + d3);
}

template <int K>
__global__ void __launch_bounds__(SEG_THREADS) seg_fwd_kernel(SegK k)
{
    extern __shared__ __align__(16) float seg_sm[];
    constexpr int LD = SegTile<K>::LD, P = SEG_THREADS;
    float *tile = seg_sm, *mt = seg_sm + P * LD, *fx = mt + P, *fy = fx + P;
    const int g = blockIdx.x, i = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = k.cnt[4 * i + 3];
    if (n == 0) return;
    const int o0 = k.off[i];
    const int ntiles = (k.HW + P - 1) / P;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const bool valid = j < n && o0 + j < k.cap;
        SegEntry e;
        e.row = 0; e.s = 0; e.tmask = -2.0f; e.area = 1.0f; e.x1 = e.y1 = 1.0f; e.x2 = e.y2 = 0.0f;
        float c[K];
#pragma unroll
        for (int kk = 0; kk < K; ++kk) c[kk] = 0.0f;
        if (valid) {
            e = k.list[o0 + j];
            const float *cp = k.preds[e.s] + (long long)e.row * k.D + k.coef_off;
#pragma unroll
            for (int kk = 0; kk < K; ++kk) c[kk] = cp[kk];
        }
        float a[SEG_FWD_Q];
#pragma unroll
        for (int q = 0; q < SEG_FWD_Q; ++q) a[q] = 0.0f;
        for (int tl = g; tl < ntiles; tl += k.G) {
            __syncthreads();
            seg_load_tile<K>(k, i, tl, tile, mt, fx, fy);
            __syncthreads();
            const int lim = min(32, k.HW - tl * P - wid * 32);
            for (int p = 0; p < lim; ++p) {
                const int pl = wid * 32 + p;
                float v[K];
                const float dot = seg_dot<K>(c, tile + pl * LD, v);
                const float sg = __fdividef(1.0f, 1.0f + __expf(-dot));
                const bool t = mt[pl] == e.tmask;
                a[0] += sg;
                if (t) { a[1] += sg; a[2] += 1.0f; }
                if (dot > 0.0f) { a[3] += 1.0f; if (t) a[4] += 1.0f; }   // round(sigmoid) = 1
                const float x = fx[pl], y = fy[pl];
                if (x >= e.x1 && x < e.x2 && y >= e.y1 && y < e.y2) a[5] += bce_logits(dot, t ? 1.0f : 0.0f);
            }
        }
        // the eight warps' sums of the same 32 matches, added in warp order
        __syncthreads();
        float *red = tile;
#pragma unroll
        for (int q = 0; q < SEG_FWD_Q; ++q) red[(wid * SEG_FWD_Q + q) * 32 + lane] = a[q];
        __syncthreads();
        if (wid == 0 && valid) {
#pragma unroll
            for (int q = 0; q < SEG_FWD_Q; ++q) {
                float sum = 0.0f;
                for (int w = 0; w < SEG_THREADS / 32; ++w) sum += red[(w * SEG_FWD_Q + q) * 32 + lane];
                k.part[((long long)g * k.cap + o0 + j) * SEG_FWD_Q + q] = sum;
            }
        }
    }
}

__device__ __forceinline__ double seg_block_sum(double v, double *s_red /*[8]*/)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_red[wid] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < SEG_THREADS / 32; ++w) t += s_red[w];
    return t;
}

__global__ void __launch_bounds__(SEG_THREADS) seg_reduce_kernel(SegK k)
{
    __shared__ double s_red[SEG_THREADS / 32];
    const int i = blockIdx.x, tid = threadIdx.x;
    const int o0 = k.off[i];
    int seg0 = 0;
    for (int s = 0; s < 3; ++s) {
        const int ns = k.cnt[4 * i + s];
        double sd = 0.0, sl = 0.0, sr = 0.0;
        for (int j = seg0 + tid; j < seg0 + ns; j += SEG_THREADS) {
            if (o0 + j >= k.cap) break;
            double q[SEG_FWD_Q];
#pragma unroll
            for (int u = 0; u < SEG_FWD_Q; ++u) q[u] = 0.0;
            for (int g = 0; g < k.G; ++g)
#pragma unroll
                for (int u = 0; u < SEG_FWD_Q; ++u) q[u] += (double)k.part[((long long)g * k.cap + o0 + j) * SEG_FWD_Q + u];
            const SegEntry e = k.list[o0 + j];
            const float num = 2.0f * (float)q[1] + SEG_DICE_E, den = (float)q[0] + (float)q[2] + SEG_DICE_E;
            const float dice = num / den;
            const float dice_r = (2.0f * (float)q[4] + SEG_DICE_E) / ((float)q[3] + (float)q[2] + SEG_DICE_E);
            const float L = ((float)q[5] / (float)k.HW) / e.area;
            k.mstat[o0 + j] = make_float4(0.0f, 0.0f, num, den);
            sd += (double)dice; sl += (double)(1.0f - L); sr += (double)dice_r;
        }
        sd = seg_block_sum(sd, s_red);
        sl = seg_block_sum(sl, s_red);
        sr = seg_block_sum(sr, s_red);
        double A = 0.0, Dl = 0.0;
        if (ns > 0) { A = sl / ns; Dl = 1.0 - sd / ns; }
        if (tid == 0) {
            k.img_s[(i * 3 + s) * 2 + 0] = ns > 0 ? A * Dl : 0.0;
            k.img_s[(i * 3 + s) * 2 + 1] = ns > 0 ? sr / ns : 0.0;
        }
        for (int j = seg0 + tid; j < seg0 + ns; j += SEG_THREADS) {
            if (o0 + j >= k.cap) break;
            const SegEntry e = k.list[o0 + j];
            float4 ms = k.mstat[o0 + j];
            ms.x = (float)(-Dl / ns) / ((float)k.HW * e.area);      // d sl / d L_m  *  d L_m / d (bce sum)
            ms.y = (float)(-A / ns) / (ms.w * ms.w);                 // d sl / d dice_m  /  Den^2
            k.mstat[o0 + j] = ms;
        }
        seg0 += ns;
    }
}

__global__ void __launch_bounds__(SEG_THREADS) seg_final_kernel(SegK k)
{
    __shared__ double s_red[SEG_THREADS / 32];
    double lseg = 0.0;
    for (int s = 0; s < 3; ++s) {
        double a = 0.0, d = 0.0;
        for (int i = threadIdx.x; i < k.B; i += SEG_THREADS) { a += k.img_s[(i * 3 + s) * 2]; d += k.img_s[(i * 3 + s) * 2 + 1]; }
        a = seg_block_sum(a, s_red) / k.B;
        d = seg_block_sum(d, s_red) / k.B;
        if (threadIdx.x == 0) { k.scalars[2 * s] = a; k.scalars[2 * s + 1] = d; }
        lseg += k.scale_w[s] * a;
    }
    if (threadIdx.x == 0) k.loss[0] = (float)((double)k.loss[0] + k.seg_w * lseg);
}

// d sl / d pred of (match, pixel), without the factor go * seg_w * scale_w[s] / B
__device__ __forceinline__ float seg_dpred(const float4 ms, float sg, bool t, bool in_crop)
{
    float d = ms.y * sg * (1.0f - sg) * ((t ? 2.0f * ms.w : 0.0f) - ms.z);
    if (in_crop) d += ms.x * (sg - (t ? 1.0f : 0.0f));
    return d;
}

template <int K>
__global__ void __launch_bounds__(SEG_THREADS) seg_bwd_coef_kernel(SegK k)
{
    extern __shared__ __align__(16) float seg_sm[];
    constexpr int LD = SegTile<K>::LD, P = SEG_THREADS;
    float *tile = seg_sm, *mt = seg_sm + P * LD, *fx = mt + P, *fy = fx + P;
    const int g = blockIdx.x, i = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = k.cnt[4 * i + 3];
    if (n == 0) return;
    const int o0 = k.off[i];
    const int ntiles = (k.HW + P - 1) / P;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const bool valid = j < n && o0 + j < k.cap;
        SegEntry e;
        e.row = 0; e.s = 0; e.tmask = -2.0f; e.area = 1.0f; e.x1 = e.y1 = 1.0f; e.x2 = e.y2 = 0.0f;
        float4 ms = make_float4(0.f, 0.f, 0.f, 1.f);
        float c[K], gc[K];
#pragma unroll
        for (int kk = 0; kk < K; ++kk) { c[kk] = 0.0f; gc[kk] = 0.0f; }
        if (valid) {
            e = k.list[o0 + j];
            ms = k.mstat[o0 + j];
            const float *cp = k.preds[e.s] + (long long)e.row * k.D + k.coef_off;
#pragma unroll
            for (int kk = 0; kk < K; ++kk) c[kk] = cp[kk];
        }
        for (int tl = g; tl < ntiles; tl += k.G) {
            __syncthreads();
            seg_load_tile<K>(k, i, tl, tile, mt, fx, fy);
            __syncthreads();
            const int lim = min(32, k.HW - tl * P - wid * 32);
            for (int p = 0; p < lim; ++p) {
                const int pl = wid * 32 + p;
                float v[K];
                const float dot = seg_dot<K>(c, tile + pl * LD, v);
                const float sg = __fdividef(1.0f, 1.0f + __expf(-dot));
                const bool t = mt[pl] == e.tmask;
                const float x = fx[pl], y = fy[pl];
                const float d = seg_dpred(ms, sg, t, x >= e.x1 && x < e.x2 && y >= e.y1 && y < e.y2);
#pragma unroll
                for (int kk = 0; kk < K; ++kk) gc[kk] = __fmaf_rn(d, v[kk], gc[kk]);
            }
        }
        __syncthreads();
        float *red = tile;   // [8][K][32] <= P * LD floats
#pragma unroll
        for (int kk = 0; kk < K; ++kk) red[(wid * K + kk) * 32 + lane] = gc[kk];
        __syncthreads();
        for (int idx = tid; idx < 32 * K; idx += SEG_THREADS) {
            const int kk = idx >> 5, ln = idx & 31;
            if (j0 + ln < n && o0 + j0 + ln < k.cap) {
                float sum = 0.0f;
                for (int w = 0; w < SEG_THREADS / 32; ++w) sum += red[(w * K + kk) * 32 + ln];
                k.part[((long long)g * k.cap + o0 + j0 + ln) * K + kk] = sum;
            }
        }
    }
}

// adds go * seg_w * scale_w[s] / B * (sum over the pixel groups, in order) to the coefficient columns of the matched rows;
// two matches of one cell (rare) meet in one address: float atomics, like the reference's index_put(accumulate=True)
__global__ void __launch_bounds__(SEG_THREADS) seg_bwd_scatter_kernel(SegK k)
{
    const long long total = (long long)min(k.off[k.B], k.cap) * k.K;
    const float go = k.go_dev ? k.go_dev[0] : 1.0f;
    for (long long idx = (long long)blockIdx.x * SEG_THREADS + threadIdx.x; idx < total; idx += (long long)gridDim.x * SEG_THREADS) {
        const int j = (int)(idx / k.K), kk = (int)(idx - (long long)j * k.K);
        double sum = 0.0;
        for (int g = 0; g < k.G; ++g) sum += (double)k.part[((long long)g * k.cap + j) * k.K + kk];
        const SegEntry e = k.list[j];
        const float w = (float)((double)go * k.seg_w * k.scale_w[e.s] / k.B);
        atomicAdd(k.gpreds[e.s] + (long long)e.row * k.D + k.coef_off + kk, (float)sum * w);
    }
}

template <int K>
__global__ void __launch_bounds__(SEG_THREADS) seg_bwd_protos_kernel(SegK k)
{
    __shared__ __align__(16) float cs[32 * K];
    __shared__ SegEntry s_e[32];
    __shared__ float4 s_ms[32];
    __shared__ float s_w[32];
    const int i = blockIdx.y, tid = threadIdx.x;
    const int px = blockIdx.x * SEG_THREADS + tid;
    const int n = min(k.cnt[4 * i + 3], k.cap - k.off[i]);
    const int o0 = k.off[i];
    const bool live = px < k.HW;
    const float *P = k.protos + (long long)i * K * k.HW;
    float pr[K], gp[K];
#pragma unroll
    for (int kk = 0; kk < K; ++kk) { pr[kk] = live ? P[(long long)kk * k.HW + px] : 0.0f; gp[kk] = 0.0f; }
    float mv = -1.0f, x = -1e30f, y = -1e30f;
    if (live) {
        const int yy = px / k.Wp, xx = px - yy * k.Wp;
        const int ys = min((int)floorf(__fmul_rn((float)yy, k.sy)), k.Hm - 1), xs = min((int)floorf(__fmul_rn((float)xx, k.sx)), k.Wm - 1);
        mv = k.masks[((long long)i * k.Hm + ys) * k.Wm + xs];
        x = (float)xx; y = (float)yy;
    }
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int nj = min(32, n - j0);
        __syncthreads();
        if (tid < nj) {
            const SegEntry e = k.list[o0 + j0 + tid];
            s_e[tid] = e;
            s_ms[tid] = k.mstat[o0 + j0 + tid];
            s_w[tid] = (float)(k.seg_w * k.scale_w[e.s] / k.B);
        }
        __syncthreads();
        for (int idx = tid; idx < nj * K; idx += SEG_THREADS) {
            const int jj = idx / K, kk = idx - jj * K;
            cs[idx] = k.preds[s_e[jj].s][(long long)s_e[jj].row * k.D + k.coef_off + kk];
        }
        __syncthreads();
        for (int jj = 0; jj < nj; ++jj) {
            float v[K];
            const float dot = seg_dot<K>(pr, cs + jj * K, v);
            const float sg = __fdividef(1.0f, 1.0f + __expf(-dot));
            const SegEntry &e = s_e[jj];
            const bool t = mv == e.tmask;
            const float d = seg_dpred(s_ms[jj], sg, t, x >= e.x1 && x < e.x2 && y >= e.y1 && y < e.y2) * s_w[jj];
#pragma unroll
            for (int kk = 0; kk < K; ++kk) gp[kk] = __fmaf_rn(d, v[kk], gp[kk]);
        }
    }
    if (live) {
        const float go = k.go_dev ? k.go_dev[0] : 1.0f;
        float *G = k.gprotos + (long long)i * K * k.HW;
#pragma unroll
        for (int kk = 0; kk < K; ++kk) G[(long long)kk * k.HW + px] = gp[kk] * go;
    }
}

// ------------------------------------------------------------------------------------------------
// inference_seg.post_process_preds lines 115-117: masks = sigmoid(coefs @ protos_i) on the protos' grid, bilinear resize
// (align_corners=False) to the image, > 0.5.  Rows are the kept detections, image by image (row_off [B+1]).
//   seg_lowres_kernel     thread = proto pixel, eight rows at a time (coefficients broadcast from shared memory, each
//                         proto value read once per eight rows), accurate sigmoid -> low [n, Hp*Wp]
//   seg_upsample_kernel   thread = output pixel (four of a row where the width allows): ATen's source index and weights
//                         (upsample_bilinear2d: src = scale*(dst+0.5)-0.5 clamped at 0, scale = in/out in fp32), threshold,
//                         one byte per pixel
// ------------------------------------------------------------------------------------------------
struct SegMaskK {
    int B, K, Hp, Wp, HW, H, W;
    long long n;
    const float *coefs;        // [n, K]
    const int *row_off;        // [B+1]
    const float *protos;       // [B, K, HW]
    float *low;                // [n, HW]
    unsigned char *out;        // [n, H, W] 0 / 1
    float ry, rx;              // Hp / H, Wp / W in fp32
};
constexpr int SEGM_ROWS = 16;
constexpr int SEGM_KMAX = 64;

__global__ void __launch_bounds__(SEG_THREADS) seg_lowres_kernel(SegMaskK k)
{
    __shared__ __align__(16) float cs[SEGM_ROWS][SEGM_KMAX];   // the coefficients of 16 rows, zero-padded to a multiple of four
    const int i = blockIdx.y, tid = threadIdx.x;
    const int px = blockIdx.x * SEG_THREADS + tid;
    const int r0 = k.row_off[i], r1 = k.row_off[i + 1];
    const float *P = k.protos + (long long)i * k.K * k.HW;
    const int K4 = (k.K + 3) & ~3;
    for (int r = r0; r < r1; r += SEGM_ROWS) {
        const int nr = min(SEGM_ROWS, r1 - r);
        __syncthreads();
        for (int idx = tid; idx < SEGM_ROWS * K4; idx += SEG_THREADS) {
            const int rr = idx / K4, kk = idx - rr * K4;
            cs[rr][kk] = (rr < nr && kk < k.K) ? k.coefs[(long long)(r + rr) * k.K + kk] : 0.0f;
        }
        __syncthreads();
        if (px < k.HW) {
            float acc[SEGM_ROWS];
#pragma unroll
            for (int rr = 0; rr < SEGM_ROWS; ++rr) acc[rr] = 0.0f;
            for (int kk = 0; kk < K4; kk += 4) {   // products added in ascending k, as a row-times-matrix product does
                float p[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) p[j] = kk + j < k.K ? P[(long long)(kk + j) * k.HW + px] : 0.0f;
#pragma unroll
                for (int rr = 0; rr < SEGM_ROWS; ++rr) {
                    const float4 c = *reinterpret_cast<const float4 *>(&cs[rr][kk]);   // one broadcast load per four products
                    acc[rr] = __fmaf_rn(c.x, p[0], acc[rr]);
                    acc[rr] = __fmaf_rn(c.y, p[1], acc[rr]);
                    acc[rr] = __fmaf_rn(c.z, p[2], acc[rr]);
                    acc[rr] = __fmaf_rn(c.w, p[3], acc[rr]);
                }
            }
#pragma unroll
            for (int rr = 0; rr < SEGM_ROWS; ++rr)
                if (rr < nr) k.low[(long long)(r + rr) * k.HW + px] = sigmoid_acc(acc[rr]);
        }
    }
}

__device__ __forceinline__ void segm_axis(float scale, int dst, int n_in, int &i0, int &i1, float &l0, float &l1)
{
    const float src = fmaxf(__fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f), 0.0f);
    i0 = (int)src;
    i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    l1 = __fsub_rn(src, (float)i0);
    l0 = __fsub_rn(1.0f, l1);
}

constexpr int SEGM_YCHUNK = 16;      // output lines per block: a thread keeps the x-axis indices / weights of its pixels for all of them
constexpr int SEGM_STAGE = 2560;     // floats of the low-resolution lines a block stages in shared memory (else it reads them through L1)

// the lines of one block: VEC output pixels per thread along x; the low-resolution lines come from the block's
// shared-memory window (32-bit shared addresses) or, when the window does not fit, from the global array
template <int VEC, bool STAGED>
__device__ __forceinline__ void segm_lines(const SegMaskK &k, const float *__restrict__ L, const float *s_low, int ylo,
                                           const int4 *s_y, int nl, unsigned char *out0)
{
    const int WV = k.W / VEC;
    for (int xv = threadIdx.x; xv < WV; xv += blockDim.x) {
        int x0[VEC], x1[VEC];
        float w0[VEC], w1[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            segm_axis(k.rx, xv * VEC + v, k.Wp, x0[v], x1[v], w0[v], w1[v]);
            asm volatile("" : "+r"(x0[v]), "+r"(x1[v]));   // keep the indices (else they are recomputed from the floats per use)
        }
        // the horizontal interpolation of a source line serves every output line between the same two source lines
        // (four of them when the masks are enlarged four times): kept in registers, and a line that was the lower one
        // becomes the upper one without being read again.  Same operations in the same order for every pixel.
        float top[VEC], bot[VEC];
        int py0 = -1, py1 = -1;
        unsigned char *o = out0 + (long long)xv * VEC;
        for (int l = 0; l < nl; ++l, o += k.W) {
            const int4 yy = s_y[l];
            const int y0 = yy.x, y1 = yy.y;
            const float h0 = __int_as_float(yy.z), h1 = __int_as_float(yy.w);
            if (y0 != py0) {
                if (y0 == py1) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) top[v] = bot[v];
                } else {
                    const float *L0 = STAGED ? s_low + (y0 - ylo) * k.Wp : L + (long long)y0 * k.Wp;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) top[v] = __fadd_rn(__fmul_rn(w0[v], L0[x0[v]]), __fmul_rn(w1[v], L0[x1[v]]));
                }
                py0 = y0;
            }
            if (y1 != py1) {
                if (y1 == y0) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) bot[v] = top[v];
                } else {
                    const float *L1 = STAGED ? s_low + (y1 - ylo) * k.Wp : L + (long long)y1 * k.Wp;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) bot[v] = __fadd_rn(__fmul_rn(w0[v], L1[x0[v]]), __fmul_rn(w1[v], L1[x1[v]]));
                }
                py1 = y1;
            }
            unsigned word = 0;
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                word |= (__fadd_rn(__fmul_rn(h0, top[v]), __fmul_rn(h1, bot[v])) > 0.5f ? 1u : 0u) << (8 * v);
            if (VEC == 4) *reinterpret_cast<unsigned *>(o) = word;
            else o[0] = (unsigned char)word;
        }
    }
}

template <int VEC>  // output pixels per thread along x (4 when W % 4 == 0: one 32-bit store)
__global__ void __launch_bounds__(SEG_THREADS) seg_upsample_kernel(SegMaskK k)
{
    __shared__ float s_low[SEGM_STAGE];
    __shared__ int4 s_y[SEGM_YCHUNK];   // per output line of the block: the two source lines and their weights, worked out once
    const int ychunks = (k.H + SEGM_YCHUNK - 1) / SEGM_YCHUNK;
    const long long r = blockIdx.x / ychunks;
    const int oy0 = (int)(blockIdx.x - r * ychunks) * SEGM_YCHUNK, oy1 = min(oy0 + SEGM_YCHUNK, k.H);
    const float *L = k.low + r * k.HW;
    // the low-resolution lines this block's output lines interpolate between (the source index is monotone in the line)
    int ylo, yhi, t0, t1;
    float u0, u1;
    segm_axis(k.ry, oy0, k.Hp, ylo, t1, u0, u1);
    segm_axis(k.ry, oy1 - 1, k.Hp, t0, yhi, u0, u1);
    const bool staged = (yhi - ylo + 1) * k.Wp <= SEGM_STAGE;
    if (staged) {
        const int cnt = (yhi - ylo + 1) * k.Wp;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) s_low[i] = L[(long long)ylo * k.Wp + i];
    }
    if (threadIdx.x < oy1 - oy0) {
        int y0, y1;
        float h0, h1;
        segm_axis(k.ry, oy0 + threadIdx.x, k.Hp, y0, y1, h0, h1);
        s_y[threadIdx.x] = make_int4(y0, y1, __float_as_int(h0), __float_as_int(h1));
    }
    __syncthreads();
    unsigned char *out0 = k.out + (r * k.H + oy0) * k.W;
    if (staged) segm_lines<VEC, true>(k, L, s_low, ylo, s_y, oy1 - oy0, out0);
    else segm_lines<VEC, false>(k, L, s_low, ylo, s_y, oy1 - oy0, out0);
}

}  // namespace bg
