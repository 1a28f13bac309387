// Segmented greedy NMS.  Two modes share the per-segment score sort and the emission:
//
//  * grid mode (0.05 <= thr < 1): IoU > t bounds the centre distance of a pair by (1-t)/t times the smaller
//    extent per axis (DESIGN.md has the derivation), so every segment gets a uniform grid over its box centres,
//    boxes are put in cell order by a counting sort and nms_grid_pairs_kernel tests a box only against the cells
//    in its reach (forward half of the neighbourhood, bf16 extent prefilter).  Overlapping pairs go to an edge
//    list (earlier score position -> later one) and nms_resolve_kernel settles the greedy outcome by rounds:
//    a box whose earlier overlapping neighbours are all decided is kept iff none of them is kept -- the same
//    fixed point as the sequential scan.  Work and memory are O(edges), so segments of any size stay cheap.
//  * dense mode (any other threshold): nms_mask_kernel fills a K x K suppression bit matrix in 64x64 tiles and
//    nms_reduce_kernel scans it in score order.
//
// Semantics follow torchvision's CPU nms kernel (the arithmetic behind the reference's
// torchvision.ops.batched_nms call, inference_det.py:77-82): candidates in stable descending score
// order, box j is suppressed by a kept box i iff  (double)(inter / (area_i + area_j - inter)) > thr
// with every operation rounded to fp32 individually (no FMA contraction).
#pragma once
#include "nms.cuh"

namespace bg {

// ------------------------------------------------------------------------------------------------
// exact IoU decision
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool iou_suppresses(const float4 a, const float aa, const float4 b, const float ab,
                                               const IouThr &t)
{
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    float w = __fsub_rn(xx2, xx1), h = __fsub_rn(yy2, yy1);
    if (!(w > 0.0f && h > 0.0f)) {
        if (!t.zero_suppresses) return false;  // inter == 0 -> ovr is 0, -0 or NaN: never > a non-negative threshold
        w = fmaxf(w, 0.0f);
        h = fmaxf(h, 0.0f);
    }
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(aa, ab), inter);
    if (t.fast_ok && uni > 0.0f) {
        // division-free filter: outside the guard band the rounded quotient cannot land on the other side of tdn
        if (inter > __fmul_rn(t.hi, uni)) return true;
        if (inter < __fmul_rn(t.lo, uni)) return false;
    }
    return __fdiv_rn(inter, uni) > t.tdn;
}

// boxes without a positive finite extent have IoU 0 or NaN with everything: they never suppress nor get suppressed
__device__ __forceinline__ bool seg_box_valid(const float4 bx, float &w, float &h, float &cx, float &cy)
{
    w = __fsub_rn(bx.z, bx.x); h = __fsub_rn(bx.w, bx.y);
    cx = 0.5f * bx.x + 0.5f * bx.z; cy = 0.5f * bx.y + 0.5f * bx.w;
    return (w > 0.0f) && (h > 0.0f) && (w < INFINITY) && (h < INFINITY) && (fabsf(cx) < INFINITY) && (fabsf(cy) < INFINITY);
}
// cell of a centre coordinate: monotone in the coordinate, so a conservative interval maps to a conservative cell range
__device__ __forceinline__ int seg_cell_x(const SegGrid &g, float x) { return (int)fminf(fmaxf(floorf(__fmul_rn(__fsub_rn(x, g.mnx), g.invx)), 0.0f), (float)(g.G - 1)); }
__device__ __forceinline__ int seg_cell_y(const SegGrid &g, float y) { return (int)fminf(fmaxf(floorf(__fmul_rn(__fsub_rn(y, g.mny), g.invy)), 0.0f), (float)(g.G - 1)); }
constexpr int SEG_GMAX = 2048;
__host__ __device__ __forceinline__ int seg_grid_dim(long long K)  // about two boxes per cell
{
    int G = 1;
    while (G < SEG_GMAX && 2ll * (G + 1) * (G + 1) <= K) ++G;
    return G;
}

// ------------------------------------------------------------------------------------------------
// block-wide exclusive scan of (long long, long long) pairs, 1024 threads
// ------------------------------------------------------------------------------------------------
struct ScanPair { long long a, b; };
__device__ __forceinline__ ScanPair block_excl_scan_1024(ScanPair v, ScanPair *s_warp /*[33]*/, ScanPair &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ScanPair inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long ua = __shfl_up_sync(0xffffffffu, inc.a, o), ub = __shfl_up_sync(0xffffffffu, inc.b, o);
        if (lane >= o) { inc.a += ua; inc.b += ub; }
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        ScanPair w = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : ScanPair{0, 0};
        ScanPair winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long ua = __shfl_up_sync(0xffffffffu, winc.a, o), ub = __shfl_up_sync(0xffffffffu, winc.b, o);
            if (lane >= o) { winc.a += ua; winc.b += ub; }
        }
        s_warp[lane] = ScanPair{winc.a - w.a, winc.b - w.b};  // exclusive warp offsets
        if (lane == 31) s_warp[32] = winc;                     // block total
    }
    __syncthreads();
    ScanPair off = s_warp[wid];
    total = s_warp[32];
    ScanPair r{off.a + inc.a - v.a, off.b + inc.b - v.b};
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------------------------------------
// kernel T: per-segment prefix tables (one CTA)
// ------------------------------------------------------------------------------------------------
constexpr int PAIR_THREADS = 128;
constexpr int PAIR_CTAS_PER_SM = 12;   // resident CTAs of the pair kernel per SM (<= 42 registers per thread)
constexpr int PAIR_LONG = 24;      // cell-order entries of one (box, grid row) from which the warp shares the run out

__global__ void __launch_bounds__(1024) seg_tables_kernel(SegNms p)
{
    __shared__ ScanPair s_scan[33];
    __shared__ long long s_total;
    const int S = p.hdr->S;
    ScanPair carry{0, 0};
    ScanPair icarry{0, 0};
    long long kcarry = 0;
    for (int base = 0; base < S; base += 1024) {
        const int sg = base + threadIdx.x;
        const long long K = (sg < S) ? p.seg_count[sg] : 0;
        const long long T = (K + 63) >> 6;
        long long cells = 0;
        if (sg < S) { const long long G = seg_grid_dim(K); cells = G * G + 1; }
        ScanPair tot, tot2;
        const ScanPair ex = block_excl_scan_1024(ScanPair{T, K * T}, s_scan, tot);
        const ScanPair ex2 = block_excl_scan_1024(ScanPair{(K + PAIR_THREADS - 1) / PAIR_THREADS, cells}, s_scan, tot2);
        ScanPair tot3;
        const ScanPair ex3 = block_excl_scan_1024(ScanPair{K, 0}, s_scan, tot3);
        if (sg < S) {
            p.tile_prefix[sg] = (int)(carry.a + ex.a);
            p.mask_off[sg] = carry.b + ex.b;
            p.item_prefix[sg] = (int)(icarry.a + ex2.a);
            p.cell_off[sg] = icarry.b + ex2.b;
            p.edge_off[sg] = kcarry + ex3.a;  // candidate prefix, scaled to edge regions below
            p.edge_count[sg] = 0ull;
        }
        carry.a += tot.a;
        carry.b += tot.b;
        icarry.a += tot2.a;
        icarry.b += tot2.b;
        kcarry += tot3.a;
    }
    if (threadIdx.x == 0) {
        p.tile_prefix[S] = (int)carry.a;
        p.mask_off[S] = carry.b;
        p.item_prefix[S] = (int)icarry.a;
        p.cell_off[S] = icarry.b;
        p.edge_off[S] = kcarry;
        s_total = kcarry;
        p.hdr->dense_fits = carry.b <= p.mask_words;
        if (!p.sparse && carry.b > p.mask_words) atomicOr(&p.hdr->status, BG_STATUS_MASK_SPACE);
    }
    if (!p.sparse) return;
    __syncthreads();
    // grid mode: every segment gets a share of the edge list proportional to its candidate count (the double
    // product is monotone in the prefix, so the regions are disjoint)
    const double scale = s_total > 0 ? (double)p.mask_words / (double)s_total : 0.0;
    for (int sg = threadIdx.x; sg <= S; sg += 1024) {
        long long o = (long long)((double)p.edge_off[sg] * scale);
        p.edge_off[sg] = o > p.mask_words ? p.mask_words : o;
    }
}

// ------------------------------------------------------------------------------------------------
// kernel S: per-segment score sort, gather, and (grid mode) the grid over the box centres with its cell order
// ------------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 1024;
constexpr int SORT_CHUNK = 8192;  // keys held in shared memory (64 KB)

__device__ __forceinline__ void bitonic_step_smem(u64 *s, int cnt, int gbase, int k, int j)
{
    for (int t = threadIdx.x; t < (cnt >> 1); t += SORT_THREADS) {
        const int i = 2 * t - (t & (j - 1));
        const int l = i + j;
        const bool asc = (((gbase + i) & k) == 0);
        const u64 a = s[i], b = s[l];
        if ((a > b) == asc) { s[i] = b; s[l] = a; }
    }
}

// sort of the P = E*1024 keys in shared memory by the register-blocked network of common.cuh; descending order is
// the ascending order of the complements
template <int E>
__device__ __forceinline__ void sort_chunk_smem(u64 *s, bool desc, bool merge_only = false)
{
    if (desc) {
        for (int i = threadIdx.x; i < E * 1024; i += SORT_THREADS) s[i] = ~s[i];
        __syncthreads();
    }
    sort_reg_1024<E>(s, merge_only ? E * 1024 : 2);
    if (desc) {
        for (int i = threadIdx.x; i < E * 1024; i += SORT_THREADS) s[i] = ~s[i];
        __syncthreads();
    }
}

// ascending sort of keys[0..K) (global, with room for next_pow2(K) entries) by the whole CTA
__device__ void block_sort_u64(u64 *keys, int K, u64 *s)
{
    const int P = (int)next_pow2((u32)K);
    if (P <= SORT_CHUNK) {
        for (int i = threadIdx.x; i < P; i += SORT_THREADS) s[i] = (i < K) ? keys[i] : ~0ull;
        __syncthreads();
        if (P == 8192) sort_chunk_smem<8>(s, false);
        else if (P == 4096) sort_chunk_smem<4>(s, false);
        else if (P == 2048) sort_chunk_smem<2>(s, false);
        else if (P == 1024) sort_chunk_smem<1>(s, false);
        else {
            for (int k = 2; k <= P; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    bitonic_step_smem(s, P, 0, k, j);
                    __syncthreads();
                }
        }
        for (int i = threadIdx.x; i < K; i += SORT_THREADS) keys[i] = s[i];
    } else {
        for (int i = K + threadIdx.x; i < P; i += SORT_THREADS) keys[i] = ~0ull;
        __syncthreads();
        for (int c = 0; c < P; c += SORT_CHUNK) {  // phase 1: chunks sorted in alternating directions (every stage k <= SORT_CHUNK)
            for (int i = threadIdx.x; i < SORT_CHUNK; i += SORT_THREADS) s[i] = keys[c + i];
            __syncthreads();
            sort_chunk_smem<SORT_CHUNK / 1024>(s, (c & SORT_CHUNK) != 0);
            for (int i = threadIdx.x; i < SORT_CHUNK; i += SORT_THREADS) keys[c + i] = s[i];
            __syncthreads();
        }
        for (int k = 2 * SORT_CHUNK; k <= P; k <<= 1) {  // phase 2: wide strides in global memory (L2 resident)
            for (int j = k >> 1; j >= SORT_CHUNK; j >>= 1) {
                for (int t = threadIdx.x; t < (P >> 1); t += SORT_THREADS) {
                    const int i = 2 * t - (t & (j - 1));
                    const int l = i + j;
                    const bool asc = ((i & k) == 0);
                    const u64 a = keys[i], b = keys[l];
                    if ((a > b) == asc) { keys[i] = b; keys[l] = a; }
                }
                __syncthreads();
            }
            for (int c = 0; c < P; c += SORT_CHUNK) {
                for (int i = threadIdx.x; i < SORT_CHUNK; i += SORT_THREADS) s[i] = keys[c + i];
                __syncthreads();
                sort_chunk_smem<SORT_CHUNK / 1024>(s, (c & k) != 0, true);  // the chunk is bitonic: one merge, direction of stage k
                for (int i = threadIdx.x; i < SORT_CHUNK; i += SORT_THREADS) keys[c + i] = s[i];
                __syncthreads();
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SORT_THREADS, 1) seg_sort_kernel(SegNms p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s = reinterpret_cast<u64 *>(smem_raw);
    __shared__ float s_red[4][32];
    __shared__ int s_wsum[33];
    const int S = p.hdr->S;
    const bool bad = (p.hdr->status & (BG_STATUS_MASK_SPACE | BG_STATUS_GROUP_RANGE)) != 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    for (int seg = blockIdx.x; seg < S; seg += gridDim.x) {
        const int K = p.seg_count[seg];
        if (K <= 0 || bad) continue;
        const long long off = p.seg_off[seg];
        u64 *keys = p.keys + off;
        block_sort_u64(keys, K, s);
        // gather boxes into score order; area rounded exactly like the CPU kernel: (x2-x1)*(y2-y1)
        const long long bbase = (long long)seg * p.box_seg_stride;
        float mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
        for (int i = threadIdx.x; i < K; i += SORT_THREADS) {
            const float4 b = p.boxes[bbase + key_id(keys[i])];
            p.sorted_box[off + i] = b;
            p.sorted_area[off + i] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
            float w, h, cx, cy;
            if (p.sparse && seg_box_valid(b, w, h, cx, cy)) { mnx = fminf(mnx, cx); mxx = fmaxf(mxx, cx); mny = fminf(mny, cy); mxy = fmaxf(mxy, cy); }
        }
        if (!p.sparse) { __syncthreads(); continue; }

        // ---- grid mode: uniform grid over the valid centres, counting sort into cell order ----
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        }
        if (lane == 0) { s_red[0][wid] = mnx; s_red[1][wid] = mxx; s_red[2][wid] = mny; s_red[3][wid] = mxy; }
        __syncthreads();
        mnx = s_red[0][lane]; mxx = s_red[1][lane]; mny = s_red[2][lane]; mxy = s_red[3][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        }
        SegGrid gr;
        gr.G = seg_grid_dim(K);
        gr.mnx = mnx; gr.mny = mny;
        gr.invx = (mxx > mnx) ? (float)gr.G / (mxx - mnx) : 0.0f;
        gr.invy = (mxy > mny) ? (float)gr.G / (mxy - mny) : 0.0f;
        gr.pad = 1e-6f * fmaxf(fmaxf(fabsf(mnx), fabsf(mxx)), fmaxf(fabsf(mny), fabsf(mxy)));  // fp32 rounding of the centres
        gr.nvalid = 0; gr.pad2 = 0;
        const int G = gr.G, ncell = G * G;
        int *cs = p.cell_start + p.cell_off[seg];
        u64 *tmp = p.emit_key + off;  // (cell << 32 | rank inside the cell); free until the emission
        for (int c = threadIdx.x; c <= ncell; c += SORT_THREADS) cs[c] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < K; i += SORT_THREADS) {
            float w, h, cx, cy;
            u64 t = ~0ull;
            if (seg_box_valid(p.sorted_box[off + i], w, h, cx, cy)) {
                const u32 cell = (u32)(seg_cell_y(gr, cy) * G + seg_cell_x(gr, cx));
                t = ((u64)cell << 32) | (u32)atomicAdd(&cs[cell], 1);
            }
            tmp[i] = t;
        }
        __syncthreads();
        int carry = 0;  // exclusive scan of the cell counts, 4 consecutive cells per thread and pass
        for (int base = 0; base < ncell; base += 4 * SORT_THREADS) {
            const int c0 = base + threadIdx.x * 4;
            int v[4], sum = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) { v[q] = (c0 + q < ncell) ? cs[c0 + q] : 0; sum += v[q]; }
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
            if (lane == 31) s_wsum[wid] = inc;
            __syncthreads();
            if (wid == 0) {
                const int x = s_wsum[lane];
                int xi = x;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += u; }
                s_wsum[lane] = xi - x;
                if (lane == 31) s_wsum[32] = xi;
            }
            __syncthreads();
            int ex = carry + s_wsum[wid] + inc - sum;
#pragma unroll
            for (int q = 0; q < 4; ++q) { if (c0 + q < ncell) cs[c0 + q] = ex; ex += v[q]; }
            carry += s_wsum[32];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            cs[ncell] = carry;
            gr.nvalid = carry;
            p.grid[seg] = gr;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < K; i += SORT_THREADS) {
            const u64 t = tmp[i];
            if (t == ~0ull) continue;
            const u32 cell = (u32)(t >> 32);
            const long long q = off + cs[cell] + (int)(u32)t;
            const float4 bx = p.sorted_box[off + i];
            p.bkeys[q] = ((u64)cell << 32) | (u32)i;
            p.bbox[q] = bx;
            p.barea[q] = p.sorted_area[off + i];
            // extents rounded toward zero to bf16: stored <= true < stored * (1 + 2^-7)
            p.bwh[q] = (__float_as_uint(__fsub_rn(bx.z, bx.x)) >> 16) | (__float_as_uint(__fsub_rn(bx.w, bx.y)) & 0xffff0000u);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// kernel P (grid mode): spatially pruned pair tests.  Work item = 128 consecutive boxes of one segment in cell
// order.  A pair in different cells is tested by the box whose cell comes first in row-major order, a pair inside
// one cell by the box that is earlier in cell order (each box of an overlapping pair lies in the other's reach).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PAIR_THREADS, PAIR_CTAS_PER_SM) nms_grid_pairs_kernel(SegNms p)
{
    __shared__ int s_item;
    if (p.hdr->status & (BG_STATUS_MASK_SPACE | BG_STATUS_GROUP_RANGE)) return;
    const int S = p.hdr->S;
    const int total = p.item_prefix[S];
    const IouThr thr = p.thr;
    while (true) {
        __syncthreads();
        // (once some edge list is full the dense fallback redoes everything: stop early)
        if (threadIdx.x == 0) s_item = ((volatile SegHdr *)p.hdr)->overflow ? total : (int)atomicAdd(&p.hdr->item_ctr, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= total) break;
        int lo = 0, hi = S;  // largest seg with item_prefix[seg] <= item
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (p.item_prefix[mid] <= item) lo = mid; else hi = mid;
        }
        const int seg = lo;
        const SegGrid gr = p.grid[seg];
        const int q = (item - p.item_prefix[seg]) * PAIR_THREADS + threadIdx.x;
        const bool live = q < gr.nvalid;      // (idle lanes stay in the loop: the warp shares out long cell runs below)
        const long long off = p.seg_off[seg];
        const u64 *bk = p.bkeys + off;
        const u32 *wh = p.bwh + off;
        const int *cs = p.cell_start + p.cell_off[seg];
        const int G = gr.G;
        const u64 key = live ? bk[q] : 0ull;
        const u32 cell = (u32)(key >> 32), pos = (u32)key;
        const float4 a = live ? p.bbox[off + q] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float aa = live ? p.barea[off + q] : 0.f;
        const float w = __fsub_rn(a.z, a.x), h = __fsub_rn(a.w, a.y);
        const float cx = 0.5f * a.x + 0.5f * a.z, cy = 0.5f * a.y + 0.5f * a.w;
        const float rx = p.reach * w + gr.pad, ry = p.reach * h + gr.pad;
        const int ay = (int)cell / G, ax = (int)cell - ay * G;
        const int y1 = live ? seg_cell_y(gr, cy + ry) : -1, xl = seg_cell_x(gr, cx - rx), x1 = seg_cell_x(gr, cx + rx);
        u64 *edges = p.mask + p.edge_off[seg];
        const unsigned long long ecap = (unsigned long long)(p.edge_off[seg + 1] - p.edge_off[seg]);
        // IoU > t needs both extent ratios above t: conservative reject on the bf16 extents (1 % slack for the fp32
        // rounding of the exact test, 2^-7 for the truncation) before the partner's box is touched
        auto test = [&](const float4 ba, float baa, float bw, float bh, u32 bpos, int q2) {
            const float wlo = 0.99f * thr.tdn * bw, hlo = 0.99f * thr.tdn * bh, tsc = 0.99f * thr.tdn;
            const u32 v = wh[q2];
            const float wt = __uint_as_float(v << 16), ht = __uint_as_float(v & 0xffff0000u);
            if (wt * 1.008f < wlo || ht * 1.008f < hlo || tsc * wt > bw || tsc * ht > bh) return;
            if (!iou_suppresses(ba, baa, p.bbox[off + q2], p.barea[off + q2], thr)) return;
            const u32 pos2 = (u32)bk[q2];
            const u64 ed = bpos < pos2 ? (((u64)bpos << 32) | pos2) : (((u64)pos2 << 32) | bpos);
            // one atomic per group of lanes that arrive here together (every lane of the warp appends to the same
            // segment's counter: crowded segments would otherwise serialise millions of atomics on one address)
            const u32 am = __activemask();
            const int leader = __ffs(am) - 1;
            unsigned long long base = 0;
            if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(&p.edge_count[seg], (unsigned long long)__popc(am));
            base = __shfl_sync(am, base, leader);
            const unsigned long long e = base + (unsigned long long)__popc(am & lanemask_lt());
            if (e < ecap) edges[e] = ed;
            else if (e == ecap) atomicOr(&p.hdr->overflow, 1);
        };
        // Row by row.  The runs of cell-order entries a box has to look at differ by orders of magnitude (a large box
        // reaches hundreds of cells of a row, a small one a handful), and the 32 boxes of a warp are neighbours in cell
        // order, not in size: long runs are shared out over the lanes of the warp, the others are scanned by their own lane.
        const int lane = threadIdx.x & 31;
        for (int gy = ay; __any_sync(0xffffffffu, gy <= y1); ++gy) {
            const bool mine = gy <= y1;
            int qa = 0, qb = 0;
            if (mine) {
                const bool own_row = gy == ay;
                qa = cs[gy * G + (own_row ? ax : xl)];
                qb = cs[gy * G + x1 + 1];
                if (own_row) qa = max(qa, q + 1);  // own cell: only the boxes after this one in cell order
            }
            // shared out: the runs of the lowest length class (24 / 96 / 384 entries and more) that at most eight lanes
            // reach -- when most lanes have long runs they are busy in parallel anyway, and one lane per run is cheaper
            const int len = qb - qa;
            u32 lm = __ballot_sync(0xffffffffu, len >= PAIR_LONG);
            if (__popc(lm) > 8) lm = __ballot_sync(0xffffffffu, len >= 4 * PAIR_LONG);
            if (__popc(lm) > 8) lm = __ballot_sync(0xffffffffu, len >= 16 * PAIR_LONG);
            if (__popc(lm) > 8) lm = 0u;
            const bool is_long = (lm >> lane) & 1u;
            while (lm) {
                const int src = __ffs(lm) - 1;
                lm &= lm - 1;
                float4 ba;
                ba.x = __shfl_sync(0xffffffffu, a.x, src); ba.y = __shfl_sync(0xffffffffu, a.y, src);
                ba.z = __shfl_sync(0xffffffffu, a.z, src); ba.w = __shfl_sync(0xffffffffu, a.w, src);
                const float baa = __shfl_sync(0xffffffffu, aa, src);
                const u32 bpos = __shfl_sync(0xffffffffu, pos, src);
                const int ba_q = __shfl_sync(0xffffffffu, qa, src), bb_q = __shfl_sync(0xffffffffu, qb, src);
                const float bw = __fsub_rn(ba.z, ba.x), bh = __fsub_rn(ba.w, ba.y);
                for (int q2 = ba_q + lane; q2 < bb_q; q2 += 32) test(ba, baa, bw, bh, bpos, q2);
            }
            if (!is_long)
                for (int q2 = qa; q2 < qb; ++q2) test(a, aa, w, h, pos, q2);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// kernel M: dense suppression bit matrix.  Persistent CTAs pull (segment, row-tile) items from an
// atomic counter; a CTA is 4 groups of 64 threads, each group sweeps every 4th column tile.
// ------------------------------------------------------------------------------------------------
constexpr int MASK_THREADS = 256;

__global__ void __launch_bounds__(MASK_THREADS) nms_mask_kernel(SegNms p)
{
    __shared__ float4 cbox[4][64];
    __shared__ float carea[4][64];
    __shared__ int s_item;
    if (p.hdr->status & (BG_STATUS_MASK_SPACE | BG_STATUS_GROUP_RANGE)) return;
    if (p.sparse) {  // launched behind the grid pair kernel: only needed when an edge list overflowed
        if (!p.hdr->overflow) return;
        if (!p.hdr->dense_fits) return;  // (nms_reduce_kernel reports BG_STATUS_MASK_SPACE)
    }
    unsigned *ctr = p.sparse ? &p.hdr->item_ctr2 : &p.hdr->item_ctr;
    const int S = p.hdr->S;
    const int total = p.tile_prefix[S];
    const int sub = threadIdx.x >> 6, rl = threadIdx.x & 63;
    const IouThr thr = p.thr;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = (int)atomicAdd(ctr, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= total) break;
        int lo = 0, hi = S;  // largest seg with tile_prefix[seg] <= item
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (p.tile_prefix[mid] <= item) lo = mid; else hi = mid;
        }
        const int seg = lo;
        const int r = item - p.tile_prefix[seg];
        const int K = p.seg_count[seg];
        const int T = (K + 63) >> 6;
        const long long off = p.seg_off[seg];
        u64 *mseg = p.mask + p.mask_off[seg];
        const int row = r * 64 + rl;
        const bool valid = row < K;
        float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);
        float ra = 0.f;
        if (valid) { rb = p.sorted_box[off + row]; ra = p.sorted_area[off + row]; }
        for (int cg = r; cg < T; cg += 4) {
            const int ct = cg + sub;
            __syncthreads();
            if (ct < T) {
                const int col = ct * 64 + rl;
                if (col < K) { cbox[sub][rl] = p.sorted_box[off + col]; carea[sub][rl] = p.sorted_area[off + col]; }
            }
            __syncthreads();
            if (ct < T && valid) {
                const int ncol = min(64, K - ct * 64);
                u64 bits = 0;
                for (int j = (ct == r) ? rl + 1 : 0; j < ncol; ++j)
                    if (iou_suppresses(rb, ra, cbox[sub][j], carea[sub][j], thr)) bits |= 1ull << j;
                mseg[(long long)ct * K + row] = bits;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// kernel R: ordered reduce of the bit matrix, one CTA per segment.  Warp 0 resolves one 64-row chunk
// per iteration (serial only over bit operations, operands prefetched into shared memory); warps
// 1..7 pre-reduce the next chunk's column of the matrix over the rows already known to be kept.
// Emission (class filter, ranks, compact lists) is done in parallel after the loop.
// ------------------------------------------------------------------------------------------------
constexpr int REDUCE_THREADS = 256;
constexpr int REDUCE_SMEM_ROWS = 4096;  // rows whose diagonal / next-column words are prefetched (2 x 32 KB)

__device__ __forceinline__ bool class_tracked(const SegNms &p, int c)
{
    for (int i = 0; i < p.n_tracked; ++i)
        if (p.tracked[i] == c) return true;
    return false;
}

// emission of one segment from its kept bitmap `kb` (score order): class filter, ranks, compact lists
template <int NT>
__device__ __forceinline__ void segnms_emit_segment(const SegNms &p, int seg, int K, long long off, const u64 *kb,
                                                    long long *s_scan /*[NT/32]*/, int32_t *out_counts, int counts_per_seg)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int G = (K + 31) >> 5;  // 32-bit words
    u32 *ew = p.ew32 + 2 * (long long)p.tile_prefix[seg];
    u32 *rk = p.rank32 + 2 * (long long)p.tile_prefix[seg];
    const long long cbase = (long long)seg * p.box_seg_stride;
    long long carry = 0;
    for (int base = 0; base < G; base += NT) {
        const int g = base + threadIdx.x;
        u32 bits = 0;
        if (g < G) {
            bits = (u32)(kb[g >> 1] >> ((g & 1) * 32));
            if (p.n_tracked > 0) {
                u32 rest = bits;
                while (rest) {
                    const int bit = __ffs(rest) - 1;
                    rest &= rest - 1;
                    if (!class_tracked(p, p.cls[cbase + key_id(p.keys[off + g * 32 + bit])])) bits &= ~(1u << bit);
                }
            }
            ew[g] = bits;
        }
        long long inc = __popc(bits);
        const long long v = inc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_scan[wid] = inc;
        __syncthreads();
        long long woff = 0, tot = 0;
        for (int q = 0; q < NT / 32; ++q) {
            const long long x = s_scan[q];
            if (q < wid) woff += x;
            tot += x;
        }
        if (g < G) rk[g] = (u32)(carry + woff + inc - v);
        carry += tot;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < K; i += NT) {
        const u32 bits = ew[i >> 5];
        if ((bits >> (i & 31)) & 1u) {
            const u32 r = rk[i >> 5] + __popc(bits & ((1u << (i & 31)) - 1u));
            p.emit_pos[off + r] = (u32)i;
            p.emit_key[off + r] = p.keys[off + i];
        }
    }
    if (threadIdx.x == 0) {
        p.emit_count[seg] = (int)carry;
        if (counts_per_seg) out_counts[2 + seg] = (int)carry;
    }
    __syncthreads();
}

// last CTA of the kernel: exclusive prefix of the emitted rows -> output offsets, total, status
template <int NT>
__device__ __forceinline__ void segnms_finish(const SegNms &p, int S, long long *s_scan, unsigned *s_ticket, int32_t *out_counts)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __threadfence();
    if (threadIdx.x == 0) *s_ticket = atomicAdd(&p.hdr->reduce_done, 1u);
    __syncthreads();
    if (*s_ticket != gridDim.x - 1) return;
    __threadfence();
    long long carry = 0;
    for (int base = 0; base < S; base += NT) {
        const int sg = base + threadIdx.x;
        const long long v = (sg < S) ? (long long)((volatile int *)p.emit_count)[sg] : 0;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_scan[wid] = inc;
        __syncthreads();
        long long woff = 0, tot = 0;
        for (int q = 0; q < NT / 32; ++q) {
            const long long x = s_scan[q];
            if (q < wid) woff += x;
            tot += x;
        }
        if (sg < S) p.out_prefix[sg] = carry + woff + inc - v;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.out_prefix[S] = carry;
        p.hdr->total_out = carry;
        out_counts[0] = (int)carry;
        out_counts[1] = p.hdr->status;
    }
}

__global__ void __launch_bounds__(REDUCE_THREADS) nms_reduce_kernel(SegNms p, int32_t *out_counts, int counts_per_seg)
{
    extern __shared__ __align__(16) unsigned char red_smem[];
    u64 *s_diag_all = reinterpret_cast<u64 *>(red_smem);       // [REDUCE_SMEM_ROWS] word of row i in its own column tile
    u64 *s_fix_all = s_diag_all + REDUCE_SMEM_ROWS;            // [REDUCE_SMEM_ROWS] word of row i in the next column tile
    __shared__ u64 s_diag[64];
    __shared__ u64 s_part[2][8];
    __shared__ long long s_scan[REDUCE_THREADS / 32];
    __shared__ unsigned s_ticket;
    const int S = p.hdr->S;
    if (p.sparse && !p.hdr->overflow) return;  // grid mode went through: nms_resolve_kernel has done everything
    if (p.sparse && !p.hdr->dense_fits && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&p.hdr->status, BG_STATUS_MASK_SPACE);
    const bool bad = (p.hdr->status & (BG_STATUS_MASK_SPACE | BG_STATUS_GROUP_RANGE)) != 0 || (p.sparse && !p.hdr->dense_fits);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    for (int seg = blockIdx.x; seg < S; seg += gridDim.x) {
        const int K = bad ? 0 : p.seg_count[seg];
        const int T = (K + 63) >> 6;
        const long long off = p.seg_off[seg];
        const u64 *mseg = p.mask + p.mask_off[seg];
        u64 *kb = p.keepbits + p.tile_prefix[seg];
        const bool pre = K <= REDUCE_SMEM_ROWS;
        if (pre) {
            for (int i = threadIdx.x; i < K; i += REDUCE_THREADS) {
                const int c = i >> 6;
                s_diag_all[i] = mseg[(long long)c * K + i];
                s_fix_all[i] = (c + 1 < T) ? mseg[(long long)(c + 1) * K + i] : 0ull;
            }
        }
        if (threadIdx.x < 16) s_part[threadIdx.x >> 3][threadIdx.x & 7] = 0;
        __syncthreads();
        u64 kept_prev = 0;  // warp 0 only
        for (int c = 0; c < T; ++c) {
            if (wid == 0) {
                u64 rem = 0;
                if (c >= 1) {
                    if (lane < 8) rem = s_part[c & 1][lane];  // rows < 64(c-1), reduced during the previous iteration
                    const int r0 = 64 * (c - 1) + lane, r1 = r0 + 32;  // rows of chunk c-1 (all < K since c-1 < T-1)
                    if (pre) {
                        if ((kept_prev >> lane) & 1ull) rem |= s_fix_all[r0];
                        if ((kept_prev >> (lane + 32)) & 1ull) rem |= s_fix_all[r1];
                    } else {
                        const u64 *mc = mseg + (long long)c * K;
                        if ((kept_prev >> lane) & 1ull) rem |= mc[r0];
                        if ((kept_prev >> (lane + 32)) & 1ull) rem |= mc[r1];
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) rem |= __shfl_xor_sync(0xffffffffu, rem, o);
                }
                const int ncol = min(64, K - 64 * c);
                const u64 *sd;
                if (pre) {
                    sd = s_diag_all + 64 * c;  // entries >= ncol are never consulted (those rows are pre-removed)
                } else {
                    const u64 *md = mseg + (long long)c * K + 64 * c;
                    s_diag[lane] = (lane < ncol) ? md[lane] : 0ull;
                    s_diag[lane + 32] = (lane + 32 < ncol) ? md[lane + 32] : 0ull;
                    __syncwarp();
                    sd = s_diag;
                }
                if (ncol < 64) rem |= ~((1ull << ncol) - 1ull);
                u32 rlo = (u32)rem, rhi = (u32)(rem >> 32), klo = 0, khi = 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (!((rlo >> i) & 1u)) {
                        const u64 d = sd[i];
                        klo |= 1u << i;
                        rlo |= (u32)d;
                        rhi |= (u32)(d >> 32);
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (!((rhi >> i) & 1u)) {
                        khi |= 1u << i;
                        rhi |= (u32)(sd[32 + i] >> 32);
                    }
                }
                const u64 kept = ((u64)khi << 32) | klo;
                if (lane == 0) kb[c] = kept;
                kept_prev = kept;
            } else if (c + 1 < T && c >= 1) {
                // column c+1 of the matrix, OR-ed over the kept rows of chunks < c (final since iteration c-1)
                const u64 *mc = mseg + (long long)(c + 1) * K;
                u64 acc = 0;
                for (int i = threadIdx.x - 32; i < 64 * c; i += REDUCE_THREADS - 32)
                    if ((kb[i >> 6] >> (i & 63)) & 1ull) acc |= mc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc |= __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) s_part[(c + 1) & 1][wid] = acc;
            }
            __syncthreads();
        }

        segnms_emit_segment<REDUCE_THREADS>(p, seg, K, off, kb, s_scan, out_counts, counts_per_seg);
    }
    segnms_finish<REDUCE_THREADS>(p, S, s_scan, &s_ticket, out_counts);
}

// ------------------------------------------------------------------------------------------------
// kernel G (grid mode): greedy outcome from the edge list, one CTA per segment.  state: 0 undecided, 1 kept,
// 2 suppressed.  Per round every edge (i -> j, i earlier) of an undecided j either suppresses j (i kept) or blocks
// it (i undecided); undecided boxes that were not blocked are kept.  The earliest undecided box is never blocked,
// so every round decides at least one box, and a box is decided only from decided predecessors: the result is
// the sequential greedy scan's.  Then the same emission as the dense path.
// ------------------------------------------------------------------------------------------------
constexpr int RESOLVE_THREADS = 1024;
constexpr int RESOLVE_SMEM_BOXES = 32768;  // segments up to this size keep their state in shared memory (2 x 32 KB)

__global__ void __launch_bounds__(RESOLVE_THREADS) nms_resolve_kernel(SegNms p, int32_t *out_counts, int counts_per_seg)
{
    extern __shared__ __align__(16) unsigned char res_smem[];
    __shared__ long long s_scan[RESOLVE_THREADS / 32];
    __shared__ unsigned s_ticket;
    if (p.hdr->overflow) return;  // the dense fallback (nms_mask_kernel + nms_reduce_kernel) takes over
    const int S = p.hdr->S;
    const bool bad = (p.hdr->status & (BG_STATUS_MASK_SPACE | BG_STATUS_GROUP_RANGE)) != 0;
    const int lane = threadIdx.x & 31;

    for (int seg = blockIdx.x; seg < S; seg += gridDim.x) {
        const int K = bad ? 0 : p.seg_count[seg];
        const long long off = p.seg_off[seg];
        u64 *kb = p.keepbits + p.tile_prefix[seg];
        unsigned char *state = (K <= RESOLVE_SMEM_BOXES) ? res_smem : p.gstate + 2 * off;
        unsigned char *blocked = state + ((K <= RESOLVE_SMEM_BOXES) ? RESOLVE_SMEM_BOXES : K);
        const u64 *edges = p.mask + p.edge_off[seg];
        const long long ne = bad ? 0 : (long long)p.edge_count[seg];
        for (int i = threadIdx.x; i < K; i += RESOLVE_THREADS) { state[i] = 0; blocked[i] = 0; }
        __syncthreads();
        // Thread t owns the edges t, t + NT, t + 2 NT, ... and keeps only the live ones, packed to the front of its own
        // positions: an edge i -> j is dead for good once j is decided or i is suppressed, so later rounds sweep a fraction
        // of the list (crowded segments need dozens of rounds).  Eight edges per thread are in flight: the list lives in L2.
        u64 *ew = const_cast<u64 *>(edges);
        long long mine = ne > (long long)threadIdx.x ? (ne - threadIdx.x + RESOLVE_THREADS - 1) / RESOLVE_THREADS : 0;
        while (ne > 0) {
            long long w = 0;
            for (long long k0 = 0; k0 < mine; k0 += 8) {
                u64 ed[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    ed[u] = k0 + u < mine ? __ldcg(edges + threadIdx.x + (k0 + u) * RESOLVE_THREADS) : ~0ull;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (ed[u] == ~0ull) continue;
                    const u32 i = (u32)(ed[u] >> 32), j = (u32)ed[u];
                    if (state[j] == 0) {
                        const unsigned char si = state[i];
                        if (si == 1) state[j] = 2;
                        else if (si == 0) {
                            blocked[j] = 1;
                            ew[threadIdx.x + w * RESOLVE_THREADS] = ed[u];   // (w <= k0 + u: never ahead of the reads)
                            ++w;
                        }
                    }
                }
            }
            mine = w;
            __syncthreads();
            int pending = 0;
            for (int j = threadIdx.x; j < K; j += RESOLVE_THREADS) {
                if (state[j] == 0) {
                    if (blocked[j]) { blocked[j] = 0; pending = 1; }
                    else state[j] = 1;
                }
            }
            if (!__syncthreads_or(pending)) break;
        }
        // kept bitmap in score order (boxes without edges are still in state 0: kept)
        const int T = (K + 63) >> 6;
        for (int c = threadIdx.x >> 5; c < T; c += RESOLVE_THREADS / 32) {  // one 64-box word per warp and pass
            const int i0 = 64 * c + lane, i1 = i0 + 32;
            const u32 lo = __ballot_sync(0xffffffffu, i0 < K && state[i0] != 2);
            const u32 hi = __ballot_sync(0xffffffffu, i1 < K && state[i1] != 2);
            if (lane == 0) kb[c] = ((u64)hi << 32) | lo;
        }
        __syncthreads();
        segnms_emit_segment<RESOLVE_THREADS>(p, seg, K, off, kb, s_scan, out_counts, counts_per_seg);
    }
    segnms_finish<RESOLVE_THREADS>(p, S, s_scan, &s_ticket, out_counts);
}

// rank of an emitted row in the global (score desc, flat index asc) order across all segments:
// own rank plus, for every other segment, the number of its emitted rows that precede the key.
__device__ __forceinline__ long long segnms_global_rank(const SegNms &p, int S, int seg, long long r, u64 key)
{
    const u32 hi = (u32)(key >> 32);
    const long long flat = (long long)seg * p.box_seg_stride + key_id(key);
    long long rank = r;
    for (int s2 = 0; s2 < S; ++s2) {
        if (s2 == seg) continue;
        const int cnt = p.emit_count[s2];
        if (cnt == 0) continue;
        const u64 *ek = p.emit_key + p.seg_off[s2];
        const long long fb = (long long)s2 * p.box_seg_stride;
        int lo = 0, hi_i = cnt;  // first element that does NOT precede
        while (lo < hi_i) {
            const int mid = (lo + hi_i) >> 1;
            const u64 k2 = ek[mid];
            const u32 h2 = (u32)(k2 >> 32);
            const bool before = (h2 < hi) || (h2 == hi && fb + key_id(k2) < flat);
            if (before) lo = mid + 1; else hi_i = mid;
        }
        rank += lo;
    }
    return rank;
}

static inline size_t segnms_sort_smem() { return (size_t)SORT_CHUNK * sizeof(u64); }
static inline size_t segnms_reduce_smem() { return (size_t)2 * REDUCE_SMEM_ROWS * sizeof(u64); }
static inline size_t segnms_resolve_smem() { return (size_t)2 * RESOLVE_SMEM_BOXES; }

// grid mode needs a positive threshold: the reach bound is (1-t)/t extents, useless for tiny t
static inline void segnms_configure(SegNms &p)
{
    p.sparse = (p.thr.fast_ok && !p.thr.zero_suppresses && p.thr.tdn >= 0.05f && p.thr.tdn < 1.0f) ? 1 : 0;
    p.reach = p.sparse ? (1.0f - p.thr.tdn) / p.thr.tdn * 1.01f + 0.01f : 0.0f;
}

// Enqueue tables -> sort -> [grid pairs] -> mask -> [resolve] -> reduce.  In grid mode the mask and reduce kernels
// return at once unless an edge list overflowed.  `S_launch` is a host-side upper bound of the segment count.
static int segnms_run(const SegNms &p, int S_launch, int32_t *out_counts, int counts_per_seg, int num_sms,
                      cudaStream_t st)
{
    static bool attr_set_dev[64] = {false};  // function attributes are per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { (void)cudaGetLastError(); dev = 0; }
    bool &attr_set = attr_set_dev[dev];
    if (!attr_set) {
        cudaFuncSetAttribute(seg_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)segnms_sort_smem());
        cudaFuncSetAttribute(nms_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)segnms_reduce_smem());
        cudaFuncSetAttribute(nms_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)segnms_resolve_smem());
        attr_set = true;
    }
    seg_tables_kernel<<<1, 1024, 0, st>>>(p);
    BG_LAUNCH_CHECK();
    int g = S_launch < 1 ? 1 : S_launch;
    if (g > num_sms * 2) g = num_sms * 2;
    seg_sort_kernel<<<g, SORT_THREADS, segnms_sort_smem(), st>>>(p);
    BG_LAUNCH_CHECK();
    if (p.sparse) {
        nms_grid_pairs_kernel<<<num_sms * PAIR_CTAS_PER_SM, PAIR_THREADS, 0, st>>>(p);
        BG_LAUNCH_CHECK();
    }
    nms_mask_kernel<<<num_sms * 8, MASK_THREADS, 0, st>>>(p);
    BG_LAUNCH_CHECK();
    int gr = S_launch < 1 ? 1 : S_launch;
    if (gr > num_sms * 3) gr = num_sms * 3;
    if (p.sparse) {
        nms_resolve_kernel<<<gr, RESOLVE_THREADS, segnms_resolve_smem(), st>>>(p, out_counts, counts_per_seg);
        BG_LAUNCH_CHECK();
    }
    nms_reduce_kernel<<<gr, REDUCE_THREADS, segnms_reduce_smem(), st>>>(p, out_counts, counts_per_seg);
    BG_LAUNCH_CHECK();
    return BG_OK;
}

// ------------------------------------------------------------------------------------------------
// generic batched_nms glue: group ids -> segments
// ------------------------------------------------------------------------------------------------
__global__ void gnms_init_kernel(SegNms p, long long max_groups)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        SegHdr h;
        h.S = 0; h.status = 0; h.item_ctr = 0; h.reduce_done = 0;
        h.gmin = 0x7fffffffffffffffLL; h.gmax = -0x7fffffffffffffffLL - 1; h.total_out = 0;
        h.item_ctr2 = 0; h.overflow = 0; h.dense_fits = 0; h.pad0 = 0; h.pad[0] = 0;
        *p.hdr = h;
    }
    for (long long k = i; k < max_groups; k += (long long)gridDim.x * blockDim.x) p.seg_count[k] = 0;
}

__global__ void gnms_minmax_kernel(SegNms p, const long long *idxs, long long n)
{
    long long mn = 0x7fffffffffffffffLL, mx = -0x7fffffffffffffffLL - 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long v = idxs[i];
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&p.hdr->gmin, mn);
        atomicMax(&p.hdr->gmax, mx);
    }
}

__global__ void gnms_count_kernel(SegNms p, const long long *idxs, long long n, long long max_groups, u32 *slot)
{
    const long long gmin = p.hdr->gmin, gmax = p.hdr->gmax;
    const bool bad = (unsigned long long)(gmax - gmin) >= (unsigned long long)max_groups;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (bad) { atomicOr(&p.hdr->status, BG_STATUS_GROUP_RANGE); p.hdr->S = 0; }
        else p.hdr->S = (int)(gmax - gmin + 1);
    }
    if (bad) return;
    // one atomic per distinct group and warp (group ids usually come in long runs)
    const int lane = threadIdx.x & 31;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i - lane < n; i += (long long)gridDim.x * blockDim.x) {
        const bool active = i < n;
        const long long g = active ? idxs[i] - gmin : -1 - lane;
        const u32 peers = __match_any_sync(0xffffffffu, g);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (active && lane == leader) base = atomicAdd(&p.seg_count[g], __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (active) slot[i] = (u32)(base + __popc(peers & ((1u << lane) - 1u)));
    }
}

__global__ void __launch_bounds__(1024) gnms_offsets_kernel(SegNms p)
{
    __shared__ ScanPair s_scan[33];
    const int S = p.hdr->S;
    long long carry = 0;
    for (int base = 0; base < S; base += 1024) {
        const int sg = base + threadIdx.x;
        const int K = (sg < S) ? p.seg_count[sg] : 0;
        const long long room = K > 0 ? (long long)next_pow2((u32)K) : 0;
        ScanPair tot;
        const ScanPair ex = block_excl_scan_1024(ScanPair{room, 0}, s_scan, tot);
        if (sg < S) p.seg_off[sg] = carry + ex.a;
        carry += tot.a;
    }
    if (threadIdx.x == 0) p.seg_off[S] = carry;
}

__global__ void gnms_scatter_kernel(SegNms p, const long long *idxs, const float *scores, long long n, const u32 *slot)
{
    if (p.hdr->status & BG_STATUS_GROUP_RANGE) return;
    const long long gmin = p.hdr->gmin;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p.keys[p.seg_off[idxs[i] - gmin] + slot[i]] = make_key(scores[i], (u32)i);
}

// Global (score desc, index asc) order of the kept candidates: the per-segment lists are sorted already, so a merge
// tree over adjacent runs finishes in ceil(log2 S) passes.  Level L merges the runs of 2^L segments pairwise; an
// element's place is its index in its own run plus the number of smaller keys in the sibling run (keys are unique:
// they carry the candidate index).  Lanes of a warp that sit in the same run share the search: the first and last
// of them search the whole sibling run, the others only between those two results.
__global__ void gnms_gather_kernel(SegNms p, u64 *dst)
{
    const int S = p.hdr->S;
    for (int seg = blockIdx.y; seg < S; seg += gridDim.y) {
        const int cnt = p.emit_count[seg];
        const u64 *src = p.emit_key + p.seg_off[seg];
        u64 *d = dst + p.out_prefix[seg];
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < cnt; r += gridDim.x * blockDim.x) d[r] = src[r];
    }
}

__device__ __forceinline__ long long key_lower_bound(const u64 *a, long long lo, long long hi, u64 key)
{
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) gnms_merge_level_kernel(SegNms p, const u64 *src, u64 *dst, int L)
{
    const int S = p.hdr->S;
    if ((1ll << L) >= S) return;  // a single run already
    const long long total = p.out_prefix[S];
    const int lane = threadIdx.x & 31;
    auto bnd = [&](long long run) -> long long { const long long sg = run << L; return p.out_prefix[sg < S ? sg : S]; };
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g - lane < total; g += (long long)gridDim.x * blockDim.x) {
        const bool active = g < total;
        long long run = -1 - lane;
        u64 key = 0;
        if (active) {
            int lo = 0, hi = S;  // largest segment with out_prefix[seg] <= g (the non-empty one that holds g)
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (p.out_prefix[mid] <= g) lo = mid; else hi = mid;
            }
            run = lo >> L;
            key = src[g];
        }
        const u32 peers = __match_any_sync(0xffffffffu, run);
        const int first = __ffs(peers) - 1, last = 31 - __clz(peers);
        long long a0 = 0, b0 = 0, b1 = 0, m0 = 0, lb = 0;
        const bool edge = lane == first || lane == last;
        if (active) {
            a0 = bnd(run); b0 = bnd(run ^ 1); b1 = bnd((run ^ 1) + 1); m0 = bnd(run & ~1ll);
            if (edge) lb = key_lower_bound(src, b0, b1, key);
        }
        const long long lbf = __shfl_sync(0xffffffffu, lb, first), lbl = __shfl_sync(0xffffffffu, lb, last);
        if (active) {
            if (!edge) lb = key_lower_bound(src, lbf, lbl, key);
            dst[m0 + (g - a0) + (lb - b0)] = key;
        }
    }
}

__global__ void gnms_output_kernel(SegNms p, const u64 *ping, const u64 *pong, long long *out_keep)
{
    const int S = p.hdr->S;
    int levels = 0;
    while ((1ll << levels) < S) ++levels;
    const u64 *src = (levels & 1) ? pong : ping;
    const long long total = p.out_prefix[S];
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x)
        out_keep[r] = (long long)key_id(src[r]);
}

}  // namespace bg
