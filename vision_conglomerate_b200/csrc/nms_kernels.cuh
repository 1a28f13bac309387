// Segmented greedy NMS: per-segment bitonic sort -> suppression bit matrix -> ordered reduce.
//
// Semantics follow torchvision's CPU nms kernel (the arithmetic behind the reference's
// torchvision.ops.batched_nms call, inference_det.py:77-82): candidates in stable descending score
// order, box j is suppressed by a kept box i iff  (double)(inter / (area_i + area_j - inter)) > thr
// with every operation rounded to fp32 individually (no FMA contraction).
//
// The bit matrix is filled by one of two kernels:
//  * nms_pairs_kernel (thr > 0): IoU > t requires width AND height ratios above t, so boxes are
//    bucketed by (log2 w, log2 h) in bins one threshold-ratio wide and every box is only tested
//    against the 3x3 neighbouring bins (three contiguous ranges of the bin-sorted order).  Bits are
//    OR-ed into a zeroed matrix.  For trained-like detections this visits ~15 % of the pairs.
//  * nms_mask_kernel (any threshold): dense 64x64 tiles, every word written exactly once.
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

// size bin of a box: (log2 h bin) * nb + (log2 w bin); degenerate boxes (w <= 0, h <= 0, NaN) can neither
// suppress nor be suppressed under a non-negative threshold and get the sentinel bin.
__device__ __forceinline__ u32 size_bin(const SegNms &p, const float4 b)
{
    const float w = __fsub_rn(b.z, b.x), h = __fsub_rn(b.w, b.y);
    if (!(w > 0.0f && h > 0.0f)) return 0xffffffffu;
    const float top = (float)(p.nb - 1);
    const int bw = (int)fminf(fmaxf(floorf((log2f(w) + 16.0f) * p.inv_delta), 0.0f), top);
    const int bh = (int)fminf(fmaxf(floorf((log2f(h) + 16.0f) * p.inv_delta), 0.0f), top);
    return (u32)(bh * p.nb + bw);
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

__global__ void __launch_bounds__(1024) seg_tables_kernel(SegNms p)
{
    __shared__ ScanPair s_scan[33];
    const int S = p.hdr->S;
    ScanPair carry{0, 0};
    long long icarry = 0;
    for (int base = 0; base < S; base += 1024) {
        const int sg = base + threadIdx.x;
        const long long K = (sg < S) ? p.seg_count[sg] : 0;
        const long long T = (K + 63) >> 6;
        ScanPair tot, tot2;
        const ScanPair ex = block_excl_scan_1024(ScanPair{T, K * T}, s_scan, tot);
        const ScanPair ex2 = block_excl_scan_1024(ScanPair{(K + PAIR_THREADS - 1) / PAIR_THREADS, 0}, s_scan, tot2);
        if (sg < S) {
            p.tile_prefix[sg] = (int)(carry.a + ex.a);
            p.mask_off[sg] = carry.b + ex.b;
            p.item_prefix[sg] = (int)(icarry + ex2.a);
        }
        carry.a += tot.a;
        carry.b += tot.b;
        icarry += tot2.a;
    }
    if (threadIdx.x == 0) {
        p.tile_prefix[S] = (int)carry.a;
        p.mask_off[S] = carry.b;
        p.item_prefix[S] = (int)icarry;
        if (carry.b > p.mask_words) atomicOr(&p.hdr->status, BG_STATUS_MASK_SPACE);
    }
}

// ------------------------------------------------------------------------------------------------
// kernel S: per-segment sorts (score order, then bin order) and matrix zeroing
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

// ascending sort of keys[0..K) (global, with room for next_pow2(K) entries) by the whole CTA
__device__ void block_sort_u64(u64 *keys, int K, u64 *s)
{
    const int P = (int)next_pow2((u32)K);
    if (P <= SORT_CHUNK) {
        for (int i = threadIdx.x; i < P; i += SORT_THREADS) s[i] = (i < K) ? keys[i] : ~0ull;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                bitonic_step_smem(s, P, 0, k, j);
                __syncthreads();
            }
        for (int i = threadIdx.x; i < K; i += SORT_THREADS) keys[i] = s[i];
    } else {
        for (int i = K + threadIdx.x; i < P; i += SORT_THREADS) keys[i] = ~0ull;
        __syncthreads();
        for (int c = 0; c < P; c += SORT_CHUNK) {  // phase 1: every stage k <= SORT_CHUNK, chunk by chunk
            for (int i = threadIdx.x; i < SORT_CHUNK; i += SORT_THREADS) s[i] = keys[c + i];
            __syncthreads();
            for (int k = 2; k <= SORT_CHUNK; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    bitonic_step_smem(s, SORT_CHUNK, c, k, j);
                    __syncthreads();
                }
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
                for (int j = SORT_CHUNK >> 1; j > 0; j >>= 1) {
                    bitonic_step_smem(s, SORT_CHUNK, c, k, j);
                    __syncthreads();
                }
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
    const int S = p.hdr->S;
    const bool bad = (p.hdr->status & (BG_STATUS_MASK_SPACE | BG_STATUS_GROUP_RANGE)) != 0;

    for (int seg = blockIdx.x; seg < S; seg += gridDim.x) {
        const int K = p.seg_count[seg];
        if (K <= 0 || bad) continue;
        const long long off = p.seg_off[seg];
        u64 *keys = p.keys + off;
        block_sort_u64(keys, K, s);
        // gather boxes into score order; area rounded exactly like the CPU kernel: (x2-x1)*(y2-y1)
        const long long bbase = (long long)seg * p.box_seg_stride;
        for (int i = threadIdx.x; i < K; i += SORT_THREADS) {
            const float4 b = p.boxes[bbase + key_id(keys[i])];
            p.sorted_box[off + i] = b;
            p.sorted_area[off + i] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
            if (p.sparse) p.bkeys[off + i] = ((u64)size_bin(p, b) << 32) | (u32)i;
        }
        __syncthreads();
        if (p.sparse) {
            u64 *bk = p.bkeys + off;
            block_sort_u64(bk, K, s);
            for (int q = threadIdx.x; q < K; q += SORT_THREADS) {
                const u32 pos = (u32)bk[q];
                p.bbox[off + q] = p.sorted_box[off + pos];
                p.barea[off + q] = p.sorted_area[off + pos];
            }
            // the pair kernel ORs bits into the matrix: clear this segment's words
            const long long words = (long long)K * ((K + 63) >> 6);
            ulonglong2 *m2 = reinterpret_cast<ulonglong2 *>(p.mask + p.mask_off[seg]);
            if ((p.mask_off[seg] & 1) == 0) {
                for (long long w = threadIdx.x; w < (words >> 1); w += SORT_THREADS) m2[w] = make_ulonglong2(0ull, 0ull);
                if ((words & 1) && threadIdx.x == 0) p.mask[p.mask_off[seg] + words - 1] = 0ull;
            } else {
                u64 *m = p.mask + p.mask_off[seg];
                for (long long w = threadIdx.x; w < words; w += SORT_THREADS) m[w] = 0ull;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// kernel P: bin-pruned pair tests.  Work item = 128 consecutive boxes of one segment in bin order.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int bin_lower_bound(const u64 *bk, int K, u32 bin)  // first q with bin(q) >= bin
{
    int lo = 0, hi = K;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((u32)(bk[mid] >> 32) < bin) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(PAIR_THREADS) nms_pairs_kernel(SegNms p)
{
    __shared__ int s_item;
    if (p.hdr->status & (BG_STATUS_MASK_SPACE | BG_STATUS_GROUP_RANGE)) return;
    const int S = p.hdr->S;
    const int total = p.item_prefix[S];
    const IouThr thr = p.thr;
    const int nb = p.nb;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = (int)atomicAdd(&p.hdr->item_ctr, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= total) break;
        int lo = 0, hi = S;  // largest seg with item_prefix[seg] <= item
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (p.item_prefix[mid] <= item) lo = mid; else hi = mid;
        }
        const int seg = lo;
        const int K = p.seg_count[seg];
        const long long off = p.seg_off[seg];
        const u64 *bk = p.bkeys + off;
        const int q = (item - p.item_prefix[seg]) * PAIR_THREADS + threadIdx.x;
        if (q >= K) continue;
        const u64 key = bk[q];
        const u32 bin = (u32)(key >> 32), pos = (u32)key;
        if (bin == 0xffffffffu) continue;
        const float4 rb = p.bbox[off + q];
        const float ra = p.barea[off + q];
        u64 *mrow = p.mask + p.mask_off[seg] + pos;
        const int bh = (int)bin / nb, bw = (int)bin % nb;
        for (int dh = -1; dh <= 1; ++dh) {
            const int bh2 = bh + dh;
            if (bh2 < 0 || bh2 >= nb) continue;
            const u32 klo = (u32)(bh2 * nb + max(bw - 1, 0)), khi = (u32)(bh2 * nb + min(bw + 1, nb - 1));
            const int a = bin_lower_bound(bk, K, klo), b = bin_lower_bound(bk, K, khi + 1);
            for (int q2 = a; q2 < b; ++q2) {
                const u32 pos2 = (u32)bk[q2];
                if (pos2 <= pos) continue;  // the earlier box of a pair owns the test
                if (iou_suppresses(rb, ra, p.bbox[off + q2], p.barea[off + q2], thr))
                    atomicOr(mrow + (long long)(pos2 >> 6) * K, 1ull << (pos2 & 63));
            }
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
    const int S = p.hdr->S;
    const int total = p.tile_prefix[S];
    const int sub = threadIdx.x >> 6, rl = threadIdx.x & 63;
    const IouThr thr = p.thr;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = (int)atomicAdd(&p.hdr->item_ctr, 1u);
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
    const bool bad = (p.hdr->status & (BG_STATUS_MASK_SPACE | BG_STATUS_GROUP_RANGE)) != 0;
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

        // ---- emission: rows that are kept and (optionally) of a tracked class, compacted in score order ----
        const int G = (K + 31) >> 5;  // 32-bit words
        u32 *ew = p.ew32 + 2 * (long long)p.tile_prefix[seg];
        u32 *rk = p.rank32 + 2 * (long long)p.tile_prefix[seg];
        const long long cbase = (long long)seg * p.box_seg_stride;
        long long carry = 0;
        for (int base = 0; base < G; base += REDUCE_THREADS) {
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
            for (int q = 0; q < REDUCE_THREADS / 32; ++q) {
                const long long x = s_scan[q];
                if (q < wid) woff += x;
                tot += x;
            }
            if (g < G) rk[g] = (u32)(carry + woff + inc - v);
            carry += tot;
            __syncthreads();
        }
        for (int i = threadIdx.x; i < K; i += REDUCE_THREADS) {
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

    // last CTA: exclusive prefix of the emitted rows -> output offsets, total, status
    __threadfence();
    if (threadIdx.x == 0) s_ticket = atomicAdd(&p.hdr->reduce_done, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    long long carry = 0;
    for (int base = 0; base < S; base += REDUCE_THREADS) {
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
        for (int q = 0; q < REDUCE_THREADS / 32; ++q) {
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

// bin geometry for the pair kernel: bins one threshold-ratio wide (plus a safety margin for the fp32
// log2 and the few-ulp slack of the rounded IoU), at most 64 per axis over log2 size in [-16, 16)
static inline void segnms_configure(SegNms &p)
{
    p.sparse = (p.thr.fast_ok && !p.thr.zero_suppresses && p.thr.tdn < 1.0f) ? 1 : 0;
    p.nb = 1;
    p.inv_delta = 0.0f;
    if (p.sparse) {
        double delta = log2(1.0 / (double)p.thr.tdn) * 1.0005 + 2e-4;
        if (delta < 32.0 / 63.0) delta = 32.0 / 63.0;
        int nb = (int)ceil(32.0 / delta) + 1;
        if (nb > 64) nb = 64;
        p.nb = nb;
        p.inv_delta = (float)(1.0 / delta);
    }
}

// Enqueue tables -> sort -> pairs|mask -> reduce.  `S_launch` is a host-side upper bound of the segment count.
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
        attr_set = true;
    }
    seg_tables_kernel<<<1, 1024, 0, st>>>(p);
    BG_LAUNCH_CHECK();
    int g = S_launch < 1 ? 1 : S_launch;
    if (g > num_sms * 2) g = num_sms * 2;
    seg_sort_kernel<<<g, SORT_THREADS, segnms_sort_smem(), st>>>(p);
    BG_LAUNCH_CHECK();
    if (p.sparse) nms_pairs_kernel<<<num_sms * 8, PAIR_THREADS, 0, st>>>(p);
    else nms_mask_kernel<<<num_sms * 8, MASK_THREADS, 0, st>>>(p);
    BG_LAUNCH_CHECK();
    int gr = S_launch < 1 ? 1 : S_launch;
    if (gr > num_sms * 3) gr = num_sms * 3;
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
        h.pad[0] = h.pad[1] = h.pad[2] = 0;
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
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        slot[i] = (u32)atomicAdd(&p.seg_count[idxs[i] - gmin], 1);
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

__global__ void gnms_output_kernel(SegNms p, long long *out_keep)
{
    const int S = p.hdr->S;
    for (int seg = blockIdx.y; seg < S; seg += gridDim.y) {
        const int cnt = p.emit_count[seg];
        const long long off = p.seg_off[seg];
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < cnt; r += gridDim.x * blockDim.x) {
            const u64 key = p.emit_key[off + r];
            out_keep[segnms_global_rank(p, S, seg, r, key)] = (long long)key_id(key);
        }
    }
}

}  // namespace bg
