// Segmented greedy NMS: per-segment bitonic sort -> tiled suppression bit matrix -> ordered reduce.
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

// ------------------------------------------------------------------------------------------------
// block-wide exclusive scan of (int, long long) pairs, 1024 threads
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
// kernel S: per-segment sort (+ prefix tables computed by CTA 0)
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

__global__ void __launch_bounds__(SORT_THREADS, 1) seg_sort_kernel(SegNms p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s = reinterpret_cast<u64 *>(smem_raw);
    __shared__ ScanPair s_scan[33];
    const int S = p.hdr->S;

    if (blockIdx.x == 0) {  // prefix tables for the mask / reduce kernels
        ScanPair carry{0, 0};
        for (int base = 0; base < S; base += SORT_THREADS) {
            const int sg = base + threadIdx.x;
            long long K = (sg < S) ? p.seg_count[sg] : 0;
            long long T = (K + 63) >> 6;
            ScanPair tot;
            ScanPair ex = block_excl_scan_1024(ScanPair{T, K * T}, s_scan, tot);
            if (sg < S) {
                p.tile_prefix[sg] = (int)(carry.a + ex.a);
                p.mask_off[sg] = carry.b + ex.b;
            }
            carry.a += tot.a;
            carry.b += tot.b;
        }
        if (threadIdx.x == 0) {
            p.tile_prefix[S] = (int)carry.a;
            p.mask_off[S] = carry.b;
            if (carry.b > p.mask_words) atomicOr(&p.hdr->status, BG_STATUS_MASK_SPACE);
        }
    }

    for (int seg = blockIdx.x; seg < S; seg += gridDim.x) {
        const int K = p.seg_count[seg];
        if (K <= 0) continue;
        const long long off = p.seg_off[seg];
        u64 *keys = p.keys + off;
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
        // gather boxes into sorted order; area rounded exactly like the CPU kernel: (x2-x1)*(y2-y1)
        const long long bbase = (long long)seg * p.box_seg_stride;
        for (int i = threadIdx.x; i < K; i += SORT_THREADS) {
            const float4 b = p.boxes[bbase + key_id(keys[i])];
            p.sorted_box[off + i] = b;
            p.sorted_area[off + i] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// kernel M: suppression bit matrix.  Persistent CTAs pull (segment, row-tile) items from an atomic
// counter; a CTA is 4 groups of 64 threads, each group sweeps every 4th column tile of the row tile.
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
// per iteration (serial only over bit operations); warps 1..7 pre-reduce the next chunk's column of
// the matrix over the rows already known to be kept.
// ------------------------------------------------------------------------------------------------
constexpr int REDUCE_THREADS = 256;

__device__ __forceinline__ bool class_tracked(const SegNms &p, int c)
{
    for (int i = 0; i < p.n_tracked; ++i)
        if (p.tracked[i] == c) return true;
    return false;
}

__global__ void __launch_bounds__(REDUCE_THREADS) nms_reduce_kernel(SegNms p, int32_t *out_counts, int counts_per_seg)
{
    __shared__ u64 s_diag[64];
    __shared__ u64 s_part[2][8];
    __shared__ ScanPair s_scan[33];
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
        const long long cbase = (long long)seg * p.box_seg_stride;
        int emitted = 0;       // warp 0 only
        u64 kept_prev = 0;     // warp 0 only
        if (threadIdx.x < 16) s_part[threadIdx.x >> 3][threadIdx.x & 7] = 0;
        __syncthreads();
        for (int c = 0; c < T; ++c) {
            if (wid == 0) {
                u64 rem = 0;
                if (c >= 1) {
                    if (lane < 8) rem = s_part[c & 1][lane];  // rows < 64(c-1), reduced during the previous iteration
                    const int r0 = 64 * (c - 1) + lane, r1 = r0 + 32;  // rows of chunk c-1 (all < K since c-1 < T-1)
                    const u64 *mc = mseg + (long long)c * K;
                    if ((kept_prev >> lane) & 1ull) rem |= mc[r0];
                    if ((kept_prev >> (lane + 32)) & 1ull) rem |= mc[r1];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) rem |= __shfl_xor_sync(0xffffffffu, rem, o);
                }
                const int ncol = min(64, K - 64 * c);
                const u64 *md = mseg + (long long)c * K + 64 * c;
                s_diag[lane] = (lane < ncol) ? md[lane] : 0ull;
                s_diag[lane + 32] = (lane + 32 < ncol) ? md[lane + 32] : 0ull;
                __syncwarp();
                if (ncol < 64) rem |= ~((1ull << ncol) - 1ull);
                u64 kept = 0;
#pragma unroll 16
                for (int i = 0; i < 64; ++i) {
                    if (!((rem >> i) & 1ull)) { kept |= 1ull << i; rem |= s_diag[i]; }
                }
                if (lane == 0) kb[c] = kept;
                // rows to emit: kept, and (optionally) of a tracked class
                const int p0 = 64 * c + lane, p1 = p0 + 32;
                u64 k0 = 0, k1 = 0;
                bool ok0 = (kept >> lane) & 1ull, ok1 = (kept >> (lane + 32)) & 1ull;
                if (ok0) k0 = p.keys[off + p0];
                if (ok1) k1 = p.keys[off + p1];
                if (p.n_tracked > 0) {
                    if (ok0) ok0 = class_tracked(p, p.cls[cbase + key_id(k0)]);
                    if (ok1) ok1 = class_tracked(p, p.cls[cbase + key_id(k1)]);
                }
                const u32 e0 = __ballot_sync(0xffffffffu, ok0), e1 = __ballot_sync(0xffffffffu, ok1);
                if (ok0) {
                    const int r = emitted + __popc(e0 & lanemask_lt());
                    p.emit_pos[off + r] = (u32)p0;
                    p.emit_key[off + r] = k0;
                }
                if (ok1) {
                    const int r = emitted + __popc(e0) + __popc(e1 & lanemask_lt());
                    p.emit_pos[off + r] = (u32)p1;
                    p.emit_key[off + r] = k1;
                }
                emitted += __popc(e0) + __popc(e1);
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
        if (threadIdx.x == 0) {
            p.emit_count[seg] = emitted;
            if (counts_per_seg) out_counts[2 + seg] = emitted;
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
        // 256-thread exclusive scan via the pair scan helper (b unused)
        const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, inc, o);
            if (l >= o) inc += u;
        }
        if (l == 31) s_scan[w].a = inc;
        __syncthreads();
        long long woff = 0, tot = 0;
        for (int q = 0; q < REDUCE_THREADS / 32; ++q) {
            const long long x = s_scan[q].a;
            if (q < w) woff += x;
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

// Enqueue sort -> mask -> reduce.  `S_launch` is a host-side upper bound of the segment count.
static int segnms_run(const SegNms &p, int S_launch, int32_t *out_counts, int counts_per_seg, int num_sms,
                      cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(seg_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)segnms_sort_smem());
        attr_set = true;
    }
    int g = S_launch < 1 ? 1 : S_launch;
    if (g > num_sms * 2) g = num_sms * 2;
    seg_sort_kernel<<<g, SORT_THREADS, segnms_sort_smem(), st>>>(p);
    BG_LAUNCH_CHECK();
    nms_mask_kernel<<<num_sms * 8, MASK_THREADS, 0, st>>>(p);
    BG_LAUNCH_CHECK();
    int gr = S_launch < 1 ? 1 : S_launch;
    if (gr > num_sms * 8) gr = num_sms * 8;
    nms_reduce_kernel<<<gr, REDUCE_THREADS, 0, st>>>(p, out_counts, counts_per_seg);
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
    for (int seg = blockIdx.x; seg < S; seg += gridDim.x) {
        const int cnt = p.emit_count[seg];
        const long long off = p.seg_off[seg];
        for (int r = threadIdx.x; r < cnt; r += blockDim.x) {
            const u64 key = p.emit_key[off + r];
            out_keep[segnms_global_rank(p, S, seg, r, key)] = (long long)key_id(key);
        }
    }
}

}  // namespace bg
