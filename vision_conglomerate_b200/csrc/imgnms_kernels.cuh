// Per-image greedy NMS (the NMS stage of the fused detect path).
//
// A CTA of 1024 threads owns one image and keeps its working set in shared memory; while the GPU has room every
// image gets a second, helper CTA that takes a share of the pair tests (step 3) and the sort (step 5) and hands
// its results over through global memory.  Survivors are numbered p = 0..K-1 in candidate order (the order of
// the decode kernel's tile slots):
//   1. load keys, boxes and classes from the tile slots (coalesced; prefix over the tile counts);
//   2. bucket the box centres on a uniform grid (counting sort), bounds from a block reduction;
//   3. pair tests.  IoU > t needs |dcx| < (1-t)/t * w and |dcy| < (1-t)/t * h for EITHER box of the pair
//      (DESIGN.md has the derivation), so a box only meets the boxes of the grid cells inside that reach.
//      The work is cut into (box, grid row) items spread evenly over the threads -- a tall box has hundreds
//      of cells in reach, a small one a handful.  A pair is tested once (by the box whose cell comes first; inside
//      a cell by the lower-numbered box), after a cheap reject on bf16 extents; the exact fp32 IoU decision is
//      iou_suppresses() of the segmented engine.  A hit becomes an edge from the box
//      that comes first in score order to the other one;
//   4. greedy resolution by rounds over the edge list: a box is suppressed once a kept earlier neighbour is
//      known, kept once all its earlier neighbours are known to be suppressed -- the fixed point is exactly
//      the sequential greedy scan;
//   5. bitonic sort of the keys (score desc, candidate index asc -- torchvision's stable order);
//   6. class filter, ranks, output offset by decoupled look-back over the images, rows.
// Overlap edges beyond the shared-memory list spill to a per-image list in global memory (L2 resident).  Images
// with more survivors than the kernel variant holds (4,096 / 8,192), or more edges than the spill list, are left to
// the general segmented engine: the kernel raises BG_STATUS_NEED_GENERAL and the caller re-enqueues accordingly.
#pragma once
#include "detect_kernels.cuh"

namespace bg {

constexpr int INMS_THREADS = 1024;
constexpr int INMS_MAXT = 1024;      // tiles per image
constexpr int INMS_GMAX = 64;        // grid cells per axis
constexpr int INMS_TINY_K = 128;     // images with at most this many survivors: all pairs, rank by counting, no helper (video frames)
constexpr int INMS_HELPER_SHARE_32 = 7;  // the helper takes 7/32 of the pair-test items (it also does the sort)

// Three sizes of the kernel.  Small: up to 4 096 survivors per image, boxes in shared memory (the latency case).
// Large: up to 8 192 survivors; the boxes no longer fit next to the keys, so they live in a compact global
// array (L2-resident) and only the bf16 extents used by the pre-filter stay in shared memory.
// Lean (throughput mode, several batches in flight): up to 2 048 survivors, 512 threads, <= 56 registers and 46 KB of
// shared memory, so that a CTA fits on an SM NEXT TO the two resident CTAs of the next batch's decode kernel
// (2 x 90.7 KB, 2 x 128 threads x 144 registers) instead of waiting for -- or taking -- a whole SM.  Boxes, classes and
// (until the sort) the keys live in the image's L2-resident scratch; the keys' shared-memory home is shared with the
// arrays of the pair-test stage.
template <int CAP_, bool BOX_SMEM_, int THREADS_ = 1024, bool LEAN_ = false>
struct InmsCfg {
    static constexpr int CAP = CAP_;                 // survivors per image
    static constexpr bool BOX_SMEM = BOX_SMEM_;
    static constexpr int THREADS = THREADS_;
    static constexpr bool LEAN = LEAN_;
    static constexpr int PBITS = CAP_ == 2048 ? 11 : (CAP_ == 4096 ? 12 : 13);   // bits of p inside the sort key
    static constexpr int ECAP = LEAN_ ? 4096 : (BOX_SMEM_ ? 12288 : 8192);  // overlap edges in shared memory (more spill to global)
    static constexpr int ITEMS = 2 * CAP_;           // (box, grid row) work items per pass
    static constexpr int PER = CAP_ / THREADS_;      // boxes / sort keys per thread
    static constexpr int MAX_N = 1 << (32 - PBITS);  // candidates per image (key = score | idx | p)
    static constexpr int GMAX = LEAN_ ? 32 : INMS_GMAX;  // grid cells per axis (2 (G+1)^2 <= K bounds G by 31 at 2 048)
    static constexpr int MAXT = THREADS_;            // tiles per image (one tile count per thread in stage 1)
    // split mode: the helper CTA takes this many 32nds of the pair-test items (it also does the sort; with 512 threads
    // the pair tests are 3.5x the sort, with 1024 threads 1.7x)
    static constexpr int HELPER_SHARE_32 = LEAN_ ? 11 : INMS_HELPER_SHARE_32;
};
typedef InmsCfg<4096, true> InmsSmall;
typedef InmsCfg<8192, false> InmsLarge;
typedef InmsCfg<2048, false, 512, true> InmsLean;
constexpr int INMS_CAP_MAX = InmsLarge::CAP;
constexpr int INMS_HCAP = 12288;     // edges a helper CTA hands over from its shared-memory list (>= any ECAP)

struct ImgNmsK {
    int B, N, TR, tpi_total;
    int tpi[3], img_off[3];
    const int *tile_count;
    const u64 *keys;           // tile slots
    const float4 *box_slots;
    const int *cls_slots;
    IouThr thr;
    float reach;               // (1-t)/t plus margin
    int n_tracked;
    int tracked[BG_MAX_TRACKED];
    FusedHdr *hdr;
    u64 *chain;
    int order;                 // 0: rows written here, image-major; 1: emit lists only (rows by detect_output_kernel)
    u64 *emit_key;             // [B*N] order 1: kept keys of image b at b*N in score order, with their boxes / classes
    float4 *emit_box;
    int *emit_cls;
    int *emit_count;           // [B]
    int *cand_count;           // [B]
    float *out_boxes;
    long long *out_img, *out_keep;
    int32_t *out_counts;
    int32_t *host_flag;   // optional: word in mapped host memory, written last (bg_detect_params.host_flag)
    int host_flag_value;
    unsigned long long *stamps;  // optional [B, INMS_STAMPS] globaltimer (ns) at the stage boundaries (profiling hook)
    // split mode (two CTAs per image when the GPU has room): the helper tests a share of the pairs and sorts the keys
    int split;
    u32 *gflag;                // [B,4]: helper -> main edge count + 1, sorted keys ready, spilled-edge counter, pad
    u32 *gedges;               // [B, INMS_HCAP] the helper's shared-memory edge list
    u32 *gspill;               // [B, gcap] edges that did not fit a CTA's shared-memory list (main and helper)
    int gcap;
    u64 *gsorted;              // [B, INMS_CAP_MAX]
    float4 *gboxp;             // [B, INMS_CAP_MAX] large / lean variant: boxes in survivor (p) order
    u64 *gkeyp;                // [B, InmsLean::CAP] lean variant: keys in survivor order (until the sort)
    u32 *gclsp;                // [B, InmsLean::CAP] lean variant: classes in survivor order
};
constexpr int INMS_STAMPS = 10;

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define INMS_STAMP(i) do { if (k.stamps && tid == 0 && role == 1) k.stamps[(long long)b * INMS_STAMPS + (i)] = globaltimer_ns(); } while (0)

template <class Cfg, bool LEAN = Cfg::LEAN>
struct ImgNmsSmem {
    u64 keys[Cfg::CAP];                          // (~score | idx | p); indexed by p until the sort
    float4 box[Cfg::BOX_SMEM ? Cfg::CAP : 1];    // by p (small variant only)
    u32 edges[Cfg::ECAP];                        // (from << 16) | to, in p numbers
    int cell_start[INMS_GMAX * INMS_GMAX + 1];
    union {
        int tile_pref[INMS_MAXT + 1];            // stage 1
        unsigned short cellord[Cfg::CAP];        // stage 2..3: box numbers in cell order
    };
    union {
        struct { unsigned short cell_of[Cfg::CAP], rank_in_cell[Cfg::CAP]; };  // stage 2
        unsigned short item_owner[Cfg::ITEMS];                                 // stage 3
    };
    union {
        unsigned char item_row[Cfg::ITEMS];                                    // stage 3
        struct { unsigned char state[Cfg::CAP], blocked[Cfg::CAP]; };          // stage 4..6: 0 undecided, 1 kept, 2 suppressed
    };
    unsigned short cls[Cfg::CAP];                // by p
    u32 whc[Cfg::CAP];                           // stage 3: (w, h) of the boxes in cell order, truncated to bf16 pairs
    int wsum[33];
    float red[4][32];
    int n_edges;
    int next_item;                               // stage 3: warps pull 32 work items at a time
    int img;
    long long base;
};
// lean layout: the keys are parked in global memory from stage 1 to the sort and share their home with the
// stage 2..3 arrays; classes stay in global memory
template <class Cfg>
struct ImgNmsSmem<Cfg, true> {
    union {
        u64 keys[Cfg::CAP];                      // stage 5..6
        struct {
            u32 whc[Cfg::CAP];                   // stage 2..3
            union {
                struct { unsigned short cell_of[Cfg::CAP], rank_in_cell[Cfg::CAP]; };  // stage 2
                unsigned short item_owner[Cfg::ITEMS];                                 // stage 3
            };
        };
    };
    u32 edges[Cfg::ECAP];
    int cell_start[Cfg::GMAX * Cfg::GMAX + 1];
    union {
        int tile_pref[Cfg::MAXT + 1];            // stage 1
        unsigned short cellord[Cfg::CAP];        // stage 2..3
    };
    union {
        unsigned char item_row[Cfg::ITEMS];                                    // stage 3
        struct { unsigned char state[Cfg::CAP], blocked[Cfg::CAP]; };          // stage 4..6
    };
    int wsum[33];
    float red[4][32];
    int n_edges;
    int next_item;
    int img;
    long long base;
};
static_assert(sizeof(ImgNmsSmem<InmsSmall>) <= 227 * 1024 && sizeof(ImgNmsSmem<InmsLarge>) <= 227 * 1024,
              "per-image NMS state must fit one SM's shared memory");
// 228 KB per SM - 2 x (87 040 B ring + 3 712 B static and reserved) of the decode CTAs - 1 KB reserved for this CTA
static_assert(sizeof(ImgNmsSmem<InmsLean>) <= 233472 - 2 * (2 * 43520 + 3712) - 1024,
              "the lean per-image NMS must fit next to two decode CTAs");

template <int NT>
__device__ __forceinline__ int inms_block_excl_scan(int v, int *wsum /*[33]*/, int &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const int w = lane < NT / 32 ? wsum[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += u;
        }
        wsum[lane] = winc - w;
        if (lane == 31) wsum[32] = winc;
    }
    __syncthreads();
    const int r = wsum[wid] + inc - v;
    total = wsum[32];
    __syncthreads();
    return r;
}

// pads keys[K..P) with ~0 and sorts keys[0..P), P = the smallest of 1, 2, 4 (, 8) x THREADS >= K
template <class Cfg>
__device__ __forceinline__ void inms_sort_keys(u64 *keys, int K)
{
    constexpr int NT = Cfg::THREADS;
    int P = NT;
    while (P < K) P <<= 1;
    for (int j = K + threadIdx.x; j < P; j += NT) keys[j] = ~0ull;
    __syncthreads();
    if (P == NT) sort_reg_1024<1, NT>(keys);
    else if (P == 2 * NT) sort_reg_1024<2, NT>(keys);
    else if (P == 4 * NT || Cfg::CAP <= 4 * NT) sort_reg_1024<4, NT>(keys);
    else sort_reg_1024<8, NT>(keys);
}

__device__ __forceinline__ bool inms_tracked(const ImgNmsK &k, int c)
{
    for (int i = 0; i < k.n_tracked; ++i)
        if (k.tracked[i] == c) return true;
    return false;
}

struct InmsGrid {
    float mnx, mny, invx, invy, pad, reach;
    int G;
    __device__ __forceinline__ int cx(float x) const { return (int)fminf(fmaxf(floorf(__fmul_rn(__fsub_rn(x, mnx), invx)), 0.0f), (float)(G - 1)); }
    __device__ __forceinline__ int cy(float y) const { return (int)fminf(fmaxf(floorf(__fmul_rn(__fsub_rn(y, mny), invy)), 0.0f), (float)(G - 1)); }
};

__device__ __forceinline__ bool inms_box_valid(const float4 bx, float &w, float &h, float &cx, float &cy)
{
    w = __fsub_rn(bx.z, bx.x); h = __fsub_rn(bx.w, bx.y);
    cx = 0.5f * bx.x + 0.5f * bx.z; cy = 0.5f * bx.y + 0.5f * bx.w;
    // boxes without a positive finite extent have IoU 0 or NaN with everything: they never suppress nor get suppressed
    return (w > 0.0f) && (h > 0.0f) && (w < INFINITY) && (h < INFINITY) && (fabsf(cx) < INFINITY) && (fabsf(cy) < INFINITY);
}

template <class Cfg>
// (lean: bounds of 2 x 576 threads cap the kernel at 56 registers -- 512 x 56 is what two decode CTAs leave of an SM's file)
__global__ void __launch_bounds__(Cfg::LEAN ? 576 : Cfg::THREADS, Cfg::LEAN ? 2 : 1) image_nms_kernel(ImgNmsK k)
{
    extern __shared__ __align__(16) unsigned char inms_raw[];
    ImgNmsSmem<Cfg> &S = *reinterpret_cast<ImgNmsSmem<Cfg> *>(inms_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int CAP = Cfg::CAP, PER = Cfg::PER, INMS_THREADS = Cfg::THREADS, INMS_GMAX = Cfg::GMAX;

    // launched with programmatic stream serialization: the CTA may become resident while the decode kernel
    // drains; everything it reads is produced by that kernel, so wait for it here
    cudaGridDependencySynchronize();
    if (tid == 0) S.img = (int)atomicAdd(&k.hdr->ticket, 1u);  // dynamic order: whoever a CTA waits for holds an earlier ticket
    __syncthreads();
    // split mode: tickets 2b (helper: a share of the pair tests, then the sort) and 2b+1 (main: everything else).
    // The main CTA waits for its helper and for the main CTAs of earlier images only -- all earlier tickets.
    const int b = k.split ? (S.img >> 1) : S.img;
    const int role = k.split ? (S.img & 1) : 1;
    if (b >= k.B) return;
    const long long ibase = (long long)b * k.N;
    // boxes by survivor number p: shared memory (small variant) or the image's compact global array (large variant;
    // both CTAs of an image write identical values there, then read their own writes)
    float4 *gbx = k.gboxp + (long long)b * INMS_CAP_MAX;
    auto box_at = [&](int i) -> float4 { if constexpr (Cfg::BOX_SMEM) return S.box[i]; else return gbx[i]; };
    // lean variant: keys (until the sort) and classes by survivor number in the image's global scratch (both CTAs of
    // an image write identical values there, then read their own writes)
    u64 *gkey = k.gkeyp + (long long)b * InmsLean::CAP;
    u32 *gcls = k.gclsp + (long long)b * InmsLean::CAP;
    auto key_at = [&](int i) -> u64 { if constexpr (Cfg::LEAN) return gkey[i]; else return S.keys[i]; };
    auto cls_at = [&](int i) -> int { if constexpr (Cfg::LEAN) return (int)gcls[i]; else return (int)S.cls[i]; };
    INMS_STAMP(0);

    // ---- 1. tile counts -> prefix; keys, boxes, classes into shared memory (slot order = candidate order) ----
    int K;
    {
        const int c = (tid < k.tpi_total) ? k.tile_count[(long long)b * k.tpi_total + tid] : 0;
        const int ex = inms_block_excl_scan<INMS_THREADS>(c, S.wsum, K);
        if (tid <= k.tpi_total) S.tile_pref[tid] = ex;  // tid == tpi_total holds the total (c = 0 there)
        if (tid == 0) S.n_edges = 0;
    }
    __syncthreads();
    const int K_all = K;
    bool over = K > CAP;
    if (over) K = 0;  // this image is left to the general path; it still takes part in the look-back chain
    float mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
    for (int j = tid; j < K; j += INMS_THREADS) {
        int lo = 0, hi = k.tpi_total;  // largest tile with tile_pref[tile] <= j
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (S.tile_pref[mid] <= j) lo = mid; else hi = mid;
        }
        int si = 0, lr = lo;
        if (lr >= k.tpi[0]) { lr -= k.tpi[0]; si = 1; if (lr >= k.tpi[1]) { lr -= k.tpi[1]; si = 2; } }
        const long long slot = ibase + k.img_off[si] + (long long)lr * k.TR + (j - S.tile_pref[lo]);
        const u64 key = k.keys[slot];
        const float4 bx = k.box_slots[slot];
        const int cl = k.cls_slots[slot];
        const u64 pkey = (key & 0xffffffff00000000ull) | ((u64)key_id(key) << Cfg::PBITS) | (u64)j;
        if constexpr (Cfg::LEAN) { gkey[j] = pkey; gcls[j] = (u32)cl; } else { S.keys[j] = pkey; S.cls[j] = (unsigned short)cl; }
        if constexpr (Cfg::BOX_SMEM) S.box[j] = bx; else gbx[j] = bx;
        float w, h, cx, cy;
        if (inms_box_valid(bx, w, h, cx, cy)) { mnx = fminf(mnx, cx); mxx = fmaxf(mxx, cx); mny = fminf(mny, cy); mxy = fmaxf(mxy, cy); }
    }
    if (!Cfg::BOX_SMEM) __syncthreads();  // the global box (key, class) arrays are complete (block-scope visibility)
    INMS_STAMP(1);
    // A few dozen survivors (batch-1 video frames with a real score threshold): the grid, the work items and the
    // 1024-key sorting network are all fixed cost there.  Test all pairs, rank the keys by counting, skip the helper
    // (both CTAs of an image see the same K, so the main CTA knows not to wait).
    const bool tiny = K <= INMS_TINY_K;
    if (tiny && role == 0) return;
    const IouThr thr = k.thr;
    if (tiny) {
        __syncthreads();  // keys / boxes of stage 1 are in place
        for (int t = tid; t < K * K; t += INMS_THREADS) {
            const int i = t / K, j = t - i * K;
            if (i >= j) continue;
            const float4 a = box_at(i), c = box_at(j);
            float w, h, cx, cy;
            if (!inms_box_valid(a, w, h, cx, cy)) continue;
            const float aa = __fmul_rn(w, h);
            if (!inms_box_valid(c, w, h, cx, cy)) continue;
            if (iou_suppresses(a, aa, c, __fmul_rn(w, h), thr)) {
                const bool i_first = key_at(i) < key_at(j);
                const u32 ed = i_first ? (((u32)i << 16) | (u32)j) : (((u32)j << 16) | (u32)i);
                const int e = atomicAdd(&S.n_edges, 1);
                if (e < Cfg::ECAP) S.edges[e] = ed;
                else {
                    const u32 o = atomicAdd(&k.gflag[4 * b + 2], 1u);
                    if (o < (u32)k.gcap) k.gspill[(long long)b * k.gcap + o] = ed;
                }
            }
        }
        __syncthreads();
        INMS_STAMP(2);
    } else {

    // ---- 2. grid over the valid centres, counting sort by cell --------------------------------------------------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if (lane == 0) { S.red[0][wid] = mnx; S.red[1][wid] = mxx; S.red[2][wid] = mny; S.red[3][wid] = mxy; }
    __syncthreads();  // also: every thread is done with tile_pref (aliases cellord)
    {   // (only THREADS / 32 entries were written)
        const bool w = lane < INMS_THREADS / 32;
        mnx = w ? S.red[0][lane] : INFINITY; mxx = w ? S.red[1][lane] : -INFINITY;
        mny = w ? S.red[2][lane] : INFINITY; mxy = w ? S.red[3][lane] : -INFINITY;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    // about two boxes per cell; the cell of a centre is a monotone function of the coordinate, so a conservative
    // coordinate interval maps to a conservative cell interval
    InmsGrid gr;
    gr.G = 1;
    while (gr.G < INMS_GMAX && 2 * (gr.G + 1) * (gr.G + 1) <= K) ++gr.G;
    gr.mnx = mnx; gr.mny = mny;
    gr.invx = (mxx > mnx) ? (float)gr.G / (mxx - mnx) : 0.0f;
    gr.invy = (mxy > mny) ? (float)gr.G / (mxy - mny) : 0.0f;
    gr.pad = 1e-6f * fmaxf(fmaxf(fabsf(mnx), fabsf(mxx)), fmaxf(fabsf(mny), fabsf(mxy)));  // fp32 rounding of the centres
    gr.reach = k.reach;
    const int G = gr.G, ncell = G * G;
    for (int c = tid; c <= ncell; c += INMS_THREADS) S.cell_start[c] = 0;
    __syncthreads();
    for (int i = tid; i < K; i += INMS_THREADS) {
        float w, h, cx, cy;
        unsigned short cell = 0xffff;
        if (inms_box_valid(box_at(i), w, h, cx, cy)) {
            cell = (unsigned short)(gr.cy(cy) * G + gr.cx(cx));
            S.rank_in_cell[i] = (unsigned short)atomicAdd(&S.cell_start[cell], 1);
        }
        S.cell_of[i] = cell;
    }
    __syncthreads();
    {   // exclusive scan of the cell counts (4 consecutive cells per thread)
        const int c0 = tid * 4;
        int v[4], sum = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) { v[q] = (c0 + q < ncell) ? S.cell_start[c0 + q] : 0; sum += v[q]; }
        int tot;
        int ex = inms_block_excl_scan<INMS_THREADS>(sum, S.wsum, tot);
#pragma unroll
        for (int q = 0; q < 4; ++q) { if (c0 + q < ncell) S.cell_start[c0 + q] = ex; ex += v[q]; }
        if (tid == 0) S.cell_start[ncell] = tot;
    }
    __syncthreads();
    for (int i = tid; i < K; i += INMS_THREADS) {
        const unsigned short cell = S.cell_of[i];
        if (cell != 0xffff) {
            const int pos = S.cell_start[cell] + S.rank_in_cell[i];
            const float4 bx = box_at(i);
            S.cellord[pos] = (unsigned short)i;
            // extents rounded toward zero to bf16: stored <= true < stored * (1 + 2^-7)
            S.whc[pos] = (__float_as_uint(__fsub_rn(bx.z, bx.x)) >> 16) | (__float_as_uint(__fsub_rn(bx.w, bx.y)) & 0xffff0000u);
        }
    }
    // Work items = (box, grid row) for the rows from the box's own row down to the end of its reach: a pair in
    // different cells is tested by the box whose cell comes first in row-major order, a pair inside one cell
    // by the lower-numbered box (each box of a pair lies in the other's reach, so either side finds it).
    // Thread t owns boxes PER*t .. PER*t + PER-1.
    int y0[PER], nrow[PER], items = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int i = tid * PER + q;
        y0[q] = 0; nrow[q] = 0;
        float w, h, cx, cy;
        if (i < K && inms_box_valid(box_at(i), w, h, cx, cy)) {
            const float ry = gr.reach * h + gr.pad;
            y0[q] = gr.cy(cy);
            nrow[q] = gr.cy(cy + ry) - y0[q] + 1;
        }
        items += nrow[q];
    }
    int T;
    const int item0 = inms_block_excl_scan<INMS_THREADS>(items, S.wsum, T);  // (its barriers also publish cellord)
    INMS_STAMP(2);

    // ---- 3. pair tests, item-parallel ------------------------------------------------------------------------
    const int T_helper = k.split ? (int)(((long long)T * Cfg::HELPER_SHARE_32) >> 5) : 0;
    const int it_lo = role == 1 ? T_helper : 0, it_hi = role == 1 ? T : T_helper;
    for (int c0 = it_lo; c0 < it_hi; c0 += Cfg::ITEMS) {
        const int nit = min(Cfg::ITEMS, it_hi - c0);
        {   // publish the items of this pass
            int it = item0;
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                for (int r = 0; r < nrow[q]; ++r, ++it) {
                    const int rel = it - c0;
                    if (rel >= 0 && rel < nit) {
                        S.item_owner[rel] = (unsigned short)(tid * PER + q);
                        S.item_row[rel] = (unsigned char)(y0[q] + r);
                    }
                }
            }
        }
        if (tid == 0) S.next_item = 0;
        __syncthreads();
        // a warp pulls 32 consecutive items at a time: the items differ a lot in length, and a fixed share per warp
        // left the CTA waiting for its unluckiest warp
        for (;;) {
            int ibase = 0;
            if (lane == 0) ibase = atomicAdd(&S.next_item, 32);
            ibase = __shfl_sync(0xffffffffu, ibase, 0);
            if (ibase >= nit) break;
            const int it = ibase + lane;
            if (it >= nit) continue;
            const int i = S.item_owner[it];
            const int gy = S.item_row[it];
            const float4 a = box_at(i);
            const float w = __fsub_rn(a.z, a.x), h = __fsub_rn(a.w, a.y);
            const float aa = __fmul_rn(w, h);
            const float cx = 0.5f * a.x + 0.5f * a.z, cy = 0.5f * a.y + 0.5f * a.w;
            const float rx = gr.reach * w + gr.pad;
            const int ax = gr.cx(cx), ay = gr.cy(cy);
            const bool own_row = gy == ay;
            const int x0 = own_row ? ax : gr.cx(cx - rx), x1 = gr.cx(cx + rx);
            const int qa = S.cell_start[gy * G + x0], qb = S.cell_start[gy * G + x1 + 1];
            const int q_own = own_row ? S.cell_start[gy * G + ax + 1] : qa;  // entries below q_own share my cell
            // IoU > t needs both extent ratios above t.  Conservative reject on the bf16 extents kept in cell
            // order (1 % slack for the fp32 rounding of the exact test, 2^-7 for the truncation), so most
            // entries cost one 4-byte load and never touch the partner's box.
            const float wlo = 0.99f * thr.tdn * w, hlo = 0.99f * thr.tdn * h, tsc = 0.99f * thr.tdn;
            for (int q = qa; q < qb; q += 4) {
                u32 wh[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) wh[u] = (q + u < qb) ? S.whc[q + u] : 0u;  // padding: zero extents, rejected
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float wt = __uint_as_float(wh[u] << 16), ht = __uint_as_float(wh[u] & 0xffff0000u);
                    if (wt * 1.008f < wlo || ht * 1.008f < hlo || tsc * wt > w || tsc * ht > h) continue;
                    const int j = S.cellord[q + u];
                    if (q + u < q_own && j <= i) continue;  // same cell: the lower-numbered box owns the test
                    const float4 c = box_at(j);
                    const float ac = __fmul_rn(__fsub_rn(c.z, c.x), __fsub_rn(c.w, c.y));
                    if (iou_suppresses(a, aa, c, ac, thr)) {
                        const bool i_first = key_at(i) < key_at(j);  // earlier in (score desc, index asc) order
                        const u32 ed = i_first ? (((u32)i << 16) | (u32)j) : (((u32)j << 16) | (u32)i);
                        const int e = atomicAdd(&S.n_edges, 1);
                        if (e < Cfg::ECAP) S.edges[e] = ed;
                        else {  // shared-memory list full: spill to the image's global list
                            const u32 o = atomicAdd(&k.gflag[4 * b + 2], 1u);
                            if (o < (u32)k.gcap) k.gspill[(long long)b * k.gcap + o] = ed;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    }  // !tiny
    int ne = min(S.n_edges, Cfg::ECAP);  // edges in this CTA's shared-memory list (the rest were spilled)
    if (role == 0) {
        // ---- helper: hand the edges over, sort the keys, hand them over, done ----
        u32 *ge = k.gedges + (long long)b * INMS_HCAP;
        for (int e = tid; e < ne; e += INMS_THREADS) ge[e] = S.edges[e];
        __threadfence();
        __syncthreads();
        if (tid == 0) ((volatile u32 *)k.gflag)[4 * b] = (u32)ne + 1u;
        if constexpr (Cfg::LEAN) {   // the keys come home (their place held the pair-test arrays)
            for (int j = tid; j < K; j += INMS_THREADS) S.keys[j] = gkey[j];
        }
        inms_sort_keys<Cfg>(S.keys, K);
        u64 *gs = k.gsorted + (long long)b * INMS_CAP_MAX;
        for (int j = tid; j < K; j += INMS_THREADS) gs[j] = S.keys[j];
        __threadfence();
        __syncthreads();
        if (tid == 0) ((volatile u32 *)k.gflag)[4 * b + 1] = 1u;
        return;
    }
    int nh = 0;  // edges in the helper's list
    if (k.split && !tiny) {
        if (tid == 0) {
            u32 v;
            do { v = ((volatile u32 *)k.gflag)[4 * b]; } while (v == 0u);
            S.img = (int)v - 1;
        }
        __syncthreads();
        nh = S.img;
        __threadfence();
    }
    __syncthreads();  // (also: this CTA's own spills are complete)
    int nsp = (int)min(((volatile u32 *)k.gflag)[4 * b + 2], 0x7fffffffu);  // spilled edges, main's and the helper's
    if (nsp > k.gcap) { over = true; ne = nh = nsp = 0; K = 0; }
    const u32 *ge = k.gedges + (long long)b * INMS_HCAP;
    const u32 *gsp = k.gspill + (long long)b * k.gcap;
    INMS_STAMP(3);

    // ---- 4. greedy resolution by rounds -------------------------------------------------------------------------------
    for (int i = tid; i < CAP; i += INMS_THREADS) { S.state[i] = 0; S.blocked[i] = 0; }  // (aliases item_row)
    __syncthreads();
    auto relax = [&](u32 ed) {
        const int i = (int)(ed >> 16), j = (int)(ed & 0xffffu);
        if (S.state[j] == 0) {
            const unsigned char si = S.state[i];
            if (si == 1) S.state[j] = 2;
            else if (si == 0) S.blocked[j] = 1;
        }
    };
    while (true) {
        for (int e = tid; e < ne; e += INMS_THREADS) relax(S.edges[e]);
        for (int e = tid; e < nh; e += INMS_THREADS) relax(__ldcg(ge + e));
        for (int e = tid; e < nsp; e += INMS_THREADS) relax(__ldcg(gsp + e));
        __syncthreads();
        int pending = 0;
        for (int j = tid; j < K; j += INMS_THREADS) {
            if (S.state[j] == 0) {
                if (S.blocked[j]) { S.blocked[j] = 0; pending = 1; }
                else S.state[j] = 1;
            }
        }
        if (!__syncthreads_or(pending)) break;
    }
    INMS_STAMP(4);

    // ---- 5. sort ------------------------------------------------------------------------------------------------------
    if (tiny) {      // rank by counting: the keys are distinct (they end in the survivor number)
        if constexpr (Cfg::LEAN) {
            for (int j = tid; j < K; j += INMS_THREADS) S.keys[j] = gkey[j];
            __syncthreads();
        }
        u64 mine = 0;
        int r = 0;
        if (tid < K) {
            mine = S.keys[tid];
            for (int j = 0; j < K; ++j) r += S.keys[j] < mine ? 1 : 0;
        }
        __syncthreads();
        if (tid < K) S.keys[r] = mine;
        __syncthreads();
    } else if (k.split) {   // the helper sorted the keys meanwhile
        if (tid == 0) { while (((volatile u32 *)k.gflag)[4 * b + 1] == 0u) { } }
        __syncthreads();
        __threadfence();
        const u64 *gs = k.gsorted + (long long)b * INMS_CAP_MAX;
        for (int j = tid; j < K; j += INMS_THREADS) S.keys[j] = __ldcg(gs + j);
        __syncthreads();
    } else {
        if constexpr (Cfg::LEAN) {   // the keys come home (their place held the pair-test arrays)
            for (int j = tid; j < K; j += INMS_THREADS) S.keys[j] = gkey[j];
        }
        inms_sort_keys<Cfg>(S.keys, K);
    }
    INMS_STAMP(5);

    // ---- 6. emission ---------------------------------------------------------------------------------------------------
    // thread t owns the PER consecutive score positions PER*t.., so ranks follow the score order
    int flags = 0, cnt = 0;
    u64 mykeys[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int i = tid * PER + q;
        mykeys[q] = 0;
        if (i < K) {
            const u64 key = S.keys[i];
            const int p = (int)(key & (CAP - 1));
            if (S.state[p] == 1 && (k.n_tracked == 0 || inms_tracked(k, cls_at(p)))) {
                mykeys[q] = key;
                flags |= 1 << q;
                ++cnt;
            }
        }
    }
    int total;
    int rank = inms_block_excl_scan<INMS_THREADS>(cnt, S.wsum, total);
    // rows of image b start after the rows of images < b: decoupled look-back over the per-image words
    // (CHAIN_AGG | own count, later CHAIN_PREFIX | inclusive prefix); predecessors hold earlier tickets, so
    // they are running or done and the wait cannot deadlock
    if (wid == 0) {
        if (lane == 0) {
            ((volatile u64 *)k.chain)[b + 1] = CHAIN_AGG | (u64)total;
            k.emit_count[b] = total;
            k.cand_count[b] = K_all;
            k.out_counts[2 + b] = total;
            k.out_counts[2 + k.B + b] = K_all;  // survivors of the score threshold, also when the image was left out
            if (over) atomicOr(&k.hdr->status, BG_STATUS_NEED_GENERAL);
        }
        long long base = 0;
        int hi = b;  // words 1..b belong to images 0..b-1; word 0 is the constant prefix 0
        while (true) {
            const int j = hi - lane;  // word index, newest first
            u64 v = CHAIN_PREFIX;     // lanes past word 0 read as "prefix 0"
            if (j >= 0) {
                do { v = ((volatile u64 *)k.chain)[j]; } while (!(v & (CHAIN_AGG | CHAIN_PREFIX)));
            }
            const u32 pm = __ballot_sync(0xffffffffu, (v & CHAIN_PREFIX) != 0);
            const int stop = pm ? (__ffs(pm) - 1) : 32;  // nearest word that already holds an inclusive prefix
            long long part = (lane <= stop) ? (long long)(v & CHAIN_VALUE) : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            base += part;
            if (pm) break;
            hi -= 32;
        }
        if (lane == 0) {
            S.base = base;
            ((volatile u64 *)k.chain)[b + 1] = CHAIN_PREFIX | (u64)(base + total);
        }
    }
    __syncthreads();
    INMS_STAMP(6);
    const long long base = S.base;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        if (!((flags >> q) & 1)) continue;
        const u64 key = mykeys[q];
        const int p = (int)(key & (CAP - 1));
        const u32 id = (u32)((key & 0xffffffffull) >> Cfg::PBITS);
        const float score = from_orderable(~(u32)(key >> 32));
        const float4 bx = box_at(p);
        if (k.order == 0) {
            const long long dst = base + rank;
            float2 *o = reinterpret_cast<float2 *>(k.out_boxes + dst * 6);  // rows are 24 bytes: 8-byte aligned
            o[0] = make_float2(score, (float)cls_at(p));
            o[1] = make_float2(bx.x, bx.y);
            o[2] = make_float2(bx.z, bx.w);
            k.out_img[dst] = b;
            k.out_keep[dst] = ibase + id;
        } else {
            k.emit_key[ibase + rank] = (key & 0xffffffff00000000ull) | (u64)id;
            k.emit_box[ibase + rank] = bx;
            k.emit_cls[ibase + rank] = cls_at(p);
        }
        ++rank;
    }
    INMS_STAMP(7);

    // last image to finish: totals and status
    if (k.host_flag) __threadfence_system();   // (the outputs may live in mapped host memory)
    else __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned d = atomicAdd(&k.hdr->done, 1u);
        if (d == (unsigned)k.B - 1) {
            __threadfence();
            const u64 v = *((volatile u64 *)(k.chain + k.B));
            k.out_counts[0] = (int)(v & CHAIN_VALUE);
            k.out_counts[1] = *((volatile int *)&k.hdr->status);
            if (k.host_flag && k.order == 0) {   // everything of this call is written: tell the polling host thread
                __threadfence_system();
                *((volatile int32_t *)k.host_flag) = k.host_flag_value;
            }
        }
    }
}

}  // namespace bg
