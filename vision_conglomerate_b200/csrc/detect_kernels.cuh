// Inference side: fused head decode + score + strict threshold + per-image compaction, the row
// assembly after NMS, and the stand-alone decode (DetectionNet._get_scale_pred replacement).
//
// Reference semantics: modules/detection.py:98-190, inference_det.py:57-97,
// utils/utils.py:215-226.  All box arithmetic is fp32 in the reference's operation order with every
// product/sum rounded separately (this TU is built with -fmad=false; the intrinsics make it explicit).
#pragma once
#include "nms.cuh"

namespace bg {

struct ScaleDesc {
    const float *raw;
    int ny, nx;
    int cells_na;      // ny*nx*na  (candidates per image on this scale)
    int img_off;       // flat index of this scale's first candidate inside an image
    long long rows;    // B*ny*nx*na
    float s0, s1;      // float(H/ny) (multiplies x), float(W/nx) (multiplies y)  -- detection.py:147-154
    float fnx, fny;
    float aw[BG_MAX_ANCHORS], ah[BG_MAX_ANCHORS];
};

struct DetectK {
    ScaleDesc sc[3];
    int B, C, D, na, N;  // N = candidates per image over the three scales
    int rescale;         // apply _bbox_to_size (guard at detection.py:76 evaluated on the host)
    float fW, fH, fW0, fH0;
    int use_allowance;
    float allowance;
    float score_thr;
    // outputs of the filter stage
    int *seg_count;          // [B] survivors per image (atomic slots)
    const long long *seg_off;
    u64 *keys;
    float4 *box_dense;       // [B*N] xyxy of survivors, at b*N + idx
    int *cls_dense;          // [B*N]
};

// Decode one candidate's box to xyxy (image or original-frame pixels).
__device__ __forceinline__ float4 decode_xyxy(const DetectK &k, const ScaleDesc &s, float tx, float ty, float tw,
                                              float th, int x, int y, int a)
{
    float bx = __fsub_rn(__fmul_rn(sigmoid_acc(tx), 2.0f), 0.5f);
    float by = __fsub_rn(__fmul_rn(sigmoid_acc(ty), 2.0f), 0.5f);
    float bw = __fmul_rn(sigmoid_acc(tw), 2.0f); bw = __fmul_rn(bw, bw);
    float bh = __fmul_rn(sigmoid_acc(th), 2.0f); bh = __fmul_rn(bh, bh);
    bx = __fmul_rn(__fadd_rn(bx, (float)x), s.s0);
    by = __fmul_rn(__fadd_rn(by, (float)y), s.s1);
    bw = __fmul_rn(__fmul_rn(__fmul_rn(bw, s.aw[a]), s.fnx), s.s0);
    bh = __fmul_rn(__fmul_rn(__fmul_rn(bh, s.ah[a]), s.fny), s.s1);
    if (k.rescale) {
        bx = __fmul_rn(__fdiv_rn(bx, k.fW), k.fW0);
        by = __fmul_rn(__fdiv_rn(by, k.fH), k.fH0);
        bw = __fmul_rn(__fdiv_rn(bw, k.fW), k.fW0);
        bh = __fmul_rn(__fdiv_rn(bh, k.fH), k.fH0);
    }
    if (k.use_allowance) { bw = __fadd_rn(bw, k.allowance); bh = __fadd_rn(bh, k.allowance); }
    const float x1 = __fsub_rn(bx, __fdiv_rn(bw, 2.0f)), y1 = __fsub_rn(by, __fdiv_rn(bh, 2.0f));
    return make_float4(x1, y1, __fadd_rn(x1, bw), __fadd_rn(y1, bh));
}

// The reference takes argmax over sigmoid(cls) (inference_det.py:58,95): distinct logits whose fp32
// sigmoids coincide tie, and the first index wins.  Logits further than this window below the maximum
// cannot share its sigmoid (ulp(p) <= 1.2e-7*p and sigmoid' = p(1-p)); p == 1 gives an infinite window.
__device__ __forceinline__ float tie_window(float pm) { return __fdividef(5e-7f, 1.0f - pm); }

__device__ __forceinline__ void row_coords(const ScaleDesc &s, int na, long long row, int &b, int &x, int &y, int &a,
                                           int &idx)
{
    b = (int)(row / s.cells_na);
    const int rl = (int)(row - (long long)b * s.cells_na);
    a = rl % na;
    const int cell = rl / na;
    x = cell % s.nx;
    y = cell / s.nx;
    idx = s.img_off + rl;
}

// ---------------------------------------------------------------------------------------------
// init: header, per-image counters and fixed segment offsets
// ---------------------------------------------------------------------------------------------
__global__ void detect_init_kernel(SegNms p, int B, long long seg_stride, int32_t *out_counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        SegHdr h;
        h.S = B; h.status = 0; h.item_ctr = 0; h.reduce_done = 0; h.gmin = 0; h.gmax = 0; h.total_out = 0;
        h.pad[0] = h.pad[1] = h.pad[2] = 0;
        *p.hdr = h;
        out_counts[0] = 0;
        out_counts[1] = 0;
    }
    if (i < B) p.seg_count[i] = 0;
    if (i <= B) p.seg_off[i] = (long long)i * seg_stride;
}

// ---------------------------------------------------------------------------------------------
// variant 1: one warp per candidate row, plain coalesced loads
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) decode_filter_warp_kernel(DetectK k)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = k.sc[0].rows + k.sc[1].rows + k.sc[2].rows;
    const int C = k.C, D = k.D;
    for (long long g = warp; g < total; g += nwarps) {
        int si = 0;
        long long row = g;
        if (row >= k.sc[0].rows) { row -= k.sc[0].rows; si = 1; }
        if (si == 1 && row >= k.sc[1].rows) { row -= k.sc[1].rows; si = 2; }
        const ScaleDesc &s = k.sc[si];
        const float *rp = s.raw + row * D;
        float m = -INFINITY;
        float obj = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
        for (int e = lane; e < D; e += 32) {
            const float v = __ldg(rp + e);
            if (e == 0) obj = v;
            else if (e <= C) m = fmaxf(m, v);
            else if (e == C + 1) t0 = v;
            else if (e == C + 2) t1 = v;
            else if (e == C + 3) t2 = v;
            else t3 = v;
        }
        // max logit across the warp (sigmoid is monotone: max_c sig(cls_c) == sig(max_c cls_c))
        float wm = m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
        obj = __shfl_sync(0xffffffffu, obj, 0);
        const float pm = sigmoid_acc(wm);
        const float score = __fmul_rn(pm, sigmoid_acc(obj));
        if (!(score > k.score_thr)) continue;  // warp-uniform
        // class id = first index whose sigmoid equals the maximum sigmoid (torch argmax over probabilities)
        int ci = 0x7fffffff;
        const float win = tie_window(pm);
        for (int e = 1 + lane; e <= C; e += 32) {
            const float v = __ldg(rp + e);
            if (v == wm || (v >= wm - win && sigmoid_acc(v) == pm)) { ci = min(ci, e - 1); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ci = min(ci, __shfl_xor_sync(0xffffffffu, ci, o));
        t0 = __shfl_sync(0xffffffffu, t0, (C + 1) & 31);
        t1 = __shfl_sync(0xffffffffu, t1, (C + 2) & 31);
        t2 = __shfl_sync(0xffffffffu, t2, (C + 3) & 31);
        t3 = __shfl_sync(0xffffffffu, t3, (C + 4) & 31);
        if (lane == 0) {
            int b, x, y, a, idx;
            row_coords(s, k.na, row, b, x, y, a, idx);
            const float4 bx = decode_xyxy(k, s, t0, t1, t2, t3, x, y, a);
            const int slot = atomicAdd(&k.seg_count[b], 1);
            k.keys[k.seg_off[b] + slot] = make_key(score, (u32)idx);
            k.box_dense[(long long)b * k.N + idx] = bx;
            k.cls_dense[(long long)b * k.N + idx] = ci;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// variant 2: persistent CTAs, several per SM, each streaming tiles of TR candidate rows into shared
// memory with one TMA bulk copy (cp.async.bulk + mbarrier) per tile.  Phase 1: one thread per row
// takes the class maximum and the score from shared memory (row stride D = 5+C words: conflict-free
// when D is odd, e.g. 85) and the survivors of the threshold are compacted into a small list;
// phase 2: one thread per *survivor* finds the class id, decodes the box and writes the candidate.
// The co-resident CTAs of an SM overlap each other's copy latency (4 x 43.5 KB in flight per SM).
// ---------------------------------------------------------------------------------------------
constexpr int TMA_THREADS = 128;
constexpr int TMA_CTAS_PER_SM = 4;
constexpr int TMA_TILE_BYTES = 44 * 1024;

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct TileMap { int tiles[3]; int total; int TR; };

__device__ __forceinline__ void tile_locate(const DetectK &k, const TileMap &tm, int tile, int &si, long long &row0,
                                            int &rows)
{
    si = 0;
    int t = tile;
    if (t >= tm.tiles[0]) { t -= tm.tiles[0]; si = 1; }
    if (si == 1 && t >= tm.tiles[1]) { t -= tm.tiles[1]; si = 2; }
    row0 = (long long)t * tm.TR;
    const long long rem = k.sc[si].rows - row0;
    rows = rem < tm.TR ? (int)rem : tm.TR;
}

struct Surv { int row; float score, wm, pm; };

template <int CT>  // compile-time class count (fully unrolled row scan); 0 = take it from the parameters
__global__ void __launch_bounds__(TMA_THREADS, TMA_CTAS_PER_SM) decode_filter_tma_kernel(DetectK k, TileMap tm)
{
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ __align__(8) u64 full_bar;
    __shared__ Surv s_surv[TMA_THREADS];
    __shared__ int s_n;
    const int C = CT ? CT : k.C;
    const int D = C + 5;
    float *tile = reinterpret_cast<float *>(tma_smem);
    const int tid = threadIdx.x, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_n = 0;
    }
    __syncthreads();

    u32 parity = 0;
    for (int t = blockIdx.x; t < tm.total; t += gridDim.x) {
        int si, rows;
        long long row0;
        tile_locate(k, tm, t, si, row0, rows);
        const ScaleDesc &s = k.sc[si];
        const float *src = s.raw + row0 * D;
        const int nfl = rows * D;
        if ((nfl & 3) == 0) {  // the tile is a whole number of 16-byte units: one bulk copy
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the tile buffer
                mbar_expect_tx(&full_bar, (u32)nfl * 4);
                bulk_g2s(tile, src, (u32)nfl * 4, &full_bar);
            }
            mbar_wait(&full_bar, parity);
            parity ^= 1;
        } else {  // ragged last tile of a scale: plain coalesced loads
            for (int i = tid; i < nfl; i += TMA_THREADS) tile[i] = __ldg(src + i);
            __syncthreads();
        }

        // phase 1: one thread per row -- class maximum, score, threshold
        bool alive = false;
        float score = 0.f, wm = -INFINITY, pm = 0.f;
        if (tid < rows) {
            const float *sr = tile + tid * D;
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
            if (CT) {
#pragma unroll
                for (int c = 0; c + 3 < CT; c += 4) {
                    m0 = fmaxf(m0, sr[1 + c]); m1 = fmaxf(m1, sr[2 + c]);
                    m2 = fmaxf(m2, sr[3 + c]); m3 = fmaxf(m3, sr[4 + c]);
                }
#pragma unroll
                for (int c = CT & ~3; c < CT; ++c) m0 = fmaxf(m0, sr[1 + c]);
            } else {
                int c = 0;
                for (; c + 3 < C; c += 4) {
                    m0 = fmaxf(m0, sr[1 + c]); m1 = fmaxf(m1, sr[2 + c]);
                    m2 = fmaxf(m2, sr[3 + c]); m3 = fmaxf(m3, sr[4 + c]);
                }
                for (; c < C; ++c) m0 = fmaxf(m0, sr[1 + c]);
            }
            wm = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            pm = sigmoid_acc(wm);
            score = __fmul_rn(pm, sigmoid_acc(sr[0]));
            alive = score > k.score_thr;
        }
        {   // compact the survivors of this tile (order inside the list is irrelevant)
            const u32 am = __ballot_sync(0xffffffffu, alive);
            int base = 0;
            if (am) {
                const int leader = __ffs(am) - 1;
                if (lane == leader) base = atomicAdd(&s_n, __popc(am));
                base = __shfl_sync(0xffffffffu, base, leader);
            }
            if (alive) s_surv[base + __popc(am & lanemask_lt())] = Surv{tid, score, wm, pm};
        }
        __syncthreads();
        const int n = s_n;

        // phase 2: one thread per survivor -- class id, box decode, per-image slot, writes
        for (int q0 = 0; q0 < n; q0 += TMA_THREADS) {
            const int q = q0 + tid;
            const bool act = q < n;
            const u32 amask = __ballot_sync(0xffffffffu, act);
            if (act) {
                const Surv sv = s_surv[q];
                const float *sr = tile + sv.row * D;
                int ci = 0;
                const float lo = sv.wm - tie_window(sv.pm);
                for (int c = 0; c < C; ++c) {
                    const float v = sr[1 + c];
                    if (v >= lo && (v == sv.wm || sigmoid_acc(v) == sv.pm)) { ci = c; break; }
                }
                int b, x, y, a, idx;
                row_coords(s, k.na, row0 + sv.row, b, x, y, a, idx);
                const float4 bx = decode_xyxy(k, s, sr[C + 1], sr[C + 2], sr[C + 3], sr[C + 4], x, y, a);
                // warp-aggregated slot allocation, one atomic per (warp, image)
                const u32 peers = __match_any_sync(amask, b);
                const int leader = __ffs(peers) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(&k.seg_count[b], __popc(peers));
                base = __shfl_sync(peers, base, leader);
                const int slot = base + __popc(peers & lanemask_lt());
                k.keys[k.seg_off[b] + slot] = make_key(sv.score, (u32)idx);
                k.box_dense[(long long)b * k.N + idx] = bx;
                k.cls_dense[(long long)b * k.N + idx] = ci;
            }
        }
        __syncthreads();  // the tile buffer and the survivor list are free again
        if (tid == 0) s_n = 0;  // ordered before the next phase 1 by the mbarrier (release/acquire) or the barrier above
    }
}

// ---------------------------------------------------------------------------------------------
// row assembly after NMS: pred_boxes = [score, class, x1,y1,x2,y2], sample index, flat keep index
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) detect_output_kernel(SegNms p, DetectK k, int order, float *out_boxes,
                                                            long long *out_img, long long *out_keep,
                                                            int32_t *out_counts)
{
    const int S = k.B;
    for (int seg = blockIdx.y; seg < S; seg += gridDim.y) {
        const int cnt = p.emit_count[seg];
        const long long off = p.seg_off[seg];
        const long long base = p.out_prefix[seg];
        if (threadIdx.x == 0 && blockIdx.x == 0) out_counts[2 + S + seg] = p.seg_count[seg];
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < cnt; r += gridDim.x * blockDim.x) {
            const u64 key = p.emit_key[off + r];
            const u32 pos = p.emit_pos[off + r];
            const u32 id = key_id(key);
            const float4 b = p.sorted_box[off + pos];
            const long long dst = order ? segnms_global_rank(p, S, seg, r, key) : base + r;
            float *o = out_boxes + dst * 6;
            o[0] = key_score(key);
            o[1] = (float)k.cls_dense[(long long)seg * k.N + id];
            o[2] = b.x; o[3] = b.y; o[4] = b.z; o[5] = b.w;
            out_img[dst] = seg;
            out_keep[dst] = (long long)seg * k.N + id;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// stand-alone decode of one scale (DetectionNet._get_scale_pred [+ _bbox_to_size])
// ---------------------------------------------------------------------------------------------
struct DecodeK {
    const float *raw;
    float *out;
    long long rows;
    int ny, nx, na, C, D;
    int inference, rescale;
    float s0, s1, fnx, fny, fW, fH, fW0, fH0;
    float aw[BG_MAX_ANCHORS], ah[BG_MAX_ANCHORS];
};

__global__ void __launch_bounds__(256) decode_scale_kernel(DecodeK k)
{
    const long long total = k.rows * k.D;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / k.D;
        const int c = (int)(e - row * k.D);
        float v = __ldg(k.raw + e);
        if (c > k.C) {
            const int q = c - k.C - 1;  // 0:x 1:y 2:w 3:h
            const int rl = (int)(row % ((long long)k.ny * k.nx * k.na));
            const int a = rl % k.na;
            const int cell = rl / k.na;
            const float sg = __fmul_rn(sigmoid_acc(v), 2.0f);
            if (q < 2) {
                v = __fsub_rn(sg, 0.5f);
                if (k.inference) {
                    const float g = (q == 0) ? (float)(cell % k.nx) : (float)(cell / k.nx);
                    v = __fmul_rn(__fadd_rn(v, g), q == 0 ? k.s0 : k.s1);
                }
            } else {
                v = __fmul_rn(sg, sg);
                if (k.inference)
                    v = __fmul_rn(__fmul_rn(__fmul_rn(v, q == 2 ? k.aw[a] : k.ah[a]), q == 2 ? k.fnx : k.fny),
                                  q == 2 ? k.s0 : k.s1);
            }
            if (k.inference && k.rescale) {
                const bool isx = (q == 0 || q == 2);
                v = __fmul_rn(__fdiv_rn(v, isx ? k.fW : k.fH), isx ? k.fW0 : k.fH0);
            }
        }
        k.out[e] = v;
    }
}

}  // namespace bg
