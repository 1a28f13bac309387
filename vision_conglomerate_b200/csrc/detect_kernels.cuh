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
    long long img_stride;  // rows between consecutive images: cells_na (own tensor per scale) or N (slices of [B,N,D])
    int img_off;       // flat index of this scale's first candidate inside an image
    long long rows;    // B*ny*nx*na
    u32 magic_nx;      // ceil(2^32 / nx): n / nx == umulhi(n, magic_nx) for the cell counts in range (n * nx < 2^32)
    float s0, s1;      // float(H/ny) (multiplies x), float(W/nx) (multiplies y)  -- detection.py:147-154
    float fnx, fny;
    float aw[BG_MAX_ANCHORS], ah[BG_MAX_ANCHORS];
};

struct DetectK {
    ScaleDesc sc[3];
    int B, C, D, na, N;  // N = candidates per image over the three scales
    u32 magic_na;        // ceil(2^32 / na)
    int predecoded;      // rows already hold decoded pixel xywh (DetectionNet.forward output): no sigmoid / grid / anchor step
    int rescale;         // apply _bbox_to_size (guard at detection.py:76 evaluated on the host)
    float fW, fH, fW0, fH0;
    int use_allowance;
    float allowance;
    float score_thr;
    // outputs of the filter stage (see decode_filter_kernel): the survivors of a tile sit, in candidate order,
    // in the tile's own slots b*N + (first candidate of the tile) + j
    u64 *keys;               // [B*N] sort keys (score desc, candidate index asc)
    float4 *box_slots;       // [B*N] xyxy
    int *cls_slots;          // [B*N] class id
    // general NMS path only: the same data addressed by candidate index b*N + idx (filled by detect_compact_kernel)
    float4 *box_dense;
    int *cls_dense;
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
    // w / 2 == w * 0.5 exactly (power-of-two scaling)
    const float x1 = __fsub_rn(bx, __fmul_rn(bw, 0.5f)), y1 = __fsub_rn(by, __fmul_rn(bh, 0.5f));
    return make_float4(x1, y1, __fadd_rn(x1, bw), __fadd_rn(y1, bh));
}

// The reference takes argmax over sigmoid(cls) (inference_det.py:58,95): distinct logits whose fp32
// sigmoids coincide tie, and the first index wins.  Logits further than this window below the maximum
// cannot share its sigmoid (ulp(p) <= 1.2e-7*p and sigmoid' = p(1-p)); p == 1 gives an infinite window.
__device__ __forceinline__ float tie_window(float pm) { return __fdividef(5e-7f, 1.0f - pm); }

// ---------------------------------------------------------------------------------------------
// tile plan: the three head tensors are cut into tiles of TR candidate rows that never straddle an
// image; tile t = b * tpi_total + r, r running over (scale, tile inside the image's slice).
// ---------------------------------------------------------------------------------------------
struct TilePlan {
    int TR;          // rows per tile (multiple of 4, <= DEC_THREADS)
    int tpi[3];      // tiles per image on each scale
    int tpi_total;   // tiles per image
    int total;       // B * tpi_total
};

// (b, r) = (image, tile inside the image) of tile t = b * tpi_total + r; the kernel advances them incrementally
__host__ __device__ inline void tile_locate(const DetectK &k, const TilePlan &tp, int b, int r, int &si, int &lrow0, int &rows)
{
    (void)b;
    si = 0;
    if (r >= tp.tpi[0]) { r -= tp.tpi[0]; si = 1; if (r >= tp.tpi[1]) { r -= tp.tpi[1]; si = 2; } }
    lrow0 = r * tp.TR;
    const int rem = k.sc[si].cells_na - lrow0;
    rows = rem < tp.TR ? rem : tp.TR;
}

// ---------------------------------------------------------------------------------------------
// decode + score + threshold + per-tile compaction.
// Persistent CTAs (DEC_CTAS_PER_SM per SM), each owning a ring of DEC_STAGES shared-memory tile
// buffers filled by TMA bulk copies (cp.async.bulk + mbarrier): while a tile is being processed the
// next DEC_STAGES-1 tiles of the CTA are already in flight, so every SM keeps >= 2 x 43.5 KB of
// reads outstanding (HBM latency x bandwidth needs ~45 KB per SM).
//   phase 1  one thread per row: class maximum and score from shared memory (row stride D = 5+C
//            words, conflict-free when D is odd), strict threshold, survivors listed in smem;
//   phase 2  each warp finishes its own survivors, four at a time: eight lanes scan the class logits for the
//            class id (first index whose sigmoid equals the maximum sigmoid), four lanes decode one box
//            coordinate each and combine them with two shuffles; no barrier inside the phase.
// Survivors of tile t go to the tile's own slots (b*N + first candidate of the tile + j, candidate order)
// and their count to tile_count[t]: no global atomics, nothing to zero between calls.
// ---------------------------------------------------------------------------------------------
constexpr int DEC_THREADS = 128;
constexpr int DEC_STAGES = 2;
constexpr int DEC_CTAS_PER_SM = 2;
constexpr int DEC_TILE_BYTES = 43520;  // 128 rows x 85 floats

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

struct Surv { int row; float score, wm, pm; int ci; };

struct FusedHdr {            // scratch of the fused path, (re)initialised by the decode kernel of every call
    unsigned ticket;         // image scheduler of image_nms_kernel
    unsigned done;           // images finished
    int status;              // BG_STATUS_* bits
    int pad;
};
// per-image words of the output-offset look-back chain (word 0 = prefix 0, word b+1 = image b)
constexpr u64 CHAIN_PREFIX = 1ull << 63;  // value = rows emitted by images <= b
constexpr u64 CHAIN_AGG = 1ull << 62;     // value = rows emitted by image b alone
constexpr u64 CHAIN_VALUE = (1ull << 62) - 1;

struct DecodeOut {
    int *tile_count;         // [B * tpi_total]
    FusedHdr *hdr;
    u64 *chain;              // [B+1] look-back words
    long long *seg_off;      // [B+1] b*N (row offset of the image's emit list)
    int force_plain;         // 1: never use TMA (unaligned inputs, variant 1)
    u32 *gflag;              // [4B] hand-over flags and spill counters of the per-image NMS (zeroed here)
    unsigned long long *cycles;  // optional profiling hook: [gridDim.x, 8] SM clock cycles thread 0 spent per phase
};

template <int CT>  // compile-time class count (fully unrolled row scan); 0 = take it from the parameters
__global__ void __launch_bounds__(DEC_THREADS, DEC_CTAS_PER_SM) decode_filter_kernel(DetectK k, TilePlan tp, DecodeOut o)
{
    extern __shared__ __align__(128) unsigned char dec_smem[];
    __shared__ __align__(8) u64 full_bar[DEC_STAGES];
    __shared__ Surv s_surv[DEC_THREADS];
    __shared__ int s_wcnt[DEC_THREADS / 32];
    const int C = CT ? CT : k.C;
    const int D = CT ? CT + 5 : k.D;  // row stride; the runtime variant also serves rows with trailing extra columns
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int stage_floats = tp.TR * D;
    float *ring = reinterpret_cast<float *>(dec_smem);

    // let the dependent per-image NMS kernel be scheduled as SMs free up (it waits for this grid to finish)
    cudaTriggerProgrammaticLaunchCompletion();
    // scratch of the per-image NMS kernel that follows in the stream
    if (blockIdx.x == 0) {
        if (tid == 0) { o.hdr->ticket = 0; o.hdr->done = 0; o.hdr->status = 0; o.hdr->pad = 0; }
        for (int b = tid; b <= k.B; b += DEC_THREADS) {
            o.chain[b] = (b == 0) ? CHAIN_PREFIX : 0ull;
            o.seg_off[b] = (long long)b * k.N;
        }
        for (int i = tid; i < 4 * k.B; i += DEC_THREADS) o.gflag[i] = 0u;
    }

    auto tile_src = [&](int b, int r, int &si, int &lrow0, int &rows) -> const float * {
        tile_locate(k, tp, b, r, si, lrow0, rows);
        return k.sc[si].raw + ((long long)b * k.sc[si].img_stride + lrow0) * D;
    };
    // a CTA steps its tile number by gridDim.x: (image, tile in image) advance without divisions
    const int step_b = (int)gridDim.x / tp.tpi_total, step_r = (int)gridDim.x - step_b * tp.tpi_total;
    auto advance = [&](int &b, int &r) { b += step_b; r += step_r; if (r >= tp.tpi_total) { r -= tp.tpi_total; ++b; } };
    auto tma_ok = [&](const float *src, int rows) -> bool {
        return !o.force_plain && (((rows * D) & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    };

    if (tid == 0) {
        for (int s = 0; s < DEC_STAGES; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        int pb = (int)blockIdx.x / tp.tpi_total, pr = (int)blockIdx.x - pb * tp.tpi_total;
        for (int s = 0; s < DEC_STAGES; ++s) {  // prologue: fill the ring
            if (pb < k.B) {
                int si, lrow0, rows;
                const float *src = tile_src(pb, pr, si, lrow0, rows);
                if (tma_ok(src, rows)) {
                    mbar_expect_tx(&full_bar[s], (u32)(rows * D) * 4);
                    bulk_g2s(ring + (size_t)s * stage_floats, src, (u32)(rows * D) * 4, &full_bar[s]);
                }
            }
            advance(pb, pr);
        }
    }
    __syncthreads();

    u32 phases = 0;  // bit s = parity the next wait on stage s expects
    int it = 0;
    long long cyc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c_prev = 0;
    const bool prof = o.cycles != nullptr && tid == 0;
#define DEC_MARK(i) do { if (prof) { const long long c_now = clock64(); cyc[i] += c_now - c_prev; c_prev = c_now; } } while (0)
    if (prof) c_prev = clock64();
    int b = (int)blockIdx.x / tp.tpi_total, tr = (int)blockIdx.x - b * tp.tpi_total;  // this tile
    int nb = b, nr = tr;                                                             // the tile DEC_STAGES steps ahead
    for (int s = 0; s < DEC_STAGES; ++s) advance(nb, nr);
    for (; b < k.B; advance(b, tr), advance(nb, nr), ++it) {
        const int t = b * tp.tpi_total + tr;
        const int stage = it % DEC_STAGES;
        float *tile = ring + (size_t)stage * stage_floats;
        int si, lrow0, rows;
        const float *src = tile_src(b, tr, si, lrow0, rows);
        const ScaleDesc &s = k.sc[si];
        // thread 0 works out the refill of this stage (the tile DEC_STAGES iterations ahead) now, while the CTA is
        // about to wait for data anyway; at the end of the iteration it only has to issue the copy
        const float *rsrc = nullptr;
        u32 rbytes = 0;
        if (tid == 0 && nb < k.B) {
            int si2, l2, r2;
            const float *src2 = tile_src(nb, nr, si2, l2, r2);
            if (tma_ok(src2, r2)) { rsrc = src2; rbytes = (u32)(r2 * D) * 4; }
        }
        if (tma_ok(src, rows)) {
            mbar_wait(&full_bar[stage], (phases >> stage) & 1u);
            phases ^= 1u << stage;
        } else {  // ragged / unaligned tile: plain coalesced loads
            const int nfl = rows * D;
            for (int i = tid; i < nfl; i += DEC_THREADS) tile[i] = __ldg(src + i);
            __syncthreads();
        }

        DEC_MARK(0);  // tile location + wait for the tile
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
        // list the survivors of this tile in row order (warp ballots + a prefix over the four warps), so that
        // slot order == candidate order and ties in score keep torchvision's lower-index-first rule downstream
        DEC_MARK(1);  // phase 1
        const u32 am = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) s_wcnt[wid] = __popc(am);
        __syncthreads();
        DEC_MARK(2);  // barrier: slowest warp of phase 1
        int base = 0, n = 0;
#pragma unroll
        for (int w2 = 0; w2 < DEC_THREADS / 32; ++w2) {
            const int c = s_wcnt[w2];
            if (w2 < wid) base += c;
            n += c;
        }
        if (alive) s_surv[base + __popc(am & lanemask_lt())] = Surv{tid, score, wm, pm, 0};
        __syncthreads();
        DEC_MARK(3);  // survivor list

        // phase 2: every warp finishes its own survivors (q = wid, wid + 4, ...), four at a time -- no barrier
        // between the class scan and the box decode.
        //  class id: eight lanes per survivor scan the class logits for the first one equal to the maximum (the
        //   first index whose sigmoid equals the maximum sigmoid; a smaller logit can only tie inside the tie
        //   window, which is rare and then re-checked with sigmoids);
        //  box: four lanes per survivor, one coordinate each (x, y, w, h), combined with two shuffles.
        if (tid == 0) o.tile_count[t] = n;
        {
            const int grp = lane >> 3, gl = lane & 7;
            for (int q0 = wid; q0 < n; q0 += 4 * (DEC_THREADS / 32)) {   // warp-uniform trip count
                const int q = q0 + grp * (DEC_THREADS / 32);
                const bool act = q < n;
                const Surv sv = s_surv[act ? q : 0];
                const float *sr = tile + sv.row * D;
                const float lo = sv.wm - tie_window(sv.pm);
                int first = 0x7fffffff, near = 0;
                if (CT) {   // lane's classes gl, gl+8, ...: all loads in flight, then one select chain (descending)
                    constexpr int NI = (CT + 7) / 8;
                    float v[NI ? NI : 1];
#pragma unroll
                    for (int i = 0; i < NI; ++i) { const int c = gl + 8 * i; v[i] = c < CT ? sr[1 + c] : -INFINITY; }
#pragma unroll
                    for (int i = NI - 1; i >= 0; --i) {
                        first = (v[i] == sv.wm) ? gl + 8 * i : first;
                        near |= (int)(v[i] >= lo) & (int)(v[i] < sv.wm);
                    }
                } else {
                    for (int c = C - 1 - ((C - 1 - gl) & 7); c >= 0; c -= 8) {
                        const float v = sr[1 + c];
                        first = (v == sv.wm) ? c : first;
                        near |= (int)(v >= lo) & (int)(v < sv.wm);
                    }
                }
#pragma unroll
                for (int o2 = 1; o2 < 8; o2 <<= 1) {
                    first = min(first, __shfl_xor_sync(0xffffffffu, first, o2));
                    near |= __shfl_xor_sync(0xffffffffu, near, o2);
                }
                if (near && act) {   // some logit ties the maximum after rounding through the sigmoid?  (rare)
                    for (int c = 0; c < first && c < C; ++c) {
                        const float v = sr[1 + c];
                        if (v >= lo && sigmoid_acc(v) == sv.pm) { first = c; break; }
                    }
                }
                if (first == 0x7fffffff) first = 0;
                // box: lane gl < 4 owns coordinate gl
                const int comp = gl & 3;
                const int rl = lrow0 + sv.row;
                const int cell = (int)__umulhi((u32)rl, k.magic_na), a = rl - cell * k.na;      // rl / na, rl % na
                const int gy = (int)__umulhi((u32)cell, s.magic_nx), gx = cell - gy * s.nx;     // cell / nx, cell % nx
                const bool isx = (comp & 1) == 0;
                const float tv = sr[C + 1 + comp];
                const float sg2 = k.predecoded ? 0.f : __fmul_rn(sigmoid_acc(tv), 2.0f);
                float val;
                if (k.predecoded) val = tv;
                else if (comp < 2) val = __fmul_rn(__fadd_rn(__fsub_rn(sg2, 0.5f), isx ? (float)gx : (float)gy), isx ? s.s0 : s.s1);
                else val = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(sg2, sg2), isx ? s.aw[a] : s.ah[a]), isx ? s.fnx : s.fny), isx ? s.s0 : s.s1);
                if (k.rescale) val = __fmul_rn(__fdiv_rn(val, isx ? k.fW : k.fH), isx ? k.fW0 : k.fH0);
                if (comp >= 2 && k.use_allowance) val = __fadd_rn(val, k.allowance);
                const float other = __shfl_xor_sync(0xffffffffu, val, 2);                   // x <-> w, y <-> h
                const float lo_c = __fsub_rn(val, __fmul_rn(other, 0.5f));                  // lanes 0,1: x1 = x - w/2, y1 = y - h/2
                const float lo_o = __shfl_xor_sync(0xffffffffu, lo_c, 2);                   // lanes 2,3 receive x1, y1
                const float outv = comp < 2 ? lo_c : __fadd_rn(lo_o, val);                  // x2 = x1 + w, y2 = y1 + h
                if (act) {
                    const long long slot = (long long)b * k.N + s.img_off + lrow0 + q;
                    if (gl < 4) reinterpret_cast<float *>(k.box_slots + slot)[comp] = outv;
                    if (gl == 4) k.keys[slot] = make_key(sv.score, (u32)(s.img_off + rl));
                    if (gl == 5) k.cls_slots[slot] = first;
                }
            }
        }
        DEC_MARK(5);  // phase 2 (this warp)
        __syncthreads();  // the tile buffer and the survivor list are free again
        DEC_MARK(6);  // barrier: slowest thread of phase 2b

        if (tid == 0 && rsrc) {   // refill this stage with the tile DEC_STAGES iterations ahead
            // the generic-proxy reads of this buffer are ordered before the copy by the barrier above (write after
            // read needs no proxy fence; only generic writes must be fenced before an async-proxy access)
            mbar_expect_tx(&full_bar[stage], rbytes);
            bulk_g2s(tile, rsrc, rbytes, &full_bar[stage]);
        }
        DEC_MARK(7);  // refill
    }
    if (prof) for (int i = 0; i < 8; ++i) o.cycles[(long long)blockIdx.x * 8 + i] = (unsigned long long)cyc[i];
#undef DEC_MARK
}

// ---------------------------------------------------------------------------------------------
// row assembly after NMS from the per-image emit lists (general NMS path, and the globally ordered
// output of the fused path): pred_boxes = [score, class, x1,y1,x2,y2], sample index, flat keep index
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) detect_output_kernel(SegNms p, DetectK k, int order, const int *cand_count,
                                                            const float4 *emit_box, const int *emit_cls,
                                                            float *out_boxes, long long *out_img, long long *out_keep,
                                                            int32_t *out_counts)
{
    const int S = k.B;
    for (int seg = blockIdx.y; seg < S; seg += gridDim.y) {
        const int cnt = p.emit_count[seg];
        const long long off = p.seg_off[seg];
        const long long base = order ? 0 : p.out_prefix[seg];
        if (cand_count && threadIdx.x == 0 && blockIdx.x == 0) out_counts[2 + S + seg] = cand_count[seg];
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < cnt; r += gridDim.x * blockDim.x) {
            const u64 key = p.emit_key[off + r];
            const u32 id = key_id(key);
            // box / class of the row: emit-list aligned (per-image NMS path) or by candidate index (general path)
            const float4 b = emit_box ? emit_box[off + r] : k.box_dense[(long long)seg * k.N + id];
            const int cl = emit_cls ? emit_cls[off + r] : k.cls_dense[(long long)seg * k.N + id];
            const long long dst = order ? segnms_global_rank(p, S, seg, r, key) : base + r;
            float *o = out_boxes + dst * 6;
            o[0] = key_score(key);
            o[1] = (float)cl;
            o[2] = b.x; o[3] = b.y; o[4] = b.z; o[5] = b.w;
            out_img[dst] = seg;
            out_keep[dst] = (long long)seg * k.N + id;
        }
    }
}

// general NMS path: gather the tile slots of every image into the segment layout of the segmented NMS
// engine (keys of image b compacted at seg_off[b] = b * seg_stride) and initialise its header.
// bg_detect_params.host_flag for the paths whose last kernel has no single last writer: everything the stream did before
// this launch is complete; make it visible to the host and tell the polling thread
__global__ void host_flag_kernel(int32_t *flag, int value)
{
    __threadfence_system();
    *((volatile int32_t *)flag) = value;
}

__global__ void __launch_bounds__(1024) detect_compact_kernel(SegNms p, DetectK k, TilePlan tp, const int *tile_count,
                                                              long long seg_stride, int32_t *out_counts)
{
    __shared__ int s_pref[1025];
    __shared__ int s_wsum[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (blockIdx.x == 0 && tid == 0) {
        SegHdr h;
        h.S = k.B; h.status = 0; h.item_ctr = 0; h.reduce_done = 0; h.gmin = 0; h.gmax = 0; h.total_out = 0;
        h.item_ctr2 = 0; h.overflow = 0; h.dense_fits = 0; h.pad0 = 0; h.pad[0] = 0;
        *p.hdr = h;
        out_counts[0] = 0;
        out_counts[1] = 0;
    }
    if (blockIdx.x == 0)
        for (int b = tid; b <= k.B; b += 1024) p.seg_off[b] = (long long)b * seg_stride;
    for (int b = blockIdx.x; b < k.B; b += gridDim.x) {
        if (tid == 0) s_carry = 0;
        __syncthreads();
        u64 *dst = p.keys + (long long)b * seg_stride;
        for (int r0 = 0; r0 < tp.tpi_total; r0 += 1024) {
            const int r = r0 + tid;
            const int c = (r < tp.tpi_total) ? tile_count[(long long)b * tp.tpi_total + r] : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            if (lane == 31) s_wsum[wid] = inc;
            __syncthreads();
            if (wid == 0) {
                int w = s_wsum[lane], winc = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, winc, o);
                    if (lane >= o) winc += u;
                }
                s_wsum[lane] = winc - w;
            }
            __syncthreads();
            const int carry = s_carry;
            s_pref[tid] = carry + s_wsum[wid] + inc - c;
            if (tid == 1023) s_pref[1024] = carry + s_wsum[wid] + inc;
            __syncthreads();
            // one warp per tile: coalesced copy of its slots
            const int nt = min(1024, tp.tpi_total - r0);
            for (int q = wid; q < nt; q += 32) {
                const int rr = r0 + q;
                int si = 0, lr = rr;
                if (lr >= tp.tpi[0]) { lr -= tp.tpi[0]; si = 1; if (lr >= tp.tpi[1]) { lr -= tp.tpi[1]; si = 2; } }
                const long long src = (long long)b * k.N + k.sc[si].img_off + (long long)lr * tp.TR;
                const int o0 = s_pref[q], cnt = s_pref[q + 1] - o0;
                for (int j = lane; j < cnt; j += 32) {
                    const u64 key = k.keys[src + j];
                    dst[o0 + j] = key;
                    k.box_dense[(long long)b * k.N + key_id(key)] = k.box_slots[src + j];
                    k.cls_dense[(long long)b * k.N + key_id(key)] = k.cls_slots[src + j];
                }
            }
            __syncthreads();
            if (tid == 0) s_carry = s_pref[1024];
            __syncthreads();
        }
        if (tid == 0) p.seg_count[b] = s_carry;
        __syncthreads();
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
    int tanh_cols;   // the first tanh_cols columns behind the box go through tanh (mask coefficients, detection.py:131-134)
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
        if (c > k.C + 4) {
            if (c - k.C - 5 < k.tanh_cols) v = tanhf(v);
        } else if (c > k.C) {
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

// Backward of the training-mode decode (modules/detection.py:122,125): grad_raw = grad_out on the objectness / class
// columns, grad_out * 2s(1-s) on x, y and grad_out * 8s^2(1-s) on w, h, with s = sigmoid(raw).
__global__ void __launch_bounds__(256) decode_train_bwd_kernel(const float *raw, const float *go, float *gr, long long rows, int C,
                                                               int extra_cols, int tanh_cols)
{
    const int D = C + 5 + extra_cols;
    const long long total = rows * D;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % D);
        float g = __ldg(go + e);
        if (c > C + 4) {
            if (c - C - 5 < tanh_cols) {                                         // d tanh = 1 - tanh^2 (mask coefficients)
                const float t = tanhf(__ldg(raw + e));
                g = __fmul_rn(g, __fsub_rn(1.0f, __fmul_rn(t, t)));
            }
        } else if (c > C) {
            const float s = sigmoid_acc(__ldg(raw + e));
            const float t = __fmul_rn(s, __fsub_rn(1.0f, s));                    // sigmoid'
            g = (c - C - 1 < 2) ? __fmul_rn(__fmul_rn(g, 2.0f), t)              // d(2s - 0.5)
                                : __fmul_rn(__fmul_rn(g, __fmul_rn(4.0f, s)), __fmul_rn(2.0f, t));  // d((2s)^2) = 2*(2s) * 2s'
        }
        gr[e] = g;
    }
}

// DetectionNet._bbox_to_size (modules/detection.py:175-190) on decoded rows, in place: box = (box / from) * to with
// from = [W,H,W,H], to = [W0,H0,W0,H0] read from the int64 device tensors the reference builds (:77-78).
__global__ void __launch_bounds__(256) bbox_to_size_kernel(float *pred, long long rows, int C, int D, const long long *from4,
                                                           const long long *to4)
{
    const long long total = rows * 4;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(e & 3);
        float *p = pred + (e >> 2) * D + C + 1 + q;
        *p = __fmul_rn(__fdiv_rn(*p, (float)from4[q]), (float)to4[q]);
    }
}

// Rows of DetectionNet.forward(x, inference=True) (modules/detection.py:69-91) for SELECTED candidates only: out[r] =
// [obj logit, class logits, x, y, w, h (+ trailing columns)] of flat candidate idx[r] = b*N + i, the box decoded (and
// rescaled to the original frame) by the same arithmetic as decode_xyxy -- before the allowance and the xyxy step,
// which the caller's post-processing applies itself (inference_det.py:73-76).  One warp per row.
__global__ void __launch_bounds__(256) decode_rows_kernel(DetectK k, const long long *idx, long long n, float *out)
{
    const long long r = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const long long f = idx[r];
    const int b = (int)(f / k.N);
    int i = (int)(f - (long long)b * k.N), si = 0;
    if (i >= k.sc[1].img_off) si = (i >= k.sc[2].img_off) ? 2 : 1;
    const ScaleDesc &s = k.sc[si];
    i -= s.img_off;
    const float *row = s.raw + ((long long)b * s.img_stride + i) * k.D;
    float *o = out + r * k.D;
    for (int c = lane; c < k.D; c += 32) {
        float v = row[c];
        const int q = c - k.C - 1;
        if (q >= 0 && q < 4) {
            const int a = i % k.na, cell = i / k.na;
            const int x = cell % s.nx, y = cell / s.nx;
            const float sg = __fmul_rn(sigmoid_acc(v), 2.0f);
            if (q < 2) v = __fmul_rn(__fadd_rn(__fsub_rn(sg, 0.5f), q == 0 ? (float)x : (float)y), q == 0 ? s.s0 : s.s1);
            else v = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(sg, sg), q == 2 ? s.aw[a] : s.ah[a]), q == 2 ? s.fnx : s.fny), q == 2 ? s.s0 : s.s1);
            if (k.rescale) { const bool isx = (q == 0 || q == 2); v = __fmul_rn(__fdiv_rn(v, isx ? k.fW : k.fH), isx ? k.fW0 : k.fH0); }
        }
        o[c] = v;
    }
}

}  // namespace bg
