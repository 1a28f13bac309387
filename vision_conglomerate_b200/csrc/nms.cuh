// Segmented greedy-NMS engine shared by bg_batched_nms (generic groups) and bg_detect (one segment
// per image).  Kernels are in nms_kernels.cuh (single translation unit: boxgeom.cu).
#pragma once
#include "common.cuh"

namespace bg {

struct SegHdr {  // lives at the start of the workspace; (re)initialised by the first kernel of every call
    int S;       // number of segments
    int status;  // BG_STATUS_* bits
    unsigned item_ctr;     // work-item scheduler of the pair / mask kernel
    unsigned reduce_done;  // "last CTA" ticket of the reduce kernel
    long long gmin, gmax;  // batched_nms: range of idxs
    long long total_out;
    unsigned item_ctr2;    // work-item scheduler of the dense mask kernel when it runs as the grid mode's fallback
    int overflow;          // grid mode: the edge list of some segment ran out of room
    int dense_fits;        // the dense bit matrix fits the mask workspace
    int pad0;
    long long pad[1];
};

struct SegGrid {  // uniform grid over the valid box centres of one segment
    float mnx, mny, invx, invy, pad;
    int G;       // cells per axis
    int nvalid;  // boxes with a positive finite extent (only those are in cell order)
    int pad2;
};

struct SegNms {
    SegHdr *hdr;
    int *seg_count;         // [S_max]   candidates per segment
    long long *seg_off;     // [S_max+1] element offset of the segment in keys / sorted_* / b* / emit_* (room for next_pow2(count))
    int *tile_prefix;       // [S_max+1] exclusive prefix of ceil(count/64); also the word offset into keepbits
    long long *mask_off;    // [S_max+1] exclusive prefix of count*ceil(count/64) (u64 words of the dense bit matrix)
    long long *edge_off;    // [S_max+1] grid mode: start of the segment's region of the edge list (u64 entries of mask)
    int *item_prefix;       // [S_max+1] exclusive prefix of the pair-kernel work items
    int *emit_count;        // [S_max]   rows emitted per segment (after NMS and the class filter)
    long long *out_prefix;  // [S_max+1]
    u64 *keys;              // (score desc, id asc) keys, sorted in place per segment
    float4 *sorted_box;     // boxes gathered into score order
    float *sorted_area;     // (x2-x1)*(y2-y1), pre-rounded like torchvision's CPU kernel
    u64 *bkeys;             // grid mode: (grid cell << 32 | score-order position) per segment in "cell order"
    float4 *bbox;           // boxes in cell order
    float *barea;
    u32 *bwh;               // (w, h) in cell order, truncated to a bf16 pair: the pair kernel's prefilter
    long long *cell_off;    // [S_max+1] offset of the segment's cell table in cell_start (G*G + 1 entries)
    int *cell_start;        // first cell-order index of every grid cell
    SegGrid *grid;          // [S_max]
    unsigned long long *edge_count;  // [S_max] overlap edges found (may exceed the segment's region: hdr->overflow)
    unsigned char *gstate;  // [2*elems] resolve state of segments too large for shared memory
    u64 *keepbits;          // kept bitmap in score order, word offset tile_prefix[s]
    u32 *ew32;              // emitted bitmap (32-bit words), offset 2*tile_prefix[s]
    u32 *rank32;            // exclusive rank of each 32-bit word of ew32
    u64 *mask;              // dense mode: suppression bit matrix, per segment column-tile-major: word(ct,row) at
                            // mask_off[s] + ct*K + row; grid mode: overlap edges (earlier position << 32 | later position)
    long long mask_words;   // capacity of mask
    u32 *emit_pos;          // score-order position of each emitted row (compact per segment, at seg_off)
    u64 *emit_key;          // its key
    const float4 *boxes;    // source boxes; box of (segment s, id) is boxes[s*box_seg_stride + id]
    long long box_seg_stride;
    const int *cls;         // optional class of (s, id) at the same index, for the tracked-class filter
    int n_tracked;
    int tracked[BG_MAX_TRACKED];
    IouThr thr;
    int sparse;             // 1: grid mode (spatially pruned pair tests -> edge list -> rounds); 0: dense bit matrix
    float reach;            // grid mode: (1-t)/t plus margin -- centre distance bound per unit of the smaller extent
};

}  // namespace bg
