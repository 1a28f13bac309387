// Shared device/host helpers for libboxgeom (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

#include "../../include/boxgeom.h"

namespace bg {

extern std::atomic<unsigned long long> g_launches;  // host-side count of kernels enqueued by this library

#define BG_LAUNCH_CHECK()                                   \
    do {                                                    \
        ++::bg::g_launches;                                 \
        if (cudaPeekAtLastError() != cudaSuccess) {         \
            (void)cudaGetLastError();                       \
            return BG_ERR_LAUNCH;                           \
        }                                                   \
    } while (0)

typedef unsigned long long u64;
typedef unsigned int u32;

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline u32 next_pow2(u32 x)
{
    if (x <= 1) return 1;
    --x;
    x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16;
    return x + 1;
}

// fp32 -> u32 whose unsigned order equals the float order (-0 canonicalised to +0).
__device__ __forceinline__ u32 orderable(float f)
{
    if (f == 0.0f) f = 0.0f;
    u32 u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(u32 u)
{
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}
// sort key: ascending u64 order == (score descending, id ascending)
__device__ __forceinline__ u64 make_key(float score, u32 id) { return ((u64)(~orderable(score)) << 32) | id; }
__device__ __forceinline__ float key_score(u64 k) { return from_orderable(~(u32)(k >> 32)); }
__device__ __forceinline__ u32 key_id(u64 k) { return (u32)k; }

// sigmoid with IEEE divide and the accurate expf (parity tolerance is rtol 1e-5 against ATen's CPU sigmoid)
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Threshold constants for the exact IoU test (see nms.cu: iou_suppresses).
struct IouThr {
    float tdn;  // largest float <= the double threshold: (double)q > thr  <=>  q > tdn for every float q
    float lo, hi;  // guard band around tdn for the division-free fast path
    int fast_ok;   // guard band valid (tdn is a positive normal number)
    int zero_suppresses;  // 0.0 > thr: non-overlapping pairs suppress too (negative thresholds)
};
IouThr make_iou_thr(double thr);

// Ascending bitonic sort of s[0..P), P = E * NT keys (callers pad with ~0), by NT threads (1024, or 512 in the lean
// per-image NMS).  Thread t holds the E consecutive keys t*E..t*E+E-1 in registers: strides below E are register
// compare-exchanges, strides below 32*E are warp shuffles with lane ^ (j/E), and only the larger strides go through
// shared memory (staged key-index-major, s[e*NT + t], so the exchange reads are conflict-free).
template <int E, int NT = 1024>
__device__ __forceinline__ void sort_reg_1024(u64 *s, int k_start = 2)  // k_start = P: only the last merge (input bitonic)
{
    const int t = threadIdx.x;
    constexpr int P = E * NT;
    u64 v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = s[t * E + e];
    __syncthreads();
    for (int k = k_start; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32 * E) {
#pragma unroll
                for (int e = 0; e < E; ++e) s[e * NT + t] = v[e];
                __syncthreads();
                const int pt = t ^ (j / E);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = t * E + e;
                    const u64 o = s[e * NT + pt];
                    const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                    v[e] = keep_min ? (o < v[e] ? o : v[e]) : (o > v[e] ? o : v[e]);
                }
                __syncthreads();
            } else if (j >= E) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = t * E + e;
                    const u64 o = __shfl_xor_sync(0xffffffffu, v[e], j / E);
                    const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                    v[e] = keep_min ? (o < v[e] ? o : v[e]) : (o > v[e] ? o : v[e]);
                }
            } else {
                // register compare-exchange; j is 1 (E >= 2), 2 (E >= 4) or 4 (E == 8): static indices keep v[] in registers
#pragma unroll
                for (int jj = 1; jj < E; jj <<= 1) {
                    if (jj != j) continue;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        if ((e & jj) == 0) {
                            const int i = t * E + e;
                            const bool up = (i & k) == 0;
                            const u64 a = v[e], c = v[e | jj];
                            const bool sw = (a > c) == up;
                            v[e] = sw ? c : a;
                            v[e | jj] = sw ? a : c;
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) s[t * E + e] = v[e];
    __syncthreads();
}


}  // namespace bg
