// Microbenchmark: read-only streaming of a large buffer with the access pattern of decode_filter_kernel
// (persistent CTAs, ring of TMA bulk copies) and with plain vector loads, to find the read-only HBM ceiling
// the decode kernel should be judged against.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/stream_read scripts/micro/stream_read.cu && /tmp/stream_read
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    asm volatile("{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WD;\nbra WL;\nWD:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// TMA ring: STAGES buffers of tile_bytes; `work` = dummy smem reads per thread per tile (emulates phase 1)
template <int STAGES>
__global__ void __launch_bounds__(128) tma_ring(const char *src, long long ntiles, int tile_bytes, int work, float *sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) u64 bar[STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < STAGES; ++s) {
            long long t = blockIdx.x + (long long)s * gridDim.x;
            if (t < ntiles) { mbar_expect_tx(&bar[s], tile_bytes); bulk_g2s(smem + (size_t)s * tile_bytes, src + t * tile_bytes, tile_bytes, &bar[s]); }
        }
    }
    __syncthreads();
    u32 phases = 0; int it = 0; float acc = 0.f;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int st = it % STAGES;
        mbar_wait(&bar[st], (phases >> st) & 1u); phases ^= 1u << st;
        const float *tile = reinterpret_cast<const float *>(smem + (size_t)st * tile_bytes);
        for (int i = 0; i < work; ++i) acc = fmaxf(acc, tile[(tid * 85 + i) % (tile_bytes / 4)]);
        __syncthreads();
        if (tid == 0) {
            long long tn = t + (long long)STAGES * gridDim.x;
            if (tn < ntiles) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); mbar_expect_tx(&bar[st], tile_bytes); bulk_g2s(smem + (size_t)st * tile_bytes, src + tn * tile_bytes, tile_bytes, &bar[st]); }
        }
    }
    if (acc == 12345.f) *sink = acc;
}

__global__ void __launch_bounds__(256) ldg_stream(const float4 *src, long long n4, float *sink)
{
    float acc = 0.f;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
        acc += a.x + b.y + c.z + d.w;
    }
    for (; i < n4; i += stride) acc += __ldcs(src + i).x;
    if (acc == 12345.f) *sink = acc;
}

template <typename F> float time_ms(F f, int reps)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main()
{
    const long long bytes = 548352000LL;  // one config-2 batch
    char *buf; float *sink;
    cudaMalloc(&buf, bytes + (1 << 20)); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d, buffer %.1f MB\n", sms, bytes / 1e6);
    {
        float ms = time_ms([&] { ldg_stream<<<sms * 8, 256>>>((const float4 *)buf, bytes / 16, sink); }, 20);
        printf("ldg float4 x4 unrolled, %d CTAs x256: %.1f us  %.0f GB/s\n", sms * 8, ms * 1e3, bytes / ms / 1e6);
        ms = time_ms([&] { ldg_stream<<<sms * 16, 256>>>((const float4 *)buf, bytes / 16, sink); }, 20);
        printf("ldg float4 x4 unrolled, %d CTAs x256: %.1f us  %.0f GB/s\n", sms * 16, ms * 1e3, bytes / ms / 1e6);
    }
    const int tiles_bytes[] = {43520, 32640, 21760, 65280};
    for (int tb : tiles_bytes) {
        const long long ntiles = bytes / tb;
        for (int stages = 2; stages <= 4; ++stages) {
            for (int cps = 1; cps <= 4; ++cps) {
                const size_t smem = (size_t)stages * tb;
                if (smem * cps > 220 * 1024 || smem > 227 * 1024) continue;
                for (int work : {0, 80}) {
                    float ms = 0;
                    auto run = [&](auto kern) {
                        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                        ms = time_ms([&] { kern<<<sms * cps, 128, smem>>>(buf, ntiles, tb, work, sink); }, 20);
                    };
                    if (stages == 2) run(tma_ring<2>); else if (stages == 3) run(tma_ring<3>); else run(tma_ring<4>);
                    cudaError_t e = cudaGetLastError();
                    printf("tma tile %5d B, %d stages, %d CTA/SM, work %2d: %.1f us  %.0f GB/s %s\n", tb, stages, cps, work, ms * 1e3,
                           (double)ntiles * tb / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
            }
        }
    }
    return 0;
}
