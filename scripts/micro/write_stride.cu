// Microbenchmarks behind the training-side roofline: (1) write-only streaming of a grad_preds-sized buffer,
// (2) the objectness read of loss_dense_kernel: one float every 340 bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/write_stride scripts/micro/write_stride.cu && /tmp/write_stride
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) fill4(float4 *dst, long long n4)
{
    const float4 v = make_float4(0.f, 1.f, 0.f, 0.f);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        if (MODE == 0) dst[i] = v;
        else if (MODE == 1) __stcs(dst + i, v);
        else if (MODE == 2) __stwt(dst + i, v);
        else __stcg(dst + i, v);
    }
}
// warp per 10,880-byte chunk (the backward kernel's pattern)
__global__ void __launch_bounds__(256) fill_chunks(float4 *dst, long long nchunks)
{
    const int lane = threadIdx.x & 31;
    long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const float4 v = make_float4(0.f, 1.f, 0.f, 0.f);
    for (; w < nchunks; w += nw) {
        float4 *p = dst + w * 680;
        for (int f = lane; f < 680; f += 32) p[f] = v;
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) strided_read(const float *src, long long n, int stride_f, float *sink)
{
    float acc = 0.f;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += st) {
        const float *p = src + i * stride_f;
        float v;
        if (MODE == 0) v = __ldg(p);
        else if (MODE == 1) v = __ldcs(p);
        else if (MODE == 2) asm volatile("ld.global.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
        else if (MODE == 3) asm volatile("ld.global.L1::no_allocate.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
        else asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
        acc += v;
    }
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
    const long long cells = 256LL * 25200, D = 85;
    const long long bytes = cells * D * 4;  // 2.19 GB
    char *buf; float *sink;
    cudaMalloc(&buf, bytes + (1 << 20)); cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, bytes);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("buffer %.1f MB\n", bytes / 1e6);
    float ms;
    ms = time_ms([&] { cudaMemsetAsync(buf, 0, bytes); }, 10);
    printf("cudaMemsetAsync: %.1f us %.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    for (int g : {8, 16, 32}) {
        ms = time_ms([&] { fill4<0><<<sms * g, 256>>>((float4 *)buf, bytes / 16); }, 10);
        printf("fill float4 plain, %d CTAs/SM: %.1f us %.0f GB/s\n", g, ms * 1e3, bytes / ms / 1e6);
    }
    ms = time_ms([&] { fill4<1><<<sms * 16, 256>>>((float4 *)buf, bytes / 16); }, 10);
    printf("fill float4 __stcs: %.1f us %.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    ms = time_ms([&] { fill4<2><<<sms * 16, 256>>>((float4 *)buf, bytes / 16); }, 10);
    printf("fill float4 __stwt: %.1f us %.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    ms = time_ms([&] { fill4<3><<<sms * 16, 256>>>((float4 *)buf, bytes / 16); }, 10);
    printf("fill float4 __stcg: %.1f us %.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    for (int g : {4, 8, 16}) {
        ms = time_ms([&] { fill_chunks<<<sms * g, 256>>>((float4 *)buf, cells / 32); }, 10);
        printf("fill warp-per-chunk, %d CTAs/SM: %.1f us %.0f GB/s\n", g, ms * 1e3, bytes / ms / 1e6);
    }
    const char *names[] = {"__ldg", "__ldcs", "L2::64B", "no_alloc+L2::64B", "volatile"};
    for (int mode = 0; mode < 5; ++mode) {
        auto run = [&](auto kern) { ms = time_ms([&] { kern<<<sms * 16, 256>>>((const float *)buf, cells, (int)D, sink); }, 10); };
        if (mode == 0) run(strided_read<0>); else if (mode == 1) run(strided_read<1>); else if (mode == 2) run(strided_read<2>);
        else if (mode == 3) run(strided_read<3>); else run(strided_read<4>);
        printf("strided read (1 float / 340 B, %lld cells) %-18s: %.1f us  -> %.1f B of DRAM time per cell at 6.5 TB/s\n", cells, names[mode], ms * 1e3,
               ms * 1e-3 * 6.5e12 / cells);
    }
    return 0;
}
