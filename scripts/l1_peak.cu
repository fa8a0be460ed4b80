// Micro-benchmark of the SM's load data path on B200 (sm_100a): what the fused cost-volume kernel is bound by.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/l1_peak scripts/l1_peak.cu && gpurun_out/l1_peak
//
// (the per-iteration stride is a kernel argument so the compiler cannot merge repeated addresses)
// Every test streams L1-resident data into registers with all 148 SMs busy (8 CTAs x 256 threads per SM) and reports
// bytes per clock per SM (CUDA-event time x the SM clock the driver reports).  Tests:
//   ldg128_aligned    warp reads 512 contiguous bytes, 128-byte aligned     (4 lines per request)
//   ldg128_shift16    the same, shifted by one 16-byte pixel                 (5 lines per request: the usual tap)
//   lds128            the same bytes from shared memory
//   shfl              4 x SHFL.IDX per iteration (moving a float4 to the neighbouring lane)
//   ldg128+shfl       one aligned LDG.128 and 4 SHFL per iteration: do the two share a pipe (times add) or overlap?
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kIters = 4096;
constexpr int kThreads = 256;
constexpr int kWordsPerCta = 1024;   // float4 words each CTA cycles over: 16 KB, L1 resident

template <int SHIFT>
__global__ void __launch_bounds__(kThreads) ldg128_kernel(const float4 *__restrict__ buf, float *out, int stride)
{
    const float4 *base = buf + (size_t)blockIdx.x * (kWordsPerCta + 8) + SHIFT;
    const int lane_off = threadIdx.x;                       // warp w reads words [32w, 32w+32): 512 contiguous bytes
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int it = 0; it < kIters; ++it) {
        const float4 v = __ldg(base + ((lane_off + it * stride) & (kWordsPerCta - 1)));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[blockIdx.x * kThreads + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

__global__ void __launch_bounds__(kThreads) lds128_kernel(const float4 *__restrict__ buf, float *out, int stride)
{
    __shared__ float4 sm[kWordsPerCta];
    for (int i = threadIdx.x; i < kWordsPerCta; i += kThreads) sm[i] = buf[i];
    __syncthreads();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int it = 0; it < kIters; ++it) {
        const float4 v = sm[(threadIdx.x + it * stride) & (kWordsPerCta - 1)];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[blockIdx.x * kThreads + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <bool WITH_LDG>
__global__ void __launch_bounds__(kThreads) shfl_kernel(const float4 *__restrict__ buf, float *out, int stride)
{
    const float4 *base = buf + (size_t)blockIdx.x * (kWordsPerCta + 8);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 cur = make_float4((float)threadIdx.x, 1.f, 2.f, 3.f);
    const int src = (threadIdx.x + 1) & 31;
#pragma unroll 8
    for (int it = 0; it < kIters; ++it) {
        if (WITH_LDG) {
            const float4 v = __ldg(base + ((threadIdx.x + it * stride) & (kWordsPerCta - 1)));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        cur.x = __shfl_sync(0xffffffffu, cur.x, src);
        cur.y = __shfl_sync(0xffffffffu, cur.y, src);
        cur.z = __shfl_sync(0xffffffffu, cur.z, src);
        cur.w = __shfl_sync(0xffffffffu, cur.w, src);
    }
    out[blockIdx.x * kThreads + threadIdx.x] = acc.x + acc.y + acc.z + acc.w + cur.x + cur.y + cur.z + cur.w;
}

// LDG.128 with an arbitrary per-lane word offset (a tap pattern): cycles per warp request for each pattern
__global__ void __launch_bounds__(kThreads) pattern_kernel(const float4 *__restrict__ buf, float *out, const int *lane_off, int stride)
{
    const float4 *base = buf + (size_t)blockIdx.x * (kWordsPerCta + 8) * 2;
    const int off = lane_off[threadIdx.x & 31] + (threadIdx.x >> 5) * 64;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int it = 0; it < kIters; ++it) {
        const float4 v = __ldg(base + off + ((it * stride) & 255));   // stride = multiple of 8 words: keeps the line phase; 20 KB per CTA stays L1 resident
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[blockIdx.x * kThreads + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <typename F>
static float time_ms(F launch)
{
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CHECK(cudaEventRecord(e0));
    for (int i = 0; i < 10; ++i) launch();
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / 10.f;
}

int main()
{
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    int clock_khz = 0;
    CHECK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    const int sms = prop.multiProcessorCount;
    const int ctas = sms * 8;
    float4 *buf;
    float *out;
    CHECK(cudaMalloc(&buf, (size_t)ctas * (kWordsPerCta + 8) * sizeof(float4)));
    CHECK(cudaMemset(buf, 0, (size_t)ctas * (kWordsPerCta + 8) * sizeof(float4)));
    CHECK(cudaMalloc(&out, (size_t)ctas * kThreads * sizeof(float)));
    const double bytes = (double)ctas * kThreads * kIters * 16.0;
    const double clk_hz = clock_khz * 1e3;
    struct { const char *name; float ms; double bytes; } rows[5];
    rows[0] = {"ldg128_aligned", time_ms([&] { ldg128_kernel<0><<<ctas, kThreads>>>(buf, out, kThreads); }), bytes};
    rows[1] = {"ldg128_shift16", time_ms([&] { ldg128_kernel<1><<<ctas, kThreads>>>(buf, out, kThreads); }), bytes};
    rows[2] = {"lds128", time_ms([&] { lds128_kernel<<<ctas, kThreads>>>(buf, out, kThreads); }), bytes};
    rows[3] = {"shfl_x4", time_ms([&] { shfl_kernel<false><<<ctas, kThreads>>>(buf, out, kThreads); }), bytes};
    rows[4] = {"ldg128+shfl_x4", time_ms([&] { shfl_kernel<true><<<ctas, kThreads>>>(buf, out, kThreads); }), bytes};
    CHECK(cudaDeviceSynchronize());
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %.0f, \"tests\": [", prop.name, sms, clock_khz / 1e3);
    for (int i = 0; i < 5; ++i) {
        const double s = rows[i].ms * 1e-3;
        printf("%s{\"name\": \"%s\", \"ms\": %.4f, \"GBps\": %.1f, \"bytes_per_clk_per_sm\": %.1f}", i ? ", " : "",
               rows[i].name, rows[i].ms, rows[i].bytes / s / 1e9, rows[i].bytes / s / clk_hz / sms);
    }
    printf("], \"patterns\": [");
    // tap patterns: word offset of each lane (1 word = one 16-byte pixel; 8 words = one 128-byte line)
    CHECK(cudaFree(buf));
    CHECK(cudaMalloc(&buf, (size_t)ctas * (kWordsPerCta + 8) * 2 * sizeof(float4)));
    CHECK(cudaMemset(buf, 0, (size_t)ctas * (kWordsPerCta + 8) * 2 * sizeof(float4)));
    int *d_off;
    CHECK(cudaMalloc(&d_off, 32 * sizeof(int)));
    struct Pat { const char *name; int off[32]; } pats[24];
    int np = 0;
    auto add = [&](const char *name, auto f) { pats[np].name = name; for (int i = 0; i < 32; ++i) pats[np].off[i] = f(i); ++np; };
    add("contig_aligned", [](int i) { return i; });
    add("contig_shift1px", [](int i) { return i + 1; });
    add("contig_shift4px", [](int i) { return i + 4; });
    add("scale1.1_shift3", [](int i) { return 3 + (int)(i * 1.1f); });
    add("scale0.9_shift3", [](int i) { return 3 + (int)(i * 0.9f); });
    add("blocked_c4x4_shift3", [](int i) { int x = i + 3; return (x >> 3) * 32 + (x & 7); });          // lines 512 B apart
    add("blocked_c4x4_scale1.1", [](int i) { int x = 3 + (int)(i * 1.1f); return (x >> 3) * 32 + (x & 7); });
    add("two_rows_split16", [](int i) { return (i < 16 ? 0 : 512) + i + 3; });                        // row change mid-warp
    add("two_rows_split5", [](int i) { return (i < 5 ? 0 : 512) + i + 3; });
    add("same_word_bcast", [](int i) { return 5; });
    add("stride2px", [](int i) { return 2 * i; });
    add("pairs_dup", [](int i) { return i / 2 + 3; });
    // candidate mapping: lane 2j = NW tap of pixel j, lane 2j+1 = its NE tap (16 pixels per request), source spacing s
    add("nw_ne_pairs_s1.0", [](int i) { return 3 + (int)((i / 2) * 1.0f) + (i & 1); });
    add("nw_ne_pairs_s1.1", [](int i) { return 3 + (int)((i / 2) * 1.1f) + (i & 1); });
    add("nw_ne_pairs_s1.5", [](int i) { return 3 + (int)((i / 2) * 1.5f) + (i & 1); });
    add("nw_ne_pairs_s0.6", [](int i) { return 3 + (int)((i / 2) * 0.6f) + (i & 1); });
    add("nw_ne_pairs_s1.1_rowsplit", [](int i) { return (i < 12 ? 0 : 512) + 3 + (int)((i / 2) * 1.1f) + (i & 1); });
    add("scale1.5_shift3", [](int i) { return 3 + (int)(i * 1.5f); });
    add("scale1.03_shift3", [](int i) { return 3 + (int)(i * 1.03f); });
    // candidate mapping: 4 pixels x 2 depth planes per quarter (plane slide 0.8 px)
    add("4px_x_2planes_s1.1", [](int i) { return 3 + (int)(((i >> 3) * 4 + (i & 3)) * 1.1f + ((i >> 2) & 1) * 0.8f); });
    for (int k = 0; k < np; ++k) {
        CHECK(cudaMemcpy(d_off, pats[k].off, sizeof(pats[k].off), cudaMemcpyHostToDevice));
        const float ms = time_ms([&] { pattern_kernel<<<ctas, kThreads>>>(buf, out, d_off, 8); });
        const double req = (double)ctas * (kThreads / 32) * kIters / sms;      // warp requests per SM
        printf("%s{\"name\": \"%s\", \"ms\": %.4f, \"clk_per_request\": %.2f}", k ? ", " : "", pats[k].name, ms,
               ms * 1e-3 * clk_hz / req);
    }
    printf("]}\n");
    return 0;
}
