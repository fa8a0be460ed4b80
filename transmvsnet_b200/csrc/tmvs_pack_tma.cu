// Layout pre-pass through the TMA engine (sm_100a): NCHW source feature maps -> the kernel-native blocked
// channel-last layout [H][Wb][C4][8 px][4 ch].
//
// The register-transpose kernel (tmvs_pack.cu pack_sources_nchw4_kernel) issues one 128-bit load per channel plane
// and thread and is throttled by the LSU queue (ncu: lg_throttle 20 per issue at 3.4-4.5 TB/s).  Here one elected
// thread per CTA asks the TMA engine for an 8 KB box of [C planes][1 row][64-256 pixels] (cp.async.bulk.tensor.3d,
// completion on an mbarrier): the loads never enter the LSU queue, and the CTA only does the transposing half -- conflict-free
// shared-memory reads (a warp = 32 x-adjacent pixels of one channel group) and fully coalesced 128-bit stores
// (4 x 128 contiguous bytes per warp).  Applies to contiguous NCHW maps with W % 4 == 0 and C % 4 == 0 (what
// FeatureNet / FMT produce, models/module.py:399-422); anything else takes the kernels of tmvs_pack.cu.
#include <cuda.h>

#include "tmvs_common.cuh"

#ifndef TMVS_PACK_PX16
#define TMVS_PACK_PX16 128
#endif
#ifndef TMVS_PACK_PX8
#define TMVS_PACK_PX8 256
#endif

namespace {

// pixels per box: 8 KB boxes (C x PX x 4 bytes), at most 256 pixels (the TMA box limit per dimension)
template <int C4T> struct BoxPx { static constexpr int value = C4T >= 8 ? 64 : (C4T == 4 ? TMVS_PACK_PX16 : TMVS_PACK_PX8); };
constexpr int kThreads = 128;

struct PackMaps {
    CUtensorMap m[TMVS_MAX_SRC_VIEWS];
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int C4T>
__global__ void __launch_bounds__(kThreads)
pack_sources_tma_kernel(const __grid_constant__ PackMaps maps, float4 *__restrict__ packed, int B, int H, int W)
{
    constexpr int C = 4 * C4T;
    constexpr int kPX = BoxPx<C4T>::value;
    __shared__ __align__(128) float tile[C][kPX];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kPX, y = blockIdx.y, vb = blockIdx.z;
    const int view = vb / B, b = vb - view * B;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)),
                     "r"((unsigned)(C * kPX * sizeof(float))) : "memory");
        // box {kPX pixels, 1 row, C planes} at (x0, y, b * C); pixels beyond W are zero-filled by the engine
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(&tile[0][0])), "l"(&maps.m[view]), "r"(smem_u32(&bar)), "r"(x0), "r"(y), "r"(b * C) : "memory");
    }
    // bounded wait: a copy that never completes (a descriptor bug) must trap, not hang the GPU
    {
        bool done = false;
#pragma unroll 1
        for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
            unsigned ok;
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            done = ok != 0;
        }
        if (!done) __trap();
    }
    const TmvsPacked pk = tmvs_packed_layout(C4T, H, W);
    const int wb = (W + 7) >> 3;
    float4 *row = packed + (size_t)vb * pk.slice + (size_t)y * pk.row;
#pragma unroll
    for (int k = 0; k < (kPX * C4T) / kThreads; ++k) {
        const int idx = k * kThreads + tid;
        const int px = idx & (kPX - 1), g = idx / kPX;          // a warp = 32 x-adjacent pixels of one channel group
        const int blk = (x0 + px) >> 3;
        if (blk < wb) {
            const float4 v = make_float4(tile[4 * g][px], tile[4 * g + 1][px], tile[4 * g + 2][px], tile[4 * g + 3][px]);
            row[blk * pk.c4x8 + g * 8 + (px & 7)] = v;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn pack_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace

// TMVS_E_UNSUPPORTED when the inputs are not contiguous NCHW with W % 4 == 0, C in {8, 16, 32, 64}
int tmvs_pack_sources_tma(const float *const *src, int n_src, int64_t sB, int64_t sC, int64_t sH, int64_t sW,
                          float *packed, int B, int C, int H, int W, cudaStream_t st)
{
    if (sW != 1 || sH != W || sC != (int64_t)H * W || sB != (int64_t)C * H * W || (W & 3) != 0) return TMVS_E_UNSUPPORTED;
    if (C != 8 && C != 16 && C != 32 && C != 64) return TMVS_E_UNSUPPORTED;
    if ((size_t)B * C > 0x7fffffffu || H > 65535 || (size_t)n_src * B > 65535) return TMVS_E_UNSUPPORTED;
    EncodeTiledFn encode = pack_encode_fn();
    if (!encode) return TMVS_E_UNSUPPORTED;
    PackMaps maps;
    for (int i = 0; i < n_src; ++i) {
        if (((uintptr_t)src[i] & 15) != 0) return TMVS_E_UNSUPPORTED;
        const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * C};
        const cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
        const int px = C == 8 ? BoxPx<2>::value : (C == 16 ? BoxPx<4>::value : 64);
        const cuuint32_t box[3] = {(cuuint32_t)px, 1u, (cuuint32_t)C};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult r = encode(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)src[i], gdim, gstr, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return TMVS_E_UNSUPPORTED;
    }
    const int kPX = C == 8 ? BoxPx<2>::value : (C == 16 ? BoxPx<4>::value : 64);
    dim3 grid((W + kPX - 1) / kPX, H, n_src * B);
    float4 *out = (float4 *)packed;
    switch (C) {
    case 8: pack_sources_tma_kernel<2><<<grid, kThreads, 0, st>>>(maps, out, B, H, W); break;
    case 16: pack_sources_tma_kernel<4><<<grid, kThreads, 0, st>>>(maps, out, B, H, W); break;
    case 32: pack_sources_tma_kernel<8><<<grid, kThreads, 0, st>>>(maps, out, B, H, W); break;
    default: pack_sources_tma_kernel<16><<<grid, kThreads, 0, st>>>(maps, out, B, H, W); break;
    }
    return tmvs_launch_status();
}
