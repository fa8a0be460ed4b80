// Fused cost-volume forward with TMA-staged source tiles (sm_100a).
//
// Same arithmetic as costvol_fwd_kernel (tmvs_costvol.cu) -- replaces models/TransMVSNet.py:71-93 -- but the
// bilinear taps are read from SHARED MEMORY: adjacent depth planes map to neighbouring source pixels, so the
// footprint of a 32x8 reference tile over an 8-plane chunk is a small source window.  Per (CTA, view):
//   pass 1  every thread computes its 8 sample positions (the reference's arithmetic, tmvs_coords), parks them in
//           shared memory, and the CTA reduces the exact bounding boxes of the in-bounds taps of the next
//           8 / 4 / 2 / 1 planes (REDUX + one barrier) and takes the longest span whose box fits a window;
//   load    one elected thread issues ONE TMA tensor copy (cp.async.bulk.tensor.4d, SASS UTMALDG) of the box
//           [rows][8-px blocks][C4][8 px x 4 ch] from the packed layout into shared memory and the CTA waits
//           on the mbarrier the copy completes on.  The box shape is picked from a small menu of tensor maps
//           (wide, square-ish, tall) so horizontal, diagonal and vertical epipolar geometry all fit; a window
//           that fits none (extreme geometry) takes the global-memory path of the L1 kernel for that view;
//   pass 2  4*C4 LDS.128 per voxel-view at immediate offsets -- bank-exact (8 x-adjacent pixels = 128 B), i.e.
//           4 wavefronts per request instead of the ~6 lines an unaligned global gather touches -- then the
//           same dot-first correlation and aggregation as the L1 kernel.
// Two to four CTAs are resident per SM, so one CTA's copy is hidden behind the others' math.
#include <cuda.h>

#include "tmvs_common.cuh"

namespace {

constexpr int kTileX = 32, kTileY = 8, kThreads = kTileX * kTileY;
constexpr int kDC = 8;
constexpr int kCapBlocks = 80;       // 8-pixel blocks (all channel groups) the shared-memory window holds
constexpr int kMenu = 5;
constexpr int kSpans = 4;            // plane spans tried per load: all remaining, 4, 2, 1
constexpr int kEmpty = 0x7fffffff;
// window shapes (8-px blocks wide x rows high), tried in order: typical first, then wide / tall variants
__constant__ int c_menu_bw[kMenu] = {7, 8, 10, 6, 5};
__constant__ int c_menu_bh[kMenu] = {10, 10, 8, 13, 16};
const int h_menu_bw[kMenu] = {7, 8, 10, 6, 5};
const int h_menu_bh[kMenu] = {10, 10, 8, 13, 16};

struct TmvsTmaMaps {
    CUtensorMap m[kMenu];
};

template <int C4T> struct TmaMinBlocks { static constexpr int value = C4T >= 8 ? 2 : (C4T >= 4 ? 3 : 4); };

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    // bounded: a copy that never completes (a descriptor/byte-count bug) must trap, not hang the GPU
#pragma unroll 1
    for (int spin = 0; spin < (1 << 22); ++spin) {
        unsigned done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, unsigned long long *bar, int c0, int c1,
                                            int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int C4T, bool PER_PIXEL, bool VIEWS, bool AGG>
__global__ void __launch_bounds__(kThreads, TmaMinBlocks<C4T>::value)
costvol_tma_kernel(const float *__restrict__ ref, int64_t rB, int64_t rC, int64_t rH, int64_t rW,
                   const float4 *__restrict__ packed, const float *__restrict__ depth, const float *__restrict__ vw,
                   float *__restrict__ sim_views, float *__restrict__ agg, int b_total, int b_first, int b_chunk, int C,
                   int D, int H, int W, int n_src, int n_dchunks, const __grid_constant__ TmvsGeom geom,
                   const __grid_constant__ TmvsTmaMaps maps)
{
    constexpr int kTileWords = kCapBlocks * C4T * 8;   // float4 words
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float4 *tile = reinterpret_cast<float4 *>(smem_raw);
    float *crd_s = reinterpret_cast<float *>(smem_raw + (size_t)kTileWords * 16);          // [2][kDC][kThreads]
    int *red_all = reinterpret_cast<int *>(crd_s + 2 * kDC * kThreads);                    // [2][kSpans*4][kTileY]
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(red_all + 2 * kSpans * 4 * kTileY);
    float *acc_s = reinterpret_cast<float *>(bar + 2);                                     // [kDC][kThreads] (AGG)

    const int tid = threadIdx.y * kTileX + threadIdx.x;
    const int chunk = blockIdx.x % n_dchunks;
    const int x = (blockIdx.x / n_dchunks) * kTileX + threadIdx.x;
    const int y = blockIdx.y * kTileY + threadIdx.y;
    const bool active = x < W && y < H;
    const int bl = blockIdx.z;
    const int d0 = chunk * kDC;
    const int nd = min(kDC, D - d0);
    const int b = b_first + bl;
    const int HW = H * W;
    const int pix = min(y, H - 1) * W + min(x, W - 1);

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }

    float4 r[C4T];
    {
        const float *rp = ref + b * rB + min(y, H - 1) * rH + min(x, W - 1) * rW;
#pragma unroll
        for (int g = 0; g < C4T; ++g) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 4 * g + j;
                v[j] = (c < C) ? __ldg(rp + c * rC) : 0.0f;
            }
            r[g] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    const float *dep_base = PER_PIXEL ? depth + ((size_t)b * D + d0) * HW + pix : depth + (size_t)b * D + d0;
    const int dep_stride = PER_PIXEL ? HW : 1;
    if (AGG) {
#pragma unroll
        for (int k = 0; k < kDC; ++k) acc_s[k * kThreads + tid] = 0.0f;
    }
    float wsum = 1e-5f;                                // TransMVSNet.py:72
    const float inv_c = 1.0f / (float)C;
    const TmvsDims dims = tmvs_dims(H, W, geom.arith);
    const TmvsPacked pk = tmvs_packed_layout(C4T, H, W);
    const float xf = (float)x, yf = (float)y;
    unsigned phase = 0, round = 0;
    __syncthreads();                                   // mbarrier initialised

    for (int i = 0; i < n_src; ++i) {
        float rt[12];
        tmvs_geom_rt(geom, i, bl, b_chunk, rt);
        const TmvsRay ray = tmvs_ray(rt, xf, yf, geom.ray_unfused);
        float wi = 0.0f;
        if (AGG) wi = __ldg(vw + ((size_t)b * n_src + i) * HW + pix);
        const float4 *img = packed + ((size_t)i * b_total + b) * pk.slice;     // global path (window too large)
        float *out_v = VIEWS ? sim_views + (((size_t)i * b_total + b) * D + d0) * HW + pix : nullptr;

        // ---- pass 1: the sample positions of this thread's planes (the reference's arithmetic), parked in smem
        if (active) {
            const float *dep_p = dep_base;
#pragma unroll 2
            for (int k = 0; k < nd; ++k, dep_p += dep_stride) {
                const float2 c = tmvs_coords(ray, rt, __ldg(dep_p), dims);
                crd_s[k * kThreads + tid] = c.x;
                crd_s[(kDC + k) * kThreads + tid] = c.y;
            }
        }

        // ---- the planes are consumed in spans: as many as fit one shared-memory window (all 8 for the usual
        //      geometry; 4, 2 or 1 when the epipolar walk is long or diagonal)
        int k0 = 0;
        while (k0 < nd) {
            const int rem = nd - k0;
            const int span_len[kSpans] = {rem, min(rem, 4), min(rem, 2), 1};
            int lo_x[kSpans], lo_y[kSpans], hi_x[kSpans], hi_y[kSpans];
#pragma unroll
            for (int sp = 0; sp < kSpans; ++sp) { lo_x[sp] = kEmpty; lo_y[sp] = kEmpty; hi_x[sp] = -1; hi_y[sp] = -1; }
            if (active) {
                for (int k = 0; k < rem; ++k) {
                    const float cx = crd_s[(k0 + k) * kThreads + tid], cy = crd_s[(kDC + k0 + k) * kThreads + tid];
                    const int x0 = (int)floorf(cx), y0 = (int)floorf(cy);
                    const bool xin = ((unsigned)x0 < (unsigned)W) | ((unsigned)(x0 + 1) < (unsigned)W);
                    const bool yin = ((unsigned)y0 < (unsigned)H) | ((unsigned)(y0 + 1) < (unsigned)H);
                    if (xin & yin) {
                        const int ax = max(x0, 0), bx = min(x0 + 1, W - 1), ay = max(y0, 0), by = min(y0 + 1, H - 1);
#pragma unroll
                        for (int sp = 0; sp < kSpans; ++sp) {
                            if (k < span_len[sp]) {
                                lo_x[sp] = min(lo_x[sp], ax); hi_x[sp] = max(hi_x[sp], bx);
                                lo_y[sp] = min(lo_y[sp], ay); hi_y[sp] = max(hi_y[sp], by);
                            }
                        }
                    }
                }
            }
            int *red = red_all + (round & 1u) * (kSpans * 4 * kTileY);   // double-buffered across rounds
            ++round;
#pragma unroll
            for (int sp = 0; sp < kSpans; ++sp) {
                const int a = __reduce_min_sync(0xffffffffu, lo_x[sp]), bq = __reduce_min_sync(0xffffffffu, lo_y[sp]);
                const int c = __reduce_max_sync(0xffffffffu, hi_x[sp]), dq = __reduce_max_sync(0xffffffffu, hi_y[sp]);
                if (threadIdx.x == 0) {
                    red[(sp * 4 + 0) * kTileY + threadIdx.y] = a; red[(sp * 4 + 1) * kTileY + threadIdx.y] = bq;
                    red[(sp * 4 + 2) * kTileY + threadIdx.y] = c; red[(sp * 4 + 3) * kTileY + threadIdx.y] = dq;
                }
            }
            __syncthreads();    // boxes visible; every thread is also done reading the previous window
            int span = 1, shape = -1, bx0 = 0, by0 = 0;
            bool empty = true;
            {
                bool chosen = false;
#pragma unroll
                for (int sp = 0; sp < kSpans; ++sp) {
                    int ax = kEmpty, ay = kEmpty, cx = -1, cy = -1;
#pragma unroll
                    for (int w = 0; w < kTileY; ++w) {
                        ax = min(ax, red[(sp * 4 + 0) * kTileY + w]); ay = min(ay, red[(sp * 4 + 1) * kTileY + w]);
                        cx = max(cx, red[(sp * 4 + 2) * kTileY + w]); cy = max(cy, red[(sp * 4 + 3) * kTileY + w]);
                    }
                    if (!chosen) {
                        const bool emp = cx < 0;           // no tap of the whole tile lands inside the source image
                        int fit = -1;
                        if (!emp) {
                            const int nbx = (cx >> 3) - (ax >> 3) + 1, nby = cy - ay + 1;
#pragma unroll
                            for (int m = kMenu - 1; m >= 0; --m)
                                if (nbx <= c_menu_bw[m] && nby <= c_menu_bh[m]) fit = m;
                        }
                        if (emp || fit >= 0 || sp == kSpans - 1) {     // the last span (1 plane) may take the global path
                            chosen = true;
                            span = span_len[sp]; shape = fit; empty = emp;
                            bx0 = emp ? 0 : (ax >> 3); by0 = emp ? 0 : ay;
                        }
                    }
                }
            }
            const int bw = shape >= 0 ? c_menu_bw[shape] : 0;
            if (shape >= 0) {
                if (tid == 0) {
                    mbar_expect_tx(bar, (unsigned)(bw * c_menu_bh[shape] * C4T * 128));
                    tma_load_4d(tile, &maps.m[shape], bar, 0, 0, bx0, ((i * b_total + b) * H) + by0);
                }
                mbar_wait(bar, phase);
                phase ^= 1u;
            }
            const int tile_base = -(by0 * bw + bx0) * (C4T * 8);

            // ---- pass 2: taps from shared memory (or from global memory if even one plane's window is too large)
            if (active) {
#pragma unroll 2
                for (int k = k0; k < k0 + span; ++k) {
                    float s = 0.0f;
                    if (!empty) {
                        const TmvsTaps t = tmvs_footprint(crd_s[k * kThreads + tid], crd_s[(kDC + k) * kThreads + tid], dims);
                        if (t.any) {
                            const int xa = min(max(t.x0, 0), W - 1), xb = min(max(t.x0 + 1, 0), W - 1);
                            const int ya = min(max(t.y0, 0), H - 1), yb = min(max(t.y0 + 1, 0), H - 1);
                            float s00 = 0.0f, s01 = 0.0f, s10 = 0.0f, s11 = 0.0f;
                            if (shape >= 0) {
                                const int ra = ya * bw * (C4T * 8) + tile_base, rb = yb * bw * (C4T * 8) + tile_base;
                                const int ca = (xa >> 3) * (C4T * 8) + (xa & 7), cb = (xb >> 3) * (C4T * 8) + (xb & 7);
                                const float4 *p00 = tile + (ra + ca), *p01 = tile + (ra + cb);
                                const float4 *p10 = tile + (rb + ca), *p11 = tile + (rb + cb);
#pragma unroll
                                for (int g = 0; g < C4T; ++g) {
                                    const float4 a = p00[g * 8], bq = p01[g * 8], cq = p10[g * 8], dq = p11[g * 8];
                                    s00 = dot4(r[g], a, s00);
                                    s01 = dot4(r[g], bq, s01);
                                    s10 = dot4(r[g], cq, s10);
                                    s11 = dot4(r[g], dq, s11);
                                }
                            } else {
                                const int ra = ya * pk.row, rb = yb * pk.row;
                                const float4 *p00 = tmvs_pk_ptr(img, tmvs_pk_off(pk, xa, ra));
                                const float4 *p01 = tmvs_pk_ptr(img, tmvs_pk_off(pk, xb, ra));
                                const float4 *p10 = tmvs_pk_ptr(img, tmvs_pk_off(pk, xa, rb));
                                const float4 *p11 = tmvs_pk_ptr(img, tmvs_pk_off(pk, xb, rb));
#pragma unroll
                                for (int g = 0; g < C4T; ++g) {
                                    const float4 a = ldg4(p00 + g * 8), bq = ldg4(p01 + g * 8);
                                    const float4 cq = ldg4(p10 + g * 8), dq = ldg4(p11 + g * 8);
                                    s00 = dot4(r[g], a, s00);
                                    s01 = dot4(r[g], bq, s01);
                                    s10 = dot4(r[g], cq, s10);
                                    s11 = dot4(r[g], dq, s11);
                                }
                            }
                            s = t.ok00 ? t.w00 * s00 : 0.0f;
                            s += t.ok01 ? t.w01 * s01 : 0.0f;
                            s += t.ok10 ? t.w10 * s10 : 0.0f;
                            s += t.ok11 ? t.w11 * s11 : 0.0f;
                            s *= inv_c;                               // .mean(1), TransMVSNet.py:80
                        }
                    }
                    if (VIEWS) __stcs(out_v + (size_t)k * HW, s);
                    if (AGG) acc_s[k * kThreads + tid] = __fadd_rn(acc_s[k * kThreads + tid], __fmul_rn(s, wi));   // :88
                }
            }
            k0 += span;
        }
        wsum = __fadd_rn(wsum, wi);                                       // TransMVSNet.py:89
    }
    if (AGG && active) {
        float *out_a = agg + ((size_t)b * D + d0) * HW + pix;
        for (int k = 0; k < nd; ++k) __stcs(out_a + (size_t)k * HW, __fdiv_rn(acc_s[k * kThreads + tid], wsum));   // :93
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
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

template <int C4T>
constexpr size_t tma_smem_bytes(bool agg)
{
    return (size_t)kCapBlocks * C4T * 128 + 2 * kDC * kThreads * 4 + 2 * kSpans * 4 * kTileY * 4 + 16 +
           (agg ? kDC * kThreads * 4 : 0);
}

template <int C4T, bool PER_PIXEL>
int launch_tma(bool views, bool do_agg, dim3 grid, cudaStream_t st, const float *ref, int64_t rB, int64_t rC, int64_t rH,
               int64_t rW, const float4 *packed, const float *depth, const float *vw, float *sim_views, float *agg,
               int b_total, int b_first, int b_chunk, int C, int D, int H, int W, int n_src, int n_dchunks,
               const TmvsGeom &geom, const TmvsTmaMaps &maps)
{
    dim3 block(kTileX, kTileY);
    const size_t smem = tma_smem_bytes<C4T>(do_agg);
#define TMVS_TMA_LAUNCH(V, A)                                                                                       \
    do {                                                                                                            \
        auto kern = costvol_tma_kernel<C4T, PER_PIXEL, V, A>;                                                       \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
        if (e != cudaSuccess) return (int)e;                                                                        \
        kern<<<grid, block, smem, st>>>(ref, rB, rC, rH, rW, packed, depth, vw, sim_views, agg, b_total, b_first,   \
                                        b_chunk, C, D, H, W, n_src, n_dchunks, geom, maps);                         \
    } while (0)
    if (views && do_agg) TMVS_TMA_LAUNCH(true, true);
    else if (views) TMVS_TMA_LAUNCH(true, false);
    else TMVS_TMA_LAUNCH(false, true);
#undef TMVS_TMA_LAUNCH
    return tmvs_launch_status();
}

}  // namespace

// Internal entry (called by tmvs_costvol_fwd): returns TMVS_E_UNSUPPORTED when the TMA path does not apply
// (C/4 not in {2,4,8}, or the driver entry point is unavailable) so the caller uses the L1 kernel.
int tmvs_costvol_fwd_tma(const float *ref, int64_t rB, int64_t rC, int64_t rH, int64_t rW, const float *packed,
                         const float *rot_trans, const float *depth, int per_pixel, const float *view_weights,
                         float *sim_views, float *agg, int B, int C, int D, int H, int W, int n_src, unsigned flags,
                         cudaStream_t st)
{
    if ((C & 3) != 0) return TMVS_E_UNSUPPORTED;
    const int c4 = C / 4;
    if (c4 != 2 && c4 != 4 && c4 != 8) return TMVS_E_UNSUPPORTED;
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return TMVS_E_UNSUPPORTED;
    const int wb = (W + 7) / 8;
    const unsigned long long rows = (unsigned long long)n_src * B * H;
    if (rows > 0x7fffffffull) return TMVS_E_UNSUPPORTED;
    TmvsTmaMaps maps;
    for (int s = 0; s < kMenu; ++s) {
        // packed tensor as seen by TMA (fp32): [rows = Nsrc*B*H][Wb][C4][32 = 8 px x 4 ch], innermost first
        cuuint64_t gdim[4] = {32, (cuuint64_t)c4, (cuuint64_t)wb, (cuuint64_t)rows};
        cuuint64_t gstr[3] = {128, (cuuint64_t)c4 * 128, (cuuint64_t)wb * c4 * 128};
        cuuint32_t box[4] = {32, (cuuint32_t)c4, (cuuint32_t)h_menu_bw[s], (cuuint32_t)h_menu_bh[s]};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&maps.m[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)packed, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return TMVS_E_UNSUPPORTED;
    }
    const int n_dchunks = (D + kDC - 1) / kDC;
    const int b_per_launch = TMVS_GEOM_SLOTS / n_src;
    for (int b0 = 0; b0 < B; b0 += b_per_launch) {
        const int bc = (B - b0 < b_per_launch) ? B - b0 : b_per_launch;
        TmvsGeom geom;
        tmvs_geom_fill(geom, rot_trans, flags, n_src, B, b0, bc);
        dim3 grid(((W + kTileX - 1) / kTileX) * n_dchunks, (H + kTileY - 1) / kTileY, bc);
        int rc;
#define TMVS_TMA_C4(C4T)                                                                                              \
    rc = per_pixel ? launch_tma<C4T, true>(sim_views != nullptr, agg != nullptr, grid, st, ref, rB, rC, rH, rW,        \
                                           (const float4 *)packed, depth, view_weights, sim_views, agg, B, b0, bc, C, \
                                           D, H, W, n_src, n_dchunks, geom, maps)                                      \
                   : launch_tma<C4T, false>(sim_views != nullptr, agg != nullptr, grid, st, ref, rB, rC, rH, rW,       \
                                            (const float4 *)packed, depth, view_weights, sim_views, agg, B, b0, bc, C, \
                                            D, H, W, n_src, n_dchunks, geom, maps)
        if (c4 == 2) { TMVS_TMA_C4(2); }
        else if (c4 == 4) { TMVS_TMA_C4(4); }
        else { TMVS_TMA_C4(8); }
#undef TMVS_TMA_C4
        if (rc != TMVS_OK) return rc;
    }
    return TMVS_OK;
}
