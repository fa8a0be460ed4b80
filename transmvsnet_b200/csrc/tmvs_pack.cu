// Layout pre-pass + drop-in homo_warping for sm_100a.
//
// tmvs_pack_sources: source feature maps (NCHW as produced by the reference's FeatureNet/FMT,
// models/module.py:399-422, or any strided view such as channels_last) -> the kernel-native
// blocked channel-last layout [Nsrc][B][H][Wb][C4][8 px][4 ch].  Reads are coalesced along x per channel
// plane, writes are 128-bit and contiguous.
//
// tmvs_homo_warp_fwd: models/module.py:284-322 for one view, materialising [B][C][D][H][W]
// (the reference's own interface; the fused kernels in tmvs_costvol.cu never form this volume).
#include "tmvs_common.cuh"

// TMA-engine variant (tmvs_pack_tma.cu); TMVS_E_UNSUPPORTED when it does not apply
int tmvs_pack_sources_tma(const float *const *src, int n_src, int64_t sB, int64_t sC, int64_t sH, int64_t sW,
                          float *packed, int B, int C, int H, int W, cudaStream_t st);

namespace {

struct SrcPtrs {
    const float *p[TMVS_MAX_SRC_VIEWS];
};

__global__ void __launch_bounds__(256)
pack_sources_kernel(SrcPtrs src, int64_t sB, int64_t sC, int64_t sH, int64_t sW, float4 *__restrict__ packed,
                    int B, int C, int c4, int H, int W)
{
    const size_t HW = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int g = blockIdx.y;
    const int vb = blockIdx.z;                 // view * B + b
    const int view = vb / B, b = vb - view * B;
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    const float *base = src.p[view] + b * sB + y * sH + x * sW;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = 4 * g + j;
        v[j] = (c < C) ? __ldg(base + c * sC) : 0.0f;
    }
    const TmvsPacked pk = tmvs_packed_layout(c4, H, W);
    packed[(size_t)vb * pk.slice + tmvs_pk_off(pk, x, y * pk.row) + g * 8] = make_float4(v[0], v[1], v[2], v[3]);
}

// NCHW fast path (x contiguous, W % 4 == 0, 16-byte aligned rows): one thread moves a 4-pixel x 4-channel
// block of TMVS_PACK_G consecutive channel groups -- 4 * G independent 128-bit streaming loads (one per channel
// plane), a register transpose, 4 * G 128-bit stores.
#ifndef TMVS_PACK_G
#define TMVS_PACK_G 2
#endif
#ifndef TMVS_PACK_THREADS
#define TMVS_PACK_THREADS 128
#endif
__global__ void __launch_bounds__(TMVS_PACK_THREADS)
pack_sources_nchw4_kernel(SrcPtrs src, int64_t sB, int64_t sC, int64_t sH, float4 *__restrict__ packed,
                          int B, int C, int c4, int H, int W)
{
    const int wq = W >> 2;
    const size_t nquads = (size_t)H * wq;
    const size_t qd = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qd >= nquads) return;
    const int g0 = blockIdx.y * TMVS_PACK_G;
    const int vb = blockIdx.z;
    const int view = vb / B, b = vb - view * B;
    const int y = (int)(qd / wq), xq = (int)(qd - (size_t)y * wq);
    const float *base = src.p[view] + b * sB + y * sH + 4 * xq;
    float4 v[TMVS_PACK_G][4];
#pragma unroll
    for (int gg = 0; gg < TMVS_PACK_G; ++gg) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 4 * (g0 + gg) + j;
            v[gg][j] = (c < C) ? __ldcs(reinterpret_cast<const float4 *>(base + c * sC)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const TmvsPacked pk = tmvs_packed_layout(c4, H, W);
    float4 *o0 = packed + (size_t)vb * pk.slice + tmvs_pk_off(pk, 4 * xq, y * pk.row);   // 4 | 8: same block
#pragma unroll
    for (int gg = 0; gg < TMVS_PACK_G; ++gg) {
        if (g0 + gg < c4) {
            float4 *o = o0 + (g0 + gg) * 8;
            o[0] = make_float4(v[gg][0].x, v[gg][1].x, v[gg][2].x, v[gg][3].x);
            o[1] = make_float4(v[gg][0].y, v[gg][1].y, v[gg][2].y, v[gg][3].y);
            o[2] = make_float4(v[gg][0].z, v[gg][1].z, v[gg][2].z, v[gg][3].z);
            o[3] = make_float4(v[gg][0].w, v[gg][1].w, v[gg][2].w, v[gg][3].w);
        }
    }
}

// channels_last fast path (channel stride 1, C % 4 == 0, aligned): a pure 128-bit permuting copy.
__global__ void __launch_bounds__(256)
pack_sources_nhwc_kernel(SrcPtrs src, int64_t sB, int64_t sH, int64_t sW, float4 *__restrict__ packed,
                         int B, int c4, int H, int W)
{
    const size_t HW = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int g = blockIdx.y;
    const int vb = blockIdx.z;
    const int view = vb / B, b = vb - view * B;
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    const float *base = src.p[view] + b * sB + y * sH + x * sW + 4 * g;
    const TmvsPacked pk = tmvs_packed_layout(c4, H, W);
    packed[(size_t)vb * pk.slice + tmvs_pk_off(pk, x, y * pk.row) + g * 8] = __ldg(reinterpret_cast<const float4 *>(base));
}

constexpr int kWarpDC = 4;   // depth planes per thread in the drop-in warp

template <bool PER_PIXEL>
__global__ void __launch_bounds__(256)
homo_warp_fwd_kernel(const float4 *__restrict__ packed, const float *__restrict__ depth, float *__restrict__ out,
                     int b_first, int b_chunk, int C, int c4, int D, int H, int W, int n_dchunks,
                     const __grid_constant__ TmvsGeom geom)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const int bl = blockIdx.z / n_dchunks;
    const int d0 = (blockIdx.z - bl * n_dchunks) * kWarpDC;
    const int b = b_first + bl;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)y * W + x;
    float rt[12];
    tmvs_geom_rt(geom, 0, bl, b_chunk, rt);
    const TmvsRay ray = tmvs_ray(rt, (float)x, (float)y, geom.ray_unfused);
    const TmvsDims dims = tmvs_dims(H, W, geom.arith);
    const TmvsPacked pk = tmvs_packed_layout(c4, H, W);
    const float4 *img = packed + (size_t)b * pk.slice;
#pragma unroll
    for (int k = 0; k < kWarpDC; ++k) {
        const int d = d0 + k;
        if (d >= D) break;
        const float dep = PER_PIXEL ? __ldg(depth + ((size_t)b * D + d) * HW + pix) : __ldg(depth + (size_t)b * D + d);
        const TmvsTaps t = tmvs_taps(ray, rt, dep, dims);
        float *o = out + (((size_t)b * C) * D + d) * HW + pix;
        const size_t c_stride = (size_t)D * HW;
        if (!t.any) {
            for (int c = 0; c < C; ++c) o[c * c_stride] = 0.0f;
            continue;
        }
        const int xa = min(max(t.x0, 0), W - 1), xb = min(max(t.x0 + 1, 0), W - 1);
        const int ra = min(max(t.y0, 0), H - 1) * pk.row, rb = min(max(t.y0 + 1, 0), H - 1) * pk.row;
        const float4 *p00 = tmvs_pk_ptr(img, tmvs_pk_off(pk, xa, ra));
        const float4 *p01 = tmvs_pk_ptr(img, tmvs_pk_off(pk, xb, ra));
        const float4 *p10 = tmvs_pk_ptr(img, tmvs_pk_off(pk, xa, rb));
        const float4 *p11 = tmvs_pk_ptr(img, tmvs_pk_off(pk, xb, rb));
        for (int g = 0; g < c4; ++g) {
            const float4 a = ldg4(p00 + g * 8), bq = ldg4(p01 + g * 8);
            const float4 cq = ldg4(p10 + g * 8), dq = ldg4(p11 + g * 8);
            float va[4] = {a.x, a.y, a.z, a.w}, vb[4] = {bq.x, bq.y, bq.z, bq.w};
            float vc[4] = {cq.x, cq.y, cq.z, cq.w}, vd[4] = {dq.x, dq.y, dq.z, dq.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 4 * g + j;
                if (c < C) {
                    // ATen accumulation order: nw, ne, sw, se; out-of-bounds taps are skipped
                    float acc = 0.0f;
                    if (t.ok00) acc = fmaf(va[j], t.w00, acc);
                    if (t.ok01) acc = fmaf(vb[j], t.w01, acc);
                    if (t.ok10) acc = fmaf(vc[j], t.w10, acc);
                    if (t.ok11) acc = fmaf(vd[j], t.w11, acc);
                    o[c * c_stride] = acc;
                }
            }
        }
    }
}

}  // namespace

extern "C" size_t tmvs_packed_bytes(int n_src, int B, int C, int H, int W)
{
    if (n_src <= 0 || B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)n_src * B * tmvs_packed_layout((C + 3) / 4, H, W).slice * 4 * sizeof(float);
}

extern "C" int tmvs_pack_sources(const float *const *src, int n_src, int64_t sB, int64_t sC, int64_t sH, int64_t sW,
                                 float *packed, int B, int C, int H, int W, unsigned flags, tmvs_stream_t stream)
{
    if (!src || !packed) return TMVS_E_NULL;
    if (n_src <= 0 || n_src > TMVS_MAX_SRC_VIEWS || B <= 0 || C <= 0 || H <= 0 || W <= 0) return TMVS_E_SHAPE;
    if ((size_t)n_src * B > 65535) return TMVS_E_SHAPE;
    if (((uintptr_t)packed & 15) != 0) return TMVS_E_ALIGN;
    SrcPtrs ptrs;
    for (int i = 0; i < TMVS_MAX_SRC_VIEWS; ++i) ptrs.p[i] = nullptr;
    for (int i = 0; i < n_src; ++i) {
        if (!src[i]) return TMVS_E_NULL;
        ptrs.p[i] = src[i];
    }
    const int c4 = (C + 3) / 4;
    const size_t HW = (size_t)H * W;
    bool aligned = true;
    for (int i = 0; i < n_src; ++i) aligned = aligned && (((uintptr_t)src[i] & 15) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    // contiguous NCHW maps go through the TMA engine (TMVS_F_PACK_LDG keeps the register-transpose kernel)
    if (!(flags & TMVS_F_PACK_LDG)) {
        const int rc = tmvs_pack_sources_tma(src, n_src, sB, sC, sH, sW, packed, B, C, H, W, st);
        if (rc != TMVS_E_UNSUPPORTED) return rc;
    }
    // (a whole-line variant -- 8 pixels per thread -- measured ~8 % slower on B200 and was dropped; two channel
    //  groups per thread and 128-thread CTAs measured 13 % faster: scripts/tune_pack.py)
    if (aligned && sW == 1 && (W & 3) == 0 && (sH & 3) == 0 && (sC & 3) == 0 && (sB & 3) == 0) {
        dim3 grid((unsigned)((HW / 4 + TMVS_PACK_THREADS - 1) / TMVS_PACK_THREADS), (c4 + TMVS_PACK_G - 1) / TMVS_PACK_G, n_src * B);
        pack_sources_nchw4_kernel<<<grid, TMVS_PACK_THREADS, 0, st>>>(ptrs, sB, sC, sH, (float4 *)packed, B, C, c4, H, W);
    } else if (aligned && sC == 1 && (C & 3) == 0 && (sW & 3) == 0 && (sH & 3) == 0 && (sB & 3) == 0) {
        dim3 grid((unsigned)((HW + 255) / 256), c4, n_src * B);
        pack_sources_nhwc_kernel<<<grid, 256, 0, st>>>(ptrs, sB, sH, sW, (float4 *)packed, B, c4, H, W);
    } else {
        dim3 grid((unsigned)((HW + 255) / 256), c4, n_src * B);
        pack_sources_kernel<<<grid, 256, 0, st>>>(ptrs, sB, sC, sH, sW, (float4 *)packed, B, C, c4, H, W);
    }
    return tmvs_launch_status();
}

extern "C" int tmvs_homo_warp_fwd(const float *packed_view, const float *rot_trans, const float *depth, int per_pixel,
                                  float *out, int B, int C, int D, int H, int W, unsigned flags, tmvs_stream_t stream)
{
    if (!packed_view || !rot_trans || !depth || !out) return TMVS_E_NULL;
    if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return TMVS_E_SHAPE;
    if (((uintptr_t)packed_view & 15) != 0) return TMVS_E_ALIGN;
    const int c4 = (C + 3) / 4;
    const int n_dchunks = (D + kWarpDC - 1) / kWarpDC;
    const int b_per_launch = 65535 / n_dchunks < TMVS_GEOM_SLOTS ? 65535 / n_dchunks : TMVS_GEOM_SLOTS;
    if (b_per_launch <= 0) return TMVS_E_SHAPE;
    dim3 block(32, 8);
    for (int b0 = 0; b0 < B; b0 += b_per_launch) {
        const int bc = (B - b0 < b_per_launch) ? B - b0 : b_per_launch;
        TmvsGeom geom;
        tmvs_geom_fill(geom, rot_trans, flags, 1, B, b0, bc);
        dim3 grid((W + 31) / 32, (H + 7) / 8, bc * n_dchunks);
        if (per_pixel)
            homo_warp_fwd_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(
                (const float4 *)packed_view, depth, out, b0, bc, C, c4, D, H, W, n_dchunks, geom);
        else
            homo_warp_fwd_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(
                (const float4 *)packed_view, depth, out, b0, bc, C, c4, D, H, W, n_dchunks, geom);
        int rc = tmvs_launch_status();
        if (rc != TMVS_OK) return rc;
    }
    return TMVS_OK;
}
