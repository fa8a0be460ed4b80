// Depth read-out kernels for sm_100a (pure streaming: HBM-bound).
//
// tmvs_softmax_wta_fwd fuses models/TransMVSNet.py:99-103 and models/module.py:474-482:
// one pass over the logits gives prob = exp(log_softmax), the winner-take-all index (int64,
// first maximal, as torch.argmax), the depth gathered at that index and the max-probability
// confidence.  One thread per pixel; every depth plane is read/written as a coalesced 128-byte
// row per warp, and the D logits of a pixel stay in registers between the three sweeps.
#include <string.h>

#include "tmvs_common.cuh"

namespace {

// log_softmax exactly as ATen writes it: x - max - log(sum(exp(x - max)))
template <int DT>
__global__ void __launch_bounds__(256)
softmax_wta_kernel(const float *__restrict__ logits, const float *__restrict__ dv, float *__restrict__ prob,
                   int64_t *__restrict__ index, float *__restrict__ depth, float *__restrict__ conf, int D,
                   size_t HW)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int b = blockIdx.y;
    const float *x = logits + (size_t)b * D * HW + p;
    float v[DT];
#pragma unroll
    for (int d = 0; d < DT; ++d) v[d] = (d < D) ? __ldg(x + (size_t)d * HW) : 0.0f;
    float m = v[0];
#pragma unroll
    for (int d = 1; d < DT; ++d)
        if (d < D && tmvs_gt(v[d], m)) m = v[d];
    float s = 0.0f;
#pragma unroll
    for (int d = 0; d < DT; ++d)
        if (d < D) s += expf(v[d] - m);
    const float ls = logf(s);
    float best = 0.0f;
    int bi = 0;
    float *pr = prob ? prob + (size_t)b * D * HW + p : nullptr;
#pragma unroll
    for (int d = 0; d < DT; ++d) {
        if (d < D) {
            const float pv = expf((v[d] - m) - ls);
            if (pr) pr[(size_t)d * HW] = pv;
            if (d == 0 || tmvs_gt(pv, best)) { best = pv; bi = d; }
        }
    }
    index[(size_t)b * HW + p] = bi;
    depth[(size_t)b * HW + p] = __ldg(dv + ((size_t)b * D + bi) * HW + p);
    conf[(size_t)b * HW + p] = best;
}

// (p, i) replaces (bp, bi)?  torch.argmax order: larger wins, NaN is the maximum, ties -> smaller index
__device__ __forceinline__ bool tmvs_takes(float p, int i, float bp, int bi)
{
    const bool pn = p != p, bn = bp != bp;
    if (pn | bn) return pn && (!bn || i < bi);
    return p > bp || (p == bp && i < bi);
}

// Small maps (stage 1/2 of the cascade) have too few pixels to fill 148 SMs with one thread per pixel, so the
// D planes of a pixel are split over SUB threads (different warps: every load is still a coalesced 128-byte
// row) that combine max / sum / argmax through shared memory.
template <int PER, int SUB>
__global__ void __launch_bounds__(32 * SUB)
softmax_wta_split_kernel(const float *__restrict__ logits, const float *__restrict__ dv, float *__restrict__ prob,
                         int64_t *__restrict__ index, float *__restrict__ depth, float *__restrict__ conf, int D,
                         size_t HW)
{
    __shared__ float sh_a[SUB][32];
    __shared__ float sh_b[SUB][32];
    __shared__ int sh_i[SUB][32];
    const int lane = threadIdx.x, sub = threadIdx.y;
    const size_t p = (size_t)blockIdx.x * 32 + lane;
    const bool live = p < HW;
    const int b = blockIdx.y;
    const float *x = logits + (size_t)b * D * HW + (live ? p : 0);
    float v[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int d = sub + SUB * k;
        v[k] = (live && d < D) ? __ldcs(x + (size_t)d * HW) : 0.0f;
    }
    // ---- max
    float m = 0.0f;
    bool have = false;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int d = sub + SUB * k;
        if (d < D && (!have || tmvs_gt(v[k], m))) { m = v[k]; have = true; }
    }
    sh_a[sub][lane] = m;
    sh_i[sub][lane] = have ? 1 : 0;
    __syncthreads();
    {
        float mm = 0.0f;
        bool h2 = false;
#pragma unroll
        for (int j = 0; j < SUB; ++j)
            if (sh_i[j][lane] && (!h2 || tmvs_gt(sh_a[j][lane], mm))) { mm = sh_a[j][lane]; h2 = true; }
        m = mm;
    }
    // ---- sum of exp
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < PER; ++k)
        if (sub + SUB * k < D) s += expf(v[k] - m);
    sh_b[sub][lane] = s;
    __syncthreads();
    s = 0.0f;
#pragma unroll
    for (int j = 0; j < SUB; ++j) s += sh_b[j][lane];
    const float ls = logf(s);
    // ---- probabilities + winner
    float best = 0.0f;
    int bi = 0x7fffffff;
    float *pr = (prob && live) ? prob + (size_t)b * D * HW + p : nullptr;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int d = sub + SUB * k;
        if (d < D) {
            const float pv = expf((v[k] - m) - ls);
            if (pr) __stcs(pr + (size_t)d * HW, pv);
            if (bi == 0x7fffffff || tmvs_takes(pv, d, best, bi)) { best = pv; bi = d; }
        }
    }
    __syncthreads();                   // sh_a / sh_i are reused
    sh_a[sub][lane] = best;
    sh_i[sub][lane] = bi;
    __syncthreads();
    if (sub == 0 && live) {
#pragma unroll
        for (int j = 1; j < SUB; ++j)
            if (sh_i[j][lane] != 0x7fffffff && (bi == 0x7fffffff || tmvs_takes(sh_a[j][lane], sh_i[j][lane], best, bi))) {
                best = sh_a[j][lane];
                bi = sh_i[j][lane];
            }
        index[(size_t)b * HW + p] = bi;
        depth[(size_t)b * HW + p] = __ldg(dv + ((size_t)b * D + bi) * HW + p);
        conf[(size_t)b * HW + p] = best;
    }
}

// ---- lean forms for the cascade's own plane counts (D == PER * SUB exactly, B * D * HW < 2^32) --------------------
// Same arithmetic as the kernels above -- exp(x - max - log(sum exp(x - max))), first maximal index -- with the
// bookkeeping stripped: no d < D guards, 32-bit offsets, and plain comparisons instead of the NaN-aware ones.  That is
// safe HERE because one NaN (or inf - inf) among a pixel's logits makes log(sum) NaN and with it EVERY probability of
// the pixel; torch.argmax of an all-NaN row is its first element, which is what "keep the first candidate unless a
// later one compares greater" returns.  (depth_wta below, whose input is an arbitrary volume, keeps the NaN ordering.)
// The read-out is issue-bound (ncu: 81-88 % issue slots, 20-46 % DRAM), so fewer instructions is the lever.
template <int PER, int SUB, bool PROB>
__global__ void __launch_bounds__(32 * SUB)
softmax_wta_split_lean_kernel(const float *__restrict__ logits, const float *__restrict__ dv, float *__restrict__ prob,
                              int64_t *__restrict__ index, float *__restrict__ depth, float *__restrict__ conf,
                              unsigned HW)
{
    constexpr int D = PER * SUB;
    __shared__ float sh_a[SUB][32];
    __shared__ float sh_b[SUB][32];
    __shared__ int sh_i[SUB][32];
    const int lane = threadIdx.x, sub = threadIdx.y;
    const unsigned p = blockIdx.x * 32u + lane;
    const bool live = p < HW;
    const unsigned base = blockIdx.y * (unsigned)D * HW + (live ? p : 0u) + (unsigned)sub * HW;   // plane sub, then += SUB planes
    const unsigned step = (unsigned)SUB * HW;
    float v[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) v[k] = __ldcs(logits + base + k * step);
    float m = v[0];
#pragma unroll
    for (int k = 1; k < PER; ++k) m = fmaxf(m, v[k]);
    sh_a[sub][lane] = m;
    __syncthreads();
    m = sh_a[0][lane];
#pragma unroll
    for (int j = 1; j < SUB; ++j) m = fmaxf(m, sh_a[j][lane]);
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < PER; ++k) { v[k] -= m; s += expf(v[k]); }
    sh_b[sub][lane] = s;
    __syncthreads();
    s = 0.0f;
#pragma unroll
    for (int j = 0; j < SUB; ++j) s += sh_b[j][lane];
    const float ls = logf(s);
    float best = expf(v[0] - ls);
    int bk = 0;
    if (PROB && live) prob[base] = best;
#pragma unroll
    for (int k = 1; k < PER; ++k) {
        const float pv = expf(v[k] - ls);
        if (PROB && live) prob[base + k * step] = pv;
        if (pv > best) { best = pv; bk = k; }
    }
    sh_a[sub][lane] = best;            // every thread passed the first barrier: its max slot is free again
    sh_i[sub][lane] = sub + SUB * bk;
    __syncthreads();
    if (sub == 0 && live) {
        int bi = sh_i[0][lane];
#pragma unroll
        for (int j = 1; j < SUB; ++j) {
            const float pj = sh_a[j][lane];
            const int ij = sh_i[j][lane];
            if (pj > best || (pj == best && ij < bi)) { best = pj; bi = ij; }
        }
        const unsigned o = blockIdx.y * HW + p;
        index[o] = bi;
        depth[o] = __ldg(dv + blockIdx.y * (unsigned)D * HW + (unsigned)bi * HW + p);
        conf[o] = best;
    }
}

template <int DT, bool PROB>
__global__ void __launch_bounds__(256)
softmax_wta_lean_kernel(const float *__restrict__ logits, const float *__restrict__ dv, float *__restrict__ prob,
                        int64_t *__restrict__ index, float *__restrict__ depth, float *__restrict__ conf, unsigned HW)
{
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const unsigned base = blockIdx.y * (unsigned)DT * HW + p;
    float v[DT];
#pragma unroll
    for (int d = 0; d < DT; ++d) v[d] = __ldcs(logits + base + d * HW);
    float m = v[0];
#pragma unroll
    for (int d = 1; d < DT; ++d) m = fmaxf(m, v[d]);
    float s = 0.0f;
#pragma unroll
    for (int d = 0; d < DT; ++d) { v[d] -= m; s += expf(v[d]); }
    const float ls = logf(s);
    float best = expf(v[0] - ls);
    int bi = 0;
    if (PROB) prob[base] = best;
#pragma unroll
    for (int d = 1; d < DT; ++d) {
        const float pv = expf(v[d] - ls);
        if (PROB) prob[base + d * HW] = pv;
        if (pv > best) { best = pv; bi = d; }
    }
    const unsigned o = blockIdx.y * HW + p;
    index[o] = bi;
    depth[o] = __ldg(dv + base + (unsigned)bi * HW);
    conf[o] = best;
}

// Any D (<= TMVS_MAX_DEPTH): re-reads the logits for each sweep (they sit in L1/L2).
__global__ void __launch_bounds__(256)
softmax_wta_generic_kernel(const float *__restrict__ logits, const float *__restrict__ dv, float *__restrict__ prob,
                           int64_t *__restrict__ index, float *__restrict__ depth, float *__restrict__ conf,
                           int D, size_t HW)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int b = blockIdx.y;
    const float *x = logits + (size_t)b * D * HW + p;
    float m = x[0];
    for (int d = 1; d < D; ++d) {
        const float v = x[(size_t)d * HW];
        if (tmvs_gt(v, m)) m = v;
    }
    float s = 0.0f;
    for (int d = 0; d < D; ++d) s += expf(x[(size_t)d * HW] - m);
    const float ls = logf(s);
    float best = 0.0f;
    int bi = 0;
    float *pr = prob ? prob + (size_t)b * D * HW + p : nullptr;
    for (int d = 0; d < D; ++d) {
        const float pv = expf((x[(size_t)d * HW] - m) - ls);
        if (pr) pr[(size_t)d * HW] = pv;
        if (d == 0 || tmvs_gt(pv, best)) { best = pv; bi = d; }
    }
    index[(size_t)b * HW + p] = bi;
    depth[(size_t)b * HW + p] = __ldg(dv + ((size_t)b * D + bi) * HW + p);
    conf[(size_t)b * HW + p] = best;
}

__global__ void __launch_bounds__(256)
depth_wta_kernel(const float *__restrict__ pvol, const float *__restrict__ dv, int64_t *__restrict__ index,
                 float *__restrict__ depth, int D, size_t HW)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int b = blockIdx.y;
    const float *x = pvol + (size_t)b * D * HW + p;
    float best = __ldg(x);
    int bi = 0;
#pragma unroll 8
    for (int d = 1; d < D; ++d) {
        const float v = __ldg(x + (size_t)d * HW);
        if (tmvs_gt(v, best)) { best = v; bi = d; }
    }
    if (index) index[(size_t)b * HW + p] = bi;
    depth[(size_t)b * HW + p] = __ldg(dv + ((size_t)b * D + bi) * HW + p);
}

template <bool PER_PIXEL>
__global__ void __launch_bounds__(256)
depth_regression_fwd_kernel(const float *__restrict__ pvol, const float *__restrict__ dv, float *__restrict__ depth,
                            int D, size_t HW)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int b = blockIdx.y;
    const float *x = pvol + (size_t)b * D * HW + p;
    float acc = 0.0f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
        const float z = PER_PIXEL ? __ldg(dv + ((size_t)b * D + d) * HW + p) : __ldg(dv + (size_t)b * D + d);
        acc = __fadd_rn(acc, __fmul_rn(__ldg(x + (size_t)d * HW), z));
    }
    depth[(size_t)b * HW + p] = acc;
}

template <bool PER_PIXEL>
__global__ void __launch_bounds__(256)
depth_regression_bwd_kernel(const float *__restrict__ gdepth, const float *__restrict__ dv, float *__restrict__ gp,
                            int D, size_t HW)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int b = blockIdx.y;
    const float g = __ldg(gdepth + (size_t)b * HW + p);
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
        const float z = PER_PIXEL ? __ldg(dv + ((size_t)b * D + d) * HW + p) : __ldg(dv + (size_t)b * D + d);
        gp[((size_t)b * D + d) * HW + p] = g * z;
    }
}

// ATen upsample (align_corners=False): source index of destination index i for ratio r = in/out
__device__ __forceinline__ float area_src(int i, float r)
{
    const float s = r * ((float)i + 0.5f) - 0.5f;
    return s < 0.0f ? 0.0f : s;
}

// value of the bilinearly upsampled previous depth at image pixel (iy, ix)   (TransMVSNet.py:175-178)
__device__ __forceinline__ float upsampled_prev(const float *__restrict__ prev, int hp, int wp, float ry, float rx,
                                                int iy, int ix)
{
    const float sy = area_src(iy, ry), sx = area_src(ix, rx);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < hp - 1 ? 1 : 0), x1 = x0 + (x0 < wp - 1 ? 1 : 0);
    const float ly1 = sy - (float)y0, lx1 = sx - (float)x0, ly0 = 1.0f - ly1, lx0 = 1.0f - lx1;
    return ly0 * (lx0 * __ldg(prev + (size_t)y0 * wp + x0) + lx1 * __ldg(prev + (size_t)y0 * wp + x1)) +
           ly1 * (lx0 * __ldg(prev + (size_t)y1 * wp + x0) + lx1 * __ldg(prev + (size_t)y1 * wp + x1));
}

__global__ void __launch_bounds__(256)
depth_hypotheses_kernel(const float *__restrict__ prev, int prev_planes, int hp, int wp, float interval,
                        float *__restrict__ out, int D, int h, int w, int scale)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const int b = blockIdx.z;
    const size_t hw = (size_t)h * w;
    float *o = out + (size_t)b * D * hw + (size_t)y * w + x;
    if (prev_planes > 0) {
        // 2-D branch of get_depth_samples (module.py:616-623): the global range cut into D planes
        const float lo = __ldg(prev + (size_t)b * prev_planes), hi = __ldg(prev + (size_t)b * prev_planes + prev_planes - 1);
        const float step = (hi - lo) / (float)(D - 1);
        for (int d = 0; d < D; ++d) o[(size_t)d * hw] = __fadd_rn(lo, __fmul_rn((float)d, step));   // module.py:618-621: product and sum rounded separately
        return;
    }
    // 3-D branch (module.py:624-632) at the image pixels the trilinear resample (TransMVSNet.py:202-204,
    // align_corners=False) blends: for scale 2 / 4 the two centre pixels of each axis with weight 1/2 each
    const int himg = h * scale, wimg = w * scale;
    const float ry = (float)hp / (float)himg, rx = (float)wp / (float)wimg;
    const float *pv = prev + (size_t)b * hp * wp;
    const int off = scale == 1 ? 0 : scale / 2 - 1;           // first of the two blended pixels
    const int n = scale == 1 ? 1 : 2;
    const float half_span = (float)D / 2.0f * interval;
    // the (up to) four upsampled depths this output pixel blends do not depend on the plane: fetch them once
    float lo[2][2], st[2][2];
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            const float up = upsampled_prev(pv, hp, wp, ry, rx, y * scale + off + j, x * scale + off + i);
            const float cur_min = up - half_span, cur_max = up + half_span;
            lo[j][i] = cur_min;
            st[j][i] = (cur_max - cur_min) / (float)(D - 1);
        }
    for (int d = 0; d < D; ++d) {
        const float fd = (float)d;
        float v;
        if (n == 1) {
            v = __fadd_rn(lo[0][0], __fmul_rn(fd, st[0][0]));     // module.py:630-632: no contraction into an FMA
        } else {
            const float s00 = __fadd_rn(lo[0][0], __fmul_rn(fd, st[0][0])), s01 = __fadd_rn(lo[0][1], __fmul_rn(fd, st[0][1]));
            const float s10 = __fadd_rn(lo[1][0], __fmul_rn(fd, st[1][0])), s11 = __fadd_rn(lo[1][1], __fmul_rn(fd, st[1][1]));
            const float r0 = __fadd_rn(__fmul_rn(0.5f, s00), __fmul_rn(0.5f, s01));
            const float r1 = __fadd_rn(__fmul_rn(0.5f, s10), __fmul_rn(0.5f, s11));
            v = __fadd_rn(__fmul_rn(0.5f, r0), __fmul_rn(0.5f, r1));
        }
        __stcs(o + (size_t)d * hw, v);
    }
}

// cv2.resize(..., INTER_LINEAR) tap of a float32 map: source index / weight exactly as OpenCV forms them
// (fx = (float)((dx + 0.5) * scale - 0.5), scale = src / dst in double; borders replicate)
struct CvTap {
    int i0, i1;
    float a0, a1;
};

__device__ __forceinline__ CvTap cv_tap(int d, int src, int dst)
{
    const double scale = (double)src / (double)dst;
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { f = 0.0f; s = 0; }
    if (s >= src - 1) { f = 0.0f; s = src - 1; }
    CvTap t;
    t.i0 = s;
    t.i1 = min(s + 1, src - 1);
    t.a0 = 1.0f - f;
    t.a1 = f;
    return t;
}

__device__ __forceinline__ float cv_bilinear(const float *__restrict__ m, int h, int w, int y, int x, int H, int W)
{
    const CvTap tx = cv_tap(x, w, W), ty = cv_tap(y, h, H);
    // OpenCV: horizontal pass per source row, then the vertical blend
    const float r0 = __fadd_rn(__fmul_rn(__ldg(m + (size_t)ty.i0 * w + tx.i0), tx.a0), __fmul_rn(__ldg(m + (size_t)ty.i0 * w + tx.i1), tx.a1));
    const float r1 = __fadd_rn(__fmul_rn(__ldg(m + (size_t)ty.i1 * w + tx.i0), tx.a0), __fmul_rn(__ldg(m + (size_t)ty.i1 * w + tx.i1), tx.a1));
    return __fadd_rn(__fmul_rn(r0, ty.a0), __fmul_rn(r1, ty.a1));
}

__global__ void __launch_bounds__(256)
finalize_maps_kernel(const float *__restrict__ depth, const float *__restrict__ conf3, const float *__restrict__ conf1,
                     int h1, int w1, const float *__restrict__ conf2, int h2, int w2, float thr, float dmin, float dmax,
                     float *__restrict__ depth_out, float *__restrict__ conf_out, uint8_t *__restrict__ alpha_out,
                     int H, int W)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const int b = blockIdx.z;
    const size_t p = ((size_t)b * H + y) * W + x;
    const float c1 = cv_bilinear(conf1 + (size_t)b * h1 * w1, h1, w1, y, x, H, W);
    const float c2 = cv_bilinear(conf2 + (size_t)b * h2 * w2, h2, w2, y, x, H, W);
    const float cf = __fmul_rn(__fmul_rn(__ldg(conf3 + p), c1), c2);       // test.py:132
    float d = __ldg(depth + p);
    if (cf < thr) d = 0.0f;                                                  // test.py:144
    if (depth_out) depth_out[p] = d;
    if (conf_out) conf_out[p] = cf;
    if (alpha_out) {                                                         // utils.py:11-21 depth_normal
        float q = d < dmin ? dmin : d;
        q = q > dmax ? dmax : q;
        const float sc = __fmul_rn(__fdiv_rn(__fsub_rn(q, dmin), __fsub_rn(dmax, dmin)), 255.0f);
        alpha_out[p] = (uint8_t)(int)sc;
    }
}

inline bool bad_dims(int B, int D, int H, int W)
{
    return B <= 0 || D <= 0 || H <= 0 || W <= 0 || B > 65535 || D > TMVS_MAX_DEPTH;
}

}  // namespace

extern "C" int tmvs_softmax_wta_fwd(const float *logits, const float *depth_values, float *prob, int64_t *index,
                                    float *depth, float *conf, int B, int D, int H, int W, tmvs_stream_t stream)
{
    if (!logits || !depth_values || !index || !depth || !conf) return TMVS_E_NULL;
    if (bad_dims(B, D, H, W)) return TMVS_E_SHAPE;
    const size_t HW = (size_t)H * W;
    dim3 grid((unsigned)((HW + 255) / 256), B);
    cudaStream_t st = (cudaStream_t)stream;
#define TMVS_RO(DT) softmax_wta_kernel<DT><<<grid, 256, 0, st>>>(logits, depth_values, prob, index, depth, conf, D, HW)
#define TMVS_RO_SPLIT(PER, SUB)                                                                                   \
    softmax_wta_split_kernel<PER, SUB><<<dim3((unsigned)((HW + 31) / 32), B), dim3(32, SUB), 0, st>>>(            \
        logits, depth_values, prob, index, depth, conf, D, HW)
    // one thread per pixel saturates the machine from ~1 M pixels; below that the planes are split over threads
    const bool small = (size_t)B * HW < ((size_t)1 << 20);
    // the cascade's own plane counts (48 / 32 / 8) with 32-bit offsets: the lean kernels
    if ((size_t)B * D * HW < 0xffffffffull && (D == 8 || (small && (D == 32 || D == 48)))) {
        const unsigned hw = (unsigned)HW;
        const dim3 sgrid((unsigned)((HW + 31) / 32), B), sblock(32, 8);
#define TMVS_RO_LEAN(PROB)                                                                                              \
        if (D == 8) softmax_wta_lean_kernel<8, PROB><<<grid, 256, 0, st>>>(logits, depth_values, prob, index, depth, conf, hw); \
        else if (D == 32) softmax_wta_split_lean_kernel<4, 8, PROB><<<sgrid, sblock, 0, st>>>(logits, depth_values, prob, index, depth, conf, hw); \
        else softmax_wta_split_lean_kernel<6, 8, PROB><<<sgrid, sblock, 0, st>>>(logits, depth_values, prob, index, depth, conf, hw)
        if (prob) { TMVS_RO_LEAN(true); } else { TMVS_RO_LEAN(false); }
#undef TMVS_RO_LEAN
        return tmvs_launch_status();
    }
    if (D <= 8) TMVS_RO(8);
    else if (small && D <= 32 && D > 16) TMVS_RO_SPLIT(4, 8);
    else if (small && D <= 48 && D > 32) TMVS_RO_SPLIT(6, 8);
    else if (small && D <= 64 && D > 48) TMVS_RO_SPLIT(8, 8);
    else if (D <= 16) TMVS_RO(16);
    else if (D <= 32) TMVS_RO(32);
    else if (D <= 48) TMVS_RO(48);
    else if (D <= 64) TMVS_RO(64);
    else softmax_wta_generic_kernel<<<grid, 256, 0, st>>>(logits, depth_values, prob, index, depth, conf, D, HW);
#undef TMVS_RO_SPLIT
#undef TMVS_RO
    return tmvs_launch_status();
}

extern "C" int tmvs_depth_wta(const float *p, const float *depth_values, int64_t *index, float *depth, int B, int D,
                              int H, int W, tmvs_stream_t stream)
{
    if (!p || !depth_values || !depth) return TMVS_E_NULL;
    if (bad_dims(B, D, H, W)) return TMVS_E_SHAPE;
    const size_t HW = (size_t)H * W;
    dim3 grid((unsigned)((HW + 255) / 256), B);
    depth_wta_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, depth_values, index, depth, D, HW);
    return tmvs_launch_status();
}

extern "C" int tmvs_depth_regression_fwd(const float *p, const float *depth_values, int per_pixel, float *depth,
                                         int B, int D, int H, int W, tmvs_stream_t stream)
{
    if (!p || !depth_values || !depth) return TMVS_E_NULL;
    if (bad_dims(B, D, H, W)) return TMVS_E_SHAPE;
    const size_t HW = (size_t)H * W;
    dim3 grid((unsigned)((HW + 255) / 256), B);
    if (per_pixel)
        depth_regression_fwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(p, depth_values, depth, D, HW);
    else
        depth_regression_fwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(p, depth_values, depth, D, HW);
    return tmvs_launch_status();
}

extern "C" int tmvs_depth_regression_bwd(const float *grad_depth, const float *depth_values, int per_pixel,
                                         float *grad_p, int B, int D, int H, int W, tmvs_stream_t stream)
{
    if (!grad_depth || !depth_values || !grad_p) return TMVS_E_NULL;
    if (bad_dims(B, D, H, W)) return TMVS_E_SHAPE;
    const size_t HW = (size_t)H * W;
    dim3 grid((unsigned)((HW + 255) / 256), B);
    if (per_pixel)
        depth_regression_bwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(grad_depth, depth_values, grad_p, D, HW);
    else
        depth_regression_bwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(grad_depth, depth_values, grad_p, D, HW);
    return tmvs_launch_status();
}

extern "C" int tmvs_finalize_maps_fwd(const float *depth, const float *conf3, const float *conf1, int h1, int w1,
                                      const float *conf2, int h2, int w2, float conf_threshold, float depth_min,
                                      float depth_max, float *depth_out, float *conf_out, uint8_t *alpha_out, int B,
                                      int H, int W, tmvs_stream_t stream)
{
    if (!depth || !conf3 || !conf1 || !conf2) return TMVS_E_NULL;
    if (!depth_out && !conf_out && !alpha_out) return TMVS_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || h1 <= 0 || w1 <= 0 || h2 <= 0 || w2 <= 0 || B > 65535) return TMVS_E_SHAPE;
    dim3 grid((W + 31) / 32, (H + 7) / 8, B);
    finalize_maps_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(depth, conf3, conf1, h1, w1, conf2, h2, w2,
                                                                        conf_threshold, depth_min, depth_max, depth_out,
                                                                        conf_out, alpha_out, H, W);
    return tmvs_launch_status();
}

extern "C" int tmvs_depth_hypotheses_fwd(const float *prev_depth, int prev_planes, int hp, int wp, float interval,
                                         float *out, int B, int D, int h, int w, int scale, tmvs_stream_t stream)
{
    if (!prev_depth || !out) return TMVS_E_NULL;
    if (bad_dims(B, D, h, w) || D < 2) return TMVS_E_SHAPE;
    if (scale != 1 && scale != 2 && scale != 4) return TMVS_E_UNSUPPORTED;
    if (prev_planes < 0 || (prev_planes == 0 && (hp <= 0 || wp <= 0)) || prev_planes == 1) return TMVS_E_SHAPE;
    dim3 grid((w + 31) / 32, (h + 7) / 8, B);
    depth_hypotheses_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(prev_depth, prev_planes, hp, wp, interval,
                                                                           out, D, h, w, scale);
    return tmvs_launch_status();
}

extern "C" int tmvs_version(void) { return TMVS_VERSION; }

// ---- peer-mapped gather buffer (sharding.PeerMapSink) -----------------------------------------------------------------
// The one exchange of the multi-GPU path is "every view's depth + confidence map ends up on one rank".  Instead of a
// collective, that rank exports ONE buffer through CUDA IPC; every other rank opens it with its own GPU current and
// cudaIpcMemLazyEnablePeerAccess, which maps the buffer into that GPU's address space over NVLink / NVSwitch.  The
// returned pointer is an ordinary output pointer for tmvs_softmax_wta_fwd: the kernel's stores cross the link.
// These three calls are the only ones in the library that allocate or free device memory.
extern "C" int tmvs_peer_buffer_create(size_t bytes, void **ptr, unsigned char *handle64)
{
    if (!ptr || !handle64) return TMVS_E_NULL;
    if (bytes == 0) return TMVS_E_SHAPE;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, bytes);
    // the fill must have landed before a peer can be handed the buffer: a memset is not ordered against other GPUs' stores
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return (int)e; }
    memcpy(handle64, &h, 64);
    return TMVS_OK;
}

extern "C" int tmvs_peer_buffer_open(const unsigned char *handle64, void **ptr)
{
    if (!ptr || !handle64) return TMVS_E_NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    const cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? TMVS_OK : (int)e;
}

// Copy-engine transport into a peer-mapped buffer: an asynchronous device-to-device copy whose destination lives on
// another GPU of the box.  It runs on the DMA engines over NVLink, not on the SMs, so on a side stream it overlaps the
// next reference view's kernels completely.  Measured at 8 GPUs it lands on the same step time as letting the read-out
// kernel store into the peer mapping itself (1.8064 vs 1.8063 ms): neither is a bottleneck.
extern "C" int tmvs_peer_copy_async(void *dst, const void *src, size_t bytes, tmvs_stream_t stream)
{
    if (!dst || !src) return TMVS_E_NULL;
    if (bytes == 0) return TMVS_E_SHAPE;
    const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
    return e == cudaSuccess ? TMVS_OK : (int)e;
}

extern "C" int tmvs_peer_buffer_release(void *ptr, int owner)
{
    if (!ptr) return TMVS_E_NULL;
    const cudaError_t e = owner ? cudaFree(ptr) : cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? TMVS_OK : (int)e;
}

extern "C" const char *tmvs_error_string(int code)
{
    switch (code) {
    case TMVS_OK: return "ok";
    case TMVS_E_NULL: return "tmvs: required pointer is NULL";
    case TMVS_E_SHAPE: return "tmvs: dimension out of range";
    case TMVS_E_ALIGN: return "tmvs: pointer not 16-byte aligned";
    case TMVS_E_UNSUPPORTED: return "tmvs: unsupported combination";
    default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "tmvs: unknown error";
}
