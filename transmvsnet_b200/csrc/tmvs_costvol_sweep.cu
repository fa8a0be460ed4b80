// Fused cost-volume forward, "epipolar sweep" variant for sm_100a (stages with a short hypothesis step: 2 and 3 of the
// cascade).  Same results as costvol_fwd_kernel (tmvs_costvol.cu) BIT FOR BIT -- same sample positions, same
// per-tap channel dot products in the same order, same bilinear blend -- with fewer loads and fewer instructions.
//
// Why.  costvol_fwd_kernel is bound by the SM's L1 -> register data pipe (ncu: 85-90 % of peak at stages 2/3): every
// (pixel, plane, view) loads its four bilinear taps, 4 * C4 LDG.128.  But the hypotheses of ONE reference pixel all lie
// on its epipolar line in the source image, and where consecutive planes are less than a pixel apart (stages 2/3:
// ~0.85 px per plane) their 2 x 2 footprints overlap: the dot product  t(q) = <ref[p], src[q]>  of a source pixel q is
// needed by up to four planes.  A value in a register serves only its own thread, so the reuse has to happen INSIDE the
// thread, and every lane of a warp has to run the same loads: the loop is re-indexed from planes to the integer
// COLUMNS the epipolar walk crosses.  Per column the thread loads a window of three pixels across the line (x-major walk:
// column c, rows r..r+2; y-major walk: row c, columns r..r+2), takes their three dot products, and keeps the previous
// column's three.  Every plane whose footprint sits on the (previous, current) column pair is then blended from those six
// scalars -- 0, 1 or 2 planes per column, a short divergent loop over scalar work only.  Loads per plane: 3 * C4 * (px
// per plane along the major axis) instead of 4 * C4: 2.5-2.9 instead of 4 (x C4) on the config-2 geometry
// (scripts/sim_l1_banks.py).
//
// MEASURED (B200, config 2, profiles/r2_sweep_ncu_v3.csv, r2_ab_forward_sweep.json): the loads do drop -- stage 2: 23.8 M
// global load requests instead of 32.0 M (-26 %), 139.9 M L1 wavefronts instead of 179.5 M (-22 %); stage 3 (D = 8): -8 % /
// -3 % -- but the re-indexed loop costs instructions (stage 2: 483 M warp instructions instead of 314 M, 25.5 instead of
// 31.7 active lanes per instruction; 80 registers instead of 64) and the kernel turns from pipe-bound (91 % of the L1 data
// pipe, 41 % issue) into issue/latency-bound (59 % pipe, 54 % issue): 0.820 ms instead of 0.683 ms at stage 2, 0.621
// instead of 0.392 ms at stage 3.  Hence OPT-IN (TMVS_F_FWD_SWEEP), not the default; kept because it is exact and wins
// on load traffic, i.e. it is the starting point wherever hypotheses are denser than ~0.5 px per plane.
//
// Always exact: a plane whose footprint is not wholly inside the image, not on the current column pair, or not inside
// the three-pixel windows (walks steeper than 1:2, walks that jump or turn back, z < 1e-6 ...) takes the generic four-tap
// path inline.  A warp whose lanes disagree about the walk (major axis, direction) or whose hypothesis step is long
// (stage 1: ~2-4 px per plane, where the sweep would load MORE) runs the generic path for that view.
#include <atomic>

#include "tmvs_common.cuh"

namespace {

constexpr int kTileX = 32, kTileY = 8, kThreads = kTileX * kTileY;

#ifndef TMVS_SWEEP_MINB4
#define TMVS_SWEEP_MINB4 3
#endif
#ifndef TMVS_SWEEP_MINB2
#define TMVS_SWEEP_MINB2 4
#endif
template <int C4T> struct SweepMinBlocks { static constexpr int value = C4T >= 4 ? TMVS_SWEEP_MINB4 : TMVS_SWEEP_MINB2; };

struct Plane {          // one hypothesis of this thread's pixel in the current view
    float ax, bx, ay, by;   // ATen corner-weight factors: (x0 + 1 - ix), (ix - x0), likewise in y
    int x0, y0;             // north-west corner of the footprint (clamped sample position: |.| small)
};

template <bool RECIP>
__device__ __forceinline__ Plane make_plane(float rx, float ry, float rz, float tx, float ty, float tz, float dep,
                                            const TmvsFwdConst &kc)
{
    const float2 pos = tmvs_coords_lean<RECIP>(rx, ry, rz, tx, ty, tz, dep, kc);
    const float fx0 = floorf(pos.x), fy0 = floorf(pos.y);
    Plane p;
    p.x0 = (int)fx0; p.y0 = (int)fy0;
    p.ax = __fsub_rn(fx0 + 1.0f, pos.x); p.bx = __fsub_rn(pos.x, fx0);
    p.ay = __fsub_rn(fy0 + 1.0f, pos.y); p.by = __fsub_rn(pos.y, fy0);
    return p;
}

// <ref, src[pixel at word offset `off`]> exactly as costvol_fwd_kernel accumulates it (two FFMA2 chains, then their sum)
template <int C4T>
__device__ __forceinline__ float tap_dot(const float4 *img, unsigned off, const float2 (&r)[2 * C4T])
{
    const float4 *p = tmvs_pk_ptr(img, off);
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int g = 0; g < C4T; ++g) {
        const float4 a = ldg4(p + g * 8);
        s = tmvs_fma2(make_float2(a.x, a.y), r[2 * g], s);
        s = tmvs_fma2(make_float2(a.z, a.w), r[2 * g + 1], s);
    }
    return s.x + s.y;
}

__device__ __forceinline__ float blend4(const Plane &p, float t00, float t01, float t10, float t11)
{
    float s = __fmul_rn(p.ax, p.ay) * t00;
    s = fmaf(__fmul_rn(p.bx, p.ay), t01, s);
    s = fmaf(__fmul_rn(p.ax, p.by), t10, s);
    s = fmaf(__fmul_rn(p.bx, p.by), t11, s);
    return s;
}

// The generic path of costvol_fwd_kernel for one plane: four taps, per-tap zero padding at the rim.
template <int C4T>
__device__ __forceinline__ float plane_generic(const float4 *img, Plane p, const float2 (&r)[2 * C4T], const TmvsFwdConst &kc)
{
    constexpr unsigned c4x8 = C4T * 8;
    const int x0 = p.x0, y0 = p.y0;
    unsigned o00, o01, o10, o11;
    if ((unsigned)x0 < (unsigned)kc.wm1 && (unsigned)y0 < (unsigned)kc.hm1) {
        const unsigned dx = ((x0 & 7) == 7) ? c4x8 - 7u : 1u;
        o00 = (unsigned)y0 * (unsigned)kc.row + ((unsigned)x0 >> 3) * c4x8 + ((unsigned)x0 & 7u);
        o01 = o00 + dx;
        o10 = o00 + (unsigned)kc.row;
        o11 = o10 + dx;
    } else {
        const bool xin0 = (unsigned)x0 <= (unsigned)kc.wm1, xin1 = (unsigned)(x0 + 1) <= (unsigned)kc.wm1;
        const bool yin0 = (unsigned)y0 <= (unsigned)kc.hm1, yin1 = (unsigned)(y0 + 1) <= (unsigned)kc.hm1;
        if (!((xin0 | xin1) & (yin0 | yin1))) return 0.0f;
        p.ax = xin0 ? p.ax : 0.0f; p.bx = xin1 ? p.bx : 0.0f;
        p.ay = yin0 ? p.ay : 0.0f; p.by = yin1 ? p.by : 0.0f;
        const unsigned xa = (unsigned)min(max(x0, 0), kc.wm1), xb = (unsigned)min(max(x0 + 1, 0), kc.wm1);
        const unsigned ra = (unsigned)min(max(y0, 0), kc.hm1) * (unsigned)kc.row;
        const unsigned rb = (unsigned)min(max(y0 + 1, 0), kc.hm1) * (unsigned)kc.row;
        const unsigned oa = (xa >> 3) * c4x8 + (xa & 7u), ob = (xb >> 3) * c4x8 + (xb & 7u);
        o00 = ra + oa; o01 = ra + ob; o10 = rb + oa; o11 = rb + ob;
    }
    TMVS_ASSERT(max(max(o00, o01), max(o10, o11)) + (C4T - 1) * 8u < (unsigned)((kc.hm1 + 1) * kc.row));
    // the four channel dots interleaved per group, as in costvol_fwd_kernel (the sums are the same either way: each
    // chain only ever adds its own tap's products)
    return blend4(p, tap_dot<C4T>(img, o00, r), tap_dot<C4T>(img, o01, r), tap_dot<C4T>(img, o10, r),
                  tap_dot<C4T>(img, o11, r));
}

struct Window {         // the three dot products of one column of the walk
    float t0, t1, t2;
    int base;           // minor-axis coordinate of t0
};

// column `c` of the walk (major-axis coordinate), three pixels from minor coordinate `base`; coordinates are clamped into
// the image so the loads are always legal -- a clamped value is never used (only planes whose footprint is wholly inside
// the image read the windows, and their taps are in range)
template <int C4T, bool XMAJOR>
__device__ __forceinline__ Window load_window(const float4 *img, int c, int base, const float2 (&r)[2 * C4T],
                                              const TmvsFwdConst &kc)
{
    constexpr unsigned c4x8 = C4T * 8;
    Window w;
    w.base = base;
    if (XMAJOR) {
        const unsigned xc = (unsigned)min(max(c, 0), kc.wm1);
        const unsigned ox = (xc >> 3) * c4x8 + (xc & 7u);
        const unsigned r0 = (unsigned)min(max(base, 0), kc.hm1), r1 = (unsigned)min(max(base + 1, 0), kc.hm1),
                       r2 = (unsigned)min(max(base + 2, 0), kc.hm1);
        TMVS_ASSERT(r2 * (unsigned)kc.row + ox + (C4T - 1) * 8u < (unsigned)((kc.hm1 + 1) * kc.row));
        w.t0 = tap_dot<C4T>(img, r0 * (unsigned)kc.row + ox, r);
        w.t1 = tap_dot<C4T>(img, r1 * (unsigned)kc.row + ox, r);
        w.t2 = tap_dot<C4T>(img, r2 * (unsigned)kc.row + ox, r);
    } else {
        const unsigned ro = (unsigned)min(max(c, 0), kc.hm1) * (unsigned)kc.row;
        const unsigned x0 = (unsigned)min(max(base, 0), kc.wm1), x1 = (unsigned)min(max(base + 1, 0), kc.wm1),
                       x2 = (unsigned)min(max(base + 2, 0), kc.wm1);
        TMVS_ASSERT(ro + (x2 >> 3) * c4x8 + (x2 & 7u) + (C4T - 1) * 8u < (unsigned)((kc.hm1 + 1) * kc.row));
        w.t0 = tap_dot<C4T>(img, ro + (x0 >> 3) * c4x8 + (x0 & 7u), r);
        w.t1 = tap_dot<C4T>(img, ro + (x1 >> 3) * c4x8 + (x1 & 7u), r);
        w.t2 = tap_dot<C4T>(img, ro + (x2 >> 3) * c4x8 + (x2 & 7u), r);
    }
    return w;
}

// One view of one thread's planes through the sweep.  SGN = +1: the walk moves towards larger major coordinates,
// -1: towards smaller ones.  emit(k, s) receives the per-view similarity sum (before the 1/C of the channel mean).
//
// The loop is WARP-SYNCHRONOUS: every iteration all lanes that still have planes advance their windows by exactly one
// column -- one window load executed by the whole warp -- and then emit the 0..2 planes that sit on the new column pair.
// A lane whose next plane is two columns ahead simply emits nothing for one iteration; only a gap of three or more
// columns (clamped or invalid positions) re-anchors, as a divergent extra load.  (A first version let each lane jump
// ahead on its own: with per-pixel hypotheses some lane of nearly every warp did, and the extra partial-warp loads
// cost more requests than the reuse saved -- profiles/r2_sweep_ncu_v1.csv.)
template <int C4T, bool XMAJOR, bool RECIP, bool PER_PIXEL, typename Emit>
__device__ __forceinline__ void sweep_view(const float4 *img, const float2 (&r)[2 * C4T], const TmvsFwdConst &kc,
                                           float rx, float ry, float rz, float tx, float ty, float tz,
                                           const float *dep_base, int dep_stride, int nd, int sgn, int minor_up,
                                           unsigned lanes, Emit emit)
{
    const bool fwd = sgn > 0;
    int k = 0;
    Plane p = make_plane<RECIP>(rx, ry, rz, tx, ty, tz, __ldg(dep_base), kc);
    // mirrored major coordinate: g grows along the walk; the footprint of a plane sits on columns (g, g + 1)
    auto g_of = [&](const Plane &q) { const int fu = XMAJOR ? q.x0 : q.y0; return fwd ? fu : -(fu + 1); };
    auto col_of = [&](int g) { return fwd ? g : -g; };
    auto base_of = [&](const Plane &q) { const int fv = XMAJOR ? q.y0 : q.x0; return minor_up ? fv : fv - 1; };
    int gA = g_of(p) - 1;                       // the first iteration shifts onto the first plane's column
    Window A, B = load_window<C4T, XMAJOR>(img, col_of(gA + 1), base_of(p), r, kc);
    A = B;
    while (__any_sync(lanes, k < nd)) {
        if (k < nd) {
            // ---- advance by one column (whole warp), or re-anchor across a gap
            const int gap = g_of(p) - gA;
            if (gap >= 3 || gap < 0) {
                gA = g_of(p);
                A = load_window<C4T, XMAJOR>(img, col_of(gA), base_of(p), r, kc);
            } else if (gap >= 1) {
                A = B;
                ++gA;
            }
            if (gap != 0) B = load_window<C4T, XMAJOR>(img, col_of(gA + 1), base_of(p), r, kc);
            // ---- the planes on this column pair (usually one, sometimes none or two)
#pragma unroll 1
            while (k < nd && g_of(p) == gA) {
                const int fv = XMAJOR ? p.y0 : p.x0;
                // lo = the window on major coordinate fu, hi = the one on fu + 1 (selected by value: a reference to one
                // of two locals would force both windows into local memory, i.e. back onto the load pipe)
                const int lo_base = fwd ? A.base : B.base, hi_base = fwd ? B.base : A.base;
                const unsigned il = (unsigned)(fv - lo_base), ih = (unsigned)(fv - hi_base);
                const bool interior = (unsigned)p.x0 < (unsigned)kc.wm1 && (unsigned)p.y0 < (unsigned)kc.hm1;
                float s;
                if (interior && il <= 1u && ih <= 1u) {
                    const float lo0 = fwd ? A.t0 : B.t0, lo1 = fwd ? A.t1 : B.t1, lo2 = fwd ? A.t2 : B.t2;
                    const float hi0 = fwd ? B.t0 : A.t0, hi1 = fwd ? B.t1 : A.t1, hi2 = fwd ? B.t2 : A.t2;
                    const float l0 = il ? lo1 : lo0, l1 = il ? lo2 : lo1;
                    const float h0 = ih ? hi1 : hi0, h1 = ih ? hi2 : hi1;
                    // taps (x0,y0), (x0+1,y0), (x0,y0+1), (x0+1,y0+1)
                    s = XMAJOR ? blend4(p, l0, h0, l1, h1) : blend4(p, l0, l1, h0, h1);
                } else {
                    s = plane_generic<C4T>(img, p, r, kc);
                }
                emit(k, s);
                ++k;
                if (k < nd) p = make_plane<RECIP>(rx, ry, rz, tx, ty, tz, __ldg(dep_base + (size_t)k * dep_stride), kc);
            }
        }
    }
}

template <int C4T, int DC, bool PER_PIXEL, bool RECIP>
__global__ void __launch_bounds__(kThreads, SweepMinBlocks<C4T>::value)
costvol_fwd_sweep_kernel(const float *__restrict__ ref, int64_t rB, int64_t rC, int64_t rH, int64_t rW,
                         const float *__restrict__ depth, const float *__restrict__ vw, int vw_shift, int vw_w, int vw_hw,
                         float *__restrict__ agg, int b_total, int b_first, int b_chunk, int C, int D, int H, int W,
                         int n_src, int n_dchunks, const __grid_constant__ TmvsFwdConst kc,
                         const __grid_constant__ TmvsGeom geom)
{
    __shared__ float acc_s[DC][kThreads];
    const int chunk = blockIdx.x % n_dchunks;
    const int x = (blockIdx.x / n_dchunks) * kTileX + threadIdx.x;
    const int y = blockIdx.y * kTileY + threadIdx.y;
    if (x >= W || y >= H) return;
    const unsigned lanes = __activemask();              // the lanes of this warp that own a pixel
    const int tid = threadIdx.y * kTileX + threadIdx.x;
    const int bl = blockIdx.z;
    const int d0 = chunk * DC;
    const int nd = min(DC, D - d0);
    const int b = b_first + bl;
    const int HW = H * W;
    const int pix = y * W + x;

    float2 r[2 * C4T];
    {
        const float *rp = ref + b * rB + y * rH + x * rW;
#pragma unroll
        for (int g = 0; g < 2 * C4T; ++g) {
            r[g].x = __ldg(rp + (2 * g) * rC);
            r[g].y = __ldg(rp + (2 * g + 1) * rC);
        }
    }
    const float *dep_base = PER_PIXEL ? depth + ((size_t)b * D + d0) * HW + pix : depth + (size_t)b * D + d0;
    const int dep_stride = PER_PIXEL ? kc.hw : 1;
#pragma unroll
    for (int k = 0; k < DC; ++k) acc_s[k][tid] = 0.0f;
    float wsum = 1e-5f;                                // TransMVSNet.py:72
    const size_t slice = (size_t)H * kc.row;
    const float xf = (float)x, yf = (float)y;
    const float *vw_p = vw + (size_t)b * n_src * vw_hw + (size_t)(y >> vw_shift) * vw_w + (x >> vw_shift);

    for (int i = 0; i < n_src; ++i) {
        float rt[12];
        tmvs_geom_rt(geom, i, bl, b_chunk, rt);
        const TmvsRay ray = tmvs_ray(rt, xf, yf, geom.ray_unfused);
        float tx = rt[9], ty = rt[10], tz = rt[11];
        asm volatile("" : "+f"(tx), "+f"(ty), "+f"(tz));
        const float wi = __ldg(vw_p + (size_t)i * vw_hw);
        const float4 *img = geom.img[i] + (size_t)b * slice;
        asm volatile("" : "+l"(img));
        auto emit = [&](int k, float s) {
            s *= kc.inv_c;                                                          // .mean(1), TransMVSNet.py:80
            acc_s[k][tid] = __fadd_rn(acc_s[k][tid], __fmul_rn(s, wi));             // TransMVSNet.py:88
        };
        // ---- how does this pixel's epipolar walk run in this view?  (first and last plane of the chunk)
        const float2 pa = tmvs_coords_lean<RECIP>(ray.rx, ray.ry, ray.rz, tx, ty, tz, __ldg(dep_base), kc);
        const float2 pb = tmvs_coords_lean<RECIP>(ray.rx, ray.ry, ray.rz, tx, ty, tz,
                                                  __ldg(dep_base + (size_t)(nd - 1) * dep_stride), kc);
        const float dx = pb.x - pa.x, dy = pb.y - pa.y;
        const bool xmajor = fabsf(dx) >= fabsf(dy);
        const float dM = xmajor ? dx : dy, dm = xmajor ? dy : dx;
        const int sgn = dM >= 0.0f ? 1 : -1;
        const int minor_up = dm >= 0.0f;                            // in plane order the minor coordinate grows
        // worth sweeping: at most ~1.2 columns per plane, and a walk no steeper than ~1:2 (three-pixel windows)
        const bool fits = fabsf(dM) <= 1.2f * (float)nd + 1.0f && fabsf(dm) <= 0.5f * fabsf(dM) + 1.0f;
        const bool uniform = __all_sync(lanes, fits) &&
                             (__all_sync(lanes, xmajor) || __all_sync(lanes, !xmajor)) &&
                             (__all_sync(lanes, sgn > 0) || __all_sync(lanes, sgn < 0)) &&
                             (__all_sync(lanes, minor_up) || __all_sync(lanes, !minor_up));
        if (uniform) {
            if (xmajor)
                sweep_view<C4T, true, RECIP, PER_PIXEL>(img, r, kc, ray.rx, ray.ry, ray.rz, tx, ty, tz, dep_base, dep_stride,
                                                        nd, sgn, minor_up, lanes, emit);
            else
                sweep_view<C4T, false, RECIP, PER_PIXEL>(img, r, kc, ray.rx, ray.ry, ray.rz, tx, ty, tz, dep_base, dep_stride,
                                                         nd, sgn, minor_up, lanes, emit);
        } else {
            const float *dep_p = dep_base;
            for (int k = 0; k < nd; ++k, dep_p += dep_stride)
                emit(k, plane_generic<C4T>(img, make_plane<RECIP>(ray.rx, ray.ry, ray.rz, tx, ty, tz, __ldg(dep_p), kc), r, kc));
        }
        wsum = __fadd_rn(wsum, wi);                                                 // TransMVSNet.py:89
    }
    float *out_a = agg + ((size_t)b * D + d0) * HW + pix;
    for (int k = 0; k < nd; ++k) __stcs(out_a + (size_t)k * HW, __fdiv_rn(acc_s[k][tid], wsum));   // :93
}

template <int C4T, int DC, bool PER_PIXEL, bool RECIP>
int launch_sweep(cudaStream_t st, const float *ref, int64_t rB, int64_t rC, int64_t rH, int64_t rW, const float *depth,
                 const float *vw, int vw_shift, int vw_w, int vw_hw, float *agg, int b_total, int b_first, int bc, int C,
                 int D, int H, int W, int n_src, const TmvsFwdConst &kc, const TmvsGeom &geom)
{
    const int n_dchunks = (D + DC - 1) / DC;
    {   // shared-memory carveout sized to the resident CTAs: the rest of the unified 228 KB is L1 for the gather
        static std::atomic<bool> done[64];
        int dev_id = 0;
        cudaGetDevice(&dev_id);
        if (dev_id < 0 || dev_id >= 64 || !done[dev_id].load(std::memory_order_acquire)) {
            const int pct = (SweepMinBlocks<C4T>::value * (DC * kThreads * 4 + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024);
            cudaFuncSetAttribute(costvol_fwd_sweep_kernel<C4T, DC, PER_PIXEL, RECIP>,
                                 cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            if (dev_id >= 0 && dev_id < 64) done[dev_id].store(true, std::memory_order_release);
        }
    }
    dim3 grid(((W + kTileX - 1) / kTileX) * n_dchunks, (H + kTileY - 1) / kTileY, bc), block(kTileX, kTileY);
    costvol_fwd_sweep_kernel<C4T, DC, PER_PIXEL, RECIP><<<grid, block, 0, st>>>(
        ref, rB, rC, rH, rW, depth, vw, vw_shift, vw_w, vw_hw, agg, b_total, b_first, bc, C, D, H, W, n_src, n_dchunks, kc,
        geom);
    return tmvs_launch_status();
}

}  // namespace

// Internal entry (called by costvol_fwd_impl, tmvs_costvol.cu).  Returns TMVS_E_UNSUPPORTED when the variant does not
// apply (per-view output wanted, C not 8 or 16), so the caller runs costvol_fwd_kernel.
int tmvs_costvol_fwd_sweep(const float *ref, int64_t rB, int64_t rC, int64_t rH, int64_t rW, const float *depth,
                           int per_pixel, const float *vw, int vw_shift, int vw_w, int vw_hw, float *agg, int b_total,
                           int b_first, int bc, int C, int D, int H, int W, int n_src, bool recip, const TmvsFwdConst &kc,
                           const TmvsGeom &geom, cudaStream_t st)
{
#define TMVS_SWEEP(C4T, DC)                                                                                             \
    do {                                                                                                                \
        if (per_pixel)                                                                                                  \
            return recip ? launch_sweep<C4T, DC, true, true>(st, ref, rB, rC, rH, rW, depth, vw, vw_shift, vw_w, vw_hw, \
                                                             agg, b_total, b_first, bc, C, D, H, W, n_src, kc, geom)    \
                         : launch_sweep<C4T, DC, true, false>(st, ref, rB, rC, rH, rW, depth, vw, vw_shift, vw_w, vw_hw,\
                                                              agg, b_total, b_first, bc, C, D, H, W, n_src, kc, geom);  \
        return recip ? launch_sweep<C4T, DC, false, true>(st, ref, rB, rC, rH, rW, depth, vw, vw_shift, vw_w, vw_hw,    \
                                                          agg, b_total, b_first, bc, C, D, H, W, n_src, kc, geom)       \
                     : launch_sweep<C4T, DC, false, false>(st, ref, rB, rC, rH, rW, depth, vw, vw_shift, vw_w, vw_hw,   \
                                                           agg, b_total, b_first, bc, C, D, H, W, n_src, kc, geom);     \
    } while (0)
    if (C == 16) { if (D > 8) TMVS_SWEEP(4, 16); else TMVS_SWEEP(4, 8); }
    if (C == 8) { if (D > 8) TMVS_SWEEP(2, 16); else TMVS_SWEEP(2, 8); }
#undef TMVS_SWEEP
    return TMVS_E_UNSUPPORTED;
}
