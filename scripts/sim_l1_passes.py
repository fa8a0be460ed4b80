"""CPU model of the L1 cost of candidate lane mappings for the cost-volume gather, on the synthetic config-2 geometry.

Pass rule measured by scripts/l1_peak.cu: a 128-bit warp load is served in four passes of 8 consecutive lanes, one
clock each only if the 8 addresses fall inside one contiguous 128-byte window (here: <= 8 consecutive pixels of one
source row).  For each sampled (view, row, 32-pixel run, plane) the script counts the passes of the NW-tap request under
   A_8x1          one lane per reference pixel (the kernel's mapping),
   pair_4x2taps   even lane = west column, odd lane = east column of the footprint,
   B_4px2planes   4 pixels x 2 consecutive planes per pass.
Result (passes per 32 pixel-taps, ideal 4): stage 1  4.41 / 4.11 / 7.18;  stage 2  7.86 / 5.43 / 7.60;  stage 3
7.91 / 5.48 / 7.81.  The model matches the measured 4.46 wavefronts per request at stage 1 but OVERESTIMATES stages 2/3
(measured 5.9), and the pair mapping it favours measured slower in the real kernel with the same wavefront total
(profiles/r1_costvol_pair_ncu_summary.txt): the rule is necessary, not sufficient.  Kept as the starting point for the
next mapping experiment.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transmvsnet_b200 import synthetic, geometry
def passes_quarter(xs, ys):
    # xs, ys: integer source pixel coords of 8 lanes. model: one pass per distinct (row, 128B window); window = 8 contiguous px (any alignment)
    n = 0
    for r in np.unique(ys):
        x = np.sort(xs[ys == r]); 
        i = 0
        while i < len(x):
            j = i
            while j + 1 < len(x) and x[j+1] - x[i] <= 7: j += 1
            n += 1; i = j + 1
    return n
for stage in (1, 2, 3):
    st = synthetic.make_stage(stage, batch=1, n_views=5, height=1152, width=1600, seed=0)
    rt = geometry.stage_rot_trans(st.proj_matrix).double().numpy()   # [N,B,12]
    dv = st.depth_values[0].double().numpy()                          # [D,h,w]
    D, h, w = dv.shape
    rng = np.random.default_rng(0)
    res = {"A_8x1": [], "pair_4x2taps": [], "B_4px2planes": []}
    for _ in range(300):
        v = rng.integers(0, 4); y = rng.integers(0, h); x0 = rng.integers(0, w // 32) * 32; d = rng.integers(0, D - 1)
        R = rt[v, 0, :9].reshape(3, 3); t = rt[v, 0, 9:]
        xs = np.arange(x0, x0 + 32)
        def proj(dd):
            p = (R @ np.stack([xs, np.full(32, y), np.ones(32)])) * dv[dd, y, xs] + t[:, None]
            return p[0] / p[2], p[1] / p[2]
        ix, iy = proj(d); ix2, iy2 = proj(d + 1)
        X0, Y0 = np.floor(ix).astype(int), np.floor(iy).astype(int)
        X02, Y02 = np.floor(ix2).astype(int), np.floor(iy2).astype(int)
        # mapping A: request = NW tap of 32 px (row Y0); 4 quarters
        a = sum(passes_quarter(X0[q*8:(q+1)*8], Y0[q*8:(q+1)*8]) for q in range(4))
        res["A_8x1"].append(a / 32)                  # passes per pixel-tap (NW); NE same statistics
        # pair: request covers 16 px x (NW, NE): quarter = 4 px x 2 taps ; per pixel-tap: passes / (16*2)
        xs16 = np.repeat(X0[:16], 2) + np.tile([0, 1], 16); ys16 = np.repeat(Y0[:16], 2)
        p = sum(passes_quarter(xs16[q*8:(q+1)*8], ys16[q*8:(q+1)*8]) for q in range(4))
        res["pair_4x2taps"].append(p / 32)
        # B: 16 px x 2 planes (NW tap): quarter = 4 px x 2 planes
        xb = np.stack([X0[:16], X02[:16]], 1).reshape(8, 4) ; 
        xsB = np.concatenate([np.concatenate([X0[4*g:4*g+4], X02[4*g:4*g+4]]) for g in range(4)])
        ysB = np.concatenate([np.concatenate([Y0[4*g:4*g+4], Y02[4*g:4*g+4]]) for g in range(4)])
        b = sum(passes_quarter(xsB[q*8:(q+1)*8], ysB[q*8:(q+1)*8]) for q in range(4))
        res["B_4px2planes"].append(b / 32)
    print("stage", stage, {k: round(float(np.mean(v)) * 32, 2) for k, v in res.items()}, "(passes per 32 pixel-taps; ideal 4)")
