"""Generate golden vectors by running the REAL reference (read-only at /root/reference).

Run in the build container only:   python tests/golden/make_golden.py
The reference has no tests or fixtures of its own, so these files are the
pin for the oracle and for the CUDA path: inputs are seeded synthetic tensors, outputs come
from the reference's unmodified functions executed on the CPU by the installed torch:
  models.module.homo_warping            (models/module.py:284-322)
  models.module.depth_wta               (models/module.py:474-482)
  models.TransMVSNet.DepthNet.forward   (models/TransMVSNet.py:38-109), incl. PixelwiseNet
and their autograd.  Nothing from the reference is copied into the repo; only its outputs.
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

import warnings  # noqa: E402

warnings.filterwarnings("ignore")

from models.module import homo_warping, depth_wta  # noqa: E402  (reference)
from models.TransMVSNet import DepthNet  # noqa: E402  (reference)

from transmvsnet_b200 import synthetic  # noqa: E402
from transmvsnet_b200.geometry import compose_projection, relative_rot_trans  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)


def save(name, **arrays):
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = v
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **conv)
    print(f"{name:28s} {os.path.getsize(path) / 1024:8.1f} KiB")


def warp_with_grad(src, src_proj, ref_proj, depth, seed):
    """The reference's homo_warping and its autograd wrt src_fea (through F.grid_sample, models/module.py:318-320) for a
    seeded, arbitrary upstream gradient [B,C,D,H,W]."""
    s = src.clone().requires_grad_(True)
    out = homo_warping(s, src_proj, ref_proj, depth)
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed))
    (gsrc,) = torch.autograd.grad(out, s, gout)
    return out.detach(), gout, gsrc


def projections(pm):
    views = torch.unbind(pm, 1)
    return [compose_projection(v) for v in views]


def warp_cases():
    # (a) per-pixel hypotheses, DTU-like cameras, some samples out of bounds
    st = synthetic.make_stage(2, batch=2, n_views=3, height=32, width=48, channels=8, num_depth=8, seed=1)
    projs = projections(st.proj_matrix)
    for tag, depth in (("perpixel", st.depth_values), ("planes", st.depth_values[:, :, 3, 5].contiguous())):
        src = st.features[1]
        out, gout, gsrc = warp_with_grad(src, projs[1], projs[0], depth, 31)
        save(f"warp_{tag}", src=src, src_proj=projs[1], ref_proj=projs[0], depth=depth,
             rot_trans=relative_rot_trans(projs[1], projs[0]), out=out, grad_out=gout, grad_src=gsrc)
    # (b) identity warp: src_proj == ref_proj  =>  every depth plane reproduces the source
    src = st.features[2]
    out, gout, gsrc = warp_with_grad(src, projs[0], projs[0], st.depth_values, 32)
    save("warp_identity", src=src, src_proj=projs[0], ref_proj=projs[0], depth=st.depth_values,
         rot_trans=relative_rot_trans(projs[0], projs[0]), out=out, grad_out=gout, grad_src=gsrc)
    # (c) exact-integer sample coordinates: pure x/y translation, power-of-two depths
    b, c, h, w = 1, 4, 8, 12
    g = torch.Generator().manual_seed(3)
    src = torch.randn(b, c, h, w, generator=g)
    ref_p = torch.eye(4)[None]
    src_p = torch.eye(4)[None].clone()
    src_p[0, 0, 3] = 8.0
    src_p[0, 1, 3] = -16.0
    depth = torch.tensor([[1.0, 2.0, 4.0, 8.0, 16.0]])
    out, gout, gsrc = warp_with_grad(src, src_p, ref_p, depth, 33)
    save("warp_integer", src=src, src_proj=src_p, ref_proj=ref_p, depth=depth,
         rot_trans=relative_rot_trans(src_p, ref_p), out=out, grad_out=gout, grad_src=gsrc)
    # (d) points behind the source camera (z < 1e-6) and huge coordinates (z tiny positive)
    src_p = torch.eye(4)[None].clone()
    src_p[0, 2, 3] = -4.0          # z = depth - 4
    src_p[0, 0, 3] = 1.0
    depth = torch.tensor([[1.0, 3.999999, 4.0, 4.0000005, 4.000001, 4.5, 6.0, 1e-9]])
    out, gout, gsrc = warp_with_grad(src, src_p, ref_p, depth, 34)
    save("warp_behind", src=src, src_proj=src_p, ref_proj=ref_p, depth=depth,
         rot_trans=relative_rot_trans(src_p, ref_p), out=out, grad_out=gout, grad_src=gsrc)


class _Capture(torch.nn.Module):
    """Stand-in for the 3-D CNN: records the aggregated similarity, returns scaled logits."""

    def __init__(self, gain):
        super().__init__()
        self.gain = gain
        self.seen = None

    def forward(self, x):
        self.seen = x
        return x * self.gain


def depthnet_cases():
    specs = [
        # name, stage, C, D, (H, W) full-res image, views, batch, with given view weights?
        ("depthnet_s1_learned", 1, 32, 48, (48, 80), 4, 1, False),
        ("depthnet_s1_given", 1, 32, 48, (48, 80), 4, 1, True),
        ("depthnet_s2_given", 2, 16, 32, (40, 56), 4, 2, True),
        ("depthnet_s3_given", 3, 8, 8, (24, 40), 7, 1, True),
        ("depthnet_odd_given", 2, 12, 5, (22, 38), 3, 2, True),   # ragged: C%8!=0, odd sizes
    ]
    for name, stage, c, d, (hh, ww), n, b, given in specs:
        torch.manual_seed(11)
        st = synthetic.make_stage(stage, batch=b, n_views=n, height=hh, width=ww, channels=c,
                                  num_depth=d, seed=5)
        feats = [f.clone().requires_grad_(True) for f in st.features]
        net = DepthNet().eval()
        cap = _Capture(gain=25.0)
        if given:
            out = net(feats, st.proj_matrix, st.depth_values, d, cap, view_weights=st.view_weights)
            vw = st.view_weights
        else:
            out, vw = net(feats, st.proj_matrix, st.depth_values, d, cap, view_weights=None)
        sim = cap.seen                                       # [B,1,D,h,w]
        # backward of the cost volume wrt the features, through the reference's autograd
        gsim = torch.randn(sim.shape, generator=torch.Generator().manual_seed(23))
        grads = torch.autograd.grad(sim, feats, gsim, retain_graph=False, allow_unused=True)
        projs = projections(st.proj_matrix)
        rts = torch.stack([relative_rot_trans(p, projs[0]) for p in projs[1:]], 0)
        arrays = dict(
            features=torch.stack([f.detach() for f in feats], 0), proj_matrix=st.proj_matrix,
            depth_values=st.depth_values, view_weights=vw.detach(), rot_trans=rts,
            similarity=sim.detach(), prob_volume=out["prob_volume"].detach(), depth=out["depth"].detach(),
            photo_confidence=out["photo_confidence"].detach(),
            index=torch.argmax(out["prob_volume"], dim=1).detach(), grad_similarity=gsim,
            gain=np.float32(25.0))
        if given:
            arrays["grad_features"] = torch.stack([g for g in grads], 0)
        else:
            sd = net.pixel_wise_net.state_dict()
            for k, v in sd.items():
                arrays["pwn." + k] = v
        save(name, **arrays)


def readout_cases():
    g = torch.Generator().manual_seed(31)
    # depth_wta on a probability volume with exact ties and a NaN-free plateau
    p = torch.rand(2, 6, 5, 7, generator=g)
    p[:, 1] = p[:, 4]                                  # ties between planes 1 and 4 wherever they win
    p[0, :, 0, 0] = 0.25                               # full plateau -> index 0
    p[1, 5, 2, 3] = 2.0
    p[1, 2, 2, 3] = 2.0                                # tie of the maximum -> first (2)
    dv = 400.0 + 10.0 * torch.rand(2, 6, 5, 7, generator=g).cumsum(1)
    out = depth_wta(p, dv)
    idx = torch.argmax(p, dim=1)
    save("wta_ties", p=p, depth_values=dv, depth=out, index=idx)


def regression_case():
    # depth_regression does not exist in this fork.  The vector below is produced by the
    # upstream 3-line definition, NOT by code under /root/reference -> "parity unpinned".
    g = torch.Generator().manual_seed(41)
    p = torch.softmax(3 * torch.randn(2, 9, 6, 10, generator=g), 1)
    dv4 = 425.0 + 2.5 * torch.arange(9, dtype=torch.float32)[None, :, None, None] + torch.rand(2, 9, 6, 10, generator=g)
    dv2 = 425.0 + 7.5 * torch.arange(9, dtype=torch.float32)[None].repeat(2, 1)
    save("regression_unpinned", p=p, depth_values_4d=dv4, depth_values_2d=dv2,
         depth_4d=torch.sum(p * dv4, 1), depth_2d=torch.sum(p * dv2.view(2, 9, 1, 1), 1))


def hypotheses_case():
    """Stage hypotheses through the reference's own ops: F.interpolate (bilinear) -> get_depth_samples
    (models/module.py:606-634) -> F.interpolate (trilinear), exactly as models/TransMVSNet.py:174-190, 202-204."""
    import torch.nn.functional as F
    from models.module import get_depth_samples
    g = torch.Generator().manual_seed(51)
    b, h_img, w_img = 2, 48, 64
    dv = (425.0 + 2.5 * torch.arange(192, dtype=torch.float32))[None].repeat(b, 1)
    dv[1] += 7.0
    interval = float((dv[0, -1] - dv[0, 0]) / 192)
    arrays = {"depth_values": dv, "image_hw": np.array([h_img, w_img]), "depth_interval": np.float32(interval)}
    prev = {2: 500 + 300 * torch.rand(b, h_img // 4, w_img // 4, generator=g),
            3: 500 + 300 * torch.rand(b, h_img // 2, w_img // 2, generator=g)}
    for stage, (nd, ratio, scale) in enumerate(((48, 4.0, 4), (32, 1.0, 2), (8, 0.5, 1)), start=1):
        if stage == 1:
            cur = dv
        else:
            cur = F.interpolate(prev[stage].unsqueeze(1), [h_img, w_img], mode="bilinear", align_corners=False).squeeze(1)
            arrays[f"prev{stage}"] = prev[stage]
        samples = get_depth_samples(cur_depth=cur, ndepth=nd, depth_inteval_pixel=ratio * interval, dtype=torch.float32,
                                    device=torch.device("cpu"), shape=[b, h_img, w_img])
        out = F.interpolate(samples.unsqueeze(1), [nd, h_img // scale, w_img // scale], mode="trilinear",
                            align_corners=False).squeeze(1)
        arrays[f"hyp{stage}"] = out
    save("hypotheses", **arrays)


def finalize_case():
    """test.py:126-144 (cv2.resize of the stage-1/2 confidences, product, depth[conf < 0.01] = 0) followed by the
    reference's utils.depth_normal -- the numpy/cv2 statements of save_depth, run as they stand."""
    import cv2
    from utils import depth_normal      # reference utils.py:11-21
    g = torch.Generator().manual_seed(61)
    b, h, w = 2, 40, 56
    depth = (400 + 560 * torch.rand(b, h, w, generator=g)).numpy()          # some beyond [425, 935]
    conf3 = torch.rand(b, h, w, generator=g).numpy() ** 2
    conf2 = torch.rand(b, h // 2, w // 2, generator=g).numpy()
    conf1 = torch.rand(b, h // 4, w // 4, generator=g).numpy()
    d_out, c_out, a_out = [], [], []
    for i in range(b):
        c1 = cv2.resize(conf1[i], (w, h))
        c2 = cv2.resize(conf2[i], (w, h))
        conf_final = conf3[i] * c1 * c2
        d = depth[i].copy()
        d[conf_final < 0.01] = 0.0
        d_out.append(d.copy())
        c_out.append(conf_final)
        a_out.append(depth_normal(d, depth_min=425.0, depth_max=935.0))
    save("finalize", depth=depth, conf3=conf3, conf2=conf2, conf1=conf1, depth_out=np.stack(d_out),
         conf_out=np.stack(c_out), alpha_out=np.stack(a_out))


if __name__ == "__main__":
    finalize_case()
    hypotheses_case()
    warp_cases()
    depthnet_cases()
    readout_cases()
    regression_case()
