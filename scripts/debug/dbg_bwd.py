import sys, copy, torch, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import build as ob, torch_port
from transmvsnet_b200 import DepthNet, synthetic, ops, geometry, _lib
from conftest import rel_err
import test_gpu_dropin_reference as T
DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = False
mod, net = ob.import_reference()
torch.manual_seed(2)
model = net.TransMVSNet().to(DEV).eval()
imgs, proj, dv = T._inputs(seed=3)
with torch.no_grad():
    calls, _ = T._capture_depthnet_calls(model, imgs, proj, dv)
for stage in (2, 3):
    kw, _ = calls[stage - 1]
    feats = kw["features"]; pm = kw["proj_matrix"]; dvv = kw["depth_values"]; vw = kw["view_weights"]
    b, d, h, w = dvv.shape
    g = torch.Generator(device=DEV).manual_seed(5)
    G = torch.randn(b, d, h, w, device=DEV, generator=g)
    # reference autograd (stock CUDA ops)
    fs = [f.clone().requires_grad_(True) for f in feats]
    agg, _ = torch_port.cost_volume(fs, pm, dvv, vw)
    gr = torch.autograd.grad(agg.squeeze(1), fs, G)
    rt = geometry.stage_rot_trans(pm)
    for name, extra in (("cells", 0), ("scan", _lib.F_BWD_SCAN)):
        fs2 = [f.clone().requires_grad_(True) for f in feats]
        with ops.extra_flags(extra):
            agg2, _ = ops.cost_volume(fs2[0], fs2[1:], rt, dvv, vw)
            go = torch.autograd.grad(agg2, fs2, G)
        print(f"stage {stage} {name}: fwd {rel_err(agg2.detach().cpu().numpy(), agg.detach().squeeze(1).cpu().numpy())[0]:.2e} grads",
              " ".join(f"{rel_err(a.cpu().numpy(), b_.cpu().numpy())[0]:.1e}/{rel_err(a.cpu().numpy(), b_.cpu().numpy())[1]:.1e}" for a, b_ in zip(go, gr)))
    # where is the error? for view 1 (first source)
    diff = (go[1] - gr[1]).abs().amax(1)[0]      # [H,W]
    thr = 1e-3 * float(gr[1].abs().max())
    bad = (diff > thr)
    print("   bad source pixels:", int(bad.sum()), "of", bad.numel(), "rows of first few:", bad.nonzero()[:8].tolist())
    # multiplicity statistics: how rough is the depth map
    print("   depth_values plane0 std of horizontal neighbours diff:", float((dvv[:, 0, :, 1:] - dvv[:, 0, :, :-1]).abs().mean()))
