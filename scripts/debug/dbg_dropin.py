import sys, copy, torch, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import build as ob
from transmvsnet_b200 import DepthNet, synthetic, ops
from conftest import rel_err
import test_gpu_dropin_reference as T
DEV = "cuda:0"
mod, net = ob.import_reference()
torch.manual_seed(2)
model = net.TransMVSNet().to(DEV).eval()
imgs, proj, dv = T._inputs(seed=3)
with torch.no_grad():
    calls, _ = T._capture_depthnet_calls(model, imgs, proj, dv)
model.train()
ours = DepthNet().to(DEV).train(); ours.load_state_dict(model.DepthNet.state_dict())
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    print("=== cudnn tf32", tf32)
    for stage, (kw, _) in enumerate(calls, start=1):
        res = []
        for name, depthnet in (("ref", model.DepthNet), ("ref again", model.DepthNet), ("ours", ours)):
            reg = copy.deepcopy(kw["cost_regularization"]).train()
            feats = [f.clone().requires_grad_(True) for f in kw["features"]]
            seen = {}
            def hook(m, inp, out, seen=seen): seen["sim"] = inp[0].detach().clone()
            h = reg.register_forward_hook(hook)
            out = depthnet(feats, kw["proj_matrix"], depth_values=kw["depth_values"], num_depth=kw["num_depth"],
                           cost_regularization=reg, view_weights=kw["view_weights"])
            h.remove()
            vw = None
            if isinstance(out, tuple): out, vw = out
            g = torch.Generator(device=DEV).manual_seed(7)
            loss = (out["prob_volume"] * torch.randn(out["prob_volume"].shape, device=DEV, generator=g)).sum()
            grads = torch.autograd.grad(loss, feats)
            res.append((name, seen["sim"], vw, out["prob_volume"].detach(), grads))
        base = res[0]
        for name, sim, vw, prob, grads in res[1:]:
            line = f"stage {stage} {name}: sim {rel_err(sim.cpu().numpy(), base[1].cpu().numpy())[0]:.2e} prob {float((prob-base[3]).abs().max()):.2e}"
            if vw is not None: line += f" vw {float((vw-base[2]).abs().max()):.2e}"
            line += " grads " + " ".join(f"{rel_err(a.cpu().numpy(), b.cpu().numpy())[0]:.1e}" for a, b in zip(grads, base[4]))
            print(line)
