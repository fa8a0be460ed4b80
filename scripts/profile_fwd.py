"""One launch of each forward cost-volume variant at the config-2 stage sizes, for ncu:
   ncu --metrics <...> -k regex:costvol_fwd python scripts/profile_fwd.py [stages]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from transmvsnet_b200 import _lib, ops, pipeline, synthetic  # noqa: E402

dev = torch.device("cuda:0")
stages = [int(a) for a in sys.argv[1:]] or [2, 3]
for stage in stages:
    st = synthetic.make_stage(stage, batch=1, n_views=5, height=1152, width=1600, seed=0)
    d = pipeline.stage_to_device(st, dev)
    packed = ops.pack_sources(d["features"][1:])
    for bits in (0, _lib.F_FWD_SWEEP):
        for _ in range(2):      # the second launch of each is the warm one
            with ops.extra_flags(bits):
                ops.cost_volume_packed(d["features"][0], packed, d["rot_trans"], d["depth_values"], d["view_weights"], False, True)
    torch.cuda.synchronize()
