"""Multi-GPU check of sharding.PeerMapSink (run under torchrun on >= 2 GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_peer_sink.py

Every rank runs a small cascade twice: once with the read-out kernel writing the stage-3 maps into its slot of rank 0's
buffer over NVLink peer memory, once locally; the local maps then travel to rank 0 through the NCCL all_gather of
sharding.gather_maps and must equal what the kernels stored through the peer mapping, bit for bit."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from transmvsnet_b200 import pipeline, sharding, synthetic  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
stages = synthetic.make_cascade(batch=1, n_views=4, height=256, width=384, seed=100 + rank)
dev_stages = [pipeline.stage_to_device(s, dev) for s in stages]
h, w = stages[-1].depth_values.shape[2:]
sink = sharding.PeerMapSink(2 * world, (h, w), dev, dst=0)
for _ in range(3):                                     # repeated writes into the same slot
    pipeline.run_cascade(dev_stages, out_maps=sink.slot(rank))
local_out = pipeline.run_cascade(dev_stages)[-1]
# the copy-engine transport into the second half of the buffer (PeerMapSink.push on a side stream)
side = torch.cuda.Stream()
ready = torch.cuda.Event()
ready.record()
side.wait_event(ready)
sink.push(world + rank, local_out["depth"], local_out["photo_confidence"], side)
side.synchronize()
maps = torch.stack([local_out["depth"], local_out["photo_confidence"]], 1)          # [1,2,H,W]
torch.cuda.synchronize()
dist.barrier()
gathered = sharding.gather_maps(maps, world)
ok = torch.ones(1, device=dev)
if rank == 0:
    got = sink.result()
    same = torch.equal(got[:world], gathered) and torch.equal(got[world:], gathered)
    print(f"peer stores / DMA push vs NCCL gather on {world} GPUs: {'bit-identical' if same else 'MISMATCH'}; "
          f"slot means {[round(float(got[r].mean()), 3) for r in range(world)]}", flush=True)
    ok[0] = 1.0 if same else 0.0
dist.broadcast(ok, 0)
dist.destroy_process_group()
sys.exit(0 if float(ok.item()) == 1.0 else 1)
