"""World-size-2 gloo test (CPU) of the reference-view sharding and the depth/confidence gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from transmvsnet_b200 import sharding


def test_shard_views_partition():
    for n, w in ((49, 8), (5, 2), (3, 4), (0, 2), (16, 1)):
        shards = [sharding.shard_views(n, r, w) for r in range(w)]
        assert sorted(v for s in shards for v in s) == list(range(n))
        assert max(map(len, shards)) - min(map(len, shards)) <= 1
        assert [len(s) for s in shards] == sharding.views_per_rank(n, w)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_views, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_views(n_views, rank, world)
        # "depth" map of view v is filled with v, "confidence" with v + 0.5
        local = torch.stack([torch.stack([torch.full((3, 4), float(v)), torch.full((3, 4), v + 0.5)]) for v in mine]) \
            if mine else torch.zeros(0, 2, 3, 4)
        out = sharding.gather_maps(local, n_views, dst=0)
        if rank == 0:
            ok = out.shape == (n_views, 2, 3, 4)
            for v in range(n_views):
                ok = ok and bool((out[v, 0] == v).all()) and bool((out[v, 1] == v + 0.5).all())
            ret[rank] = ok
        else:
            ret[rank] = out is None
    finally:
        dist.destroy_process_group()


def test_gather_maps_world2_gloo():
    for n_views in (5, 4):
        mgr = mp.Manager()
        ret = mgr.dict()
        port = _free_port()
        mp.spawn(_worker, args=(2, port, n_views, ret), nprocs=2, join=True)
        assert ret[0] is True and ret[1] is True
