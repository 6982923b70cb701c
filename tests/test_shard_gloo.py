"""N>1 host logic on CPU (gloo, world_size 2): every rank derives the same LPT partition from the index pass, takes its
own entries, and the union over ranks is an exact, balanced cover -- with no data-path collective (the all_gather here is
the TEST's check, not part of the path)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("portable-network-archive_b200.shard")
    host = importlib.import_module("portable-network-archive_b200._host")
    # index pass (C++ host layer, no GPU needed) over a golden archive + a synthetic size list
    buf = np.fromfile(os.path.join(ROOT, "tests", "golden", "ref", "zstd_with_raw_file_size.pna"), dtype=np.uint8)
    ents = [e for e in host.HostArchive(buf).entries() if e["kind"] == 0 and e["data_kind"] == 0]
    rng = np.random.default_rng(5)
    weights = [e["compressed_size"] for e in ents] + [int(x) for x in rng.integers(1, 1 << 22, 1000)]
    mine = shard.rank_entries(weights, rank, world)
    # gather for the check
    t = torch.full((len(weights),), -1, dtype=torch.int64)
    t[torch.tensor(mine, dtype=torch.int64)] = rank
    got = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(got, t)
    owner = torch.stack(got).max(dim=0).values
    claimed = torch.stack([(g >= 0).to(torch.int64) for g in got]).sum(dim=0)
    loads = [sum(weights[i] for i in range(len(weights)) if int(owner[i]) == r) for r in range(world)]
    ok = bool((claimed == 1).all()) and mine == sorted(mine) and max(loads) - min(loads) <= max(weights)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok, loads))


def test_entry_sharding_world2_gloo(built):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res


def test_lpt_partition_properties():
    shard = importlib.import_module("portable-network-archive_b200.shard")
    for world in (1, 2, 4, 8):
        w = [4 << 20] * 8192
        parts = shard.lpt_partition(w, world)
        assert sorted(i for p in parts for i in p) == list(range(8192))
        assert {len(p) for p in parts} == {8192 // world}          # cfg2: 1024 entries per GPU at 8 GPUs
    assert shard.lpt_partition([], 3) == [[], [], []]
    assert shard.lpt_partition([5, 1, 1, 1, 1, 1], 2) == [[0], [1, 2, 3, 4, 5]]
