"""N > 1 host logic on CPU: world-size-2 gloo run of the frame-sharded path (shard -> per-rank top-K -> one all-gather),
with the oracle standing in for the per-rank compute, checked against the single-process result."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_range_partitions():
    from cube_slam_b200.sharding import shard_range
    for n in [0, 1, 7, 8, 256, 1023]:
        for world in [1, 2, 3, 8]:
            cover = []
            for r in range(world):
                lo, hi = shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                cover += list(range(lo, hi))
            assert cover == list(range(n))
            sizes = [shard_range(n, world, r)[1] - shard_range(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from cube_slam_b200 import synthetic as S
    from cube_slam_b200.sharding import pad_records, records_per_rank, shard_range, unpack_gathered
    from oracle import pyoracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    imgs, Ts, boxes, lines, K = S.make_batch(77, 5, 320, 240, 2, poisson=False)
    nb = [len(b) for b in boxes]
    topk = 2
    p = O.default_params(max_cuboid_num=topk)
    lo, hi = shard_range(len(imgs), world, rank)
    mine = []
    for f in range(lo, hi):
        r = O.detect_cuboid(imgs[f], K, Ts[f], boxes[f], lines[f], p)
        rec = np.zeros((nb[f], topk), O.CUBOID_DTYPE)
        for b, cl in enumerate(r["cuboids"]):
            rec[b, :len(cl)] = cl
        mine.append(rec)
    slots = records_per_rank(nb, world, topk)
    flat = pad_records(np.concatenate([m.reshape(-1) for m in mine]) if mine else np.zeros(0, O.CUBOID_DTYPE), slots)
    send = torch.from_numpy(flat.view(np.uint8).copy())
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    gathered = np.stack([t.numpy().view(O.CUBOID_DTYPE) for t in recv])
    per_frame = unpack_gathered(gathered, nb, world, topk)
    if rank == 0:
        ok = True
        for f in range(len(imgs)):
            r = O.detect_cuboid(imgs[f], K, Ts[f], boxes[f], lines[f], p)
            for b, cl in enumerate(r["cuboids"]):
                got = per_frame[f][b]
                ok &= int(got["valid"].sum()) == len(cl)
                for k in range(len(cl)):
                    ok &= int(got[k]["proposal_index"]) == int(cl[k]["proposal_index"])
                    ok &= float(got[k]["normalized_error"]) == float(cl[k]["normalized_error"])
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allgather_matches_single_process(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
