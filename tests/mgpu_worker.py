"""Worker of tests/test_gpu_multigpu.py: run under torchrun with WORLD_SIZE ranks, one GPU each.

Every rank takes its contiguous shard of a synthetic batch (uneven on purpose: the shards differ in frames and in boxes), runs the CUDA
path on it, all-gathers the top-K records through the library's NCCL communicator (cs_allgather_topk), fetches the WHOLE gathered buffer
(cs_fetch_gathered) and compares every rank's slice with the CPU oracle run on that rank's frames.  Exit code 0 = all slices agree."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import cube_slam_b200 as cs
    from cube_slam_b200 import _lib, sharding
    from cube_slam_b200 import synthetic as S
    from oracle import pyoracle as O
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    F, topk = 7, 2  # 7 frames over 2 ranks: 4 + 3
    imgs, Ts, boxes, lines, K = S.make_batch(4242, F, 640, 480, 3, poisson=True)
    boxes[F - 1] = np.zeros((0, 5))  # a frame without boxes at the end of the last shard
    nb = [len(b) for b in boxes]
    lo, hi = sharding.shard_range(F, world, rank)
    p = cs.default_params(max_cuboid_num=topk)
    ctx = cs.Context(local, 640, 480, hi - lo, 16, 4096)
    ctx.set_calibration(K)
    nccl_path = ""
    for d in sys.path:
        cand = os.path.join(d, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            nccl_path = cand
            break
    uid = np.zeros(128, np.uint8)
    if rank == 0:
        ctx.check(ctx.L.cs_comm_unique_id(ctx.h, nccl_path.encode(), _lib.ptr(uid, C.c_uint8)))
    t = torch.from_numpy(uid).cuda()
    dist.broadcast(t, 0)
    uid = t.cpu().numpy()
    ctx.check(ctx.L.cs_comm_init(ctx.h, nccl_path.encode(), _lib.ptr(uid, C.c_uint8), world, rank))
    recs_per_rank = sharding.records_per_rank(nb, world, topk)
    ctx.upload(imgs[lo:hi], Ts[lo:hi], boxes[lo:hi], lines[lo:hi], p)
    bad = 0
    for rep in range(2):  # twice: the second gather reuses the buffers of the first
        ctx.run_async()
        gathered = C.c_void_p()
        ctx.check(ctx.L.cs_allgather_topk(ctx.h, recs_per_rank, C.byref(gathered)))
        buf = np.zeros((world, recs_per_rank), cs.CUBOID_DTYPE)
        ctx.check(ctx.L.cs_fetch_gathered(ctx.h, buf.ctypes.data_as(C.c_void_p), world * recs_per_rank))
        per_frame = sharding.unpack_gathered(buf, nb, world, topk)
        for f in range(F):
            ref = O.detect_cuboid(imgs[f], K, Ts[f], boxes[f], lines[f], O.default_params(max_cuboid_num=topk))
            for b in range(nb[f]):
                got = per_frame[f][b]
                want = ref["cuboids"][b]
                n_valid = int((got["valid"] == 1).sum())
                if n_valid != len(want):
                    print("rank %d: frame %d box %d: %d gathered records, oracle %d" % (rank, f, b, n_valid, len(want)), flush=True)
                    bad += 1
                    continue
                for k in range(len(want)):
                    if int(got[k]["proposal_index"]) != int(want[k]["proposal_index"]) or float(got[k]["normalized_error"]) != float(want[k]["normalized_error"]) \
                            or not np.allclose(got[k]["pos"], want[k]["pos"], rtol=1e-9, atol=1e-9):
                        print("rank %d: frame %d box %d top%d differs" % (rank, f, b, k), flush=True)
                        bad += 1
        # the padding past a rank's own records must read valid == 0 on every rank
        for r in range(world):
            rlo, rhi = sharding.shard_range(F, world, r)
            used = sum(nb[rlo:rhi]) * topk
            if (buf[r, used:]["valid"] != 0).any():
                print("rank %d: padding of rank %d's slice is not clear" % (rank, r), flush=True)
                bad += 1
    tot = torch.tensor([bad], device="cuda")
    dist.all_reduce(tot)
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("mgpu_worker: %d mismatches over %d ranks" % (int(tot.item()), world), flush=True)
    sys.exit(1 if int(tot.item()) else 0)


if __name__ == "__main__":
    main()
