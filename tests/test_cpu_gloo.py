"""CPU tier: the N>1 path (ensemble sharding + max/sum over ranks) on world_size 2 with the gloo backend."""
import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spatialpy_b200.ensemble import reduce_scalar, shard_trajectories
    mine = shard_trajectories(11, rank, world)
    # every rank derives the same global seed for trajectory k (seed + k), whatever the world size
    seeds = [1000 + k for k in mine]
    t_max = reduce_scalar(10.0 + rank, "max")
    n_sum = reduce_scalar(float(len(mine)), "sum")
    q.put((rank, mine, seeds, t_max, n_sum))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ensemble_bookkeeping():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(out[0][1] + out[1][1]) == list(range(11))
    assert out[0][2] == [1000 + k for k in out[0][1]]
    assert out[0][3] == out[1][3] == 11.0            # max over ranks
    assert out[0][4] == out[1][4] == 11.0            # total trajectories
