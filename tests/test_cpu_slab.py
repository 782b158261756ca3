"""CPU tier: slab partition logic and the halo exchange routine (world size 2, gloo)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _model():
    from spatialpy_b200 import configs
    return configs.tank_sdpd(n=14, nt=10, output_every=10, dt=2e-5)


def test_partition_covers_domain_and_neighbourhoods():
    from scipy.spatial import cKDTree
    from spatialpy_b200.slab import partition
    fm = _model()
    world = 3
    parts = [partition(fm, r, world) for r in range(world)]
    owned_g = np.concatenate([p.gids[p.owned == 1] for p in parts])
    assert sorted(owned_g.tolist()) == list(range(fm.num_particles))          # every particle owned exactly once
    tree = cKDTree(fm.x)
    for r, p in enumerate(parts):
        local = set(p.gids.tolist())
        # every owned particle finds its complete candidate neighbourhood (h * 1.1) among owned + ghosts
        for g in p.gids[p.owned == 1][::7]:
            assert set(tree.query_ball_point(fm.x[g], fm.h * 1.1)) <= local
        # exchange lists are mirror images: what I send to nb is what nb receives from me, in the same (global id) order
        for nb, ids in p.send_ids.items():
            q = parts[nb]
            np.testing.assert_array_equal(p.gids[ids], q.gids[q.recv_ids[r]])
            assert (p.owned[ids] == 1).all() and (q.owned[q.recv_ids[r]] == 0).all()
        np.testing.assert_array_equal(p.local.x, fm.x[p.gids])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spatialpy_b200.slab import exchange, partition
    fm = _model()
    p = partition(fm, rank, world)
    # ship the global ids of my boundary particles; the neighbour must receive exactly the ids of its ghosts
    send = {nb: torch.as_tensor(p.gids[ids].astype(np.float64)) for nb, ids in p.send_ids.items()}
    recv = {nb: torch.empty(len(ids), dtype=torch.float64) for nb, ids in p.recv_ids.items()}
    exchange(send, recv, rank)
    ok = all(np.array_equal(recv[nb].numpy().astype(np.int64), p.gids[ids]) for nb, ids in p.recv_ids.items())
    q.put((rank, ok, p.n_owned, len(p.gids)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_halo_exchange_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(o[1] for o in out)
    assert out[0][2] + out[1][2] == _model().num_particles
    assert all(o[3] > o[2] for o in out)          # both ranks carry ghosts
