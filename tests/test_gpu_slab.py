"""GPU tier, 2 GPUs (skipped on a 1-GPU box): a slab-decomposed run over NCCL reproduces the single-GPU engine."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _model(reactive):
    from spatialpy_b200 import configs
    fm = configs.tank_sdpd(n=18, nt=30, output_every=30, dt=2e-5)
    rng = np.random.default_rng(3)
    fm.x = fm.x + rng.uniform(-0.004, 0.004, size=fm.x.shape)
    if not reactive:
        fm.parameters = {"P0": 0.0, "P1": 0.0}        # pure diffusion: the molecule count is conserved exactly
    return fm


def _worker(rank, world, port, q, steps):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from spatialpy_b200.slab import SlabEngine, partition
    fm = _model(False)
    part = partition(fm, rank, world)
    se = SlabEngine(part, rank, world, device=rank)
    se.reset(11)
    se.step(steps)
    out = {"gid": part.gids[part.owned == 1]}
    for f in ("x", "v", "rho", "F", "bvf_phi", "C", "xx"):
        out[f] = se.owned_field(f)[1]
    out["counters"] = se.eng.counters()
    q.put((rank, out))
    dist.barrier()
    se.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpu_slab_matches_single_gpu():
    import torch.multiprocessing as mp
    from spatialpy_b200.engine import Engine
    steps = 22                                     # through the Shepard filter at steps 0 and 20
    fm = _model(False)
    with Engine(fm, device=0) as eng:
        eng.reset(11)
        eng.step(steps)
        ref = {f: eng.get(f) for f in ("x", "v", "rho", "F", "bvf_phi", "C", "xx")}
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, steps)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    gid = np.concatenate([outs[r]["gid"] for r in range(2)])
    assert sorted(gid.tolist()) == list(range(fm.num_particles))
    for f in ("x", "v", "rho", "F", "bvf_phi", "C"):
        got = np.empty_like(ref[f])
        got[gid] = np.concatenate([outs[r][f] for r in range(2)])
        scale = max(float(np.abs(ref[f]).max()), 1e-300)
        err = float(np.abs(got - ref[f]).max()) / scale
        assert err <= 1e-9, f"{f}: {err:.3e}"          # same physics; only the summation order inside a sweep differs
    total = sum(int(outs[r]["xx"].sum()) for r in range(2))
    assert total == int(fm.u0.sum())                  # molecules that crossed the slab face were delivered, none lost
    assert sum(outs[r]["counters"]["diffusions"] for r in range(2)) > 0
