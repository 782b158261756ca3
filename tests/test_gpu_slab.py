"""GPU tier: a slab-decomposed run reproduces the single-GPU engine — ranks as processes on 2 GPUs (windows shared through CUDA
IPC; skipped on a 1-GPU box) and ranks as threads on one GPU (windows shared by pointer)."""
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


# ----------------------------------------------------------------------------------------------------------------------
# state hand-over and re-partition (ssb_set_field / ssb_set_step, SlabEngine.repartition)
# ----------------------------------------------------------------------------------------------------------------------
_FIELDS = ("x", "v", "rho", "F", "bvf_phi", "C")


def _assert_close(got, ref, tol=1e-9):
    for f in _FIELDS:
        scale = max(float(np.abs(ref[f]).max()), 1e-300)
        err = float(np.abs(got[f] - ref[f]).max()) / scale
        assert err <= tol, f"{f}: {err:.3e}"


def test_set_field_round_trip_and_step_counter():
    """ssb_set_field is the inverse of ssb_get_field (particle-id order, whatever the storage order is)."""
    from spatialpy_b200.engine import Engine, EngineError
    fm = _model(False)
    rng = np.random.default_rng(1)
    with Engine(fm, device=0) as eng:
        eng.reset(11)
        eng.step(3)                                   # storage is cell-sorted by now
        for name in ("v", "F", "Fbp", "vt", "rho", "old_rho", "Frho", "bvf_phi", "nu", "C", "Q"):
            a = eng.get(name)
            b = a + rng.normal(size=a.shape)
            eng.set(name, b)
            np.testing.assert_array_equal(eng.get(name), b)
            eng.set(name, a)
        xx = eng.get("xx")
        eng.set("xx", xx[::-1].copy())
        np.testing.assert_array_equal(eng.get("xx"), xx[::-1])
        eng.set("xx", xx)
        step, epoch = eng.get_step()
        assert step == 3 and epoch > 0
        eng.set_step(7, epoch + 5)
        assert eng.get_step() == (7, epoch + 5)
        with pytest.raises(EngineError):               # derived fields cannot be set
            eng._check(eng.lib.ssb_set_field(eng._h, b"nbr_count", xx.ctypes.data, xx.nbytes))
        with pytest.raises(ValueError):
            eng.set("rho", np.zeros(3))


def test_trajectory_continues_in_a_fresh_handle():
    """The hand-over a re-partition performs, on one rank: pack the state after k steps, build a new model + engine from the
    rows, restore the fields and the step / epoch counters, continue — and land where the uninterrupted run lands (the
    neighbour lists are rebuilt at the hand-over, so only the summation order inside a sweep differs)."""
    from spatialpy_b200.engine import Engine
    from spatialpy_b200.slab import StateLayout, assemble_partition, pack_state, partition
    fm = _model(False)
    k, steps = 21, 26                                # hand over right after the Shepard-filter step 20
    with Engine(fm, device=0) as eng:
        eng.reset(11)
        eng.step(steps)
        ref = {f: eng.get(f) for f in _FIELDS + ("xx",)}
    part = partition(fm, 0, 1)
    lay = StateLayout.of(fm)
    with Engine(fm, device=0) as eng:
        eng.reset(11)
        eng.step(k)
        rows = pack_state(eng.get, part, lay)
        step, epoch = eng.get_step()
    new, fields = assemble_partition(part.local, lay, rows, part.edges, part.halo, 0, 1)
    np.testing.assert_array_equal(new.gids, np.arange(fm.num_particles))
    with Engine(new.local, device=0, owned=new.owned, rng_id=new.gids.astype(np.int32)) as eng:
        eng.reset(11)
        for name, val in fields.items():
            eng.set(name, val)
        eng.set_step(step, epoch)
        eng.step(steps - k)
        got = {f: eng.get(f) for f in _FIELDS + ("xx",)}
        assert eng.get_step()[0] == steps
    _assert_close(got, ref)
    assert int(got["xx"].sum()) == int(fm.u0.sum())


def _loopback_run(world, steps, every, transport="native", reactive=False):
    """`world` slab ranks as threads of this process on cuda:0 (LoopbackComm), re-partitioning every `every` steps.
    transport "native": halo messages written by the pack kernels into the neighbour's receive window, ordered by the engine
    streams (ssb_slab_*; same process => the windows are shared by pointer); "host": every exchange orchestrated from Python."""
    import threading
    import torch
    from spatialpy_b200.slab import LoopbackComm, LoopbackHub, SlabEngine, partition
    fm = _model(reactive)
    hub = LoopbackHub(world, timeout=300.0)
    outs, errs = {}, {}
    from spatialpy_b200 import codegen
    for r in range(world):                            # compile (or find) the model units before the threads race for them
        codegen.build_model_unit(partition(fm, r, world).local)

    def body(rank):
        try:
            torch.cuda.set_device(0)
            part = partition(fm, rank, world)
            se = SlabEngine(part, rank, world, device=0, comm=LoopbackComm(hub, rank, torch.device("cuda", 0)),
                            auto_repartition=True, repartition_every=every, transport=transport)
            se.reset(11)
            se.step(steps)
            out = {"gid": se.part.gids[se.part.owned == 1], "repartitions": se.repartitions, "counters": se.counters()}
            for f in _FIELDS + ("xx",):
                out[f] = se.owned_field(f)[1]
            outs[rank] = out
            se.close()
        except BaseException as err:  # noqa: BLE001 - reported by the main thread
            errs[rank] = err
            hub.barrier.abort()

    ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(600)
    if errs:          # the first real failure (the other ranks only see the barrier it broke)
        real = [e for e in errs.values() if not isinstance(e, threading.BrokenBarrierError)]
        raise (real or list(errs.values()))[0]
    return fm, outs


def _check_against_single(fm, outs, steps):
    from spatialpy_b200.engine import Engine
    with Engine(fm, device=0) as eng:
        eng.reset(11)
        eng.step(steps)
        ref = {f: eng.get(f) for f in _FIELDS}
    world = len(outs)
    gid = np.concatenate([outs[r]["gid"] for r in range(world)])
    assert sorted(gid.tolist()) == list(range(fm.num_particles))          # every particle owned exactly once
    got = {}
    for f in _FIELDS:
        got[f] = np.empty_like(ref[f])
        got[f][gid] = np.concatenate([outs[r][f] for r in range(world)])
    _assert_close(got, ref)
    assert sum(int(outs[r]["xx"].sum()) for r in range(world)) == int(fm.u0.sum())
    assert sum(outs[r]["counters"]["diffusions"] for r in range(world)) > 0


@pytest.mark.parametrize("transport", ["native", "host"])
@pytest.mark.parametrize("world,every", [(2, 0), (3, 0), (2, 5), (3, 4)])
def test_loopback_slabs_with_repartition_match_single_gpu(world, every, transport):
    """The whole slab code path on ONE GPU: `world` ranks as threads.  every = 0 is the fixed partition (same as the 2-GPU
    test); every > 0 hands the trajectory over to fresh handles several times."""
    steps = 22
    fm, outs = _loopback_run(world, steps, every, transport)
    _check_against_single(fm, outs, steps)
    want = 0 if every == 0 else (steps // every)
    assert all(outs[r]["repartitions"] == want for r in range(world))


@pytest.mark.parametrize("world", [2, 3])
def test_native_transport_equals_host_orchestrated_exchange_bit_for_bit(world):
    """Same kernels, same messages, same Philox numbering — only who moves the bytes differs (pack kernels writing into
    peer-mapped windows behind sequence flags vs. Python-orchestrated pack / copy / unpack): every field and every
    population of a REACTIVE run (births, deaths, jumps across the faces, the step-end overshoot events) must be identical."""
    steps = 22
    _, a = _loopback_run(world, steps, 0, "native", reactive=True)
    _, b = _loopback_run(world, steps, 0, "host", reactive=True)
    for r in range(world):
        np.testing.assert_array_equal(a[r]["gid"], b[r]["gid"])
        for f in _FIELDS + ("xx",):
            np.testing.assert_array_equal(a[r][f], b[r][f], err_msg=f"rank {r} {f}")
        assert a[r]["counters"]["reactions"] == b[r]["counters"]["reactions"]
        assert a[r]["counters"]["diffusions"] == b[r]["counters"]["diffusions"]
    assert sum(a[r]["counters"]["reactions"] for r in range(world)) > 0


def _worker_repart(rank, world, port, q, steps, every):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from spatialpy_b200.slab import SlabEngine, partition
    fm = _model(False)
    se = SlabEngine(partition(fm, rank, world), rank, world, device=rank, auto_repartition=True, repartition_every=every)
    se.reset(11)
    se.step(steps)
    out = {"gid": se.part.gids[se.part.owned == 1], "repartitions": se.repartitions, "counters": se.counters()}
    for f in _FIELDS + ("xx",):
        out[f] = se.owned_field(f)[1]
    q.put((rank, out))
    dist.barrier()
    se.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpu_slab_with_repartition_matches_single_gpu():
    import torch.multiprocessing as mp
    steps, every = 22, 5
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_repart, args=(r, 2, port, q, steps, every)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    _check_against_single(_model(False), outs, steps)
    assert all(outs[r]["repartitions"] == steps // every for r in range(2))


def test_solver_slab_decomposition_leaves_the_single_gpu_file_set():
    """Solver.run(decomposition="slab", devices=[0, 0]): two slabs (two engine handles on cuda:0) leave the same files as the
    single-handle run — same names, same file -> step map, every field within 1e-9, the molecule count conserved."""
    from spatialpy_b200 import Solver
    from spatialpy_b200.vtk import read_ssb, read_vtk
    fm = _model(False)
    fm.nt, fm.output_steps = 22, np.array([0, 11, 22], dtype=np.uint32)
    one = Solver(fm).run(seed=11, binary_store=True)
    two = Solver(fm).run(seed=11, binary_store=True, decomposition="slab", devices=[0, 0])
    assert sorted(os.listdir(one.result_dir)) == sorted(os.listdir(two.result_dir))
    nfiles = len([f for f in os.listdir(one.result_dir) if f.endswith(".ssb")])
    assert nfiles == 4                                   # steps 0, 1 (the reference's off-by-one), 11, 22
    for k in range(nfiles):
        pa, a = read_ssb(os.path.join(one.result_dir, f"output{k}.ssb"))
        pb, b = read_ssb(os.path.join(two.result_dir, f"output{k}.ssb"))
        assert list(a) == list(b) and a["__nfields_header__"] == b["__nfields_header__"]
        np.testing.assert_allclose(pb, pa, rtol=1e-6)
        for key in a:
            if key.startswith("D[") or key.startswith("__"):
                continue
            scale = max(float(np.abs(a[key]).max()), 1e-300)
            assert float(np.abs(np.asarray(b[key], dtype=float) - a[key]).max()) / scale <= 1e-9, (k, key)
        for sp in fm.species_names:
            assert int(b[f"D[{sp}]"].sum()) == int(a[f"D[{sp}]"].sum()) == int(fm.u0.sum())
        ta = open(os.path.join(one.result_dir, f"output{k}.vtk")).read().split("\n")
        tb = open(os.path.join(two.result_dir, f"output{k}.vtk")).read().split("\n")
        assert len(ta) == len(tb) and ta[:5] == tb[:5] and [l for l in ta if l[:1].isalpha()] == [l for l in tb if l[:1].isalpha()]
        _, va = read_vtk(os.path.join(two.result_dir, f"output{k}.vtk"))
        np.testing.assert_allclose(va["rho"], b["rho"], atol=5e-7)
    assert open(os.path.join(one.result_dir, "output0_boundingBox.vtk")).read() == \
        open(os.path.join(two.result_dir, "output0_boundingBox.vtk")).read()
