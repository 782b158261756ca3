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
    # the variable-length exchange and the scalar reductions the re-partition uses
    from spatialpy_b200.slab import DistComm
    comm = DistComm(rank, world)
    rows = np.arange((3 + 2 * rank) * 4, dtype=np.float64).reshape(-1, 4) + 100 * rank
    got = comm.exchange_rows({1 - rank: rows}, 4)
    other = np.arange((3 + 2 * (1 - rank)) * 4, dtype=np.float64).reshape(-1, 4) + 100 * (1 - rank)
    ok = ok and np.array_equal(got[1 - rank], other)
    ok = ok and comm.allreduce(float(rank), "max") == 1.0 and comm.allreduce(float(rank) + 2.0, "min") == 2.0
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


def _moved_state(fm, seed=9):
    """A global mid-trajectory state defined per global id: positions displaced by up to 0.3 h (some particles cross a
    slab face), every other field an arbitrary function of the id."""
    rng = np.random.default_rng(seed)
    n = fm.num_particles
    st = {"x": fm.x + rng.uniform(-0.3, 0.3, size=(n, 3)) * fm.h, "v": rng.normal(size=(n, 3)), "F": rng.normal(size=(n, 3)),
          "Fbp": rng.normal(size=(n, 3)), "rho": 1.0 + 0.01 * rng.random(n), "Frho": rng.normal(size=n), "nu": fm.nu * 1.5,
          "mass": fm.mass, "bvf_phi": rng.random(n), "type": fm.type,
          "C": rng.random((n, fm.num_chem_species)), "Q": rng.normal(size=(n, fm.num_chem_species)),
          "xx": rng.integers(0, 50, size=(n, fm.num_stoch_species)).astype(np.uint32)}
    return st


def test_repartition_equals_a_fresh_partition_of_the_moved_domain():
    """pack_state -> rows_for_neighbour -> assemble_partition on every rank == partition() of the moved global model:
    same owned / ghost sets in the same order, mirror-image exchange lists, and every field handed over intact."""
    import copy
    from spatialpy_b200.slab import StateLayout, assemble_partition, pack_state, partition, rows_for_neighbour, slab_bounds
    fm = _model()
    world = 3
    edges = slab_bounds(fm.x[:, 0], world)
    parts = [partition(fm, r, world, edges=edges) for r in range(world)]
    st = _moved_state(fm)
    lay = StateLayout.of(fm)
    assert lay.width == 3 + 12 + 6 + 2 * fm.num_chem_species + fm.num_stoch_species + fm.num_data_fn
    rows = [pack_state(lambda name, p=p: st[name][p.gids], p, lay) for p in parts]
    assert sum(len(r) for r in rows) == fm.num_particles
    moved = copy.copy(fm)
    moved.x, moved.rho, moved.nu, moved.u0 = st["x"], st["rho"], st["nu"], st["xx"]
    crossed = 0
    for r in range(world):
        nbs = [nb for nb in (r - 1, r + 1) if 0 <= nb < world]
        got = [rows_for_neighbour(rows[nb], lay, edges, parts[r].halo, r) for nb in nbs]     # what the neighbours send to r
        new, fields = assemble_partition(parts[r].local, lay, np.concatenate([rows[r]] + got), edges, parts[r].halo, r, world)
        want = partition(moved, r, world, halo=parts[r].halo, edges=edges)
        np.testing.assert_array_equal(new.gids, want.gids)
        np.testing.assert_array_equal(new.owned, want.owned)
        assert set(new.send_ids) == set(want.send_ids) and set(new.recv_ids) == set(want.recv_ids)
        for nb in new.send_ids:
            np.testing.assert_array_equal(new.send_ids[nb], want.send_ids[nb])
            np.testing.assert_array_equal(new.recv_ids[nb], want.recv_ids[nb])
        for name in ("x", "type", "nu", "mass", "c", "rho", "solid", "u0", "data_fn"):
            np.testing.assert_array_equal(getattr(new.local, name), getattr(want.local, name), err_msg=name)
        for name, val in fields.items():
            np.testing.assert_array_equal(val, st[name][new.gids], err_msg=name)
        assert new.local.reactions == fm.reactions and new.local.h == fm.h and new.edges is not None
        crossed += int((~np.isin(new.gids[new.owned == 1], parts[r].gids[parts[r].owned == 1])).sum())
    assert crossed > 0                                   # the test did move particles across slab faces


def test_assemble_rejects_duplicate_ownership():
    import pytest
    from spatialpy_b200.slab import StateLayout, assemble_partition, pack_state, partition, slab_bounds
    fm = _model()
    edges = slab_bounds(fm.x[:, 0], 2)
    p = partition(fm, 0, 2, edges=edges)
    st = _moved_state(fm)
    lay = StateLayout.of(fm)
    rows = pack_state(lambda name: st[name][p.gids], p, lay)
    with pytest.raises(RuntimeError, match="twice"):
        assemble_partition(p.local, lay, np.concatenate([rows, rows[:3]]), edges, p.halo, 0, 2)


def test_loopback_comm_matches_the_collective_semantics():
    """LoopbackComm (ranks = threads of one process; the one-GPU test double of NCCL): neighbour exchange, scalar reductions
    and the variable-length row exchange."""
    import threading
    from spatialpy_b200.slab import LoopbackComm, LoopbackHub
    world = 3
    hub = LoopbackHub(world, timeout=60.0)
    out = {}

    def body(rank):
        comm = LoopbackComm(hub, rank)
        nbs = [nb for nb in (rank - 1, rank + 1) if 0 <= nb < world]
        res = []
        for rep in range(3):                              # repeated rounds must not see stale posts
            send = {nb: torch.full((4,), 10.0 * rank + nb + 100 * rep, dtype=torch.float64) for nb in nbs}
            recv = {nb: torch.zeros(4, dtype=torch.float64) for nb in nbs}
            comm.exchange(send, recv)
            res.append(all(float(recv[nb][0]) == 10.0 * nb + rank + 100 * rep for nb in nbs))
            res.append(comm.allreduce(rank + rep, "max") == world - 1 + rep and comm.allreduce(rank + rep, "min") == rep)
            rows = comm.exchange_rows({nb: np.full((rank + 1 + rep, 2), float(rank)) for nb in nbs}, 2)
            res.append(all(rows[nb].shape == (nb + 1 + rep, 2) and (rows[nb] == nb).all() for nb in nbs))
        out[rank] = all(res)

    ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(120)
    assert out == {0: True, 1: True, 2: True}


class _FakeRank:
    """Stands in for SlabEngine in run_slab_trajectory: every field is a known function of (global id, engine step)."""

    def __init__(self, part, rank, world, device, hub, flags, rdme_epsilon, fail_at=None):
        self.part, self.rank, self.t, self.fail_at = part, rank, 0, fail_at
        self.m = part.owned.astype(bool)

    def reset(self, seed):
        self.t = 0

    def step(self, n=1):
        self.t += n
        if self.fail_at is not None and self.rank == 1 and self.t == self.fail_at:
            raise RuntimeError("boom")

    @staticmethod
    def field(fm, name, gid, t):
        g = gid.astype(np.float64)
        if name == "x":
            return fm.x[gid] + 1e-3 * t
        if name == "v":
            return np.stack([g, -g, g * 0 + t], axis=1) * 1e-2
        if name in ("rho", "mass", "bvf_phi", "nu"):
            return {"rho": 1.0, "mass": 2.0, "bvf_phi": 0.0, "nu": 3.0}[name] + 1e-4 * g + t
        if name == "type":
            return fm.type[gid]
        if name == "C":
            return np.stack([g + 0.5 * t + s for s in range(fm.num_chem_species)], axis=1)
        if name == "xx":
            return np.stack([(gid + 3 * t + s) % 17 for s in range(fm.num_stoch_species)], axis=1).astype(np.uint32)
        raise KeyError(name)

    def owned_field(self, name):
        gid = self.part.gids[self.m]
        return gid, self.field(self._fm, name, gid, self.t)

    def counters(self):
        return {"reactions": 1, "diffusions": 2 + self.rank, "seconds": 0.5 * self.rank, "windows": 7}

    def close(self):
        pass


def test_slab_trajectory_driver_writes_the_reference_file_set(tmp_path):
    """run_slab_trajectory with fake rank engines (3 ranks as threads): the file -> step map of the reference's output gate,
    host-assembled snapshots in global-id order, both writers, FIELD undercount in output0, counters summed."""
    from spatialpy_b200.slab import output_schedule, run_slab_trajectory
    from spatialpy_b200.vtk import read_ssb, read_vtk
    fm = _model()
    fm.nt, fm.output_steps = 12, np.array([0, 5, 10], dtype=np.uint32)
    _FakeRank._fm = fm
    assert output_schedule(12, [0, 5, 10]) == [(0, 0), (1, 1), (2, 5), (3, 10), (4, 12)]     # simulate_threads.cpp:231-247,283-288
    # SSB_FLAG_CORRECTED_OUTPUT_STEPS: file k <-> step output_steps[k], no extra files (mirrors the branch in ssb_run)
    assert output_schedule(12, [0, 5, 10], corrected=True) == [(0, 0), (1, 5), (2, 10)]
    assert output_schedule(10, [0, 5, 10], corrected=True) == [(0, 0), (1, 5), (2, 10)]
    assert output_schedule(10, [0, 5, 10, 20], corrected=True) == [(0, 0), (1, 5), (2, 10)]
    total = run_slab_trajectory(fm, [0, 0, 0], 3, str(tmp_path), vtk=True, binary_store=True, rank_engine=_FakeRank)
    assert total == {"reactions": 3, "diffusions": 9, "seconds": 1.0, "windows": 7}
    names = sorted(os.listdir(tmp_path))
    assert names == sorted(["output0_boundingBox.vtk"] + [f"output{k}.{e}" for k in range(5) for e in ("vtk", "ssb")])
    gid = np.arange(fm.num_particles)
    for k, step in output_schedule(12, [0, 5, 10]):
        pts, arr = read_ssb(str(tmp_path / f"output{k}.ssb"))
        np.testing.assert_array_equal(pts, _FakeRank.field(fm, "x", gid, step).astype(np.float32))
        np.testing.assert_array_equal(arr["v"], _FakeRank.field(fm, "v", gid, step))
        for name in ("rho", "mass", "bvf_phi", "nu"):
            np.testing.assert_array_equal(arr[name], _FakeRank.field(fm, name, gid, step))
        np.testing.assert_array_equal(arr["type"], fm.type)
        for s, sp in enumerate(fm.species_names):
            np.testing.assert_array_equal(arr[f"C[{sp}]"], _FakeRank.field(fm, "C", gid, step)[:, s])
            np.testing.assert_array_equal(arr[f"D[{sp}]"], _FakeRank.field(fm, "xx", gid, step)[:, s])
        pv, av = read_vtk(str(tmp_path / f"output{k}.vtk"))
        assert av["__nfields_header__"] == arr["__nfields_header__"] == 7 + fm.num_chem_species + (fm.num_stoch_species if step else 0)
        np.testing.assert_allclose(av["rho"], arr["rho"], atol=5e-7)
        np.testing.assert_array_equal(av[f"D[{fm.species_names[0]}]"], arr[f"D[{fm.species_names[0]}]"])
        np.testing.assert_allclose(pv, pts, rtol=1e-6)


def test_slab_trajectory_driver_propagates_a_rank_failure(tmp_path):
    import functools
    import pytest
    from spatialpy_b200.slab import run_slab_trajectory
    fm = _model()
    fm.nt, fm.output_steps = 6, np.array([0, 3, 6], dtype=np.uint32)
    _FakeRank._fm = fm
    with pytest.raises(RuntimeError, match="boom"):
        run_slab_trajectory(fm, [0, 0], 3, str(tmp_path), rank_engine=functools.partial(_FakeRank, fail_at=2))
    polled = []
    with pytest.raises(InterruptedError):
        run_slab_trajectory(fm, [0, 0], 3, str(tmp_path), rank_engine=_FakeRank, cancelled=lambda: polled.append(1) or len(polled) > 3)


def test_box_slab_builder_follows_the_partition_rule_and_survives_a_repartition():
    """configs.box_slab (rank-local builder of the 64 M-particle workload) = the owner/ghost rule of partition() for its slab
    faces, and a re-partition of the unmoved domain gives the same sets and exchange lists back."""
    from spatialpy_b200 import configs
    from spatialpy_b200.slab import StateLayout, assemble_partition, pack_state, rows_for_neighbour
    world = 3
    parts = [configs.box_slab(r, world, nx_per_rank=12, ny=6, nz=8) for r in range(world)]
    lay = StateLayout.of(parts[0].local)

    def getter(p):
        fm = p.local
        n = fm.num_particles
        zero3, zero = np.zeros((n, 3)), np.zeros(n)
        fields = {"x": fm.x, "v": zero3, "F": zero3, "Fbp": zero3, "rho": fm.rho, "Frho": zero, "nu": fm.nu, "mass": fm.mass,
                  "bvf_phi": zero, "type": fm.type, "C": fm.u0.astype(float), "Q": np.zeros((n, lay.Sc)), "xx": fm.u0}
        return lambda name: fields[name]

    rows = [pack_state(getter(p), p, lay) for p in parts]
    for r, p in enumerate(parts):
        x = p.local.x[:, 0]
        owner = np.clip(np.searchsorted(p.edges, x, side="right") - 1, 0, world - 1)
        assert ((owner == r) == (p.owned == 1)).all()
        nbs = [nb for nb in (r - 1, r + 1) if 0 <= nb < world]
        got = [rows_for_neighbour(rows[nb], lay, p.edges, p.halo, r) for nb in nbs]
        new, _ = assemble_partition(p.local, lay, np.concatenate([rows[r]] + got), p.edges, p.halo, r, world)
        np.testing.assert_array_equal(new.gids, p.gids)
        np.testing.assert_array_equal(new.owned, p.owned)
        for nb in nbs:
            np.testing.assert_array_equal(new.send_ids[nb], p.send_ids[nb])
            np.testing.assert_array_equal(new.recv_ids[nb], p.recv_ids[nb])
        np.testing.assert_array_equal(new.local.x, p.local.x)
        np.testing.assert_array_equal(new.local.u0, p.local.u0)


# ----------------------------------------------------------------------------------------------------------------------
# the stepping protocol itself over gloo: SlabEngine drives a numpy stand-in for the engine handle (tests/fake_engine.py)
# ----------------------------------------------------------------------------------------------------------------------
_PROTO_FIELDS = ("x", "v", "F", "rho", "bvf_phi", "C", "xx")


def _proto_model():
    from spatialpy_b200 import configs
    return configs.tank_sdpd(n=12, nt=10, output_every=10, dt=2e-5)


def _proto_engine_factory(fm, **kw):
    """FakeEngine whose particles start with a per-id velocity of 0.05 h per step along +-x, so that they cross slab faces."""
    from fake_engine import FakeEngine

    class Drifting(FakeEngine):
        def reset(self, seed):
            super().reset(seed)
            sign = np.where(self.gid % 2 == 0, 1.0, -1.0)
            self.v[:, 0] = sign * 0.05 * self.fm.h / self.fm.dt * (self.fm.solid == 0)
    return Drifting(fm, **kw)


def _proto_run(rank, world, steps, every):
    from spatialpy_b200.slab import SlabEngine, partition
    fm = _proto_model()
    se = SlabEngine(partition(fm, rank, world), rank, world, flags=128 | 8, engine_factory=_proto_engine_factory,
                    torch_device=torch.device("cpu"), auto_repartition=True, repartition_every=every)
    se.reset(1)
    se.step(steps)
    out = {"gid": se.part.gids[se.part.owned == 1], "repartitions": se.repartitions, "jumps": se.counters()["diffusions"]}
    for f in _PROTO_FIELDS:
        out[f] = se.owned_field(f)[1]
    se.close()
    return out


def _proto_worker(rank, world, port, q, steps, every):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, _proto_run(rank, world, steps, every)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_slab_stepping_protocol_over_gloo_reproduces_the_single_rank_run():
    """SlabEngine.step + repartition over gloo (2 ranks) with the numpy stand-in engine: ghost synchronisation after every
    sweep, inbox traffic across the face, and the re-partition hand-over must give the single-rank result BIT FOR BIT (the
    stand-in sums neighbours in global-id order), with particles migrating between the slabs on the way."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    steps, every = 8, 2
    ref = _proto_run(0, 1, steps, 0)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_proto_worker, args=(r, 2, port, q, steps, every)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    fm = _proto_model()
    gid = np.concatenate([outs[r]["gid"] for r in range(2)])
    assert sorted(gid.tolist()) == list(range(fm.num_particles))
    order = np.argsort(ref["gid"])
    for f in _PROTO_FIELDS:
        got = np.empty_like(ref[f])
        got[gid] = np.concatenate([outs[r][f] for r in range(2)])
        np.testing.assert_array_equal(got, ref[f][order], err_msg=f)
    assert all(outs[r]["repartitions"] == steps // every for r in range(2))
    assert int(sum(outs[r]["xx"].sum() for r in range(2))) == int(fm.u0.sum())
    assert sum(outs[r]["jumps"] for r in range(2)) == ref["jumps"] > 0
    # particles did change owner on the way
    from spatialpy_b200.slab import partition
    first = partition(fm, 0, 2)
    assert set(outs[0]["gid"].tolist()) != set(first.gids[first.owned == 1].tolist())


def test_solver_slab_keyword_plumbing(monkeypatch):
    """Solver.run(decomposition="slab"): one driver call per trajectory with seed+k (solver.py:558-559), the output flags turned
    into the driver's vtk / binary_store arguments, success flags set, a model the decomposition cannot run reported as the
    reference reports engine failures (SimulationError with a return code, solver.py:595-597)."""
    import pytest
    import spatialpy_b200.slab as slab
    from spatialpy_b200 import SimulationError, Solver, configs
    from spatialpy_b200.engine import FLAG_BINARY_STORE, FLAG_NO_VTK
    calls = []

    def stub(fm, devices, seed, out_dir, flags=0, rdme_epsilon=0.0, vtk=True, binary_store=False, cancelled=None, **kw):
        assert os.path.isdir(out_dir) and not (flags & (FLAG_BINARY_STORE | FLAG_NO_VTK)) and not cancelled()
        calls.append((list(devices), seed, vtk, binary_store, rdme_epsilon))
        if fm.static_domain:
            raise ValueError("slab decomposition is implemented for moving domains")
        return {}

    monkeypatch.setattr(slab, "run_slab_trajectory", stub)
    sol = Solver(_model())
    res = sol.run(number_of_trajectories=3, seed=40, decomposition="slab", devices=[0, 1], binary_store=True, vtk=False,
                  rdme_epsilon=0.1)
    assert calls == [([0, 1], 40 + k, False, True, 0.1) for k in range(3)]
    assert len(res) == 3 and all(r.success and not r.timeout for r in res)
    assert len({r.result_dir for r in res}) == 3
    with pytest.raises(SimulationError, match="unknown decomposition"):
        sol.run(decomposition="pencil")
    static = Solver(configs.cylinder_rdme(delta=0.25, nt=10, output_every=10))
    with pytest.raises(SimulationError, match="return code = 4"):
        static.run(seed=1, decomposition="slab", devices=[0, 1])
