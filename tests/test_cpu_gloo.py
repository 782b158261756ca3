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


# ----------------------------------------------------------------------------------------------------------------------
# batched ensembles: many copies of a small static model in one engine handle (spatialpy_b200/ensemble.py)
# ----------------------------------------------------------------------------------------------------------------------
def _oracle_paths():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "oracle"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)


def test_replicated_model_is_block_diagonal_and_copies_are_independent_trajectories():
    """replicate_model: no neighbour list crosses copies, every copy has the original's lists and D_i_j; and run through the
    serial NSM restatement, the copies of ONE run are distributed like independent trajectories of the original model — their
    totals pass the KS test against the reference ensemble (tests/golden/birth_death.ens.npz) and are uncorrelated."""
    import numpy as np
    from scipy import stats
    _oracle_paths()
    import nsm_oracle
    import sdpd_oracle
    from util import load_ens, load_model
    from spatialpy_b200.ensemble import replicate_model
    fm = load_model("birth_death")
    N, copies = fm.num_particles, 50
    rep = replicate_model(fm, copies)
    assert rep.num_particles == copies * N and rep.static_domain and rep.h == fm.h
    np.testing.assert_array_equal(rep.u0.reshape(copies, N, -1)[7], fm.u0)
    o1 = sdpd_oracle.SdpdOracle(fm)
    n1 = o1.find_neighbors(o1.x, o1.x)
    o = sdpd_oracle.SdpdOracle(rep)
    nb = o.find_neighbors(o.x, o.x)
    own = np.repeat(np.arange(rep.num_particles), np.diff(nb["ptr"]))
    assert ((own // N) == (nb["j"] // N)).all()                                   # block diagonal
    for r in (0, copies // 2, copies - 1):
        lo, hi = nb["ptr"][r * N], nb["ptr"][(r + 1) * N]
        np.testing.assert_array_equal(nb["ptr"][r * N:(r + 1) * N + 1] - lo, n1["ptr"])
        np.testing.assert_array_equal(np.sort(nb["j"][lo:hi] - r * N), np.sort(n1["j"]))
        assert abs(nb["Dij"][lo:hi].sum() - n1["Dij"].sum()) <= 1e-9 * abs(n1["Dij"].sum())
    ens = load_ens("birth_death")
    t_end = int(ens["steps"][1]) * fm.dt
    lib = nsm_oracle.build(rep)
    totals = []
    for run in range(6):                                                          # 6 runs x 50 copies = 300 trajectories
        xx, _, _ = nsm_oracle.run(lib, rep, nb, 4000 + run, t_end)
        totals.append(xx.reshape(copies, N, -1).sum(axis=1)[:, 0])
    totals = np.array(totals).astype(np.int64)                                    # [runs, copies]
    p = stats.ks_2samp(totals.ravel(), ens["t1_totals"][:, 0]).pvalue
    assert p > 0.01, f"copies of a replicated run vs reference trajectories: KS p = {p:.4f}"
    # independence: neighbouring copies of the same run are no more alike than copies of different runs
    r_adj = np.corrcoef(totals[:, :-1].ravel(), totals[:, 1:].ravel())[0, 1]
    assert abs(r_adj) < 4.0 / np.sqrt(totals[:, 1:].size), r_adj
    import pytest
    with pytest.raises(ValueError):
        replicate_model(load_model("tank3d"), 2)                                  # moving domains are excluded


class _FakeBatchEngine:
    """Stands in for Engine in run_ensemble_batched: fields are known functions of (particle id, step, seed)."""
    created = []

    def __init__(self, fm, device=0, flags=0, rdme_epsilon=0.0):
        self.fm, self.t, self.seed = fm, 0, None
        _FakeBatchEngine.created.append(fm.num_particles)

    def reset(self, seed):
        self.t, self.seed = 0, seed

    def step(self, n=1):
        self.t += n

    def get(self, name):
        import numpy as np
        n = self.fm.num_particles
        pid = np.arange(n)
        if name == "v":
            return np.zeros((n, 3))
        if name in ("rho", "mass", "bvf_phi", "nu"):
            return {"rho": 1.0, "mass": 2.0, "bvf_phi": 0.0, "nu": 3.0}[name] + 1e-3 * pid + self.t
        if name == "C":
            return np.stack([pid + 0.5 * self.t + s for s in range(self.fm.num_chem_species)], axis=1).astype(float)
        if name == "xx":
            return np.stack([(pid + 3 * self.t + self.seed + s) % 19 for s in range(self.fm.num_stoch_species)], axis=1).astype(np.uint32)
        raise KeyError(name)

    def counters(self):
        return {"reactions": 2, "diffusions": 3, "seconds": 0.25, "windows": 10}

    def close(self):
        pass


def test_batched_ensemble_driver_splits_batches_seeds_and_files(tmp_path):
    """run_ensemble_batched with a fake engine: 10 trajectories in batches of 4 (4 + 4 + 2 copies, handles reused for equal
    sizes), batch b seeded seed + 4 b, every trajectory gets the reference's file set with ITS copy's slice of the state."""
    import numpy as np
    _oracle_paths()
    from util import load_model
    from spatialpy_b200.ensemble import default_batch, run_ensemble_batched
    from spatialpy_b200.slab import output_schedule
    from spatialpy_b200.vtk import read_ssb, read_vtk
    fm = load_model("birth_death")
    N = fm.num_particles
    dirs = [str(tmp_path / f"traj{k}") for k in range(10)]
    for d in dirs:
        os.makedirs(d)
    _FakeBatchEngine.created.clear()
    res = run_ensemble_batched(fm, 10, 100, out_dirs=dirs, batch=4, binary_store=True, engine_factory=_FakeBatchEngine)
    assert _FakeBatchEngine.created == [4 * N, 2 * N]
    assert res["counters"] == {"reactions": 6, "diffusions": 9, "seconds": 0.75, "windows": 30}
    sched = output_schedule(fm.nt, fm.output_steps)
    for k in (0, 3, 4, 9):
        b0, r = (k // 4) * 4, k % 4
        pid = np.arange(r * N, (r + 1) * N)
        names = sorted(os.listdir(dirs[k]))
        assert names == sorted(["output0_boundingBox.vtk"] + [f"output{f}.{e}" for f, _ in sched for e in ("vtk", "ssb")])
        for f, step in sched:
            pts, arr = read_ssb(os.path.join(dirs[k], f"output{f}.ssb"))
            np.testing.assert_array_equal(pts, fm.x.astype(np.float32))                   # the ORIGINAL coordinates, not the shifted copy
            np.testing.assert_array_equal(arr["D[" + fm.species_names[0] + "]"], (pid + 3 * step + 100 + b0) % 19)
            np.testing.assert_array_equal(arr["rho"], 1.0 + 1e-3 * pid + step)
            np.testing.assert_array_equal(arr["type"], fm.type)
            _, av = read_vtk(os.path.join(dirs[k], f"output{f}.vtk"))
            assert av["__nfields_header__"] == arr["__nfields_header__"] == 7 + fm.num_chem_species + (fm.num_stoch_species if step else 0)
        np.testing.assert_array_equal(res[k]["xx_final"][:, 0], (pid + 3 * fm.nt + 100 + b0) % 19)
    assert default_batch(121, 1024) == 1024 and default_batch(2500, 1024) == 104 and default_batch(10 ** 6, 8) == 1
    # the product path writes through the engine's C++ writers (ssb_write_snapshot): same bytes as the Python twins used above
    import filecmp
    from spatialpy_b200.vtk import write_snapshot
    dirs2 = [str(tmp_path / f"cxx{k}") for k in range(10)]
    for d in dirs2:
        os.makedirs(d)
    run_ensemble_batched(fm, 10, 100, out_dirs=dirs2, batch=4, binary_store=True, engine_factory=_FakeBatchEngine, writer=write_snapshot)
    for a, b in zip(dirs, dirs2):
        cmp = filecmp.dircmp(a, b)
        assert not cmp.left_only and not cmp.right_only
        match, mismatch, errors = filecmp.cmpfiles(a, b, cmp.common_files, shallow=False)
        assert not mismatch and not errors and len(match) == len(cmp.common_files)


def test_solver_batch_keyword_plumbing(monkeypatch):
    """Solver.run(batch=...): one batched-ensemble call for the whole ensemble, output flags turned into the driver's arguments,
    success flags per trajectory; a moving domain is refused the way engine failures are reported (solver.py:595-597).  A
    replicated model needs no new model unit (units depend on the reactions, not on the particles)."""
    import pytest
    _oracle_paths()
    from util import load_model
    import spatialpy_b200.ensemble as ensemble
    from spatialpy_b200 import SimulationError, Solver, codegen
    from spatialpy_b200.ensemble import replicate_model
    fm = load_model("birth_death")
    assert codegen.build_model_unit(replicate_model(fm, 3)) == codegen.build_model_unit(fm)
    calls = []

    def stub(flat, ntraj, seed, device=0, out_dirs=None, batch=None, flags=0, rdme_epsilon=0.0, vtk=True, binary_store=False, **kw):
        calls.append((ntraj, seed, device, len(out_dirs), batch, vtk, binary_store))
        if not flat.static_domain:
            raise ValueError("replicate_model is for static domains")
        return {**{k: {} for k in range(ntraj)}, "counters": {}}

    monkeypatch.setattr(ensemble, "run_ensemble_batched", stub)
    res = Solver(fm).run(number_of_trajectories=5, seed=9, batch=True, devices=[2], binary_store=True)
    assert calls == [(5, 9, 2, 5, None, True, True)] and len(res) == 5 and all(r.success for r in res)
    Solver(fm).run(number_of_trajectories=5, seed=9, batch=2)
    assert calls[-1][4] == 2
    # several devices: contiguous shards, each seeded with seed + its first trajectory index
    calls.clear()
    res = Solver(fm).run(number_of_trajectories=7, seed=100, batch=True, devices=[0, 1, 3])
    assert sorted(calls) == [(1, 106, 3, 1, None, True, False), (3, 100, 0, 3, None, True, False), (3, 103, 1, 3, None, True, False)]
    assert len(res) == 7 and all(r.success for r in res)
    with pytest.raises(SimulationError, match="return code = 4"):
        Solver(load_model("tank3d")).run(seed=1, batch=True)


def test_ensemble_lane_that_fails_to_start_does_not_hang_the_others():
    """A lane whose engine cannot be created (out of memory on one of 24 handles, a bad device ordinal) never arrives at the
    lanes' meeting points; the healthy lanes must not wait for it for ever: run_ensemble raises the lane's error promptly."""
    import threading
    import time
    from spatialpy_b200 import FlatModel
    from spatialpy_b200.ensemble import run_ensemble
    from conftest import GOLDEN
    fm = FlatModel.load(os.path.join(GOLDEN, "birth_death.model.npz"))
    made = []

    class Stub:
        def __init__(self, fm, device=0, **kw):
            made.append(device)
            if len(made) == 2:
                raise MemoryError("lane 2 could not allocate")
        def reset(self, seed): pass
        def step(self, n): pass
        def run_no_files(self, seed, n, first_traj=0): time.sleep(0.01)
        def counters(self): return {"reactions": 0, "diffusions": 0, "seconds": 0.0, "windows": 0}
        def close(self): pass

    out = {}
    def go():
        try:
            run_ensemble(fm, 8, 1, devices=[0], lanes=4, engine_factory=Stub)
        except BaseException as err:  # noqa: BLE001
            out["err"] = err
    t = threading.Thread(target=go)
    t.start()
    t.join(20)
    assert not t.is_alive(), "run_ensemble hangs when one lane fails to start"
    assert isinstance(out.get("err"), MemoryError)
