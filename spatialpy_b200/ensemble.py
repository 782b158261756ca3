"""Ensemble sharding over GPUs: trajectories are independent units (the reference runs them as separate processes with
seed+k, spatialpy/solvers/solver.py:547-605), so the multi-GPU path has no data-path collective — only the
bookkeeping below and a max/sum over ranks for timing and counters."""


def shard_trajectories(number_of_trajectories, rank, world_size):
    """Trajectory indices owned by `rank`: k -> rank k mod world_size.  Trajectory k always uses seed + k, whatever the
    world size, so an ensemble is reproducible across GPU counts."""
    return list(range(rank, number_of_trajectories, world_size))


def reduce_scalar(value, op="max", device=None):
    """max / sum of a python float over the ranks of the default torch.distributed group (no-op when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def default_lanes(num_particles):
    """Concurrent engine handles per GPU: small models cannot fill a B200 with one trajectory (a 144-voxel model occupies two
    CTAs), so several trajectories run side by side on separate streams; large models get the GPU to themselves."""
    if num_particles >= 200_000:
        return 1
    if num_particles >= 20_000:
        return 4
    return 24


def run_ensemble(fm, number_of_trajectories, seed, devices=(0,), lanes=None, out_dirs=None, flags=None, rdme_epsilon=0.0,
                 unit_path=None, on_engine=None, rank=0, world_size=1, engine_factory=None):
    """Run trajectories k = rank, rank+world_size, ... (seed + k each) on `devices`, `lanes` engine handles per device, each
    driven by its own host thread (ctypes releases the GIL during ssb_run).  Returns {k: counters} for the local trajectories;
    raises the first failure of any lane.  `out_dirs[k]` (optional) receives trajectory k's VTK files."""
    import threading
    from .engine import Engine, EngineError, FLAG_NO_VTK, FLAG_SKIP_STATIC_FORCES
    lanes = lanes or default_lanes(fm.num_particles)
    if flags is None:
        flags = FLAG_SKIP_STATIC_FORCES | (0 if out_dirs is not None else FLAG_NO_VTK)
    mine = shard_trajectories(number_of_trajectories, rank, world_size)
    workers = [(d, l) for l in range(lanes) for d in devices][:max(1, len(mine))]
    results, errors, lock = {}, [], threading.Lock()
    # engines are torn down only after EVERY lane has finished: cudaFree synchronises the whole device, so a lane that closed
    # early would stall ~100 times behind the other lanes' running sSSA kernels (measured: 11 s instead of 1.5 s)
    done_barrier = threading.Barrier(len(workers))
    # ... and they are all created (and have done their one-off allocations: neighbour lists, cached coefficients) BEFORE any
    # lane starts its trajectories: cudaMalloc / cudaMallocHost also synchronise with running kernels, so a lane that is still
    # allocating would wait ~100 times for the other lanes' sSSA kernels (measured: 4 s of start-up for 16 lanes)
    ready_barrier = threading.Barrier(len(workers))
    created_barrier = threading.Barrier(len(workers))

    def work(w):
        dev, _ = workers[w]
        ks = mine[w::len(workers)]
        eng = None
        try:
            try:
                if ks:
                    eng = (engine_factory or Engine)(fm, device=dev, flags=flags, rdme_epsilon=rdme_epsilon, unit_path=unit_path)
                    if on_engine:
                        on_engine(eng)
                try:
                    created_barrier.wait()
                except threading.BrokenBarrierError:
                    pass
                if eng is not None:
                    eng.reset(seed)          # one engine step sizes every lazily allocated buffer
                    eng.step(1)
            except BaseException:
                # this lane will never arrive at the meeting points: break them BEFORE waiting at one, or the healthy lanes wait
                # for ever (threading.Barrier needs all parties) and Solver.run hangs in join()
                for b in (created_barrier, ready_barrier, done_barrier):
                    b.abort()
                raise
            finally:
                try:
                    ready_barrier.wait()
                except threading.BrokenBarrierError:
                    pass
            if ks:
                for k in ks:
                    if errors:
                        break
                    if out_dirs is not None:
                        eng.run(seed, [out_dirs[k]], first_traj=k)
                    else:
                        eng.run_no_files(seed, 1, first_traj=k)
                    c = eng.counters()
                    with lock:
                        results[k] = c
        except BaseException as err:      # noqa: BLE001 - any failure of a lane is the ensemble's failure (re-raised by the caller)
            with lock:
                errors.append(err)
            # a lane that failed will never arrive at the meeting points: break them, or the healthy lanes wait for ever
            # (threading.Barrier needs all parties) and Solver.run hangs in join()
            for b in (created_barrier, ready_barrier, done_barrier):
                b.abort()
        finally:
            try:
                done_barrier.wait()
            except threading.BrokenBarrierError:
                pass
            if eng is not None:
                eng.close()

    threads = [threading.Thread(target=work, args=(w,)) for w in range(len(workers))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


# ----------------------------------------------------------------------------------------------------------------------
# Batched ensembles (opt-in): many trajectories of a small STATIC model in ONE engine handle
# ----------------------------------------------------------------------------------------------------------------------
def replicate_model(fm, copies, gap=None):
    """`copies` disjoint copies of a static-domain model laid out on a grid of tiles as ONE FlatModel (pitch = extent + gap,
    gap >= 2 h so no neighbour list crosses copies).  Copy r owns particle ids r*N .. (r+1)*N-1.

    Why: a 121-voxel model occupies one CTA; even with 24 engine handles side by side (run_ensemble) a B200 runs ~50 of its
    ~600 resident CTAs.  The copies share nothing — the reference runs trajectories as separate processes
    (spatialpy/solvers/solver.py:547-605) — and nothing in a static-domain step couples them: events are per voxel, jumps go
    to neighbours within h (E/src/simulate_rdme.cpp:359-366), the sSSA window length comes from max Ddiag, which every copy
    shares, and the Philox counters are keyed by particle id, which differs between copies.  So every copy is a statistically
    independent trajectory of the original model (moving domains are excluded: their step-end overshoot event is one per
    SYSTEM, simulate_rdme.cpp:233-238)."""
    import numpy as np
    from .flatmodel import FlatModel
    fm = fm.finalize()
    if not fm.static_domain:
        raise ValueError("replicate_model is for static domains (a moving domain's step-end event couples the copies)")
    copies = int(copies)
    if copies < 1:
        raise ValueError("copies must be >= 1")
    # reference boundary conditions are coordinate predicates (`me->x[0] >= xmin && ...`, spatialpy/core/boundarycondition.py:121-146):
    # in a translated copy they would select the wrong region, or nothing at all — silently wrong C, v, nu in every copy but the first
    if (fm.bc_source or "").strip() and "me->x" in fm.bc_source:
        raise ValueError("batched ensembles translate the copies of the model, but its boundary conditions test particle "
                         "coordinates (me->x[...]); run this model with lanes (batch=None)")
    gap = 2.0 * fm.h if gap is None else float(gap)
    if gap < 1.5 * fm.h:
        raise ValueError("gap must be at least 1.5 h so that no neighbour list crosses copies")
    # copies sit on a near-square grid of tiles in x (and y for 2-D / 3-D models): the engine's cell list has at most 4096 cells
    # per axis (ssb_core.cu setup_grid), so a single row of a thousand copies would coarsen the cells
    nx = copies if fm.dimension < 2 else int(np.ceil(np.sqrt(copies)))
    ny = -(-copies // nx)
    pitch = [float(fm.x[:, d].max() - fm.x[:, d].min()) + gap for d in range(2)]
    x = np.tile(fm.x, (copies, 1))
    r = np.repeat(np.arange(copies), fm.num_particles)
    x[:, 0] += (r % nx) * pitch[0]
    x[:, 1] += (r // nx) * pitch[1]
    t1 = lambda a: np.tile(a, copies)                       # noqa: E731  per-particle vectors
    return FlatModel(
        name=f"{fm.name}_x{copies}", x=x, type=t1(fm.type), nu=t1(fm.nu), mass=t1(fm.mass), c=t1(fm.c), rho=t1(fm.rho),
        solid=t1(fm.solid), species_names=list(fm.species_names), reactions=list(fm.reactions), parameters=dict(fm.parameters),
        type_constants=dict(fm.type_constants), u0=np.tile(fm.u0, (copies, 1)), N_dense=fm.N_dense, irN=fm.irN, jcN=fm.jcN,
        prN=fm.prN, irG=fm.irG, jcG=fm.jcG, diffusion_matrix=fm.diffusion_matrix, data_fn=np.tile(fm.data_fn, (1, copies)),
        bc_source=fm.bc_source, enable_pde=fm.enable_pde, enable_rdme=fm.enable_rdme, static_domain=True, dt=fm.dt, nt=fm.nt,
        output_steps=fm.output_steps, h=fm.h, rho0=fm.rho0, c0=fm.c0, P0=fm.P0,
        xlim=(fm.xlim[0], fm.xlim[1] + (nx - 1) * pitch[0]), ylim=(fm.ylim[0], fm.ylim[1] + (ny - 1) * pitch[1]), zlim=fm.zlim,
        dimension=fm.dimension,
        gravity=fm.gravity).finalize()


def default_batch(num_particles, number_of_trajectories):
    """Copies per engine handle: enough voxels (~2^18) to give every SM several chunks of the sSSA window kernel."""
    return max(1, min(int(number_of_trajectories), 262144 // max(int(num_particles), 1)))


def run_ensemble_batched(fm, number_of_trajectories, seed, device=0, out_dirs=None, batch=None, flags=None, rdme_epsilon=0.0,
                         vtk=True, binary_store=False, engine_factory=None, on_engine=None, writer=None):
    """The ensemble as batches of `batch` trajectories per engine handle (`replicate_model`).  Trajectory k is copy k mod batch
    of batch k // batch; batch b is seeded with seed + b * batch, so a run is reproducible for a given (seed, batch) — the
    reference's "trajectory k uses seed + k" mapping (solver.py:558-559) holds for the ENSEMBLE LAW, not trajectory by trajectory.
    With `out_dirs`, every trajectory gets the reference's own file set (output%u.vtk / .ssb, output0_boundingBox.vtk, same
    file -> step map), written by the engine's own writers (`ssb_write_snapshot`) from the copies' slices of the state.
    Returns {k: {"xx_final": [N, S_d] populations}} plus the summed counters under key "counters"."""
    import numpy as np
    from .engine import Engine, FLAG_CORRECTED_OUTPUT_STEPS, FLAG_NO_VTK, FLAG_SKIP_STATIC_FORCES
    from .slab import output_schedule
    from .vtk import write_snapshot, write_snapshot_py
    if writer is None:      # the engine's C++ writers; the Python twins when a fake engine stands in (CPU tier)
        writer = write_snapshot if engine_factory is None else write_snapshot_py
    fm = fm.finalize()
    ntraj = int(number_of_trajectories)
    B = int(batch) if batch else default_batch(fm.num_particles, ntraj)
    N, Sc, Sd = fm.num_particles, fm.num_chem_species, fm.num_stoch_species
    flags = (FLAG_SKIP_STATIC_FORCES if flags is None else flags) | FLAG_NO_VTK
    factory = engine_factory or Engine
    schedule = output_schedule(fm.nt, fm.output_steps, corrected=bool(flags & FLAG_CORRECTED_OUTPUT_STEPS))
    results, totals = {}, {"reactions": 0, "diffusions": 0, "seconds": 0.0, "windows": 0}
    eng, eng_copies = None, 0
    try:
        for b0 in range(0, ntraj, B):
            nb = min(B, ntraj - b0)
            if eng is None or eng_copies != nb:           # the last batch may be smaller
                if eng is not None:
                    eng.close()
                eng = factory(replicate_model(fm, nb), device=device, flags=flags, rdme_epsilon=rdme_epsilon)
                eng_copies = nb
                if on_engine:
                    on_engine(eng)
            eng.reset(seed + b0)
            done = 0
            for file_index, step in schedule:
                if step > done:
                    eng.step(step - done)
                    done = step
                if out_dirs is None:
                    continue
                v = eng.get("v").reshape(nb, N, 3)
                scal = np.stack([eng.get(name).reshape(nb, N) for name in ("rho", "mass", "bvf_phi", "nu")], axis=1)    # [nb, 4, N]
                C = eng.get("C").reshape(nb, N, Sc) if Sc else None
                D = eng.get("xx").reshape(nb, N, Sd) if Sd else None
                init = 1 if (Sd > 0 and step > 0) else 0          # output.cpp:151-154: output0 undercounts FIELD
                for r in range(nb):
                    writer(out_dirs[b0 + r], file_index, fm.x, v[r], scal[r], C[r].T if Sc else None, fm.type, D[r].T if Sd else None,
                           fm.species_names, (fm.xlim, fm.ylim, fm.zlim), step=step, rdme_initialized=init, vtk=vtk, binary=binary_store)
            xx = eng.get("xx").reshape(nb, N, Sd) if Sd else np.zeros((nb, N, 0), np.uint32)
            for r in range(nb):
                results[b0 + r] = {"xx_final": xx[r].copy()}
            c = eng.counters()
            for key in ("reactions", "diffusions", "windows"):
                totals[key] += c[key]
            totals["seconds"] += c["seconds"]
    finally:
        if eng is not None:
            eng.close()
    results["counters"] = totals
    return results
