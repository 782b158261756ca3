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
                 unit_path=None, on_engine=None, rank=0, world_size=1):
    """Run trajectories k = rank, rank+world_size, ... (seed + k each) on `devices`, `lanes` engine handles per device, each
    driven by its own host thread (ctypes releases the GIL during ssb_run).  Returns {k: counters} for the local trajectories;
    raises the first EngineError.  `out_dirs[k]` (optional) receives trajectory k's VTK files."""
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
                    eng = Engine(fm, device=dev, flags=flags, rdme_epsilon=rdme_epsilon, unit_path=unit_path)
                    if on_engine:
                        on_engine(eng)
                try:
                    created_barrier.wait()
                except threading.BrokenBarrierError:
                    pass
                if eng is not None:
                    eng.reset(seed)          # one engine step sizes every lazily allocated buffer
                    eng.step(1)
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
        except EngineError as err:
            with lock:
                errors.append(err)
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
