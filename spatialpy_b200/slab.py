"""Spatial slab decomposition of one ssa_sdpd domain over several GPUs (one process per GPU, NCCL send/recv over NVLink).

The reference has no domain decomposition (one shared-memory process); interactions are short range (radius h,
and sSSA jumps go to neighbours within h, E/src/simulate_rdme.cpp:359-366), so the domain splits into 1-D slabs along x:

* every rank's model = the particles it OWNS (x inside its slab) + GHOST copies of the neighbours' particles within
  `halo` of the slab faces (halo >= h*(1+skin) + the distance particles may travel before a re-partition);
* ghosts run the same per-particle kernels (predictor / corrector are deterministic functions of synchronised inputs, so
  they reproduce the owner's values bit for bit) but skip the neighbour sweeps; after each sweep the owner's results are
  sent to the ghost copies — three small messages per step to each slab neighbour:
      after the force sweep   F[3] Fbp[3] Frho Q[S_c]      (group 0)
      after the corrector     rho_new                      (group 1; needed by the serial-order rule of the BVF sweep)
      after the BVF sweep     v[3] bvf_phi                 (group 2)
* sSSA: a ghost voxel is simulated only by its owner; molecules that jump INTO a ghost voxel are read-and-cleared from
  the local inbox after every window and added to the owner's inbox, which sees them as ordinary mail at its next window;
  the window count per step comes from the GLOBAL maximum jump rate (one scalar all-reduce), so every rank uses the same
  windows.  Philox counters and the serial particle order use GLOBAL particle ids, so results do not depend on the
  partition (up to the summation order inside a neighbour sweep).

Re-partitioning (migrating particles): ownership and ghost sets are valid while nobody has travelled further than the halo
allows (a pair within h*(1+skin) must have both members present).  `SlabEngine.step` tracks the global maximum displacement;
when the bound is used up it calls `repartition()`: every rank packs the full inter-step state of the particles it owns
(`StateLayout`), sends each slab neighbour the rows that now lie inside that neighbour's slab + halo, re-derives owned / ghost
sets and the exchange lists from the received rows with the same rule as `partition()` (slab faces stay where they are), and
continues the trajectory in a fresh engine handle (`ssb_set_field` / `ssb_set_step`, include/ssb.h) at the same step and
Philox epoch.  Host-orchestrated and rare (hundreds of steps apart at SDPD time steps); the set logic is pure numpy and tested on
the CPU (tests/test_cpu_slab.py), the hand-over on the GPU (tests/test_gpu_slab.py, loopback ranks on one GPU or NCCL on two).
"""
import os
import threading

import numpy as np

from .engine import (Engine, FLAG_CORRECTED_OUTPUT_STEPS, FLAG_SKIP_STATIC_FORCES, PH_CORRECTOR, PH_END, PH_FINISH, PH_PRE, PH_RDME_CLOSE, PH_RDME_INIT,
                     PH_RDME_PREP, PH_RDME_WINDOW, PH_RDME_MIN, PH_RDME_EXTRA)
from .flatmodel import FlatModel


class SlabPartition:
    """What one rank needs: its local model (owned first, then ghosts, each sorted by global id) and the exchange lists."""

    def __init__(self, local, gids, owned, send_ids, recv_ids, bounds, halo, edges=None):
        self.edges = edges            # all slab faces along x ([world+1], outer faces infinite); kept across re-partitions
        self.local = local            # FlatModel of owned + ghost particles
        self.gids = gids              # [n_local] global particle id of every local particle
        self.owned = owned            # [n_local] int32 1/0
        self.send_ids = send_ids      # {neighbour rank: local ids of MY owned particles that are ghosts there (by global id)}
        self.recv_ids = recv_ids      # {neighbour rank: local ids of MY ghosts owned by that rank (by global id)}
        self.bounds = bounds          # (lo, hi) of the slab along x
        self.halo = halo

    @property
    def n_owned(self):
        return int(self.owned.sum())


def slab_bounds(x, world):
    """Slab faces along x with (nearly) equal particle counts per slab; the outer faces are infinite."""
    q = np.quantile(x, np.linspace(0.0, 1.0, world + 1))
    edges = [-np.inf] + [0.5 * (np.max(x[x <= q[r]]) + np.min(x[x > q[r]])) if (x > q[r]).any() else q[r] for r in range(1, world)] + [np.inf]
    return np.array(edges)


def subset_model(fm, idx, name):
    """FlatModel restricted to the global particle indices `idx` (order preserved)."""
    return FlatModel(
        name=name, x=fm.x[idx], type=fm.type[idx], nu=fm.nu[idx], mass=fm.mass[idx], c=fm.c[idx], rho=fm.rho[idx],
        solid=fm.solid[idx], species_names=list(fm.species_names), reactions=list(fm.reactions), parameters=dict(fm.parameters),
        type_constants=dict(fm.type_constants), u0=fm.u0[idx], N_dense=fm.N_dense, irN=fm.irN, jcN=fm.jcN, prN=fm.prN,
        irG=fm.irG, jcG=fm.jcG, diffusion_matrix=fm.diffusion_matrix, data_fn=fm.data_fn[:, idx], bc_source=fm.bc_source,
        enable_pde=fm.enable_pde, enable_rdme=fm.enable_rdme, static_domain=fm.static_domain, dt=fm.dt, nt=fm.nt,
        output_steps=fm.output_steps, h=fm.h, rho0=fm.rho0, c0=fm.c0, P0=fm.P0, xlim=fm.xlim, ylim=fm.ylim, zlim=fm.zlim,
        dimension=fm.dimension, gravity=fm.gravity).finalize()


def partition(fm, rank, world, halo=None, edges=None, gids=None):
    """Partition a (global or pre-cut) model for `rank`.  `halo` defaults to 1.6 h (h * 1.1 skin + 0.5 h of travel)."""
    x = fm.x[:, 0]
    halo = float(halo) if halo is not None else 1.6 * fm.h
    edges = slab_bounds(x, world) if edges is None else np.asarray(edges)
    gids = np.arange(fm.num_particles, dtype=np.int64) if gids is None else np.asarray(gids, dtype=np.int64)
    owner = np.clip(np.searchsorted(edges, x, side="right") - 1, 0, world - 1)
    lo, hi = edges[rank], edges[rank + 1]
    mine = owner == rank
    ghost = (~mine) & (x >= lo - halo) & (x < hi + halo)
    if world > 1 and np.isfinite(lo) and np.isfinite(hi) and (hi - lo) < halo:
        raise ValueError(f"slab {rank} is thinner ({hi - lo:g}) than the halo ({halo:g}); use fewer ranks")
    own_idx = np.nonzero(mine)[0]
    gh_idx = np.nonzero(ghost)[0]
    idx = np.concatenate([own_idx, gh_idx])
    local = subset_model(fm, idx, f"{fm.name}_slab{rank}of{world}")
    owned = np.concatenate([np.ones(len(own_idx), np.int32), np.zeros(len(gh_idx), np.int32)])
    lg = gids[idx]
    send_ids, recv_ids = {}, {}
    for nb in (rank - 1, rank + 1):
        if nb < 0 or nb >= world:
            continue
        nlo, nhi = edges[nb], edges[nb + 1]
        # my owned particles inside the neighbour's halo region  <->  the neighbour's ghosts owned by me
        s_loc = np.nonzero((x[own_idx] >= nlo - halo) & (x[own_idx] < nhi + halo))[0]
        send_ids[nb] = s_loc[np.argsort(gids[own_idx][s_loc], kind="stable")].astype(np.int32)
        r_loc = np.nonzero(owner[gh_idx] == nb)[0]
        recv_ids[nb] = (len(own_idx) + r_loc[np.argsort(gids[gh_idx][r_loc], kind="stable")]).astype(np.int32)
    return SlabPartition(local, lg, owned, send_ids, recv_ids, (lo, hi), halo, edges=edges)


# ----------------------------------------------------------------------------------------------------------------------
# re-partition: the state of a particle between two engine steps, as one float64 row (ints are exact below 2^53)
# ----------------------------------------------------------------------------------------------------------------------
class StateLayout:
    """Column map of a state row: what a particle carries when its owner or its ghost copies change.

    Between two engine steps a particle is described by its identity (global id, type, solidTag, mass, c, data_fn), the
    integrator state the next predictor reads (x, v, rho, F, Fbp, Frho, nu — take_step1, E/src/simulate.cpp:56-109; C and
    the flux Q of the last sweep, simulate.cpp:81-85), bvf_phi (output only) and the populations xx.  On a moving domain the
    NSM is rebuilt every step (E/src/simulate_rdme.cpp:54-65), so propensities and event clocks are not state."""

    def __init__(self, Sc, Sd, ndf):
        self.Sc, self.Sd, self.ndf = int(Sc), int(Sd), int(ndf)
        self.cols, k = {}, 0
        for name, w in (("gid", 1), ("type", 1), ("solid", 1), ("x", 3), ("v", 3), ("F", 3), ("Fbp", 3), ("rho", 1), ("Frho", 1),
                        ("nu", 1), ("mass", 1), ("c", 1), ("bvf_phi", 1), ("C", self.Sc), ("Q", self.Sc), ("xx", self.Sd),
                        ("data_fn", self.ndf)):
            self.cols[name] = slice(k, k + w)
            k += w
        self.width = k

    @classmethod
    def of(cls, fm):
        return cls(fm.num_chem_species, fm.num_stoch_species, fm.num_data_fn)

    def col(self, rows, name):
        a = rows[:, self.cols[name]]
        return a[:, 0] if name in ("gid", "type", "solid", "rho", "Frho", "nu", "mass", "c", "bvf_phi") else a


def pack_state(get, part, lay):
    """[n_owned, width] state rows of the particles this rank owns.  `get(name)` returns a field in local particle order
    (Engine.get)."""
    fm = part.local
    n = fm.num_particles
    rows = np.empty((n, lay.width), dtype=np.float64)
    rows[:, lay.cols["gid"]] = part.gids.reshape(n, 1)
    rows[:, lay.cols["type"]] = np.asarray(get("type")).reshape(n, 1)
    rows[:, lay.cols["solid"]] = np.asarray(fm.solid).reshape(n, 1)
    for name in ("x", "v", "F", "Fbp"):
        rows[:, lay.cols[name]] = np.asarray(get(name)).reshape(n, 3)
    for name in ("rho", "Frho", "nu", "mass", "bvf_phi"):
        rows[:, lay.cols[name]] = np.asarray(get(name)).reshape(n, 1)
    rows[:, lay.cols["c"]] = np.asarray(fm.c).reshape(n, 1)
    if lay.Sc:
        rows[:, lay.cols["C"]] = np.asarray(get("C")).reshape(n, lay.Sc)
        rows[:, lay.cols["Q"]] = np.asarray(get("Q")).reshape(n, lay.Sc)
    if lay.Sd:
        rows[:, lay.cols["xx"]] = np.asarray(get("xx")).reshape(n, lay.Sd)
    if lay.ndf:
        rows[:, lay.cols["data_fn"]] = np.asarray(fm.data_fn).T
    return rows[part.owned.astype(bool)]


def rows_for_neighbour(rows, lay, edges, halo, nb):
    """The rows whose position lies inside slab `nb` extended by the halo: nb's future owned particles and ghosts."""
    x = rows[:, lay.cols["x"]][:, 0]
    return rows[(x >= edges[nb] - halo) & (x < edges[nb + 1] + halo)]


def model_from_rows(template, lay, rows, name):
    """FlatModel of the particles in `rows` (current state as the initial condition); species, reactions, parameters, the
    stoichiometry and every other per-model table are the template's."""
    fm = template
    n = rows.shape[0]
    S = fm.num_species
    u0 = np.zeros((n, S), np.uint32)
    if lay.Sd:
        u0[:, :lay.Sd] = rows[:, lay.cols["xx"]].astype(np.uint32)
    return FlatModel(
        name=name, x=rows[:, lay.cols["x"]].copy(), type=lay.col(rows, "type").astype(np.int32), nu=lay.col(rows, "nu").copy(),
        mass=lay.col(rows, "mass").copy(), c=lay.col(rows, "c").copy(), rho=lay.col(rows, "rho").copy(),
        solid=lay.col(rows, "solid").astype(np.int32), species_names=list(fm.species_names), reactions=list(fm.reactions),
        parameters=dict(fm.parameters), type_constants=dict(fm.type_constants), u0=u0, N_dense=fm.N_dense, irN=fm.irN, jcN=fm.jcN,
        prN=fm.prN, irG=fm.irG, jcG=fm.jcG, diffusion_matrix=fm.diffusion_matrix,
        data_fn=np.ascontiguousarray(rows[:, lay.cols["data_fn"]].T), bc_source=fm.bc_source, enable_pde=fm.enable_pde,
        enable_rdme=fm.enable_rdme, static_domain=fm.static_domain, dt=fm.dt, nt=fm.nt, output_steps=fm.output_steps, h=fm.h,
        rho0=fm.rho0, c0=fm.c0, P0=fm.P0, xlim=fm.xlim, ylim=fm.ylim, zlim=fm.zlim, dimension=fm.dimension,
        gravity=fm.gravity).finalize()


def assemble_partition(template, lay, rows, edges, halo, rank, world):
    """New SlabPartition of `rank` from the state rows it holds after the exchange (its own owned particles + what the slab
    neighbours sent), by the same rule as `partition()`: owner = the slab containing x, ghosts = non-owned rows within the halo,
    owned first then ghosts, each sorted by global id.  Returns (partition, fields) with `fields` = the state the fresh engine
    needs on top of the model's initial condition, in the new local order."""
    edges = np.asarray(edges)
    gid = lay.col(rows, "gid").astype(np.int64)
    if np.unique(gid).size != gid.size:
        raise RuntimeError("re-partition received a particle twice (ownership was not unique)")
    x = rows[:, lay.cols["x"]][:, 0]
    owner = np.clip(np.searchsorted(edges, x, side="right") - 1, 0, world - 1)
    lo, hi = edges[rank], edges[rank + 1]
    mine = owner == rank
    ghost = (~mine) & (x >= lo - halo) & (x < hi + halo)
    own_idx = np.nonzero(mine)[0]
    own_idx = own_idx[np.argsort(gid[own_idx], kind="stable")]
    gh_idx = np.nonzero(ghost)[0]
    gh_idx = gh_idx[np.argsort(gid[gh_idx], kind="stable")]
    idx = np.concatenate([own_idx, gh_idx])
    sel = rows[idx]
    local = model_from_rows(template, lay, sel, template.name)
    owned = np.concatenate([np.ones(len(own_idx), np.int32), np.zeros(len(gh_idx), np.int32)])
    send_ids, recv_ids = {}, {}
    for nb in (rank - 1, rank + 1):
        if nb < 0 or nb >= world:
            continue
        nlo, nhi = edges[nb], edges[nb + 1]
        xo = x[own_idx]
        send_ids[nb] = np.nonzero((xo >= nlo - halo) & (xo < nhi + halo))[0].astype(np.int32)     # already in global-id order
        recv_ids[nb] = (len(own_idx) + np.nonzero(owner[gh_idx] == nb)[0]).astype(np.int32)
    part = SlabPartition(local, gid[idx], owned, send_ids, recv_ids, (lo, hi), halo, edges=edges)
    fields = {name: np.ascontiguousarray(sel[:, lay.cols[name]]) for name in ("v", "F", "Fbp")}
    fields.update({name: np.ascontiguousarray(lay.col(sel, name)) for name in ("Frho", "bvf_phi")})
    if lay.Sc:
        fields["C"] = np.ascontiguousarray(sel[:, lay.cols["C"]])
        fields["Q"] = np.ascontiguousarray(sel[:, lay.cols["Q"]])
    return part, fields


def exchange(send, recv, rank, tag_base=0):
    """Point-to-point halo exchange with the slab neighbours: `send`/`recv` map neighbour rank -> contiguous torch tensor.
    Works on any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    ops = []
    for nb in sorted(set(send) | set(recv)):
        if nb in send and send[nb].numel():
            ops.append(dist.P2POp(dist.isend, send[nb], nb))
        if nb in recv and recv[nb].numel():
            ops.append(dist.P2POp(dist.irecv, recv[nb], nb))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class DistComm:
    """The communication a slab rank needs, over torch.distributed (NCCL between GPUs; gloo in the CPU tests)."""

    def __init__(self, rank, world, device=None):
        import torch
        self.torch, self.rank, self.world = torch, rank, world
        self.device = device if device is not None else torch.device("cpu")

    def _sync(self):
        if self.device.type == "cuda":
            self.torch.cuda.current_stream(self.device).synchronize()

    def exchange(self, send, recv):
        exchange(send, recv, self.rank)
        self._sync()

    def allreduce(self, value, op):
        """Scalar max / min over all ranks."""
        if self.world == 1:
            return float(value)
        import torch.distributed as dist
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.MIN)
        return float(t.item())

    def allgather_bytes(self, blob):
        """Every rank's `blob` (bytes), in rank order — control plane of the native transport (IPC handles travel once per partition)."""
        if self.world == 1:
            return [blob]
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, blob)
        return out

    def exchange_rows(self, send, width):
        """Variable-length exchange with the slab neighbours: {nb: float64 [n_nb, width]} -> {nb: float64 [m_nb, width]}
        (row counts first, then the rows)."""
        torch = self.torch
        nbs = sorted(send)
        cs = {nb: torch.tensor([float(send[nb].shape[0])], dtype=torch.float64, device=self.device) for nb in nbs}
        cr = {nb: torch.zeros(1, dtype=torch.float64, device=self.device) for nb in nbs}
        self.exchange(cs, cr)
        ps = {nb: torch.as_tensor(np.ascontiguousarray(send[nb], dtype=np.float64)).to(self.device) for nb in nbs}
        pr = {nb: torch.empty((int(cr[nb].item()), width), dtype=torch.float64, device=self.device) for nb in nbs}
        self.exchange(ps, pr)
        return {nb: pr[nb].cpu().numpy() for nb in nbs}


class LoopbackHub:
    """Shared mailbox of `world` slab ranks that run as threads of ONE process (several engine handles on one GPU): the
    test double of the NCCL transport — same SlabEngine code path, exchanges become device-to-device copies."""

    def __init__(self, world, timeout=300.0):
        self.world = world
        self.barrier = threading.Barrier(world, timeout=timeout)
        self.box = {}


class LoopbackComm:
    def __init__(self, hub, rank, device=None):
        import torch
        self.torch, self.hub, self.rank, self.world = torch, hub, rank, hub.world
        self.device = device if device is not None else torch.device("cpu")

    def _round(self, key, send, take):
        """Post `send` {nb: obj}, wait for everybody, collect take(obj) from the neighbours' posts, wait again (nobody may
        overwrite a post before it has been read)."""
        box = self.hub.box
        for nb, obj in send.items():
            box[(key, self.rank, nb)] = obj
        self.hub.barrier.wait()
        out = take(lambda nb: box[(key, nb, self.rank)])
        if self.device.type == "cuda":
            self.torch.cuda.current_stream(self.device).synchronize()
        self.hub.barrier.wait()
        return out

    def exchange(self, send, recv):
        def take(post):
            for nb, t in recv.items():
                if t.numel():
                    t.copy_(post(nb))
        self._round("x", send, take)

    def allreduce(self, value, op):
        vals = self._round("r", {nb: float(value) for nb in range(self.world)},
                           lambda post: [post(nb) for nb in range(self.world)])
        return max(vals) if op == "max" else min(vals)

    def exchange_rows(self, send, width):
        return self._round("s", {nb: np.array(a, dtype=np.float64, copy=True) for nb, a in send.items()},
                           lambda post: {nb: post(nb).reshape(-1, width) for nb in send})

    def allgather_bytes(self, blob):
        return self._round("g", {nb: blob for nb in range(self.world)}, lambda post: [post(nb) for nb in range(self.world)])


class SlabEngine:
    """One rank of a slab-decomposed run: an Engine on the local model + the halo exchanges between the step phases."""

    def __init__(self, part, rank, world, device=0, flags=FLAG_SKIP_STATIC_FORCES, rdme_epsilon=0.0, comm=None,
                 auto_repartition=True, repartition_every=0, engine_factory=None, torch_device=None, transport=None):
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        # transport "native" (default with the CUDA engine): halo messages are written by the engine's pack kernels straight into
        # the neighbour's receive window over NVLink and ordered by the engine stream (ssb_slab_*, include/ssb.h); `comm` is only the
        # control plane (IPC handles once per partition, re-partition rows).  transport "host": every exchange is orchestrated
        # from here through `comm` (phase API) — the protocol the CPU tier pins with a stand-in engine, and the cross-check of
        # the native path on the GPU.
        self.transport = transport or ("host" if engine_factory is not None else "native")
        if part.local.static_domain:
            raise ValueError("slab decomposition is implemented for moving domains (static ensembles shard by trajectory)")
        self.device_index = device
        # engine_factory / torch_device exist for the CPU tier only (tests/test_cpu_slab.py drives this class over gloo with a
        # numpy stand-in for the engine handle to exercise the exchange protocol); the product path is Engine on a CUDA device
        self.engine_factory = engine_factory or Engine
        self.dev = torch_device if torch_device is not None else torch.device("cuda", device)
        self.comm = comm if comm is not None else DistComm(rank, world, self.dev)
        self.flags, self.rdme_epsilon = flags, rdme_epsilon
        self.overshoot = not (flags & 129)     # FLAG_CORRECTED_NSM_SELECT | FLAG_NO_STEP_OVERSHOOT switch the extra event off
        self.auto_repartition = auto_repartition
        self.repartition_every = int(repartition_every)      # > 0: also re-partition every that many steps (tests)
        self.repartitions = 0
        self.steps_since_partition = 0
        self.seed = 0
        self.halo_bytes_per_step = 0
        self._carry = {"reactions": 0, "diffusions": 0, "seconds": 0.0, "windows": 0}
        self.eng = None
        self.part0 = part             # the partition of the initial condition (reset() returns to it)
        self._build(part)

    def _build(self, part):
        """Engine handle + exchange buffers for a partition (the first one, and every re-partition)."""
        torch = self.torch
        self.part = part
        self.eng = self.engine_factory(part.local, device=self.device_index, flags=self.flags, rdme_epsilon=self.rdme_epsilon,
                                       owned=part.owned, rng_id=part.gids.astype(np.int32))
        self.Sd = part.local.num_stoch_species
        self.travel_bound = 0.0       # upper bound on how far any particle has moved since the partition was made
        self.steps_since_partition = 0
        if self.transport == "native":
            self.eng.slab_setup(self.rank, self.world, part.send_ids, part.recv_ids)
            self.eng.slab_connect(self.comm.allgather_bytes(self.eng.slab_export()))
            return
        self.send_ids = {nb: torch.as_tensor(v, device=self.dev) for nb, v in part.send_ids.items()}
        self.recv_ids = {nb: torch.as_tensor(v, device=self.dev) for nb, v in part.recv_ids.items()}
        self.buf = {}
        for g in range(4):
            w = self.eng.halo_width(g)
            self.buf[g] = ({nb: torch.empty((len(v), w), dtype=torch.float64, device=self.dev) for nb, v in self.send_ids.items()},
                           {nb: torch.empty((len(v), w), dtype=torch.float64, device=self.dev) for nb, v in self.recv_ids.items()})
        # inbox traffic flows the other way: from my ghosts (recv_ids) to their owners (the neighbour's send_ids)
        self.ibuf = ({nb: torch.empty((len(v), max(self.Sd, 1)), dtype=torch.int32, device=self.dev) for nb, v in self.recv_ids.items()},
                     {nb: torch.empty((len(v), max(self.Sd, 1)), dtype=torch.int32, device=self.dev) for nb, v in self.send_ids.items()})

    def _retire_engine(self):
        """Collective: unmap the neighbours' windows, meet, then free (nobody may still be writing into a window that goes away)."""
        if self.eng is None:
            return
        if self.transport == "native":
            self.eng.slab_disconnect()
            if self.world > 1:
                try:
                    self.comm.allreduce(0.0, "max")
                except Exception:       # noqa: BLE001 - a rank that already failed broke the meeting point; freeing is all that is left
                    pass
        self.eng.close()
        self.eng = None

    def close(self):
        self._retire_engine()

    def reset(self, seed):
        self.seed = seed
        self._carry = {"reactions": 0, "diffusions": 0, "seconds": 0.0, "windows": 0}
        if self.part is not self.part0:        # a re-partition replaced the model by a mid-trajectory state
            self._retire_engine()
            self._build(self.part0)
        self.travel_bound = 0.0
        self.steps_since_partition = 0
        self.eng.reset(seed)

    def counters(self):
        """Engine counters of the whole trajectory (summed over the handles a re-partition retired)."""
        c = self.eng.counters()
        return {k: c[k] + self._carry[k] for k in self._carry}

    def _sync_group(self, g):
        send, recv = self.buf[g]
        for nb, ids in self.send_ids.items():
            self.eng.halo_pack(g, ids.data_ptr(), ids.numel(), send[nb].data_ptr())
        self.comm.exchange(send, recv)
        for nb, ids in self.recv_ids.items():
            self.eng.halo_unpack(g, ids.data_ptr(), ids.numel(), recv[nb].data_ptr())

    def _sync_inbox(self):
        send, recv = self.ibuf
        for nb, ids in self.recv_ids.items():          # my ghosts' mail -> owner
            self.eng.inbox_pack(ids.data_ptr(), ids.numel(), send[nb].data_ptr())
        self.comm.exchange(send, recv)
        for nb, ids in self.send_ids.items():          # mail for my owned particles that are ghosts over there
            self.eng.inbox_add(ids.data_ptr(), ids.numel(), recv[nb].data_ptr())

    def _step_native(self, n):
        """n engine steps through ssb_slab_step; the engine stops early (on every rank after the same step) when the travel bound
        is used up, and the re-partition runs here."""
        limit = 0.5 * (self.part.halo - 1.1 * self.part.local.h)
        stale_is_error = (not self.auto_repartition and self.repartition_every <= 0) or self.part.edges is None
        left = int(n)
        while left > 0:
            chunk = left
            if self.repartition_every > 0:
                chunk = min(chunk, max(1, self.repartition_every - self.steps_since_partition))
            done, self.travel_bound = self.eng.slab_step(chunk, limit if self.world > 1 else 0.0)
            left -= done
            self.steps_since_partition += done
            if done < chunk:
                if stale_is_error:
                    raise RuntimeError(f"slab partition is stale: particles may have travelled {self.travel_bound:g} since the "
                                       f"partition (halo {self.part.halo:g}, h {self.part.local.h:g}); re-partition needed "
                                       "(SlabEngine(auto_repartition=True))")
                self.repartition()
            elif self.world > 1 and self.repartition_every > 0 and self.steps_since_partition >= self.repartition_every:
                self.repartition()

    def step(self, n=1):
        if self.transport == "native":
            return self._step_native(n)
        for _ in range(n):
            e = self.eng
            e.phase(PH_PRE)
            self._sync_group(0)
            e.phase(PH_CORRECTOR)
            self._sync_group(1)
            e.phase(PH_FINISH)
            self._sync_group(2)
            if self.Sd > 0:
                mx = self.comm.allreduce(e.phase(PH_RDME_PREP), "max")
                nwin = int(e.phase(PH_RDME_INIT, mx))
                for w in range(nwin):
                    e.phase(PH_RDME_WINDOW, w)
                    self._sync_inbox()
                e.phase(PH_RDME_CLOSE)
                # the reference's one event past the end of every step (simulate_rdme.cpp:233-238): globally earliest pending event
                if self.overshoot:
                    tmin = self.comm.allreduce(e.phase(PH_RDME_MIN), "min")
                    e.phase(PH_RDME_EXTRA, tmin)
                    self._sync_inbox()
                    e.phase(PH_RDME_CLOSE, -1.0)       # delivers under the epoch the overshoot reserved (same numbering as one GPU)
            e.phase(PH_END)
            self.steps_since_partition += 1
            # a pair within h*(1+skin) must have both members present on the owner's rank, so nobody may travel further than
            # half of what the halo leaves beyond the candidate radius
            disp = e.skin_stats()["step_disp_max"]
            limit = 0.5 * (self.part.halo - 1.1 * self.part.local.h)
            if (not self.auto_repartition and self.repartition_every <= 0) or self.part.edges is None:
                self.travel_bound += disp
                if self.travel_bound > limit:
                    raise RuntimeError(f"slab partition is stale: particles may have travelled {self.travel_bound:g} since the "
                                       f"partition (halo {self.part.halo:g}, h {self.part.local.h:g}); re-partition needed "
                                       "(SlabEngine(auto_repartition=True))")
                continue
            if self.world == 1:
                continue              # no ghosts, nothing can go stale
            # re-partitioning is collective, so the bound is the global maximum and every rank acts after the same step
            self.travel_bound += self.comm.allreduce(disp, "max")
            if self.travel_bound > limit or (self.repartition_every > 0 and self.steps_since_partition >= self.repartition_every):
                self.repartition()

    def repartition(self):
        """Move particles to the slab they are in now and rebuild the ghost sets (collective: every rank calls it after the
        same step).  The trajectory continues in a fresh engine handle at the same step and Philox epoch."""
        e, old = self.eng, self.part
        lay = StateLayout.of(old.local)
        rows = pack_state(e.get, old, lay)
        nbs = [nb for nb in (self.rank - 1, self.rank + 1) if 0 <= nb < self.world]
        got = self.comm.exchange_rows({nb: rows_for_neighbour(rows, lay, old.edges, old.halo, nb) for nb in nbs}, lay.width)
        rows = np.concatenate([rows] + [got[nb] for nb in nbs], axis=0)
        part, fields = assemble_partition(old.local, lay, rows, old.edges, old.halo, self.rank, self.world)
        # both sides of every face must have derived the same exchange lists (they would not if a particle crossed a whole
        # slab between two re-partitions): compare the global ids, fail loudly
        echo = self.comm.exchange_rows({nb: part.gids[part.send_ids[nb]].astype(np.float64).reshape(-1, 1) for nb in nbs}, 1)
        for nb in nbs:
            if not np.array_equal(echo[nb][:, 0].astype(np.int64), part.gids[part.recv_ids[nb]]):
                raise RuntimeError(f"re-partition: rank {self.rank} and rank {nb} disagree about the ghosts on their face "
                                   "(a particle crossed more than one slab since the last re-partition)")
        step, epoch = e.get_step()
        c = e.counters()
        for k in self._carry:
            self._carry[k] += c[k]
        self._retire_engine()
        self._build(part)
        self.eng.reset(self.seed)
        for name, val in fields.items():
            self.eng.set(name, val)
        self.eng.set_step(step, epoch)
        self.repartitions += 1

    # -- gather helpers for tests -------------------------------------------------------------------------------
    def owned_field(self, name):
        """(global ids, values) of the owned particles for a tap field."""
        a = self.eng.get(name)
        m = self.part.owned.astype(bool)
        return self.part.gids[m], a[m]


# ----------------------------------------------------------------------------------------------------------------------
# a whole trajectory with output files, slab-decomposed over the GPUs of one process (what Solver.run(decomposition="slab") calls)
# ----------------------------------------------------------------------------------------------------------------------
def output_schedule(nt, output_steps, corrected=False):
    """[(file index, engine step)] of one trajectory: the output gate of run_simulation (E/src/simulate_threads.cpp:231-247,
    283-288) exactly as ssb_run walks it — `next_output_step` starts at 0 and get_next_output() hands out the table from its
    first entry (so step 0 is written twice when the table starts with 0: the reference's file->step off-by-one), and the
    final state is written once more after the last step.  `corrected` (SSB_FLAG_CORRECTED_OUTPUT_STEPS): file k holds step
    output_steps[k] and nothing else is written."""
    steps = [int(v) for v in output_steps]
    if corrected:
        out, k = [], 0
        for step in range(int(nt)):
            while k < len(steps) and steps[k] <= step:
                if steps[k] == step:
                    out.append((len(out), step))
                k += 1
        if any(v == int(nt) for v in steps[k:]):
            out.append((len(out), int(nt)))
        return out
    out, nxt, k, f = [], 0, 0, 0
    for step in range(int(nt)):
        if step >= nxt:
            out.append((f, step))
            f += 1
            nxt = steps[k] if k < len(steps) else 0xFFFFFFFF
            k += 1
    out.append((f, int(nt)))
    return out


def _default_rank_engine(part, rank, world, device, hub, flags, rdme_epsilon):
    import torch
    torch.cuda.set_device(device)             # thread-local: this rank's copies and synchronisations go to its own GPU
    return SlabEngine(part, rank, world, device=device, flags=flags, rdme_epsilon=rdme_epsilon,
                      comm=LoopbackComm(hub, rank, torch.device("cuda", device)))


def run_slab_trajectory(fm, devices, seed, out_dir, flags=FLAG_SKIP_STATIC_FORCES, rdme_epsilon=0.0, vtk=True, binary_store=False,
                        halo=None, cancelled=None, rank_engine=_default_rank_engine, writer=None):
    """One trajectory of a moving-domain model split into len(devices) slabs, driven from ONE process: rank r is a thread
    with its own engine handle on GPU devices[r]; the halo exchanges are device-to-device copies (`LoopbackComm`; peer copies
    over NVLink between distinct GPUs, plain copies when a device is listed twice).  At the reference's output steps
    (`output_schedule`) every rank drops its owned particles into one host snapshot in global-id order and rank 0 writes
    `output%u.vtk` (+ `output0_boundingBox.vtk`) and/or `output%u.ssb` through the engine's own writers applied to host memory
    (`ssb_write_snapshot`) — the same files, names and step map a single-GPU `ssb_run` leaves behind.  Returns the summed engine counters.
    `cancelled`: optional callable polled once per step (Solver timeout)."""
    from .vtk import write_snapshot, write_snapshot_py
    # the engine's own C++ writers on the host snapshot; the CPU tier (fake rank engines, maybe no libssb_core.so) uses the Python twins
    if writer is None:
        writer = write_snapshot if rank_engine is _default_rank_engine else write_snapshot_py
    fm = fm.finalize()
    world = len(devices)
    if fm.static_domain:
        raise ValueError("slab decomposition is implemented for moving domains (static ensembles shard by trajectory)")
    N, Sc, Sd = fm.num_particles, fm.num_chem_species, fm.num_stoch_species
    edges = slab_bounds(fm.x[:, 0], world)
    hub = LoopbackHub(world, timeout=3600.0)
    snap = {"x": np.empty((N, 3)), "v": np.empty((N, 3)), "scal": np.empty((4, N)), "C": np.empty((Sc, N)),
            "type": np.empty(N, np.int32), "D": np.empty((Sd, N), np.uint32)}
    schedule = output_schedule(fm.nt, fm.output_steps, corrected=bool(flags & FLAG_CORRECTED_OUTPUT_STEPS))
    errs, counters = {}, {}
    if rank_engine is _default_rank_engine:
        from . import codegen
        codegen.build_core()
        codegen.build_model_unit(fm)          # model units are particle independent: compile once, before the ranks race for it

    def fill(se):
        gid, x = se.owned_field("x")
        snap["x"][gid] = x
        snap["v"][gid] = se.owned_field("v")[1]
        for k, name in enumerate(("rho", "mass", "bvf_phi", "nu")):
            snap["scal"][k, gid] = se.owned_field(name)[1]
        snap["type"][gid] = se.owned_field("type")[1]
        if Sc:
            snap["C"][:, gid] = se.owned_field("C")[1].T
        if Sd:
            snap["D"][:, gid] = se.owned_field("xx")[1].T

    def write(file_index, step):
        init = 1 if (Sd > 0 and step > 0) else 0          # output0 is staged before the first RDME step (output.cpp:151-154)
        writer(out_dir, file_index, snap["x"], snap["v"], snap["scal"], snap["C"] if Sc else None, snap["type"],
               snap["D"] if Sd else None, fm.species_names, (fm.xlim, fm.ylim, fm.zlim), step=step, rdme_initialized=init,
               vtk=vtk, binary=binary_store)

    def body(rank):
        se = None
        try:
            part = partition(fm, rank, world, halo=halo, edges=edges)
            se = rank_engine(part, rank, world, devices[rank], hub, flags, rdme_epsilon)
            se.reset(seed)
            done = 0
            for file_index, step in schedule:
                if step > done:
                    for _ in range(step - done):
                        if cancelled is not None and cancelled():
                            raise InterruptedError("cancelled")
                        se.step(1)
                    done = step
                fill(se)
                hub.barrier.wait()                      # every rank has dropped its rows
                if rank == 0:
                    write(file_index, step)
                hub.barrier.wait()                      # the snapshot may be overwritten again
            counters[rank] = se.counters()
        except BaseException as err:  # noqa: BLE001 - re-raised by the caller's thread
            errs[rank] = err
            hub.barrier.abort()
        finally:
            if se is not None:
                se.close()

    threads = [threading.Thread(target=body, args=(r,), name=f"slab-rank-{r}") for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:          # the first real failure (the other ranks only see the barrier it broke)
        real = [e for e in errs.values() if not isinstance(e, threading.BrokenBarrierError)]
        raise (real or list(errs.values()))[0]
    total = {"reactions": 0, "diffusions": 0, "seconds": 0.0, "windows": 0}
    for c in counters.values():
        total["reactions"] += c["reactions"]
        total["diffusions"] += c["diffusions"]
        total["seconds"] = max(total["seconds"], c["seconds"])
        total["windows"] = max(total["windows"], c["windows"])
    return total
