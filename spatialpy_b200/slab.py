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

Round-1 limitation (stated in DESIGN.md): ownership and ghost sets are fixed at partition time; `SlabEngine.step` raises
if a particle has travelled further than the halo allows.  Re-partitioning at list rebuilds is the next step.
"""
import numpy as np

from .engine import (Engine, FLAG_SKIP_STATIC_FORCES, PH_CORRECTOR, PH_END, PH_FINISH, PH_PRE, PH_RDME_CLOSE, PH_RDME_INIT,
                     PH_RDME_PREP, PH_RDME_WINDOW, PH_RDME_MIN, PH_RDME_EXTRA)
from .flatmodel import FlatModel


class SlabPartition:
    """What one rank needs: its local model (owned first, then ghosts, each sorted by global id) and the exchange lists."""

    def __init__(self, local, gids, owned, send_ids, recv_ids, bounds, halo):
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
    return SlabPartition(local, lg, owned, send_ids, recv_ids, (lo, hi), halo)


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


class SlabEngine:
    """One rank of a slab-decomposed run: an Engine on the local model + the halo exchanges between the step phases."""

    def __init__(self, part, rank, world, device=0, flags=FLAG_SKIP_STATIC_FORCES, rdme_epsilon=0.0):
        import torch
        self.torch = torch
        self.part, self.rank, self.world = part, rank, world
        if part.local.static_domain:
            raise ValueError("slab decomposition is implemented for moving domains (static ensembles shard by trajectory)")
        self.dev = torch.device("cuda", device)
        self.overshoot = not (flags & 129)     # FLAG_CORRECTED_NSM_SELECT | FLAG_NO_STEP_OVERSHOOT switch the extra event off
        self.eng = Engine(part.local, device=device, flags=flags, rdme_epsilon=rdme_epsilon, owned=part.owned,
                          rng_id=part.gids.astype(np.int32))
        self.Sd = part.local.num_stoch_species
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
        self.halo_bytes_per_step = 0
        self.travel_bound = 0.0       # upper bound on how far any particle has moved since the partition was made

    def close(self):
        self.eng.close()

    def reset(self, seed):
        self.eng.reset(seed)

    def _sync_group(self, g):
        send, recv = self.buf[g]
        for nb, ids in self.send_ids.items():
            self.eng.halo_pack(g, ids.data_ptr(), ids.numel(), send[nb].data_ptr())
        exchange(send, recv, self.rank)
        self.torch.cuda.current_stream(self.dev).synchronize()
        for nb, ids in self.recv_ids.items():
            self.eng.halo_unpack(g, ids.data_ptr(), ids.numel(), recv[nb].data_ptr())

    def _sync_inbox(self):
        send, recv = self.ibuf
        for nb, ids in self.recv_ids.items():          # my ghosts' mail -> owner
            self.eng.inbox_pack(ids.data_ptr(), ids.numel(), send[nb].data_ptr())
        exchange(send, recv, self.rank)
        self.torch.cuda.current_stream(self.dev).synchronize()
        for nb, ids in self.send_ids.items():          # mail for my owned particles that are ghosts over there
            self.eng.inbox_add(ids.data_ptr(), ids.numel(), recv[nb].data_ptr())

    def step(self, n=1):
        import torch.distributed as dist
        e = self.eng
        for _ in range(n):
            e.phase(PH_PRE)
            self._sync_group(0)
            e.phase(PH_CORRECTOR)
            self._sync_group(1)
            e.phase(PH_FINISH)
            self._sync_group(2)
            if self.Sd > 0:
                mx = e.phase(PH_RDME_PREP)
                if self.world > 1:
                    t = self.torch.tensor([mx], dtype=self.torch.float64, device=self.dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    mx = float(t.item())
                nwin = int(e.phase(PH_RDME_INIT, mx))
                for w in range(nwin):
                    e.phase(PH_RDME_WINDOW, w)
                    self._sync_inbox()
                e.phase(PH_RDME_CLOSE)
                # the reference's one event past the end of every step (simulate_rdme.cpp:233-238): globally earliest pending event
                if self.overshoot:
                    tmin = e.phase(PH_RDME_MIN)
                    if self.world > 1:
                        t = self.torch.tensor([tmin], dtype=self.torch.float64, device=self.dev)
                        dist.all_reduce(t, op=dist.ReduceOp.MIN)
                        tmin = float(t.item())
                    e.phase(PH_RDME_EXTRA, tmin)
                    self._sync_inbox()
                    e.phase(PH_RDME_CLOSE)
            e.phase(PH_END)
            # fixed ghost sets: a pair within h*(1+skin) must have both members present, so nobody may travel further than
            # half of what the halo leaves beyond the candidate radius
            self.travel_bound += e.skin_stats()["step_disp_max"]
            if self.travel_bound > 0.5 * (self.part.halo - 1.1 * self.part.local.h):
                raise RuntimeError(f"slab partition is stale: particles may have travelled {self.travel_bound:g} since the partition "
                                   f"(halo {self.part.halo:g}, h {self.part.local.h:g}); re-partition needed")

    # -- gather helpers for tests -------------------------------------------------------------------------------
    def owned_field(self, name):
        """(global ids, values) of the owned particles for a tap field."""
        a = self.eng.get(name)
        m = self.part.owned.astype(bool)
        return self.part.gids[m], a[m]
