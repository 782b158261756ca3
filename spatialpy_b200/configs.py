"""Synthetic BASELINE configurations built straight into arrays (no spatialpy import, no per-particle Python loops).

The reference front-end cannot construct these sizes: `Domain.add_point` is O(N^2) (spatialpy/core/domain.py:247-255)
and the reference codegen emits one C++ source line per particle (spatialpy/solvers/solver.py:312-331).  Definitions
follow SURVEY.md §8(d); every builder takes a size knob so the same geometry also exists at oracle-checkable sizes.
"""
import numpy as np

from .flatmodel import FlatModel, ReactionSource


def _find_h(x, dim):
    """Domain.find_h (spatialpy/core/domain.py:751-769): 2.2 x the largest nearest-neighbour distance."""
    from scipy.spatial import cKDTree
    pts = x[:, :dim] if dim < 3 else x
    d, _ = cKDTree(pts).query(pts, 2)
    return 2.2 * float(d[:, 1].max())


def _output_steps(nt, every):
    # TimeSpan.output_steps for an output every `every` steps, including 0 and nt (core/timespan.py)
    return np.arange(0, nt + 1, every, dtype=np.uint32)


def cylinder_rdme(delta=0.03155, nt=1000, output_every=100, dt=1e-3, enable_pde=True, seed=2):
    """BASELINE config 2b: 3D_Cylinder_Demo geometry (axis x in [-5,5], radius 1) refined to a jittered cubic lattice.

    delta=0.03155 gives N ~ 1.0 M; delta=0.25 gives ~2 k for oracle-sized checks.  A + B -> 0 with sources at the two
    end caps, species restricted as in examples/3D_Cylinder_Demo.ipynb cell 11 / test/models/cylinder_demo3D.py.
    """
    rng = np.random.default_rng(seed)
    nx = int(round(10.0 / delta)) + 1
    nr = int(np.floor(1.0 / delta))
    gx = -5.0 + delta * np.arange(nx)
    gy = delta * np.arange(-nr, nr + 1)
    X, Y, Z = np.meshgrid(gx, gy, gy, indexing="ij")
    keep = (Y * Y + Z * Z) <= 1.0
    x = np.stack([X[keep], Y[keep], Z[keep]], axis=1)
    x += rng.uniform(-0.1 * delta, 0.1 * delta, size=x.shape)
    N = x.shape[0]
    # types (fixed numbering; the reference's numbering depends on PYTHONHASHSEED, domain.py:141-146)
    T_EDGE2, T_EDGE1, T_MIDDLE = 1, 2, 3
    width = max(0.05, 2.0 * delta)            # edge slabs widened to 2*delta so they are never empty
    typ = np.full(N, T_MIDDLE, np.int32)
    typ[np.abs(x[:, 0] - 5.0) < width] = T_EDGE1
    typ[np.abs(x[:, 0] + 5.0) < width] = T_EDGE2
    vol = delta ** 3
    mass = np.full(N, vol)
    left = vol * np.count_nonzero(typ == T_EDGE1)
    right = vol * np.count_nonzero(typ == T_EDGE2)
    D = 0.1
    dmat = np.zeros((2, 3))
    dmat[0, [T_MIDDLE - 1, T_EDGE1 - 1]] = D      # A lives in Middle, Edge1
    dmat[1, [T_MIDDLE - 1, T_EDGE2 - 1]] = D      # B lives in Middle, Edge2
    reactions = [
        ReactionSource("R1", "(P1*vol)", "P1", ["type_Edge1"]),
        ReactionSource("R2", "(P2*vol)", "P2", ["type_Edge2"]),
        ReactionSource("R3", "(((P0*x[0])*x[1])/vol)", "((P0*x[0])*x[1])", None),
    ]
    N_dense = np.array([[1, 0, -1], [0, 1, -1]], np.int32)
    # dependency graph (model.py:187-250): species -> reactions that read it; reaction -> reactions it disturbs
    G = np.zeros((3, 5))
    G[2, 0] = 1; G[2, 1] = 1                      # R3 depends on A and on B
    G[2, 2] = 1; G[2, 3] = 1; G[2, 4] = 1         # R1, R2, R3 each change a reactant of R3
    import scipy.sparse
    Gc = scipy.sparse.csc_matrix(G)
    fm = FlatModel(
        name=f"cylinder_rdme_{N}", x=x, type=typ, nu=np.ones(N), mass=mass, c=np.zeros(N), rho=np.ones(N),
        solid=np.ones(N, np.int32), species_names=["A", "B"], reactions=reactions,
        parameters={"P0": 1.0, "P1": 100.0 / left, "P2": 100.0 / right},
        type_constants={"type_UnAssigned": 0, "type_Edge2": T_EDGE2, "type_Edge1": T_EDGE1, "type_Middle": T_MIDDLE},
        u0=np.zeros((N, 2), np.uint32), N_dense=N_dense, irG=Gc.indices, jcG=Gc.indptr, diffusion_matrix=dmat,
        enable_pde=enable_pde, enable_rdme=True, static_domain=True, dt=dt, nt=nt,
        output_steps=_output_steps(nt, output_every), h=_find_h(x, 3), rho0=1.0, c0=10.0, P0=100.0,
        xlim=(float(x[:, 0].min()), float(x[:, 0].max())), ylim=(float(x[:, 1].min()), float(x[:, 1].max())),
        zlim=(float(x[:, 2].min()), float(x[:, 2].max())), dimension=3, gravity=(0.0, 0.0, 0.0))
    return fm.finalize()


def tank_sdpd(n=120, nt=1000, output_every=100, dt=1e-5, with_species=True, seed=3, fill=0.5):
    """BASELINE config 3 stand-in (Weir/Gravity notebooks are missing from the checkout): 3-D tank [0,1]^3 on an n^3 cubic
    lattice, outer 3 layers = fixed Walls, fluid column filling z < fill, the rest empty; gravity -z; one advected species
    A with D = 0.01, 10 molecules per voxel, A -> 0 @0.1 and 0 -> A @1.0.  n=120 gives ~1.0 M particles."""
    delta = 1.0 / (n - 1)
    g = delta * np.arange(n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    I, J, K = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    wall = (I < 3) | (I >= n - 3) | (J < 3) | (J >= n - 3) | (K < 3) | (K >= n - 3)
    fluid = (~wall) & (Z < fill)
    keep = wall | fluid
    x = np.stack([X[keep], Y[keep], Z[keep]], axis=1)
    solid = wall[keep].astype(np.int32)
    N = x.shape[0]
    T_WALLS, T_FLUID = 1, 2
    typ = np.where(solid == 1, T_WALLS, T_FLUID).astype(np.int32)
    rho0, c0 = 1.0, 10.0
    mass = np.full(N, rho0 * delta ** 3)
    kw = {}
    species, reactions, params = [], [], {}
    if with_species:
        species = ["A"]
        reactions = [ReactionSource("decay", "(P0*x[0])", "(P0*x[0])", None),
                     ReactionSource("create", "(P1*vol)", "P1", None)]
        params = {"P0": 0.1, "P1": 1.0}
        import scipy.sparse
        G = np.zeros((2, 3))
        G[0, 0] = 1; G[0, 1] = 1; G[0, 2] = 1
        Gc = scipy.sparse.csc_matrix(G)
        kw = dict(u0=np.full((N, 1), 10, np.uint32), N_dense=np.array([[-1, 1]], np.int32), irG=Gc.indices, jcG=Gc.indptr,
                  diffusion_matrix=np.full((1, 2), 0.01))
    fm = FlatModel(
        name=f"tank_sdpd_{N}", x=x, type=typ, nu=np.full(N, 0.01), mass=mass, c=np.zeros(N), rho=np.full(N, rho0),
        solid=solid, species_names=species, reactions=reactions, parameters=params,
        type_constants={"type_UnAssigned": 0, "type_Walls": T_WALLS, "type_Fluid": T_FLUID},
        enable_pde=True, enable_rdme=True, static_domain=False, dt=dt, nt=nt, output_steps=_output_steps(nt, output_every),
        h=2.2 * delta, rho0=rho0, c0=c0, P0=rho0 * c0 ** 2, xlim=(0.0, 1.0), ylim=(0.0, 1.0), zlim=(0.0, 1.0),
        dimension=3, gravity=(0.0, 0.0, -1.0), **kw)
    return fm.finalize()


def box_sdpd_rdme(nx=200, ny=200, nz=200, nt=100, output_every=100, dt=1e-5, seed=5, x_offset=0.0):
    """BASELINE config 5 building block: jittered cubic lattice box, delta = 0.01, all fluid except 3-layer walls on the
    z faces, gravity -z, species A, B (D = 0.01, 10/voxel each), A+B -> 0 @1e-3, 0 -> A, 0 -> B @1.0.
    200^3 = 8 M is the per-GPU weak-scaling unit (the 64 M case is 8 slabs of 50 x 400 x 400 along x)."""
    rng = np.random.default_rng(seed)
    delta = 0.01
    X, Y, Z = np.meshgrid(x_offset + delta * np.arange(nx), delta * np.arange(ny), delta * np.arange(nz), indexing="ij")
    x = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    K = np.broadcast_to(np.arange(nz)[None, None, :], (nx, ny, nz)).ravel()
    x += rng.uniform(-0.05 * delta, 0.05 * delta, size=x.shape)
    N = x.shape[0]
    solid = ((K < 3) | (K >= nz - 3)).astype(np.int32)
    T_WALLS, T_FLUID = 1, 2
    typ = np.where(solid == 1, T_WALLS, T_FLUID).astype(np.int32)
    rho0, c0 = 1.0, 10.0
    reactions = [ReactionSource("annihilate", "(((P0*x[0])*x[1])/vol)", "((P0*x[0])*x[1])", None),
                 ReactionSource("createA", "(P1*vol)", "P1", None),
                 ReactionSource("createB", "(P1*vol)", "P1", None)]
    import scipy.sparse
    G = np.zeros((3, 5))
    G[0, :] = 1
    Gc = scipy.sparse.csc_matrix(G)
    fm = FlatModel(
        name=f"box_sdpd_rdme_{N}", x=x, type=typ, nu=np.full(N, 0.1), mass=np.full(N, delta ** 3), c=np.zeros(N),
        rho=np.full(N, rho0), solid=solid, species_names=["A", "B"], reactions=reactions,
        parameters={"P0": 1e-3, "P1": 1.0},
        type_constants={"type_UnAssigned": 0, "type_Walls": T_WALLS, "type_Fluid": T_FLUID},
        u0=np.full((N, 2), 10, np.uint32), N_dense=np.array([[-1, 1, 0], [-1, 0, 1]], np.int32), irG=Gc.indices, jcG=Gc.indptr,
        diffusion_matrix=np.full((2, 2), 0.01), enable_pde=True, enable_rdme=True, static_domain=False, dt=dt, nt=nt,
        output_steps=_output_steps(nt, output_every), h=2.2 * delta * 1.1, rho0=rho0, c0=c0, P0=rho0 * c0 ** 2,
        xlim=(float(x[:, 0].min()), float(x[:, 0].max())), ylim=(float(x[:, 1].min()), float(x[:, 1].max())),
        zlim=(float(x[:, 2].min()), float(x[:, 2].max())), dimension=3, gravity=(0.0, 0.0, -1.0))
    return fm.finalize()


def box_slab(rank, world, nx_per_rank=200, ny=200, nz=200, ghost_cols=5, nt=100, dt=1e-5, seed=5):
    """Rank-local piece of the BASELINE config-5 box (box_sdpd_rdme) for a slab-decomposed run: the rank's own x-columns plus
    `ghost_cols` ghost columns on each interior face, generated without ever materialising the global model.  Jitter is drawn
    per x-plane from a generator seeded by (seed, plane), so neighbouring ranks agree on the shared columns.
    Returns a spatialpy_b200.slab.SlabPartition."""
    from .slab import SlabPartition
    delta = 0.01
    nx = nx_per_rank * world
    i0, i1 = rank * nx_per_rank, (rank + 1) * nx_per_rank
    lo, hi = max(0, i0 - ghost_cols), min(nx, i1 + ghost_cols)
    cols = np.arange(lo, hi)
    J, K = np.meshgrid(np.arange(ny), np.arange(nz), indexing="ij")
    xs = []
    for i in cols:
        rng = np.random.default_rng([seed, int(i)])
        plane = np.stack([np.full(J.size, i * delta), J.ravel() * delta, K.ravel() * delta], axis=1)
        plane += rng.uniform(-0.05 * delta, 0.05 * delta, size=plane.shape)
        xs.append(plane)
    x = np.concatenate(xs)
    col_of = np.repeat(cols, ny * nz)
    kk = np.tile(K.ravel(), len(cols))
    gid = (col_of.astype(np.int64) * ny * nz + np.tile((J * nz + K).ravel(), len(cols)))
    owned = ((col_of >= i0) & (col_of < i1)).astype(np.int32)
    order = np.argsort(1 - owned, kind="stable")          # owned first, each group already sorted by global id
    x, col_of, kk, gid, owned = x[order], col_of[order], kk[order], gid[order], owned[order]
    N = x.shape[0]
    solid = ((kk < 3) | (kk >= nz - 3)).astype(np.int32)
    base = box_sdpd_rdme(2, 2, 8, nt=nt, output_every=nt, dt=dt)       # reactions / parameters / tables of the workload
    fm = FlatModel(
        name=f"box_slab_r{rank}of{world}", x=x, type=np.where(solid == 1, 1, 2).astype(np.int32), nu=np.full(N, 0.1),
        mass=np.full(N, delta ** 3), c=np.zeros(N), rho=np.full(N, 1.0), solid=solid, species_names=list(base.species_names),
        reactions=list(base.reactions), parameters=dict(base.parameters), type_constants=dict(base.type_constants),
        u0=np.full((N, 2), 10, np.uint32), N_dense=base.N_dense, irG=base.irG, jcG=base.jcG, diffusion_matrix=base.diffusion_matrix,
        enable_pde=True, enable_rdme=True, static_domain=False, dt=dt, nt=nt, output_steps=_output_steps(nt, nt),
        h=base.h, rho0=1.0, c0=10.0, P0=100.0, xlim=(0.0, nx * delta), ylim=(0.0, ny * delta), zlim=(0.0, nz * delta),
        dimension=3, gravity=(0.0, 0.0, -1.0)).finalize()
    send_ids, recv_ids = {}, {}
    n_own = int(owned.sum())
    loc = np.arange(N)
    if rank > 0:
        send_ids[rank - 1] = loc[(owned == 1) & (col_of < i0 + ghost_cols)].astype(np.int32)
        recv_ids[rank - 1] = loc[(owned == 0) & (col_of < i0)].astype(np.int32)
    if rank < world - 1:
        send_ids[rank + 1] = loc[(owned == 1) & (col_of >= i1 - ghost_cols)].astype(np.int32)
        recv_ids[rank + 1] = loc[(owned == 0) & (col_of >= i1)].astype(np.int32)
    assert n_own == nx_per_rank * ny * nz
    # slab faces half a lattice spacing before each rank's first column (jitter is 0.05 delta): the same owner / ghost rule as
    # slab.partition(), so a re-partition of this domain re-derives consistent sets
    edges = np.array([-np.inf] + [(k * nx_per_rank - 0.5) * delta for k in range(1, world)] + [np.inf])
    return SlabPartition(fm, gid, owned, send_ids, recv_ids, (i0 * delta, i1 * delta), ghost_cols * delta, edges=edges)
