"""Flat (array-only) description of an ssa_sdpd model — everything that crosses the C-ABI.

The reference embeds every input as a C++ literal in a generated translation unit
(spatialpy/solvers/solver.py:100-158: one `init_create_particle(...)` source line per particle,
`input_u0[]`, `input_irN[]`, ... literals).  Here the same inputs are contiguous numpy arrays handed
to the engine by pointer (include/ssb.h `ssb_model`), and only the three *code* inputs — reaction
propensities, deterministic right-hand sides and boundary-condition snippets — are compiled
(spatialpy_b200/codegen.py).

`FlatModel.from_spatialpy(model)` flattens a `spatialpy.Model` exactly the way
`Solver.__create_propensity_file` does, field for field (citations inline), so the two engines are fed
identical inputs from the same Python process (type indices come from iterating a Python `set`,
spatialpy/core/domain.py:141-146, and vary with PYTHONHASHSEED).  `FlatModel` itself needs no
spatialpy import: synthetic large domains are built straight into arrays (spatialpy_b200/configs.py)
and fixtures round-trip through `.save()` / `.load()` (npz).
"""
import json
from dataclasses import dataclass, field

import numpy as np


@dataclass
class ReactionSource:
    """One reaction's compiled inputs (solver.py:344-373 stochastic, :160-190 deterministic)."""
    name: str
    propensity: str            # C expression over x[], P<i>, vol, t, data_fn[], sd
    ode_propensity: str        # C expression for the deterministic RHS
    restrict_to: list = None   # list of type-constant names/ints, or None


@dataclass
class FlatModel:
    name: str
    # particles (solver.py:312-331)
    x: np.ndarray                      # [N,3] f64
    type: np.ndarray                   # [N] i32, 1-based (0 = UnAssigned is rejected)
    nu: np.ndarray                     # [N] f64
    mass: np.ndarray                   # [N] f64
    c: np.ndarray                      # [N] f64
    rho: np.ndarray                    # [N] f64
    solid: np.ndarray                  # [N] i32  (domain.fixed)
    # species / reactions
    species_names: list = field(default_factory=list)
    reactions: list = field(default_factory=list)          # [ReactionSource]
    parameters: dict = field(default_factory=dict)          # sanitized name ("P0") -> float (solver.py:301-306)
    type_constants: dict = field(default_factory=dict)      # "type_<name>" -> int       (solver.py:308-309)
    u0: np.ndarray = None              # [N,S] u32 voxel-major (solver.py:211-220)
    N_dense: np.ndarray = None         # [S,R] i32 row-major (solver.py:223-233)
    irN: np.ndarray = None             # CSC of N: row indices   (solver.py:235)
    jcN: np.ndarray = None             # CSC of N: column ptrs   (solver.py:236)
    prN: np.ndarray = None             # CSC of N: values        (solver.py:237)
    irG: np.ndarray = None             # dependency graph CSC rows (solver.py:256; columns [species.., reactions..])
    jcG: np.ndarray = None
    diffusion_matrix: np.ndarray = None  # [S, num_types] f64 (solver.py:269-286), num_types excludes UnAssigned
    data_fn: np.ndarray = None         # [ndf, N] f64 (solver.py:239-249)
    bc_source: str = ""                # concatenated BoundaryCondition.expression() text (solver.py:131-133)
    enable_pde: bool = True            # model.py:111 -> num_chem_species = S else 0 (solver.py:106-111)
    enable_rdme: bool = True           # model.py:110 -> num_stoch_species = S else 0 (solver.py:112-117)
    # system config (solver.py:375-419)
    static_domain: bool = True
    dt: float = 1.0
    nt: int = 1
    output_steps: np.ndarray = None    # u32 (solver.py:290-299)
    h: float = 0.0
    rho0: float = 1.0
    c0: float = 10.0
    P0: float = 10.0
    xlim: tuple = (0.0, 1.0)
    ylim: tuple = (0.0, 1.0)
    zlim: tuple = (0.0, 1.0)
    dimension: int = 3
    gravity: tuple = (0.0, 0.0, 0.0)

    # ------------------------------------------------------------------ derived sizes
    @property
    def num_particles(self):
        return int(self.x.shape[0])

    @property
    def num_species(self):
        return len(self.species_names)

    @property
    def num_reactions(self):
        return len(self.reactions)

    @property
    def num_types(self):
        return int(self.diffusion_matrix.shape[1]) if self.num_species else max(1, int(self.type.max()))

    @property
    def num_chem_species(self):
        return self.num_species if self.enable_pde else 0

    @property
    def num_chem_rxns(self):
        return self.num_reactions if self.enable_pde else 0

    @property
    def num_stoch_species(self):
        return self.num_species if self.enable_rdme else 0

    @property
    def num_stoch_rxns(self):
        return self.num_reactions if self.enable_rdme else 0

    @property
    def num_data_fn(self):
        return 0 if self.data_fn is None else int(self.data_fn.shape[0])

    # ------------------------------------------------------------------ normalisation
    def finalize(self):
        """Coerce dtypes/shapes to what the C-ABI expects; fill empty tables."""
        N = self.num_particles
        S, R = self.num_species, self.num_reactions
        self.x = np.ascontiguousarray(self.x, dtype=np.float64).reshape(N, 3)
        self.type = np.ascontiguousarray(self.type, dtype=np.int32)
        for name in ("nu", "mass", "c", "rho"):
            setattr(self, name, np.ascontiguousarray(getattr(self, name), dtype=np.float64))
        self.solid = np.ascontiguousarray(self.solid, dtype=np.int32)
        if (self.type < 1).any():
            raise ValueError("Not all particles have been defined in a type (solver.py:317-319)")
        self.u0 = (np.zeros((N, S), np.uint32) if self.u0 is None
                   else np.ascontiguousarray(self.u0, dtype=np.uint32).reshape(N, S))
        self.N_dense = (np.zeros((S, R), np.int32) if self.N_dense is None
                        else np.ascontiguousarray(self.N_dense, dtype=np.int32).reshape(S, R))
        if self.irN is None:
            import scipy.sparse
            csc = scipy.sparse.csc_matrix(self.N_dense.astype(np.float64))
            self.irN, self.jcN, self.prN = csc.indices, csc.indptr, csc.data
            if csc.indptr.shape[0] != R + 1:
                self.jcN = np.zeros(R + 1, np.int64)
        self.irN = np.ascontiguousarray(self.irN, dtype=np.int64)
        self.jcN = np.ascontiguousarray(self.jcN, dtype=np.int64)
        self.prN = np.ascontiguousarray(self.prN, dtype=np.int32)
        if self.irG is None:  # conservative: everything depends on everything (model.py:190-195)
            self.irG = np.tile(np.arange(R, dtype=np.int64), S + R)
            self.jcG = np.arange(0, (S + R) * R + 1, R, dtype=np.int64) if R else np.zeros(S + R + 1, np.int64)
        self.irG = np.ascontiguousarray(self.irG, dtype=np.int64)
        self.jcG = np.ascontiguousarray(self.jcG, dtype=np.int64)
        if self.diffusion_matrix is None:
            self.diffusion_matrix = np.zeros((S, max(1, int(self.type.max()))), np.float64)
        self.diffusion_matrix = np.ascontiguousarray(self.diffusion_matrix, dtype=np.float64)
        self.data_fn = (np.zeros((0, N), np.float64) if self.data_fn is None
                        else np.ascontiguousarray(self.data_fn, dtype=np.float64).reshape(-1, N))
        self.output_steps = np.ascontiguousarray(self.output_steps, dtype=np.uint32)
        self.gravity = tuple(float(g) for g in self.gravity)
        return self

    # ------------------------------------------------------------------ (de)serialisation
    _ARRAYS = ("x", "type", "nu", "mass", "c", "rho", "solid", "u0", "N_dense", "irN", "jcN", "prN",
               "irG", "jcG", "diffusion_matrix", "data_fn", "output_steps")
    _SCALARS = ("name", "species_names", "parameters", "type_constants", "bc_source", "enable_pde",
                "enable_rdme", "static_domain", "dt", "nt", "h", "rho0", "c0", "P0", "xlim", "ylim",
                "zlim", "dimension", "gravity")

    def save(self, path):
        meta = {k: getattr(self, k) for k in self._SCALARS}
        meta["reactions"] = [dict(name=r.name, propensity=r.propensity, ode_propensity=r.ode_propensity,
                                  restrict_to=r.restrict_to) for r in self.reactions]
        np.savez_compressed(path, __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
                            **{k: getattr(self, k) for k in self._ARRAYS})

    @classmethod
    def load(cls, path):
        z = np.load(path)
        meta = json.loads(bytes(z["__meta__"]).decode())
        reactions = [ReactionSource(**r) for r in meta.pop("reactions")]
        for k in ("xlim", "ylim", "zlim", "gravity"):
            meta[k] = tuple(meta[k])
        fm = cls(reactions=reactions, **meta, **{k: z[k] for k in cls._ARRAYS})
        return fm.finalize()

    # ------------------------------------------------------------------ from the reference front-end
    @classmethod
    def from_spatialpy(cls, model, h=None):
        """Flatten a spatialpy.Model the way the reference's codegen does (solver.py:100-419)."""
        import numpy
        stoich, dep = model.compile_prep()                       # model.py:982-1019
        dom = model.domain
        N = dom.get_num_voxels()
        S = len(model.listOfSpecies)
        R = len(model.listOfReactions)
        # types: names -> indices via the same mapping the reference emits (solver.py:308-309,316)
        type_constants = {str(k): int(v) for k, v in dom.typeNdxMapping.items()}
        if dom.type_id is None:
            dom.type_id = ["type_1"] * N
        types = numpy.array([type_constants.get(str(t), -1) if not str(t).isdigit() else int(t)
                             for t in dom.type_id], dtype=numpy.int32)
        if any("UnAssigned" in str(t) for t in dom.type_id):
            from spatialpy.core.spatialpyerror import SimulationError
            raise SimulationError("Not all particles have been defined in a type. "
                                  "Mass and other properties must be defined")
        coords = numpy.asarray(dom.coordinates(), dtype=numpy.float64)
        num_types = len(dom.listOfTypeIDs) - 1                    # solver.py:266 (UnAssigned dropped)
        # species / reactions
        species_names = list(model.listOfSpecies.keys())
        reactions = []
        for rname, reac in model.listOfReactions.items():
            restrict = None
            if not (reac.restrict_to is None or (isinstance(reac.restrict_to, list) and len(reac.restrict_to) == 0)):
                restrict = [str(t) for t in reac.restrict_to]
            reactions.append(ReactionSource(
                name=rname,
                propensity=model.expr.getexpr_cpp(reac.propensity_function),       # solver.py:350
                ode_propensity=model.expr.getexpr_cpp(reac.ode_propensity_function),  # solver.py:166
                restrict_to=restrict))
        sanitized = model.sanitized_parameter_names()
        parameters = {sanitized[p]: float(model.listOfParameters[p].value) for p in model.listOfParameters}
        u0 = numpy.ascontiguousarray(numpy.asarray(model.u0).T).astype(numpy.uint32) if S else None  # [N,S]
        kw = {}
        if S and min(stoich.shape) > 0:
            kw = dict(N_dense=numpy.asarray(stoich.todense()).astype(numpy.int32),
                      irN=stoich.indices, jcN=stoich.indptr, prN=stoich.data.astype(numpy.int32))
        if S:
            kw.update(irG=dep.indices, jcG=dep.indptr)
        # diffusion matrix, species-major [S, num_types] (solver.py:269-286)
        dmat = numpy.zeros((S, max(num_types, 1)))
        for i, species in enumerate(model.listOfSpecies.values()):
            for j, type_id in enumerate(dom.typeNdxMapping.keys()):
                if j == 0:
                    continue
                if species not in model.listOfDiffusionRestrictions or \
                        type_id in model.listOfDiffusionRestrictions[species]:
                    dmat[i, j - 1] = float(species.diffusion_coefficient)
        # data functions [ndf, N] (solver.py:239-249; the reference's comma bug for ndf>=2 is not mirrored)
        ndf = len(model.listOfDataFunctions)
        data_fn = numpy.zeros((ndf, N))
        for k, dfn in enumerate(model.listOfDataFunctions.values()):
            for i in range(N):
                data_fn[k, i] = dfn.map([coords[i, 0], coords[i, 1], coords[i, 2]])
        bc = "".join(b.expression() for b in model.listOfBoundaryConditions)   # solver.py:131-133
        # dimension inference (solver.py:407-413)
        if not numpy.count_nonzero(coords[:, 1]):
            dim = 1
        elif not numpy.count_nonzero(coords[:, 2]):
            dim = 2
        else:
            dim = 3
        if h is None:
            h = dom.find_h()                                       # solver.py:391-392
        grav = dom.gravity if dom.gravity is not None else (0.0, 0.0, 0.0)
        fm = cls(
            name=model.name, x=coords, type=types, nu=dom.nu, mass=dom.mass, c=dom.c, rho=dom.rho,
            solid=numpy.asarray(dom.fixed).astype(numpy.int32),
            species_names=species_names, reactions=reactions, parameters=parameters,
            type_constants=type_constants, u0=u0, diffusion_matrix=dmat, data_fn=data_fn, bc_source=bc,
            enable_pde=bool(model.enable_pde), enable_rdme=bool(model.enable_rdme),
            static_domain=bool(model.staticDomain), dt=float(model.tspan.timestep_size),
            nt=int(model.tspan.num_timesteps), output_steps=numpy.asarray(model.tspan.output_steps),
            h=float(h), rho0=float(dom.rho0), c0=float(dom.c0), P0=float(dom.P0),
            xlim=tuple(float(v) for v in dom.xlim), ylim=tuple(float(v) for v in dom.ylim),
            zlim=tuple(float(v) for v in dom.zlim), dimension=dim, gravity=tuple(grav), **kw)
        return fm.finalize()
