"""spatialpy.Model builders for the parity fixtures (run ONLY in the dev container, where the reference
checkout is importable; the GPU box consumes the generated .npz fixtures, never this file's imports).

Every builder seeds numpy first so scatter initial conditions are reproducible.
"""
import os
import sys

import numpy


def birth_death():
    """BASELINE config 1 verbatim: /root/reference/test/models/birth_death.py:19-50."""
    import spatialpy
    numpy.random.seed(1)
    model = spatialpy.Model(name='Spatial Birth-Death')
    model.HABITAT = "Habitat"
    domain = spatialpy.Domain.create_2D_domain(xlim=(0, 1), ylim=(0, 1), numx=10, numy=10,
                                               type_id=model.HABITAT, fixed=True)
    model.add_domain(domain)
    model.add_species(spatialpy.Species(name='Rabbits', diffusion_coefficient=0.1))
    model.add_initial_condition(spatialpy.ScatterInitialCondition(species='Rabbits', count=100))
    model.add_parameter([spatialpy.Parameter(name='k_birth', expression=10),
                         spatialpy.Parameter(name='k_death', expression=0.1)])
    model.add_reaction([
        spatialpy.Reaction(name='birth', reactants={}, products={"Rabbits": 1}, rate="k_birth"),
        spatialpy.Reaction(name='death', reactants={"Rabbits": 1}, products={}, rate="k_death")])
    model.timespan(spatialpy.TimeSpan.linspace(t=10, num_points=11, timestep_size=1))
    return model


def diffusion3d(n=7, steps=10):
    """Static 3-D lattice, two diffusing species, NO reactions: gates the deterministic C[] diffusion path
    (model.cpp:152-170) and D_i_j / Ddiag, without touching the reference's stoichiometry-indexing defect."""
    import spatialpy
    numpy.random.seed(2)
    model = spatialpy.Model(name='diffusion3d')
    domain = spatialpy.Domain.create_3D_domain(xlim=(0, 1), ylim=(0, 1), zlim=(0, 1), numx=n, numy=n, numz=n,
                                               type_id="Box", fixed=True)
    # two types so the type-indexed diffusion matrix is exercised (equal D keeps the reference's transposed
    # PDE index benign)
    class Left(spatialpy.Geometry):
        def inside(self, point, on_boundary):
            return point[0] < 0.5
    domain.set_properties(Left(), "Left", mass=1.0)
    model.add_domain(domain)
    model.add_species([spatialpy.Species(name='A', diffusion_coefficient=0.01),
                       spatialpy.Species(name='B', diffusion_coefficient=0.01)])
    model.add_initial_condition(spatialpy.ScatterInitialCondition(species='A', count=5000))
    model.add_initial_condition(spatialpy.PlaceInitialCondition(species='B', count=2000, location=[0.5, 0.5, 0.5]))
    model.timespan(spatialpy.TimeSpan.linspace(t=steps * 0.01, num_points=steps + 1, timestep_size=0.01))
    return model


def cavity2d(nf=14, steps=45, gravity=(0.0, -1.0, 0.0), with_species=False):
    """Moving-domain SDPD: the lid-driven cavity of examples/Under_Construction/'Lid driven cavity.ipynb' cell 3,
    shrunk to nf x nf fluid particles + 3 wall layers, with gravity so every force term is non-zero."""
    import spatialpy

    class Walls(spatialpy.Geometry):
        def inside(self, point, on_boundary):
            return point[0] < 0.0 or point[0] > 1.0 or point[1] < 0.0 or point[1] > 1.0

    numpy.random.seed(3)
    model = spatialpy.Model(name='cavity2d')
    nu, L, nW, rho0, c0 = 0.01, 1.0, 3, 1.0, 10.0
    P0 = rho0 * c0 ** 2
    ntot = nf + 2 * nW
    dx = L / (nf - 1)
    lim = ((0 - (nW - 1) * dx), 1 + (nW - 1) * dx)
    vol = (lim[1] - lim[0]) ** 2
    mpp = rho0 * vol / (ntot * ntot)
    domain = spatialpy.Domain.create_2D_domain(lim, lim, ntot, ntot, type_id="Fluid", mass=mpp, nu=nu,
                                               rho0=rho0, c0=c0, P0=P0, fixed=False, gravity=list(gravity))
    domain.set_properties(Walls(), "Walls", mass=mpp, fixed=True)
    model.add_domain(domain)
    model.add_boundary_condition(spatialpy.BoundaryCondition(ymin=lim[1] - (nW - 1) * dx - 1e-9, target='v',
                                                             value=[1.0, 0.0, 0.0]))
    if with_species:
        model.add_species(spatialpy.Species(name='A', diffusion_coefficient=0.01))
        model.add_initial_condition(spatialpy.UniformInitialCondition(species='A', count=10))
    dt = 1e-4
    model.timespan(spatialpy.TimeSpan.linspace(t=steps * dt, num_points=steps + 1, timestep_size=dt))
    model.staticDomain = False
    return model


def cavity2d_rdme(steps=45):
    """Moving-domain RDME: the cavity with one advected species (10 molecules per voxel, D = 0.01) that decays at rate 20.  On a
    moving domain the reference rebuilds the NSM every step and executes one event past each step's end
    (simulate_rdme.cpp:54-65,233-238) — about 8 % of all events here, so the ensemble totals pin that behaviour."""
    import spatialpy
    model = cavity2d(steps=steps, with_species=True)
    model.name = "cavity2d_rdme"
    model.add_parameter(spatialpy.Parameter(name="k_decay", expression=20.0))
    model.add_reaction(spatialpy.Reaction(name="decay", reactants={"A": 1}, products={}, rate="k_decay"))
    return model


def tank3d(n=8, steps=25):
    """Moving 3-D SDPD tank: n^3 lattice, outer layer fixed walls, gravity, one advected diffusing species."""
    import spatialpy

    class Walls(spatialpy.Geometry):
        def inside(self, point, on_boundary):
            lo, hi = 0.2, 0.8
            return any(p < lo or p > hi for p in point[:3])

    numpy.random.seed(4)
    model = spatialpy.Model(name='tank3d')
    rho0, c0 = 1.0, 10.0
    P0 = rho0 * c0 ** 2
    dx = 1.0 / (n - 1)
    mpp = rho0 * dx ** 3
    domain = spatialpy.Domain.create_3D_domain((0, 1), (0, 1), (0, 1), n, n, n, type_id="Fluid", mass=mpp, nu=0.05,
                                               rho0=rho0, c0=c0, P0=P0, fixed=False, gravity=[0.0, 0.0, -1.0])
    domain.set_properties(Walls(), "Walls", mass=mpp, fixed=True)
    model.add_domain(domain)
    model.add_species(spatialpy.Species(name='A', diffusion_coefficient=0.01))
    model.add_initial_condition(spatialpy.UniformInitialCondition(species='A', count=10))
    dt = 1e-4
    model.timespan(spatialpy.TimeSpan.linspace(t=steps * dt, num_points=steps + 1, timestep_size=dt))
    model.staticDomain = False
    return model


def cylinder(steps=5, dt=1.0):
    """BASELINE config 2a: the shipped 1 059-vertex cylinder mesh (/root/reference/test/models/cylinder_demo3D.py)."""
    import spatialpy
    numpy.random.seed(5)
    MAX_X, MIN_X = 5.0, -5.0

    class Edge1(spatialpy.Geometry):
        def inside(self, x, on_boundary):
            return abs(x[0] - MAX_X) < 0.05

    class Edge2(spatialpy.Geometry):
        def inside(self, x, on_boundary):
            return abs(x[0] - MIN_X) < 0.05

    class Middle(spatialpy.Geometry):
        def inside(self, x, on_boundary):
            return abs(x[0] - MIN_X) >= 0.05

    model = spatialpy.Model("cylinder_demo3d")
    ref_root = os.environ.get("SSB_REFERENCE_ROOT", "/root/reference")
    domain = spatialpy.Domain.read_xml_mesh(os.path.join(ref_root, "test/models/data/cylinder.xml"))
    domain.set_properties(Middle(), "Middle")
    domain.set_properties(Edge1(), "Edge1")
    domain.set_properties(Edge2(), "Edge2")
    model.add_domain(domain)
    A = spatialpy.Species(name="A", diffusion_coefficient=0.1, restrict_to=["Middle", "Edge1"])
    B = spatialpy.Species(name="B", diffusion_coefficient=0.1, restrict_to=["Middle", "Edge2"])
    model.add_species([A, B])
    vol = model.domain.get_vol()
    type_id = model.domain.type_id
    left = numpy.sum(vol[numpy.array([t == model.domain.get_type_def("Edge1") for t in type_id])])
    right = numpy.sum(vol[numpy.array([t == model.domain.get_type_def("Edge2") for t in type_id])])
    model.add_parameter([spatialpy.Parameter(name="k_react", expression=1.0),
                         spatialpy.Parameter(name="k_creat1", expression=100 / left),
                         spatialpy.Parameter(name="k_creat2", expression=100 / right)])
    model.add_reaction([
        spatialpy.Reaction(name="R1", reactants=None, products={A: 1}, rate="k_creat1", restrict_to="Edge1"),
        spatialpy.Reaction(name="R2", reactants=None, products={B: 1}, rate="k_creat2", restrict_to="Edge2"),
        spatialpy.Reaction(name="R3", reactants={A: 1, B: 1}, products=None, rate="k_react")])
    model.timespan(spatialpy.TimeSpan.linspace(t=steps * dt, num_points=steps + 1, timestep_size=dt))
    return model


def cdc42(DX=12, end_time=0.02, steps=2):
    """BASELINE config 4 (examples/Yeast_Polarization/Cdc42.ipynb cell 12, create_cdc42_model) on a coarse DX x DX lattice and a
    short horizon so a >= 600-trajectory reference ensemble is affordable: 3 types with different masses (=> different voxel
    volumes across the membrane/cytoplasm interface), 9 species restricted by type, 13 reactions (12 mass action + CR0 with a
    custom propensity reading the data function GbgGradient), scatter initial conditions restricted by type."""
    import math
    import spatialpy
    numpy.random.seed(6)

    class Membrane(spatialpy.Geometry):
        def inside(self, point, on_boundary):
            r = numpy.sqrt(point[0] ** 2 + point[1] ** 2)
            return 0.6 >= r >= 0.4

    class Cytoplasm(spatialpy.Geometry):
        def inside(self, point, on_boundary):
            return numpy.sqrt(point[0] ** 2 + point[1] ** 2) < 0.4

    class GbgGradient(spatialpy.DataFunction):
        def __init__(self, Gbg_mid=5000, Gbg_slope=0.0, mem_vol=1.0):
            spatialpy.DataFunction.__init__(self, name="GbgGradient")
            self.Gbg_mid, self.Gbg_slope, self.mem_vol = Gbg_mid, Gbg_slope, mem_vol

        def map(self, x):
            return (self.Gbg_slope * x[1] + self.Gbg_mid) / self.mem_vol

    model = spatialpy.Model("Cdc42_2D")
    EXTRA, MEM, CYT = "Extra_Cellular", "Membrane", "Cytoplasm"
    D_membrane, D_bulk = 0.0053, 10.0
    domain = spatialpy.Domain.create_2D_domain(xlim=(-1, 1), ylim=(-1, 1), numx=DX, numy=DX, rho0=1.0, c0=10, P0=10, type_id=EXTRA)
    domain.set_properties(Membrane(), type_id=MEM, mass=4.0, nu=1.0, fixed=False)
    domain.set_properties(Cytoplasm(), type_id=CYT, mass=2.0, nu=1.0, fixed=False)
    model.add_domain(domain)
    S = spatialpy.Species
    model.add_species([
        S(name="Cdc24_m", diffusion_coefficient=D_membrane, restrict_to=MEM),
        S(name="Cdc24_c", diffusion_coefficient=D_bulk, restrict_to=[MEM, CYT]),
        S(name="Cdc42", diffusion_coefficient=D_membrane, restrict_to=MEM),
        S(name="Cdc42_a", diffusion_coefficient=D_membrane, restrict_to=MEM),
        S(name="Bem1_m", diffusion_coefficient=D_membrane, restrict_to=MEM),
        S(name="Bem1_c", diffusion_coefficient=D_bulk, restrict_to=[MEM, CYT]),
        S(name="Cla4", diffusion_coefficient=D_bulk, restrict_to=[MEM, CYT]),
        S(name="Cla4_a", diffusion_coefficient=D_membrane, restrict_to=MEM),
        S(name="Cdc42_c", diffusion_coefficient=D_bulk, restrict_to=[MEM, CYT])])
    IC = spatialpy.ScatterInitialCondition
    model.add_initial_condition([IC("Cdc42", 2700, [CYT]), IC("Cdc24_c", 1000, [MEM]), IC("Bem1_c", 3000, [MEM]),
                                 IC("Cla4", 5000, [MEM]), IC("Cdc42_a", 300, [CYT])])
    P = spatialpy.Parameter
    model.add_parameter([P(name="k_42a", expression=0.2), P(name="k_42d", expression=1.0), P(name="k_24cm1", expression=0.00297),
                         P(name="k_24mc", expression=0.35), P(name="k_B1mc", expression=0.35), P(name="k_B1cm", expression=0.2667),
                         P(name="k_Cla4a", expression=0.006), P(name="k_Cla4d", expression=0.01), P(name="k_24d", expression=1.0 / 30000),
                         P(name="beta1", expression=0.266), P(name="beta2", expression=0.28), P(name="beta3", expression=1.0),
                         P(name="delta1_gbg", expression=0.00297)])
    R = spatialpy.Reaction
    model.add_reaction([
        R(name="CR0", reactants={'Cdc24_c': 1}, products={'Cdc24_m': 1},
          propensity_function="delta1_gbg * Cdc24_c * GbgGradient * vol", restrict_to=MEM),
        R(name="CR1", reactants={'Cdc24_c': 1, 'Bem1_m': 1}, products={'Cdc24_m': 1, 'Bem1_m': 1}, rate='k_24cm1', restrict_to=MEM),
        R(name="CR2", reactants={'Cdc24_m': 1}, products={'Cdc24_c': 1}, rate='k_24mc', restrict_to=MEM),
        R(name="CR3", reactants={'Cdc24_m': 1, 'Cla4_a': 1}, products={'Cdc24_c': 1, 'Cla4_a': 1}, rate='k_24d', restrict_to=MEM),
        R(name="CR4", reactants={'Cdc24_m': 1, 'Cdc42': 1}, products={'Cdc24_m': 1, 'Cdc42_a': 1}, rate='k_42a', restrict_to=MEM),
        R(name="CR5", reactants={'Cdc42_a': 1}, products={'Cdc42': 1}, rate='k_42d', restrict_to=MEM),
        R(name="CR6", reactants={'Cdc42_a': 1, 'Bem1_c': 1}, products={'Cdc42_a': 1, 'Bem1_m': 1}, rate='k_B1cm', restrict_to=MEM),
        R(name="CR7", reactants={'Bem1_m': 1}, products={'Bem1_c': 1}, rate='k_B1mc', restrict_to=MEM),
        R(name="CR8", reactants={'Cdc42_a': 1, 'Cla4': 1}, products={'Cdc42_a': 1, 'Cla4_a': 1}, rate='k_Cla4a', restrict_to=MEM),
        R(name="CR9", reactants={'Cla4_a': 1}, products={'Cla4': 1}, rate='k_Cla4d', restrict_to=MEM),
        R(name="CR10", reactants={'Cdc42_c': 1}, products={'Cdc42': 1}, rate='beta2', restrict_to=MEM),
        R(name="CR11", reactants={'Cdc42': 1}, products={'Cdc42_c': 1}, rate='beta3', restrict_to=MEM),
        R(name="CR12", reactants={'Cdc42_c': 1, 'Cdc24_m': 1}, products={'Cdc42_a': 1, 'Cdc24_m': 1}, rate='beta1', restrict_to=MEM)])
    membrane_volume = sum(model.domain.vol[i] for i, t in enumerate(model.domain.type_id) if t == model.domain.get_type_def(MEM))
    model.add_data_function(GbgGradient(Gbg_mid=5000.0, Gbg_slope=0.0, mem_vol=membrane_volume))
    model.timespan(spatialpy.TimeSpan.linspace(t=end_time, num_points=steps + 1, timestep_size=end_time / steps))
    return model


def line1d(steps=4, isolated=False):
    """Edge cases in one model, fed to the reference through its own template (oracle/emit_ref_model.py): a 1-D MOVING domain
    (all y = z = 0 => dimension 1, solver.py:407-413; kernel normalisation alpha = h, particle.cpp:174), ragged neighbour lists
    (end particles), fixed end walls, one diffusing species with populations from 0 upwards.  isolated=True moves the last
    particle out of everybody's reach: its Shepard-filtered density is 0/0 at step 0 (model.cpp:194-233 runs for every
    particle of a moving domain when step % 20 == 0), so the REFERENCE aborts at step 1 with "nan/inf detected"
    (particle.cpp:88-126) — the error-behaviour fixture."""
    from spatialpy_b200 import FlatModel, ReactionSource
    n = 14
    x = numpy.zeros((n, 3))
    x[:, 0] = numpy.arange(n) * 0.1 + 0.003 * numpy.sin(numpy.arange(n) * 1.7)
    if isolated:
        x[-1, 0] = 5.0
    solid = numpy.zeros(n, numpy.int32)
    solid[[0, 1, n - 2, n - 1]] = 1
    u0 = (numpy.arange(n) % 5 * 3).astype(numpy.uint32).reshape(n, 1)
    dt = 1e-3
    return FlatModel(
        name="line1d_isolated" if isolated else "line1d", x=x, type=numpy.where(solid == 1, 1, 2).astype(numpy.int32), nu=numpy.full(n, 0.05), mass=numpy.full(n, 0.1),
        c=numpy.zeros(n), rho=numpy.full(n, 1.0), solid=solid, species_names=["A"],
        reactions=[ReactionSource(name="decay", propensity="P0*x[0]", ode_propensity="P0*x[0]", restrict_to=None)],
        parameters={"P0": 0.0}, type_constants={"type_Walls": 1, "type_Fluid": 2}, u0=u0,
        N_dense=numpy.array([[-1]], numpy.int32), diffusion_matrix=numpy.array([[0.02, 0.02]]), enable_pde=True, enable_rdme=True,
        static_domain=False, dt=dt, nt=steps, output_steps=numpy.arange(steps + 1, dtype=numpy.uint32), h=0.25, rho0=1.0, c0=10.0,
        P0=100.0, xlim=(0.0, 5.0), ylim=(0.0, 0.0), zlim=(0.0, 0.0), dimension=1, gravity=(0.3, 0.0, 0.0)).finalize()


def line1d_isolated():
    return line1d(isolated=True)


def letters():
    """test/integration_tests/test_model.py:60-78: every single-letter species name that is not reserved (51 species) must
    survive the compilation namespace — here: the generated device code (x, t, vol ... are identifiers of the propensity ABI)."""
    import string
    import spatialpy
    model = spatialpy.Model("letters")
    for ch in string.ascii_letters:
        if ch not in spatialpy.Model.reserved_names:
            model.add_species(spatialpy.Species(ch, 0))
    model.set_timesteps(output_interval=1, num_steps=1, timestep_size=1)
    model.add_domain(spatialpy.Domain.create_2D_domain([0, 1], [0, 1], 2, 2))
    return model


def datafn():
    """test/integration_tests/test_model.py:81-104: a data function (10000 * x) used as a propensity must reach the engine:
    no births where x = 0, births where x = 1."""
    import spatialpy
    model = spatialpy.Model("datafn")
    df = spatialpy.DataFunction(name="df")
    df.map = lambda x: x[0] * 10000
    model.add_data_function(df)
    model.set_timesteps(output_interval=1, num_steps=1, timestep_size=1)
    model.add_domain(spatialpy.Domain.create_2D_domain([0, 1], [0, 1], 2, 2))
    model.add_species(spatialpy.Species('A', 0))
    model.add_reaction(spatialpy.Reaction(products={'A': 1}, propensity_function="df"))
    return model


def cdc42_full():
    """BASELINE config 4 at its NAMED size — create_cdc42_model(DX=50): 2 500 particles — on a horizon the reference finishes in
    seconds (the notebook's end_time=100 is ~7e10 events per trajectory): full-size parity taps for the deterministic parts
    (neighbour lists, D_i_j, Ddiag, propensities) and the model the `ens_cdc42_full` bench workload runs."""
    return cdc42(DX=50, end_time=0.001, steps=2)


def cavity2d_bc(steps=22):
    """Every boundary-condition target the reference emits (spatialpy/core/boundarycondition.py:147-166) on one moving domain:
    `v` on the lid (as cavity2d), `rho` on the bottom wall rows, `nu` on the left wall columns and a deterministic species
    concentration `C[0]` on a band of the fluid — so me->rho / me->nu / me->C[k] assignments are pinned against the reference's
    taps, through the predictor, the corrector (rho_new carries the BC) and the final BC of a step."""
    import spatialpy
    model = cavity2d(steps=steps, with_species=True)
    model.name = "cavity2d_bc"
    model.add_boundary_condition(spatialpy.BoundaryCondition(ymax=-1e-9, target='rho', value=1.02))
    model.add_boundary_condition(spatialpy.BoundaryCondition(xmax=-1e-9, target='nu', value=0.02))
    model.add_boundary_condition(spatialpy.BoundaryCondition(xmin=0.4, xmax=0.6, ymin=0.2, ymax=0.8, target='A', deterministic=True,
                                                             value=25.0, model=model))
    return model


def mid_cylinder():
    """MID-SIZE fixture (15 957 particles, 125 CTAs): the bounded cylinder instance of the bench's reference arm
    (oracle/oracle_build.py BENCH_REF) — multi-CTA cell lists, crowded clamped boundary cells, lists of up to ~70 entries."""
    from spatialpy_b200 import configs
    return configs.cylinder_rdme(delta=0.125, nt=25, output_every=25, dt=1e-3)


def mid_tank():
    """MID-SIZE fixture (13 576 moving particles): the bounded SDPD tank of the bench's reference arm, 25 steps through the
    Shepard filter at steps 0 and 20 and several candidate-list rebuilds."""
    from spatialpy_b200 import configs
    return configs.tank_sdpd(n=26, nt=25, output_every=1, dt=1e-5)


BUILDERS = {"cavity2d_bc": cavity2d_bc, "mid_cylinder": mid_cylinder, "mid_tank": mid_tank, "birth_death": birth_death, "diffusion3d": diffusion3d, "cavity2d": cavity2d, "tank3d": tank3d,
            "cylinder": cylinder, "cdc42": cdc42, "cavity2d_rdme": cavity2d_rdme, "cdc42_full": cdc42_full, "line1d": line1d, "line1d_isolated": line1d_isolated,
            "letters": letters, "datafn": datafn}
