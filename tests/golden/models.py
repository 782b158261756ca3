"""spatialpy.Model builders for the parity fixtures (run ONLY in the dev container, where the reference
checkout is importable; the GPU box consumes the generated .npz fixtures, never this file's imports).

Every builder seeds numpy first so scatter initial conditions are reproducible.
"""
import os

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


BUILDERS = {"birth_death": birth_death, "diffusion3d": diffusion3d, "cavity2d": cavity2d, "tank3d": tank3d,
            "cylinder": cylinder}
