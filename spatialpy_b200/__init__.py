"""B200-native ssa_sdpd engine for SpatialPy (drop-in `Solver`; CUDA sm_100a, fp64, no CPU fallback)."""
from .flatmodel import FlatModel, ReactionSource  # noqa: F401

__version__ = "0.1.0"
from .solver import Solver, install, SimulationError  # noqa: E402,F401
from .builders import domain_from_arrays, import_meshio_object, read_msh_file, read_stochss_domain, read_xml_mesh  # noqa: E402,F401
