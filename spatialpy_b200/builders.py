"""Array-at-a-time construction of the reference's `Domain` for GPU-sized inputs.

`Domain.add_point` (spatialpy/core/domain.py:203-255) appends to eight numpy arrays per particle — every call copies all of
them, so building N particles costs O(N²) (minutes at 10⁵, hopeless at 10⁶).  `domain_from_arrays` applies the same
per-particle rules (volume sign, type-id characters, `type_` prefix, `rho = mass/vol` default) to whole arrays and fills a
`Domain` in one pass; the result is attribute-for-attribute what the add_point loop produces (tests/test_cpu_abi.py), so
`Model.add_domain`, `compile_prep` and both solvers accept it unchanged.
"""
import string

import numpy as np

_BAD_CHARS = set(string.punctuation.replace("_", "") + " ")


def _per_particle(value, n, dtype, name):
    a = np.asarray(value, dtype=dtype)
    if a.ndim == 0:
        return np.full(n, a, dtype=dtype)
    if a.shape != (n,):
        raise ValueError(f"{name} must be a scalar or have one entry per point ({n}), got shape {a.shape}")
    return np.ascontiguousarray(a)


def domain_from_arrays(points, type_id="UnAssigned", vol=1.0, mass=1.0, nu=0.0, fixed=False, rho=None, c=10.0,
                       xlim=None, ylim=None, zlim=None, rho0=1.0, c0=10, P0=None, gravity=None):
    """A `spatialpy.Domain` holding `points` ([N,3] or [N,2]; 2-D points get z = 0) with per-particle properties given as
    scalars or length-N arrays — the vectorised equivalent of
    `d = Domain(0, xlim, ylim, zlim, ...); for p in points: d.add_point(p, vol, mass, type_id, nu, fixed, rho, c)`.

    `type_id` entries are ints (> 0) or strings without punctuation/space other than `_` (domain.py:236-242).  Limits default
    to the bounding box of the points."""
    from spatialpy.core.domain import Domain
    from spatialpy.core.spatialpyerror import DomainError
    pts = np.asarray(points, dtype=float)
    if pts.ndim != 2 or pts.shape[1] not in (2, 3):
        raise ValueError(f"points must be [N,2] or [N,3], got shape {pts.shape}")
    n = pts.shape[0]
    if pts.shape[1] == 2:
        pts = np.concatenate([pts, np.zeros((n, 1))], axis=1)
    vol = _per_particle(vol, n, float, "vol")
    if (vol < 0).any():
        raise DomainError("Volume must be a positive value.")                       # domain.py:232-233
    mass = _per_particle(mass, n, float, "mass")
    # type ids: validate each distinct value once, then prefix (domain.py:235-242)
    tid = np.asarray(type_id, dtype=object)
    tid = np.full(n, type_id, dtype=object) if tid.ndim == 0 else tid
    if tid.shape != (n,):
        raise ValueError(f"type_id must be a scalar or have one entry per point ({n})")
    names = np.empty(n, dtype=object)
    uniq = {}
    for t in tid:                                # one dict lookup per particle; validation once per distinct id
        if t not in uniq:
            if isinstance(t, (int, np.integer)) and not isinstance(t, bool) and t <= 0:
                raise DomainError("Type_id must be a non-zero positive integer or a string.")
            if isinstance(t, str):
                for ch in t:
                    if ch in _BAD_CHARS:
                        raise DomainError(f"Type_id cannot contain '{ch}'")
            uniq[t] = f"type_{t}"
    if len(uniq) == 1:
        names[:] = next(iter(uniq.values()))
    else:
        names[:] = [uniq[t] for t in tid]
    lim = [(float(pts[:, k].min()), float(pts[:, k].max())) if n else (0.0, 0.0) for k in range(3)]
    dom = Domain(0, xlim if xlim is not None else lim[0], ylim if ylim is not None else lim[1],
                 zlim if zlim is not None else lim[2], rho0=rho0, c0=c0, P0=P0, gravity=gravity)
    dom.vertices = np.ascontiguousarray(pts)
    dom.vol = vol
    dom.mass = mass
    dom.type_id = names
    dom.nu = _per_particle(nu, n, float, "nu")
    dom.c = _per_particle(c, n, float, "c")
    dom.rho = mass / vol if rho is None else _per_particle(rho, n, float, "rho")  # domain.py:245-246
    dom.fixed = _per_particle(fixed, n, bool, "fixed")
    return dom


def _tetrahedron_volumes(vertices, tets):
    """|(a-d)·((b-d)x(c-d))| / 6 per tetrahedron (domain.py:446-454)."""
    a, b, c, d = (vertices[tets[:, k]] for k in range(4))
    return np.abs(np.einsum("ij,ij->i", a - d, np.cross(b - d, c - d)) / 6)


def read_xml_mesh(filename, subdomain_file=None, type_ids=None):
    """Array-at-a-time `Domain.read_xml_mesh` (domain.py:1148-1180; `XMLMeshLattice.apply`, lattice.py:595-680): a
    FEniCS/dolfin tetrahedral XML mesh -> Domain with one particle per vertex, `tetrahedrons`, per-vertex volume = a quarter
    of every adjacent tetrahedron (`calculate_vol`, domain.py:439-458, accumulated in the reference's order), mass = vol,
    rho = 1 (volumes agree with the reference's to one ulp: its `numpy.dot` of 3-vectors goes through BLAS).  The reference adds the vertices one `add_point` at a time (O(N²)) and loops over the tetrahedra in Python;
    this reader is linear, so refined meshes of 10⁶ vertices load in seconds.  `subdomain_file` lines are `index,type`; as in
    the reference, a vertex without an entry inherits the type of the last vertex that had one (lattice.py:644-645)."""
    import xml.etree.ElementTree as ET
    from spatialpy.core.spatialpyerror import DomainError, LatticeError
    root = ET.parse(filename).getroot()
    if root.tag != "dolfin":
        raise LatticeError(f"{filename} is not a FEniCS/dolfin xml mesh.")
    mesh = root[0]
    if mesh.tag != "mesh" or mesh.attrib["celltype"] != "tetrahedron" or mesh.attrib["dim"] != "3":
        raise LatticeError("XML mesh format error.")
    vertices, cells = mesh[0], mesh[1]
    n = len(vertices)
    idx = np.fromiter((int(v.attrib["index"]) for v in vertices), dtype=np.int64, count=n)
    pts = np.zeros((n, 3))
    for k, key in enumerate("xyz"):
        pts[idx, k] = np.fromiter((float(v.attrib[key]) for v in vertices), dtype=float, count=n)
    m = len(cells)
    cidx = np.fromiter((int(c.attrib["index"]) for c in cells), dtype=np.int64, count=m)
    tets = np.zeros((m, 4), dtype=int)
    for k in range(4):
        tets[cidx, k] = np.fromiter((int(c.attrib[f"v{k}"]) for c in cells), dtype=np.int64, count=m)
    type_id = "UnAssigned"
    if subdomain_file is not None:
        names = {}
        with open(subdomain_file, "r", encoding="utf-8") as f:
            for lnum, line in enumerate(f):
                try:
                    ndx, t = line.rstrip().split(",")
                    names[int(ndx)] = type_ids[t] if type_ids is not None else t
                except ValueError as err:
                    raise LatticeError(f"Could not read in subdomain file, error on line {lnum}: {line}") from err
        type_id = np.empty(n, dtype=object)
        last = "UnAssigned"
        for i in range(n):                      # sticky: kwargs['type_id'] survives to the next vertex (lattice.py:644-645)
            last = names.get(i, last)
            type_id[i] = last
    lims = [(pts[:, k].min(), pts[:, k].max()) for k in range(3)]      # apply_actions leaves the bounding box (get_bounding_box, domain.py:771-784)
    dom = domain_from_arrays(pts, type_id=type_id, vol=1.0, mass=1.0, nu=0.0, fixed=False, c=10.0, xlim=lims[0], ylim=lims[1],
                             zlim=lims[2])
    dom.tetrahedrons = tets
    dom.tetrahedron_vol = _tetrahedron_volumes(pts, tets)
    vol = np.zeros(n)
    np.add.at(vol, tets.reshape(-1), np.repeat(dom.tetrahedron_vol / 4, 4))     # same order as the reference's loop
    if not np.count_nonzero(vol):
        raise DomainError("Paritcles cannot have 0 volume")
    dom.vol = vol
    dom.mass = dom.vol
    with np.errstate(invalid="ignore", divide="ignore"):
        dom.rho = dom.mass / dom.vol
    return dom


def _subdomain_types(subdomain_file, type_ids, n):
    """Per-vertex type ids from a StochSS v1.x subdomain file (`index,type` lines) with the reference's sticky rule: a vertex
    without an entry inherits the type of the last vertex that had one (lattice.py:644-645, 776-777)."""
    from spatialpy.core.spatialpyerror import LatticeError
    names = {}
    with open(subdomain_file, "r", encoding="utf-8") as f:
        for lnum, line in enumerate(f):
            try:
                ndx, t = line.rstrip().split(",")
                names[int(ndx)] = type_ids[t] if type_ids is not None else t
            except ValueError as err:
                raise LatticeError(f"Could not read in subdomain file, error on line {lnum}: {line}") from err
    type_id = np.empty(n, dtype=object)
    last = "UnAssigned"
    for i in range(n):
        last = names.get(i, last)
        type_id[i] = last
    return type_id


def parse_msh(filename):
    """Minimal Gmsh ASCII reader (format 2.2 and 4.1) -> (points [N,3], [(cell type, connectivity)] in file order), cell types named
    as meshio names them ("vertex", "line", "triangle", "tetra").  Blocks follow meshio's grouping — 2.2: maximal runs of
    consecutive elements of one type; 4.1: one block per entity block — because `MeshIOLattice.apply` (lattice.py:783-803) keeps
    only the FIRST triangle block and the FIRST tetra block.  Node tags are mapped to 0-based positions in file order."""
    from spatialpy.core.spatialpyerror import LatticeError
    names = {15: ("vertex", 1), 1: ("line", 2), 2: ("triangle", 3), 4: ("tetra", 4)}
    with open(filename, "r", encoding="utf-8") as f:
        lines = f.read().split("\n")
    sections, k = {}, 0
    while k < len(lines):
        tag = lines[k].strip()
        if tag.startswith("$") and not tag.startswith("$End"):
            end = "$End" + tag[1:]
            j = k + 1
            while j < len(lines) and lines[j].strip() != end:
                j += 1
            sections[tag[1:]] = lines[k + 1:j]
            k = j
        k += 1
    try:
        version = float(sections["MeshFormat"][0].split()[0])
        if int(sections["MeshFormat"][0].split()[1]) != 0:
            raise LatticeError("binary .msh files are not supported; export the mesh as ASCII")
        nodes, elems = sections["Nodes"], sections["Elements"]
    except (KeyError, IndexError, ValueError) as err:
        raise LatticeError(f"{filename} is not a Gmsh .msh file") from err
    blocks = []
    if version < 4:
        n = int(nodes[0])
        body = np.array([ln.split() for ln in nodes[1:1 + n]], dtype=float)
        tags, pts = body[:, 0].astype(np.int64), np.ascontiguousarray(body[:, 1:4])
        cur_type, cur = None, []
        for ln in elems[1:1 + int(elems[0])]:
            t = ln.split()
            et, ntags = int(t[1]), int(t[2])
            if et != cur_type:
                if cur:
                    blocks.append((cur_type, cur))
                cur_type, cur = et, []
            cur.append(t[3 + ntags:])
        if cur:
            blocks.append((cur_type, cur))
    else:
        nblocks, n = int(nodes[0].split()[0]), int(nodes[0].split()[1])
        tags, pts, k = np.empty(n, np.int64), np.empty((n, 3)), 1
        filled = 0
        for _ in range(nblocks):
            cnt = int(nodes[k].split()[3])
            tags[filled:filled + cnt] = [int(v) for v in nodes[k + 1:k + 1 + cnt]]
            pts[filled:filled + cnt] = [[float(v) for v in ln.split()[:3]] for ln in nodes[k + 1 + cnt:k + 1 + 2 * cnt]]
            filled += cnt
            k += 1 + 2 * cnt
        k = 1
        for _ in range(int(elems[0].split()[0])):
            _, _, et, cnt = (int(v) for v in elems[k].split())
            blocks.append((et, [ln.split()[1:] for ln in elems[k + 1:k + 1 + cnt]]))
            k += 1 + cnt
    lookup = np.full(int(tags.max()) + 1, -1, dtype=np.int64)
    lookup[tags] = np.arange(len(tags))
    cells = []
    for et, rows in blocks:
        if et in names:
            cells.append((names[et][0], lookup[np.array(rows, dtype=np.int64)[:, :names[et][1]]]))
    return pts, cells


def _domain_from_mesh(pts, cells, subdomain_file, type_ids):
    from spatialpy.core.spatialpyerror import DomainError
    pts = np.ascontiguousarray(np.asarray(pts, dtype=float)[:, :3])
    n = len(pts)
    type_id = "UnAssigned" if subdomain_file is None else _subdomain_types(subdomain_file, type_ids, n)
    lims = [(pts[:, k].min(), pts[:, k].max()) for k in range(3)]
    dom = domain_from_arrays(pts, type_id=type_id, vol=1.0, mass=1.0, nu=0.0, fixed=False, c=10.0, xlim=lims[0], ylim=lims[1],
                             zlim=lims[2])
    tris = [c for name, c in cells if name == "triangle"]
    tets = [c for name, c in cells if name == "tetra"]
    if tris:
        dom.triangles = tris[0]
    vol = np.zeros(n)
    if tets:
        dom.tetrahedrons = tets[0]
        dom.tetrahedron_vol = _tetrahedron_volumes(pts, tets[0])
        np.add.at(vol, np.asarray(tets[0]).reshape(-1), np.repeat(dom.tetrahedron_vol / 4, 4))
    if not np.count_nonzero(vol):
        raise DomainError("Paritcles cannot have 0 volume")
    dom.vol = vol
    dom.mass = dom.vol
    with np.errstate(invalid="ignore", divide="ignore"):
        dom.rho = dom.mass / dom.vol
    return dom


def read_msh_file(filename, subdomain_file=None, type_ids=None):
    """Array-at-a-time `Domain.read_msh_file` (domain.py:1063-1096; `MeshIOLattice.apply`, lattice.py:732-806) for Gmsh ASCII
    meshes, without the `meshio` package the reference needs for this entry point: one particle per node in file order,
    `triangles` / `tetrahedrons` = the first block of each kind, per-vertex volume = a quarter of every adjacent tetrahedron
    (`calculate_vol`, domain.py:439-458), mass = vol, rho = 1.  Linear in the mesh size; the reference adds the nodes one
    `add_point` at a time (O(N²))."""
    pts, cells = parse_msh(filename)
    return _domain_from_mesh(pts, cells, subdomain_file, type_ids)


def import_meshio_object(mesh_obj, subdomain_file=None, type_ids=None):
    """Array-at-a-time `Domain.import_meshio_object` (domain.py:865-898): any object with meshio's `points` ([N,3]) and `cells`
    (blocks with `.type` / `.data`) -> Domain, same rules as `read_msh_file`."""
    return _domain_from_mesh(mesh_obj.points, [(c.type, np.asarray(c.data)) for c in mesh_obj.cells], subdomain_file, type_ids)


def read_stochss_domain(filename):
    """Array-at-a-time `Domain.read_stochss_domain` (domain.py:1098-1118; `StochSSLattice.apply`, lattice.py:845-903): a StochSS
    Domain (.domn) file, or the `domain` member of a StochSS Spatial Model (.smdl) file -> Domain.  Per particle: `point`, `type`
    (looked up in `types[].typeID -> name`, '-' removed), `volume`, `mass`, `nu`, `fixed`, optional `rho` (default mass/volume)
    and `c` (default 0); domain-wide `rho_0`, `c_0`, `p_0`, `gravity`; limits = the bounding box of the particles."""
    import json
    from spatialpy.core.spatialpyerror import LatticeError
    try:
        with open(filename, "r", encoding="utf-8") as f:
            s = json.load(f)
        if "domain" in s:
            s = s["domain"]
        names = {t["typeID"]: t["name"].replace("-", "") for t in s["types"]}
        parts = s["particles"]
        n = len(parts)
        pts = np.array([p["point"][:3] for p in parts], dtype=float).reshape(n, 3)
        mass = np.array([p["mass"] for p in parts], dtype=float)
        vol = np.array([p["volume"] for p in parts], dtype=float)
        rho = np.array([mass[i] / vol[i] if p.get("rho") is None else p["rho"] for i, p in enumerate(parts)], dtype=float)
        tid = np.empty(n, dtype=object)
        tid[:] = [names[p["type"]] for p in parts]
        lims = [(pts[:, k].min(), pts[:, k].max()) if n else (0.0, 0.0) for k in range(3)]
        dom = domain_from_arrays(pts, type_id=tid if n else "UnAssigned", vol=vol, mass=mass, nu=np.array([p["nu"] for p in parts], dtype=float),
                                 fixed=np.array([p["fixed"] for p in parts], dtype=bool), rho=rho,
                                 c=np.array([p.get("c", 0) for p in parts], dtype=float), xlim=lims[0], ylim=lims[1], zlim=lims[2],
                                 rho0=s["rho_0"], c0=s["c_0"], P0=s["p_0"], gravity=s["gravity"])
    except KeyError as err:
        raise LatticeError("The file is not a StochSS Domain (.domn) or a StochSS Spatial Model (.smdl).") from err
    return dom
