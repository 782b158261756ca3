"""CPU tier: the native slab transport kernels (spatialpy_b200/csrc/ssb_core.cu: k_halo_send/_recv/_wait, k_inbox_send/_recv,
k_board_post/_reduce) run as SOURCE on the host by the block emulator (tests/cuda_emu): message layout in the receive window,
last-CTA publish of the sequence flags, compaction of the sSSA inbox, the scalar all-reduce boards; further down, k_lookahead
against the oracle's next predictor and a property test of the Verlet keep rule it feeds.  The reference has no domain
decomposition; what these kernels must preserve is that a ghost copy carries its owner's values after every
exchange at the substep boundaries of E/src/simulate_threads.cpp:232-281 and that no molecule is lost or duplicated."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P64 = ctypes.POINTER(ctypes.c_double)


class _Rank(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("Sc", ctypes.c_int), ("Sd", ctypes.c_int), ("F", ctypes.c_void_p * 3), ("Fbp", ctypes.c_void_p * 3),
                ("Frho", ctypes.c_void_p), ("Q", ctypes.c_void_p), ("rho_new", ctypes.c_void_p), ("v", ctypes.c_void_p * 3),
                ("bvf", ctypes.c_void_p), ("inbox", ctypes.c_void_p * 2), ("inbox_src", ctypes.c_void_p * 2),
                ("blk_mail", ctypes.c_void_p * 2), ("slot_of_id", ctypes.c_void_p)]


class _Side(ctypes.Structure):
    _fields_ = [("ids", ctypes.c_void_p), ("n", ctypes.c_int), ("buf", ctypes.c_void_p), ("flag", ctypes.c_void_p), ("count", ctypes.c_void_p)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    from spatialpy_b200 import codegen
    tmp = tmp_path_factory.mktemp("halo_emu")
    src = open(os.path.join(codegen.CSRC, "ssb_core.cu")).read()
    a = src.index("__device__ __forceinline__ int halo_width(")
    b = src.index("// =====", src.index("__global__ void k_board_reduce"))
    text = src[a:b]
    # the two inline-PTX helpers have host versions in the harness; everything else is the shipped text
    text, n1 = re.subn(r"__device__ __forceinline__ unsigned long long ssb_globaltimer\(\) \{.*?\n", "", text)
    text, n2 = re.subn(r"__device__ __forceinline__ unsigned long long ssb_ld_flag\(const unsigned long long \*p\) \{.*?\n\}\n", "", text, flags=re.S)
    assert n1 == 1 and n2 == 1
    kern = tmp / "halo_kernels.inc"
    kern.write_text(text)
    so = tmp / "halo_emu.so"
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(codegen.nvcc_path())), "include")
    cmd = ["g++", "-std=c++20", "-O1", "-w", "-pthread", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I", cuda_inc, "-I", codegen.CSRC,
           "-I", os.path.join(ROOT, "tests", "cuda_emu"), f"-DEMU_KERNELS=\"{kern}\"",
           os.path.join(ROOT, "tests", "cuda_emu", "halo_emu.cpp"), "-o", str(so)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return ctypes.CDLL(str(so))


class Rank:
    """One slab rank's state in host memory, in a scrambled storage order (slot_of_id)."""
    def __init__(self, n, Sc, Sd, seed):
        rng = np.random.default_rng(seed)
        self.n, self.Sc, self.Sd = n, Sc, Sd
        self.slot = rng.permutation(n).astype(np.int32)
        self.f = {k: rng.normal(size=(3, n)) for k in ("F", "Fbp", "v")}
        self.f.update({k: rng.normal(size=n) for k in ("Frho", "rho_new", "bvf")})
        self.f["Q"] = rng.normal(size=(max(Sc, 1), n))
        self.inbox = [np.zeros((max(Sd, 1), n), np.uint32) for _ in range(2)]
        self.inbox_src = [np.zeros(n, np.uint64) for _ in range(2)]
        self.blk_mail = [np.zeros((n + 127) // 128, np.int32) for _ in range(2)]
        r = _Rank()
        r.N, r.Sc, r.Sd = n, Sc, Sd
        for d in range(3):
            r.F[d], r.Fbp[d], r.v[d] = self.f["F"][d].ctypes.data, self.f["Fbp"][d].ctypes.data, self.f["v"][d].ctypes.data
        r.Frho, r.Q, r.rho_new, r.bvf = (self.f[k].ctypes.data for k in ("Frho", "Q", "rho_new", "bvf"))
        for b in range(2):
            r.inbox[b], r.inbox_src[b], r.blk_mail[b] = self.inbox[b].ctypes.data, self.inbox_src[b].ctypes.data, self.blk_mail[b].ctypes.data
        r.slot_of_id = self.slot.ctypes.data
        self.c = r

    def rows(self, group, ids):
        """The [W, len(ids)] message the transport must carry for `group` (halo_width: 7 + Sc | 1 | 4)."""
        p = self.slot[ids]
        if group == 0:
            return np.concatenate([self.f["F"][:, p], self.f["Fbp"][:, p], self.f["Frho"][None, p], self.f["Q"][:self.Sc, p]])
        if group == 1:
            return self.f["rho_new"][None, p]
        return np.concatenate([self.f["v"][:, p], self.f["bvf"][None, p]])


class Window:
    def __init__(self, words):
        self.buf = np.zeros(words, np.float64)
        self.flag = np.zeros(1, np.uint64)
        self.count = np.zeros(1, np.uint64)


def _sides(specs):
    arr = (_Side * 2)()
    keep = []
    for k, sp in enumerate(specs):
        if sp is None:
            arr[k].n = 0
            continue
        ids, win = sp
        ids = np.ascontiguousarray(ids, np.int32)
        keep.append(ids)
        arr[k].ids, arr[k].n, arr[k].buf = ids.ctypes.data, len(ids), win.buf.ctypes.data
        arr[k].flag, arr[k].count = win.flag.ctypes.data, win.count.ctypes.data
    return arr, keep


@pytest.mark.parametrize("group", [0, 1, 2])
@pytest.mark.parametrize("two_sided", [True, False])
def test_halo_messages_carry_the_owner_rows_into_the_ghost_copies(emu, group, two_sided):
    """k_halo_send writes the column-major message [field][row] into the neighbour's window and its LAST CTA raises the sequence
    flag of every connected side (and rewinds the CTA counter for the next message); k_halo_wait passes once the flag is there;
    k_halo_recv moves the rows into the receiver's ghost slots and touches nothing else.  Two CTAs (220 rows) and the one-sided
    end-of-chain case (side without a peer)."""
    Sc = 2
    A, B, Cc = Rank(700, Sc, 1, 1), Rank(500, Sc, 1, 2), Rank(400, Sc, 1, 3)
    rng = np.random.default_rng(10 + group)
    n0, n1 = 150, 70
    W = emu.emu_halo_width(group, Sc)
    assert W == {0: 7 + Sc, 1: 1, 2: 4}[group]
    a_left, a_right = rng.choice(A.n, n0, replace=False), rng.choice(A.n, n1, replace=False)      # A's owned rows next to each face
    b_ghost, c_ghost = rng.choice(B.n, n0, replace=False), rng.choice(Cc.n, n1, replace=False)    # the neighbours' copies, same order
    wb, wc = Window(W * n0 + 8), Window(W * n1 + 8)
    done = np.zeros(1, np.uint32)
    send, keep = _sides([(a_left, wb), (a_right, wc) if two_sided else None])
    seq = 5
    assert emu.emu_halo_send(ctypes.byref(A.c), group, send, ctypes.c_ulonglong(seq), done.ctypes.data_as(ctypes.c_void_p)) == 0
    np.testing.assert_array_equal(wb.buf[:W * n0].reshape(W, n0), A.rows(group, a_left))
    assert wb.buf[W * n0:].sum() == 0 and int(wb.flag[0]) == seq and int(done[0]) == 0
    if two_sided:
        np.testing.assert_array_equal(wc.buf[:W * n1].reshape(W, n1), A.rows(group, a_right))
        assert int(wc.flag[0]) == seq
    else:
        assert int(wc.flag[0]) == 0 and not wc.buf.any()
    err = np.zeros(4, np.int32)
    emu.emu_halo_wait(wb.flag.ctypes.data_as(ctypes.c_void_p), wb.flag.ctypes.data_as(ctypes.c_void_p), ctypes.c_ulonglong(seq),
                      err.ctypes.data_as(ctypes.c_void_p))
    assert err[0] == 0
    before = {k: v.copy() for k, v in B.f.items()}
    recv, keep2 = _sides([(b_ghost, wb), None])
    assert emu.emu_halo_recv(ctypes.byref(B.c), group, recv) == 0
    np.testing.assert_array_equal(B.rows(group, b_ghost), A.rows(group, a_left))
    untouched = np.ones(B.n, bool)
    untouched[B.slot[b_ghost]] = False
    for k, v in B.f.items():
        np.testing.assert_array_equal(v[..., untouched], before[k][..., untouched])
    groups_fields = {0: ("F", "Fbp", "Frho", "Q"), 1: ("rho_new",), 2: ("v", "bvf")}[group]
    for k, v in B.f.items():
        if k not in groups_fields:
            np.testing.assert_array_equal(v, before[k])


def test_inbox_entries_are_compacted_and_no_molecule_is_lost(emu):
    """k_inbox_send reads-and-clears the ghost voxels' arrivals and appends one (row, species, count) entry per non-zero pair to
    the owner's window, publishes the entry count with the flag and rewinds its counters; k_inbox_recv adds the entries to the
    owner's voxels and raises the chunk mail flags.  Molecules are conserved; nothing but the listed rows changes."""
    Sd = 3
    A, B = Rank(900, 1, Sd, 4), Rank(600, 1, Sd, 5)
    rng = np.random.default_rng(21)
    n = 260                                                   # three CTAs
    a_ghost, b_owned = rng.choice(A.n, n, replace=False), rng.choice(B.n, n, replace=False)
    buf = 1
    arrivals = (rng.random((Sd, n)) < 0.07) * rng.integers(1, 5, size=(Sd, n))
    A.inbox[buf][:, A.slot[a_ghost]] = arrivals
    A.inbox_src[buf][A.slot[a_ghost]] = 77
    elsewhere = rng.integers(0, 3, size=(Sd, A.n)).astype(np.uint32)      # mail for A's own voxels must stay
    mask = np.ones(A.n, bool)
    mask[A.slot[a_ghost]] = False
    A.inbox[buf][:, mask] = elsewhere[:, mask]
    B.inbox[buf][:] = rng.integers(0, 2, size=(Sd, B.n))
    b_before = B.inbox[buf].copy()
    win = Window(4 * n * Sd + 8)                             # uint4 entries = 2 doubles each
    done, icount = np.zeros(1, np.uint32), np.zeros(2, np.uint32)
    send, keep = _sides([(a_ghost, win), None])
    seq = 9
    emu.emu_inbox_send(ctypes.byref(A.c), buf, send, ctypes.c_ulonglong(seq), done.ctypes.data_as(ctypes.c_void_p),
                       icount.ctypes.data_as(ctypes.c_void_p))
    nent = int(np.count_nonzero(arrivals))
    assert nent > 10 and int(win.count[0]) == nent and int(win.flag[0]) == seq and int(done[0]) == 0 and not icount.any()
    assert not A.inbox[buf][:, A.slot[a_ghost]].any() and not A.inbox_src[buf][A.slot[a_ghost]].any()
    np.testing.assert_array_equal(A.inbox[buf][:, mask], elsewhere[:, mask])
    ent = win.buf.view(np.uint32)[:4 * nent].reshape(nent, 4)
    got = np.zeros((Sd, n), np.int64)
    np.add.at(got, (ent[:, 1], ent[:, 0]), ent[:, 2])
    np.testing.assert_array_equal(got, arrivals)              # every non-zero pair exactly once, whatever the order
    recv, keep2 = _sides([(b_owned, win), None])
    emu.emu_inbox_recv(ctypes.byref(B.c), buf, recv, 128)
    want = b_before.astype(np.int64)
    want[:, B.slot[b_owned]] += arrivals
    np.testing.assert_array_equal(B.inbox[buf], want)
    mail = np.zeros_like(B.blk_mail[buf])
    mail[np.unique(B.slot[b_owned][arrivals.any(axis=0)] // 128)] = 1
    np.testing.assert_array_equal(B.blk_mail[buf], mail)
    assert int(B.inbox[buf].sum()) == int(b_before.sum()) + int(arrivals.sum())


@pytest.mark.parametrize("world", [2, 3, 8])
def test_board_all_reduce_gives_every_rank_the_same_extreme(emu, world):
    """k_board_post / k_board_reduce: every rank stores (value, sequence) into its slot of EVERY rank's board; each rank reduces
    its own board.  Maximum (Ddiag, displacement) and minimum (earliest pending event) of bit patterns of non-negative doubles;
    consecutive sequence numbers use the two parities of a channel and do not disturb each other."""
    words = emu.emu_board_words()
    boards = [np.zeros(words, np.uint64) for _ in range(world)]
    ptrs = (ctypes.c_void_p * world)(*[b.ctypes.data for b in boards])
    rng = np.random.default_rng(world)
    err = np.zeros(4, np.int32)
    for seq, ch, take_min in ((1, 0, 0), (2, 0, 0), (1, 1, 1), (1, 2, 0), (3, 0, 0)):
        vals = np.abs(rng.normal(size=world)) * 10.0 ** rng.integers(-3, 4, size=world)
        bits = vals.view(np.uint64).copy()
        out = np.zeros(world, np.uint64)
        emu.emu_board_allreduce(ptrs, world, ch, ctypes.c_ulonglong(seq), take_min, bits.ctypes.data_as(ctypes.c_void_p),
                                out.ctypes.data_as(ctypes.c_void_p), err.ctypes.data_as(ctypes.c_void_p))
        want = vals.min() if take_min else vals.max()
        assert (out.view(np.float64) == want).all() and err[0] == 0, (seq, ch)


# ----------------------------------------------------------------------------------------------------------------------
# k_lookahead: the displacement bookkeeping that lets a moving step decide "keep the candidate lists or rebuild" without a
# read-back after the predictor
# ----------------------------------------------------------------------------------------------------------------------
class _LookArgs(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("dt", ctypes.c_double), ("x", ctypes.c_void_p * 3), ("xref", ctypes.c_void_p * 3),
                ("v", ctypes.c_void_p * 3), ("F", ctypes.c_void_p * 3), ("Fbp", ctypes.c_void_p * 3), ("solid", ctypes.c_void_p),
                ("out", ctypes.c_void_p)]


@pytest.fixture(scope="module")
def look_emu(tmp_path_factory):
    from spatialpy_b200 import codegen
    tmp = tmp_path_factory.mktemp("look_emu")
    src = open(os.path.join(codegen.CSRC, "ssb_core.cu")).read()
    a = src.index("__global__ void k_lookahead(")
    b = src.index("// neighbour search: query = live x_i")
    kern = tmp / "look_kernels.inc"
    kern.write_text(src[a:b])
    so = tmp / "lookahead_emu.so"
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(codegen.nvcc_path())), "include")
    cmd = ["g++", "-std=c++20", "-O1", "-w", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I", cuda_inc,
           "-I", codegen.CSRC, "-I", os.path.join(ROOT, "tests", "cuda_emu"), f"-DEMU_KERNELS=\"{kern}\"",
           os.path.join(ROOT, "tests", "cuda_emu", "lookahead_emu.cpp"), "-o", str(so)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return ctypes.CDLL(str(so))


@pytest.mark.parametrize("which", ["tank", "cavity2d"])
def test_lookahead_predicts_what_the_next_predictor_does(look_emu, which):
    """After a step is complete, k_lookahead's three maxima — |x' - xref|^2, |x' - x|^2, |x - xref|^2 with x' = the position the
    NEXT predictor will produce — must be exactly what the oracle's take_step1 (E/src/simulate.cpp:68-79; the boundary conditions
    reassign v only after the position update) then does, on a 3-D tank and on the lid-driven cavity whose BC overwrites v."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sdpd_oracle
    from spatialpy_b200 import configs
    from util import load_model
    if which == "tank":
        fm = configs.tank_sdpd(n=9, nt=10, output_every=10, dt=2e-5)
        fm.x = fm.x + np.random.default_rng(3).uniform(-0.01, 0.01, fm.x.shape)
    else:
        fm = load_model("cavity2d")
    o = sdpd_oracle.SdpdOracle(fm)
    o.step()
    xref = o.x.copy()                                        # "the lists were built here"
    for _ in range(3):
        o.step()
    st = {k: np.ascontiguousarray(getattr(o, k).T.copy()) for k in ("x", "v", "F", "Fbp")}     # [3, N] component arrays
    xr = np.ascontiguousarray(xref.T.copy())
    solid = np.ascontiguousarray(o.solid.astype(np.int32))
    out = np.zeros(3, np.uint64)
    a = _LookArgs()
    a.N, a.dt = o.N, o.dt
    for d in range(3):
        a.x[d], a.xref[d], a.v[d], a.F[d], a.Fbp[d] = (arr[d].ctypes.data for arr in (st["x"], xr, st["v"], st["F"], st["Fbp"]))
    a.solid, a.out = solid.ctypes.data, out.ctypes.data
    assert look_emu.emu_lookahead(ctypes.byref(a), 3) == 0   # 3 CTAs of 256 threads, grid-stride
    x_now = o.x.copy()
    o.take_step1()
    want = [((o.x - xref) ** 2).sum(axis=1).max(), ((o.x - x_now) ** 2).sum(axis=1).max(), ((x_now - xref) ** 2).sum(axis=1).max()]
    got = out.view(np.float64)
    assert want[1] > 0 and want[0] > want[1]
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)


def _verlet_run(x, moves, h, skin, charge_build_step):
    """Host restatement of the moving-domain list bookkeeping (ssb_core.cu: mv_decide_keep / mv_pre): lists of radius h(1+skin)
    are built around the PREDICTED queries against the start-of-step snapshot; a later step keeps them while
    D(x') + D(x0) [+ the build step's own predictor displacement] <= skin*h.  Returns (missed pairs, rebuilds)."""
    from scipy.spatial import cKDTree
    x = x.copy()
    xref = lists = None
    d_build = 0.0
    missed = rebuilds = 0
    for mv in moves:
        q = x + mv                                           # what the predictor will produce
        keep = False
        if lists is not None:
            d_next = np.sqrt(((q - xref) ** 2).sum(axis=1).max())
            d_cur = np.sqrt(((x - xref) ** 2).sum(axis=1).max())
            keep = d_next + d_cur + (d_build if charge_build_step else 0.0) <= skin * h
        if not keep:
            xref = x.copy()
            d_build = np.sqrt((mv ** 2).sum(axis=1).max())
            lists = [set(l) for l in cKDTree(x).query_ball_point(q, h * (1 + skin))]
            rebuilds += 1
        need = cKDTree(x).query_ball_point(q, h)             # the reference's rule: live query against this step's snapshot
        missed += sum(len(set(l) - lists[i]) for i, l in enumerate(need))
        x = q                                                # (the corrector does not move particles)
    return missed, rebuilds


def test_verlet_keep_rule_never_loses_a_pair_and_needs_the_build_step_term():
    """Property behind the skin logic: with the build step's predictor displacement charged, no pair the reference would find is
    ever missing from a kept candidate list — random walks with reversals, 40 steps, several seeds; and a two-particle
    counter-example shows the term is needed (a particle that steps forward in the build step and back afterwards)."""
    h, skin = 1.0, 0.05
    rng = np.random.default_rng(0)
    total_rebuilds = 0
    for seed in range(4):
        rng = np.random.default_rng(seed)
        x = rng.uniform(0, 6, size=(400, 3))
        moves = [rng.normal(size=x.shape) * 0.004 * (1 + 3 * (k % 7 == 0)) * (-1) ** (k // 3) for k in range(40)]
        missed, rebuilds = _verlet_run(x, moves, h, skin, True)
        assert missed == 0
        total_rebuilds += rebuilds
    assert 4 < total_rebuilds < 4 * 40                       # lists are reused, and rebuilt when the budget runs out
    # counter-example (eps = skin*h): particle 1 starts at h + 0.45 eps from particle 0 and steps 0.6 eps AWAY in the build step, so
    # its predicted query lies outside the candidate radius and its list lacks particle 0; it then comes back by 0.85 eps and by
    # 0.22 eps.  Measured from xref the two displacement terms never exceed 0.85 eps, so the rule without the build-step term keeps
    # the lists — and at the third step the pair is within h but not on the list
    eps = skin * h
    x = np.array([[0.0, 0.0, 0.0], [h + 0.45 * eps, 0.0, 0.0]])
    away = np.array([[0.0, 0.0, 0.0], [0.6 * eps, 0.0, 0.0]])
    moves = [away, -away - np.array([[0.0, 0, 0], [0.25 * eps, 0, 0]]), np.array([[0.0, 0, 0], [-0.22 * eps, 0, 0]])]
    missed_old, _ = _verlet_run(x, moves, h, skin, False)
    missed_new, rebuilds_new = _verlet_run(x, moves, h, skin, True)
    assert missed_old > 0 and missed_new == 0 and rebuilds_new >= 2
