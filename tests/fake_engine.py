"""A numpy stand-in for `spatialpy_b200.engine.Engine` with the slab-phase interface (CPU tier only).

It is NOT the engine and computes no SDPD: it is a toy particle system whose every quantity depends on the neighbours within
h, so that the exchange protocol of `SlabEngine` (ghost synchronisation after each sweep, inbox traffic of the windowed
reaction-diffusion step, re-partition hand-over) can be exercised over gloo without a GPU.  All neighbour sums run in ascending
GLOBAL id order, so a slab-decomposed run must reproduce the single-rank run bit for bit — any ghost that is missing, stale or
delivered twice shows up as a difference."""
import ctypes

import numpy as np
from scipy.spatial import cKDTree

PH_PRE, PH_CORRECTOR, PH_FINISH, PH_RDME_PREP, PH_RDME_INIT, PH_RDME_WINDOW, PH_RDME_CLOSE, PH_END, PH_RDME_MIN, PH_RDME_EXTRA = range(10)


def _view(ptr, count, ctype, dtype):
    if count == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array((ctype * count).from_address(ptr)).view(dtype)


class FakeEngine:
    def __init__(self, fm, device=0, flags=0, rdme_epsilon=0.0, owned=None, rng_id=None, **_):
        self.fm = fm.finalize()
        self.N = fm.num_particles
        self.Sc, self.Sd = fm.num_chem_species, fm.num_stoch_species
        self.owned = np.ones(self.N, bool) if owned is None else np.asarray(owned).astype(bool)
        self.gid = np.arange(self.N) if rng_id is None else np.asarray(rng_id).astype(np.int64)
        self.closed = False
        self.reset(0)

    # ------------------------------------------------------------------ lifecycle / taps
    def reset(self, seed):
        fm, N = self.fm, self.N
        self.x = fm.x.copy()
        self.v, self.F, self.Fbp = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros((N, 3))
        self.rho, self.nu, self.mass = fm.rho.copy(), fm.nu.copy(), fm.mass.copy()
        self.Frho, self.bvf, self.rho_new = np.zeros(N), np.zeros(N), np.zeros(N)
        self.C = fm.u0[:, :self.Sc].astype(np.float64)
        self.Q = np.zeros((N, self.Sc))
        self.xx = fm.u0[:, :self.Sd].astype(np.uint32).copy()
        self.inbox = np.zeros((N, self.Sd), np.uint32)
        self.step_no, self.epoch = 0, 0
        self.disp = 0.0
        self.n_jumps = 0
        self.nbr = None

    def close(self):
        self.closed = True

    def get(self, name):
        return {"x": self.x, "v": self.v, "F": self.F, "Fbp": self.Fbp, "rho": self.rho, "Frho": self.Frho, "nu": self.nu,
                "mass": self.mass, "bvf_phi": self.bvf, "type": self.fm.type, "C": self.C, "Q": self.Q, "xx": self.xx}[name].copy()

    def set(self, name, values):
        dst = {"v": self.v, "F": self.F, "Fbp": self.Fbp, "Frho": self.Frho, "bvf_phi": self.bvf, "C": self.C, "Q": self.Q}[name]
        dst[...] = np.asarray(values).reshape(dst.shape)

    def get_step(self):
        return self.step_no, self.epoch

    def set_step(self, step, epoch):
        self.step_no, self.epoch = int(step), int(epoch)

    def counters(self):
        return {"reactions": 0, "diffusions": self.n_jumps, "seconds": 0.0, "windows": 0}

    def skin_stats(self):
        return {"skin": 0.1, "step_disp_max": self.disp, "rebuilds": 0}

    def halo_width(self, group):
        return 7 + self.Sc if group == 0 else (4 if group == 2 else 1)

    # ------------------------------------------------------------------ toy physics
    def _neighbours(self):
        """Per particle: local indices of the others within h, in ascending global id order."""
        tree = cKDTree(self.x)
        out = []
        for i, lst in enumerate(tree.query_ball_point(self.x, self.fm.h)):
            js = np.array([j for j in lst if j != i], dtype=np.int64)
            out.append(js[np.argsort(self.gid[js], kind="stable")])
        return out

    def phase(self, phase, arg=0.0):
        dt, g = self.fm.dt, np.asarray(self.fm.gravity)
        mobile = (self.fm.solid == 0)
        if phase == PH_PRE:
            self.v[mobile] += 0.5 * dt * self.F[mobile]
            move = dt * self.v
            move[~mobile] = 0.0
            self.x += move
            self.disp = max(self.disp, float(np.sqrt((move ** 2).sum(axis=1)).max()))
            self.nbr = self._neighbours()
            for i in np.nonzero(self.owned)[0]:
                f, q, fr = g.copy(), np.zeros(self.Sc), 0.0
                for j in self.nbr[i]:
                    d = self.x[i] - self.x[j]
                    w = self.mass[j] / (d @ d + 0.01 * self.fm.h ** 2)
                    f = f + w * d + 0.1 * (self.v[j] - self.v[i])
                    fr = fr + w * (self.rho[j] - self.rho[i])
                    q = q + (self.C[j] - self.C[i]) * w
                self.F[i], self.Frho[i], self.Q[i] = f, fr, q
                self.Fbp[i] = 0.5 * f
        elif phase == PH_CORRECTOR:
            self.v[mobile] += 0.5 * dt * self.F[mobile]
            self.rho_new = self.rho + dt * self.Frho
            # like the Shepard filter, the owner adds a neighbour sweep that its ghost copies do not run: only exchange
            # group 1 gives the ghosts the right value
            for i in np.nonzero(self.owned)[0]:
                if len(self.nbr[i]):
                    self.rho_new[i] += 1.0e-3 * sum(self.rho[j] for j in self.nbr[i]) / len(self.nbr[i])
        elif phase == PH_FINISH:
            for i in np.nonzero(self.owned)[0]:
                # like the engine's BVF sweep this reads the neighbours' NEW density (exchange group 1) ...
                s = sum(self.rho_new[j] * self.fm.solid[j] for j in self.nbr[i])
                self.bvf[i] = s / (1.0 + len(self.nbr[i]))
                # ... and, like the bounce-back, changes the owner's velocity in a way its ghosts cannot (exchange group 2)
                if mobile[i]:
                    for j in self.nbr[i]:
                        if self.fm.solid[j]:
                            self.v[i] = self.v[i] + 1.0e-4 * (self.x[i] - self.x[j]) / dt * self.rho_new[j]
            self.C += dt * self.Q
            self.rho = self.rho_new.copy()
        elif phase == PH_RDME_PREP:
            return 1.0 + 0.001 * float(self.gid[self.owned].max())     # rank dependent on purpose: the caller must all-reduce
        elif phase == PH_RDME_INIT:
            self.mx = arg
            self.inbox[...] = 0
            self.epoch += 1
            return 2.0
        elif phase in (PH_RDME_WINDOW, PH_RDME_CLOSE):
            own = np.nonzero(self.owned)[0]
            self.xx[own] += self.inbox[own]                 # mail for ghost voxels stays put for the halo exchange
            self.inbox[own] = 0
            self.epoch += 1
            if phase == PH_RDME_WINDOW:
                w = int(arg)
                for i in own:
                    if len(self.nbr[i]) == 0:
                        continue
                    for s in range(self.Sd):
                        if self.xx[i, s] > 0 and (self.gid[i] + 2 * self.step_no + w + s) % 3 == 0:
                            j = self.nbr[i][(self.gid[i] + self.step_no + w) % len(self.nbr[i])]
                            self.xx[i, s] -= 1
                            self.inbox[j, s] += 1
                            self.n_jumps += 1
        elif phase == PH_RDME_MIN:
            return float("inf")
        elif phase == PH_RDME_EXTRA:
            self.epoch += 2
        elif phase == PH_END:
            self.step_no += 1
        return 0.0

    # ------------------------------------------------------------------ halo primitives (raw pointers, like the C-ABI)
    def _cols(self, group):
        if group == 0:
            return [self.F, self.Fbp, self.Frho[:, None], self.Q]
        if group == 1:
            return [self.rho_new[:, None]]
        if group == 2:
            return [self.v, self.bvf[:, None]]
        return [self.rho[:, None]]

    def halo_pack(self, group, ids_ptr, n, out_ptr):
        ids = _view(ids_ptr, n, ctypes.c_int32, np.int32)
        w = self.halo_width(group)
        out = _view(out_ptr, n * w, ctypes.c_double, np.float64).reshape(n, w)
        if n:
            out[...] = np.concatenate([c[ids] for c in self._cols(group)], axis=1)

    def halo_unpack(self, group, ids_ptr, n, in_ptr):
        ids = _view(ids_ptr, n, ctypes.c_int32, np.int32)
        w = self.halo_width(group)
        data = _view(in_ptr, n * w, ctypes.c_double, np.float64).reshape(n, w)
        k = 0
        for c in self._cols(group):
            if n:
                c[ids] = data[:, k:k + c.shape[1]]
            k += c.shape[1]
        if group == 1:
            pass            # rho_new is a plain array here (self._cols returned a view of it)

    def inbox_pack(self, ids_ptr, n, out_ptr):
        ids = _view(ids_ptr, n, ctypes.c_int32, np.int32)
        out = _view(out_ptr, n * max(self.Sd, 1), ctypes.c_int32, np.int32).reshape(n, max(self.Sd, 1))
        if n and self.Sd:
            out[:, :self.Sd] = self.inbox[ids].astype(np.int32)
            self.inbox[ids] = 0

    def inbox_add(self, ids_ptr, n, in_ptr):
        ids = _view(ids_ptr, n, ctypes.c_int32, np.int32)
        data = _view(in_ptr, n * max(self.Sd, 1), ctypes.c_int32, np.int32).reshape(n, max(self.Sd, 1))
        if n and self.Sd:
            self.inbox[ids] += data[:, :self.Sd].astype(np.uint32)
