"""TEST INFRASTRUCTURE — CPU restatement (numpy, fp64) of the reference's SDPD path.

This is the ORACLE of the parity tests, not product code: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import it; the product path (spatialpy_b200/) never does and has no CPU fallback.

It restates, function by function, the serial (`-t 1`) semantics of
    E/src/particle.cpp       find_neighbors / add_to_neighbor_list            (:150-210, :240-294)
    E/src/simulate.cpp       take_step1 / compute_forces / take_step2          (:56-173)
    E/src/model.cpp          pairwiseForce / filterDensity / computeBoundaryVolumeFraction / applyBoundaryVolumeFraction
    E/src/simulate_rdme.cpp  nsm_core__initialize_diff_propensities            (:131-152)
with E = /root/reference/spatialpy/solvers/c_base/ssa_sdpd-c-simulation-engine.
PINNED: tests/test_cpu_oracle.py checks every function here against the full-precision taps of the unmodified
reference engine (tests/golden/*.ref.npz, produced by oracle/_ref built from /root/reference).
Neighbour search is brute force O(N^2) with ANN's inclusion arithmetic; use it for N up to a few thousand.
"""
import re

import numpy as np


def alpha(dim, h):
    """particle.cpp:169-175 — note `5 / 4 * h` is integer division in C: alpha_1D = h."""
    if dim == 3:
        return 105 / (16 * np.pi * h * h * h)
    if dim == 2:
        return 5 / (np.pi * h * h)
    return (5 // 4) * h


def parse_bc(bc_source, type_constants):
    """Interpret the reference's generated BC text (boundarycondition.py:121-167):
    if((me->x[0] >= a)&&(me->type == type_T)){me->v[0]=..;me->rho=..;me->C[k] = ..;} ..."""
    rules = []
    for cond, body in re.findall(r"if\((.*?)\)\{(.*?)\}", bc_source or ""):
        conds = []
        for c in cond.split("&&"):
            m = re.match(r"\(me->x\[(\d)\] (>=|<=) (.*)\)", c.strip())
            if m:
                conds.append(("x", int(m.group(1)), m.group(2), float(m.group(3))))
                continue
            m = re.match(r"\(me->type == (.*)\)", c.strip())
            if m:
                t = m.group(1)
                conds.append(("type", type_constants.get(t, None) if not t.isdigit() else int(t)))
                continue
            raise ValueError(f"unsupported BC condition {c!r}")
        assigns = []
        for a in body.split(";"):
            a = a.strip()
            if not a:
                continue
            m = re.match(r"me->(v|C)\[(\d+)\]\s*=\s*(.*)", a)
            if m:
                assigns.append((m.group(1), int(m.group(2)), float(m.group(3))))
                continue
            m = re.match(r"me->(nu|rho)\s*=\s*(.*)", a)
            if m:
                assigns.append((m.group(1), None, float(m.group(2))))
                continue
            raise ValueError(f"unsupported BC assignment {a!r}")
        rules.append((conds, assigns))
    return rules


class SdpdOracle:
    def __init__(self, fm):
        self.fm = fm
        N = fm.num_particles
        self.N, self.dim, self.h, self.dt = N, fm.dimension, float(fm.h), float(fm.dt)
        self.static = bool(fm.static_domain)
        self.x = fm.x.copy()
        self.v = np.zeros((N, 3)); self.vt = np.zeros((N, 3))
        self.F = np.zeros((N, 3)); self.Fbp = np.zeros((N, 3)); self.Frho = np.zeros(N)
        self.rho = fm.rho.copy(); self.old_rho = np.zeros(N); self.bvf = np.zeros(N)
        self.mass, self.nu = fm.mass.copy(), fm.nu.copy()
        self.type, self.solid = fm.type.copy(), fm.solid.copy()
        Sc = fm.num_chem_species
        self.Sc = Sc
        self.C = fm.u0[:, :Sc].astype(np.float64).copy() if Sc else np.zeros((N, 0))
        self.Q = np.zeros((N, Sc))
        self.step_no = 0
        self.bc_rules = parse_bc(fm.bc_source, fm.type_constants)
        self.nbr = None

    # ------------------------------------------------------------------ boundary conditions
    def apply_bc(self, idx=None):
        for conds, assigns in self.bc_rules:
            m = np.ones(self.N, bool)
            for c in conds:
                if c[0] == "x":
                    m &= (self.x[:, c[1]] >= c[3]) if c[2] == ">=" else (self.x[:, c[1]] <= c[3])
                else:
                    m &= (self.type == c[1])
            for tgt, k, val in assigns:
                if tgt == "v":
                    self.v[m, k] = val
                elif tgt == "C":
                    self.C[m, k] = val
                elif tgt == "nu":
                    self.nu[m] = val
                elif tgt == "rho":
                    self.rho[m] = val

    # ------------------------------------------------------------------ neighbour search
    def find_neighbors(self, xq, xdata):
        """ANN fixed-radius semantics (kd_fix_rad_search.cpp:160-178): include iff 0 < sum_d (q_d-p_d)^2 <= h*h, summed in
        axis order with separate multiply/add; then particle.cpp:160-162 drops sqrt(r2) > h.  Returns CSR sorted by r2."""
        N, dim, h = self.N, self.dim, self.h
        h2 = h * h
        rows, cols, d2s = [], [], []
        B = max(1, 4_000_000 // max(N, 1))
        for b in range(0, N, B):
            q = xq[b:b + B]
            t = q[:, None, 0] - xdata[None, :, 0]
            d2 = t * t
            for d in range(1, dim):
                t = q[:, None, d] - xdata[None, :, d]
                d2 = d2 + t * t
            ok = (d2 <= h2) & (d2 != 0.0)
            ok &= ~(np.sqrt(d2) > h)
            r, c = np.nonzero(ok)
            rows.append(r + b); cols.append(c); d2s.append(d2[r, c])
        rows, cols, d2s = np.concatenate(rows), np.concatenate(cols), np.concatenate(d2s)
        order = np.lexsort((cols, d2s, rows))
        rows, cols, d2s = rows[order], cols[order], d2s[order]
        ptr = np.zeros(N + 1, np.int64)
        np.add.at(ptr, rows + 1, 1)
        ptr = np.cumsum(ptr)
        r = np.sqrt(d2s)
        # add_to_neighbor_list (particle.cpp:164-187)
        R = r / h
        a = alpha(dim, h)
        dWdr = a * (-12 * r / (h * h)) * ((1 - R) * (1 - R))
        ih = 1.0 / h
        ihsq = ih * ih
        dhr = h - r
        wfd = -25.066903536973515383e0 * dhr * dhr * ihsq * ihsq * ihsq * ih
        mi, mj = self.mass[rows], self.mass[cols]
        ri, rj = self.rho[rows], self.rho[cols]
        Dij = -2.0 * (mi * mj) / (mi + mj) * (ri + rj) / (ri * rj) * d2s * wfd / (d2s + 0.01 * h * h)
        self.nbr = dict(ptr=ptr, i=rows, j=cols, dist=r, dWdr=dWdr, Dij=Dij)
        return self.nbr

    # ------------------------------------------------------------------ substeps
    def take_step1(self):
        """simulate.cpp:56-109."""
        dt = self.dt
        if not self.static:
            mv = self.solid == 0
            self.v[mv] = self.v[mv] + 0.5 * dt * self.F[mv]
            self.vt[mv] = self.v[mv] + 0.5 * dt * self.Fbp[mv]
            self.x[mv] = self.x[mv] + dt * self.vt[mv]
            self.rho[mv] = self.rho[mv] + 0.5 * dt * self.Frho[mv]
        if self.step_no > 0:
            self.C += self.Q * dt * 0.5
        self.apply_bc()
        self.F[:] = np.asarray(self.fm.gravity)[None, :]
        self.Fbp[:] = 0.0
        self.Frho[:] = 0.0
        self.Q[:] = 0.0
        self.old_rho = self.rho.copy()

    def pairwise_force(self):
        """model.cpp:39-191 over the frozen neighbour list; dx, dv, rho, C live."""
        nb, h, dim = self.nbr, self.h, self.dim
        i, j, r, dWdr = nb["i"], nb["j"], nb["dist"], nb["dWdr"]
        rho0, P0 = self.fm.rho0, self.fm.P0
        dx = np.zeros((len(i), 3)); dv = np.zeros((len(i), 3))
        dx[:, :dim] = self.x[i, :dim] - self.x[j, :dim]
        dv[:, :dim] = self.v[i, :dim] - self.v[j, :dim]
        dv_dx = np.zeros(len(i))
        for d in range(dim):
            dv_dx += dv[:, d] * dx[:, d]
        rho_i, rho_j, m_i, m_j = self.rho[i], self.rho[j], self.mass[i], self.mass[j]
        Pi = P0 * (rho_i / rho0 - 1.0)
        Pj = P0 * (rho_j / rho0 - 1.0)
        pg = Pi / (rho_i * rho_i) + Pj / (rho_j * rho_j)
        pg = np.where(pg < 0, -Pi / (rho_i * rho_i) + Pj / (rho_j * rho_j), pg)
        reg = r + 0.001 * h
        fp = -1.0 * m_j * pg * dWdr / reg
        nu_i, nu_j = self.nu[i], self.nu[j]
        fv = m_j * (2.0 * (nu_i * nu_j) / (nu_i + nu_j)) * (1 / reg) * dWdr / (rho_i * rho_j)
        vv = (m_i / rho_i) ** 2 + (m_j / rho_j) ** 2
        fbp = -10.0 * P0 * (1.0 / m_i) * vv * dWdr / reg
        vi, vj, vti, vtj = self.v[i], self.v[j], self.vt[i], self.vt[j]
        ft = np.zeros((len(i), 3))
        for a in range(3):
            acc = np.zeros(len(i))
            for b in range(3):
                T = 0.5 * ((rho_i * vi[:, a] * (vti[:, b] - vi[:, b])) + (rho_j * vj[:, a] * (vtj[:, b] - vj[:, b])))
                acc = acc + T * dx[:, b]
            ft[:, a] = (1.0 / m_i) * vv * acc * dWdr / reg
        for d in range(dim):
            np.add.at(self.F[:, d], i, fp * dx[:, d] + fv * dv[:, d] + ft[:, d])
            np.add.at(self.Fbp[:, d], i, fbp * dx[:, d])
        w = vi - vti
        wj = vj - vtj
        frho = rho_i * (m_j / rho_j) * dv_dx * (1 / reg) * dWdr \
            - (m_j / rho_j) * (rho_i * (w[:, 0] * dx[:, 0] + w[:, 1] * dx[:, 1] + w[:, 2] * dx[:, 2])
                               + rho_j * (wj[:, 0] * dx[:, 0] + wj[:, 1] * dx[:, 1] + wj[:, 2] * dx[:, 2])) * (1.0 / reg) * dWdr
        np.add.at(self.Frho, i, frho)
        if self.Sc:
            wfd = (1.0 / reg) * dWdr
            base = 2.0 * ((m_i * m_j) / (m_i + m_j)) * ((rho_i + rho_j) / (rho_i * rho_j)) * (r * r) * wfd / ((r * r) + 0.01 * h * h)
            dm = self.fm.diffusion_matrix.reshape(-1)
            for s in range(self.Sc):
                k = self.Sc * (self.type[i] - 1) + s          # model.cpp:163 (transposed index, mirrored)
                if getattr(self, "corrected_pde_index", False):   # SSB_FLAG_CORRECTED_PDE_INDEX: the entry simulate_rdme.cpp:146 reads
                    k = s * self.fm.num_types + (self.type[i] - 1)
                Dk = np.where(k < dm.size, dm[np.minimum(k, dm.size - 1)], 0.0)
                np.add.at(self.Q[:, s], i, Dk * (self.C[i, s] - self.C[j, s]) * base)
            self.det_reactions(getattr(self, "corrected_stoich", False))

    def det_reactions(self, corrected=False):
        """model.cpp:181-189: Q[s] += stoichiometric_matrix[num_chem_rxns*rxn + s] * det_rxn(C, t, vol, data_fn, type).
        The index is the reference's (transposed on the species x rxn row-major table; out-of-bounds reads defined 0)."""
        fm = self.fm
        Rc, Sc, S, R = fm.num_chem_rxns, self.Sc, fm.num_species, fm.num_reactions
        if Rc == 0:
            return
        dense = fm.N_dense.reshape(-1)
        vol = self.mass / self.rho
        t = self.step_no * self.dt
        env = {k: v for k, v in fm.parameters.items()}
        env.update({k: v for k, v in fm.type_constants.items()})
        env.update(vol=vol, t=t, pow=np.power, exp=np.exp, log=np.log, sqrt=np.sqrt, sin=np.sin, cos=np.cos,
                   x=[self.C[:, q] for q in range(Sc)], data_fn=[fm.data_fn[q] for q in range(fm.num_data_fn)])
        for rxn, r in enumerate(fm.reactions):
            flux = eval(r.ode_propensity, {"__builtins__": {}}, env) * np.ones(self.N)     # C arithmetic text == Python text here
            if r.restrict_to:
                ok = np.zeros(self.N, bool)
                for tname in r.restrict_to:
                    ok |= self.type == (fm.type_constants[tname] if not str(tname).isdigit() else int(tname))
                flux = np.where(ok, flux, 0.0)
            for s_ in range(Sc):
                k = (s_ * R + rxn) if corrected else (Rc * rxn + s_)
                nval = dense[k] if k < S * R else 0
                self.Q[:, s_] += nval * flux

    def _W(self, r):
        R = r / self.h
        return alpha(self.dim, self.h) * ((1 + 3 * R) * (1 - R) ** 3)

    def take_step2(self):
        """simulate.cpp:132-173 with the serial Gauss–Seidel visibility of rho in the BVF sweep (model.cpp:285-293)."""
        dt, nb, h, dim = self.dt, self.nbr, self.h, self.dim
        i, j, r, dWdr = nb["i"], nb["j"], nb["dist"], nb["dWdr"]
        rho_pre = self.rho.copy()
        if not self.static:
            fl = self.solid == 0
            self.v[fl] = self.v[fl] + 0.5 * dt * self.F[fl]
            rho_c = self.rho.copy()
            if self.step_no % 20 == 0:                      # filterDensity (model.cpp:194-233)
                W = self._W(r)
                num = np.zeros(self.N); den = np.zeros(self.N)
                np.add.at(num, i, self.old_rho[j] * W)
                np.add.at(den, i, W)
                with np.errstate(invalid="ignore", divide="ignore"):
                    rho_c = num / den
            rho_c = np.where(fl, rho_c + 0.5 * dt * self.Frho, rho_c)
            # density each particle shows to LATER particles: post corrector and post trailing BC
            keep = (self.v.copy(), self.C.copy(), self.nu.copy())
            self.rho = rho_c.copy()
            self.apply_bc()
            rho_post = self.rho.copy()
            self.v, self.C, self.nu = keep                  # only the effect on rho is visible early
            # BVF for fluid particles
            W = self._W(r)
            rho_seen = np.where(j <= i, rho_post[j], rho_pre[j])
            volj2 = (self.mass[j] / rho_seen) ** 2
            sj = self.solid[j] != 0
            vos = np.zeros(self.N); vtot = np.zeros(self.N); nw = np.zeros((self.N, 3))
            np.add.at(vos, i, np.where(sj, volj2 * W, 0.0))
            np.add.at(vtot, i, volj2 * W)
            dx = np.zeros((len(i), 3))
            dx[:, :dim] = self.x[i, :dim] - self.x[j, :dim]
            for d in range(3):
                np.add.at(nw[:, d], i, np.where(sj, volj2 * dx[:, d] * dWdr / (r + 0.001 * h), 0.0))
            with np.errstate(invalid="ignore", divide="ignore"):
                nw = nw / vtot[:, None]
                norm = np.sqrt(nw[:, 0] ** 2 + nw[:, 1] ** 2 + nw[:, 2] ** 2)
                normal = -nw / norm[:, None]
                bvf = np.abs(vos / vtot)
            self.bvf = np.where(fl, bvf, self.bvf)
            self.vt[fl, :dim] = 0.0
            vdn = self.v[:, 0] * normal[:, 0] + self.v[:, 1] * normal[:, 1] + self.v[:, 2] * normal[:, 2]
            bounce = fl & (self.bvf >= 0.5)
            for d in range(dim):
                self.v[bounce, d] = -self.v[bounce, d] + 2.0 * np.fmax(0.0, vdn[bounce]) * normal[bounce, d]
        self.C += self.Q * dt * 0.5
        self.apply_bc()

    def step(self):
        """One engine step without the RDME (simulate_threads.cpp:232-281)."""
        x0 = self.x.copy()                                  # kd-tree snapshot (simulate_threads.cpp:100-104)
        if self.step_no == 0:
            self.find_neighbors(self.x, x0)                 # simulate.cpp:61-63 (list used by this step's forces)
        self.take_step1()
        if self.step_no > 0 and not self.static:
            self.find_neighbors(self.x, x0)                 # simulate.cpp:121-123: live query vs stale snapshot
        self.pairwise_force()
        self.take_step2()
        self.step_no += 1

    def ddiag(self):
        """simulate_rdme.cpp:131-152."""
        nb = self.nbr
        dm = self.fm.diffusion_matrix
        S = self.fm.num_stoch_species
        out = np.zeros((self.N, S))
        for s in range(S):
            np.add.at(out[:, s], nb["i"], dm[s, self.type[nb["j"]] - 1] * nb["Dij"])
        return out
