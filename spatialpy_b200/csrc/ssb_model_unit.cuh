// ssb_model_unit.cuh — the model-specialised kernels of the ssa_sdpd hot path (sm_100a).
//
// Included at the END of a generated model .cu (spatialpy_b200/codegen.py), after it has defined
//   SSB_SC, SSB_RC, SSB_SD, SSB_RD, SSB_NDF, SSB_NTYPES, SSB_S, SSB_R            (compile-time sizes)
//   namespace ssb_gen { parameters P<i>, type_<name> constants, rxn_<j>(), det_<j>(), N_dense[], dep graph,
//                       struct Particle (BC proxy), applyBoundaryConditions(Particle*, const System*) }
// so that species loops unroll and x[] / C[] / rrate[] live in registers (SURVEY.md §8b "device-function ABI").
// The kernels restate, per substep, E/src/simulate.cpp (take_step1 / compute_forces / take_step2),
// E/src/model.cpp (pairwiseForce / filterDensity / computeBoundaryVolumeFraction / applyBoundaryVolumeFraction)
// and E/src/simulate_rdme.cpp (propensity init, NSM event execution) — citations at each block.
// E = /root/reference/spatialpy/solvers/c_base/ssa_sdpd-c-simulation-engine.
#pragma once
#include <math.h>
#include <cooperative_groups.h>
#include "ssb_device.cuh"
#include "ssb_unit_abi.h"

#ifndef SSB_BLOCK
#define SSB_BLOCK 128
#endif
// moving domains: a second, 32-byte gather record per particle — {1/rho, P/rho^2, C[0], C[1]} — written by the predictor, so that
// the force sweep fetches the neighbour's concentrations and pressure term with ONE load instead of two gathers and a division
#ifndef SSB_REC2
#define SSB_REC2 (SSB_SC <= 2)
#endif

namespace ssb_unit {

using ssb_gen::Particle;
using ssb_gen::System;

// ---------------------------------------------------------------------------------------------
// Boundary-condition proxy: load a particle into registers, run the user text, store what it may assign
// (v, nu, rho, C — spatialpy/core/boundarycondition.py:147-166).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bc_load(const SsbView &V, int i, Particle &p) {
#pragma unroll
    for (int d = 0; d < 3; d++) { p.x[d] = V.x[d][i]; p.v[d] = V.v[d][i]; }
    p.nu = V.nu[i]; p.rho = V.rho[i]; p.mass = V.mass[i]; p.type = V.type[i]; p.id = V.id[i];
    p.solidTag = V.solid[i];
#pragma unroll
    for (int s = 0; s < SSB_SC; s++) p.C[s] = V.C[(size_t) s * V.N + i];
}

__device__ __forceinline__ System make_system(const SsbView &V, unsigned step) {
    System sys;
    sys.dt = V.dt; sys.h = V.h; sys.rho0 = V.rho0; sys.c0 = V.c0; sys.P0 = V.P0;
    sys.dimension = V.dim; sys.current_step = step; sys.static_domain = V.static_domain;
    return sys;
}

// ---------------------------------------------------------------------------------------------
// K2  take_step1 (E/src/simulate.cpp:56-109) + check_particle_nan (E/src/particle.cpp:88-134)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SSB_BLOCK) k_predictor(SsbView V, unsigned step) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V.N) {
    Particle p;
    bc_load(V, i, p);
    // NaN / Inf guard (particle.cpp:89-96) — the reference exit(1)s; we raise the device error flag.
    bool bad = !isfinite(p.x[0]) || !isfinite(p.x[1]) || !isfinite(p.x[2]) ||
               !isfinite(p.v[0]) || !isfinite(p.v[1]) || !isfinite(p.v[2]) || !isfinite(p.rho);
    if (bad) atomicCAS(V.err_flag, 0, 1 /*SSB_ERR_NAN*/);
    const double dt = V.dt;
    if (!V.static_domain) {
        // snapshot = what the kd-tree of this step holds (simulate_threads.cpp:100-104)
#pragma unroll
        for (int d = 0; d < 3; d++) V.x0[d][i] = p.x[d];
    }
    if (p.solidTag == 0 && V.static_domain == 0) {     // simulate.cpp:68-79
#pragma unroll
        for (int d = 0; d < 3; d++) {
            p.v[d] = p.v[d] + 0.5 * dt * V.F[d][i];
            double vt = p.v[d] + 0.5 * dt * V.Fbp[d][i];
            V.vt[d][i] = vt;
            p.x[d] = p.x[d] + dt * vt;      // (the Verlet-skin displacement bookkeeping of this update ran ahead of it: k_lookahead, ssb_core.cu)
            V.x[d][i] = p.x[d];
        }
        p.rho = p.rho + 0.5 * dt * V.Frho[i];
    }
    if (step > 0) {                                     // simulate.cpp:81-85
#pragma unroll
        for (int s = 0; s < SSB_SC; s++) p.C[s] += V.Q[(size_t) s * V.N + i] * dt * 0.5;
    }
    System sys = make_system(V, step);
    ssb_gen::applyBoundaryConditions(&p, &sys);         // simulate.cpp:88
#pragma unroll
    for (int d = 0; d < 3; d++) {                       // simulate.cpp:92-97
        V.v[d][i] = p.v[d];
        V.F[d][i] = V.gravity[d];
        V.Fbp[d][i] = 0.0;
    }
    V.Frho[i] = 0.0;
    V.nu[i] = p.nu;
    V.rho[i] = p.rho;
    V.old_rho[i] = p.rho;                               // simulate.cpp:106
    if (!V.static_domain && V.rec) {
        double2 *r2 = reinterpret_cast<double2 *>(V.rec + (size_t) i * 16);
        const long long bits = ((long long) (unsigned) V.gid[i]) | ((long long) (p.type & 0xffff) << 32) | ((long long) (p.solidTag & 0xffff) << 48);
        r2[0] = make_double2(V.x0[0][i], V.x0[1][i]);
        r2[1] = make_double2(V.x0[2][i], p.x[0]);
        r2[2] = make_double2(p.x[1], p.x[2]);
        r2[3] = make_double2(p.v[0], p.v[1]);
        r2[4] = make_double2(p.v[2], V.vt[0][i]);
        r2[5] = make_double2(V.vt[1][i], V.vt[2][i]);
        r2[6] = make_double2(p.rho, p.mass);
        r2[7] = make_double2(p.nu, __longlong_as_double(bits));
#if SSB_REC2
        if (V.rec2) {
            const double inv_rho = 1.0 / p.rho;
            double2 *e2 = reinterpret_cast<double2 *>(V.rec2 + (size_t) i * 4);
            e2[0] = make_double2(inv_rho, (V.P0 * (p.rho / V.rho0 - 1.0)) * inv_rho * inv_rho);
            e2[1] = make_double2(SSB_SC > 0 ? p.C[0] : 0.0, SSB_SC > 1 ? p.C[SSB_SC > 1 ? 1 : 0] : 0.0);
        }
#endif
    }
#pragma unroll
    for (int s = 0; s < SSB_SC; s++) {
        V.C[(size_t) s * V.N + i] = p.C[s];
        V.Q[(size_t) s * V.N + i] = 0.0;                // simulate.cpp:101-103
    }
    }
}

// ---------------------------------------------------------------------------------------------
// K3  compute_forces -> pairwiseForce (E/src/model.cpp:39-191).  One thread per particle over its
// index-only neighbour list; r and dWdr are recomputed from (live x_i, snapshot x0_j) exactly as
// add_to_neighbor_list froze them (particle.cpp:160-178); dx, dv, rho, C are live (model.cpp:102-108).
// FULL = false on static domains: only the chemistry flux Q is consumed there (simulate.cpp:68,137).
// ---------------------------------------------------------------------------------------------
template <bool FULL>
__global__ void __launch_bounds__(SSB_BLOCK) k_force(SsbView V, unsigned step) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V.N) return;
    const int N = V.N, dim = V.dim;
    const double h = V.h, rho0 = V.rho0, P0 = V.P0;
    const double alpha = ssb_alpha(dim, h);
    const double xi0 = V.x[0][i], xi1 = V.x[1][i], xi2 = V.x[2][i];
    const double rho_i = V.rho[i], m_i = V.mass[i];
    const int type_i = V.type[i];
    double Ci[SSB_SC > 0 ? SSB_SC : 1], Qi[SSB_SC > 0 ? SSB_SC : 1], Dk[SSB_SC > 0 ? SSB_SC : 1];
#pragma unroll
    for (int s = 0; s < SSB_SC; s++) {
        Ci[s] = V.C[(size_t) s * N + i];
        Qi[s] = V.Q[(size_t) s * N + i];
        // NOTE the reference indexes the species-major table as [S_c*(type-1)+s] (model.cpp:163) — mirrored,
        // with a bounds guard (the read is in-bounds whenever S == num_types or D is type-independent).
        int k = ssb_pde_dindex(V, SSB_SC, type_i, s);
        Dk[s] = (k >= 0 && k < SSB_S * V.num_types) ? V.dmat[k] : 0.0;
    }
    double vi[3], vti[3], nu_i = 0, Pi = 0, Fa[3], Fb[3], Frho = 0;
    if (FULL) {
#pragma unroll
        for (int d = 0; d < 3; d++) { vi[d] = V.v[d][i]; vti[d] = V.vt[d][i]; Fa[d] = V.F[d][i]; Fb[d] = V.Fbp[d][i]; }
        nu_i = V.nu[i];
        Pi = P0 * (rho_i / rho0 - 1.0);                 // model.cpp:55
        Frho = V.Frho[i];
    }
    const int cnt = V.nbr_count[i];
    for (int k = 0; k < cnt; k++) {
        const int j = V.nbr[(size_t) k * N + i];
        const double d2 = ssb_dist2(dim, xi0, xi1, xi2, V.x0[0][j], V.x0[1][j], V.x0[2][j]);
        const double r = sqrt(d2);                      // n->dist (particle.cpp:160)
        const double dWdr = ssb_dWdr(alpha, r, h);      // n->dWdr (particle.cpp:178)
        const double rho_j = V.rho[j], m_j = V.mass[j];
        double dx[3] = {0.0, 0.0, 0.0};
        dx[0] = xi0 - V.x[0][j];
        if (dim > 1) dx[1] = xi1 - V.x[1][j];
        if (dim > 2) dx[2] = xi2 - V.x[2][j];
        const double inv_reg = 1.0 / (r + 0.001 * h);
        if (FULL) {
            double dv[3] = {0.0, 0.0, 0.0}, vj[3], vtj[3];
#pragma unroll
            for (int d = 0; d < 3; d++) { vj[d] = V.v[d][j]; vtj[d] = V.vt[d][j]; }
            double dv_dx = 0.0;
#pragma unroll
            for (int d = 0; d < 3; d++) if (d < dim) { dv[d] = vi[d] - vj[d]; dv_dx += dv[d] * dx[d]; }
            const double nu_j = V.nu[j];
            const double Pj = P0 * (rho_j / rho0 - 1.0);                        // model.cpp:108
            double pg = Pi / (rho_i * rho_i) + Pj / (rho_j * rho_j);            // model.cpp:111
            if (pg < 0) pg = -Pi / (rho_i * rho_i) + Pj / (rho_j * rho_j);      // model.cpp:112
            const double fp = -1.0 * m_j * pg * dWdr / (r + 0.001 * h);         // model.cpp:115
            const double fv = m_j * (2.0 * (nu_i * nu_j) / (nu_i + nu_j)) * inv_reg * dWdr / ((rho_i * rho_j));  // :118
            const double voli = m_i / rho_i, volj = m_j / rho_j;
            const double vv = voli * voli + volj * volj;                        // pow(.,2)+pow(.,2)
            const double fbp = -10.0 * P0 * (1.0 / m_i) * vv * dWdr / (r + 0.001 * h);   // model.cpp:121
            double ft[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {                                       // model.cpp:124-132
                double acc = 0.0;
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    double T = 0.5 * ((rho_i * vi[a] * (vti[b] - vi[b])) + (rho_j * vj[a] * (vtj[b] - vj[b])));
                    acc += T * dx[b];
                }
                ft[a] = (1.0 / m_i) * vv * acc * dWdr / (r + 0.001 * h);
            }
#pragma unroll
            for (int d = 0; d < 3; d++) if (d < dim) {                          // model.cpp:135-138
                Fa[d] += fp * dx[d] + fv * dv[d] + ft[d];
                Fb[d] += fbp * dx[d];
            }
            // model.cpp:143-146 (the c0 term is multiplied by 0.0 in the reference and dropped here)
            Frho = Frho + rho_i * volj * dv_dx * inv_reg * dWdr
                   - volj * (rho_i * ((vi[0] - vti[0]) * dx[0] + (vi[1] - vti[1]) * dx[1] + (vi[2] - vti[2]) * dx[2])
                             + rho_j * ((vj[0] - vtj[0]) * dx[0] + (vj[1] - vtj[1]) * dx[1] + (vj[2] - vtj[2]) * dx[2])) * inv_reg * dWdr;
        }
        if (SSB_SC > 0) {                                                       // model.cpp:152-170
            const double wfd = inv_reg * dWdr;
            const double dQc_base = 2.0 * ((m_i * m_j) / (m_i + m_j)) * ((rho_i + rho_j) / (rho_i * rho_j)) * (r * r) * wfd / ((r * r) + 0.01 * h * h);
#pragma unroll
            for (int s = 0; s < SSB_SC; s++) Qi[s] += Dk[s] * (Ci[s] - V.C[(size_t) s * N + j]) * dQc_base;
        }
    }
    if (FULL) {
#pragma unroll
        for (int d = 0; d < 3; d++) { V.F[d][i] = Fa[d]; V.Fbp[d][i] = Fb[d]; }
        V.Frho[i] = Frho;
    }
    if (SSB_SC > 0) {
        // deterministic reaction right-hand side (model.cpp:181-189)
        if (SSB_RC > 0) {
            const double vol = m_i / rho_i;
            const double cur_time = step * V.dt;
            double df[SSB_NDF > 0 ? SSB_NDF : 1];
#pragma unroll
            for (int q = 0; q < SSB_NDF; q++) df[q] = V.data_fn[(size_t) q * N + i];
            double flux[SSB_RC > 0 ? SSB_RC : 1];
            ssb_gen::eval_det(Ci, cur_time, vol, df, type_i, flux);
#pragma unroll
            for (int rxn = 0; rxn < SSB_RC; rxn++) {
#pragma unroll
                for (int s = 0; s < SSB_SC; s++) {
                    int nval;
                    if (V.flags & 2u /*SSB_FLAG_CORRECTED_STOICH*/) {
                        nval = ssb_gen::N_dense(s * SSB_R + rxn);
                    } else {
                        // reference: stoichiometric_matrix[num_chem_rxns*rxn + s] on the species x rxn row-major table
                        // (model.cpp:186) — transposed, and out of bounds when rxn >= S; out-of-bounds reads are defined 0 here.
                        int kk = SSB_RC * rxn + s;
                        nval = (kk < SSB_S * SSB_R) ? ssb_gen::N_dense(kk) : 0;
                    }
                    Qi[s] += nval * flux[rxn];
                }
            }
        }
#pragma unroll
        for (int s = 0; s < SSB_SC; s++) V.Q[(size_t) s * N + i] = Qi[s];
    }
}

// ---------------------------------------------------------------------------------------------
// K3 (moving domains, optimised form of k_force<true>): same physics as model.cpp:39-191, restructured for B200:
//   * neighbour data comes from ONE 128-byte gather record per neighbour (4 sectors instead of 17 scattered 8-byte gathers);
//   * the three per-pair reciprocals 1/(r+0.001h), 1/(nu_i+nu_j), 1/((m_i+m_j)(r^2+0.01h^2)) share ONE division;
//   * the 3x3 transport tensor contraction T.dx is factorised as 0.5*(rho_i v_i (w_i.dx) + rho_j v_j (w_j.dx)), w = vt - v;
//   * K6 (D_i_j, Ddiag, max for the window controller; simulate_rdme.cpp:131-152) is fused in — it shares r and the
//     (m, rho, r^2) factor with the chemistry flux.
// Results differ from the literal evaluation order by a few ulp per pair (parity gate: 1e-12 of the field scale).
//
//
// What bounds it (ncu, box at 1 M particles, profiles/r2c_force_coop1_metrics.csv): the L1 data pipe — 74 % busy with ~190
// wavefronts per warp and candidate (every lane's record is its own 128-byte line: an LDG.E.256 of a warp replays 32 times, and the
// two C_j gathers another 2 x 32) — and, behind it, the long scoreboard (9.5 of 13 warp cycles per issued instruction); fp64 pipe
// 32 %, DRAM 4 %, 16 warps per SM at 128 registers.  Measured alternatives, both deleted (one pair body stays):
//   * quad-cooperative gather (4 lanes fetch 4 records together, 4x4 shuffle transpose): L1 wavefronts 74 % -> 47 %, but twice the
//     instructions (issue 28 % -> 52 %) and the C_j gathers untouched: same time (1.79 vs 1.84 ms, profiles/r2c_force_coop4_metrics.csv);
//   * shared-memory staging per row segment (<= 128 particles of one cell row, nine staged slot ranges, cursor over the lists):
//     L1 34 %, long scoreboard 2.8 — but the nine phases leave 14 of 32 lanes active on average and double the instruction
//     count: 2.66 ms (profiles/r2e_force_rows_metrics.csv); the round-1 bitmap tile form was 2x slower still.
// What helped: the neighbour's concentrations and its 1/rho, P/rho^2 now come from ONE extra 32-byte sector (rec2, written by the
// predictor) instead of two 8-byte gathers + a division per pair: 5 replayed loads per candidate instead of 6.
// ---------------------------------------------------------------------------------------------

// per-particle state of the sweep + the pair body (ONE copy, whatever feeds it the records)
struct ForceSweep {
    int N, dim, filter, num_types;
    double h, P0, inv_h, c12, eps_r, eps2, h2x, ih7c, rho0, inv_rho0;
    bool inv_exact;
    double xi0, xi1, xi2, vi0, vi1, vi2, wi0, wi1, wi2, rho_i, m_i, nu_i;
    double inv_rho_i, aP_i, volsq_i, inv_m_i, rv0, rv1, rv2;
    int type_i;
    double Ci[SSB_SC > 0 ? SSB_SC : 1], Qi[SSB_SC > 0 ? SSB_SC : 1], Dk[SSB_SC > 0 ? SSB_SC : 1], Dd[SSB_SD > 0 ? SSB_SD : 1];
    double F0, F1, F2, B0, B1, B2, Frho;
    const double *Cg, *dmat;

    __device__ __forceinline__ void init(const SsbView &V, int i) {
        N = V.N; dim = V.dim; filter = V.filter; num_types = V.num_types;
        h = V.h; P0 = V.P0;
        inv_h = 1.0 / h;
        c12 = ssb_alpha(dim, h) * (-12.0) / (h * h);
        eps_r = 0.001 * h; eps2 = 0.01 * h * h;
        h2x = __dmul_rn(h, h);
        ih7c = 25.066903536973515383e0 * inv_h * inv_h * inv_h * inv_h * inv_h * inv_h * inv_h;
        const double *ri = V.rec + (size_t) i * 16;
        const ssb_d4 a0 = ssb_ld256(ri), a1 = ssb_ld256(ri + 4), a2 = ssb_ld256(ri + 8), a3 = ssb_ld256(ri + 12);
        xi0 = a0.d; xi1 = a1.a; xi2 = a1.b;
        vi0 = a1.c; vi1 = a1.d; vi2 = a2.a;
        wi0 = a2.b - vi0; wi1 = a2.c - vi1; wi2 = a2.d - vi2;
        rho_i = a3.a; m_i = a3.b; nu_i = a3.c;
        // 1/rho, P/rho^2 and m/rho of the neighbour are recomputed from its record (two divisions per pair) rather than gathered
        // from a second per-particle record: a fifth sector per pair cost more than the divisions (measured 0.77 -> 0.70 ms at 1 M).
        rho0 = V.rho0; inv_rho0 = 1.0 / V.rho0;
        inv_exact = (__double_as_longlong(rho0) & 0x000fffffffffffffLL) == 0;    // power of two: rho * (1/rho0) == rho / rho0 bit for bit
        inv_rho_i = 1.0 / rho_i;
        aP_i = (P0 * (rho_i / rho0 - 1.0)) * inv_rho_i * inv_rho_i;
        const double vol_i = m_i * inv_rho_i;
        volsq_i = vol_i * vol_i; inv_m_i = 1.0 / m_i;
        rv0 = rho_i * vi0; rv1 = rho_i * vi1; rv2 = rho_i * vi2;
        type_i = (int) ((__double_as_longlong(a3.d) >> 32) & 0xffff);
        Cg = V.C; dmat = V.dmat;
#pragma unroll
        for (int s = 0; s < SSB_SC; s++) {
            Ci[s] = V.C[(size_t) s * N + i];
            Qi[s] = V.Q[(size_t) s * N + i];
            int k = ssb_pde_dindex(V, SSB_SC, type_i, s);
            Dk[s] = (k >= 0 && k < SSB_S * V.num_types) ? V.dmat[k] : 0.0;
        }
#pragma unroll
        for (int s = 0; s < SSB_SD; s++) Dd[s] = 0.0;
        F0 = V.F[0][i]; F1 = V.F[1][i]; F2 = V.F[2][i];
        B0 = V.Fbp[0][i]; B1 = V.Fbp[1][i]; B2 = V.Fbp[2][i];
        Frho = V.Frho[i];
    }

    // one candidate j with its record in c0..c3 (rejected unless it is in ANN's exact set for THIS step's snapshot)
    // c0..c3 = the neighbour's gather record; e = its derived sector rec2 = {1/rho, P/rho^2, C[0], C[1]} when SSB_REC2, else unused
    // (1/rho and P/rho^2 are then recomputed here and the concentrations gathered from V.C)
    __device__ __forceinline__ void pair(const ssb_d4 &c0, const ssb_d4 &c1, const ssb_d4 &c2, const ssb_d4 &c3, const ssb_d4 &e, const int j) {
#if SSB_REC2
        const double inv_rho_j = e.a, aP_j = e.b;
#else
        const double inv_rho_j = 1.0 / c3.a;
        const double aP_j = (P0 * ((inv_exact ? c3.a * inv_rho0 : c3.a / rho0) - 1.0)) * inv_rho_j * inv_rho_j;
#endif
        const double vol_j = c3.b * inv_rho_j;
        const double d2 = ssb_dist2(dim, xi0, xi1, xi2, c0.a, c0.b, c0.c);      // live x_i vs snapshot x0_j (particle.cpp:160)
        const double r = sqrt(d2);
        // candidate list -> ANN's exact set (the record is fetched before this test on purpose: a rejected candidate wastes
        // 4 sectors, a dependent second round trip per accepted neighbour would cost far more)
        if (filter && !((d2 <= h2x) && (d2 != 0.0) && !(r > h))) return;
        double dx0 = xi0 - c0.d, dx1 = 0.0, dx2 = 0.0;
        if (dim > 1) dx1 = xi1 - c1.a;
        if (dim > 2) dx2 = xi2 - c1.b;
        const double vj0 = c1.c, vj1 = c1.d, vj2 = c2.a;
        const double wj0 = c2.b - vj0, wj1 = c2.c - vj1, wj2 = c2.d - vj2;
        const double rho_j = c3.a, m_j = c3.b, nu_j = c3.c;
        // one division for three reciprocals
        const double reg = r + eps_r, nusum = nu_i + nu_j, r2 = r * r;
        const double md = (m_i + m_j) * (r2 + eps2);
        const double rn = reg * nusum;
        const double q = 1.0 / (rn * md);
        const double inv_reg = q * (nusum * md), inv_nusum = q * (reg * md), inv_md = q * rn;
        const double R = r * inv_h, omR = 1.0 - R;
        const double dWdr = c12 * r * (omR * omR);                           // particle.cpp:178
        const double wr = dWdr * inv_reg;                                    // dWdr / (r + 0.001 h)
        double dv0 = vi0 - vj0, dv1 = 0.0, dv2 = 0.0;
        if (dim > 1) dv1 = vi1 - vj1;
        if (dim > 2) dv2 = vi2 - vj2;
        const double dv_dx = dv0 * dx0 + dv1 * dx1 + dv2 * dx2;
        double pg = aP_i + aP_j;                                             // model.cpp:111
        if (pg < 0) pg = -aP_i + aP_j;                                       // model.cpp:112
        const double fp = -m_j * pg * wr;                                    // model.cpp:115
        const double fv = m_j * (2.0 * (nu_i * nu_j) * inv_nusum) * wr * (inv_rho_i * inv_rho_j);   // model.cpp:118
        const double vv = volsq_i + vol_j * vol_j;
        const double fbp = -10.0 * P0 * inv_m_i * vv * wr;                   // model.cpp:121
        const double widx = wi0 * dx0 + wi1 * dx1 + wi2 * dx2;
        const double wjdx = wj0 * dx0 + wj1 * dx1 + wj2 * dx2;
        const double ftc = 0.5 * inv_m_i * vv * wr;                          // model.cpp:124-132
        const double rj_w = rho_j * wjdx;
        const double ft0 = ftc * (rv0 * widx + vj0 * rj_w);
        const double ft1 = ftc * (rv1 * widx + vj1 * rj_w);
        const double ft2 = ftc * (rv2 * widx + vj2 * rj_w);
        F0 += fp * dx0 + fv * dv0 + ft0;                                     // model.cpp:135-138
        B0 += fbp * dx0;
        if (dim > 1) { F1 += fp * dx1 + fv * dv1 + ft1; B1 += fbp * dx1; }
        if (dim > 2) { F2 += fp * dx2 + fv * dv2 + ft2; B2 += fbp * dx2; }
        Frho += wr * vol_j * (rho_i * dv_dx + rho_i * widx + rj_w);          // model.cpp:143-146
        if (SSB_SC > 0 || SSB_SD > 0) {
            const double G = 2.0 * (m_i * m_j) * (inv_rho_i + inv_rho_j) * r2 * inv_md;   // shared by model.cpp:155 and particle.cpp:187
            if (SSB_SC > 0) {
                const double base = G * wr;
#pragma unroll
#if SSB_REC2
                Qi[0] += Dk[0] * (Ci[0] - e.c) * base;
                if (SSB_SC > 1) Qi[SSB_SC > 1 ? 1 : 0] += Dk[SSB_SC > 1 ? 1 : 0] * (Ci[SSB_SC > 1 ? 1 : 0] - e.d) * base;
#else
#pragma unroll
                for (int s = 0; s < SSB_SC; s++) Qi[s] += Dk[s] * (Ci[s] - Cg[(size_t) s * N + j]) * base;
#endif
            }
            if (SSB_SD > 0) {
                const double hr = h - r;
                const double Dij = G * (ih7c * hr * hr);                     // particle.cpp:182-187 (sign folded)
                const int tj = (int) ((__double_as_longlong(c3.d) >> 32) & 0xffff) - 1;
#pragma unroll
                for (int s = 0; s < SSB_SD; s++) Dd[s] += dmat[s * num_types + tj] * Dij;
            }
        }
    }

    // results of particle i; returns max_s Ddiag_i[s] (window controller)
    __device__ __forceinline__ double store(const SsbView &V, int i, unsigned step) {
        V.F[0][i] = F0; V.F[1][i] = F1; V.F[2][i] = F2;
        V.Fbp[0][i] = B0; V.Fbp[1][i] = B1; V.Fbp[2][i] = B2;
        V.Frho[i] = Frho;
        if (SSB_SC > 0) {
            if (SSB_RC > 0) {                                                    // model.cpp:181-189
                const double vol = m_i / rho_i;
                const double cur_time = step * V.dt;
                double df[SSB_NDF > 0 ? SSB_NDF : 1];
#pragma unroll
                for (int qd = 0; qd < SSB_NDF; qd++) df[qd] = V.data_fn[(size_t) qd * N + i];
                double flux[SSB_RC > 0 ? SSB_RC : 1];
                ssb_gen::eval_det(Ci, cur_time, vol, df, type_i, flux);
#pragma unroll
                for (int rxn = 0; rxn < SSB_RC; rxn++) {
#pragma unroll
                    for (int s = 0; s < SSB_SC; s++) {
                        int nval;
                        if (V.flags & 2u) nval = ssb_gen::N_dense(s * SSB_R + rxn);
                        else { int kk = SSB_RC * rxn + s; nval = (kk < SSB_S * SSB_R) ? ssb_gen::N_dense(kk) : 0; }
                        Qi[s] += nval * flux[rxn];
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < SSB_SC; s++) V.Q[(size_t) s * N + i] = Qi[s];
        }
        double mx = 0.0;
#pragma unroll
        for (int s = 0; s < SSB_SD; s++) { V.Ddiag[(size_t) s * N + i] = Dd[s]; mx = fmax(mx, Dd[s]); }
        return mx;
    }
};

#ifndef SSB_FORCE_MINB
#define SSB_FORCE_MINB 4          // resident CTAs per SM the register budget is held to (4 x 128 threads: 128 registers)
#endif
__global__ void __launch_bounds__(SSB_BLOCK, SSB_FORCE_MINB) k_force_mv(SsbView V, unsigned step, unsigned long long *max_ddiag_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double mx = 0.0;
    if (i < V.N) {
        ForceSweep S;
        S.init(V, i);
        const int N = V.N;
        const int cnt = V.owned[i] ? V.nbr_count[i] : 0;     // ghost copies receive F, Fbp, Frho, Q from their owner (halo exchange)
#pragma unroll 2
        for (int k = 0; k < cnt; k++) {
            const int j = V.nbr[(size_t) k * N + i];
            const double *rj = V.rec + (size_t) j * 16;
            const ssb_d4 c0 = ssb_ld256(rj), c1 = ssb_ld256(rj + 4), c2 = ssb_ld256(rj + 8), c3 = ssb_ld256(rj + 12);
#if SSB_REC2
            const ssb_d4 e = ssb_ld256(V.rec2 + (size_t) j * 4);
#else
            const ssb_d4 e = c3;
#endif
            S.pair(c0, c1, c2, c3, e, j);
        }
        mx = S.store(V, i, step);
    }
    if (SSB_SD > 0) {
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(max_ddiag_bits, (unsigned long long) __double_as_longlong(mx));
    }
}

// ---------------------------------------------------------------------------------------------
// Static-domain fast path (positions, masses and densities never change: simulate.cpp:68,137 skip the integrator).
//   k_static_coef  once: the chemistry pair coefficient dQc_base of model.cpp:155 per ELL entry.
//   k_static_step  one launch per engine step, fusing  [compute_forces: Q = sweep(C)]  [take_step2: C += dt/2 Q; BC]
//                  with the NEXT step's [take_step1: C += dt/2 Q; BC; Q = 0]  — legal because the RDME that runs in between
//                  never touches C.  Reads the half-stepped concentrations of all particles from Cpre[in], writes the
//                  state after this step to C (what outputs and taps see), Q, and the next half-step to Cpre[in^1].
// Per particle-step the kernel streams 12 B per neighbour (index + coefficient) and gathers C_j from L2.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SSB_BLOCK) k_static_coef(SsbView V) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V.N) return;
    const int N = V.N, dim = V.dim;
    const double h = V.h;
    const double alpha = ssb_alpha(dim, h);
    const double xi0 = V.x[0][i], xi1 = V.x[1][i], xi2 = V.x[2][i];
    const double rho_i = V.rho[i], m_i = V.mass[i];
    const int cnt = V.nbr_count[i];
    for (int k = 0; k < cnt; k++) {
        const int j = V.nbr[(size_t) k * N + i];
        const double d2 = ssb_dist2(dim, xi0, xi1, xi2, V.x0[0][j], V.x0[1][j], V.x0[2][j]);
        const double r = sqrt(d2);
        const double dWdr = ssb_dWdr(alpha, r, h);
        const double rho_j = V.rho[j], m_j = V.mass[j];
        const double inv_reg = 1.0 / (r + 0.001 * h);
        const double wfd = inv_reg * dWdr;
        V.coef[(size_t) k * N + i] = 2.0 * ((m_i * m_j) / (m_i + m_j)) * ((rho_i + rho_j) / (rho_i * rho_j)) * (r * r) * wfd / ((r * r) + 0.01 * h * h);
    }
}

// Cpre is voxel-major ([j*S_c + s]) so that a neighbour's concentrations arrive with ONE gather (one 16-byte load for S_c = 2)
__global__ void __launch_bounds__(SSB_BLOCK) k_static_seed(SsbView V, int buf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V.N) return;
#pragma unroll
    for (int s = 0; s < SSB_SC; s++) V.Cpre[buf][(size_t) i * SSB_SC + s] = V.C[(size_t) s * V.N + i];
}

__device__ __forceinline__ void load_cvec(const double *base, size_t j, double *out) {
    if (SSB_SC == 2) {
        const double2 v = *reinterpret_cast<const double2 *>(base + j * 2);
        out[0] = v.x; out[SSB_SC > 1 ? 1 : 0] = v.y;
    } else {
#pragma unroll
        for (int s = 0; s < SSB_SC; s++) out[s] = base[j * SSB_SC + s];
    }
}

__global__ void __launch_bounds__(SSB_BLOCK) k_static_step(SsbView V, unsigned step, int in_buf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V.N) return;
    const int N = V.N;
    const double dt = V.dt;
    const double *Cin = V.Cpre[in_buf];
    double *Cnext = V.Cpre[in_buf ^ 1];
    const int type_i = V.type[i];
    double Ci[SSB_SC > 0 ? SSB_SC : 1], Qi[SSB_SC > 0 ? SSB_SC : 1], Dk[SSB_SC > 0 ? SSB_SC : 1];
    load_cvec(Cin, (size_t) i, Ci);
#pragma unroll
    for (int s = 0; s < SSB_SC; s++) {
        Qi[s] = 0.0;
        int k = ssb_pde_dindex(V, SSB_SC, type_i, s);                                  // model.cpp:163 (mirrored index)
        Dk[s] = (k >= 0 && k < SSB_S * V.num_types) ? V.dmat[k] : 0.0;
    }
    const int cnt = V.nbr_count[i];
    int k = 0;
    // four neighbours per trip: the index/coefficient streams and the four gathers are issued before any arithmetic, so
    // every thread keeps 12 loads in flight (the sweep is latency bound otherwise); accumulation order is unchanged
    for (; k + 4 <= cnt; k += 4) {
        int jq[4];
        double cq[4], Cj[4][SSB_SC > 0 ? SSB_SC : 1];
#pragma unroll
        for (int q = 0; q < 4; q++) { jq[q] = V.nbr[(size_t) (k + q) * N + i]; cq[q] = V.coef[(size_t) (k + q) * N + i]; }
#pragma unroll
        for (int q = 0; q < 4; q++) load_cvec(Cin, (size_t) jq[q], Cj[q]);
#pragma unroll
        for (int q = 0; q < 4; q++) {
#pragma unroll
            for (int s = 0; s < SSB_SC; s++) Qi[s] += Dk[s] * (Ci[s] - Cj[q][s]) * cq[q];
        }
    }
    for (; k < cnt; k++) {
        const int j = V.nbr[(size_t) k * N + i];
        const double cf = V.coef[(size_t) k * N + i];
        double Cj[SSB_SC > 0 ? SSB_SC : 1];
        load_cvec(Cin, (size_t) j, Cj);
#pragma unroll
        for (int s = 0; s < SSB_SC; s++) Qi[s] += Dk[s] * (Ci[s] - Cj[s]) * cf;
    }
    if (SSB_RC > 0) {                                                       // model.cpp:181-189
        const double vol = V.mass[i] / V.rho[i];
        const double cur_time = step * dt;
        double df[SSB_NDF > 0 ? SSB_NDF : 1];
#pragma unroll
        for (int q = 0; q < SSB_NDF; q++) df[q] = V.data_fn[(size_t) q * N + i];
        double flux[SSB_RC > 0 ? SSB_RC : 1];
        ssb_gen::eval_det(Ci, cur_time, vol, df, type_i, flux);
#pragma unroll
        for (int rxn = 0; rxn < SSB_RC; rxn++) {
#pragma unroll
            for (int s = 0; s < SSB_SC; s++) {
                int nval;
                if (V.flags & 2u) nval = ssb_gen::N_dense(s * SSB_R + rxn);
                else { int kk = SSB_RC * rxn + s; nval = (kk < SSB_S * SSB_R) ? ssb_gen::N_dense(kk) : 0; }
                Qi[s] += nval * flux[rxn];
            }
        }
    }
    Particle p;
#if SSB_HAS_BC
    bc_load(V, i, p);
    System sys = make_system(V, step);
#endif
    // take_step2 of this step: C += dt/2 Q; BC      (simulate.cpp:167-171)
#pragma unroll
    for (int s = 0; s < SSB_SC; s++) p.C[s] = Ci[s] + Qi[s] * dt * 0.5;
#if SSB_HAS_BC
    ssb_gen::applyBoundaryConditions(&p, &sys);
#endif
#pragma unroll
    for (int s = 0; s < SSB_SC; s++) { V.C[(size_t) s * N + i] = p.C[s]; V.Q[(size_t) s * N + i] = Qi[s]; }
    // take_step1 of the next step: C += dt/2 Q; BC   (simulate.cpp:81-88)
#pragma unroll
    for (int s = 0; s < SSB_SC; s++) p.C[s] += Qi[s] * dt * 0.5;
#if SSB_HAS_BC
    sys.current_step = step + 1;
    ssb_gen::applyBoundaryConditions(&p, &sys);
#endif
#pragma unroll
    for (int s = 0; s < SSB_SC; s++) Cnext[(size_t) i * SSB_SC + s] = p.C[s];
}

// ---------------------------------------------------------------------------------------------
// K4  take_step2 part 1 (E/src/simulate.cpp:136-155): corrector + Shepard filter (model.cpp:194-233).
// Writes the post-corrector density to rho_new so that the BVF sweep can serve `rho_j` with the
// reference's serial Gauss–Seidel visibility (model.cpp:285-293).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SSB_BLOCK) k_corrector(SsbView V, unsigned step) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V.N) return;
    const int N = V.N, dim = V.dim;
    const double h = V.h, dt = V.dt;
    const int solid = V.solid[i];
    double rho = V.rho[i];
    if (solid == 0) {
#pragma unroll
        for (int d = 0; d < 3; d++) V.v[d][i] = V.v[d][i] + 0.5 * dt * V.F[d][i];      // simulate.cpp:139-141
    }
    if (step % 20 == 0 && V.owned[i]) {                 // filterDensity (model.cpp:194-233); ghosts receive rho_new from their owner
        const double alpha = ssb_alpha(dim, h);
        const double xi0 = V.x[0][i], xi1 = V.x[1][i], xi2 = V.x[2][i];
        double num = 0.0, den = 0.0;
        const int cnt = V.nbr_count[i];
        for (int k = 0; k < cnt; k++) {
            const int j = V.nbr[(size_t) k * N + i];
            const double d2 = ssb_dist2(dim, xi0, xi1, xi2, V.x0[0][j], V.x0[1][j], V.x0[2][j]);
            if (V.filter && !ssb_in_range(d2, h, __dmul_rn(h, h))) continue;
            const double r = sqrt(d2);
            const double Wij = ssb_W(alpha, r, h);
            num += V.old_rho[j] * Wij;
            den += Wij;
        }
        rho = num / den;
    }
    if (solid == 0) rho = rho + 0.5 * dt * V.Frho[i];   // simulate.cpp:147
    // effect of the trailing applyBoundaryConditions (simulate.cpp:171) on rho, so later particles see it
    Particle p;
    bc_load(V, i, p);
    p.rho = rho;
    System sys = make_system(V, step);
    ssb_gen::applyBoundaryConditions(&p, &sys);
    V.rho_new[i] = p.rho;
}

// ---------------------------------------------------------------------------------------------
// K5  take_step2 part 2 (E/src/simulate.cpp:158-171): boundary volume fraction (model.cpp:236-313),
// bounce-back (model.cpp:316-329), chemistry half step, boundary conditions.
// MOVING = false on static domains: only the C half step and the BCs remain.
// ---------------------------------------------------------------------------------------------
template <bool MOVING>
__global__ void __launch_bounds__(SSB_BLOCK) k_finish(SsbView V, unsigned step) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V.N) return;
    const int N = V.N, dim = V.dim;
    const double h = V.h, dt = V.dt;
    Particle p;
    bc_load(V, i, p);
    if (MOVING) p.rho = V.rho_new[i];
    if (MOVING && p.solidTag == 0 && V.owned[i]) {      // ghosts receive v and bvf_phi from their owner
        const double alpha = ssb_alpha(dim, h);
        const int my_id = V.gid[i];
        const double inv_h = 1.0 / h;
        double nw[3] = {0.0, 0.0, 0.0};
        double vos = 0.0, vtot = 0.0;
        const int cnt = V.nbr_count[i];
        const bool use_rec = (V.rec != nullptr) && !(V.flags & 16u /*SSB_FLAG_LITERAL_KERNELS*/);
        // A fluid particle with no solid particle among its candidates has vos = 0 and a zero wall normal whatever the rest of the
        // sum is: bvf_phi = |0 / vtot| only needs to know whether vtot > 0, i.e. the sweep may stop at the first neighbour that
        // contributes.  The list build recorded that fact (V.solid_nbr, k_search); in a tank ~90 % of the fluid is bulk.
        const bool bulk = V.solid_nbr != nullptr && V.solid_nbr[i] == 0;
        for (int k = 0; k < cnt; k++) {
            if (bulk && vtot > 0.0) break;
            const int j = V.nbr[(size_t) k * N + i];
            double x0j0, x0j1, x0j2, xj0 = 0.0, xj1 = 0.0, xj2 = 0.0, m_j, rho_pre_j;
            int id_j, solid_j;
            if (use_rec) {      // one 128-byte gather record per neighbour: sectors 0 and 3, plus sector 1 (live x_j) for solid neighbours
                const double *rj = V.rec + (size_t) j * 16;
                const ssb_d4 c0 = ssb_ld256(rj), c3 = ssb_ld256(rj + 12);
                x0j0 = c0.a; x0j1 = c0.b; x0j2 = c0.c; xj0 = c0.d;
                rho_pre_j = c3.a; m_j = c3.b;
                const long long bits = __double_as_longlong(c3.d);
                id_j = (int) (bits & 0xffffffffll);                       // global particle id
                solid_j = (int) ((bits >> 48) & 0xffff);
                if (solid_j) { const ssb_d4 c1 = ssb_ld256(rj + 4); xj1 = c1.a; xj2 = c1.b; }
            } else {
                x0j0 = V.x0[0][j]; x0j1 = V.x0[1][j]; x0j2 = V.x0[2][j];
                xj0 = V.x[0][j]; xj1 = V.x[1][j]; xj2 = V.x[2][j];
                rho_pre_j = V.rho[j]; m_j = V.mass[j]; id_j = V.gid[j]; solid_j = V.solid[j];
            }
            const double d2 = ssb_dist2(dim, p.x[0], p.x[1], p.x[2], x0j0, x0j1, x0j2);
            if (V.filter && !ssb_in_range(d2, h, __dmul_rn(h, h))) continue;
            const double r = sqrt(d2);
            const double R = use_rec ? r * inv_h : r / h;
            const double q1 = 1 - R;
            const double Wij = alpha * ((1 + 3 * R) * (q1 * q1 * q1));         // model.cpp:275
            // serial (-t 1) visibility: particles earlier in the vector already ran their corrector
            const double rho_j = (id_j <= my_id) ? V.rho_new[j] : rho_pre_j;
            const double volj = m_j / rho_j;
            const double w2 = volj * volj * Wij;
            if (solid_j) vos += w2;
            vtot += w2;
            if (solid_j) {
                const double dWdr = ssb_dWdr(alpha, r, h);
                const double f = volj * volj * dWdr / (r + 0.001 * h);
                nw[0] += f * (p.x[0] - xj0);
                if (dim > 1) nw[1] += f * (p.x[1] - xj1);
                if (dim > 2) nw[2] += f * (p.x[2] - xj2);
            }
        }
#pragma unroll
        for (int d = 0; d < 3; d++) nw[d] = nw[d] / vtot;
        const double norm_nw = sqrt(nw[0] * nw[0] + nw[1] * nw[1] + nw[2] * nw[2]);
        double normal[3];
#pragma unroll
        for (int d = 0; d < 3; d++) normal[d] = -nw[d] / norm_nw;
        const double bvf = fabs(vos / vtot);            // fluid particle (model.cpp:309-312)
        V.bvf[i] = bvf;
#pragma unroll
        for (int d = 0; d < 3; d++) if (d < dim) V.vt[d][i] = 0.0;              // model.cpp:257-260
        // applyBoundaryVolumeFraction (model.cpp:316-329)
        const double vdn = p.v[0] * normal[0] + p.v[1] * normal[1] + p.v[2] * normal[2];
        if (bvf >= 0.5) {
#pragma unroll
            for (int d = 0; d < 3; d++) if (d < dim) p.v[d] = -p.v[d] + 2.0 * fmax(0.0, vdn) * normal[d];
        }
    }
#pragma unroll
    for (int s = 0; s < SSB_SC; s++) p.C[s] += V.Q[(size_t) s * N + i] * dt * 0.5;   // simulate.cpp:167-169
    System sys = make_system(V, step);
    ssb_gen::applyBoundaryConditions(&p, &sys);         // simulate.cpp:171
#pragma unroll
    for (int d = 0; d < 3; d++) V.v[d][i] = p.v[d];
    V.nu[i] = p.nu;
    if (!MOVING) V.rho[i] = p.rho;   // moving: rho_new already carries the BC effect (k_corrector)
#pragma unroll
    for (int s = 0; s < SSB_SC; s++) V.C[(size_t) s * N + i] = p.C[s];
}

// ---------------------------------------------------------------------------------------------
// K6  diffusion-matrix assembly: D_i_j (particle.cpp:182-187) and Ddiag / sdrate
// (E/src/simulate_rdme.cpp:131-152).  Caches D_i_j on static domains (V.Dij != nullptr).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double pair_Dij(const SsbView &V, int i, int j, double xi0, double xi1, double xi2,
                                           double m_i, double rho_i) {
    const double d2 = ssb_dist2(V.dim, xi0, xi1, xi2, V.x0[0][j], V.x0[1][j], V.x0[2][j]);
    const double r = sqrt(d2);
    double rho_j;
    ssb_search_rho(V, i, j, rho_i, rho_j);
    return ssb_Dij(d2, r, V.h, m_i, V.mass[j], rho_i, rho_j);
}

__global__ void __launch_bounds__(SSB_BLOCK) k_diff_init(SsbView V, unsigned long long *max_ddiag_bits) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double mx = 0.0;
    if (i < V.N) {
        const int N = V.N;
        const double xi0 = V.x[0][i], xi1 = V.x[1][i], xi2 = V.x[2][i];
        const double m_i = V.mass[i], rho_i = V.rho_search[i];
        double Dd[SSB_SD > 0 ? SSB_SD : 1];
#pragma unroll
        for (int s = 0; s < SSB_SD; s++) Dd[s] = 0.0;
        const int cnt = V.nbr_count[i];
        for (int k = 0; k < cnt; k++) {
            const int j = V.nbr[(size_t) k * N + i];
            if (V.filter && !ssb_in_range(ssb_dist2(V.dim, xi0, xi1, xi2, V.x0[0][j], V.x0[1][j], V.x0[2][j]), V.h, __dmul_rn(V.h, V.h))) continue;
            const double Dij = pair_Dij(V, i, j, xi0, xi1, xi2, m_i, rho_i);
            if (V.Dij) V.Dij[(size_t) k * N + i] = Dij;
            const int tj = V.type[j] - 1;
#pragma unroll
            for (int s = 0; s < SSB_SD; s++) Dd[s] += V.dmat[s * V.num_types + tj] * Dij;   // simulate_rdme.cpp:146-147
        }
#pragma unroll
        for (int s = 0; s < SSB_SD; s++) { V.Ddiag[(size_t) s * N + i] = Dd[s]; mx = fmax(mx, Dd[s]); }
    }
    // block max -> global max (positive doubles order like their bit patterns)
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(max_ddiag_bits, (unsigned long long) __double_as_longlong(mx));
}

// ---------------------------------------------------------------------------------------------
// K7a  RDME (re)initialisation: reaction propensities (simulate_rdme.cpp:113-128), sdrate (:149),
// first event time Exp(1)/(srrate+sdrate)+t0 (simulate_rdme.cpp:155-195, NRMConstant_v5.cpp:52-59).
// ---------------------------------------------------------------------------------------------
// minimum over the block (all threads must call); result valid in thread 0
__device__ __forceinline__ double block_min(double v) {
    __shared__ double sh_min[SSB_BLOCK / 32];
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sh_min[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double w = (threadIdx.x < SSB_BLOCK / 32) ? sh_min[threadIdx.x] : INFINITY;
        for (int o = 16; o > 0; o >>= 1) w = fmin(w, __shfl_xor_sync(0xffffffffu, w, o));
        v = w;
    }
    return v;
}

struct VoxelRates {
    double rr[SSB_RD > 0 ? SSB_RD : 1];
    double sr, sd;
};

// lag compensation of the windowed scheme: a molecule that jumps inside a window only becomes mobile again at the
// window end (mean lag tau/2), so its jump cadence would be 1/d + tau/2; using d' = d / (1 - d tau/2) restores 1/d.
__device__ __forceinline__ double lag_comp(double d, double tau) {
    const double q = fmin(0.5 * d * tau, 0.5);
    return d / (1.0 - q);
}

// xr = molecules that can REACT here (present + departing), xx = molecules that can still JUMP (present)
__device__ __forceinline__ void eval_rates(const SsbView &V, int i, const int *xr, const int *xx, double t, double vol,
                                           const double *df, int type, double tau, VoxelRates &R) {
    ssb_gen::eval_propensities(xr, t, vol, df, type, R.rr);
    double sr = 0.0;
#pragma unroll
    for (int r = 0; r < SSB_RD; r++) sr += R.rr[r];
    double sd = 0.0;
#pragma unroll
    for (int s = 0; s < SSB_SD; s++) sd += lag_comp(V.Ddiag[(size_t) s * V.N + i], tau) * xx[s];
    R.sr = sr; R.sd = sd;
}

__global__ void __launch_bounds__(SSB_BLOCK) k_rdme_init(SsbView V, double t0, double t_eval, double tau, uint64_t seed, uint64_t epoch) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double tn = INFINITY;
    if (i < V.N) {
        const int N = V.N;
        int xx[SSB_SD > 0 ? SSB_SD : 1];
#pragma unroll
        for (int s = 0; s < SSB_SD; s++) xx[s] = (int) V.xx[(size_t) s * N + i];
        double df[SSB_NDF > 0 ? SSB_NDF : 1];
#pragma unroll
        for (int q = 0; q < SSB_NDF; q++) df[q] = V.data_fn[(size_t) q * N + i];
        const double vol = V.mass[i] / V.rho[i];
        VoxelRates R;
        eval_rates(V, i, xx, xx, t_eval, vol, df, V.type[i], tau, R);
#pragma unroll
        for (int r = 0; r < SSB_RD; r++) V.rrate[(size_t) r * N + i] = R.rr[r];
        V.srrate[i] = R.sr;
        V.sdrate[i] = R.sd;
        const double tot = R.sr + R.sd;
        double u0, u1;
        philox_uniform2((uint32_t) V.gid[i], 0u, epoch, seed, u0, u1);
        tn = (tot > 0.0 && V.owned[i]) ? t0 + (-log(u0)) / tot : INFINITY;      // ghost voxels are simulated by their owner
        if ((V.flags & 32u) && V.owned[i]) {      // leap form: the stored clock is the REACTION clock; mobile molecules make a voxel due every window
            tn = (R.sd > 0.0) ? -INFINITY : ((R.sr > 0.0) ? t0 + (-log(u0)) / R.sr : INFINITY);
        }
        V.tnext[i] = tn;
#pragma unroll
        for (int s = 0; s < SSB_SD; s++) { V.inbox[0][(size_t) s * N + i] = 0u; V.inbox[1][(size_t) s * N + i] = 0u; }
        V.inbox_src[0][i] = 0ull; V.inbox_src[1][i] = 0ull;
    }
    tn = block_min(tn);
    if (threadIdx.x == 0) { V.blk_tmin[blockIdx.x] = tn; V.blk_mail[0][blockIdx.x] = 0; V.blk_mail[1][blockIdx.x] = 0; }
}

// ---------------------------------------------------------------------------------------------
// K7b  one sSSA window [t_lo, t_hi] — one thread per voxel.
// The reference's NSM (simulate_rdme.cpp:211-472) executes events in global time order, which is inherently
// serial.  Here every voxel runs an exact SSA of ITS OWN channels (reactions + outgoing diffusion jumps) inside
// the window.  A molecule that jumps is posted to the destination's inbox (integer atomicAdd: order-independent, so
// results are bit-reproducible) and is picked up by the destination at the start of the next window, when the
// destination re-draws its next event time (memoryless, exactly what NRMConstant_v5::update does:
// Exp(1)/a + t, NRMConstant_v5.cpp:98).  Until the window closes the jumped molecule stays REACTIVE in its source
// voxel (`dep`), so no molecule is ever invisible to the reaction channels, and the jump propensities carry the
// lag compensation above; what remains is an O((tau * jump rate)^2) splitting error bounded by the window
// controller (ssb_model.rdme_epsilon = tau * max jump rate).
// Event execution mirrors the reference branch by branch:
//   channel draw  rand1 <= srrate/totrate ? reaction : diffusion                     (simulate_rdme.cpp:253-255)
//   reaction pick rand1*srrate against the running sum of rrate[]                     (:260-261)  [reference rule]
//   species pick  rand1*sdrate against the running sum of Ddiag[s]*xx[s]              (:317-321)  [reference rule]
//   direction     r2*Ddiag[spec] against the running sum of D_i_j*D[spec,type(dest)]  (:353-367)
// SSB_FLAG_CORRECTED_NSM_SELECT switches the two picks to the textbook rule (rand*totrate, subtract srrate).
// ---------------------------------------------------------------------------------------------
// warp-cooperative direction pick (simulate_rdme.cpp:353-380) for the voxel `il` of one lane: the 32 lanes evaluate 32
// neighbours' weights D_i_j * D[spec, type(dest)] at once, a shuffle prefix sum forms the running sums, and the first
// neighbour whose running sum exceeds `target` is the destination.  Returns -1 when no neighbour accepts the species.
__device__ __forceinline__ int coop_pick_direction(const SsbView &V, int il, int spec, double target) {
    const int N = V.N, lane = threadIdx.x & 31;
    const int cnt = V.nbr_count[il];
    const bool cached = V.Dij != nullptr;
    const bool use_rec = !cached && V.rec != nullptr && !V.static_domain && !(V.flags & 16u /*SSB_FLAG_LITERAL_KERNELS*/);
    double xl0 = 0, xl1 = 0, xl2 = 0, m_l = 0, rho_l = 0;
    if (!cached) { xl0 = V.x[0][il]; xl1 = V.x[1][il]; xl2 = V.x[2][il]; m_l = V.mass[il]; rho_l = V.rho_search[il]; }
    double base = 0.0;
    int dest = -1, last_ok = -1;
    for (int k0 = 0; k0 < cnt; k0 += 32) {
        const int k = k0 + lane;
        int j = -1;
        double w = 0.0;
        bool ok = false;
        if (k < cnt) {
            j = V.nbr[(size_t) k * N + il];
            if (use_rec) {
                // moving domains: the search-time snapshot of j (x0, rho, mass, type) sits in two sectors of its gather record
                // instead of seven separate arrays — the pick is bound by scattered sectors, not by arithmetic
                const double *rj = V.rec + (size_t) j * 16;
                const ssb_d4 c0 = ssb_ld256(rj), c3 = ssb_ld256(rj + 12);
                const int tj = (int) ((__double_as_longlong(c3.d) >> 32) & 0xffff) - 1;
                const double dc = V.dmat[spec * V.num_types + tj];
                const double d2 = ssb_dist2(V.dim, xl0, xl1, xl2, c0.a, c0.b, c0.c);
                const bool in = !V.filter || ssb_in_range(d2, V.h, __dmul_rn(V.h, V.h));
                if (dc != 0.0 && in) {
                    ok = true;
                    w = ssb_Dij(d2, sqrt(d2), V.h, m_l, c3.b, rho_l, c3.a) * dc;
                }
            } else {
            const double dc = V.dmat[spec * V.num_types + (V.type[j] - 1)];
            bool in = true;
            if (V.filter) in = ssb_in_range(ssb_dist2(V.dim, xl0, xl1, xl2, V.x0[0][j], V.x0[1][j], V.x0[2][j]), V.h, __dmul_rn(V.h, V.h));
            if (dc != 0.0 && in) {
                ok = true;
                const double Dij = cached ? V.Dij[(size_t) k * N + il] : pair_Dij(V, il, j, xl0, xl1, xl2, m_l, rho_l);
                w = Dij * dc;
            }
            }
        }
        double incl = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        incl += base;
        const unsigned hit = __ballot_sync(0xffffffffu, ok && incl > target);
        const unsigned oks = __ballot_sync(0xffffffffu, ok);
        if (hit) { dest = __shfl_sync(0xffffffffu, j, __ffs(hit) - 1); break; }
        if (oks) last_ok = __shfl_sync(0xffffffffu, j, 31 - __clz(oks));
        base = __shfl_sync(0xffffffffu, incl, 31);
    }
    return dest >= 0 ? dest : last_ok;                                          // round-off overflow (:368-380)
}

// the same pick done by ONE lane over its own neighbour row (serial running sum).  Used when several lanes of a warp need a
// direction at once (crowded voxels): the lanes then work in parallel instead of queueing for the warp-cooperative pick.
__device__ __forceinline__ int serial_pick_direction(const SsbView &V, int il, int spec, double target) {
    const int N = V.N;
    const int cnt = V.nbr_count[il];
    const bool cached = V.Dij != nullptr;
    double xl0 = 0, xl1 = 0, xl2 = 0, m_l = 0, rho_l = 0;
    if (!cached) { xl0 = V.x[0][il]; xl1 = V.x[1][il]; xl2 = V.x[2][il]; m_l = V.mass[il]; rho_l = V.rho_search[il]; }
    const bool use_rec = !cached && V.rec != nullptr && !V.static_domain && !(V.flags & 16u);
    double cum = 0.0;
    int last_ok = -1;
    for (int k = 0; k < cnt; k++) {
        const int j = V.nbr[(size_t) k * N + il];
        if (use_rec) {
            const double *rj = V.rec + (size_t) j * 16;
            const ssb_d4 c0 = ssb_ld256(rj), c3 = ssb_ld256(rj + 12);
            const double dc = V.dmat[spec * V.num_types + ((int) ((__double_as_longlong(c3.d) >> 32) & 0xffff) - 1)];
            if (dc == 0.0) continue;
            const double d2 = ssb_dist2(V.dim, xl0, xl1, xl2, c0.a, c0.b, c0.c);
            if (V.filter && !ssb_in_range(d2, V.h, __dmul_rn(V.h, V.h))) continue;
            cum += ssb_Dij(d2, sqrt(d2), V.h, m_l, c3.b, rho_l, c3.a) * dc;
            last_ok = j;
            if (cum > target) return j;
            continue;
        }
        const double dc = V.dmat[spec * V.num_types + (V.type[j] - 1)];
        if (dc == 0.0) continue;
        if (V.filter && !ssb_in_range(ssb_dist2(V.dim, xl0, xl1, xl2, V.x0[0][j], V.x0[1][j], V.x0[2][j]), V.h, __dmul_rn(V.h, V.h))) continue;
        const double Dij = cached ? V.Dij[(size_t) k * N + il] : pair_Dij(V, il, j, xl0, xl1, xl2, m_l, rho_l);
        cum += Dij * dc;
        last_ok = j;
        if (cum > target) return j;
    }
    return last_ok;
}

__device__ __forceinline__ void rdme_window_body(const SsbView &V, double t_lo, double t_hi, double tau, uint64_t seed,
                                                 uint64_t epoch, int buf, unsigned &n_rx, unsigned &n_df) {
    // Persistent grid: each CTA owns the chunks c = blockIdx.x + t*gridDim.x (a chunk = SSB_BLOCK consecutive voxels).
    // Parallel triage first — thread t inspects chunk t's summary (earliest tnext, mail flag) — so a window in which
    // nothing is due costs two loads per chunk instead of a block dispatch per chunk; then only the active chunks run.
    // The order of the active list depends on atomics, the result does not (chunks are independent in a window).
    // Inside a chunk the event loop is warp-synchronous: every lane runs the SSA of its own voxel, and the one long
    // operation of an event — the scan of the neighbour row for the jump direction — is done by the whole warp for one
    // lane at a time (coop_pick_direction), so an event costs a few memory round trips instead of one per neighbour.
    // The active chunks are then cut into warp-sized slices that the CTA's warps take from a shared counter: a warp whose slice has
    // no event moves on at once instead of waiting at a block barrier for the one warp of the chunk that is executing events, so
    // the cost of a window is (events x event latency) / (resident warps), not the sum over chunks of their slowest warp.
    __shared__ int sh_act[SSB_BLOCK];
    __shared__ unsigned long long sh_tmin[SSB_BLOCK];    // earliest tnext per active chunk (bit pattern of a non-negative double)
    __shared__ int sh_nact, sh_next;
    constexpr int WPB = SSB_BLOCK / 32;
    const int N = V.N;
    const int nchunks = (N + SSB_BLOCK - 1) / SSB_BLOCK;
    const int lane = threadIdx.x & 31;
    for (int round0 = 0; blockIdx.x + (long long) round0 * gridDim.x < nchunks; round0 += SSB_BLOCK) {
    __syncthreads();
    if (threadIdx.x == 0) { sh_nact = 0; sh_next = 0; }
    __syncthreads();
    {
        const long long c = blockIdx.x + (long long) (round0 + threadIdx.x) * gridDim.x;
        if (c < nchunks) {
            const int mail = __ldcg(&V.blk_mail[buf ^ 1][c]);      // written by other CTAs: read through L2
            if (mail) V.blk_mail[buf ^ 1][c] = 0;
            if (mail != 0 || V.blk_tmin[c] <= t_hi) {
                const int slot = atomicAdd(&sh_nact, 1);
                sh_act[slot] = (int) c;
                sh_tmin[slot] = 0x7ff0000000000000ull;             // +inf
            }
        }
    }
    __syncthreads();
    const int nact = sh_nact;
    const int nitems = nact * WPB;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(&sh_next, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nitems) break;
        const int a = item / WPB;
        const int chunk = sh_act[a];
        const int i = chunk * SSB_BLOCK + (item % WPB) * 32 + lane;
        const bool valid = i < N;
        const int ii = valid ? i : N - 1;                          // clamp so idle lanes can run the same code path
        const unsigned *in_prev = V.inbox[buf ^ 1];
        unsigned *out_box = V.inbox[buf];
        const bool mine = valid && V.owned[ii];                    // mail addressed to ghost voxels stays in the inbox for the halo exchange
        double tnext = mine ? V.tnext[ii] : INFINITY;
        bool arrived = false;
        unsigned inc[SSB_SD > 0 ? SSB_SD : 1];
#pragma unroll
        for (int s = 0; s < SSB_SD; s++) { inc[s] = mine ? __ldcg(&in_prev[(size_t) s * N + ii]) : 0u; arrived |= (inc[s] != 0u); }
        const bool touched = mine && (arrived || tnext <= t_hi);
        int xx[SSB_SD > 0 ? SSB_SD : 1];     // present: may react and jump
        int xr[SSB_SD > 0 ? SSB_SD : 1];     // present + departing: may react
        double Dd[SSB_SD > 0 ? SSB_SD : 1];  // lag-compensated jump propensity per molecule
        double df[SSB_NDF > 0 ? SSB_NDF : 1];
        VoxelRates R;
        R.sr = 0.0; R.sd = 0.0;
        int type_i = 1;
        double vol = 1.0;
        uint32_t vid = 0, draw = 0;
        if (touched) {
#pragma unroll
            for (int s = 0; s < SSB_SD; s++) {
                xx[s] = (int) V.xx[(size_t) s * N + i];
                Dd[s] = lag_comp(V.Ddiag[(size_t) s * N + i], tau);
            }
#pragma unroll
            for (int q = 0; q < SSB_NDF; q++) df[q] = V.data_fn[(size_t) q * N + i];
            type_i = V.type[i];
            vol = V.mass[i] / V.rho[i];
            vid = (uint32_t) V.gid[i];
            if (arrived) {
                // stored propensities; only the reactions that depend on an arrived species are re-evaluated
                // (dependency graph columns [0,S), simulate_rdme.cpp:419-435)
#pragma unroll
                for (int r = 0; r < SSB_RD; r++) R.rr[r] = V.rrate[(size_t) r * N + i];
                unsigned long long mask = 0ull;
#pragma unroll
                for (int s = 0; s < SSB_SD; s++) {
                    if (inc[s]) {
                        xx[s] += (int) inc[s];
                        ((unsigned *) in_prev)[(size_t) s * N + i] = 0u;     // nobody writes buf^1 during this window
                        mask |= ssb_gen::dep_mask_species(s);
                    }
                }
                // reference quirk: the destination's propensities are re-evaluated with the SOURCE voxel's vol
                // (simulate_rdme.cpp:433) and stay that way until its next own event.
                double vol_dest = vol;
                const int src = (int) (__ldcg(&V.inbox_src[buf ^ 1][i]) & 0xffffffffull);
                if (src > 0 && !(V.flags & 1u)) { vol_dest = V.mass[src - 1] / V.rho[src - 1]; }
                V.inbox_src[buf ^ 1][i] = 0ull;
                double tmp[SSB_RD > 0 ? SSB_RD : 1];
                ssb_gen::eval_propensities(xx, t_lo, vol_dest, df, type_i, tmp);
                double sr = 0.0, sd = 0.0;
#pragma unroll
                for (int r = 0; r < SSB_RD; r++) { if ((mask >> r) & 1ull) R.rr[r] = tmp[r]; sr += R.rr[r]; }
#pragma unroll
                for (int s = 0; s < SSB_SD; s++) sd += Dd[s] * xx[s];
                R.sr = sr; R.sd = sd;
                const double tot = R.sr + R.sd;
                double u0, u1;
                philox_uniform2(vid, draw++, epoch, seed, u0, u1);
                tnext = (tot > 0.0) ? t_lo + (-log(u0)) / tot : INFINITY;
            } else {
                R.sr = V.srrate[i]; R.sd = V.sdrate[i];
#pragma unroll
                for (int r = 0; r < SSB_RD; r++) R.rr[r] = V.rrate[(size_t) r * N + i];
            }
#pragma unroll
            for (int s = 0; s < SSB_SD; s++) xr[s] = xx[s];
        }
        int guard = 0;
        bool failed = false;
        // ---- warp-synchronous event loop ------------------------------------------------------------------------
        for (;;) {
            const bool ev = touched && !failed && tnext <= t_hi;
            if (!__any_sync(0xffffffffu, ev)) break;
            double tt = tnext, rand2 = 0.0, pick = 0.0;
            bool is_rxn = false;
            int spec = 0, ev_re = 0;
            if (ev) {
                const double tot = R.sr + R.sd;
                double rand1;
                philox_uniform2(vid, draw++, epoch, seed, rand1, rand2);
                if (V.flags & 1u /*SSB_FLAG_CORRECTED_NSM_SELECT*/) {
                    pick = rand1 * tot;
                    is_rxn = pick <= R.sr;
                    if (!is_rxn) pick -= R.sr;
                } else {
                    is_rxn = rand1 <= R.sr / tot;                               // simulate_rdme.cpp:253-255
                    pick = is_rxn ? rand1 * R.sr : rand1 * R.sd;                // :260, :317
                }
                if (!is_rxn) {
                    double cum = Dd[0] * xx[0];
#pragma unroll
                    for (int q = 1; q < SSB_SD; q++) {
                        if (pick > cum) { spec = q; cum += Dd[q] * xx[q]; } else break;
                    }
                    while (spec > 0 && xx[spec] <= 0) spec--;                   // simulate_rdme.cpp:338-347
                    if (xx[spec] <= 0) { atomicCAS(V.err_flag, 0, 2); failed = true; }
                }
            }
            // the warp serves, one lane at a time, every lane that needs a jump direction
            const bool need_dir = ev && !is_rxn && !failed;
            int dest = -1;
            unsigned todo = __ballot_sync(0xffffffffu, need_dir);
            if (__popc(todo) >= 4) {           // crowded: every lane scans its own row, all lanes in parallel
                if (need_dir) dest = serial_pick_direction(V, ii, spec, rand2 * V.Ddiag[(size_t) spec * N + ii]);
                todo = 0u;
            }
            while (todo) {
                const int leader = __ffs(todo) - 1;
                todo &= todo - 1;
                const int il = __shfl_sync(0xffffffffu, ii, leader);
                const int sl = __shfl_sync(0xffffffffu, spec, leader);
                const double dd_l = V.Ddiag[(size_t) sl * N + il];
                const double target = __shfl_sync(0xffffffffu, rand2, leader) * dd_l;     // simulate_rdme.cpp:353-354
                const int d = coop_pick_direction(V, il, sl, target);
                if (lane == leader) dest = d;
            }
            if (ev && !failed) {
                if (is_rxn) {
                    int re = 0;
                    double cum = R.rr[0];
#pragma unroll
                    for (int q = 1; q < SSB_RD; q++) { if (pick > cum) { re = q; cum += R.rr[q]; } else break; }
                    // fell off the end with a zero-propensity tail: step back to the last live reaction (:262-281)
                    while (re > 0 && R.rr[re] <= 0.0) re--;
                    ev_re = re;
                    // the reaction acts on the reactive population; consumption is taken from the present molecules.
                    int xn[SSB_SD > 0 ? SSB_SD : 1];
#pragma unroll
                    for (int s = 0; s < SSB_SD; s++) xn[s] = xr[s];
                    bool neg = false;
                    ssb_gen::apply_stoich(re, xn, neg);                         // simulate_rdme.cpp:285-296
                    if (neg) atomicCAS(V.err_flag, 0, 2 /*SSB_ERR_RDME*/);
                    bool feasible = true;
#pragma unroll
                    for (int s = 0; s < SSB_SD; s++) feasible &= (xx[s] + (xn[s] - xr[s]) >= 0);
                    if (feasible) {
#pragma unroll
                        for (int s = 0; s < SSB_SD; s++) { xx[s] += xn[s] - xr[s]; xr[s] = xn[s]; }
                        n_rx++;
                    }
                    // else: every present reactant already departed in this window — the event is dropped
                    // (second order in tau; keeps posted jumps final and the result deterministic)
                } else if (dest < 0) {
                    atomicCAS(V.err_flag, 0, 2);
                    failed = true;
                } else {
                    if (dest != i) {                                            // (dest == i: stale self-neighbour on moving domains)
                        // no longer mobile here, still reactive (xr) until the window closes
#pragma unroll
                        for (int s = 0; s < SSB_SD; s++) if (s == spec) xx[s]--;
                        atomicAdd(&out_box[(size_t) spec * N + dest], 1u);
                        {   // random priority from this voxel's Philox stream (bits of rand2 not used by the direction pick)
                            const unsigned long long pri = (unsigned long long) (__double_as_longlong(rand2 * 4294967296.0 * 4096.0) & 0xffffffffll);
                            atomicMax(&V.inbox_src[buf][dest], (pri << 32) | (unsigned long long) (unsigned) (i + 1));
                        }
                        V.blk_mail[buf][dest / SSB_BLOCK] = 1;
                    }
                    n_df++;
                }
                if (!failed) {
                    if (V.flags & 1u) {
                        eval_rates(V, i, xr, xx, tt, vol, df, type_i, tau, R);      // textbook mode: everything fresh, own vol
                    } else {
                        // reference: only the reactions the dependency graph lists for this event are re-evaluated
                        // (simulate_rdme.cpp:299-308 after a reaction, :419-437 after a jump); the others keep their stored
                        // value — including one computed with a neighbour's vol on arrival (:433)
                        const unsigned long long mask = is_rxn ? ssb_gen::dep_mask_reaction(ev_re) : ssb_gen::dep_mask_species(spec);
                        double tmp[SSB_RD > 0 ? SSB_RD : 1];
                        ssb_gen::eval_propensities(xr, tt, vol, df, type_i, tmp);
                        double sr = 0.0, sd = 0.0;
#pragma unroll
                        for (int r = 0; r < SSB_RD; r++) { if ((mask >> r) & 1ull) R.rr[r] = tmp[r]; sr += R.rr[r]; }
#pragma unroll
                        for (int s = 0; s < SSB_SD; s++) sd += Dd[s] * xx[s];
                        R.sr = sr; R.sd = sd;
                    }
                    const double tot2 = R.sr + R.sd;
                    double u0, u1;
                    philox_uniform2(vid, draw++, epoch, seed, u0, u1);
                    tnext = (tot2 > 0.0) ? tt + (-log(u0)) / tot2 : INFINITY;   // NRMConstant_v5.cpp:92-99
                    if (++guard > 100000000) { atomicCAS(V.err_flag, 0, 2); failed = true; }
                }
            }
        }
        if (touched) {
            // window closes: departed molecules leave the reactive population; rates for the next window
            bool departed = false;
#pragma unroll
            for (int s = 0; s < SSB_SD; s++) departed |= (xr[s] != xx[s]);
            if (departed) {
                eval_rates(V, i, xx, xx, t_hi, vol, df, type_i, tau, R);
                const double tot3 = R.sr + R.sd;
                double u0, u1;
                philox_uniform2(vid, draw++, epoch, seed, u0, u1);
                tnext = (tot3 > 0.0) ? t_hi + (-log(u0)) / tot3 : INFINITY;
            }
#pragma unroll
            for (int s = 0; s < SSB_SD; s++) V.xx[(size_t) s * N + i] = (unsigned) xx[s];
#pragma unroll
            for (int r = 0; r < SSB_RD; r++) V.rrate[(size_t) r * N + i] = R.rr[r];
            V.srrate[i] = R.sr;
            V.sdrate[i] = R.sd;
            V.tnext[i] = tnext;
        }
        double tn_final = valid ? tnext : INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tn_final = fmin(tn_final, __shfl_xor_sync(0xffffffffu, tn_final, o));
        if (lane == 0) atomicMin(&sh_tmin[a], (unsigned long long) __double_as_longlong(tn_final));
    }
    __syncthreads();
    if ((int) threadIdx.x < nact) V.blk_tmin[sh_act[threadIdx.x]] = __longlong_as_double((long long) sh_tmin[threadIdx.x]);
    }
}

// ---------------------------------------------------------------------------------------------
// K7b, leap form (SSB_FLAG_LEAP_DIFFUSION): same windows, same inboxes, same reaction SSA — but the DIFFUSION channel of a
// voxel is advanced per window instead of per jump: every molecule present at the window start jumps independently with
// probability q = 1 - exp(-lambda' tau), so the number of jumpers of a species is Binomial(n, q) and their destinations are
// multinomial over the neighbour weights D_i_j * D[spec, type(dest)].  In the windowed scheme a molecule jumps at most once per
// window anyway (it sits in the destination's inbox until the window closes), so for the diffusion channel this is the same
// law at O(1) work per species and window instead of O(jumps) — a voxel holding hundreds of molecules no longer serialises
// hundreds of events.  The reference's channel-pick distortion (rand1*sdrate on (srrate/totrate, 1], simulate_rdme.cpp:317-321)
// is carried over as the per-species effective propensity tot * |(p,1] ∩ (c_{s-1}/sd, c_s/sd]| evaluated at the window start.
// ---------------------------------------------------------------------------------------------
struct LeapRng {
    uint32_t vid, draw;
    uint64_t epoch, seed;
    double spare;
    bool has_spare;
    __device__ __forceinline__ double next() {
        if (has_spare) { has_spare = false; return spare; }
        double a, b;
        philox_uniform2(vid, draw++, epoch, seed, a, b);
        spare = b; has_spare = true;
        return a;
    }
};

// Binomial(n, p): inversion (BINV) for small means, BTRS (Hormann 1993) otherwise; both exact up to fp rounding.
__device__ int ssb_binomial(int n, double p, LeapRng &rng) {
    if (n <= 0 || !(p > 0.0)) return 0;
    if (p >= 1.0) return n;
    const bool flip = p > 0.5;
    const double pp = flip ? 1.0 - p : p;
    int x;
    if ((double) n * pp < 30.0 && (double) n * log1p(-pp) > -700.0) {
        const double q = 1.0 - pp, s = pp / q, a = (n + 1) * s;
        double r = exp((double) n * log(q)), u = rng.next();
        x = 0;
        while (u > r) {
            u -= r;
            x++;
            if (x > n) { x = n; break; }
            r *= (a / x - s);
            if (r <= 0.0) break;
        }
    } else {
        const double spq = sqrt((double) n * pp * (1.0 - pp));
        const double b = 1.15 + 2.53 * spq, a = -0.0873 + 0.0248 * b + 0.01 * pp, c = n * pp + 0.5;
        const double vr = 0.92 - 4.2 / b, alpha = (2.83 + 5.1 / b) * spq, lpq = log(pp / (1.0 - pp));
        const int m = (int) floor((n + 1) * pp);
        const double h = lgamma(m + 1.0) + lgamma(n - m + 1.0);
        for (;;) {
            double u = rng.next() - 0.5, v = rng.next();
            const double us = 0.5 - fabs(u);
            const double kd = floor((2.0 * a / us + b) * u + c);
            if (kd < 0.0 || kd > (double) n) continue;
            const int k = (int) kd;
            if (us >= 0.07 && v <= vr) { x = k; break; }
            v = log(v * alpha / (a / (us * us) + b));
            if (v <= h - lgamma(k + 1.0) - lgamma(n - k + 1.0) + (k - m) * lpq) { x = k; break; }
        }
    }
    return flip ? n - x : x;
}

__device__ __forceinline__ void rdme_window_body_leap(const SsbView &V, double t_lo, double t_hi, double tau, uint64_t seed,
                                                      uint64_t epoch, int buf, unsigned &n_rx, unsigned &n_df) {
    __shared__ int sh_act[SSB_BLOCK];
    __shared__ int sh_nact;
    const int N = V.N;
    const int nchunks = (N + SSB_BLOCK - 1) / SSB_BLOCK;
    const double tau_w = t_hi - t_lo;
    for (int round0 = 0; blockIdx.x + (long long) round0 * gridDim.x < nchunks; round0 += SSB_BLOCK) {
    __syncthreads();
    if (threadIdx.x == 0) sh_nact = 0;
    __syncthreads();
    {
        const long long c = blockIdx.x + (long long) (round0 + threadIdx.x) * gridDim.x;
        if (c < nchunks) {
            const int mail = __ldcg(&V.blk_mail[buf ^ 1][c]);
            if (mail) V.blk_mail[buf ^ 1][c] = 0;
            if (mail != 0 || V.blk_tmin[c] <= t_hi) sh_act[atomicAdd(&sh_nact, 1)] = (int) c;
        }
    }
    __syncthreads();
    const int nact = sh_nact;
    for (int a = 0; a < nact; a++) {
        const int chunk = sh_act[a];
        const int i = chunk * SSB_BLOCK + threadIdx.x;
        const bool valid = i < N;
        const int ii = valid ? i : N - 1;
        const unsigned *in_prev = V.inbox[buf ^ 1];
        unsigned *out_box = V.inbox[buf];
        const bool mine = valid && V.owned[ii];
        double tnext = mine ? V.tnext[ii] : INFINITY;      // stored: next REACTION time of a voxel that was idle (-INF: has mobile molecules)
        bool arrived = false;
        unsigned inc[SSB_SD > 0 ? SSB_SD : 1];
#pragma unroll
        for (int s = 0; s < SSB_SD; s++) { inc[s] = mine ? __ldcg(&in_prev[(size_t) s * N + ii]) : 0u; arrived |= (inc[s] != 0u); }
        const bool touched = mine && (arrived || tnext <= t_hi);
        double tn_store = tnext;
        if (touched) {
            int xx[SSB_SD > 0 ? SSB_SD : 1], xr[SSB_SD > 0 ? SSB_SD : 1];
            double Dd[SSB_SD > 0 ? SSB_SD : 1], df[SSB_NDF > 0 ? SSB_NDF : 1];
#pragma unroll
            for (int s = 0; s < SSB_SD; s++) { xx[s] = (int) V.xx[(size_t) s * N + i]; Dd[s] = lag_comp(V.Ddiag[(size_t) s * N + i], tau); }
#pragma unroll
            for (int q = 0; q < SSB_NDF; q++) df[q] = V.data_fn[(size_t) q * N + i];
            const int type_i = V.type[i];
            const double vol = V.mass[i] / V.rho[i];
            LeapRng rng;
            rng.vid = (uint32_t) V.gid[i]; rng.draw = 0; rng.epoch = epoch; rng.seed = seed; rng.has_spare = false; rng.spare = 0.0;
            VoxelRates R;
#pragma unroll
            for (int r = 0; r < SSB_RD; r++) R.rr[r] = V.rrate[(size_t) r * N + i];
            // ---- arrivals (as in the event form): only dependents are re-evaluated, with the vol of a random arrival's source
            unsigned long long amask = 0ull;
            int n_arr = 0;
            double vol_src = vol;
            if (arrived) {
#pragma unroll
                for (int s = 0; s < SSB_SD; s++) {
                    if (inc[s]) {
                        xx[s] += (int) inc[s];
                        n_arr += (int) inc[s];
                        ((unsigned *) in_prev)[(size_t) s * N + i] = 0u;
                        amask |= ssb_gen::dep_mask_species(s);
                    }
                }
                const int src = (int) (__ldcg(&V.inbox_src[buf ^ 1][i]) & 0xffffffffull);
                if (src > 0 && !(V.flags & 1u)) vol_src = V.mass[src - 1] / V.rho[src - 1];
                V.inbox_src[buf ^ 1][i] = 0ull;
            }
#pragma unroll
            for (int s = 0; s < SSB_SD; s++) xr[s] = xx[s];
            // ---- diffusion leap --------------------------------------------------------------------------------------
            double sr = 0.0, sd = 0.0;
#pragma unroll
            for (int r = 0; r < SSB_RD; r++) sr += R.rr[r];
#pragma unroll
            for (int s = 0; s < SSB_SD; s++) sd += Dd[s] * xx[s];
            unsigned long long dmask = 0ull;
            int n_dep = 0;
            if (tau_w > 0.0 && sd > 0.0) {
                const double tot = sr + sd;
                const double plo = (V.flags & 1u) ? 0.0 : (sr / tot) * sd;       // reference: species pick lives on (p*sd, sd]
                double cum = 0.0;
#pragma unroll
                for (int s = 0; s < SSB_SD; s++) {
                    const double lo_s = cum;
                    cum += Dd[s] * xx[s];
                    if (xx[s] <= 0 || !(Dd[s] > 0.0)) continue;
                    double a_eff;
                    if (V.flags & 1u) a_eff = Dd[s] * xx[s];
                    else { const double ov = fmax(0.0, cum - fmax(lo_s, plo)); a_eff = tot * ov / sd; }
                    if (!(a_eff > 0.0)) continue;
                    const double q = 1.0 - exp(-(a_eff / xx[s]) * tau_w);
                    int nj = ssb_binomial(xx[s], q, rng);
                    if (nj <= 0) continue;
                    // destinations: one pass over the neighbour row, conditional binomials (multinomial), weights in list order
                    const int cnt = V.nbr_count[i];
                    const bool cached = V.Dij != nullptr;
                    double xl0 = 0, xl1 = 0, xl2 = 0, m_l = 0, rho_l = 0;
                    if (!cached) { xl0 = V.x[0][i]; xl1 = V.x[1][i]; xl2 = V.x[2][i]; m_l = V.mass[i]; rho_l = V.rho_search[i]; }
                    double wrem = V.Ddiag[(size_t) s * N + i];
                    int rem = nj, sent = 0, last_ok = -1;
                    for (int k = 0; k < cnt && rem > 0; k++) {
                        const int j = V.nbr[(size_t) k * N + i];
                        const double dc = V.dmat[s * V.num_types + (V.type[j] - 1)];
                        if (dc == 0.0) continue;
                        if (V.filter && !ssb_in_range(ssb_dist2(V.dim, xl0, xl1, xl2, V.x0[0][j], V.x0[1][j], V.x0[2][j]), V.h, __dmul_rn(V.h, V.h))) continue;
                        const double w = (cached ? V.Dij[(size_t) k * N + i] : pair_Dij(V, i, j, xl0, xl1, xl2, m_l, rho_l)) * dc;
                        last_ok = j;
                        const double pk = (wrem > w) ? w / wrem : 1.0;
                        const int mk = (pk >= 1.0) ? rem : ssb_binomial(rem, pk, rng);
                        wrem -= w;
                        if (mk > 0) {
                            rem -= mk;
                            if (j != i) {
                                sent += mk;
                                atomicAdd(&out_box[(size_t) s * N + j], (unsigned) mk);
                                const unsigned long long pri = (unsigned long long) (__double_as_longlong(rng.next() * 4294967296.0 * 4096.0) & 0xffffffffll);
                                atomicMax(&V.inbox_src[buf][j], (pri << 32) | (unsigned long long) (unsigned) (i + 1));
                                V.blk_mail[buf][j / SSB_BLOCK] = 1;
                            }
                        }
                    }
                    if (rem > 0 && last_ok >= 0 && last_ok != i) {               // round-off left-over goes to the last eligible neighbour
                        sent += rem;
                        atomicAdd(&out_box[(size_t) s * N + last_ok], (unsigned) rem);
                        V.blk_mail[buf][last_ok / SSB_BLOCK] = 1;
                    }
                    xx[s] -= sent;                                              // gone for diffusion, still reactive (xr) until the window closes
                    n_dep += nj;
                    n_df += (unsigned) nj;
                    dmask |= ssb_gen::dep_mask_species(s);
                }
            }
            // ---- propensities after arrivals / departures: dependents only; arrival-dependents with a source vol with the
            // probability that an arrival (not a departure) was the last event to touch them (simulate_rdme.cpp:419-437)
            if ((amask | dmask) != 0ull || (V.flags & 1u)) {
                double tmp_own[SSB_RD > 0 ? SSB_RD : 1], tmp_src[SSB_RD > 0 ? SSB_RD : 1];
                ssb_gen::eval_propensities(xr, t_lo, vol, df, type_i, tmp_own);
                bool use_src = false;
                if (n_arr > 0 && vol_src != vol) use_src = rng.next() * (double) (n_arr + n_dep) < (double) n_arr;
                if (use_src) ssb_gen::eval_propensities(xr, t_lo, vol_src, df, type_i, tmp_src);
                const unsigned long long mask = (V.flags & 1u) ? ~0ull : (amask | dmask);
#pragma unroll
                for (int r = 0; r < SSB_RD; r++) {
                    if ((mask >> r) & 1ull) R.rr[r] = (use_src && ((amask >> r) & 1ull)) ? tmp_src[r] : tmp_own[r];
                }
            }
            sr = 0.0;
#pragma unroll
            for (int r = 0; r < SSB_RD; r++) sr += R.rr[r];
            // ---- reaction SSA inside the window (clock re-drawn at the window start: memoryless) -------------------------
            // a voxel that sat idle keeps the reaction clock it stored (its rates did not change); anything that changed the state in
            // this window (arrivals, departures) makes a fresh draw from the window start valid by memorylessness
            double tr;
            if (tnext != -INFINITY && !arrived && n_dep == 0) tr = tnext;
            else tr = (sr > 0.0 && tau_w > 0.0) ? t_lo + (-log(rng.next())) / sr : INFINITY;
            int guard = 0;
            while (tr <= t_hi) {
                sd = 0.0;
#pragma unroll
                for (int s = 0; s < SSB_SD; s++) sd += Dd[s] * xx[s];
                double pick;
                if (V.flags & 1u) pick = rng.next() * sr;
                else pick = rng.next() * (sr / (sr + sd)) * sr;                  // rand1 <= srrate/totrate, then rand1*srrate (:253-261)
                int re = 0;
                double cum = R.rr[0];
#pragma unroll
                for (int q = 1; q < SSB_RD; q++) { if (pick > cum) { re = q; cum += R.rr[q]; } else break; }
                while (re > 0 && R.rr[re] <= 0.0) re--;
                int xn[SSB_SD > 0 ? SSB_SD : 1];
#pragma unroll
                for (int s = 0; s < SSB_SD; s++) xn[s] = xr[s];
                bool neg = false;
                ssb_gen::apply_stoich(re, xn, neg);
                if (neg) atomicCAS(V.err_flag, 0, 2);
                bool feasible = true;
#pragma unroll
                for (int s = 0; s < SSB_SD; s++) feasible &= (xx[s] + (xn[s] - xr[s]) >= 0);
                if (feasible) {
#pragma unroll
                    for (int s = 0; s < SSB_SD; s++) { xx[s] += xn[s] - xr[s]; xr[s] = xn[s]; }
                    n_rx++;
                }
                double tmp[SSB_RD > 0 ? SSB_RD : 1];
                ssb_gen::eval_propensities(xr, tr, vol, df, type_i, tmp);
                const unsigned long long mask = (V.flags & 1u) ? ~0ull : ssb_gen::dep_mask_reaction(re);
                sr = 0.0;
#pragma unroll
                for (int r = 0; r < SSB_RD; r++) { if ((mask >> r) & 1ull) R.rr[r] = tmp[r]; sr += R.rr[r]; }
                tr = (sr > 0.0) ? tr + (-log(rng.next())) / sr : INFINITY;
                if (++guard > 100000000) { atomicCAS(V.err_flag, 0, 2); break; }
            }
            // ---- window closes: departed molecules leave the reactive population ------------------------------------------
            bool departed = false, mobile = false;
#pragma unroll
            for (int s = 0; s < SSB_SD; s++) { departed |= (xr[s] != xx[s]); mobile |= (xx[s] > 0 && Dd[s] > 0.0); }
            if (departed) {
                double tmp[SSB_RD > 0 ? SSB_RD : 1];
                ssb_gen::eval_propensities(xx, t_hi, vol, df, type_i, tmp);
                const unsigned long long mask = (V.flags & 1u) ? ~0ull : dmask;
                sr = 0.0;
#pragma unroll
                for (int r = 0; r < SSB_RD; r++) { if ((mask >> r) & 1ull) R.rr[r] = tmp[r]; sr += R.rr[r]; }
            }
            sd = 0.0;
#pragma unroll
            for (int s = 0; s < SSB_SD; s++) sd += Dd[s] * xx[s];
            // voxels with mobile molecules are due in every window; idle ones keep a reaction clock for the triage
            tn_store = mobile ? -INFINITY : ((sr > 0.0) ? t_hi + (-log(rng.next())) / sr : INFINITY);
#pragma unroll
            for (int s = 0; s < SSB_SD; s++) V.xx[(size_t) s * N + i] = (unsigned) xx[s];
#pragma unroll
            for (int r = 0; r < SSB_RD; r++) V.rrate[(size_t) r * N + i] = R.rr[r];
            V.srrate[i] = sr;
            V.sdrate[i] = sd;
            V.tnext[i] = tn_store;
        }
        const double tn_final = block_min(valid ? tn_store : INFINITY);
        if (threadIdx.x == 0) V.blk_tmin[chunk] = tn_final;
        __syncthreads();
    }
    }
}

// event counters (ParticleSystem::total_reactions / total_diffusion): warp reduce, one atomic per warp
__device__ __forceinline__ void flush_event_counters(const SsbView &V, unsigned n_rx, unsigned n_df) {
    for (int o = 16; o > 0; o >>= 1) {
        n_rx += __shfl_xor_sync(0xffffffffu, n_rx, o);
        n_df += __shfl_xor_sync(0xffffffffu, n_df, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_rx) atomicAdd(&V.counters[0], (unsigned long long) n_rx);
        if (n_df) atomicAdd(&V.counters[1], (unsigned long long) n_df);
    }
}

template <bool LEAP>
__global__ void __launch_bounds__(SSB_BLOCK) k_rdme_window(SsbView V, double t_lo, double t_hi, double tau, uint64_t seed,
                                                          uint64_t epoch, int buf) {
    unsigned n_rx = 0, n_df = 0;
    if (LEAP) rdme_window_body_leap(V, t_lo, t_hi, tau, seed, epoch, buf, n_rx, n_df);
    else rdme_window_body(V, t_lo, t_hi, tau, seed, epoch, buf, n_rx, n_df);
    flush_event_counters(V, n_rx, n_df);
}

// The step-end overshoot event of a SLAB-decomposed run (simulate_rdme.cpp:233-238; single-GPU twin: windows nwin+1 / nwin+2 of
// k_rdme_windows_coop): the globally earliest pending clock t_min is a minimum over ranks that only exists in DEVICE memory (reduced
// over peer-mapped boards, ssb_core.cu), so the window bounds are read from there — no host round trip.  deliver = 0: the event
// window [t_end, t_min]; deliver = 1: the zero-length window (t_min, t_min) that delivers the molecule if it jumped.
template <bool LEAP>
__global__ void __launch_bounds__(SSB_BLOCK) k_rdme_window_dev(SsbView V, const unsigned long long *tmin_bits, double te, int deliver,
                                                              double tau, uint64_t seed, uint64_t epoch, int buf) {
    const double tmin = __longlong_as_double((long long) __ldcg(tmin_bits));
    if (!(tmin < INFINITY) || !(tmin > te)) return;                  // nothing pending anywhere (uniform over the grid and over the ranks)
    unsigned n_rx = 0, n_df = 0;
    const double lo = deliver ? tmin : te;
    if (LEAP) rdme_window_body_leap(V, lo, tmin, tau, seed, epoch, buf, n_rx, n_df);
    else rdme_window_body(V, lo, tmin, tau, seed, epoch, buf, n_rx, n_df);
    flush_event_counters(V, n_rx, n_df);
}

// All sSSA windows of one engine step in ONE cooperative launch: windows w = 0..nwin-1 of [t0, t0+dt], then the
// zero-length closing window that delivers in-flight molecules; a grid-wide barrier separates consecutive windows
// (a window's inbox writes must be complete before the next window reads them).  Removes the per-window launch
// latency that dominates small systems (config 1: ~1900 windows per step) and idle windows of large ones.
// SINGLE = true: the whole model fits the chunks of ONE CTA (small ensembles: birth-death, Cdc42): the barrier between windows is a
// block barrier, the launch is an ordinary one, and many trajectories run side by side on separate streams.
template <bool SINGLE, bool LEAP>
__global__ void __launch_bounds__(SSB_BLOCK) k_rdme_windows_coop(SsbView V, double t0, double dt, long long nwin, double tau,
                                                                uint64_t seed, uint64_t epoch0, int buf0) {
    // Moving domains, parity mode: the reference's `while(tt <= end_time)` tests the PREVIOUS event's time, so each take_step runs
    // one event past the step's end (simulate_rdme.cpp:233-238) — and the NSM is rebuilt every step (:54-65), so that overshoot is a
    // real extra event.  After the closing window every CTA reduces the chunk minima to the earliest pending clock t_min, window
    // nwin+1 = [t_end, t_min] executes exactly that event and window nwin+2 = (t_min, t_min) delivers it.
    // (Host-driven twin for slabs, where t_min is a minimum over ranks: rdme_min_time / rdme_extra_event in ssb_core.cu.)
    __shared__ double sh_tmin_all;
    const bool overshoot = !V.static_domain && !(V.flags & (1u | 128u));
    const long long wlast = overshoot ? nwin + 2 : nwin;
    const double te = t0 + dt;
    unsigned n_rx = 0, n_df = 0;
    int buf = buf0;
    double tmin = INFINITY;
    for (long long w = 0; w <= wlast; w++) {
        double lo, hi;
        if (w < nwin) {
            lo = t0 + dt * ((double) w / (double) nwin);
            hi = (w + 1 == nwin) ? te : t0 + dt * ((double) (w + 1) / (double) nwin);
        } else if (w == nwin) {
            lo = hi = te;
        } else if (w == nwin + 1) {
            const int nchunks = (V.N + SSB_BLOCK - 1) / SSB_BLOCK;
            double m = INFINITY;
            for (int c = threadIdx.x; c < nchunks; c += SSB_BLOCK) m = fmin(m, __ldcg(&V.blk_tmin[c]));
            m = block_min(m);
            if (threadIdx.x == 0) sh_tmin_all = m;
            __syncthreads();
            tmin = sh_tmin_all;
            if (!SINGLE) cooperative_groups::this_grid().sync();   // nobody may update a chunk minimum while another CTA is still reducing
            if (!(tmin < INFINITY && tmin > te)) break;            // uniform over the grid
            lo = te; hi = tmin;
        } else {
            lo = hi = tmin;
        }
        if (LEAP) rdme_window_body_leap(V, lo, hi, tau, seed, epoch0 + (uint64_t) w, buf, n_rx, n_df);
        else rdme_window_body(V, lo, hi, tau, seed, epoch0 + (uint64_t) w, buf, n_rx, n_df);
        buf ^= 1;
        if (w < wlast) {
            if (SINGLE) __syncthreads();       // one CTA: a block barrier orders the inbox traffic, and (unlike a device-scope fence) keeps L1 warm
            else cooperative_groups::this_grid().sync();
        }
    }
    flush_event_counters(V, n_rx, n_df);
}

// ---------------------------------------------------------------------------------------------
// host launchers (the table the core library calls through)
// ---------------------------------------------------------------------------------------------
static inline unsigned grid_for(int n) { return (unsigned) ((n + SSB_BLOCK - 1) / SSB_BLOCK); }

static int l_predictor(const SsbView *V, unsigned step, cudaStream_t st) {
    k_predictor<<<grid_for(V->N), SSB_BLOCK, 0, st>>>(*V, step);
    return (int) cudaGetLastError();
}
static int l_force(const SsbView *V, unsigned step, int full, cudaStream_t st) {
    if (full) k_force<true><<<grid_for(V->N), SSB_BLOCK, 0, st>>>(*V, step);
    else k_force<false><<<grid_for(V->N), SSB_BLOCK, 0, st>>>(*V, step);
    return (int) cudaGetLastError();
}
static int l_force_mv(const SsbView *V, unsigned step, unsigned long long *max_bits, cudaStream_t st) {
    k_force_mv<<<grid_for(V->N), SSB_BLOCK, 0, st>>>(*V, step, max_bits);
    return (int) cudaGetLastError();
}
static int l_corrector(const SsbView *V, unsigned step, cudaStream_t st) {
    k_corrector<<<grid_for(V->N), SSB_BLOCK, 0, st>>>(*V, step);
    return (int) cudaGetLastError();
}
static int l_finish(const SsbView *V, unsigned step, int moving, cudaStream_t st) {
    if (moving) k_finish<true><<<grid_for(V->N), SSB_BLOCK, 0, st>>>(*V, step);
    else k_finish<false><<<grid_for(V->N), SSB_BLOCK, 0, st>>>(*V, step);
    return (int) cudaGetLastError();
}
static int l_diff_init(const SsbView *V, unsigned long long *max_bits, cudaStream_t st) {
    k_diff_init<<<grid_for(V->N), SSB_BLOCK, 0, st>>>(*V, max_bits);
    return (int) cudaGetLastError();
}
static int l_static_coef(const SsbView *V, cudaStream_t st) {
    k_static_coef<<<grid_for(V->N), SSB_BLOCK, 0, st>>>(*V);
    return (int) cudaGetLastError();
}
static int l_static_step(const SsbView *V, unsigned step, int in_buf, cudaStream_t st) {
    if (step == 0) k_static_seed<<<grid_for(V->N), SSB_BLOCK, 0, st>>>(*V, in_buf);     // C (species-major) -> Cpre[in] (voxel-major)
    k_static_step<<<grid_for(V->N), SSB_BLOCK, 0, st>>>(*V, step, in_buf);
    return (int) cudaGetLastError();
}
static int l_rdme_init(const SsbView *V, double t0, double t_eval, double tau, uint64_t seed, uint64_t epoch, cudaStream_t st) {
    k_rdme_init<<<grid_for(V->N), SSB_BLOCK, 0, st>>>(*V, t0, t_eval, tau, seed, epoch);
    return (int) cudaGetLastError();
}
// all windows of a step; returns the number of kernel launches used through *launches
static int l_rdme_windows(const SsbView *V, double t0, double dt, long long nwin, double tau, uint64_t seed, uint64_t epoch0,
                          int buf0, int *launches, cudaStream_t st) {
    static int coop_blocks_of[2] = {-1, -1};   // co-resident CTAs of the cooperative kernel on this device (0 = unsupported), [leap]
    const bool leap = (V->flags & 32u) != 0;
    if (coop_blocks_of[leap] < 0) {
        int dev = 0, coop = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (leap) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rdme_windows_coop<false, true>, SSB_BLOCK, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rdme_windows_coop<false, false>, SSB_BLOCK, 0);
        coop_blocks_of[leap] = coop ? sms * per_sm : 0;
    }
    const int coop_blocks = coop_blocks_of[leap];
    const unsigned nchunks = grid_for(V->N);
    if (nchunks <= 8) {      // small model: one CTA walks all chunks; ordinary launch, block barrier between windows
        if (leap) k_rdme_windows_coop<true, true><<<1, SSB_BLOCK, 0, st>>>(*V, t0, dt, nwin, tau, seed, epoch0, buf0);
        else k_rdme_windows_coop<true, false><<<1, SSB_BLOCK, 0, st>>>(*V, t0, dt, nwin, tau, seed, epoch0, buf0);
        if (launches) *launches = 1;
        return (int) cudaGetLastError();
    }
    if (coop_blocks > 0) {
        unsigned grid = (unsigned) coop_blocks;
        if (grid > nchunks) grid = nchunks;
        SsbView view = *V;
        void *args[] = {&view, &t0, &dt, &nwin, &tau, &seed, &epoch0, &buf0};
        cudaError_t e = cudaLaunchCooperativeKernel(leap ? (const void *) k_rdme_windows_coop<false, true> : (const void *) k_rdme_windows_coop<false, false>, dim3(grid), dim3(SSB_BLOCK), args, 0, st);
        if (launches) *launches = 1;
        return (int) e;
    }
    int buf = buf0;
    unsigned grid = 148u * 8u;
    if (grid > nchunks) grid = nchunks;
    for (long long w = 0; w <= nwin; w++) {
        double lo = (w < nwin) ? t0 + dt * ((double) w / (double) nwin) : t0 + dt;
        double hi = (w < nwin) ? ((w + 1 == nwin) ? t0 + dt : t0 + dt * ((double) (w + 1) / (double) nwin)) : t0 + dt;
        if (leap) k_rdme_window<true><<<grid, SSB_BLOCK, 0, st>>>(*V, lo, hi, tau, seed, epoch0 + (uint64_t) w, buf);
        else k_rdme_window<false><<<grid, SSB_BLOCK, 0, st>>>(*V, lo, hi, tau, seed, epoch0 + (uint64_t) w, buf);
        buf ^= 1;
    }
    if (launches) *launches = (int) (nwin + 1);
    const int e = (int) cudaGetLastError();
    if (e) return e;
    return (!V->static_domain && !(V->flags & (1u | 128u))) ? -1 : 0;      // -1: fine, but the step-end overshoot event is left to the host
}
static int l_rdme_window(const SsbView *V, double t_lo, double t_hi, double tau, uint64_t seed, uint64_t epoch, int buf, cudaStream_t st) {
    const unsigned nchunks = grid_for(V->N);
    unsigned grid = 148u * 8u;         // persistent grid: a multiple of the 148 SMs
    if (grid > nchunks) grid = nchunks;
    if (V->flags & 32u) k_rdme_window<true><<<grid, SSB_BLOCK, 0, st>>>(*V, t_lo, t_hi, tau, seed, epoch, buf);
    else k_rdme_window<false><<<grid, SSB_BLOCK, 0, st>>>(*V, t_lo, t_hi, tau, seed, epoch, buf);
    return (int) cudaGetLastError();
}

static int l_rdme_window_dev(const SsbView *V, const unsigned long long *tmin_bits, double te, int deliver, double tau, uint64_t seed,
                             uint64_t epoch, int buf, cudaStream_t st) {
    const unsigned nchunks = grid_for(V->N);
    unsigned grid = 148u * 8u;
    if (grid > nchunks) grid = nchunks;
    if (V->flags & 32u) k_rdme_window_dev<true><<<grid, SSB_BLOCK, 0, st>>>(*V, tmin_bits, te, deliver, tau, seed, epoch, buf);
    else k_rdme_window_dev<false><<<grid, SSB_BLOCK, 0, st>>>(*V, tmin_bits, te, deliver, tau, seed, epoch, buf);
    return (int) cudaGetLastError();
}

static int l_warm(const SsbView *V0, cudaStream_t st) {
    SsbView V = *V0;
    V.N = 0;                                         // every kernel below bounds its work by V.N (or by the chunk count derived from it)
    k_predictor<<<1, SSB_BLOCK, 0, st>>>(V, 0u);
    k_force<true><<<1, SSB_BLOCK, 0, st>>>(V, 0u);
    k_force<false><<<1, SSB_BLOCK, 0, st>>>(V, 0u);
    k_force_mv<<<1, SSB_BLOCK, 0, st>>>(V, 0u, nullptr);
    k_corrector<<<1, SSB_BLOCK, 0, st>>>(V, 0u);
    k_finish<true><<<1, SSB_BLOCK, 0, st>>>(V, 0u);
    k_finish<false><<<1, SSB_BLOCK, 0, st>>>(V, 0u);
    k_rdme_window<false><<<1, SSB_BLOCK, 0, st>>>(V, 0.0, 0.0, 1.0, 0ull, 0ull, 0);
    k_rdme_window<true><<<1, SSB_BLOCK, 0, st>>>(V, 0.0, 0.0, 1.0, 0ull, 0ull, 0);
    k_diff_init<<<1, SSB_BLOCK, 0, st>>>(V, nullptr);
    // (the bounds word of the device-bounded windows: any readable word that is not a pending time — the event counter, >= 0 and "not > te")
    k_rdme_window_dev<false><<<1, SSB_BLOCK, 0, st>>>(V, V.counters, 1e300, 0, 1.0, 0ull, 0ull, 0);
    k_rdme_window_dev<true><<<1, SSB_BLOCK, 0, st>>>(V, V.counters, 1e300, 0, 1.0, 0ull, 0ull, 0);
    return (int) cudaGetLastError();
}

}  // namespace ssb_unit

extern "C" const SsbModelUnit *ssbm_get_unit() {
    static SsbModelUnit u;
    u.abi = SSB_UNIT_ABI;
    u.Sc = SSB_SC; u.Rc = SSB_RC; u.Sd = SSB_SD; u.Rd = SSB_RD; u.ndf = SSB_NDF; u.ntypes = SSB_NTYPES;
    u.S = SSB_S; u.R = SSB_R;
    u.predictor = ssb_unit::l_predictor;
    u.force = ssb_unit::l_force;
    u.force_mv = ssb_unit::l_force_mv;
    u.corrector = ssb_unit::l_corrector;
    u.finish = ssb_unit::l_finish;
    u.diff_init = ssb_unit::l_diff_init;
    u.has_bc = SSB_HAS_BC; u.bc_touches_rho = SSB_BC_TOUCHES_RHO;
    u.static_coef = ssb_unit::l_static_coef;
    u.static_step = ssb_unit::l_static_step;
    u.rdme_init = ssb_unit::l_rdme_init;
    u.block = SSB_BLOCK;
    u.rdme_window = ssb_unit::l_rdme_window;
    u.rdme_windows = ssb_unit::l_rdme_windows;
    u.rdme_window_dev = ssb_unit::l_rdme_window_dev;
    u.warm = ssb_unit::l_warm;
    return &u;
}
