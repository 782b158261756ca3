/* TEST INFRASTRUCTURE — serial CPU restatement of the reference's next-subvolume method.
 *
 * ORACLE of the statistical parity tests, never product code: compiled by tests/ (g++) together with the generated
 * model header (spatialpy_b200/codegen.py:generate_model_header, the same propensity text the CUDA unit compiles) and
 * called through ctypes.  Restates E/src/simulate_rdme.cpp (E = /root/reference/spatialpy/solvers/c_base/
 * ssa_sdpd-c-simulation-engine):
 *     nsm_core__initialize_rxn_propensities   :113-128       nsm_core__initialize_heap  :155-195
 *     nsm_core__initialize_diff_propensities  :131-152       nsm_core__take_step        :211-472
 * including the reference's rules that shape the law of the process: the channel pick `rand1 <= srrate/totrate` with
 * rand1*srrate / rand1*sdrate as the within-channel pick (:253-261, :317-321), destination propensities evaluated with
 * the SOURCE voxel's vol (:433), dependency-graph partial updates (:299, :419), restriction by destination type (:361).
 * The event queue is an indexed binary heap instead of NRMConstant_v5's hash bins — same ordering, same
 * Exp(1)/a + t redraw on update (NRMConstant_v5.cpp:98).  RNG: std::mt19937_64 like the reference (template:99).
 * PINNED statistically against ensembles of the unmodified reference (tests/test_cpu_oracle.py::test_nsm_*).
 */
#include <math.h>
#include <stdint.h>
#include <random>
#include <vector>
#include <limits>

#include SSB_MODEL_HEADER

namespace {
struct Heap {
    std::vector<double> t; std::vector<int> node, pos;   // node[k] = voxel at heap slot k; pos[v] = slot of voxel v
    void init(int n) { t.assign(n, 0.0); node.resize(n); pos.resize(n); for (int i = 0; i < n; i++) { node[i] = i; pos[i] = i; } }
    void swp(int a, int b) { std::swap(node[a], node[b]); pos[node[a]] = a; pos[node[b]] = b; }
    bool less(int a, int b) const { return t[node[a]] < t[node[b]]; }
    void up(int k) { while (k > 0) { int p = (k - 1) / 2; if (less(k, p)) { swp(k, p); k = p; } else break; } }
    void down(int k) { int n = (int) node.size(); for (;;) { int l = 2 * k + 1, r = l + 1, m = k; if (l < n && less(l, m)) m = l; if (r < n && less(r, m)) m = r; if (m == k) break; swp(k, m); k = m; } }
    void build() { for (int k = (int) node.size() / 2 - 1; k >= 0; k--) down(k); }
    void update(int v, double tv) { t[v] = tv; up(pos[v]); down(pos[v]); }
};
}

extern "C" int nsm_oracle_run(int N, const int64_t *nbr_ptr, const int32_t *nbr_idx, const double *nbr_Dij, const int32_t *type,
                              const double *vol_in, const double *data_fn, const double *dmat, int num_types, uint32_t *xx_io,
                              double t0, double t_end, uint64_t seed, int corrected, int64_t *n_rx, int64_t *n_df) {
    const int S = SSB_SD, R = SSB_RD, NDF = SSB_NDF;
    std::mt19937_64 rng(seed);
    std::exponential_distribution<double> expo(1.0);
    const double INF = std::numeric_limits<double>::infinity();
    std::vector<int> xx((size_t) N * S);
    for (size_t k = 0; k < xx.size(); k++) xx[k] = (int) xx_io[k];
    std::vector<double> rrate((size_t) N * (R > 0 ? R : 1)), srrate(N), sdrate(N), Ddiag((size_t) N * S);
    std::vector<double> df((size_t) (NDF > 0 ? NDF : 1));
    auto load_df = [&](int v) { for (int q = 0; q < NDF; q++) df[q] = data_fn[(size_t) q * N + v]; };
    double tmp[SSB_RD > 0 ? SSB_RD : 1];
    for (int v = 0; v < N; v++) {                                   // :113-128 (t = 0.0), :131-152
        load_df(v);
        ssb_gen::eval_propensities(&xx[(size_t) v * S], 0.0, vol_in[v], df.data(), type[v], tmp);
        double sr = 0; for (int r = 0; r < R; r++) { rrate[(size_t) v * R + r] = tmp[r]; sr += tmp[r]; }
        srrate[v] = sr;
        double sd = 0;
        for (int s = 0; s < S; s++) {
            double d = 0;
            for (int64_t k = nbr_ptr[v]; k < nbr_ptr[v + 1]; k++) d += dmat[s * num_types + (type[nbr_idx[k]] - 1)] * nbr_Dij[k];
            Ddiag[(size_t) v * S + s] = d; sd += d * xx[(size_t) v * S + s];
        }
        sdrate[v] = sd;
    }
    Heap hp; hp.init(N);
    for (int v = 0; v < N; v++) { double a = srrate[v] + sdrate[v]; hp.t[v] = a > 0 ? expo(rng) / a + t0 : INF; }   // NRMConstant_v5.cpp:52-59
    hp.build();
    int64_t nrx = 0, ndf = 0;
    double tt = t0;
    while (tt <= t_end) {                                           // :233 (one event past t_end is executed, as in the reference)
        int sv = hp.node[0];
        tt = hp.t[sv];
        if (!(tt < INF)) break;
        int *x = &xx[(size_t) sv * S];
        double *rr = &rrate[(size_t) sv * (R > 0 ? R : 1)];
        const double vol = vol_in[sv];
        double totrate = srrate[sv] + sdrate[sv];
        double rand1 = rng() * 1.0 / rng.max();
        bool is_rxn; double pick;
        if (corrected) { pick = rand1 * totrate; is_rxn = pick <= srrate[sv]; if (!is_rxn) pick -= srrate[sv]; }
        else { is_rxn = rand1 <= srrate[sv] / totrate; pick = is_rxn ? rand1 * srrate[sv] : rand1 * sdrate[sv]; }
        int dest = -1;
        if (is_rxn) {
            int re = 0; double cum = rr[0];
            for (; re < R && pick > cum; ) { re++; if (re < R) cum += rr[re]; }
            if (re >= R) { re = R - 1; while (re > 0 && rr[re] <= 0.0) re--; }
            bool neg = false;
            int before[SSB_SD > 0 ? SSB_SD : 1];
            for (int s = 0; s < S; s++) before[s] = x[s];
            ssb_gen::apply_stoich(re, x, neg);
            if (neg) return 2;
            for (int s = 0; s < S; s++) sdrate[sv] += Ddiag[(size_t) sv * S + s] * (x[s] - before[s]);    // :295
            load_df(sv);
            ssb_gen::eval_propensities(x, tt, vol, df.data(), type[sv], tmp);
            unsigned long long mask = ssb_gen::dep_mask_reaction(re);                                    // :299-307
            double rdelta = 0;
            for (int r = 0; r < R; r++) if ((mask >> r) & 1ull) { rdelta += tmp[r] - rr[r]; rr[r] = tmp[r]; }
            srrate[sv] += rdelta;
            nrx++;
        } else {
            int spec = 0; double cum = Ddiag[(size_t) sv * S] * x[0];
            for (; spec < S && pick > cum; ) { spec++; if (spec < S) cum += Ddiag[(size_t) sv * S + spec] * x[spec]; }
            if (spec >= S) { spec = S - 1; while (spec > 0 && x[spec] <= 0) spec--; }
            if (x[spec] <= 0) return 2;
            double r2 = rng() * 1.0 / rng.max();
            double target = r2 * Ddiag[(size_t) sv * S + spec];
            double cum2 = 0; int last_ok = -1;
            for (int64_t k = nbr_ptr[sv]; k < nbr_ptr[sv + 1]; k++) {                                     // :359-367
                int j = nbr_idx[k];
                double dc = dmat[spec * num_types + (type[j] - 1)];
                cum2 += nbr_Dij[k] * dc;
                if (dc != 0.0) last_ok = j;
                if (cum2 > target) { dest = j; break; }
            }
            if (dest < 0) dest = last_ok;
            if (dest < 0) return 2;
            x[spec]--;
            int *xd = &xx[(size_t) dest * S];
            xd[spec]++;
            if (R > 0) {                                                                                  // :418-439
                unsigned long long mask = ssb_gen::dep_mask_species(spec);
                load_df(sv);
                ssb_gen::eval_propensities(x, tt, vol, df.data(), type[sv], tmp);
                double rdelta = 0;
                for (int r = 0; r < R; r++) if ((mask >> r) & 1ull) { rdelta += tmp[r] - rr[r]; rr[r] = tmp[r]; }
                srrate[sv] += rdelta;
                double *rd = &rrate[(size_t) dest * R];
                load_df(dest);
                ssb_gen::eval_propensities(xd, tt, corrected ? vol_in[dest] : vol /* source vol, :433 */, df.data(), type[dest], tmp);
                double rrdelta = 0;
                for (int r = 0; r < R; r++) if ((mask >> r) & 1ull) { rrdelta += tmp[r] - rd[r]; rd[r] = tmp[r]; }
                srrate[dest] += rrdelta;
            }
            sdrate[sv] -= Ddiag[(size_t) sv * S + spec];
            sdrate[dest] += Ddiag[(size_t) dest * S + spec];
            ndf++;
        }
        double a = srrate[sv] + sdrate[sv];
        hp.update(sv, a > 0 ? expo(rng) / a + tt : INF);                                                  // :450-453
        if (dest >= 0 && dest != sv) { double b = srrate[dest] + sdrate[dest]; hp.update(dest, b > 0 ? expo(rng) / b + tt : INF); }
    }
    for (size_t k = 0; k < xx.size(); k++) xx_io[k] = (uint32_t) xx[k];
    if (n_rx) *n_rx = nrx;
    if (n_df) *n_df = ndf;
    return 0;
}

/* Tap for tests/test_cpu_abi.py: the generated propensity functions evaluated on the host — the same text nvcc compiles into
 * the CUDA model unit — so the reference's expression-conversion test (test/integration_tests/test_solver.py:202-256) can be
 * replayed on this code generator. */
extern "C" void nsm_oracle_eval_propensities(const int *x, double t, double vol, const double *data_fn, int sd, double *out) {
    ssb_gen::eval_propensities(x, t, vol, data_fn, sd, out);
}

