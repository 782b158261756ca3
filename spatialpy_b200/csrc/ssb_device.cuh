// ssb_device.cuh — device-side state view and pair arithmetic shared by libssb_core and the model units.
//
// Data layout in HBM: every per-particle quantity is its own fp64/int32 array (SoA) in
// *storage order* = cell-sorted order of the most recent cell-list build; `id[i]` maps a storage slot back
// to the reference's particle id (the reference never permutes `system->particles`, E/src/output.cpp:78-84).
// Multi-component fields are split per component (x[0], x[1], x[2] are three arrays) so that a warp reading
// 32 consecutive particles issues fully coalesced 256-byte requests.  Species-indexed fields are
// species-major: C[s*N + i], xx[s*N + i], Ddiag[s*N + i], rrate[r*N + i].
// Neighbour lists are index-only ELL, transposed: nbr[k*N + i] is the k-th neighbour of particle i, so the
// k-th neighbour load of a warp is one coalesced request; r, dWdr and D_i_j are recomputed in registers.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define SSB_MAX_DIM 3

struct SsbView {
    int N;
    int dim;
    int static_domain;
    int num_types;
    int Sc, Rc, Sd, Rd, ndf;
    unsigned flags;
    double dt, h, rho0, c0, P0;
    double gravity[3];
    // particle state, storage order
    double *x[3];       // live positions
    double *x0[3];      // snapshot at step start = what the reference's kd-tree holds (simulate_threads.cpp:100-104)
    double *v[3], *vt[3], *F[3], *Fbp[3];
    double *rho;        // live density (substeps 0/1); corrector writes rho_new
    double *rho_new;    // density after the corrector (+BC), seen by later particles in the BVF sweep (model.cpp:285-293)
    double *old_rho, *Frho, *bvf, *mass, *nu;
    int *type, *solid, *id;
    int *owned;         // 1 = this rank integrates the particle, 0 = ghost copy of a particle owned by a neighbouring slab
    int *gid;           // global particle id (Philox counter), == id on a single GPU
    double *C, *Q;      // [Sc*N]
    unsigned *xx;       // [Sd*N]
    double *data_fn;    // [ndf*N]
    // neighbour lists
    int *nbr;           // [cap*N]
    int *nbr_count;     // [N]
    int nbr_cap;
    // Verlet-skin candidate lists (moving domains): nbr[] holds every particle within h*(1+skin) of the positions at the last
    // list build; each sweep re-applies the reference's exact inclusion test against THIS step's snapshot (filter = 1), so the
    // neighbour SETS are still exactly ANN's.  xref = positions at the last build; disp_bits[0] = max |x - xref|^2 since the build,
    // disp_bits[1] = max squared displacement of a single step (both as the bit patterns of non-negative doubles).
    int filter;
    double search_h2;
    double *xref[3];
    unsigned long long *disp_bits;
    double *Dij;        // [cap*N] cached D_i_j (static domains) or nullptr
    double *rho_search; // density at neighbour-search time (frozen into D_i_j, particle.cpp:187)
    // Step 0 only, models whose boundary conditions ASSIGN rho: the reference searches at the top of take_step1 (simulate.cpp:61-63),
    // BEFORE the particle's own predictor / BC, and runs take_step1 particle by particle — so the D_i_j of the step-0 lists see the
    // particle's own density as it was (rho_pre) and a neighbour's density after its BC only if the neighbour comes earlier in the
    // particle vector.  nullptr whenever the lists were built after step 0 (all of take_step1 is complete then).
    double *rho_pre;
    // moving-domain gather record (written by k_predictor, read by the neighbour sweeps): one 128-byte line per particle
    //   rec[16*j + 0..2] x0   3..5 x   6..8 v   9..11 vt   12 rho   13 mass   14 nu   15 bits(id:32 | type:16 | solid:16)
    double *rec;
    int *solid_nbr;     // [N] moving domains: 1 if any CANDIDATE neighbour is a solid particle (written by k_search, read by k_finish)
    double *rec2;       // [4*N] moving domains, S_c <= 2: {1/rho, P/rho^2, C[0], C[1]} per particle (predictor -> force sweep), else nullptr
    // static-domain fast path: cached chemistry pair coefficient dQc_base (model.cpp:155) and the double-buffered
    // half-stepped concentrations the next sweep reads (see k_static_step)
    double *coef;       // [cap*N] or nullptr
    double *Cpre[2];    // [Sc*N] each
    // RDME
    double *rrate;      // [Rd*N]
    double *srrate, *sdrate, *tnext;
    double *Ddiag;      // [Sd*N]
    unsigned *inbox[2]; // [Sd*N] each, double-buffered arrivals
    // [N] (random priority << 32 | source slot + 1) of ONE arrival of the window, 0 = none.  The reference evaluates a destination's
    // propensities with the vol of the voxel the LAST molecule came from (simulate_rdme.cpp:433); within a window the last arrival
    // is a uniformly random one, which the atomicMax over random priorities reproduces deterministically.
    unsigned long long *inbox_src[2];
    // block summaries (one entry per SSB_BLOCK consecutive voxels): earliest tnext in the block, and whether any voxel of
    // the block has mail in inbox[b]; lets a whole block leave an sSSA window after two loads when nothing is due.
    double *blk_tmin;   // [ceil(N/SSB_BLOCK)]
    int *blk_mail[2];   // [ceil(N/SSB_BLOCK)] each
    const double *dmat; // [S*num_types]
    // error / counters (device)
    int *err_flag;      // 0 ok, SSB_ERR_*
    unsigned long long *counters; // [0]=reactions [1]=diffusions
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al. 2011).  key = (seed lo, seed hi);
// counter = (voxel id, draw index, window lo, window hi) so every draw is addressable and no RNG state is stored.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
    uint64_t p0 = (uint64_t) M0 * c[0], p1 = (uint64_t) M1 * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t) p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t) p1;
#endif
    uint32_t n0 = hi1 ^ c[1] ^ k[0];
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c[3] ^ k[1];
    uint32_t n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k[2] = {k0, k1};
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// Two uniforms in (0,1) with 52 random mantissa bits each from one Philox block.
__host__ __device__ __forceinline__ void philox_uniform2(uint32_t vox, uint32_t draw, uint64_t window, uint64_t seed,
                                                         double &u0, double &u1) {
    uint32_t o[4];
    philox4x32_10(vox, draw, (uint32_t) window, (uint32_t)(window >> 32), (uint32_t) seed, (uint32_t)(seed >> 32), o);
    uint64_t a = ((uint64_t) o[0] << 32) | o[1];
    uint64_t b = ((uint64_t) o[2] << 32) | o[3];
    u0 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    u1 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// 256-bit read-only global load (sm_100a LDG.E.256): one instruction per 32-byte sector of a gather record.
struct ssb_d4 { double a, b, c, d; };
__device__ __forceinline__ ssb_d4 ssb_ld256(const double *p) {
    ssb_d4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
}

// ------------------------------------------------------------------------------------------------
// Pair arithmetic — restates E/src/particle.cpp:150-210 (add_to_neighbor_list) in registers.
// ------------------------------------------------------------------------------------------------
// kernel normalisation alpha: 3-D 105/(16 pi h^3), 2-D 5/(pi h^2), 1-D `5 / 4 * h` == h (integer division!)
// (particle.cpp:169-175, model.cpp:200-206,243-249)
__host__ __device__ __forceinline__ double ssb_alpha(int dim, double h) {
    const double PI = 3.14159265358979323846;
    if (dim == 3) return 105 / (16 * PI * h * h * h);
    if (dim == 2) return 5 / (PI * h * h);
    return (5 / 4) * h;
}

// squared distance in the first `dim` coordinates, axis order, NO fused multiply-add: the inclusion test
// `0 < d2 <= h*h` must be bit-identical to x86-64 g++ (E/external/ANN/src/kd_fix_rad_search.cpp:160-178).
__device__ __forceinline__ double ssb_dist2(int dim, double ax, double ay, double az, double bx, double by, double bz) {
    double t = __dsub_rn(ax, bx);
    double d = __dmul_rn(t, t);           // dist = 0 + t*t
    if (dim > 1) { t = __dsub_rn(ay, by); d = __dadd_rn(d, __dmul_rn(t, t)); }
    if (dim > 2) { t = __dsub_rn(az, bz); d = __dadd_rn(d, __dmul_rn(t, t)); }
    return d;
}

// the reference's neighbour rule: ANN's 0 < d2 <= h*h (kd_fix_rad_search.cpp:168-176) and particle.cpp:160-162's sqrt(d2) <= h
__device__ __forceinline__ bool ssb_in_range(double d2, double h, double h2) {
    return (d2 <= h2) && (d2 != 0.0) && !(sqrt(d2) > h);
}

// dWdr of the Wendland-type kernel, frozen at search time (particle.cpp:178)
__device__ __forceinline__ double ssb_dWdr(double alpha, double r, double h) {
    double R = r / h;
    return alpha * (-12 * r / (h * h)) * ((1 - R) * (1 - R));
}

// kernel value W (model.cpp:221,275)
__device__ __forceinline__ double ssb_W(double alpha, double r, double h) {
    double R = r / h;
    double q = 1 - R;
    return alpha * ((1 + 3 * R) * (q * q * q));
}

// index of (species s, type) in the species-major diffusion table for the PDE flux (model.cpp:163).  The reference reads
// [S_c*(type-1)+s] — a type-major index into a species-major table (solver.py:269-286) — mirrored by default;
// SSB_FLAG_CORRECTED_PDE_INDEX (512) reads [s*num_types + type-1], the entry simulate_rdme.cpp:146 uses for the same pair.
__device__ __forceinline__ int ssb_pde_dindex(const SsbView &V, int sc, int type_i, int s) {
    return (V.flags & 512u) ? s * V.num_types + (type_i - 1) : sc * (type_i - 1) + s;
}

// D_i_j (particle.cpp:182-187); always the 3-D constant, whatever the dimension
// the densities D_i_j freezes for the pair (i, j) (see SsbView::rho_pre)
__device__ __forceinline__ void ssb_search_rho(const SsbView &V, int i, int j, double &rho_i, double &rho_j) {
    rho_i = V.rho_search[i]; rho_j = V.rho_search[j];
    if (V.rho_pre) { rho_i = V.rho_pre[i]; if (!(V.gid[j] < V.gid[i])) rho_j = V.rho_pre[j]; }
}

__device__ __forceinline__ double ssb_Dij(double r2, double r, double h, double mi, double mj, double rhoi, double rhoj) {
    double ih = 1.0 / h;
    double ihsq = ih * ih;
    double dhr = h - r;
    double wfd = -25.066903536973515383e0 * dhr * dhr * ihsq * ihsq * ihsq * ih;
    return -2.0 * (mi * mj) / (mi + mj) * (rhoi + rhoj) / (rhoi * rhoj) * r2 * wfd / (r2 + 0.01 * h * h);
}
