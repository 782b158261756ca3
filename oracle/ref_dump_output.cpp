/* TEST SCAFFOLDING — not product code, never shipped, never linked into the CUDA path.
 *
 * Replacement translation unit for the reference engine's output stage
 * (E/src/output.cpp, E = spatialpy/solvers/c_base/ssa_sdpd-c-simulation-engine).
 * The reference prints v/rho/C with "%lf" (6 decimals, E/src/output.cpp:173-208), which cannot
 * support a 1e-12 comparison, so the parity build of oracle/_ref links this file INSTEAD of
 * output.cpp.  It defines the same three entry points (E/include/output.hpp:31-33) and writes, at
 * every output step, a raw little-endian dump of the full fp64 particle state, the neighbour lists
 * and the RDME rate tables.  Nothing of the reference's arithmetic is touched.
 *
 * One normalisation: Particle::Particle (E/src/particle.cpp:70-83) never initialises F, Fbp, Frho, vt,
 * bvf_phi, normal, old_x, old_v, old_rho, so the reference starts from stack noise (observed here:
 * denormals in one run, 1e241 and a NaN abort in the next; -ftrivial-auto-var-init=zero does not reach
 * the temporary).  The first tap (step 0, which run_simulation issues BEFORE the first substep,
 * E/src/simulate_threads.cpp:238-247) therefore sets exactly those fields to 0 — the value the new engine
 * defines for them — making the oracle deterministic.
 *
 * File "dump_<step>.bin" layout (all little endian):
 *   int64  N, Sc, Sd, step, initialized, nnz
 *   f64    x[N*3] v[N*3] vt[N*3] F[N*3] Fbp[N*3] normal[N*3]
 *   f64    rho[N] old_rho[N] Frho[N] bvf_phi[N] mass[N] nu[N]
 *   int32  type[N] solid[N] id[N]
 *   f64    C[N*Sc] Q[N*Sc]
 *   uint32 xx[N*Sd]
 *   int64  nbr_ptr[N+1]
 *   int32  nbr_idx[nnz]            (index into the particle vector == id, particles are never permuted)
 *   f64    nbr_dist[nnz] nbr_dWdr[nnz] nbr_Dij[nnz]
 *   if initialized: f64 srrate[N] sdrate[N] Ddiag[N*Sd] rrate[N*R]  (R appended as int64 before them)
 *   int64  total_reactions, total_diffusion
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <stdint.h>

#include "output.hpp"
#include "particle_system.hpp"

namespace Spatialpy{

    static void wr(FILE *fp, const void *p, size_t n){ if(n) fwrite(p, 1, n, fp); }

    void output_csv(ParticleSystem*, int){}

    void output_vtk__sync_step(ParticleSystem *system, int current_step){
        /* run_simulation releases the output thread one extra time at exit
         * (E/src/simulate_threads.cpp:288) while main() calls exit(0): never rewrite a finished dump. */
        static int last_step_dumped = -1;
        if(current_step == last_step_dumped) return;
        last_step_dumped = current_step;
        char filename[256];
        sprintf(filename, "dump_%06d.bin", current_step);
        FILE *fp = fopen(filename, "wb");
        if(!fp){ perror("dump"); exit(1); }
        if(current_step == 0){
            for(auto &p : system->particles){
                for(int d=0;d<3;d++){ p.F[d]=0.0; p.Fbp[d]=0.0; p.vt[d]=0.0; p.normal[d]=0.0; p.old_x[d]=0.0; p.old_v[d]=0.0; }
                p.Frho = 0.0; p.bvf_phi = 0.0; p.old_rho = 0.0;
            }
        }
        const int64_t N = (int64_t) system->particles.size();
        const int64_t Sc = (int64_t) system->num_chem_species;
        const int64_t Sd = (int64_t) system->num_stoch_species;
        const int64_t R = (int64_t) system->num_stoch_rxns;
        const int64_t init = (Sd > 0 && system->initialized) ? 1 : 0;
        int64_t nnz = 0;
        for(auto &p : system->particles) nnz += (int64_t) p.neighbors.size();
        int64_t hdr[6] = {N, Sc, Sd, (int64_t) current_step, init, nnz};
        wr(fp, hdr, sizeof(hdr));
        std::vector<double> b3(N*3), b1(N);
        std::vector<int32_t> i1(N);
#define DUMP3(field) for(int64_t i=0;i<N;i++){ for(int d=0;d<3;d++) b3[i*3+d] = system->particles[i].field[d]; } wr(fp, b3.data(), sizeof(double)*N*3);
#define DUMP1(field) for(int64_t i=0;i<N;i++){ b1[i] = system->particles[i].field; } wr(fp, b1.data(), sizeof(double)*N);
#define DUMPI(field) for(int64_t i=0;i<N;i++){ i1[i] = (int32_t) system->particles[i].field; } wr(fp, i1.data(), sizeof(int32_t)*N);
        DUMP3(x) DUMP3(v) DUMP3(vt) DUMP3(F) DUMP3(Fbp) DUMP3(normal)
        DUMP1(rho) DUMP1(old_rho) DUMP1(Frho) DUMP1(bvf_phi) DUMP1(mass) DUMP1(nu)
        DUMPI(type) DUMPI(solidTag) DUMPI(id)
        for(int64_t i=0;i<N;i++) wr(fp, system->particles[i].C, sizeof(double)*Sc);
        for(int64_t i=0;i<N;i++) wr(fp, system->particles[i].Q, sizeof(double)*Sc);
        if(Sd > 0){ for(int64_t i=0;i<N;i++) wr(fp, system->particles[i].xx, sizeof(unsigned int)*Sd); }
        std::vector<int64_t> ptr(N+1);
        ptr[0] = 0;
        for(int64_t i=0;i<N;i++) ptr[i+1] = ptr[i] + (int64_t) system->particles[i].neighbors.size();
        wr(fp, ptr.data(), sizeof(int64_t)*(N+1));
        std::vector<int32_t> idx(nnz);
        std::vector<double> nd(nnz), nw(nnz), nD(nnz);
        int64_t k = 0;
        Particle *base = &system->particles[0];
        for(int64_t i=0;i<N;i++){
            for(auto &n : system->particles[i].neighbors){
                idx[k] = (int32_t)(n.data - base);
                nd[k] = n.dist; nw[k] = n.dWdr; nD[k] = n.D_i_j;
                k++;
            }
        }
        wr(fp, idx.data(), sizeof(int32_t)*nnz);
        wr(fp, nd.data(), sizeof(double)*nnz);
        wr(fp, nw.data(), sizeof(double)*nnz);
        wr(fp, nD.data(), sizeof(double)*nnz);
        if(init){
            wr(fp, &R, sizeof(R));
            DUMP1(srrate) DUMP1(sdrate)
            for(int64_t i=0;i<N;i++) wr(fp, system->particles[i].Ddiag, sizeof(double)*Sd);
            for(int64_t i=0;i<N;i++) wr(fp, system->particles[i].rrate, sizeof(double)*R);
        }
        int64_t cnt[2] = {(int64_t) system->total_reactions, (int64_t) system->total_diffusion};
        if(Sd == 0){ cnt[0] = 0; cnt[1] = 0; }
        wr(fp, cnt, sizeof(cnt));
        fclose(fp);
    }

    void output_vtk__async_step(ParticleSystem*){}
}
