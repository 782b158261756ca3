// force_emu.cpp — TEST INFRASTRUCTURE: the SOURCE of the moving-domain force sweep k_force_mv (extracted verbatim from
// spatialpy_b200/csrc/ssb_model_unit.cuh by tests/test_cpu_abi.py into EMU_KERNELS) run on the host through emu_shim.h, so that
// its two record layouts (SSB_REC2: the neighbour's 1/rho, P/rho^2 and concentrations from one derived 32-byte sector, or
// recomputed / gathered per pair) can be compared without a GPU.
#include "emu_shim.h"

#include SSB_MODEL_HEADER          // generated ssb_gen namespace + SSB_* sizes of the test model (same text nvcc compiles)

#define ssb_ld256 ssb_ld256_device_asm      // the device version is inline PTX; never instantiated here
#include "ssb_device.cuh"
#undef ssb_ld256
static inline ssb_d4 ssb_ld256(const double *p) { ssb_d4 r; r.a = p[0]; r.b = p[1]; r.c = p[2]; r.d = p[3]; return r; }

#ifndef SSB_BLOCK
#define SSB_BLOCK 128
#endif

namespace ssb_unit {
#include EMU_KERNELS
}

struct EmuArgs {
    int N, dim, num_types, filter;
    unsigned flags;
    double dt, h, rho0, P0;
    double *rec; int *nbr; int *nbr_count; int nbr_cap; int *owned;
    double *F[3], *Fbp[3], *Frho, *C, *Q, *Ddiag, *data_fn;
    const double *dmat;
    unsigned long long *max_bits;
    double *rec2;
};

extern "C" int emu_force(const EmuArgs *a, unsigned step) {
    SsbView V;
    std::memset(&V, 0, sizeof(V));
    V.N = a->N; V.dim = a->dim; V.num_types = a->num_types; V.filter = a->filter; V.flags = a->flags;
    V.Sc = SSB_SC; V.Rc = SSB_RC; V.Sd = SSB_SD; V.Rd = SSB_RD; V.ndf = SSB_NDF;
    V.dt = a->dt; V.h = a->h; V.rho0 = a->rho0; V.P0 = a->P0;
    V.rec = a->rec; V.nbr = a->nbr; V.nbr_count = a->nbr_count; V.nbr_cap = a->nbr_cap; V.owned = a->owned;
    for (int d = 0; d < 3; d++) { V.F[d] = a->F[d]; V.Fbp[d] = a->Fbp[d]; }
    V.Frho = a->Frho; V.C = a->C; V.Q = a->Q; V.Ddiag = a->Ddiag; V.data_fn = a->data_fn; V.dmat = a->dmat;
    const unsigned blocks = (unsigned) ((a->N + SSB_BLOCK - 1) / SSB_BLOCK);
    V.rec2 = a->rec2;
    emu_launch(blocks, SSB_BLOCK, ssb_unit::k_force_mv, V, step, a->max_bits);
    return 0;
}
