// lookahead_emu.cpp — TEST INFRASTRUCTURE: the SOURCE of k_lookahead (spatialpy_b200/csrc/ssb_core.cu, extracted verbatim by
// tests/test_cpu_halo_emu.py into EMU_KERNELS) run on the host through emu_shim.h, so that the displacement it predicts for the NEXT
// predictor can be compared with what the oracle's take_step1 (E/src/simulate.cpp:56-109) actually does.
#include "emu_shim.h"

#define ssb_ld256 ssb_ld256_device_asm
#include "ssb_device.cuh"
#undef ssb_ld256

#define CORE_BLOCK 256
namespace lookahead_emu {
#include EMU_KERNELS
}

struct LookArgs { int N; double dt; double *x[3], *xref[3], *v[3], *F[3], *Fbp[3]; int *solid; unsigned long long *out; };

extern "C" int emu_lookahead(const LookArgs *a, unsigned blocks) {
    SsbView V;
    std::memset(&V, 0, sizeof(V));
    V.N = a->N; V.dt = a->dt; V.solid = a->solid;
    for (int d = 0; d < 3; d++) { V.x[d] = a->x[d]; V.xref[d] = a->xref[d]; V.v[d] = a->v[d]; V.F[d] = a->F[d]; V.Fbp[d] = a->Fbp[d]; }
    emu_launch(blocks, CORE_BLOCK, lookahead_emu::k_lookahead, V, a->out);
    return 0;
}
