// halo_emu.cpp — TEST INFRASTRUCTURE: the SOURCE of the native slab transport kernels (k_halo_send/_recv/_wait, k_inbox_send/_recv,
// k_board_post/_reduce; extracted verbatim from spatialpy_b200/csrc/ssb_core.cu by tests/test_cpu_abi.py into EMU_KERNELS, minus the
// two inline-PTX helpers replaced below) run on the host through emu_shim.h: two "ranks" are two SsbViews in host memory and a
// "peer-mapped window" is a plain buffer, so the message layout, the last-CTA publish, the inbox compaction and the board
// all-reduce can be exercised without a GPU.  Says nothing about NVLink ordering or timing.
#include "emu_shim.h"

#include <chrono>

static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicCAS(int *p, int c, int v) { __atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return c; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
static inline unsigned long long ssb_globaltimer() {
    return (unsigned long long) std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline unsigned long long ssb_ld_flag(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline unsigned long long __shfl_xor_sync(unsigned m, unsigned long long v, int x) { return emu_shfl(m, v, (threadIdx.x & 31u) ^ (unsigned) x); }

#define SSB_S 1
#define SSB_SC 1
#define SSB_SD 1
#define ssb_ld256 ssb_ld256_device_asm      // inline PTX in the device header; never instantiated here
#include "ssb_device.cuh"
#undef ssb_ld256

// own namespace: libssb_core.so exports nvcc's host stubs of the very same kernels under the same mangled names, and it may
// already be loaded (RTLD_GLOBAL) in the test process — a plain call to k_halo_send would bind to the stub
namespace halo_emu {
#include EMU_KERNELS
}
using namespace halo_emu;

struct HaloEmuRank {
    int N, Sc, Sd;
    double *F[3], *Fbp[3], *Frho, *Q, *rho_new, *v[3], *bvf;
    unsigned *inbox[2];
    unsigned long long *inbox_src[2];
    int *blk_mail[2];
    const int *slot_of_id;
};
static SsbView view_of(const HaloEmuRank *r) {
    SsbView V;
    std::memset(&V, 0, sizeof(V));
    V.N = r->N; V.Sc = r->Sc; V.Sd = r->Sd;
    for (int d = 0; d < 3; d++) { V.F[d] = r->F[d]; V.Fbp[d] = r->Fbp[d]; V.v[d] = r->v[d]; }
    V.Frho = r->Frho; V.Q = r->Q; V.rho_new = r->rho_new; V.bvf = r->bvf;
    for (int b = 0; b < 2; b++) { V.inbox[b] = r->inbox[b]; V.inbox_src[b] = r->inbox_src[b]; V.blk_mail[b] = r->blk_mail[b]; }
    return V;
}
struct HaloEmuSide { const int *ids; int n; char *buf; unsigned long long *flag; unsigned long long *count; };

static unsigned grid_for(int n, unsigned block) { return (unsigned) std::max(1, (n + (int) block - 1) / (int) block); }

// pack rank r's rows of both faces into the two windows and publish `seq`
extern "C" int emu_halo_send(const HaloEmuRank *r, int group, const HaloEmuSide *s, unsigned long long seq, unsigned *done) {
    HaloPackArgs A;
    std::memset(&A, 0, sizeof(A));
    for (int k = 0; k < 2; k++) { A.side[k].ids = s[k].ids; A.side[k].n = s[k].n; A.side[k].peer_buf = s[k].buf; A.side[k].peer_flag = s[k].flag; A.side[k].peer_count = s[k].count; }
    A.seq = seq; A.done = done;
    emu_launch(grid_for(s[0].n + s[1].n, 128), 128, k_halo_send, view_of(r), group, A, r->slot_of_id);
    return 0;
}
extern "C" int emu_halo_recv(const HaloEmuRank *r, int group, const HaloEmuSide *s) {
    HaloUnpackArgs A;
    std::memset(&A, 0, sizeof(A));
    for (int k = 0; k < 2; k++) { A.side[k].ids = s[k].ids; A.side[k].n = s[k].n; A.side[k].buf = s[k].buf; A.side[k].count = s[k].count; }
    emu_launch(grid_for(s[0].n + s[1].n, 128), 128, k_halo_recv, view_of(r), group, A, r->slot_of_id);
    return 0;
}
extern "C" int emu_halo_wait(const unsigned long long *f0, const unsigned long long *f1, unsigned long long seq, int *err_flag) {
    emu_launch(1, 32, k_halo_wait, f0, f1, seq, err_flag);
    return 0;
}
extern "C" int emu_inbox_send(const HaloEmuRank *r, int buf, const HaloEmuSide *s, unsigned long long seq, unsigned *done, unsigned *icount) {
    HaloPackArgs A;
    std::memset(&A, 0, sizeof(A));
    for (int k = 0; k < 2; k++) { A.side[k].ids = s[k].ids; A.side[k].n = s[k].n; A.side[k].peer_buf = s[k].buf; A.side[k].peer_flag = s[k].flag; A.side[k].peer_count = s[k].count; }
    A.seq = seq; A.done = done; A.icount = icount;
    emu_launch(grid_for(s[0].n + s[1].n, 128), 128, k_inbox_send, view_of(r), buf, A, r->slot_of_id);
    return 0;
}
extern "C" int emu_inbox_recv(const HaloEmuRank *r, int buf, const HaloEmuSide *s, int block) {
    HaloUnpackArgs A;
    std::memset(&A, 0, sizeof(A));
    for (int k = 0; k < 2; k++) { A.side[k].ids = s[k].ids; A.side[k].n = s[k].n; A.side[k].buf = s[k].buf; A.side[k].count = s[k].count; }
    emu_launch(2, 128, k_inbox_recv, view_of(r), buf, A, r->slot_of_id, block);
    return 0;
}
// all-reduce of one scalar per rank over `world` boards (each SSB_BOARD_NCH * 2 * SSB_BOARD_MAXW * 2 words)
extern "C" int emu_board_allreduce(unsigned long long **boards, int world, int ch, unsigned long long seq, int take_min,
                                   const unsigned long long *values, unsigned long long *out, int *err_flag) {
    BoardPeers P;
    std::memset(&P, 0, sizeof(P));
    for (int r = 0; r < world; r++) P.b[r] = boards[r];
    for (int me = 0; me < world; me++) emu_launch(1, 32, k_board_post, P, world, me, ch, seq, values + me);
    for (int me = 0; me < world; me++) emu_launch(1, 32, k_board_reduce, (const unsigned long long *) boards[me], world, ch, seq, take_min, out + me, err_flag);
    return 0;
}
extern "C" int emu_board_words() { return (int) board_slot(SSB_BOARD_NCH, 0, 0); }
extern "C" int emu_halo_width(int group, int Sc) { return halo_width(group, Sc); }
