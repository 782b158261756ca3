// ssb_unit_abi.h — the table a compiled model unit (one .so per model, built by codegen.py with
// nvcc -arch=sm_100a) hands to libssb_core.  Replaces the reference's function-pointer tables
// ALLOC_propensities()/ALLOC_ChemRxnFun() and the generated applyBoundaryConditions()
// (E/propensity_file_template.cpp:47-66,92-94; E/include/propensities.hpp:44-53).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

struct SsbView;

#define SSB_UNIT_ABI 12

struct SsbModelUnit {
    int abi;
    int Sc, Rc, Sd, Rd, ndf, ntypes, S, R;
    int has_bc, bc_touches_rho;
    int block;                 // voxels per sSSA chunk (index unit of blk_tmin / blk_mail)
    int (*predictor)(const SsbView *, unsigned step, cudaStream_t);
    int (*force)(const SsbView *, unsigned step, int full, cudaStream_t);
    int (*force_mv)(const SsbView *, unsigned step, unsigned long long *max_ddiag_bits, cudaStream_t);
    int (*corrector)(const SsbView *, unsigned step, cudaStream_t);
    int (*finish)(const SsbView *, unsigned step, int moving, cudaStream_t);
    int (*diff_init)(const SsbView *, unsigned long long *max_ddiag_bits, cudaStream_t);
    int (*rdme_init)(const SsbView *, double t0, double t_eval, double tau, uint64_t seed, uint64_t epoch, cudaStream_t);
    int (*static_coef)(const SsbView *, cudaStream_t);
    int (*static_step)(const SsbView *, unsigned step, int in_buf, cudaStream_t);
    int (*rdme_window)(const SsbView *, double t_lo, double t_hi, double tau, uint64_t seed, uint64_t epoch, int buf, cudaStream_t);
    int (*rdme_windows)(const SsbView *, double t0, double dt, long long nwin, double tau, uint64_t seed, uint64_t epoch0, int buf0,
                        int *launches, cudaStream_t);
    // slab runs: the step-end overshoot windows with the (global) earliest pending clock read from device memory
    int (*rdme_window_dev)(const SsbView *, const unsigned long long *tmin_bits, double te, int deliver, double tau, uint64_t seed,
                           uint64_t epoch, int buf, cudaStream_t);
    // launch every kernel of the unit once on an EMPTY view (N = 0: all of them return at once).  Loading a kernel for the first
    // time, or growing the device's local-memory pool for it, synchronises the whole CUDA context; slab ranks that share a context
    // (ranks as threads) must have that behind them before one of them spins on a neighbour's message.
    int (*warm)(const SsbView *, cudaStream_t);
};

extern "C" const SsbModelUnit *ssbm_get_unit();
