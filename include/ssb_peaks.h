/* ssb_peaks.h — measurement helper beside the engine's C-ABI (libssb_peaks.so; not on the product path).
 *
 * the neighbour sweeps of ssa_sdpd (pairwiseForce, E/src/model.cpp:39-191; find_neighbors,
 * E/src/particle.cpp:240-294) do ~70 flop per compulsory byte in fp64, so next to the HBM roofline the bench reports
 * their flop rate against the device's MEASURED fp64 FMA peak.  The reference has no counterpart (CPU code). */
#ifndef SSB_PEAKS_H
#define SSB_PEAKS_H
#ifdef __cplusplus
extern "C" {
#endif
/* Issue-bound DFMA loop on every SM of `device`; *tflops = 2 flop per FMA / best of 4 timed launches (CUDA events);
 * *best_ms (optional) = that launch's duration.  Returns 0 on success, 3 on a CUDA error, 4 on a bad argument. */
int ssb_fp64_peak(int device, double *tflops, double *best_ms);
#ifdef __cplusplus
}
#endif
#endif
