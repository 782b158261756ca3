/* Plain-C consumer of the C-ABI: proves include/ssb.h and include/ssb_peaks.h compile as C (not only as C++) and that the
 * libraries link from C.  Without a GPU it exercises the no-compute entry points and the loud failure of ssb_create;
 * with `run` as argv[1] it also creates an engine for a 2-particle model (GPU box only). */
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include "ssb.h"
#include "ssb_peaks.h"

int main(int argc, char **argv) {
    if (ssb_abi_version() != SSB_ABI_VERSION) { printf("abi mismatch\n"); return 1; }
    ssb_model m;
    memset(&m, 0, sizeof(m));
    m.abi_version = SSB_ABI_VERSION + 1;                 /* wrong version: must be refused before anything is touched */
    ssb_handle *h = NULL;
    if (ssb_create(&m, &h) != SSB_ERR_ARG || h != NULL) { printf("bad-version model accepted\n"); return 2; }
    if (ssb_reset(NULL, 1) != SSB_ERR_ARG || ssb_step(NULL, 1) != SSB_ERR_ARG || ssb_set_field(NULL, "x", NULL, 0) != SSB_ERR_ARG) {
        printf("NULL handle accepted\n");
        return 3;
    }
    int count = -1;
    int rc = ssb_device_count(&count);
    printf("abi %d, sizeof(ssb_model) %zu, device_count rc %d count %d\n", ssb_abi_version(), sizeof(ssb_model), rc, count);
    printf("offsets output_steps %zu h %zu gravity %zu x %zu u0 %zu species_names %zu rdme_epsilon %zu device %zu owned %zu rng_id %zu\n",
           offsetof(ssb_model, output_steps), offsetof(ssb_model, h), offsetof(ssb_model, gravity), offsetof(ssb_model, x),
           offsetof(ssb_model, u0), offsetof(ssb_model, species_names), offsetof(ssb_model, rdme_epsilon), offsetof(ssb_model, device),
           offsetof(ssb_model, owned), offsetof(ssb_model, rng_id));
    if (argc > 2 && strcmp(argv[1], "snapshot") == 0) {   /* the engine's writers from C: a 3-particle snapshot into argv[2] */
        const double x[9] = {0.0, 0.0, 0.0, 0.5, 0.25, 0.0, 1.0, 0.5, 0.0}, v[9] = {0}, lims[6] = {0.0, 1.0, 0.0, 0.5, 0.0, 0.0};
        const double scal[12] = {1.0, 1.0, 1.0, 0.1, 0.1, 0.1, 0.0, 0.0, 0.0, 2.0, 2.0, 2.0}, conc[3] = {0.5, 1.5, 2.5};
        const int32_t type[3] = {1, 2, 1};
        const uint32_t pop[3] = {7, 0, 42};
        const char *names[1] = {"A"};
        return ssb_write_snapshot(argv[2], 1, 10, 1, 3, 1, 1, names, lims, x, v, scal, conc, type, pop, 3u);
    }
    if (argc > 1 && strcmp(argv[1], "peak") == 0) {
        double tf = 0.0, ms = 0.0;
        rc = ssb_fp64_peak(0, &tf, &ms);
        printf("fp64 peak rc %d: %.2f TFLOP/s (%.3f ms)\n", rc, tf, ms);
        return rc;
    }
    return 0;
}
