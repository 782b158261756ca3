// ssb_core.cu — libssb_core.so: the model-independent half of the B200-native ssa_sdpd engine.
//
//   * K1  cell-list neighbour search: cell keys -> counting sort (one radix digit = the cell key, ties broken by
//         previous storage index so the order is deterministic) -> SoA permutation -> index-only ELL neighbour lists.
//         Replaces buildKDTree (E/src/simulate_threads.cpp:80-108) + Particle::find_neighbors (E/src/particle.cpp:240-294)
//         with ANN's inclusion rule 0 < sum_d (q_d-p_d)^2 <= h*h (E/external/ANN/src/kd_fix_rad_search.cpp:160-178).
//   * step scheduler: run_simulation (E/src/simulate_threads.cpp:171-312) as a single CUDA stream of kernels; the
//         model-specialised kernels come from the per-model unit (ssb_model_unit.cuh) through SsbModelUnit.
//   * K8  output staging: gather to particle-id order on the device, one D2H into pinned memory, and a host writer
//         thread producing the reference's ASCII VTK byte format (E/src/output.cpp:104-229) while the GPU steps on
//         (mirrors output_system_thread, simulate_threads.cpp:60-73).
//   * the C-ABI of include/ssb.h.
// E = /root/reference/spatialpy/solvers/c_base/ssa_sdpd-c-simulation-engine.  No CPU fallback exists: every entry point
// fails with SSB_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX 3: ranges cost a no-op call unless a profiler is attached

#include "../../include/ssb.h"
#include "ssb_device.cuh"
#include "ssb_unit_abi.h"

#define CORE_BLOCK 256
#define SSB_NCAT 10
enum { CAT_CELLS = 0, CAT_PREDICTOR, CAT_SEARCH, CAT_FORCE, CAT_CORRECTOR, CAT_FINISH, CAT_DIFF_INIT, CAT_RDME_INIT, CAT_RDME_WINDOW, CAT_OUTPUT };

// =====================================================================================================
// device kernels
// =====================================================================================================
struct CellGrid {
    double lo[3];
    double inv_cell[3];
    int n[3];
    int ncells;
    int dim;
};

__device__ __forceinline__ int cell_coord(double x, double lo, double inv, int n) {
    int c = (int) floor((x - lo) * inv);
    return min(max(c, 0), n - 1);
}

__device__ __forceinline__ int cell_key(const CellGrid &g, double x, double y, double z, int &cx, int &cy, int &cz) {
    cx = cell_coord(x, g.lo[0], g.inv_cell[0], g.n[0]);
    cy = cell_coord(y, g.lo[1], g.inv_cell[1], g.n[1]);
    cz = cell_coord(z, g.lo[2], g.inv_cell[2], g.n[2]);
    return (cz * g.n[1] + cy) * g.n[0] + cx;
}

__global__ void k_cell_keys(int N, CellGrid g, const double *x, const double *y, const double *z, int *key, int *cell_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int cx, cy, cz;
    int k = cell_key(g, x[i], y[i], z[i], cx, cy, cz);
    key[i] = k;
    atomicAdd(&cell_count[k], 1);
}

// exclusive scan, three phases (tile = 1024 elements per block)
#define SCAN_TILE 1024
__global__ void k_scan_tiles(int n, const int *in, int *out, int *tile_sums) {
    __shared__ int sh[SCAN_TILE];
    int base = blockIdx.x * SCAN_TILE;
    for (int t = threadIdx.x; t < SCAN_TILE; t += blockDim.x) sh[t] = (base + t < n) ? in[base + t] : 0;
    __syncthreads();
    // each thread scans 4 consecutive items serially, then a block scan of the per-thread sums
    int t4 = threadIdx.x * 4;
    int a0 = sh[t4], a1 = sh[t4 + 1], a2 = sh[t4 + 2], a3 = sh[t4 + 3];
    int sum = a0 + a1 + a2 + a3;
    __shared__ int warp_sums[8];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = sum;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < 8) ? warp_sums[lane] : 0;
        for (int o = 1; o < 8; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        if (lane < 8) warp_sums[lane] = w;
    }
    __syncthreads();
    int excl = incl - sum + (wid > 0 ? warp_sums[wid - 1] : 0);
    if (base + t4 < n) out[base + t4] = excl;
    if (base + t4 + 1 < n) out[base + t4 + 1] = excl + a0;
    if (base + t4 + 2 < n) out[base + t4 + 2] = excl + a0 + a1;
    if (base + t4 + 3 < n) out[base + t4 + 3] = excl + a0 + a1 + a2;
    if (threadIdx.x == blockDim.x - 1) tile_sums[blockIdx.x] = excl + sum;
}
__global__ void k_scan_sums(int ntiles, int *tile_sums, int *total_out) {
    // single block: serial-in-chunks exclusive scan of the tile sums
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += blockDim.x) {
        int idx = base + threadIdx.x;
        int v = (idx < ntiles) ? tile_sums[idx] : 0;
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        __shared__ int ws[32];
        if (lane == 31) ws[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = (lane < (blockDim.x >> 5)) ? ws[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            ws[lane] = w;
        }
        __syncthreads();
        int excl = incl - v + (wid > 0 ? ws[wid - 1] : 0) + carry;
        if (idx < ntiles) tile_sums[idx] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}
__global__ void k_scan_add(int n, int *out, const int *tile_sums) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_sums[i / SCAN_TILE];
}

__global__ void k_scatter(int N, const int *key, const int *cell_start, int *cursor, int *perm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int k = key[i];
    int slot = atomicAdd(&cursor[k], 1);
    perm[cell_start[k] + slot] = i;
}

// one thread per cell: order the cell's members by previous storage index (deterministic result whatever the
// atomic arrival order was), and flag a non-identity permutation.
__global__ void k_sort_cells(int ncells, const int *cell_start, int N, int *perm, int *nonidentity) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    int b = cell_start[c], e = (c + 1 < ncells) ? cell_start[c + 1] : N;
    int n = e - b;
    int *a = perm + b;
    if (n > 1) {
        if (n <= 48) {
            for (int i = 1; i < n; i++) { int v = a[i]; int j = i - 1; while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; } a[j + 1] = v; }
        } else {  // heap sort for crowded (clamped boundary) cells
            for (int start = n / 2 - 1; start >= 0; start--) {
                int root = start;
                while (2 * root + 1 < n) { int ch = 2 * root + 1; if (ch + 1 < n && a[ch] < a[ch + 1]) ch++; if (a[root] < a[ch]) { int t = a[root]; a[root] = a[ch]; a[ch] = t; root = ch; } else break; }
            }
            for (int end = n - 1; end > 0; end--) {
                int t = a[0]; a[0] = a[end]; a[end] = t;
                int root = 0;
                while (2 * root + 1 < end) { int ch = 2 * root + 1; if (ch + 1 < end && a[ch] < a[ch + 1]) ch++; if (a[root] < a[ch]) { int t2 = a[root]; a[root] = a[ch]; a[ch] = t2; root = ch; } else break; }
            }
        }
    }
    bool moved = false;
    for (int i = 0; i < n; i++) moved |= (a[i] != b + i);
    if (moved) *nonidentity = 1;
}

// Permutation into cell-sorted storage order.  The fixed per-particle fields travel in one kernel whose pointer table is a kernel
// ARGUMENT (no device-side table, no host synchronisation for its lifetime); the species-indexed blocks (C, Q, data_fn, xx — any
// number of rows: the reference has no species limit, test/integration_tests/test_model.py:61-79 runs 51) are permuted row by row
// by a 2-D launch and swapped with their alternates.
#define PERM_MAX64 32
#define PERM_MAX32 8
struct PermTable {
    int n64, n32;
    const double *src64[PERM_MAX64];
    double *dst64[PERM_MAX64];
    const int *src32[PERM_MAX32];
    int *dst32[PERM_MAX32];
};
__global__ void k_permute(int N, const int *perm, const PermTable T) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int src = perm[p];
#pragma unroll
    for (int f = 0; f < PERM_MAX64; f++) if (f < T.n64) T.dst64[f][p] = T.src64[f][src];
#pragma unroll
    for (int f = 0; f < PERM_MAX32; f++) if (f < T.n32) T.dst32[f][p] = T.src32[f][src];
}
// rows of a species-major block: dst[r*N + p] = src[r*N + perm[p]], blockIdx.y = row
__global__ void k_permute_rows64(int N, const int *perm, const double *src, double *dst) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) dst[(size_t) blockIdx.y * N + p] = src[(size_t) blockIdx.y * N + perm[p]];
}
__global__ void k_permute_rows32(int N, const int *perm, const int *src, int *dst) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) dst[(size_t) blockIdx.y * N + p] = src[(size_t) blockIdx.y * N + perm[p]];
}

// Look-ahead displacement (moving domains with a Verlet skin).  After a step is complete, everything the NEXT predictor will do to
// the positions is known: x' = x + dt*((v + dt/2 F) + dt/2 Fbp) for non-solid particles (take_step1, simulate.cpp:68-79; v is only
// reassigned by the boundary conditions AFTER the position update).  out[0] = max |x' - xref|^2, out[1] = max |x' - x|^2,
// out[2] = max |x - xref|^2 (bit patterns of non-negative doubles, atomicMax).  The host reads them while the sSSA of the same step
// is still queued, and decides whether the next step keeps its candidate lists — no blocking read-back after the predictor and no
// "skin exceeded" failure mode: a step whose displacement would not fit simply rebuilds.
__global__ void k_lookahead(SsbView V, unsigned long long *out) {
    double a = 0.0, b = 0.0, c = 0.0;
    const double dt = V.dt;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < V.N; i += gridDim.x * blockDim.x) {     // persistent grid: 3 atomics per CTA
        const bool moves = V.solid[i] == 0;
        double pa = 0.0, pb = 0.0, pc = 0.0;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double x = V.x[d][i], xr = V.xref[d][i];
            double xn = x;
            if (moves) {
                const double v = V.v[d][i] + 0.5 * dt * V.F[d][i];
                const double vt = v + 0.5 * dt * V.Fbp[d][i];
                xn = x + dt * vt;
            }
            pa += (xn - xr) * (xn - xr); pb += (xn - x) * (xn - x); pc += (x - xr) * (x - xr);
        }
        a = fmax(a, pa); b = fmax(b, pb); c = fmax(c, pc);
    }
    __shared__ double sh[3][CORE_BLOCK / 32];
    for (int o = 16; o > 0; o >>= 1) {
        a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
        c = fmax(c, __shfl_xor_sync(0xffffffffu, c, o));
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; sh[2][threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double m = 0.0;
        for (int w = 0; w < CORE_BLOCK / 32; w++) m = fmax(m, sh[threadIdx.x][w]);
        if (m > 0.0) atomicMax(&out[threadIdx.x], (unsigned long long) __double_as_longlong(m));
    }
}

// neighbour search: query = live x_i, data = snapshot x0 (cell-sorted).  One thread per particle; the three
// x-adjacent cells of a (cy,cz) row form one contiguous particle range.
__global__ void k_search(SsbView V, CellGrid g, const int *cell_start, int *max_count, unsigned long long *total_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (i < V.N && !V.owned[i]) {                       // ghost copies of a neighbouring slab's particles: no sweep reads their lists
        V.nbr_count[i] = 0;
        if (V.solid_nbr) V.solid_nbr[i] = 0;
    } else if (i < V.N) {
        int any_solid = 0;
        const int N = V.N, dim = V.dim, cap = V.nbr_cap;
        const double h = V.h;
        const double h2 = __dmul_rn(h, h);             // ANNdist dist = system->h * system->h (particle.cpp:253)
        const double qx = V.x[0][i], qy = V.x[1][i], qz = V.x[2][i];
        int cx, cy, cz;
        cell_key(g, qx, qy, qz, cx, cy, cz);
        const int x_lo = max(cx - 1, 0), x_hi = min(cx + 1, g.n[0] - 1);
        for (int zz = max(cz - 1, 0); zz <= min(cz + 1, g.n[2] - 1); zz++) {
            for (int yy = max(cy - 1, 0); yy <= min(cy + 1, g.n[1] - 1); yy++) {
                const int row = (zz * g.n[1] + yy) * g.n[0];
                const int b = cell_start[row + x_lo];
                const int e = (row + x_hi + 1 < g.ncells) ? cell_start[row + x_hi + 1] : N;
                for (int j = b; j < e; j++) {
                    const double d2 = ssb_dist2(dim, qx, qy, qz, V.x0[0][j], V.x0[1][j], V.x0[2][j]);
                    // exact lists: kd_fix_rad_search.cpp:168-176 (ANN_ALLOW_SELF_MATCH = false) + particle.cpp:160-162;
                    // candidate lists (Verlet skin): everything within h*(1+skin), self included (the stale self-neighbour)
                    const bool take = V.filter ? (d2 <= V.search_h2) : ssb_in_range(d2, h, h2);
                    if (take) {
                        if (cnt < cap) V.nbr[(size_t) cnt * N + i] = j;
                        cnt++;
                        any_solid |= V.solid[j];
                    }
                }
            }
        }
        V.nbr_count[i] = min(cnt, cap);
        if (V.solid_nbr) V.solid_nbr[i] = any_solid != 0;      // lets the BVF sweep of bulk fluid stop early (k_finish)
    }
    int tot = cnt;
    for (int o = 16; o > 0; o >>= 1) { cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, o)); tot += __shfl_xor_sync(0xffffffffu, tot, o); }
    if ((threadIdx.x & 31) == 0 && cnt > 0) { atomicMax(max_count, cnt); atomicAdd(total_count, (unsigned long long) tot); }
}

__global__ void k_iota(int n, int *a) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = i; }
__global__ void k_copy64(int n, const double *src, double *dst) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) dst[i] = src[i]; }

// gather storage order -> id order (output staging and parity taps)
__global__ void k_unperm64(int N, const int *id, const double *src, double *dst, int stride, int offset) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) dst[(size_t) id[i] * stride + offset] = src[i];
}
__global__ void k_unperm32(int N, const int *id, const int *src, int *dst, int stride, int offset) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) dst[(size_t) id[i] * stride + offset] = src[i];
}
// scatter id order -> storage order (ssb_set_field: state handed over at a slab re-partition)
__global__ void k_perm_in64(int N, const int *id, const double *src, double *dst, int stride, int offset) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) dst[i] = src[(size_t) id[i] * stride + offset];
}
__global__ void k_perm_in32(int N, const int *id, const int *src, int *dst, int stride, int offset) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) dst[i] = src[(size_t) id[i] * stride + offset];
}
// neighbour list taps in id space
__global__ void k_nbr_count_by_id(SsbView V, long long *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V.N) return;
    int c = V.nbr_count[i];
    if (V.filter) {     // candidate list -> exact count
        const double xi0 = V.x[0][i], xi1 = V.x[1][i], xi2 = V.x[2][i];
        int e = 0;
        for (int k = 0; k < c; k++) {
            int j = V.nbr[(size_t) k * V.N + i];
            e += ssb_in_range(ssb_dist2(V.dim, xi0, xi1, xi2, V.x0[0][j], V.x0[1][j], V.x0[2][j]), V.h, __dmul_rn(V.h, V.h)) ? 1 : 0;
        }
        c = e;
    }
    out[V.id[i]] = c;
}
__global__ void k_nbr_export(SsbView V, const long long *ptr, int *idx, double *dist, double *dWdr, double *Dij) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V.N) return;
    const int N = V.N;
    const double alpha = ssb_alpha(V.dim, V.h);
    long long base = ptr[V.id[i]];
    const double xi0 = V.x[0][i], xi1 = V.x[1][i], xi2 = V.x[2][i];
    int o = 0;
    for (int k = 0; k < V.nbr_count[i]; k++) {
        int j = V.nbr[(size_t) k * N + i];
        double d2 = ssb_dist2(V.dim, xi0, xi1, xi2, V.x0[0][j], V.x0[1][j], V.x0[2][j]);
        if (V.filter && !ssb_in_range(d2, V.h, __dmul_rn(V.h, V.h))) continue;
        double r = sqrt(d2);
        idx[base + o] = V.id[j];
        dist[base + o] = r;
        dWdr[base + o] = ssb_dWdr(alpha, r, V.h);
        double rho_i, rho_j;
        ssb_search_rho(V, i, j, rho_i, rho_j);
        Dij[base + o] = V.Dij ? V.Dij[(size_t) k * N + i] : ssb_Dij(d2, r, V.h, V.mass[i], V.mass[j], rho_i, rho_j);
        o++;
    }
}

// earliest pending sSSA event over all chunks (non-negative doubles order like their bit patterns)
__global__ void k_min_time(int nchunks, const double *blk_tmin, unsigned long long *out_bits) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    double v = (c < nchunks) ? blk_tmin[c] : INFINITY;
    if (!(v >= 0.0)) v = 0.0;
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) atomicMin(out_bits, (unsigned long long) __double_as_longlong(v));
}

// ---- slab decomposition: halo pack / unpack (ids are particle ids of THIS rank's model; slot_of_id maps to storage) ----
__global__ void k_slot_of_id(int N, const int *id, int *slot_of_id) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) slot_of_id[id[p]] = p;
}
// group 0 (after the force sweep): F[3] Fbp[3] Frho Q[Sc]   group 1 (after the corrector): rho_new
// group 2 (after the BVF sweep): v[3] bvf_phi                group 3 (initial consistency): rho
__device__ __forceinline__ int halo_width(int group, int Sc) { return group == 0 ? 7 + Sc : (group == 2 ? 4 : 1); }
__global__ void k_halo_pack(SsbView V, int group, const int *ids, int n, const int *slot_of_id, double *out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int p = slot_of_id[ids[t]];
    const int W = halo_width(group, V.Sc);
    double *o = out + (size_t) t * W;
    if (group == 0) {
        for (int d = 0; d < 3; d++) { o[d] = V.F[d][p]; o[3 + d] = V.Fbp[d][p]; }
        o[6] = V.Frho[p];
        for (int s = 0; s < V.Sc; s++) o[7 + s] = V.Q[(size_t) s * V.N + p];
    } else if (group == 1) {
        o[0] = V.rho_new[p];
    } else if (group == 2) {
        for (int d = 0; d < 3; d++) o[d] = V.v[d][p];
        o[3] = V.bvf[p];
    } else {
        o[0] = V.rho[p];
    }
}
__global__ void k_halo_unpack(SsbView V, int group, const int *ids, int n, const int *slot_of_id, const double *in) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int p = slot_of_id[ids[t]];
    const int W = halo_width(group, V.Sc);
    const double *o = in + (size_t) t * W;
    if (group == 0) {
        for (int d = 0; d < 3; d++) { V.F[d][p] = o[d]; V.Fbp[d][p] = o[3 + d]; }
        V.Frho[p] = o[6];
        for (int s = 0; s < V.Sc; s++) V.Q[(size_t) s * V.N + p] = o[7 + s];
    } else if (group == 1) {
        V.rho_new[p] = o[0];
    } else if (group == 2) {
        for (int d = 0; d < 3; d++) V.v[d][p] = o[d];
        V.bvf[p] = o[3];
    } else {
        V.rho[p] = o[0];
    }
}
// molecules that jumped into ghost voxels during the last sSSA window: read-and-clear on the sender ...
__global__ void k_inbox_pack(SsbView V, int buf, const int *ids, int n, const int *slot_of_id, unsigned *out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int p = slot_of_id[ids[t]];
    for (int s = 0; s < V.Sd; s++) {
        unsigned *q = &V.inbox[buf][(size_t) s * V.N + p];
        out[(size_t) t * V.Sd + s] = *q;
        *q = 0u;
    }
    V.inbox_src[buf][p] = 0;
}
// ... and add on the owner, which sees them as ordinary mail at the start of its next window
__global__ void k_inbox_add(SsbView V, int buf, const int *ids, int n, const int *slot_of_id, const unsigned *in, int block) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int p = slot_of_id[ids[t]];
    bool any = false;
    for (int s = 0; s < V.Sd; s++) {
        const unsigned v = in[(size_t) t * V.Sd + s];
        if (v) { atomicAdd(&V.inbox[buf][(size_t) s * V.N + p], v); any = true; }
    }
    if (any) V.blk_mail[buf][p / block] = 1;
}

// ---- slab decomposition, native transport: pack kernels that WRITE INTO THE NEIGHBOUR'S RECEIVE WINDOW over NVLink (peer-mapped
// memory), a sequence flag per channel raised by the last CTA of the pack kernel, a one-warp wait kernel, unpack from the local
// window.  Everything is ordered by the engine stream; the host never waits for a message.  Messages are column-major
// (buf[f*n + t]) so that a warp stores / loads whole 256-byte runs.
#define HALO_NCH 4                   // channels 0..2 = field groups 0..2, 3 = sSSA inbox (compacted entries)
#define HALO_CH_INBOX 3
#define HALO_HDR_BYTES 1024          // window header: flag[ch] at 64*ch, inbox entry count of parity q at 512 + 64*q
#define HALO_TIMEOUT_NS 10000000000ull
#define SSB_BOARD_NCH 3              // scalar all-reduce boards: 0 max Ddiag (max), 1 earliest pending event (min), 2 step displacement (max)
#define SSB_BOARD_MAXW 16
struct HaloPackSide { const int *ids; int n; char *peer_buf; unsigned long long *peer_flag; unsigned long long *peer_count; };
struct HaloPackArgs { HaloPackSide side[2]; unsigned long long seq; unsigned *done; unsigned *icount; };
struct HaloUnpackSide { const int *ids; int n; const char *buf; const unsigned long long *count; };
struct HaloUnpackArgs { HaloUnpackSide side[2]; };

__device__ __forceinline__ unsigned long long ssb_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned long long ssb_ld_flag(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
// last CTA of a pack kernel: everything every CTA wrote to the peer is fenced, raise the flags (and the inbox entry counts)
__device__ __forceinline__ void halo_publish(const HaloPackArgs &A, bool with_counts) {
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) s_last = (atomicAdd(A.done, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last && threadIdx.x < 2) {
        const HaloPackSide &S = A.side[threadIdx.x];
        if (S.peer_flag) {
            if (with_counts) { *(volatile unsigned long long *) S.peer_count = (unsigned long long) atomicExch(&A.icount[threadIdx.x], 0u); }
            __threadfence_system();
            *(volatile unsigned long long *) S.peer_flag = A.seq;
        }
        if (threadIdx.x == 0) *A.done = 0u;
    }
}
__global__ void k_halo_send(SsbView V, int group, HaloPackArgs A, const int *slot_of_id) {
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int sd = (t0 < A.side[0].n) ? 0 : 1;
    const int t = sd ? t0 - A.side[0].n : t0;
    const HaloPackSide &S = A.side[sd];
    if (t < S.n) {
        const int p = slot_of_id[S.ids[t]];
        double *o = (double *) S.peer_buf;
        const size_t n = (size_t) S.n;
        if (group == 0) {
            for (int d = 0; d < 3; d++) { o[d * n + t] = V.F[d][p]; o[(3 + d) * n + t] = V.Fbp[d][p]; }
            o[6 * n + t] = V.Frho[p];
            for (int sp = 0; sp < V.Sc; sp++) o[(7 + sp) * n + t] = V.Q[(size_t) sp * V.N + p];
        } else if (group == 1) {
            o[t] = V.rho_new[p];
        } else {
            for (int d = 0; d < 3; d++) o[d * n + t] = V.v[d][p];
            o[3 * n + t] = V.bvf[p];
        }
    }
    halo_publish(A, false);
}
__global__ void k_halo_recv(SsbView V, int group, HaloUnpackArgs A, const int *slot_of_id) {
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int sd = (t0 < A.side[0].n) ? 0 : 1;
    const int t = sd ? t0 - A.side[0].n : t0;
    const HaloUnpackSide &S = A.side[sd];
    if (t >= S.n) return;
    const int p = slot_of_id[S.ids[t]];
    const double *o = (const double *) S.buf;      // written by the neighbour's pack kernel: read through L2 (__ldcg), never from L1
    const size_t n = (size_t) S.n;
    if (group == 0) {
        for (int d = 0; d < 3; d++) { V.F[d][p] = __ldcg(&o[d * n + t]); V.Fbp[d][p] = __ldcg(&o[(3 + d) * n + t]); }
        V.Frho[p] = __ldcg(&o[6 * n + t]);
        for (int sp = 0; sp < V.Sc; sp++) V.Q[(size_t) sp * V.N + p] = __ldcg(&o[(7 + sp) * n + t]);
    } else if (group == 1) {
        V.rho_new[p] = __ldcg(&o[t]);
    } else {
        for (int d = 0; d < 3; d++) V.v[d][p] = __ldcg(&o[d * n + t]);
        V.bvf[p] = __ldcg(&o[3 * n + t]);
    }
}
// one warp: lane s < 2 waits for the flag of side s.  A peer that never arrives raises SSB_ERR_HALO instead of hanging the GPU.
__global__ void k_halo_wait(const unsigned long long *flag0, const unsigned long long *flag1, unsigned long long seq, int *err_flag) {
    const unsigned long long *f = threadIdx.x == 0 ? flag0 : (threadIdx.x == 1 ? flag1 : nullptr);
    if (f && *(volatile int *) err_flag != 8) {          // (a run that already lost a message fails fast instead of timing out again and again)
        const unsigned long long t0 = ssb_globaltimer();
        while (ssb_ld_flag(f) < seq) {
            __nanosleep(100);
            if (ssb_globaltimer() - t0 > HALO_TIMEOUT_NS) {
                if (atomicCAS(err_flag, 0, 8 /*SSB_ERR_HALO*/) == 0) { err_flag[1] = 100 + (int) threadIdx.x; err_flag[2] = (int) seq; err_flag[3] = (int) ssb_ld_flag(f); }
                break;
            }
        }
    }
    __threadfence_system();
}
// molecules that jumped into ghost voxels in the last sSSA window travel as (row in the exchange list, species, count) entries: the
// pack kernel reads-and-clears the ghosts' inboxes and appends only what is non-zero — a handful of entries instead of a dense
// [ghosts x S_d] array per window
__global__ void k_inbox_send(SsbView V, int buf, HaloPackArgs A, const int *slot_of_id) {
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int sd = (t0 < A.side[0].n) ? 0 : 1;
    const int t = sd ? t0 - A.side[0].n : t0;
    const HaloPackSide &S = A.side[sd];
    if (t < S.n) {
        const int p = slot_of_id[S.ids[t]];
        for (int sp = 0; sp < V.Sd; sp++) {
            unsigned *q = &V.inbox[buf][(size_t) sp * V.N + p];
            const unsigned c = *q;
            if (c) {
                *q = 0u;
                const unsigned e = atomicAdd(&A.icount[sd], 1u);
                ((uint4 *) S.peer_buf)[e] = make_uint4((unsigned) t, (unsigned) sp, c, 0u);
            }
        }
        V.inbox_src[buf][p] = 0;
    }
    halo_publish(A, true);
}
__global__ void k_inbox_recv(SsbView V, int buf, HaloUnpackArgs A, const int *slot_of_id, int block) {
    for (int sd = 0; sd < 2; sd++) {
        const HaloUnpackSide &S = A.side[sd];
        if (!S.count) continue;
        const unsigned n = (unsigned) ssb_ld_flag(S.count);
        for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
            const uint4 m = ((const uint4 *) S.buf)[e];
            const int p = slot_of_id[S.ids[m.x]];
            atomicAdd(&V.inbox[buf][(size_t) m.y * V.N + p], m.z);
            V.blk_mail[buf][p / block] = 1;
        }
    }
}
// scalar all-reduce over peer-mapped boards: every rank writes its value into slot `me` of EVERY rank's board, then each rank
// reduces its own board once all slots carry the sequence number.  Values are bit patterns of non-negative doubles (they order
// like unsigned integers).  Board entry of (channel, parity, rank) = {value, flag}.
__device__ __forceinline__ size_t board_slot(int ch, int parity, int r) { return ((size_t) (ch * 2 + parity) * SSB_BOARD_MAXW + r) * 2; }
struct BoardPeers { unsigned long long *b[SSB_BOARD_MAXW]; };
__global__ void k_board_post(BoardPeers P, int world, int me, int ch, unsigned long long seq, const unsigned long long *src) {
    const int r = threadIdx.x;
    if (r >= world) return;
    const unsigned long long v = *src;
    unsigned long long *slot = P.b[r] + board_slot(ch, (int) (seq & 1ull), me);
    *(volatile unsigned long long *) slot = v;
    __threadfence_system();
    *(volatile unsigned long long *) (slot + 1) = seq;
}
__global__ void k_board_reduce(const unsigned long long *board, int world, int ch, unsigned long long seq, int take_min,
                               unsigned long long *out_dev, int *err_flag) {
    const int r = threadIdx.x;
    unsigned long long v = take_min ? ~0ull : 0ull;
    if (r < world) {
        const unsigned long long *slot = board + board_slot(ch, (int) (seq & 1ull), r);
        const unsigned long long t0 = ssb_globaltimer();
        while (ssb_ld_flag(slot + 1) < seq && *(volatile int *) err_flag != 8) {
            __nanosleep(100);
            if (ssb_globaltimer() - t0 > HALO_TIMEOUT_NS) {
                if (atomicCAS(err_flag, 0, 8 /*SSB_ERR_HALO*/) == 0) { err_flag[1] = 200 + 10 * ch + r; err_flag[2] = (int) seq; err_flag[3] = (int) ssb_ld_flag(slot + 1); }
                break;
            }
        }
        __threadfence_system();
        v = ssb_ld_flag(slot);
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = take_min ? (w < v ? w : v) : (w > v ? w : v);
    }
    if (r == 0) *out_dev = v;
}

// =====================================================================================================
// host side
// =====================================================================================================
struct OutputJob {
    int valid = 0;
    unsigned step = 0;
    unsigned file_index = 0;
    int rdme_initialized = 0;
    int write_file = 1;
    int write_bin = 0;
    std::string dir;
    cudaEvent_t ready = nullptr;
    // pinned staging (id order)
    double *x = nullptr;     // N*3
    double *v = nullptr;     // N*3
    double *scal = nullptr;  // 4*N: rho, mass, bvf, nu
    int *type = nullptr;     // N
    double *C = nullptr;     // Sc*N (species-major)
    unsigned *xx = nullptr;  // Sd*N (species-major)
};

struct SlabComm;
struct ssb_handle {
    ssb_model m;       // scalars (pointers inside are NOT retained)
    int N = 0, S = 0, R = 0;
    // host copies of the initial condition
    std::vector<double> hx, hnu, hmass, hrho, hdata_fn, hdmat;
    std::vector<int> htype, hsolid, howned, hgid;
    std::vector<unsigned> hu0, hout_steps;
    std::vector<std::string> species_names;
    // device
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    std::vector<void *> allocs;
    SsbView V;
    // permuted field tables (current / alternate buffers)
    std::vector<double **> f64_slots;  // addresses of the view pointers
    std::vector<double *> f64_alt;
    std::vector<int **> i32_slots;
    std::vector<int *> i32_alt;
    double *C_alt = nullptr, *Q_alt = nullptr, *df_alt = nullptr;   // alternates of the species-major blocks (swapped by a permutation)
    unsigned *xx_alt = nullptr;
    // pinned scalars the device reports through (async copies + events instead of blocking read-backs):
    //   [0..2] look-ahead displacement (k_lookahead)  [3] max Ddiag of this step's force sweep  [4..7] slab scalar reductions
    unsigned long long *pin = nullptr;
    unsigned long long *d_look = nullptr;     // device side of pin[0..2]
    cudaEvent_t ev_look = nullptr, ev_maxd = nullptr;
    int look_valid = 0;           // pin[0..2] describe the step that is about to run
    // cell list
    CellGrid grid;
    int *d_key = nullptr, *d_cell_count = nullptr, *d_cell_start = nullptr, *d_cursor = nullptr, *d_perm = nullptr;
    int *d_tile_sums = nullptr, *d_flags = nullptr;  // d_flags[0]=nonidentity [1]=max nbr count
    unsigned long long *d_maxbits = nullptr;
    double *d_stage = nullptr;   // device staging for output/taps (id order)
    size_t stage_bytes = 0;
    double *d_dmat = nullptr;
    double *init_f64 = nullptr;   // pinned
    int *init_i32 = nullptr;      // pinned
    double *rho_buf[2] = {nullptr, nullptr};
    double *d_rho_pre = nullptr;  // step-0 densities before the predictor / BCs (models whose BC assigns rho; SsbView::rho_pre)
    // model unit
    void *unit_dl = nullptr;
    const SsbModelUnit *unit = nullptr;
    // run state
    unsigned current_step = 0;
    int rdme_initialized = 0;
    uint64_t seed = 0, epoch = 0;
    int inbox_buf = 0;
    double tau = 0.0;
    long long nwin = 1;
    int nbr_valid = 0;
    int ddiag_fresh = 0;
    double skin = 0.0;            // Verlet skin as a fraction of h (moving domains; 0 = exact lists rebuilt every step)
    int lists_valid = 0;          // candidate lists + storage order from an earlier step are still usable
    double disp_prev = 0.0;       // max displacement from xref after the previous step
    double step_disp_max = 0.0;   // largest single-step displacement seen in this trajectory
    double disp_step_next = 0.0;  // single-step displacement of the step that is about to run (k_lookahead out[1])
    double disp_build = 0.0;      // the same quantity at the step that built the current candidate lists
    int64_t rebuilds = 0;
    int skin_chosen = 0;
    // slab decomposition support
    int *d_slot_of_id = nullptr;  // particle id -> storage slot (rebuilt lazily after a permutation)
    int slot_dirty = 1;
    struct SlabComm *slab = nullptr;   // native transport (ssb_slab_*): receive windows, peer mappings, boards
    cudaEvent_t mark_a = nullptr, mark_b = nullptr, sync_ev = nullptr;
    int static_cached = 0;        // static domain: storage order, neighbour lists, coefficients and Ddiag survive ssb_reset
    int *d_static_perm = nullptr; // slot -> particle id of the cached storage order
    double max_ddiag_cached = 0.0;
    int64_t launches = 0, windows = 0, h2d_bytes = 0, d2h_bytes = 0;
    int64_t total_reactions = 0, total_diffusion = 0;
    double step_seconds = 0.0;
    std::atomic<int> cancel{0};
    std::string err;
    // per-category device timing (ssb_profile): CUDA event pairs on the engine stream around each launch group
    int profile = 0;
    std::vector<cudaEvent_t> ev_a, ev_b;
    std::vector<int> ev_cat;
    double cat_ms[SSB_NCAT] = {0};
    int64_t cat_launches[SSB_NCAT] = {0};
    // output
    OutputJob jobs[2];
    int job_cursor = 0;
    std::thread writer;
    std::mutex mu;
    std::condition_variable cv;
    int writer_pending[2] = {0, 0};
    int writer_quit = 0;
    int writer_error = 0;
};

static int fail(ssb_handle *h, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(h, SSB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T>
static cudaError_t dalloc(ssb_handle *h, T **p, size_t count) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, (count > 0 ? count : 1) * sizeof(T));
    if (e == cudaSuccess) { h->allocs.push_back(q); *p = (T *) q; }
    return e;
}

static inline unsigned gridN(int n) { return (unsigned) ((n + CORE_BLOCK - 1) / CORE_BLOCK); }

// Host wait for the engine stream through a BLOCKING-SYNC event: the calling thread sleeps instead of spinning, so many
// engine handles (ensemble lanes, one host thread each) do not fight over the host cores while their kernels run.
static cudaError_t ssb_sync(ssb_handle *h) {
    if (!h->sync_ev) {
        cudaError_t e = cudaEventCreateWithFlags(&h->sync_ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(h->sync_ev, h->stream);
    if (e != cudaSuccess) return e;
    // short waits (a scalar read-back behind a kernel that is about to finish) are polled: waking a sleeping thread costs far more
    // than the wait itself (measured: ~70 us per blocking wait on a 1.6 ms moving-domain step)
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        e = cudaEventQuery(h->sync_ev);
        if (e != cudaErrorNotReady) return e;
        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(40)) break;
    }
    return cudaEventSynchronize(h->sync_ev);
}

// ---- per-category device timing -------------------------------------------------------------------------
static void prof_harvest(ssb_handle *h) {
    if (h->ev_cat.empty()) return;
    ssb_sync(h);
    for (size_t k = 0; k < h->ev_cat.size(); k++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev_a[k], h->ev_b[k]) == cudaSuccess) h->cat_ms[h->ev_cat[k]] += ms;
    }
    h->ev_cat.clear();
}
// kernel categories as NVTX ranges (nsys / ncu --nvtx timelines; SURVEY.md section 5 "tracing") and, with ssb_profile(1), as CUDA-event
// pairs on the engine stream
static const char *const CAT_NAMES[SSB_NCAT] = {"ssb:cell_list", "ssb:predictor", "ssb:neighbour_search", "ssb:force_sweep", "ssb:corrector",
                                                "ssb:finish_bvf", "ssb:diffusion_matrix", "ssb:rdme_init", "ssb:sssa_windows", "ssb:output_staging"};
static int prof_begin(ssb_handle *h, int cat, int nlaunch) {
    nvtxRangePushA(CAT_NAMES[cat]);
    if (!h->profile) return -1;
    if (h->ev_cat.size() >= 4096) prof_harvest(h);
    size_t k = h->ev_cat.size();
    if (k >= h->ev_a.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        h->ev_a.push_back(a); h->ev_b.push_back(b);
    }
    h->ev_cat.push_back(cat);
    h->cat_launches[cat] += nlaunch;
    cudaEventRecord(h->ev_a[k], h->stream);
    return (int) k;
}
static void prof_end(ssb_handle *h, int slot) { if (slot >= 0) cudaEventRecord(h->ev_b[slot], h->stream); nvtxRangePop(); }

// ----------------------------------------------------------------------------------------------------
// VTK writer (host thread) — byte format of E/src/output.cpp:104-229
// ----------------------------------------------------------------------------------------------------
struct Buf {
    std::vector<char> d;
    size_t n = 0;
    void reserve(size_t extra) { if (n + extra > d.size()) d.resize((n + extra) * 2 + 4096); }
    void putf(const char *fmt, ...) {
        reserve(512);
        va_list ap;
        va_start(ap, fmt);
        n += (size_t) vsnprintf(d.data() + n, 512, fmt, ap);
        va_end(ap);
    }
};

// "%lf " with exactly printf's rounding: fast path through integer arithmetic when the scaled value is far from a
// rounding tie, otherwise snprintf.
static inline void put_lf(Buf &b, double v) {
    b.reserve(400);
    char *p = b.d.data() + b.n;
    double a = fabs(v);
    if (a < 1e9 && a == a) {
        double scaled = a * 1e6;
        double fl = floor(scaled);
        double frac = scaled - fl;
        // a*1e6 carries at most ~1 ulp of error (~1.2e-7 at 1e9*1e6 scale is too coarse, so restrict the band)
        double tol = scaled * 4.5e-16 + 1e-300;
        if (fabs(frac - 0.5) > tol * 4 + 1e-9) {   // only a near-tie can round differently from printf
            unsigned long long q = (unsigned long long) fl + (frac > 0.5 ? 1ull : 0ull);
            unsigned long long ip = q / 1000000ull, fp = q % 1000000ull;
            char tmp[32];
            int len = 0;
            if (signbit(v)) *p++ = '-';
            if (ip == 0) tmp[len++] = '0';
            while (ip) { tmp[len++] = (char) ('0' + ip % 10); ip /= 10; }
            while (len) *p++ = tmp[--len];
            *p++ = '.';
            for (int k = 5; k >= 0; k--) { p[k] = (char) ('0' + fp % 10); fp /= 10; }
            p += 6;
            *p++ = ' ';
            b.n = (size_t) (p - b.d.data());
            return;
        }
    }
    b.n += (size_t) snprintf(p, 400, "%lf ", v);
}

static inline void put_u(Buf &b, unsigned v) {
    b.reserve(16);
    char *p = b.d.data() + b.n;
    char tmp[12];
    int len = 0;
    if (v == 0) tmp[len++] = '0';
    while (v) { tmp[len++] = (char) ('0' + v % 10); v /= 10; }
    while (len) *p++ = tmp[--len];
    *p++ = ' ';
    b.n = (size_t) (p - b.d.data());
}

// what the two file writers need besides the staged arrays: sizes, bounding box, species names (a handle's, or a caller's for
// snapshots assembled on the host: ssb_write_snapshot)
struct SnapMeta {
    int np, Sc, Sd;
    double xlo, xhi, ylo, yhi, zlo, zhi;
    const std::vector<std::string> *names;
};

static int write_vtk_impl(const SnapMeta &M, const OutputJob &J) {
    const int np = M.np;
    const int Sc = M.Sc, Sd = M.Sd;
    char filename[4096];
    if (J.step == 0 && J.file_index == 0) {
        snprintf(filename, sizeof(filename), "%s/output0_boundingBox.vtk", J.dir.c_str());
        FILE *fp = fopen(filename, "w+");
        if (!fp) return SSB_ERR_IO;
        fprintf(fp, "# vtk DataFile Version 4.1\n");
        fprintf(fp, "Generated by ssa_sdpd\n");
        fprintf(fp, "ASCII\n");
        fprintf(fp, "DATASET RECTILINEAR_GRID\n");
        fprintf(fp, "DIMENSIONS 2 2 2\n");
        fprintf(fp, "X_COORDINATES 2 double\n");
        fprintf(fp, "%lf %lf\n", M.xlo, M.xhi);
        fprintf(fp, "Y_COORDINATES 2 double\n");
        fprintf(fp, "%lf %lf\n", M.ylo, M.yhi);
        fprintf(fp, "Z_COORDINATES 2 double\n");
        fprintf(fp, "%lf %lf\n", M.zlo, M.zhi);
        fclose(fp);
    }
    snprintf(filename, sizeof(filename), "%s/output%u.vtk", J.dir.c_str(), J.file_index);
    FILE *fp = fopen(filename, "w+");
    if (!fp) return SSB_ERR_IO;
    Buf b;
    b.d.resize(65536);
    auto flush = [&]() { if (b.n) fwrite(b.d.data(), 1, b.n, fp); b.n = 0; };
    // One section of the file = header line (already in `b`) + np formatted items.  Large snapshots are formatted by several host
    // threads over contiguous particle ranges (multiples of 9, so every range starts at a line start for both the 3-per-line and
    // the 9-per-line layouts) into private buffers that are written in order: the text of a 1 M-particle snapshot is ~150 MB and
    // single-threaded formatting, not the GPU, bounds the drop-in path.
    int nthreads = (np >= 200000) ? (int) std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())) : 1;
    if (const char *e = getenv("SSB_VTK_THREADS")) nthreads = std::max(1, std::min(64, atoi(e)));
    auto section = [&](auto &&fmt_item) {
        flush();
        if (nthreads == 1) {
            for (int i = 0; i < np; i++) { fmt_item(b, i); if (b.n > (1u << 22)) flush(); }
            flush();
            return;
        }
        const int per = ((np + nthreads - 1) / nthreads + 8) / 9 * 9;
        std::vector<Buf> bufs((size_t) nthreads);
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) {
            th.emplace_back([&, t]() {
                const int lo = std::min(np, t * per), hi = std::min(np, lo + per);
                Buf &q = bufs[(size_t) t];
                q.d.resize((size_t) (hi - lo) * 48 + 4096);
                for (int i = lo; i < hi; i++) fmt_item(q, i);
            });
        }
        for (auto &t : th) t.join();
        for (auto &q : bufs) if (q.n) fwrite(q.d.data(), 1, q.n, fp);
    };
    b.putf("# vtk DataFile Version 4.1\n");
    b.putf("Generated by SpatialPy\n");
    b.putf("ASCII\n");
    b.putf("DATASET POLYDATA\n");
    b.putf("POINTS %i float\n", np);
    section([&](Buf &q, int i) {
        q.reserve(128);
        q.n += (size_t) snprintf(q.d.data() + q.n, 128, "%.10e %.10e %.10e ", J.x[i * 3], J.x[i * 3 + 1], J.x[i * 3 + 2]);
        if ((i + 1) % 3 == 0) q.d[q.n++] = '\n';
    });
    b.putf("\n");
    b.putf("VERTICES %i %i\n", np, 2 * np);
    section([&](Buf &q, int i) {
        q.reserve(32);
        q.d[q.n++] = '1'; q.d[q.n++] = ' ';
        put_u(q, (unsigned) i);
        q.d[q.n - 1] = '\n';
    });
    b.putf("\n");
    b.putf("POINT_DATA %i\n", np);
    int num_fields = 7;
    if (J.rdme_initialized) num_fields += Sd;      // output.cpp:151-154: undercounts in output0
    if (Sc > 0) num_fields += Sc;
    b.putf("FIELD FieldData %i\n", num_fields);
    b.putf("id 1 %i int\n", np);
    section([&](Buf &q, int i) { put_u(q, (unsigned) i); if ((i + 1) % 9 == 0) { q.reserve(2); q.d[q.n++] = '\n'; } });
    b.putf("\n");
    b.putf("type 1 %i int\n", np);
    section([&](Buf &q, int i) { put_u(q, (unsigned) J.type[i]); if ((i + 1) % 9 == 0) { q.reserve(2); q.d[q.n++] = '\n'; } });
    b.putf("\n");
    b.putf("v 3 %i double\n", np);
    section([&](Buf &q, int i) {
        put_lf(q, J.v[i * 3]); put_lf(q, J.v[i * 3 + 1]); put_lf(q, J.v[i * 3 + 2]);
        if ((i + 1) % 3 == 0) { q.reserve(2); q.d[q.n++] = '\n'; }
    });
    b.putf("\n");
    const char *scal_names[4] = {"rho", "mass", "bvf_phi", "nu"};
    for (int f = 0; f < 4; f++) {
        b.putf("%s 1 %i double\n", scal_names[f], np);
        const double *a = J.scal + (size_t) f * np;
        section([&](Buf &q, int i) { put_lf(q, a[i]); if ((i + 1) % 9 == 0) { q.reserve(2); q.d[q.n++] = '\n'; } });
        b.putf("\n");
    }
    for (int s = 0; s < Sc; s++) {
        b.putf("C[%s] 1 %i double\n", (*M.names)[(size_t) s].c_str(), np);
        const double *a = J.C + (size_t) s * np;
        section([&](Buf &q, int i) { put_lf(q, a[i]); if ((i + 1) % 9 == 0) { q.reserve(2); q.d[q.n++] = '\n'; } });
        b.putf("\n");
    }
    for (int s = 0; s < Sd; s++) {
        b.putf("D[%s] 1 %i int\n", (*M.names)[(size_t) s].c_str(), np);
        const unsigned *a = J.xx + (size_t) s * np;
        section([&](Buf &q, int i) { put_u(q, a[i]); if ((i + 1) % 9 == 0) { q.reserve(2); q.d[q.n++] = '\n'; } });
        b.putf("\n");
    }
    flush();
    fclose(fp);
    return 0;
}

// Binary side-store (SSB_FLAG_BINARY_STORE): outputN.ssb = 8-byte magic, u64 length of an ASCII JSON header, the header padded with
// blanks to a multiple of 64 bytes, then the raw little-endian arrays in id order:
//   x f64[np*3]  v f64[np*3]  rho,mass,bvf_phi,nu f64[4*np]  C f64[Sc*np]  type i32[np]  D u32[Sd*np]
static int write_bin_impl(const SnapMeta &M, const OutputJob &J) {
    const size_t np = (size_t) M.np;
    const int Sc = M.Sc, Sd = M.Sd;
    char filename[4096];
    snprintf(filename, sizeof(filename), "%s/output%u.ssb", J.dir.c_str(), J.file_index);
    FILE *fp = fopen(filename, "wb");
    if (!fp) return SSB_ERR_IO;
    std::string hdr = "{\"np\": " + std::to_string(np) + ", \"Sc\": " + std::to_string(Sc) + ", \"Sd\": " + std::to_string(Sd) +
                      ", \"step\": " + std::to_string(J.step) + ", \"rdme_initialized\": " + std::to_string(J.rdme_initialized) + ", \"species\": [";
    const int ns = std::max(Sc, Sd);
    for (int s = 0; s < ns; s++) hdr += std::string(s ? ", " : "") + "\"" + (*M.names)[(size_t) s] + "\"";
    hdr += "]}";
    while ((16 + hdr.size()) % 64) hdr += ' ';
    const unsigned long long hl = hdr.size();
    size_t ok = fwrite("SSBOUT1\0", 1, 8, fp) == 8;
    ok &= fwrite(&hl, 8, 1, fp) == 1;
    ok &= fwrite(hdr.data(), 1, hdr.size(), fp) == hdr.size();
    ok &= fwrite(J.x, sizeof(double), 3 * np, fp) == 3 * np;
    ok &= fwrite(J.v, sizeof(double), 3 * np, fp) == 3 * np;
    ok &= fwrite(J.scal, sizeof(double), 4 * np, fp) == 4 * np;
    if (Sc > 0) ok &= fwrite(J.C, sizeof(double), (size_t) Sc * np, fp) == (size_t) Sc * np;
    ok &= fwrite(J.type, sizeof(int), np, fp) == np;
    if (Sd > 0) ok &= fwrite(J.xx, sizeof(unsigned), (size_t) Sd * np, fp) == (size_t) Sd * np;
    ok &= fclose(fp) == 0;
    return ok ? 0 : SSB_ERR_IO;
}

static SnapMeta meta_of(const ssb_handle *h) {
    return SnapMeta{h->N, h->V.Sc, h->V.Sd, h->m.xlo, h->m.xhi, h->m.ylo, h->m.yhi, h->m.zlo, h->m.zhi, &h->species_names};
}
static int write_vtk(ssb_handle *h, const OutputJob &J) { return write_vtk_impl(meta_of(h), J); }
static int write_bin(ssb_handle *h, const OutputJob &J) { return write_bin_impl(meta_of(h), J); }

static void writer_main(ssb_handle *h) {
    cudaSetDevice(h->device);
    int next = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(h->mu);
            h->cv.wait(lk, [&] { return h->writer_pending[next] || h->writer_quit; });
            if (!h->writer_pending[next] && h->writer_quit) return;
        }
        OutputJob &J = h->jobs[next];
        cudaEventSynchronize(J.ready);
        int rc = J.write_file ? write_vtk(h, J) : 0;
        if (!rc && J.write_bin) rc = write_bin(h, J);
        {
            std::lock_guard<std::mutex> lk(h->mu);
            if (rc) h->writer_error = rc;
            h->writer_pending[next] = 0;
        }
        h->cv.notify_all();
        next ^= 1;
    }
}

// ----------------------------------------------------------------------------------------------------
// cell grid
// ----------------------------------------------------------------------------------------------------
static void setup_grid(ssb_handle *h) {
    const int N = h->N;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < N; i++)
        for (int d = 0; d < 3; d++) { double v = h->hx[(size_t) i * 3 + d]; if (v < lo[d]) lo[d] = v; if (v > hi[d]) hi[d] = v; }
    CellGrid &g = h->grid;
    g.dim = h->m.dimension;
    long long total = 1;
    const double hh = h->m.h * (1.0 + h->skin);       // cell edge >= candidate radius
    for (int d = 0; d < 3; d++) {
        double L = hi[d] - lo[d];
        int n = 1;
        if (d < g.dim && L > 0 && hh > 0) {
            double q = floor(L / hh);
            n = (q < 1) ? 1 : (q > 4096 ? 4096 : (int) q);
        }
        g.n[d] = n;
        total *= n;
    }
    // keep the cell table proportionate to the particle count (sparse / elongated domains)
    long long cap = (long long) 4 * N + 1024;
    while (total > cap) {
        int big = 0;
        for (int d = 1; d < 3; d++) if (g.n[d] > g.n[big]) big = d;
        total /= g.n[big];
        g.n[big] = (g.n[big] + 1) / 2;
        total *= g.n[big];
    }
    for (int d = 0; d < 3; d++) {
        double L = hi[d] - lo[d];
        g.lo[d] = lo[d];
        // cell edge L/n >= h by construction; the last cell is closed by clamping
        g.inv_cell[d] = (g.n[d] > 1 && L > 0) ? (double) g.n[d] / L : 0.0;
    }
    g.ncells = (int) total;
}

// ----------------------------------------------------------------------------------------------------
// C-ABI
// ----------------------------------------------------------------------------------------------------
extern "C" int ssb_abi_version(void) { return SSB_ABI_VERSION; }

extern "C" int ssb_device_count(int *count) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (count) *count = (e == cudaSuccess) ? c : 0;
    return e == cudaSuccess ? SSB_OK : SSB_ERR_CUDA;
}

extern "C" const char *ssb_last_error(ssb_handle *h) { return h ? h->err.c_str() : "null handle"; }

static void register_f64(ssb_handle *h, double **slot) { h->f64_slots.push_back(slot); }
static void register_i32(ssb_handle *h, int **slot) { h->i32_slots.push_back(slot); }

extern "C" int ssb_create(const ssb_model *m, ssb_handle **out) {
    if (!m || !out) return SSB_ERR_ARG;
    *out = nullptr;
    if (m->abi_version != SSB_ABI_VERSION) return SSB_ERR_ARG;
    if (m->n_particles <= 0 || m->n_particles > 2000000000LL) return SSB_ERR_ARG;
    if (m->dimension < 1 || m->dimension > 3) return SSB_ERR_ARG;
    ssb_handle *h = new ssb_handle();
    *out = h;  // returned even on failure so the caller can read ssb_last_error, then ssb_destroy
    h->m = *m;
    const int N = h->N = (int) m->n_particles;
    const int S = h->S = (m->num_stoch_species > m->num_chem_species) ? m->num_stoch_species : m->num_chem_species;
    const int Sc = m->num_chem_species, Sd = m->num_stoch_species;
    const int Rd = m->num_stoch_rxns, ndf = m->num_data_fn;
    if (!(m->h > 0.0)) return fail(h, SSB_ERR_ARG, "h (basis function width) can not be zero.");
    // host copies (the caller's buffers are only borrowed for this call)
    h->hx.assign(m->x, m->x + (size_t) N * 3);
    h->htype.assign(m->type, m->type + N);
    h->hnu.assign(m->nu, m->nu + N);
    h->hmass.assign(m->mass, m->mass + N);
    h->hrho.assign(m->rho, m->rho + N);
    h->hsolid.assign(m->solid, m->solid + N);
    if (m->owned) h->howned.assign(m->owned, m->owned + N); else h->howned.assign((size_t) N, 1);
    if (m->rng_id) h->hgid.assign(m->rng_id, m->rng_id + N); else { h->hgid.resize((size_t) N); for (int i = 0; i < N; i++) h->hgid[i] = i; }
    if (S > 0) h->hu0.assign(m->u0, m->u0 + (size_t) N * S);
    if (ndf > 0) h->hdata_fn.assign(m->data_fn, m->data_fn + (size_t) N * ndf);
    if (S > 0) h->hdmat.assign(m->diffusion_matrix, m->diffusion_matrix + (size_t) S * m->num_types);
    h->hout_steps.assign(m->output_steps, m->output_steps + m->n_output_steps);
    for (int s = 0; s < S; s++) h->species_names.push_back(m->species_names ? m->species_names[s] : "S");
    for (int i = 0; i < N; i++)
        if (h->htype[i] < 1 || h->htype[i] > m->num_types) return fail(h, SSB_ERR_ARG, "particle %d has type %d outside 1..%d", i, h->htype[i], m->num_types);
    h->m.x = nullptr; h->m.type = nullptr; h->m.nu = h->m.mass = h->m.c = h->m.rho = nullptr; h->m.solid = nullptr;
    h->m.u0 = nullptr; h->m.data_fn = nullptr; h->m.N_dense = nullptr; h->m.irN = h->m.jcN = nullptr; h->m.prN = nullptr;
    h->m.irG = h->m.jcG = nullptr; h->m.diffusion_matrix = nullptr; h->m.species_names = nullptr; h->m.output_steps = nullptr;
    h->m.owned = nullptr; h->m.rng_id = nullptr;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(h, SSB_ERR_CUDA, "no CUDA device available: the ssa_sdpd B200 engine has no CPU fallback");
    h->device = m->device;
    if (h->device < 0 || h->device >= ndev) return fail(h, SSB_ERR_ARG, "device %d out of range (%d devices)", h->device, ndev);
    CK(cudaSetDevice(h->device));
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));

    SsbView &V = h->V;
    memset(&V, 0, sizeof(V));
    V.N = N; V.dim = m->dimension; V.static_domain = m->static_domain; V.num_types = m->num_types;
    V.Sc = Sc; V.Rc = m->num_chem_rxns; V.Sd = Sd; V.Rd = Rd; V.ndf = ndf;
    V.flags = m->flags;
    V.dt = m->dt; V.h = m->h; V.rho0 = m->rho0; V.c0 = m->c0; V.P0 = m->P0;
    for (int d = 0; d < 3; d++) V.gravity[d] = m->gravity[d];

    // permuted per-particle fields (each gets a current and an alternate buffer)
    for (int d = 0; d < 3; d++) { register_f64(h, &V.x[d]); register_f64(h, &V.v[d]); register_f64(h, &V.vt[d]); register_f64(h, &V.F[d]); register_f64(h, &V.Fbp[d]); }
    register_f64(h, &V.rho); register_f64(h, &V.old_rho); register_f64(h, &V.Frho); register_f64(h, &V.bvf);
    register_f64(h, &V.mass); register_f64(h, &V.nu);
    if (h->f64_slots.size() > PERM_MAX64) return fail(h, SSB_ERR_ARG, "too many fields");
    for (auto slot : h->f64_slots) { CK(dalloc(h, slot, (size_t) N)); double *alt; CK(dalloc(h, &alt, (size_t) N)); h->f64_alt.push_back(alt); }
    // species-major blocks are permuted row by row; allocate as blocks and register rows through shadow pointers
    CK(dalloc(h, &V.C, (size_t) Sc * N)); CK(dalloc(h, &V.Q, (size_t) Sc * N));
    CK(dalloc(h, &V.xx, (size_t) Sd * N)); CK(dalloc(h, &V.data_fn, (size_t) ndf * N));
    register_i32(h, &V.type); register_i32(h, &V.solid); register_i32(h, &V.id); register_i32(h, &V.owned); register_i32(h, &V.gid);
    for (auto slot : h->i32_slots) { CK(dalloc(h, slot, (size_t) N)); int *alt; CK(dalloc(h, &alt, (size_t) N)); h->i32_alt.push_back(alt); }
    if (h->i32_slots.size() > PERM_MAX32) return fail(h, SSB_ERR_ARG, "too many fields");
    // non-permuted scratch
    if (V.static_domain) { for (int d = 0; d < 3; d++) V.x0[d] = nullptr; }   // aliased to x after allocation (below)
    else { for (int d = 0; d < 3; d++) CK(dalloc(h, &V.x0[d], (size_t) N)); }
    CK(dalloc(h, &V.rho_new, (size_t) N));
    if (!V.static_domain) {
        CK(dalloc(h, &V.rec, (size_t) 16 * N)); CK(dalloc(h, &V.solid_nbr, (size_t) N));
        if (Sc <= 2) CK(dalloc(h, &V.rec2, (size_t) 4 * N));
    }
    if (V.static_domain && Sc > 0) { CK(dalloc(h, &V.Cpre[0], (size_t) Sc * N)); CK(dalloc(h, &V.Cpre[1], (size_t) Sc * N)); }
    CK(dalloc(h, &V.nbr_count, (size_t) N));
    CK(cudaMemsetAsync(V.nbr_count, 0, sizeof(int) * N, h->stream));
    CK(dalloc(h, &V.rrate, (size_t) Rd * N)); CK(dalloc(h, &V.srrate, (size_t) N)); CK(dalloc(h, &V.sdrate, (size_t) N));
    CK(dalloc(h, &V.tnext, (size_t) N)); CK(dalloc(h, &V.Ddiag, (size_t) Sd * N));
    CK(dalloc(h, &V.inbox[0], (size_t) Sd * N)); CK(dalloc(h, &V.inbox[1], (size_t) Sd * N));
    CK(dalloc(h, &V.inbox_src[0], (size_t) N)); CK(dalloc(h, &V.inbox_src[1], (size_t) N));
    {
        const size_t nblk = (size_t) (N + 31) / 32 + 1;     // enough for any model-unit block size >= 32
        CK(dalloc(h, &V.blk_tmin, nblk)); CK(dalloc(h, &V.blk_mail[0], nblk)); CK(dalloc(h, &V.blk_mail[1], nblk));
    }
    CK(dalloc(h, &V.err_flag, 4)); CK(dalloc(h, &V.counters, 4));
    CK(dalloc(h, &h->d_dmat, (size_t) S * m->num_types));
    if (S > 0) CK(cudaMemcpyAsync(h->d_dmat, h->hdmat.data(), sizeof(double) * S * m->num_types, cudaMemcpyHostToDevice, h->stream));
    V.dmat = h->d_dmat;
    // alternates for the species blocks
    {
        double *alt;
        CK(dalloc(h, &alt, (size_t) (2 * Sc + ndf) * N));   // one block: [C | Q | data_fn]
        h->C_alt = alt; h->Q_alt = alt + (size_t) Sc * N; h->df_alt = alt + (size_t) 2 * Sc * N;
        CK(dalloc(h, &h->xx_alt, (size_t) Sd * N));
    }
    CK(cudaHostAlloc((void **) &h->pin, sizeof(unsigned long long) * 16, cudaHostAllocDefault));
    memset(h->pin, 0, sizeof(unsigned long long) * 16);
    CK(dalloc(h, &h->d_look, 4));
    CK(cudaEventCreateWithFlags(&h->ev_look, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_maxd, cudaEventDisableTiming));
    // Verlet skin: moving domains keep candidate lists across steps unless the literal kernels are requested
    h->skin = (!V.static_domain && !(m->flags & SSB_FLAG_LITERAL_KERNELS)) ? 0.1 : 0.0;
    if (const char *e = getenv("SSB_SKIN")) { if (!V.static_domain && !(m->flags & SSB_FLAG_LITERAL_KERNELS)) { h->skin = atof(e); h->skin_chosen = 1; } }
    V.filter = h->skin > 0.0 ? 1 : 0;
    V.search_h2 = (m->h * (1.0 + h->skin)) * (m->h * (1.0 + h->skin));
    if (V.filter) { for (int d = 0; d < 3; d++) CK(dalloc(h, &V.xref[d], (size_t) N)); }
    CK(dalloc(h, &V.disp_bits, 2));
    // cell list
    setup_grid(h);
    CK(dalloc(h, &h->d_key, (size_t) N)); CK(dalloc(h, &h->d_perm, (size_t) N));
    CK(dalloc(h, &h->d_cell_count, (size_t) h->grid.ncells + 1)); CK(dalloc(h, &h->d_cell_start, (size_t) h->grid.ncells + 1));
    CK(dalloc(h, &h->d_cursor, (size_t) h->grid.ncells + 1));
    CK(dalloc(h, &h->d_tile_sums, (size_t) (h->grid.ncells / SCAN_TILE + 2)));
    CK(dalloc(h, &h->d_flags, 8));
    CK(dalloc(h, &h->d_maxbits, 2));
    // staging: the largest of an output snapshot and any single tap
    size_t per = (size_t) 3 * 8 * 2 + 4 * 8 + 8 + (size_t) Sc * 8 + (size_t) Sd * 8 + (size_t) Rd * 8 + 64;
    h->stage_bytes = per * N + 4096;
    CK(cudaMalloc((void **) &h->d_stage, h->stage_bytes));
    h->allocs.push_back(h->d_stage);
    for (int b = 0; b < 2; b++) {
        OutputJob &J = h->jobs[b];
        CK(cudaEventCreateWithFlags(&J.ready, cudaEventDisableTiming | cudaEventBlockingSync));
        CK(cudaMallocHost((void **) &J.x, sizeof(double) * 3 * N));
        CK(cudaMallocHost((void **) &J.v, sizeof(double) * 3 * N));
        CK(cudaMallocHost((void **) &J.scal, sizeof(double) * 4 * N));
        CK(cudaMallocHost((void **) &J.type, sizeof(int) * N));
        CK(cudaMallocHost((void **) &J.C, sizeof(double) * (Sc > 0 ? Sc : 1) * N));
        CK(cudaMallocHost((void **) &J.xx, sizeof(unsigned) * (Sd > 0 ? Sd : 1) * N));
    }
    {   // pinned SoA image of the initial condition (device layout), uploaded by every ssb_reset
        const size_t nd = (size_t) (6 + Sc + ndf) * N, ni = (size_t) (4 + Sd) * N;
        CK(cudaMallocHost((void **) &h->init_f64, sizeof(double) * (nd > 0 ? nd : 1)));
        CK(cudaMallocHost((void **) &h->init_i32, sizeof(int) * (ni > 0 ? ni : 1)));
        double *pd = h->init_f64;
        for (int d = 0; d < 3; d++) for (int i = 0; i < N; i++) pd[(size_t) d * N + i] = h->hx[(size_t) i * 3 + d];
        for (int i = 0; i < N; i++) { pd[(size_t) 3 * N + i] = h->hrho[i]; pd[(size_t) 4 * N + i] = h->hmass[i]; pd[(size_t) 5 * N + i] = h->hnu[i]; }
        for (int sp = 0; sp < Sc; sp++) for (int i = 0; i < N; i++) pd[(size_t) (6 + sp) * N + i] = (double) h->hu0[(size_t) i * S + sp];
        for (size_t k = 0; k < (size_t) ndf * N; k++) pd[(size_t) (6 + Sc) * N + k] = h->hdata_fn[k];
        int *pi = h->init_i32;
        for (int i = 0; i < N; i++) { pi[i] = h->htype[i]; pi[(size_t) N + i] = h->hsolid[i]; pi[(size_t) 2 * N + i] = h->howned[i]; pi[(size_t) 3 * N + i] = h->hgid[i]; }
        for (int sp = 0; sp < Sd; sp++) for (int i = 0; i < N; i++) pi[(size_t) (4 + sp) * N + i] = (int) h->hu0[(size_t) i * S + sp];
    }
    CK(ssb_sync(h));
    h->writer = std::thread(writer_main, h);
    return SSB_OK;
}

extern "C" int ssb_load_kernels(ssb_handle *h, const char *path) {
    if (!h || !path) return SSB_ERR_ARG;
    void *dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!dl) return fail(h, SSB_ERR_MODEL_UNIT, "dlopen(%s): %s", path, dlerror());
    typedef const SsbModelUnit *(*getter)();
    getter g = (getter) dlsym(dl, "ssbm_get_unit");
    if (!g) { dlclose(dl); return fail(h, SSB_ERR_MODEL_UNIT, "%s does not export ssbm_get_unit", path); }
    const SsbModelUnit *u = g();
    if (!u || u->abi != SSB_UNIT_ABI) { dlclose(dl); return fail(h, SSB_ERR_MODEL_UNIT, "model unit ABI mismatch"); }
    const SsbView &V = h->V;
    if (u->Sc != V.Sc || u->Rc != V.Rc || u->Sd != V.Sd || u->Rd != V.Rd || u->ndf != V.ndf || u->ntypes != V.num_types) {
        dlclose(dl);
        return fail(h, SSB_ERR_MODEL_UNIT, "model unit compiled for (Sc=%d Rc=%d Sd=%d Rd=%d ndf=%d types=%d), model has (%d %d %d %d %d %d)",
                    u->Sc, u->Rc, u->Sd, u->Rd, u->ndf, u->ntypes, V.Sc, V.Rc, V.Sd, V.Rd, V.ndf, V.num_types);
    }
    if (h->unit_dl) dlclose(h->unit_dl);
    h->unit_dl = dl;
    h->unit = u;
    if (u->bc_touches_rho && !h->d_rho_pre) { CK(cudaSetDevice(h->device)); CK(dalloc(h, &h->d_rho_pre, (size_t) h->N)); }
    return SSB_OK;
}

static void slab_free(ssb_handle *h);
static void slab_reset(ssb_handle *h);

static int drain_writer(ssb_handle *h) {
    std::unique_lock<std::mutex> lk(h->mu);
    h->cv.wait(lk, [&] { return !h->writer_pending[0] && !h->writer_pending[1]; });
    int rc = h->writer_error;
    h->writer_error = 0;
    return rc;
}

extern "C" int ssb_destroy(ssb_handle *h) {
    if (!h) return SSB_OK;
    if (h->writer.joinable()) {
        { std::lock_guard<std::mutex> lk(h->mu); h->writer_quit = 1; }
        h->cv.notify_all();
        h->writer.join();
    }
    cudaSetDevice(h->device);
    if (h->stream) ssb_sync(h);
    slab_free(h);
    if (h->V.nbr) cudaFreeAsync(h->V.nbr, h->stream);           // (grown with cudaMallocAsync, neighbour_search)
    if (h->V.coef) cudaFreeAsync(h->V.coef, h->stream);
    if (h->V.Dij) cudaFreeAsync(h->V.Dij, h->stream);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void *p : h->allocs) cudaFree(p);
    for (int b = 0; b < 2; b++) {
        OutputJob &J = h->jobs[b];
        if (J.ready) cudaEventDestroy(J.ready);
        cudaFreeHost(J.x); cudaFreeHost(J.v); cudaFreeHost(J.scal); cudaFreeHost(J.type); cudaFreeHost(J.C); cudaFreeHost(J.xx);
    }
    if (h->mark_a) { cudaEventDestroy(h->mark_a); cudaEventDestroy(h->mark_b); }
    if (h->sync_ev) cudaEventDestroy(h->sync_ev);
    if (h->ev_look) cudaEventDestroy(h->ev_look);
    if (h->ev_maxd) cudaEventDestroy(h->ev_maxd);
    if (h->pin) cudaFreeHost(h->pin);
    for (auto e : h->ev_a) cudaEventDestroy(e);
    for (auto e : h->ev_b) cudaEventDestroy(e);
    if (h->init_f64) cudaFreeHost(h->init_f64);
    if (h->init_i32) cudaFreeHost(h->init_i32);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->unit_dl) dlclose(h->unit_dl);
    delete h;
    return SSB_OK;
}

extern "C" int ssb_cancel(ssb_handle *h) { if (h) h->cancel.store(1); return SSB_OK; }

// ----------------------------------------------------------------------------------------------------
// state reset (replaces init_all_particles + initialize_rdme, template:134-138)
// ----------------------------------------------------------------------------------------------------
static int apply_permutation(ssb_handle *h, const int *d_perm);

extern "C" int ssb_reset(ssb_handle *h, uint64_t seed) {
    if (!h) return SSB_ERR_ARG;
    if (!h->unit) return fail(h, SSB_ERR_MODEL_UNIT, "no model unit loaded (ssb_load_kernels)");
    CK(cudaSetDevice(h->device));
    SsbView &V = h->V;
    const int N = h->N, Sc = V.Sc, Sd = V.Sd, S = h->S, ndf = V.ndf;
    cudaStream_t st = h->stream;
    // the initial condition lives in one pinned SoA block built once by ssb_create: reset = a handful of async H2D copies
    {
        const double *pd = h->init_f64;
        for (int d = 0; d < 3; d++) {
            CK(cudaMemcpyAsync(V.x[d], pd + (size_t) d * N, sizeof(double) * N, cudaMemcpyHostToDevice, st));
            // fields the reference leaves uninitialised are defined as 0 (particle.cpp:70-83)
            CK(cudaMemsetAsync(V.v[d], 0, sizeof(double) * N, st));
            CK(cudaMemsetAsync(V.vt[d], 0, sizeof(double) * N, st));
            CK(cudaMemsetAsync(V.F[d], 0, sizeof(double) * N, st));
            CK(cudaMemsetAsync(V.Fbp[d], 0, sizeof(double) * N, st));
        }
        CK(cudaMemcpyAsync(V.rho, pd + (size_t) 3 * N, sizeof(double) * N, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(V.mass, pd + (size_t) 4 * N, sizeof(double) * N, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(V.nu, pd + (size_t) 5 * N, sizeof(double) * N, cudaMemcpyHostToDevice, st));
        // species: u0 is voxel-major [N][S] (solver.py:211-220); C starts as (double) u0 (template:77-81)
        if (Sc > 0) CK(cudaMemcpyAsync(V.C, pd + (size_t) 6 * N, sizeof(double) * Sc * N, cudaMemcpyHostToDevice, st));
        if (ndf > 0) CK(cudaMemcpyAsync(V.data_fn, pd + (size_t) (6 + Sc) * N, sizeof(double) * ndf * N, cudaMemcpyHostToDevice, st));
        const int *pi = h->init_i32;
        CK(cudaMemcpyAsync(V.type, pi, sizeof(int) * N, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(V.solid, pi + (size_t) N, sizeof(int) * N, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(V.owned, pi + (size_t) 2 * N, sizeof(int) * N, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(V.gid, pi + (size_t) 3 * N, sizeof(int) * N, cudaMemcpyHostToDevice, st));
        if (Sd > 0) CK(cudaMemcpyAsync(V.xx, pi + (size_t) 4 * N, sizeof(unsigned) * Sd * N, cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemsetAsync(V.old_rho, 0, sizeof(double) * N, st));
    CK(cudaMemsetAsync(V.Frho, 0, sizeof(double) * N, st));
    CK(cudaMemsetAsync(V.bvf, 0, sizeof(double) * N, st));
    CK(cudaMemsetAsync(V.rho_new, 0, sizeof(double) * N, st));
    k_iota<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id);
    if (Sc > 0) CK(cudaMemsetAsync(V.Q, 0, sizeof(double) * Sc * N, st));
    if (Sd > 0) {
        CK(cudaMemsetAsync(V.inbox[0], 0, sizeof(unsigned) * Sd * N, st));
        CK(cudaMemsetAsync(V.inbox[1], 0, sizeof(unsigned) * Sd * N, st));
        CK(cudaMemsetAsync(V.inbox_src[0], 0, sizeof(unsigned long long) * N, st));
        CK(cudaMemsetAsync(V.inbox_src[1], 0, sizeof(unsigned long long) * N, st));
    }
    CK(cudaMemsetAsync(V.err_flag, 0, 16, st));
    CK(cudaMemsetAsync(V.counters, 0, 32, st));
    CK(cudaMemsetAsync(V.disp_bits, 0, 16, st));
    if (!h->static_cached) CK(cudaMemsetAsync(V.nbr_count, 0, sizeof(int) * N, st));
    if (V.static_domain) for (int d = 0; d < 3; d++) V.x0[d] = V.x[d];
    V.rho_search = V.rho;
    CK(ssb_sync(h));
    if (h->static_cached) {      // put the freshly uploaded (id-ordered) state into the cached storage order
        int rcp = apply_permutation(h, h->d_static_perm);
        if (rcp) return rcp;
        CK(ssb_sync(h));
    }
    h->current_step = 0;
    h->rdme_initialized = 0;
    h->seed = seed;
    h->epoch = 0;
    h->inbox_buf = 0;
    h->nbr_valid = 0;
    h->lists_valid = 0;
    h->look_valid = 0;
    h->disp_prev = 0.0;
    h->step_disp_max = 0.0;
    h->ddiag_fresh = 0;
    h->launches = 0;
    h->windows = 0;
    h->h2d_bytes = (int64_t) N * (3 * 8 + 3 * 8 + 4 * 4 + 8 * Sc + 4 * Sd + 8 * ndf);
    h->slot_dirty = 1;
    h->d2h_bytes = 0;
    h->total_reactions = h->total_diffusion = 0;
    h->step_seconds = 0.0;
    h->cancel.store(0);
    slab_reset(h);
    return SSB_OK;
}

// ----------------------------------------------------------------------------------------------------
// K1: cell list build + permutation into cell-sorted storage order
// ----------------------------------------------------------------------------------------------------
static int build_cells(ssb_handle *h) {
    SsbView &V = h->V;
    const int N = h->N, nc = h->grid.ncells;
    cudaStream_t st = h->stream;
    CK(cudaMemsetAsync(h->d_cell_count, 0, sizeof(int) * (nc + 1), st));
    CK(cudaMemsetAsync(h->d_cursor, 0, sizeof(int) * (nc + 1), st));
    CK(cudaMemsetAsync(h->d_flags, 0, sizeof(int) * 2, st));
    k_cell_keys<<<gridN(N), CORE_BLOCK, 0, st>>>(N, h->grid, V.x[0], V.x[1], V.x[2], h->d_key, h->d_cell_count);
    const int ntiles = (nc + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tiles<<<ntiles, 256, 0, st>>>(nc, h->d_cell_count, h->d_cell_start, h->d_tile_sums);
    k_scan_sums<<<1, 256, 0, st>>>(ntiles, h->d_tile_sums, nullptr);
    k_scan_add<<<gridN(nc), CORE_BLOCK, 0, st>>>(nc, h->d_cell_start, h->d_tile_sums);
    k_scatter<<<gridN(N), CORE_BLOCK, 0, st>>>(N, h->d_key, h->d_cell_start, h->d_cursor, h->d_perm);
    k_sort_cells<<<gridN(nc), CORE_BLOCK, 0, st>>>(nc, h->d_cell_start, N, h->d_perm, h->d_flags);
    h->launches += 6;
    // always permute (an identity permutation costs one pass over the state at a list build; testing for it would cost a
    // host round trip on every build)
    return apply_permutation(h, h->d_perm);
}

// dst[p] = src[perm[p]] for every per-particle field (alternate buffers, then swap); stream-ordered, no host synchronisation
static int apply_permutation(ssb_handle *h, const int *d_perm) {
    SsbView &V = h->V;
    const int N = h->N;
    cudaStream_t st = h->stream;
    PermTable T;
    memset(&T, 0, sizeof(T));
    const size_t nf = h->f64_slots.size(), ni = h->i32_slots.size();
    for (size_t f = 0; f < nf; f++) { T.src64[T.n64] = *h->f64_slots[f]; T.dst64[T.n64] = h->f64_alt[f]; T.n64++; }
    for (size_t f = 0; f < ni; f++) { T.src32[T.n32] = *h->i32_slots[f]; T.dst32[T.n32] = h->i32_alt[f]; T.n32++; }
    k_permute<<<gridN(N), CORE_BLOCK, 0, st>>>(N, d_perm, T);
    const int Sc = V.Sc, Sd = V.Sd, ndf = V.ndf;
    if (Sc > 0) {
        k_permute_rows64<<<dim3(gridN(N), (unsigned) Sc), CORE_BLOCK, 0, st>>>(N, d_perm, V.C, h->C_alt);
        k_permute_rows64<<<dim3(gridN(N), (unsigned) Sc), CORE_BLOCK, 0, st>>>(N, d_perm, V.Q, h->Q_alt);
        std::swap(V.C, h->C_alt); std::swap(V.Q, h->Q_alt);
    }
    if (ndf > 0) { k_permute_rows64<<<dim3(gridN(N), (unsigned) ndf), CORE_BLOCK, 0, st>>>(N, d_perm, V.data_fn, h->df_alt); std::swap(V.data_fn, h->df_alt); }
    if (Sd > 0) { k_permute_rows32<<<dim3(gridN(N), (unsigned) Sd), CORE_BLOCK, 0, st>>>(N, d_perm, (const int *) V.xx, (int *) h->xx_alt); std::swap(V.xx, h->xx_alt); }
    CK(cudaGetLastError());
    h->slot_dirty = 1;
    h->launches += 1 + (Sc > 0 ? 2 : 0) + (ndf > 0) + (Sd > 0);
    for (size_t f = 0; f < nf; f++) { double *cur = *h->f64_slots[f]; *h->f64_slots[f] = h->f64_alt[f]; h->f64_alt[f] = cur; }
    for (size_t f = 0; f < ni; f++) { int *cur = *h->i32_slots[f]; *h->i32_slots[f] = h->i32_alt[f]; h->i32_alt[f] = cur; }
    if (V.static_domain) for (int d = 0; d < 3; d++) V.x0[d] = V.x[d];
    V.rho_search = V.rho;
    return SSB_OK;
}

// count-only pass (capacity 0 stores nothing): total number of list entries the current (filter, search_h2) would produce
static int count_candidates(ssb_handle *h, double *total) {
    SsbView V = h->V;
    V.nbr_cap = 0;
    cudaStream_t st = h->stream;
    CK(cudaMemsetAsync(h->d_flags + 1, 0, sizeof(int), st));
    CK(cudaMemsetAsync(h->d_maxbits + 1, 0, sizeof(unsigned long long), st));
    k_search<<<gridN(h->N), CORE_BLOCK, 0, st>>>(V, h->grid, h->d_cell_start, h->d_flags + 1, h->d_maxbits + 1);
    unsigned long long t = 0;
    CK(cudaMemcpyAsync(&t, h->d_maxbits + 1, sizeof(t), cudaMemcpyDeviceToHost, st));
    CK(ssb_sync(h));
    *total = (double) t;
    h->launches += 1;
    return SSB_OK;
}

// Verlet skin selection (first list build of a handle): the widest skin whose candidate lists are at most 15 % longer than
// the exact lists.  On lattice-like clouds the neighbour shells make this a step function of the skin, so it is measured.
static int choose_skin(ssb_handle *h) {
    SsbView &V = h->V;
    const double skin_max = h->skin;
    int rc;
    double exact = 0.0;
    V.filter = 0;
    if ((rc = count_candidates(h, &exact))) return rc;
    V.filter = 1;
    double pick = 0.0;
    // (lattice-like clouds have a neighbour shell just outside h — sqrt(5) d for h = 2.2 d, sqrt(6) d for h = 2.42 d — so the skin that
    // qualifies is small, ~0.6-1.2 % of h; it still saves 10-20 list builds per rebuild at SDPD time steps, and a step that would
    // not fit simply rebuilds (look-ahead displacement), so a small skin has no failure mode)
    for (double sk = skin_max; sk >= 0.004; sk *= 0.5) {
        V.search_h2 = (V.h * (1.0 + sk)) * (V.h * (1.0 + sk));
        double cand = 0.0;
        if ((rc = count_candidates(h, &cand))) return rc;
        if (cand <= 1.15 * exact + (double) h->N) { pick = sk; break; }     // (+N: the self entry every particle carries)
    }
    h->skin = pick;
    V.filter = pick > 0.0 ? 1 : 0;
    V.search_h2 = (V.h * (1.0 + pick)) * (V.h * (1.0 + pick));
    h->skin_chosen = 1;
    return SSB_OK;
}

static int neighbour_search(ssb_handle *h) {
    SsbView &V = h->V;
    const int N = h->N;
    cudaStream_t st = h->stream;
    if (V.filter && !h->skin_chosen) { int rcs = choose_skin(h); if (rcs) return rcs; }
    for (int attempt = 0; attempt < 8; attempt++) {
        // (first build: the capacity is 0, so this pass only counts; the rows are then sized from the maximum)
        CK(cudaMemsetAsync(h->d_flags + 1, 0, sizeof(int), st));
        CK(cudaMemsetAsync(h->d_maxbits + 1, 0, sizeof(unsigned long long), st));
        k_search<<<gridN(N), CORE_BLOCK, 0, st>>>(V, h->grid, h->d_cell_start, h->d_flags + 1, h->d_maxbits + 1);
        h->launches += 1;
        int mx = 0;
        CK(cudaMemcpyAsync(&mx, h->d_flags + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(ssb_sync(h));
        if (mx <= V.nbr_cap) return SSB_OK;
        // grow (with head-room on moving domains) and search again
        // stream-ordered allocation (cudaMallocAsync / cudaFreeAsync): cudaMalloc and cudaFree synchronise the whole DEVICE, and
        // on a device shared by several slab ranks (ranks as threads, one GPU) another rank's stream may at that moment hold a kernel
        // that waits for THIS rank's halo message — a deadlock until the device-side timeout.  Also keeps the step loop free of a
        // device-wide stall on every capacity growth.
        int cap = V.static_domain ? mx : (mx + mx / 4 + 8);
        int *nb = nullptr;
        CK(cudaMallocAsync((void **) &nb, sizeof(int) * (size_t) cap * N, st));
        if (V.nbr) CK(cudaFreeAsync(V.nbr, st));
        V.nbr = nb;
        V.nbr_cap = cap;
        if (V.static_domain && V.Sc > 0) {
            double *cf = nullptr;
            CK(cudaMallocAsync((void **) &cf, sizeof(double) * (size_t) cap * N, st));
            if (V.coef) CK(cudaFreeAsync(V.coef, st));
            V.coef = cf;
        }
        if (V.static_domain && V.Sd > 0) {
            double *dj = nullptr;
            CK(cudaMallocAsync((void **) &dj, sizeof(double) * (size_t) cap * N, st));
            if (V.Dij) CK(cudaFreeAsync(V.Dij, st));
            V.Dij = dj;
        }
    }
    return fail(h, SSB_ERR_CUDA, "neighbour list capacity did not converge");
}

// ----------------------------------------------------------------------------------------------------
// one engine step: simulate_threads.cpp:232-281 (three substeps, then the RDME)
// ----------------------------------------------------------------------------------------------------
static int check_device_error(ssb_handle *h) {
    int flags[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(flags, h->V.err_flag, sizeof(flags), cudaMemcpyDeviceToHost, h->stream));
    CK(ssb_sync(h));
    const int flag = flags[0];
    if (flag == SSB_ERR_NAN) return fail(h, SSB_ERR_NAN, "ERROR: nan/inf detected!!! (step %u)", h->current_step);
    if (flag == SSB_ERR_RDME) return fail(h, SSB_ERR_RDME, "RDME state error (negative population or propensity overflow) at step %u", h->current_step);
    if (flag == SSB_ERR_HALO && getenv("SSB_SLAB_DEBUG"))
        fprintf(stderr, "[slab] HALO timeout at step %u: wait %d expected seq %d found %d\n", h->current_step, flags[1], flags[2], flags[3]);
    if (flag == SSB_ERR_HALO)
        return fail(h, SSB_ERR_HALO, "halo exchange timed out: a neighbouring slab rank did not deliver its message (step %u; wait %d: 10x = window flag "
                    "of side x, 2cr = board channel c rank r; expected sequence %d, found %d)", h->current_step, flags[1], flags[2], flags[3]);
    return SSB_OK;
}

// The reference's NSM loop tests `tt <= end_time` BEFORE it pops the next event (simulate_rdme.cpp:233-238), so every call executes
// exactly one event that lies beyond the end of the step.  On static domains that event is merely executed early; on moving
// domains the queue is rebuilt from scratch at the next step (simulate_rdme.cpp:54-65), so it is one EXTRA event per step —
// dominant in small, quiet systems (the 512-particle tank fixture: 37 reference events in 25 steps).  Parity mode mirrors it:
// after the windows of the step, the globally earliest pending event is executed (a window that ends exactly at its time).
static int rdme_min_time(ssb_handle *h, double *tmin) {
    const int nchunks = (h->N + h->unit->block - 1) / h->unit->block;
    unsigned long long bits = 0x7ff0000000000000ull;      // +inf
    CK(cudaMemcpyAsync(h->d_maxbits + 1, &bits, sizeof(bits), cudaMemcpyHostToDevice, h->stream));
    k_min_time<<<gridN(nchunks), CORE_BLOCK, 0, h->stream>>>(nchunks, h->V.blk_tmin, h->d_maxbits + 1);
    CK(cudaMemcpyAsync(&bits, h->d_maxbits + 1, sizeof(bits), cudaMemcpyDeviceToHost, h->stream));
    CK(ssb_sync(h));
    memcpy(tmin, &bits, sizeof(bits));
    h->launches += 1;
    return SSB_OK;
}
// deliver = false (slab decomposition): only the event window runs here; if the event is a jump into a ghost voxel the molecule
// must first travel to its owner (ssb_halo_inbox_pack/add read the buffer the LAST window wrote), so the caller exchanges the
// inboxes and then runs the delivering zero-length window itself (PH_RDME_CLOSE).
static int rdme_extra_event(ssb_handle *h, double tmin, bool deliver = true) {
    SsbView &V = h->V;
    const SsbModelUnit *u = h->unit;
    const double te = V.dt * (h->current_step + 1);
    h->epoch += 2;                 // the two window epochs are consumed whether or not an event is pending (same numbering on every path)
    if (!(tmin < INFINITY) || !(tmin > te)) return SSB_OK;
    if (u->rdme_window(&V, te, tmin, h->tau, h->seed, h->epoch - 2, h->inbox_buf, h->stream)) return fail(h, SSB_ERR_CUDA, "rdme_window launch failed");
    h->inbox_buf ^= 1;
    h->launches += 1;
    if (!deliver) return SSB_OK;
    if (u->rdme_window(&V, tmin, tmin, h->tau, h->seed, h->epoch - 1, h->inbox_buf, h->stream)) return fail(h, SSB_ERR_CUDA, "rdme_window launch failed");
    h->inbox_buf ^= 1;
    h->launches += 1;
    return SSB_OK;
}

// Host wait for ONE event recorded earlier on the engine stream (a scalar the device reports through pinned memory) while later
// kernels are already queued behind it: the GPU never idles for the wait.  Short waits are polled, long ones sleep (see ssb_sync).
static cudaError_t wait_event(cudaEvent_t ev) {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        cudaError_t e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady) return e;
        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) break;
    }
    return cudaEventSynchronize(ev);
}

// window controller: tau * (largest per-molecule jump rate) <= rdme_epsilon, and an integer number of windows per step
static int set_windows(ssb_handle *h, double max_ddiag) {
    const SsbView &V = h->V;
    // Default 0.0125.  The windowed scheme delays every jump to the end of its window, so while the spatial distribution is still
    // relaxing the reaction counts carry a FIRST-order error in tau: measured on Cdc42 (D = 10, membrane reaction fed by cytoplasmic
    // diffusion, 1000 vs 1000 reference trajectories, profiles/eps_sweep.py) +5.7 % at 0.05 (5.2 standard errors, KS p < 1e-4),
    // +3.0 % at 0.025 — 0.0125 keeps it near one standard error of a 1000-trajectory ensemble.  Moving domains at SDPD time steps
    // have one window per step either way (tau is capped by dt).
    const double eps = (h->m.rdme_epsilon > 0.0) ? h->m.rdme_epsilon : 0.0125;
    const double tau = (max_ddiag > 0.0) ? eps / max_ddiag : V.dt;
    double nwin_d = ceil(V.dt / tau);
    if (!(nwin_d >= 1.0)) nwin_d = 1.0;
    if (nwin_d > 5.0e7) return fail(h, SSB_ERR_ARG, "sSSA window count per step (%g) too large; raise rdme_epsilon", nwin_d);
    h->nwin = (long long) nwin_d;
    h->tau = V.dt / (double) h->nwin;
    return SSB_OK;
}

// largest Ddiag of this step -> *mx.  Moving domains: the force sweep assembled Ddiag and its maximum, and its copy into pinned memory
// was queued right behind the sweep (ev_maxd) — by now the corrector and the BVF sweep are queued behind that, so the wait is free.
static int step_max_ddiag(ssb_handle *h, double *mx) {
    SsbView &V = h->V;
    cudaStream_t st = h->stream;
    unsigned long long bits = 0;
    if (h->static_cached && V.static_domain) h->ddiag_fresh = 2;   // Ddiag, D_ij, tau from the first trajectory are still valid
    if (h->ddiag_fresh == 2) {
        memcpy(&bits, &h->max_ddiag_cached, sizeof(bits));
    } else if (h->ddiag_fresh == 1) {
        CK(wait_event(h->ev_maxd));
        bits = h->pin[3];
    } else {
        CK(cudaMemsetAsync(h->d_maxbits, 0, sizeof(unsigned long long), st));
        int ps0 = prof_begin(h, CAT_DIFF_INIT, 1);
        if (h->unit->diff_init(&V, h->d_maxbits, st)) return fail(h, SSB_ERR_CUDA, "diff_init launch failed");
        prof_end(h, ps0);
        h->launches += 1;
        CK(cudaMemcpyAsync(&bits, h->d_maxbits, sizeof(bits), cudaMemcpyDeviceToHost, st));
        CK(ssb_sync(h));
    }
    h->ddiag_fresh = 0;
    memcpy(mx, &bits, sizeof(*mx));
    h->max_ddiag_cached = *mx;
    return SSB_OK;
}

static int rdme_step(ssb_handle *h) {
    SsbView &V = h->V;
    const SsbModelUnit *u = h->unit;
    cudaStream_t st = h->stream;
    if (V.Sd == 0) return SSB_OK;
    const double t0 = V.dt * h->current_step;
    if (!V.static_domain || !h->rdme_initialized) {      // simulate_rdme.cpp:54-65
        double mx = 0.0;
        int rc = step_max_ddiag(h, &mx);
        if (!rc) rc = set_windows(h, mx);
        if (rc) return rc;
        // propensities are (re)initialised at t = 0.0 in the reference (simulate_rdme.cpp:124)
        int ps1 = prof_begin(h, CAT_RDME_INIT, 1);
        if (u->rdme_init(&V, t0, 0.0, h->tau, h->seed, h->epoch++, st)) return fail(h, SSB_ERR_CUDA, "rdme_init launch failed");
        prof_end(h, ps1);
        h->launches += 1;
        h->rdme_initialized = 1;
        h->inbox_buf = 0;
    }
    const long long nwin = h->nwin;
    int psw = prof_begin(h, CAT_RDME_WINDOW, (int) nwin + 1);
    // all windows of the step plus the zero-length closing window (delivers in-flight molecules so that the state read at
    // the step boundary conserves molecules) — one cooperative launch when the device supports it
    int nl = 0;
    const int rw = u->rdme_windows(&V, t0, V.dt, nwin, h->tau, h->seed, h->epoch, h->inbox_buf, &nl, st);
    if (rw > 0) return fail(h, SSB_ERR_CUDA, "rdme_windows launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    h->epoch += (uint64_t) nwin + 1;
    h->inbox_buf ^= (int) ((nwin + 1) & 1);
    if (!V.static_domain && !(V.flags & (SSB_FLAG_CORRECTED_NSM_SELECT | SSB_FLAG_NO_STEP_OVERSHOOT))) {
        // the step-end overshoot event (simulate_rdme.cpp:233-238): done inside the cooperative launch (rw == 0), else from the host
        if (rw < 0) {
            double tmin = INFINITY;
            int rcx = rdme_min_time(h, &tmin);
            if (!rcx) rcx = rdme_extra_event(h, tmin);
            if (rcx) return rcx;
        } else {
            h->epoch += 2;
        }
    }
    prof_end(h, psw);
    h->launches += nl;
    if (h->cancel.load()) return fail(h, SSB_ERR_CANCELLED, "cancelled");
    h->windows += nwin;
    return SSB_OK;
}

// ---- the pieces of a MOVING-domain step (shared by engine_step, the phase API and the native slab step) -----------------------
// Verlet skin: the candidate lists hold every j with |q_i - xref_j| <= h*(1+skin), q_i = the predicted position of the step that
// built them and xref = the positions at the start of that step (the snapshot the build step searched).  A pair that is within h
// NOW, |x_i - x0_j| <= h with x0 = the previous step's x, is on the list when D(x_i) + D(x0_j) + |q_i - xref_i| <= skin*h, D =
// displacement from xref.  All three terms are known before the step is queued: the first two from k_lookahead, the last one is
// the single-step displacement of the build step (`disp_build`, the same kernel's second output at that step).
static int mv_lookahead(ssb_handle *h);
static int mv_decide_keep(ssb_handle *h, bool *keep) {
    SsbView &V = h->V;
    *keep = false;
    if (!V.filter) return SSB_OK;
    if (!h->look_valid) {                 // first step of a trajectory, or the state was set from outside: look ahead now
        int rc = mv_lookahead(h);
        if (rc) return rc;
    }
    CK(wait_event(h->ev_look));
    double d2next, s2next, d2cur;
    memcpy(&d2next, &h->pin[0], 8); memcpy(&s2next, &h->pin[1], 8); memcpy(&d2cur, &h->pin[2], 8);
    h->disp_step_next = sqrt(s2next);
    if (h->disp_step_next > h->step_disp_max) h->step_disp_max = h->disp_step_next;
    if (!h->lists_valid) return SSB_OK;   // (xref is not meaningful yet: d2next, d2cur are unused)
    const double Dnext = sqrt(d2next), Dcur = sqrt(d2cur), budget = h->skin * V.h * (1.0 - 1e-9);
    h->disp_prev = Dcur;
    *keep = (Dnext + Dcur + h->disp_build <= budget);
    return SSB_OK;
}

// cell list (if due) + predictor + neighbour search (if due) + force sweep; queues the copy of max Ddiag behind the sweep
static int mv_pre(ssb_handle *h) {
    SsbView &V = h->V;
    const SsbModelUnit *u = h->unit;
    cudaStream_t st = h->stream;
    const unsigned step = h->current_step;
    int rc, ps;
    bool keep_lists = false;
    if ((rc = mv_decide_keep(h, &keep_lists))) return rc;
    if (!keep_lists) {                                                       // buildKDTree (simulate_threads.cpp:80-108)
        ps = prof_begin(h, CAT_CELLS, 7);
        rc = build_cells(h);
        if (!rc && V.filter) {
            for (int d = 0; d < 3; d++) CK(cudaMemcpyAsync(V.xref[d], V.x[d], sizeof(double) * V.N, cudaMemcpyDeviceToDevice, st));
            h->disp_prev = 0.0;
            h->disp_build = h->disp_step_next;      // |q_i - xref_i| of the lists this step builds (mv_decide_keep)
        }
        h->rebuilds++;
        prof_end(h, ps);
        if (rc) return rc;
    }
    V.rho_pre = nullptr;
    if (step == 0 && h->d_rho_pre) {      // (see SsbView::rho_pre: the step-0 lists freeze densities from before the predictor / BCs)
        CK(cudaMemcpyAsync(h->d_rho_pre, V.rho, sizeof(double) * V.N, cudaMemcpyDeviceToDevice, st));
        V.rho_pre = h->d_rho_pre;
    }
    ps = prof_begin(h, CAT_PREDICTOR, 1);
    if (u->predictor(&V, step, st)) return fail(h, SSB_ERR_CUDA, "predictor launch failed");
    prof_end(h, ps);
    h->launches++;
    V.rho_search = V.rho;
    if (!keep_lists) {                                                       // find_neighbors (simulate.cpp:61-63,121-123)
        ps = prof_begin(h, CAT_SEARCH, 1);
        rc = neighbour_search(h);
        prof_end(h, ps);
        if (rc) return rc;
        h->lists_valid = 1;
    }
    if (!(V.flags & SSB_FLAG_LITERAL_KERNELS)) {
        // optimised moving-domain sweep; also produces Ddiag + its maximum for the sSSA window controller
        if (V.Sd > 0) CK(cudaMemsetAsync(h->d_maxbits, 0, sizeof(unsigned long long), st));
        ps = prof_begin(h, CAT_FORCE, 1);
        if (u->force_mv(&V, step, h->d_maxbits, st)) return fail(h, SSB_ERR_CUDA, "force launch failed");
        prof_end(h, ps);
        h->launches++;
        if (V.Sd > 0 && V.rho_pre) {
            // step 0 with density-assigning BCs: the fused Ddiag used post-BC densities for both particles of a pair; redo it with
            // the reference's step-0 rule (k_diff_init -> pair_Dij -> ssb_search_rho).  One extra sweep, once per trajectory.
            CK(cudaMemsetAsync(h->d_maxbits, 0, sizeof(unsigned long long), st));
            if (u->diff_init(&V, h->d_maxbits, st)) return fail(h, SSB_ERR_CUDA, "diff_init launch failed");
            h->launches++;
        }
        if (V.Sd > 0) {
            CK(cudaMemcpyAsync(&h->pin[3], h->d_maxbits, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(h->ev_maxd, st));
            h->ddiag_fresh = 1;
        }
    } else {
        ps = prof_begin(h, CAT_FORCE, 1);
        if (u->force(&V, step, 1, st)) return fail(h, SSB_ERR_CUDA, "force launch failed");
        prof_end(h, ps);
        h->launches++;
    }
    return SSB_OK;
}
static int mv_corrector(ssb_handle *h) {
    int ps = prof_begin(h, CAT_CORRECTOR, 1);
    if (h->unit->corrector(&h->V, h->current_step, h->stream)) return fail(h, SSB_ERR_CUDA, "corrector launch failed");
    prof_end(h, ps);
    h->launches++;
    return SSB_OK;
}
static int mv_finish(ssb_handle *h) {
    SsbView &V = h->V;
    int ps = prof_begin(h, CAT_FINISH, 1);
    if (h->unit->finish(&V, h->current_step, 1, h->stream)) return fail(h, SSB_ERR_CUDA, "finish launch failed");
    prof_end(h, ps);
    h->launches++;
    // rho <- post-corrector density; the old buffer keeps the search-time density frozen into D_i_j
    double *pre = V.rho;
    V.rho = V.rho_new;
    V.rho_new = pre;
    V.rho_search = pre;
    return SSB_OK;
}
// displacement bookkeeping of the NEXT step (after the last kernel that changes v, F or Fbp of any particle, ghost copies included)
static int mv_lookahead(ssb_handle *h) {
    SsbView &V = h->V;
    if (!V.filter) return SSB_OK;
    cudaStream_t st = h->stream;
    int ps = prof_begin(h, CAT_FINISH, 1);
    CK(cudaMemsetAsync(h->d_look, 0, sizeof(unsigned long long) * 3, st));
    k_lookahead<<<std::min(gridN(V.N), 148u * 8u), CORE_BLOCK, 0, st>>>(V, h->d_look);
    CK(cudaMemcpyAsync(h->pin, h->d_look, sizeof(unsigned long long) * 3, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(h->ev_look, st));
    prof_end(h, ps);
    h->look_valid = 1;
    h->launches++;
    return SSB_OK;
}

static int engine_step(ssb_handle *h) {
    SsbView &V = h->V;
    const SsbModelUnit *u = h->unit;
    cudaStream_t st = h->stream;
    const unsigned step = h->current_step;
    const bool moving = !V.static_domain;
    int rc;
    int ps;
    if (moving) {
        // simulate_threads.cpp:232-281: substeps 0-2, then the RDME.  Steady state: no blocking synchronisation — two event waits
        // (look-ahead displacement at the start, max Ddiag before the sSSA), both behind kernels that are already queued.
        if ((rc = mv_pre(h))) return rc;
        if ((rc = mv_corrector(h))) return rc;
        if ((rc = mv_finish(h))) return rc;
        if ((rc = mv_lookahead(h))) return rc;
        if ((rc = rdme_step(h))) return rc;
        h->current_step++;
        return SSB_OK;
    }
    // static-domain fast path: one fused chemistry kernel per step (k_static_step); with no continuous species a static
    // step has no SDPD work at all after step 0 (boundary conditions are idempotent assignments)
    const bool fast_static = (V.flags & SSB_FLAG_SKIP_STATIC_FORCES) && !u->bc_touches_rho;
    if (fast_static && step > 0) {
        if (V.Sc > 0) {
            ps = prof_begin(h, CAT_FORCE, 1);
            if (u->static_step(&V, step, (int) ((step + 1) & 1), st)) return fail(h, SSB_ERR_CUDA, "static_step launch failed");
            prof_end(h, ps);
            h->launches++;
        }
        if ((rc = rdme_step(h))) return rc;
        h->current_step++;
        return SSB_OK;
    }
    const bool reuse = h->static_cached;                                     // geometry work of step 0 already done by an earlier trajectory
    if (step == 0 && !reuse) {                                               // buildKDTree (simulate_threads.cpp:80-108)
        ps = prof_begin(h, CAT_CELLS, 7);
        rc = build_cells(h);
        prof_end(h, ps);
        if (rc) return rc;
    }
    V.rho_pre = nullptr;
    if (step == 0 && h->d_rho_pre) {      // static domains keep their step-0 lists (and the D_i_j frozen then) for the whole run
        CK(cudaMemcpyAsync(h->d_rho_pre, V.rho, sizeof(double) * V.N, cudaMemcpyDeviceToDevice, st));
        V.rho_pre = h->d_rho_pre;
    }
    ps = prof_begin(h, CAT_PREDICTOR, 1);
    if (u->predictor(&V, step, st)) return fail(h, SSB_ERR_CUDA, "predictor launch failed");
    prof_end(h, ps);
    h->launches++;
    if (step == 0 && !reuse) {                                               // find_neighbors (simulate.cpp:61-63,121-123)
        V.rho_search = V.rho;
        ps = prof_begin(h, CAT_SEARCH, 1);
        rc = neighbour_search(h);
        prof_end(h, ps);
        if (rc) return rc;
    }
    if (fast_static) {
        // step 0 of the fast path: cache the pair coefficients, then run the fused kernel from a copy of C
        if (V.Sc > 0) {
            if (!reuse && u->static_coef(&V, st)) return fail(h, SSB_ERR_CUDA, "static_coef launch failed");
            ps = prof_begin(h, CAT_FORCE, 1);
            if (u->static_step(&V, step, 1, st)) return fail(h, SSB_ERR_CUDA, "static_step launch failed");
            prof_end(h, ps);
            h->launches += 2;
        }
        if ((rc = rdme_step(h))) return rc;
        if (!reuse) {             // remember the storage order so later trajectories skip the geometry work
            if (!h->d_static_perm) CK(dalloc(h, &h->d_static_perm, (size_t) V.N));
            CK(cudaMemcpyAsync(h->d_static_perm, V.id, sizeof(int) * V.N, cudaMemcpyDeviceToDevice, st));
            h->static_cached = 1;
        }
        h->current_step++;
        return SSB_OK;
    }
    const bool full = !(V.flags & SSB_FLAG_SKIP_STATIC_FORCES);
    if (full || V.Sc > 0) {
        ps = prof_begin(h, CAT_FORCE, 1);
        if (u->force(&V, step, full ? 1 : 0, st)) return fail(h, SSB_ERR_CUDA, "force launch failed");
        prof_end(h, ps);
        h->launches++;
    }
    ps = prof_begin(h, CAT_FINISH, 1);
    if (u->finish(&V, step, 0, st)) return fail(h, SSB_ERR_CUDA, "finish launch failed");
    prof_end(h, ps);
    h->launches++;
    if ((rc = rdme_step(h))) return rc;
    h->current_step++;
    return SSB_OK;
}

extern "C" int ssb_step(ssb_handle *h, uint32_t nsteps) {
    if (!h) return SSB_ERR_ARG;
    if (!h->unit) return fail(h, SSB_ERR_MODEL_UNIT, "no model unit loaded (ssb_load_kernels)");
    CK(cudaSetDevice(h->device));
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t s = 0; s < nsteps; s++) {
        int rc = engine_step(h);
        if (rc) return rc;
        if ((s & 15) == 15 || s + 1 == nsteps) { if ((rc = check_device_error(h))) return rc; }
        if (h->cancel.load()) return fail(h, SSB_ERR_CANCELLED, "cancelled");
    }
    CK(ssb_sync(h));
    h->step_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return SSB_OK;
}

// ----------------------------------------------------------------------------------------------------
// K8 output staging
// ----------------------------------------------------------------------------------------------------
static int stage_output(ssb_handle *h, const char *dir, unsigned file_index) {
    SsbView &V = h->V;
    const int N = h->N, Sc = V.Sc, Sd = V.Sd;
    cudaStream_t st = h->stream;
    const int b = h->job_cursor;
    {   // the slot must have been written out before it is overwritten
        std::unique_lock<std::mutex> lk(h->mu);
        h->cv.wait(lk, [&] { return !h->writer_pending[b]; });
        if (h->writer_error) { int rc = h->writer_error; h->writer_error = 0; return fail(h, rc, "could not write VTK output into %s", dir ? dir : "(none)"); }
    }
    OutputJob &J = h->jobs[b];
    // device staging layout (id order): x[3N] v[3N] scal[4N] C[Sc*N] | type[N] xx[Sd*N]
    double *sx = h->d_stage, *sv = sx + (size_t) 3 * N, *ss = sv + (size_t) 3 * N, *sC = ss + (size_t) 4 * N;
    int *stype = (int *) (sC + (size_t) Sc * N);
    int *sxx = stype + N;
    for (int d = 0; d < 3; d++) {
        k_unperm64<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, V.x[d], sx, 3, d);
        k_unperm64<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, V.v[d], sv, 3, d);
    }
    const double *scal[4] = {V.rho, V.mass, V.bvf, V.nu};
    for (int f = 0; f < 4; f++) k_unperm64<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, scal[f], ss + (size_t) f * N, 1, 0);
    for (int s = 0; s < Sc; s++) k_unperm64<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, V.C + (size_t) s * N, sC + (size_t) s * N, 1, 0);
    k_unperm32<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, V.type, stype, 1, 0);
    for (int s = 0; s < Sd; s++) k_unperm32<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, (int *) V.xx + (size_t) s * N, sxx + (size_t) s * N, 1, 0);
    h->launches += 11 + Sc + Sd;
    h->d2h_bytes += (int64_t) N * (3 * 8 * 2 + 4 * 8 + 4 + 8 * Sc + 4 * Sd);
    CK(cudaMemcpyAsync(J.x, sx, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(J.v, sv, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(J.scal, ss, sizeof(double) * 4 * N, cudaMemcpyDeviceToHost, st));
    if (Sc > 0) CK(cudaMemcpyAsync(J.C, sC, sizeof(double) * Sc * N, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(J.type, stype, sizeof(int) * N, cudaMemcpyDeviceToHost, st));
    if (Sd > 0) CK(cudaMemcpyAsync(J.xx, sxx, sizeof(unsigned) * Sd * N, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(J.ready, st));
    // d_stage is reused by the next snapshot: the copies above are stream-ordered before any later gather
    J.step = h->current_step;
    J.file_index = file_index;
    J.rdme_initialized = h->rdme_initialized;
    J.write_file = (dir && !(h->m.flags & SSB_FLAG_NO_VTK)) ? 1 : 0;
    J.write_bin = (dir && (h->m.flags & SSB_FLAG_BINARY_STORE)) ? 1 : 0;
    J.dir = dir ? dir : "";
    {
        std::lock_guard<std::mutex> lk(h->mu);
        h->writer_pending[b] = 1;
    }
    h->cv.notify_all();
    h->job_cursor ^= 1;
    return SSB_OK;
}

extern "C" int ssb_run(ssb_handle *h, uint64_t seed, int32_t ntraj, int32_t first_traj, const char *const *out_dirs,
                       ssb_progress_cb cb, void *cb_user) {
    if (!h || ntraj < 0) return SSB_ERR_ARG;
    if (!h->unit) return fail(h, SSB_ERR_MODEL_UNIT, "no model unit loaded (ssb_load_kernels)");
    const bool write_files = !(h->m.flags & SSB_FLAG_NO_VTK) || (h->m.flags & SSB_FLAG_BINARY_STORE);
    if (write_files && !out_dirs) return SSB_ERR_ARG;
    const unsigned nt = h->m.nt;
    for (int k = 0; k < ntraj; k++) {
        int rc = ssb_reset(h, seed + (uint64_t) (first_traj + k));      // solver.py:558-559
        if (rc) return rc;
        auto t0 = std::chrono::steady_clock::now();
        // output gate of run_simulation (simulate_threads.cpp:231-247): next_output_step starts at 0 and
        // get_next_output() walks the table from its first entry, which yields the reference's file->step map.
        unsigned next_output_step = 0;
        size_t out_index = 0;
        unsigned file_index = 0;
        // SSB_FLAG_CORRECTED_OUTPUT_STEPS: file k holds step output_steps[k], nothing else is written
        const bool exact_steps = (h->m.flags & SSB_FLAG_CORRECTED_OUTPUT_STEPS) != 0;
        for (unsigned step = 0; step < nt; step++) {
            if (exact_steps) {
                while (out_index < h->hout_steps.size() && h->hout_steps[out_index] <= step) {
                    if (h->hout_steps[out_index] == step) {
                        if ((rc = stage_output(h, write_files ? out_dirs[k] : nullptr, file_index))) return rc;
                        file_index++;
                    }
                    out_index++;
                }
            } else if (step >= next_output_step) {
                if ((rc = stage_output(h, write_files ? out_dirs[k] : nullptr, file_index))) return rc;
                file_index++;
                next_output_step = (out_index < h->hout_steps.size()) ? h->hout_steps[out_index] : 0xffffffffu;
                out_index++;
            }
            if ((rc = engine_step(h))) return rc;
            if ((step & 15) == 15 || step + 1 == nt) { if ((rc = check_device_error(h))) return rc; }
            if (h->cancel.load()) { drain_writer(h); return fail(h, SSB_ERR_CANCELLED, "cancelled"); }
            if (cb && cb(cb_user, step + 1, nt)) { drain_writer(h); return fail(h, SSB_ERR_CANCELLED, "cancelled by callback"); }
        }
        bool final_file = !exact_steps;                                                           // final timepoint (:283-285)
        for (; exact_steps && out_index < h->hout_steps.size(); out_index++) final_file |= (h->hout_steps[out_index] == nt);
        if (final_file && (rc = stage_output(h, write_files ? out_dirs[k] : nullptr, file_index))) return rc;
        CK(ssb_sync(h));
        h->step_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        unsigned long long cnt[2] = {0, 0};
        CK(cudaMemcpyAsync(cnt, h->V.counters, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream));
        CK(ssb_sync(h));
        h->total_reactions = (int64_t) cnt[0];
        h->total_diffusion = (int64_t) cnt[1];
        if ((rc = drain_writer(h))) return fail(h, rc, "could not write VTK output into %s", write_files ? out_dirs[k] : "(none)");
    }
    return SSB_OK;
}

extern "C" int ssb_counters(ssb_handle *h, int64_t *reactions, int64_t *diffusions, double *seconds, int64_t *windows) {
    if (!h) return SSB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(ssb_sync(h));
    unsigned long long cnt[2] = {0, 0};
    CK(cudaMemcpyAsync(cnt, h->V.counters, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream));
    CK(ssb_sync(h));
    if (reactions) *reactions = (int64_t) cnt[0];
    if (diffusions) *diffusions = (int64_t) cnt[1];
    if (seconds) *seconds = h->step_seconds;
    if (windows) *windows = h->windows;
    return SSB_OK;
}

// Host-assembled snapshots (slab runs gather the owned particles of every rank; batched ensembles cut one state into its copies):
// the SAME writers as the engine's own output thread, fed from caller memory.  No device work, no handle.
extern "C" int ssb_write_snapshot(const char *dir, uint32_t file_index, uint32_t step, int32_t rdme_initialized, int64_t np, int32_t Sc,
                                  int32_t Sd, const char *const *species_names, const double *lims6, const double *x, const double *v,
                                  const double *scal, const double *C, const int32_t *type, const uint32_t *xx, uint32_t what) {
    if (!dir || !lims6 || !x || !v || !scal || !type || np < 0 || np > 0x7fffffff || Sc < 0 || Sd < 0) return SSB_ERR_ARG;
    if ((Sc > 0 && !C) || (Sd > 0 && !xx) || ((Sc > 0 || Sd > 0) && !species_names)) return SSB_ERR_ARG;
    std::vector<std::string> names;
    for (int s = 0; s < std::max(Sc, Sd); s++) names.emplace_back(species_names[s] ? species_names[s] : "");
    SnapMeta M{(int) np, Sc, Sd, lims6[0], lims6[1], lims6[2], lims6[3], lims6[4], lims6[5], &names};
    OutputJob J;
    J.step = step; J.file_index = file_index; J.rdme_initialized = rdme_initialized; J.dir = dir;
    J.x = const_cast<double *>(x); J.v = const_cast<double *>(v); J.scal = const_cast<double *>(scal);
    J.C = const_cast<double *>(C); J.type = const_cast<int *>(type); J.xx = const_cast<unsigned *>(xx);
    int rc = 0;
    if (what & 1u) rc = write_vtk_impl(M, J);
    if (!rc && (what & 2u)) rc = write_bin_impl(M, J);
    return rc;
}

extern "C" int ssb_launch_count(ssb_handle *h, int64_t *launches) {
    if (!h || !launches) return SSB_ERR_ARG;
    *launches = h->launches;
    return SSB_OK;
}

// ----------------------------------------------------------------------------------------------------
// parity taps
// ----------------------------------------------------------------------------------------------------
extern "C" int ssb_get_field(ssb_handle *h, const char *name, void *dst, int64_t bytes) {
    if (!h || !name || !dst) return SSB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    SsbView &V = h->V;
    const int N = h->N;
    cudaStream_t st = h->stream;
    std::string n(name);
    double **v3 = nullptr;
    if (n == "x") v3 = V.x; else if (n == "v") v3 = V.v; else if (n == "vt") v3 = V.vt; else if (n == "F") v3 = V.F; else if (n == "Fbp") v3 = V.Fbp;
    else if (n == "x0") v3 = V.x0;
    if (v3) {
        if (bytes != (int64_t) sizeof(double) * 3 * N) return fail(h, SSB_ERR_ARG, "field %s needs %lld bytes", name, (long long) sizeof(double) * 3 * N);
        for (int d = 0; d < 3; d++) k_unperm64<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, v3[d], h->d_stage, 3, d);
        CK(cudaMemcpyAsync(dst, h->d_stage, bytes, cudaMemcpyDeviceToHost, st));
        CK(ssb_sync(h));
        return SSB_OK;
    }
    const double *s1 = nullptr;
    if (n == "rho") s1 = V.rho; else if (n == "old_rho") s1 = V.old_rho; else if (n == "Frho") s1 = V.Frho; else if (n == "bvf_phi") s1 = V.bvf;
    else if (n == "mass") s1 = V.mass; else if (n == "nu") s1 = V.nu; else if (n == "srrate") s1 = V.srrate; else if (n == "sdrate") s1 = V.sdrate;
    else if (n == "tnext") s1 = V.tnext; else if (n == "rho_search") s1 = V.rho_search;
    if (s1) {
        if (bytes != (int64_t) sizeof(double) * N) return fail(h, SSB_ERR_ARG, "field %s needs %lld bytes", name, (long long) sizeof(double) * N);
        k_unperm64<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, s1, h->d_stage, 1, 0);
        CK(cudaMemcpyAsync(dst, h->d_stage, bytes, cudaMemcpyDeviceToHost, st));
        CK(ssb_sync(h));
        return SSB_OK;
    }
    const int *i1 = nullptr;
    if (n == "type") i1 = V.type; else if (n == "solid") i1 = V.solid; else if (n == "nbr_count") i1 = V.nbr_count; else if (n == "id") i1 = V.id;
    if (i1) {
        if (bytes != (int64_t) sizeof(int) * N) return fail(h, SSB_ERR_ARG, "field %s needs %lld bytes", name, (long long) sizeof(int) * N);
        k_unperm32<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, i1, (int *) h->d_stage, 1, 0);
        CK(cudaMemcpyAsync(dst, h->d_stage, bytes, cudaMemcpyDeviceToHost, st));
        CK(ssb_sync(h));
        return SSB_OK;
    }
    // species-major blocks -> voxel-major [N][K] on the host side
    const double *blk = nullptr;
    int K = 0;
    if (n == "C") { blk = V.C; K = V.Sc; } else if (n == "Q") { blk = V.Q; K = V.Sc; } else if (n == "Ddiag") { blk = V.Ddiag; K = V.Sd; }
    else if (n == "rrate") { blk = V.rrate; K = V.Rd; }
    if (blk || n == "C" || n == "Q" || n == "Ddiag" || n == "rrate") {
        if (bytes != (int64_t) sizeof(double) * K * N) return fail(h, SSB_ERR_ARG, "field %s needs %lld bytes", name, (long long) sizeof(double) * K * N);
        for (int s = 0; s < K; s++) k_unperm64<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, blk + (size_t) s * N, h->d_stage, K, s);
        if (K > 0) CK(cudaMemcpyAsync(dst, h->d_stage, bytes, cudaMemcpyDeviceToHost, st));
        CK(ssb_sync(h));
        return SSB_OK;
    }
    if (n == "xx") {
        K = V.Sd;
        if (bytes != (int64_t) sizeof(unsigned) * K * N) return fail(h, SSB_ERR_ARG, "field xx needs %lld bytes", (long long) sizeof(unsigned) * K * N);
        for (int s = 0; s < K; s++) k_unperm32<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, (int *) V.xx + (size_t) s * N, (int *) h->d_stage, K, s);
        if (K > 0) CK(cudaMemcpyAsync(dst, h->d_stage, bytes, cudaMemcpyDeviceToHost, st));
        CK(ssb_sync(h));
        return SSB_OK;
    }
    return fail(h, SSB_ERR_ARG, "unknown field '%s'", name);
}

// ----------------------------------------------------------------------------------------------------
// state hand-over (slab re-partition, spatialpy_b200/slab.py): the inverse of ssb_get_field for the fields that make up
// a particle's state between two engine steps, and the step / Philox-epoch counters that go with them
// ----------------------------------------------------------------------------------------------------
extern "C" int ssb_set_field(ssb_handle *h, const char *name, const void *src, int64_t bytes) {
    if (!h || !name || !src) return SSB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    SsbView &V = h->V;
    const int N = h->N;
    cudaStream_t st = h->stream;
    std::string n(name);
    if ((size_t) bytes > h->stage_bytes) return fail(h, SSB_ERR_ARG, "field %s: %lld bytes exceed the staging buffer", name, (long long) bytes);
    double **v3 = nullptr;
    if (n == "x") v3 = V.x; else if (n == "v") v3 = V.v; else if (n == "vt") v3 = V.vt; else if (n == "F") v3 = V.F; else if (n == "Fbp") v3 = V.Fbp;
    if (v3) {
        h->look_valid = 0;       // (the look-ahead displacement was computed from the old x, v, F, Fbp)
        if (bytes != (int64_t) sizeof(double) * 3 * N) return fail(h, SSB_ERR_ARG, "field %s needs %lld bytes", name, (long long) sizeof(double) * 3 * N);
        if (n == "x" && V.static_domain) return fail(h, SSB_ERR_ARG, "positions of a static domain cannot be replaced (cached geometry)");
        CK(cudaMemcpyAsync(h->d_stage, src, bytes, cudaMemcpyHostToDevice, st));
        for (int d = 0; d < 3; d++) k_perm_in64<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, h->d_stage, v3[d], 3, d);
        if (n == "x") { h->lists_valid = 0; h->nbr_valid = 0; h->look_valid = 0; }       // candidate lists and storage order refer to the old positions
        CK(ssb_sync(h));
        return SSB_OK;
    }
    double *s1 = nullptr;
    if (n == "rho") s1 = V.rho; else if (n == "old_rho") s1 = V.old_rho; else if (n == "Frho") s1 = V.Frho; else if (n == "bvf_phi") s1 = V.bvf;
    else if (n == "nu") s1 = V.nu;
    if (s1) {
        if (bytes != (int64_t) sizeof(double) * N) return fail(h, SSB_ERR_ARG, "field %s needs %lld bytes", name, (long long) sizeof(double) * N);
        CK(cudaMemcpyAsync(h->d_stage, src, bytes, cudaMemcpyHostToDevice, st));
        k_perm_in64<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, h->d_stage, s1, 1, 0);
        CK(ssb_sync(h));
        return SSB_OK;
    }
    if (n == "C" || n == "Q") {            // host side voxel-major [N][S_c] -> species-major blocks
        double *blk = (n == "C") ? V.C : V.Q;
        const int K = V.Sc;
        if (bytes != (int64_t) sizeof(double) * K * N) return fail(h, SSB_ERR_ARG, "field %s needs %lld bytes", name, (long long) sizeof(double) * K * N);
        if (K > 0) CK(cudaMemcpyAsync(h->d_stage, src, bytes, cudaMemcpyHostToDevice, st));
        for (int s = 0; s < K; s++) k_perm_in64<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, h->d_stage, blk + (size_t) s * N, K, s);
        CK(ssb_sync(h));
        return SSB_OK;
    }
    if (n == "xx") {
        const int K = V.Sd;
        if (bytes != (int64_t) sizeof(unsigned) * K * N) return fail(h, SSB_ERR_ARG, "field xx needs %lld bytes", (long long) sizeof(unsigned) * K * N);
        if (K > 0) CK(cudaMemcpyAsync(h->d_stage, src, bytes, cudaMemcpyHostToDevice, st));
        for (int s = 0; s < K; s++) k_perm_in32<<<gridN(N), CORE_BLOCK, 0, st>>>(N, V.id, (const int *) h->d_stage, (int *) V.xx + (size_t) s * N, K, s);
        if (V.static_domain) h->rdme_initialized = 0;                  // propensities and clocks refer to the old populations
        CK(ssb_sync(h));
        return SSB_OK;
    }
    return fail(h, SSB_ERR_ARG, "field '%s' cannot be set", name);
}

extern "C" int ssb_get_step(ssb_handle *h, uint32_t *step, uint64_t *epoch) {
    if (!h || !step || !epoch) return SSB_ERR_ARG;
    *step = h->current_step;
    *epoch = h->epoch;
    return SSB_OK;
}

extern "C" int ssb_set_step(ssb_handle *h, uint32_t step, uint64_t epoch) {
    if (!h) return SSB_ERR_ARG;
    if (h->V.static_domain) return fail(h, SSB_ERR_ARG, "ssb_set_step is implemented for moving domains (a static domain keeps its step-0 geometry cache)");
    h->current_step = step;
    h->epoch = epoch;
    return SSB_OK;
}

extern "C" int ssb_get_neighbors(ssb_handle *h, int64_t *ptr, int32_t *idx, double *dist, double *dWdr, double *Dij, int64_t *nnz_out) {
    if (!h || !nnz_out) return SSB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    SsbView &V = h->V;
    const int N = h->N;
    cudaStream_t st = h->stream;
    std::vector<long long> cnt((size_t) N + 1, 0);
    long long *d_cnt = nullptr;
    CK(cudaMalloc((void **) &d_cnt, sizeof(long long) * (N + 1)));
    k_nbr_count_by_id<<<gridN(N), CORE_BLOCK, 0, st>>>(V, d_cnt);
    CK(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(long long) * N, cudaMemcpyDeviceToHost, st));
    CK(ssb_sync(h));
    std::vector<long long> p((size_t) N + 1, 0);
    for (int i = 0; i < N; i++) p[i + 1] = p[i] + cnt[i];
    const long long nnz = p[N];
    *nnz_out = nnz;
    if (!idx) { cudaFree(d_cnt); return SSB_OK; }
    if (!ptr || !dist || !dWdr || !Dij) { cudaFree(d_cnt); return SSB_ERR_ARG; }
    for (int i = 0; i <= N; i++) ptr[i] = p[i];
    CK(cudaMemcpyAsync(d_cnt, p.data(), sizeof(long long) * (N + 1), cudaMemcpyHostToDevice, st));
    int *d_idx = nullptr;
    double *d_a = nullptr;
    CK(cudaMalloc((void **) &d_idx, sizeof(int) * (size_t) (nnz + 1)));
    CK(cudaMalloc((void **) &d_a, sizeof(double) * 3 * (size_t) (nnz + 1)));
    k_nbr_export<<<gridN(N), CORE_BLOCK, 0, st>>>(V, d_cnt, d_idx, d_a, d_a + nnz, d_a + 2 * nnz);
    CK(cudaMemcpyAsync(idx, d_idx, sizeof(int) * nnz, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(dist, d_a, sizeof(double) * nnz, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(dWdr, d_a + nnz, sizeof(double) * nnz, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(Dij, d_a + 2 * nnz, sizeof(double) * nnz, cudaMemcpyDeviceToHost, st));
    CK(ssb_sync(h));
    cudaFree(d_cnt); cudaFree(d_idx); cudaFree(d_a);
    return SSB_OK;
}

// ----------------------------------------------------------------------------------------------------
// measurement hooks (bench.py): device-timed stepping and per-kernel-category timers
// ----------------------------------------------------------------------------------------------------
extern "C" int ssb_step_timed(ssb_handle *h, uint32_t nsteps, double *device_ms) {
    if (!h || !device_ms) return SSB_ERR_ARG;
    if (!h->unit) return fail(h, SSB_ERR_MODEL_UNIT, "no model unit loaded (ssb_load_kernels)");
    CK(cudaSetDevice(h->device));
    cudaEvent_t a, b;
    CK(cudaEventCreateWithFlags(&a, cudaEventBlockingSync)); CK(cudaEventCreateWithFlags(&b, cudaEventBlockingSync));
    CK(ssb_sync(h));
    CK(cudaEventRecord(a, h->stream));
    int rc = SSB_OK;
    for (uint32_t s = 0; s < nsteps && !rc; s++) rc = engine_step(h);
    CK(cudaEventRecord(b, h->stream));
    CK(ssb_sync(h));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    *device_ms = ms;
    if (rc) return rc;
    if (h->profile) prof_harvest(h);
    return check_device_error(h);
}

extern "C" int ssb_profile(ssb_handle *h, int enable) {
    if (!h) return SSB_ERR_ARG;
    prof_harvest(h);
    h->profile = enable ? 1 : 0;
    for (int c = 0; c < SSB_NCAT; c++) { h->cat_ms[c] = 0.0; h->cat_launches[c] = 0; }
    return SSB_OK;
}

extern "C" int ssb_profile_read(ssb_handle *h, int category, double *ms_total, int64_t *launches) {
    if (!h || category < 0 || category >= SSB_NCAT) return SSB_ERR_ARG;
    prof_harvest(h);
    if (ms_total) *ms_total = h->cat_ms[category];
    if (launches) *launches = h->cat_launches[category];
    return SSB_OK;
}

extern "C" int ssb_io_bytes(ssb_handle *h, int64_t *h2d, int64_t *d2h) {
    if (!h) return SSB_ERR_ARG;
    if (h2d) *h2d = h->h2d_bytes;
    if (d2h) *d2h = h->d2h_bytes;
    return SSB_OK;
}

extern "C" int ssb_nbr_stats(ssb_handle *h, int32_t *capacity, int64_t *total) {
    if (!h) return SSB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (capacity) *capacity = h->V.nbr_cap;
    if (total) {
        std::vector<int> c((size_t) h->N);
        CK(cudaMemcpy(c.data(), h->V.nbr_count, sizeof(int) * h->N, cudaMemcpyDeviceToHost));
        long long t = 0;
        for (int v : c) t += v;
        *total = t;
    }
    return SSB_OK;
}

// ----------------------------------------------------------------------------------------------------
// Slab decomposition support.  A rank's model = its owned particles + ghost copies (ssb_model.owned).  Ghosts run the same
// per-particle kernels as everyone else (predictor / corrector updates are deterministic functions of synced inputs), but their
// neighbour sweeps are skipped; after each sweep the owner's results overwrite the ghost copies:
//     PRE        cell list (if due) + predictor + neighbour search (if due) + force sweep   -> exchange group 0
//     CORRECTOR  corrector (+ Shepard filter)                                               -> exchange group 1
//     FINISH     BVF sweep, bounce-back, chemistry half step, BCs                           -> exchange group 2
//     RDME       global max Ddiag -> windows; one inbox exchange after every sSSA window; the step-end overshoot event
// Two drivers share these pieces: ssb_slab_step (native transport below: the product path) and ssb_step_phase + ssb_halo_*
// (caller-orchestrated, any transport — the CPU-tier protocol tests and the cross-check of the native path).
// ----------------------------------------------------------------------------------------------------
enum { PH_PRE = 0, PH_CORRECTOR = 1, PH_FINISH = 2, PH_RDME_PREP = 3, PH_RDME_INIT = 4, PH_RDME_WINDOW = 5, PH_RDME_CLOSE = 6, PH_END = 7,
       PH_RDME_MIN = 8, PH_RDME_EXTRA = 9 };

static int ensure_slot_map(ssb_handle *h) {
    if (!h->d_slot_of_id) CK(dalloc(h, &h->d_slot_of_id, (size_t) h->N));
    if (h->slot_dirty) {
        k_slot_of_id<<<gridN(h->N), CORE_BLOCK, 0, h->stream>>>(h->N, h->V.id, h->d_slot_of_id);
        h->slot_dirty = 0;
    }
    return SSB_OK;
}

static int rdme_window_phase(ssb_handle *h, bool closing, long long w) {
    SsbView &V = h->V;
    const double t0 = V.dt * h->current_step;
    const long long nwin = h->nwin;
    double lo, hi;
    if (!closing) {
        lo = t0 + V.dt * ((double) w / (double) nwin);
        hi = (w + 1 == nwin) ? t0 + V.dt : t0 + V.dt * ((double) (w + 1) / (double) nwin);
        h->windows++;
    } else { lo = hi = t0 + V.dt; }
    // closing window with w < 0: the delivery of the step-end overshoot event, which runs under the epoch rdme_extra_event reserved
    // for it (the numbering of the Philox epochs is the same on every path: +2 per step for the overshoot)
    const uint64_t epoch = (closing && w < 0) ? h->epoch - 1 : h->epoch++;
    if (h->unit->rdme_window(&V, lo, hi, h->tau, h->seed, epoch, h->inbox_buf, h->stream)) return fail(h, SSB_ERR_CUDA, "rdme_window launch failed");
    h->inbox_buf ^= 1;
    h->launches++;
    return SSB_OK;
}

extern "C" int ssb_step_phase(ssb_handle *h, int phase, double arg, double *out) {
    if (!h) return SSB_ERR_ARG;
    if (!h->unit) return fail(h, SSB_ERR_MODEL_UNIT, "no model unit loaded (ssb_load_kernels)");
    CK(cudaSetDevice(h->device));
    SsbView &V = h->V;
    const SsbModelUnit *u = h->unit;
    cudaStream_t st = h->stream;
    const unsigned step = h->current_step;
    if (V.static_domain) return fail(h, SSB_ERR_ARG, "phase stepping (slab decomposition) is implemented for moving domains");
    int rc;
    switch (phase) {
    case PH_PRE:
        if ((rc = mv_pre(h))) return rc;
        break;
    case PH_CORRECTOR:
        if ((rc = mv_corrector(h))) return rc;
        break;
    case PH_FINISH:
        if ((rc = mv_finish(h))) return rc;
        break;
    case PH_RDME_PREP: {
        double mx = 0.0;
        if ((rc = step_max_ddiag(h, &mx))) return rc;
        if (out) *out = mx;
        break;
    }
    case PH_RDME_INIT: {
        if ((rc = set_windows(h, arg))) return rc;       // arg = GLOBAL max Ddiag: every rank must use the same windows
        if (u->rdme_init(&V, V.dt * step, 0.0, h->tau, h->seed, h->epoch++, st)) return fail(h, SSB_ERR_CUDA, "rdme_init launch failed");
        h->launches++;
        h->rdme_initialized = 1;
        h->inbox_buf = 0;
        if (out) *out = (double) h->nwin;
        break;
    }
    case PH_RDME_WINDOW:
    case PH_RDME_CLOSE:
        if ((rc = rdme_window_phase(h, phase == PH_RDME_CLOSE, (long long) arg))) return rc;
        break;
    case PH_RDME_MIN: {          // -> *out = this rank's earliest pending event (the caller all-reduces the minimum)
        double tmin = INFINITY;
        if ((rc = rdme_min_time(h, &tmin))) return rc;
        if (out) *out = tmin;
        break;
    }
    case PH_RDME_EXTRA:          // arg = GLOBAL earliest pending event: only its owner fires; the caller syncs the inboxes and then
                                 // delivers (PH_RDME_CLOSE with arg < 0 reuses the epoch rdme_extra_event reserved for the delivery)
        if (!(V.flags & (SSB_FLAG_CORRECTED_NSM_SELECT | SSB_FLAG_NO_STEP_OVERSHOOT))) { if ((rc = rdme_extra_event(h, arg, false))) return rc; }
        break;
    case PH_END:
        if ((rc = mv_lookahead(h))) return rc;            // (after the caller's last halo unpack: ghost v, F, Fbp are final)
        h->current_step++;
        return check_device_error(h);
    default:
        return fail(h, SSB_ERR_ARG, "unknown phase %d", phase);
    }
    return SSB_OK;
}

// pack `n` particles (ids = particle ids of this rank's model, device array) of field group `group` into dev_out
// (n * width doubles, device memory owned by the caller, e.g. a torch tensor); synchronises the engine stream.
extern "C" int ssb_halo_pack(ssb_handle *h, int group, const int32_t *dev_ids, int32_t n, double *dev_out) {
    if (!h || group < 0 || group > 3) return SSB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    int rc = ensure_slot_map(h);
    if (rc) return rc;
    if (n > 0) k_halo_pack<<<gridN(n), CORE_BLOCK, 0, h->stream>>>(h->V, group, dev_ids, n, h->d_slot_of_id, dev_out);
    h->launches++;
    CK(ssb_sync(h));
    return SSB_OK;
}
extern "C" int ssb_halo_unpack(ssb_handle *h, int group, const int32_t *dev_ids, int32_t n, const double *dev_in) {
    if (!h || group < 0 || group > 3) return SSB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    int rc = ensure_slot_map(h);
    if (rc) return rc;
    if (n > 0) k_halo_unpack<<<gridN(n), CORE_BLOCK, 0, h->stream>>>(h->V, group, dev_ids, n, h->d_slot_of_id, dev_in);
    h->launches++;
    CK(ssb_sync(h));
    return SSB_OK;
}
// inbox of the sSSA window that just ran (the buffer the next window will read)
extern "C" int ssb_halo_inbox_pack(ssb_handle *h, const int32_t *dev_ids, int32_t n, uint32_t *dev_out) {
    if (!h) return SSB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    int rc = ensure_slot_map(h);
    if (rc) return rc;
    if (n > 0 && h->V.Sd > 0) k_inbox_pack<<<gridN(n), CORE_BLOCK, 0, h->stream>>>(h->V, h->inbox_buf ^ 1, dev_ids, n, h->d_slot_of_id, dev_out);
    h->launches++;
    CK(ssb_sync(h));
    return SSB_OK;
}
extern "C" int ssb_halo_inbox_add(ssb_handle *h, const int32_t *dev_ids, int32_t n, const uint32_t *dev_in) {
    if (!h) return SSB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    int rc = ensure_slot_map(h);
    if (rc) return rc;
    if (n > 0 && h->V.Sd > 0) k_inbox_add<<<gridN(n), CORE_BLOCK, 0, h->stream>>>(h->V, h->inbox_buf ^ 1, dev_ids, n, h->d_slot_of_id, dev_in, h->unit->block);
    h->launches++;
    CK(ssb_sync(h));
    return SSB_OK;
}
extern "C" int ssb_halo_width(ssb_handle *h, int group, int32_t *width) {
    if (!h || !width) return SSB_ERR_ARG;
    *width = group == 0 ? 7 + h->V.Sc : (group == 2 ? 4 : 1);
    return SSB_OK;
}

// ----------------------------------------------------------------------------------------------------
// Native slab transport (include/ssb.h: ssb_slab_*).  Every rank owns one RECEIVE WINDOW per slab face (device memory, exported
// through CUDA IPC, or by raw pointer when the neighbour rank lives in the same process) and one scalar BOARD (mapped by all ranks).
// A message is written by the SENDER's pack kernel straight into the receiver's window over NVLink; buffers are double-buffered by
// the parity of the channel's sequence number (a sender can be at most one message ahead of the receiver's consumption: it only
// sends message k+2 after it has received the neighbour's message k+1, which the neighbour sent after consuming message k).
// ----------------------------------------------------------------------------------------------------
struct HaloSide {
    int active = 0;
    int n_send = 0, n_recv = 0;
    int *d_send_ids = nullptr, *d_recv_ids = nullptr;
    char *win = nullptr;          // my window facing this neighbour
    size_t win_bytes = 0;
    char *peer = nullptr;         // the neighbour's window facing me
    int peer_ipc = 0;
    size_t off[HALO_NCH][2];      // buffer offsets in MY window   (group rows = n_recv, inbox entries <= n_send * Sd)
    size_t poff[HALO_NCH][2];     // buffer offsets in the PEER's  (group rows = n_send, inbox entries <= n_recv * Sd)
};
struct SlabComm {
    int rank = 0, world = 1, connected = 0;
    HaloSide side[2];             // 0 = rank-1, 1 = rank+1
    unsigned long long seq[HALO_NCH] = {0, 0, 0, 0};
    char *board = nullptr;
    unsigned long long *peer_board[SSB_BOARD_MAXW] = {nullptr};
    int peer_board_ipc[SSB_BOARD_MAXW] = {0};
    unsigned long long bseq[SSB_BOARD_NCH] = {0, 0, 0};
    unsigned *d_done = nullptr, *d_icount = nullptr;
    unsigned long long *d_red = nullptr;      // [SSB_BOARD_NCH] reduced scalars (device)
    unsigned long long *d_inf = nullptr;      // bit pattern of +inf
    cudaEvent_t ev_post[SSB_BOARD_NCH] = {nullptr}, ev_red[SSB_BOARD_NCH] = {nullptr}, ev_disp[2] = {nullptr, nullptr};
    double travel = 0.0, gdisp_max = 0.0;
    uint32_t steps = 0;           // steps run since the windows were set up
};
struct SlabBlob {                 // what a rank publishes about itself (exchanged by the caller: torch.distributed / a thread hub)
    int64_t pid;
    int32_t device, rank;
    int32_t n_send[2], n_recv[2];
    uint64_t ptr[3];              // window 0, window 1, board
    cudaIpcMemHandle_t ipc[3];
};

static size_t halo_layout(int rows, int inbox_entries, int Sc, size_t off[HALO_NCH][2]) {
    size_t o = HALO_HDR_BYTES;
    const int width[3] = {7 + Sc, 1, 4};
    for (int ch = 0; ch < 3; ch++)
        for (int q = 0; q < 2; q++) { off[ch][q] = o; o += (((size_t) rows * width[ch] * sizeof(double)) + 255) & ~(size_t) 255; }
    for (int q = 0; q < 2; q++) { off[HALO_CH_INBOX][q] = o; o += (((size_t) inbox_entries * sizeof(uint4)) + 255) & ~(size_t) 255; }
    return o;
}

extern "C" int ssb_slab_setup(ssb_handle *h, int32_t rank, int32_t world, const int32_t *send_lo, int32_t n_send_lo, const int32_t *recv_lo,
                              int32_t n_recv_lo, const int32_t *send_hi, int32_t n_send_hi, const int32_t *recv_hi, int32_t n_recv_hi) {
    if (!h || rank < 0 || rank >= world || world > SSB_BOARD_MAXW) return SSB_ERR_ARG;
    if (h->V.static_domain) return fail(h, SSB_ERR_ARG, "slab decomposition is implemented for moving domains");
    if (h->slab) return fail(h, SSB_ERR_ARG, "ssb_slab_setup: already set up");
    CK(cudaSetDevice(h->device));
    SlabComm *c = h->slab = new SlabComm();
    c->rank = rank; c->world = world;
    const int32_t *ids[2][2] = {{send_lo, recv_lo}, {send_hi, recv_hi}};
    const int cnt[2][2] = {{n_send_lo, n_recv_lo}, {n_send_hi, n_recv_hi}};
    for (int sd = 0; sd < 2; sd++) {
        HaloSide &S = c->side[sd];
        const int nb = sd ? rank + 1 : rank - 1;
        S.active = (nb >= 0 && nb < world) ? 1 : 0;
        if (!S.active) continue;
        S.n_send = cnt[sd][0]; S.n_recv = cnt[sd][1];
        if (S.n_send < 0 || S.n_recv < 0 || (S.n_send && !ids[sd][0]) || (S.n_recv && !ids[sd][1])) return SSB_ERR_ARG;
        CK(dalloc(h, &S.d_send_ids, (size_t) S.n_send)); CK(dalloc(h, &S.d_recv_ids, (size_t) S.n_recv));
        if (S.n_send) CK(cudaMemcpyAsync(S.d_send_ids, ids[sd][0], sizeof(int) * S.n_send, cudaMemcpyHostToDevice, h->stream));
        if (S.n_recv) CK(cudaMemcpyAsync(S.d_recv_ids, ids[sd][1], sizeof(int) * S.n_recv, cudaMemcpyHostToDevice, h->stream));
        S.win_bytes = halo_layout(S.n_recv, S.n_send * std::max(h->V.Sd, 1), h->V.Sc, S.off);
        halo_layout(S.n_send, S.n_recv * std::max(h->V.Sd, 1), h->V.Sc, S.poff);
        CK(cudaMalloc((void **) &S.win, S.win_bytes));          // (own allocation: exported whole through CUDA IPC)
        h->allocs.push_back(S.win);
        CK(cudaMemsetAsync(S.win, 0, HALO_HDR_BYTES, h->stream));
    }
    const size_t board_bytes = sizeof(unsigned long long) * 2 * SSB_BOARD_NCH * 2 * SSB_BOARD_MAXW;
    CK(cudaMalloc((void **) &c->board, board_bytes));
    h->allocs.push_back(c->board);
    CK(cudaMemsetAsync(c->board, 0, board_bytes, h->stream));
    if (!h->d_slot_of_id) CK(dalloc(h, &h->d_slot_of_id, (size_t) h->N));      // (no cudaMalloc once ranks may be waiting for each other)
    CK(dalloc(h, &c->d_done, 4)); CK(dalloc(h, &c->d_icount, 4)); CK(dalloc(h, &c->d_red, SSB_BOARD_NCH + 1)); CK(dalloc(h, &c->d_inf, 1));
    CK(cudaMemsetAsync(c->d_done, 0, 16, h->stream)); CK(cudaMemsetAsync(c->d_icount, 0, 16, h->stream));
    CK(cudaMemsetAsync(c->d_red, 0, sizeof(unsigned long long) * (SSB_BOARD_NCH + 1), h->stream));       // ([3] stays 0: the constant a rank without a value posts)
    const unsigned long long inf_bits = 0x7ff0000000000000ull;
    CK(cudaMemcpyAsync(c->d_inf, &inf_bits, sizeof(inf_bits), cudaMemcpyHostToDevice, h->stream));
    for (int k = 0; k < SSB_BOARD_NCH; k++) {
        CK(cudaEventCreateWithFlags(&c->ev_post[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_red[k], cudaEventDisableTiming));
    }
    for (int k = 0; k < 2; k++) CK(cudaEventCreateWithFlags(&c->ev_disp[k], cudaEventDisableTiming));
    // First launches behind us (see SsbModelUnit::warm): every kernel of a slab step once, on empty inputs.  Measured without this on
    // a B200 with two ranks as threads of one process: whichever rank reached a not-yet-loaded kernel while its neighbour was already
    // spinning on a flag waited for the neighbour's kernel to end — i.e. for the 10 s timeout of a message it had not sent yet.
    {
        cudaStream_t st = h->stream;
        if (h->unit && h->unit->warm(&h->V, st)) return fail(h, SSB_ERR_CUDA, "kernel warm-up failed: %s", cudaGetErrorString(cudaGetLastError()));
        HaloPackArgs P; HaloUnpackArgs U; BoardPeers B;
        memset(&P, 0, sizeof(P)); memset(&U, 0, sizeof(U)); memset(&B, 0, sizeof(B));
        P.done = c->d_done; P.icount = c->d_icount;
        k_slot_of_id<<<1, CORE_BLOCK, 0, st>>>(0, h->V.id, h->d_slot_of_id);
        k_halo_send<<<1, CORE_BLOCK, 0, st>>>(h->V, 0, P, h->d_slot_of_id);
        k_inbox_send<<<1, CORE_BLOCK, 0, st>>>(h->V, 0, P, h->d_slot_of_id);
        k_halo_wait<<<1, 32, 0, st>>>(nullptr, nullptr, 0ull, h->V.err_flag);
        k_halo_recv<<<1, CORE_BLOCK, 0, st>>>(h->V, 0, U, h->d_slot_of_id);
        k_inbox_recv<<<1, CORE_BLOCK, 0, st>>>(h->V, 0, U, h->d_slot_of_id, 128);
        k_board_post<<<1, 32, 0, st>>>(B, 0, 0, 0, 0ull, c->d_red + 3);
        k_board_reduce<<<1, 32, 0, st>>>((const unsigned long long *) c->board, 0, 0, 0ull, 0, c->d_red + 3, h->V.err_flag);
        k_min_time<<<1, CORE_BLOCK, 0, st>>>(0, h->V.blk_tmin, c->d_red + 3);
        SsbView V0 = h->V;
        V0.N = 0;
        k_lookahead<<<1, CORE_BLOCK, 0, st>>>(V0, h->d_look);
        // ... and the list build (a rank may rebuild while its neighbour already waits for the step's first message)
        PermTable T;
        memset(&T, 0, sizeof(T));
        k_cell_keys<<<1, CORE_BLOCK, 0, st>>>(0, h->grid, h->V.x[0], h->V.x[1], h->V.x[2], h->d_key, h->d_cell_count);
        k_scan_tiles<<<1, 256, 0, st>>>(0, h->d_cell_count, h->d_cell_start, h->d_tile_sums);
        k_scan_sums<<<1, 256, 0, st>>>(0, h->d_tile_sums, nullptr);
        k_scan_add<<<1, CORE_BLOCK, 0, st>>>(0, h->d_cell_start, h->d_tile_sums);
        k_scatter<<<1, CORE_BLOCK, 0, st>>>(0, h->d_key, h->d_cell_start, h->d_cursor, h->d_perm);
        k_sort_cells<<<1, CORE_BLOCK, 0, st>>>(0, h->d_cell_start, 0, h->d_perm, h->d_flags);
        k_permute<<<1, CORE_BLOCK, 0, st>>>(0, h->d_perm, T);
        k_permute_rows64<<<1, CORE_BLOCK, 0, st>>>(0, h->d_perm, h->V.C, h->C_alt);
        k_permute_rows32<<<1, CORE_BLOCK, 0, st>>>(0, h->d_perm, (const int *) h->V.xx, (int *) h->xx_alt);
        k_search<<<1, CORE_BLOCK, 0, st>>>(V0, h->grid, h->d_cell_start, h->d_flags + 1, h->d_maxbits + 1);
        k_iota<<<1, CORE_BLOCK, 0, st>>>(0, h->V.id);
        k_unperm64<<<1, CORE_BLOCK, 0, st>>>(0, h->V.id, h->V.rho, h->d_stage, 1, 0);
        k_unperm32<<<1, CORE_BLOCK, 0, st>>>(0, h->V.id, h->V.type, (int *) h->d_stage, 1, 0);
        CK(cudaMemsetAsync(c->d_red, 0, sizeof(unsigned long long) * (SSB_BOARD_NCH + 1), st));
        CK(cudaMemsetAsync(c->d_done, 0, 16, st));
        CK(cudaGetLastError());
    }
    CK(ssb_sync(h));
    return SSB_OK;
}

extern "C" int ssb_slab_blob_bytes(void) { return (int) sizeof(SlabBlob); }

extern "C" int ssb_slab_export(ssb_handle *h, void *blob, int64_t bytes) {
    if (!h || !h->slab || !blob || bytes < (int64_t) sizeof(SlabBlob)) return SSB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    SlabComm *c = h->slab;
    SlabBlob B;
    memset(&B, 0, sizeof(B));
    B.pid = (int64_t) getpid(); B.device = h->device; B.rank = c->rank;
    char *ptr[3] = {c->side[0].win, c->side[1].win, c->board};
    for (int k = 0; k < 3; k++) {
        B.ptr[k] = (uint64_t) (uintptr_t) ptr[k];
        if (ptr[k]) CK(cudaIpcGetMemHandle(&B.ipc[k], ptr[k]));
    }
    for (int sd = 0; sd < 2; sd++) { B.n_send[sd] = c->side[sd].n_send; B.n_recv[sd] = c->side[sd].n_recv; }
    memcpy(blob, &B, sizeof(B));
    return SSB_OK;
}

static int slab_map(ssb_handle *h, const SlabBlob &B, int k, char **out, int *is_ipc) {
    *out = nullptr; *is_ipc = 0;
    if (!B.ptr[k]) return fail(h, SSB_ERR_ARG, "rank %d exported no window %d", B.rank, k);
    if (B.pid == (int64_t) getpid()) {                   // same process (ranks as threads): the pointer itself, peer access if another GPU
        if (B.device != h->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(B.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(h, SSB_ERR_CUDA, "no peer access from GPU %d to GPU %d: %s", h->device, B.device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        *out = (char *) (uintptr_t) B.ptr[k];
        return SSB_OK;
    }
    void *q = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&q, B.ipc[k], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(h, SSB_ERR_CUDA, "cudaIpcOpenMemHandle (rank %d, window %d): %s", B.rank, k, cudaGetErrorString(e));
    *out = (char *) q; *is_ipc = 1;
    return SSB_OK;
}

// blobs = the SlabBlob of every rank, in rank order
extern "C" int ssb_slab_connect(ssb_handle *h, const void *blobs, int64_t bytes) {
    if (!h || !h->slab || !blobs) return SSB_ERR_ARG;
    SlabComm *c = h->slab;
    if (bytes != (int64_t) sizeof(SlabBlob) * c->world) return fail(h, SSB_ERR_ARG, "ssb_slab_connect: expected %d blobs of %d bytes", c->world, (int) sizeof(SlabBlob));
    CK(cudaSetDevice(h->device));
    const SlabBlob *B = (const SlabBlob *) blobs;
    int rc;
    for (int sd = 0; sd < 2; sd++) {
        HaloSide &S = c->side[sd];
        if (!S.active) continue;
        const SlabBlob &P = B[sd ? c->rank + 1 : c->rank - 1];
        const int theirs = sd ? 0 : 1;                   // my upper face is the neighbour's lower face
        if (P.n_recv[theirs] != S.n_send || P.n_send[theirs] != S.n_recv)
            return fail(h, SSB_ERR_ARG, "slab faces disagree: rank %d sends %d / expects %d, rank %d expects %d / sends %d", c->rank, S.n_send, S.n_recv,
                        P.rank, P.n_recv[theirs], P.n_send[theirs]);
        if ((rc = slab_map(h, P, theirs, &S.peer, &S.peer_ipc))) return rc;
    }
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank) { c->peer_board[r] = (unsigned long long *) c->board; continue; }
        char *q; int ipc;
        if ((rc = slab_map(h, B[r], 2, &q, &ipc))) return rc;
        c->peer_board[r] = (unsigned long long *) q; c->peer_board_ipc[r] = ipc;
    }
    c->connected = 1;
    return SSB_OK;
}

// unmap the neighbours' memory (every rank calls this, then all ranks meet, then the handles may be destroyed)
extern "C" int ssb_slab_disconnect(ssb_handle *h) {
    if (!h || !h->slab) return SSB_OK;
    SlabComm *c = h->slab;
    cudaSetDevice(h->device);
    if (h->stream) ssb_sync(h);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    for (int sd = 0; sd < 2; sd++) { if (c->side[sd].peer && c->side[sd].peer_ipc) cudaIpcCloseMemHandle(c->side[sd].peer); c->side[sd].peer = nullptr; }
    for (int r = 0; r < c->world; r++) { if (c->peer_board[r] && c->peer_board_ipc[r]) cudaIpcCloseMemHandle(c->peer_board[r]); c->peer_board[r] = nullptr; }
    c->connected = 0;
    return SSB_OK;
}

static void slab_reset(ssb_handle *h) {        // ssb_reset: the particles are back where the partition was made
    if (!h->slab) return;
    h->slab->travel = 0.0; h->slab->gdisp_max = 0.0; h->slab->steps = 0;
}
static void slab_free(ssb_handle *h) {
    if (!h->slab) return;
    ssb_slab_disconnect(h);
    SlabComm *c = h->slab;
    for (int k = 0; k < SSB_BOARD_NCH; k++) { if (c->ev_post[k]) cudaEventDestroy(c->ev_post[k]); if (c->ev_red[k]) cudaEventDestroy(c->ev_red[k]); }
    for (int k = 0; k < 2; k++) if (c->ev_disp[k]) cudaEventDestroy(c->ev_disp[k]);
    delete c;
    h->slab = nullptr;
}

// one field group to both neighbours and back: pack+send (peer writes, flags raised by the last CTA), wait, unpack
struct NvtxScope { explicit NvtxScope(const char *n) { nvtxRangePushA(n); } ~NvtxScope() { nvtxRangePop(); } };

static int slab_exchange(ssb_handle *h, int group) {
    NvtxScope nv("ssb:halo_exchange");
    SlabComm *c = h->slab;
    cudaStream_t st = h->stream;
    int rc = ensure_slot_map(h);
    if (rc) return rc;
    const unsigned long long seq = ++c->seq[group];
    const int q = (int) (seq & 1ull);
    HaloPackArgs P;
    HaloUnpackArgs U;
    memset(&P, 0, sizeof(P)); memset(&U, 0, sizeof(U));
    P.seq = seq; P.done = c->d_done; P.icount = c->d_icount;
    const unsigned long long *flag[2] = {nullptr, nullptr};
    for (int sd = 0; sd < 2; sd++) {
        HaloSide &S = c->side[sd];
        if (!S.active) continue;
        P.side[sd].ids = S.d_send_ids; P.side[sd].n = S.n_send;
        P.side[sd].peer_buf = S.peer + S.poff[group][q];
        P.side[sd].peer_flag = (unsigned long long *) (S.peer + 64 * group);
        U.side[sd].ids = S.d_recv_ids; U.side[sd].n = S.n_recv; U.side[sd].buf = S.win + S.off[group][q];
        flag[sd] = (const unsigned long long *) (S.win + 64 * group);
    }
    const int ns = P.side[0].n + P.side[1].n, nr = U.side[0].n + U.side[1].n;
    k_halo_send<<<std::max(gridN(ns), 1u), CORE_BLOCK, 0, st>>>(h->V, group, P, h->d_slot_of_id);
    k_halo_wait<<<1, 32, 0, st>>>(flag[0], flag[1], seq, h->V.err_flag);
    if (nr > 0) k_halo_recv<<<gridN(nr), CORE_BLOCK, 0, st>>>(h->V, group, U, h->d_slot_of_id);
    CK(cudaGetLastError());
    h->launches += 3;
    return SSB_OK;
}
// mail of the sSSA window that just ran: my ghosts' inboxes -> their owners (the neighbour's inbox of the same buffer)
static int slab_exchange_inbox(ssb_handle *h) {
    NvtxScope nv("ssb:halo_inbox");
    SlabComm *c = h->slab;
    cudaStream_t st = h->stream;
    if (h->V.Sd == 0) return SSB_OK;
    int rc = ensure_slot_map(h);
    if (rc) return rc;
    const unsigned long long seq = ++c->seq[HALO_CH_INBOX];
    const int q = (int) (seq & 1ull), buf = h->inbox_buf ^ 1;
    HaloPackArgs P;
    HaloUnpackArgs U;
    memset(&P, 0, sizeof(P)); memset(&U, 0, sizeof(U));
    P.seq = seq; P.done = c->d_done; P.icount = c->d_icount;
    const unsigned long long *flag[2] = {nullptr, nullptr};
    for (int sd = 0; sd < 2; sd++) {
        HaloSide &S = c->side[sd];
        if (!S.active) continue;
        P.side[sd].ids = S.d_recv_ids; P.side[sd].n = S.n_recv;              // my ghosts
        P.side[sd].peer_buf = S.peer + S.poff[HALO_CH_INBOX][q];
        P.side[sd].peer_flag = (unsigned long long *) (S.peer + 64 * HALO_CH_INBOX);
        P.side[sd].peer_count = (unsigned long long *) (S.peer + 512 + 64 * q);
        U.side[sd].ids = S.d_send_ids; U.side[sd].n = S.n_send; U.side[sd].buf = S.win + S.off[HALO_CH_INBOX][q];
        U.side[sd].count = (const unsigned long long *) (S.win + 512 + 64 * q);
        flag[sd] = (const unsigned long long *) (S.win + 64 * HALO_CH_INBOX);
    }
    const int ns = P.side[0].n + P.side[1].n;
    k_inbox_send<<<std::max(gridN(ns), 1u), CORE_BLOCK, 0, st>>>(h->V, buf, P, h->d_slot_of_id);
    k_halo_wait<<<1, 32, 0, st>>>(flag[0], flag[1], seq, h->V.err_flag);
    k_inbox_recv<<<64, CORE_BLOCK, 0, st>>>(h->V, buf, U, h->d_slot_of_id, h->unit->block);
    CK(cudaGetLastError());
    h->launches += 3;
    return SSB_OK;
}
// scalar all-reduce, device to device: post my value to every board (engine stream), reduce my own board on `on` (the engine
// stream, or the side stream when the host wants the result without stalling the engine stream) -> c->d_red[ch]
static int slab_allreduce(ssb_handle *h, int ch, const unsigned long long *src, bool take_min, cudaStream_t on) {
    NvtxScope nv("ssb:board_allreduce");
    SlabComm *c = h->slab;
    cudaStream_t st = h->stream;
    const unsigned long long seq = ++c->bseq[ch];
    BoardPeers P;
    memset(&P, 0, sizeof(P));
    for (int r = 0; r < c->world; r++) P.b[r] = c->peer_board[r];
    if (seq > 1) CK(cudaStreamWaitEvent(st, c->ev_red[ch], 0));          // my previous reduction of this channel has read its slots
    k_board_post<<<1, 32, 0, st>>>(P, c->world, c->rank, ch, seq, src);
    if (on != st) { CK(cudaEventRecord(c->ev_post[ch], st)); CK(cudaStreamWaitEvent(on, c->ev_post[ch], 0)); }
    k_board_reduce<<<1, 32, 0, on>>>((const unsigned long long *) c->board, c->world, ch, seq, take_min ? 1 : 0, c->d_red + ch, h->V.err_flag);
    CK(cudaEventRecord(c->ev_red[ch], on));
    CK(cudaGetLastError());
    h->launches += 2;
    return SSB_OK;
}

// n engine steps of one slab rank, everything stream-ordered; per step the host waits for two scalars only (its own look-ahead
// displacement, the global max Ddiag), both behind queued work.  Stops early — on every rank after the same step — when particles
// may have travelled `travel_limit` since the windows were set up (the caller re-partitions); *done = steps executed.
extern "C" int ssb_slab_step(ssb_handle *h, uint32_t nsteps, double travel_limit, uint32_t *done, double *travel) {
    if (!h || !h->slab) return SSB_ERR_ARG;
    if (!h->unit) return fail(h, SSB_ERR_MODEL_UNIT, "no model unit loaded (ssb_load_kernels)");
    SlabComm *c = h->slab;
    if (!c->connected && c->world > 1) return fail(h, SSB_ERR_ARG, "ssb_slab_step before ssb_slab_connect");
    CK(cudaSetDevice(h->device));
    SsbView &V = h->V;
    const SsbModelUnit *u = h->unit;
    cudaStream_t st = h->stream, aux = h->copy_stream;
    const bool overshoot = !(V.flags & (SSB_FLAG_CORRECTED_NSM_SELECT | SSB_FLAG_NO_STEP_OVERSHOOT));
    auto t0w = std::chrono::steady_clock::now();
    int rc;
    uint32_t s = 0;
    for (; s < nsteps; s++) {
        // global step displacement of the step before the previous one (two steps of lag keep the read free of any wait)
        if (c->steps >= 2) {
            CK(wait_event(c->ev_disp[c->steps & 1]));
            double g2;
            memcpy(&g2, &h->pin[8 + (c->steps & 1)], 8);
            const double g = sqrt(g2);
            c->travel += g;
            if (g > c->gdisp_max) c->gdisp_max = g;
            if (travel_limit > 0.0 && c->travel + 3.0 * c->gdisp_max > travel_limit) break;
        }
        const unsigned step = h->current_step;
        if ((rc = mv_pre(h))) return rc;
        if (V.Sd > 0) {
            if ((rc = slab_allreduce(h, 0, h->d_maxbits, false, aux))) return rc;
            CK(cudaMemcpyAsync(&h->pin[4], c->d_red + 0, sizeof(unsigned long long), cudaMemcpyDeviceToHost, aux));
            CK(cudaEventRecord(h->ev_maxd, aux));
        }
        if ((rc = slab_exchange(h, 0))) return rc;
        if ((rc = mv_corrector(h))) return rc;
        if ((rc = slab_exchange(h, 1))) return rc;
        if ((rc = mv_finish(h))) return rc;
        if ((rc = slab_exchange(h, 2))) return rc;
        if ((rc = mv_lookahead(h))) return rc;
        {
            // (a COLLECTIVE: every rank takes part whatever its own Verlet skin is — the skin is chosen per rank from local candidate
            // statistics, and a rank that rebuilds exact lists every step has no look-ahead: it contributes 0)
            if ((rc = slab_allreduce(h, 2, V.filter ? h->d_look + 1 : c->d_red + 3, false, aux))) return rc;
            CK(cudaMemcpyAsync(&h->pin[8 + (c->steps & 1)], c->d_red + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, aux));
            CK(cudaEventRecord(c->ev_disp[c->steps & 1], aux));
        }
        if (V.Sd > 0) {
            CK(wait_event(h->ev_maxd));                      // every rank's force sweep is done; corrector .. look-ahead are still queued
            double mx;
            memcpy(&mx, &h->pin[4], 8);
            if (getenv("SSB_SLAB_DEBUG")) {
                double loc; memcpy(&loc, &h->pin[3], 8);
                fprintf(stderr, "[slab %d/%d] step %u local max Ddiag %.6g global %.6g filter %d skin %g\n", c->rank, c->world, step, loc, mx, V.filter, h->skin);
            }
            h->ddiag_fresh = 0;
            if ((rc = set_windows(h, mx))) {                  // GLOBAL max Ddiag: every rank uses the same windows
                            if (getenv("SSB_SLAB_DEBUG")) fprintf(stderr, "[slab %d/%d] step %u: global max Ddiag %g is not usable\n", c->rank, c->world, step, mx);
                const int rc2 = check_device_error(h);       // (a lost halo message upstream explains a garbage maximum: report that)
                return rc2 ? rc2 : rc;
            }
            if (u->rdme_init(&V, V.dt * step, 0.0, h->tau, h->seed, h->epoch++, st)) return fail(h, SSB_ERR_CUDA, "rdme_init launch failed");
            h->launches++;
            h->rdme_initialized = 1;
            h->inbox_buf = 0;
            for (long long w = 0; w < h->nwin; w++) {
                if ((rc = rdme_window_phase(h, false, w))) return rc;
                if ((rc = slab_exchange_inbox(h))) return rc;
            }
            if ((rc = rdme_window_phase(h, true, 0))) return rc;
            if (overshoot) {
                // the reference's one event past the end of the step (simulate_rdme.cpp:233-238): the globally earliest pending clock
                // is reduced on the device and read by the two windows from device memory
                const int nchunks = (h->N + u->block - 1) / u->block;
                CK(cudaMemcpyAsync(h->d_maxbits + 1, c->d_inf, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
                k_min_time<<<gridN(nchunks), CORE_BLOCK, 0, st>>>(nchunks, V.blk_tmin, h->d_maxbits + 1);
                if ((rc = slab_allreduce(h, 1, h->d_maxbits + 1, true, st))) return rc;
                const double te = V.dt * (step + 1);
                h->epoch += 2;
                if (u->rdme_window_dev(&V, c->d_red + 1, te, 0, h->tau, h->seed, h->epoch - 2, h->inbox_buf, st)) return fail(h, SSB_ERR_CUDA, "rdme_window launch failed");
                h->inbox_buf ^= 1;
                if ((rc = slab_exchange_inbox(h))) return rc;
                if (u->rdme_window_dev(&V, c->d_red + 1, te, 1, h->tau, h->seed, h->epoch - 1, h->inbox_buf, st)) return fail(h, SSB_ERR_CUDA, "rdme_window launch failed");
                h->inbox_buf ^= 1;
                h->launches += 3;
            }
        }
        h->current_step++;
        c->steps++;
        if ((c->steps & 15) == 0) { if ((rc = check_device_error(h))) return rc; }
        if (h->cancel.load()) return fail(h, SSB_ERR_CANCELLED, "cancelled");
    }
    if ((rc = check_device_error(h))) return rc;
    h->step_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0w).count();
    if (done) *done = s;
    if (travel) *travel = c->travel;
    return SSB_OK;
}

// CUDA-event markers on the engine stream (device timing of a phased step loop)
extern "C" int ssb_mark(ssb_handle *h, int which) {
    if (!h) return SSB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (!h->mark_a) { CK(cudaEventCreate(&h->mark_a)); CK(cudaEventCreate(&h->mark_b)); }
    CK(cudaEventRecord(which == 0 ? h->mark_a : h->mark_b, h->stream));
    return SSB_OK;
}
extern "C" int ssb_mark_elapsed_ms(ssb_handle *h, double *ms) {
    if (!h || !ms || !h->mark_a) return SSB_ERR_ARG;
    CK(cudaEventSynchronize(h->mark_b));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, h->mark_a, h->mark_b));
    *ms = f;
    return SSB_OK;
}

extern "C" int ssb_skin_stats(ssb_handle *h, double *skin, double *step_disp_max, int64_t *rebuilds) {
    if (!h) return SSB_ERR_ARG;
    if (skin) *skin = h->skin;
    if (step_disp_max) *step_disp_max = h->step_disp_max;
    if (rebuilds) *rebuilds = h->rebuilds;
    return SSB_OK;
}
