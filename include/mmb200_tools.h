/* mmb200_tools.h - bench / test utilities exported by libmmb200.so.
 *
 * NOT part of the reference-facing boundary (that is include/mmb200.h): nothing in
 * sanderslab/magellanmapper corresponds to these.  They exist so that bench.py and the
 * parity tests can build the synthetic inputs BASELINE.json names at sizes that never fit
 * one host array (2048 x 8192 x 8192 uint16 = 275 GB).
 */
#ifndef MMB200_TOOLS_H
#define MMB200_TOOLS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Synthetic nuclei volume (SURVEY.md section 8d recipe: N(400, 30) background, Gaussian
 * spots sigma0 ~ U(2.5, 4.5), amplitude ~ U(0.3, 0.9) * 65535, `density` centres per
 * voxel) as a pure function of (seed, global z, y, x): fills the dense (Z, Y, X) uint16
 * DEVICE box whose first voxel sits at (z_off, y_off, x_off) of the unbounded volume.
 * Any two boxes agree bit for bit where they overlap.  Asynchronous on `stream`.        */
int mmb_synth_nuclei(uint16_t* out, int Z, int Y, int X, int64_t z_off, int64_t y_off,
                     int64_t x_off, uint64_t seed, double density, void* stream);

/* The fused x -> y sweep of mmb_log_scale on its own (csrc/log_xy.cu): C = g_y*(g_x*in),
 * D = h_y*(g_x*in) + g_y*(h_x*in).  MMB_ERR_UNSUPPORTED for the shapes and radii that
 * mmb_log_scale serves with the two separate sweeps (radius > 20, Y < 32, X < 32).     */
int mmb_log_xy_fused(const float* in, float* outC, float* outD, int Z, int Y, int X,
                     int64_t pitch, double sigma, void* stream);

/* Developer check: while on, every kernel launch of the library is followed (on the legacy
 * default stream) by a kernel that fills the shared memory of every SM with 0xFFFFFFFF.
 * A kernel that reads shared memory it never wrote then changes its results - shared
 * memory has no initcheck.  Single-stream runs only (tools/poison_check.py).            */
int mmb_debug_smem_poison(int on);

#ifdef __cplusplus
}
#endif
#endif /* MMB200_TOOLS_H */
