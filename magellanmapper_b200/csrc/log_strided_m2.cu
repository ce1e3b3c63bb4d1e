// Strided-axis (y / z) convolution sweep, mode 2 instantiations.
#include <stdlib.h>
#include "log_kernels.cuh"

namespace mmb {

constexpr int kNB = 16;
constexpr int kThreads = 128;

constexpr int kCols = 128;
#ifndef MMB_MARCH_COLS
#define MMB_MARCH_COLS 64
#endif
constexpr int kMarchCols = MMB_MARCH_COLS;     // columns per CTA of the marching kernel
constexpr int kTileThreads = 256;

// marching kernel: ring of 2*RP + 2*STEP rows of both inputs, 512 threads per SM
template <int R, int G>
static int march(const float* in0, const float* in1, float* out0, float* out1, int n_axis,
                 int64_t inner, int64_t outer, const LogWeights& w, float scale,
                 cudaStream_t st) {
  constexpr int NBM = MMB_MARCH_NB;
  constexpr size_t smem = (size_t)2 * (2 * ((R + 7) / 8 * 8) + 2 * NBM * G) * kMarchCols * sizeof(float);
  auto kern = conv_march_kernel<R, 2, NBM, G, kMarchCols>;
  static bool configured = false;
  if (!configured) {
    MMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    configured = true;
  }
  // split the marched axis when columns x outer slices alone cannot fill the GPU
  // (thin chunks, small ROIs): segments of a multiple of STEP rows, >= 4 steps each
  constexpr int STEP = NBM * G;
  const int64_t ctas = cdiv(inner, kMarchCols) * outer;
  const int64_t want = (512 / (G * kMarchCols / 2)) * (int64_t)num_sms();
  int nseg = 1;
  if (ctas < want) {
    nseg = (int)cdiv(want, ctas);
    const int max_seg = (int)cdiv(n_axis, 4 * STEP);
    if (nseg > max_seg) nseg = max_seg;
    if (nseg < 1) nseg = 1;
  }
  const int seg_len = (int)cdiv(cdiv(n_axis, nseg), STEP) * STEP;
  dim3 grid((unsigned)cdiv(inner, kMarchCols), (unsigned)cdiv(n_axis, seg_len), (unsigned)outer);
  kern<<<grid, G * kMarchCols / 2, smem, st>>>(in0, in1, out0, out1, n_axis, inner,
                                               (int64_t)n_axis * inner, w, scale, seg_len);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

template <int R>
static int run(const float* in0, const float* in1, float* out0, float* out1, int n_axis,
               int64_t inner, int64_t outer, const LogWeights& w, float scale,
               cudaStream_t st) {
  // outer == 1 is the z sweep (whole y-x plane contiguous), otherwise the y sweep
  ProfScope ps(outer == 1 ? PROF_LOG_Z : PROF_LOG_Y, (double)inner * n_axis * outer, st);
  const uintptr_t bits = (uintptr_t)in0 | (uintptr_t)in1 | (uintptr_t)out0 | (uintptr_t)out1;
  const bool aligned = (bits & 15) == 0 && inner % 4 == 0;
  if constexpr (R <= 20) {
    if (aligned) {
      // thin volumes (the 12-plane trailing chunk row): half-size steps, so that 12 of 16
      // rows of a step are outputs instead of 12 of 32
      if (n_axis <= 16)
        return march<R, 2>(in0, in1, out0, out1, n_axis, inner, outer, w, scale, st);
      return march<R, MMB_MARCH_G>(in0, in1, out0, out1, n_axis, inner, outer, w, scale, st);
    }
  }
  if constexpr (R > 20 && R <= 32) {      // wider radii: direct kernel below
    if (aligned) {
      constexpr int NSEGS = R <= 24 ? 4 : 2;
      constexpr int ROWS = kNB * NSEGS + 2 * R;
      constexpr int NARR = 2 == 0 ? 1 : 2;
      constexpr size_t smem = (size_t)NARR * ROWS * kCols * sizeof(float);
      auto kern = conv_strided_tile_kernel<R, 2, kNB, NSEGS, kCols, kTileThreads>;
      static bool configured = false;
      if (!configured) {
        MMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        configured = true;
      }
      dim3 grid((unsigned)cdiv(inner, kCols), (unsigned)cdiv(n_axis, kNB * NSEGS), (unsigned)outer);
      kern<<<grid, kTileThreads, smem, st>>>(in0, in1, out0, out1, n_axis, inner,
                                             (int64_t)n_axis * inner, w, scale);
      MMB_CHECK_LAUNCH();
      return MMB_OK;
    }
  }
  {
    dim3 grid((unsigned)cdiv(inner, kThreads), (unsigned)cdiv(n_axis, kNB), (unsigned)outer);
    conv_strided_kernel<R, 2, kNB, kThreads><<<grid, kThreads, 0, st>>>(
        in0, in1, out0, out1, n_axis, inner, (int64_t)n_axis * inner, w, scale);
  }
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

int launch_strided_m2(int r, const float* in0, const float* in1, float* out0, float* out1,
                      int n_axis, int64_t inner, int64_t outer, const LogWeights& w,
                      float scale, cudaStream_t st) {
#define X(RR) if (r <= RR) return run<RR>(in0, in1, out0, out1, n_axis, inner, outer, w, scale, st);
  MMB_RADIUS_BUCKETS(X)
#undef X
  set_error("radius %d has no compiled bucket", r);
  return MMB_ERR_UNSUPPORTED;
}

}  // namespace mmb
