// Strided-axis (y / z) convolution sweep, mode 0 instantiations.
#include "log_kernels.cuh"

namespace mmb {

constexpr int kNB = 16;
constexpr int kThreads = 128;


template <int R>
static int run(const float* in0, const float* in1, float* out0, float* out1, int n_axis,
               int64_t inner, int64_t outer, const LogWeights& w, float scale,
               cudaStream_t st) {
  // outer == 1 is the z sweep (whole y-x plane contiguous), otherwise the y sweep
  ProfScope ps(outer == 1 ? PROF_LOG_Z : PROF_LOG_Y, (double)inner * n_axis * outer, st);
  // MODE_FIRST along a strided axis is not on the detector's path (x is always the
  // first sweep); mmb_log_pass still offers it, served by the direct kernel only
  {
    dim3 grid((unsigned)cdiv(inner, kThreads), (unsigned)cdiv(n_axis, kNB), (unsigned)outer);
    conv_strided_kernel<R, 0, kNB, kThreads><<<grid, kThreads, 0, st>>>(
        in0, in1, out0, out1, n_axis, inner, (int64_t)n_axis * inner, w, scale);
  }
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

int launch_strided_m0(int r, const float* in0, const float* in1, float* out0, float* out1,
                      int n_axis, int64_t inner, int64_t outer, const LogWeights& w,
                      float scale, cudaStream_t st) {
#define X(RR) if (r <= RR) return run<RR>(in0, in1, out0, out1, n_axis, inner, outer, w, scale, st);
  MMB_RADIUS_BUCKETS(X)
#undef X
  set_error("radius %d has no compiled bucket", r);
  return MMB_ERR_UNSUPPORTED;
}

}  // namespace mmb
