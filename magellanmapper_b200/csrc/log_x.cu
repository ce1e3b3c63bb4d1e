// Contiguous-axis (x) first sweep instantiations.
#include "log_kernels.cuh"

namespace mmb {

constexpr int kNBx = 16;
constexpr int kNSEG = 8;

template <int R>
static int run_x(const float* in, float* outA, float* outB, int64_t nrows, int X,
                 int64_t pitch, const LogWeights& w, cudaStream_t st) {
  dim3 grid((unsigned)cdiv(nrows, 32), (unsigned)cdiv(X, kNBx * kNSEG));
  ProfScope ps(PROF_LOG_X, (double)nrows * pitch, st);
  conv_x_first_kernel<R, kNBx, kNSEG><<<grid, 32 * kNSEG, 0, st>>>(in, outA, outB, nrows, X,
                                                                  pitch, w);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

int launch_x_first(int r, const float* in, float* outA, float* outB, int64_t nrows, int X,
                   int64_t pitch, const LogWeights& w, cudaStream_t st) {
#define X_(RR) if (r <= RR) return run_x<RR>(in, outA, outB, nrows, X, pitch, w, st);
  MMB_RADIUS_BUCKETS(X_)
#undef X_
  set_error("radius %d has no compiled bucket", r);
  return MMB_ERR_UNSUPPORTED;
}

}  // namespace mmb
