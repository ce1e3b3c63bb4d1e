// LoG scale entry points: weights, sweep dispatch, generic large-radius fallback.
#include <math.h>
#include <stdlib.h>
#include <vector>
#include "log_kernels.cuh"

namespace mmb {

int gaussian_radius(double sigma) { return (int)(4.0 * sigma + 0.5); }

// scipy.ndimage._filters._gaussian_kernel1d(sigma, order, radius) for order 0
// and 2, in float64, then rounded to float32.
static void kernel1d(double sigma, int r, std::vector<double>& g, std::vector<double>& h) {
  const double sigma2 = sigma * sigma;
  g.assign(r + 1, 0.0);
  h.assign(r + 1, 0.0);
  double sum = 0.0;
  for (int x = -r; x <= r; ++x) sum += exp(-0.5 / sigma2 * (double)(x * x));
  const double c0 = 1.0 / -sigma2;   // q(x) = c0 + c2 x^2
  const double c2 = c0 * c0;
  for (int x = 0; x <= r; ++x) {
    const double phi = exp(-0.5 / sigma2 * (double)(x * x)) / sum;
    g[x] = phi;
    h[x] = (c0 + (double)(x * x) * c2) * phi;
  }
}

int make_log_weights(double sigma, LogWeights* w) {
  const int r = gaussian_radius(sigma);
  if (r > kMaxRadius) return -1;
  std::vector<double> g, h;
  kernel1d(sigma, r, g, h);
  for (int k = 0; k <= kMaxRadius; ++k) {
    w->g[k] = k <= r ? (float)g[k] : 0.f;
    w->h[k] = k <= r ? (float)h[k] : 0.f;
    w->gh[k] = make_float2(w->g[k], w->h[k]);
  }
  return r;
}

// One thread per output voxel; any axis, any mode, any radius.
__global__ void conv_generic_kernel(const float* __restrict__ in0,
                                    const float* __restrict__ in1,
                                    float* __restrict__ out0, float* __restrict__ out1,
                                    int Z, int Y, int X, int64_t pitch, int axis, int mode,
                                    const float* __restrict__ g, const float* __restrict__ h,
                                    int r, float scale) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int z = blockIdx.z;
  if (x >= X) return;
  const int n = axis == 0 ? Z : (axis == 1 ? Y : X);
  const int c = axis == 0 ? z : (axis == 1 ? y : x);
  const int64_t stride = axis == 0 ? (int64_t)Y * pitch : (axis == 1 ? pitch : 1);
  const int64_t here = ((int64_t)z * Y + y) * pitch + x;
  const int64_t base = here - (int64_t)c * stride;
  float a0 = 0.f, a1 = 0.f;
  for (int t = -r; t <= r; ++t) {
    const int q = reflect_index(c + t, n);
    const int wi = t < 0 ? -t : t;
    const float v0 = in0[base + (int64_t)q * stride];
    if (mode == MODE_FIRST) {
      a0 = fmaf(g[wi], v0, a0);
      a1 = fmaf(h[wi], v0, a1);
    } else {
      const float v1 = in1[base + (int64_t)q * stride];
      if (mode == MODE_MID) {
        a0 = fmaf(g[wi], v0, a0);
        a1 = fmaf(h[wi], v0, a1);
        a1 = fmaf(g[wi], v1, a1);
      } else {
        a0 = fmaf(h[wi], v0, a0);
        a0 = fmaf(g[wi], v1, a0);
      }
    }
  }
  if (mode == MODE_LAST) {
    out0[here] = a0 * scale;
  } else {
    out0[here] = a0;
    out1[here] = a1;
  }
}

static int run_generic(const float* in0, const float* in1, float* out0, float* out1, int Z,
                       int Y, int X, int64_t pitch, int axis, int mode, double sigma,
                       float scale, cudaStream_t st) {
  const int r = gaussian_radius(sigma);
  std::vector<double> g, h;
  kernel1d(sigma, r, g, h);
  std::vector<float> gh(2 * (r + 1));
  for (int k = 0; k <= r; ++k) { gh[k] = (float)g[k]; gh[r + 1 + k] = (float)h[k]; }
  float* d = nullptr;
  MMB_CHECK_CUDA(cudaMallocAsync((void**)&d, gh.size() * sizeof(float), st));
  MMB_CHECK_CUDA(cudaMemcpyAsync(d, gh.data(), gh.size() * sizeof(float),
                                 cudaMemcpyHostToDevice, st));
  MMB_CHECK_CUDA(cudaStreamSynchronize(st));   // gh is a stack-lifetime host buffer
  dim3 grid((unsigned)cdiv(X, 128), (unsigned)Y, (unsigned)Z);
  conv_generic_kernel<<<grid, 128, 0, st>>>(in0, in1, out0, out1, Z, Y, X, pitch, axis, mode,
                                            d, d + r + 1, r, scale);
  MMB_CHECK_LAUNCH();
  MMB_CHECK_CUDA(cudaFreeAsync(d, st));
  return MMB_OK;
}

int launch_strided(int mode, int r, const float* in0, const float* in1, float* out0,
                   float* out1, int n_axis, int64_t inner, int64_t outer,
                   const LogWeights& w, float scale, cudaStream_t st) {
  switch (mode) {
    case MODE_FIRST: return launch_strided_m0(r, in0, in1, out0, out1, n_axis, inner, outer, w, scale, st);
    case MODE_MID:   return launch_strided_m1(r, in0, in1, out0, out1, n_axis, inner, outer, w, scale, st);
    default:         return launch_strided_m2(r, in0, in1, out0, out1, n_axis, inner, outer, w, scale, st);
  }
}

static int log_pass(const float* in0, const float* in1, float* out0, float* out1, int Z, int Y,
                    int X, int64_t pitch, int axis, int mode, double sigma, double scale,
                    cudaStream_t st) {
  LogWeights w;
  const int r = make_log_weights(sigma, &w);
  const bool fast = r >= 0 && Y <= 65535 && Z <= 65535 &&
                    !(axis == 2 && mode != MODE_FIRST);
  if (!fast)
    return run_generic(in0, in1, out0, out1, Z, Y, X, pitch, axis, mode, sigma, (float)scale, st);
  // the sweeps see pitch-padded rows; the profile counts the true voxels
  prof_set_unit_scale((double)X / (double)pitch);
  int rc;
  if (axis == 2) rc = launch_x_first(r, in0, out0, out1, (int64_t)Z * Y, X, pitch, w, st);
  else if (axis == 1)
    rc = launch_strided(mode, r, in0, in1, out0, out1, Y, pitch, Z, w, (float)scale, st);
  else
    rc = launch_strided(mode, r, in0, in1, out0, out1, Z, (int64_t)Y * pitch, 1, w,
                        (float)scale, st);
  prof_set_unit_scale(1.0);
  return rc;
}

int log_scale_impl(const float* in, float* out, float* work, int Z, int Y, int X, int64_t pitch,
                   double sigma, cudaStream_t st) {
  const int64_t vol = (int64_t)Z * Y * pitch;
  float* A = work;
  float* B = work + vol;
  float* C = work + 2 * vol;
  float* D = work + 3 * vol;
  int rc = MMB_ERR_UNSUPPORTED;
  static const bool no_fused = getenv("MMB_NO_FUSED_XY") != nullptr;   // developer A/B switch
  if (!no_fused) {
    LogWeights w;
    const int r = make_log_weights(sigma, &w);
    if (r >= 0 && Y <= 65535) {
      prof_set_unit_scale((double)X / (double)pitch);
      rc = launch_xy_fused(r, in, C, D, Z, Y, X, pitch, w, st);
      prof_set_unit_scale(1.0);
    }
  }
  if (rc == MMB_ERR_UNSUPPORTED) {
    rc = log_pass(in, nullptr, A, B, Z, Y, X, pitch, 2, MODE_FIRST, sigma, 1.0, st);
    if (rc) return rc;
    rc = log_pass(A, B, C, D, Z, Y, X, pitch, 1, MODE_MID, sigma, 1.0, st);
  }
  if (rc) return rc;
  return log_pass(C, D, out, nullptr, Z, Y, X, pitch, 0, MODE_LAST, sigma, -(sigma * sigma), st);
}

}  // namespace mmb

using namespace mmb;

extern "C" int64_t mmb_log_work_bytes(int Z, int Y, int64_t pitch) {
  return 4 * (int64_t)Z * Y * pitch * (int64_t)sizeof(float);
}

extern "C" int mmb_log_pass(const float* in0, const float* in1, float* out0, float* out1, int Z,
                            int Y, int X, int64_t pitch, int axis, int mode, double sigma,
                            double scale, void* stream) {
  MMB_REQUIRE(in0 && out0, "null buffer");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0 && pitch >= X, "bad shape");
  MMB_REQUIRE(axis >= 0 && axis <= 2 && mode >= 0 && mode <= 2, "bad axis/mode");
  MMB_REQUIRE(mode == MODE_FIRST || in1, "mode needs two inputs");
  MMB_REQUIRE(mode == MODE_LAST || out1, "mode needs two outputs");
  MMB_REQUIRE(sigma > 0, "sigma must be positive");
  return log_pass(in0, in1, out0, out1, Z, Y, X, pitch, axis, mode, sigma, scale,
                  (cudaStream_t)stream);
}

extern "C" int mmb_log_xy_fused(const float* in, float* outC, float* outD, int Z, int Y, int X,
                                int64_t pitch, double sigma, void* stream) {
  MMB_REQUIRE(in && outC && outD, "null buffer");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0 && pitch >= X, "bad shape");
  MMB_REQUIRE(sigma > 0, "sigma must be positive");
  LogWeights w;
  const int r = make_log_weights(sigma, &w);
  if (r < 0 || Y > 65535) {
    set_error("radius of sigma %g is outside the fused sweep", sigma);
    return MMB_ERR_UNSUPPORTED;
  }
  prof_set_unit_scale((double)X / (double)pitch);
  const int rc = launch_xy_fused(r, in, outC, outD, Z, Y, X, pitch, w, (cudaStream_t)stream);
  prof_set_unit_scale(1.0);
  if (rc == MMB_ERR_UNSUPPORTED) set_error("shape or radius outside the fused sweep");
  return rc;
}

extern "C" int mmb_log_scale(const float* in, float* out, void* work, int Z, int Y, int X,
                             int64_t pitch, double sigma, void* stream) {
  MMB_REQUIRE(in && out && work, "null buffer");
  MMB_REQUIRE(in != out, "out may not alias in");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0 && pitch >= X, "bad shape");
  MMB_REQUIRE(sigma > 0, "sigma must be positive");
  return log_scale_impl(in, out, (float*)work, Z, Y, X, pitch, sigma, (cudaStream_t)stream);
}
