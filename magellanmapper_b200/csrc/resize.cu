// Isotropic resize and spectral unmixing: the two steps of detector.detect_blobs that sit
// between preprocessing and blob_log (magmap/cv/detector.py:893-897, 910-921).
//
// make_isotropic (magmap/cv/cv_nd.py:1071-1106) calls rescale_resize (:1109-1167), i.e.
//   skimage.transform.resize(roi, isotropic_shape, mode='reflect', preserve_range=True)
// followed by .astype(roi.dtype).  scikit-image (0.19+) evaluates that as
//   image -> float64 (float32 stays float32)
//   anti_aliasing = any(out < in)  (never for bool):  scipy.ndimage.gaussian_filter with
//       sigma = max(0, (in / out - 1) / 2) per axis, truncate 4, mode 'mirror'
//       ('reflect' of np.pad is 'mirror' of scipy.ndimage; 'edge' is 'nearest')
//   scipy.ndimage.zoom(order=1, mode, grid_mode=True): output o samples the input at
//       (o + 0.5) * in / out - 0.5, mapped back into range by the boundary mode, with the
//       two linear weights (1 - t, t) per axis, products accumulated tap by tap in C order
//   np.clip to the input's range (a no-op for convex weights up to rounding).
// Everything here is evaluated in float64 with separately rounded multiplies and adds, as
// the C code of scipy does, so that the truncating cast back to an integer dtype lands on
// the same integer; the result is stored as float32 (exact for uint8 / uint16 values).
#include <math.h>
#include <type_traits>
#include <vector>
#include "common.cuh"

namespace mmb {

enum { RESIZE_MIRROR = 0, RESIZE_NEAREST = 1 };

// scipy's map_coordinate for 'mirror' / 'nearest' (ni_interpolation.c)
__device__ __forceinline__ double map_coord(double in, int len, int mode) {
  if (mode == RESIZE_NEAREST) return in < 0 ? 0.0 : (in > len - 1 ? (double)(len - 1) : in);
  if (in < 0) {
    if (len <= 1) return 0.0;
    const int sz2 = 2 * len - 2;
    in = sz2 * (double)(long long)(-in / sz2) + in;
    return in <= 1 - len ? in + sz2 : -in;
  }
  if (in > len - 1) {
    if (len <= 1) return 0.0;
    const int sz2 = 2 * len - 2;
    in -= sz2 * (double)(long long)(in / sz2);
    if (in >= len) in = sz2 - in;
    return in;
  }
  return in;
}

__device__ __forceinline__ int map_index(int i, int len, int mode) {   // taps next to a face
  if (i < 0) return mode == RESIZE_NEAREST ? 0 : (len > 1 ? -i : 0);
  if (i > len - 1) return mode == RESIZE_NEAREST ? len - 1 : (len > 1 ? 2 * (len - 1) - i : 0);
  return i;
}

struct ZoomGeom {
  int Zi, Yi, Xi, Zo, Yo, Xo;
  int64_t sz, sy, sx;          // input element strides
  int64_t pitch;               // output: [Zo][Yo][pitch] float32
  double fz, fy, fx;           // in / out per axis
  int mode, truncate;
};

template <typename T>
__global__ void __launch_bounds__(256)
zoom_linear_kernel(const T* __restrict__ in, float* __restrict__ out,
                   const __grid_constant__ ZoomGeom g) {
  const int xo = blockIdx.x * 256 + threadIdx.x;
  const int yo = blockIdx.y, zo = blockIdx.z;
  if (xo >= g.Xo) return;
  int i0[3], i1[3];
  double w0[3], w1[3];
  const int o[3] = {zo, yo, xo}, n[3] = {g.Zi, g.Yi, g.Xi};
  const double f[3] = {g.fz, g.fy, g.fx};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double cc = __dadd_rn(__dmul_rn((double)o[a] + 0.5, f[a]), -0.5);
    cc = map_coord(cc, n[a], g.mode);
    const double fl = floor(cc);
    const double t = cc - fl;
    i0[a] = map_index((int)fl, n[a], g.mode);
    i1[a] = map_index((int)fl + 1, n[a], g.mode);
    w0[a] = 1.0 - t; w1[a] = t;
  }
  double acc = 0.0;
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int64_t at = (int64_t)(dz ? i1[0] : i0[0]) * g.sz +
                           (int64_t)(dy ? i1[1] : i0[1]) * g.sy + (int64_t)(dx ? i1[2] : i0[2]) * g.sx;
        double c = (double)in[at];
        c = __dmul_rn(c, dz ? w1[0] : w0[0]);
        c = __dmul_rn(c, dy ? w1[1] : w0[1]);
        c = __dmul_rn(c, dx ? w1[2] : w0[2]);
        acc = __dadd_rn(acc, c);
      }
  if (g.truncate) acc = trunc(acc);          // .astype(integer dtype)
  out[((int64_t)zo * g.Yo + yo) * g.pitch + xo] = (float)acc;
}

// strided T -> dense float64 [Z][Y][X]
template <typename T>
__global__ void __launch_bounds__(256)
to_double_kernel(const T* __restrict__ in, int64_t sz, int64_t sy, int64_t sx, int Y, int X,
                 double* __restrict__ out) {
  const int x = blockIdx.x * 256 + threadIdx.x;
  if (x >= X) return;
  out[((int64_t)blockIdx.z * Y + blockIdx.y) * X + x] =
      (double)in[(int64_t)blockIdx.z * sz + (int64_t)blockIdx.y * sy + (int64_t)x * sx];
}

// scipy.ndimage.correlate1d of a dense float64 volume with a symmetric kernel w[0..r]
// along `axis`, accumulated from the lowest tap upwards
__global__ void __launch_bounds__(256)
blur_f64_kernel(const double* __restrict__ in, double* __restrict__ out, int Z, int Y, int X,
                int axis, const double* __restrict__ w, int r, int mode) {
  const int x = blockIdx.x * 256 + threadIdx.x;
  const int y = blockIdx.y, z = blockIdx.z;
  if (x >= X) return;
  const int n = axis == 0 ? Z : (axis == 1 ? Y : X);
  const int c = axis == 0 ? z : (axis == 1 ? y : x);
  const int64_t stride = axis == 0 ? (int64_t)Y * X : (axis == 1 ? X : 1);
  const int64_t here = ((int64_t)z * Y + y) * X + x;
  const double* base = in + (here - (int64_t)c * stride);
  double acc = 0.0;
  for (int t = -r; t <= r; ++t) {
    int q = c + t;
    if (q < 0 || q > n - 1) q = (int)map_coord((double)q, n, mode);
    acc = __dadd_rn(acc, __dmul_rn(w[t < 0 ? -t : t], base[(int64_t)q * stride]));
  }
  out[here] = acc;
}

// target = max(target - factor * other, 0): one step of the spectral unmixing loop
__global__ void __launch_bounds__(256)
unmix_kernel(float* __restrict__ target, const float* __restrict__ other, int Y, int X,
             int64_t pitch, float factor) {
  const int x = blockIdx.x * 256 + threadIdx.x;
  if (x >= X) return;
  const int64_t at = ((int64_t)blockIdx.z * Y + blockIdx.y) * pitch + x;
  const float v = target[at] - factor * other[at];
  target[at] = v < 0.f ? 0.f : v;
}

template <typename T>
static int run_resize(const T* in, ZoomGeom g, float* out, cudaStream_t st) {
  const int ni[3] = {g.Zi, g.Yi, g.Xi}, no[3] = {g.Zo, g.Yo, g.Xo};
  const bool anti_alias = no[0] < ni[0] || no[1] < ni[1] || no[2] < ni[2];
  dim3 ogrid((unsigned)cdiv(g.Xo, 256), (unsigned)g.Yo, (unsigned)g.Zo);
  if (!anti_alias) {
    zoom_linear_kernel<T><<<ogrid, 256, 0, st>>>(in, out, g);
    MMB_CHECK_LAUNCH();
    return MMB_OK;
  }
  // gaussian_filter walks the axes in order; an axis with sigma <= 1e-15 is skipped
  const int64_t nvox = (int64_t)g.Zi * g.Yi * g.Xi;
  double *d0 = nullptr, *d1 = nullptr, *dw = nullptr;
  MMB_CHECK_CUDA(cudaMallocAsync((void**)&d0, (size_t)nvox * 8, st));
  MMB_CHECK_CUDA(cudaMallocAsync((void**)&d1, (size_t)nvox * 8, st));
  dim3 igrid((unsigned)cdiv(g.Xi, 256), (unsigned)g.Yi, (unsigned)g.Zi);
  to_double_kernel<T><<<igrid, 256, 0, st>>>(in, g.sz, g.sy, g.sx, g.Yi, g.Xi, d0);
  MMB_CHECK_LAUNCH();
  const double f[3] = {g.fz, g.fy, g.fx};
  int rc = MMB_OK;
  for (int a = 0; a < 3 && rc == MMB_OK; ++a) {
    double sigma = (f[a] - 1.0) / 2.0;
    if (sigma < 0) sigma = 0;
    if (sigma <= 1e-15) continue;
    const int r = (int)(4.0 * sigma + 0.5);
    std::vector<double> w(r + 1);
    double sum = 0.0;
    for (int t = -r; t <= r; ++t) sum += exp(-0.5 / (sigma * sigma) * (double)(t * t));
    for (int t = 0; t <= r; ++t) w[t] = exp(-0.5 / (sigma * sigma) * (double)(t * t)) / sum;
    if (dw) MMB_CHECK_CUDA(cudaFreeAsync(dw, st));
    MMB_CHECK_CUDA(cudaMallocAsync((void**)&dw, (size_t)(r + 1) * 8, st));
    MMB_CHECK_CUDA(cudaMemcpyAsync(dw, w.data(), (size_t)(r + 1) * 8, cudaMemcpyHostToDevice, st));
    MMB_CHECK_CUDA(cudaStreamSynchronize(st));          // w is a stack-lifetime host buffer
    blur_f64_kernel<<<igrid, 256, 0, st>>>(d0, d1, g.Zi, g.Yi, g.Xi, a, dw, r, g.mode);
    MMB_CHECK_LAUNCH();
    double* t = d0; d0 = d1; d1 = t;
  }
  ZoomGeom gd = g;
  gd.sz = (int64_t)g.Yi * g.Xi; gd.sy = g.Xi; gd.sx = 1;
  zoom_linear_kernel<double><<<ogrid, 256, 0, st>>>(d0, out, gd);
  MMB_CHECK_LAUNCH();
  MMB_CHECK_CUDA(cudaFreeAsync(d0, st));
  MMB_CHECK_CUDA(cudaFreeAsync(d1, st));
  if (dw) MMB_CHECK_CUDA(cudaFreeAsync(dw, st));
  return rc;
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_resize_linear(const void* in, int dtype, const int64_t in_strides[3], int Z,
                                 int Y, int X, float* out, int Zo, int Yo, int Xo,
                                 int64_t pitch_out, int edge_mode, void* stream) {
  MMB_REQUIRE(in && out && in_strides, "null buffer");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0 && Zo > 0 && Yo > 0 && Xo > 0 && pitch_out >= Xo,
              "bad shape");
  MMB_REQUIRE(Y <= 65535 && Z <= 65535 && Yo <= 65535 && Zo <= 65535,
              "Y and Z must be <= 65535");
  MMB_REQUIRE(edge_mode == RESIZE_MIRROR || edge_mode == RESIZE_NEAREST, "unknown edge mode");
  ZoomGeom g;
  g.Zi = Z; g.Yi = Y; g.Xi = X; g.Zo = Zo; g.Yo = Yo; g.Xo = Xo;
  g.sz = in_strides[0]; g.sy = in_strides[1]; g.sx = in_strides[2];
  g.pitch = pitch_out;
  g.fz = (double)Z / (double)Zo; g.fy = (double)Y / (double)Yo; g.fx = (double)X / (double)Xo;
  g.mode = edge_mode;
  g.truncate = dtype == MMB_U8 || dtype == MMB_U16;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case MMB_U8:  return run_resize<uint8_t>((const uint8_t*)in, g, out, st);
    case MMB_U16: return run_resize<uint16_t>((const uint16_t*)in, g, out, st);
    case MMB_F32: return run_resize<float>((const float*)in, g, out, st);
    case MMB_F64: return run_resize<double>((const double*)in, g, out, st);
    default: set_error("unknown dtype %d", dtype); return MMB_ERR_INVALID;
  }
}

extern "C" int mmb_unmix_subtract(float* target, const float* other, int Z, int Y, int X,
                                  int64_t pitch, double factor, void* stream) {
  MMB_REQUIRE(target && other, "null buffer");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0 && pitch >= X, "bad shape");
  MMB_REQUIRE(Y <= 65535 && Z <= 65535, "Y and Z must be <= 65535");
  dim3 grid((unsigned)cdiv(X, 256), (unsigned)Y, (unsigned)Z);
  unmix_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(target, other, Y, X, pitch, (float)factor);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}
