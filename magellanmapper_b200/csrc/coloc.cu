// Intensity co-localisation of blobs across channels: the voxel work of
// colocalizer.colocalize_blobs (magmap/cv/colocalizer.py:340-441).
//
// The reference labels a mask of the ROI with the index of every blob of one channel at
// the blob's voxel (a later blob overwrites an earlier one at the same voxel), grey-dilates
// the mask with skimage.morphology.ball(2) - so a voxel reached by several blobs takes the
// LARGEST index - and then, blob by blob, averages roi[mask == b, channel] for every channel.
// Here: (1) every blob of the channel writes its index into the 33 voxels of its ball with
// atomicMax (the same maximum; voxels outside the ROI are skipped, which equals the
// dilation's 'reflect' border because a reflected ball offset is again a ball offset);
// (2) every blob revisits its ball and adds the voxels it still owns, for every channel, to
// an exact integer sum (uint8 / uint16) or a float64 sum.  The host divides, takes the
// thresholds and compares, exactly as the reference's numpy does.
#include "common.cuh"

namespace mmb {

struct BallOffsets { signed char d[33][3]; };

static BallOffsets make_ball2() {          // skimage.morphology.ball(2): dz^2 + dy^2 + dx^2 <= 4
  BallOffsets b;
  int n = 0;
  for (int z = -2; z <= 2; ++z)
    for (int y = -2; y <= 2; ++y)
      for (int x = -2; x <= 2; ++x)
        if (z * z + y * y + x * x <= 4) { b.d[n][0] = z; b.d[n][1] = y; b.d[n][2] = x; ++n; }
  return b;                                 // n == 33
}

// blobs: (n, 4) int32 rows z, y, x, channel.  One thread per (blob, ball voxel).
__global__ void coloc_label_kernel(const int* __restrict__ blobs, int n, int chl, int Z, int Y,
                                   int X, const __grid_constant__ BallOffsets ball,
                                   int* __restrict__ label) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = i / 33, k = i - b * 33;
  if (b >= n || blobs[4 * b + 3] != chl) return;
  const int z = blobs[4 * b] + ball.d[k][0], y = blobs[4 * b + 1] + ball.d[k][1],
            x = blobs[4 * b + 2] + ball.d[k][2];
  if (z < 0 || z >= Z || y < 0 || y >= Y || x < 0 || x >= X) return;
  atomicMax(label + ((int64_t)z * Y + y) * X + x, b);
}

template <typename T>
__global__ void coloc_sum_kernel(const T* __restrict__ roi, int64_t sz, int64_t sy, int64_t sx,
                                 int64_t sc, const int* __restrict__ blobs, int n, int chl, int Z,
                                 int Y, int X, int C, const __grid_constant__ BallOffsets ball,
                                 const int* __restrict__ label, double* __restrict__ sums,
                                 unsigned long long* __restrict__ isums,
                                 int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = i / 33, k = i - b * 33;
  if (b >= n || blobs[4 * b + 3] != chl) return;
  const int z = blobs[4 * b] + ball.d[k][0], y = blobs[4 * b + 1] + ball.d[k][1],
            x = blobs[4 * b + 2] + ball.d[k][2];
  if (z < 0 || z >= Z || y < 0 || y >= Y || x < 0 || x >= X) return;
  if (__ldcg(label + ((int64_t)z * Y + y) * X + x) != b) return;
  // a voxel is reached once per blob: two blobs at one voxel share every offset, but only
  // the larger index owns the voxel
  atomicAdd(counts + b, 1);
  const T* p = roi + z * sz + y * sy + x * sx;
  for (int c = 0; c < C; ++c) {
    if constexpr (sizeof(T) <= 2)
      atomicAdd(isums + (int64_t)b * C + c, (unsigned long long)p[c * sc]);
    else
      atomicAdd(sums + (int64_t)b * C + c, (double)p[c * sc]);
  }
}

// exact integer sums (below 2^53) to float64, in place
__global__ void coloc_to_f64_kernel(double* __restrict__ sums, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long v = reinterpret_cast<unsigned long long*>(sums)[i];
  sums[i] = (double)v;
}

}  // namespace mmb

using namespace mmb;

extern "C" int64_t mmb_coloc_work_bytes(int Z, int Y, int X) {
  return (int64_t)Z * Y * X * (int64_t)sizeof(int);
}

extern "C" int mmb_coloc_sums(const void* roi, int dtype, const int64_t strides[4], int Z, int Y,
                              int X, int C, const int32_t* blobs, int n, double* sums,
                              int32_t* counts, void* work, void* stream) {
  MMB_REQUIRE(roi && strides && blobs && sums && counts && work, "null buffer");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0 && C > 0 && n >= 0, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  MMB_CHECK_CUDA(cudaMemsetAsync(sums, 0, (size_t)n * C * sizeof(double), st));
  MMB_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)n * sizeof(int), st));
  if (n == 0) return MMB_OK;
  static const BallOffsets ball = make_ball2();
  int* label = (int*)work;
  const bool integer = dtype == MMB_U8 || dtype == MMB_U16;
  const unsigned grid = (unsigned)cdiv((int64_t)n * 33, 256);
  // integer sums are exact in 64 bits and are converted in place at the end
  unsigned long long* isums = (unsigned long long*)sums;
  for (int chl = 0; chl < C; ++chl) {
    MMB_CHECK_CUDA(cudaMemsetAsync(label, 0xFF, (size_t)mmb_coloc_work_bytes(Z, Y, X), st));
    coloc_label_kernel<<<grid, 256, 0, st>>>(blobs, n, chl, Z, Y, X, ball, label);
    MMB_CHECK_LAUNCH();
#define COLOC_(T)                                                                              \
  coloc_sum_kernel<T><<<grid, 256, 0, st>>>((const T*)roi, strides[0], strides[1], strides[2], \
                                            strides[3], blobs, n, chl, Z, Y, X, C, ball, label, \
                                            sums, isums, counts)
    switch (dtype) {
      case MMB_U8: COLOC_(uint8_t); break;
      case MMB_U16: COLOC_(uint16_t); break;
      case MMB_F32: COLOC_(float); break;
      case MMB_F64: COLOC_(double); break;
      default: set_error("unknown dtype %d", dtype); return MMB_ERR_INVALID;
    }
#undef COLOC_
    MMB_CHECK_LAUNCH();
  }
  if (integer) {
    coloc_to_f64_kernel<<<(unsigned)cdiv((int64_t)n * C, 256), 256, 0, st>>>(sums, (int64_t)n * C);
    MMB_CHECK_LAUNCH();
  }
  return MMB_OK;
}
