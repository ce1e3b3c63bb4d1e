// img_as_float: strided u8/u16/f32/f64 -> pitched float32, times a scale.
#include "common.cuh"

namespace mmb {

template <typename T>
__global__ void to_float_kernel(const T* __restrict__ in, int64_t sz, int64_t sy, int64_t sx,
                                int Y, int X, float* __restrict__ out, int64_t pitch,
                                float scale, double dscale) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= X) return;
  const int y = blockIdx.y;
  const int z = blockIdx.z;
  const T v = in[(int64_t)z * sz + (int64_t)y * sy + (int64_t)x * sx];
  float f;
  if (sizeof(T) == 8) f = (float)((double)v * dscale);
  else f = (float)v * scale;
  out[((int64_t)z * Y + y) * pitch + x] = f;
}

int to_float_impl(const void* in, int dtype, const int64_t st[3], int Z, int Y, int X,
                  float* out, int64_t pitch, double scale, cudaStream_t s) {
  dim3 grid((unsigned)cdiv(X, 256), (unsigned)Y, (unsigned)Z);
  const float fs = (float)scale;
  ProfScope ps(PROF_TO_FLOAT, (double)Z * Y * X, s);
  switch (dtype) {
    case MMB_U8:
      to_float_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t*)in, st[0], st[1], st[2], Y, X, out, pitch, fs, scale);
      break;
    case MMB_U16:
      to_float_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)in, st[0], st[1], st[2], Y, X, out, pitch, fs, scale);
      break;
    case MMB_F32:
      to_float_kernel<float><<<grid, 256, 0, s>>>((const float*)in, st[0], st[1], st[2], Y, X, out, pitch, fs, scale);
      break;
    case MMB_F64:
      to_float_kernel<double><<<grid, 256, 0, s>>>((const double*)in, st[0], st[1], st[2], Y, X, out, pitch, fs, scale);
      break;
    default:
      set_error("unknown dtype %d", dtype);
      return MMB_ERR_INVALID;
  }
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

}  // namespace mmb

extern "C" int mmb_to_float(const void* in, int dtype, const int64_t in_strides[3], int Z, int Y,
                            int X, float* out, int64_t pitch, double scale, void* stream) {
  MMB_REQUIRE(in && out && in_strides, "null buffer");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0 && pitch >= X, "bad shape");
  MMB_REQUIRE(Y <= 65535 && Z <= 65535, "Y and Z must be <= 65535");
  return mmb::to_float_impl(in, dtype, in_strides, Z, Y, X, out, pitch, scale,
                            (cudaStream_t)stream);
}

extern "C" int mmb_upload_pieces(void* dst_device, const void* src_host, int64_t planes,
                                 int64_t piece_bytes, int64_t src_pitch_bytes, void* stream) {
  MMB_REQUIRE(dst_device && src_host, "null buffer");
  MMB_REQUIRE(planes > 0 && piece_bytes > 0 && src_pitch_bytes >= piece_bytes, "bad geometry");
  MMB_CHECK_CUDA(cudaMemcpy2DAsync(dst_device, (size_t)piece_bytes, src_host,
                                   (size_t)src_pitch_bytes, (size_t)piece_bytes, (size_t)planes,
                                   cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return MMB_OK;
}
