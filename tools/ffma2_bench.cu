// Micro-benchmark of the FP32 FMA pipe of sm_100a: scalar FFMA vs packed FFMA2 in the
// operand forms the sweeps use.  Developer tool:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench tools/ffma2_bench.cu
// Prints lane-FMAs per cycle per SM (peak of the scalar pipe = 128).
#include <cstdio>
#include <cuda_runtime.h>

struct W { float g[32]; float2 gh[32]; };

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, int iters, const __grid_constant__ W w,
                                         long long* cycles) {
  constexpr int NA = 16;
  float2 acc[NA];
  float sacc[2 * NA];
  float2 v = make_float2(threadIdx.x * 1e-3f, threadIdx.x * 2e-3f);
#pragma unroll
  for (int j = 0; j < NA; ++j) { acc[j] = make_float2(j, -j); sacc[2 * j] = j; sacc[2 * j + 1] = -j; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 16; ++t) {
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        if (MODE == 0) {          // scalar FFMA, constant-bank weight
          sacc[2 * j] = fmaf(w.g[t], v.x, sacc[2 * j]);
          sacc[2 * j + 1] = fmaf(w.g[t + 16], v.y, sacc[2 * j + 1]);
        } else if (MODE == 1) {   // FFMA2, uniform scalar broadcast weight
          acc[j] = __ffma2_rn(v, make_float2(w.g[t], w.g[t]), acc[j]);
        } else if (MODE == 2) {   // FFMA2, scalar broadcast input, uniform pair weight
          acc[j] = __ffma2_rn(make_float2(v.x, v.x), w.gh[t], acc[j]);
        } else if (MODE == 3) {   // FFMA2, all register operands
          acc[j] = __ffma2_rn(v, acc[(j + 1) % NA], acc[j]);
        } else if (MODE == 4) {   // 2 FFMA2 : 1 scalar FFMA mix
          acc[j] = __ffma2_rn(v, make_float2(w.g[t], w.g[t]), acc[j]);
          if (j % 2 == 0) sacc[j] = fmaf(w.g[t + 16], v.x, sacc[j]);
        }
      }
      v.x += 1e-7f;
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NA; ++j) s += acc[j].x + acc[j].y + sacc[2 * j] + sacc[2 * j + 1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, double fma_per_iter_thread) {
  int nsm = 148;
  float* out; long long* cyc;
  cudaMalloc(&out, nsm * 1024 * sizeof(float));
  cudaMalloc(&cyc, nsm * sizeof(long long));
  W w;
  for (int i = 0; i < 32; ++i) { w.g[i] = 1e-3f * i; w.gh[i] = make_float2(1e-3f * i, -1e-3f * i); }
  const int iters = 2000;
  k<MODE><<<nsm, threads>>>(out, 10, w, cyc);
  k<MODE><<<nsm, threads>>>(out, iters, w, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < nsm; ++i) c += h[i]; c /= nsm;
  printf("%-52s threads/SM %4d  lane-FMA/cycle/SM %7.1f\n", name, threads,
         fma_per_iter_thread * iters * threads / c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 512, 1024}) {
    run<0>("scalar FFMA, constant weight", threads, 16 * 16 * 2);
    run<1>("FFMA2  Rpair * UR.F32(bcast) + Rpair", threads, 16 * 16 * 2);
    run<2>("FFMA2  R.F32(bcast) * URpair + Rpair", threads, 16 * 16 * 2);
    run<3>("FFMA2  Rpair * Rpair + Rpair", threads, 16 * 16 * 2);
    run<4>("mix 2 FFMA2 : 1 FFMA", threads, 16 * 16 * 2 + 16 * 8);
  }
  return 0;
}
