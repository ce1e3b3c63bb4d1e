// Micro-benchmark: the x phase's arithmetic (16 outputs x (A, B) per thread from a register
// window of 16 + 2 R values) in the scatter form the kernels use today and in the
// symmetric-pair form (FADD2 on aligned window pairs, then one FFMA2 per pair sum).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xform_bench tools/xform_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int R = 16, R4 = 16, W = 16 + 2 * R4;
struct Wt { float2 gh[R + 1]; };

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const float* in, float* out, int iters,
                                            const __grid_constant__ Wt w, long long* cycles) {
  float win[W];
#pragma unroll
  for (int i = 0; i < W; ++i) win[i] = in[(threadIdx.x * W + i) & 4095];
  float2 tot = make_float2(0.f, 0.f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float2 acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = make_float2(0.f, 0.f);
    if (MODE == 0) {
#pragma unroll
      for (int i = R4 - R; i < R4 + 16 + R; ++i) {
        const float2 vv = make_float2(win[i], win[i]);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int t = i - (R4 + j);
          if (t >= -R && t <= R) acc[j] = __ffma2_rn(vv, w.gh[t < 0 ? -t : t], acc[j]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        acc[j] = __ffma2_rn(make_float2(win[R4 + j], win[R4 + j]), w.gh[0], acc[j]);
#pragma unroll
      for (int t = 1; t <= R; ++t) {
        const int j0 = t & 1;                 // pairs start where the window index is even
        if (j0) {
          const float s0 = win[R4 - t] + win[R4 + t];
          acc[0] = __ffma2_rn(make_float2(s0, s0), w.gh[t], acc[0]);
          const float s15 = win[R4 + 15 - t] + win[R4 + 15 + t];
          acc[15] = __ffma2_rn(make_float2(s15, s15), w.gh[t], acc[15]);
        }
#pragma unroll
        for (int j = j0; j + 1 < 16; j += 2) {
          const float2 s = __fadd2_rn(make_float2(win[R4 + j - t], win[R4 + j - t + 1]),
                                      make_float2(win[R4 + j + t], win[R4 + j + t + 1]));
          acc[j] = __ffma2_rn(make_float2(s.x, s.x), w.gh[t], acc[j]);
          acc[j + 1] = __ffma2_rn(make_float2(s.y, s.y), w.gh[t], acc[j + 1]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) { tot.x += acc[j].x; tot.y += acc[j].y; }
#pragma unroll
    for (int i = 0; i < W; ++i) win[i] = fmaf(tot.x, 1e-9f, win[i]);   // every value changes: nothing hoists
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = tot.x + tot.y;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name) {
  const int nsm = 148 * 2, threads = 256, iters = 2000;
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
  cudaMalloc(&out, nsm * threads * 4); cudaMalloc(&cyc, nsm * 8);
  Wt w; for (int i = 0; i <= R; ++i) w.gh[i] = make_float2(1e-3f * i, -1e-3f * i);
  k<MODE><<<nsm, threads>>>(in, out, 10, w, cyc);
  k<MODE><<<nsm, threads>>>(in, out, iters, w, cyc);
  cudaDeviceSynchronize();
  long long h[296]; cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < nsm; ++i) c += h[i]; c /= nsm;
  printf("%-28s cycles per tile-row per thread %.1f  (%s)\n", name, c / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  run<0>("scatter form (today)");
  run<1>("symmetric pairs, FADD2");
  return 0;
}
