// 4-D (z, y, x, scale) local maxima above threshold, warp-ballot compacted.
//
// peak_local_max(cube, threshold_abs=thr, footprint=ones((3,3,3,3)),
// exclude_border=False): voxel is a peak iff it equals the maximum of its 3^4
// neighbourhood ('nearest' padding = clamp at every face and at the ends of the
// scale axis) and is strictly above thr.  A neighbour can only beat a voxel that
// is above thr by being above thr itself, so the 80-neighbour test runs only for
// the sparse above-threshold voxels.
#include "common.cuh"

namespace mmb {

__global__ void __launch_bounds__(256)
localmax_kernel(const float* __restrict__ prev, const float* __restrict__ cur,
                const float* __restrict__ next, int Z, int Y, int X, int64_t pitch, int s,
                float thr, int z_lo, int z_hi, mmb_cand* __restrict__ out, int capacity,
                int* __restrict__ counter) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int z = z_lo + blockIdx.z;
  bool peak = false;
  float v = 0.f;
  if (x < X && z < z_hi) {
    const int64_t plane = (int64_t)Y * pitch;
    v = cur[(int64_t)z * plane + (int64_t)y * pitch + x];
    if (v > thr) {
      peak = true;
      const int zs[3] = {clamp_index(z - 1, Z), z, clamp_index(z + 1, Z)};
      const int ys[3] = {clamp_index(y - 1, Y), y, clamp_index(y + 1, Y)};
      const int xs[3] = {clamp_index(x - 1, X), x, clamp_index(x + 1, X)};
      const float* vols[3] = {prev, cur, next};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* vol = vols[c];
        if (vol == nullptr) continue;
        for (int a = 0; a < 3 && peak; ++a)
          for (int b = 0; b < 3; ++b) {
            const float* row = vol + (int64_t)zs[a] * plane + (int64_t)ys[b] * pitch;
            const float m = fmaxf(fmaxf(row[xs[0]], row[xs[1]]), row[xs[2]]);
            if (m > v) peak = false;
          }
      }
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, peak);
  if (ballot) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(counter, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (peak) {
      const int idx = base + __popc(ballot & ((1u << lane) - 1u));
      if (idx < capacity) {
        mmb_cand c;
        c.z = z; c.y = y; c.x = x; c.s = s; c.resp = v;
        out[idx] = c;
      }
    }
  }
}

int localmax_impl(const float* prev, const float* cur, const float* next, int Z, int Y, int X,
                  int64_t pitch, int s, float thr, int z_lo, int z_hi, mmb_cand* out,
                  int capacity, int* counter, cudaStream_t st) {
  if (z_hi <= z_lo) return MMB_OK;
  const int nz = z_hi - z_lo;
  for (int z0 = 0; z0 < nz; z0 += 65535) {
    const int zn = nz - z0 < 65535 ? nz - z0 : 65535;
    dim3 grid((unsigned)cdiv(X, 256), (unsigned)Y, (unsigned)zn);
    ProfScope ps(PROF_LOCALMAX, (double)zn * Y * X, st);
    localmax_kernel<<<grid, 256, 0, st>>>(prev, cur, next, Z, Y, X, pitch, s, thr, z_lo + z0,
                                          z_hi, out, capacity, counter);
    MMB_CHECK_LAUNCH();
  }
  return MMB_OK;
}

}  // namespace mmb

extern "C" int mmb_localmax_compact(const float* prev, const float* cur, const float* next,
                                    int Z, int Y, int X, int64_t pitch, int s, float thr,
                                    int z_lo, int z_hi, mmb_cand* out, int capacity,
                                    int* counter, void* stream) {
  MMB_REQUIRE(cur && out && counter, "null buffer");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0 && pitch >= X, "bad shape");
  MMB_REQUIRE(Y <= 65535, "Y must be <= 65535");
  MMB_REQUIRE(z_lo >= 0 && z_hi <= Z, "bad z range");
  MMB_REQUIRE(capacity >= 0, "bad capacity");
  return mmb::localmax_impl(prev, cur, next, Z, Y, X, pitch, s, thr, z_lo, z_hi, out, capacity,
                            counter, (cudaStream_t)stream);
}
