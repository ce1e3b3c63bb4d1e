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

// A warp owns an 8-row x 128-column strip of one z-plane of scale `cur`; every
// thread owns four consecutive x and keeps its 10 rows (8 + one halo row above
// and below, 'nearest' = clamped row indices) in registers: ten 16-byte loads in
// flight per thread, no shared memory.  The in-plane 3x3 maximum is separable:
// a vertical max of three rows per column, then a horizontal max of three
// columns, with the two columns next to a thread's four taken from the
// neighbouring lanes by shuffle (lanes 0 and 31 only carry those columns for
// lanes 1 and 30, so a warp emits 120 of the 128 columns it loads).  Cold voxels cost about seven instructions each.  The few in-plane
// maxima above threshold (about one per blob per plane) are then tested against
// the two neighbouring planes and the 54 voxels of the adjacent scales with
// global loads.
constexpr int kLmThreads = 128;
constexpr int kLmRows = 8;                     // rows per warp
constexpr int kLmWarps = kLmThreads / 32;
constexpr int kLmCols = 120;                   // output columns per warp: 30 lanes x 4

__device__ __forceinline__ bool row_beats(const float* __restrict__ row, int xl, int x, int xr,
                                          float v) {
  return fmaxf(fmaxf(__ldcg(row + xl), __ldcg(row + x)), __ldcg(row + xr)) > v;
}

// planes z-1, z+1 of the same scale and all 27 voxels of each adjacent scale
__device__ __noinline__ bool survives_3d_and_scales(const float* __restrict__ prev,
                                       const float* __restrict__ cur,
                                       const float* __restrict__ next, int Z, int Y, int X,
                                       int64_t pitch, int z, int y, int x, float v) {
  const int64_t plane = (int64_t)Y * pitch;
  const int xl = x > 0 ? x - 1 : 0, xr = x < X - 1 ? x + 1 : X - 1;
  const int64_t ys[3] = {(int64_t)(y > 0 ? y - 1 : 0) * pitch, (int64_t)y * pitch,
                         (int64_t)(y < Y - 1 ? y + 1 : Y - 1) * pitch};
  const int64_t zl = (int64_t)(z > 0 ? z - 1 : 0) * plane;
  const int64_t zr = (int64_t)(z < Z - 1 ? z + 1 : Z - 1) * plane;
  const int64_t zs[3] = {zl, (int64_t)z * plane, zr};
  // Cheap pre-filter: the four neighbours straight above / below in z and in scale, loaded
  // together.  An in-plane maximum that is not a 4-D peak (about 99 of 100: a blob leaves one
  // in every plane and scale it spans) is almost always beaten by one of them, so most
  // candidates leave after ONE round trip instead of a chain of dependent row tests.
  {
    const int64_t at = ys[1] + x;
    const float a = __ldcg(cur + zl + at), b = __ldcg(cur + zr + at);
    const float p = prev ? __ldcg(prev + zs[1] + at) : -INFINITY;
    const float n = next ? __ldcg(next + zs[1] + at) : -INFINITY;
    if (fmaxf(fmaxf(a, b), fmaxf(p, n)) > v) return false;
  }
  for (int b = 0; b < 3; ++b)
    if (row_beats(cur + zl + ys[b], xl, x, xr, v) || row_beats(cur + zr + ys[b], xl, x, xr, v))
      return false;
  const float* others[2] = {prev, next};
  for (int c = 0; c < 2; ++c) {
    const float* vol = others[c];
    if (vol == nullptr) continue;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        if (row_beats(vol + zs[a] + ys[b], xl, x, xr, v)) return false;
  }
  return true;
}

__global__ void __launch_bounds__(kLmThreads, 8)
localmax_kernel(const float* __restrict__ prev, const float* __restrict__ cur,
                const float* __restrict__ next, int Z, int Y, int X, int64_t pitch, int s,
                float thr, int z_lo, int z_hi, mmb_cand* __restrict__ out, int capacity,
                int* __restrict__ counter) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int z = z_lo + blockIdx.z;
  const int y0 = (blockIdx.y * kLmWarps + warp) * kLmRows;
  if (y0 >= Y) return;                                  // warp-uniform
  // lanes 1..30 own outputs; lanes 0 and 31 only carry the neighbouring columns
  const int x0 = blockIdx.x * kLmCols - 4 + lane * 4;
  const float* plane = cur + (int64_t)z * Y * pitch;
  const float NEG = -INFINITY;

  // rows y0-1 .. y0+8 (clamped row indices = 'nearest'); everything outside the
  // volume is -inf so it never wins a maximum
  float4 v[kLmRows + 2];
#pragma unroll
  for (int rr = 0; rr < kLmRows + 2; ++rr) {
    const int y = min(max(y0 - 1 + rr, 0), Y - 1);
    v[rr] = make_float4(NEG, NEG, NEG, NEG);
    if (x0 >= 0 && x0 < X)
      v[rr] = __ldcg(reinterpret_cast<const float4*>(plane + (int64_t)y * pitch + x0));
  }
  // columns past X inside a loaded float4 hold row padding: mask them
  if (x0 + 3 >= X) {
#pragma unroll
    for (int rr = 0; rr < kLmRows + 2; ++rr) {
      if (x0 + 1 >= X) v[rr].y = NEG;
      if (x0 + 2 >= X) v[rr].z = NEG;
      if (x0 + 3 >= X) v[rr].w = NEG;
    }
  }
  const bool owner = lane >= 1 && lane <= 30;

  // phase 1 (unrolled, registers only): one bit per (row, column) that is above the
  // threshold and not beaten inside its 3x3 in-plane neighbourhood
  unsigned mask = 0;
#pragma unroll
  for (int r = 1; r <= kLmRows; ++r) {
    // vertical maxima of rows r-1, r, r+1
    const float c0 = fmaxf(fmaxf(v[r - 1].x, v[r].x), v[r + 1].x);
    const float c1 = fmaxf(fmaxf(v[r - 1].y, v[r].y), v[r + 1].y);
    const float c2 = fmaxf(fmaxf(v[r - 1].z, v[r].z), v[r + 1].z);
    const float c3 = fmaxf(fmaxf(v[r - 1].w, v[r].w), v[r + 1].w);
    const float cl = __shfl_up_sync(0xffffffffu, c3, 1);
    const float cr = __shfl_down_sync(0xffffffffu, c0, 1);
    // 3x3 maxima (the centre is included, so "is a maximum" is !(m > v))
    const float m0 = fmaxf(fmaxf(cl, c0), c1);
    const float m1 = fmaxf(fmaxf(c0, c1), c2);
    const float m2 = fmaxf(fmaxf(c1, c2), c3);
    const float m3 = fmaxf(fmaxf(c2, c3), cr);
    unsigned bits = 0;
    if (v[r].x > thr && !(m0 > v[r].x)) bits |= 1u;
    if (v[r].y > thr && !(m1 > v[r].y)) bits |= 2u;
    if (v[r].z > thr && !(m2 > v[r].z)) bits |= 4u;
    if (v[r].w > thr && !(m3 > v[r].w)) bits |= 8u;
    if (y0 + r - 1 < Y) mask |= bits << (4 * (r - 1));
  }
  if (!owner) mask = 0;

  // phase 2 (a loop, not unrolled: the straight-line form of this kernel was 58 KB of
  // code and stalled on instruction fetch): the few in-plane maxima are re-read from
  // memory (an L1 hit) and tested against the planes above and below and the two
  // adjacent scales; survivors are compacted with a warp ballot
  while (__any_sync(0xffffffffu, mask != 0)) {
    bool peak = false;
    int y = 0, x = 0;
    float val = 0.f;
    if (mask) {
      const int bit = __ffs(mask) - 1;
      mask &= mask - 1;
      y = y0 + (bit >> 2);
      x = x0 + (bit & 3);
      val = __ldcg(plane + (int64_t)y * pitch + x);
      peak = survives_3d_and_scales(prev, cur, next, Z, Y, X, pitch, z, y, x, val);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, peak);
    if (ballot) {
      int base = 0;
      if (lane == 0) base = atomicAdd(counter, __popc(ballot));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (peak) {
        const int idx = base + __popc(ballot & ((1u << lane) - 1u));
        if (idx < capacity) {
          mmb_cand c;
          c.z = z; c.y = y; c.x = x; c.s = s; c.resp = val;
          out[idx] = c;
        }
      }
    }
  }
}

int localmax_impl(const float* prev, const float* cur, const float* next, int Z, int Y, int X,
                  int64_t pitch, int s, float thr, int z_lo, int z_hi, mmb_cand* out,
                  int capacity, int* counter, cudaStream_t st) {
  if (z_hi <= z_lo) return MMB_OK;
  if (pitch % 4 != 0 || (reinterpret_cast<uintptr_t>(cur) & 15) != 0) {
    set_error("localmax needs 16-byte aligned rows (pitch %% 4 == 0)");
    return MMB_ERR_INVALID;
  }
  const int nz = z_hi - z_lo;
  for (int zb = 0; zb < nz; zb += 65535) {
    const int zn = nz - zb < 65535 ? nz - zb : 65535;
    dim3 grid((unsigned)cdiv(X, kLmCols), (unsigned)cdiv(Y, kLmRows * kLmWarps), (unsigned)zn);
    ProfScope ps(PROF_LOCALMAX, (double)zn * Y * X, st);
    localmax_kernel<<<grid, kLmThreads, 0, st>>>(prev, cur, next, Z, Y, X, pitch, s, thr,
                                                 z_lo + zb, z_hi, out, capacity, counter);
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
  MMB_REQUIRE(Y <= 65535 * 32, "Y too large");
  MMB_REQUIRE(z_lo >= 0 && z_hi <= Z, "bad z range");
  MMB_REQUIRE(capacity >= 0, "bad capacity");
  return mmb::localmax_impl(prev, cur, next, Z, Y, X, pitch, s, thr, z_lo, z_hi, out, capacity,
                            counter, (cudaStream_t)stream);
}
