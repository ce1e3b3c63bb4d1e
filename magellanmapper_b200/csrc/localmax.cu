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

// A CTA owns an 8-row x 512-column tile of one z-plane of scale `cur`, staged in
// shared memory with a one-voxel halo ('nearest' = clamped indices at the faces).
// Every thread owns four consecutive x (16-byte loads).  Cold voxels cost a
// quarter of a load and a quarter of a shared store each.  An above-threshold
// voxel is first tested against its eight in-plane neighbours from shared memory;
// the few 2-D maxima that survive (about one per blob per plane) are then tested
// against the two neighbouring planes and the 54 voxels of the adjacent scales
// with global loads.
constexpr int kLmThreads = 128;
constexpr int kLmRows = 8;
constexpr int kLmCols = kLmThreads * 4;
constexpr int kLmPitch = kLmCols + 8;        // data starts at column 4 (keeps float4 alignment)

__device__ __forceinline__ bool row_beats(const float* __restrict__ row, int xl, int x, int xr,
                                          float v) {
  return fmaxf(fmaxf(__ldg(row + xl), __ldg(row + x)), __ldg(row + xr)) > v;
}

// planes z-1, z+1 of the same scale and all 27 voxels of each adjacent scale
__device__ bool survives_3d_and_scales(const float* __restrict__ prev,
                                       const float* __restrict__ cur,
                                       const float* __restrict__ next, int Z, int Y, int X,
                                       int64_t pitch, int z, int y, int x, float v) {
  const int64_t plane = (int64_t)Y * pitch;
  const int xl = x > 0 ? x - 1 : 0, xr = x < X - 1 ? x + 1 : X - 1;
  const int64_t ys[3] = {(int64_t)(y > 0 ? y - 1 : 0) * pitch, (int64_t)y * pitch,
                         (int64_t)(y < Y - 1 ? y + 1 : Y - 1) * pitch};
  const int64_t zl = (int64_t)(z > 0 ? z - 1 : 0) * plane;
  const int64_t zr = (int64_t)(z < Z - 1 ? z + 1 : Z - 1) * plane;
  for (int b = 0; b < 3; ++b)
    if (row_beats(cur + zl + ys[b], xl, x, xr, v) || row_beats(cur + zr + ys[b], xl, x, xr, v))
      return false;
  const int64_t zs[3] = {zl, (int64_t)z * plane, zr};
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

__global__ void __launch_bounds__(kLmThreads)
localmax_kernel(const float* __restrict__ prev, const float* __restrict__ cur,
                const float* __restrict__ next, int Z, int Y, int X, int64_t pitch, int s,
                float thr, int z_lo, int z_hi, mmb_cand* __restrict__ out, int capacity,
                int* __restrict__ counter) {
  __shared__ __align__(16) float tile[(kLmRows + 2) * kLmPitch];
  const int tid = threadIdx.x, lane = tid & 31;
  const int z = z_lo + blockIdx.z;
  const int y0 = blockIdx.y * kLmRows;
  const int xt = blockIdx.x * kLmCols;          // first column of the tile
  const int x0 = xt + tid * 4;
  const float* plane = cur + (int64_t)z * Y * pitch;
#pragma unroll
  for (int rr = 0; rr < kLmRows + 2; ++rr) {
    const int y = min(max(y0 - 1 + rr, 0), Y - 1);
    const float* row = plane + (int64_t)y * pitch;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x0 < X) q = __ldg(reinterpret_cast<const float4*>(row + x0));
    *reinterpret_cast<float4*>(&tile[rr * kLmPitch + 4 + tid * 4]) = q;
    if (tid == 0) tile[rr * kLmPitch + 3] = xt > 0 ? __ldg(row + xt - 1) : 0.f;
    if (tid == 1) tile[rr * kLmPitch + 4 + kLmCols] = xt + kLmCols < X ? __ldg(row + xt + kLmCols) : 0.f;
  }
  __syncthreads();
#pragma unroll 1
  for (int r = 1; r <= kLmRows; ++r) {
    const int y = y0 + r - 1;
    const bool row_ok = y < Y;                         // uniform across the CTA
    const float* t = &tile[r * kLmPitch + 4 + tid * 4];
    const float4 q = *reinterpret_cast<const float4*>(t);
    const float vs[4] = {q.x, q.y, q.z, q.w};
    unsigned hot = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (row_ok && x0 + k < X && vs[k] > thr) hot |= 1u << k;
    if (!__any_sync(0xffffffffu, hot != 0)) continue;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      bool peak = false;
      if (hot & (1u << k)) {
        const int x = x0 + k;
        const float v = vs[k];
        // clamped in-tile neighbour columns ('nearest' at the x faces)
        const int cl = x > 0 ? k - 1 : k, cr = x < X - 1 ? k + 1 : k;
        // rows r-1 / r+1 already hold clamped y (loaded with clamped indices)
        const float* up = t - kLmPitch;
        const float* dn = t + kLmPitch;
        float m = fmaxf(t[cl], t[cr]);
        m = fmaxf(m, fmaxf(fmaxf(up[cl], up[k]), up[cr]));
        m = fmaxf(m, fmaxf(fmaxf(dn[cl], dn[k]), dn[cr]));
        peak = !(m > v);
        if (peak) peak = survives_3d_and_scales(prev, cur, next, Z, Y, X, pitch, z, y, x, v);
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
            c.z = z; c.y = y; c.x = x0 + k; c.s = s; c.resp = vs[k];
            out[idx] = c;
          }
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
    dim3 grid((unsigned)cdiv(X, kLmCols), (unsigned)cdiv(Y, kLmRows), (unsigned)zn);
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
  MMB_REQUIRE(Y <= 65535 * 8, "Y too large");
  MMB_REQUIRE(z_lo >= 0 && z_hi <= Z, "bad z range");
  MMB_REQUIRE(capacity >= 0, "bad capacity");
  return mmb::localmax_impl(prev, cur, next, Z, Y, X, pitch, s, thr, z_lo, z_hi, out, capacity,
                            counter, (cudaStream_t)stream);
}
