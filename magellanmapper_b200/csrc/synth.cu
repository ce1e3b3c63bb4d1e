// Counter-based synthetic nuclei volumes, generated on the device.
//
// Bench / test utility, not part of the reference-facing boundary (declared in
// include/mmb200_tools.h): the whole-brain configuration (2048 x 8192 x 8192 uint16,
// 275 GB) can only exist as per-GPU slabs, so every voxel is a pure function of
// (seed, global z, y, x) and any sub-box of the volume - a slab, a halo, the box an
// oracle spot check recomputes - can be produced on its own, bit for bit.
//
// Recipe (the one magellanmapper_b200/synth.py draws with numpy's generator): background
// ~ N(400, 30^2); nuclei = isotropic Gaussian spots, sigma0 ~ U(2.5, 4.5), peak
// amplitude ~ U(0.3, 0.9) * 65535, Poisson(16^3 * density) centres per 16^3 cell of a
// grid anchored at the volume origin, each spot cut off 16 voxels from its centre;
// clipped to [0, 65535] and rounded half to even.
#include "common.cuh"

namespace mmb {

constexpr int kCell = 16;
constexpr int kMaxPerCell = 6;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {   // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ float u01(uint64_t h) {                 // (0, 1]
  return ((float)(uint32_t)(h >> 40) + 1.0f) * (1.0f / 16777216.0f);
}

struct Nucleus { float z, y, x, inv2s2, amp; };

// nuclei of cell (cz, cy, cx): count by inverse Poisson CDF, then five uniforms each
__device__ int cell_nuclei(uint64_t seed, int64_t cz, int64_t cy, int64_t cx, float lambda,
                           Nucleus* out) {
  const uint64_t cid = mix64(seed ^ mix64((uint64_t)cz * 0x100000001B3ull ^
                                          mix64((uint64_t)cy * 0x9E3779B1ull ^ mix64((uint64_t)cx))));
  const float u = u01(mix64(cid));
  float p = __expf(-lambda), cdf = p;
  int n = 0;
  while (u > cdf && n < kMaxPerCell) { ++n; p *= lambda / (float)n; cdf += p; }
  for (int k = 0; k < n; ++k) {
    const uint64_t h = mix64(cid + 0x632BE59BD9B4E019ull * (uint64_t)(k + 1));
    const float s0 = 2.5f + 2.0f * u01(mix64(h ^ 4));
    out[k].z = (float)(cz * kCell) + kCell * u01(mix64(h ^ 1));
    out[k].y = (float)(cy * kCell) + kCell * u01(mix64(h ^ 2));
    out[k].x = (float)(cx * kCell) + kCell * u01(mix64(h ^ 3));
    out[k].inv2s2 = 0.5f / (s0 * s0);
    out[k].amp = (0.3f + 0.6f * u01(mix64(h ^ 5))) * 65535.0f;
  }
  return n;
}

// one CTA per 16^3 cell of the OUTPUT box's cell cover; 256 threads = (y, x), loop z
__global__ void __launch_bounds__(256)
synth_kernel(uint16_t* __restrict__ out, int Z, int Y, int X, int64_t z_off, int64_t y_off,
             int64_t x_off, uint64_t seed, float lambda, int ncy, int ncx, int64_t c0z,
             int64_t c0y, int64_t c0x) {
  __shared__ Nucleus s_n[27 * kMaxPerCell];
  __shared__ int s_c[27];
  int b = blockIdx.x;
  const int ix = b % ncx; b /= ncx;
  const int iy = b % ncy; b /= ncy;
  const int64_t cz = c0z + b, cy = c0y + iy, cx = c0x + ix;
  // the 27 neighbouring cells list their nuclei in a FIXED order (cell index, then
  // nucleus index), so the float sum below is bit-reproducible from any sub-box
  Nucleus loc[kMaxPerCell];
  int n_loc = 0;
  if (threadIdx.x < 27) {
    const int dz = threadIdx.x / 9 - 1, dy = (threadIdx.x / 3) % 3 - 1, dx = threadIdx.x % 3 - 1;
    n_loc = cell_nuclei(seed, cz + dz, cy + dy, cx + dx, lambda, loc);
    s_c[threadIdx.x] = n_loc;
  }
  __syncthreads();
  if (threadIdx.x < 27) {
    int at = 0;
    for (int u = 0; u < (int)threadIdx.x; ++u) at += s_c[u];
    for (int k = 0; k < n_loc; ++k) s_n[at + k] = loc[k];
  }
  __syncthreads();
  int n = 0;
  for (int u = 0; u < 27; ++u) n += s_c[u];
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int64_t gy = cy * kCell + ty, gx = cx * kCell + tx;
  const int64_t ly = gy - y_off, lx = gx - x_off;
  if (ly < 0 || ly >= Y || lx < 0 || lx >= X) return;
  for (int tz = 0; tz < kCell; ++tz) {
    const int64_t gz = cz * kCell + tz;
    const int64_t lz = gz - z_off;
    if (lz < 0 || lz >= Z) continue;
    // background: Box-Muller on two hashes of the global voxel index
    const uint64_t vid = mix64(seed * 0xD6E8FEB86659FD93ull ^
                               (((uint64_t)gz << 42) ^ ((uint64_t)gy << 21) ^ (uint64_t)gx));
    const float u1 = u01(vid), u2 = u01(mix64(vid));
    float v = 400.0f + 30.0f * sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853f * u2);
    float acc = 0.f;
    for (int k = 0; k < n; ++k) {
      const float dz = (float)gz - s_n[k].z, dy = (float)gy - s_n[k].y, dx = (float)gx - s_n[k].x;
      const float d2 = dz * dz + dy * dy + dx * dx;
      if (d2 <= 256.0f) acc += s_n[k].amp * __expf(-d2 * s_n[k].inv2s2);
    }
    v += acc;
    v = fminf(fmaxf(v, 0.f), 65535.f);
    out[((int64_t)lz * Y + ly) * X + lx] = (uint16_t)__float2int_rn(v);
  }
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_synth_nuclei(uint16_t* out, int Z, int Y, int X, int64_t z_off, int64_t y_off,
                                int64_t x_off, uint64_t seed, double density, void* stream) {
  MMB_REQUIRE(out, "null buffer");
  MMB_REQUIRE(Z > 0 && Y > 0 && X > 0, "bad shape");
  MMB_REQUIRE(z_off >= 0 && y_off >= 0 && x_off >= 0, "offsets must be non-negative");
  MMB_REQUIRE(density > 0 && density * kCell * kCell * kCell < 3.0, "density out of range");
  const int64_t c0z = z_off / kCell, c0y = y_off / kCell, c0x = x_off / kCell;
  const int64_t ncz = (z_off + Z - 1) / kCell - c0z + 1;
  const int64_t ncy = (y_off + Y - 1) / kCell - c0y + 1;
  const int64_t ncx = (x_off + X - 1) / kCell - c0x + 1;
  MMB_REQUIRE(ncz * ncy * ncx < ((int64_t)1 << 31), "box too large for one launch");
  synth_kernel<<<(unsigned)(ncz * ncy * ncx), 256, 0, (cudaStream_t)stream>>>(
      out, Z, Y, X, z_off, y_off, x_off, seed, (float)(density * kCell * kCell * kCell),
      (int)ncy, (int)ncx, c0z, c0y, c0x);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}
