// Blob tables on the device: merge ordering, seam pruning and the final table of a
// chunked stack as hand-written kernels (no library sort / select on the product path).
//
// Reference: chunking.merge_blobs (magmap/cv/chunking.py:410-445) over
// Blobs.format_blobs tables (detector.py:325-364), StackPruner.prune_blobs_mp
// (stack_detect.py:680-861) with prune_overlap (:644-677) and remove_close_blobs
// (detector.py:1009-1085), then replace_rel_with_abs_blob_coords +
// remove_abs_blob_coords (stack_detect.py:458-467).
//
// A blob travels as a 32-byte mmb_row (chunk-local voxel, sigma index, response, chunk,
// channel).  Everything the reference keeps in (N, 14) float64 tables is index
// bookkeeping here:
//   * merge order = stable LSD radix sort of a row permutation by (chunk, channel,
//     descending response, C-order index in the chunk) - peak_local_max order inside a
//     detection, channels in request order inside a chunk, chunks in grid order;
//   * per axis, every row is classified by position into exactly one of "kept section
//     j", "master of seam j", "check of seam j" or "dropped" (the slabs and the kept
//     ranges tile the axis), one stable sort by that class puts the table into the
//     reference's order (kept sections, then per seam master rows and check rows),
//     check rows are bucketed into 64-voxel strips of an in-plane axis, each master
//     scans the three strips around it with the inclusive box test, matched masters
//     take the rounded mean of the absolute coordinates with their LAST match, and a
//     stable partition drops the matched checks;
//   * the final (N', 8) or (N', 11) float64 table is written once.
// Nothing synchronises with the host: counts live in device memory and grids are sized
// for the capacity.
#include <math.h>
#include <string.h>
#include "common.cuh"

namespace mmb {

constexpr int kRsThreads = 256;
constexpr int kRsRounds = 8;
constexpr int kRsItems = kRsThreads * kRsRounds;      // rows per block per pass
constexpr int kMaxSec = 128;                          // chunk sections per axis

// ---- stable LSD radix sort of a permutation, 8 bits per pass -------------------------
// digit of position i = (key[perm_in ? perm_in[i] : i] >> shift) & 255
__global__ void __launch_bounds__(kRsThreads)
rs_hist_kernel(const uint32_t* __restrict__ key, int shift, const int32_t* __restrict__ perm,
               const int* __restrict__ n_ptr, int n_max, int nblk, int* __restrict__ hist) {
  __shared__ int h[256];
  const int n = min(__ldcg(n_ptr), n_max);
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * kRsItems;
  for (int r = 0; r < kRsRounds; ++r) {
    const int i = base + r * kRsThreads + threadIdx.x;
    if (i < n) {
      const int row = perm ? __ldcg(perm + i) : i;
      atomicAdd(&h[(__ldcg(key + row) >> shift) & 255u], 1);
    }
  }
  __syncthreads();
  hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of hist[256][nblk] in digit-major order (single CTA);
// digit_start[d] = first output position of digit d, digit_start[256] = n
__global__ void __launch_bounds__(1024)
rs_scan_kernel(int* __restrict__ hist, int nblk, int* __restrict__ digit_start) {
  __shared__ int part[1024];
  const int total = 256 * nblk;
  const int per = (total + 1023) / 1024;
  const int lo = min(total, (int)threadIdx.x * per), hi = min(total, lo + per);
  int sum = 0;
  for (int k = lo; k < hi; ++k) sum += hist[k];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const int v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = part[threadIdx.x] - sum;
  for (int k = lo; k < hi; ++k) {
    const int c = hist[k];
    hist[k] = run;
    if (digit_start && k % nblk == 0) digit_start[k / nblk] = run;
    run += c;
  }
  if (digit_start && threadIdx.x == 1023) digit_start[256] = part[1023];
}

__global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const uint32_t* __restrict__ key, int shift, const int32_t* __restrict__ perm_in,
                  int32_t* __restrict__ perm_out, const int* __restrict__ n_ptr, int n_max,
                  int nblk, const int* __restrict__ hist) {
  __shared__ int running[256];
  const int n = min(__ldcg(n_ptr), n_max);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  running[threadIdx.x] = hist[threadIdx.x * nblk + blockIdx.x];
  __syncthreads();
  const int base = blockIdx.x * kRsItems;
  if (base >= n) return;
  for (int r = 0; r < kRsRounds; ++r) {
    const int i = base + r * kRsThreads + threadIdx.x;
    const bool on = i < n;
    int row = 0;
    unsigned d = 256u + (unsigned)threadIdx.x;          // unique: matches nobody
    if (on) {
      row = perm_in ? __ldcg(perm_in + i) : i;
      d = (__ldcg(key + row) >> shift) & 255u;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    const int leader = __ffs(peers) - 1;
    int at = 0;
    // warps claim their output ranges in warp order, which keeps the pass stable
    for (int w = 0; w < kRsThreads / 32; ++w) {
      if (warp == w && on && lane == leader) {
        at = running[d];
        running[d] = at + __popc(peers);
      }
      __syncthreads();
    }
    at = __shfl_sync(0xffffffffu, at, leader);
    if (on) perm_out[at + rank] = row;
  }
}

struct Sorter {
  int32_t* a;          // ping
  int32_t* b;          // pong
  int* hist;           // 256 * nblk
  int nblk;
  int n_max;
  cudaStream_t st;
};

// one pass; returns the buffer that holds the result.  perm_in may be null (identity).
static int radix_pass(const Sorter& S, const uint32_t* key, int shift, const int32_t* perm_in,
                      int32_t* perm_out, const int* n_ptr, int* digit_start) {
  rs_hist_kernel<<<S.nblk, kRsThreads, 0, S.st>>>(key, shift, perm_in, n_ptr, S.n_max, S.nblk,
                                                  S.hist);
  MMB_CHECK_LAUNCH();
  rs_scan_kernel<<<1, 1024, 0, S.st>>>(S.hist, S.nblk, digit_start);
  MMB_CHECK_LAUNCH();
  rs_scatter_kernel<<<S.nblk, kRsThreads, 0, S.st>>>(key, shift, perm_in, perm_out, n_ptr,
                                                     S.n_max, S.nblk, S.hist);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

// ---- geometry in kernel parameter space -------------------------------------------------
struct StackGeom {
  int grid[3];
  int tol[3];
  int n_channels;
  int num_sigma;
  int start[3][kMaxSec];      // first voxel of section j along axis a
  int size[3][kMaxSec];       // extent of section j along axis a
};

struct AxisPlan {
  int axis, n_sec;
  int upper[kMaxSec];         // kept range of section j is [lower[j], upper[j])
  int lower[kMaxSec];
  int slab_hi[kMaxSec];       // slab of seam j is [upper[j], slab_hi[j]), j < n_sec - 1
  int nlo[kMaxSec], nhi[kMaxSec];   // the ratio metric's region past slab j; nlo > nhi: none
  int u_axis;                 // in-plane axis the checks are bucketed along
  int n_strips;
  int strip_shift;            // strip of a position u = u >> strip_shift (>= 64 voxels wide)
};

__device__ __forceinline__ void chunk_coord(const StackGeom& g, int chunk, int c[3]) {
  c[2] = chunk % g.grid[2];
  chunk /= g.grid[2];
  c[1] = chunk % g.grid[1];
  c[0] = chunk / g.grid[1];
}

__device__ __forceinline__ uint32_t desc_key(float f) {      // ascending key = descending float
  const uint32_t u = __float_as_uint(f);
  const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~asc;
}

// sort keys of the merge order and the absolute coordinates of every row
__global__ void row_keys_kernel(const mmb_row* __restrict__ rows, int n,
                                const __grid_constant__ StackGeom g,
                                uint32_t* __restrict__ k_lin_lo, uint32_t* __restrict__ k_lin_hi,
                                uint32_t* __restrict__ k_resp, uint32_t* __restrict__ k_chunk,
                                int32_t* __restrict__ pos, int32_t* __restrict__ absz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const mmb_row r = rows[i];
  int c[3];
  chunk_coord(g, r.chunk, c);
  const long long Y = g.size[1][c[1]], X = g.size[2][c[2]];
  const long long lin = (((long long)r.z * Y + r.y) * X + r.x) * g.num_sigma + r.s;
  k_lin_lo[i] = (uint32_t)lin;
  k_lin_hi[i] = (uint32_t)(lin >> 32);
  k_resp[i] = desc_key(r.resp);
  k_chunk[i] = (uint32_t)r.chunk * (uint32_t)g.n_channels + (uint32_t)r.channel;
  const int z = r.z + g.start[0][c[0]], y = r.y + g.start[1][c[1]], x = r.x + g.start[2][c[2]];
  pos[i] = z; pos[n + i] = y; pos[2 * n + i] = x;
  absz[i] = z; absz[n + i] = y; absz[2 * n + i] = x;
}

__global__ void channel_flag_kernel(const mmb_row* __restrict__ rows, int n, int channel,
                                    uint32_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = rows[i].channel == channel ? 0u : 1u;
}

// class of every row of `cur` along one axis (see the file comment); counters:
// cnt[j * 4 + 0] rows in slab j (any chunk), cnt[j * 4 + 2] rows in the ratio region
__global__ void classify_kernel(const mmb_row* __restrict__ rows, const int32_t* __restrict__ cur,
                                const int* __restrict__ n_ptr, int n_max, int n_rows,
                                const int32_t* __restrict__ pos,
                                const __grid_constant__ StackGeom g,
                                const __grid_constant__ AxisPlan p, uint32_t* __restrict__ cls,
                                int* __restrict__ cnt) {
  const int n = min(__ldcg(n_ptr), n_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int row = __ldcg(cur + i);
  const int q = __ldcg(pos + (size_t)p.axis * n_rows + row);
  int c[3];
  chunk_coord(g, rows[row].chunk, c);
  const int tag = c[p.axis];
  const uint32_t drop = 3u * p.n_sec - 2u;
  uint32_t k = drop;
  for (int j = 0; j < p.n_sec; ++j) {
    if (q >= p.lower[j] && q < p.upper[j]) { k = (uint32_t)j; break; }
    if (j < p.n_sec - 1 && q >= p.upper[j] && q < p.slab_hi[j]) {
      atomicAdd(&cnt[j * 4 + 0], 1);
      if (tag == j) k = (uint32_t)(p.n_sec + 2 * j);
      else if (tag == j + 1) k = (uint32_t)(p.n_sec + 2 * j + 1);
      break;
    }
  }
  for (int j = 0; j < p.n_sec - 1; ++j)
    if (q >= p.nlo[j] && q < p.nhi[j]) atomicAdd(&cnt[j * 4 + 2], 1);
  cls[row] = k;
}

// seg[k] = first position of class k in the class-sorted list (classes 0 .. n_cls,
// seg[n_cls] = n); absent classes inherit the start of the next present one
__global__ void seg_mark_kernel(const int32_t* __restrict__ cur, const int* __restrict__ n_ptr,
                                int n_max, const uint32_t* __restrict__ cls,
                                int* __restrict__ seg) {
  const int n = min(__ldcg(n_ptr), n_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = cls[__ldcg(cur + i)];
  if (i == 0 || cls[__ldcg(cur + i - 1)] != k) seg[k] = i;
}
__global__ void seg_fill_kernel(int* __restrict__ seg, int n_cls, const int* __restrict__ n_ptr,
                                int n_max) {
  if (threadIdx.x || blockIdx.x) return;
  seg[n_cls] = min(__ldcg(n_ptr), n_max);
  for (int k = n_cls - 1; k >= 0; --k)
    if (seg[k] < 0) seg[k] = seg[k + 1];
}

// check rows -> strip buckets (counting sort: histogram, scan, scatter of positions)
__device__ __forceinline__ bool check_bucket(const AxisPlan& p, uint32_t k, int u, int* bucket) {
  if (k < (uint32_t)p.n_sec || k >= 3u * p.n_sec - 2u) return false;
  const int t = (int)k - p.n_sec;
  if (!(t & 1)) return false;
  *bucket = (t >> 1) * p.n_strips + (u >> p.strip_shift);
  return true;
}
__global__ void bucket_hist_kernel(const int32_t* __restrict__ cur, const int* __restrict__ seg,
                                   int n_max, int n_rows, const uint32_t* __restrict__ cls,
                                   const int32_t* __restrict__ pos,
                                   const __grid_constant__ AxisPlan p, int* __restrict__ bhist) {
  const int lo = __ldcg(seg + p.n_sec), hi = min(__ldcg(seg + 3 * p.n_sec - 2), n_max);
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const int row = __ldcg(cur + i);
  int b;
  if (check_bucket(p, cls[row], __ldcg(pos + (size_t)p.u_axis * n_rows + row), &b))
    atomicAdd(&bhist[b], 1);
}
__global__ void __launch_bounds__(1024)
bucket_scan_kernel(int* __restrict__ bhist, int nb) {
  __shared__ int part[1024];
  const int per = (nb + 1023) / 1024;
  const int lo = min(nb, (int)threadIdx.x * per), hi = min(nb, lo + per);
  int sum = 0;
  for (int k = lo; k < hi; ++k) sum += bhist[k];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const int v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = part[threadIdx.x] - sum;
  for (int k = lo; k < hi; ++k) { const int c = bhist[k]; bhist[k] = run; run += c; }
}
__global__ void bucket_scatter_kernel(const int32_t* __restrict__ cur, const int* __restrict__ seg,
                                      int n_max, int n_rows, const uint32_t* __restrict__ cls,
                                      const int32_t* __restrict__ pos,
                                      const __grid_constant__ AxisPlan p, int* __restrict__ cursor,
                                      int32_t* __restrict__ bucketed) {
  const int lo = __ldcg(seg + p.n_sec), hi = min(__ldcg(seg + 3 * p.n_sec - 2), n_max);
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const int row = __ldcg(cur + i);
  int b;
  if (check_bucket(p, cls[row], __ldcg(pos + (size_t)p.u_axis * n_rows + row), &b))
    bucketed[atomicAdd(&cursor[b], 1)] = i;       // position in the class-sorted list
}

// every master against the checks of its seam (after the scatter bend[b] = one past the
// last entry of bucket b): last[row] = largest matching position or -1, hit[check] = 1
__global__ void seam_match_rows_kernel(const int32_t* __restrict__ cur,
                                       const int* __restrict__ seg, int n_max, int n_rows,
                                       const uint32_t* __restrict__ cls,
                                       const int32_t* __restrict__ pos,
                                       const __grid_constant__ StackGeom g,
                                       const __grid_constant__ AxisPlan p,
                                       const int* __restrict__ bend,
                                       const int32_t* __restrict__ bucketed,
                                       int32_t* __restrict__ last, uint32_t* __restrict__ hit,
                                       int* __restrict__ cnt) {
  const int lo = __ldcg(seg + p.n_sec), hi = min(__ldcg(seg + 3 * p.n_sec - 2), n_max);
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const int row = __ldcg(cur + i);
  const uint32_t k = cls[row];
  const int t = (int)k - p.n_sec;
  if (t & 1) return;                               // a check row
  const int j = t >> 1;
  const int mz = __ldcg(pos + row), my = __ldcg(pos + n_rows + row),
            mx = __ldcg(pos + 2 * (size_t)n_rows + row);
  const int mu = p.u_axis == 0 ? mz : (p.u_axis == 1 ? my : mx);
  const int s0 = max((mu >> p.strip_shift) - 1, 0),
            s1 = min((mu >> p.strip_shift) + 1, p.n_strips - 1);
  const int b0 = j * p.n_strips + s0, b1 = j * p.n_strips + s1;
  int e = b0 > 0 ? __ldcg(bend + b0 - 1) : 0;
  const int e_end = __ldcg(bend + b1);
  int best = -1;
  for (; e < e_end; ++e) {
    const int ci = __ldcg(bucketed + e);
    const int crow = __ldcg(cur + ci);
    const int dz = abs(__ldcg(pos + crow) - mz), dy = abs(__ldcg(pos + n_rows + crow) - my),
              dx = abs(__ldcg(pos + 2 * (size_t)n_rows + crow) - mx);
    if (dz <= g.tol[0] && dy <= g.tol[1] && dx <= g.tol[2]) {
      best = max(best, ci);
      if (atomicExch(&hit[crow], 1u) == 0u) atomicAdd(&cnt[j * 4 + 3], 1);
    }
  }
  last[row] = best;
}

__device__ __forceinline__ int mean_half_even(int a, int b) {
  // np.around((a + b) / 2): halves go to the even neighbour
  const int s = a + b;
  if (!(s & 1)) return s / 2;
  const int f = (s - 1) / 2;                       // floor for s >= 0
  return (f & 1) ? f + 1 : f;
}

__global__ void seam_apply_kernel(const int32_t* __restrict__ cur, const int* __restrict__ seg,
                                  int n_max, int n_rows, const uint32_t* __restrict__ cls,
                                  const __grid_constant__ AxisPlan p,
                                  const int32_t* __restrict__ last, int32_t* __restrict__ absz) {
  const int lo = __ldcg(seg + p.n_sec), hi = min(__ldcg(seg + 3 * p.n_sec - 2), n_max);
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const int row = __ldcg(cur + i);
  if (((int)cls[row] - p.n_sec) & 1) return;
  const int ci = last[row];
  if (ci < 0) return;
  const int crow = __ldcg(cur + ci);
  for (int a = 0; a < 3; ++a)
    absz[(size_t)a * n_rows + row] =
        mean_half_even(absz[(size_t)a * n_rows + row], absz[(size_t)a * n_rows + crow]);
}

// cnt[j*4+1] = rows of seam j after pruning (masters + unmatched checks)
__global__ void seam_counts_kernel(const int* __restrict__ seg, int n_sec, int* __restrict__ cnt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_sec - 1) return;
  const int nm = seg[n_sec + 2 * j + 1] - seg[n_sec + 2 * j];
  const int nc = seg[n_sec + 2 * j + 2] - seg[n_sec + 2 * j + 1];
  cnt[j * 4 + 1] = nm + nc - cnt[j * 4 + 3];
}

__global__ void reset_rows_kernel(const int32_t* __restrict__ cur, const int* __restrict__ n_ptr,
                                  int n_max, uint32_t* __restrict__ hit,
                                  int32_t* __restrict__ last) {
  const int n = min(__ldcg(n_ptr), n_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int row = __ldcg(cur + i);
  hit[row] = 0u;
  last[row] = -1;
}

// final rows of one channel -> the float64 table, appended at *total
__global__ void table_kernel(const mmb_row* __restrict__ rows, const int32_t* __restrict__ cur,
                             const int* __restrict__ n_ptr, int n_max, int n_rows,
                             const int32_t* __restrict__ pos, const int32_t* __restrict__ absz,
                             const double* __restrict__ sigmas, int num_sigma,
                             const int32_t* __restrict__ channel_ids, int final_layout,
                             const int* __restrict__ total, double* __restrict__ out) {
  const int n = min(__ldcg(n_ptr), n_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int row = __ldcg(cur + i);
  const mmb_row r = rows[row];
  const double radius = __dmul_rn(sigmas[r.channel * num_sigma + r.s], 1.7320508075688772);
  const double az = absz[row], ay = absz[n_rows + row], ax = absz[2 * (size_t)n_rows + row];
  const double chl = (double)channel_ids[r.channel];
  const size_t at = (size_t)(__ldcg(total) + i);
  if (final_layout) {
    // z, y, x (seam-averaged absolute), radius, confirmed, truth, channel, region
    double* o = out + at * 8;
    o[0] = az; o[1] = ay; o[2] = ax; o[3] = radius; o[4] = -1.0; o[5] = -1.0; o[6] = chl;
    o[7] = -1.0;
  } else {
    double* o = out + at * 11;
    o[0] = pos[row]; o[1] = pos[n_rows + row]; o[2] = pos[2 * (size_t)n_rows + row];
    o[3] = radius; o[4] = -1.0; o[5] = -1.0; o[6] = chl;
    o[7] = az; o[8] = ay; o[9] = ax; o[10] = -1.0;
  }
}
__global__ void add_total_kernel(int* __restrict__ total, const int* __restrict__ n_ptr,
                                 int n_max) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *total += min(__ldcg(n_ptr), n_max);
}

__global__ void rows_from_cands_kernel(const mmb_cand* __restrict__ cand, int n, int chunk,
                                       int channel, mmb_row* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const mmb_cand c = load_cand(cand + i);
  mmb_row r;
  r.z = c.z; r.y = c.y; r.x = c.x; r.s = c.s; r.resp = c.resp;
  r.chunk = chunk; r.channel = channel; r.reserved = 0;
  out[i] = r;
}

// candidates inside a box (and, optionally, flagged by `keep`), shifted, appended to a
// shared list with one atomic per warp; the counter keeps counting past the capacity
__global__ void cands_append_kernel(const mmb_cand* __restrict__ cand, int n,
                                    const uint8_t* __restrict__ keep, int sz, int sy, int sx,
                                    int lz, int ly, int lx, int hz, int hy, int hx,
                                    mmb_cand* __restrict__ dst, int capacity,
                                    int* __restrict__ counter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  mmb_cand c;
  bool ok = false;
  if (i < n && (!keep || __ldcg(keep + i))) {
    c = load_cand(cand + i);
    c.z += sz; c.y += sy; c.x += sx;
    ok = c.z >= lz && c.z < hz && c.y >= ly && c.y < hy && c.x >= lx && c.x < hx;
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, ok);
  if (!ballot) return;
  int base = 0;
  if (lane == __ffs(ballot) - 1) base = atomicAdd(counter, __popc(ballot));
  base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
  if (ok) {
    const int at = base + __popc(ballot & ((1u << lane) - 1u));
    if (at < capacity) dst[at] = c;
  }
}

static inline int64_t al(int64_t x) { return (x + 255) / 256 * 256; }
constexpr int64_t kSigArea = 8192;       // sigma ladders + channel ids at the workspace tail

struct TableLayout {
  int nblk;
  int64_t perm, keys, pos, hist, bucketed, small, total;
};
static TableLayout table_layout(int64_t n) {
  TableLayout L;
  if (n < 1) n = 1;
  L.nblk = (int)cdiv(n, kRsItems);
  L.perm = al(n * 4);                 // x3: ping, pong, merged order
  L.keys = al(n * 4);                 // x6: lin lo/hi, resp, chunk, class, hit/flag
  L.pos = al(3 * n * 4);              // x2: pos, abs  (+ last: n ints in `bucketed` slot 2)
  L.hist = al((int64_t)256 * L.nblk * 4);
  L.bucketed = al(n * 4);             // x2: bucketed positions, last match
  L.small = al((int64_t)(kMaxSec * 129 + 3 * kMaxSec + 16 + 257 + 3 * kMaxSec * 4 + 64) * 4);
  L.total = 3 * L.perm + 6 * L.keys + 2 * L.pos + L.hist + 2 * L.bucketed + L.small + kSigArea;
  return L;
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_rows_from_cands(const mmb_cand* cand, int n, int chunk, int channel,
                                   mmb_row* out, void* stream) {
  MMB_REQUIRE(n >= 0 && chunk >= 0 && channel >= 0, "bad arguments");
  if (n == 0) return MMB_OK;
  MMB_REQUIRE(cand && out, "null buffer");
  rows_from_cands_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
      cand, n, chunk, channel, out);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

extern "C" int mmb_cands_append(const mmb_cand* cand, int n, const uint8_t* keep,
                                const int32_t shift[3], const int32_t lo[3],
                                const int32_t hi[3], mmb_cand* dst, int capacity,
                                int32_t* counter, void* stream) {
  MMB_REQUIRE(n >= 0 && capacity >= 0 && shift && lo && hi, "bad arguments");
  if (n == 0) return MMB_OK;
  MMB_REQUIRE(cand && dst && counter, "null buffer");
  cands_append_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
      cand, n, keep, shift[0], shift[1], shift[2], lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], dst,
      capacity, counter);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

extern "C" int64_t mmb_stack_tables_work_bytes(int n_rows) { return table_layout(n_rows).total; }

extern "C" int mmb_stack_tables(const mmb_row* rows, int n, const mmb_stack_geom* geom,
                                int final_layout, double* out_table, int32_t* n_out,
                                int32_t* seam_counts, void* work, void* stream) {
  MMB_REQUIRE(geom && n_out && work, "null buffer");
  MMB_REQUIRE(n >= 0, "bad row count");
  cudaStream_t st = (cudaStream_t)stream;
  MMB_CHECK_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int32_t), st));
  if (seam_counts)
    MMB_CHECK_CUDA(cudaMemsetAsync(seam_counts, 0,
                                   (size_t)geom->n_channels * 3 * kMaxSec * 4 * sizeof(int32_t), st));
  if (n == 0) return MMB_OK;
  MMB_REQUIRE(rows && out_table, "null buffer");
  MMB_REQUIRE(geom->n_channels >= 1 && geom->n_channels <= 64, "1..64 channels");
  MMB_REQUIRE(geom->num_sigma >= 1 && geom->sigmas && geom->channel_ids, "sigma ladders");
  StackGeom g;
  memset(&g, 0, sizeof(g));
  int64_t n_chunks = 1;
  for (int a = 0; a < 3; ++a) {
    MMB_REQUIRE(geom->grid[a] >= 1 && geom->start[a] && geom->size[a], "chunk grid");
    if (geom->grid[a] > kMaxSec) {
      set_error("%d chunk sections along axis %d exceed the supported %d", geom->grid[a], a, kMaxSec);
      return MMB_ERR_UNSUPPORTED;
    }
    g.grid[a] = geom->grid[a];
    g.tol[a] = geom->tol[a];
    n_chunks *= geom->grid[a];
    for (int j = 0; j < geom->grid[a]; ++j) {
      g.start[a][j] = geom->start[a][j];
      g.size[a][j] = geom->size[a][j];
    }
  }
  g.n_channels = geom->n_channels;
  g.num_sigma = geom->num_sigma;
  MMB_REQUIRE(n_chunks * geom->n_channels < ((int64_t)1 << 31), "chunk x channel keys overflow");

  const TableLayout L = table_layout(n);
  char* w = (char*)work;
  int32_t* perm_a = (int32_t*)w;                  w += L.perm;
  int32_t* perm_b = (int32_t*)w;                  w += L.perm;
  int32_t* merged = (int32_t*)w;                  w += L.perm;
  uint32_t* k_lin_lo = (uint32_t*)w;              w += L.keys;
  uint32_t* k_lin_hi = (uint32_t*)w;              w += L.keys;
  uint32_t* k_resp = (uint32_t*)w;                w += L.keys;
  uint32_t* k_chunk = (uint32_t*)w;               w += L.keys;
  uint32_t* cls = (uint32_t*)w;                   w += L.keys;
  uint32_t* hit = (uint32_t*)w;                   w += L.keys;
  int32_t* pos = (int32_t*)w;                     w += L.pos;
  int32_t* absz = (int32_t*)w;                    w += L.pos;
  int* hist = (int*)w;                            w += L.hist;
  int32_t* bucketed = (int32_t*)w;                w += L.bucketed;
  int32_t* last = (int32_t*)w;                    w += L.bucketed;
  int* small = (int*)w;
  int* bhist = small;                                   // kMaxSec * 129
  int* seg = bhist + kMaxSec * 129;                     // 3 * kMaxSec + 16
  int* digit_start = seg + 3 * kMaxSec + 16;            // 257
  int* cnt = digit_start + 257;                         // kMaxSec * 4 per axis (reused)
  int* n_all = cnt + 3 * kMaxSec * 4;                   // [0] = n, [1] = running count
  double* d_sig = (double*)((char*)small + L.small);    // 256-byte aligned tail area
  // sigma ladders and channel ids go up once (tiny, from the caller's host arrays)
  const size_t sig_bytes = (size_t)geom->n_channels * geom->num_sigma * sizeof(double);
  MMB_REQUIRE((int64_t)(sig_bytes + (size_t)geom->n_channels * 4) <= kSigArea,
              "sigma table too large for the workspace tail");
  int32_t* d_chl = (int32_t*)((char*)d_sig + sig_bytes);
  MMB_CHECK_CUDA(cudaMemcpyAsync(d_sig, geom->sigmas, sig_bytes, cudaMemcpyHostToDevice, st));
  MMB_CHECK_CUDA(cudaMemcpyAsync(d_chl, geom->channel_ids, (size_t)geom->n_channels * 4,
                                 cudaMemcpyHostToDevice, st));
  MMB_CHECK_CUDA(cudaMemcpyAsync(n_all, &n, sizeof(int), cudaMemcpyHostToDevice, st));
  // the host buffers above must outlive the copies: pageable copies are staged before
  // cudaMemcpyAsync returns, so they do

  Sorter S{perm_a, perm_b, hist, L.nblk, n, st};
  const unsigned nb = (unsigned)cdiv(n, 256);
  row_keys_kernel<<<nb, 256, 0, st>>>(rows, n, g, k_lin_lo, k_lin_hi, k_resp, k_chunk, pos, absz);
  MMB_CHECK_LAUNCH();

  // ---- merge order: LSD passes over the bytes that can differ ------------------------
  long long max_lin = 1;
  {
    long long mz = 1, my = 1, mx = 1;
    for (int j = 0; j < g.grid[0]; ++j) mz = g.size[0][j] > mz ? g.size[0][j] : mz;
    for (int j = 0; j < g.grid[1]; ++j) my = g.size[1][j] > my ? g.size[1][j] : my;
    for (int j = 0; j < g.grid[2]; ++j) mx = g.size[2][j] > mx ? g.size[2][j] : mx;
    max_lin = mz * my * mx * g.num_sigma;
  }
  const long long max_chunk_key = n_chunks * g.n_channels - 1;
  const int32_t* src = nullptr;
  int32_t* dst = perm_a;
  auto pass = [&](const uint32_t* key, int shift) -> int {
    int rc = radix_pass(S, key, shift, src, dst, n_all, nullptr);
    src = dst;
    dst = dst == perm_a ? perm_b : perm_a;
    return rc;
  };
  int rc = MMB_OK;
  for (int b = 0; b < 4 && !rc; ++b)
    if (b == 0 || (max_lin >> (8 * b)) != 0) rc = pass(k_lin_lo, 8 * b);
  for (int b = 0; b < 4 && !rc; ++b)
    if ((max_lin >> (32 + 8 * b)) != 0) rc = pass(k_lin_hi, 8 * b);
  for (int b = 0; b < 4 && !rc; ++b) rc = pass(k_resp, 8 * b);
  for (int b = 0; b < 4 && !rc; ++b)
    if (b == 0 || (max_chunk_key >> (8 * b)) != 0) rc = pass(k_chunk, 8 * b);
  if (rc) return rc;
  MMB_CHECK_CUDA(cudaMemcpyAsync(merged, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));

  // ---- per channel: seam pruning axis by axis, then the table ---------------------------
  for (int chl = 0; chl < g.n_channels; ++chl) {
    // rows of this channel in merge order = stable partition by a 1-bit key
    channel_flag_kernel<<<nb, 256, 0, st>>>(rows, n, chl, hit);
    MMB_CHECK_LAUNCH();
    int32_t* cur = perm_a;
    int32_t* oth = perm_b;
    rc = radix_pass(S, hit, 0, merged, cur, n_all, digit_start);
    if (rc) return rc;
    // n_cur lives on the device: count of flag 0 = digit_start[1]
    MMB_CHECK_CUDA(cudaMemcpyAsync(n_all + 1, digit_start + 1, sizeof(int),
                                   cudaMemcpyDeviceToDevice, st));
    const int* n_cur = n_all + 1;
    for (int axis = 0; axis < 3; ++axis) {
      const int n_sec = g.grid[axis];
      if (n_sec <= 1) continue;
      AxisPlan p;
      memset(&p, 0, sizeof(p));
      p.axis = axis; p.n_sec = n_sec;
      p.u_axis = axis == 1 ? 2 : 1;
      int u_extent = 0;
      for (int j = 0; j < g.grid[p.u_axis]; ++j) {
        const int e = g.start[p.u_axis][j] + g.size[p.u_axis][j];
        u_extent = e > u_extent ? e : u_extent;
      }
      // strips at least 64 voxels wide (far above any tolerance), at most 128 of them
      p.strip_shift = 6;
      while ((u_extent >> p.strip_shift) + 1 > 128) ++p.strip_shift;
      p.n_strips = (u_extent >> p.strip_shift) + 1;
      if (geom->tol[0] >= 64 || geom->tol[1] >= 64 || geom->tol[2] >= 64) {
        set_error("seam tolerance above 63 voxels is not supported");
        return MMB_ERR_UNSUPPORTED;
      }
      const int ov = geom->overlap[axis], pad = geom->pad[axis], tl = geom->tol[axis];
      const int shift = ov + pad;
      const int total_base = g.start[axis][n_sec - 1];
      for (int j = 0; j < n_sec; ++j) {
        const int start = g.start[axis][j], size = g.size[axis][j], end = start + size;
        p.lower[j] = start + (j > 0 ? shift : 0);
        if (j < n_sec - 1) {
          p.upper[j] = end - shift;
          p.slab_hi[j] = end + pad;
          const int nlo = end + tl, nhi = nlo + ov + 2 * pad, total = total_base + size;
          if (nlo < total && nhi < total) { p.nlo[j] = nlo; p.nhi[j] = nhi; }
          else { p.nlo[j] = 1; p.nhi[j] = 0; }
          // the classes must tile the axis (chunking.stack_splitter geometry)
          if (p.slab_hi[j] != g.start[axis][j + 1] + shift || p.upper[j] < p.lower[j]) {
            set_error("chunk sections along axis %d do not tile (section %d)", axis, j);
            return MMB_ERR_UNSUPPORTED;
          }
        } else {
          p.upper[j] = end;
        }
      }
      const int n_cls = 3 * n_sec - 2;              // class ids 0 .. n_cls (n_cls = dropped)
      int* acnt = cnt + axis * kMaxSec * 4;
      MMB_CHECK_CUDA(cudaMemsetAsync(acnt, 0, (size_t)kMaxSec * 4 * sizeof(int), st));
      MMB_CHECK_CUDA(cudaMemsetAsync(seg, 0xff, (size_t)(3 * kMaxSec + 16) * sizeof(int), st));
      MMB_CHECK_CUDA(cudaMemsetAsync(bhist, 0, (size_t)kMaxSec * 129 * sizeof(int), st));
      reset_rows_kernel<<<nb, 256, 0, st>>>(cur, n_cur, n, hit, last);
      MMB_CHECK_LAUNCH();
      classify_kernel<<<nb, 256, 0, st>>>(rows, cur, n_cur, n, n, pos, g, p, cls, acnt);
      MMB_CHECK_LAUNCH();
      rc = radix_pass(S, cls, 0, cur, oth, n_cur, nullptr);
      if (rc) return rc;
      rc = radix_pass(S, cls, 8, oth, cur, n_cur, nullptr);
      if (rc) return rc;
      seg_mark_kernel<<<nb, 256, 0, st>>>(cur, n_cur, n, cls, seg);
      MMB_CHECK_LAUNCH();
      seg_fill_kernel<<<1, 32, 0, st>>>(seg, n_cls + 1, n_cur, n);
      MMB_CHECK_LAUNCH();
      bucket_hist_kernel<<<nb, 256, 0, st>>>(cur, seg, n, n, cls, pos, p, bhist);
      MMB_CHECK_LAUNCH();
      bucket_scan_kernel<<<1, 1024, 0, st>>>(bhist, (n_sec - 1) * p.n_strips);
      MMB_CHECK_LAUNCH();
      bucket_scatter_kernel<<<nb, 256, 0, st>>>(cur, seg, n, n, cls, pos, p, bhist, bucketed);
      MMB_CHECK_LAUNCH();
      {
        ProfScope ps(PROF_SEAM, (double)n, st);
        seam_match_rows_kernel<<<nb, 256, 0, st>>>(cur, seg, n, n, cls, pos, g, p, bhist,
                                                   bucketed, last, hit, acnt);
      }
      MMB_CHECK_LAUNCH();
      seam_apply_kernel<<<nb, 256, 0, st>>>(cur, seg, n, n, cls, p, last, absz);
      MMB_CHECK_LAUNCH();
      seam_counts_kernel<<<1, kMaxSec, 0, st>>>(seg, n_sec, acnt);
      MMB_CHECK_LAUNCH();
      if (seam_counts)
        MMB_CHECK_CUDA(cudaMemcpyAsync(seam_counts + ((size_t)chl * 3 + axis) * kMaxSec * 4, acnt,
                                       (size_t)kMaxSec * 4 * sizeof(int), cudaMemcpyDeviceToDevice,
                                       st));
      // drop the matched checks and the rows of other chunks: stable partition of the
      // first seg[n_cls] rows (everything but the dropped class) by the hit flag
      rc = radix_pass(S, hit, 0, cur, oth, seg + n_cls, digit_start);
      if (rc) return rc;
      MMB_CHECK_CUDA(cudaMemcpyAsync(n_all + 1, digit_start + 1, sizeof(int),
                                     cudaMemcpyDeviceToDevice, st));
      int32_t* t = cur; cur = oth; oth = t;
    }
    table_kernel<<<nb, 256, 0, st>>>(rows, cur, n_cur, n, n, pos, absz, d_sig, g.num_sigma,
                                     d_chl, final_layout, n_out, out_table);
    MMB_CHECK_LAUNCH();
    add_total_kernel<<<1, 32, 0, st>>>(n_out, n_cur, n);
    MMB_CHECK_LAUNCH();
  }
  return MMB_OK;
}
