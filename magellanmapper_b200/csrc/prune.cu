// Overlap pruning inside a volume (_prune_blobs) and seam matching across chunks
// (remove_close_blobs).
//
// _prune_blobs walks the pairs closer than 2*sigma_max*sqrt(3) and, when the
// smaller sphere (radius sigma*sqrt(3)) has more than `overlap` of its volume
// inside the larger one, zeroes the smaller-sigma blob (on equal sigma the one
// listed first = stronger response).  A zeroed blob never removes another one,
// which makes the sequential result depend on the iteration order of a Python
// set.  Here the pair tests produce a kill graph (killer -> victim, acyclic) and
// the graph is resolved order-independently: a blob survives iff none of its
// killers survives.  DESIGN.md states where this can differ from one particular
// scikit-image run (chains only) and how the tests bound it.
#include <math.h>
#include <vector>
#include "common.cuh"

namespace mmb {

__device__ __forceinline__ bool listed_first(float ra, long long la, float rb, long long lb) {
  // order of peak_local_max: descending response, ties in C order
  return ra > rb || (ra == rb && la < lb);
}

// sphere-overlap fraction as skimage.feature.blob._blob_overlap computes it in
// float64: same operation order, no FMA contraction
__device__ double overlap_f64(const mmb_cand& a, const mmb_cand& b, double s1, double s2) {
  const double root = sqrt(3.0);
  double big, r1, r2;
  if (s1 > s2) { big = s1; r1 = 1.0; r2 = __ddiv_rn(s2, s1); }
  else         { big = s2; r1 = __ddiv_rn(s1, s2); r2 = 1.0; }
  const double den = __dmul_rn(big, root);
  const double ez = __dadd_rn(__ddiv_rn((double)b.z, den), -__ddiv_rn((double)a.z, den));
  const double ey = __dadd_rn(__ddiv_rn((double)b.y, den), -__ddiv_rn((double)a.y, den));
  const double ex = __dadd_rn(__ddiv_rn((double)b.x, den), -__ddiv_rn((double)a.x, den));
  const double ss = __dadd_rn(__dadd_rn(__dmul_rn(ez, ez), __dmul_rn(ey, ey)), __dmul_rn(ex, ex));
  const double d = __dsqrt_rn(ss);
  const double rs = __dadd_rn(r1, r2);
  if (d > rs) return 0.0;
  if (d <= fabs(__dadd_rn(r1, -r2))) return 1.0;
  const double pi = 3.141592653589793;
  // vol = pi/(12 d) * (r1+r2-d)^2 * (d^2 + 2 d (r1+r2) - 3 (r1^2 + r2^2) + 6 r1 r2)
  const double t0 = __ddiv_rn(pi, __dmul_rn(12.0, d));
  const double t1 = __dadd_rn(rs, -d);
  const double t1s = __dmul_rn(t1, t1);
  double poly = __dadd_rn(__dmul_rn(d, d), __dmul_rn(__dmul_rn(2.0, d), rs));
  poly = __dadd_rn(poly, -__dmul_rn(3.0, __dadd_rn(__dmul_rn(r1, r1), __dmul_rn(r2, r2))));
  poly = __dadd_rn(poly, __dmul_rn(__dmul_rn(6.0, r1), r2));
  const double vol = __dmul_rn(__dmul_rn(t0, t1s), poly);
  const double m = r1 < r2 ? r1 : r2;
  const double small = __dmul_rn(__dmul_rn(__ddiv_rn(4.0, 3.0), pi), __dmul_rn(__dmul_rn(m, m), m));
  return __ddiv_rn(vol, small);
}

constexpr int kTile = 256;

// Cell grid for the pair search: candidates are listed cell by cell (counting sort,
// below) with cubic cells at least as wide as the cut-off distance, so every partner of
// a candidate lies in the 27 cells around its own; cells are numbered z, y, x-major,
// which makes the three x-neighbours of a (z, y) cell row one contiguous range.
__host__ __device__ __forceinline__ int cell_of(const CellGrid& g, int z, int y, int x) {
  return ((z / g.cs) * g.ncy + y / g.cs) * g.ncx + x / g.cs;
}

// pairs i < j (positions in the cell-sorted list) closer than the cut-off; emits
// (killer, victim).  cell_end[c] = one past the last candidate of cell c.
__global__ void __launch_bounds__(kTile)
prune_edges_kernel(const mmb_cand* __restrict__ cand, const int* __restrict__ n_ptr, int n_max,
                   const __grid_constant__ SigmaLadder ladder, double overlap, int Y, int X,
                   int2* __restrict__ edges, int edge_cap, int* __restrict__ edge_count,
                   const __grid_constant__ CellGrid grid, const int* __restrict__ cell_end) {
  __shared__ double sigmas[kMaxSigmas];
  // the candidate count lives on the device (no host round trip); the grid is
  // sized for the buffer capacity and surplus CTAs leave at once
  const int n = min(__ldcg(n_ptr), n_max);
  if (blockIdx.x * kTile >= n) return;
  const int num_sigma = ladder.n;
  for (int k = threadIdx.x; k < num_sigma; k += kTile) sigmas[k] = ladder.s[k];
  __syncthreads();
  const int i = blockIdx.x * kTile + threadIdx.x;
  if (i >= n) return;
  const mmb_cand me = load_cand(cand + i);
  const double my_sigma = sigmas[me.s];
  const double smax = sigmas[num_sigma - 1] > sigmas[0] ? sigmas[num_sigma - 1] : sigmas[0];
  // spheres can only touch when |d| <= (s1+s2)*sqrt(3) <= 2*smax*sqrt(3)
  const float cut = (float)(2.0 * smax * 1.7320508075688772) + 1.0f;
  const float cut2 = cut * cut;
  const int cz = me.z / grid.cs, cy = me.y / grid.cs, cx = me.x / grid.cs;
  const int x_lo = max(cx - 1, 0), x_hi = min(cx + 1, grid.ncx - 1);
  for (int dz = -1; dz <= 1; ++dz) {
    const int zz = cz + dz;
    if (zz < 0 || zz >= grid.ncz) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = cy + dy;
      if (yy < 0 || yy >= grid.ncy) continue;
      const int c_first = (zz * grid.ncy + yy) * grid.ncx + x_lo;
      const int c_last = c_first + (x_hi - x_lo);
      int j = c_first > 0 ? __ldcg(cell_end + c_first - 1) : 0;
      const int j_end = __ldcg(cell_end + c_last);
      if (j <= i) j = i + 1;                        // pairs are emitted once, by the lower index
      for (; j < j_end; ++j) {
        const mmb_cand o = load_cand(cand + j);
        const float dzf = (float)(o.z - me.z);
        const float dyf = (float)(o.y - me.y), dxf = (float)(o.x - me.x);
        if (dzf * dzf + dyf * dyf + dxf * dxf > cut2) continue;
        const double so = sigmas[o.s];
        if (overlap_f64(me, o, my_sigma, so) > overlap) {
          int killer, victim;
          if (my_sigma > so) { killer = i; victim = j; }
          else if (so > my_sigma) { killer = j; victim = i; }
          else {
            const long long li = (((long long)me.z * Y + me.y) * X + me.x) * num_sigma + me.s;
            const long long lj = (((long long)o.z * Y + o.y) * X + o.x) * num_sigma + o.s;
            // equal sigma: the blob listed first by peak_local_max is removed
            if (listed_first(me.resp, li, o.resp, lj)) { victim = i; killer = j; }
            else { victim = j; killer = i; }
          }
          const int e = atomicAdd(edge_count, 1);
          if (e < edge_cap) edges[e] = make_int2(killer, victim);
        }
      }
    }
  }
}

// single-CTA fixed point over the kill graph. state: 0 unknown, 1 alive, 2 dead
__global__ void __launch_bounds__(1024)
prune_resolve_kernel(const int* __restrict__ n_ptr, int n_max, const int2* __restrict__ edges,
                     const int* __restrict__ n_edges_ptr, int edge_cap,
                     unsigned char* __restrict__ state, unsigned char* __restrict__ mark,
                     unsigned char* __restrict__ keep, int* __restrict__ od_count) {
  __shared__ int remaining;
  const int n = min(__ldcg(n_ptr), n_max);
  const int n_edges = min(__ldcg(n_edges_ptr), edge_cap);
  for (int v = threadIdx.x; v < n; v += blockDim.x) state[v] = 0;
  __syncthreads();
  while (true) {
    for (int v = threadIdx.x; v < n; v += blockDim.x) mark[v] = 0;
    if (threadIdx.x == 0) remaining = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < n_edges; e += blockDim.x) {
      const int2 kv = __ldcg(&edges[e]);
      if (state[kv.y] == 0) {
        const unsigned char sk = state[kv.x];
        if (sk == 1) atomicOr((unsigned int*)(mark + (kv.y & ~3)), 1u << (8 * (kv.y & 3)));
        else if (sk == 0) atomicOr((unsigned int*)(mark + (kv.y & ~3)), 2u << (8 * (kv.y & 3)));
      }
    }
    __syncthreads();
    int mine = 0;
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
      if (state[v] == 0) {
        const unsigned char m = mark[v];
        if (m & 1) state[v] = 2;
        else if (!(m & 2)) state[v] = 1;
        else ++mine;
      }
    }
    if (mine) atomicAdd(&remaining, mine);
    __syncthreads();
    const int r = remaining;
    __syncthreads();
    if (r == 0) break;
  }
  for (int v = threadIdx.x; v < n; v += blockDim.x) keep[v] = state[v] == 1 ? 1 : 0;
  // size of the set whose fate depends on scikit-image's pair iteration order: blobs
  // with at least one killer, none of which is a root (a blob nobody kills).  mark
  // bit 0 = has a killer, bit 1 = killed by a root.
  if (od_count) {
    __syncthreads();
    for (int v = threadIdx.x; v < n; v += blockDim.x) mark[v] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < n_edges; e += blockDim.x) {
      const int2 kv = __ldcg(&edges[e]);
      atomicOr((unsigned int*)(mark + (kv.y & ~3)), 1u << (8 * (kv.y & 3)));
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n_edges; e += blockDim.x) {
      const int2 kv = __ldcg(&edges[e]);
      if (!(mark[kv.x] & 1)) atomicOr((unsigned int*)(mark + (kv.y & ~3)), 2u << (8 * (kv.y & 3)));
    }
    __syncthreads();
    int mine = 0;
    for (int v = threadIdx.x; v < n; v += blockDim.x) mine += (mark[v] & 3) == 1 ? 1 : 0;
    if (mine) atomicAdd(od_count, mine);
  }
}

// ---- counting sort of the candidates by cell -----------------------------------------
// Local maxima arrive in atomic order; the pair search wants them cell by cell
// (CellGrid above): histogram, single-CTA exclusive scan, scatter.  The order inside a
// cell is arbitrary - the kill-graph resolution does not depend on the listing order.
// After the scatter hist[c] = one past the last candidate of cell c.
__global__ void cell_hist_kernel(const mmb_cand* __restrict__ cand, const int* __restrict__ n_ptr,
                                 int n_max, const __grid_constant__ CellGrid grid,
                                 int* __restrict__ hist) {
  const int n = min(__ldcg(n_ptr), n_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const mmb_cand c = load_cand(cand + i);
    atomicAdd(&hist[cell_of(grid, c.z, c.y, c.x)], 1);
  }
}

__global__ void __launch_bounds__(1024)
cell_scan_kernel(int* __restrict__ hist, int ncells) {
  __shared__ int part[1024];
  const int per = (ncells + 1023) / 1024;
  const int lo = min(ncells, (int)threadIdx.x * per), hi = min(ncells, lo + per);
  int sum = 0;
  for (int z = lo; z < hi; ++z) sum += __ldcg(&hist[z]);
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {          // Hillis-Steele inclusive scan
    const int v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = part[threadIdx.x] - sum;             // exclusive prefix of this thread's range
  for (int z = lo; z < hi; ++z) {
    const int c = __ldcg(&hist[z]);
    hist[z] = run;
    run += c;
  }
}

__global__ void cell_scatter_kernel(const mmb_cand* __restrict__ cand,
                                    const int* __restrict__ n_ptr, int n_max,
                                    const __grid_constant__ CellGrid grid,
                                    int* __restrict__ cursor, mmb_cand* __restrict__ out) {
  const int n = min(__ldcg(n_ptr), n_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const mmb_cand c = load_cand(cand + i);
    out[atomicAdd(&cursor[cell_of(grid, c.z, c.y, c.x)], 1)] = c;
  }
}

// cells at least `min_edge` voxels wide over a (Z, Y, X) volume; the edge grows until
// the grid fits `max_cells` (the caller's scratch)
CellGrid make_cell_grid(int Z, int Y, int X, int min_edge, int64_t max_cells) {
  CellGrid g;
  g.cs = min_edge < 1 ? 1 : min_edge;
  for (;;) {
    g.ncz = (int)cdiv(Z, g.cs); g.ncy = (int)cdiv(Y, g.cs); g.ncx = (int)cdiv(X, g.cs);
    if ((int64_t)g.ncz * g.ncy * g.ncx <= max_cells) break;
    g.cs += (g.cs + 3) / 4;
  }
  return g;
}
int64_t cell_count(const CellGrid& g) { return (int64_t)g.ncz * g.ncy * g.ncx; }

// cand[0 .. min(*n_ptr, n_max)) -> out, listed cell by cell; hist = cell_count ints
int sort_by_cell_enqueue(const mmb_cand* cand, const int* n_ptr, int n_max, const CellGrid& grid,
                         int* hist, mmb_cand* out, cudaStream_t st) {
  if (n_max <= 0) return MMB_OK;
  const int ncells = (int)cell_count(grid);
  MMB_CHECK_CUDA(cudaMemsetAsync(hist, 0, (size_t)ncells * sizeof(int), st));
  cell_hist_kernel<<<(unsigned)cdiv(n_max, 256), 256, 0, st>>>(cand, n_ptr, n_max, grid, hist);
  MMB_CHECK_LAUNCH();
  cell_scan_kernel<<<1, 1024, 0, st>>>(hist, ncells);
  MMB_CHECK_LAUNCH();
  cell_scatter_kernel<<<(unsigned)cdiv(n_max, 256), 256, 0, st>>>(cand, n_ptr, n_max, grid, hist,
                                                                  out);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

// cell edge that covers the pair cut-off 2 * sigma_max * sqrt(3) + 1 of a ladder
int prune_cell_edge(const SigmaLadder& ladder) {
  double smax = 0.0;
  for (int k = 0; k < ladder.n; ++k) smax = ladder.s[k] > smax ? ladder.s[k] : smax;
  return (int)ceil(2.0 * smax * 1.7320508075688772 + 1.0) + 1;
}

int make_ladder(const double* sigmas_host, int num_sigma, SigmaLadder* out) {
  if (num_sigma < 1 || num_sigma > kMaxSigmas) {
    set_error("num_sigma %d outside 1..%d", num_sigma, kMaxSigmas);
    return MMB_ERR_UNSUPPORTED;
  }
  out->n = num_sigma;
  for (int k = 0; k < kMaxSigmas; ++k) out->s[k] = k < num_sigma ? sigmas_host[k] : 0.0;
  return MMB_OK;
}

__global__ void max_z_kernel(const mmb_cand* __restrict__ cand, int n, int* __restrict__ zmax) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int z = i < n ? __ldcg(&cand[i].z) : 0;
  z = __reduce_max_sync(0xffffffffu, z);
  if ((threadIdx.x & 31) == 0 && z > 0) atomicMax(zmax, z);
}

__global__ void keep_all_kernel(const int* __restrict__ n_ptr, int n_max, uint8_t* __restrict__ keep) {
  const int n = min(__ldcg(n_ptr), n_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = 1;
}

// Fully asynchronous pruning of the first min(*n_ptr, n_max) candidates, which must be
// listed cell by cell (sort_by_cell_enqueue with `grid`, cell_end = its histogram).
// Scratch is caller-provided: edges[edge_cap], edge_count (1 int), state[2 * (n_max + 8)]
// bytes.  *edge_count may exceed edge_cap afterwards: the caller must check and redo.
// od_count (may be NULL): incremented by the size of the order-dependent set.
int prune_within_enqueue(const mmb_cand* cand, const int* n_ptr, int n_max,
                         const SigmaLadder& ladder, double overlap, int Y, int X, int2* edges,
                         int edge_cap, int* edge_count, unsigned char* state, uint8_t* keep,
                         cudaStream_t st, const CellGrid& grid, const int* cell_end,
                         int* od_count) {
  if (n_max <= 0) return MMB_OK;
  const int64_t npad = (n_max + 3) / 4 * 4 + 4;
  MMB_CHECK_CUDA(cudaMemsetAsync(edge_count, 0, sizeof(int), st));
  if (overlap >= 1.0) {
    // the overlap fraction never exceeds 1, so nothing can be removed: the callers
    // that prune later over a larger candidate set (seamless slabs) land here
    keep_all_kernel<<<(unsigned)cdiv(n_max, 256), 256, 0, st>>>(n_ptr, n_max, keep);
    MMB_CHECK_LAUNCH();
    return MMB_OK;
  }
  {
    ProfScope ps(PROF_PRUNE_EDGES, n_max, st);
    prune_edges_kernel<<<(unsigned)cdiv(n_max, kTile), kTile, 0, st>>>(
        cand, n_ptr, n_max, ladder, overlap, Y, X, edges, edge_cap, edge_count, grid, cell_end);
  }
  MMB_CHECK_LAUNCH();
  {
    ProfScope ps(PROF_PRUNE_RESOLVE, n_max, st);
    prune_resolve_kernel<<<1, 1024, 0, st>>>(n_ptr, n_max, edges, edge_count, edge_cap, state,
                                             state + npad, keep, od_count);
  }
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

// keep flags refer to the CALLER's order: the candidates are bucketed into a scratch
// copy and pruned there; the scatter records where every candidate went, and each
// original candidate then looks its own flag up by that position.
__global__ void cell_scatter_track_kernel(const mmb_cand* __restrict__ cand, int n,
                                          const __grid_constant__ CellGrid grid,
                                          int* __restrict__ cursor, mmb_cand* __restrict__ out,
                                          int* __restrict__ where) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const mmb_cand c = load_cand(cand + i);
    const int at = atomicAdd(&cursor[cell_of(grid, c.z, c.y, c.x)], 1);
    out[at] = c;
    where[i] = at;
  }
}
__global__ void gather_keep_kernel(const uint8_t* __restrict__ keep_sorted,
                                   const int* __restrict__ where, int n,
                                   uint8_t* __restrict__ keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = keep_sorted[where[i]];
}

int prune_within_impl(const mmb_cand* cand, int n, const double* sigmas_host, int num_sigma,
                      double overlap, int Y, int X, uint8_t* keep, cudaStream_t st, int z_sorted) {
  (void)z_sorted;       // any listing order is accepted: candidates are bucketed here
  if (n == 0) return MMB_OK;
  SigmaLadder ladder;
  int rc = make_ladder(sigmas_host, num_sigma, &ladder);
  if (rc) return rc;
  // the ABI passes Y and X only: the z extent of the coordinates is found on the device
  int* d_counts = nullptr;          // [0] = n, [1] = edge count, [2] = max z
  int2* d_edges = nullptr;
  unsigned char* d_state = nullptr;
  mmb_cand* d_sorted = nullptr;
  int* d_where = nullptr;
  int* d_cells = nullptr;
  uint8_t* d_keep = nullptr;
  const int64_t npad = (n + 3) / 4 * 4 + 4;
  MMB_CHECK_CUDA(cudaMallocAsync((void**)&d_counts, 4 * sizeof(int), st));
  MMB_CHECK_CUDA(cudaMemsetAsync(d_counts, 0, 4 * sizeof(int), st));
  MMB_CHECK_CUDA(cudaMemcpyAsync(d_counts, &n, sizeof(int), cudaMemcpyHostToDevice, st));
  max_z_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(cand, n, d_counts + 2);
  MMB_CHECK_LAUNCH();
  int zmax = 0;
  MMB_CHECK_CUDA(cudaMemcpyAsync(&zmax, d_counts + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
  MMB_CHECK_CUDA(cudaStreamSynchronize(st));
  const CellGrid grid = make_cell_grid(zmax + 1, Y, X, prune_cell_edge(ladder),
                                       (int64_t)1 << 26);
  const int64_t ncells = cell_count(grid);
  MMB_CHECK_CUDA(cudaMallocAsync((void**)&d_state, 2 * npad, st));
  MMB_CHECK_CUDA(cudaMallocAsync((void**)&d_sorted, (size_t)n * sizeof(mmb_cand), st));
  MMB_CHECK_CUDA(cudaMallocAsync((void**)&d_where, (size_t)n * sizeof(int), st));
  MMB_CHECK_CUDA(cudaMallocAsync((void**)&d_keep, (size_t)n, st));
  MMB_CHECK_CUDA(cudaMallocAsync((void**)&d_cells, (size_t)ncells * sizeof(int), st));
  MMB_CHECK_CUDA(cudaMemsetAsync(d_cells, 0, (size_t)ncells * sizeof(int), st));
  cell_hist_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(cand, d_counts, n, grid, d_cells);
  MMB_CHECK_LAUNCH();
  cell_scan_kernel<<<1, 1024, 0, st>>>(d_cells, (int)ncells);
  MMB_CHECK_LAUNCH();
  cell_scatter_track_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(cand, n, grid, d_cells,
                                                                    d_sorted, d_where);
  MMB_CHECK_LAUNCH();
  int64_t edge_cap = 4 * (int64_t)n + 4096;
  if (edge_cap > 0x7fffffff) edge_cap = 0x7fffffff;
  int n_edges = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    MMB_CHECK_CUDA(cudaMallocAsync((void**)&d_edges, (size_t)edge_cap * sizeof(int2), st));
    rc = prune_within_enqueue(d_sorted, d_counts, n, ladder, overlap, Y, X, d_edges,
                              (int)edge_cap, d_counts + 1, d_state, d_keep, st, grid, d_cells,
                              nullptr);
    if (rc) return rc;
    MMB_CHECK_CUDA(cudaMemcpyAsync(&n_edges, d_counts + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    MMB_CHECK_CUDA(cudaStreamSynchronize(st));
    MMB_CHECK_CUDA(cudaFreeAsync(d_edges, st));
    if (n_edges <= edge_cap) break;
    edge_cap = n_edges;
  }
  gather_keep_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(d_keep, d_where, n, keep);
  MMB_CHECK_LAUNCH();
  MMB_CHECK_CUDA(cudaFreeAsync(d_state, st));
  MMB_CHECK_CUDA(cudaFreeAsync(d_sorted, st));
  MMB_CHECK_CUDA(cudaFreeAsync(d_where, st));
  MMB_CHECK_CUDA(cudaFreeAsync(d_keep, st));
  MMB_CHECK_CUDA(cudaFreeAsync(d_cells, st));
  MMB_CHECK_CUDA(cudaFreeAsync(d_counts, st));
  MMB_CHECK_CUDA(cudaStreamSynchronize(st));
  if (n_edges > edge_cap) {
    set_error("kill-edge buffer overflow (%d > %lld)", n_edges, (long long)edge_cap);
    return MMB_ERR_OVERFLOW;
  }
  return MMB_OK;
}

// ---- seam matching -----------------------------------------------------------

__global__ void __launch_bounds__(kTile)
seam_match_kernel(const int32_t* __restrict__ master, int nm, const int32_t* __restrict__ check,
                  int nc, int tz, int ty, int tx, int32_t* __restrict__ master_last,
                  unsigned char* __restrict__ check_hit) {
  __shared__ int32_t tile[kTile * 3];
  const int i = blockIdx.x * kTile + threadIdx.x;
  int mz = 0, my = 0, mx = 0;
  if (i < nm) { mz = master[3 * i]; my = master[3 * i + 1]; mx = master[3 * i + 2]; }
  int last = -1;
  for (int j0 = 0; j0 < nc; j0 += kTile) {
    __syncthreads();
    const int cnt = nc - j0 < kTile ? nc - j0 : kTile;
    for (int t = threadIdx.x; t < cnt * 3; t += kTile) tile[t] = check[3 * j0 + t];
    __syncthreads();
    if (i >= nm) continue;
    for (int t = 0; t < cnt; ++t) {
      const int dz = abs(tile[3 * t] - mz), dy = abs(tile[3 * t + 1] - my),
                dx = abs(tile[3 * t + 2] - mx);
      if (dz <= tz && dy <= ty && dx <= tx) {
        last = j0 + t;            // ascending j: the last hit is the largest index
        check_hit[j0 + t] = 1;
      }
    }
  }
  if (i < nm) master_last[i] = last;
}

int prune_seams_impl(const int32_t* master, int nm, const int32_t* check, int nc,
                     const int32_t tol[3], int32_t* master_last, uint8_t* check_hit,
                     cudaStream_t st) {
  if (nc > 0) MMB_CHECK_CUDA(cudaMemsetAsync(check_hit, 0, nc, st));
  if (nm == 0) return MMB_OK;
  {
    ProfScope ps(PROF_SEAM, (double)nm * nc, st);
    seam_match_kernel<<<(unsigned)cdiv(nm, kTile), kTile, 0, st>>>(
        master, nm, check, nc, tol[0], tol[1], tol[2], master_last, check_hit);
  }
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

}  // namespace mmb

extern "C" int mmb_prune_within(const mmb_cand* cand, int n, const double* sigmas, int num_sigma,
                                double overlap, int Y, int X, uint8_t* keep, void* stream) {
  MMB_REQUIRE(n >= 0 && num_sigma > 0 && sigmas, "bad arguments");
  MMB_REQUIRE(n == 0 || (cand && keep), "null buffer");
  return mmb::prune_within_impl(cand, n, sigmas, num_sigma, overlap, Y, X, keep,
                                (cudaStream_t)stream, 0);
}

extern "C" int mmb_prune_within_zsorted(const mmb_cand* cand, int n, const double* sigmas,
                                        int num_sigma, double overlap, int Y, int X,
                                        uint8_t* keep, void* stream) {
  MMB_REQUIRE(n >= 0 && num_sigma > 0 && sigmas, "bad arguments");
  MMB_REQUIRE(n == 0 || (cand && keep), "null buffer");
  return mmb::prune_within_impl(cand, n, sigmas, num_sigma, overlap, Y, X, keep,
                                (cudaStream_t)stream, 1);
}

extern "C" int mmb_prune_seams(const int32_t* master_zyx, int n_master, const int32_t* check_zyx,
                               int n_check, const int32_t tol[3], int32_t* master_last,
                               uint8_t* check_hit, void* stream) {
  MMB_REQUIRE(n_master >= 0 && n_check >= 0 && tol, "bad arguments");
  MMB_REQUIRE(n_master == 0 || (master_zyx && master_last), "null master buffer");
  MMB_REQUIRE(n_check == 0 || (check_zyx && check_hit), "null check buffer");
  return mmb::prune_seams_impl(master_zyx, n_master, check_zyx, n_check, tol, master_last,
                               check_hit, (cudaStream_t)stream);
}
