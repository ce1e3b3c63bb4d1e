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

// all pairs i < j, tile of j staged in shared memory; emits (killer, victim)
__global__ void __launch_bounds__(kTile)
prune_edges_kernel(const mmb_cand* __restrict__ cand, const int* __restrict__ n_ptr, int n_max,
                   const __grid_constant__ SigmaLadder ladder, double overlap, int Y, int X,
                   int2* __restrict__ edges, int edge_cap, int* __restrict__ edge_count,
                   int z_sorted) {
  __shared__ mmb_cand tile[kTile];
  __shared__ double sigmas[kMaxSigmas];
  // the candidate count lives on the device (no host round trip); the grid is
  // sized for the buffer capacity and surplus CTAs leave at once
  const int n = min(__ldcg(n_ptr), n_max);
  if (blockIdx.x * kTile >= n) return;
  const int num_sigma = ladder.n;
  for (int k = threadIdx.x; k < num_sigma; k += kTile) sigmas[k] = ladder.s[k];
  __syncthreads();
  const int i = blockIdx.x * kTile + threadIdx.x;
  mmb_cand me;
  me.z = me.y = me.x = 0; me.s = 0; me.resp = 0.f;
  double my_sigma = 0.0;
  if (i < n) { me = load_cand(cand + i); my_sigma = sigmas[me.s]; }
  const double smax = sigmas[num_sigma - 1] > sigmas[0] ? sigmas[num_sigma - 1] : sigmas[0];
  // spheres can only touch when |d| <= (s1+s2)*sqrt(3) <= 2*smax*sqrt(3)
  const float cut = (float)(2.0 * smax * 1.7320508075688772) + 1.0f;
  const float cut2 = cut * cut;
  // candidates listed by ascending z (the global prune of the seamless mode): once
  // a tile starts farther above this block's last candidate than the cut-off, so
  // does every later tile
  const int z_block_max =
      z_sorted ? __ldcg(&cand[min(n, (int)(blockIdx.x + 1) * kTile) - 1].z) : 0;
  // tiles with j > i only: start at this block's own tile
  for (int j0 = blockIdx.x * kTile; j0 < n; j0 += kTile) {
    __syncthreads();
    const int jl = j0 + threadIdx.x;
    if (jl < n) tile[threadIdx.x] = load_cand(cand + jl);
    __syncthreads();
    if (z_sorted && (float)(tile[0].z - z_block_max) > cut) break;      // block-uniform
    if (i >= n) continue;
    const int cnt = n - j0 < kTile ? n - j0 : kTile;
    for (int t = 0; t < cnt; ++t) {
      const int j = j0 + t;
      if (j <= i) continue;
      const mmb_cand o = tile[t];
      const float dz = (float)(o.z - me.z);
      if (fabsf(dz) > cut) continue;
      const float dy = (float)(o.y - me.y), dx = (float)(o.x - me.x);
      if (dz * dz + dy * dy + dx * dx > cut2) continue;
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

// single-CTA fixed point over the kill graph. state: 0 unknown, 1 alive, 2 dead
__global__ void __launch_bounds__(1024)
prune_resolve_kernel(const int* __restrict__ n_ptr, int n_max, const int2* __restrict__ edges,
                     const int* __restrict__ n_edges_ptr, int edge_cap,
                     unsigned char* __restrict__ state, unsigned char* __restrict__ mark,
                     unsigned char* __restrict__ keep) {
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
}

// ---- counting sort of the candidates by z plane -------------------------------------
// The pair search only has to look 2*sigma_max*sqrt(3)+1 planes up once the
// candidates are listed by ascending z (prune_edges_kernel's z_sorted path); local
// maxima arrive in atomic order, so they are bucketed by plane first: histogram,
// single-CTA exclusive scan, scatter.  The order inside a plane is arbitrary - the
// kill-graph resolution does not depend on the listing order.
__global__ void z_hist_kernel(const mmb_cand* __restrict__ cand, const int* __restrict__ n_ptr,
                              int n_max, int* __restrict__ hist) {
  const int n = min(__ldcg(n_ptr), n_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&hist[__ldcg(&cand[i].z)], 1);
}

__global__ void __launch_bounds__(1024)
z_scan_kernel(int* __restrict__ hist, int Z) {
  __shared__ int part[1024];
  const int per = (Z + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(Z, lo + per);
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

__global__ void z_scatter_kernel(const mmb_cand* __restrict__ cand, const int* __restrict__ n_ptr,
                                 int n_max, int* __restrict__ cursor,
                                 mmb_cand* __restrict__ out) {
  const int n = min(__ldcg(n_ptr), n_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const mmb_cand c = load_cand(cand + i);
    out[atomicAdd(&cursor[c.z], 1)] = c;
  }
}

// cand[0 .. min(*n_ptr, n_max)) -> out, listed by ascending z; hist = Z ints of scratch
int sort_by_z_enqueue(const mmb_cand* cand, const int* n_ptr, int n_max, int Z, int* hist,
                      mmb_cand* out, cudaStream_t st) {
  if (n_max <= 0) return MMB_OK;
  MMB_CHECK_CUDA(cudaMemsetAsync(hist, 0, (size_t)Z * sizeof(int), st));
  z_hist_kernel<<<(unsigned)cdiv(n_max, 256), 256, 0, st>>>(cand, n_ptr, n_max, hist);
  MMB_CHECK_LAUNCH();
  z_scan_kernel<<<1, 1024, 0, st>>>(hist, Z);
  MMB_CHECK_LAUNCH();
  z_scatter_kernel<<<(unsigned)cdiv(n_max, 256), 256, 0, st>>>(cand, n_ptr, n_max, hist, out);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
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

__global__ void keep_all_kernel(const int* __restrict__ n_ptr, int n_max, uint8_t* __restrict__ keep) {
  const int n = min(__ldcg(n_ptr), n_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = 1;
}

// Fully asynchronous pruning of the first min(*n_ptr, n_max) candidates.  Scratch is
// caller-provided: edges[edge_cap], edge_count (1 int), state[2 * (n_max + 8)] bytes.
// *edge_count may exceed edge_cap afterwards: the caller must check and redo.
int prune_within_enqueue(const mmb_cand* cand, const int* n_ptr, int n_max,
                         const SigmaLadder& ladder, double overlap, int Y, int X, int2* edges,
                         int edge_cap, int* edge_count, unsigned char* state, uint8_t* keep,
                         cudaStream_t st, int z_sorted) {
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
        cand, n_ptr, n_max, ladder, overlap, Y, X, edges, edge_cap, edge_count, z_sorted);
  }
  MMB_CHECK_LAUNCH();
  {
    ProfScope ps(PROF_PRUNE_RESOLVE, n_max, st);
    prune_resolve_kernel<<<1, 1024, 0, st>>>(n_ptr, n_max, edges, edge_count, edge_cap, state,
                                             state + npad, keep);
  }
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

int prune_within_impl(const mmb_cand* cand, int n, const double* sigmas_host, int num_sigma,
                      double overlap, int Y, int X, uint8_t* keep, cudaStream_t st, int z_sorted) {
  if (n == 0) return MMB_OK;
  SigmaLadder ladder;
  int rc = make_ladder(sigmas_host, num_sigma, &ladder);
  if (rc) return rc;
  int* d_counts = nullptr;          // [0] = n, [1] = edge count
  int2* d_edges = nullptr;
  unsigned char* d_state = nullptr;
  const int64_t npad = (n + 3) / 4 * 4 + 4;
  MMB_CHECK_CUDA(cudaMallocAsync((void**)&d_counts, 2 * sizeof(int), st));
  MMB_CHECK_CUDA(cudaMallocAsync((void**)&d_state, 2 * npad, st));
  MMB_CHECK_CUDA(cudaMemcpyAsync(d_counts, &n, sizeof(int), cudaMemcpyHostToDevice, st));
  int edge_cap = 4 * n + 4096;
  int n_edges = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    MMB_CHECK_CUDA(cudaMallocAsync((void**)&d_edges, (size_t)edge_cap * sizeof(int2), st));
    rc = prune_within_enqueue(cand, d_counts, n, ladder, overlap, Y, X, d_edges, edge_cap,
                              d_counts + 1, d_state, keep, st, z_sorted);
    if (rc) return rc;
    MMB_CHECK_CUDA(cudaMemcpyAsync(&n_edges, d_counts + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    MMB_CHECK_CUDA(cudaStreamSynchronize(st));
    MMB_CHECK_CUDA(cudaFreeAsync(d_edges, st));
    if (n_edges <= edge_cap) break;
    edge_cap = n_edges;
  }
  MMB_CHECK_CUDA(cudaFreeAsync(d_state, st));
  MMB_CHECK_CUDA(cudaFreeAsync(d_counts, st));
  MMB_CHECK_CUDA(cudaStreamSynchronize(st));
  if (n_edges > edge_cap) {
    set_error("kill-edge buffer overflow (%d > %d)", n_edges, edge_cap);
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
