// Fused x -> y sweep: C = g_y * (g_x * I), D = h_y * (g_x * I) + g_y * (h_x * I) in ONE kernel.
//
// The separate sweeps (log_x.cu: A = g*I, B = h*I; log_kernels.cuh conv_march_kernel<MID>:
// C = g*A, D = h*A + g*B) move A and B through HBM: 8 B/voxel written and 8 B/voxel read
// back per scale, 16 of the 40 B/voxel of a LoG scale, and each of the two kernels leaves the
// FP32 pipe idle while it waits on that traffic.  Here a CTA owns 128 x-columns of one
// z-plane and MARCHES along y in steps of 32 rows, alternating two phases over shared memory:
//
//   x phase  (the tile arithmetic of conv_x_tma_kernel): the 32-row x (128 + 2 R4)-column
//            window of F arrives by TMA (32 x 32-float boxes, 128-byte swizzle, zero fill
//            outside the plane, scipy 'reflect' patched in at the x faces); thread (lane =
//            row, warp = 16-output segment) copies its window to registers, scatters every value into
//            the outputs it touches (FFMA2 on the (A, B) pair with the (g, h) weight pair as
//            the uniform operand) and writes A and B rows into a shared-memory RING;
//   y phase  (the arithmetic of conv_march_kernel<MID>): thread (column pair, row group)
//            computes 8 rows x 2 columns of C and D from the ring rows a - R .. a + 7 + R
//            (FFMA2 on the column pair with the tap weight as the broadcast scalar) and
//            stores them to global memory.
//
// A and B never leave the SM: the kernel reads 4 B/voxel (plus the x halo, an L2 hit) and
// writes 8, and is bound by the FP32 pipe alone ((10 r + 5) lane-FMAs per voxel).
//
// Ring: RR = 2 RP + 32 rows of A and of B (RP = R rounded up to 4), row pitch 512 B, 16-byte
// chunks XOR-swizzled with (slot & 7) so that both the x phase's stores (lane = row) and the
// y phase's loads (lane = column pair) are bank-conflict free.  Slot of row q is
// (q - (a - RP)) mod RR, a = first output row of the CTA's segment.  The y phase's windows
// start at multiples of 8 slots, so an 8-row block never straddles the ring end and its
// swizzle key is a compile-time constant.  Rows outside [0, Y) are 'reflect' copies of rows
// inside, made in shared memory after the x phase that produced their sources.
#include <stdlib.h>
#include "log_kernels.cuh"
#include "tma.cuh"

namespace mmb {

// x columns per CTA: 64 columns = 128 threads and ~45 KB of shared memory, i.e. four CTAs per
// SM whose barriers are independent of one another; with 128 columns (two CTAs of 256
// threads) the FP32 pipe idles at every phase change (measured: 0.84 vs 0.xx ms at r = 16)
#ifndef MMB_XY_COLS
#define MMB_XY_COLS 64
#endif
constexpr int kFCols = MMB_XY_COLS;
constexpr int kFRows = 32;             // rows per step (= lanes of the x phase)
constexpr int kFThreads = 2 * kFCols;  // x phase: 32 lanes x kFCols/16 segments; y phase: kFCols/2 pairs x 4 groups
constexpr int kFPairs = kFCols / 2;
constexpr int kFChunks = kFCols / 4;   // 16-byte chunks per ring row
constexpr int kFBox = 32 * 32 * 4;     // bytes of one TMA box
constexpr int kFRowB = kFCols * 4;     // ring row pitch in bytes

// byte offset of window position p of row rr inside the TMA stage (128-byte swizzle)
__device__ __forceinline__ int x_sw_off_f(int rr, int p) {
  const int col = p & 31;
  return (p >> 5) * kFBox + rr * 128 + ((((col >> 2) ^ (rr & 7)) << 4) | ((col & 3) << 2));
}

// radius buckets of the fused sweep (a request uses the smallest bucket >= r)
#ifdef MMB_DEV_BUCKETS
#define MMB_XY_BUCKETS(X) X(12) X(16) X(20)
#else
#define MMB_XY_BUCKETS(X) X(8) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) X(20)
#endif

template <int R>
struct FGeom {
  static constexpr int R4 = (R + 3) / 4 * 4;
  // ring halo: R rounded up to 4 keeps RR and the y window (8 + 2 RP rows) multiples of 8
  static constexpr int RP = (R + 3) / 4 * 4;
  static constexpr int WIN = kFCols + 2 * R4;          // x window of a tile, floats
  static constexpr int NBOX = (WIN + 31) / 32;
  static constexpr int W = 16 + 2 * R4;                // per-thread x window, floats
  static constexpr int STAGE = NBOX * kFBox;
  static constexpr int RR = 2 * RP + kFRows;           // ring rows
  static constexpr int RINGB = RR * kFRowB;            // bytes of one ring (A or B)
  static constexpr int P = (2 * RP + 31) / 32;         // prologue tiles
  static constexpr int NB = 8;                         // y outputs per thread
  static constexpr int NBLK = (NB + 2 * RP) / 8;       // 8-row blocks in a y window
  static constexpr size_t SMEM = (size_t)STAGE + 2 * RINGB + 64 + 1024;
};

// y phase: one ring row pair scattered into the 8 accumulators it touches (MODE_MID)
template <int R, int RP, int NBLK, int RINGB, int K>
__device__ __forceinline__ void xy_scatter(const unsigned char* const (&bp)[NBLK],
                                           const uint32_t (&xo)[8], float2 (&acc0)[8],
                                           float2 (&acc1)[8], const LogWeights& w) {
  if constexpr (K < 8 + 2 * RP) {
    constexpr int KT = K - (RP - R);           // row index relative to a_out - R
    if constexpr (KT >= 0 && KT < 8 + 2 * R) {
      // block base + swizzled column offset, then compile-time offsets: one add per row
      const unsigned char* p = bp[K / 8] + xo[K & 7];
      const float2 v0 = *reinterpret_cast<const float2*>(p + (K & 7) * kFRowB);
      const float2 v1 = *reinterpret_cast<const float2*>(p + (K & 7) * kFRowB + RINGB);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t = KT - j;
        if (t >= 0 && t <= 2 * R) {
          const int wi = t >= R ? t - R : R - t;
          const float2 g2 = make_float2(w.g[wi], w.g[wi]);
          const float2 h2 = make_float2(w.h[wi], w.h[wi]);
          acc0[j] = ffma2(v0, g2, acc0[j]);
          acc1[j] = ffma2(v0, h2, acc1[j]);
          acc1[j] = ffma2(v1, g2, acc1[j]);
        }
      }
    }
    xy_scatter<R, RP, NBLK, RINGB, K + 1>(bp, xo, acc0, acc1, w);
  }
}

// grid = (ceil(pitch / kFCols), segments of the y axis, Z); block = 2 kFCols; 512 threads per SM
template <int R>
__global__ void __launch_bounds__(kFThreads, 512 / kFThreads)
conv_xy_fused_kernel(const __grid_constant__ CUtensorMap tm_in, const float* __restrict__ in,
                     float* __restrict__ outC, float* __restrict__ outD, int Y, int X,
                     int64_t pitch, int seg_len, const __grid_constant__ LogWeights w) {
  using G = FGeom<R>;
  constexpr int RP = G::RP, RR = G::RR, R4 = G::R4;
  extern __shared__ unsigned char smem_raw[];
  // aligned up on the 32-bit SHARED address, as an offset from smem_raw: a round trip through
  // uintptr_t would make every later access a generic LD / ST with 64-bit address arithmetic
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* s_in = base;                           // one TMA stage (1024-aligned)
  unsigned char* ringA = base + G::STAGE;
  uint64_t* bar = reinterpret_cast<uint64_t*>(ringA + 2 * G::RINGB);

  const int tid = threadIdx.x;
  const int lane = tid & 31, seg = tid >> 5;
  const int key = (lane & 7) << 4;
  const int x0 = blockIdx.x * kFCols;
  const int z = blockIdx.z;
  const int a0 = blockIdx.y * seg_len;                  // first output row of this segment
  const int a_end = min(Y, a0 + seg_len);
  const int nsteps = (a_end - a0 + kFRows - 1) / kFRows;
  const int ntiles = G::P + nsteps;
  const int row_first = a0 - RP;                        // lowest row the ring ever holds
  const int org0 = a0 + RP - kFRows * G::P;             // first row of tile 0
  const int xs = x0 - R4;                               // global x of window position 0
  const int ncols = (int)((pitch - x0) < kFCols ? (pitch - x0) : kFCols);   // multiple of 4
  const int64_t plane = (int64_t)z * Y * pitch;

  auto issue_load = [&](int j) {                        // one thread
    mbar_arrive_expect_tx(bar, G::STAGE);
#pragma unroll
    for (int b = 0; b < G::NBOX; ++b)
      tma_load_3d(s_in + b * kFBox, &tm_in, bar, xs + 32 * b, org0 + kFRows * j, z);
  };
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) issue_load(0);

  // 'reflect' rows: ring row q (outside [0, Y)) := ring row reflect(q), rows [q_lo, q_hi)
  auto mirror_rows = [&](int q_lo, int q_hi) {
    const int nq = q_hi - q_lo;
    for (int i = tid; i < nq * 2 * kFChunks; i += kFThreads) {
      const int q = q_lo + i / (2 * kFChunks), c = i % (2 * kFChunks);   // chunks of A, then of B
      const int src = reflect_index(q, Y);
      int sd = (q - row_first) % RR, ss = (src - row_first) % RR;
      const uint32_t arr = (c / kFChunks) * G::RINGB;
      const int ch = c % kFChunks;
      const float4 v = *reinterpret_cast<const float4*>(
          ringA + arr + ss * kFRowB + ((ch ^ (ss & 7)) << 4));
      *reinterpret_cast<float4*>(ringA + arr + sd * kFRowB + ((ch ^ (sd & 7)) << 4)) = v;
    }
  };

  // x faces: window positions outside [0, X) that feed valid outputs (CTA-uniform)
  const int n_left = xs < 0 ? -xs : 0;
  const int pr0 = X - xs;                               // first window position with x >= X
  const int pr1 = min(G::WIN, X + R - xs);
  const int n_right = pr1 > pr0 ? pr1 - pr0 : 0;
  const int n_fix = n_left + n_right;
  const bool x_active = seg * 16 < ncols;               // this segment's 16 columns exist

  // x phase of tile j: rows [org0 + 32 j, + 32) of A and B into the ring
  auto x_phase = [&](int j) {
    const int org = org0 + kFRows * j;
    mbar_wait(bar, (uint32_t)j & 1u);
    if (n_fix > 0) {
      for (int i = tid; i < kFRows * n_fix; i += kFThreads) {
        const int rr = i / n_fix, q = i - rr * n_fix;
        const int p = q < n_left ? q : pr0 + (q - n_left);
        const int src = reflect_index(xs + p, X) - xs;
        float v = 0.f;
        if (src >= 0 && src < G::WIN) {
          v = *reinterpret_cast<const float*>(s_in + x_sw_off_f(rr, src));
        } else if (org + rr >= 0 && org + rr < Y) {     // volume narrower than the window
          v = __ldcg(in + plane + (int64_t)(org + rr) * pitch + (xs + src));
        }
        *reinterpret_cast<float*>(s_in + x_sw_off_f(rr, p)) = v;
      }
      __syncthreads();
    }
    float win[G::W];
    if (x_active) {
      const unsigned char* rowp = s_in + lane * 128;
#pragma unroll
      for (int c = 0; c < G::W / 4; ++c) {
        const int ci = seg * 4 + c;
        const float4 v = *reinterpret_cast<const float4*>(rowp + (ci >> 3) * kFBox +
                                                          (((ci & 7) << 4) ^ key));
        win[4 * c + 0] = v.x; win[4 * c + 1] = v.y; win[4 * c + 2] = v.z; win[4 * c + 3] = v.w;
      }
    }
    // every thread has finished the previous y phase: the ring rows this tile replaces are
    // free.  The TMA load of the NEXT tile into the (single) stage is NOT issued here: a
    // barrier orders the window reads above only among the threads, not against the async
    // proxy, and a read still in flight when the bulk copy lands would return the next
    // tile's rows (seen as run-to-run differences of a few blobs per chunk whenever other
    // work shared the SMs: tools/concurrency_probe.py).  It is issued after the barrier that
    // follows the arithmetic, by which every window value has been consumed.
    __syncthreads();
    const int row = org + lane;
    if (x_active && row >= row_first) {
      float2 acc[16];
      x_taps_sym<R, R4, G::W>(win, acc, w);
      const int slot = (row - row_first) % RR;
      unsigned char* ra = ringA + slot * kFRowB;
      const int skey = slot & 7;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int off = ((seg * 4 + q) ^ skey) << 4;
        *reinterpret_cast<float4*>(ra + off) =
            make_float4(acc[4 * q].x, acc[4 * q + 1].x, acc[4 * q + 2].x, acc[4 * q + 3].x);
        *reinterpret_cast<float4*>(ra + G::RINGB + off) =
            make_float4(acc[4 * q].y, acc[4 * q + 1].y, acc[4 * q + 2].y, acc[4 * q + 3].y);
      }
    }
  };

  const int grp = tid / kFPairs;                         // 4 row groups of 8 outputs
  const int col = (tid % kFPairs) * 2;                   // column pair
  uint32_t xo[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) xo[k] = (uint32_t)((((col >> 2) ^ k) << 4) + ((col & 3) << 2));
  const bool y_active = col < ncols;

  // tiles 0 .. P-1 are the prologue (rows [a0 - RP, a0 + RP)); tile P + s brings the rows
  // [a0 + RP + 32 s, + 32) that step s (outputs [a0 + 32 s, + 32)) still lacks
#pragma unroll 1
  for (int j = 0; j < ntiles; ++j) {
    x_phase(j);
    const int s = j - G::P;
    if (s < 0) {
      // prologue tile: no y phase follows, so the barrier that proves every window value
      // consumed is this one
      __syncthreads();
      if (tid == 0 && j + 1 < ntiles) { fence_proxy_async(); issue_load(j + 1); }
    }
    if (j == G::P - 1) {
      // 'reflect' rows the prologue covers: above the first row, and - for a segment that
      // starts within RP rows of the end - below the last one
      if (row_first < 0 || a0 + RP > Y) {
        if (row_first < 0) mirror_rows(row_first, min(0, a0 + RP));
        if (a0 + RP > Y) mirror_rows(Y, min(a0 + RP, Y + R));
      }
      continue;
    }
    if (s < 0) continue;
    const int org = a0 + RP + kFRows * s;
    if (org + kFRows > Y) {                              // rows past the last one: reflect
      __syncthreads();
      const int q_lo = max(org, Y), q_hi = min(org + kFRows, Y + R);
      if (q_hi > q_lo) mirror_rows(q_lo, q_hi);
    }
    __syncthreads();                                     // the ring rows of this step are in
    // ... and every window of this tile has been consumed: the stage may be refilled
    if (tid == 0 && j + 1 < ntiles) { fence_proxy_async(); issue_load(j + 1); }
    const int a_out = a0 + kFRows * s + grp * 8;
    if (y_active && a_out < a_end) {
      float2 acc0[8], acc1[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) { acc0[jj] = make_float2(0.f, 0.f); acc1[jj] = make_float2(0.f, 0.f); }
      // window row 0 = row a_out - RP = slot (32 s + 8 grp) mod RR, a multiple of 8
      const int cslot = (kFRows * s + grp * 8) % RR;
      const unsigned char* bp[G::NBLK];
#pragma unroll
      for (int i = 0; i < G::NBLK; ++i) {
        int sl = cslot + 8 * i;
        if (sl >= RR) sl -= RR;
        bp[i] = ringA + sl * kFRowB;
      }
      xy_scatter<R, RP, G::NBLK, G::RINGB, 0>(bp, xo, acc0, acc1, w);
      float* o0 = outC + plane + (int64_t)a_out * pitch + x0 + col;
      float* o1 = outD + plane + (int64_t)a_out * pitch + x0 + col;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        if (a_out + jj < a_end) {
          *reinterpret_cast<float2*>(o0) = acc0[jj];
          *reinterpret_cast<float2*>(o1) = acc1[jj];
        }
        o0 += pitch;
        o1 += pitch;
      }
    }
  }
}

template <int R>
static int run_xy(const float* in, float* outC, float* outD, int Z, int Y, int X, int64_t pitch,
                  const LogWeights& w, cudaStream_t st) {
  using G = FGeom<R>;
  ProfScope ps(PROF_LOG_XY, (double)Z * Y * pitch, st);
  auto kern = conv_xy_fused_kernel<R>;
  static bool configured = false;
  if (!configured) {
    MMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)G::SMEM));
    configured = true;
  }
  CUtensorMap tin;
  const uint64_t dims[3] = {(uint64_t)X, (uint64_t)Y, (uint64_t)Z};
  const uint64_t strides[2] = {(uint64_t)pitch * sizeof(float),
                               (uint64_t)Y * (uint64_t)pitch * sizeof(float)};
  const uint32_t box[3] = {32, 32, 1};
  if (encode_tensor_map_f32(&tin, in, 3, dims, strides, box, true)) return MMB_ERR_CUDA;
  // split the marched axis when column strips x planes alone cannot fill the GPU
  const int64_t ctas = cdiv(pitch, kFCols) * Z;
  const int64_t want = (512 / kFThreads) * (int64_t)num_sms();
  int nseg = 1;
  if (ctas < want) {
    // as many segments as still fit ONE wave (a second, partly filled wave costs a whole
    // march), of at least two steps each: a segment pays one or two prologue tiles
    nseg = (int)(want / ctas);
    const int max_seg = (int)cdiv(Y, 2 * kFRows);
    if (nseg > max_seg) nseg = max_seg;
    if (nseg < 1) nseg = 1;
  }
  const int seg_len = (int)cdiv(cdiv(Y, nseg), kFRows) * kFRows;
  dim3 grid((unsigned)cdiv(pitch, kFCols), (unsigned)cdiv(Y, seg_len), (unsigned)Z);
  kern<<<grid, kFThreads, G::SMEM, st>>>(tin, in, outC, outD, Y, X, pitch, seg_len, w);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

// the fused sweep serves radii up to 20 on volumes whose rows are 16-byte aligned and at least
// 32 voxels long in x and y (the 'reflect' sources of every ring row then lie inside the ring); everything else takes the two
// separate sweeps.  Returns MMB_ERR_UNSUPPORTED when the caller should do that.
int launch_xy_fused(int r, const float* in, float* outC, float* outD, int Z, int Y, int X,
                    int64_t pitch, const LogWeights& w, cudaStream_t st) {
  const uintptr_t bits = (uintptr_t)in | (uintptr_t)outC | (uintptr_t)outD;
  // measured on a 505^3 chunk (profiles/r02_kbench.jsonl), fused vs x + y sweeps: 0.58 vs 0.71 ms
  // at r = 12, 0.74 vs 0.82 at r = 16, 0.93 vs 1.06 at r = 20.  MMB_XY_RMAX lowers the limit
  // for A/B runs.
  static const int r_max = getenv("MMB_XY_RMAX") ? atoi(getenv("MMB_XY_RMAX")) : 20;
  if (r > r_max || r > 20 || (bits & 15) != 0 || pitch % 4 != 0 || X < 32 || Y < 32 || Z > 65535)
    return MMB_ERR_UNSUPPORTED;
#define XY_(RR) if (r <= RR) return run_xy<RR>(in, outC, outD, Z, Y, X, pitch, w, st);
  MMB_XY_BUCKETS(XY_)
#undef XY_
  return MMB_ERR_UNSUPPORTED;
}

}  // namespace mmb
