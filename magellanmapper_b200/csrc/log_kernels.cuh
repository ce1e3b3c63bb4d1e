// Separable Gaussian / Gaussian-second-derivative passes of the LoG cube.
//
// scipy.ndimage.gaussian_laplace(img, s) = sum over axes a of
// gaussian_filter(img, s, order=2 on a, 0 elsewhere) - nine correlate1d sweeps
// per scale (scipy/ndimage/_filters.py:1120-1139, 845-848).  With g the sampled
// Gaussian and h its second derivative the same sum factors into seven 1-D
// convolutions over three sweeps:
//     x:  A = g*I          B = h*I                 (MODE_FIRST)
//     y:  C = g*A          D = h*A + g*B           (MODE_MID)
//     z:  out = scale * (h*C + g*D)                (MODE_LAST, scale = -sigma^2)
// Each sweep is FP32-issue bound (about 2r+1 FMAs per convolution per voxel,
// r = 12..20 for sigma 3..5), so every thread owns NB consecutive outputs along
// the filtered axis and streams NB+2r inputs through registers once, scattering
// each into the accumulators it touches.  Tap weights are read from the kernel
// parameter block, i.e. constant-bank operands of the FMAs.
#pragma once
#include <cuda_pipeline.h>
#include "common.cuh"

namespace mmb {

enum { MODE_FIRST = 0, MODE_MID = 1, MODE_LAST = 2 };

// marching sweep shape: NB outputs per thread, G row groups per CTA (STEP = NB * G)
#ifndef MMB_MARCH_NB
#define MMB_MARCH_NB 8
#endif
#ifndef MMB_MARCH_G
#define MMB_MARCH_G 4
#endif

// Packed FP32 FMA of sm_100 (SASS FFMA2): two independent IEEE fp32 fmas per lane
// in one issue slot.  The sweeps are FP32-issue bound, so every inner loop below
// is written on float2 values; ptxas folds a (w, w) or (v, v) operand into the
// scalar-broadcast form (`UR.F32` / `R.F32`), so no register is spent on the copy.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  return __ffma2_rn(a, b, c);
}

// The x taps of 16 consecutive outputs held by one thread: win[R4 + j] is the input at output
// j, acc[j] = (A_j, B_j).  Symmetric-pair form, as scipy's correlate1d evaluates a symmetric
// kernel: acc_j = w[0] v_j, then for t = 1..R  acc_j += w[t] (v_{j-t} + v_{j+t}).  The pair
// sums of two neighbouring outputs are one packed FADD2 on register pairs of the window that
// start at an even index (even taps pair the outputs (0,1), (2,3), ..; odd taps (1,2), (3,4),
// .. with outputs 0 and 15 on scalar adds), and each sum feeds one FFMA2 on the (A, B) pair
// with the (g, h) weight pair as the uniform operand: (3R + 2) pipe slots per output instead
// of the (4R + 2) of one FFMA2 per tap and side.  Measured (tools/xform_bench.cu, r = 16):
// 2742 vs 3475 cycles per 16 outputs.  EVERY x kernel uses this order, so their results are
// bit-identical.
template <int R, int R4, int WN>
__device__ __forceinline__ void x_taps_sym(const float (&win)[WN], float2 (&acc)[16],
                                           const LogWeights& w) {
  static_assert(R4 % 2 == 0 && R4 >= R && WN >= R4 + 16 + R, "window geometry");
#pragma unroll
  for (int j = 0; j < 16; ++j)
    acc[j] = ffma2(make_float2(win[R4 + j], win[R4 + j]), w.gh[0], make_float2(0.f, 0.f));
#pragma unroll
  for (int t = 1; t <= R; ++t) {
    const int j0 = t & 1;
    if (j0) {
      const float s0 = win[R4 - t] + win[R4 + t];
      acc[0] = ffma2(make_float2(s0, s0), w.gh[t], acc[0]);
      const float s15 = win[R4 + 15 - t] + win[R4 + 15 + t];
      acc[15] = ffma2(make_float2(s15, s15), w.gh[t], acc[15]);
    }
#pragma unroll
    for (int j = j0; j + 1 < 16; j += 2) {
      const float2 s = __fadd2_rn(make_float2(win[R4 + j - t], win[R4 + j - t + 1]),
                                  make_float2(win[R4 + j + t], win[R4 + j + t + 1]));
      acc[j] = ffma2(make_float2(s.x, s.x), w.gh[t], acc[j]);
      acc[j + 1] = ffma2(make_float2(s.y, s.y), w.gh[t], acc[j + 1]);
    }
  }
}

// Convolution along a strided axis (y or z).  The volume is viewed as
// [outer][n_axis][inner] with `inner` contiguous: (Z, Y, pitch) for the y sweep,
// (1, Z, Y*pitch) for the z sweep.  Thread = one inner position, NB outputs.
// grid = (ceil(inner / threads), ceil(n_axis / NB), outer).
template <int R, int MODE, int NB, int THREADS>
__global__ void __launch_bounds__(THREADS)
conv_strided_kernel(const float* __restrict__ in0, const float* __restrict__ in1,
                    float* __restrict__ out0, float* __restrict__ out1, int n_axis,
                    int64_t inner, int64_t outer_stride,
                    const __grid_constant__ LogWeights w, float scale) {
  constexpr int NIN = NB + 2 * R;
  __shared__ int64_t s_off[NIN];
  const int a0 = blockIdx.y * NB;
  for (int k = threadIdx.x; k < NIN; k += THREADS)
    s_off[k] = (int64_t)reflect_index(a0 - R + k, n_axis) * inner;
  __syncthreads();
  const int64_t xi = (int64_t)blockIdx.x * THREADS + threadIdx.x;
  if (xi >= inner) return;
  const int64_t base = (int64_t)blockIdx.z * outer_stride + xi;
  const float* p0 = in0 + base;
  const float* p1 = (MODE == MODE_FIRST) ? nullptr : in1 + base;

  float acc0[NB], acc1[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }

#pragma unroll
  for (int k = 0; k < NIN; ++k) {
    const int64_t off = s_off[k];
    const float v0 = __ldcg(p0 + off);
    const float v1 = (MODE == MODE_FIRST) ? 0.f : __ldcg(p1 + off);
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int t = k - j;
      if (t >= 0 && t <= 2 * R) {
        const int wi = t >= R ? t - R : R - t;
        if (MODE == MODE_FIRST) {
          acc0[j] = fmaf(w.g[wi], v0, acc0[j]);
          acc1[j] = fmaf(w.h[wi], v0, acc1[j]);
        } else if (MODE == MODE_MID) {
          acc0[j] = fmaf(w.g[wi], v0, acc0[j]);
          acc1[j] = fmaf(w.h[wi], v0, acc1[j]);
          acc1[j] = fmaf(w.g[wi], v1, acc1[j]);
        } else {
          acc0[j] = fmaf(w.h[wi], v0, acc0[j]);
          acc0[j] = fmaf(w.g[wi], v1, acc0[j]);
        }
      }
    }
  }

#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int a = a0 + j;
    if (a < n_axis) {
      const int64_t o = base + (int64_t)a * inner;
      if (MODE == MODE_LAST) {
        out0[o] = acc0[j] * scale;
      } else {
        out0[o] = acc0[j];
        out1[o] = acc1[j];
      }
    }
  }
}

// One input row pair (v0, v1) of the strided sweeps scattered into the NB
// accumulators it touches, then the next row: template recursion instead of a
// loop so that every tap index is a compile-time constant whatever the unroller's
// size heuristics decide.
template <int R, int MODE, int NB, int COLS, int K>
__device__ __forceinline__ void scatter_rows(const float* __restrict__ p0,
                                             const float* __restrict__ p1, float2 (&acc0)[NB],
                                             float2 (&acc1)[NB], const LogWeights& w) {
  if constexpr (K < NB + 2 * R) {
    const float2 v0 = *reinterpret_cast<const float2*>(p0 + K * COLS);
    float2 v1 = make_float2(0.f, 0.f);
    if (MODE != MODE_FIRST) v1 = *reinterpret_cast<const float2*>(p1 + K * COLS);
    // three passes over the NB accumulators so that the two FFMA2s that update the
    // same accumulator (h*v0 then g*v1) are NB issue slots apart, not back to back
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int t = K - j;
        if (t >= 0 && t <= 2 * R) {
          const int wi = t >= R ? t - R : R - t;
          const float2 g2 = make_float2(w.g[wi], w.g[wi]);
          const float2 h2 = make_float2(w.h[wi], w.h[wi]);
          if (MODE == MODE_FIRST) {
            if (pass == 0) acc0[j] = ffma2(v0, g2, acc0[j]);
            if (pass == 1) acc1[j] = ffma2(v0, h2, acc1[j]);
          } else if (MODE == MODE_MID) {
            if (pass == 0) acc1[j] = ffma2(v0, h2, acc1[j]);
            if (pass == 1) acc0[j] = ffma2(v0, g2, acc0[j]);
            if (pass == 2) acc1[j] = ffma2(v1, g2, acc1[j]);
          } else {
            if (pass == 0) acc0[j] = ffma2(v0, h2, acc0[j]);
            if (pass == 2) acc0[j] = ffma2(v1, g2, acc0[j]);
          }
        }
      }
    }
    scatter_rows<R, MODE, NB, COLS, K + 1>(p0, p1, acc0, acc1, w);
  }
}


// Tiled variant of the strided sweep (the default when rows are 16-byte aligned).
// A CTA stages a tall input tile - (NB*NSEGS + 2R) rows x COLS columns of each
// input - in shared memory with 16-byte cp.async copies (all in flight at once),
// then every thread runs the register-blocked scatter over NB outputs of TWO
// adjacent columns at once: inputs are 8-byte shared loads at compile-time
// offsets, the arithmetic is packed FFMA2 with the tap weight as the broadcast
// scalar operand (a uniform register), i.e. half the FP32 issue slots of the
// scalar form.  Compared with reading global memory directly the tile cuts the
// L2->SM traffic from (NB+2R)/NB to (NB*NSEGS+2R)/(NB*NSEGS) times the input.
// grid = (ceil(inner / COLS), ceil(n_axis / (NB*NSEGS)), outer), block = THREADS.
template <int R, int MODE, int NB, int NSEGS, int COLS, int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
conv_strided_tile_kernel(const float* __restrict__ in0, const float* __restrict__ in1,
                         float* __restrict__ out0, float* __restrict__ out1, int n_axis,
                         int64_t inner, int64_t outer_stride,
                         const __grid_constant__ LogWeights w, float scale) {
  constexpr int NBT = NB * NSEGS;
  constexpr int ROWS = NBT + 2 * R;
  constexpr int CH = COLS / 4;                 // 16-byte chunks per row
  constexpr int PAIRS = COLS / 2;              // column pairs per tile
  constexpr int GROUPS = THREADS / PAIRS;      // threads sharing a column pair
  static_assert(THREADS % PAIRS == 0, "THREADS must be a multiple of COLS / 2");
  extern __shared__ __align__(16) float tile[];
  float* t0 = tile;
  float* t1 = tile + ROWS * COLS;
  __shared__ int64_t s_off[ROWS];
  const int a0 = blockIdx.y * NBT;
  const int64_t c0 = (int64_t)blockIdx.x * COLS;
  for (int k = threadIdx.x; k < ROWS; k += THREADS)
    s_off[k] = (int64_t)reflect_index(a0 - R + k, n_axis) * inner;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.z * outer_stride + c0;
  const int ncols = (int)((inner - c0) < COLS ? (inner - c0) : COLS);   // multiple of 4
  for (int i = threadIdx.x; i < ROWS * CH; i += THREADS) {
    const int row = i / CH, ch = i - row * CH;
    if (ch * 4 < ncols) {
      const int64_t g = base + s_off[row] + ch * 4;
      __pipeline_memcpy_async(t0 + row * COLS + ch * 4, in0 + g, 16);
      if (MODE != MODE_FIRST) __pipeline_memcpy_async(t1 + row * COLS + ch * 4, in1 + g, 16);
    }
  }
  __pipeline_commit();
  __pipeline_wait_prior(0);
  __syncthreads();

  const int col = (threadIdx.x % PAIRS) * 2;
  if (col >= ncols) return;
  for (int seg = threadIdx.x / PAIRS; seg < NSEGS; seg += GROUPS) {
    if (a0 + seg * NB >= n_axis) break;
    float2 acc0[NB], acc1[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) { acc0[j] = make_float2(0.f, 0.f); acc1[j] = make_float2(0.f, 0.f); }
    const float* p0 = t0 + seg * NB * COLS + col;
    const float* p1 = t1 + seg * NB * COLS + col;
    scatter_rows<R, MODE, NB, COLS, 0>(p0, p1, acc0, acc1, w);
    const int64_t ob = base + col;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int a = a0 + seg * NB + j;
      if (a < n_axis) {
        const int64_t o = ob + (int64_t)a * inner;
        if (MODE == MODE_LAST) {
          *reinterpret_cast<float2*>(out0 + o) = make_float2(acc0[j].x * scale, acc0[j].y * scale);
        } else {
          *reinterpret_cast<float2*>(out0 + o) = acc0[j];
          *reinterpret_cast<float2*>(out1 + o) = acc1[j];
        }
      }
    }
  }
}

// Marching variant of the strided sweep (the default for radii <= 20): a CTA owns
// COLS columns of one outer slice and walks the whole filtered axis in steps of
// STEP = NB*G outputs, keeping the input rows it needs in a shared-memory ring.
// Every input row is read from L2/HBM exactly once (the tile kernel re-reads its 2R
// halo rows per tile), and the cp.async copies of the NEXT step's rows are in
// flight while the current step is computed, so the FMA pipe never waits on a tile
// load.  One __syncthreads per step.
//
// Ring geometry is kept in units of 8 rows: RP = R rounded up to 8, RING = 2 RP +
// 2 STEP rows, slot of global row a = (a + RP) mod RING.  A thread's window starts
// at a multiple of 8, so each block of 8 window rows wraps as a whole and the wrap
// costs one select per 8 rows (the shared loads use immediate offsets from one of
// two base pointers).  Rows between R and RP are loaded but carry no taps.
// grid = (ceil(inner / COLS), 1, outer), block = G * COLS / 2, 2 CTAs per SM.
template <int R, int RP, int MODE, int NB, int COLS, int NBLK, int K>
__device__ __forceinline__ void scatter_rows_ring(const float* const (&b0)[NBLK],
                                                  const float* const (&b1)[NBLK],
                                                  float2 (&acc0)[NB], float2 (&acc1)[NB],
                                                  const LogWeights& w) {
  // K counts window rows from a_out - RP; taps exist for rows a_out - R .. a_out + NB - 1 + R
  if constexpr (K < NB + 2 * RP) {
    constexpr int KT = K - (RP - R);           // row index relative to a_out - R
    if constexpr (KT >= 0 && KT < NB + 2 * R) {
      // block base pointer + compile-time offset: one LDS.64 with an immediate
      const float2 v0 = *reinterpret_cast<const float2*>(b0[K / 8] + K * COLS);
      float2 v1 = make_float2(0.f, 0.f);
      if (MODE != MODE_FIRST) v1 = *reinterpret_cast<const float2*>(b1[K / 8] + K * COLS);
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int t = KT - j;
        if (t >= 0 && t <= 2 * R) {
          const int wi = t >= R ? t - R : R - t;
          const float2 g2 = make_float2(w.g[wi], w.g[wi]);
          const float2 h2 = make_float2(w.h[wi], w.h[wi]);
          if (MODE == MODE_FIRST) {
            acc0[j] = ffma2(v0, g2, acc0[j]);
            acc1[j] = ffma2(v0, h2, acc1[j]);
          } else if (MODE == MODE_MID) {
            acc0[j] = ffma2(v0, g2, acc0[j]);
            acc1[j] = ffma2(v0, h2, acc1[j]);
            acc1[j] = ffma2(v1, g2, acc1[j]);
          } else {
            acc0[j] = ffma2(v0, h2, acc0[j]);
            acc0[j] = ffma2(v1, g2, acc0[j]);
          }
        }
      }
    }
    scatter_rows_ring<R, RP, MODE, NB, COLS, NBLK, K + 1>(b0, b1, acc0, acc1, w);
  }
}

template <int R, int MODE, int NB, int G, int COLS>
__global__ void __launch_bounds__(G * COLS / 2, 512 / (G * COLS / 2))
conv_march_kernel(const float* __restrict__ in0, const float* __restrict__ in1,
                  float* __restrict__ out0, float* __restrict__ out1, int n_axis,
                  int64_t inner, int64_t outer_stride, const __grid_constant__ LogWeights w,
                  float scale, int seg_len) {
  constexpr int PAIRS = COLS / 2;
  constexpr int THREADS = G * PAIRS;
  constexpr int STEP = NB * G;
  constexpr int RP = (R + 7) / 8 * 8;
  constexpr int RING = 2 * RP + 2 * STEP;
  constexpr int NBLK = (NB + 2 * RP) / 8;      // 8-row blocks in a thread's window
  constexpr int CH = COLS / 4;                 // 16-byte chunks per row
  constexpr int RPT = THREADS / CH;            // rows covered by one pass of the CTA
  static_assert(THREADS % CH == 0 && STEP % RPT == 0 && (2 * RP) % RPT == 0, "load geometry");
  static_assert(NB % 8 == 0 && STEP % 8 == 0, "ring geometry is in units of 8 rows");
  extern __shared__ __align__(16) float ring[];
  float* r0 = ring;
  float* r1 = ring + RING * COLS;
  const int tid = threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.x * COLS;
  const int64_t base = (int64_t)blockIdx.z * outer_stride + c0;
  const int ncols = (int)((inner - c0) < COLS ? (inner - c0) : COLS);   // multiple of 4
  const int lch = (tid % CH) * 4;              // this thread's 16-byte column chunk
  const bool lact = lch < ncols;

  // blockIdx.y splits the marched axis into segments of seg_len rows (a multiple of
  // STEP) when the other two grid dimensions alone would leave SMs idle; a segment
  // loads its own 2*RP halo rows.  Ring slots count rows from the segment's start.
  const int a_base = blockIdx.y * seg_len;
  const int a_end = min(n_axis, a_base + seg_len);
  // loader state: next row this thread copies, its ring slot, its global pointers
  int la = a_base - RP + tid / CH;
  int lslot = tid / CH;
  const int64_t lstride = (int64_t)RPT * inner;
  const float* lg0 = in0 + base + lch + (int64_t)la * inner;   // dereferenced only when 0 <= la < n_axis
  const float* lg1 = (MODE != MODE_FIRST ? in1 : in0) + base + lch + (int64_t)la * inner;

  // the next `count` rows (a multiple of RPT) of both inputs -> their ring slots
  auto load_rows = [&](int count) {
    const int a_first = la - tid / CH;                                   // CTA-uniform
    const bool interior = a_first >= 0 && a_first + count <= n_axis;
    if (lact) {
      if (interior) {
        for (int i = 0; i < count / RPT; ++i) {
          __pipeline_memcpy_async(r0 + lslot * COLS + lch, lg0, 16);
          if (MODE != MODE_FIRST) __pipeline_memcpy_async(r1 + lslot * COLS + lch, lg1, 16);
          lg0 += lstride;
          lg1 += lstride;
          lslot += RPT;
          if (lslot >= RING) lslot -= RING;
        }
      } else {
        for (int i = 0; i < count / RPT; ++i) {
          const int64_t g = base + lch + (int64_t)reflect_index(la + RPT * i, n_axis) * inner;
          __pipeline_memcpy_async(r0 + lslot * COLS + lch, in0 + g, 16);
          if (MODE != MODE_FIRST) __pipeline_memcpy_async(r1 + lslot * COLS + lch, in1 + g, 16);
          lslot += RPT;
          if (lslot >= RING) lslot -= RING;
        }
        lg0 += (int64_t)(count / RPT) * lstride;
        lg1 += (int64_t)(count / RPT) * lstride;
      }
    }
    la += count;
    __pipeline_commit();
  };

  load_rows(STEP + 2 * RP);
  const int nsteps = (a_end - a_base + STEP - 1) / STEP;
  const int grp = tid / PAIRS;
  const int col = (tid - grp * PAIRS) * 2;
  int a_out = a_base + grp * NB;
  int cslot = grp * NB;                        // ring slot of input row a_out - RP
  float* o0 = out0 + base + col + (int64_t)a_out * inner;
  float* o1 = (MODE != MODE_LAST ? out1 : out0) + base + col + (int64_t)a_out * inner;
  const int64_t ostep = (int64_t)(STEP - NB) * inner;
  for (int s = 0; s < nsteps; ++s) {
    __pipeline_wait_prior(0);
    __syncthreads();          // this step's rows have landed; step s-1's rows are free
    if (s + 1 < nsteps) load_rows(STEP);
    if (a_out < a_end && col < ncols) {
      float2 acc0[NB], acc1[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) { acc0[j] = make_float2(0.f, 0.f); acc1[j] = make_float2(0.f, 0.f); }
      // a window starts at a multiple of 8 rows, so each 8-row block wraps as a whole
      const float* b0[NBLK];
      const float* b1[NBLK];
#pragma unroll
      for (int i = 0; i < NBLK; ++i) {
        const int sl = cslot + 8 * i < RING ? cslot : cslot - RING;
        b0[i] = r0 + sl * COLS + col;
        b1[i] = r1 + sl * COLS + col;
      }
      scatter_rows_ring<R, RP, MODE, NB, COLS, NBLK, 0>(b0, b1, acc0, acc1, w);
      if (a_out + NB <= a_end) {               // full block: no per-row bound checks
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          if (MODE == MODE_LAST) {
            *reinterpret_cast<float2*>(o0) = make_float2(acc0[j].x * scale, acc0[j].y * scale);
          } else {
            *reinterpret_cast<float2*>(o0) = acc0[j];
            *reinterpret_cast<float2*>(o1) = acc1[j];
            o1 += inner;
          }
          o0 += inner;
        }
      } else {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          if (a_out + j < a_end) {
            if (MODE == MODE_LAST) {
              *reinterpret_cast<float2*>(o0) = make_float2(acc0[j].x * scale, acc0[j].y * scale);
            } else {
              *reinterpret_cast<float2*>(o0) = acc0[j];
              *reinterpret_cast<float2*>(o1) = acc1[j];
            }
          }
          o0 += inner;
          if (MODE != MODE_LAST) o1 += inner;
        }
      }
      o0 += ostep;
      if (MODE != MODE_LAST) o1 += ostep;
    }
    a_out += STEP;
    cslot += STEP;
    if (cslot >= RING) cslot -= RING;
  }
}

// First sweep along the contiguous x axis.  A CTA owns 32 rows x (NB*NSEG)
// outputs; the input tile (with its 2R halo) is staged in shared memory with an
// odd row pitch, so "lane = row" accesses are bank-conflict free; thread
// (lane = row, warp = segment) owns NB consecutive x outputs.  Results go back
// through the same shared tile so global stores are coalesced along x.
// grid = (ceil(nrows / 32), ceil(X / (NB*NSEG))).
template <int R, int NB, int NSEG>
__global__ void __launch_bounds__(32 * NSEG)
conv_x_first_kernel(const float* __restrict__ in, float* __restrict__ outA,
                    float* __restrict__ outB, int64_t nrows, int X, int64_t pitch,
                    const __grid_constant__ LogWeights w) {
  constexpr int WT = NB * NSEG;
  constexpr int WIN = WT + 2 * R;
  constexpr int SP = WIN | 1;       // odd pitch for the input tile
  constexpr int OP = WT | 1;        // odd pitch for the output staging tile
  __shared__ float s[32 * SP];
  const int lane = threadIdx.x & 31;
  const int seg = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int x0 = blockIdx.y * WT;
  const bool x_interior = (x0 - R >= 0) && (x0 + WT + R <= X);

  // Tile load: 4-byte cp.async per element so that every thread has all of its
  // loads in flight at once (a plain load/store loop here is latency bound).
  const bool full_rows = r0 + 32 <= nrows;
  if (x_interior && full_rows) {
    const float* src0 = in + r0 * pitch + (x0 - R);
    for (int i = threadIdx.x; i < 32 * WIN; i += 32 * NSEG) {
      const int rr = i / WIN, c = i - rr * WIN;
      __pipeline_memcpy_async(&s[rr * SP + c], src0 + (int64_t)rr * pitch + c, 4);
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
  } else {
    for (int i = threadIdx.x; i < 32 * WIN; i += 32 * NSEG) {
      const int rr = i / WIN, c = i - rr * WIN;
      const int64_t row = r0 + rr;
      float v = 0.f;
      if (row < nrows) v = __ldcg(in + row * pitch + reflect_index(x0 - R + c, X));
      s[rr * SP + c] = v;
    }
  }
  __syncthreads();

  // same order of operations as x_taps_sym (symmetric pairs, t ascending), read from the tile
  float accA[NB], accB[NB];
  const float* srow = s + lane * SP + seg * NB + R;      // srow[j] = input at output j
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const float v = srow[j];
    accA[j] = fmaf(v, w.g[0], 0.f);
    accB[j] = fmaf(v, w.h[0], 0.f);
  }
#pragma unroll 4
  for (int t = 1; t <= R; ++t) {
    const float g = w.g[t], h = w.h[t];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const float sum = srow[j - t] + srow[j + t];
      accA[j] = fmaf(sum, g, accA[j]);
      accB[j] = fmaf(sum, h, accB[j]);
    }
  }
  __syncthreads();

#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? outA : outB;
#pragma unroll
    for (int j = 0; j < NB; ++j)
      s[lane * OP + seg * NB + j] = pass == 0 ? accA[j] : accB[j];
    __syncthreads();
    for (int rr = seg; rr < 32; rr += NSEG) {
      const int64_t row = r0 + rr;
      if (row < nrows) {
        for (int c = lane; c < WT; c += 32) {
          const int x = x0 + c;
          if (x < X) dst[row * pitch + x] = s[rr * OP + c];
        }
      }
    }
    __syncthreads();
  }
}

// Slow generic fallbacks for radii above kMaxRadius (weights in global memory).
__global__ void conv_generic_kernel(const float* __restrict__ in0,
                                    const float* __restrict__ in1,
                                    float* __restrict__ out0, float* __restrict__ out1,
                                    int Z, int Y, int X, int64_t pitch, int axis, int mode,
                                    const float* __restrict__ g, const float* __restrict__ h,
                                    int r, float scale);

int launch_strided(int mode, int r, const float* in0, const float* in1, float* out0,
                   float* out1, int n_axis, int64_t inner, int64_t outer,
                   const LogWeights& w, float scale, cudaStream_t st);
int launch_strided_m0(int r, const float*, const float*, float*, float*, int, int64_t,
                      int64_t, const LogWeights&, float, cudaStream_t);
int launch_strided_m1(int r, const float*, const float*, float*, float*, int, int64_t,
                      int64_t, const LogWeights&, float, cudaStream_t);
int launch_strided_m2(int r, const float*, const float*, float*, float*, int, int64_t,
                      int64_t, const LogWeights&, float, cudaStream_t);
int launch_x_first(int r, const float* in, float* outA, float* outB, int64_t nrows, int X,
                   int64_t pitch, const LogWeights& w, cudaStream_t st);
// fused x -> y sweep (log_xy.cu); MMB_ERR_UNSUPPORTED = take the two separate sweeps
int launch_xy_fused(int r, const float* in, float* outC, float* outD, int Z, int Y, int X,
                    int64_t pitch, const LogWeights& w, cudaStream_t st);

// Radius buckets with a compiled kernel; a request is served by the smallest
// bucket >= r (taps beyond r carry zero weight).
#ifdef MMB_DEV_BUCKETS      /* quick developer builds: the radii of sigma 3..5 only */
#define MMB_RADIUS_BUCKETS(X) X(12) X(16) X(20)
#else
#define MMB_RADIUS_BUCKETS(X) \
  X(2) X(4) X(6) X(8) X(10) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) X(20) \
  X(24) X(32) X(48) X(64)
#endif

}  // namespace mmb
