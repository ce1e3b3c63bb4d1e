// Contiguous-axis (x) first sweep: A = g * I, B = h * I.
//
// conv_x_tma_kernel (default): a tile is 32 rows x 128 outputs.  The input
// window (128 + 2*R4 columns, R4 = R rounded up to a multiple of 4) arrives as
// 32x32-float TMA boxes with the 128-byte swizzle, so thread (lane = row,
// warp = 16-output segment) reads its whole window with conflict-free 16-byte
// shared loads; out-of-volume columns are zero-filled by TMA and then patched
// with scipy's 'reflect' by the few tiles that touch an x face.  The arithmetic
// is the register-blocked scatter of log_kernels.cuh (the symmetric-pair form
// sum_k w[k] * (v[c-k] + v[c+k]) was measured slower: both addends of every pair
// sit in the same register bank, so each FADD costs two issue cycles and the
// dependent FADD->FFMA pairs starve the pipe).  Results go to swizzled output
// boxes (16-byte conflict-free stores) and leave through TMA stores, which clip
// at the volume edge - no per-thread global stores at all.
//
// conv_x_first_kernel (log_kernels.cuh) remains for unaligned rows, tiny volumes
// and radii above 24.
#include "log_kernels.cuh"
#include "tma.cuh"

namespace mmb {

constexpr int kXRows = 32;        // rows per CTA (= lanes)
constexpr int kXCols = 128;       // outputs per row per CTA (8 warps x 16)
constexpr int kXThreads = 256;
constexpr int kBoxBytes = 32 * 32 * 4;

constexpr int kXStages = 3;       // input ring depth of the persistent kernel

template <int R>
struct XGeom {
  static constexpr int R4 = (R + 3) / 4 * 4;
  static constexpr int WIN = kXCols + 2 * R4;          // tile window, floats
  static constexpr int NBOX = (WIN + 31) / 32;         // input boxes per stage
  static constexpr int W = 16 + 2 * R4;                // per-thread window, floats
  static constexpr int STAGE_BYTES = NBOX * kBoxBytes;
  static constexpr size_t SMEM = (size_t)kXStages * STAGE_BYTES + 8 * kBoxBytes + 1024 + 64;
};

// byte offset of window position p of row rr inside a stage (128-byte swizzle)
__device__ __forceinline__ int x_sw_off(int rr, int p) {
  const int col = p & 31;
  return (p >> 5) * kBoxBytes + rr * 128 + ((((col >> 2) ^ (rr & 7)) << 4) | ((col & 3) << 2));
}

// Persistent: grid = 2 CTAs per SM, each walks tiles t = blockIdx.x, +gridDim.x, ...
// (x tile fastest, so tiles sharing halo columns run on neighbouring CTAs at the
// same time and the halo is an L2 hit).  Input tiles flow through a kXStages-deep
// TMA ring; a thread copies its whole window to registers first, which frees the
// stage for the load of tile t + kXStages while tile t is still being computed.
template <int R>
__global__ void __launch_bounds__(kXThreads, 2)
conv_x_tma_kernel(const __grid_constant__ CUtensorMap tm_in,
                  const __grid_constant__ CUtensorMap tm_a,
                  const __grid_constant__ CUtensorMap tm_b, const float* __restrict__ in,
                  int64_t nrows, int X, int64_t pitch, int n_xt, int64_t n_tiles,
                  const __grid_constant__ LogWeights w) {
  using G = XGeom<R>;
  extern __shared__ unsigned char smem_raw[];
  // swizzled boxes need 1024-byte alignment
  // aligned up on the 32-bit SHARED address, as an offset from smem_raw: a round trip through
  // uintptr_t would make every later access a generic LD / ST with 64-bit address arithmetic
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* s_in = base;
  unsigned char* s_a = base + kXStages * G::STAGE_BYTES;
  unsigned char* s_b = s_a + 4 * kBoxBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_b + 4 * kBoxBytes);    // [kXStages]

  const int tid = threadIdx.x;
  const int lane = tid & 31, seg = tid >> 5;
  const int key = (lane & 7) << 4;

  auto issue_load = [&](int64_t t, int stage) {          // one thread
    const int xt = (int)(t % n_xt);
    const int64_t rt = t / n_xt;
    mbar_arrive_expect_tx(&bar[stage], G::STAGE_BYTES);
#pragma unroll
    for (int b = 0; b < G::NBOX; ++b)
      tma_load_2d(s_in + stage * G::STAGE_BYTES + b * kBoxBytes, &tm_in, &bar[stage],
                  xt * kXCols - G::R4 + 32 * b, (int)(rt * kXRows));
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kXStages; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kXStages; ++s) {
      const int64_t t = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (t < n_tiles) issue_load(t, s);
    }
  }

  int it = 0;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
    const int stage = it % kXStages;
    const uint32_t parity = (uint32_t)(it / kXStages) & 1u;
    const int x0 = (int)(t % n_xt) * kXCols;
    const int64_t r0 = (t / n_xt) * kXRows;
    const int xs = x0 - G::R4;                      // global x of window position 0
    unsigned char* st_in = s_in + stage * G::STAGE_BYTES;
    mbar_wait(&bar[stage], parity);

    // 'reflect' at the x faces: patch the zero-filled window positions that feed
    // valid outputs (CTA-uniform condition; interior tiles skip this entirely)
    const int n_left = xs < 0 ? -xs : 0;
    const int pr0 = X - xs;                                   // first position with x >= X
    const int pr1 = min(G::WIN, X + R - xs);                  // one past the last one needed
    const int n_right = pr1 > pr0 ? pr1 - pr0 : 0;           // pr0 > 0 because x0 < X
    const int n_fix = n_left + n_right;
    if (n_fix > 0) {
      for (int i = tid; i < kXRows * n_fix; i += kXThreads) {
        const int rr = i / n_fix, q = i - rr * n_fix;
        const int p = q < n_left ? q : pr0 + (q - n_left);
        const int src = reflect_index(xs + p, X) - xs;        // window position of the source
        float v = 0.f;
        if (src >= 0 && src < G::WIN) {
          v = *reinterpret_cast<const float*>(st_in + x_sw_off(rr, src));
        } else if (r0 + rr < nrows) {                         // volume narrower than the window
          v = __ldcg(in + (r0 + rr) * pitch + (xs + src));
        }
        *reinterpret_cast<float*>(st_in + x_sw_off(rr, p)) = v;
      }
      __syncthreads();
    }

    // segments entirely past the right face have nothing to compute (warp-uniform)
    const bool active = x0 + seg * 16 < X;
    float win[G::W];
    if (active) {
      const unsigned char* rowp = st_in + lane * 128;
#pragma unroll
      for (int c = 0; c < G::W / 4; ++c) {
        const int ci = seg * 4 + c;
        const float4 v = *reinterpret_cast<const float4*>(rowp + (ci >> 3) * kBoxBytes +
                                                          (((ci & 7) << 4) ^ key));
        win[4 * c + 0] = v.x; win[4 * c + 1] = v.y; win[4 * c + 2] = v.z; win[4 * c + 3] = v.w;
      }
    }
    // the previous tile's stores must have finished reading the output boxes
    if (tid == 0) tma_wait_read<0>();
    __syncthreads();
    if (active) {
      // scatter form: one window value feeds every output it touches, so it stays
      // in the operand-reuse cache across the FFMAs and each FFMA reads a single
      // register (its accumulator) from the banks - no bank conflicts, and 32
      // independent accumulation chains per thread
      // (A, B) of one output share an FFMA2: the pair (g, h) is a 64-bit uniform
      // operand, the input value the broadcast scalar
      float2 acc[16];
      x_taps_sym<R, G::R4, G::W>(win, acc, w);
      unsigned char* oa = s_a + lane * 128;
      unsigned char* ob = s_b + lane * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int co = seg * 4 + q;
        const int off = (co >> 3) * kBoxBytes + (((co & 7) << 4) ^ key);
        *reinterpret_cast<float4*>(oa + off) =
            make_float4(acc[4 * q].x, acc[4 * q + 1].x, acc[4 * q + 2].x, acc[4 * q + 3].x);
        *reinterpret_cast<float4*>(ob + off) =
            make_float4(acc[4 * q].y, acc[4 * q + 1].y, acc[4 * q + 2].y, acc[4 * q + 3].y);
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      // every window value of this tile has been consumed by the arithmetic above: only now
      // may the stage be refilled (a barrier right after the window copy does not order
      // reads still in flight against the async proxy)
      const int64_t tn = t + (int64_t)kXStages * gridDim.x;
      if (tn < n_tiles) issue_load(tn, stage);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (x0 + 32 * b < X) {
          tma_store_2d(&tm_a, s_a + b * kBoxBytes, x0 + 32 * b, (int)r0);
          tma_store_2d(&tm_b, s_b + b * kBoxBytes, x0 + 32 * b, (int)r0);
        }
      }
      tma_commit();
    }
  }
  // the stores must have LANDED before the CTA exits: the y sweep reads them
  if (tid == 0) tma_wait_all<0>();
}

// ---- warp-specialised x sweep ------------------------------------------------------
// conv_x_ws_kernel: one persistent CTA per SM, three roles that only meet at mbarriers:
//   warp 0          producer: TMA loads of the input window into a kWsIn-deep ring
//   warps 1..NCW    consumers: lane = row, warp = 16-output segment; copy the window to
//                   registers (which frees the ring slot), FFMA2 scatter, write the
//                   (A, B) results into one of kWsOut staging buffers
//   warp NCW+1      store issuer: TMA stores of a full staging buffer, frees it when the
//                   stores have read it
// Nothing is CTA-wide, so loads, arithmetic and stores of different tiles overlap inside
// one SM.  (The CTA-synchronous kernel above adds its arithmetic to a 0.33 ms
// load/window/store skeleton instead of hiding it: 0.39 / 0.46 / 0.56 ms at r = 12 / 16 /
// 20 on a 505^3 chunk; this one: 0.37 / 0.43 / 0.49 ms with 16 consumer warps, and it
// keeps improving with the consumer count - 8: 0.62, 12: 0.43 at r = 16 - until shared
// memory for the staging buffers runs out.  Writing results straight from registers,
// 64 bytes per lane, frees that memory but costs more than it buys: 0.68 ms.)  'reflect'
// at the x faces is applied by the consumers while they fill their register window.
#ifndef MMB_WS_IN
#define MMB_WS_IN 2
#endif
#ifndef MMB_WS_OUT
#define MMB_WS_OUT 2
#endif
#ifndef MMB_WS_NCW
#define MMB_WS_NCW 16
#endif
constexpr int kWsIn = MMB_WS_IN;          // input ring depth
constexpr int kWsOut = MMB_WS_OUT;         // output staging buffers
constexpr int kWsNCW = MMB_WS_NCW;         // consumer warps = 16-column segments per tile
constexpr int kWsCols = 16 * kWsNCW;
constexpr int kWsThreads = 32 * (kWsNCW + 2);
constexpr int kWsOutBoxes = (kWsCols + 31) / 32;

template <int R>
struct WsGeom {
  static constexpr int R4 = (R + 3) / 4 * 4;
  static constexpr int WIN = kWsCols + 2 * R4;
  static constexpr int NBOX = (WIN + 31) / 32;
  static constexpr int W = 16 + 2 * R4;                // per-thread window, floats
  static constexpr int STAGE_BYTES = NBOX * kBoxBytes;
  static constexpr int OUT_BYTES = 2 * kWsOutBoxes * kBoxBytes;     // A boxes then B boxes
#ifdef MMB_WS_DIRECT
  static constexpr size_t SMEM = (size_t)kWsIn * STAGE_BYTES + 1024 + 256;
#else
  static constexpr size_t SMEM = (size_t)kWsIn * STAGE_BYTES + kWsOut * OUT_BYTES + 1024 + 256;
#endif
};

template <int R>
__global__ void __launch_bounds__(kWsThreads, 1)
conv_x_ws_kernel(const __grid_constant__ CUtensorMap tm_in,
                 const __grid_constant__ CUtensorMap tm_a,
                 const __grid_constant__ CUtensorMap tm_b, int64_t nrows, int X, int n_xt,
                 int n_tiles, const __grid_constant__ LogWeights w, float* __restrict__ outA,
                 float* __restrict__ outB, int64_t pitch) {
  using G = WsGeom<R>;
  extern __shared__ unsigned char smem_raw[];
  // aligned up on the 32-bit SHARED address, as an offset from smem_raw: a round trip through
  // uintptr_t would make every later access a generic LD / ST with 64-bit address arithmetic
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* s_in = base;
  unsigned char* s_out = base + kWsIn * G::STAGE_BYTES;
#ifdef MMB_WS_DIRECT
  uint64_t* full_in = reinterpret_cast<uint64_t*>(s_out);
#else
  uint64_t* full_in = reinterpret_cast<uint64_t*>(s_out + kWsOut * G::OUT_BYTES);
#endif
  uint64_t* empty_in = full_in + kWsIn;
  uint64_t* full_out = empty_in + kWsIn;
  uint64_t* empty_out = full_out + kWsOut;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kWsIn; ++s) { mbar_init(&full_in[s], 1); mbar_init(&empty_in[s], kWsNCW); }
    for (int o = 0; o < kWsOut; ++o) { mbar_init(&full_out[o], kWsNCW); mbar_init(&empty_out[o], 1); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ---- producer ----------------------------------------------------------------
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it % kWsIn;
        if (it >= kWsIn) mbar_wait(&empty_in[s], (uint32_t)((it / kWsIn) - 1) & 1u);
        const int rt = t / n_xt, xt = t - rt * n_xt;
        mbar_arrive_expect_tx(&full_in[s], G::STAGE_BYTES);
#pragma unroll
        for (int b = 0; b < G::NBOX; ++b)
          tma_load_2d(s_in + s * G::STAGE_BYTES + b * kBoxBytes, &tm_in, &full_in[s],
                      xt * kWsCols - G::R4 + 32 * b, rt * kXRows);
      }
    }
  } else if (warp <= kWsNCW) {
    // ---- consumers ---------------------------------------------------------------
    const int seg = warp - 1;
    const int key = (lane & 7) << 4;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int s = it % kWsIn;
      const int rt = t / n_xt, xt = t - rt * n_xt;
      const int x0 = xt * kWsCols;
      const int xs = x0 - G::R4 + 16 * seg;            // global x of this thread's win[0]
      const unsigned char* st_in = s_in + s * G::STAGE_BYTES;
      mbar_wait(&full_in[s], (uint32_t)(it / kWsIn) & 1u);
      float win[G::W];
      const unsigned char* rowp = st_in + lane * 128;
#pragma unroll
      for (int c = 0; c < G::W / 4; ++c) {
        const int ci = seg * 4 + c;
        const float4 v = *reinterpret_cast<const float4*>(rowp + (ci >> 3) * kBoxBytes +
                                                          (((ci & 7) << 4) ^ key));
        win[4 * c + 0] = v.x; win[4 * c + 1] = v.y; win[4 * c + 2] = v.z; win[4 * c + 3] = v.w;
      }
      // scipy 'reflect' at the x faces (TMA zero-filled those positions); both tests are
      // warp-uniform and false for interior windows.  Left face: only the window that
      // starts at x = -R4, mirrored inside the registers.  Right face: the mirrored
      // sample x' = 2X - 1 - x is read back from the tile (the host checks X >= WIN, so
      // it is always inside it).
      if (xs == -G::R4) {
#pragma unroll
        for (int i = 0; i < G::R4; ++i) win[i] = win[2 * G::R4 - 1 - i];
      }
      if (G::R4 > 16 && xs == 16 - G::R4) {            // the second segment also starts left of 0
#pragma unroll
        for (int i = 0; i < G::R4 - 16; ++i) win[i] = win[2 * (G::R4 - 16) - 1 - i];
      }
      if (xs + G::W > X) {
#pragma unroll
        for (int i = 0; i < G::W; ++i) {
          const int x = xs + i;
          if (x >= X) {
            const int p = 2 * X - 1 - x - (x0 - G::R4);
            float v = 0.f;
            if (p >= 0 && p < G::WIN)
              v = *reinterpret_cast<const float*>(st_in + x_sw_off(lane, p));
            win[i] = v;
          }
        }
      }
      float2 acc[16];
      x_taps_sym<R, G::R4, G::W>(win, acc, w);
      // the slot may be refilled only once every window value has been consumed: an arrival
      // right after the window copy would not wait for reads still in flight
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_in[s]);

#ifdef MMB_WS_DIRECT
      {
        // results leave straight from the registers: a lane owns 64 contiguous bytes of
        // its row in A and in B (no staging buffer, no store warp)
        const int64_t row = (int64_t)rt * kXRows + lane;
        const int xo = x0 + 16 * seg;
        if (row < nrows) {
          float* pa = outA + row * pitch + xo;
          float* pb = outB + row * pitch + xo;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (xo + 4 * q < (int)pitch) {
              *reinterpret_cast<float4*>(pa + 4 * q) = make_float4(
                  acc[4 * q].x, acc[4 * q + 1].x, acc[4 * q + 2].x, acc[4 * q + 3].x);
              *reinterpret_cast<float4*>(pb + 4 * q) = make_float4(
                  acc[4 * q].y, acc[4 * q + 1].y, acc[4 * q + 2].y, acc[4 * q + 3].y);
            }
          }
        }
        continue;
      }
#endif
      const int o = it % kWsOut;
      if (it >= kWsOut) mbar_wait(&empty_out[o], (uint32_t)((it / kWsOut) - 1) & 1u);
      unsigned char* oa = s_out + o * G::OUT_BYTES + lane * 128;
      unsigned char* ob = oa + kWsOutBoxes * kBoxBytes;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int co = seg * 4 + q;
        const int off = (co >> 3) * kBoxBytes + (((co & 7) << 4) ^ key);
        *reinterpret_cast<float4*>(oa + off) =
            make_float4(acc[4 * q].x, acc[4 * q + 1].x, acc[4 * q + 2].x, acc[4 * q + 3].x);
        *reinterpret_cast<float4*>(ob + off) =
            make_float4(acc[4 * q].y, acc[4 * q + 1].y, acc[4 * q + 2].y, acc[4 * q + 3].y);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_out[o]);
    }
  } else {
    // ---- store issuer --------------------------------------------------------------
#ifdef MMB_WS_DIRECT
    return;
#endif
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int o = it % kWsOut;
        const int rt = t / n_xt, xt = t - rt * n_xt;
        const int x0 = xt * kWsCols;
        mbar_wait(&full_out[o], (uint32_t)(it / kWsOut) & 1u);
        const unsigned char* oa = s_out + o * G::OUT_BYTES;
        const unsigned char* ob = oa + kWsOutBoxes * kBoxBytes;
#pragma unroll
        for (int b = 0; b < kWsOutBoxes; ++b) {
          if (x0 + 32 * b < X) {
            tma_store_2d(&tm_a, oa + b * kBoxBytes, x0 + 32 * b, rt * kXRows);
            tma_store_2d(&tm_b, ob + b * kBoxBytes, x0 + 32 * b, rt * kXRows);
          }
        }
        tma_commit();
        tma_wait_read<0>();
        mbar_arrive(&empty_out[o]);
      }
    }
    // every store of this CTA has landed before it exits (the y sweep reads them)
    if (lane == 0) tma_wait_all<0>();
  }
}

#ifdef MMB_X_WARP_KERNEL
// ---- warp-autonomous x sweep (experimental, not the default) -------------------------
// Measured on B200, 505^3, r = 12/16/20: 0.47/0.49/0.61 ms single-stage (17 warps/SM),
// 0.48/0.59/0.79 ms with a per-warp two-stage ring (10 warps/SM) - no better than the
// CTA-synchronous TMA kernel (0.39/0.46/0.56 ms), so the limit of the x sweep is not its
// barriers; kept for the next round's experiments.
// conv_x_warp_kernel: every WARP walks its own sequence of 32-row x 32-output tiles
// with its own two-stage cp.async ring and its own output staging, so nothing in the
// main loop is CTA-wide: no __syncthreads, no shared barrier, no store drain that a
// whole CTA waits for.  (The TMA kernel above is CTA-synchronous, three barriers per
// tile, and sat at 60 % of the FMA pipe whatever its CTA shape.)  Lane = row: a lane
// reads its (16 + 2R)-float window with 16-byte shared loads - the row pitch is 4 mod
// 8 floats, which spreads the eight lanes of a quarter-warp over all banks - and
// scatters it into 16 (A, B) accumulators with FFMA2; two such passes cover the 32
// outputs.  Results go through a small staging tile so that global stores are
// 128-byte row segments.  Tiles at an x face (window partly outside the volume) and
// row tails take a scalar load path with scipy's 'reflect'.
constexpr int kWTile = 32;                    // outputs per row per tile
constexpr int kWWarps = 1;                    // warps per CTA: warps are autonomous, so one-warp CTAs pack shared memory best
constexpr int kWOutPitch = 36;                // staging pitch, floats (4 mod 8)
#ifndef MMB_XW_STAGES
#define MMB_XW_STAGES 1
#endif
constexpr int kWStages = MMB_XW_STAGES;       // 1: latency hidden by other warps (17 per SM at r = 16); 2: per-warp prefetch

template <int R>
struct WGeom {
  static constexpr int R4 = (R + 3) / 4 * 4;
  static constexpr int WIN = kWTile + 2 * R4;                       // floats per row
  static constexpr int PITCH = (WIN / 4) % 2 == 1 ? WIN : WIN + 4;  // 4 mod 8
  static constexpr int STAGE = 32 * PITCH;                          // floats
  static constexpr int WARP_FLOATS = kWStages * STAGE + 32 * kWOutPitch;
  static constexpr size_t SMEM = (size_t)kWWarps * WARP_FLOATS * sizeof(float);
};

template <int R>
__global__ void __launch_bounds__(32 * kWWarps)
conv_x_warp_kernel(const float* __restrict__ in, float* __restrict__ outA,
                   float* __restrict__ outB, int64_t nrows, int X, int64_t pitch, int n_xt,
                   int n_tiles, const __grid_constant__ LogWeights w) {
  using G = WGeom<R>;
  extern __shared__ __align__(16) float wsm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* s_in = wsm + warp * G::WARP_FLOATS;
  float* s_out = s_in + kWStages * G::STAGE;
  const int n_warps = gridDim.x * kWWarps;
  const int gw = blockIdx.x * kWWarps + warp;

  // stage <- tile t (rows r0.., window columns x0 - R4 ..)
  auto load_tile = [&](int t, int stage) {
    const int rt = t / n_xt, xt = t - rt * n_xt;
    const int64_t r0 = (int64_t)rt * 32;
    const int xs = xt * kWTile - G::R4;
    float* dst = s_in + stage * G::STAGE;
    const bool fast = xs >= 0 && xs + G::WIN <= (int)pitch && xs + G::WIN - G::R4 + R <= X &&
                      r0 + 32 <= nrows;
    if (fast) {
      constexpr int CPR = G::WIN / 4;                 // 16-byte chunks per row
#pragma unroll 4
      for (int i = lane; i < 32 * CPR; i += 32) {
        const int rr = i / CPR, c = i - rr * CPR;
        __pipeline_memcpy_async(dst + rr * G::PITCH + 4 * c,
                                in + (r0 + rr) * pitch + xs + 4 * c, 16);
      }
    } else {
      // face or tail tile: 16-byte copies wherever a chunk lies inside the row, scalar
      // loads with scipy's 'reflect' for the few columns outside [0, X)
      constexpr int CPR = G::WIN / 4;
      for (int i = lane; i < 32 * CPR; i += 32) {
        const int rr = i / CPR, c = i - rr * CPR;
        const int x = xs + 4 * c;
        float* d = dst + rr * G::PITCH + 4 * c;
        if (r0 + rr >= nrows) {
          *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (x >= 0 && x + 4 <= X) {
          __pipeline_memcpy_async(d, in + (r0 + rr) * pitch + x, 16);
        } else {
          const float* src = in + (r0 + rr) * pitch;
#pragma unroll
          for (int k = 0; k < 4; ++k) d[k] = __ldcg(src + reflect_index(x + k, X));
        }
      }
    }
    __pipeline_commit();
  };

  int t = gw;
  if (kWStages == 2 && t < n_tiles) load_tile(t, 0);
  int stage = 0;
  for (; t < n_tiles; t += n_warps, stage ^= (kWStages - 1)) {
    if (kWStages == 2) {
      const int tn = t + n_warps;
      if (tn < n_tiles) load_tile(tn, stage ^ 1);
      else __pipeline_commit();
      __pipeline_wait_prior(1);              // this tile's copies (all but the newest group)
    } else {
      load_tile(t, 0);
      __pipeline_wait_prior(0);
    }
    __syncwarp();
    const int rt = t / n_xt, xt = t - rt * n_xt;
    const int64_t r0 = (int64_t)rt * 32;
    const int x0 = xt * kWTile;
    const float* row = s_in + stage * G::STAGE + lane * G::PITCH;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      // window of this pass: positions half*16 .. half*16 + 16 + 2*R4 of the row
      float win[16 + 2 * G::R4];
#pragma unroll
      for (int c = 0; c < (16 + 2 * G::R4) / 4; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(row + half * 16 + 4 * c);
        win[4 * c + 0] = v.x; win[4 * c + 1] = v.y; win[4 * c + 2] = v.z; win[4 * c + 3] = v.w;
      }
      float2 acc[16];
      x_taps_sym<R, G::R4, G::W>(win, acc, w);
      // A then B through the staging tile: lane = row writes, row-segment reads
#pragma unroll
      for (int ab = 0; ab < 2; ++ab) {
        float* dstg = ab == 0 ? outA : outB;
        __syncwarp();                         // previous read-back of the staging tile done
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = ab == 0
              ? make_float4(acc[4 * q].x, acc[4 * q + 1].x, acc[4 * q + 2].x, acc[4 * q + 3].x)
              : make_float4(acc[4 * q].y, acc[4 * q + 1].y, acc[4 * q + 2].y, acc[4 * q + 3].y);
          *reinterpret_cast<float4*>(s_out + lane * kWOutPitch + 4 * q) = v;
        }
        __syncwarp();
        // 4 lanes cover the 16 floats of a row; 8 rows per instruction
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = i * 8 + (lane >> 2), c4 = lane & 3;
          const float4 v = *reinterpret_cast<const float4*>(s_out + rr * kWOutPitch + 4 * c4);
          const int x = x0 + half * 16 + 4 * c4;
          if (r0 + rr < nrows && x < (int)pitch)
            *reinterpret_cast<float4*>(dstg + (r0 + rr) * pitch + x) = v;
        }
      }
    }
    __syncwarp();                             // every lane is done with this stage
  }
  __pipeline_wait_prior(0);
}

#endif  // MMB_X_WARP_KERNEL

constexpr int kNBx = 16;
constexpr int kNSEG = 8;

template <int R>
static int run_x(const float* in, float* outA, float* outB, int64_t nrows, int X,
                 int64_t pitch, const LogWeights& w, cudaStream_t st) {
  ProfScope ps(PROF_LOG_X, (double)nrows * pitch, st);
  const uintptr_t bits = (uintptr_t)in | (uintptr_t)outA | (uintptr_t)outB;
  const bool tma_ok = R <= 24 && (bits & 15) == 0 && pitch % 4 == 0 && X >= 32 && nrows >= 32 &&
                      nrows < (int64_t)1 << 31;
#ifdef MMB_X_WARP_KERNEL
  if constexpr (R <= 24) {
    const int n_xt = (int)cdiv(X, kWTile);
    const int64_t n_tiles64 = cdiv(nrows, 32) * n_xt;
    if ((bits & 15) == 0 && pitch % 4 == 0 && X >= 1 && n_tiles64 < ((int64_t)1 << 31)) {
      using G = WGeom<R>;
      auto kern = conv_x_warp_kernel<R>;
      static bool configured = false;
      static int ctas_per_sm = 1;
      if (!configured) {
        MMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)G::SMEM));
        MMB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern,
                                                                     32 * kWWarps, G::SMEM));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        configured = true;
      }
      const int64_t want = cdiv(n_tiles64, kWWarps);
      int64_t full = (int64_t)ctas_per_sm * num_sms();
      // a warp strides through the tiles by the total warp count: keep that coprime with
      // the tiles per row so every warp meets its share of the (dearer) x-face tiles
      auto gcd = [](int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; };
      while (full > 1 && gcd(full * kWWarps, n_xt) != 1) --full;
      kern<<<(unsigned)(want < full ? want : full), 32 * kWWarps, G::SMEM, st>>>(
          in, outA, outB, nrows, X, pitch, n_xt, (int)n_tiles64, w);
      MMB_CHECK_LAUNCH();
      return MMB_OK;
    }
  }
#endif
#ifndef MMB_X_NO_WS
  if constexpr (R <= 24) {
    using G = WsGeom<R>;
    const int n_xt = (int)cdiv(X, kWsCols);
    const int64_t n_tiles64 = cdiv(nrows, kXRows) * n_xt;
    // long launches only: with a few tiles per SM the pipeline never fills and the
    // CTA-synchronous kernel below (two CTAs per SM, smaller tiles) is faster
    if (tma_ok && X >= G::WIN && n_tiles64 >= 16 * (int64_t)num_sms() &&
        n_tiles64 < ((int64_t)1 << 31)) {
      auto kern = conv_x_ws_kernel<R>;
      static bool configured = false;
      if (!configured) {
        MMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)G::SMEM));
        configured = true;
      }
      CUtensorMap tin, ta, tb;
      const uint64_t dims[2] = {(uint64_t)X, (uint64_t)nrows};
      const uint64_t strides[1] = {(uint64_t)pitch * sizeof(float)};
      const uint32_t box[2] = {32, 32};
      if (encode_tensor_map_f32(&tin, in, 2, dims, strides, box, true) ||
          encode_tensor_map_f32(&ta, outA, 2, dims, strides, box, true) ||
          encode_tensor_map_f32(&tb, outB, 2, dims, strides, box, true))
        return MMB_ERR_CUDA;
      const int64_t grid = n_tiles64 < num_sms() ? n_tiles64 : num_sms();
      kern<<<(unsigned)grid, kWsThreads, G::SMEM, st>>>(tin, ta, tb, nrows, X, n_xt,
                                                        (int)n_tiles64, w, outA, outB, pitch);
      MMB_CHECK_LAUNCH();
      return MMB_OK;
    }
  }
#endif
  if constexpr (R <= 24) {
    if (tma_ok) {
      using G = XGeom<R>;
      auto kern = conv_x_tma_kernel<R>;
      static bool configured = false;
      if (!configured) {
        MMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)G::SMEM));
        configured = true;
      }
      CUtensorMap tin, ta, tb;
      const uint64_t dims[2] = {(uint64_t)X, (uint64_t)nrows};
      const uint64_t strides[1] = {(uint64_t)pitch * sizeof(float)};
      const uint32_t box[2] = {32, 32};
      if (encode_tensor_map_f32(&tin, in, 2, dims, strides, box, true) ||
          encode_tensor_map_f32(&ta, outA, 2, dims, strides, box, true) ||
          encode_tensor_map_f32(&tb, outB, 2, dims, strides, box, true))
        return MMB_ERR_CUDA;
      const int n_xt = (int)cdiv(X, kXCols);
      const int64_t n_tiles = cdiv(nrows, kXRows) * n_xt;
      const int64_t grid = n_tiles < 2 * num_sms() ? n_tiles : 2 * num_sms();
      kern<<<(unsigned)grid, kXThreads, G::SMEM, st>>>(tin, ta, tb, in, nrows, X, pitch, n_xt,
                                                       n_tiles, w);
      MMB_CHECK_LAUNCH();
      return MMB_OK;
    }
  }
  dim3 grid((unsigned)cdiv(nrows, 32), (unsigned)cdiv(X, kNBx * kNSEG));
  conv_x_first_kernel<R, kNBx, kNSEG><<<grid, 32 * kNSEG, 0, st>>>(in, outA, outB, nrows, X,
                                                                  pitch, w);
  MMB_CHECK_LAUNCH();
  return MMB_OK;
}

int launch_x_first(int r, const float* in, float* outA, float* outB, int64_t nrows, int X,
                   int64_t pitch, const LogWeights& w, cudaStream_t st) {
#define X_(RR) if (r <= RR) return run_x<RR>(in, outA, outB, nrows, X, pitch, w, st);
  MMB_RADIUS_BUCKETS(X_)
#undef X_
  set_error("radius %d has no compiled bucket", r);
  return MMB_ERR_UNSUPPORTED;
}

}  // namespace mmb
