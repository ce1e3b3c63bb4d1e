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
  unsigned char* base = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
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
          v = __ldg(in + (r0 + rr) * pitch + (xs + src));
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
    __syncthreads();                 // every window is in registers: the stage is free
    if (tid == 0) {
      const int64_t tn = t + (int64_t)kXStages * gridDim.x;
      if (tn < n_tiles) issue_load(tn, stage);
    }
    if (active) {
      // scatter form: one window value feeds every output it touches, so it stays
      // in the operand-reuse cache across the FFMAs and each FFMA reads a single
      // register (its accumulator) from the banks - no bank conflicts, and 32
      // independent accumulation chains per thread
      // (A, B) of one output share an FFMA2: the pair (g, h) is a 64-bit uniform
      // operand, the input value the broadcast scalar
      float2 acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = G::R4 - R; i < G::R4 + 16 + R; ++i) {
        const float2 vv = make_float2(win[i], win[i]);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int t = i - (G::R4 + j);
          if (t >= -R && t <= R) {
            const int wi = t < 0 ? -t : t;
            acc[j] = ffma2(vv, w.gh[wi], acc[j]);
          }
        }
      }
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
  if (tid == 0) tma_wait_read<0>();
}

constexpr int kNBx = 16;
constexpr int kNSEG = 8;

template <int R>
static int run_x(const float* in, float* outA, float* outB, int64_t nrows, int X,
                 int64_t pitch, const LogWeights& w, cudaStream_t st) {
  ProfScope ps(PROF_LOG_X, (double)nrows * pitch, st);
  const uintptr_t bits = (uintptr_t)in | (uintptr_t)outA | (uintptr_t)outB;
  const bool tma_ok = R <= 24 && (bits & 15) == 0 && pitch % 4 == 0 && X >= 32 && nrows >= 32 &&
                      nrows < (int64_t)1 << 31;
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
