// Depthwise k x k convolution ('SAME', NHWC fp16) + bias + activation, with the SE squeeze
// (per-image spatial sums) produced in the same pass, and the SE gate kernel.
//
// Memory-bound op: algorithmic bytes per launch =
//   2 * n * c * (h*w + ho*wo) + 2 * k*k*c      (+ 4 * n * tiles * c when the squeeze is emitted).
//
// Mapping: one thread owns 8 consecutive channels (one 128-bit access) of one output column
// and walks a strip of ROWS output rows; the k*k*8 folded weights live in registers as half2,
// accumulation is fp32.  Consecutive threads cover consecutive (x, channel-group) positions,
// i.e. consecutive 16-byte pieces of the NHWC row, so every warp access is fully coalesced;
// the kx re-reads of neighbouring pixels are served by L1.
#include "common.cuh"

namespace edet {

constexpr int kDwThreads = 256;

template <int K, int S>
struct DwCfg {
  // 5x5: 4 channels per thread (64-bit accesses) so the 25-tap fp32 weight set fits in
  // registers; 3x3: 8 channels per thread (128-bit accesses).
  static constexpr int CPT = (K == 5) ? 4 : 8;
  static constexpr int ROWS = (S == 2) ? 4 : 8;             // output rows per thread
  static constexpr int IN_ROWS = (ROWS - 1) * S + K;        // input rows touched
};

template <int CPT> struct VecT;
template <> struct VecT<8> { using type = uint4; };
template <> struct VecT<4> { using type = uint2; };

// The arithmetic runs on channel PAIRS with the packed fp32 FMA of sm_100 (FFMA2): a 5x5
// depthwise conv needs 25 MACs per output element, which at the HBM rate is more than the
// scalar FFMA pipe can issue, so the packed form is what keeps this kernel memory-bound.
template <int CPT>
__device__ __forceinline__ void vec_to_float2(const typename VecT<CPT>::type& v, float2* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < CPT / 2; ++i) f[i] = __half22float2(h[i]);
}
template <int CPT>
__device__ __forceinline__ typename VecT<CPT>::type float_to_vec(const float* f) {
  typename VecT<CPT>::type v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < CPT / 2; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

template <int K, int S, int ACT, bool HAS_BIAS, bool HAS_SE>
__global__ void __launch_bounds__(kDwThreads)
depthwise_kernel(const __half* __restrict__ in, __half* __restrict__ out,
                 const __half* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ se_partial, int h, int wd, int c, int ho, int wo, int pad_t,
                 int pad_l) {
  constexpr int CPT = DwCfg<K, S>::CPT;
  constexpr int P = CPT / 2;  // channel pairs
  constexpr int ROWS = DwCfg<K, S>::ROWS;
  constexpr int IN_ROWS = DwCfg<K, S>::IN_ROWS;
  using Vec = typename VecT<CPT>::type;
  const int cg_count = c / CPT;
  const int e = blockIdx.x * kDwThreads + threadIdx.x;  // (x_out, channel group) flattened
  const int n = blockIdx.z;
  const int oy0 = blockIdx.y * ROWS;
  const bool active = e < wo * cg_count;
  const int ox = active ? e / cg_count : 0;
  const int cg = active ? e - ox * cg_count : 0;
  const int ch = cg * CPT;

  float ssum[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) ssum[i] = 0.f;

  if (active) {
    // folded weights for this channel group, converted once to fp32 pairs
    float2 wreg[K * K][P];
#pragma unroll
    for (int t = 0; t < K * K; ++t) {
      const Vec wv = __ldg(reinterpret_cast<const Vec*>(w + static_cast<size_t>(t) * c + ch));
      vec_to_float2<CPT>(wv, wreg[t]);
    }

    float2 acc[ROWS][P];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int i = 0; i < P; ++i) acc[r][i] = make_float2(0.f, 0.f);

    const __half* in_n = in + static_cast<size_t>(n) * h * wd * c;
    const int iy0 = oy0 * S - pad_t;
    const int ix0 = ox * S - pad_l;

#pragma unroll
    for (int ir = 0; ir < IN_ROWS; ++ir) {
      const int iy = iy0 + ir;
      const bool row_ok = (iy >= 0) && (iy < h);
      float2 xv[K][P];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int ix = ix0 + kx;
        Vec v;
        memset(&v, 0, sizeof(v));
        if (row_ok && ix >= 0 && ix < wd)
          v = __ldg(reinterpret_cast<const Vec*>(in_n + (static_cast<size_t>(iy) * wd + ix) * c + ch));
        vec_to_float2<CPT>(v, xv[kx]);
      }
      // input row ir feeds output row r through tap ky = ir - r*S
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const int ky = ir - r * S;
        if (ky >= 0 && ky < K) {
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int i = 0; i < P; ++i)
              acc[r][i] = __ffma2_rn(xv[kx][i], wreg[ky * K + kx][i], acc[r][i]);
        }
      }
    }

    float bv[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) bv[i] = HAS_BIAS ? __ldg(bias + ch + i) : 0.f;
    __half* out_n = out + static_cast<size_t>(n) * ho * wo * c;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int oy = oy0 + r;
      if (oy < ho) {
        float o[CPT];
#pragma unroll
        for (int i = 0; i < P; ++i) {
          o[2 * i] = apply_act_t<ACT>(acc[r][i].x + bv[2 * i]);
          o[2 * i + 1] = apply_act_t<ACT>(acc[r][i].y + bv[2 * i + 1]);
        }
        if (HAS_SE) {
#pragma unroll
          for (int i = 0; i < CPT; ++i) ssum[i] += o[i];
        }
        *reinterpret_cast<Vec*>(out_n + (static_cast<size_t>(oy) * wo + ox) * c + ch) =
            float_to_vec<CPT>(o);
      }
    }
  }

  if (HAS_SE) {
    // Deterministic block reduction: every thread parks its sums, then one thread per
    // channel adds the contributions of the threads that own that channel in index order.
    __shared__ float red[kDwThreads][CPT + 1];
#pragma unroll
    for (int i = 0; i < CPT; ++i) red[threadIdx.x][i] = ssum[i];
    __syncthreads();
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int tiles = gridDim.x * gridDim.y;
    float* dst = se_partial + (static_cast<size_t>(n) * tiles + tile) * c;
    const int e0 = blockIdx.x * kDwThreads;
    for (int ch_o = threadIdx.x; ch_o < c; ch_o += kDwThreads) {
      const int g = ch_o / CPT, lane_c = ch_o % CPT;
      // first thread t in this block with (e0 + t) % cg_count == g
      int t = (g - (e0 % cg_count) + cg_count) % cg_count;
      float s = 0.f;
      for (; t < kDwThreads; t += cg_count) s += red[t][lane_c];
      dst[ch_o] = s;
    }
  }
}

// SE gate: gridDim.x = images, gridDim.y = slices of the project-weight slab.
//   mean[c] = inv_hw * sum_t partial[n][t][c]
//   r[j]    = act(b1[j] + sum_c w1[j][c] * mean[c])
//   gate[c] = sigmoid(b2[c] + sum_j w2[c][j] * r[j])
//   wt_scaled[n][o][c] = wt[o][c] * gate[c]
__global__ void __launch_bounds__(256)
se_fc_kernel(const float* __restrict__ partial, int tiles, float inv_hw,
             const float* __restrict__ w1, const float* __restrict__ b1,
             const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ gate,
             const __half* __restrict__ wt, __half* __restrict__ wt_scaled, int c, int se, int nout,
             int act) {
  extern __shared__ float sm[];
  float* mean = sm;        // [c]
  float* red = sm + c;     // [se]
  float* g = red + se;     // [c]
  float* red_scratch = g + c;  // [8][c] worst case is bounded by blockDim.x floats
  const int n = blockIdx.x;
  const float* pn = partial + static_cast<size_t>(n) * tiles * c;
  // Squeeze: G thread groups split the tile range of every channel; loads are issued 8 at a
  // time (this loop is pure latency), partial sums are combined in a fixed order.
  {
    const int groups = max(1, min(static_cast<int>(blockDim.x) / c, 8));
    const int gi = threadIdx.x / c, ch = threadIdx.x % c;
    float* gsum = g;  // reuse the gate buffer as [groups][c] scratch when it fits
    if (groups > 1 && gi < groups) {
      float s = 0.f;
      for (int t0 = gi; t0 < tiles; t0 += groups * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int t = t0 + u * groups;
          v[u] = t < tiles ? pn[static_cast<size_t>(t) * c + ch] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
      }
      red_scratch[gi * c + ch] = s;
    }
    (void)gsum;
    if (groups > 1) {
      __syncthreads();
      for (int cc = threadIdx.x; cc < c; cc += blockDim.x) {
        float s = 0.f;
        for (int q = 0; q < groups; ++q) s += red_scratch[q * c + cc];
        mean[cc] = s * inv_hw;
      }
    } else {
      for (int cc = threadIdx.x; cc < c; cc += blockDim.x) {
        float s = 0.f;
        for (int t0 = 0; t0 < tiles; t0 += 8) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = (t0 + u) < tiles ? pn[static_cast<size_t>(t0 + u) * c + cc] : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) s += v[u];
        }
        mean[cc] = s * inv_hw;
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < se; j += nwarps) {
    float s = 0.f;
    for (int ch = lane; ch < c; ch += 32) s = fmaf(w1[static_cast<size_t>(j) * c + ch], mean[ch], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[j] = apply_act(s + b1[j], act);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = b2[ch];
    for (int j = 0; j < se; ++j) s = fmaf(w2[static_cast<size_t>(ch) * se + j], red[j], s);
    const float gv = 1.0f / (1.0f + expf(-s));
    g[ch] = gv;
    if (blockIdx.y == 0) gate[static_cast<size_t>(n) * c + ch] = gv;
  }
  __syncthreads();
  if (wt != nullptr) {
    const int cg = c >> 3;
    const int total = nout * cg;
    __half* dst = wt_scaled + static_cast<size_t>(n) * nout * c;
    // the weight slab is split over gridDim.y CTAs (each recomputed the tiny gate above)
    const int per = (total + gridDim.y - 1) / gridDim.y;
    const int i_end = min(total, static_cast<int>(blockIdx.y + 1) * per);
    for (int i = blockIdx.y * per + threadIdx.x; i < i_end; i += blockDim.x) {
      const int gidx = i % cg;
      float f[8];
      half8_to_float(__ldg(reinterpret_cast<const uint4*>(wt) + i), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] *= g[gidx * 8 + e];
      reinterpret_cast<uint4*>(dst)[i] = float_to_half8(f);
    }
  }
}

template <int K, int S>
static int launch_dw(const __half* in, __half* out, const __half* w, const float* bias,
                     float* se_partial, int n, int h, int wd, int c, int act, cudaStream_t stream) {
  const int ho = ceil_div(h, S), wo = ceil_div(wd, S);
  const int pad_t = same_pad_before(h, K, S), pad_l = same_pad_before(wd, K, S);
  dim3 grid(ceil_div(wo * (c / DwCfg<K, S>::CPT), kDwThreads), ceil_div(ho, DwCfg<K, S>::ROWS), n);
#define EDET_DW_LAUNCH(ACT, HB, HS)                                                        \
  depthwise_kernel<K, S, ACT, HB, HS><<<grid, kDwThreads, 0, stream>>>(                    \
      in, out, w, bias, se_partial, h, wd, c, ho, wo, pad_t, pad_l)
  const bool hb = bias != nullptr, hs = se_partial != nullptr;
  if (act == EDET_ACT_SWISH && hb && hs) EDET_DW_LAUNCH(EDET_ACT_SWISH, true, true);
  else if (act == EDET_ACT_SWISH && hb && !hs) EDET_DW_LAUNCH(EDET_ACT_SWISH, true, false);
  else if (act == EDET_ACT_RELU6 && hb && !hs) EDET_DW_LAUNCH(EDET_ACT_RELU6, true, false);
  else if (act == EDET_ACT_RELU6 && hb && hs) EDET_DW_LAUNCH(EDET_ACT_RELU6, true, true);
  else if (act == EDET_ACT_NONE && !hb && !hs) EDET_DW_LAUNCH(EDET_ACT_NONE, false, false);
  else if (act == EDET_ACT_NONE && hb && !hs) EDET_DW_LAUNCH(EDET_ACT_NONE, true, false);
  else {
    set_error("depthwise: unsupported combination act=%d bias=%d se=%d", act, (int)hb, (int)hs);
    return EDET_ERR_UNSUPPORTED;
  }
#undef EDET_DW_LAUNCH
  EDET_CHECK_LAUNCH();
  return EDET_OK;
}

}  // namespace edet

extern "C" int edet_depthwise_tiles(int h, int wd, int c, int k, int stride) {
  using namespace edet;
  const int ho = ceil_div(h, stride), wo = ceil_div(wd, stride);
  const int rows = (stride == 2) ? 4 : 8;
  const int cpt = (k == 5) ? 4 : 8;
  return ceil_div(wo * (c / cpt), kDwThreads) * ceil_div(ho, rows);
}

extern "C" int edet_depthwise_conv(const edet_half* in, edet_half* out, const edet_half* w,
                                   const float* bias, float* se_partial, int n, int h, int wd,
                                   int c, int k, int stride, int act, edet_stream_t stream) {
  using namespace edet;
  EDET_CHECK_ARG(in && out && w, "depthwise: null pointer");
  EDET_CHECK_ARG(n > 0 && h > 0 && wd > 0 && c > 0 && c % 8 == 0, "depthwise: bad shape (c%%8)");
  EDET_CHECK_ARG((k == 3 || k == 5) && (stride == 1 || stride == 2),
                 "depthwise: k must be 3 or 5 and stride 1 or 2 (got %d, %d)", k, stride);
  const __half* hi = reinterpret_cast<const __half*>(in);
  const __half* hw = reinterpret_cast<const __half*>(w);
  __half* ho = reinterpret_cast<__half*>(out);
  cudaStream_t s = as_stream(stream);
  if (k == 3 && stride == 1) return launch_dw<3, 1>(hi, ho, hw, bias, se_partial, n, h, wd, c, act, s);
  if (k == 3 && stride == 2) return launch_dw<3, 2>(hi, ho, hw, bias, se_partial, n, h, wd, c, act, s);
  if (k == 5 && stride == 1) return launch_dw<5, 1>(hi, ho, hw, bias, se_partial, n, h, wd, c, act, s);
  return launch_dw<5, 2>(hi, ho, hw, bias, se_partial, n, h, wd, c, act, s);
}

extern "C" int edet_se_fc(const float* partial, int tiles, float inv_hw, const float* w1,
                          const float* b1, const float* w2, const float* b2, float* gate,
                          const edet_half* wt, edet_half* wt_scaled, int n, int c, int se, int nout,
                          int act, edet_stream_t stream) {
  using namespace edet;
  EDET_CHECK_ARG(partial && w1 && b1 && w2 && b2 && gate, "se_fc: null pointer");
  EDET_CHECK_ARG(n > 0 && c > 0 && c % 8 == 0 && se > 0 && tiles > 0, "se_fc: bad shape");
  EDET_CHECK_ARG(!wt || (wt_scaled && nout > 0), "se_fc: wt given without wt_scaled/nout");
  const size_t smem = static_cast<size_t>(2 * c + se + 256) * sizeof(float);
  EDET_CHECK_ARG(smem <= 48 * 1024, "se_fc: c too large");
  int nsplit = 1;
  if (wt) {
    nsplit = (nout * (c >> 3)) / (256 * 8);
    nsplit = nsplit < 1 ? 1 : (nsplit > 64 ? 64 : nsplit);
  }
  se_fc_kernel<<<dim3(n, nsplit), 256, smem, as_stream(stream)>>>(
      partial, tiles, inv_hw, w1, b1, w2, b2, gate, reinterpret_cast<const __half*>(wt),
      reinterpret_cast<__half*>(wt_scaled), c, se, nout, act);
  EDET_CHECK_LAUNCH();
  return EDET_OK;
}
