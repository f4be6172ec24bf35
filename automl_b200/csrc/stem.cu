// Stem: Conv2D 3x3 stride 2 'SAME', 3 -> cout (32..64), + folded BN + activation.
// Reads the float32 NHWC image the reference feeds the network (efficientnet_model.py:526-527)
// and writes NHWC fp16.  Memory-bound (K = 27): bytes = 12*n*h*w + 2*n*ho*wo*cout.
//
// One thread = one output pixel x all output channels; the [27][cout] weights sit in shared
// memory as fp32 and are read as warp-wide broadcasts.
#include "common.cuh"

namespace edet {

constexpr int kStemThreads = 128;
constexpr int kStemMaxC = 64;

template <int COUT, int ACT>
__global__ void __launch_bounds__(kStemThreads)
stem_kernel(const float* __restrict__ in, __half* __restrict__ out, const __half* __restrict__ w,
            const float* __restrict__ bias, int h, int wd, int ho, int wo, int pad_t, int pad_l) {
  __shared__ float ws[27][COUT];
  __shared__ float bs[COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += kStemThreads)
    ws[i / COUT][i % COUT] = __half2float(w[i]);
  for (int i = threadIdx.x; i < COUT; i += kStemThreads) bs[i] = bias[i];
  __syncthreads();
  const int n = blockIdx.z;
  const int p = blockIdx.x * kStemThreads + threadIdx.x;
  if (p >= ho * wo) return;
  const int oy = p / wo, ox = p - oy * wo;
  float x[27];
  const float* in_n = in + static_cast<size_t>(n) * h * wd * 3;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - pad_t + ky;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - pad_l + kx;
      const bool ok = iy >= 0 && iy < h && ix >= 0 && ix < wd;
      const float* px = in_n + (static_cast<size_t>(iy) * wd + ix) * 3;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) x[(ky * 3 + kx) * 3 + ci] = ok ? __ldg(px + ci) : 0.f;
    }
  }
  __half* o = out + (static_cast<size_t>(n) * ho * wo + p) * COUT;
#pragma unroll
  for (int c0 = 0; c0 < COUT; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bs[c0 + j];
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float4 w0 = *reinterpret_cast<const float4*>(&ws[t][c0]);
      const float4 w1 = *reinterpret_cast<const float4*>(&ws[t][c0 + 4]);
      acc[0] = fmaf(x[t], w0.x, acc[0]); acc[1] = fmaf(x[t], w0.y, acc[1]);
      acc[2] = fmaf(x[t], w0.z, acc[2]); acc[3] = fmaf(x[t], w0.w, acc[3]);
      acc[4] = fmaf(x[t], w1.x, acc[4]); acc[5] = fmaf(x[t], w1.y, acc[5]);
      acc[6] = fmaf(x[t], w1.z, acc[6]); acc[7] = fmaf(x[t], w1.w, acc[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = apply_act_t<ACT>(acc[j]);
    *reinterpret_cast<uint4*>(o + c0) = float_to_half8(acc);
  }
}

template <int COUT>
static int launch_stem(const float* in, __half* out, const __half* w, const float* bias, int n,
                       int h, int wd, int act, cudaStream_t s) {
  const int ho = ceil_div(h, 2), wo = ceil_div(wd, 2);
  const int pad_t = same_pad_before(h, 3, 2), pad_l = same_pad_before(wd, 3, 2);
  dim3 grid(ceil_div(ho * wo, kStemThreads), 1, n);
  if (act == EDET_ACT_SWISH)
    stem_kernel<COUT, EDET_ACT_SWISH><<<grid, kStemThreads, 0, s>>>(in, out, w, bias, h, wd, ho, wo, pad_t, pad_l);
  else if (act == EDET_ACT_RELU6)
    stem_kernel<COUT, EDET_ACT_RELU6><<<grid, kStemThreads, 0, s>>>(in, out, w, bias, h, wd, ho, wo, pad_t, pad_l);
  else if (act == EDET_ACT_NONE)
    stem_kernel<COUT, EDET_ACT_NONE><<<grid, kStemThreads, 0, s>>>(in, out, w, bias, h, wd, ho, wo, pad_t, pad_l);
  else {
    set_error("stem: unsupported activation %d", act);
    return EDET_ERR_UNSUPPORTED;
  }
  EDET_CHECK_LAUNCH();
  return EDET_OK;
}

}  // namespace edet

extern "C" int edet_stem_conv(const float* in, edet_half* out, const edet_half* w,
                              const float* bias, int n, int h, int wd, int cout, int act,
                              edet_stream_t stream) {
  using namespace edet;
  EDET_CHECK_ARG(in && out && w && bias, "stem: null pointer");
  EDET_CHECK_ARG(n > 0 && h > 0 && wd > 0, "stem: bad shape");
  const __half* hw = reinterpret_cast<const __half*>(w);
  __half* ho = reinterpret_cast<__half*>(out);
  cudaStream_t s = as_stream(stream);
  switch (cout) {
    case 24: return launch_stem<24>(in, ho, hw, bias, n, h, wd, act, s);
    case 32: return launch_stem<32>(in, ho, hw, bias, n, h, wd, act, s);
    case 40: return launch_stem<40>(in, ho, hw, bias, n, h, wd, act, s);
    case 48: return launch_stem<48>(in, ho, hw, bias, n, h, wd, act, s);
    case 56: return launch_stem<56>(in, ho, hw, bias, n, h, wd, act, s);
    case 64: return launch_stem<64>(in, ho, hw, bias, n, h, wd, act, s);
    default:
      set_error("stem: unsupported cout %d (24/32/40/48/56/64)", cout);
      return EDET_ERR_UNSUPPORTED;
  }
}
