// Post-processing on the device: pre-NMS (class max/argmax + sigmoid + anchor decode) and a
// bit-faithful NonMaxSuppressionV5 (hard / gaussian soft) followed by the serving-layout gather.
#include <math_constants.h>

#include "common.cuh"

namespace edet {

// ------------------------------------------------------------------------------------------
// pre-NMS
// ------------------------------------------------------------------------------------------
constexpr int kPreMaxLevels = 8;
constexpr int kPrePix = 16;       // pixels per CTA
constexpr int kPreThreads = 256;

struct PreLevel {
  const __half* cls;
  const __half* box;
  int pixels;        // h*w
  int block_begin;   // first CTA (blockIdx.x) of this level
  int anchor_begin;  // first flattened anchor of this level
};
struct PreParams {
  PreLevel lv[kPreMaxLevels];
  int levels, ld_cls, ld_box, num_anchors, num_classes, total_anchors;
};

__global__ void __launch_bounds__(kPreThreads)
pre_nms_kernel(const PreParams p, const float* __restrict__ anchors, float* __restrict__ boxes,
               float* __restrict__ scores, int32_t* __restrict__ classes) {
  extern __shared__ __align__(16) uint8_t pre_smem[];
  __half* cls_s = reinterpret_cast<__half*>(pre_smem);
  int l = 0;
  while (l + 1 < p.levels && static_cast<int>(blockIdx.x) >= p.lv[l + 1].block_begin) ++l;
  const PreLevel lv = p.lv[l];
  const int n = blockIdx.y;
  const int pix0 = (blockIdx.x - lv.block_begin) * kPrePix;
  const int npix = min(kPrePix, lv.pixels - pix0);
  // coalesced copy of npix * ld_cls halves (contiguous in NHWC)
  const uint4* src = reinterpret_cast<const uint4*>(
      lv.cls + (static_cast<size_t>(n) * lv.pixels + pix0) * p.ld_cls);
  const int nvec = npix * p.ld_cls / 8;
  for (int i = threadIdx.x; i < nvec; i += kPreThreads)
    reinterpret_cast<uint4*>(cls_s)[i] = ldg_nc_v4(src + i);
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= npix * p.num_anchors) return;
  const int pl = t / p.num_anchors, a = t - pl * p.num_anchors;
  const __half* row = cls_s + pl * p.ld_cls + a * p.num_classes;
  float best = __half2float(row[0]);
  int best_c = 0;
  for (int c = 1; c < p.num_classes; ++c) {
    const float v = __half2float(row[c]);
    if (v > best) {  // strict: first maximum wins, like tf.argmax
      best = v;
      best_c = c;
    }
  }
  const int anchor = lv.anchor_begin + (pix0 + pl) * p.num_anchors + a;
  const size_t o = static_cast<size_t>(n) * p.total_anchors + anchor;
  scores[o] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-best)));
  classes[o] = best_c;
  // box decode (tf2/anchors.py:30-58), float32, no FMA contraction
  const uint2 bv = __ldg(reinterpret_cast<const uint2*>(
      lv.box + (static_cast<size_t>(n) * lv.pixels + pix0 + pl) * p.ld_box + a * 4));
  const float2 t01 = __half22float2(*reinterpret_cast<const __half2*>(&bv.x));
  const float2 t23 = __half22float2(*reinterpret_cast<const __half2*>(&bv.y));
  const float ty = t01.x, tx = t01.y, th = t23.x, tw = t23.y;
  const float4 an = __ldg(reinterpret_cast<const float4*>(anchors) + anchor);
  const float ycenter_a = __fmul_rn(__fadd_rn(an.x, an.z), 0.5f);
  const float xcenter_a = __fmul_rn(__fadd_rn(an.y, an.w), 0.5f);
  const float ha = __fsub_rn(an.z, an.x), wa = __fsub_rn(an.w, an.y);
  const float w = __fmul_rn(expf(tw), wa), h = __fmul_rn(expf(th), ha);
  const float yc = __fadd_rn(__fmul_rn(ty, ha), ycenter_a);
  const float xc = __fadd_rn(__fmul_rn(tx, wa), xcenter_a);
  const float hh = __fmul_rn(h, 0.5f), hw = __fmul_rn(w, 0.5f);
  reinterpret_cast<float4*>(boxes)[o] =
      make_float4(__fsub_rn(yc, hh), __fsub_rn(xc, hw), __fadd_rn(yc, hh), __fadd_rn(xc, hw));
}

// ------------------------------------------------------------------------------------------
// NonMaxSuppressionV5 (TensorFlow core/kernels/image/non_max_suppression_op.cc semantics)
// ------------------------------------------------------------------------------------------
constexpr int kNmsThreads = 1024;
constexpr int kNmsMaxOut = 512;

__device__ __forceinline__ bool better(float sa, int ia, float sb, int ib) {
  return sa > sb || (sa == sb && ia < ib);
}

__device__ __forceinline__ float iou_tf(const float4 a, const float4 b) {
  const float ymin_i = fminf(a.x, a.z), xmin_i = fminf(a.y, a.w);
  const float ymax_i = fmaxf(a.x, a.z), xmax_i = fmaxf(a.y, a.w);
  const float ymin_j = fminf(b.x, b.z), xmin_j = fminf(b.y, b.w);
  const float ymax_j = fmaxf(b.x, b.z), xmax_j = fmaxf(b.y, b.w);
  const float area_i = __fmul_rn(__fsub_rn(ymax_i, ymin_i), __fsub_rn(xmax_i, xmin_i));
  const float area_j = __fmul_rn(__fsub_rn(ymax_j, ymin_j), __fsub_rn(xmax_j, xmin_j));
  if (area_i <= 0.f || area_j <= 0.f) return 0.f;
  const float iymin = fmaxf(ymin_i, ymin_j), ixmin = fmaxf(xmin_i, xmin_j);
  const float iymax = fminf(ymax_i, ymax_j), ixmax = fminf(xmax_i, xmax_j);
  const float inter = __fmul_rn(fmaxf(__fsub_rn(iymax, iymin), 0.f),
                                fmaxf(__fsub_rn(ixmax, ixmin), 0.f));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_i, area_j), inter));
}

__global__ void __launch_bounds__(kNmsThreads)
nms_v5_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
              const int32_t* __restrict__ classes, const float* __restrict__ image_scales,
              int image_id_base, int k, int max_out, float iou_thr, float score_thr, float sigma,
              float clip_h, float clip_w, float* __restrict__ detections,
              int32_t* __restrict__ sel_index, int32_t* __restrict__ valid,
              float* __restrict__ work_scores, int32_t* __restrict__ work_begin,
              const int32_t* __restrict__ need_full) {
  // Full-queue path: only runs for images the shared-memory fast path could not prove exact.
  if (need_full != nullptr && need_full[blockIdx.x] == 0) return;
  __shared__ float4 sel_box[kNmsMaxOut];
  __shared__ int sel_idx[kNmsMaxOut];
  __shared__ float sel_score[kNmsMaxOut];
  __shared__ float sim_s[kNmsMaxOut];
  __shared__ float wgt_s[kNmsMaxOut];
  __shared__ float red_s[32];
  __shared__ int red_i[32];
  __shared__ int best_i;
  __shared__ int nsel_s;

  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + static_cast<size_t>(n) * k;
  const float* sc = scores + static_cast<size_t>(n) * k;
  float* ws = work_scores + static_cast<size_t>(n) * k;
  int32_t* wb = work_begin + static_cast<size_t>(n) * k;
  const bool soft = sigma > 0.f;
  const float scale = soft ? __fdiv_rn(-0.5f, sigma) : 0.f;

  // candidate queue = every box with score > threshold; removed entries become -inf
  float my_s = -CUDART_INF_F;
  int my_i = 0x7fffffff;
  for (int i = tid; i < k; i += kNmsThreads) {
    float s = sc[i];
    if (!(s > score_thr)) s = -CUDART_INF_F;
    ws[i] = s;
    wb[i] = 0;
    if (better(s, i, my_s, my_i)) {
      my_s = s;
      my_i = i;
    }
  }
  if (tid == 0) nsel_s = 0;
  __syncthreads();

  while (true) {
    // ---- pop: block-wide arg-max, ties to the lower index ----
    float s = my_s;
    int i = my_i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float so = __shfl_xor_sync(0xffffffffu, s, o);
      const int io = __shfl_xor_sync(0xffffffffu, i, o);
      if (better(so, io, s, i)) {
        s = so;
        i = io;
      }
    }
    if (lane == 0) {
      red_s[warp] = s;
      red_i[warp] = i;
    }
    __syncthreads();
    if (warp == 0) {
      s = red_s[lane];
      i = red_i[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float so = __shfl_xor_sync(0xffffffffu, s, o);
        const int io = __shfl_xor_sync(0xffffffffu, i, o);
        if (better(so, io, s, i)) {
          s = so;
          i = io;
        }
      }
      const int nsel = nsel_s;
      if (s == -CUDART_INF_F || nsel >= max_out) {
        if (lane == 0) best_i = -1;
      } else {
        // ---- lazily apply the suppression of the boxes selected since the last visit ----
        const int begin = wb[i];
        const float4 cb = bx[i];
        for (int j = begin + lane; j < nsel; j += 32) {
          const float sim = iou_tf(cb, sel_box[j]);
          float wgt = static_cast<float>(exp(static_cast<double>(__fmul_rn(__fmul_rn(scale, sim), sim))));
          if (!(soft || sim <= iou_thr)) wgt = 0.f;
          sim_s[j] = sim;
          wgt_s[j] = wgt;
        }
        __syncwarp();
        if (lane == 0) {
          float cur = s;
          bool hard = false;
          for (int j = nsel - 1; j >= begin; --j) {
            cur = __fmul_rn(cur, wgt_s[j]);
            if (!soft && sim_s[j] > iou_thr) {
              hard = true;
              break;
            }
            if (cur <= score_thr) break;
          }
          float new_s = -CUDART_INF_F;
          if (!hard) {
            if (cur == s) {
              sel_box[nsel] = cb;
              sel_idx[nsel] = i;
              sel_score[nsel] = cur;
              nsel_s = nsel + 1;
            } else if (cur > score_thr) {
              new_s = cur;
              wb[i] = nsel;
            }
          }
          ws[i] = new_s;
          best_i = i;
        }
      }
    }
    __syncthreads();
    const int popped = best_i;
    if (popped < 0) break;
    if ((popped % kNmsThreads) == tid) {
      // owner refreshes its cached local best
      my_s = -CUDART_INF_F;
      my_i = 0x7fffffff;
      for (int j0 = tid; j0 < k; j0 += 8 * kNmsThreads) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {   // issue the loads together: this is latency bound
          const int j = j0 + u * kNmsThreads;
          v[u] = j < k ? ws[j] : -CUDART_INF_F;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + u * kNmsThreads;
          if (j < k && better(v[u], j, my_s, my_i)) {
            my_s = v[u];
            my_i = j;
          }
        }
      }
    }
    // (no barrier needed here: the next iteration's first __syncthreads orders best_i reuse)
  }

  // ---- gather into the serving layout (padded with index 0 / score 0, like TF) ----
  const int nsel = nsel_s;
  const float scale_img = image_scales ? image_scales[n] : 1.f;
  for (int r = tid; r < max_out; r += kNmsThreads) {
    const int idx = r < nsel ? sel_idx[r] : 0;
    const float score = r < nsel ? sel_score[r] : 0.f;
    const float4 b = bx[idx];
    float* d = detections + (static_cast<size_t>(n) * max_out + r) * 7;
    d[0] = static_cast<float>(image_id_base + n);
    d[1] = __fmul_rn(fminf(fmaxf(b.x, 0.f), clip_h), scale_img);
    d[2] = __fmul_rn(fminf(fmaxf(b.y, 0.f), clip_w), scale_img);
    d[3] = __fmul_rn(fminf(fmaxf(b.z, 0.f), clip_h), scale_img);
    d[4] = __fmul_rn(fminf(fmaxf(b.w, 0.f), clip_w), scale_img);
    d[5] = score;
    d[6] = static_cast<float>(classes[static_cast<size_t>(n) * k + idx] + 1);
    sel_index[static_cast<size_t>(n) * max_out + r] = idx;
  }
  if (tid == 0) valid[n] = nsel;
}


// ------------------------------------------------------------------------------------------
// Fast path: the same algorithm on the top candidates only, entirely in shared memory.
//
// Only the highest-scoring candidates are ever popped before max_output_size boxes are selected
// (a few hundred for a 76 725-anchor image), so each image first compacts its top <= kFastCap
// candidates into shared memory (adaptive two-level histogram threshold on the score bits; the
// boxes stay in global memory and are fetched when a candidate is popped), then runs the
// exact lazy-suppression loop there.  Exactness is PROVEN per image at run time: every popped
// (stale) score must be strictly greater than the best excluded score; otherwise the image is
// flagged and the full-queue kernel above recomputes it.
// ------------------------------------------------------------------------------------------
constexpr int kFastThreads = 512;
constexpr int kFastCap = 16384;
constexpr int kFastPer = kFastCap / kFastThreads;  // 32 slots per thread
constexpr int kFastBins = 2048;
constexpr int kFastWarps = kFastThreads / 32;

__device__ __forceinline__ uint32_t score_key(float s) {
  const uint32_t b = __float_as_uint(s);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // monotone float -> uint
}

struct FastSmem {
  float score[kFastCap];
  int idx[kFastCap];
  unsigned short begin[kFastCap];
  int hist[kFastBins];
  float4 sel_box[kNmsMaxOut];
  int sel_idx[kNmsMaxOut];
  float sel_score[kNmsMaxOut];
  float sim[kNmsMaxOut];
  float wgt[kNmsMaxOut];
  float red_s[kFastWarps];
  int red_i[kFastWarps];
  int red_slot[kFastWarps];
  uint32_t kmin, kmax;
  int count, nsel, bstar, sstar, fail;
  float excl_max;
};

__global__ void __launch_bounds__(kFastThreads)
nms_v5_fast_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                   const int32_t* __restrict__ classes, const float* __restrict__ image_scales,
                   int image_id_base, int k, int max_out, float iou_thr, float score_thr,
                   float sigma, float clip_h, float clip_w, float* __restrict__ detections,
                   int32_t* __restrict__ sel_index, int32_t* __restrict__ valid,
                   int32_t* __restrict__ need_full) {
  extern __shared__ __align__(16) uint8_t fast_raw[];
  FastSmem& sm = *reinterpret_cast<FastSmem*>(fast_raw);
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + static_cast<size_t>(n) * k;
  const float* sc = scores + static_cast<size_t>(n) * k;
  const bool soft = sigma > 0.f;
  const float scale = soft ? __fdiv_rn(-0.5f, sigma) : 0.f;

  // ---- A. key range of the valid candidates ----
  uint32_t kmin = 0xffffffffu, kmax = 0u;
  for (int i = tid; i < k; i += kFastThreads) {
    const float s = sc[i];
    if (s > score_thr) {
      const uint32_t key = score_key(s);
      kmin = min(kmin, key);
      kmax = max(kmax, key);
    }
  }
  if (tid == 0) {
    sm.kmin = 0xffffffffu; sm.kmax = 0u; sm.count = 0; sm.nsel = 0; sm.fail = 0;
    sm.excl_max = -CUDART_INF_F; sm.bstar = 0; sm.sstar = 0;
  }
  for (int i = tid; i < kFastBins; i += kFastThreads) sm.hist[i] = 0;
  __syncthreads();
  atomicMin(&sm.kmin, kmin);
  atomicMax(&sm.kmax, kmax);
  __syncthreads();
  kmin = sm.kmin; kmax = sm.kmax;
  const bool any_valid = kmax >= kmin;
  // bin = floor((key - kmin) * inv) in [0, kFastBins); the fraction inside the bin gives a
  // second-level sub-bin.  Both are monotone in the key, which is all the selection needs.
  const double inv = static_cast<double>(kFastBins) /
                     (static_cast<double>(any_valid ? kmax - kmin : 0u) + 1.0);
  auto bin_of = [&](uint32_t key, int& sub) -> int {
    const double x = static_cast<double>(key - kmin) * inv;
    int b = static_cast<int>(x);
    b = b > kFastBins - 1 ? kFastBins - 1 : b;
    int sb = static_cast<int>((x - static_cast<double>(b)) * kFastBins);
    sub = sb > kFastBins - 1 ? kFastBins - 1 : (sb < 0 ? 0 : sb);
    return b;
  };
  // ---- B. coarse histogram -> threshold bin ----
  if (any_valid) {
    for (int i = tid; i < k; i += kFastThreads) {
      const float s = sc[i];
      int sub;
      if (s > score_thr) atomicAdd(&sm.hist[bin_of(score_key(s), sub)], 1);
    }
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0, b = kFastBins;
    while (b > 0 && acc + sm.hist[b - 1] <= kFastCap) acc += sm.hist[--b];
    sm.bstar = b;       // bins >= bstar are taken whole; bin bstar-1 is refined below
    sm.count = acc;     // (reused as the running total for the refinement)
  }
  __syncthreads();
  const int bstar = sm.bstar;
  const int coarse_total = sm.count;
  __syncthreads();
  // ---- B2. refine the boundary bin with a second-level histogram ----
  for (int i = tid; i < kFastBins; i += kFastThreads) sm.hist[i] = 0;
  __syncthreads();
  if (any_valid && bstar > 0) {
    for (int i = tid; i < k; i += kFastThreads) {
      const float s = sc[i];
      if (s > score_thr) {
        int sub;
        if (bin_of(score_key(s), sub) == bstar - 1) atomicAdd(&sm.hist[sub], 1);
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    int acc = coarse_total, sb = kFastBins;
    if (bstar > 0)
      while (sb > 0 && acc + sm.hist[sb - 1] <= kFastCap) acc += sm.hist[--sb];
    sm.sstar = sb;
    sm.count = 0;
    if (any_valid && acc == 0) sm.fail = 1;  // one score value alone overflows the capacity
  }
  __syncthreads();
  const int sstar = sm.sstar;
  // ---- D. compaction into shared memory; best excluded score ----
  float excl = -CUDART_INF_F;
  if (any_valid && !sm.fail) {
    for (int i = tid; i < k; i += kFastThreads) {
      const float s = sc[i];
      if (s > score_thr) {
        int sub;
        const int b = bin_of(score_key(s), sub);
        if (b >= bstar || (b == bstar - 1 && sub >= sstar)) {
          const int slot = atomicAdd(&sm.count, 1);
          sm.score[slot] = s;
          sm.idx[slot] = i;
          sm.begin[slot] = 0;
        } else {
          excl = fmaxf(excl, s);
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) excl = fmaxf(excl, __shfl_xor_sync(0xffffffffu, excl, o));
  if (lane == 0) sm.red_s[warp] = excl;
  __syncthreads();
  if (tid == 0) {
    float e = -CUDART_INF_F;
    for (int w = 0; w < kFastWarps; ++w) e = fmaxf(e, sm.red_s[w]);
    sm.excl_max = e;
  }
  __syncthreads();
  const int count = sm.count;
  const float excl_max = sm.excl_max;
  for (int sl = count + tid; sl < kFastCap; sl += kFastThreads) sm.score[sl] = -CUDART_INF_F;
  __syncthreads();

  // thread-local best over its interleaved slots (slot = tid + u * kFastThreads)
  float my_s = -CUDART_INF_F;
  int my_i = 0x7fffffff, my_slot = -1;
  auto rescan = [&]() {
    my_s = -CUDART_INF_F; my_i = 0x7fffffff; my_slot = -1;
#pragma unroll 8
    for (int u = 0; u < kFastPer; ++u) {
      const int sl = tid + u * kFastThreads;
      const float v = sm.score[sl];
      if (v > -CUDART_INF_F) {
        const int id = sm.idx[sl];
        if (better(v, id, my_s, my_i)) { my_s = v; my_i = id; my_slot = sl; }
      }
    }
  };
  rescan();

  // ---- E. exact lazy-suppression loop ----
  while (!sm.fail) {
    const int nsel = sm.nsel;   // stable here: only written between the two barriers below
    float s = my_s;
    int i = my_i, slot = my_slot;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float so = __shfl_xor_sync(0xffffffffu, s, o);
      const int io = __shfl_xor_sync(0xffffffffu, i, o);
      const int lo = __shfl_xor_sync(0xffffffffu, slot, o);
      if (better(so, io, s, i)) { s = so; i = io; slot = lo; }
    }
    if (lane == 0) { sm.red_s[warp] = s; sm.red_i[warp] = i; sm.red_slot[warp] = slot; }
    __syncthreads();
    // every warp reduces the per-warp partials itself: no second barrier for the broadcast
    s = lane < kFastWarps ? sm.red_s[lane] : -CUDART_INF_F;
    i = lane < kFastWarps ? sm.red_i[lane] : 0x7fffffff;
    slot = lane < kFastWarps ? sm.red_slot[lane] : -1;
#pragma unroll
    for (int o = kFastWarps / 2; o > 0; o >>= 1) {
      const float so = __shfl_xor_sync(0xffffffffu, s, o);
      const int io = __shfl_xor_sync(0xffffffffu, i, o);
      const int lo = __shfl_xor_sync(0xffffffffu, slot, o);
      if (better(so, io, s, i)) { s = so; i = io; slot = lo; }
    }
    s = __shfl_sync(0xffffffffu, s, 0);
    i = __shfl_sync(0xffffffffu, i, 0);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (nsel >= max_out) break;
    if (s == -CUDART_INF_F) {           // queue exhausted
      if (excl_max > -CUDART_INF_F) { if (tid == 0) sm.fail = 2; __syncthreads(); }
      break;
    }
    if (!(s > excl_max)) {              // an excluded candidate could be next: not provable
      if (tid == 0) sm.fail = 3;
      __syncthreads();
      break;
    }
    const int owner = slot % kFastThreads;
    if (warp == (owner >> 5)) {
      const int begin = sm.begin[slot];
      const float4 cb = bx[i];
      // Newest -> oldest over the boxes selected since this candidate's last visit, 32 at a
      // time.  A non-overlapping box has weight exp(0) == 1 exactly and cannot trigger either
      // break, so only the overlapping ones are evaluated (fp64 exp) and applied, in order.
      float cur = s;
      bool hard = false, done = false;
      for (int hi = nsel; hi > begin && !done; hi -= 32) {
        const int j = hi - 1 - lane;                  // lane 0 = newest of this chunk
        float simv = 0.f;
        if (j >= begin) simv = iou_tf(cb, sm.sel_box[j]);
        unsigned mask = __ballot_sync(0xffffffffu, simv > 0.f);
        float wgt = 1.f;
        if (simv > 0.f) {
          wgt = static_cast<float>(exp(static_cast<double>(__fmul_rn(__fmul_rn(scale, simv), simv))));
          if (!(soft || simv <= iou_thr)) wgt = 0.f;
        }
        while (mask) {                                // ascending lane = descending j
          const int l = __ffs(mask) - 1;
          mask &= mask - 1;
          const float wl = __shfl_sync(0xffffffffu, wgt, l);
          const float sl = __shfl_sync(0xffffffffu, simv, l);
          cur = __fmul_rn(cur, wl);
          if (!soft && sl > iou_thr) { hard = true; done = true; break; }
          if (cur <= score_thr) { done = true; break; }
        }
      }
      if (tid == owner) {
        float new_s = -CUDART_INF_F;
        if (!hard) {
          if (cur == s) {
            sm.sel_box[nsel] = cb;
            sm.sel_idx[nsel] = i;
            sm.sel_score[nsel] = cur;
            sm.nsel = nsel + 1;
          } else if (cur > score_thr) {
            new_s = cur;
            sm.begin[slot] = static_cast<unsigned short>(nsel);
          }
        }
        sm.score[slot] = new_s;
        rescan();
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (sm.fail) {
    if (tid == 0) need_full[n] = sm.fail;   // reason code (1: tie overflow, 2: exhausted, 3: bound)
    return;
  }
  if (tid == 0) need_full[n] = 0;
  const int nsel = sm.nsel;
  const float scale_img = image_scales ? image_scales[n] : 1.f;
  for (int r = tid; r < max_out; r += kFastThreads) {
    const int idx = r < nsel ? sm.sel_idx[r] : 0;
    const float score = r < nsel ? sm.sel_score[r] : 0.f;
    const float4 b = bx[idx];
    float* d = detections + (static_cast<size_t>(n) * max_out + r) * 7;
    d[0] = static_cast<float>(image_id_base + n);
    d[1] = __fmul_rn(fminf(fmaxf(b.x, 0.f), clip_h), scale_img);
    d[2] = __fmul_rn(fminf(fmaxf(b.y, 0.f), clip_w), scale_img);
    d[3] = __fmul_rn(fminf(fmaxf(b.z, 0.f), clip_h), scale_img);
    d[4] = __fmul_rn(fminf(fmaxf(b.w, 0.f), clip_w), scale_img);
    d[5] = score;
    d[6] = static_cast<float>(classes[static_cast<size_t>(n) * k + idx] + 1);
    sel_index[static_cast<size_t>(n) * max_out + r] = idx;
  }
  if (tid == 0) valid[n] = nsel;
}

}  // namespace edet

extern "C" int edet_pre_nms(const edet_half* const* h_cls, const edet_half* const* h_box,
                            const int* h_level_hw, int levels, int ld_cls, int ld_box,
                            int num_anchors, int num_classes, const float* anchors, float* boxes,
                            float* scores, int32_t* classes, int n, edet_stream_t stream) {
  using namespace edet;
  EDET_CHECK_ARG(h_cls && h_box && h_level_hw && anchors && boxes && scores && classes,
                 "pre_nms: null pointer");
  EDET_CHECK_ARG(levels >= 1 && levels <= kPreMaxLevels, "pre_nms: 1..8 levels");
  EDET_CHECK_ARG(ld_cls % 8 == 0 && ld_cls >= num_anchors * num_classes && ld_box % 4 == 0 &&
                     ld_box >= num_anchors * 4,
                 "pre_nms: bad leading dims (ld_cls=%d ld_box=%d)", ld_cls, ld_box);
  EDET_CHECK_ARG(kPrePix * num_anchors <= kPreThreads, "pre_nms: too many anchors per location");
  PreParams p;
  p.levels = levels; p.ld_cls = ld_cls; p.ld_box = ld_box;
  p.num_anchors = num_anchors; p.num_classes = num_classes;
  int blocks = 0, anchors_total = 0;
  for (int l = 0; l < levels; ++l) {
    PreLevel& lv = p.lv[l];
    lv.cls = reinterpret_cast<const __half*>(h_cls[l]);
    lv.box = reinterpret_cast<const __half*>(h_box[l]);
    EDET_CHECK_ARG(lv.cls && lv.box, "pre_nms: level %d pointer is null", l);
    lv.pixels = h_level_hw[2 * l] * h_level_hw[2 * l + 1];
    lv.block_begin = blocks;
    lv.anchor_begin = anchors_total;
    blocks += ceil_div(lv.pixels, kPrePix);
    anchors_total += lv.pixels * num_anchors;
  }
  p.total_anchors = anchors_total;
  const size_t smem = static_cast<size_t>(kPrePix) * ld_cls * sizeof(__half);
  EDET_CHECK_ARG(smem <= 96 * 1024, "pre_nms: ld_cls too large");
  if (smem > 48 * 1024)
    EDET_CHECK_CUDA(cudaFuncSetAttribute(pre_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
  pre_nms_kernel<<<dim3(blocks, n), kPreThreads, smem, as_stream(stream)>>>(p, anchors, boxes,
                                                                            scores, classes);
  EDET_CHECK_LAUNCH();
  return EDET_OK;
}

extern "C" size_t edet_nms_work_bytes(int n, int k) {
  return static_cast<size_t>(n) * k * (sizeof(float) + sizeof(int32_t)) + static_cast<size_t>(n) * sizeof(int32_t);
}

extern "C" int edet_nms_v5(const float* boxes, const float* scores, const int32_t* classes,
                           const float* image_scales, int image_id_base, int n, int k,
                           int max_output_size, float iou_threshold, float score_threshold,
                           float soft_nms_sigma, float clip_h, float clip_w, float* detections,
                           int32_t* sel_index, int32_t* valid, void* work, edet_stream_t stream) {
  using namespace edet;
  EDET_CHECK_ARG(boxes && scores && classes && detections && sel_index && valid && work,
                 "nms_v5: null pointer");
  EDET_CHECK_ARG(n > 0 && k > 0, "nms_v5: bad shape");
  EDET_CHECK_ARG(max_output_size > 0 && max_output_size <= kNmsMaxOut,
                 "nms_v5: max_output_size must be in 1..%d", kNmsMaxOut);
  float* ws = reinterpret_cast<float*>(work);
  int32_t* wb = reinterpret_cast<int32_t*>(ws + static_cast<size_t>(n) * k);
  int32_t* need_full = wb + static_cast<size_t>(n) * k;
  static bool configured = false;
  if (!configured) {
    EDET_CHECK_CUDA(cudaFuncSetAttribute(nms_v5_fast_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(sizeof(FastSmem))));
    configured = true;
  }
  // fast path (top candidates in shared memory, exactness proven per image) ...
  nms_v5_fast_kernel<<<n, kFastThreads, sizeof(FastSmem), as_stream(stream)>>>(
      boxes, scores, classes, image_scales, image_id_base, k, max_output_size, iou_threshold,
      score_threshold, soft_nms_sigma, clip_h, clip_w, detections, sel_index, valid, need_full);
  EDET_CHECK_LAUNCH();
  // ... and the full-queue kernel, which returns immediately for images the fast path settled.
  nms_v5_kernel<<<n, kNmsThreads, 0, as_stream(stream)>>>(
      boxes, scores, classes, image_scales, image_id_base, k, max_output_size, iou_threshold,
      score_threshold, soft_nms_sigma, clip_h, clip_w, detections, sel_index, valid, ws, wb,
      need_full);
  EDET_CHECK_LAUNCH();
  return EDET_OK;
}
