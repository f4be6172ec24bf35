// CUDA replacement for nms_np.per_class_nms (hard / DIoU NMS here, the soft methods further down), the numpy post-process the
// reference calls through tf.numpy_function (tf2/postprocess.py:541-556, nms_np.py:28-126,
// 220-264): per class, greedy NMS in descending score order with the "+1 pixel" float32 IoU, all
// survivors of all classes merged, the top max_boxes by score emitted as
// [image_id, x1, y1, x2, y2, score, class + 1] (x scale), padded with dummy rows of score -1e5.
//
// The per-class greedy loops and the final global sort are restated as ONE scan in globally
// descending score order: a candidate survives iff no already-kept box OF ITS CLASS has
// metric > iou_thresh.  (Restricted to one class this is exactly nms_np's loop, and the first
// max_boxes survivors of the global scan are the top max_boxes of the merged list.)  The scan
// stops at max_boxes survivors, so only the highest-scoring candidates are ever touched:
//   rounds of  radix-select of the next <= 2048 highest (score, index) keys  ->  bitonic sort in
//   shared memory  ->  gather their boxes  ->  warp scan in chunks of 32 (each lane tests its
//   candidate against the kept list, then the chunk is resolved lane by lane with shuffles).
// All arithmetic is float32 with explicit round-to-nearest intrinsics (no FMA contraction), in
// nms_np's order of operations, so keep decisions are bit-identical to NumPy's.
// Tie order of EQUAL scores: numpy's argsort is not stable (implementation defined); here the
// higher index comes first (what a stable argsort followed by [::-1] gives).
#include "common.cuh"

namespace edet {
namespace pcn {

constexpr int kThreads = 512;
constexpr int kRound = 2048;     // candidates per round
constexpr int kMaxKeep = 256;    // max_boxes_to_draw supported
constexpr float kDummyScore = -1e5f;   // nms_np._DUMMY_DETECTION_SCORE

struct Smem {
  unsigned long long keys[kRound];
  float4 cbox[kRound];            // x1, y1, x2, y2 of the sorted candidates
  int ccls[kRound];
  float4 kbox[kMaxKeep];
  float karea[kMaxKeep];
  float kscore[kMaxKeep];
  int kcls[kMaxKeep];
  int kidx[kMaxKeep];
  unsigned hist[256];
  int count;
  int nkept;
  unsigned long long sel_prefix;
  int sel_rank;
};

// larger score -> larger key; equal scores -> larger index first
__device__ __forceinline__ unsigned long long make_key(float s, int idx) {
  unsigned u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned>(idx);
}
__device__ __forceinline__ float key_score(unsigned long long k) {
  unsigned u = static_cast<unsigned>(k >> 32);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}

__device__ __forceinline__ float area_plus1(const float4& b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

// nms_np.py:61-83 / 113-121, float32, `i` = the kept box, `r` = the candidate
template <bool DIOU>
__device__ __forceinline__ float metric(const float4& bi, float area_i, const float4& br,
                                        float area_r) {
  const float xx1 = fmaxf(bi.x, br.x), yy1 = fmaxf(bi.y, br.y);
  const float xx2 = fminf(bi.z, br.z), yy2 = fminf(bi.w, br.w);
  const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
  const float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
  const float inter = __fmul_rn(w, h);
  float m = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_i, area_r), inter));
  if (DIOU) {
    const float ex1 = fminf(bi.x, br.x), ex2 = fmaxf(bi.z, br.z);
    const float ey1 = fminf(bi.y, br.y), ey2 = fmaxf(bi.w, br.w);
    const float dx = __fsub_rn(ex2, ex1), dy = __fsub_rn(ey2, ey1);
    const float diag = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const float cxi = __fdiv_rn(__fadd_rn(bi.x, bi.z), 2.0f), cyi = __fdiv_rn(__fadd_rn(bi.y, bi.w), 2.0f);
    const float cxr = __fdiv_rn(__fadd_rn(br.x, br.z), 2.0f), cyr = __fdiv_rn(__fadd_rn(br.y, br.w), 2.0f);
    const float ddx = __fsub_rn(cxi, cxr), ddy = __fsub_rn(cyi, cyr);
    const float cd = __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
    m = __fsub_rn(m, __fdiv_rn(cd, __fadd_rn(diag, 1e-10f)));
  }
  return m;
}

template <bool DIOU>
__global__ void __launch_bounds__(kThreads)
per_class_nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                     const int32_t* __restrict__ classes, const float* __restrict__ image_ids,
                     const float* __restrict__ image_scales, int k, int num_classes, int max_boxes,
                     float iou_thresh, float* __restrict__ detections,
                     int32_t* __restrict__ keep_index, int32_t* __restrict__ num_valid) {
  extern __shared__ __align__(16) uint8_t pcn_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(pcn_raw);
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + static_cast<size_t>(n) * k;
  const float* sc = scores + static_cast<size_t>(n) * k;
  const int32_t* cl = classes + static_cast<size_t>(n) * k;
  if (tid == 0) sm.nkept = 0;
  __syncthreads();

  // candidates = anchors whose class is a real class (nms_np loops c in range(num_classes))
  auto valid = [&](int i) { const int c = cl[i]; return c >= 0 && c < num_classes; };

  unsigned long long upper = ~0ull;   // exclusive upper bound of the keys not yet processed
  bool first = true;
  while (true) {
    // ---- how many candidates are left below `upper` ----
    if (tid == 0) sm.count = 0;
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < k; i += kThreads)
      if (valid(i) && (first || make_key(sc[i], i) < upper)) ++mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0 && mine) atomicAdd(&sm.count, mine);
    __syncthreads();
    const int remaining = sm.count;
    __syncthreads();
    if (remaining == 0) break;
    // ---- threshold = the kRound-th largest remaining key (radix select, 8 bits per pass) ----
    unsigned long long thr = 0ull;
    if (remaining > kRound) {
      if (tid == 0) { sm.sel_prefix = 0ull; sm.sel_rank = kRound; }
      for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += kThreads) sm.hist[i] = 0u;
        __syncthreads();
        const unsigned long long prefix = sm.sel_prefix;
        const unsigned long long hi_mask = shift == 56 ? 0ull : (~0ull << (shift + 8));
        for (int i = tid; i < k; i += kThreads) {
          if (!valid(i)) continue;
          const unsigned long long key = make_key(sc[i], i);
          if ((first || key < upper) && (key & hi_mask) == prefix)
            atomicAdd(&sm.hist[(key >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (tid == 0) {
          int rank = sm.sel_rank;   // the rank-th largest key among the matching ones
          int d = 255;
          for (; d > 0; --d) {
            const int cnt = static_cast<int>(sm.hist[d]);
            if (rank <= cnt) break;
            rank -= cnt;
          }
          sm.sel_prefix = prefix | (static_cast<unsigned long long>(d) << shift);
          sm.sel_rank = rank;
        }
        __syncthreads();
      }
      thr = sm.sel_prefix;   // keys are unique: exactly kRound remaining keys are >= thr
    }
    // ---- compaction + sort (descending) ----
    if (tid == 0) sm.count = 0;
    for (int i = tid; i < kRound; i += kThreads) sm.keys[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < k; i += kThreads) {
      if (!valid(i)) continue;
      const unsigned long long key = make_key(sc[i], i);
      if ((first || key < upper) && key >= thr) {
        const int pos = atomicAdd(&sm.count, 1);
        if (pos < kRound) sm.keys[pos] = key;
      }
    }
    __syncthreads();
    const int m = min(sm.count, kRound);
    for (int size = 2; size <= kRound; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = tid; i < kRound / 2; i += kThreads) {
          const int lo = 2 * i - (i & (stride - 1));
          const int hi = lo + stride;
          const bool desc = ((lo & size) == 0);
          const unsigned long long a = sm.keys[lo], b = sm.keys[hi];
          if ((a < b) == desc) { sm.keys[lo] = b; sm.keys[hi] = a; }
        }
        __syncthreads();
      }
    }
    // ---- gather the sorted candidates ([y1,x1,y2,x2] -> [x1,y1,x2,y2], nms_np.py:223) ----
    for (int i = tid; i < m; i += kThreads) {
      const int idx = static_cast<int>(sm.keys[i] & 0xffffffffu);
      const float4 b = bx[idx];
      sm.cbox[i] = make_float4(b.y, b.x, b.w, b.z);
      sm.ccls[i] = cl[idx];
    }
    __syncthreads();
    // ---- scan (warp 0): 32 candidates at a time ----
    if (tid < 32) {
      int nkept = sm.nkept;
      for (int base = 0; base < m && nkept < max_boxes; base += 32) {
        const int ci = base + lane;
        const bool have = ci < m;
        const float4 b = have ? sm.cbox[ci] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = have ? sm.ccls[ci] : -1;
        const float area = area_plus1(b);
        bool alive = have;
        for (int j = 0; j < nkept && alive; ++j) {
          if (sm.kcls[j] == c) {
            const float mt = metric<DIOU>(sm.kbox[j], sm.karea[j], b, area);
            if (!(mt <= iou_thresh)) alive = false;
          }
        }
        for (int i = 0; i < 32; ++i) {
          const bool kept_i = __shfl_sync(0xffffffffu, alive ? 1 : 0, i) != 0;   // final for lane i
          if (!kept_i) continue;
          if (nkept >= max_boxes) { if (lane >= i) alive = false; continue; }
          const float4 bi = make_float4(__shfl_sync(0xffffffffu, b.x, i), __shfl_sync(0xffffffffu, b.y, i),
                                        __shfl_sync(0xffffffffu, b.z, i), __shfl_sync(0xffffffffu, b.w, i));
          const float ai = __shfl_sync(0xffffffffu, area, i);
          const int cls_i = __shfl_sync(0xffffffffu, c, i);
          if (lane == i) {
            sm.kbox[nkept] = b;
            sm.karea[nkept] = area;
            sm.kcls[nkept] = c;
            const unsigned long long key = sm.keys[ci];
            sm.kidx[nkept] = static_cast<int>(key & 0xffffffffu);
            sm.kscore[nkept] = key_score(key);
          }
          ++nkept;
          if (lane > i && alive && c == cls_i) {
            const float mt = metric<DIOU>(bi, ai, b, area);
            if (!(mt <= iou_thresh)) alive = false;
          }
        }
        __syncwarp();
      }
      if (lane == 0) sm.nkept = nkept;
    }
    __syncthreads();
    if (sm.nkept >= max_boxes || remaining <= kRound) break;
    upper = thr;
    first = false;
  }
  __syncthreads();
  // ---- rows: [image_id, x1, y1, x2, y2, score, class + 1], boxes x image_scale; dummies after ----
  const int nkept = min(sm.nkept, max_boxes);
  const float id = image_ids ? image_ids[n] : static_cast<float>(n);
  const float scale = image_scales ? image_scales[n] : 1.0f;
  for (int i = tid; i < max_boxes; i += kThreads) {
    float* d = detections + (static_cast<size_t>(n) * max_boxes + i) * 7;
    if (i < nkept) {
      const float4 b = sm.kbox[i];
      d[0] = id;
      d[1] = __fmul_rn(b.x, scale); d[2] = __fmul_rn(b.y, scale);
      d[3] = __fmul_rn(b.z, scale); d[4] = __fmul_rn(b.w, scale);
      d[5] = sm.kscore[i];
      d[6] = static_cast<float>(sm.kcls[i] + 1);
      keep_index[static_cast<size_t>(n) * max_boxes + i] = sm.kidx[i];
    } else {
      d[0] = id;
      d[1] = __fmul_rn(0.f, scale); d[2] = d[1]; d[3] = d[1]; d[4] = d[1];
      d[5] = kDummyScore;
      d[6] = 0.f;
      keep_index[static_cast<size_t>(n) * max_boxes + i] = -1;
    }
  }
  if (tid == 0) num_valid[n] = nkept;
}

// ---- soft NMS (nms_np.py:129-191: gaussian / linear) -------------------------------------------
// nms_np's per-class loop -- arg-max of the decayed scores, decay every other box of the class by
// exp(-iou^2 / sigma) (gaussian) or 1 - iou where iou > iou_thresh (linear), drop boxes whose
// score falls below score_thresh -- emits each class's boxes in non-increasing score order, and
// classes do not interact.  The merged top max_boxes are therefore the first max_boxes GLOBAL
// arg-max selections: one pass over the K current scores per selection (decay of the winner's
// class fused with the arg-max for the next selection), max_boxes selections in all.
//   cur: float32 [n][k] workspace holding the current (decayed) scores, -inf = removed.
// Equal scores: the lower anchor index is selected first.  gaussian: exp is evaluated in double
// and rounded to float32; NumPy's SIMD float32 exp is within 2 ulp of that, so scores can differ
// from a given NumPy build in the last bits (the linear method is bit-identical).
constexpr int kSoftThreads = 512;

struct SoftSmem {
  float4 kbox[kMaxKeep];
  float kscore[kMaxKeep];
  int kcls[kMaxKeep];
  int kidx[kMaxKeep];
  float red_v[kSoftThreads / 32];
  int red_i[kSoftThreads / 32];
  float win_v;
  int win_i;
};

template <int METHOD>
__global__ void __launch_bounds__(kSoftThreads)
per_class_soft_nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                          const int32_t* __restrict__ classes, const float* __restrict__ image_ids,
                          const float* __restrict__ image_scales, int k, int num_classes,
                          int max_boxes, float iou_thresh, float sigma, float score_thresh,
                          float* __restrict__ cur_all, float* __restrict__ detections,
                          int32_t* __restrict__ keep_index, int32_t* __restrict__ num_valid) {
  __shared__ SoftSmem sm;
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + static_cast<size_t>(n) * k;
  const float* sc = scores + static_cast<size_t>(n) * k;
  const int32_t* cl = classes + static_cast<size_t>(n) * k;
  float* cur = cur_all + static_cast<size_t>(n) * k;
  const float kNegInf = __int_as_float(0xff800000);

  float4 wbox = make_float4(0.f, 0.f, 0.f, 0.f);
  float warea = 0.f;
  int wcls = -1, widx = -1, nkept = 0;
  for (int t = 0; t < max_boxes; ++t) {
    // one pass: (t > 0) decay the class of the previous winner; always: arg-max of what remains.
    // Element i is always handled by thread i % kSoftThreads, so `cur` needs no global ordering.
    float best_v = kNegInf;
    int best_i = 0x7fffffff;
    for (int i = tid; i < k; i += kSoftThreads) {
      float v;
      if (t == 0) {
        const int c = cl[i];
        v = (c >= 0 && c < num_classes) ? sc[i] : kNegInf;
        cur[i] = v;
      } else {
        v = cur[i];
        if (i == widx) {
          v = kNegInf;
          cur[i] = v;
        } else if (v != kNegInf && cl[i] == wcls) {
          const float4 r = bx[i];
          const float4 b = make_float4(r.y, r.x, r.w, r.z);
          const float iou = metric<false>(wbox, warea, b, area_plus1(b));
          float wgt = 1.0f;
          if (METHOD == EDET_NMS_LINEAR) {
            if (iou > iou_thresh) wgt = __fsub_rn(1.0f, iou);
          } else {
            const float e = -__fdiv_rn(__fmul_rn(iou, iou), sigma);
            wgt = static_cast<float>(exp(static_cast<double>(e)));
          }
          v = __fmul_rn(v, wgt);
          if (!(v >= score_thresh)) v = kNegInf;
          cur[i] = v;
        }
      }
      if (v > best_v) { best_v = v; best_i = i; }   // ascending i: ties keep the lower index
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
    }
    if (lane == 0) { sm.red_v[warp] = best_v; sm.red_i[warp] = best_i; }
    __syncthreads();
    if (warp == 0) {
      float v = lane < kSoftThreads / 32 ? sm.red_v[lane] : kNegInf;
      int i = lane < kSoftThreads / 32 ? sm.red_i[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
      }
      if (lane == 0) { sm.win_v = v; sm.win_i = i; }
    }
    __syncthreads();
    const float wv = sm.win_v;
    widx = sm.win_i;
    if (wv == kNegInf) break;     // nothing left
    const float4 r = bx[widx];
    wbox = make_float4(r.y, r.x, r.w, r.z);
    warea = area_plus1(wbox);
    wcls = cl[widx];
    if (tid == 0) {
      sm.kbox[nkept] = wbox;
      sm.kscore[nkept] = wv;
      sm.kcls[nkept] = wcls;
      sm.kidx[nkept] = widx;
    }
    ++nkept;
    __syncthreads();              // red_v / win_* are rewritten by the next selection
  }
  __syncthreads();
  const float id = image_ids ? image_ids[n] : static_cast<float>(n);
  const float scale = image_scales ? image_scales[n] : 1.0f;
  for (int i = tid; i < max_boxes; i += kSoftThreads) {
    float* d = detections + (static_cast<size_t>(n) * max_boxes + i) * 7;
    if (i < nkept) {
      const float4 b = sm.kbox[i];
      d[0] = id;
      d[1] = __fmul_rn(b.x, scale); d[2] = __fmul_rn(b.y, scale);
      d[3] = __fmul_rn(b.z, scale); d[4] = __fmul_rn(b.w, scale);
      d[5] = sm.kscore[i];
      d[6] = static_cast<float>(sm.kcls[i] + 1);
      keep_index[static_cast<size_t>(n) * max_boxes + i] = sm.kidx[i];
    } else {
      d[0] = id;
      d[1] = __fmul_rn(0.f, scale); d[2] = d[1]; d[3] = d[1]; d[4] = d[1];
      d[5] = kDummyScore;
      d[6] = 0.f;
      keep_index[static_cast<size_t>(n) * max_boxes + i] = -1;
    }
  }
  if (tid == 0) num_valid[n] = nkept;
}

}  // namespace pcn
}  // namespace edet

extern "C" int edet_per_class_nms(const float* boxes, const float* scores, const int32_t* classes,
                                  const float* image_ids, const float* image_scales, int n, int k,
                                  int num_classes, int max_boxes_to_draw, int method,
                                  float iou_thresh, float sigma, float score_thresh, float* work,
                                  float* detections, int32_t* keep_index, int32_t* num_valid,
                                  edet_stream_t stream) {
  using namespace edet;
  using namespace edet::pcn;
  EDET_CHECK_ARG(boxes && scores && classes && detections && keep_index && num_valid,
                 "per_class_nms: null pointer");
  EDET_CHECK_ARG(n > 0 && k >= 0 && num_classes > 0, "per_class_nms: bad shape");
  EDET_CHECK_ARG(max_boxes_to_draw > 0 && max_boxes_to_draw <= kMaxKeep,
                 "per_class_nms: max_boxes_to_draw must be in 1..%d (got %d)", kMaxKeep,
                 max_boxes_to_draw);
  cudaStream_t s = as_stream(stream);
  if (method == EDET_NMS_GAUSSIAN || method == EDET_NMS_LINEAR) {
    EDET_CHECK_ARG(work != nullptr, "per_class_nms: the soft methods need the [n][k] float workspace");
    EDET_CHECK_ARG(sigma > 0.f, "per_class_nms: sigma must be positive");
    if (method == EDET_NMS_GAUSSIAN)
      per_class_soft_nms_kernel<EDET_NMS_GAUSSIAN><<<n, kSoftThreads, 0, s>>>(
          boxes, scores, classes, image_ids, image_scales, k, num_classes, max_boxes_to_draw,
          iou_thresh, sigma, score_thresh, work, detections, keep_index, num_valid);
    else
      per_class_soft_nms_kernel<EDET_NMS_LINEAR><<<n, kSoftThreads, 0, s>>>(
          boxes, scores, classes, image_ids, image_scales, k, num_classes, max_boxes_to_draw,
          iou_thresh, sigma, score_thresh, work, detections, keep_index, num_valid);
    EDET_CHECK_LAUNCH();
    return EDET_OK;
  }
  if (method != EDET_NMS_HARD && method != EDET_NMS_DIOU) {
    set_error("per_class_nms: unknown method %d", method);
    return EDET_ERR_INVALID;
  }
  static int configured[2][kMaxDevices];
  if (int rc = ensure_dynamic_smem(per_class_nms_kernel<false>, static_cast<int>(sizeof(Smem)), configured[0])) return rc;
  if (int rc = ensure_dynamic_smem(per_class_nms_kernel<true>, static_cast<int>(sizeof(Smem)), configured[1])) return rc;
  if (method == EDET_NMS_DIOU)
    per_class_nms_kernel<true><<<n, kThreads, sizeof(Smem), s>>>(
        boxes, scores, classes, image_ids, image_scales, k, num_classes, max_boxes_to_draw,
        iou_thresh, detections, keep_index, num_valid);
  else
    per_class_nms_kernel<false><<<n, kThreads, sizeof(Smem), s>>>(
        boxes, scores, classes, image_ids, image_scales, k, num_classes, max_boxes_to_draw,
        iou_thresh, detections, keep_index, num_valid);
  EDET_CHECK_LAUNCH();
  return EDET_OK;
}
