// Pointwise (1x1) convolution as an NHWC GEMM on the 5th-generation tensor cores (sm_100a).
//
//   out[b, r, n] = act( sum_k A[b, r, k] * Wt[b|0, n, k] + bias[n] ) (+ residual[b, r, n])
//
// Structure (persistent CTAs, two per SM so one CTA's epilogue overlaps the other's loads;
// 192 threads each):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor (3-D maps, 128B swizzle) for the A tile
//                 [128 x 64] and the W tile [block_n x 64] into an N-stage smem ring (mbarrier
//                 complete_tx).  Out-of-bounds rows / K tail are zero-filled by TMA.
//   warp 1      : allocates TMEM, then one lane issues tcgen05.mma (kind::f16, M=128,
//                 N=block_n, K=16) with fp32 accumulators in TMEM, double-buffered across tiles;
//                 tcgen05.commit releases smem stages / publishes the accumulator.
//   warps 2..9  : epilogue -- tcgen05.ld (32x32b) -> +bias -> activation -> (+residual) -> fp16 ->
//                 swizzled st.shared -> TMA store (cp.async.bulk.tensor ... bulk_group), which
//                 also clips the ragged M / N edges.  Two warps share each TMEM lane quarter and
//                 split every 64-column store chunk into two 32-column halves (the swish
//                 epilogue is MUFU / issue bound, so it gets 8 warps and packed fp32 math).
//
// Replaces Conv2D 1x1 (+BN, +swish, +skip) at the reference call sites listed in
// include/automl_b200.h (edet_pointwise_conv).  Algorithmic HBM bytes per launch:
//   2*(batch*rows*k + batch*rows*nout [+ same for residual]) + 2*wbatch*nout*k   (SURVEY 8d).
#include <math_constants.h>

#include "tc_common.cuh"

namespace edet {
namespace pwtc {

// warp 0 TMA, warp 1 MMA, then TEAMS x 4 epilogue warps (one per TMEM lane quarter and team):
//   TEAMS == 2: 320 threads, <= 96 registers; every 64-column store chunk is split between the
//               two teams' staging slabs (4 KiB each)
//   TEAMS == 3: 448 threads, <= 72 registers; the unit of epilogue work is 32 columns (one
//               2 KiB [32 rows x 32 cols] TMA store), handed to the teams round robin -- 50 % more
//               warps to hide the MUFU / dependent-issue latency of the swish epilogue, balanced
//               for the expand widths (N = 96, 144, 240 and tiles of 96)
template <int TEAMS>
struct Epi {
  static constexpr int kWarps = 4 * TEAMS;
  static constexpr int kThreads = 64 + 32 * kWarps;
  static constexpr int kSlabBytes = TEAMS == 2 ? 4096 : 2048;
};
constexpr int kStoreCols = 64;
constexpr int kMaxStages = 8;
constexpr int kRing = 4;          // tile-index ring entries (power of two)
constexpr int kMaxAccum = 4;      // TMEM accumulator stages (thin N tiles)
constexpr int kSmemLimit = 113 * 1024;                    // two CTAs per SM share the 227 KiB
constexpr int kSmemCoResident = 99 * 1024;               // one CTA next to an nms_v5_fast CTA

struct Params {
  int batch, rows, k, nout, nout_pad8;
  int block_n, num_m_blocks, num_n_blocks, num_k_blocks, num_stages;
  int block_k;        // 64 / 32 / 16 halves per k-block == 128B / 64B / 32B swizzled smem rows
  int a_stage_bytes, b_stage_bytes;
  int desc_sbo;       // byte distance between 8-row groups in smem (8 * row pitch)
  int desc_layout;    // UMMA layout type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
  int slabs_per_warp; // private TMA-store staging slabs per epilogue warp (1 or 2)
  int accum_stages;   // TMEM accumulator stages: 1..kMaxAccum, accum_stages * block_n <= 256 columns
  int wbatch, ldr, tmem_cols;
  int total_tiles;
  int bias_floats;    // floats of the zero-padded bias staged in shared memory: whole N tiles
  const float* bias;
  const __half* residual;
  unsigned* sched;    // dynamic tile scheduler slot (tc_common.cuh)
  // EPI_ARGMAX (class head -> pre-NMS): N tile n_blk = anchor n_blk, rows = pixels of one level
  float* am_scores;       // [batch][am_total] sigmoid of the best class logit
  int32_t* am_classes;    // [batch][am_total] its class index
  int am_anchor_begin, am_total, am_num_anchors;
};
constexpr int EPI_STORE = 0, EPI_ARGMAX = 1;
constexpr int kArgmaxCols = 96;   // columns per anchor in the padded class-head weights

struct TileCoord {
  int b, m_blk, n_blk;
};
// (accumulator stage, mbarrier phase) of consecutive tiles without a division per tile
struct StageCounter {
  int as = -1;
  uint32_t phase = 1u;
  __device__ __forceinline__ void next(int stages) {
    if (++as == stages) as = 0;
    if (as == 0) phase ^= 1u;
  }
};
// Biases of 8 consecutive output columns from the copy of the (zero-padded) bias vector that the
// prologue stages in shared memory: two ld.shared.v4 instead of two global loads plus the
// ragged-edge branches in the middle of the MUFU / issue-bound epilogue.
constexpr int kMaxBiasSmem = 8192;
__device__ __forceinline__ void load_bias8(uint32_t smem_bias_u32, int col, float4& b0, float4& b1) {
  const uint32_t a = smem_bias_u32 + static_cast<uint32_t>(col) * 4u;
  // not volatile: the staged bias is constant for the kernel, so the loads may move freely
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w) : "r"(a));
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(b1.x), "=f"(b1.y), "=f"(b1.z), "=f"(b1.w) : "r"(a + 16u));
}

__device__ __forceinline__ TileCoord decode_tile(int t, const Params& p) {
  TileCoord c;
  c.n_blk = 0;
  if (p.num_n_blocks != 1) {       // (uniform) most layers are one N tile wide ...
    c.n_blk = t % p.num_n_blocks;
    t /= p.num_n_blocks;
  }
  c.m_blk = t;
  c.b = 0;
  if (p.batch != 1) {              // ... and one batch entry long: no division at all
    c.m_blk = t % p.num_m_blocks;
    c.b = t / p.num_m_blocks;
  }
  return c;
}

template <int ACT, bool HAS_RES, int TEAMS, int EPI>
__global__ void __launch_bounds__(Epi<TEAMS>::kThreads, 2)
pointwise_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                    const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_o, const Params p) {
  pdl_launch_dependents();   // the next kernel may start its prologue while this one runs
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for the swizzle atoms.
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  uint8_t* smem_store = smem + p.num_stages * stage_bytes;
  constexpr int kEpiWarps = Epi<TEAMS>::kWarps;
  constexpr int kSlabBytes = Epi<TEAMS>::kSlabBytes;
  float* smem_bias = reinterpret_cast<float*>(smem_store + p.slabs_per_warp * kEpiWarps * kSlabBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bias + p.bias_floats);
  const uint32_t smem_bias_u32 = smem_u32(smem_bias);
  uint64_t* full_bar = bars;                       // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;         // [kMaxStages]
  uint64_t* tmem_full_bar = bars + 2 * kMaxStages;                 // [kMaxAccum]
  uint64_t* tmem_empty_bar = bars + 2 * kMaxStages + kMaxAccum;    // [kMaxAccum]
  // Tile indices travel from the producer to the MMA warp and the epilogue warps through a small
  // ring: CTA i owns tile i, every further tile comes from a global counter (sched_next_tile), so
  // a CTA that gets its SM late -- another stream's kernel, e.g. the NMS of the previous batch,
  // was holding it -- simply finds less work instead of owning a full static share.
  uint64_t* ring_full = bars + 2 * kMaxStages + 2 * kMaxAccum;   // [kRing] producer -> consumers
  uint64_t* ring_empty = ring_full + kRing;              // [kRing] consumers -> producer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ring_empty + kRing);
  // [kRing] x {tile, batch entry, M block, N block}: the producer decodes each tile once (its two
  // integer divisions) and the 13 consumer warps read the coordinates with one 16-byte load
  volatile int4* tile_ring = reinterpret_cast<volatile int4*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < kMaxAccum; ++s) {
      mbar_init(smem_u32(&tmem_full_bar[s]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[s]), kEpiWarps);  // one arrive per epilogue warp
    }
    for (int s = 0; s < kRing; ++s) {
      mbar_init(smem_u32(&ring_full[s]), 1);
      mbar_init(smem_u32(&ring_empty[s]), 1 + kEpiWarps);  // the MMA warp + every epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_o)) : "memory");
  }
  // bias (a constant, like the weights: read before the PDL wait) -> shared memory, zero padded
  for (int i = threadIdx.x; i < p.bias_floats; i += blockDim.x)
    smem_bias[i] = i < p.nout ? __ldg(p.bias + i) : 0.f;
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait_prior();          // everything above overlapped the previous kernel's tail

  // consumer side of the tile ring: whole warp (or the single MMA lane) waits, reads, releases
  auto ring_get = [&](int i, bool whole_warp, TileCoord* tc) -> int {
    const int slot = i & (kRing - 1);
    mbar_wait(smem_u32(&ring_full[slot]), static_cast<uint32_t>(i / kRing) & 1u);
    const int4 e = const_cast<const int4*>(tile_ring)[slot];
    if (whole_warp) __syncwarp();
    if (!whole_warp || lane == 0) mbar_arrive(smem_u32(&ring_empty[slot]));
    if (tc) {
      tc->b = e.y;
      tc->m_blk = e.z;
      tc->n_blk = e.w;
    }
    return e.x;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // bytes the two TMA boxes deliver (the B slot may be padded to 1 KiB)
      const uint32_t tx_bytes = static_cast<uint32_t>(p.a_stage_bytes + p.block_n * p.block_k * 2);
      bool exhausted = false;
      for (int i = 0;; ++i) {
        int t = p.total_tiles;
        if (i == 0) {
          t = blockIdx.x;
        } else if (!exhausted) {
          t = sched_next_tile(p.sched, p.total_tiles);
          exhausted = t >= p.total_tiles;
        }
        const int slot = i & (kRing - 1);
        mbar_wait(smem_u32(&ring_empty[slot]), (static_cast<uint32_t>(i / kRing) & 1u) ^ 1u);
        TileCoord tc;
        tc.b = tc.m_blk = tc.n_blk = 0;
        if (t < p.total_tiles) tc = decode_tile(t, p);
        const_cast<int4*>(tile_ring)[slot] = make_int4(t, tc.b, tc.m_blk, tc.n_blk);
        mbar_arrive(smem_u32(&ring_full[slot]));     // release: the entry is visible to waiters
        if (t >= p.total_tiles) break;
        const int wb = (p.wbatch > 1) ? tc.b : 0;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, tx_bytes);
          uint8_t* sa = smem + stage * stage_bytes;
          tma_load_3d(smem_u32(sa), &map_a, fb, kb * p.block_k, tc.m_blk * BLOCK_M, tc.b);
          tma_load_3d(smem_u32(sa + p.a_stage_bytes), &map_w, fb, kb * p.block_k,
                      tc.n_blk * p.block_n, wb);
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b F16 (0), K-major both,
      // n_dim = N>>3 at bit 17, m_dim = M>>4 at bit 24.
      const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(p.block_n >> 3) << 17) |
                             (static_cast<uint32_t>(BLOCK_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      StageCounter sc;
      for (int iter = 0;; ++iter) {
        if (ring_get(iter, false, nullptr) >= p.total_tiles) break;
        sc.next(p.accum_stages);
        const int as = sc.as;
        const uint32_t aphase = sc.phase;
        mbar_wait(smem_u32(&tmem_empty_bar[as]), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * p.block_n);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          uint8_t* sa = smem + stage * stage_bytes;
          const uint64_t da = make_smem_desc(smem_u32(sa), p.desc_sbo, p.desc_layout);
          const uint64_t db = make_smem_desc(smem_u32(sa + p.a_stage_bytes), p.desc_sbo, p.desc_layout);
          const int k_rem = p.k - kb * p.block_k;
          const int ksteps = k_rem >= p.block_k ? p.block_k / UMMA_K : (k_rem + UMMA_K - 1) / UMMA_K;
          for (int ks = 0; ks < ksteps; ++ks) {
            // advance 16 halves = 32 bytes inside the swizzle atom: +2 in the >>4 address field
            tc_mma_f16(tmem_d, da + static_cast<uint64_t>(ks * 2), db + static_cast<uint64_t>(ks * 2),
                       idesc, (kb > 0 || ks > 0) ? 1u : 0u);
          }
          tc_commit(smem_u32(&empty_bar[stage]));  // frees the smem stage when the MMAs retire
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(smem_u32(&tmem_full_bar[as]));  // accumulator ready for the epilogue
      }
    }
  } else {
    // ===================== Epilogue =====================
    if constexpr (TEAMS == 2) {
    // Fully decoupled warps: warp (quarter, team) owns rows quarter*32..+31 of the tile and the
    // 64-column store chunks c == team (mod 2).  Each warp has a private 4 KiB staging slab and
    // issues its own TMA stores ([32 rows x 64 cols] boxes), so the only synchronisation in the
    // epilogue is the TMEM full/empty handshake with the MMA warp.
    const int e_warp = warp - 2;              // 0..7
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int team = e_warp >> 2;             // even / odd store chunks
    const int row_in_tile = quarter * 32 + lane;
    uint8_t* my_slabs = smem_store + e_warp * p.slabs_per_warp * (32 * 128);
    int store_cnt = 0;
    StageCounter sc;
    for (int iter = 0;; ++iter) {
      TileCoord tc;
      const int t = ring_get(iter, true, &tc);
      if (t >= p.total_tiles) break;
      sc.next(p.accum_stages);
      const int as = sc.as;
      const uint32_t aphase = sc.phase;
      const int n0 = tc.n_blk * p.block_n;
      const int row = tc.m_blk * BLOCK_M + row_in_tile;
      const bool row_ok = row < p.rows;
      const __half* res_row = nullptr;
      if (HAS_RES) {
        res_row = p.residual + (static_cast<size_t>(tc.b) * p.rows + (row_ok ? row : 0)) * p.ldr;
      }
      mbar_wait(smem_u32(&tmem_full_bar[as]), aphase);
      tc_fence_after();
      // only the columns that exist in the output are worth an epilogue (rounded up to the
      // 16-column TMEM load granule); the rest of a ragged last N tile is skipped
      const int n_valid = min(p.block_n, ((p.nout - n0 + 15) >> 4) << 4);
      const int num_chunks = (n_valid + kStoreCols - 1) / kStoreCols;
      // last chunk this warp reads from TMEM (then the accumulator can be handed back)
      int my_last = -1;
      for (int c = team; c < num_chunks; c += 2) my_last = c;
      if (my_last < 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[as]));
      }
      for (int c = team; c < num_chunks; c += 2) {
        const int cols = min(kStoreCols, n_valid - c * kStoreCols);  // multiple of 16
        uint8_t* my_stage = my_slabs + (p.slabs_per_warp == 2 ? (store_cnt & 1) * 4096 : 0);
        ++store_cnt;
        if (lane == 0) {                           // the store that last used this slab has left it
          if (p.slabs_per_warp == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
        }
        __syncwarp();
        uint8_t* row_base = my_stage + lane * 128;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int c_lo = hf * 32;
          if (c_lo >= cols) break;
          float v[32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                 static_cast<uint32_t>(as * p.block_n + c * kStoreCols + c_lo);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (c_lo + g * 16 < cols) tc_ld16(taddr + g * 16, v + g * 16);
          }
          tc_wait_ld();
          if (c == my_last && (hf == 1 || c_lo + 32 >= cols)) {
            // all TMEM reads of this warp for this accumulator are done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[as]));
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            if (c_lo + jj * 8 < cols) {
              const int col = n0 + c * kStoreCols + c_lo + jj * 8;
              float4 b0, b1;
              load_bias8(smem_bias_u32, col, b0, b1);
              float2 o2[4];
              o2[0] = __fadd2_rn(make_float2(v[jj * 8 + 0], v[jj * 8 + 1]), make_float2(b0.x, b0.y));
              o2[1] = __fadd2_rn(make_float2(v[jj * 8 + 2], v[jj * 8 + 3]), make_float2(b0.z, b0.w));
              o2[2] = __fadd2_rn(make_float2(v[jj * 8 + 4], v[jj * 8 + 5]), make_float2(b1.x, b1.y));
              o2[3] = __fadd2_rn(make_float2(v[jj * 8 + 6], v[jj * 8 + 7]), make_float2(b1.z, b1.w));
              apply_act4<ACT>(o2[0], o2[1]);
              apply_act4<ACT>(o2[2], o2[3]);
              float o[8] = {o2[0].x, o2[0].y, o2[1].x, o2[1].y, o2[2].x, o2[2].y, o2[3].x, o2[3].y};
              if (HAS_RES) {
                if (row_ok && col < p.nout) {
                  float r[8];
                  half8_to_float(ldg_nc_v4(res_row + col), r);
#pragma unroll
                  for (int e = 0; e < 8; ++e) o[e] += r[e];
                }
              }
              const uint4 packed = float_to_half8(o);
              const int chunk16 = hf * 4 + jj;   // 16-byte piece inside the 128-byte row
              *reinterpret_cast<uint4*>(row_base + ((chunk16 ^ (lane & 7)) << 4)) = packed;
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&map_o, smem_u32(my_stage), n0 + c * kStoreCols,
                       tc.m_blk * BLOCK_M + quarter * 32, tc.b);
          tma_store_commit();
        }
      }
    }
    } else if constexpr (EPI == EPI_ARGMAX) {
      // Class head fused with the class half of pre-NMS (tf2/postprocess.py:88-156 with
      // max_nms_inputs == 0): an N tile is ONE anchor (90 class columns + 6 pad columns whose
      // staged bias is -inf); the team (iter % TEAMS) owns the whole tile, lane = pixel.  Each
      // logit is rounded to fp16 exactly as the storing epilogue would store it, then max / first
      // arg-max / sigmoid as pre_nms_kernel does -> bit-identical scores and classes, without the
      // [N, H, W, 810] logits ever reaching HBM.
      const int e_warp = warp - 2;
      const int quarter = warp & 3;
      const int team = e_warp >> 2;
      const int row_in_tile = quarter * 32 + lane;
      StageCounter sc;
      int rot = TEAMS - 1;                      // iter % TEAMS, kept incrementally
      for (int iter = 0;; ++iter) {
        TileCoord tc;
        const int t = ring_get(iter, true, &tc);
        if (t >= p.total_tiles) break;
        sc.next(p.accum_stages);
        if (++rot == TEAMS) rot = 0;
        const int as = sc.as;
        const uint32_t aphase = sc.phase;
        mbar_wait(smem_u32(&tmem_full_bar[as]), aphase);
        tc_fence_after();
        if (rot != team) {                       // not this team's tile: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[as]));
          continue;
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(as * p.block_n);
        const uint32_t bias_u32 = smem_bias_u32 + static_cast<uint32_t>(tc.n_blk * kArgmaxCols) * 4u;
        // Running maximum over 32-bit keys (fp16 logit mapped to an order-preserving uint16) << 16 |
        // (0xFFFF - column): one unsigned max per column pair gives the maximum AND its first
        // column (equal logits: the lower column has the larger key), with no serial
        // compare / select chain.  -0 is canonicalised to +0 first (the float compare of the
        // stored-logits path treats them as equal).
        uint32_t best_key = 0u;
#pragma unroll
        for (int g = 0; g < kArgmaxCols / 16; ++g) {
          float v[16];
          tc_ld16(taddr + g * 16, v);
          tc_wait_ld();
          if (g == kArgmaxCols / 16 - 1) {       // all TMEM reads of this warp are done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[as]));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 b;
            asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                : "r"(bias_u32 + static_cast<uint32_t>(g * 16 + q * 4) * 4u));
            const float2 s01 = __fadd2_rn(make_float2(v[q * 4 + 0], v[q * 4 + 1]), make_float2(b.x, b.y));
            const float2 s23 = __fadd2_rn(make_float2(v[q * 4 + 2], v[q * 4 + 3]), make_float2(b.z, b.w));
            const float2 pairs[2] = {s01, s23};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              // the two logits exactly as the storing epilogue rounds them; -0 -> +0
              const __half2 h = __hadd2(__floats2half2_rn(pairs[e].x, pairs[e].y), __float2half2_rn(0.f));
              const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
              uint32_t sgn;      // 0xFFFF in the halves that are negative
              asm("prmt.b32 %0, %1, %2, 0xBB99;" : "=r"(sgn) : "r"(hb), "r"(0u));
              const uint32_t ord = hb ^ (sgn | 0x80008000u);       // order-preserving uint16 x 2
              // (0xFFFF - col) of the pair's two columns in one register; selectors as immediates
              const uint32_t col = static_cast<uint32_t>(g * 16 + q * 4 + e * 2);
              const uint32_t codes = ((0xFFFFu - (col + 1u)) << 16) | (0xFFFFu - col);
              uint32_t k0, k1;
              asm("prmt.b32 %0, %1, %2, 0x1054;" : "=r"(k0) : "r"(ord), "r"(codes));
              asm("prmt.b32 %0, %1, %2, 0x3276;" : "=r"(k1) : "r"(ord), "r"(codes));
              best_key = max(best_key, max(k0, k1));
            }
          }
        }
        const uint32_t ord_best = best_key >> 16;
        const unsigned short hbits = static_cast<unsigned short>(
            (ord_best & 0x8000u) ? (ord_best ^ 0x8000u) : (~ord_best & 0xFFFFu));
        const float best = __half2float(__ushort_as_half(hbits));
        const int best_c = static_cast<int>(0xFFFFu - (best_key & 0xFFFFu));
        const int row = tc.m_blk * BLOCK_M + row_in_tile;
        if (row < p.rows) {
          const size_t o = static_cast<size_t>(tc.b) * p.am_total + p.am_anchor_begin +
                           static_cast<size_t>(row) * p.am_num_anchors + tc.n_blk;
          p.am_scores[o] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-best)));
          p.am_classes[o] = best_c;
        }
      }
    } else {
      // Units of 32 columns: warp (quarter, team) owns rows quarter*32..+31 of the tile and the
      // units u == team (mod TEAMS).  Each warp has private 2 KiB staging slabs ([32 rows] x 64
      // bytes, 64B-swizzled) and issues its own [32 x 32] TMA stores; the only synchronisation is
      // the TMEM full / empty handshake with the MMA warp.
      const int e_warp = warp - 2;
      const int quarter = warp & 3;             // TMEM lane quarter this warp may access
      const int team = e_warp >> 2;
      const int row_in_tile = quarter * 32 + lane;
      uint8_t* my_slabs = smem_store + e_warp * p.slabs_per_warp * kSlabBytes;
      int store_cnt = 0;
      StageCounter sc;
      int rot = TEAMS - 1;                      // iter % TEAMS, kept incrementally
      for (int iter = 0;; ++iter) {
        TileCoord tc;
        const int t = ring_get(iter, true, &tc);
        if (t >= p.total_tiles) break;
        sc.next(p.accum_stages);
        if (++rot == TEAMS) rot = 0;
        const int as = sc.as;
        const uint32_t aphase = sc.phase;
        const int n0 = tc.n_blk * p.block_n;
        const int row = tc.m_blk * BLOCK_M + row_in_tile;
        const bool row_ok = row < p.rows;
        const __half* res_row = nullptr;
        if (HAS_RES) {
          res_row = p.residual + (static_cast<size_t>(tc.b) * p.rows + (row_ok ? row : 0)) * p.ldr;
        }
        mbar_wait(smem_u32(&tmem_full_bar[as]), aphase);
        tc_fence_after();
        const int n_valid = min(p.block_n, ((p.nout - n0 + 15) >> 4) << 4);
        const int num_units = (n_valid + 31) >> 5;
        // unit u of tile `iter` belongs to team (u + iter) % TEAMS: thin layers (one or two units
        // per tile) keep all three teams busy on consecutive tiles (up to kMaxAccum accumulators
        // in flight) instead of leaving two of them idle
        const int u_first = team >= rot ? team - rot : team + TEAMS - rot;
        int my_last = -1;
        for (int u = u_first; u < num_units; u += TEAMS) my_last = u;
        if (my_last < 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[as]));
        }
        for (int u = u_first; u < num_units; u += TEAMS) {
          const int cols = min(32, n_valid - u * 32);      // 16 or 32
          uint8_t* my_stage = my_slabs + (p.slabs_per_warp == 2 ? (store_cnt & 1) * kSlabBytes : 0);
          ++store_cnt;
          if (lane == 0) {                        // the store that last used this slab has left it
            if (p.slabs_per_warp == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
          }
          __syncwarp();
          const uint32_t row_base = smem_u32(my_stage) + lane * 64;
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                 static_cast<uint32_t>(as * p.block_n + u * 32);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (g * 16 < cols) {
              float v[16];
              tc_ld16(taddr + g * 16, v);
              tc_wait_ld();
              if (u == my_last && (g == 1 || cols == 16)) {
                // all TMEM reads of this warp for this accumulator are done
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[as]));
              }
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                const int col = n0 + u * 32 + g * 16 + jj * 8;
                float4 b0, b1;
                load_bias8(smem_bias_u32, col, b0, b1);
                float2 o2[4];
                o2[0] = __fadd2_rn(make_float2(v[jj * 8 + 0], v[jj * 8 + 1]), make_float2(b0.x, b0.y));
                o2[1] = __fadd2_rn(make_float2(v[jj * 8 + 2], v[jj * 8 + 3]), make_float2(b0.z, b0.w));
                o2[2] = __fadd2_rn(make_float2(v[jj * 8 + 4], v[jj * 8 + 5]), make_float2(b1.x, b1.y));
                o2[3] = __fadd2_rn(make_float2(v[jj * 8 + 6], v[jj * 8 + 7]), make_float2(b1.z, b1.w));
                apply_act4<ACT>(o2[0], o2[1]);
                apply_act4<ACT>(o2[2], o2[3]);
                float o[8] = {o2[0].x, o2[0].y, o2[1].x, o2[1].y, o2[2].x, o2[2].y, o2[3].x, o2[3].y};
                if (HAS_RES) {
                  if (row_ok && col < p.nout) {
                    float r[8];
                    half8_to_float(ldg_nc_v4(res_row + col), r);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] += r[e];
                  }
                }
                const uint4 packed = float_to_half8(o);
                const int chunk16 = g * 2 + jj;    // 16-byte piece inside the 64-byte row
                // 64B swizzle: the 16-byte piece index is XORed with bits [7:8] of the address
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(
                                 row_base + (((chunk16 ^ ((lane >> 1) & 3))) << 4)),
                             "r"(packed.x), "r"(packed.y), "r"(packed.z), "r"(packed.w)
                             : "memory");
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_o, smem_u32(my_stage), n0 + u * 32, tc.m_blk * BLOCK_M + quarter * 32,
                         tc.b);
            tma_store_commit();
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
  }
}

// N tile.  Two CTAs are resident per SM (<=113 KiB smem, <=256 TMEM columns each):
//   nout <= 128 : one tile, two accumulator stages;
//   nout <= 256 : one tile, ONE accumulator stage (the other CTA of the SM hides the gap) --
//                 avoids re-reading A and a ragged second tile for N = 144 / 240 (splitting these
//                 into two double-buffered tiles measured 1.5 % slower on the D0 step);
//   wider       : tiles of 128 columns; the A tile of the extra tiles comes from L2 and the
//                 epilogue skips the columns past nout.
// With three epilogue teams (32-column units) wide layers use tiles of 96 columns (3 units).
static int pick_block_n(int nout, int teams) {
  if (nout <= 128) return ((nout + 15) / 16) * 16;
  if (nout <= 256) return ((nout + 15) / 16) * 16;
  return teams == 3 ? 96 : 128;
}

template <int ACT, bool HAS_RES, int TEAMS, int EPI = EPI_STORE>
static int launch(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mo,
                  const Params& p, int grid, int smem_bytes, cudaStream_t stream) {
  auto kern = pointwise_tc_kernel<ACT, HAS_RES, TEAMS, EPI>;
  static int configured[kMaxDevices];   // per instantiation and device; no API call once set
  if (int rc = ensure_dynamic_smem(kern, kSmemLimit, configured)) return rc;
  EDET_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(Epi<TEAMS>::kThreads), smem_bytes, stream, ma,
                             mw, mo, p));
  return EDET_OK;
}

int run(const __half* a, int lda, const __half* wt, int wbatch, const float* bias,
        const __half* residual, int ldr, __half* out, int ldo, int batch, int rows, int k, int nout,
        int act, cudaStream_t stream, const ArgmaxArgs* am) {
  Params p;
  p.am_scores = nullptr; p.am_classes = nullptr;
  p.am_anchor_begin = p.am_total = p.am_num_anchors = 0;
  if (am) {
    EDET_CHECK_ARG(nout == am->num_anchors * kArgmaxCols && !residual && act == EDET_ACT_NONE,
                   "class_argmax: nout must be num_anchors * %d", kArgmaxCols);
    p.am_scores = am->scores; p.am_classes = am->classes;
    p.am_anchor_begin = am->anchor_begin; p.am_total = am->total_anchors;
    p.am_num_anchors = am->num_anchors;
  }
  p.batch = batch;
  p.rows = rows;
  p.k = k;
  p.nout = nout;
  p.nout_pad8 = nout & ~7;   // whole float4 pairs of bias that are in bounds
  // Epilogue teams: three teams of four warps (448 threads at <= 72 registers) by default -- the
  // epilogues are bound by MUFU / dependent-issue latency, not bytes, and 50 % more warps hide it
  // (measured on the D0 step: 4.68 ms with two teams, 4.53 ms with three, all layers);
  // edet_set_option("pw_teams", 2 | 3) forces a variant for A/B measurements.
  const int opt_teams = option_pw_teams();
  const int teams = am ? 3 : (opt_teams ? opt_teams : 3);
  const int epi_warps = 4 * teams, slab_bytes = teams == 2 ? 4096 : 2048;
  p.block_n = pick_block_n(nout, teams);
  p.num_m_blocks = ceil_div(rows, BLOCK_M);
  p.num_n_blocks = ceil_div(nout, p.block_n);
  // k-block / smem row pitch: 16 halves (32B swizzle) for K <= 16, 32 (64B) for K <= 32, else
  // 64 (128B) -- unless two 64-wide stages of this (wide) N tile would not fit next to the store
  // slabs, in which case the narrower 32-wide stages keep the pipeline >= 3 deep.
  p.wbatch = wbatch;
  p.ldr = ldr;
  p.bias = bias;
  p.residual = residual;
  // accumulator stages: as many as fit in this CTA's 256 TMEM columns (two CTAs per SM), up to 4
  // with three teams (so that every team can work on its own tile of a thin layer), 2 with two
  p.accum_stages = 256 / p.block_n;
  if (p.accum_stages < 1) p.accum_stages = 1;
  if (p.accum_stages > (teams == 3 ? kMaxAccum : 2)) p.accum_stages = teams == 3 ? kMaxAccum : 2;
  int cols = 32;
  while (cols < p.accum_stages * p.block_n) cols *= 2;
  p.tmem_cols = cols;
  p.total_tiles = batch * p.num_m_blocks * p.num_n_blocks;
  p.sched = next_sched_slot();
  if (!p.sched) return EDET_ERR_CUDA;

  // k-block / smem row pitch: 16 halves (32B swizzle) for K <= 16, 32 (64B) for K <= 32, else 64
  // (128B).  Two store slabs per epilogue warp when they fit; a wide N tile falls back first to
  // one slab, then to 32-wide k-blocks, so that the TMA ring stays >= 3 deep (>= 2 at worst).
  // every column the epilogue can touch: whole N tiles (n_valid is rounded up to 16 inside a tile)
  const int bias_cols = p.num_n_blocks * p.block_n;
  EDET_CHECK_ARG(bias_cols <= kMaxBiasSmem, "pointwise_tc: nout %d too wide", nout);
  p.bias_floats = bias_cols;
  const int ctrl = p.bias_floats * 4 + (2 * kMaxStages + 2 * kMaxAccum + 2 * kRing) * 8 + 16 + 16 * kRing;
  int best_stages = 0, fixed = 0, stage_bytes = 0;
  auto plan = [&](int limit) {
    best_stages = 0;
    for (int attempt = 0; attempt < 4 && best_stages < 3; ++attempt) {
      const int bk = k <= 16 ? 16 : (k <= 32 ? 32 : (attempt >= 2 ? 32 : 64));
      const int slabs = (attempt & 1) ? 1 : 2;
      const int a_bytes = BLOCK_M * bk * 2;
      const int b_bytes = ((p.block_n * bk * 2 + 1023) / 1024) * 1024;
      const int fx = slabs * epi_warps * slab_bytes + ctrl;
      int st = (limit - 1024 - fx) / (a_bytes + b_bytes);
      if (st > kMaxStages) st = kMaxStages;
      if (st > best_stages) {
        best_stages = st;
        p.block_k = bk;
        p.slabs_per_warp = slabs;
        p.a_stage_bytes = a_bytes;
        p.b_stage_bytes = b_bytes;
        fixed = fx;
        stage_bytes = a_bytes + b_bytes;
      }
    }
  };
  // Shared-memory budget.  The NMS of the previous batch (one 126 KiB CTA per image, own stream)
  // runs under the first layers of the next network: with 99 KiB a pointwise CTA fits next to it
  // (99 + 126 + 2 KiB <= 228 KiB), with 113 KiB those SMs stay empty.  So 99 KiB wherever that
  // still gives a ring of >= 4 stages (every thin-K layer), the full half SM otherwise.
  const int opt_kb = option_pw_smem_kb();
  if (opt_kb) {
    plan(opt_kb * 1024);
  } else {
    plan(kSmemCoResident);
    if (best_stages < 4) plan(kSmemLimit);
  }
  EDET_CHECK_ARG(best_stages >= 2, "pointwise_tc: block_n %d leaves <2 pipeline stages", p.block_n);
  p.num_stages = best_stages;
  p.desc_layout = p.block_k == 64 ? 2 : (p.block_k == 32 ? 4 : 6);
  p.desc_sbo = 8 * p.block_k * 2;
  p.num_k_blocks = ceil_div(k, p.block_k);
  const int smem_bytes = 1024 + best_stages * stage_bytes + fixed;

  CUtensorMap ma, mw, mo;
  int rc;
  if ((rc = make_map(&ma, a, k, rows, batch, lda, static_cast<uint64_t>(rows) * lda, BLOCK_M,
                     p.block_k)))
    return rc;
  if ((rc = make_map(&mw, wt, k, nout, wbatch, k, static_cast<uint64_t>(nout) * k, p.block_n,
                     p.block_k)))
    return rc;
  // store box: [32 rows] x 64 columns (128B swizzle) for two teams, x 32 columns (64B) for three
  if (am) {
    mo = mw;   // the arg-max epilogue stores nothing through TMA
  } else if ((rc = make_map(&mo, out, nout, rows, batch, ldo, static_cast<uint64_t>(rows) * ldo, 32,
                            teams == 2 ? 64 : 32))) {
    return rc;
  }

  const int sm_count = device_sm_count();
  if (!sm_count) return EDET_ERR_CUDA;
  int grid = 2 * sm_count - option_persist_slack();
  if (grid < sm_count) grid = sm_count;
  if (p.total_tiles < grid) grid = p.total_tiles;
  const bool has_res = residual != nullptr;
  if (am) return launch<EDET_ACT_NONE, false, 3, EPI_ARGMAX>(ma, mw, mo, p, grid, smem_bytes, stream);

#define EDET_PW_CASE(A)                                                                       \
  if (teams == 3)                                                                             \
    return has_res ? launch<A, true, 3>(ma, mw, mo, p, grid, smem_bytes, stream)              \
                   : launch<A, false, 3>(ma, mw, mo, p, grid, smem_bytes, stream);            \
  return has_res ? launch<A, true, 2>(ma, mw, mo, p, grid, smem_bytes, stream)                \
                 : launch<A, false, 2>(ma, mw, mo, p, grid, smem_bytes, stream)
  switch (act) {
    case EDET_ACT_NONE: EDET_PW_CASE(EDET_ACT_NONE);
    case EDET_ACT_SWISH: EDET_PW_CASE(EDET_ACT_SWISH);
    case EDET_ACT_RELU: EDET_PW_CASE(EDET_ACT_RELU);
    case EDET_ACT_RELU6: EDET_PW_CASE(EDET_ACT_RELU6);
    case EDET_ACT_HSWISH: EDET_PW_CASE(EDET_ACT_HSWISH);
    case EDET_ACT_SIGMOID: EDET_PW_CASE(EDET_ACT_SIGMOID);
    default:
      set_error("pointwise_tc: bad activation %d", act);
      return EDET_ERR_INVALID;
  }
#undef EDET_PW_CASE
}

}  // namespace pwtc
}  // namespace edet
