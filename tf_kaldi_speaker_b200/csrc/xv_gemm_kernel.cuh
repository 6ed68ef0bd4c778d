// Persistent, warp-specialised implicit-GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> 5-stage smem ring -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM,
//   256 x 256 tile per CTA pair or 128 x 128 per CTA, double-buffered accumulator) -> tcgen05.ld epilogue ->
//   swizzled smem staging -> TMA store (bf16 / f32) or TMA reduce-add (f32 split-K).
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM
// allocator, warp 3 idle, warps 4..11 = epilogue (TMEM lane quadrant = warp_idx % 4, column half = (warp_idx-4)/4).
//
// One kernel serves every contraction of the x-vector step (see include/xvector_b200.h):
//   forward  conv/dense : A = activations (K-major, tap rows +j), B = kernel [k*Cin, Cout] (MN-major)
//   dgrad               : A = dY (K-major, tap rows -j),          B = kernel viewed [k*Cin, Cout] (K-major, taps)
//   wgrad               : A = activations (MN-major, tap rows +j), B = dY (MN-major), split-K + fp32 atomics
//   head                : A = embeddings, B = normalised speaker matrix, fused margin / online-LSE epilogues
#pragma once
#include "xv_internal.h"
#include "xv_ptx.cuh"

namespace xv {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
// CG = CTAs cooperating on one output tile.
//   CG = 2 (tcgen05 cta_group::2, a 2-CTA cluster on one TPC): 256 x 256 tile; each CTA stages its 128 rows of A and
//           HALF of B (the MMA reads the other half from the peer's shared memory) -- a third less L2 -> SM traffic and
//           shared-memory fill per FLOP than a 1-CTA 128 x 256 tile.  The frame-level fwd / dgrad / wgrad launches.
//   CG = 1: 128 x 128 tile, one CTA; the latency-bound launches (utterance level, head, small problems), where twice as
//           many tiles means twice as many SMs pulling operands (57 CTAs instead of 29 for the 7200-speaker head).
// Both stage (16 KB A + 16 KB B) per 64-deep k-block in a 5-stage ring.
template <int CG>
struct GemmCfg {
  static constexpr int BN = CG == 2 ? 256 : 128;                    // tile columns
  static constexpr int B_ROWS = BN / CG;                            // B (N) rows staged per CTA = 128
  static constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = 5;
  static constexpr int TILE_M = BLOCK_M * CG;
  static constexpr int EPI_COLS = BN / 2;                           // columns per epilogue warp (column half)
};
constexpr int MAX_BN = 256;
constexpr int MAX_STAGES = 5;
constexpr int PIPE_BYTES = MAX_STAGES * (A_STAGE_BYTES + 128 * BLOCK_K * 2);   // 160 KB
constexpr int CHUNK_BYTES = 64 * BLOCK_K * 2;          // one MN-major 64x64 box = 8 KB
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 384;
constexpr int NUM_EPI_WARPS = 8;
constexpr int STAGING_BYTES = 32 * 128;                // per epilogue warp: one 32-row x 128-byte TMA store box
// Column statistics of a tile: every epilogue warp STORES the sums of its 32 rows, [quadrant][sum|sumsq][col]; after a
// named barrier one thread per column adds the four quadrants in a fixed order (no shared-memory atomics: the per-tile
// result is deterministic) and issues one global atomic.  Double-buffered by tile parity: one barrier per tile.
constexpr int PART_FLOATS = 4 * 2 * MAX_BN;
constexpr int STATS_BYTES = 2 * PART_FLOATS * 4;       // 16 KB
constexpr int BAR_BYTES = 256;
// The dynamic shared-memory window is declared 1024-byte aligned (128B-swizzle atom); no slack is reserved.
constexpr int SMEM_BYTES = PIPE_BYTES + NUM_EPI_WARPS * STAGING_BYTES + STATS_BYTES + BAR_BYTES;
static_assert(GemmCfg<1>::STAGES * GemmCfg<1>::STAGE_BYTES == PIPE_BYTES && GemmCfg<2>::STAGES * GemmCfg<2>::STAGE_BYTES == PIPE_BYTES,
              "both pipeline configurations fill the same 160 KB");
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");

struct alignas(64) GemmKernelParams {
  CUtensorMap tma_a;
  CUtensorMap tma_b;
  CUtensorMap tma_out;    // output matrix (box 32 rows x 128 bytes, 128B swizzle); valid when use_tma_out
  int use_tma_out;
  int M, N, K;
  int a_div, a_tap, b_div, b_tap;
  int a_mn, b_mn;
  int num_m, num_n, splits, num_kb, kb_per_split;
  int seg_len, seg_valid;
  int accumulate;
  void* out;
  long long ldc;
  const float* bias;
  float* col_sum;
  float* col_sumsq;
  xv_head_args head;
  xv_bn_bwd_args bnb;     // y != nullptr: BN-backward reductions of the layer whose activation gradient this GEMM emits
  const float* aff_scale; // != nullptr (bf16 TMA epilogue): out = act(acc * aff_scale[n] + aff_shift[n]) -- inference-mode
  const float* aff_shift; //   batch-norm (moving statistics folded into scale / shift) + activation in the epilogue
  float aff_neg_slope;    //   act(z) = z > 0 ? z : neg_slope * z   (relu 0, leaky_relu 0.2, identity 1)
};

// Sum over the 32 lanes of a warp of v[j] for each of 32 columns j, in 31 shuffles (recursive halving).
// On return v[0] of lane l holds the total of column l.
__device__ __forceinline__ void warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool upper = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      const float send = upper ? v[i] : v[i + step];
      const float keep = upper ? v[i + step] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
}

// ---- margin transforms of the target logit (model/loss.py:129-137, 225, 314-323) -----------------------
struct MarginOut {
  float zprime;   // fs*z + fa*n*phi(c)
  float dz;       // d z'/d z   = fs + fa*phi'(c)*k
  float dn;       // d z'/d n   = fa*(phi(c) - phi'(c)*k*z/n)
};
__device__ __forceinline__ MarginOut margin_target(const xv_head_args& h, float z, float n) {
  MarginOut o;
  const float ratio = z / n;
  // clip_by_value(cos, -1+1e-12, 1-1e-12): in fp32 the bounds round to -1 / 1; gradient passes inside only
  const float c = fminf(fmaxf(ratio, -1.0f), 1.0f);
  const float k = (ratio >= -1.0f && ratio <= 1.0f) ? 1.0f : 0.0f;
  float phi, dphi;
  if (h.type == XV_HEAD_AM) {
    phi = c - h.margin;
    dphi = 1.0f;
  } else if (h.type == XV_HEAD_AAM) {
    const float cm = h.cos_m, sm = h.sin_m;
    const float s2 = 1.0f - c * c;
    const float s = sqrtf(fmaxf(s2, 1e-12f));
    const float u = c * cm - s * sm;
    const float du = cm + ((s2 > 1e-12f) ? (c * sm / s) : 0.0f);
    const bool easy = c > h.threshold;
    phi = easy ? u : (-u - 2.0f);
    dphi = easy ? du : -du;
  } else {  // XV_HEAD_ASOFTMAX, m = 2 or 4
    const float s0 = (c > 0.f) ? 1.f : ((c < 0.f) ? -1.f : 0.f);
    const float c2 = c * c;
    if (h.asoftmax_m == 2) {
      phi = 2.0f * s0 * c2 - 1.0f;
      dphi = 4.0f * s0 * c;
    } else {
      const float t = 2.0f * c2 - 1.0f;
      const float s3 = ((t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : 0.f)) * s0;
      const float s4 = 2.0f * s0 + s3 - 3.0f;
      phi = s3 * (8.0f * c2 * c2 - 8.0f * c2 + 1.0f) + s4;
      dphi = s3 * (32.0f * c2 * c - 16.0f * c);
    }
  }
  const float fa = __ldg(h.sched), fs = __ldg(h.sched + 1);
  o.zprime = fs * z + fa * n * phi;
  o.dz = fs + fa * dphi * k;
  o.dn = fa * (phi - dphi * k * ratio);
  return o;
}

// 128B-swizzled staging tile of one epilogue warp (32 rows x 128 bytes): logical 16-byte chunk j of row r lives at
// chunk position j ^ (r & 7) -- the layout CU_TENSOR_MAP_SWIZZLE_128B expects, and conflict-free for a warp
// whose lanes each write 16 bytes of their own row.
__device__ __forceinline__ void stage_store16(uint8_t* stg, int row, int chunk, uint4 v) {
  *reinterpret_cast<uint4*>(stg + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
}

template <int CG>
__device__ __forceinline__ void tma_load(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  if (CG == 2) tma_load_2d_cg2(dst, map, bar, c0, c1);
  else tma_load_2d(dst, map, bar, c0, c1);
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (CG == 2) umma_bf16_cg2(d, da, db, idesc, acc);
  else umma_bf16(d, da, db, idesc, acc);
}
template <int CG>
__device__ __forceinline__ void umma_commit_n(uint64_t* bar) {
  if (CG == 2) umma_commit_cg2(bar);
  else umma_commit(bar);
}

// Packed fp32 pairs (sm_100 add / fma .f32x2): two columns per instruction in the statistics column pass.
__device__ __forceinline__ uint64_t pack_f32x2(uint32_t lo_bits, uint32_t hi_bits) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo_bits), "r"(hi_bits));
  return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// One statistic of one column leaves the CTA: a single global atomic per (tile, column).
__device__ __forceinline__ void stat_emit(float* dst, int n, float v) { atomicAdd(dst + n, v); }

// Release the TMEM accumulator of this tile to the MMA warp: called by every epilogue warp as soon as its tcgen05.ld
// traffic for the tile has completed (the arithmetic / stores that follow overlap the next tile's MMAs).
template <int CG>
__device__ __forceinline__ void release_accumulator(uint64_t* bar, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    if (CG == 2) mbar_arrive_leader(bar);
    else mbar_arrive(bar);
  }
}

template <int EPI, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_kernel(const __grid_constant__ GemmKernelParams p) {
  using Cfg = GemmCfg<CG>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int B_STAGE_BYTES = Cfg::B_STAGE_BYTES;
  constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int BN = Cfg::BN;
  constexpr int EPI_COLS = Cfg::EPI_COLS;
  const bool A_MN = p.a_mn != 0, B_MN = p.b_mn != 0;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smem_stg = smem + PIPE_BYTES;
  float* s_part = reinterpret_cast<float*>(smem_stg + NUM_EPI_WARPS * STAGING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_part) + STATS_BYTES);
  uint64_t* full_bar = bars;                            // [MAX_STAGES]  (CG = 2: only the leader's are used)
  uint64_t* empty_bar = bars + MAX_STAGES;              // [MAX_STAGES]
  uint64_t* tmem_full = bars + 2 * MAX_STAGES;          // [2]
  uint64_t* tmem_empty = bars + 2 * MAX_STAGES + 2;     // [2]           (CG = 2: the leader's collect both CTAs)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  // CTA pair bookkeeping: rank 0 ("leader") issues the MMAs for both CTAs
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const int unit_id = CG == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);   // tile-walking unit
  const int num_units = CG == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();      // the next kernel of the stream may be scheduled as SMs free up
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("xv: dynamic shared memory base is not 1024-byte aligned (0x%x)\n", smem_u32(smem));
    __trap();
  }
  if (warp_idx == 0 && lane == 0) {
    prefetch_tmap(&p.tma_a);
    prefetch_tmap(&p.tma_b);
    if (p.use_tma_out) prefetch_tmap(&p.tma_out);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], NUM_EPI_WARPS * CG);   // one arrive per epilogue warp (of both CTAs of a pair)
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    if (CG == 2) {
      tmem_alloc_cg2(tmem_ptr, TMEM_COLS);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_ptr, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();                      // publishes tmem_ptr inside the CTA (compute-sanitizer's racecheck does not model
                                        // barrier.cluster as a shared-memory synchronisation point; once per kernel, free)
  if (CG == 2) cluster_sync_all();      // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // Everything above (barrier init, TMEM allocation, tensor-map prefetch) is independent of earlier kernels; global
  // memory is first touched below, after the previous kernel of the stream has completed.
  pdl_wait();

  const int total_tiles = p.num_m * p.num_n * p.splits;     // num_m counts TILE_M-row blocks

  if (warp_idx == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit_id; tile < total_tiles; tile += num_units) {
        const int n_blk = tile % p.num_n;
        const int m_blk = (tile / p.num_n) % p.num_m;
        const int split = tile / (p.num_n * p.num_m);
        const int m0 = m_blk * Cfg::TILE_M + static_cast<int>(cta_rank) * BLOCK_M;   // this CTA's 128 rows of A
        const int n0 = n_blk * BN + static_cast<int>(cta_rank) * Cfg::B_ROWS;        // this CTA's share of B
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        // Tile coordinates are walked incrementally: the only integer divisions happen here, once per tile
        // (a division per k-block made this single thread slower than the MMAs it feeds).
        int a_col, a_row, b_col, b_row;
        if (!A_MN) {
          const int kk = kb0 * BLOCK_K;
          a_col = p.a_div ? (kk % p.a_div) : kk;
          a_row = m0 + (p.a_div ? (kk / p.a_div) * p.a_tap : 0);
        } else {
          a_col = p.a_div ? (m0 % p.a_div) : m0;
          a_row = kb0 * BLOCK_K + (p.a_div ? (m0 / p.a_div) * p.a_tap : 0);
        }
        if (!B_MN) {
          const int kk = kb0 * BLOCK_K;
          b_col = p.b_div ? (kk % p.b_div) : kk;
          b_row = n0 + (p.b_div ? (kk / p.b_div) * p.b_tap : 0);
        } else {
          b_col = p.b_div ? (n0 % p.b_div) : n0;
          b_row = kb0 * BLOCK_K + (p.b_div ? (n0 / p.b_div) * p.b_tap : 0);
        }
        const int a_wrap = (!A_MN && p.a_div) ? p.a_div : 0x7fffffff;
        const int b_wrap = (!B_MN && p.b_div) ? p.b_div : 0x7fffffff;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          // CG = 2: both CTAs' loads complete on the LEADER's barrier, which its producer arms for all the bytes
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES * CG);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          if (!A_MN) {
            tma_load<CG>(sa, &p.tma_a, &full_bar[stage], a_col, a_row);
            a_col += BLOCK_K;
            if (a_col >= a_wrap) { a_col = 0; a_row += p.a_tap; }
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load<CG>(sa + c * CHUNK_BYTES, &p.tma_a, &full_bar[stage], a_col + 64 * c, a_row);
            a_row += BLOCK_K;
          }
          if (!B_MN) {
            tma_load<CG>(sb, &p.tma_b, &full_bar[stage], b_col, b_row);
            b_col += BLOCK_K;
            if (b_col >= b_wrap) { b_col = 0; b_row += p.b_tap; }
          } else {
#pragma unroll
            for (int c = 0; c < Cfg::B_ROWS / 64; ++c)
              tma_load<CG>(sb + c * CHUNK_BYTES, &p.tma_b, &full_bar[stage], b_col + 64 * c, b_row);
            b_row += BLOCK_K;
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ============================== MMA issuer (CG = 2: leader CTA only) ==============================
    if (lane == 0 && cta_rank == 0) {
      const uint32_t idesc = make_idesc_bf16(Cfg::TILE_M, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // K-major SW128: 8-row groups 1024 B apart (SBO), LBO unused.  MN-major SW128: 64-wide MN chunks
      // CHUNK_BYTES apart (LBO), 8-k-row groups 1024 B apart (SBO).
      const uint32_t a_lbo = A_MN ? CHUNK_BYTES : 16, b_lbo = B_MN ? CHUNK_BYTES : 16;
      const uint32_t a_kstep = A_MN ? (UMMA_K * 128) : (UMMA_K * 2);   // bytes per UMMA_K
      const uint32_t b_kstep = B_MN ? (UMMA_K * 128) : (UMMA_K * 2);
      // descriptors of stage 0 / k = 0; the start-address field (bits 0..13, 16-byte units) is advanced by adds
      const uint64_t da0 = make_smem_desc(smem_u32(smem_a), a_lbo, 1024);
      const uint64_t db0 = make_smem_desc(smem_u32(smem_b), b_lbo, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = unit_id; tile < total_tiles; tile += num_units, ++local) {
        const int split = tile / (p.num_n * p.num_m);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = da0 + static_cast<uint64_t>((stage * A_STAGE_BYTES) >> 4);
          const uint64_t db = db0 + static_cast<uint64_t>((stage * B_STAGE_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma<CG>(d_tmem, da + static_cast<uint64_t>((k * a_kstep) >> 4), db + static_cast<uint64_t>((k * b_kstep) >> 4),
                     idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit_n<CG>(&empty_bar[stage]);   // frees the smem slot (in both CTAs) once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_n<CG>(&tmem_full[acc]);       // accumulator complete -> epilogue (of both CTAs)
      }
    }
  } else if (warp_idx >= 4) {
    // ============================== epilogue ==============================
    const int e = warp_idx - 4;             // 0..7
    const int qd = warp_idx & 3;            // TMEM lane quadrant this warp may read
    const int hf = e >> 2;                  // column half of the tile
    const int et = threadIdx.x - 128;       // 0..255
    uint8_t* stg = smem_stg + e * STAGING_BYTES;
    const bool tma_out = p.use_tma_out != 0;
    int local = 0;
    for (int tile = unit_id; tile < total_tiles; tile += num_units, ++local) {
      const int n_blk = tile % p.num_n;
      const int m_blk = (tile / p.num_n) % p.num_m;
      const int split = tile / (p.num_n * p.num_m);
      const int m0 = m_blk * Cfg::TILE_M + static_cast<int>(cta_rank) * BLOCK_M, n0 = n_blk * BN;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int row0 = m0 + qd * 32;
      const int m = row0 + lane;
      const bool row_ok = m < p.M;
      const int cbase = n0 + hf * EPI_COLS;                       // first global column of this warp
      const int ncols = max(0, min(EPI_COLS, p.N - cbase));
      float* part = s_part + (local & 1) * PART_FLOATS;           // [quadrant][sum | sumsq][MAX_BN]
      if (EPI == XV_EPI_BF16 && p.bnb.y != nullptr && row_ok && ncols > 0) {
        // The fused BN-backward reductions read this thread's row of y (256 B): pull it into L2 now, while the MMAs of
        // this tile are still running, so the loads below do not expose HBM latency four times per tile.
        const char* yb = reinterpret_cast<const char*>(reinterpret_cast<const __nv_bfloat16*>(p.bnb.y) +
                                                       static_cast<long long>(m) * p.bnb.ldy + cbase);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(yb));
        if (ncols > 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(yb + 128));
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + acc * BN + hf * EPI_COLS;

      const bool do_stats = (EPI == XV_EPI_BF16 || EPI == XV_EPI_HEAD_BWD) && p.col_sum != nullptr;
      bool row_valid = row_ok;
      if (EPI == XV_EPI_BF16 && p.seg_len > 0) row_valid = row_ok && ((m % p.seg_len) < p.seg_valid);

      if (EPI == XV_EPI_BF16 && tma_out) {
        // ---------------- bf16 matrix output through TMA stores, one 32-row x 64-column box at a time ----------------
        // tcgen05.ld -> (release the accumulator after the last load) -> pack to bf16 into the swizzled staging tile ->
        // TMA store -> column pass over the STAGED tile for the statistics: lane l owns columns 2l, 2l+1 (one
        // conflict-free 4-byte word per row), so the sums over the 32 rows need no shuffles at all.  (v1 summed the fp32
        // accumulators with two 31-shuffle butterflies per 32 x 32 chunk: the K = 512 layers were bound by them.)
        constexpr int NBOX = EPI_COLS / 64;
        const bool zero_invalid = do_stats && p.bnb.y == nullptr;       // forward statistics
        const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
        bool released = false;
#pragma unroll
        for (int b = 0; b < NBOX; ++b) {
          const int bc0 = cbase + 64 * b;                            // first global column of the box (warp-uniform)
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          if (bc0 < p.N) {
            uint32_t r0[32], r1[32];
            const bool second = bc0 + 32 < p.N;
            tmem_ld_32x32(t_row + 64 * b, r0);
            if (second) tmem_ld_32x32(t_row + 64 * b + 32, r1);
            tmem_ld_wait();
            if (b == NBOX - 1 || bc0 + 64 >= p.N) {
              release_accumulator<CG>(&tmem_empty[acc], lane);
              released = true;
            }
            if (!second) {
#pragma unroll
              for (int j = 0; j < 32; ++j) r1[j] = 0u;
            }
            if (p.aff_scale != nullptr) {
              // folded batch-norm + activation: per-column constants are warp-uniform (every lane holds one ROW), so
              // they arrive as broadcast 16-byte loads
              const float ns = p.aff_neg_slope;
              if (bc0 + 64 <= p.N) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.aff_scale + bc0) + j);
                  const float4 h4 = __ldg(reinterpret_cast<const float4*>(p.aff_shift + bc0) + j);
                  const float4 s5 = __ldg(reinterpret_cast<const float4*>(p.aff_scale + bc0 + 32) + j);
                  const float4 h5 = __ldg(reinterpret_cast<const float4*>(p.aff_shift + bc0 + 32) + j);
                  const float sc[8] = {s4.x, s4.y, s4.z, s4.w, s5.x, s5.y, s5.z, s5.w};
                  const float sh[8] = {h4.x, h4.y, h4.z, h4.w, h5.x, h5.y, h5.z, h5.w};
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    float z0 = fmaf(__uint_as_float(r0[4 * j + q]), sc[q], sh[q]);
                    float z1 = fmaf(__uint_as_float(r1[4 * j + q]), sc[4 + q], sh[4 + q]);
                    z0 = z0 > 0.f ? z0 : z0 * ns;
                    z1 = z1 > 0.f ? z1 : z1 * ns;
                    r0[4 * j + q] = __float_as_uint(z0);
                    r1[4 * j + q] = __float_as_uint(z1);
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (bc0 + j < p.N) {
                    const float z = fmaf(__uint_as_float(r0[j]), __ldg(p.aff_scale + bc0 + j), __ldg(p.aff_shift + bc0 + j));
                    r0[j] = __float_as_uint(z > 0.f ? z : z * ns);
                  }
                  if (bc0 + 32 + j < p.N) {
                    const float z = fmaf(__uint_as_float(r1[j]), __ldg(p.aff_scale + bc0 + 32 + j), __ldg(p.aff_shift + bc0 + 32 + j));
                    r1[j] = __float_as_uint(z > 0.f ? z : z * ns);
                  }
                }
              }
              if (p.seg_len > 0 && !row_valid) {      // activations of invalid frames are zeros (as xv_bn_act_apply writes them)
#pragma unroll
                for (int j = 0; j < 32; ++j) { r0[j] = 0u; r1[j] = 0u; }
              }
            }
            if (zero_invalid && !row_valid) {     // rows outside the valid frames are stored as zeros: no mask in the column pass
#pragma unroll
              for (int j = 0; j < 32; ++j) { r0[j] = 0u; r1[j] = 0u; }
            }
            if (p.bias) {        // layers followed by a batch-norm pass no bias (it cancels; folded into the BN shift)
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (bc0 + j < p.N) r0[j] = __float_as_uint(__uint_as_float(r0[j]) + __ldg(p.bias + bc0 + j));
                if (bc0 + 32 + j < p.N) r1[j] = __float_as_uint(__uint_as_float(r1[j]) + __ldg(p.bias + bc0 + 32 + j));
              }
            }
            if (lane == 0) bulk_wait_read_all();      // the previous box has left the staging tile
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              stage_store16(stg, lane, j,
                            make_uint4(pack_bf16x2(__uint_as_float(r0[8 * j]), __uint_as_float(r0[8 * j + 1])),
                                       pack_bf16x2(__uint_as_float(r0[8 * j + 2]), __uint_as_float(r0[8 * j + 3])),
                                       pack_bf16x2(__uint_as_float(r0[8 * j + 4]), __uint_as_float(r0[8 * j + 5])),
                                       pack_bf16x2(__uint_as_float(r0[8 * j + 6]), __uint_as_float(r0[8 * j + 7]))));
              stage_store16(stg, lane, 4 + j,
                            make_uint4(pack_bf16x2(__uint_as_float(r1[8 * j]), __uint_as_float(r1[8 * j + 1])),
                                       pack_bf16x2(__uint_as_float(r1[8 * j + 2]), __uint_as_float(r1[8 * j + 3])),
                                       pack_bf16x2(__uint_as_float(r1[8 * j + 4]), __uint_as_float(r1[8 * j + 5])),
                                       pack_bf16x2(__uint_as_float(r1[8 * j + 6]), __uint_as_float(r1[8 * j + 7]))));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&p.tma_out, stg, bc0, row0);   // columns >= N and rows >= M are clipped by the TMA unit
              bulk_commit();
            }
            if (do_stats) {
              const int colb = bc0 + 2 * lane;                           // first of this lane's two columns
              const bool col_ok = colb < p.N;                            // N is even: both columns or neither
              const uint8_t* sbase = stg + ((lane & 3) << 2);
              const int jc = lane >> 2;
              if (p.bnb.y == nullptr) {
                // forward BN statistics of the STORED (bf16) tensor: sum and sum of squares over the 32 rows (invalid rows
                // were zeroed before staging).  Packed f32x2 arithmetic (sm_100): one add + one fma per row for the
                // lane's two columns; even / odd rows keep separate accumulators (shorter dependency chains).
                uint64_t sa = 0ull, sb = 0ull, qa = 0ull, qb = 0ull;
#pragma unroll
                for (int r = 0; r < 32; r += 2) {
                  const uint32_t w0 = *reinterpret_cast<const uint32_t*>(sbase + r * 128 + ((jc ^ (r & 7)) << 4));
                  const uint32_t w1 = *reinterpret_cast<const uint32_t*>(sbase + (r + 1) * 128 + ((jc ^ ((r + 1) & 7)) << 4));
                  const uint64_t v0 = pack_f32x2(w0 << 16, w0 & 0xffff0000u);
                  const uint64_t v1 = pack_f32x2(w1 << 16, w1 & 0xffff0000u);
                  sa = add_f32x2(sa, v0); qa = fma_f32x2(v0, v0, qa);
                  sb = add_f32x2(sb, v1); qb = fma_f32x2(v1, v1, qb);
                }
                sa = add_f32x2(sa, sb); qa = add_f32x2(qa, qb);
                if (col_ok) {
                  s0 = __uint_as_float(static_cast<uint32_t>(sa)); s1 = __uint_as_float(static_cast<uint32_t>(sa >> 32));
                  q0 = __uint_as_float(static_cast<uint32_t>(qa)); q1 = __uint_as_float(static_cast<uint32_t>(qa >> 32));
                }
              } else if (col_ok) {
                // Fused BN backward of the producer layer: g = dX * act'(y*scale + shift); dbeta += g,
                // dgamma += g * (y - mean) * rstd = (sum g*y - mean * sum g) * rstd.  y is read with coalesced 128-byte rows, the
                // per-column constants live in registers; g uses the STORED (bf16) gradient, like the stand-alone kernel.
                // Packed f32x2 arithmetic for the lane's two columns; rows >= M hold zeros (TMA zero fill), so no row mask.
                const uint64_t sc2 = pack_f32x2(__float_as_uint(__ldg(p.bnb.scale + colb)), __float_as_uint(__ldg(p.bnb.scale + colb + 1)));
                const uint64_t sh2 = pack_f32x2(__float_as_uint(__ldg(p.bnb.shift + colb)), __float_as_uint(__ldg(p.bnb.shift + colb + 1)));
                const float ns = p.bnb.neg_slope;
                const uint32_t* yp = reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(p.bnb.y) +
                                                                       static_cast<long long>(row0) * p.bnb.ldy + colb);
                const long long ystep = p.bnb.ldy >> 1;                   // row stride in 4-byte words
                uint64_t sa = 0ull, qa = 0ull;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                  uint32_t yw[16];
#pragma unroll
                  for (int r = 0; r < 16; ++r)
                    yw[r] = ((okmask >> (half * 16 + r)) & 1u) ? __ldg(yp + (half * 16 + r) * ystep) : 0u;
#pragma unroll
                  for (int rr = 0; rr < 16; ++rr) {
                    const int r = half * 16 + rr;
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(sbase + r * 128 + ((jc ^ (r & 7)) << 4));
                    const uint64_t y2 = pack_f32x2(yw[rr] << 16, yw[rr] & 0xffff0000u);
                    const uint64_t z2 = fma_f32x2(y2, sc2, sh2);
                    const float z0 = __uint_as_float(static_cast<uint32_t>(z2)), z1 = __uint_as_float(static_cast<uint32_t>(z2 >> 32));
                    const float f0 = __uint_as_float(w << 16), f1 = __uint_as_float(w & 0xffff0000u);
                    const float g0 = z0 > 0.f ? f0 : f0 * ns;
                    const float g1 = z1 > 0.f ? f1 : f1 * ns;
                    const uint64_t g2 = pack_f32x2(__float_as_uint(g0), __float_as_uint(g1));
                    sa = add_f32x2(sa, g2);
                    qa = fma_f32x2(g2, y2, qa);
                  }
                }
                s0 = __uint_as_float(static_cast<uint32_t>(sa)); s1 = __uint_as_float(static_cast<uint32_t>(sa >> 32));
                q0 = (__uint_as_float(static_cast<uint32_t>(qa)) - __ldg(p.bnb.mean + colb) * s0) * __ldg(p.bnb.rstd + colb);
                q1 = (__uint_as_float(static_cast<uint32_t>(qa >> 32)) - __ldg(p.bnb.mean + colb + 1) * s1) * __ldg(p.bnb.rstd + colb + 1);
              }
            }
          }
          if (do_stats) {
            const int pc = hf * EPI_COLS + 64 * b + 2 * lane;
            *reinterpret_cast<float2*>(part + (qd * 2 + 0) * MAX_BN + pc) = make_float2(s0, s1);
            *reinterpret_cast<float2*>(part + (qd * 2 + 1) * MAX_BN + pc) = make_float2(q0, q1);
          }
        }
        if (!released) release_accumulator<CG>(&tmem_empty[acc], lane);
      } else {
        // ---------------- 32-column chunks: f32 outputs, head epilogues, the direct (non-TMA) bf16 path ----------------
        // head state (one batch row per thread)
        float run_max = -INFINITY, run_sum = 0.f;
        int label = -1;
        float xn = 1.f, lse = 0.f;
        if (EPI == XV_EPI_HEAD_FWD || EPI == XV_EPI_HEAD_BWD) {
          if (row_ok) {
            label = p.head.labels[m];
            if (p.head.type != XV_HEAD_SOFTMAX) xn = p.head.xnorm[m];
            if (EPI == XV_EPI_HEAD_BWD) lse = p.head.lse[m];
          }
        }
        if (EPI == XV_EPI_HEAD_BWD && do_stats) {      // column sums of chunks this warp does not visit stay zero
#pragma unroll
          for (int j = lane; j < EPI_COLS; j += 32) part[(qd * 2) * MAX_BN + hf * EPI_COLS + j] = 0.f;
        }

        for (int c = 0; c * 32 < ncols; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + c * 32, r);
          tmem_ld_wait();
          const int nc0 = cbase + c * 32;
          const bool full_chunk = (nc0 + 32 <= p.N);
          const bool last_chunk = (c + 1) * 32 >= ncols;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);

          if (EPI == XV_EPI_BF16) {
            // direct bf16 path: gradient fan-in (read-modify-write) or an output the TMA unit cannot address
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += (nc0 + j < p.N) ? __ldg(p.bias + nc0 + j) : 0.f;
            }
            if (row_ok) {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(m) * p.ldc + nc0;
              if (full_chunk) {
                if (p.accumulate) {      // gradient fan-in: add the tile already in memory (tiles are exclusive)
                  const uint4* o4 = reinterpret_cast<const uint4*>(dst);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint4 o = o4[j];
                    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
                      v[8 * j + 2 * k] += f.x;
                      v[8 * j + 2 * k + 1] += f.y;
                    }
                  }
                }
                uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  d4[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                     pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (nc0 + j < p.N)
                    dst[j] = __float2bfloat16(v[j] + (p.accumulate ? __bfloat162float(dst[j]) : 0.f));
              }
            }
          } else if (EPI == XV_EPI_F32) {
            if (p.bias != nullptr && split == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += (nc0 + j < p.N) ? __ldg(p.bias + nc0 + j) : 0.f;
            }
            if (tma_out) {
              // one 32-column f32 chunk = one 128-byte store box; split-K partial tiles are combined by the TMA
              // reduce-add unit in L2 (no per-thread atomics)
              if (lane == 0) bulk_wait_read_all();
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 8; ++j)
                stage_store16(stg, lane, j,
                              make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                         __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3])));
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                if (p.splits > 1) tma_reduce_add_2d(&p.tma_out, stg, nc0, row0);
                else tma_store_2d(&p.tma_out, stg, nc0, row0);
                bulk_commit();
              }
            } else if (row_ok) {
              float* dst = reinterpret_cast<float*>(p.out) + static_cast<long long>(m) * p.ldc + nc0;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (nc0 + j < p.N) {
                  if (p.splits > 1) atomicAdd(dst + j, v[j]);
                  else dst[j] = v[j];
                }
              }
            }
          } else if (EPI == XV_EPI_HEAD_FWD) {
            if (row_ok) {
              const int jl = label - nc0;                 // position of the target column inside this chunk (or outside)
              if (p.bias) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += (nc0 + j < p.N) ? __ldg(p.bias + nc0 + j) : 0.f;
              }
              if (p.head.logits_out) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (nc0 + j < p.N) p.head.logits_out[static_cast<long long>(m) * p.ldc + nc0 + j] = v[j];
              }
              if (jl >= 0 && jl < 32) {                    // margin transform once per row, not per column
                float zl = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) zl = (j == jl) ? v[j] : zl;
                if (p.head.type != XV_HEAD_SOFTMAX) zl = margin_target(p.head, zl, xn).zprime;
                p.head.target_logit[m] = zl;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (j == jl) ? zl : v[j];
              }
              float cmax = -INFINITY;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                v[j] = (nc0 + j < p.N) ? v[j] : -INFINITY;
                cmax = fmaxf(cmax, v[j]);
              }
              const float nmax = fmaxf(run_max, cmax);
              float acc_s = run_sum * __expf(run_max - nmax);   // exp(-inf) = 0 on the first chunk
#pragma unroll
              for (int j = 0; j < 32; ++j) acc_s += __expf(v[j] - nmax);
              run_max = nmax;
              run_sum = acc_s;
            }
          } else {  // XV_EPI_HEAD_BWD
            float d[32];
            const int jl = label - nc0;
            float dz = 1.f;
            if (row_ok && p.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += (nc0 + j < p.N) ? __ldg(p.bias + nc0 + j) : 0.f;
            }
            if (row_ok && jl >= 0 && jl < 32) {
              float zl = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) zl = (j == jl) ? v[j] : zl;
              float dn = 0.f;
              if (p.head.type != XV_HEAD_SOFTMAX) {
                const MarginOut mo = margin_target(p.head, zl, xn);
                zl = mo.zprime; dz = mo.dz; dn = mo.dn;
              }
              if (p.head.gnorm) p.head.gnorm[m] = (__expf(zl - lse) - 1.0f) * p.head.inv_batch * dn;
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = (j == jl) ? zl : v[j];
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float g = 0.f;
              if (row_ok && nc0 + j < p.N) {
                const float pr = __expf(v[j] - lse);
                g = (j == jl) ? (pr - 1.0f) * p.head.inv_batch * dz : pr * p.head.inv_batch;
              }
              d[j] = g;
            }
            if (tma_out) {
              const int cc = c & 1;
              if (cc == 0) {
                if (lane == 0) bulk_wait_read_all();
                __syncwarp();
              }
#pragma unroll
              for (int j = 0; j < 4; ++j)
                stage_store16(stg, lane, cc * 4 + j,
                              make_uint4(pack_bf16x2(d[8 * j], d[8 * j + 1]), pack_bf16x2(d[8 * j + 2], d[8 * j + 3]),
                                         pack_bf16x2(d[8 * j + 4], d[8 * j + 5]), pack_bf16x2(d[8 * j + 6], d[8 * j + 7])));
              if (cc == 1 || last_chunk) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&p.tma_out, stg, nc0 - cc * 32, row0);
                  bulk_commit();
                }
              }
            } else if (row_ok) {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(m) * p.ldc + nc0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nc0 + j < p.N) dst[j] = __float2bfloat16(d[j]);
            }
            if (do_stats) {   // bias gradient of the plain softmax head: column sums of dLoss/dlogit
              warp_column_sums(d, lane);
              part[(qd * 2) * MAX_BN + hf * EPI_COLS + c * 32 + lane] = d[0];
            }
          }
        }
        // accumulator drained -> hand the TMEM buffer back to the MMA warp
        release_accumulator<CG>(&tmem_empty[acc], lane);

        if (EPI == XV_EPI_HEAD_FWD && row_ok) {
          p.head.part_max[static_cast<long long>(n_blk * 2 + hf) * p.M + m] = run_max;
          p.head.part_sum[static_cast<long long>(n_blk * 2 + hf) * p.M + m] = run_sum;
        }
      }
      if (do_stats) {   // combine the four row quadrants in a fixed order; one global atomic per (tile, column)
        named_bar_sync(1, NUM_EPI_WARPS * 32);
        if (et < BN && n0 + et < p.N) {
          const float* ps = part + et;
          stat_emit(p.col_sum, n0 + et, ((ps[0] + ps[2 * MAX_BN]) + ps[4 * MAX_BN]) + ps[6 * MAX_BN]);
          if (EPI == XV_EPI_BF16)
            stat_emit(p.col_sumsq, n0 + et, ((ps[MAX_BN] + ps[3 * MAX_BN]) + ps[5 * MAX_BN]) + ps[7 * MAX_BN]);
        }
      }
    }
    if (lane == 0) bulk_wait_all();     // outstanding TMA stores read this CTA's shared memory
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();      // neither CTA may exit (or free TMEM) while the peer's MMAs read its smem
  else __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int EPI, int CG>
int launch_gemm(const GemmKernelParams& kp, int grid, cudaStream_t stream) {
  auto kern = gemm_kernel<EPI, CG>;
  static bool configured = false;
  if (!configured) {
    XV_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  launch_cluster(kern, grid, NUM_THREADS, SMEM_BYTES, stream, CG, kp);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

// explicit instantiations live in xv_gemm_epi*.cu (one TU per epilogue keeps the parallel build short)
extern template int launch_gemm<0, 1>(const GemmKernelParams&, int, cudaStream_t);
extern template int launch_gemm<1, 1>(const GemmKernelParams&, int, cudaStream_t);
extern template int launch_gemm<2, 1>(const GemmKernelParams&, int, cudaStream_t);
extern template int launch_gemm<3, 1>(const GemmKernelParams&, int, cudaStream_t);
extern template int launch_gemm<0, 2>(const GemmKernelParams&, int, cudaStream_t);
extern template int launch_gemm<1, 2>(const GemmKernelParams&, int, cudaStream_t);

}  // namespace xv
