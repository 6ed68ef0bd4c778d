// Persistent, warp-specialised implicit-GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> 4-stage smem ring -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM,
//   128 x 256 tile, double-buffered accumulator = all 512 TMEM columns) -> tcgen05.ld epilogue.
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM
// allocator, warp 3 idle, warps 4..7 = epilogue (TMEM lane quadrant = warp_idx % 4).
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
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;   // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int CHUNK_BYTES = 64 * BLOCK_K * 2;          // one MN-major 64x64 box = 8 KB
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 256;
constexpr int STATS_BYTES = 4 * 2 * BLOCK_N * 4;       // [epilogue warp][sum|sumsq][col]
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STATS_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;

struct alignas(64) GemmKernelParams {
  CUtensorMap tma_a;
  CUtensorMap tma_b;
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

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_kernel(const __grid_constant__ GemmKernelParams p) {
  const bool A_MN = p.a_mn != 0, B_MN = p.b_mn != 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  float* s_stats = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STATS_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;      // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    prefetch_tmap(&p.tma_a);
    prefetch_tmap(&p.tma_b);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int total_tiles = p.num_m * p.num_n * p.splits;

  if (warp_idx == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.num_n;
        const int m_blk = (tile / p.num_n) % p.num_m;
        const int split = tile / (p.num_n * p.num_m);
        const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          const int kk = kb * BLOCK_K;
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          if (!A_MN) {
            const int col = p.a_div ? (kk % p.a_div) : kk;
            const int row = m0 + (p.a_div ? (kk / p.a_div) * p.a_tap : 0);
            tma_load_2d(sa, &p.tma_a, &full_bar[stage], col, row);
          } else {
            const int col = p.a_div ? (m0 % p.a_div) : m0;
            const int row = kk + (p.a_div ? (m0 / p.a_div) * p.a_tap : 0);
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_2d(sa + c * CHUNK_BYTES, &p.tma_a, &full_bar[stage], col + 64 * c, row);
          }
          if (!B_MN) {
            const int col = p.b_div ? (kk % p.b_div) : kk;
            const int row = n0 + (p.b_div ? (kk / p.b_div) * p.b_tap : 0);
            tma_load_2d(sb, &p.tma_b, &full_bar[stage], col, row);
          } else {
            const int col = p.b_div ? (n0 % p.b_div) : n0;
            const int row = kk + (p.b_div ? (n0 / p.b_div) * p.b_tap : 0);
#pragma unroll
            for (int c = 0; c < BLOCK_N / 64; ++c)
              tma_load_2d(sb + c * CHUNK_BYTES, &p.tma_b, &full_bar[stage], col + 64 * c, row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // K-major SW128: 8-row groups 1024 B apart (SBO), LBO unused.  MN-major SW128: 64-wide MN chunks
      // CHUNK_BYTES apart (LBO), 8-k-row groups 1024 B apart (SBO).
      const uint32_t a_lbo = A_MN ? CHUNK_BYTES : 16, b_lbo = B_MN ? CHUNK_BYTES : 16;
      const uint32_t a_kstep = A_MN ? (UMMA_K * 128) : (UMMA_K * 2);   // bytes per UMMA_K
      const uint32_t b_kstep = B_MN ? (UMMA_K * 128) : (UMMA_K * 2);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
        const int split = tile / (p.num_n * p.num_m);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(a_addr + k * a_kstep, a_lbo, 1024);
            const uint64_t db = make_smem_desc(b_addr + k * b_kstep, b_lbo, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);       // accumulator complete -> epilogue
      }
    }
  } else if (warp_idx >= 4) {
    // ============================== epilogue ==============================
    const int ew = warp_idx - 4;            // == warp_idx % 4 == TMEM lane quadrant
    const int et = threadIdx.x - 128;       // 0..127
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int n_blk = tile % p.num_n;
      const int m_blk = (tile / p.num_n) % p.num_m;
      const int split = tile / (p.num_n * p.num_m);
      const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int m = m0 + ew * 32 + lane;
      const bool row_ok = m < p.M;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BLOCK_N;

      const bool do_stats = (EPI == XV_EPI_BF16 || EPI == XV_EPI_HEAD_BWD) && p.col_sum != nullptr;
      bool row_valid = row_ok;
      if (EPI == XV_EPI_BF16 && p.seg_len > 0) row_valid = row_ok && ((m % p.seg_len) < p.seg_valid);

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

      const int ncols = min(BLOCK_N, p.N - n0);
      for (int c = 0; c * 32 < ncols; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tmem_ld_wait();
        const int nc0 = n0 + c * 32;
        const bool full_chunk = (nc0 + 32 <= p.N);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);

        if (EPI == XV_EPI_BF16) {
          float q[32];
          if (do_stats) {
#pragma unroll
            for (int j = 0; j < 32; ++j) q[j] = row_valid ? v[j] : 0.f;
          }
          if (row_ok) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(m) * p.ldc + nc0;
            if (full_chunk) {
              uint32_t pk[16];
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
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float lo = v[2 * j], hi = v[2 * j + 1];
                if (p.bias) { lo += __ldg(p.bias + nc0 + 2 * j); hi += __ldg(p.bias + nc0 + 2 * j + 1); }
                pk[j] = pack_bf16x2(lo, hi);
              }
              uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
              for (int j = 0; j < 4; ++j) d4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nc0 + j < p.N)
                  dst[j] = __float2bfloat16(v[j] + (p.bias ? __ldg(p.bias + nc0 + j) : 0.f) +
                                            (p.accumulate ? __bfloat162float(dst[j]) : 0.f));
            }
          }
          if (do_stats) {
            float s[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { s[j] = q[j]; q[j] = q[j] * q[j]; }
            warp_column_sums(s, lane);
            warp_column_sums(q, lane);
            s_stats[(ew * 2 + 0) * BLOCK_N + c * 32 + lane] = s[0];
            s_stats[(ew * 2 + 1) * BLOCK_N + c * 32 + lane] = q[0];
          }
        } else if (EPI == XV_EPI_F32) {
          if (row_ok) {
            float* dst = reinterpret_cast<float*>(p.out) + static_cast<long long>(m) * p.ldc + nc0;
            const bool add_bias = p.bias != nullptr && split == 0;
            if (p.splits > 1) {
              if (full_chunk) {      // 16-byte vector reductions: 8 RED instructions per 32 columns instead of 32
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                  if (add_bias) {
                    o.x += __ldg(p.bias + nc0 + 4 * j); o.y += __ldg(p.bias + nc0 + 4 * j + 1);
                    o.z += __ldg(p.bias + nc0 + 4 * j + 2); o.w += __ldg(p.bias + nc0 + 4 * j + 3);
                  }
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(o.x), "f"(o.y),
                               "f"(o.z), "f"(o.w) : "memory");
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (nc0 + j < p.N) atomicAdd(dst + j, v[j] + (add_bias ? __ldg(p.bias + nc0 + j) : 0.f));
              }
            } else if (full_chunk) {
              float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (add_bias) {
                  o.x += __ldg(p.bias + nc0 + 4 * j); o.y += __ldg(p.bias + nc0 + 4 * j + 1);
                  o.z += __ldg(p.bias + nc0 + 4 * j + 2); o.w += __ldg(p.bias + nc0 + 4 * j + 3);
                }
                d4[j] = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nc0 + j < p.N) dst[j] = v[j] + (add_bias ? __ldg(p.bias + nc0 + j) : 0.f);
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
          if (row_ok) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(m) * p.ldc + nc0;
            if (full_chunk) {
              uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                d4[j] = make_uint4(pack_bf16x2(d[8 * j], d[8 * j + 1]), pack_bf16x2(d[8 * j + 2], d[8 * j + 3]),
                                   pack_bf16x2(d[8 * j + 4], d[8 * j + 5]), pack_bf16x2(d[8 * j + 6], d[8 * j + 7]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nc0 + j < p.N) dst[j] = __float2bfloat16(d[j]);
            }
          }
          if (do_stats) {   // bias gradient of the plain softmax head: column sums of dLoss/dlogit
            warp_column_sums(d, lane);
            s_stats[(ew * 2 + 0) * BLOCK_N + c * 32 + lane] = d[0];
          }
        }
      }
      // accumulator drained -> hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);

      if (EPI == XV_EPI_HEAD_FWD && row_ok) {
        p.head.part_max[static_cast<long long>(n_blk) * p.M + m] = run_max;
        p.head.part_sum[static_cast<long long>(n_blk) * p.M + m] = run_sum;
      }
      if (do_stats) {
        named_bar_sync(1, 128);
        for (int col = et; col < ncols; col += 128) {
          float s = 0.f, q = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            s += s_stats[(w * 2 + 0) * BLOCK_N + col];
            if (EPI == XV_EPI_BF16) q += s_stats[(w * 2 + 1) * BLOCK_N + col];
          }
          atomicAdd(p.col_sum + n0 + col, s);
          if (EPI == XV_EPI_BF16) atomicAdd(p.col_sumsq + n0 + col, q);
        }
        named_bar_sync(1, 128);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int EPI>
int launch_gemm(const GemmKernelParams& kp, int grid, cudaStream_t stream) {
  auto kern = gemm_kernel<EPI>;
  static bool configured = false;
  if (!configured) {
    XV_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(kp);
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}



// explicit instantiations live in xv_gemm_epi*.cu (one TU per epilogue keeps the parallel build short)
extern template int launch_gemm<0>(const GemmKernelParams&, int, cudaStream_t);
extern template int launch_gemm<1>(const GemmKernelParams&, int, cudaStream_t);
extern template int launch_gemm<2>(const GemmKernelParams&, int, cudaStream_t);
extern template int launch_gemm<3>(const GemmKernelParams&, int, cudaStream_t);

}  // namespace xv
