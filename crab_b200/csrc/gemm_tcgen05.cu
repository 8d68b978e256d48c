// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] . B[N,K]^T)
//
//   warp 0 (1 lane)  : TMA producer  — cp.async.bulk.tensor 128x64 A tiles and BNx64 B tiles, SWIZZLE_128B
//   warp 1 (1 lane)  : MMA issuer    — tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16, accumulators in TMEM
//   warps 2..5       : epilogue      — tcgen05.ld 32x32b, bias/activation/residual in fp32, 16-byte global stores
//
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, double-buffered
// accumulator so the epilogue of tile i overlaps the MMAs of tile i+1), and a static persistent tile schedule with
// grouped rasterisation (GROUP_M m-tiles share the B panel in L2).
//
// This one kernel carries every dense contraction on Crab's hot path (see include/crab_b200.h for the list).
#include <stdlib.h>
#include <unordered_map>
#include <mutex>

#include "host_common.h"
#include "ptx.cuh"

namespace crab {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int UMMA_K = 16;
static constexpr int GROUP_M = 16;
static constexpr int GEMM_THREADS = 192;
static constexpr int GEMM2_DEFAULT = 1;  // CTA-pair kernel: auto (see use_gemm2)

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;  // double-buffered fp32 accumulator
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
  void* C;
  const float* bias;
  const __nv_bfloat16* residual;
  int M, N, K;
  int ldc, ldr;
  float res_scale, out_scale;
  int act, out_dtype;
  int lora_groups;  // CRAB_ACT_LORA_Z: number of 11-wide groups
};

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == CRAB_ACT_QUICK_GELU) return x / (1.0f + __expf(-1.702f * x));
  if (act == CRAB_ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
  return x;
}

// Store 32 consecutive output values of one row (guarded against the N edge in 8-column groups).
__device__ __forceinline__ void store_row32(const GemmParams& p, int row, int col0, int n_limit, float* v) {
  if (p.residual != nullptr) {
    const __nv_bfloat16* r = p.residual + (size_t)row * p.ldr + col0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (col0 + g * 8 < n_limit) {
        uint4 q = __ldg(reinterpret_cast<const uint4*>(r + g * 8));
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[g * 8 + 2 * j] += p.res_scale * bf16lo(w[j]);
          v[g * 8 + 2 * j + 1] += p.res_scale * bf16hi(w[j]);
        }
      }
    }
  }
  if (p.out_dtype == CRAB_BF16) {
    __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc + col0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (col0 + g * 8 < n_limit) {
        uint4 q;
        q.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
        q.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
        q.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
        q.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
        *reinterpret_cast<uint4*>(c + g * 8) = q;
      }
    }
  } else {
    float* c = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col0;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (col0 + g * 4 < n_limit) {
        *reinterpret_cast<float4*>(c + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
      }
    }
  }
}

// Epilogue of one accumulator tile for one warp: this thread owns output row `row` (TMEM lane), BN columns at taddr.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, int row, int n_blk, uint32_t taddr) {
  const bool row_ok = row < p.M;
  if (p.act == CRAB_ACT_SWIGLU) {
    // packed columns: [64 gate | 64 up] per 128; output column = (pc / 128) * 64 + pc % 128
    const int n_out = p.N >> 1;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      if ((c & 3) >= 2) continue;
      const int pc0 = n_blk * BN + c * 32;
      if (pc0 >= p.N) break;
      uint32_t g[32], u[32];
      tmem_ld_32x32b_x32(taddr + c * 32, g);
      tmem_ld_32x32b_x32(taddr + (c + 2) * 32, u);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float gv = __uint_as_float(g[j]);
        v[j] = (gv / (1.0f + __expf(-gv))) * __uint_as_float(u[j]) * p.out_scale;
      }
      const int oc0 = (pc0 >> 7) * 64 + (pc0 & 127);
      if (row_ok) store_row32(p, row, oc0, n_out, v);
    }
  } else if (p.act == CRAB_ACT_LORA_Z) {
    if constexpr (BN == 64) {
      uint32_t r[64];
      tmem_ld_32x32b_x32(taddr, r);
      tmem_ld_32x32b_x32(taddr + 32, r + 32);
      tmem_ld_wait();
      if (row_ok) {
        __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g < p.lora_groups) {
            const float l0 = __uint_as_float(r[g * 11 + 0]), l1 = __uint_as_float(r[g * 11 + 1]),
                        l2 = __uint_as_float(r[g * 11 + 2]);
            const float mx = fmaxf(l0, fmaxf(l1, l2));
            const float e0 = __expf(l0 - mx), e1 = __expf(l1 - mx), e2 = __expf(l2 - mx);
            const float inv = p.out_scale / (e0 + e1 + e2);
            const float rw[3] = {e0 * inv, e1 * inv, e2 * inv};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              uint4 q;
              q.x = pack_bf16x2(rw[i] * __uint_as_float(r[g * 11 + 3]), rw[i] * __uint_as_float(r[g * 11 + 4]));
              q.y = pack_bf16x2(rw[i] * __uint_as_float(r[g * 11 + 5]), rw[i] * __uint_as_float(r[g * 11 + 6]));
              q.z = pack_bf16x2(rw[i] * __uint_as_float(r[g * 11 + 7]), rw[i] * __uint_as_float(r[g * 11 + 8]));
              q.w = pack_bf16x2(rw[i] * __uint_as_float(r[g * 11 + 9]), rw[i] * __uint_as_float(r[g * 11 + 10]));
              *reinterpret_cast<uint4*>(c + g * 24 + i * 8) = q;
            }
          }
        }
      }
    }
  } else {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int col0 = n_blk * BN + c * 32;
      if (col0 >= p.N) break;
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr + c * 32, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (p.bias != nullptr) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (col0 + g * 4 < p.N) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + g * 4));
            v[g * 4 + 0] += b.x; v[g * 4 + 1] += b.y; v[g * 4 + 2] += b.z; v[g * 4 + 3] += b.w;
          }
        }
      }
      if (p.act != CRAB_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = act_apply(v[j], p.act);
      }
      if (p.out_scale != 1.0f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
      }
      if (row_ok) store_row32(p, row, col0, p.N, v);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  // barrier layout (8 bytes each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    const int per_group = GROUP_M * n_tiles;
    const int group = tile / per_group;
    const int first_m = group * GROUP_M;
    const int gsz = min(m_tiles - first_m, GROUP_M);
    const int local = tile - group * per_group;
    m_blk = first_m + local % gsz;
    n_blk = local / gsz;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          tma_load_2d(sa, &tmap_a, full_bar(stage), kb * BK, m_blk * BM);
          tma_load_2d(sa + Cfg::A_BYTES, &tmap_b, full_bar(stage), kb * BK, n_blk * BN);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc_bf16_f32(BM, BN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t da = make_sdesc_sw128(sa);
          const uint64_t db = make_sdesc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // +32 bytes per K=16 step inside the 128-byte swizzle row: start-address field += 2
            umma_bf16_ss(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0);
          }
          umma_commit(empty_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, m_blk, n_blk);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
      epilogue_tile<BN>(p, row, n_blk, taddr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// --------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster — the two SMs of a TPC — compute one 256 x 256 tile.
// Each CTA stages ITS 128 rows of A and ITS 128 of the 256 B rows (32 KB per k-block instead of 48 KB: the B operand is
// shared through the pair, which halves the shared-memory fill per SM and frees room for a 6-stage ring); the leader CTA
// issues ONE tcgen05.mma.cta_group::2 M=256 N=256 K=16 per step for both tensor cores; each CTA's TMEM receives its own
// 128 x 256 fp32 accumulator and each CTA runs the epilogue for its rows.
//   * TMA: both CTAs' loads signal the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2 with the leader's
//     barrier address from mapa); the leader expects the bytes of both halves.
//   * smem slot release and accumulator-ready: tcgen05.commit.cta_group::2 ... multicast::cluster to both CTAs.
//   * accumulator-drained: the peer's epilogue warps arrive on the leader's barrier through shared::cluster.
// --------------------------------------------------------------------------------------------------------------
struct Gemm2Cfg {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = BM * BK * 2;          // this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN / 2) * BK * 2;    // this CTA's 128 rows of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 6;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ uint32_t g2_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void g2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t g2_map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void g2_tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void g2_umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void g2_commit_both(uint32_t bar) {  // arrive on `bar` (same offset) in BOTH CTAs of the pair
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void g2_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                          const GemmParams p) {
  using Cfg = Gemm2Cfg;
  constexpr int BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = g2_cluster_rank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  const int m_tiles = (p.M + 2 * BM - 1) / (2 * BM);   // 256-row tiles
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);   // leader: its producer's arrive.expect_tx (bytes of both CTAs)
      mbar_init(empty_bar(s), 1);  // one multicast commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);  // leader: 4 epilogue warps of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  g2_cluster_sync();  // both CTAs' barriers are initialised before anybody signals across the pair
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    const int per_group = (GROUP_M / 2) * n_tiles;
    const int group = tile / per_group;
    const int first_m = group * (GROUP_M / 2);
    const int gsz = min(m_tiles - first_m, GROUP_M / 2);
    const int local = tile - group * per_group;
    m_blk = first_m + local % gsz;
    n_blk = local / gsz;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (both CTAs) =====================
      uint32_t stage = 0, phase = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          const uint32_t lbar = g2_map_to_rank(full_bar(stage), 0);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          g2_tma_load_2d(sa, &tmap_a, lbar, kb * BK, m_blk * 2 * BM + (int)rank * BM);
          g2_tma_load_2d(sa + Cfg::A_BYTES, &tmap_b, lbar, kb * BK, n_blk * BN + (int)rank * (BN / 2));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      constexpr uint32_t idesc = make_idesc_bf16_f32(2 * BM, BN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t da = make_sdesc_sw128(sa);
          const uint64_t db = make_sdesc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) g2_umma(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0);
          g2_commit_both(empty_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        g2_commit_both(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int quarter = warp & 3;
    uint32_t acc = 0, acc_phase = 0;
    const uint32_t tempty_leader0 = g2_map_to_rank(tempty_bar(0), 0), tempty_leader1 = g2_map_to_rank(tempty_bar(1), 0);
    for (int tile = pair; tile < num_tiles; tile += n_pairs) {
      int m_blk, n_blk;
      tile_coords(tile, m_blk, n_blk);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * 2 * BM + (int)rank * BM + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
      epilogue_tile<BN>(p, row, n_blk, taddr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) g2_arrive_remote(acc ? tempty_leader1 : tempty_leader0);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  g2_cluster_sync();  // nobody leaves (or frees TMEM) while the partner may still signal / read across the pair
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// --------------------------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------------------------
struct TmapKey {
  const void* base;
  uint64_t rows, cols, ld;
  uint32_t box_rows;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    h ^= k.rows * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= k.cols * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= k.ld * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    h ^= k.box_rows + (h << 6) + (h >> 2);
    return h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::mutex g_tmap_mutex;

static int get_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  TmapKey key{base, rows, cols, ld, box_rows};
  std::lock_guard<std::mutex> lock(g_tmap_mutex);
  auto it = g_tmap_cache.find(key);
  if (it != g_tmap_cache.end()) {
    *out = it->second;
    return 0;
  }
  int rc = encode_tmap_bf16_2d(out, base, rows, cols, ld, box_rows, BK);
  if (rc != 0) return rc;
  if (g_tmap_cache.size() > 65536) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *out);
  return 0;
}

template <int BN>
static int launch_gemm(const crab_gemm_args* a, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static DeviceOnce attr_once;
  if (first_on_device(attr_once)) {
    CRAB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES));
  }
  CUtensorMap ta, tb;
  int rc = get_tmap(&ta, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, BM);
  if (rc != 0) return rc;
  rc = get_tmap(&tb, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, BN);
  if (rc != 0) return rc;
  const int m_tiles = (a->M + BM - 1) / BM, n_tiles = (a->N + BN - 1) / BN;
  int ctas = sm_count();
  if (a->max_ctas > 0 && a->max_ctas < ctas) ctas = a->max_ctas;
  if (ctas > m_tiles * n_tiles) ctas = m_tiles * n_tiles;
  gemm_bf16_tcgen05_kernel<BN><<<ctas, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

// CTA-pair launch: one cluster of two CTAs per 256 x 256 tile, persistent over sm_count / 2 pairs.
static int launch_gemm2(const crab_gemm_args* a, const GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg;
  static DeviceOnce attr_once;
  if (first_on_device(attr_once)) {
    CRAB_CHECK_CUDA(cudaFuncSetAttribute(gemm2_bf16_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  CUtensorMap ta, tb;
  int rc = get_tmap(&ta, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, BM);
  if (rc != 0) return rc;
  rc = get_tmap(&tb, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, Cfg::BN / 2);
  if (rc != 0) return rc;
  const int m_tiles = (a->M + 2 * BM - 1) / (2 * BM), n_tiles = (a->N + Cfg::BN - 1) / Cfg::BN;
  int pairs = sm_count() / 2;
  if (a->max_ctas > 0 && a->max_ctas / 2 < pairs) pairs = a->max_ctas / 2 > 0 ? a->max_ctas / 2 : 1;
  if (pairs > m_tiles * n_tiles) pairs = m_tiles * n_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CRAB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm2_bf16_tcgen05_kernel, ta, tb, p));
  return CRAB_OK;
}

// CTA-pair policy (env CRAB_GEMM_2CTA / crab_set_gemm_2cta): 0 = never, 1 = auto (default), 2 = whenever the 256-wide
// tile is in use and the problem has at least two 256 x 256 tiles (tests).  Auto = the decoder-prefill regime where the
// A/B in profiles/r02_gemm_2cta.txt shows a win: long K (>= 2048: the mainloop, not the epilogue, paces the tile) and
// enough rows for >= 4 waves of pairs.
static thread_local int g_gemm2_mode = -1;   // per calling thread, like the PDL mask
static int gemm2_mode() {
  if (g_gemm2_mode < 0) { const char* e = getenv("CRAB_GEMM_2CTA"); g_gemm2_mode = e ? atoi(e) : GEMM2_DEFAULT; }
  return g_gemm2_mode;
}
static bool use_gemm2(const crab_gemm_args* a) {
  const int mode = gemm2_mode();
  if (mode <= 0 || a->act == CRAB_ACT_LORA_Z) return false;
  const long tiles2 = (long)((a->M + 255) / 256) * ((a->N + 255) / 256);
  if (mode >= 2) return tiles2 >= 2;
  return a->K >= 2048 && a->M >= 16384 && tiles2 >= 4L * (sm_count() / 2);
}

}  // namespace crab

extern "C" int crab_gemm_bf16(const crab_gemm_args* a, void* stream_) {
  using namespace crab;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CRAB_REQUIRE(a != nullptr, "crab_gemm_bf16: null args");
  CRAB_REQUIRE(a->A && a->B && a->C, "crab_gemm_bf16: null A/B/C");
  CRAB_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "crab_gemm_bf16: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  CRAB_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0 && a->lda >= a->K && a->ldb >= a->K,
               "crab_gemm_bf16: lda/ldb must be >= K and multiples of 8 (lda=%d ldb=%d K=%d)", a->lda, a->ldb, a->K);
  CRAB_REQUIRE(((uintptr_t)a->A % 16 == 0) && ((uintptr_t)a->B % 16 == 0) && ((uintptr_t)a->C % 16 == 0),
               "crab_gemm_bf16: A/B/C must be 16-byte aligned");
  CRAB_REQUIRE(a->act >= CRAB_ACT_NONE && a->act <= CRAB_ACT_LORA_Z, "crab_gemm_bf16: bad act %d", a->act);
  GemmParams p;
  p.C = a->C;
  p.bias = a->bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.ldc = a->ldc; p.ldr = a->ldr;
  p.res_scale = a->res_scale; p.out_scale = a->out_scale;
  p.act = a->act; p.out_dtype = a->out_dtype;
  p.lora_groups = 0;
  int bn = a->block_n;
  if (a->act == CRAB_ACT_LORA_Z) {
    CRAB_REQUIRE(a->N % 11 == 0 && a->N <= 44, "crab_gemm_bf16: LORA_Z needs N = 11*g, g<=4 (N=%d)", a->N);
    CRAB_REQUIRE(a->out_dtype == CRAB_BF16 && a->ldc % 8 == 0 && !a->bias && !a->residual,
                 "crab_gemm_bf16: LORA_Z writes bf16, ldc%%8==0, no bias/residual");
    p.lora_groups = a->N / 11;
    bn = 64;
  } else if (a->act == CRAB_ACT_SWIGLU) {
    CRAB_REQUIRE(a->N % 128 == 0 && !a->bias, "crab_gemm_bf16: SWIGLU needs N %% 128 == 0 and no bias (N=%d)", a->N);
    CRAB_REQUIRE(a->ldc % 8 == 0, "crab_gemm_bf16: ldc must be a multiple of 8");
    if (bn == 64) bn = 128;
  } else {
    CRAB_REQUIRE(a->N % 8 == 0, "crab_gemm_bf16: N must be a multiple of 8 (N=%d)", a->N);
    CRAB_REQUIRE(a->ldc % (a->out_dtype == CRAB_BF16 ? 8 : 4) == 0, "crab_gemm_bf16: ldc alignment");
    if (a->residual) CRAB_REQUIRE(a->ldr % 8 == 0 && ((uintptr_t)a->residual % 16 == 0), "crab_gemm_bf16: ldr/residual alignment");
    if (a->bias) CRAB_REQUIRE((uintptr_t)a->bias % 16 == 0, "crab_gemm_bf16: bias alignment");
  }
  if (bn == 0) {
    // Auto: the widest tile that still gives every SM work; skinny problems use narrower tiles for more CTAs.
    const int sms = sm_count();
    const long m_tiles = (a->M + BM - 1) / BM;
    if (m_tiles * ((a->N + 255) / 256) >= sms) bn = 256;
    else if (m_tiles * ((a->N + 127) / 128) >= sms || a->N < 128) bn = (a->N <= 64) ? 64 : 128;
    else bn = 64;
    if (a->act == CRAB_ACT_SWIGLU && bn == 64) bn = 128;
  }
  switch (bn) {
    case 64: return launch_gemm<64>(a, p, stream);
    case 128: return launch_gemm<128>(a, p, stream);
    case 256:
      if (use_gemm2(a)) return launch_gemm2(a, p, stream);
      return launch_gemm<256>(a, p, stream);
    default: return set_error(CRAB_ERR_INVALID, "crab_gemm_bf16: block_n must be 0/64/128/256 (got %d)", bn);
  }
}

extern "C" int crab_set_gemm_2cta(int mode) {
  crab::g_gemm2_mode = mode < 0 ? 0 : (mode > 2 ? 2 : mode);
  return CRAB_OK;
}
