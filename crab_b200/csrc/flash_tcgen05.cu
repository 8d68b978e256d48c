// FlashAttention forward on the 5th-generation tensor cores (sm_100a) for head_dim = 128: the decoder-prefill attention
// (causal, GQA-aware) and any other bias-free hd-128 attention with TMA-describable strides.
//
//   S = Q K^T      : tcgen05.mma  M=128 (queries) x N=64 (keys) x K=16, operands in SWIZZLE_128B shared memory (TMA),
//                    fp32 scores in TENSOR MEMORY (two 64-column buffers, so S_{j+1} is computed while tile j's softmax runs)
//   softmax        : 4 warps, thread = query row = TMEM lane: tcgen05.ld the row, online max / exp2 / sum in fp32 with a
//                    LAZY rescale (the reference maximum only moves when the running maximum grew by > 2^8, so the
//                    O-accumulator correction is rare), P written as bf16 into a K-major SWIZZLE_128B tile
//   O += P V       : tcgen05.mma  M=128 x N=128 (head dim) x K=16 (keys); V is consumed straight from its row-major
//                    [key, hd] TMA tile as an MN-major B operand; O lives in TMEM (128 columns) for the whole CTA and is
//                    only read back once, normalised by 1/l, and stored as bf16.
//
//   warp 0: TMA producer (Q once, then separate 2-stage rings of 64-key K and V tiles: a K slot is free as soon as its
//           S = QK^T has completed, a V slot after its PV)                  warp 1: MMA issuer + TMEM owner
//   warps 2-5: softmax / correction / epilogue
//
// 112 KB of shared memory and 256 TMEM columns per CTA: two CTAs share an SM, so one CTA's softmax overlaps the
// other's MMAs without any intra-CTA ping-pong.  Causal q-tiles are launched heaviest-first.
//
// Replaces (for hd = 128): eager softmax(QK^T / sqrt(d) + causal mask) V of LlamaAttention / Qwen2Attention prefill
// (models/modeling_llama.py:405-450, models/qwen/modeling_qwen2.py:190-316); the mma.sync kernel in attention.cu keeps
// the hd-64 encoders, the BEATs bias and odd strides.
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace crab {

static constexpr int FT_BM = 128, FT_BN = 64, FT_HD = 128;
static constexpr int FT_STAGES = 2;
static constexpr int FT_Q_BYTES = FT_BM * FT_HD * 2;          // 32 KB: two 64-column chunks of 16 KB
static constexpr int FT_KV_TILE = FT_BN * FT_HD * 2;          // 16 KB: two 64-column chunks of 8 KB
static constexpr int FT_STAGE_BYTES = 2 * FT_KV_TILE;         // K + V
static constexpr int FT_P_BYTES = FT_BM * FT_BN * 2;          // 16 KB
static constexpr int FT_SMEM = FT_Q_BYTES + FT_STAGES * FT_STAGE_BYTES + FT_P_BYTES;  // 112 KB, 1024-aligned
static constexpr int FT_THREADS = 192;
static constexpr int FT_TMEM_COLS = 256;                      // S0 [0,64) S1 [64,128) O [128,256)
static constexpr float FT_TAU = 8.0f;                         // lazy-rescale threshold (log2 units)

struct FlashTcParams {
  __nv_bfloat16* o;
  long long o_bs, o_rs, o_hs;
  int q_row0_per_b, q_col_per_h, q_col0;     // Q tile coords: row = b * q_row0_per_b + q0, col = q_col0 + h * q_col_per_h
  int k_row_per_b, k_row_per_h, k_col_per_h, k_col0;
  int v_row_per_b, v_row_per_h, v_col_per_h, v_col0;
  int B, H, KVH, Sq, Sk;
  float scale;
  int causal;
};

__device__ __forceinline__ void tmem_alloc_n(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major SWIZZLE_128B operand (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>): in 16-byte units the
// canonical layout is ((8,n),(8,k)):((1,LBO),(8,SBO)) — 64 contiguous MN elements per 128-byte row, one row per K index,
// 8-row swizzle atoms SBO bytes apart along K, 64-element MN groups LBO bytes apart.
__device__ __forceinline__ uint64_t make_sdesc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32_bmn(int umma_m, int umma_n) {
  return make_idesc_bf16_f32(umma_m, umma_n) | (1u << 16);  // b_major = MN
}

__global__ void __launch_bounds__(FT_THREADS, 2)
flash_attn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                          const __grid_constant__ CUtensorMap tmap_v, const FlashTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4 * FT_STAGES + 6];
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t sQ = smem_base;
  const uint32_t sKV = sQ + FT_Q_BYTES;
  const uint32_t sP = sKV + FT_STAGES * FT_STAGE_BYTES;
  const uint32_t bar0 = smem_u32(bars);
  auto k_full = [&](int s) { return bar0 + 8u * s; };
  auto k_empty = [&](int s) { return bar0 + 8u * (FT_STAGES + s); };
  auto v_full = [&](int s) { return bar0 + 8u * (2 * FT_STAGES + s); };
  auto v_empty = [&](int s) { return bar0 + 8u * (3 * FT_STAGES + s); };
  const uint32_t q_full = bar0 + 8u * (4 * FT_STAGES);
  auto s_full = [&](int a) { return bar0 + 8u * (4 * FT_STAGES + 1 + a); };
  const uint32_t p_full = bar0 + 8u * (4 * FT_STAGES + 3);
  const uint32_t pv_done = bar0 + 8u * (4 * FT_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = (int)gridDim.x - 1 - (int)blockIdx.x;  // heaviest (most keys under the causal mask) first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.KVH);
  const int q0 = m_blk * FT_BM;
  const int off = p.Sk - p.Sq;  // causal: key j visible to query i iff j <= i + off
  int n_tiles = (p.Sk + FT_BN - 1) / FT_BN;
  if (p.causal) {
    const int last_key = min(p.Sk - 1, q0 + FT_BM - 1 + off);
    n_tiles = min(n_tiles, last_key / FT_BN + 1);
  }

  if (threadIdx.x == 0) {
    if ((smem_base & 1023u) != 0) { printf("crab: flash_tcgen05 smem misaligned\n"); __trap(); }
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < FT_STAGES; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
    mbar_init(q_full, 1);
    mbar_init(s_full(0), 1);
    mbar_init(s_full(1), 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_n(smem_u32(&tmem_slot_s), FT_TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot_s;
  const uint32_t tS0 = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      const int qrow = b * p.q_row0_per_b + q0, qcol = p.q_col0 + h * p.q_col_per_h;
      mbar_arrive_expect_tx(q_full, FT_Q_BYTES);
      tma_load_2d(sQ, &tmap_q, q_full, qcol, qrow);
      tma_load_2d(sQ + FT_Q_BYTES / 2, &tmap_q, q_full, qcol + 64, qrow);
      const int krow = b * p.k_row_per_b + kvh * p.k_row_per_h, kcol = p.k_col0 + kvh * p.k_col_per_h;
      const int vrow = b * p.v_row_per_b + kvh * p.v_row_per_h, vcol = p.v_col0 + kvh * p.v_col_per_h;
      uint32_t stage = 0, phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t sk = sKV + stage * FT_STAGE_BYTES, sv = sk + FT_KV_TILE;
        mbar_wait(k_empty(stage), phase ^ 1);   // S_{j-2} has completed
        mbar_arrive_expect_tx(k_full(stage), FT_KV_TILE);
        tma_load_2d(sk, &tmap_k, k_full(stage), kcol, krow + j * FT_BN);
        tma_load_2d(sk + FT_KV_TILE / 2, &tmap_k, k_full(stage), kcol + 64, krow + j * FT_BN);
        mbar_wait(v_empty(stage), phase ^ 1);   // PV_{j-2} has completed
        mbar_arrive_expect_tx(v_full(stage), FT_KV_TILE);
        tma_load_2d(sv, &tmap_v, v_full(stage), vcol, vrow + j * FT_BN);
        tma_load_2d(sv + FT_KV_TILE / 2, &tmap_v, v_full(stage), vcol + 64, vrow + j * FT_BN);
        if (++stage == FT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(FT_BM, FT_BN);
      constexpr uint32_t idesc_o = make_idesc_bf16_f32_bmn(FT_BM, FT_HD);
      auto issue_s = [&](int j) {  // S_j = Q K_j^T  -> TMEM buffer j & 1
        const uint32_t stage = (uint32_t)j % FT_STAGES, phase = ((uint32_t)j / FT_STAGES) & 1;
        mbar_wait(k_full(stage), phase);
        tc_fence_after();
        const uint32_t sk = sKV + stage * FT_STAGE_BYTES;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint64_t da = make_sdesc_sw128(sQ + c * (FT_Q_BYTES / 2));
          const uint64_t db = make_sdesc_sw128(sk + c * (FT_KV_TILE / 2));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tS0 + (uint32_t)(j & 1) * FT_BN, da + 2u * k, db + 2u * k, idesc_s, (c | k) != 0);
        }
        umma_commit(k_empty(stage));
        umma_commit(s_full(j & 1));
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_s(j + 1);
        const uint32_t stage = (uint32_t)j % FT_STAGES;
        mbar_wait(v_full(stage), ((uint32_t)j / FT_STAGES) & 1);
        mbar_wait(p_full, (uint32_t)j & 1);
        tc_fence_after();
        const uint32_t sv = sKV + stage * FT_STAGE_BYTES + FT_KV_TILE;
        const uint64_t da = make_sdesc_sw128(sP);
#pragma unroll
        for (int k = 0; k < FT_BN / 16; ++k) {  // 16 keys per instruction = two 8-key swizzle atoms = 2048 B of the V tile
          const uint64_t db = make_sdesc_mn_sw128(sv + k * 2048, FT_KV_TILE / 2, 1024);
          umma_bf16_ss(tO, da + 2u * k, db, idesc_o, (j | k) != 0);
        }
        umma_commit(v_empty(stage));
        umma_commit(pv_done);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;     // query row within the tile == TMEM lane
    const int qi = q0 + row;                 // global query index
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    float m_ref = -INFINITY, l_run = 0.f;
    const uint32_t p_row = sP + (uint32_t)row * 128u;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full(j & 1), ((uint32_t)j >> 1) & 1);
      tc_fence_after();
      uint32_t sr[64];
      tmem_ld_32x32b_x32(tS0 + lane_addr + (uint32_t)(j & 1) * FT_BN, sr);
      tmem_ld_32x32b_x32(tS0 + lane_addr + (uint32_t)(j & 1) * FT_BN + 32, sr + 32);
      tmem_ld_wait();
      const int k0 = j * FT_BN;
      const bool edge = (k0 + FT_BN > p.Sk) || (p.causal && (k0 + FT_BN - 1 > q0 + off));
      float mx = -INFINITY;
      if (edge) {
        const int kmax = p.causal ? min(p.Sk - 1, qi + off) : p.Sk - 1;  // last visible key of this row
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float v = (k0 + i <= kmax) ? __uint_as_float(sr[i]) : -INFINITY;
          sr[i] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(sr[i]));
      }
      mx *= sl2;  // sl2 > 0
      // lazy rescale: keep the old reference unless the row maximum outgrew it by more than TAU
      float corr = 1.f;
      bool need = false;
      if (mx > m_ref + FT_TAU) {  // also true for the first tile (m_ref = -inf) unless the whole row is masked
        corr = exp2f(m_ref - mx);  // 0 when m_ref = -inf
        need = (j > 0);
        m_ref = mx;
        l_run *= corr;
      }
      const float mref_s = (m_ref == -INFINITY) ? 0.f : m_ref;
      float rs = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float p0 = exp2f(fmaf(__uint_as_float(sr[2 * i]), sl2, -mref_s));
        const float p1 = exp2f(fmaf(__uint_as_float(sr[2 * i + 1]), sl2, -mref_s));
        rs += p0 + p1;
        pk[i] = pack_bf16x2(p0, p1);
      }
      l_run += rs;
      if (k0 + FT_BN > p.Sk) {
        // the tile reaches past the last key: its P columns are 0, but 0 x NaN would still poison O if the memory behind the
        // tensor holds NaN bit patterns (an uninitialised cache).  Zero the V rows past Sk in shared memory before PV_j reads them.
        const uint32_t stage_v = (uint32_t)j % FT_STAGES;
        mbar_wait(v_full(stage_v), ((uint32_t)j / FT_STAGES) & 1);   // the tile has landed (a second waiter next to the MMA warp)
        const uint32_t sv = sKV + stage_v * FT_STAGE_BYTES + FT_KV_TILE;
        const int r0 = max(0, p.Sk - k0), nrows = FT_BN - r0;
        for (int i = quarter * 32 + lane; i < nrows * 16; i += 128) {
          const int r = r0 + (i >> 4), piece = i & 15;
          const uint32_t addr = sv + (uint32_t)((piece >> 3) * (FT_KV_TILE / 2) + r * 128 + (piece & 7) * 16);
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
        }
      }
      // P buffer free and O stable only once PV_{j-1} has completed
      if (j > 0) {
        mbar_wait(pv_done, (uint32_t)(j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {  // tcgen05.ld/st are warp-collective: the whole warp corrects its 32 rows
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32b_x32(tO + lane_addr + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
            tmem_st_32x32b_x32(tO + lane_addr + c * 32, ov);
          }
          tmem_st_wait();
        }
      }
      // P row -> K-major SWIZZLE_128B tile (16-byte chunk c of row r at chunk c ^ (r & 7))
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t addr = p_row + (uint32_t)((c ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * c]), "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]),
                     "r"(pk[4 * c + 3]) : "memory");
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 ----
    mbar_wait(pv_done, (uint32_t)(n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    __nv_bfloat16* og = p.o + b * p.o_bs + h * p.o_hs + (long long)qi * p.o_rs;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t ov[32];
      tmem_ld_32x32b_x32(tO + lane_addr + c * 32, ov);
      tmem_ld_wait();
      if (qi < p.Sq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv);
          q.y = pack_bf16x2(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv);
          q.z = pack_bf16x2(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv);
          q.w = pack_bf16x2(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(og + c * 32 + g * 8) = q;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, FT_TMEM_COLS); }
}

// Can (base, batch stride, row stride, head stride) be walked as a 2-D [rows, row_stride] tensor by TMA?
static bool tma_view(const void* base, long long bs, long long rs, long long hs, int heads, int B, int* row_per_b,
                     int* row_per_h, int* col_per_h, long long* rows_total) {
  if (rs < FT_HD || rs % 8 != 0 || ((uintptr_t)base % 16) != 0) return false;
  if (bs % rs != 0) return false;
  *row_per_b = (int)(bs / rs);
  if (hs % rs == 0) { *row_per_h = (int)(hs / rs); *col_per_h = 0; }            // head-major (KV cache)
  else if ((long long)heads * hs <= rs && hs >= FT_HD) { *row_per_h = 0; *col_per_h = (int)hs; }  // heads side by side in a row
  else return false;
  *rows_total = (long long)B * (bs / rs);
  return true;
}

// -1: not eligible (caller falls back to the mma.sync kernel); 0: launched; < -1: error
int flash_attn_tcgen05_try(const crab_attn_args* a, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("CRAB_FLASH_TC"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled || a->head_dim != FT_HD || a->gate != nullptr || a->bias_table != nullptr || a->Sq < FT_BM) return -1;
  if (a->o_hs % 8 != 0 || a->o_rs % 8 != 0 || a->o_bs % 8 != 0 || ((uintptr_t)a->o % 16) != 0) return -1;
  FlashTcParams p;
  long long q_rows, k_rows, v_rows;
  int dummy;
  if (!tma_view(a->q, a->q_bs, a->q_rs, a->q_hs, a->H, a->B, &p.q_row0_per_b, &dummy, &p.q_col_per_h, &q_rows)) return -1;
  if (dummy != 0 || a->q_hs % a->q_rs == 0) return -1;  // Q heads must sit side by side in the row
  if (!tma_view(a->k, a->k_bs, a->k_rs, a->k_hs, a->KVH, a->B, &p.k_row_per_b, &p.k_row_per_h, &p.k_col_per_h, &k_rows)) return -1;
  if (!tma_view(a->v, a->v_bs, a->v_rs, a->v_hs, a->KVH, a->B, &p.v_row_per_b, &p.v_row_per_h, &p.v_col_per_h, &v_rows)) return -1;
  p.q_col0 = p.k_col0 = p.v_col0 = 0;
  p.o = (__nv_bfloat16*)a->o; p.o_bs = a->o_bs; p.o_rs = a->o_rs; p.o_hs = a->o_hs;
  p.B = a->B; p.H = a->H; p.KVH = a->KVH; p.Sq = a->Sq; p.Sk = a->Sk; p.scale = a->scale; p.causal = a->causal;
  CUtensorMap tq, tk, tv;
  int rc = encode_tmap_bf16_2d(&tq, a->q, (uint64_t)q_rows, (uint64_t)a->q_rs, (uint64_t)a->q_rs, FT_BM, 64);
  if (rc != 0) return rc;
  rc = encode_tmap_bf16_2d(&tk, a->k, (uint64_t)k_rows, (uint64_t)a->k_rs, (uint64_t)a->k_rs, FT_BN, 64);
  if (rc != 0) return rc;
  rc = encode_tmap_bf16_2d(&tv, a->v, (uint64_t)v_rows, (uint64_t)a->v_rs, (uint64_t)a->v_rs, FT_BN, 64);
  if (rc != 0) return rc;
  static DeviceOnce attr_once;
  if (first_on_device(attr_once)) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    if (e != cudaSuccess) return set_error(CRAB_ERR_CUDA, "flash_attn_tcgen05: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  }
  dim3 grid((a->Sq + FT_BM - 1) / FT_BM, a->H, a->B);
  flash_attn_tcgen05_kernel<<<grid, FT_THREADS, FT_SMEM, st>>>(tq, tk, tv, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(CRAB_ERR_CUDA, "flash_attn_tcgen05 launch failed: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace crab
