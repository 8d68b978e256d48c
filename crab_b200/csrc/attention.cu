// Attention kernels.
//
//  flash_attn_kernel  — FlashAttention-style forward (online softmax, never materialises S): encoder self-attention
//                       (CLIP N=257, BEATs N<=96 with gated relative-position bias, Q-Former 32x32), Q-Former
//                       cross-attention (32 x {48,96,256}) and causal decoder prefill (hd=128, GQA-aware).
//                       K/V tiles are staged in XOR-swizzled shared memory with a cp.async double buffer; QK^T and
//                       PV run on mma.sync.m16n8k16 (bf16 in, fp32 accumulate); softmax statistics in fp32.
//                       Attention is <= 4% of the path's FLOPs; the tcgen05 budget went to the GEMMs.
//  attn_decode_kernel — single-query decode attention over the KV cache: pure HBM streaming (16-byte loads, one
//                       key per half-warp), split over the context when batch*heads cannot fill 148 SMs, context
//                       length read from device memory so the step can live in a CUDA graph.
#include "host_common.h"
#include "ptx.cuh"

namespace crab {

int flash_attn_tcgen05_try(const crab_attn_args* a, cudaStream_t st);  // flash_tcgen05.cu

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct FlashParams {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v; __nv_bfloat16* o;
  long long q_bs, q_rs, q_hs;  // element strides: batch, row (sequence), head
  long long k_bs, k_rs, k_hs;
  long long v_bs, v_rs, v_hs;
  long long o_bs, o_rs, o_hs;
  int B, H, KVH, Sq, Sk;
  float scale;
  int causal;
  const float* gate;   // [B, H, Sq] or null
  const float* table;  // [H, Sq, Sk] or null   (bias = gate * table)
  const int* sk_dev;   // null, or the number of keys in device memory (overrides Sk)
  int sk_add;          // added to *sk_dev (decode: *past_dev + 1 keys)
  int nsplit;          // > 1: split-KV (Sq <= 64 only): blockIdx.x = split, partial states -> ws, attn_decode_combine_kernel finishes
  float* ws;           // [(b * H * Sq + h * Sq + row) * nsplit + split][HD + 2]: un-normalised acc, running max (log2 domain), sum
  // TMA staging (head_dim 64, TMA-describable strides): the three tensors as 2-D [rows, row stride] maps, tile (row, col) origins
  int stages;          // K/V tile buffers of the cp.async path: 2, or 3 for the split-KV decode (one more 64-key step in flight per block)
  int use_tma;
  int q_row_per_b, q_col_per_h, k_row_per_b, k_row_per_h, k_col_per_h, v_row_per_b, v_row_per_h, v_col_per_h;
};

template <int HD, int MT>
struct FlashCfg {
  static constexpr int BM = 64 * MT, BN = 64, THREADS = 128;  // MT m16-tiles of query rows per warp
  static constexpr int TILE_BYTES = 64 * HD * 2;
  static constexpr int SMEM = TILE_BYTES * (MT + 4);  // Q (MT tiles) + 2 x (K, V)
  static constexpr int SMEM3 = TILE_BYTES * (MT + 6); // three K / V buffers (split-KV decode)
};

// smem tile: row-major [64][HD] bf16, 16-byte chunk c of row r stored at chunk (c ^ (r & 7))
template <int HD>
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int r, int c) {
  return base + (uint32_t)(r * HD * 2) + (uint32_t)(((c ^ (r & 7))) << 4);
}

template <int HD>
__device__ __forceinline__ void load_tile(uint32_t sbase, const __nv_bfloat16* g, long long rs, int row0, int nrows_valid) {
  constexpr int CH = HD / 8;
  for (int i = threadIdx.x; i < 64 * CH; i += 128) {
    const int r = i / CH, c = i % CH;
    const bool ok = (row0 + r) < nrows_valid;
    const __nv_bfloat16* src = g + (long long)(ok ? (row0 + r) : 0) * rs + c * 8;
    cp_async16(tile_addr<HD>(sbase, r, c), src, ok);
  }
}

// Each warp owns MT x 16 query rows, so every K / V fragment fetched from shared memory (ldmatrix) feeds MT MMAs.
// KS (key split, decode over a kv group: Sq <= 16 query rows): the four warps share the SAME 16 query rows and each takes 16 of
// a tile's 64 keys — with the rows split over the warps three of them would idle and the fourth would do a tile's 128 MMAs
// alone (2.2 us per 64-key step, measured: the kernel was compute-latency bound at 3.8 TB/s); the four partial softmax
// states are merged through shared memory after the last tile.
template <int HD, int MT, bool KS = false>
__global__ void __launch_bounds__(128, (HD == 64 && MT == 1) ? 4 : 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v, const FlashParams p) {
  using Cfg = FlashCfg<HD, MT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  // head_dim 64: a tile row is exactly one 128-byte swizzle row, and tile_addr's layout (16-byte chunk c of row r at c ^ (r & 7))
  // IS the TMA SWIZZLE_128B layout — so Q / K / V tiles can be staged by cp.async.bulk.tensor (one thread, two instructions per
  // 64-key step, completion on an mbarrier) instead of 16 cp.async per thread, with the ldmatrix addressing unchanged.
  const bool tma = (HD == 64) && p.use_tma;
  constexpr int NPG = KS ? 1 : 4;      // 16-key groups of a tile handled by one warp
  constexpr int NT8 = 2 * NPG;         // score n8-tiles per warp and tile
  constexpr int NKK = KS ? 1 : 4;      // 16-key PV steps per warp and tile
  __shared__ __align__(8) uint64_t fbars[3];   // Q, K/V buffer 0, K/V buffer 1
  const uint32_t q_bar = smem_u32(&fbars[0]);
  const uint32_t sQ = smem_u32(smem);
  const int NST = tma ? 2 : p.stages;
  const uint32_t sK0 = sQ + MT * Cfg::TILE_BYTES;
  const uint32_t sV0 = sK0 + NST * Cfg::TILE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = p.nsplit > 1 ? (int)blockIdx.x : 0;
  const int m_blk = p.nsplit > 1 ? 0 : (int)blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.KVH);
  const int q0 = m_blk * Cfg::BM;
  const __nv_bfloat16* qg = p.q + b * p.q_bs + h * p.q_hs;
  const __nv_bfloat16* kg = p.k + b * p.k_bs + kvh * p.k_hs;
  const __nv_bfloat16* vg = p.v + b * p.v_bs + kvh * p.v_hs;
  const int Sk = p.sk_dev ? *p.sk_dev + p.sk_add : p.Sk;
  const int off = Sk - p.Sq;  // causal: key j visible to query i iff j <= i + off
  int n_tiles = (Sk + Cfg::BN - 1) / Cfg::BN;
  if (p.causal) {
    const int last_key = min(Sk - 1, q0 + Cfg::BM - 1 + off);
    n_tiles = min(n_tiles, last_key / Cfg::BN + 1);
  }

  // split-KV: this block walks key tiles [t_begin, t_end)
  int t_begin = 0, t_end = n_tiles;
  if (p.nsplit > 1) {
    const int per = (n_tiles + p.nsplit - 1) / p.nsplit;
    t_begin = min(n_tiles, split * per);
    t_end = min(n_tiles, t_begin + per);
  }
  const int qrow = b * p.q_row_per_b + q0, qcol = h * p.q_col_per_h;
  const int krow = b * p.k_row_per_b + kvh * p.k_row_per_h, kcol = kvh * p.k_col_per_h;
  const int vrow = b * p.v_row_per_b + kvh * p.v_row_per_h, vcol = kvh * p.v_col_per_h;
  auto kv_bar = [&](int buf_) { return smem_u32(&fbars[1 + buf_]); };
  auto tma_kv = [&](int buf_, int t_) {   // one thread: K and V tile of key step t_ -> buffer buf_
    mbar_arrive_expect_tx(kv_bar(buf_), 2 * Cfg::TILE_BYTES);
    tma_load_2d(sK0 + buf_ * Cfg::TILE_BYTES, &tmap_k, kv_bar(buf_), kcol, krow + t_ * Cfg::BN);
    tma_load_2d(sV0 + buf_ * Cfg::TILE_BYTES, &tmap_v, kv_bar(buf_), vcol, vrow + t_ * Cfg::BN);
  };
  if (tma) {
    if (threadIdx.x == 0) {
      if ((sQ & 1023u) != 0) { printf("crab: flash_attn smem misaligned for TMA\n"); __trap(); }
      mbar_init(q_bar, 1); mbar_init(kv_bar(0), 1); mbar_init(kv_bar(1), 1);
      fence_barrier_init();
      mbar_arrive_expect_tx(q_bar, MT * Cfg::TILE_BYTES);
#pragma unroll
      for (int t = 0; t < MT; ++t) tma_load_2d(sQ + t * Cfg::TILE_BYTES, &tmap_q, q_bar, qcol, qrow + t * 64);
      if (t_begin < t_end) tma_kv(0, t_begin);
    }
    __syncthreads();   // barriers initialised before anybody waits on them
  } else {
#pragma unroll
    for (int t = 0; t < MT; ++t) load_tile<HD>(sQ + t * Cfg::TILE_BYTES, qg, p.q_rs, q0 + t * 64, p.Sq);
    for (int i = 0; i < NST - 1; ++i) {   // NST - 1 key steps in flight before the loop; one commit group per step (possibly empty)
      if (t_begin + i < t_end) {
        load_tile<HD>(sK0 + i * Cfg::TILE_BYTES, kg, p.k_rs, (t_begin + i) * Cfg::BN, Sk);
        load_tile<HD>(sV0 + i * Cfg::TILE_BYTES, vg, p.v_rs, (t_begin + i) * Cfg::BN, Sk);
      }
      cp_async_commit();
    }
  }

  constexpr int DT = HD / 8;  // output n8-tiles per row
  const int wrow = KS ? 0 : warp;          // 16-row block of this warp
  const int kw = KS ? warp * 16 : 0;       // first key of this warp within a tile
  float o_acc[MT][DT][4];
  float m_run[MT][2], l_run[MT][2];
  int r_lo[MT];               // global query row of c0/c1 for m-tile mt; c2/c3 are r_lo + 8
  float gate_lo[MT], gate_hi[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int i = 0; i < DT; ++i) { o_acc[mt][i][0] = o_acc[mt][i][1] = o_acc[mt][i][2] = o_acc[mt][i][3] = 0.f; }
    m_run[mt][0] = m_run[mt][1] = -INFINITY;
    l_run[mt][0] = l_run[mt][1] = 0.f;
    r_lo[mt] = q0 + (wrow * MT + mt) * 16 + (lane >> 2);
    gate_lo[mt] = gate_hi[mt] = 0.f;
    if (p.gate) {
      if (r_lo[mt] < p.Sq) gate_lo[mt] = p.gate[((size_t)b * p.H + h) * p.Sq + r_lo[mt]];
      if (r_lo[mt] + 8 < p.Sq) gate_hi[mt] = p.gate[((size_t)b * p.H + h) * p.Sq + r_lo[mt] + 8];
    }
  }
  const float sl2 = p.scale * 1.4426950408889634f;
  const bool warp_active = q0 + wrow * MT * 16 < p.Sq;

  if (t_begin >= t_end && !tma) cp_async_wait<0>();   // empty split: nothing to consume
  if (tma) mbar_wait(q_bar, 0);
  for (int t = t_begin; t < t_end; ++t) {
    const int buf = (t - t_begin) % NST;
    if (tma) {
      // buffer buf ^ 1 was last read in iteration t - 1, which ended with a block barrier: it may be refilled now
      if (t + 1 < t_end && threadIdx.x == 0) tma_kv(buf ^ 1, t + 1);
      mbar_wait(kv_bar(buf), (uint32_t)((t - t_begin) >> 1) & 1);
    } else {
      // request key step t + NST - 1 into the buffer that iteration t - 1 released (block barrier at its end), then wait until
      // only the NST - 1 youngest groups are pending: step t has landed
      const int tn = t + NST - 1;
      if (tn < t_end) {
        const int bn = (tn - t_begin) % NST;
        load_tile<HD>(sK0 + bn * Cfg::TILE_BYTES, kg, p.k_rs, tn * Cfg::BN, Sk);
        load_tile<HD>(sV0 + bn * Cfg::TILE_BYTES, vg, p.v_rs, tn * Cfg::BN, Sk);
      }
      cp_async_commit();
      if (NST == 3) cp_async_wait<2>(); else cp_async_wait<1>();
      __syncthreads();
    }
    // key split: a row offset of 16 keeps (row & 7), i.e. the swizzle, so the warp's 16 keys are simply a shifted tile base
    const uint32_t sK = sK0 + buf * Cfg::TILE_BYTES + (uint32_t)(kw * HD * 2), sV = sV0 + buf * Cfg::TILE_BYTES + (uint32_t)(kw * HD * 2);

    // a warp whose query rows all lie past Sq (decode over a kv group: Sq = G <= 16 rows, three of the four warps) only helps
    // with the loads: its MMAs / softmax would be work on padding that competes with the one useful warp for the tensor pipe
    if (warp_active) {
    // ---- S = Q K^T  (MT x 16 x 64 per warp) ----
    float s[MT][NT8][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < NT8; ++i) { s[mt][i][0] = s[mt][i][1] = s[mt][i][2] = s[mt][i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      uint32_t a[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        // A fragment (m16 x k16): matrices (rows 0-7,k0-7), (rows 8-15,k0-7), (rows 0-7,k8-15), (rows 8-15,k8-15)
        const int rr = (wrow * MT + mt) * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;  // row within the BM-row Q block
        const int c = ks * 2 + (lane >> 4);
        ldsm_x4(tile_addr<HD>(sQ + (rr >> 6) * Cfg::TILE_BYTES, rr & 63, c), a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
      }
#pragma unroll
      for (int np = 0; np < NPG; ++np) {
        uint32_t b0, b1, b2, b3;
        // matrices: (n0-7,k0-7), (n0-7,k8-15), (n8-15,k0-7), (n8-15,k8-15)
        const int n = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int c = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(tile_addr<HD>(sK, n, c), b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16(s[mt][np * 2], a[mt], b0, b1);
          mma_bf16(s[mt][np * 2 + 1], a[mt], b2, b3);
        }
      }
    }

    // ---- scale, bias, mask, online softmax (log2 domain) ----
    // Interior tiles (no key past Sk, nothing above the causal diagonal, no bias) take the fast path: the raw scores stay
    // in registers and the scale is folded into the exponent's FFMA; only edge / diagonal / biased tiles pay for the
    // per-element index arithmetic.
    const int k0 = t * Cfg::BN;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int q_min = q0 + (wrow * MT + mt) * 16;  // smallest query row of this m-tile (warp-uniform)
      const bool general = (p.table != nullptr) || (k0 + Cfg::BN > Sk) || (p.causal && (k0 + Cfg::BN - 1 > q_min + off));
      float mx[2] = {-INFINITY, -INFINITY};
      if (general) {
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int kj = k0 + kw + nt * 8 + (lane & 3) * 2 + (e & 1);
            const int hi = e >> 1;
            const int qi = r_lo[mt] + hi * 8;
            float v = s[mt][nt][e] * sl2;
            if (p.table != nullptr && qi < p.Sq && kj < Sk)
              v += (hi ? gate_hi[mt] : gate_lo[mt]) * p.table[((size_t)h * p.Sq + qi) * Sk + kj] * 1.4426950408889634f;
            const bool masked = (kj >= Sk) || (p.causal && kj > qi + off);
            v = masked ? -INFINITY : v;
            s[mt][nt][e] = v;
            mx[hi] = fmaxf(mx[hi], v);
          }
        }
      } else {
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) {
          mx[0] = fmaxf(mx[0], fmaxf(s[mt][nt][0], s[mt][nt][1]));
          mx[1] = fmaxf(mx[1], fmaxf(s[mt][nt][2], s[mt][nt][3]));
        }
        mx[0] *= sl2;  // sl2 > 0: max commutes with the scale
        mx[1] *= sl2;
      }
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        mx[hi] = fmaxf(mx[hi], __shfl_xor_sync(0xffffffffu, mx[hi], 1));
        mx[hi] = fmaxf(mx[hi], __shfl_xor_sync(0xffffffffu, mx[hi], 2));
        mx[hi] = fmaxf(mx[hi], m_run[mt][hi]);
      }
      float corr[2], msafe[2];
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        msafe[hi] = (mx[hi] == -INFINITY) ? 0.f : mx[hi];
        corr[hi] = exp2f(m_run[mt][hi] - msafe[hi]);
        m_run[mt][hi] = mx[hi];
        l_run[mt][hi] *= corr[hi];
      }
      float rs[2] = {0.f, 0.f};
      if (general) {
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float pv = exp2f(s[mt][nt][e] - msafe[e >> 1]);
            s[mt][nt][e] = pv;
            rs[e >> 1] += pv;
          }
        }
      } else {
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float pv = exp2f(fmaf(s[mt][nt][e], sl2, -msafe[e >> 1]));
            s[mt][nt][e] = pv;
            rs[e >> 1] += pv;
          }
        }
      }
      l_run[mt][0] += rs[0];
      l_run[mt][1] += rs[1];
#pragma unroll
      for (int i = 0; i < DT; ++i) {
        o_acc[mt][i][0] *= corr[0]; o_acc[mt][i][1] *= corr[0];
        o_acc[mt][i][2] *= corr[1]; o_acc[mt][i][3] *= corr[1];
      }
    }

    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < NKK; ++kk) {  // 16 keys per step
      uint32_t a[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        a[mt][0] = pack_bf16x2(s[mt][2 * kk][0], s[mt][2 * kk][1]);
        a[mt][1] = pack_bf16x2(s[mt][2 * kk][2], s[mt][2 * kk][3]);
        a[mt][2] = pack_bf16x2(s[mt][2 * kk + 1][0], s[mt][2 * kk + 1][1]);
        a[mt][3] = pack_bf16x2(s[mt][2 * kk + 1][2], s[mt][2 * kk + 1][3]);
      }
#pragma unroll
      for (int dp = 0; dp < DT / 2; ++dp) {
        uint32_t b0, b1, b2, b3;
        // transposed loads; matrices: (keys 0-7,d0-7), (keys 8-15,d0-7), (keys 0-7,d8-15), (keys 8-15,d8-15)
        const int key = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = dp * 2 + (lane >> 4);
        ldsm_x4_t(tile_addr<HD>(sV, key, c), b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16(o_acc[mt][dp * 2], a[mt], b0, b1);
          mma_bf16(o_acc[mt][dp * 2 + 1], a[mt], b2, b3);
        }
      }
    }
    }   // warp_active
    __syncthreads();
  }

  // ---- finalise ----
  if (KS) {
    // merge the four warps' partial states (same 16 rows, disjoint keys) through the K / V buffers, idle now:
    // mw[w][16] running max, lw[w][16] sums, ow[w][16][HD] accumulators
    float* mw = reinterpret_cast<float*>(smem + MT * Cfg::TILE_BYTES);
    float* lw = mw + 64;
    float* ow = lw + 64;
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      l_run[0][hi] += __shfl_xor_sync(0xffffffffu, l_run[0][hi], 1);
      l_run[0][hi] += __shfl_xor_sync(0xffffffffu, l_run[0][hi], 2);
      const int row = (lane >> 2) + 8 * hi;
      if ((lane & 3) == 0) { mw[warp * 16 + row] = m_run[0][hi]; lw[warp * 16 + row] = l_run[0][hi]; }
#pragma unroll
      for (int i = 0; i < DT; ++i)
        *reinterpret_cast<float2*>(ow + ((size_t)(warp * 16 + row) * HD + i * 8 + (lane & 3) * 2)) = make_float2(o_acc[0][i][2 * hi], o_acc[0][i][2 * hi + 1]);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 16 * HD; idx += 128) {
      const int row = idx / HD, col = idx % HD;
      if (row >= p.Sq) continue;
      float mm = -INFINITY;
#pragma unroll
      for (int w = 0; w < 4; ++w) mm = fmaxf(mm, mw[w * 16 + row]);
      float ll = 0.f, aa = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {   // fixed warp order: deterministic
        const float mv = mw[w * 16 + row];
        const float c = (mv == -INFINITY) ? 0.f : exp2f(mv - mm);
        ll += lw[w * 16 + row] * c;
        aa += ow[(size_t)(w * 16 + row) * HD + col] * c;
      }
      if (p.nsplit > 1) {
        float* w_ = p.ws + ((((size_t)b * p.H + h) * p.Sq + row) * p.nsplit + split) * (HD + 2);
        w_[col] = aa;
        if (col == 0) { w_[HD] = mm; w_[HD + 1] = ll; }
      } else {
        p.o[b * p.o_bs + h * p.o_hs + (long long)row * p.o_rs + col] = __float2bfloat16_rn(ll > 0.f ? aa / ll : 0.f);
      }
    }
    return;
  }
  if (p.nsplit > 1) {
    // partial state of this key range -> workspace (one row per query row, i.e. per q head of the kv group in GQA decode)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        l_run[mt][hi] += __shfl_xor_sync(0xffffffffu, l_run[mt][hi], 1);
        l_run[mt][hi] += __shfl_xor_sync(0xffffffffu, l_run[mt][hi], 2);
      }
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        const int row = r_lo[mt] + 8 * hi;
        if (row < p.Sq) {
          float* w = p.ws + ((((size_t)b * p.H + h) * p.Sq + row) * p.nsplit + split) * (HD + 2);
#pragma unroll
          for (int i = 0; i < DT; ++i) {
            const int col = i * 8 + (lane & 3) * 2;
            *reinterpret_cast<float2*>(w + col) = make_float2(o_acc[mt][i][2 * hi], o_acc[mt][i][2 * hi + 1]);
          }
          if ((lane & 3) == 0) { w[HD] = m_run[mt][hi]; w[HD + 1] = l_run[mt][hi]; }
        }
      }
    }
    return;
  }
  __nv_bfloat16* og = p.o + b * p.o_bs + h * p.o_hs;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      l_run[mt][hi] += __shfl_xor_sync(0xffffffffu, l_run[mt][hi], 1);
      l_run[mt][hi] += __shfl_xor_sync(0xffffffffu, l_run[mt][hi], 2);
    }
    const float inv0 = l_run[mt][0] > 0.f ? 1.f / l_run[mt][0] : 0.f;
    const float inv1 = l_run[mt][1] > 0.f ? 1.f / l_run[mt][1] : 0.f;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      const int col = i * 8 + (lane & 3) * 2;
      if (r_lo[mt] < p.Sq)
        *reinterpret_cast<uint32_t*>(og + (long long)r_lo[mt] * p.o_rs + col) = pack_bf16x2(o_acc[mt][i][0] * inv0, o_acc[mt][i][1] * inv0);
      if (r_lo[mt] + 8 < p.Sq)
        *reinterpret_cast<uint32_t*>(og + (long long)(r_lo[mt] + 8) * p.o_rs + col) = pack_bf16x2(o_acc[mt][i][2] * inv1, o_acc[mt][i][3] * inv1);
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------
// decode attention
// ----------------------------------------------------------------------------------------------------------------
struct DecodeParams {
  const __nv_bfloat16* q; int ldq;            // [B, ldq]; head h at column h*HD
  const __nv_bfloat16* kc; const __nv_bfloat16* vc;  // [B, KVH, ctx_max, HD]
  __nv_bfloat16* o; int ldo;                  // [B, ldo]
  float* ws;                                  // [B*H, nsplit, HD + 2] partials (nsplit > 1)
  int B, H, KVH, ctx_max, nsplit;
  const int* len_dev; int len_host;           // number of valid keys
  float scale;
  int late_trigger;                           // PDL: release the dependent kernel after the streaming loop, not at the top
  // ---- FUSE variant: q points at the RAW qkv row [q (H*HD) | k (KVH*HD) | v (KVH*HD)] straight from the qkv GEMM ----
  const float* cos_sin;                       // [max_pos, HD] = [cos(half) | sin(half)] per position
  const int* past_dev;                        // position of the new token; number of valid keys = past + 1
  __nv_bfloat16* kc_w; __nv_bfloat16* vc_w;   // the same caches, writable: the new K (rotated) / V row is appended here
  const __nv_bfloat16* ra; int ldra;          // optional o_proj hyper-LoRA pre-pass: [R (3); A (8)] rows over H*HD columns
  __nv_bfloat16* z; int ldz;                  // [B, >= 24] z columns (the K-extension of the o_proj GEMM)
  float lora_scale;
  float* lora_ws;                             // [B, KVH, 11] per-head-group partial dots
  int* lora_cnt;                              // [B] arrival counters, zero on entry and left zero
  unsigned long long* trace;                  // diagnostics: [ctas][16] stamps or nullptr
  int csplit;                                 // > 1: split KV over the `csplit` blocks of a CLUSTER, merged in rank 0 through DSMEM (no combine launch)
};

// o_proj hyper-LoRA pre-pass, block tail shared by the fused decode attention (one block per (b, kv head)) and the split-KV
// combine kernel (one block per (b, head)): t11 = this thread's share of the 11 router / A dots of its block's output elements.
// Block partial -> workspace slot `slot` of batch row b; the last of the row's `nblocks` blocks (arrival counter) sums the slots
// in fixed order, applies the fp32 router softmax and writes the 24 z columns (peft_hyper/tuners/lora.py:344-350).
__device__ __forceinline__ void lora_prepass_tail(float (&t11)[11], int b, int slot, int nblocks, float* __restrict__ lora_ws,
                                                  int* __restrict__ lora_cnt, __nv_bfloat16* __restrict__ z, int ldz, float lora_scale,
                                                  float (*sh_red)[11], float* sh_tot, int* sh_ticket) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int j = 0; j < 11; ++j) {
    const float v = warp_sum(t11[j]);
    if (lane == 0) sh_red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 11) {
    float v = 0.f;
    for (int w = 0; w < nwarps; ++w) v += sh_red[w][threadIdx.x];
    lora_ws[((size_t)b * nblocks + slot) * 11 + threadIdx.x] = v;
  }
  // one acq_rel ticket between two block barriers: its release covers the 11 partial stores above (cumulative over the barrier),
  // its acquire the last block's loads below — a GPU-scope fence per thread on either side costs ~2 us each on the step's critical
  // path (this tail sits between the attention's last key and o_proj's first operand)
  __syncthreads();
  if (threadIdx.x == 0) {
    int tk;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(tk) : "l"(lora_cnt + b) : "memory");
    *sh_ticket = tk;
  }
  __syncthreads();
  if (*sh_ticket == nblocks - 1) {
    if (threadIdx.x < 11) {
      const volatile float* wsp = lora_ws + (size_t)b * nblocks * 11 + threadIdx.x;
      float t = 0.f;
      for (int h2 = 0; h2 < nblocks; ++h2) t += wsp[(size_t)h2 * 11];
      sh_tot[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x < 24) {
      const int i = threadIdx.x / 8, j = threadIdx.x % 8;
      const float l0 = sh_tot[0], l1 = sh_tot[1], l2 = sh_tot[2];
      const float mx = fmaxf(l0, fmaxf(l1, l2));
      const float e0 = __expf(l0 - mx), e1 = __expf(l1 - mx), e2 = __expf(l2 - mx);
      const float ri = (i == 0 ? e0 : (i == 1 ? e1 : e2)) / (e0 + e1 + e2);
      z[(size_t)b * ldz + threadIdx.x] = __float2bfloat16_rn(lora_scale * ri * sh_tot[3 + j]);
    }
    if (threadIdx.x == 0) lora_cnt[b] = 0;  // ready for the next launch (next layer / next step)
  }
}

// FUSE = the decode step's RoPE + KV-cache append (+ the o_proj hyper-LoRA pre-pass) folded into the attention kernel:
// the block rotates its own q heads and the new k row in registers (models/modeling_llama.py:204-236), appends k / v to
// the cache, treats the new key as one more key of the stream, and — when `ra` is given — finishes with the 11 router /
// A dot products of its heads' output; the last block of a batch row (arrival counter) sums them over the heads in
// fixed order, applies the fp32 router softmax and writes the 24 z columns (peft_hyper/tuners/lora.py:344-350).
// Two launches per layer (rope_kv_kernel, row_loraz_kernel) disappear from the decode chain.
template <int HD, int G, bool FUSE>
__global__ void __launch_bounds__(128) attn_decode_kernel(const DecodeParams p) {
  constexpr int LPK = HD / 8;        // lanes per key (16 for hd=128, 8 for hd=64)
  constexpr int KPW = 32 / LPK;      // keys per warp-load
  constexpr int NSUB = 4 * KPW;      // independent softmax states per block
  __shared__ float sh_m[NSUB][G], sh_l[NSUB][G];
  __shared__ float sh_acc[NSUB][G][HD];
  __shared__ float sh_red[4][11];
  __shared__ float sh_tot[11];
  __shared__ int sh_ticket;
  __shared__ uint4 sh_new[2][LPK];
  constexpr int PEERS = (G == 1 && FUSE) ? 7 : 0;              // cluster split is offered for MHA only (host check)
  __shared__ float sh_peer[PEERS * (HD + 2) + 2];               // rank 0: the other ranks' states (acc[HD], max, sum)
  __shared__ __align__(8) uint64_t sh_peer_bar;
  if (p.csplit > 1) {
    if (threadIdx.x == 0) {
      mbar_init(smem_u32(&sh_peer_bar), 1);
      fence_barrier_init();
      mbar_arrive_expect_tx(smem_u32(&sh_peer_bar), (uint32_t)((p.csplit - 1) * (HD + 2) * 4));
    }
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");   // waited for by ranks > 0 before their first remote store
  }
  // With an early trigger the next kernels of a PDL chain (ultimately the streaming GEMM, 100 KB smem / 32 K registers per
  // CTA) become resident while this grid still streams the cache and squat on its SM slots; a late trigger releases them
  // only when every CTA is past its loop, which still hides their launch latency behind the combine/tail.
  if (threadIdx.x == 0) trace_stamp(p.trace, 0);
  if (!p.late_trigger) pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) trace_stamp(p.trace, 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPK, li = lane % LPK;
  // cluster split (small batches: one block per (b, kv head) would leave most SMs idle): the `csplit` blocks of a cluster share
  // one (b, kv head) and take a key range each; ranks > 0 hand their softmax state to rank 0 through distributed shared memory,
  // which merges in rank order and finishes (output, o_proj pre-pass) — one launch instead of split kernel + combine kernel
  // (7.5 + 7.5 us per layer at bs 1, where the whole cache read is 0.3 MB).
  uint32_t crank = 0;
  if (p.csplit > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int blk = p.csplit > 1 ? (int)blockIdx.x / p.csplit : (int)blockIdx.x;
  const int b = blk / p.KVH, kvh = blk % p.KVH;
  const int split = p.csplit > 1 ? (int)crank : (int)blockIdx.y;
  const int nsp = p.csplit > 1 ? p.csplit : p.nsplit;
  const int past = FUSE ? *p.past_dev : 0;
  const int len = FUSE ? past + 1 : (p.len_dev ? *p.len_dev : p.len_host);
  const int per = (len + nsp - 1) / nsp;
  const int k_begin = split * per, k_end = min(len, k_begin + per);
  const float sl2 = p.scale * 1.4426950408889634f;

  // FUSE: rotary factors of this lane's 8 dims at position `past` (lanes of the low half pair with lane ^ LPK/2)
  constexpr int HL = LPK / 2;
  const bool is_hi = li >= HL;
  float rc[8], rs[8];
  if (FUSE) {
    const float* cp = p.cos_sin + (size_t)past * HD + (li % HL) * 8;
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(cp)), c1 = __ldg(reinterpret_cast<const float4*>(cp) + 1);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(cp + HD / 2)), s1 = __ldg(reinterpret_cast<const float4*>(cp + HD / 2) + 1);
    rc[0] = c0.x; rc[1] = c0.y; rc[2] = c0.z; rc[3] = c0.w; rc[4] = c1.x; rc[5] = c1.y; rc[6] = c1.z; rc[7] = c1.w;
    rs[0] = s0.x; rs[1] = s0.y; rs[2] = s0.z; rs[3] = s0.w; rs[4] = s1.x; rs[5] = s1.y; rs[6] = s1.z; rs[7] = s1.w;
    if (!is_hi) {
#pragma unroll
      for (int j = 0; j < 8; ++j) rs[j] = -rs[j];  // low half: x c - partner s;  high half: x c + partner s
    }
  }
  auto rope8 = [&](const uint4& raw, float* out) {  // rotate, then round to bf16 exactly where the unfused path stored bf16
    float t[8];
    t[0] = bf16lo(raw.x); t[1] = bf16hi(raw.x); t[2] = bf16lo(raw.y); t[3] = bf16hi(raw.y);
    t[4] = bf16lo(raw.z); t[5] = bf16hi(raw.z); t[6] = bf16lo(raw.w); t[7] = bf16hi(raw.w);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float other = __shfl_xor_sync(0xffffffffu, t[j], HL);
      out[j] = __bfloat162float(__float2bfloat16_rn(t[j] * rc[j] + other * rs[j]));
    }
  };

  float qf[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const __nv_bfloat16* qp = p.q + (size_t)b * p.ldq + (size_t)(kvh * G + g) * HD + li * 8;
    float t[8];
    if (FUSE) {
      rope8(ld_dep_u4(qp), t);
    } else {
      const uint4 qq = *reinterpret_cast<const uint4*>(qp);
      t[0] = bf16lo(qq.x); t[1] = bf16hi(qq.x); t[2] = bf16lo(qq.y); t[3] = bf16hi(qq.y);
      t[4] = bf16lo(qq.z); t[5] = bf16hi(qq.z); t[6] = bf16lo(qq.w); t[7] = bf16hi(qq.w);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) qf[g][j] = t[j] * sl2;
  }
  // FUSE: warp 0 rotates the new key row of this kv head, parks it (and the new value row) in shared memory for the
  // tail below, and — in the block whose key range holds position `past` — appends both to the cache.  Nobody reads that
  // cache row in this launch: the stream below stops at `past` and the tail folds the new key in from shared memory, so
  // the streaming loop is instruction-for-instruction the unfused one.
  const bool owns_new = FUSE && past >= k_begin && past < k_end;
  if (FUSE && warp == 0 && owns_new) {
    const __nv_bfloat16* kp = p.q + (size_t)b * p.ldq + (size_t)(p.H + kvh) * HD + li * 8;
    float kr[8];
    rope8(ld_dep_u4(kp), kr);
    uint4 knew;
    knew.x = pack_bf16x2(kr[0], kr[1]); knew.y = pack_bf16x2(kr[2], kr[3]);
    knew.z = pack_bf16x2(kr[4], kr[5]); knew.w = pack_bf16x2(kr[6], kr[7]);
    const uint4 vnew = ld_dep_u4(kp + (size_t)p.KVH * HD);
    if (sub == 0) {
      sh_new[0][li] = knew;
      sh_new[1][li] = vnew;
      if (past < p.ctx_max) {
        const size_t row = (((size_t)b * p.KVH + kvh) * p.ctx_max + past) * HD + li * 8;
        *reinterpret_cast<uint4*>(p.kc_w + row) = knew;
        *reinterpret_cast<uint4*>(p.vc_w + row) = vnew;
      }
    }
    __syncwarp();
  }
  const int k_stream_end = FUSE ? min(k_end, past) : k_end;  // keys read from the cache
  float m[G], l[G], acc[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[g][j] = 0.f;
  }
  const __nv_bfloat16* kbase = p.kc + ((size_t)b * p.KVH + kvh) * p.ctx_max * HD;
  const __nv_bfloat16* vbase = p.vc + ((size_t)b * p.KVH + kvh) * p.ctx_max * HD;
  constexpr int U = 4;
  // this (warp, sub) walks keys k_begin + (warp*KPW + sub) + i * NSUB.  The loop bound depends on the warp only, so
  // every lane of a warp runs the same number of iterations (the shuffles below need the full warp).
  for (int kb = k_begin + warp * KPW; kb < k_stream_end; kb += NSUB * U) {
    const int k0 = kb + sub;
    uint4 kq[U], vq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int key = k0 + u * NSUB;
      if (key < k_stream_end) {
        kq[u] = __ldg(reinterpret_cast<const uint4*>(kbase + (size_t)key * HD) + li);
        vq[u] = __ldg(reinterpret_cast<const uint4*>(vbase + (size_t)key * HD) + li);
      } else {
        kq[u] = make_uint4(0, 0, 0, 0);
        vq[u] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int key = k0 + u * NSUB;
      const bool valid = key < k_stream_end;  // uniform across the LPK lanes of this key
      float kf[8], vf[8];
      kf[0] = bf16lo(kq[u].x); kf[1] = bf16hi(kq[u].x); kf[2] = bf16lo(kq[u].y); kf[3] = bf16hi(kq[u].y);
      kf[4] = bf16lo(kq[u].z); kf[5] = bf16hi(kq[u].z); kf[6] = bf16lo(kq[u].w); kf[7] = bf16hi(kq[u].w);
      vf[0] = bf16lo(vq[u].x); vf[1] = bf16hi(vq[u].x); vf[2] = bf16lo(vq[u].y); vf[3] = bf16hi(vq[u].y);
      vf[4] = bf16lo(vq[u].z); vf[5] = bf16hi(vq[u].z); vf[6] = bf16lo(vq[u].w); vf[7] = bf16hi(vq[u].w);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float sc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sc += qf[g][j] * kf[j];
#pragma unroll
        for (int o = LPK / 2; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
        if (valid) {
          const float mn = fmaxf(m[g], sc);
          const float c = exp2f(m[g] - mn);
          const float pe = exp2f(sc - mn);
          m[g] = mn;
          l[g] = l[g] * c + pe;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[g][j] = acc[g][j] * c + pe * vf[j];
        }
      }
    }
  }
  if (FUSE && warp == 0 && owns_new) {
    // the new key, from shared memory, joins the (warp 0, sub 0) softmax state
    const uint4 kq1 = sh_new[0][li], vq1 = sh_new[1][li];
    float kf[8], vf[8];
    kf[0] = bf16lo(kq1.x); kf[1] = bf16hi(kq1.x); kf[2] = bf16lo(kq1.y); kf[3] = bf16hi(kq1.y);
    kf[4] = bf16lo(kq1.z); kf[5] = bf16hi(kq1.z); kf[6] = bf16lo(kq1.w); kf[7] = bf16hi(kq1.w);
    vf[0] = bf16lo(vq1.x); vf[1] = bf16hi(vq1.x); vf[2] = bf16lo(vq1.y); vf[3] = bf16hi(vq1.y);
    vf[4] = bf16lo(vq1.z); vf[5] = bf16hi(vq1.z); vf[6] = bf16lo(vq1.w); vf[7] = bf16hi(vq1.w);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float sc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) sc += qf[g][j] * kf[j];
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
      if (sub == 0) {
        const float mn = fmaxf(m[g], sc);
        const float c = exp2f(m[g] - mn);
        const float pe = exp2f(sc - mn);
        m[g] = mn;
        l[g] = l[g] * c + pe;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[g][j] = acc[g][j] * c + pe * vf[j];
      }
    }
  }
  if (threadIdx.x == 0) trace_stamp(p.trace, 4);
  if (p.late_trigger) pdl_trigger();
  // combine the NSUB partial states
  const int sidx = warp * KPW + sub;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (li == 0) { sh_m[sidx][g] = m[g]; sh_l[sidx][g] = l[g]; }
#pragma unroll
    for (int j = 0; j < 8; ++j) sh_acc[sidx][g][li * 8 + j] = acc[g][j];
  }
  __syncthreads();
  const bool lora = FUSE && p.ra != nullptr && p.nsplit == 1;   // split KV: the combine kernel does the pre-pass
  float t11[11];
#pragma unroll
  for (int j = 0; j < 11; ++j) t11[j] = 0.f;
  if (p.csplit > 1 && crank != 0) asm volatile("barrier.cluster.wait.acquire;" ::: "memory");   // rank 0's barrier is initialised and armed
  for (int idx = threadIdx.x; idx < G * HD; idx += 128) {
    const int g = idx / HD, d = idx % HD;
    float mm = -INFINITY;
#pragma unroll
    for (int s2 = 0; s2 < NSUB; ++s2) mm = fmaxf(mm, sh_m[s2][g]);
    float ll = 0.f, aa = 0.f;
    if (mm != -INFINITY) {
#pragma unroll
      for (int s2 = 0; s2 < NSUB; ++s2) {
        const float c = exp2f(sh_m[s2][g] - mm);  // exp2(-inf) = 0 for empty states
        ll += sh_l[s2][g] * c;
        aa += sh_acc[s2][g][d] * c;
      }
    }
    const int hq = kvh * G + g;
    if (PEERS > 0 && p.csplit > 1) {
      if (crank != 0) {
        // state of this rank's key range -> rank 0's shared memory, bytes counted on its barrier
        uint32_t rbase, rbar;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbase) : "r"(smem_u32(sh_peer)), "r"(0u));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(&sh_peer_bar)), "r"(0u));
        rbase += (uint32_t)((crank - 1) * (HD + 2) * 4);
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(rbase + (uint32_t)(d * 4)), "f"(aa), "r"(rbar) : "memory");
        if (d == 0) {
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(rbase + (uint32_t)(HD * 4)), "f"(mm), "r"(rbar) : "memory");
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(rbase + (uint32_t)((HD + 1) * 4)), "f"(ll), "r"(rbar) : "memory");
        }
        continue;
      }
      mbar_wait(smem_u32(&sh_peer_bar), 0);
      for (int r = 1; r < p.csplit; ++r) {   // rank order: deterministic
        const float* pp = sh_peer + (r - 1) * (HD + 2);
        const float pm = pp[HD], pl = pp[HD + 1], pa = pp[d];
        const float mt = fmaxf(mm, pm);
        const float c0 = (mm == -INFINITY) ? 0.f : exp2f(mm - mt), c1 = (pm == -INFINITY) ? 0.f : exp2f(pm - mt);
        ll = ll * c0 + pl * c1;
        aa = aa * c0 + pa * c1;
        mm = mt;
      }
    }
    if (p.nsplit == 1) {
      const __nv_bfloat16 ob = __float2bfloat16_rn(ll > 0.f ? aa / ll : 0.f);
      p.o[(size_t)b * p.ldo + (size_t)hq * HD + d] = ob;
      if (lora) {
        const float of = __bfloat162float(ob);
        const __nv_bfloat16* rp = p.ra + (size_t)hq * HD + d;
#pragma unroll
        for (int j = 0; j < 11; ++j) t11[j] += of * __bfloat162float(rp[(size_t)j * p.ldra]);
      }
    } else {
      float* w = p.ws + (((size_t)b * p.H + hq) * p.nsplit + split) * (HD + 2);
      w[d] = aa;
      if (d == 0) { w[HD] = mm; w[HD + 1] = ll; }
    }
  }
  if (lora && crank == 0) lora_prepass_tail(t11, b, kvh, p.KVH, p.lora_ws, p.lora_cnt, p.z, p.ldz, p.lora_scale, sh_red, sh_tot, &sh_ticket);
  if (threadIdx.x == 0) trace_stamp(p.trace, 7);
}

template <int HD>
__global__ void attn_decode_combine_kernel(const float* __restrict__ ws, __nv_bfloat16* __restrict__ o, int ldo, int H, int nsplit,
                                           const __nv_bfloat16* __restrict__ ra, int ldra, __nv_bfloat16* __restrict__ z, int ldz,
                                           float lora_scale, float* __restrict__ lora_ws, int* __restrict__ lora_cnt) {
  __shared__ float sh_red[4][11];
  __shared__ float sh_tot[11];
  __shared__ int sh_ticket;
  pdl_trigger();
  pdl_wait();
  const int bh = blockIdx.x;
  const int b = bh / H, h = bh % H;
  const float* w = ws + (size_t)bh * nsplit * (HD + 2);
  float mm = -INFINITY;
  for (int s = 0; s < nsplit; ++s) mm = fmaxf(mm, w[s * (HD + 2) + HD]);
  float t11[11];
#pragma unroll
  for (int j = 0; j < 11; ++j) t11[j] = 0.f;
  for (int d = threadIdx.x; d < HD; d += blockDim.x) {
    float ll = 0.f, aa = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float ms = w[s * (HD + 2) + HD];
      const float c = (ms == -INFINITY) ? 0.f : exp2f(ms - mm);
      ll += w[s * (HD + 2) + HD + 1] * c;
      aa += w[s * (HD + 2) + d] * c;
    }
    const __nv_bfloat16 ob = __float2bfloat16_rn(ll > 0.f ? aa / ll : 0.f);
    o[(size_t)b * ldo + (size_t)h * HD + d] = ob;
    if (ra) {   // o_proj hyper-LoRA pre-pass on the rounded output, as the unsplit fused kernel does
      const float of = __bfloat162float(ob);
      const __nv_bfloat16* rp = ra + (size_t)h * HD + d;
#pragma unroll
      for (int j = 0; j < 11; ++j) t11[j] += of * __bfloat162float(rp[(size_t)j * ldra]);
    }
  }
  if (ra) lora_prepass_tail(t11, b, h, H, lora_ws, lora_cnt, z, ldz, lora_scale, sh_red, sh_tot, &sh_ticket);
}

}  // namespace crab

using namespace crab;

// Can (base, batch stride, row stride, head stride) of a head_dim-64 operand be walked as a 2-D [rows, row stride] tensor by TMA?
static bool flash64_tma_view(const void* base, long long bs, long long rs, long long hs, int heads, int B, int* row_per_b, int* row_per_h,
                             int* col_per_h, long long* rows_total) {
  if (rs < 64 || rs % 8 != 0 || ((uintptr_t)base % 16) != 0 || bs % rs != 0) return false;
  *row_per_b = (int)(bs / rs);
  if (hs % rs == 0) { *row_per_h = (int)(hs / rs); *col_per_h = 0; }
  else if ((long long)heads * hs <= rs && hs >= 64 && hs % 8 == 0) { *row_per_h = 0; *col_per_h = (int)hs; }
  else return false;
  *rows_total = (long long)B * (bs / rs);
  return true;
}

extern "C" int crab_flash_attn(const crab_attn_args* a, void* stream) {
  CRAB_REQUIRE(a && a->q && a->k && a->v && a->o, "crab_flash_attn: null pointer");
  CRAB_REQUIRE(a->head_dim == 64 || a->head_dim == 128, "crab_flash_attn: head_dim must be 64 or 128 (got %d)", a->head_dim);
  CRAB_REQUIRE(a->B > 0 && a->H > 0 && a->KVH > 0 && a->H % a->KVH == 0 && a->Sq > 0 && a->Sk > 0, "crab_flash_attn: bad shape");
  CRAB_REQUIRE((a->gate == nullptr) == (a->bias_table == nullptr), "crab_flash_attn: gate and bias_table go together");
  CRAB_REQUIRE(!a->causal || a->Sk >= a->Sq, "crab_flash_attn: causal needs Sk >= Sq");
  CRAB_REQUIRE(!a->sk_dev || (!a->causal && !a->bias_table), "crab_flash_attn: sk_dev is for bias-free non-causal problems");
  const long long strides[] = {a->q_bs, a->q_rs, a->q_hs, a->k_bs, a->k_rs, a->k_hs, a->v_bs, a->v_rs, a->v_hs, a->o_rs, a->o_hs, a->o_bs};
  for (long long s : strides) CRAB_REQUIRE(s % 8 == 0, "crab_flash_attn: strides must be multiples of 8 elements");
  CRAB_REQUIRE(((uintptr_t)a->q % 16 == 0) && ((uintptr_t)a->k % 16 == 0) && ((uintptr_t)a->v % 16 == 0) && ((uintptr_t)a->o % 4 == 0),
               "crab_flash_attn: pointer alignment");
  FlashParams p;
  p.q = (const __nv_bfloat16*)a->q; p.k = (const __nv_bfloat16*)a->k; p.v = (const __nv_bfloat16*)a->v; p.o = (__nv_bfloat16*)a->o;
  p.q_bs = a->q_bs; p.q_rs = a->q_rs; p.q_hs = a->q_hs;
  p.k_bs = a->k_bs; p.k_rs = a->k_rs; p.k_hs = a->k_hs;
  p.v_bs = a->v_bs; p.v_rs = a->v_rs; p.v_hs = a->v_hs;
  p.o_bs = a->o_bs; p.o_rs = a->o_rs; p.o_hs = a->o_hs;
  p.B = a->B; p.H = a->H; p.KVH = a->KVH; p.Sq = a->Sq; p.Sk = a->Sk;
  p.scale = a->scale; p.causal = a->causal; p.gate = a->gate; p.table = a->bias_table; p.sk_dev = a->sk_dev;
  p.sk_add = 0; p.nsplit = 1; p.ws = nullptr; p.stages = 2;
  p.use_tma = 0;
  p.q_row_per_b = p.q_col_per_h = p.k_row_per_b = p.k_row_per_h = p.k_col_per_h = p.v_row_per_b = p.v_row_per_h = p.v_col_per_h = 0;
  CUtensorMap tq, tk, tv;
  memset(&tq, 0, sizeof(tq)); memset(&tk, 0, sizeof(tk)); memset(&tv, 0, sizeof(tv));
  if (a->head_dim == 64 && !a->sk_dev) {
    static int tma_enabled = -1;
    if (tma_enabled < 0) { const char* e = getenv("CRAB_FLASH_TMA"); tma_enabled = (e && e[0] == '0') ? 0 : 1; }
    long long qr, kr, vr;
    int q_row_per_h = 0;
    if (tma_enabled &&
        flash64_tma_view(a->q, a->q_bs, a->q_rs, a->q_hs, a->H, a->B, &p.q_row_per_b, &q_row_per_h, &p.q_col_per_h, &qr) && q_row_per_h == 0 &&
        flash64_tma_view(a->k, a->k_bs, a->k_rs, a->k_hs, a->KVH, a->B, &p.k_row_per_b, &p.k_row_per_h, &p.k_col_per_h, &kr) &&
        flash64_tma_view(a->v, a->v_bs, a->v_rs, a->v_hs, a->KVH, a->B, &p.v_row_per_b, &p.v_row_per_h, &p.v_col_per_h, &vr)) {
      // extents = the last row the call is entitled to read + 1 (the API carries no buffer sizes): whatever a 64-row box covers
      // beyond that is zero-filled by TMA instead of being read
      qr = (long long)(a->B - 1) * p.q_row_per_b + a->Sq;
      kr = (long long)(a->B - 1) * p.k_row_per_b + (long long)(a->KVH - 1) * p.k_row_per_h + a->Sk;
      vr = (long long)(a->B - 1) * p.v_row_per_b + (long long)(a->KVH - 1) * p.v_row_per_h + a->Sk;
      int rc = encode_tmap_bf16_2d(&tq, a->q, (uint64_t)qr, (uint64_t)a->q_rs, (uint64_t)a->q_rs, 64, 64);
      if (rc == 0) rc = encode_tmap_bf16_2d(&tk, a->k, (uint64_t)kr, (uint64_t)a->k_rs, (uint64_t)a->k_rs, 64, 64);
      if (rc == 0) rc = encode_tmap_bf16_2d(&tv, a->v, (uint64_t)vr, (uint64_t)a->v_rs, (uint64_t)a->v_rs, 64, 64);
      if (rc != 0) return rc;
      p.use_tma = 1;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  {  // head_dim 128 without a bias table and with TMA-describable strides: the tcgen05 / TMEM kernel (flash_tcgen05.cu)
    const int rc = a->sk_dev ? -1 : flash_attn_tcgen05_try(a, st);
    if (rc == 0) return CRAB_OK;
    if (rc < -1) return rc;
  }
  // 32 query rows per warp (MT = 2) when there are enough rows to fill 128-row blocks; short sequences (Q-Former's 32
  // queries, BEATs' 48 tokens) keep 64-row blocks.
  const bool big = a->Sq >= 512;  // N=257 (CLIP) wastes less with 64-row blocks (5 x 64 vs 3 x 128 row slots)
#define CRAB_FLASH_LAUNCH(HD_, MT_)                                                                                        \
  {                                                                                                                        \
    static DeviceOnce set;                                                                                                 \
    if (first_on_device(set)) {                                                                                            \
      CRAB_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<HD_, MT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                           FlashCfg<HD_, MT_>::SMEM3));  /* the larger of the two launch configurations */ \
    }                                                                                                                      \
    dim3 grid((a->Sq + FlashCfg<HD_, MT_>::BM - 1) / FlashCfg<HD_, MT_>::BM, a->H, a->B);                                  \
    flash_attn_kernel<HD_, MT_><<<grid, 128, FlashCfg<HD_, MT_>::SMEM, st>>>(tq, tk, tv, p);                               \
  }
  if (a->head_dim == 64) {
    if (big) CRAB_FLASH_LAUNCH(64, 2) else CRAB_FLASH_LAUNCH(64, 1)
  } else {
    if (big) CRAB_FLASH_LAUNCH(128, 2) else CRAB_FLASH_LAUNCH(128, 1)
  }
#undef CRAB_FLASH_LAUNCH
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_attn_decode_workspace_bytes(int B, int H, int head_dim, int nsplit, int64_t* bytes) {
  CRAB_REQUIRE(bytes != nullptr, "crab_attn_decode_workspace_bytes: null");
  *bytes = (int64_t)B * H * nsplit * (head_dim + 2) * 4;
  return CRAB_OK;
}

extern "C" int crab_attn_decode(const void* q, int ldq, const void* k_cache, const void* v_cache, void* o, int ldo,
                                float* workspace, int B, int H, int KVH, int head_dim, int ctx_max, int nsplit,
                                const int* len_dev, int len_host, float scale, void* stream) {
  CRAB_REQUIRE(q && k_cache && v_cache && o, "crab_attn_decode: null pointer");
  CRAB_REQUIRE(head_dim == 128 || head_dim == 64, "crab_attn_decode: head_dim must be 64 or 128");
  CRAB_REQUIRE(B > 0 && H > 0 && KVH > 0 && H % KVH == 0, "crab_attn_decode: bad heads");
  CRAB_REQUIRE(nsplit >= 1 && (nsplit == 1 || workspace != nullptr), "crab_attn_decode: nsplit>1 needs a workspace");
  CRAB_REQUIRE(ldq % 8 == 0 && ((uintptr_t)q % 16 == 0), "crab_attn_decode: q alignment");
  const int G = H / KVH;
  DecodeParams p;
  p.q = (const __nv_bfloat16*)q; p.ldq = ldq; p.kc = (const __nv_bfloat16*)k_cache; p.vc = (const __nv_bfloat16*)v_cache;
  p.o = (__nv_bfloat16*)o; p.ldo = ldo; p.ws = workspace; p.B = B; p.H = H; p.KVH = KVH; p.ctx_max = ctx_max;
  p.nsplit = nsplit; p.len_dev = len_dev; p.len_host = len_host; p.scale = scale;
  p.late_trigger = (pdl_mask() & PDL_ATTN_LATE) ? 1 : 0;
  p.cos_sin = nullptr; p.past_dev = nullptr; p.kc_w = nullptr; p.vc_w = nullptr; p.ra = nullptr; p.ldra = 0; p.z = nullptr;
  p.ldz = 0; p.lora_scale = 0.f; p.lora_ws = nullptr; p.lora_cnt = nullptr;
  p.trace = nullptr; p.csplit = 0;
  dim3 grid(B * KVH, nsplit);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
#define CRAB_DECODE_CASE(HD_, G_) \
  if (head_dim == HD_ && G == G_) { e = launch_pdl(PDL_ATTN, attn_decode_kernel<HD_, G_, false>, grid, dim3(128), 0, st, p); } else
  CRAB_DECODE_CASE(128, 1) CRAB_DECODE_CASE(128, 2) CRAB_DECODE_CASE(128, 4) CRAB_DECODE_CASE(128, 7)
  CRAB_DECODE_CASE(128, 8) CRAB_DECODE_CASE(64, 1)
  { return set_error(CRAB_ERR_INVALID, "crab_attn_decode: unsupported head_dim=%d group=%d", head_dim, G); }
#undef CRAB_DECODE_CASE
  CRAB_CHECK_CUDA(e);
  if (nsplit > 1) {
    if (head_dim == 128) e = launch_pdl(PDL_ATTN, attn_decode_combine_kernel<128>, dim3(B * H), dim3(128), 0, st, (const float*)workspace, p.o, ldo, H, nsplit,
                                        (const __nv_bfloat16*)nullptr, 0, (__nv_bfloat16*)nullptr, 0, 0.f, (float*)nullptr, (int*)nullptr);
    else e = launch_pdl(PDL_ATTN, attn_decode_combine_kernel<64>, dim3(B * H), dim3(64), 0, st, (const float*)workspace, p.o, ldo, H, nsplit,
                        (const __nv_bfloat16*)nullptr, 0, (__nv_bfloat16*)nullptr, 0, 0.f, (float*)nullptr, (int*)nullptr);
    CRAB_CHECK_CUDA(e);
  }
  return CRAB_OK;
}

extern "C" int crab_attn_decode_fused(const crab_decode_fused_args* a, void* stream) {
  CRAB_REQUIRE(a && a->qkv && a->cos_sin && a->k_cache && a->v_cache && a->o && a->past_dev, "crab_attn_decode_fused: null pointer");
  CRAB_REQUIRE(a->head_dim == 128 || a->head_dim == 64, "crab_attn_decode_fused: head_dim must be 64 or 128");
  CRAB_REQUIRE(a->B > 0 && a->H > 0 && a->KVH > 0 && a->H % a->KVH == 0, "crab_attn_decode_fused: bad heads");
  CRAB_REQUIRE(a->nsplit >= 1 && (a->nsplit == 1 || a->workspace != nullptr), "crab_attn_decode_fused: nsplit>1 needs a workspace");
  CRAB_REQUIRE(a->ldq % 8 == 0 && ((uintptr_t)a->qkv % 16 == 0) && ((uintptr_t)a->cos_sin % 16 == 0), "crab_attn_decode_fused: alignment");
  CRAB_REQUIRE(a->ldq >= (a->H + 2 * a->KVH) * a->head_dim, "crab_attn_decode_fused: ldq smaller than the [q|k|v] row");
  if (a->lora_ra) {
    CRAB_REQUIRE(a->lora_z && a->lora_ws && a->lora_counters && a->ld_ra >= a->H * a->head_dim && a->ld_z >= 24,
                 "crab_attn_decode_fused: LoRA pre-pass needs z, workspace [B*KVH*11] floats ([B*H*11] when nsplit > 1) and counters [B] ints");
  }
  const int G = a->H / a->KVH;
  if (a->gqa_tensor_cores) {
    // Grouped-query decode on tensor cores: the G query heads of a kv group are the Sq = G rows of one flash-attention problem
    // (q row stride = head_dim), the key count is *past_dev + 1; RoPE + cache append run as their own launch.  With nsplit > 1
    // every (b, kv head) is split over nsplit blocks (128 problems alone cannot keep HBM busy: 32 KB in flight each) and the
    // combine launch finishes — and does the o_proj pre-pass.
    CRAB_REQUIRE(a->head_dim == 128 && G >= 1 && G <= 64, "crab_attn_decode_fused: gqa_tensor_cores needs head_dim 128 and a group of <= 64 heads");
    CRAB_REQUIRE(!a->lora_ra || a->nsplit > 1, "crab_attn_decode_fused: with gqa_tensor_cores the LoRA pre-pass runs in the combine launch (nsplit > 1)");
    int rc = crab_rope_kv_append(const_cast<void*>(a->qkv), a->ldq, a->cos_sin, a->k_cache, a->v_cache, a->B, 1, a->H, a->KVH, a->head_dim,
                                 a->ctx_max, a->past_dev, 0, stream);
    if (rc != CRAB_OK) return rc;
    FlashParams fp;
    fp.q = (const __nv_bfloat16*)a->qkv; fp.k = (const __nv_bfloat16*)a->k_cache; fp.v = (const __nv_bfloat16*)a->v_cache; fp.o = (__nv_bfloat16*)a->o;
    fp.q_bs = a->ldq; fp.q_rs = a->head_dim; fp.q_hs = (long long)G * a->head_dim;
    fp.k_bs = (long long)a->KVH * a->ctx_max * a->head_dim; fp.k_rs = a->head_dim; fp.k_hs = (long long)a->ctx_max * a->head_dim;
    fp.v_bs = fp.k_bs; fp.v_rs = fp.k_rs; fp.v_hs = fp.k_hs;
    fp.o_bs = a->ldo; fp.o_rs = a->head_dim; fp.o_hs = (long long)G * a->head_dim;
    fp.B = a->B; fp.H = a->KVH; fp.KVH = a->KVH; fp.Sq = G; fp.Sk = a->ctx_max;
    fp.scale = a->scale; fp.causal = 0; fp.gate = nullptr; fp.table = nullptr; fp.sk_dev = a->past_dev; fp.sk_add = 1;
    fp.nsplit = a->nsplit; fp.ws = a->workspace; fp.stages = 3;
    fp.use_tma = 0;
    fp.q_row_per_b = fp.q_col_per_h = fp.k_row_per_b = fp.k_row_per_h = fp.k_col_per_h = fp.v_row_per_b = fp.v_row_per_h = fp.v_col_per_h = 0;
    CUtensorMap tnone;
    memset(&tnone, 0, sizeof(tnone));
    static DeviceOnce fset;
    if (first_on_device(fset))
    {
      CRAB_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<128, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FlashCfg<128, 1>::SMEM3));
      CRAB_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<128, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FlashCfg<128, 1>::SMEM3));
    }
    cudaStream_t st2 = (cudaStream_t)stream;
    if (G <= 16) flash_attn_kernel<128, 1, true><<<dim3((unsigned)a->nsplit, (unsigned)a->KVH, (unsigned)a->B), 128, FlashCfg<128, 1>::SMEM3, st2>>>(tnone, tnone, tnone, fp);
    else flash_attn_kernel<128, 1><<<dim3((unsigned)a->nsplit, (unsigned)a->KVH, (unsigned)a->B), 128, FlashCfg<128, 1>::SMEM3, st2>>>(tnone, tnone, tnone, fp);
    CRAB_CHECK_CUDA(cudaGetLastError());
    if (a->nsplit > 1) {
      cudaError_t e2 = launch_pdl(PDL_ATTN, attn_decode_combine_kernel<128>, dim3(a->B * a->H), dim3(128), 0, st2, (const float*)a->workspace,
                                  (__nv_bfloat16*)a->o, a->ldo, a->H, a->nsplit, (const __nv_bfloat16*)a->lora_ra, a->ld_ra,
                                  (__nv_bfloat16*)a->lora_z, a->ld_z, a->lora_scale, a->lora_ws, a->lora_counters);
      CRAB_CHECK_CUDA(e2);
    }
    return CRAB_OK;
  }
  DecodeParams p;
  p.q = (const __nv_bfloat16*)a->qkv; p.ldq = a->ldq; p.kc = (const __nv_bfloat16*)a->k_cache; p.vc = (const __nv_bfloat16*)a->v_cache;
  p.o = (__nv_bfloat16*)a->o; p.ldo = a->ldo; p.ws = a->workspace; p.B = a->B; p.H = a->H; p.KVH = a->KVH; p.ctx_max = a->ctx_max;
  p.nsplit = a->nsplit; p.len_dev = nullptr; p.len_host = 0; p.scale = a->scale;
  p.late_trigger = (pdl_mask() & PDL_ATTN_LATE) ? 1 : 0;
  p.cos_sin = a->cos_sin; p.past_dev = a->past_dev; p.kc_w = (__nv_bfloat16*)a->k_cache; p.vc_w = (__nv_bfloat16*)a->v_cache;
  p.ra = (const __nv_bfloat16*)a->lora_ra; p.ldra = a->ld_ra; p.z = (__nv_bfloat16*)a->lora_z; p.ldz = a->ld_z;
  p.lora_scale = a->lora_scale; p.lora_ws = a->lora_ws; p.lora_cnt = a->lora_counters;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  // split KV at MHA / head_dim 128: the splits of one (b, head) are the blocks of a cluster (<= 8) and merge through DSMEM
  static int cluster_ok = -1;
  if (cluster_ok < 0) { const char* ev = getenv("CRAB_ATTN_CLUSTER"); cluster_ok = (ev && ev[0] == '0') ? 0 : 1; }
  p.csplit = 0;
  if (cluster_ok && a->nsplit > 1 && G == 1 && a->head_dim == 128) {
    p.csplit = a->nsplit > 8 ? 8 : a->nsplit;
    p.nsplit = 1;                       // the kernel finishes itself: output + pre-pass as in the unsplit launch
    p.trace = next_trace_slot(a->B * a->KVH * p.csplit);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a->B * a->KVH * p.csplit));
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)p.csplit; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_mask() & PDL_ATTN) ? 2 : 1;
    CRAB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_decode_kernel<128, 1, true>, p));
    return CRAB_OK;
  }
  p.trace = next_trace_slot(a->B * a->KVH * a->nsplit);
  dim3 grid(a->B * a->KVH, a->nsplit);
#define CRAB_DECODE_CASE(HD_, G_) \
  if (a->head_dim == HD_ && G == G_) { e = launch_pdl(PDL_ATTN, attn_decode_kernel<HD_, G_, true>, grid, dim3(128), 0, st, p); } else
  CRAB_DECODE_CASE(128, 1) CRAB_DECODE_CASE(128, 2) CRAB_DECODE_CASE(128, 4) CRAB_DECODE_CASE(128, 7)
  CRAB_DECODE_CASE(128, 8) CRAB_DECODE_CASE(64, 1)
  { return set_error(CRAB_ERR_INVALID, "crab_attn_decode_fused: unsupported head_dim=%d group=%d", a->head_dim, G); }
#undef CRAB_DECODE_CASE
  CRAB_CHECK_CUDA(e);
  if (a->nsplit > 1) {
    // split KV: the o_proj pre-pass moves to the combine kernel (one block per (b, head): lora_ws holds B * H * 11 floats)
    if (a->head_dim == 128) e = launch_pdl(PDL_ATTN, attn_decode_combine_kernel<128>, dim3(a->B * a->H), dim3(128), 0, st, (const float*)a->workspace, p.o, a->ldo, a->H, a->nsplit,
                                           p.ra, p.ldra, p.z, p.ldz, p.lora_scale, p.lora_ws, p.lora_cnt);
    else e = launch_pdl(PDL_ATTN, attn_decode_combine_kernel<64>, dim3(a->B * a->H), dim3(64), 0, st, (const float*)a->workspace, p.o, a->ldo, a->H, a->nsplit,
                        p.ra, p.ldra, p.z, p.ldz, p.lora_scale, p.lora_ws, p.lora_cnt);
    CRAB_CHECK_CUDA(e);
  }
  return CRAB_OK;
}
