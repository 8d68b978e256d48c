// Weight-streaming GEMM for the decode step (M <= 32 token rows):  C[M,N] = epilogue(X[M,K] . W[N,K]^T)
//
// At M <= 32 the op is HBM-bound (2 FLOP per weight byte): the job is to stream W once at full bandwidth.  The
// operands are swapped so the weights take the 128-row A slot of the tensor core and the batch takes the N slot:
//     D[128 weight rows, 32 batch columns] += W_tile[128, 64] . X_tile[32, 64]^T      (tcgen05.mma M=128 N=32 K=16)
// so no weight byte is fetched twice and the X tile (4 KB / k-block) is the only redundant traffic.
//
// N/128 tiles cannot fill 148 SMs for the 4096-wide projections, so K is split S ways and the S CTAs of one tile form
// a THREAD-BLOCK CLUSTER.  The split-K reduction never touches global memory: after its K-slice each CTA holds a
// 128x32 fp32 partial in registers (from TMEM); the cluster does a reduce-scatter over DISTRIBUTED SHARED MEMORY —
// rank j owns a group of rows; every rank parks its partial in its own (by then idle) TMA ring and sends the rows of
// group j as bulk shared->shared copies into rank j's dedicated receive area, reporting the bytes to an mbarrier there;
// rank j sums its S partials in rank order (deterministic) and stores.  No workspace, no atomics, no cluster barrier or
// release fence after the last MMA (both compile to MEMBAR.ALL.GPU: 1-3 us inside a saturated stream).
// What follows the last MMA runs once per launch with one warp per scheduler, i.e. at instruction-FETCH speed (~0.25 us
// per 128-byte line of cold straight-line code): bias / residual are fetched behind the weight stream and pre-added by the
// thread that owns the row, and the finishing loop is one compact rolled loop per cluster size (sk_finish_split).
// History with measurements: profiles/r01_skinny_epilogue_trace.txt (workspace + ticket -> cluster/DSMEM) and
// profiles/r04_skinny_timeline.txt (in-kernel timeline: 24 us per layer of finishing loops -> 3 us).
//
// Fused form (the decode step): the packed weight carries W.diag(gamma) so the launch reads the raw residual stream and
// multiplies its accumulator by rstd[b]; rstd and the hyper-LoRA pre-pass z' come from STATISTICS CLUSTERS, the first
// clusters of the same grid (see is_stats below), through a flag in global memory.
//
// Weights are read either through a TMA tensor map (row-major [N, ldw]) or — the decode path — from the streaming
// layout of crab_pack_skinny_weight: every (tile, k-block) is one contiguous, pre-swizzled 16 KB block, fetched with
// a 1-D cp.async.bulk, so a CTA's K-slice is one sequential span of HBM.
//
//   warp 0 : TMA producer (W 128x64 + X 32x64 per stage, SWIZZLE_128B)     warp 1 : MMA issuer + TMEM owner
//   warps 2-5 : TMEM -> registers -> (rstd, bias, residual) -> DSMEM reduce-scatter -> finishing loop
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace crab {

static constexpr int SK_BM = 128;   // weight rows per tile
static constexpr int SK_MB = 32;    // batch columns (UMMA N)
static constexpr int SK_BK = 64;
static constexpr int SK_STAGES = 4;  // 80 KB ring + 18 KB receive area: two CTAs per SM
static constexpr int SK_W_BYTES = SK_BM * SK_BK * 2;
static constexpr int SK_X_BYTES = SK_MB * SK_BK * 2;
static constexpr int SK_STAGE_BYTES = SK_W_BYTES + SK_X_BYTES;
static constexpr int SK_PART_BYTES = 20 * 1024;   // receive area of the split-K reduce-scatter: [src rank][rows per rank][36 floats]; S * R <= 140 rows (S = 7)
static constexpr int SK_SMEM = SK_STAGES * SK_STAGE_BYTES + SK_PART_BYTES + 1024 + 256;
static constexpr int SK_THREADS = 192;
static constexpr int SK_MAX_SPLIT = 8;  // portable cluster size
static constexpr int SK_PSTRIDE = 36;   // floats per partial row in smem (32 + pad, keeps 16-byte alignment)

struct SkinnyParams {
  void* C;
  const float* bias;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* w_tiled;  // non-null: weights pre-packed as contiguous, pre-swizzled 16 KB (tile, k-block) blocks
  int M, N, K, ldc, ldr;
  int act, out_dtype;
  int splits, kb_per_tile, rows_per_rank;
  // ---- in-launch statistics (RMSNorm as an epilogue scale + hyper-LoRA pre-pass), see crab_skinny_args ----
  const __nv_bfloat16* stats_w;  // packed [kb_main][40 x 64] router/A rows (gamma folded) or null
  __nv_bfloat16* zbuf;           // z' columns [M, ldz]: written by the statistics cluster, read through tmap_z
  float* rstd;                   // [32]
  int* flags;                    // [0] z / rstd published, [32] CTAs that have left (self-cleaning)
  int n_tiles, kb_main, ldz, has_stats, norm, stats_linears, ext_from_z;
  int stats_clusters;          // C: clusters that share the statistics item (0 when the launch has none)
  float* stats_scratch;        // [C][3][32][12] floats (C > 1): per-cluster partial sums, combined by the last cluster to arrive
  int* flags_clear;            // optional: flag slot of the PREVIOUS launch, zeroed here (then this launch leaves its own slot set)
  float eps, lora_scale;
  // L2 prefetch of the NEXT launch's weight stream: each CTA touches its share once its own loads are all issued, so HBM keeps
  // working through this launch's reduce / epilogue / exit and the next launch's ramp instead of idling between two kernels
  const uint8_t* pf_ptr;
  unsigned long long pf_bytes;
  unsigned long long* trace;   // diagnostics: [ctas][16] stamps or nullptr
  int debug;                   // diagnostics (env CRAB_SKINNY_DEBUG): 1 = no final stores, 2 = no partial loads
};
static constexpr int SK_SROWS = 40;                       // router/A rows per statistics k-block (33 used)
static constexpr int SK_S_BYTES = SK_SROWS * SK_BK * 2;   // 5 KB
static constexpr int SK_SSROWS = 32 + SK_SROWS;           // tile rows of a statistics item that carry data: x x^T, then the dots
__device__ __forceinline__ int sk_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void sk_wait_flag(const int* p) {
  if (sk_ld_acquire(p) > 0) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (sk_ld_acquire(p) <= 0) {
    __nanosleep(32);
    if ((++spins & 0xff) == 0 && globaltimer_ns() - t0 > CRAB_MBAR_TIMEOUT_NS) {
      printf("crab: skinny statistics flag timeout block=%d thread=%d\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Remote shared-memory store that reports its bytes to an mbarrier of the destination CTA: the receiver waits for the expected
// byte count, no fence on either side.  (barrier.cluster.arrive.release and mbarrier.arrive.release.cluster both compile to
// MEMBAR.ALL.GPU, which inside a saturated weight stream costs 1-3 us per use: profiles/r04_skinny_timeline.txt.)
__device__ __forceinline__ void st_async_f4(uint32_t remote_addr, uint32_t remote_bar, float a, float b, float c, float d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"(remote_bar) : "memory");
}
// smem -> smem of another CTA of the cluster, bytes reported to an mbarrier of the destination CTA (one transaction per call)
__device__ __forceinline__ void bulk_copy_to_rank(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(remote_dst),
               "r"(local_src), "r"(bytes), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void st_async_f1(uint32_t remote_addr, uint32_t remote_bar, float a) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(remote_addr), "f"(a), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Final step of the reduce-scatter for S_ in {1, 2, 4, 8} (R = 128 / S_ tile rows per rank).  The partials in `part` already carry
// the RMSNorm scale, the bias and the residual (added by the thread that produced the rank's own partial), so this is a pure
// fixed-order sum + convert + store.  Deliberately ROLLED loops with a small body: the code runs once per launch with one warp
// per scheduler, i.e. at instruction-fetch speed — the in-kernel timeline (profiles/r04_skinny_timeline.txt) showed ~0.25 us per
// 128-byte line of cold straight-line code and ~30 ns per line for loop bodies that do not fit the L0 instruction cache.
template <int S_>
__device__ __forceinline__ void sk_finish_split(const SkinnyParams& p, const float* __restrict__ part, int tt, int tile, int rank) {
  constexpr int R = 128 / S_;
  const int row0 = rank * R;
  if (p.act == CRAB_ACT_SWIGLU) {
    constexpr int PAIRS = R / 2, IT = R / 8, BSTEP = 128 / PAIRS;
    const int pr = tt & (PAIRS - 1), b0 = tt / PAIRS;
    const int n = tile * 64 + ((row0 + 2 * pr) >> 1);
    const float* pp = part + (2 * pr) * SK_PSTRIDE + b0;
    __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)b0 * p.ldc + n;
    const int mlim = (n < (p.N >> 1)) ? p.M : 0;
    const size_t cstep = (size_t)BSTEP * p.ldc;
#pragma unroll 2
    for (int b = b0; b < 32; b += BSTEP) {
      float g = 0.f, u = 0.f;
#pragma unroll
      for (int s = 0; s < S_; ++s) {  // fixed rank order: deterministic
        g += pp[s * R * SK_PSTRIDE];
        u += pp[s * R * SK_PSTRIDE + SK_PSTRIDE];
      }
      if (b < mlim) *cp = __float2bfloat16_rn(g / (1.0f + __expf(-g)) * u);
      pp += BSTEP;
      cp += cstep;
    }
    (void)IT;
  } else {
    constexpr int BSTEP = 128 / R;
    const int rl = tt & (R - 1), b0 = tt / R;
    const int n = tile * SK_BM + row0 + rl;
    const float* pp = part + rl * SK_PSTRIDE + b0;
    const int mlim = (n < p.N) ? p.M : 0;
    if (p.out_dtype == CRAB_BF16) {
      __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)b0 * p.ldc + n;
      const size_t cstep = (size_t)BSTEP * p.ldc;
#pragma unroll 2
      for (int b = b0; b < 32; b += BSTEP) {
        float a = 0.f;
#pragma unroll
        for (int s = 0; s < S_; ++s) a += pp[s * R * SK_PSTRIDE];
        if (b < mlim) *cp = __float2bfloat16_rn(a);
        pp += BSTEP;
        cp += cstep;
      }
    } else {
      float* cp = reinterpret_cast<float*>(p.C) + (size_t)b0 * p.ldc + n;
      const size_t cstep = (size_t)BSTEP * p.ldc;
#pragma unroll 2
      for (int b = b0; b < 32; b += BSTEP) {
        float a = 0.f;
#pragma unroll
        for (int s = 0; s < S_; ++s) a += pp[s * R * SK_PSTRIDE];
        if (b < mlim) *cp = a;
        pp += BSTEP;
        cp += cstep;
      }
    }
  }
}

__global__ void __launch_bounds__(SK_THREADS, 2)
gemm_skinny_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                           const __grid_constant__ CUtensorMap tmap_z, const SkinnyParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t part_base = smem_base + SK_STAGES * SK_STAGE_BYTES;   // dedicated receive area: peers may write it while the ring still streams
  const uint32_t bar_base = part_base + SK_PART_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SK_STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * SK_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * SK_STAGES + 1);
  const uint32_t rfree_bar = bar_base + 8u * (2 * SK_STAGES + 2);   // one arrival per rank of the cluster: "my MMAs are done, my ring may be overwritten"
  const uint32_t pfull_bar = bar_base + 8u * (2 * SK_STAGES + 3);   // bytes of the partials this CTA receives
  const uint32_t sent_bar = bar_base + 8u * (2 * SK_STAGES + 4);    // one arrival per rank: "your partials have arrived here" (source smem may go)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) trace_stamp(p.trace, 0);
  const int S = p.splits;
  const int SC = p.stats_clusters;
  const int tile = (int)blockIdx.x / S - SC;   // weight tile of this cluster (< 0: a statistics cluster)
  const int rank = (S > 1) ? (int)cluster_ctarank() : 0;
  // The first SC clusters of the launch are STATISTICS clusters (when the launch has any): their A tile is
  // [gamma*[R;A] rows ; x rows (32)], so the same MMA chain yields diag(x x^T) = sum x^2 and the hyper-LoRA router / A dots over
  // their share of K; rank 0 of each reduces its cluster, the last cluster to arrive (ticket) adds the clusters' partials in
  // cluster order and publishes rstd and z' = scale * softmax(rstd * logits)_i * u_j, then raises flags[0].  The other clusters
  // read z' (their K-extension k-blocks, the last ones of the last rank) and rstd (epilogue scale) only after that flag.
  // They come FIRST in the grid: clusters are placed in block order, so they are resident before any cluster that will wait for
  // the flag (a last-placed statistics cluster can be starved of a slot by the very CTAs that spin on it).
  const bool is_stats = blockIdx.x < (unsigned)(SC * S);
  const int KB = is_stats ? p.kb_main : p.kb_per_tile;
  // K-slice: a weight tile is split over the S ranks of its cluster, the statistics item over the SC * S CTAs of its clusters
  const int kpart = is_stats ? (int)blockIdx.x : rank, kparts = is_stats ? SC * S : S;
  const int kb0 = kpart * KB / kparts, kb1 = (kpart + 1) * KB / kparts;

  if (warp == 0 && lane == 0) {
    if (!p.w_tiled) tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < SK_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(rfree_bar, (uint32_t)S);
    mbar_init(pfull_bar, 1);
    mbar_init(sent_bar, (uint32_t)S);
    fence_barrier_init();
    // partials this CTA will receive: a weight tile's rank gets its rows from all S ranks (128 B each), rank 0 of a statistics
    // cluster gets the router/A dot rows and the 32 diagonal elements of x x^T from all S ranks
    const int rpr = p.rows_per_rank;
    const int rows_mine = max(0, min(rpr, SK_BM - rank * rpr));
    const uint32_t expect = S == 1 ? 0u : is_stats ? (rank == 0 ? (uint32_t)S * (SK_SROWS * 128u + 32u * 4u) : 0u) : (uint32_t)(S * rows_mine) * (SK_PSTRIDE * 4u);
    if (expect) mbar_arrive_expect_tx(pfull_bar, expect);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // the peers' barriers exist once every thread of the cluster has arrived here; the epilogue waits for that before its first remote op
  if (S > 1) asm volatile("barrier.cluster.arrive.release;" ::: "memory");

  pdl_trigger();
  if (warp == 0) {
    if (lane == 0) {
      // ===================== producer =====================
      // Weights never depend on an earlier kernel: the ring is filled with W tiles before the PDL wait.
      const int depth = (p.debug >> 8) > 0 ? min(SK_STAGES, p.debug >> 8) : SK_STAGES;   // diagnostics: loads in flight per CTA
      const int npre = min(depth, kb1 - kb0);
      const size_t blk0 = (size_t)tile * p.kb_per_tile;
      auto load_w = [&](int slot_i, int kb) {
        const uint32_t sw = smem_base + slot_i * SK_STAGE_BYTES;
        if (is_stats) {
          // x tile (it is both the B operand and, for x x^T, an A operand) + the router/A rows: armed here, x follows the PDL wait
          mbar_arrive_expect_tx(full_bar(slot_i), SK_X_BYTES + (p.stats_w ? SK_S_BYTES : 0));
          if (p.stats_w) bulk_load_1d_hint(sw + SK_X_BYTES, p.stats_w + (size_t)kb * (SK_SROWS * SK_BK), SK_S_BYTES, full_bar(slot_i), kEvictFirst);
        } else {
          mbar_arrive_expect_tx(full_bar(slot_i), SK_STAGE_BYTES);
          if (p.w_tiled) bulk_load_1d_hint(sw, p.w_tiled + (blk0 + kb) * (SK_BM * SK_BK), SK_W_BYTES, full_bar(slot_i), kEvictFirst);
          else tma_load_2d_hint(sw, &tmap_w, full_bar(slot_i), kb * SK_BK, tile * SK_BM, kEvictFirst);
        }
      };
      bool z_ok = false;
      auto load_x = [&](int slot_i, int kb) {
        const uint32_t sw = smem_base + slot_i * SK_STAGE_BYTES;
        if (is_stats) {
          tma_load_2d_hint(sw + SK_W_BYTES, &tmap_x, full_bar(slot_i), kb * SK_BK, 0, kEvictLast);
        } else if (p.ext_from_z && kb >= p.kb_main) {
          if (p.has_stats && !z_ok) {   // z' comes from this launch's statistics cluster
            sk_wait_flag(p.flags);
            asm volatile("fence.proxy.async;" ::: "memory");
            z_ok = true;
          }
          tma_load_2d_hint(sw + SK_W_BYTES, &tmap_z, full_bar(slot_i), (kb - p.kb_main) * SK_BK, 0, kEvictLast);
        } else {
          tma_load_2d_hint(sw + SK_W_BYTES, &tmap_x, full_bar(slot_i), kb * SK_BK, 0, kEvictLast);
        }
      };
      for (int i = 0; i < npre; ++i) load_w(i, kb0 + i);
      trace_stamp(p.trace, 1);
      pdl_wait();
      trace_stamp(p.trace, 2);
      for (int i = 0; i < npre; ++i) load_x(i, kb0 + i);
      uint32_t stage = 0, phase = 1;  // ring position after the prefill (npre == SK_STAGES wraps to stage 0)
      if (npre < SK_STAGES) { stage = (uint32_t)npre; phase = 0; }
      for (int kb = kb0 + npre; kb < kb1; ++kb) {
        if (depth < SK_STAGES) {   // wait until load (i - depth) has landed before issuing load i
          const int j = kb - kb0 - depth;
          mbar_wait(full_bar(j % SK_STAGES), (uint32_t)((j / SK_STAGES) & 1));
        }
        mbar_wait(empty_bar(stage), phase ^ 1);
        load_w((int)stage, kb);
        load_x((int)stage, kb);
        if (++stage == SK_STAGES) { stage = 0; phase ^= 1; }
      }
      if (p.pf_bytes) {
        const unsigned long long per = ((p.pf_bytes / gridDim.x) + 16383ull) & ~16383ull;
        const unsigned long long b0 = (unsigned long long)blockIdx.x * per;
        for (unsigned long long o = b0; o < b0 + per && o < p.pf_bytes; o += 16384ull) {
          const unsigned long long n = (p.pf_bytes - o < 16384ull) ? (p.pf_bytes - o) & ~15ull : 16384ull;
          if (n) bulk_prefetch_l2(p.pf_ptr + o, (uint32_t)n);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc_bf16_f32(SK_BM, SK_MB);
      uint32_t stage = 0, phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        if (kb == kb0) trace_stamp(p.trace, 3);
        tc_fence_after();
        const uint32_t sw = smem_base + stage * SK_STAGE_BYTES;
        const uint64_t da = make_sdesc_sw128(sw);
        const uint64_t db = make_sdesc_sw128(sw + SK_W_BYTES);
        // Statistics cluster: the A tile starts 4 KB into the slot, so its rows 0..39 are the gamma*[R;A] rows and its rows
        // 96..127 are the x tile itself (slot + 16 KB): ONE MMA chain gives the router / A dots (TMEM lanes 0..39) and
        // x x^T (lanes 96..127, diagonal = sum x^2) from a single copy of x.
        const uint64_t da_ = is_stats ? make_sdesc_sw128(sw + SK_X_BYTES) : da;
#pragma unroll
        for (int k = 0; k < SK_BK / 16; ++k) umma_bf16_ss(tmem_base, da_ + 2u * k, db + 2u * k, idesc, (kb > kb0) | (k > 0));
        umma_commit(empty_bar(stage));
        if (++stage == SK_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  }
  __syncwarp();

  // ===================== epilogue =====================
  const bool epi = warp >= 2;
  const int quarter = warp & 3;
  const int row = quarter * 32 + lane;  // weight row of this thread within the tile (epilogue warps)
  const bool swiglu = p.act == CRAB_ACT_SWIGLU;
  const int R = p.rows_per_rank;  // even; rank j owns tile rows [j*R, min(128, (j+1)*R))
  // a thread whose tile row stays in this CTA after the reduce-scatter adds the bias and the residual to its partial: the
  // values are fetched NOW, behind the weight stream (packed two bf16 per register), not between the last MMA and the store
  const bool own = epi && !is_stats && (S == 1 || row / R == rank) && (tile * SK_BM + row) < p.N && !swiglu;
  const bool add_res = own && p.residual != nullptr;
  uint32_t r[32];
  uint32_t resp[16];
  float bias_v = 0.f;
  if (epi) {
    pdl_wait();
    if (p.flags_clear && blockIdx.x == 0 && threadIdx.x == 64) {   // the launch that used that slot is complete
      p.flags_clear[0] = 0;
      p.flags_clear[1] = 0;
    }
    if (own && p.bias) bias_v = p.bias[tile * SK_BM + row];
    if (add_res) {
      const unsigned short* rp = reinterpret_cast<const unsigned short*>(p.residual) + (tile * SK_BM + row);
#pragma unroll
      for (int b = 0; b < 16; ++b) {
        const uint32_t lo = (2 * b < p.M) ? rp[(size_t)(2 * b) * p.ldr] : 0u;
        const uint32_t hi = (2 * b + 1 < p.M) ? rp[(size_t)(2 * b + 1) * p.ldr] : 0u;
        resp[b] = lo | (hi << 16);
      }
    }
    mbar_wait(tfull_bar, 0);
    if (threadIdx.x == 64) trace_stamp(p.trace, 4);
    tc_fence_after();
    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16), r);
    tmem_ld_wait();
  }

  __shared__ float rstd_s[32];
  if (is_stats) {
    // ---- statistics cluster: partial x x^T rows / router-A dots of every rank -> rank 0 -> rstd, z', flag ----
    if (epi) {
      if (S > 1) {
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
        // every rank tells every rank that its MMAs are done (tfull was observed above), i.e. that its TMA ring may hold partials
        if (warp == 2 && lane < S) mbar_arrive_remote_relaxed(map_to_rank(rfree_bar, (uint32_t)lane));
        mbar_wait(rfree_bar, 0);
      }
      if (threadIdx.x == 64) trace_stamp(p.trace, 8);
      const uint32_t pf0 = (S > 1) ? map_to_rank(pfull_bar, 0u) : 0u;
      if (row < SK_SROWS) {
        // dots row `row` -> slot 32 + row of rank 0's buffer
        const uint32_t local = smem_base + (uint32_t)((rank * SK_SSROWS + 32 + row) * SK_PSTRIDE * 4);
        if (S > 1) {
          const uint32_t remote = map_to_rank(local, 0u);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            st_async_f4(remote + g * 16, pf0, __uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]),
                        __uint_as_float(r[4 * g + 3]));
        } else {   // a launch without cluster dimensions has no shared::cluster window (st.async / remote arrive are illegal there)
#pragma unroll
          for (int g = 0; g < 8; ++g)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(local + g * 16), "r"(r[4 * g]), "r"(r[4 * g + 1]), "r"(r[4 * g + 2]),
                         "r"(r[4 * g + 3]) : "memory");
        }
      } else if (row >= 96) {
        // x x^T row of batch row b = row - 96: only its diagonal element is needed -> slot b, column b
        const int b = row - 96;
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) d = (c == b) ? __uint_as_float(r[c]) : d;
        const uint32_t dl = smem_base + (uint32_t)(((rank * SK_SSROWS + b) * SK_PSTRIDE + b) * 4);
        if (S > 1) st_async_f1(map_to_rank(dl, 0u), pf0, d);
        else asm volatile("st.shared.f32 [%0], %1;" ::"r"(dl), "f"(d) : "memory");
      }
      if (threadIdx.x == 64) trace_stamp(p.trace, 9);
      if (S == 1) asm volatile("bar.sync 1, 128;" ::: "memory");
      else if (rank == 0) mbar_wait(pfull_bar, 0);
      if (threadIdx.x == 64) trace_stamp(p.trace, 6);
    }
    if (rank == 0 && epi) {
      const int tt = (warp - 2) * 32 + lane;
      const int b = tt & 31, l = tt >> 5;
      const int L = p.stats_linears;
      const int nw = L > 1 ? L : 1;
      const float* part = reinterpret_cast<const float*>(smem_raw + (smem_base - smem_u32(smem_raw)));
      if (l < nw) {
        // this cluster's sums (fixed rank order): thread (b, l) owns sum x^2 of batch row b and the 11 dots of linear l
        float ss = 0.f, t[11];
        for (int s = 0; s < S; ++s) ss += part[(s * SK_SSROWS + b) * SK_PSTRIDE + b];   // diagonal of x x^T
#pragma unroll
        for (int j = 0; j < 11; ++j) t[j] = 0.f;
        if (l < L) {
#pragma unroll 1
          for (int s = 0; s < S; ++s) {
            const float* pr = part + (s * SK_SSROWS + 32 + l * 11) * SK_PSTRIDE + b;
#pragma unroll
            for (int j = 0; j < 11; ++j) t[j] += pr[j * SK_PSTRIDE];
          }
        }
        bool fin = true;
        if (SC > 1) {
          // several statistics clusters: partials meet in global memory, the last cluster to arrive adds them in cluster order.
          // scratch[c][l][b][12] = {ss, t[0..10]} of thread (b, l): three 16-byte stores / loads per thread and cluster
          float4* sc = reinterpret_cast<float4*>(p.stats_scratch) + ((size_t)(blockIdx.x / S) * (3 * 32) + l * 32 + b) * 3;
          sc[0] = make_float4(ss, t[0], t[1], t[2]);
          sc[1] = make_float4(t[3], t[4], t[5], t[6]);
          sc[2] = make_float4(t[7], t[8], t[9], t[10]);
          // one acq_rel ticket by thread 0 between two CTA barriers: its release covers the block's partial stores (cumulativity
          // through the barrier), its acquire the last block's loads below — instead of a GPU-scope fence per thread on each side
          // (each costs ~2 us while 300 CTAs stream weights)
          asm volatile("bar.sync 2, %0;" ::"r"(32 * nw) : "memory");
          __shared__ int ticket_s;
          if (tt == 0) {
            int tk;
            asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(tk) : "l"(p.flags + 1) : "memory");
            ticket_s = tk;
          }
          asm volatile("bar.sync 2, %0;" ::"r"(32 * nw) : "memory");
          fin = ticket_s == SC - 1;
          if (threadIdx.x == 64) trace_stamp(p.trace, 14);
          if (fin) {
            ss = 0.f;
#pragma unroll
            for (int j = 0; j < 11; ++j) t[j] = 0.f;
            const float4* rd = reinterpret_cast<const float4*>(p.stats_scratch) + ((size_t)l * 32 + b) * 3;
#pragma unroll 4
            for (int c = 0; c < SC; ++c) {
              const float4 v0 = __ldcg(rd + (size_t)c * (3 * 32 * 3)), v1 = __ldcg(rd + (size_t)c * (3 * 32 * 3) + 1),
                           v2 = __ldcg(rd + (size_t)c * (3 * 32 * 3) + 2);
              ss += v0.x; t[0] += v0.y; t[1] += v0.z; t[2] += v0.w;
              t[3] += v1.x; t[4] += v1.y; t[5] += v1.z; t[6] += v1.w;
              t[7] += v2.x; t[8] += v2.y; t[9] += v2.z; t[10] += v2.w;
            }
            if (threadIdx.x == 64) trace_stamp(p.trace, 15);
          }
        }
        if (fin) {
          const float rs = p.norm ? rsqrtf(ss / (float)(p.kb_main * SK_BK) + p.eps) : 1.0f;
          if (l == 0 && p.norm && b < p.M) p.rstd[b] = rs;
          if (l < L && b < p.M) {
            const float l0 = t[0] * rs, l1 = t[1] * rs, l2 = t[2] * rs;
            const float mx = fmaxf(l0, fmaxf(l1, l2));
            const float e0 = __expf(l0 - mx), e1 = __expf(l1 - mx), e2 = __expf(l2 - mx);
            const float inv = p.lora_scale / (e0 + e1 + e2);
            const float rw[3] = {e0 * inv, e1 * inv, e2 * inv};
            __nv_bfloat16* zrow = p.zbuf + (size_t)b * p.ldz + l * 24;
#pragma unroll
            for (int i = 0; i < 3; ++i) {   // z' is NOT normalised: the consumers' epilogue multiplies the whole accumulator by rstd
              uint4 v;
              v.x = pack_bf16x2(rw[i] * t[3], rw[i] * t[4]);
              v.y = pack_bf16x2(rw[i] * t[5], rw[i] * t[6]);
              v.z = pack_bf16x2(rw[i] * t[7], rw[i] * t[8]);
              v.w = pack_bf16x2(rw[i] * t[9], rw[i] * t[10]);
              *reinterpret_cast<uint4*>(zrow + i * 8) = v;
            }
          }
          // rstd / z' stores of the block -> CTA barrier -> ONE release store (cumulative over the barrier); the consumers pair it
          // with ld.acquire + fence.proxy.async before their TMA loads of z'
          asm volatile("bar.sync 2, %0;" ::"r"(32 * nw) : "memory");
          if (tt == 0) {
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.flags), "r"(1) : "memory");
            trace_stamp(p.trace, 5);
          }
        }
      }
    }
  } else if (p.norm && epi) {
    // RMSNorm as an epilogue scale: gamma is folded into the packed weights, rstd[b] comes from the statistics cluster
    if (warp == 2) {
      if (lane == 0) sk_wait_flag(p.flags);
      __syncwarp();
      float v;
      asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p.rstd + lane) : "memory");
      rstd_s[lane] = lane < p.M ? v : 0.f;
      if (lane == 0) trace_stamp(p.trace, 5);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }

  if (!is_stats) {
    // ---- reduce-scatter of the K-split partials through (distributed) shared memory; S == 1 takes the same route through
    //      its own shared memory, so that there is ONE compact finishing loop ----
    if (epi) {
      if (p.norm) {
#pragma unroll
        for (int b = 0; b < 32; ++b) r[b] = __float_as_uint(__uint_as_float(r[b]) * rstd_s[b]);
      }
      if (add_res) {
#pragma unroll
        for (int b = 0; b < 16; ++b) {
          r[2 * b] = __float_as_uint(__uint_as_float(r[2 * b]) + __uint_as_float(resp[b] << 16));
          r[2 * b + 1] = __float_as_uint(__uint_as_float(r[2 * b + 1]) + __uint_as_float(resp[b] & 0xffff0000u));
        }
      }
      if (own && p.bias) {
#pragma unroll
        for (int b = 0; b < 32; ++b) r[b] = __float_as_uint(__uint_as_float(r[b]) + bias_v);
      }
    }
    // Reduce-scatter: every thread parks its tile row (32 floats) in this CTA's own ring — idle, its MMAs are done — and one lane
    // per warp sends the rows as bulk shared-to-shared copies into the owners' dedicated receive areas ([src rank][row in group]
    // [SK_PSTRIDE]); the copies report their bytes to the owner's pfull barrier.  No cluster barrier, no fence, no handshake before
    // the send (nobody streams into a receive area), one mbarrier update per copy instead of one per 16 bytes.
    if (epi) {
      const uint32_t stage_row = smem_base + (uint32_t)(row * SK_PSTRIDE * 4);
#pragma unroll
      for (int g = 0; g < 8; ++g)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + g * 16), "r"(r[4 * g]), "r"(r[4 * g + 1]), "r"(r[4 * g + 2]),
                     "r"(r[4 * g + 3]) : "memory");
      if (threadIdx.x == 64) trace_stamp(p.trace, 8);
      if (S > 1) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores above -> async-proxy reads of the copies
        __syncwarp();
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");     // the peers' barriers are initialised (arrive: prologue)
        if (lane == 0) {
          const int r0 = quarter * 32;
          for (int d = r0 / R; d * R < r0 + 32 && d < S; ++d) {   // destination ranks of this warp's 32 rows
            const int ra = max(r0, d * R), rb = min(r0 + 32, (d + 1) * R);
            bulk_copy_to_rank(map_to_rank(part_base + (uint32_t)((rank * R + (ra - d * R)) * SK_PSTRIDE * 4), (uint32_t)d),
                              smem_base + (uint32_t)(ra * SK_PSTRIDE * 4), (uint32_t)((rb - ra) * SK_PSTRIDE * 4),
                              map_to_rank(pfull_bar, (uint32_t)d));
          }
        }
        if (threadIdx.x == 64) trace_stamp(p.trace, 9);
        mbar_wait(pfull_bar, 0);   // all partials of this rank's rows have landed
        // tell every sender that its rows are here: a CTA may exit (and its ring be reused) only after its copies were read
        if (warp == 2 && lane < S) mbar_arrive_remote_relaxed(map_to_rank(sent_bar, (uint32_t)lane));
      } else {
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (threadIdx.x == 64) trace_stamp(p.trace, 6);
    }
    if (epi) {
      const int tt = (warp - 2) * 32 + lane;
      // S == 1: the parked rows are the result; S > 1: the receive area
      const float* part = reinterpret_cast<const float*>(smem_raw + ((S > 1 ? part_base : smem_base) - smem_u32(smem_raw)));
      if (S == 1) sk_finish_split<1>(p, part, tt, tile, rank);
      else if (S == 2) sk_finish_split<2>(p, part, tt, tile, rank);
      else if (S == 4) sk_finish_split<4>(p, part, tt, tile, rank);
      else if (S == 8) sk_finish_split<8>(p, part, tt, tile, rank);
      else {
        // other cluster sizes (3, 5, 6, 7: only on request, they schedule poorly): generic index arithmetic
        const int row0 = rank * R;
        const int rows_g = max(0, min(R, SK_BM - row0));
        if (swiglu) {
          const int pairs = rows_g >> 1;
          for (int idx = tt; idx < pairs * 32; idx += 128) {
            const int b = idx / pairs, pr = idx - b * pairs;
            float g = 0.f, u = 0.f;
            for (int s2 = 0; s2 < S; ++s2) {  // fixed rank order: deterministic
              g += part[(s2 * R + 2 * pr) * SK_PSTRIDE + b];
              u += part[(s2 * R + 2 * pr + 1) * SK_PSTRIDE + b];
            }
            const int n = tile * 64 + ((row0 + 2 * pr) >> 1);
            if (b < p.M && n < (p.N >> 1))
              reinterpret_cast<__nv_bfloat16*>(p.C)[(size_t)b * p.ldc + n] = __float2bfloat16_rn(g / (1.0f + __expf(-g)) * u);
          }
        } else {
          for (int idx = tt; idx < rows_g * 32; idx += 128) {
            const int b = idx / rows_g, rl = idx - b * rows_g;
            const int n = tile * SK_BM + row0 + rl;
            float a = 0.f;
            for (int s2 = 0; s2 < S; ++s2) a += part[(s2 * R + rl) * SK_PSTRIDE + b];
            if (b < p.M && n < p.N) {
              if (p.out_dtype == CRAB_BF16) reinterpret_cast<__nv_bfloat16*>(p.C)[(size_t)b * p.ldc + n] = __float2bfloat16_rn(a);
              else reinterpret_cast<float*>(p.C)[(size_t)b * p.ldc + n] = a;
            }
          }
        }
      }
    }
  }
  if (S > 1 && !is_stats && threadIdx.x == 64) mbar_wait(sent_bar, 0);   // every owner has received this CTA's rows
  if (threadIdx.x == 64) trace_stamp(p.trace, 10);
  if (threadIdx.x == 0) trace_stamp(p.trace, 11);
  if (threadIdx.x == 32) trace_stamp(p.trace, 12);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_stamp(p.trace, 7);
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 32); }
  // self-cleaning flags: the last CTA to get here (every flag wait is over by then) zeroes them for the next launch
  if (p.has_stats && !p.flags_clear && threadIdx.x == 0) {
    __threadfence();
    const int prev = atomicAdd(p.flags + 32, 1);
    if (prev == (int)gridDim.x - 1) {
      p.flags[0] = 0;
      p.flags[1] = 0;
      p.flags[32] = 0;
      __threadfence();
    }
  }
}

// Pre-pack a row-major weight [N, ldw] into the streaming layout: block f = tile * KB + kb holds the 128 x 64 tile
// exactly as the UMMA SWIZZLE_128B smem layout wants it (row r: 128 bytes, 16-byte chunk c stored at c ^ (r & 7)),
// zero-padded past N / K.  A CTA's K-slice is then ONE contiguous span of HBM.
// swiglu_interleave: the source rows are packed [64 gate | 64 up] per 128 (the prefill GEMM's layout); the decode layout
// interleaves them (tile row 2i = gate_i, 2i+1 = up_i) so the SwiGLU partner of a row is the neighbouring TMEM lane.
__global__ void pack_skinny_weight_kernel(const __nv_bfloat16* __restrict__ w, int N, int K, int ldw,
                                          __nv_bfloat16* __restrict__ out, int kb_per_tile, long long total_chunks,
                                          int swiglu_interleave) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk per thread
  if (i >= total_chunks) return;
  const int cs = (int)(i & 7);                 // stored chunk position
  const int r = (int)((i >> 3) & 127);         // row within tile
  const long long f = i >> 10;                 // block index
  const int kb = (int)(f % kb_per_tile), tile = (int)(f / kb_per_tile);
  const int c = cs ^ (r & 7);                  // source chunk
  const int src_r = swiglu_interleave ? ((r & 1) ? 64 + (r >> 1) : (r >> 1)) : r;
  const int row = tile * SK_BM + src_r, col = kb * SK_BK + c * 8;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (row < N) {
    if (col + 8 <= K) v = *reinterpret_cast<const uint4*>(w + (size_t)row * ldw + col);
    else if (col < K) {
      __nv_bfloat16 t[8];
      for (int e = 0; e < 8; ++e) t[e] = (col + e < K) ? w[(size_t)row * ldw + col + e] : __float2bfloat16_rn(0.f);
      v = *reinterpret_cast<uint4*>(t);
    }
  }
  *reinterpret_cast<uint4*>(out + i * 8) = v;
}

// K-split (= cluster size) for (N, K): about one CTA per SM (at most 1.3 x SMs), at least 4 k-blocks per CTA.  Measured inside the
// decode step (tools/sweep_splits_timeline.sh, profiles/r04_split_sweep_timeline.txt): a second CTA per SM adds cluster-exchange
// and scheduling cost but no bandwidth — LLaMA qkv (96 tiles) 2 > 4 > 1, o / down (32 tiles) 4 > 8 > 2, Qwen2 qkv (36 tiles) 4 > 8.
int choose_splits(int N, int K) {
  const int tiles = (N + SK_BM - 1) / SK_BM;
  const int kb = (K + SK_BK - 1) / SK_BK;
  const int slots = sm_count() * 13 / 10;
  int s = 1;
  // powers of two only: odd cluster sizes schedule poorly (measured: qkv S=3 33.6 us vs S=2 26.2 us)
  while (s * 2 <= SK_MAX_SPLIT && tiles * s * 2 <= slots && kb / (s * 2) >= 4) s *= 2;
  return s;
}

}  // namespace crab

using namespace crab;

extern "C" int crab_gemm_skinny_plan(int N, int K, int* splits, int64_t* workspace_bytes, int* n_counters) {
  CRAB_REQUIRE(N > 0 && K > 0 && splits && workspace_bytes && n_counters, "crab_gemm_skinny_plan: bad args");
  *splits = choose_splits(N, K);
  *workspace_bytes = 0;  // the split-K reduction lives in distributed shared memory
  *n_counters = 0;
  return CRAB_OK;
}

extern "C" int crab_skinny_packed_bytes(int N, int K, int64_t* bytes) {
  CRAB_REQUIRE(N > 0 && K > 0 && bytes, "crab_skinny_packed_bytes: bad args");
  *bytes = (int64_t)((N + SK_BM - 1) / SK_BM) * ((K + SK_BK - 1) / SK_BK) * SK_BM * SK_BK * 2;
  return CRAB_OK;
}

extern "C" int crab_pack_skinny_weight(const void* W, int N, int K, int ldw, void* out, int swiglu_interleave, void* stream) {
  CRAB_REQUIRE(W && out && N > 0 && K > 0 && ldw >= K && ldw % 8 == 0, "crab_pack_skinny_weight: bad args");
  CRAB_REQUIRE(((uintptr_t)W % 16 == 0) && ((uintptr_t)out % 128 == 0), "crab_pack_skinny_weight: alignment (W 16 B, out 128 B)");
  CRAB_REQUIRE(!swiglu_interleave || N % 128 == 0, "crab_pack_skinny_weight: SwiGLU interleave needs N %% 128 == 0");
  const int kb = (K + SK_BK - 1) / SK_BK;
  const long long chunks = (long long)((N + SK_BM - 1) / SK_BM) * kb * SK_BM * 8;
  pack_skinny_weight_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(W), N, K, ldw, reinterpret_cast<__nv_bfloat16*>(out), kb, chunks, swiglu_interleave);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_gemm_skinny_bf16(const crab_skinny_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CRAB_REQUIRE(a && a->X && (a->W || a->W_packed) && a->C, "crab_gemm_skinny_bf16: null pointer");
  CRAB_REQUIRE(a->M > 0 && a->M <= SK_MB, "crab_gemm_skinny_bf16: M must be in 1..32 (got %d)", a->M);
  CRAB_REQUIRE(a->N > 0 && a->K > 0 && a->ldx % 8 == 0 && a->ldx >= a->K, "crab_gemm_skinny_bf16: bad shape/strides N=%d K=%d ldx=%d",
               a->N, a->K, a->ldx);
  if (a->W_packed) CRAB_REQUIRE((uintptr_t)a->W_packed % 128 == 0, "crab_gemm_skinny_bf16: W_packed must be 128-byte aligned");
  else CRAB_REQUIRE(a->ldw % 8 == 0 && a->ldw >= a->K && ((uintptr_t)a->W % 16 == 0), "crab_gemm_skinny_bf16: W alignment / ldw=%d", a->ldw);
  CRAB_REQUIRE((uintptr_t)a->X % 16 == 0, "crab_gemm_skinny_bf16: X must be 16-byte aligned");
  CRAB_REQUIRE(a->act == CRAB_ACT_NONE || a->act == CRAB_ACT_SWIGLU, "crab_gemm_skinny_bf16: act must be NONE or SWIGLU");
  if (a->act == CRAB_ACT_SWIGLU)
    CRAB_REQUIRE(a->N % 128 == 0 && !a->bias && !a->residual && a->out_dtype == CRAB_BF16 && a->W_packed,
                 "crab_gemm_skinny_bf16: SWIGLU needs N %% 128 == 0, bf16 out, no bias/residual and a weight packed with swiglu_interleave");
  // ---- optional: K-extension columns from a separate buffer, RMSNorm as an epilogue scale, in-launch hyper-LoRA pre-pass ----
  const bool ext_z = a->Z != nullptr && a->Kext > 0;
  const bool has_stats = a->norm != 0 || a->stats_linears > 0;
  if (ext_z || has_stats) {
    CRAB_REQUIRE(a->W_packed && a->K % SK_BK == 0, "crab_gemm_skinny_bf16: Z / norm / stats need a packed weight and K %% 64 == 0");
    CRAB_REQUIRE(!ext_z || (a->Kext <= 128 && a->ldz % 8 == 0 && a->ldz >= a->Kext && ((uintptr_t)a->Z % 16 == 0)),
                 "crab_gemm_skinny_bf16: bad K-extension (Kext=%d ldz=%d)", a->Kext, a->ldz);
    CRAB_REQUIRE(a->stats_linears >= 0 && a->stats_linears <= 3 &&
                 (a->stats_linears == 0 || (a->stats_packed && ext_z && a->Kext >= 24 * a->stats_linears && ((uintptr_t)a->stats_packed % 128 == 0))),
                 "crab_gemm_skinny_bf16: a LoRA pre-pass needs stats_packed, Z and Kext >= 24 per linear");
    CRAB_REQUIRE(!a->norm || a->rstd, "crab_gemm_skinny_bf16: norm needs an rstd scratch buffer (32 floats)");
    CRAB_REQUIRE(!has_stats || (a->flags && ((uintptr_t)a->flags % 128 == 0)), "crab_gemm_skinny_bf16: norm / stats need `flags` (64 zeroed ints, 128-byte aligned)");
  }
  const int tiles = (a->N + SK_BM - 1) / SK_BM;
  const int kb_main = (a->K + SK_BK - 1) / SK_BK;
  const int kb = kb_main + (ext_z ? (a->Kext + SK_BK - 1) / SK_BK : 0);
  static DeviceOnce attr_once;
  if (first_on_device(attr_once)) {
    CRAB_CHECK_CUDA(cudaFuncSetAttribute(gemm_skinny_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM));
    CRAB_CHECK_CUDA(cudaFuncSetAttribute(gemm_skinny_tcgen05_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  int splits = a->splits > 0 ? a->splits : choose_splits(a->N, kb * SK_BK);
  if (splits > SK_MAX_SPLIT) splits = SK_MAX_SPLIT;
  if (splits > kb_main) splits = kb_main;
  CUtensorMap tw, tx, tz;
  int rc = 0;
  if (a->W_packed) memset(&tw, 0, sizeof(tw));
  else rc = encode_tmap_bf16_2d(&tw, a->W, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldw, SK_BM, SK_BK);
  if (rc != 0) return rc;
  rc = encode_tmap_bf16_2d(&tx, a->X, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->ldx, SK_MB, SK_BK);
  if (rc != 0) return rc;
  if (ext_z) rc = encode_tmap_bf16_2d(&tz, a->Z, (uint64_t)a->M, (uint64_t)a->Kext, (uint64_t)a->ldz, SK_MB, SK_BK);
  else tz = tx;
  if (rc != 0) return rc;
  SkinnyParams p;
  memset(&p, 0, sizeof(p));
  p.C = a->C; p.bias = a->bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  p.w_tiled = reinterpret_cast<const __nv_bfloat16*>(a->W_packed);
  p.M = a->M; p.N = a->N; p.K = a->K; p.ldc = a->ldc; p.ldr = a->ldr;
  p.act = a->act; p.out_dtype = a->out_dtype;
  p.splits = splits; p.kb_per_tile = kb;
  p.rows_per_rank = 2 * ((64 + splits - 1) / splits);  // even, so SwiGLU (gate, up) row pairs never straddle two ranks
  p.stats_w = reinterpret_cast<const __nv_bfloat16*>(a->stats_packed);
  p.zbuf = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(a->Z));
  p.rstd = a->rstd;
  p.flags = a->flags;
  p.n_tiles = tiles; p.kb_main = kb_main; p.ldz = a->ldz;
  p.has_stats = has_stats ? 1 : 0; p.norm = a->norm != 0; p.stats_linears = a->stats_linears; p.ext_from_z = ext_z ? 1 : 0;
  p.eps = a->eps; p.lora_scale = a->lora_scale;
  p.pf_ptr = reinterpret_cast<const uint8_t*>(a->prefetch);
  p.pf_bytes = (a->prefetch && a->prefetch_bytes > 0) ? (unsigned long long)a->prefetch_bytes : 0ull;
  // clusters that share the statistics item: ~8 k-blocks per CTA (its stream is latency-bound: x from L2 + 5 KB of router/A rows per
  // k-block behind the weight streams of 300 other CTAs), at most 8; more than one needs the scratch buffer
  int sc = 0;
  if (has_stats) {
    // the cross-cluster combine costs ~8 us of dependent global round trips: worth it only when one cluster would stream long
    // (0.4 us per k-block and CTA).  Qwen2 qkv (56 k-blocks over 4 ranks): one cluster; LLaMA gate/up (64 over 1): eight
    const int per_cta = kb_main / splits;
    sc = a->stats_clusters > 0 ? a->stats_clusters : ((a->stats_scratch && per_cta > 24) ? (per_cta + 7) / 8 : 1);
    if (sc > 8) sc = 8;
    if (sc * splits > kb_main) sc = kb_main / splits > 0 ? kb_main / splits : 1;
    CRAB_REQUIRE(sc == 1 || a->stats_scratch, "crab_gemm_skinny_bf16: stats_clusters > 1 needs stats_scratch (8 x 36 x 32 floats)");
  }
  p.stats_clusters = sc;
  p.stats_scratch = a->stats_scratch;
  p.flags_clear = has_stats ? a->flags_clear : nullptr;
  p.trace = next_trace_slot((tiles + sc) * splits);
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("CRAB_SKINNY_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((tiles + sc) * splits));
  cfg.blockDim = dim3(SK_THREADS);
  cfg.dynamicSmemBytes = SK_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (splits > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)splits;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_mask() & PDL_GEMM) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  CRAB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_skinny_tcgen05_kernel, tw, tx, tz, p));
  return CRAB_OK;
}
