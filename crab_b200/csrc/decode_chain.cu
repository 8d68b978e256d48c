// Decode-step GEMM CHAIN: up to four dependent weight-streaming linears of one decoder layer in ONE persistent launch
//     o_proj (+residual)  ->  [RMSNorm] gate/up + SwiGLU  ->  down_proj (+residual)  ->  [RMSNorm] qkv of the next layer | lm_head
// for M <= 32 token rows (HF generate's per-token forward: models/modeling_llama.py:765-837 with the hyper-LoRA linears of
// peft_hyper/tuners/lora.py:338-369).  At M <= 32 every one of these is HBM-bound: the job is to keep ONE stream of weight
// bytes flowing at full bandwidth across the phase boundaries, instead of paying a ramp, a tail and three row kernels per
// linear (round-1 decode step: 8 launches per layer, 0.64 of the HBM roofline for the GEMM class).
//
// Structure
//   * persistent clusters of S CTAs (1 CTA / SM, ~220 KB smem).  A work item = one 128-row weight tile; the S ranks of the
//     cluster split its K range, accumulate D[batch (32 of 128 lanes), 128 weight rows] in TMEM (tcgen05.mma M128 N128 K16,
//     double-buffered accumulator) and reduce-scatter the partials through distributed shared memory (rank j owns 128/S
//     rows) with mbarriers instead of cluster barriers, so the pipeline never drains between items.
//     The weights are the B (N-side) operand on purpose: with the weights on the M side (swap-AB, N = 32 batch columns, as in
//     gemm_skinny.cu) every K=16 MMA re-reads its 128 x 32 B A slice from shared memory at ~32 B/clk and the issue loop, not
//     HBM, bounds the stream (measured: 0.31 us per 16 KB k-block = 52 GB/s per SM; profiles/r03_chain_trace.txt); as the B
//     operand the same 16 KB go through in 4 x 64 clk.  The 96 unused A rows are whatever follows the 4 KB X tile in shared
//     memory: their D lanes are never read.
//   * warp roles: warp 0 streams WEIGHTS (1-D cp.async.bulk of pre-swizzled 16 KB blocks, crab_pack_skinny_weight layout)
//     through a 7-stage ring and never waits for anything but a free slot — weights do not depend on activations, so the ring
//     keeps filling across phase boundaries; warp 6 loads the ACTIVATION tiles (TMA 2-D, 4 KB) into the same stages and is the
//     only producer that waits for dependencies; warp 1 issues the MMAs; warps 2-5 are the epilogue.
//   * phases are ordered by global counters: an epilogue bumps done[phase] after its outputs are globally visible, the
//     activation producer of the next phase spins on it (acquire) before its first TMA load.  All CTAs are co-resident
//     (grid = cudaOccupancyMaxActiveClusters x S), items are dealt round-robin continuing across phases, every cluster walks
//     the phases in the same order, so the waits cannot cycle.
//   * RMSNorm never materialises: gamma is folded into the packed weights at load time and rstd[b] scales the accumulator
//     column in the epilogue.  rstd and the hyper-LoRA pre-pass  t = x . [R;A]^T  (11 dots per wrapped linear) come from a
//     STATS ITEM, the first item of the phase: its B tile is [x rows (32) ; gamma*[R;A] rows (<= 33)], so one extra MMA chain
//     gives  diag(x x^T) = sum x^2  and all the dots; rank 0 of that cluster turns them into rstd and z' = scale * softmax(rstd
//     * logits)_i * u_j (un-normalised u: the epilogue's rstd multiplies the whole accumulator, z' included), writes them to
//     global memory and raises zflag[phase].  The K-EXTENSION k-blocks of every other item (B_0|B_1|B_2 columns of the packed
//     weight against z') are loaded after that flag; epilogues of normalised phases wait for it before reading rstd.
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace crab {

static constexpr int DC_BM = 128, DC_MB = 32, DC_BK = 64;
static constexpr int DC_MAX_STAGES = 10;   // ring depth is a launch parameter: as many 20 KB stages as fit next to the buffers below
static constexpr int DC_W_BYTES = DC_BM * DC_BK * 2;      // 16 KB
static constexpr int DC_X_BYTES = DC_MB * DC_BK * 2;      // 4 KB
static constexpr int DC_STAGE_BYTES = DC_W_BYTES + DC_X_BYTES;
static constexpr int DC_SROWS = 40;                       // router/A rows streamed per k-block of a stats item (33 used)
static constexpr int DC_S_BYTES = DC_SROWS * DC_BK * 2;   // 5 KB
static constexpr int DC_MAX_S = 8;
static constexpr int DC_SGROUPS = 11;                     // stats buffer per source rank: [11 groups of 4][batch row 32][4]: 40 dots, sum x^2, pad
static constexpr int DC_SBUF_PER_RANK = DC_SGROUPS * 32 * 16;
// partial buffer of a reduce-scatter owner: [source rank][R / 4 row groups][batch row 32][4 floats], R = 128 / S rows owned.
// One warp-wide 16-byte store (lane = batch row) covers 512 CONTIGUOUS bytes of the owner's shared memory: DSMEM moves whole
// 128-byte lines, a row-major [batch][row] layout sends 32 lines for the same 512 bytes (measured: 2.0 us -> per scatter).
__host__ __device__ constexpr int dc_part_bytes(int) { return DC_BM * 32 * 4; }
static constexpr int DC_THREADS = 7 * 32;
static constexpr int DC_BAR_BYTES = 256;
static constexpr int DC_SMEM_MAX = 227 * 1024;
static constexpr int DC_MAX_PHASES = 4;
// dynamic shared memory: [ring: stages x 20 KB][partials: 2 x ~18 KB][stats: S x 5.5 KB][barriers]
__host__ __device__ constexpr int dc_off_part(int stages) { return stages * DC_STAGE_BYTES; }
__host__ __device__ constexpr int dc_off_sbuf(int stages, int S) { return dc_off_part(stages) + 2 * dc_part_bytes(S); }
__host__ __device__ constexpr int dc_off_bar(int stages, int S) { return dc_off_sbuf(stages, S) + S * DC_SBUF_PER_RANK; }
__host__ __device__ constexpr int dc_smem(int stages, int S) { return dc_off_bar(stages, S) + DC_BAR_BYTES + 1024; }

struct alignas(64) DcPhase {
  CUtensorMap tmap_x;              // main activation [M, K]
  CUtensorMap tmap_z;              // K-extension activation (z' columns) [M, Kext]
  const __nv_bfloat16* w;          // packed [n_tiles][kb_total][128 x 64]
  const __nv_bfloat16* stats_w;    // packed [kb_main][40 x 64] or null
  void* out;
  const float* bias;
  const __nv_bfloat16* resid;
  float* rstd;                     // [32]
  __nv_bfloat16* zbuf;             // stats output [M, ldz]
  int N, n_tiles, kb_main, kb_total, ldc, ldr, ldz;
  int swiglu, out_f32, norm, stats_linears, has_stats;
  int first_cluster, expected_prev;
  float eps, lora_scale;
  int k_main;
};
struct DcParams {
  DcPhase ph[DC_MAX_PHASES];
  int n_phases, M;
  int stages;      // ring depth
  int lag;         // items by which the K-extension runs trail their item (env CRAB_CHAIN_LAG, default 2; 4 accumulators allow <= 2)
  int debug;       // timing experiments only (env CRAB_CHAIN_DEBUG): 1 skip publish (fence + counter), 2 skip X loads, 4 skip MMAs
  unsigned long long* trace;   // diagnostics (env CRAB_CHAIN_TRACE = device address): [cta][32 items][16] globaltimer stamps
  int* counters;   // slots of DC_CSTRIDE ints (one 128-byte line each, so pollers of one flag do not queue behind the atomics of
                   // another): [0..3] done, [4..7] zflag, [8] exit count; all zero on entry, left zero on exit
};
static constexpr int DC_CSTRIDE = 32;

__device__ __forceinline__ uint32_t dc_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t dc_nrank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void dc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dc_mapa(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void dc_st_remote_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void dc_st_remote_f1(uint32_t addr, float a) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ void dc_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a LOCAL mbarrier whose arrivals come from other CTAs of the cluster (acquire at cluster scope)
__device__ __forceinline__ void dc_mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint64_t t0 = 0;
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++spins & 0x3ff) == 0) {
      const uint64_t t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > CRAB_MBAR_TIMEOUT_NS) {
        printf("crab: decode-chain cluster barrier timeout block=%d thread=%d\n", (int)blockIdx.x, (int)threadIdx.x);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ int dc_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Bounded spin on a global counter / flag written by other CTAs of the grid.
__device__ __forceinline__ void dc_spin_ge(const int* p, int target) {
  if (dc_ld_acquire(p) >= target) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (dc_ld_acquire(p) < target) {
    __nanosleep(32);
    if ((++spins & 0xff) == 0 && globaltimer_ns() - t0 > CRAB_MBAR_TIMEOUT_NS) {
      printf("crab: decode-chain dependency timeout block=%d thread=%d target=%d have=%d\n", (int)blockIdx.x, (int)threadIdx.x, target,
             dc_ld_acquire(p));
      __trap();
    }
  }
}
__device__ __forceinline__ void dc_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float dc_ld_cg_f32(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void dc_arrive_remote_n(uint32_t cluster_addr) { dc_arrive_remote(cluster_addr); }
__device__ __forceinline__ uint2 dc_ld_cg_u2(const void* p) {
  uint2 v;
  asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void dc_named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One run of k-blocks of one item, issued back to back by this CTA.
struct DcSeg {
  int ph, tile, kb0, kb1;
  uint32_t it;                      // running item index of this CTA: accumulator it & 3, barrier parity (it >> 2) & 1
  bool stats, ext, first, last;     // first: starts the item's accumulation; last: the accumulator is complete after this run
};
// The static block sequence of one CTA, shared by the weight producer, the activation producer and the MMA issuer.
// Items are dealt to clusters round-robin, continuing across phases.  The rank that owns the K-extension k-blocks (the last
// one) runs them ONE ITEM LATE:  main(0) main(1) ext(0) main(2) ext(1) ... ext(m-1)  — the z' columns they need appear only
// after the phase's statistics item has gone through the whole MMA -> DSMEM -> softmax -> flag chain (4-7 us), and in item
// order that rank (and, through the reduce-scatter, its whole cluster) would sit idle for it at the start of every phase.
// Four TMEM accumulators keep the deferred item, the running one and the one being drained apart.
struct DcSeq {
  const DcParams* p;
  int rank, S, C, cluster_id;
  int ph, m, first, s, nseg;
  uint32_t it0;
  bool lag;
  __device__ void open_phase() {
    const DcPhase& P = p->ph[ph];
    const int n_items = P.n_tiles + P.has_stats;
    first = cluster_id - P.first_cluster;
    if (first < 0) first += C;
    m = first < n_items ? (n_items - 1 - first) / C + 1 : 0;
    lag = (rank == S - 1) && (P.kb_total > P.kb_main) && m > 0;
    nseg = lag ? 2 * m : m;
    s = 0;
  }
  __device__ void init(const DcParams* p_, int rank_, int S_, int C_, int cluster_id_) {
    p = p_; rank = rank_; S = S_; C = C_; cluster_id = cluster_id_;
    ph = 0; it0 = 0;
    open_phase();
  }
  __device__ bool next(DcSeg& g) {
    for (;;) {
      while (s >= nseg) {
        it0 += (uint32_t)m;
        if (++ph >= p->n_phases) return false;
        open_phase();
      }
      const DcPhase& P = p->ph[ph];
      // lagged order with LAG = 2:  M0 M1 M2 E0 M3 E1 M4 E2 ... E(m-1): slot s holds M(s) while s < min(m, LAG + 1); after
      // that main and extension runs alternate (extension first) until the mains run out, then the remaining extensions.
      int k;
      bool ext_part = false;
      if (!lag) {
        k = s;
      } else {
        const int lead = m < p->lag + 1 ? m : p->lag + 1;     // mains issued before the first extension
        if (s < lead) {
          k = s;
        } else {
          const int r = s - lead;                              // 0: E0, 1: M(lead), 2: E1, 3: M(lead+1), ...
          const int mains_left = m - lead;
          if (r < 2 * mains_left) {
            if (r & 1) k = lead + (r >> 1);
            else { k = r >> 1; ext_part = true; }
          } else {
            k = mains_left + (r - 2 * mains_left);
            ext_part = true;
          }
        }
      }
      ++s;
      const int item = first + k * C;
      g.ph = ph;
      g.it = it0 + (uint32_t)k;
      g.stats = P.has_stats && item == 0;
      g.tile = item - P.has_stats;
      if (g.stats) {
        if (ext_part) continue;                     // a statistics item has no K-extension
        g.kb0 = rank * P.kb_main / S; g.kb1 = (rank + 1) * P.kb_main / S;
        g.ext = false; g.first = true; g.last = true;
      } else if (!lag) {
        g.kb0 = rank * P.kb_total / S; g.kb1 = (rank + 1) * P.kb_total / S;
        g.ext = false; g.first = true; g.last = true;
      } else if (!ext_part) {
        g.kb0 = rank * P.kb_total / S; g.kb1 = P.kb_main;
        if (g.kb0 >= g.kb1) continue;               // tiny K: this rank's slice is all extension, its run below starts the item
        g.ext = false; g.first = true; g.last = false;
      } else {
        g.kb0 = P.kb_main; g.kb1 = P.kb_total;
        g.ext = true; g.first = (rank * P.kb_total / S >= P.kb_main); g.last = true;
      }
      return true;
    }
  }
};

// warps 0-3: epilogue (warp 0 owns TMEM lanes 0..31 = the batch rows and drains the accumulator), warp 4: weight producer,
// warp 5: MMA issuer + TMEM owner, warp 6: activation producer
__global__ void __launch_bounds__(DC_THREADS, 1) decode_chain_kernel(const __grid_constant__ DcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = (int)dc_nrank();
  const int NST = p.stages;
  const int OFF_PART = dc_off_part(NST), OFF_SBUF = dc_off_sbuf(NST, S), PART_BYTES = dc_part_bytes(S);
  const uint32_t bar_base = smem_base + dc_off_bar(NST, S);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (DC_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * DC_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * DC_MAX_STAGES + 4 + a); };
  auto pfull_bar = [&](int b) { return bar_base + 8u * (2 * DC_MAX_STAGES + 8 + b); };
  const uint32_t sfull_bar = bar_base + 8u * (2 * DC_MAX_STAGES + 10);
  const uint32_t tmem_slot = bar_base + 8u * (2 * DC_MAX_STAGES + 11);
  const int rank = (int)dc_rank();
  const int cluster_id = blockIdx.x / S;
  const int C = gridDim.x / S;
  const int R = DC_BM / S;    // tile rows owned by one rank in the reduce-scatter
  const int RG = R >> 2;      // 4-row groups per rank

  if (warp == 4 && lane == 0) {
    for (int i = 0; i < p.n_phases; ++i) { tma_prefetch_desc(&p.ph[i].tmap_x); tma_prefetch_desc(&p.ph[i].tmap_z); }
    for (int s = 0; s < NST; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 4; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 1); }
    for (int a = 0; a < 2; ++a) mbar_init(pfull_bar(a), (uint32_t)S);
    mbar_init(sfull_bar, (uint32_t)S);
    fence_barrier_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  dc_cluster_sync();   // every CTA's barriers exist before anybody signals across the cluster
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 4) {
    if (lane == 0) {
      // ===================== weight producer: never waits for activations =====================
      // The statistics item heads every phase's critical path and its stream is tiny: pull it into L2 now, so that its
      // ring refills do not queue behind the weight stream's HBM traffic when the phase opens.
      for (int ph = 1; ph < p.n_phases; ++ph) {
        const DcPhase& P = p.ph[ph];
        int f = cluster_id - P.first_cluster;
        if (f < 0) f += C;
        if (P.has_stats && P.stats_w && f == 0) {
          const int kb0 = rank * P.kb_main / S, kb1 = (rank + 1) * P.kb_main / S;
          for (int kb = kb0; kb < kb1; kb += 3)
            bulk_prefetch_l2(P.stats_w + (size_t)kb * (DC_SROWS * DC_BK), (uint32_t)((kb1 - kb < 3 ? kb1 - kb : 3) * DC_S_BYTES));
        }
      }
      DcSeq seq;
      seq.init(&p, rank, S, C, cluster_id);
      DcSeg g;
      uint32_t stage = 0, phase = 0;
      while (seq.next(g)) {
        const DcPhase& P = p.ph[g.ph];
        for (int kb = g.kb0; kb < g.kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t slot = smem_base + stage * DC_STAGE_BYTES;
          if (g.stats) {
            if (P.stats_w) {
              mbar_arrive_expect_tx(full_bar(stage), DC_S_BYTES);
              bulk_load_1d_hint(slot + DC_X_BYTES, P.stats_w + (size_t)kb * (DC_SROWS * DC_BK), DC_S_BYTES, full_bar(stage), kEvictFirst);
            } else {
              mbar_arrive(full_bar(stage));
            }
          } else {
            mbar_arrive_expect_tx(full_bar(stage), DC_W_BYTES);
            bulk_load_1d_hint(slot, P.w + ((size_t)g.tile * P.kb_total + kb) * (DC_BM * DC_BK), DC_W_BYTES, full_bar(stage), kEvictFirst);
          }
          if (++stage == (uint32_t)NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    if (lane == 0) {
      // ===================== activation producer: the only role that waits for other CTAs' results =====================
      DcSeq seq;
      seq.init(&p, rank, S, C, cluster_id);
      DcSeg g;
      uint32_t stage = 0, phase = 0;
      int dep_ph = 0, z_ph = -1;   // phases whose inputs / z' are known to be complete
      while (seq.next(g)) {
        const DcPhase& P = p.ph[g.ph];
        if (g.ph > dep_ph) {   // the previous phase's outputs (this phase's activations) are complete and visible
          dc_spin_ge(p.counters + DC_CSTRIDE * (g.ph - 1), P.expected_prev);
          dc_fence_proxy_async();
          dep_ph = g.ph;
        }
        if (g.ext && P.has_stats && z_ph < g.ph) {   // z' of this phase comes from its statistics item
          dc_spin_ge(p.counters + DC_CSTRIDE * (4 + g.ph), 1);
          dc_fence_proxy_async();
          z_ph = g.ph;
        }
        for (int kb = g.kb0; kb < g.kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t slot = smem_base + stage * DC_STAGE_BYTES;
          if (p.debug & 2) {
            mbar_arrive(full_bar(stage));
          } else {
            mbar_arrive_expect_tx(full_bar(stage), DC_X_BYTES);
            if (g.ext) tma_load_2d_hint(slot + DC_W_BYTES, &P.tmap_z, full_bar(stage), (kb - P.kb_main) * DC_BK, 0, kEvictLast);
            else tma_load_2d_hint(slot + DC_W_BYTES, &P.tmap_x, full_bar(stage), kb * DC_BK, 0, kEvictLast);
          }
          if (++stage == (uint32_t)NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // ===================== MMA issuer: D[batch lanes, 128 weight rows] += X_tile[128 (32 valid), 64] . W_tile[128, 64]^T ====
      constexpr uint32_t idesc = make_idesc_bf16_f32(DC_BM, DC_BM);
      constexpr uint32_t idesc_x = make_idesc_bf16_f32(DC_BM, 32);        // statistics: x x^T        -> columns 0..31
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(DC_BM, 48);        // statistics: x [R;A]^T   -> columns 32..79
      DcSeq seq;
      seq.init(&p, rank, S, C, cluster_id);
      DcSeg g;
      uint32_t stage = 0, phase = 0;
      while (seq.next(g)) {
        const uint32_t acc = g.it & 3;
        unsigned long long* tr = (p.trace && g.it < 32) ? p.trace + ((size_t)blockIdx.x * 32 + g.it) * 16 : nullptr;
        if (g.first) {
          if (tr) tr[8] = globaltimer_ns();
          mbar_wait(tempty_bar(acc), ((g.it >> 2) & 1) ^ 1);   // the epilogue has drained this accumulator
          if (tr) tr[9] = globaltimer_ns();
          tc_fence_after();
        }
        const uint32_t d_addr = tmem_base + acc * DC_BM;
        for (int kb = g.kb0; kb < g.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t slot = smem_base + stage * DC_STAGE_BYTES;
          const uint64_t da = make_sdesc_sw128(slot + DC_W_BYTES);   // activations: rows 0..31 valid
          const uint32_t fresh = (g.first && kb == g.kb0) ? 0u : 1u;
          if (!(p.debug & 4)) {
            if (g.stats) {
              const uint64_t dbs = make_sdesc_sw128(slot + DC_X_BYTES);   // gamma*[R;A] rows
#pragma unroll
              for (int k = 0; k < DC_BK / 16; ++k) {
                umma_bf16_ss(d_addr, da + 2u * k, da + 2u * k, idesc_x, fresh | (k > 0));
                umma_bf16_ss(d_addr + 32, da + 2u * k, dbs + 2u * k, idesc_s, fresh | (k > 0));
              }
            } else {
              const uint64_t db = make_sdesc_sw128(slot);
#pragma unroll
              for (int k = 0; k < DC_BK / 16; ++k) umma_bf16_ss(d_addr, da + 2u * k, db + 2u * k, idesc, fresh | (k > 0));
            }
          }
          umma_commit(empty_bar(stage));
          if (++stage == (uint32_t)NST) { stage = 0; phase ^= 1; }
        }
        if (g.last) {
          umma_commit(tfull_bar(acc));
          if (tr) { tr[10] = globaltimer_ns(); tr[12] = (unsigned long long)(g.kb1 - g.kb0); }
        }
      }
    }
  } else {
    // ===================== epilogue warps 0..3 =====================
    const int tt = threadIdx.x;                 // 0..127
    const int eb = lane;                        // batch row this thread finishes
    const int eq = warp;                        // row sub-group
    const int RP = R >> 2;                      // tile rows per thread in the reduce
    float* const part = reinterpret_cast<float*>(smem_gen + OFF_PART);
    float* const sbuf = reinterpret_cast<float*>(smem_gen + OFF_SBUF);
    uint32_t it = 0, nn = 0, ns = 0;
    for (int ph = 0; ph < p.n_phases; ++ph) {
      const DcPhase& P = p.ph[ph];
      const int n_items = P.n_tiles + P.has_stats;
      int first = cluster_id - P.first_cluster;
      if (first < 0) first += C;
      bool rstd_ok = false;
      float rstd = 1.0f;
      for (int item = first; item < n_items; item += C, ++it) {
        const bool is_stats = P.has_stats && item == 0;
        const int tile = item - P.has_stats;
        const uint32_t acc = it & 3;
        const uint32_t buf = nn & 1;
        unsigned long long* tr = (p.trace && it < 32 && tt == 0) ? p.trace + ((size_t)blockIdx.x * 32 + it) * 16 : nullptr;
        if (tr) { tr[0] = globaltimer_ns(); tr[7] = (unsigned long long)(ph * 1000 + item); }
        if (warp == 0) {
          // ---- drain the accumulator (TMEM lane = batch row) and scatter the partials to their owners ----
          mbar_wait(tfull_bar(acc), (it >> 2) & 1);
          if (tr) tr[1] = globaltimer_ns();
          tc_fence_after();
          const uint32_t t_addr = tmem_base + acc * DC_BM;
          uint32_t r[32];
          if (is_stats) {
            // columns 0..31: x x^T (the diagonal is sum x^2); columns 32..71: router/A dots.  Everything goes to rank 0.
            const uint32_t remote = dc_mapa(smem_base + OFF_SBUF + (uint32_t)(rank * DC_SBUF_PER_RANK + lane * 16), 0u);
            tmem_ld_32x32b_x32(t_addr, r);
            tmem_ld_wait();
            float d = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) d = (c == lane) ? __uint_as_float(r[c]) : d;
            tmem_ld_32x32b_x32(t_addr + 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 8; ++g)
              dc_st_remote_f4(remote + g * 512, __uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]),
                              __uint_as_float(r[4 * g + 3]));
            tmem_ld_32x32b_x32(t_addr + 64, r);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 2; ++g)
              dc_st_remote_f4(remote + (8 + g) * 512, __uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]),
                              __uint_as_float(r[4 * g + 3]));
            dc_st_remote_f4(remote + 10 * 512, d, 0.f, 0.f, 0.f);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(tempty_bar(acc)); dc_arrive_remote(dc_mapa(sfull_bar, 0u)); }
          } else {
            const uint32_t pbase = smem_base + OFF_PART + buf * PART_BYTES + (uint32_t)(rank * RG * 512 + lane * 16);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              tmem_ld_32x32b_x32(t_addr + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const int row0 = c * 32 + g * 4;
                const int dst = row0 / R;
                const uint32_t remote = dc_mapa(pbase + (uint32_t)(((row0 - dst * R) >> 2) * 512), (uint32_t)dst);
                dc_st_remote_f4(remote, __uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]),
                                __uint_as_float(r[4 * g + 3]));
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (lane < S) dc_arrive_remote(dc_mapa(pfull_bar(buf), (uint32_t)lane));
          }
          if (tr) tr[2] = globaltimer_ns();
        }
        if (is_stats) {
          const int L = P.stats_linears;
          const int nw = L > 1 ? L : 1;   // warps of rank 0 that finish the statistics
          if (rank == 0 && eq < nw) {
            dc_mbar_wait_cluster(sfull_bar, ns & 1);
            float ss = 0.f;
            auto sval = [&](int s_, int i) { return sbuf[((s_ * DC_SGROUPS + (i >> 2)) * 32 + eb) * 4 + (i & 3)]; };
            for (int s = 0; s < S; ++s) ss += sval(s, 40);
            const float rs = P.norm ? rsqrtf(ss / (float)P.k_main + P.eps) : 1.0f;
            if (eq == 0 && P.norm && eb < p.M) P.rstd[eb] = rs;
            if (eq < L && eb < p.M) {
              float t[11];
#pragma unroll
              for (int j = 0; j < 11; ++j) {
                float a = 0.f;
                for (int s = 0; s < S; ++s) a += sval(s, eq * 11 + j);
                t[j] = a;
              }
              const float l0 = t[0] * rs, l1 = t[1] * rs, l2 = t[2] * rs;
              const float mx = fmaxf(l0, fmaxf(l1, l2));
              const float e0 = __expf(l0 - mx), e1 = __expf(l1 - mx), e2 = __expf(l2 - mx);
              const float inv = P.lora_scale / (e0 + e1 + e2);
              const float rw[3] = {e0 * inv, e1 * inv, e2 * inv};
              __nv_bfloat16* zrow = P.zbuf + (size_t)eb * P.ldz + eq * 24;
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                uint4 v;
                v.x = pack_bf16x2(rw[i] * t[3], rw[i] * t[4]);
                v.y = pack_bf16x2(rw[i] * t[5], rw[i] * t[6]);
                v.z = pack_bf16x2(rw[i] * t[7], rw[i] * t[8]);
                v.w = pack_bf16x2(rw[i] * t[9], rw[i] * t[10]);
                *reinterpret_cast<uint4*>(zrow + i * 8) = v;
              }
            }
            __threadfence();
            dc_fence_proxy_async();
            dc_named_bar(2, 32 * nw);
            if (tt == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.counters + DC_CSTRIDE * (4 + ph)), "r"(1) : "memory");
          }
          // keep the four epilogue warps in the same item: a warp that ran a whole item ahead would wait on the NEXT use of
          // sfull / pfull with a parity that the not-yet-completed current use makes look satisfied
          dc_named_bar(1, 128);
          ++ns;
          continue;
        }
        if (P.norm && !rstd_ok) {   // rstd of this phase is published together with zflag; ONE thread per CTA polls
          if (tt == 0) dc_spin_ge(p.counters + DC_CSTRIDE * (4 + ph), 1);
          dc_named_bar(1, 128);
          rstd = (eb < p.M) ? dc_ld_cg_f32(P.rstd + eb) : 0.f;
          rstd_ok = true;
        }
        dc_mbar_wait_cluster(pfull_bar(buf), (nn >> 1) & 1);
        if (tr) tr[3] = globaltimer_ns();
        // ---- this rank's R rows: sum the S partials in rank order (deterministic), finish, store ----
        const float* pb = part + buf * (PART_BYTES / 4);
        const int rl0 = eq * RP;
        const int n0 = tile * DC_BM + rank * R + rl0;
        for (int c = 0; c < RP; c += 4) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int s = 0; s < S; ++s) {
            const float4 q = *reinterpret_cast<const float4*>(pb + ((s * RG + ((rl0 + c) >> 2)) * 32 + eb) * 4);
            a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
          }
          float v[4] = {a.x, a.y, a.z, a.w};
          if (P.norm) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] *= rstd;
          }
          const int n = n0 + c;
          if (eb >= p.M || n >= P.N) continue;
          if (P.swiglu) {
            // interleaved rows: 2i = gate_i, 2i+1 = up_i -> output column (tile*128 + row) / 2
            const float o0 = v[0] / (1.0f + __expf(-v[0])) * v[1];
            const float o1 = v[2] / (1.0f + __expf(-v[2])) * v[3];
            *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(P.out) + (size_t)eb * P.ldc + (n >> 1)) = pack_bf16x2(o0, o1);
            continue;
          }
          const bool vec = (n + 3 < P.N);
          if (P.bias) {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (n + e < P.N) v[e] += __ldg(P.bias + n + e);
          }
          if (P.resid) {
            const __nv_bfloat16* rr = P.resid + (size_t)eb * P.ldr + n;
            if (vec) {
              const uint2 q = dc_ld_cg_u2(rr);   // written by an earlier phase of THIS launch on another SM: L2, not L1
              v[0] += bf16lo(q.x); v[1] += bf16hi(q.x); v[2] += bf16lo(q.y); v[3] += bf16hi(q.y);
            } else {
              for (int e = 0; e < 4; ++e) if (n + e < P.N) v[e] += __bfloat162float(*reinterpret_cast<const volatile __nv_bfloat16*>(rr + e));
            }
          }
          if (P.out_f32) {
            float* o = reinterpret_cast<float*>(P.out) + (size_t)eb * P.ldc + n;
            if (vec) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            else for (int e = 0; e < 4; ++e) if (n + e < P.N) o[e] = v[e];
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(P.out) + (size_t)eb * P.ldc + n;
            if (vec) *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
            else for (int e = 0; e < 4; ++e) if (n + e < P.N) o[e] = __float2bfloat16_rn(v[e]);
          }
        }
        if (tr) tr[4] = globaltimer_ns();
        // publish (CUTLASS arrive_inc pattern): every epilogue thread's stores happen-before the barrier, ONE thread then makes
        // them visible at gpu scope (and to the async proxy: the next phase reads them with TMA) and bumps the phase counter;
        // the other 127 threads move on.  The barrier also orders this item's reads of the partial buffer before the drainer
        // warp's next scatter round can make a peer overwrite it two items later.
        dc_named_bar(1, 128);
        if (tt == 32 && !(p.debug & 1)) {
          __threadfence();
          dc_fence_proxy_async();
          atomicAdd(p.counters + DC_CSTRIDE * ph, 1);
        }
        if (tr) tr[5] = globaltimer_ns();
        ++nn;
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  dc_cluster_sync();   // nobody leaves while a peer may still store into its shared memory or signal its barriers
  if (warp == 5) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
  // self-cleaning counters: the last CTA to get here zeroes them for the next launch (every wait above is over by then)
  if (threadIdx.x == 0) {
    __threadfence();
    const int prev = atomicAdd(p.counters + DC_CSTRIDE * 8, 1);
    if (prev == (int)gridDim.x - 1) {
      for (int i = 0; i < 9; ++i) p.counters[DC_CSTRIDE * i] = 0;
      __threadfence();
    }
  }
}

static int dc_pick_stages(int S) {
  const char* e = getenv("CRAB_CHAIN_STAGES");   // read per call: tools sweep it
  const int env = e ? atoi(e) : 0;
  int st = DC_MAX_STAGES;
  while (st > 2 && dc_smem(st, S) > DC_SMEM_MAX) --st;
  if (env >= 2 && env < st) st = env;
  return st;
}

static int dc_max_clusters(int S, int smem) {
  static int cache[DC_MAX_S + 1] = {0};
  static int cache_smem[DC_MAX_S + 1] = {0};
  if (cache[S] > 0 && cache_smem[S] == smem) return cache[S];
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(S * 148));
  cfg.blockDim = dim3(DC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)S;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, decode_chain_kernel, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return 0;
  }
  cache[S] = n;
  cache_smem[S] = smem;
  return n;
}

}  // namespace crab

using namespace crab;

extern "C" int crab_decode_chain_stats_bytes(int K, int64_t* bytes) {
  CRAB_REQUIRE(K > 0 && K % DC_BK == 0 && bytes, "crab_decode_chain_stats_bytes: K must be a positive multiple of 64");
  *bytes = (int64_t)(K / DC_BK) * DC_S_BYTES;
  return CRAB_OK;
}

extern "C" int crab_decode_chain_max_clusters(int cluster, int* n) {
  CRAB_REQUIRE(n && (cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8), "crab_decode_chain_max_clusters: cluster must be 1, 2, 4 or 8");
  static DeviceOnce attr_once;
  if (first_on_device(attr_once)) {
    CRAB_CHECK_CUDA(cudaFuncSetAttribute(decode_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DC_SMEM_MAX));
  }
  *n = dc_max_clusters(cluster, dc_smem(dc_pick_stages(cluster), cluster));
  CRAB_REQUIRE(*n > 0, "crab_decode_chain_max_clusters: occupancy query failed for cluster size %d", cluster);
  return CRAB_OK;
}

extern "C" int crab_decode_chain(const crab_chain_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CRAB_REQUIRE(a && a->n_phases >= 1 && a->n_phases <= DC_MAX_PHASES, "crab_decode_chain: 1..4 phases");
  CRAB_REQUIRE(a->M > 0 && a->M <= DC_MB, "crab_decode_chain: M must be in 1..32 (got %d)", a->M);
  CRAB_REQUIRE(a->counters && ((uintptr_t)a->counters % 128 == 0), "crab_decode_chain: counters (288 zeroed ints, 128-byte aligned) required");
  const int S = a->cluster > 0 ? a->cluster : 4;
  CRAB_REQUIRE(S == 1 || S == 2 || S == 4 || S == 8, "crab_decode_chain: cluster must be 1, 2, 4 or 8");
  int maxc = 0;
  int rc = crab_decode_chain_max_clusters(S, &maxc);
  if (rc != 0) return rc;
  int C = maxc;
  if (a->max_clusters > 0 && a->max_clusters < C) C = a->max_clusters;

  DcParams p;
  memset(&p, 0, sizeof(p));
  p.stages = dc_pick_stages(S);
  {
    const char* le = getenv("CRAB_CHAIN_LAG");
    p.lag = le ? atoi(le) : 2;
    if (p.lag < 0) p.lag = 0;
    if (p.lag > 2) p.lag = 2;
  }
  {
    const char* e = getenv("CRAB_CHAIN_DEBUG");
    p.debug = e ? atoi(e) : 0;
    const char* t = getenv("CRAB_CHAIN_TRACE");
    p.trace = t ? reinterpret_cast<unsigned long long*>(strtoull(t, nullptr, 0)) : nullptr;
  }
  p.n_phases = a->n_phases;
  p.M = a->M;
  p.counters = a->counters;
  int base = 0, prev_writers = 0;
  for (int i = 0; i < a->n_phases; ++i) {
    const crab_chain_phase& s = a->phase[i];
    DcPhase& d = p.ph[i];
    CRAB_REQUIRE(s.X && s.W_packed && s.C, "crab_decode_chain: phase %d: null X / W_packed / C", i);
    CRAB_REQUIRE(s.K > 0 && s.K % DC_BK == 0 && s.ldx % 8 == 0 && s.ldx >= s.K && ((uintptr_t)s.X % 16 == 0),
                 "crab_decode_chain: phase %d: K=%d must be a multiple of 64, ldx=%d a multiple of 8, X 16-byte aligned", i, s.K, s.ldx);
    CRAB_REQUIRE(s.Kext >= 0 && s.Kext <= 128 && (s.Kext == 0 || (s.Z && s.ldz % 8 == 0 && s.ldz >= s.Kext && ((uintptr_t)s.Z % 16 == 0))),
                 "crab_decode_chain: phase %d: bad K-extension (Kext=%d ldz=%d)", i, s.Kext, s.ldz);
    CRAB_REQUIRE(s.N > 0 && (uintptr_t)s.W_packed % 128 == 0, "crab_decode_chain: phase %d: N / W_packed alignment", i);
    CRAB_REQUIRE(s.stats_linears >= 0 && s.stats_linears <= 3, "crab_decode_chain: phase %d: stats_linears in 0..3", i);
    CRAB_REQUIRE(s.stats_linears == 0 || (s.stats_packed && s.Kext >= 24 * s.stats_linears && ((uintptr_t)s.stats_packed % 128 == 0)),
                 "crab_decode_chain: phase %d: a LoRA pre-pass needs stats_packed and Kext >= 24 per linear", i);
    CRAB_REQUIRE(!s.norm || s.rstd, "crab_decode_chain: phase %d: norm needs an rstd scratch buffer (32 floats)", i);
    CRAB_REQUIRE(s.act == CRAB_ACT_NONE || s.act == CRAB_ACT_SWIGLU, "crab_decode_chain: phase %d: act must be NONE or SWIGLU", i);
    if (s.act == CRAB_ACT_SWIGLU)
      CRAB_REQUIRE(s.N % 128 == 0 && !s.bias && !s.residual && s.out_dtype == CRAB_BF16,
                   "crab_decode_chain: phase %d: SWIGLU needs N %% 128 == 0, bf16 out, no bias / residual", i);
    CRAB_REQUIRE(s.ldc % 4 == 0 && (s.residual == nullptr || s.ldr % 4 == 0), "crab_decode_chain: phase %d: ldc / ldr must be multiples of 4", i);
    d.kb_main = s.K / DC_BK;
    d.kb_total = d.kb_main + (s.Kext + DC_BK - 1) / DC_BK;
    CRAB_REQUIRE(d.kb_main >= S, "crab_decode_chain: phase %d: K too small for a %d-way split", i, S);
    rc = encode_tmap_bf16_2d(&d.tmap_x, s.X, (uint64_t)a->M, (uint64_t)s.K, (uint64_t)s.ldx, DC_MB, DC_BK);
    if (rc != 0) return rc;
    if (s.Kext > 0) rc = encode_tmap_bf16_2d(&d.tmap_z, s.Z, (uint64_t)a->M, (uint64_t)s.Kext, (uint64_t)s.ldz, DC_MB, DC_BK);
    else d.tmap_z = d.tmap_x;
    if (rc != 0) return rc;
    d.w = reinterpret_cast<const __nv_bfloat16*>(s.W_packed);
    d.stats_w = reinterpret_cast<const __nv_bfloat16*>(s.stats_packed);
    d.out = s.C;
    d.bias = s.bias;
    d.resid = reinterpret_cast<const __nv_bfloat16*>(s.residual);
    d.rstd = s.rstd;
    d.zbuf = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(s.Z));
    d.N = s.N;
    d.n_tiles = (s.N + DC_BM - 1) / DC_BM;
    d.ldc = s.ldc; d.ldr = s.ldr; d.ldz = s.ldz;
    d.swiglu = s.act == CRAB_ACT_SWIGLU;
    d.out_f32 = s.out_dtype == CRAB_F32;
    d.norm = s.norm != 0;
    d.stats_linears = s.stats_linears;
    d.has_stats = (s.norm != 0 || s.stats_linears > 0) ? 1 : 0;
    d.first_cluster = base;
    d.expected_prev = prev_writers;
    d.eps = s.eps;
    d.lora_scale = s.lora_scale;
    d.k_main = s.K;
    base = (base + d.n_tiles + d.has_stats) % C;
    prev_writers = d.n_tiles * S;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(C * S));
  cfg.blockDim = dim3(DC_THREADS);
  cfg.dynamicSmemBytes = dc_smem(p.stages, S);
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)S;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CRAB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decode_chain_kernel, p));
  return CRAB_OK;
}
