// CTA-pair (tcgen05 cta_group::2) variant of the implicit-GEMM convolution.
//
// Why: measured on B200 (profiles/r01_conv_pipeline_experiments.md) a single-CTA tcgen05.mma with both operands in shared
// memory runs at ~62 B/clk of operand fetch: 128x192x16 takes ~165 cycles instead of the 96-cycle floor, because every
// MMA re-reads the whole 192x16 weight slice.  In a CTA pair the two SMs execute ONE 256 x N x 16 MMA: each SM supplies its
// own 128 activation rows and only HALF of the weight tile (N/2 rows), so weight bytes per SM halve - in shared-memory
// reads and in L2->SMEM TMA traffic.
//
// Work split inside the cluster (2 CTAs, rank 0 = leader):
//   both ranks   TMA producer: own A halo tile (rows m0 + rank*128 - halo ...) and own half of each tap's weight tile,
//                all signalling the LEADER's "full" mbarriers (cp.async.bulk.tensor ... cta_group::2)
//   leader only  MMA issuer: tcgen05.mma.cta_group::2, UMMA 256 x n_tile x 16; tcgen05.commit multicasts the "empty" and
//                "accumulator full" arrivals to both CTAs
//   both ranks   epilogue warps drain their own TMEM half (128 rows), peers arrive remotely on the leader's
//                "accumulator empty" barrier
// Everything else (padded-row layout, halo reuse across taps, taps-per-stage, staged epilogue) is as in kgb_conv_tc.cu.
#include "kgb_conv_tc_common.cuh"

#include <cstdlib>

namespace kgb {

static constexpr int PAIR_M = 256;

// dynamic smem per CTA: [slack][2 x A halo tile][stages x tps x half weight tile][BarrierBlock2][bn][epilogue tiles]
static inline int aBufBytes2(int a_box_rows) { return (a_box_rows * BLOCK_K * 2 + 1023) / 1024 * 1024; }
int convTC2SmemBytes(int n_tile, int cout_p, int a_box_rows, int tps, int epi_warps, int* stagesOut) {
  int bStage = tps * (n_tile / 2) * BLOCK_K * 2;
  int fixed = 1024 + 2 * aBufBytes2(a_box_rows) + 512 + 8 * cout_p + epi_warps * 4096;
  int stages = (SMEM_LIMIT - fixed) / bStage;
  if(stages > MAX_STAGES) stages = MAX_STAGES;
  if(stagesOut) *stagesOut = stages;
  return stages * bStage + fixed;
}

// ---- cluster / cta_group::2 PTX -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA load whose completion bytes are credited to a barrier that may live in the peer CTA (cta_group::2)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
    "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
    ::"r"(dst), "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
    ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// commit: arrive (once) on the barrier at this shared-memory offset in BOTH CTAs of the pair when all prior MMAs are done
__device__ __forceinline__ void tcgen05_commit_2sm_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc2(int n) {  // M = 256 across the pair
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(PAIR_M >> 4) << 24);
}

struct __align__(8) BarrierBlock2 {
  uint64_t full[MAX_STAGES];    // leader: weight half-tiles of both CTAs landed (2 arrivals + tx bytes)
  uint64_t empty[MAX_STAGES];   // each CTA: stage consumed (multicast commit)
  uint64_t a_full[2];           // leader: A halo tiles of both CTAs landed
  uint64_t a_empty[2];          // each CTA
  uint64_t tmem_full[2];        // each CTA (multicast commit)
  uint64_t tmem_empty[2];       // leader: epilogue warps of both CTAs
  uint32_t tmem_base;
  uint32_t pad;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MAX_THREADS, 1)
kgb_conv_tc2_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                    const __grid_constant__ ConvParams p, int stages, int epi_per_quad, int tps, int num_pair_m_tiles, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int n_half = p.n_tile >> 1;
  const int b_tile_bytes = n_half * BLOCK_K * 2;              // this CTA's half of one tap's weight tile
  const int b_stage_bytes = tps * b_tile_bytes;
  const int halo = (p.ky / 2) * p.Wp + (p.kx / 2);
  const int a_box_rows = BLOCK_M + 2 * halo;
  const int a_tx_bytes = a_box_rows * BLOCK_K * 2;
  const int a_buf_bytes = (a_tx_bytes + 1023) / 1024 * 1024;
  const uint32_t smem_b = smem_base + 2 * a_buf_bytes;
  uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
  BarrierBlock2* bars = reinterpret_cast<BarrierBlock2*>(smem_aligned + 2 * (size_t)a_buf_bytes + (size_t)stages * b_stage_bytes);
  float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
  float* s_bias = s_scale + p.cout_p;
  uint8_t* s_epi = reinterpret_cast<uint8_t*>(s_bias + p.cout_p);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_tiles = num_pair_m_tiles * p.num_n_tiles;
  const int taps = p.ky * p.kx;
  const int kblocks = p.cin_p / BLOCK_K;
  const int parts = p.split ? 3 : 1;
  const int phases = kblocks * parts;
  const int tap_groups = (taps + tps - 1) / tps;
  const int epi_warps = 4 * epi_per_quad;

  if(warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapB) : "memory");
  }
  if(warp == 1 && lane == 0) {
    for(int s = 0; s < stages; s++) {
      mbar_init(smem_u32(&bars->full[s]), 2);        // one arrival per CTA's producer (+ tx bytes of both)
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for(int s = 0; s < 2; s++) {
      mbar_init(smem_u32(&bars->a_full[s]), 2);
      mbar_init(smem_u32(&bars->a_empty[s]), 1);
      mbar_init(smem_u32(&bars->tmem_full[s]), 1);
      mbar_init(smem_u32(&bars->tmem_empty[s]), 2 * epi_warps);   // epilogue warps of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if(p.act_out != nullptr) {
    for(int c = threadIdx.x; c < p.cout_p; c += blockDim.x) { s_scale[c] = p.bn_scale[c]; s_bias[c] = p.bn_bias[c]; }
  }
  cluster_sync_all();   // barriers of both CTAs are initialised before anyone signals across the pair
  if(warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if(warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if(lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int abuf = 0; uint32_t aphase = 0;
      for(int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = (tile / p.num_n_tiles) * PAIR_M + (int)rank * BLOCK_M;
        const int n0 = (tile % p.num_n_tiles) * p.n_tile + (int)rank * n_half;
        for(int ph = 0; ph < phases; ph++) {
          const int kb = ph / parts, part = ph - kb * parts;
          const int colA = kb * BLOCK_K + (part == 1 ? p.cin_p : 0);
          const int colB = kb * BLOCK_K + (part == 2 ? p.cin_p : 0);
          mbar_wait(smem_u32(&bars->a_empty[abuf]), aphase ^ 1);
          {
            const uint32_t afull_leader = mapa_rank(smem_u32(&bars->a_full[abuf]), 0);
            if(leader) mbar_arrive_expect_tx(smem_u32(&bars->a_full[abuf]), (uint32_t)(2 * a_tx_bytes));
            else mbar_arrive_cluster(afull_leader);
            tma_load_2d_2sm(smem_base + abuf * a_buf_bytes, &tmapA, afull_leader, colA, m0 - halo);
          }
          if(++abuf == 2) { abuf = 0; aphase ^= 1; }
          for(int tg = 0; tg < tap_groups; tg++) {
            const int nb = min(tps, taps - tg * tps);
            mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
            const uint32_t full_leader = mapa_rank(smem_u32(&bars->full[stage]), 0);
            if(leader) mbar_arrive_expect_tx(smem_u32(&bars->full[stage]), (uint32_t)(2 * nb * b_tile_bytes));
            else mbar_arrive_cluster(full_leader);
            for(int j = 0; j < nb; j++)
              tma_load_2d_2sm(smem_b + stage * b_stage_bytes + j * b_tile_bytes, &tmapB, full_leader, colB, (tg * tps + j) * p.cout_p + n0);
            if(++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  }
  else if(warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if(leader && lane == 0) {
      const uint32_t idesc = make_idesc2(p.n_tile);
      const int ry = p.ky / 2, rx = p.kx / 2;
      int stage = 0; uint32_t phase = 0;
      int abuf = 0; uint32_t aphase = 0;
      int acc_stage = 0; uint32_t acc_phase = 0;
      for(int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(smem_u32(&bars->tmem_empty[acc_stage]), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc_stage * p.n_tile;
        for(int ph = 0; ph < phases; ph++) {
          mbar_wait(smem_u32(&bars->a_full[abuf]), aphase);
          const uint32_t a_base = smem_base + abuf * a_buf_bytes;
          for(int tg = 0; tg < tap_groups; tg++) {
            const int nb = min(tps, taps - tg * tps);
            mbar_wait(smem_u32(&bars->full[stage]), phase);
            tcgen05_fence_after();
            for(int j = 0; j < nb; j++) {
              const int tap = tg * tps + j;
              const int dy = tap / p.kx - ry, dx = tap - (tap / p.kx) * p.kx - rx;
              uint32_t a_off = (uint32_t)(halo + dy * p.Wp + dx) * 128u;
              if(dbg & 16) a_off &= ~1023u;   // timing experiment: 8-row aligned tap views (wrong results)
              const uint64_t da = make_smem_desc(a_base + a_off);
              const uint64_t db = make_smem_desc(smem_b + stage * b_stage_bytes + j * b_tile_bytes);
#pragma unroll
              for(int k = 0; k < BLOCK_K / UMMA_K; k++)
                if(!(dbg & 2)) umma_f16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (ph > 0 || tap > 0 || k > 0) ? 1u : 0u);
            }
            tcgen05_commit_2sm_mc(smem_u32(&bars->empty[stage]));
            if(++stage == stages) { stage = 0; phase ^= 1; }
          }
          tcgen05_commit_2sm_mc(smem_u32(&bars->a_empty[abuf]));
          if(++abuf == 2) { abuf = 0; aphase ^= 1; }
        }
        tcgen05_commit_2sm_mc(smem_u32(&bars->tmem_full[acc_stage]));
        if(++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  }
  else if(warp >= EPI_WARP0) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int quad = warp & 3;
    const int part = (warp - EPI_WARP0) >> 2;
    const int cols_per_part = p.n_tile / epi_per_quad;
    const int nchunks = cols_per_part >> 4;
    int acc_stage = 0; uint32_t acc_phase = 0;
    for(int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m0 = (tile / p.num_n_tiles) * PAIR_M + (int)rank * BLOCK_M;
      const int n0 = (tile % p.num_n_tiles) * p.n_tile + part * cols_per_part;
      const int rowBase = m0 + quad * 32;
      const int row = rowBase + lane;
      const int rowsValid = min(32, p.M - rowBase);
      const bool valid = row < p.M;
      const float maskv = valid ? __ldg(p.mask + row) : 0.0f;
      const int img = valid ? row / p.P : 0;
      float* S = reinterpret_cast<float*>(s_epi + (size_t)(warp - EPI_WARP0) * EPI_SMEM_PER_WARP);
      uint32_t* T = reinterpret_cast<uint32_t*>(S + EPI_S_WORDS);
      mbar_wait(smem_u32(&bars->tmem_full[acc_stage]), acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc_stage * p.n_tile + part * cols_per_part;
      uint32_t accA[16], accB[16];
      if(!(dbg & 4)) tmem_ld16(taddr, accA);
      for(int c = 0; c < ((dbg & 4) ? 0 : nchunks); c += 2) {
        if(c + 1 < nchunks) tmem_ld16(taddr + (c + 1) * 16, accB);
        tmem_ld_wait(accA);
        if(rowsValid > 0)
          epilogue_chunk_staged(p, accA, rowBase, rowsValid, lane, n0 + c * 16, maskv, img, s_scale + n0 + c * 16, s_bias + n0 + c * 16, S, T);
        if(c + 1 < nchunks) {
          if(c + 2 < nchunks) tmem_ld16(taddr + (c + 2) * 16, accA);
          tmem_ld_wait(accB);
          if(rowsValid > 0)
            epilogue_chunk_staged(p, accB, rowBase, rowsValid, lane, n0 + (c + 1) * 16, maskv, img, s_scale + n0 + (c + 1) * 16,
                                  s_bias + n0 + (c + 1) * 16, S, T);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if(lane == 0) {
        if(leader) mbar_arrive(smem_u32(&bars->tmem_empty[acc_stage]));
        else mbar_arrive_cluster(mapa_rank(smem_u32(&bars->tmem_empty[acc_stage]), 0));
      }
      if(++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  // Neither CTA may exit (or free TMEM) while the pair's MMAs can still read its shared memory / write its TMEM.
  tcgen05_fence_before();
  cluster_sync_all();
  if(warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

cudaError_t convTC2Init() {
  return cudaFuncSetAttribute(kgb_conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
}

// Returns cudaErrorNotSupported when the shape does not fit the pair kernel (caller falls back to the 1-CTA kernel).
cudaError_t launchConvTC2(const CUtensorMap& tmapA, const CUtensorMap& tmapBhalf, const ConvParams& p, int numSMs, cudaStream_t stream) {
  static int envTps = -1, envEpi = 0, dbg = 0;
  if(envTps < 0) {
    const char* e = getenv("KGB_CONV_DBG"); dbg = e ? atoi(e) : 0;
    e = getenv("KGB_CONV_TPS"); envTps = e ? atoi(e) : 0;
    e = getenv("KGB_CONV_EPI"); envEpi = e ? atoi(e) : 0;
  }
  const int taps = p.ky * p.kx;
  const int halo = (p.ky / 2) * p.Wp + (p.kx / 2);
  if(BLOCK_M + 2 * halo > 256 || (p.n_tile % 32) != 0) return cudaErrorNotSupported;
  int epi_per_quad = (p.n_tile % 64 == 0) ? 4 : 2;
  if(envEpi == 2 || envEpi == 4) epi_per_quad = (p.n_tile % (16 * envEpi) == 0) ? envEpi : epi_per_quad;
  int tps = 1, stages = 0, smem = 0;
  for(int t = (envTps > 0 ? envTps : 3); t >= 1; t--) {
    if(t > taps) continue;
    smem = convTC2SmemBytes(p.n_tile, p.cout_p, BLOCK_M + 2 * halo, t, 4 * epi_per_quad, &stages);
    if(stages >= 3 || (t == 1 && stages >= 2)) { tps = t; break; }
  }
  if(stages < 2) return cudaErrorNotSupported;
  const int num_pair_m_tiles = (p.M + PAIR_M - 1) / PAIR_M;
  const int tiles = num_pair_m_tiles * p.num_n_tiles;
  int clusters = numSMs / 2;
  if(tiles < clusters) clusters = tiles;
  const int threads = 128 + 128 * epi_per_quad;
  kgb_conv_tc2_kernel<<<2 * clusters, threads, smem, stream>>>(tmapA, tmapBhalf, p, stages, epi_per_quad, tps, num_pair_m_tiles, dbg);
  return cudaGetLastError();
}

}  // namespace kgb
