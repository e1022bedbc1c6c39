// tcgen05 / TMA implicit-GEMM convolution for sm_100a: the trunk hot op of the NN evaluator
// (reference call sites: ConvLayer::apply eigenbackend.cpp:448-701,
// cudnnConvolutionForward / cublasHgemm cudabackend.cpp:788-841, BN+act+mask cudahelpers.cu:1370-2101).
//
// GEMM view:  D[M = batch*P rows, N = cout]  =  sum over taps t, k-blocks kb of  A_t[M, 64] * W_t[64, N]
//   A_t = the activation matrix shifted by the tap's row offset (see kgb_conv.cuh "padded rows").  All taps of one k-block
//         read the SAME rows of A, only shifted: the producer loads ONE halo tile {64 channels, 128 + 2*halo rows}
//         (halo = ry*(X+pad)+rx, 128B swizzle, out-of-range rows zero-filled by TMA) and the MMA issuer addresses tap t
//         by starting its shared-memory descriptor `halo + dy*(X+pad)+dx` rows (128 B each) into that tile.  A is fetched from L2 once instead of 9 times.
//   W   = packed [tap][cout_p][cin_p] fp16 (K-major), box {64, n_tile}, one TMA box per (k-block, tap).
//   D   = fp32 accumulator in TMEM, 2 stages x n_tile columns, so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Persistent, warp-specialised CTA (1 per SM, 384 threads):
//   warp 0     TMA producer            (one elected lane)
//   warp 1     tcgen05.mma issuer      (one elected lane), UMMA 128 x n_tile x 16, kind::f16, fp32 accumulate
//   warp 2     TMEM allocator
//   warps 4-11 epilogue: tcgen05.ld -> (+ncbias, +residual) -> raw store, BN+act+mask -> fp16 store (next layer's A)
// Pipelines: smem full/empty mbarriers per stage (TMA <-> MMA), tmem full/empty mbarriers per accumulator stage.
//
// "split" mode computes a*w ~= ah*wh + al*wh + ah*wl with a = ah + al, w = wh + wl in fp16 (fp32 accumulate):
// fp32-grade accuracy on the fp16 tensor pipe at 3x the MMA count (the backend's useFP16=false mode).
#include "kgb_conv_tc_common.cuh"

#include <cstdlib>

namespace kgb {

// dynamic smem: [<=1023 B slack][2 x A halo tile][stages x B tile][BarrierBlock, 512 B][bn scale | bn bias: 2 x cout_p fp32]
//               [epilogue warps x 4 KB staging tiles]
static inline int aBufBytes(int a_box_rows) { return (a_box_rows * BLOCK_K * 2 + 1023) / 1024 * 1024; }
// One pipeline stage carries the weight tiles of `tps` consecutive taps (same k-block): fewer, fatter barrier round trips
// per MMA (measured: ~400 cycles of producer<->issuer handshake per stage, as long as the 4 MMAs of one tap).
int convTCSmemBytes(int n_tile, int cout_p, int a_box_rows, int tps, int epi_warps, int* stagesOut) {
  int bStage = tps * n_tile * BLOCK_K * 2;
  int fixed = 1024 + 2 * aBufBytes(a_box_rows) + 512 + 8 * cout_p + epi_warps * 4096;
  int stages = (SMEM_LIMIT - fixed) / bStage;
  if(stages > MAX_STAGES) stages = MAX_STAGES;
  if(stagesOut) *stagesOut = stages;
  return stages * bStage + fixed;
}

struct __align__(8) BarrierBlock {
  uint64_t full[MAX_STAGES];    // B tile landed
  uint64_t empty[MAX_STAGES];   // B tile consumed
  uint64_t a_full[2];           // A halo tile landed
  uint64_t a_empty[2];          // A halo tile consumed by all of its taps
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// ------------------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MAX_THREADS, 1)
kgb_conv_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                   const __grid_constant__ ConvParams p, int stages, int epi_per_quad, int tps, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_tile_bytes = p.n_tile * BLOCK_K * 2;             // one tap's weight tile
  const int b_stage_bytes = tps * b_tile_bytes;                // a stage = `tps` taps of one k-block
  const int halo = (p.ky / 2) * p.Wp + (p.kx / 2);          // rows of A above / below the tile that the taps reach
  const int a_box_rows = BLOCK_M + 2 * halo;
  const int a_tx_bytes = a_box_rows * BLOCK_K * 2;
  const int a_buf_bytes = (a_tx_bytes + 1023) / 1024 * 1024;
  const uint32_t smem_b = smem_base + 2 * a_buf_bytes;
  uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
  BarrierBlock* bars = reinterpret_cast<BarrierBlock*>(smem_aligned + 2 * (size_t)a_buf_bytes + (size_t)stages * b_stage_bytes);
  float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
  float* s_bias = s_scale + p.cout_p;
  uint8_t* s_epi = reinterpret_cast<uint8_t*>(s_bias + p.cout_p);   // 16-byte aligned: cout_p is a multiple of 64

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int taps = p.ky * p.kx;
  const int kblocks = p.cin_p / BLOCK_K;
  const int parts = p.split ? 3 : 1;
  const int phases = kblocks * parts;                         // one A halo tile per (k-block, split part)
  const int tap_groups = (taps + tps - 1) / tps;

  if(warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapB) : "memory");
  }
  if(warp == 1 && lane == 0) {
    for(int s = 0; s < stages; s++) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for(int s = 0; s < 2; s++) {
      mbar_init(smem_u32(&bars->a_full[s]), 1);
      mbar_init(smem_u32(&bars->a_empty[s]), 1);
      mbar_init(smem_u32(&bars->tmem_full[s]), 1);
      mbar_init(smem_u32(&bars->tmem_empty[s]), 4 * epi_per_quad);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if(warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if(p.act_out != nullptr) {
    for(int c = threadIdx.x; c < p.cout_p; c += blockDim.x) { s_scale[c] = p.bn_scale[c]; s_bias[c] = p.bn_bias[c]; }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if(warp == 0) {
    // ===================== TMA producer =====================
    if(lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int abuf = 0; uint32_t aphase = 0;
      for(int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.num_n_tiles) * BLOCK_M;
        const int n0 = (tile % p.num_n_tiles) * p.n_tile;
        for(int ph = 0; ph < phases; ph++) {
          const int kb = ph / parts, part = ph - kb * parts;
          const int colA = kb * BLOCK_K + (part == 1 ? p.cin_p : 0);
          const int colB = kb * BLOCK_K + (part == 2 ? p.cin_p : 0);
          mbar_wait(smem_u32(&bars->a_empty[abuf]), aphase ^ 1);
          const uint32_t afull = smem_u32(&bars->a_full[abuf]);
          if(dbg & 1) mbar_arrive(afull);
          else {
            mbar_arrive_expect_tx(afull, (uint32_t)a_tx_bytes);
            tma_load_2d(smem_base + abuf * a_buf_bytes, &tmapA, afull, colA, m0 - halo);
          }
          if(++abuf == 2) { abuf = 0; aphase ^= 1; }
          for(int tg = 0; tg < tap_groups; tg++) {
            const int nb = min(tps, taps - tg * tps);
            mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
            const uint32_t full = smem_u32(&bars->full[stage]);
            if(dbg & 1) mbar_arrive(full);
            else {
              mbar_arrive_expect_tx(full, (uint32_t)(nb * b_tile_bytes));
              for(int j = 0; j < nb; j++)
                tma_load_2d(smem_b + stage * b_stage_bytes + j * b_tile_bytes, &tmapB, full, colB, (tg * tps + j) * p.cout_p + n0);
            }
            if(++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  }
  else if(warp == 1) {
    // ===================== MMA issuer =====================
    if(lane == 0) {
      const uint32_t idesc = make_idesc(p.n_tile);
      const int ry = p.ky / 2, rx = p.kx / 2;
      int stage = 0; uint32_t phase = 0;
      int abuf = 0; uint32_t aphase = 0;
      int acc_stage = 0; uint32_t acc_phase = 0;
      for(int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
              // tap (dy,dx) = the 128 rows starting `halo + dy*Wp + dx` rows into the halo tile (128 B per row)
              uint32_t a_off = (uint32_t)(halo + dy * p.Wp + dx) * 128u;
              if(dbg & 16) a_off &= ~1023u;   // timing experiment: 8-row aligned tap views (wrong results)
              const uint64_t da = make_smem_desc(a_base + a_off);
              const uint64_t db = make_smem_desc(smem_b + stage * b_stage_bytes + j * b_tile_bytes);
#pragma unroll
              for(int k = 0; k < BLOCK_K / UMMA_K; k++) {
                // advance 32 bytes (16 fp16) along K inside the 128B swizzle row: +2 in 16-byte descriptor units
                if(!(dbg & 2)) umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (ph > 0 || tap > 0 || k > 0) ? 1u : 0u);
              }
            }
            tcgen05_commit(smem_u32(&bars->empty[stage]));
            if(++stage == stages) { stage = 0; phase ^= 1; }
          }
          tcgen05_commit(smem_u32(&bars->a_empty[abuf]));
          if(++abuf == 2) { abuf = 0; aphase ^= 1; }
        }
        tcgen05_commit(smem_u32(&bars->tmem_full[acc_stage]));
        if(++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  }
  else if(warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int quad = warp & 3;                          // TMEM lane quadrant this warp may access
    const int part = (warp - EPI_WARP0) >> 2;           // which slice of the tile's columns
    const int cols_per_part = p.n_tile / epi_per_quad;  // multiple of 16
    const int nchunks = cols_per_part >> 4;
    int acc_stage = 0; uint32_t acc_phase = 0;
    for(int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.num_n_tiles) * BLOCK_M;
      const int n0 = (tile % p.num_n_tiles) * p.n_tile + part * cols_per_part;
      const int rowBase = m0 + quad * 32;
      const int row = rowBase + lane;
      const int rowsValid = min(32, p.M - rowBase);          // may be <= 0 in the last tile
      const bool valid = row < p.M;
      const float maskv = valid ? __ldg(p.mask + row) : 0.0f;
      const int img = valid ? row / p.P : 0;
      float* S = reinterpret_cast<float*>(s_epi + (size_t)(warp - EPI_WARP0) * EPI_SMEM_PER_WARP);
      uint32_t* T = reinterpret_cast<uint32_t*>(S + EPI_S_WORDS);
      mbar_wait(smem_u32(&bars->tmem_full[acc_stage]), acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc_stage * p.n_tile + part * cols_per_part;
      // software pipeline: the TMEM load of chunk c+1 is in flight while chunk c goes through the epilogue
      uint32_t accA[16], accB[16];
      if(!(dbg & 4)) tmem_ld16(taddr, accA);
      for(int c = 0; c < ((dbg & 4) ? 0 : nchunks); c += 2) {
        if(c + 1 < nchunks) tmem_ld16(taddr + (c + 1) * 16, accB);
        tmem_ld_wait(accA);
        if(rowsValid > 0 && !(dbg & 8))
          epilogue_chunk_staged(p, accA, rowBase, rowsValid, lane, n0 + c * 16, maskv, img, s_scale + n0 + c * 16, s_bias + n0 + c * 16, S, T);
        if(c + 1 < nchunks) {
          if(c + 2 < nchunks) tmem_ld16(taddr + (c + 2) * 16, accA);
          tmem_ld_wait(accB);
          if(rowsValid > 0 && !(dbg & 8))
            epilogue_chunk_staged(p, accB, rowBase, rowsValid, lane, n0 + (c + 1) * 16, maskv, img, s_scale + n0 + (c + 1) * 16,
                                  s_bias + n0 + (c + 1) * 16, S, T);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if(lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[acc_stage]));
      if(++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if(warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

cudaError_t convTCInit() {
  return cudaFuncSetAttribute(kgb_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
}

cudaError_t launchConvTC(const CUtensorMap& tmapA, const CUtensorMap& tmapB, const ConvParams& p, int numSMs, cudaStream_t stream) {
  static int dbg = -1, maxStages = 0, envTps = 0, envEpi = 0;
  if(dbg < 0) {  // bring-up knobs (timing experiments only): KGB_CONV_DBG bit0 = no TMA traffic, bit1 = no MMA, bit2 = no epilogue,
                 // bit3 = no epilogue stores; KGB_CONV_STAGES caps the ring; KGB_CONV_TPS / KGB_CONV_EPI override the tiling choice
    const char* e = getenv("KGB_CONV_DBG"); dbg = e ? atoi(e) : 0;
    e = getenv("KGB_CONV_STAGES"); maxStages = e ? atoi(e) : 0;
    e = getenv("KGB_CONV_TPS"); envTps = e ? atoi(e) : 0;
    e = getenv("KGB_CONV_EPI"); envEpi = e ? atoi(e) : 0;
  }
  const int taps = p.ky * p.kx;
  const int halo = (p.ky / 2) * p.Wp + (p.kx / 2);
  if(BLOCK_M + 2 * halo > 256) return cudaErrorInvalidValue;  // TMA box limit
  // 4 epilogue warps per TMEM lane quadrant when the tile's columns split evenly into 16-column chunks, else 2
  int epi_per_quad = (p.n_tile % 64 == 0) ? 4 : 2;
  if(envEpi == 2 || envEpi == 4) epi_per_quad = (p.n_tile % (16 * envEpi) == 0) ? envEpi : epi_per_quad;
  // taps per stage: as many as still leave a 2-deep ring (3x3: 2 taps per stage with 16 epilogue warps at n_tile 192)
  int tps = 1, stages = 0, smem = 0;
  for(int t = (envTps > 0 ? envTps : 3); t >= 1; t--) {
    if(t > taps) continue;
    smem = convTCSmemBytes(p.n_tile, p.cout_p, BLOCK_M + 2 * halo, t, 4 * epi_per_quad, &stages);
    if(stages >= 2) { tps = t; break; }
  }
  if(stages < 2) return cudaErrorInvalidValue;
  int threads = 128 + 128 * epi_per_quad;
  if(maxStages > 0 && stages > maxStages) stages = maxStages;
  int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = tiles < numSMs ? tiles : numSMs;
  kgb_conv_tc_kernel<<<grid, threads, smem, stream>>>(tmapA, tmapB, p, stages, epi_per_quad, tps, dbg);
  return cudaGetLastError();
}

}  // namespace kgb
