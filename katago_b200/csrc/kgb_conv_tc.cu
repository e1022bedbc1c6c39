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
#include "kgb_conv.cuh"

#include <cstdlib>

namespace kgb {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;    // fp16 elements = one 128B swizzle row
static constexpr int UMMA_K = 16;
static constexpr int EPI_WARP0 = 4;            // warps 0-3: TMA, MMA, TMEM alloc, spare; epilogue warps follow
static constexpr int MAX_THREADS = 128 + 512;   // up to 16 epilogue warps (4 per TMEM lane quadrant)
static constexpr int MAX_STAGES = 8;
static constexpr int SMEM_LIMIT = 227 * 1024;

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

// ------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while(!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
    ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
    ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
    : "r"(taddr));
}
// The wait names the destination registers as in/out operands so no consumer can be scheduled above it.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
    : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
      "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
    :: "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=1024B: 8 rows x 128B)
//   | [46,48) version=1 | [61,64) layout_type=2 (SWIZZLE_128B)
// The start address may sit on ANY 128-byte row of a 1024B-aligned swizzled tile (the per-tap views of the A halo tile):
// measured on B200, the tensor core applies the 128B XOR swizzle to absolute shared-memory address bits [4,7)^[7,10) -
// exactly how TMA wrote the tile - so base_offset [49,52) must stay 0 (setting it to (start>>7)&7 gives wrong results;
// profiles/r01_descriptor_shift_experiment.md).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (InstrDescriptor): c_format F32 (bit 4), a/b F16 (0), K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
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
// Staged epilogue for one 16-column chunk of a warp's 32 accumulator rows.
// The accumulator arrives row-per-thread (tcgen05.ld 32x32b); writing global memory in that shape touches 32 different
// 128-byte lines with 16 bytes each per instruction.  Instead every tensor goes through a small per-warp shared-memory
// tile and is moved to / from global memory with lanes laid out along the rows: 64 B (fp32) or 32 B (fp16) contiguous per
// row per instruction, i.e. full sectors and 4x / 2x fewer LSU wavefronts.
//   S: fp32 tile [32 rows][16 cols], row stride 20 words;  T: fp16 tile [32 rows][16 cols], row stride 12 words.
// (strides chosen so that both the row-wise 16-byte accesses and the piece-wise ones are bank-conflict free per quarter warp)
// ------------------------------------------------------------------------------------------------------------
static constexpr int EPI_S_WORDS = 32 * 20;
static constexpr int EPI_T_WORDS = 32 * 12;
static constexpr int EPI_SMEM_PER_WARP = (EPI_S_WORDS + EPI_T_WORDS) * 4;   // 4 KB

__device__ __forceinline__ void tile_ld_f32(float* S, const float* g, int pitch, int rowsValid, int lane) {
#pragma unroll
  for(int k = 0; k < 4; k++) {
    int piece = k * 32 + lane, r = piece >> 2, part = piece & 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if(r < rowsValid) v = *reinterpret_cast<const float4*>(g + (size_t)r * pitch + part * 4);
    *reinterpret_cast<float4*>(S + r * 20 + part * 4) = v;
  }
}
__device__ __forceinline__ void tile_st_f32(const float* S, float* g, int pitch, int rowsValid, int lane) {
#pragma unroll
  for(int k = 0; k < 4; k++) {
    int piece = k * 32 + lane, r = piece >> 2, part = piece & 3;
    if(r < rowsValid) *reinterpret_cast<float4*>(g + (size_t)r * pitch + part * 4) = *reinterpret_cast<const float4*>(S + r * 20 + part * 4);
  }
}
__device__ __forceinline__ void tile_ld_f16(uint32_t* T, const __half* g, int pitch, int rowsValid, int lane) {
#pragma unroll
  for(int k = 0; k < 2; k++) {
    int piece = k * 32 + lane, r = piece >> 1, part = piece & 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if(r < rowsValid) v = *reinterpret_cast<const uint4*>(g + (size_t)r * pitch + part * 8);
    *reinterpret_cast<uint4*>(T + r * 12 + part * 4) = v;
  }
}
__device__ __forceinline__ void tile_st_f16(const uint32_t* T, __half* g, int pitch, int rowsValid, int lane) {
#pragma unroll
  for(int k = 0; k < 2; k++) {
    int piece = k * 32 + lane, r = piece >> 1, part = piece & 1;
    if(r < rowsValid) *reinterpret_cast<uint4*>(g + (size_t)r * pitch + part * 8) = *reinterpret_cast<const uint4*>(T + r * 12 + part * 4);
  }
}

// rowBase = first of the warp's 32 rows, rowsValid = how many of them are < M; col = first of the 16 columns.
__device__ __forceinline__ void epilogue_chunk_staged(const ConvParams& p, const uint32_t (&acc)[16], int rowBase, int rowsValid, int lane, int col,
                                                      float maskv, int img, const float* sc, const float* bi, float* S, uint32_t* T) {
  float v[16];
#pragma unroll
  for(int j = 0; j < 16; j++) v[j] = __uint_as_float(acc[j]);
  const bool valid = lane < rowsValid;
  if(p.ncbias != nullptr && valid) {
    const float4* b = reinterpret_cast<const float4*>(p.ncbias + (size_t)img * p.cout_p + col);
#pragma unroll
    for(int q = 0; q < 4; q++) {
      float4 t = __ldg(b + q);
      v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
    }
  }
  const size_t off = (size_t)rowBase * p.cout_p + col;
  if(p.residual != nullptr) {
    if(p.residual_fp32) {
      tile_ld_f32(S, reinterpret_cast<const float*>(p.residual) + off, p.cout_p, rowsValid, lane);
      __syncwarp();
#pragma unroll
      for(int q = 0; q < 4; q++) {
        float4 t = *reinterpret_cast<const float4*>(S + lane * 20 + q * 4);
        v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
      }
    }
    else {
      tile_ld_f16(T, reinterpret_cast<const __half*>(p.residual) + off, p.cout_p, rowsValid, lane);
      __syncwarp();
#pragma unroll
      for(int q = 0; q < 2; q++) {
        uint4 t = *reinterpret_cast<const uint4*>(T + lane * 12 + q * 4);
        const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
        for(int e = 0; e < 4; e++) {
          float2 f = __half22float2(h[e]);
          v[8 * q + 2 * e] += f.x; v[8 * q + 2 * e + 1] += f.y;
        }
      }
    }
    __syncwarp();
  }
  if(p.raw_out != nullptr) {
    if(p.raw_fp32) {
#pragma unroll
      for(int q = 0; q < 4; q++)
        *reinterpret_cast<float4*>(S + lane * 20 + q * 4) = make_float4(v[4 * q] * maskv, v[4 * q + 1] * maskv, v[4 * q + 2] * maskv, v[4 * q + 3] * maskv);
      __syncwarp();
      tile_st_f32(S, reinterpret_cast<float*>(p.raw_out) + off, p.cout_p, rowsValid, lane);
    }
    else {
#pragma unroll
      for(int q = 0; q < 2; q++) {
        uint4 t;
        __half2* h = reinterpret_cast<__half2*>(&t);
#pragma unroll
        for(int e = 0; e < 4; e++) h[e] = __floats2half2_rn(v[8 * q + 2 * e] * maskv, v[8 * q + 2 * e + 1] * maskv);
        *reinterpret_cast<uint4*>(T + lane * 12 + q * 4) = t;
      }
      __syncwarp();
      tile_st_f16(T, reinterpret_cast<__half*>(p.raw_out) + off, p.cout_p, rowsValid, lane);
    }
    __syncwarp();
  }
  if(p.act_out != nullptr) {
    float a[16];
#pragma unroll
    for(int q = 0; q < 4; q++) {
      float4 s = *reinterpret_cast<const float4*>(sc + q * 4), b = *reinterpret_cast<const float4*>(bi + q * 4);
      a[4 * q] = kgb_activate(fmaf(v[4 * q], s.x, b.x), p.act) * maskv;
      a[4 * q + 1] = kgb_activate(fmaf(v[4 * q + 1], s.y, b.y), p.act) * maskv;
      a[4 * q + 2] = kgb_activate(fmaf(v[4 * q + 2], s.z, b.z), p.act) * maskv;
      a[4 * q + 3] = kgb_activate(fmaf(v[4 * q + 3], s.w, b.w), p.act) * maskv;
    }
    if(maskv == 0.0f) {
#pragma unroll
      for(int j = 0; j < 16; j++) a[j] = 0.0f;  // guards NaN/inf garbage at pad rows
    }
    const int ldo = p.split ? 2 * p.cout_p : p.cout_p;
    __half* dst = p.act_out + (size_t)rowBase * ldo + col;
    uint4 hi[2], lo[2];
    __half2* hh = reinterpret_cast<__half2*>(hi);
    __half2* hl = reinterpret_cast<__half2*>(lo);
#pragma unroll
    for(int e = 0; e < 8; e++) {
      __half2 h = __floats2half2_rn(a[2 * e], a[2 * e + 1]);
      hh[e] = h;
      float2 hf = __half22float2(h);
      hl[e] = __floats2half2_rn(a[2 * e] - hf.x, a[2 * e + 1] - hf.y);
    }
    *reinterpret_cast<uint4*>(T + lane * 12) = hi[0];
    *reinterpret_cast<uint4*>(T + lane * 12 + 4) = hi[1];
    __syncwarp();
    tile_st_f16(T, dst, ldo, rowsValid, lane);
    __syncwarp();
    if(p.split) {
      *reinterpret_cast<uint4*>(T + lane * 12) = lo[0];
      *reinterpret_cast<uint4*>(T + lane * 12 + 4) = lo[1];
      __syncwarp();
      tile_st_f16(T, dst + p.cout_p, ldo, rowsValid, lane);
      __syncwarp();
    }
  }
}

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
              const uint64_t da = make_smem_desc(a_base + (uint32_t)(halo + dy * p.Wp + dx) * 128u);
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
