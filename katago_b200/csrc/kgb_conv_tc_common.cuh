// Shared pieces of the tcgen05 convolution kernels (1-CTA: kgb_conv_tc.cu, CTA-pair: kgb_conv_tc2.cu): tile constants,
// PTX wrappers, UMMA descriptors and the staged (coalesced) epilogue.
#pragma once
#include "kgb_conv.cuh"

namespace kgb {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;    // fp16 elements = one 128B swizzle row
static constexpr int UMMA_K = 16;
static constexpr int EPI_WARP0 = 4;            // warps 0-3: TMA, MMA, TMEM alloc, spare; epilogue warps follow
static constexpr int MAX_THREADS = 128 + 512;   // up to 16 epilogue warps (4 per TMEM lane quadrant)
static constexpr int MAX_STAGES = 8;
static constexpr int SMEM_LIMIT = 227 * 1024;


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
    // The epilogue is instruction-bound (ncu: ~30 issued instructions per output element before this form): the activation kind
    // is decided once per chunk, mish is x - 2x / (e(e+2) + 2) (same function as x tanh(softplus x); overflow of e gives 1/inf = 0
    // -> x, underflow gives x - x = 0, so no clamps), and the row mask is a select instead of a multiply.
    float a[16];
#pragma unroll
    for(int q = 0; q < 4; q++) {
      float4 s = *reinterpret_cast<const float4*>(sc + q * 4), b = *reinterpret_cast<const float4*>(bi + q * 4);
      a[4 * q] = fmaf(v[4 * q], s.x, b.x); a[4 * q + 1] = fmaf(v[4 * q + 1], s.y, b.y);
      a[4 * q + 2] = fmaf(v[4 * q + 2], s.z, b.z); a[4 * q + 3] = fmaf(v[4 * q + 3], s.w, b.w);
    }
    if(p.act == 2) {
#pragma unroll
      for(int j = 0; j < 16; j++) {
        const float x = a[j];
        const float e = kgb_ex2(x * 1.4426950408889634f);
        const float r = kgb_rcp(fmaf(e, e + 2.0f, 2.0f));
        a[j] = fmaf(-2.0f, x * r, x);
      }
    }
    else if(p.act == 1) {
#pragma unroll
      for(int j = 0; j < 16; j++) a[j] = fmaxf(a[j], 0.0f);
    }
    else if(p.act == 3) {
#pragma unroll
      for(int j = 0; j < 16; j++) a[j] = a[j] * kgb_rcp(1.0f + kgb_ex2(a[j] * -1.4426950408889634f));
    }
    if(maskv == 0.0f) {
#pragma unroll
      for(int j = 0; j < 16; j++) a[j] = 0.0f;  // pad and off-board rows (also guards NaN/inf garbage there)
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
      if(p.split) {
        float2 hf = __half22float2(h);
        hl[e] = __floats2half2_rn(a[2 * e] - hf.x, a[2 * e + 1] - hf.y);
      }
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


}  // namespace kgb
