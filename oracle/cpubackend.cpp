// oracle/cpubackend.cpp - TEST / BASELINE INFRASTRUCTURE, never linked into the product.
//
// A CPU implementation of the reference's backend boundary (namespace NeuralNet, cpp/neuralnet/nninterface.h:32-182) so that the
// UNMODIFIED reference (board, search, NNEvaluator, self-play) runs entirely on host cores: the CPU baseline of bench.py
// (`cpu_baseline.kind = "restated Eigen path (C++), full selfplay"`, `--impl reference`).  The reference's own CPU backend
// (neuralnet/eigenbackend.cpp) cannot be compiled here - it needs Eigen3, which is neither vendored under /root/reference nor
// installed (CMakeLists.txt:516) - so its algorithm is restated: NHWC fp32 tensors, 3x3 convolutions as Winograd F(4x4,3x3) with one
// GEMM per transform point (eigenbackend.cpp:335-412,448-701), 1x1 / other convolutions and matmuls as GEMMs (:811-862), BN+activation
// +mask, global pooling (:124-197), residual / gpool / nested-bottleneck blocks (:1038-1315), policy and value heads (:1723-2216),
// getOutput with input / output symmetries and policy optimism (:2445-2628).  Like that backend it is single-threaded per compute
// handle: NNEvaluator runs one handle per server thread (nneval.cpp:562-581).  Model loading is the reference's own (desc.cpp).
// The GEMM micro-kernel is AVX-512 (12 x 32 register tile) with a plain-loop fallback; everything else is written as loops over the
// channel axis that the compiler vectorises (target_clones).
//
// Built by oracle/Makefile.drivers into oracle/_ref/{katago_cpu, kgref_nnloop_cpu, kgref_cpu_selfplay}; checked against the numpy
// oracle (oracle/kg_nn_oracle.py) by tests/test_cpu_baseline.py.
#include "neuralnet/nninterface.h"
#include "neuralnet/nninputs.h"
#include "neuralnet/nneval.h"
#include "neuralnet/modelversion.h"
#include "neuralnet/activations.h"
#include "neuralnet/desc.h"

#include <immintrin.h>

#include <algorithm>
#include <cmath>
#include <cstring>

using namespace std;

#define CLONES __attribute__((target_clones("avx512f", "avx2", "default")))

// ------------------------------------------------------------------------------------------------------------------
// GEMM: C[M x N] (+)= A[M x K] * B[K x N], B packed once into panels of 32 columns: Bp[panel][k][32]
// ------------------------------------------------------------------------------------------------------------------
namespace {

struct PackedB {
  int K = 0, N = 0, panels = 0;
  vector<float> data;   // 64-byte aligned start not guaranteed by vector: loads are unaligned-safe
  void pack(const float* B, int k, int n) {   // B row-major [k][n]
    K = k; N = n; panels = (n + 31) / 32;
    data.assign((size_t)panels * K * 32, 0.0f);
    for(int p = 0; p < panels; p++)
      for(int kk = 0; kk < K; kk++)
        for(int j = 0; j < 32; j++) {
          int col = p * 32 + j;
          if(col < N) data[((size_t)p * K + kk) * 32 + j] = B[(size_t)kk * N + col];
        }
  }
};

static bool haveAvx512() {
  static const bool v = __builtin_cpu_supports("avx512f");
  return v;
}

template <int MR>
__attribute__((target("avx512f"))) static inline void micro512(int K, const float* A, int lda, const float* Bp, float* C, int ldc, int nvalid, bool accumulate) {
  __m512 acc[MR][2];
  for(int r = 0; r < MR; r++) { acc[r][0] = _mm512_setzero_ps(); acc[r][1] = _mm512_setzero_ps(); }
  for(int k = 0; k < K; k++) {
    const __m512 b0 = _mm512_loadu_ps(Bp + (size_t)k * 32), b1 = _mm512_loadu_ps(Bp + (size_t)k * 32 + 16);
    for(int r = 0; r < MR; r++) {
      const __m512 a = _mm512_set1_ps(A[(size_t)r * lda + k]);
      acc[r][0] = _mm512_fmadd_ps(a, b0, acc[r][0]);
      acc[r][1] = _mm512_fmadd_ps(a, b1, acc[r][1]);
    }
  }
  const __mmask16 m0 = nvalid >= 16 ? (__mmask16)0xFFFF : (__mmask16)((1u << nvalid) - 1);
  const __mmask16 m1 = nvalid >= 32 ? (__mmask16)0xFFFF : (nvalid > 16 ? (__mmask16)((1u << (nvalid - 16)) - 1) : (__mmask16)0);
  for(int r = 0; r < MR; r++) {
    float* c = C + (size_t)r * ldc;
    if(accumulate) {
      acc[r][0] = _mm512_add_ps(acc[r][0], _mm512_maskz_loadu_ps(m0, c));
      acc[r][1] = _mm512_add_ps(acc[r][1], _mm512_maskz_loadu_ps(m1, c + 16));
    }
    _mm512_mask_storeu_ps(c, m0, acc[r][0]);
    _mm512_mask_storeu_ps(c + 16, m1, acc[r][1]);
  }
}

__attribute__((target("avx512f"))) static void gemm512(int M, const float* A, int lda, const PackedB& B, float* C, int ldc, bool accumulate) {
  for(int p = 0; p < B.panels; p++) {
    const float* Bp = B.data.data() + (size_t)p * B.K * 32;
    const int nvalid = min(32, B.N - p * 32);
    float* Cp = C + p * 32;
    int r = 0;
    for(; r + 12 <= M; r += 12) micro512<12>(B.K, A + (size_t)r * lda, lda, Bp, Cp + (size_t)r * ldc, ldc, nvalid, accumulate);
    const int rem = M - r;
    const float* Ar = A + (size_t)r * lda;
    float* Cr = Cp + (size_t)r * ldc;
    switch(rem) {
      case 1: micro512<1>(B.K, Ar, lda, Bp, Cr, ldc, nvalid, accumulate); break;
      case 2: micro512<2>(B.K, Ar, lda, Bp, Cr, ldc, nvalid, accumulate); break;
      case 3: micro512<3>(B.K, Ar, lda, Bp, Cr, ldc, nvalid, accumulate); break;
      case 4: micro512<4>(B.K, Ar, lda, Bp, Cr, ldc, nvalid, accumulate); break;
      case 5: micro512<5>(B.K, Ar, lda, Bp, Cr, ldc, nvalid, accumulate); break;
      case 6: micro512<6>(B.K, Ar, lda, Bp, Cr, ldc, nvalid, accumulate); break;
      case 7: micro512<7>(B.K, Ar, lda, Bp, Cr, ldc, nvalid, accumulate); break;
      case 8: micro512<8>(B.K, Ar, lda, Bp, Cr, ldc, nvalid, accumulate); break;
      case 9: micro512<9>(B.K, Ar, lda, Bp, Cr, ldc, nvalid, accumulate); break;
      case 10: micro512<10>(B.K, Ar, lda, Bp, Cr, ldc, nvalid, accumulate); break;
      case 11: micro512<11>(B.K, Ar, lda, Bp, Cr, ldc, nvalid, accumulate); break;
      default: break;
    }
  }
}

static void gemmPlain(int M, const float* A, int lda, const PackedB& B, float* C, int ldc, bool accumulate) {
  for(int p = 0; p < B.panels; p++) {
    const float* Bp = B.data.data() + (size_t)p * B.K * 32;
    const int nvalid = min(32, B.N - p * 32);
    for(int r = 0; r < M; r++) {
      float acc[32];
      for(int j = 0; j < 32; j++) acc[j] = 0.0f;
      const float* a = A + (size_t)r * lda;
      for(int k = 0; k < B.K; k++) {
        const float av = a[k];
        const float* b = Bp + (size_t)k * 32;
        for(int j = 0; j < 32; j++) acc[j] += av * b[j];
      }
      float* c = C + (size_t)r * ldc + p * 32;
      for(int j = 0; j < nvalid; j++) c[j] = accumulate ? c[j] + acc[j] : acc[j];
    }
  }
}

static inline void gemm(int M, const float* A, int lda, const PackedB& B, float* C, int ldc, bool accumulate) {
  if(M <= 0) return;
  if(haveAvx512()) gemm512(M, A, lda, B, C, ldc, accumulate);
  else gemmPlain(M, A, lda, B, C, ldc, accumulate);
}

// ------------------------------------------------------------------------------------------------------------------
// Elementwise pieces (vectorised over the channel axis by the compiler)
// ------------------------------------------------------------------------------------------------------------------
static inline float fastExp(float x) {   // 2^(x log2 e): degree-5 polynomial on the fraction, relative error ~2e-7
  x = x < -87.0f ? -87.0f : (x > 88.0f ? 88.0f : x);
  const float t = x * 1.44269504088896341f;
  const float fi = floorf(t);
  const float f = t - fi;
  const float p = 1.0f + f * (0.693147180f + f * (0.240226507f + f * (0.0555041087f + f * (0.00961812911f + f * (0.00133335581f + f * 0.000154035304f)))));
  int32_t bits = ((int32_t)fi + 127) << 23;
  float scale;
  memcpy(&scale, &bits, 4);
  return p * scale;
}

// out[i][c] = mask[i] * act(in[i][c] * scale[c] + bias[c])     (eigenbackend.cpp:739-809; mish = x tanh(softplus x) = x n / (n + 2), n = e^x (e^x + 2))
CLONES static void bnActMask(const float* in, float* out, int rows, int C, const float* scale, const float* bias, int act, const float* mask) {
  for(int i = 0; i < rows; i++) {
    const float m = mask ? mask[i] : 1.0f;
    const float* x = in + (size_t)i * C;
    float* y = out + (size_t)i * C;
    if(act == ACTIVATION_MISH) {
      for(int c = 0; c < C; c++) {
        const float v = x[c] * scale[c] + bias[c];
        const float e = fastExp(v > 20.0f ? 20.0f : v);
        const float n = e * (e + 2.0f);
        y[c] = m * (v * (n / (n + 2.0f)));
      }
    }
    else if(act == ACTIVATION_RELU) {
      for(int c = 0; c < C; c++) { const float v = x[c] * scale[c] + bias[c]; y[c] = m * (v > 0.0f ? v : 0.0f); }
    }
    else if(act == ACTIVATION_SILU) {
      for(int c = 0; c < C; c++) { const float v = x[c] * scale[c] + bias[c]; y[c] = m * (v / (1.0f + fastExp(-v))); }
    }
    else {
      for(int c = 0; c < C; c++) y[c] = m * (x[c] * scale[c] + bias[c]);
    }
  }
}

static float actScalar(float v, int act) {
  if(act == ACTIVATION_RELU) return v > 0.0f ? v : 0.0f;
  if(act == ACTIVATION_MISH) { const float e = fastExp(v > 20.0f ? 20.0f : v); const float n = e * (e + 2.0f); return v * (n / (n + 2.0f)); }
  if(act == ACTIVATION_SILU) return v / (1.0f + fastExp(-v));
  return v;
}

// Winograd F(4x4, 3x3) one-dimensional transforms over channel vectors (Lavin & Gray; the reference's transform pair, eigenbackend.cpp:335-412,
// computes the same bilinear algorithm)
CLONES static void winoInput1D(const float* __restrict d0, const float* __restrict d1, const float* __restrict d2, const float* __restrict d3,
                               const float* __restrict d4, const float* __restrict d5, float* __restrict t0, float* __restrict t1, float* __restrict t2,
                               float* __restrict t3, float* __restrict t4, float* __restrict t5, int C) {
  for(int c = 0; c < C; c++) {
    const float a0 = d0[c], a1 = d1[c], a2 = d2[c], a3 = d3[c], a4 = d4[c], a5 = d5[c];
    t0[c] = 4.0f * a0 - 5.0f * a2 + a4;
    t1[c] = -4.0f * a1 - 4.0f * a2 + a3 + a4;
    t2[c] = 4.0f * a1 - 4.0f * a2 - a3 + a4;
    t3[c] = -2.0f * a1 - a2 + 2.0f * a3 + a4;
    t4[c] = 2.0f * a1 - a2 - 2.0f * a3 + a4;
    t5[c] = 4.0f * a1 - 5.0f * a3 + a5;
  }
}
CLONES static void winoOutput1D(const float* __restrict m0, const float* __restrict m1, const float* __restrict m2, const float* __restrict m3,
                                const float* __restrict m4, const float* __restrict m5, float* __restrict y0, float* __restrict y1, float* __restrict y2,
                                float* __restrict y3, int C) {
  for(int c = 0; c < C; c++) {
    const float a0 = m0[c], a1 = m1[c], a2 = m2[c], a3 = m3[c], a4 = m4[c], a5 = m5[c];
    y0[c] = a0 + a1 + a2 + a3 + a4;
    y1[c] = a1 - a2 + 2.0f * a3 - 2.0f * a4;
    y2[c] = a1 + a2 + 4.0f * a3 + 4.0f * a4;
    y3[c] = a1 - a2 + 8.0f * a3 - 8.0f * a4 + a5;
  }
}
CLONES static void addRows(float* dst, const float* src, size_t n) { for(size_t i = 0; i < n; i++) dst[i] += src[i]; }
CLONES static void copyOrAdd(float* dst, const float* src, int C, bool accumulate) {
  if(accumulate) for(int c = 0; c < C; c++) dst[c] += src[c];
  else for(int c = 0; c < C; c++) dst[c] = src[c];
}

// ------------------------------------------------------------------------------------------------------------------
// Prepared layers
// ------------------------------------------------------------------------------------------------------------------
struct PConv {
  int ky = 0, kx = 0, cin = 0, cout = 0;
  bool wino = false;
  vector<PackedB> w;   // wino: 36 matrices [cin x cout]; otherwise one [ky*kx*cin x cout] (rows ordered (dy, dx, ic))
  void prepare(const ConvLayerDesc& d) {
    ky = d.convYSize; kx = d.convXSize; cin = d.inChannels; cout = d.outChannels;
    if(d.dilationX != 1 || d.dilationY != 1) throw StringError("cpubackend: dilated convolutions are not supported");
    auto W = [&](int oc, int ic, int y, int x) { return d.weights[(((size_t)oc * cin + ic) * ky + y) * kx + x]; };   // (oc, ic, y, x), desc.cpp:110-155
    wino = ky == 3 && kx == 3;
    if(wino) {
      static const float G[6][3] = {{0.25f, 0.0f, 0.0f}, {-1.0f / 6, -1.0f / 6, -1.0f / 6}, {-1.0f / 6, 1.0f / 6, -1.0f / 6},
                                    {1.0f / 24, 1.0f / 12, 1.0f / 6}, {1.0f / 24, -1.0f / 12, 1.0f / 6}, {0.0f, 0.0f, 1.0f}};
      vector<vector<float>> U(36, vector<float>((size_t)cin * cout));
      for(int oc = 0; oc < cout; oc++)
        for(int ic = 0; ic < cin; ic++) {
          float tmp[6][3];
          for(int i = 0; i < 6; i++) for(int x = 0; x < 3; x++) { float s = 0; for(int y = 0; y < 3; y++) s += G[i][y] * W(oc, ic, y, x); tmp[i][x] = s; }
          for(int i = 0; i < 6; i++) for(int j = 0; j < 6; j++) { float s = 0; for(int x = 0; x < 3; x++) s += tmp[i][x] * G[j][x]; U[i * 6 + j][(size_t)ic * cout + oc] = s; }
        }
      w.resize(36);
      for(int xi = 0; xi < 36; xi++) w[xi].pack(U[xi].data(), cin, cout);
    }
    else {
      vector<float> B((size_t)ky * kx * cin * cout);
      for(int y = 0; y < ky; y++) for(int x = 0; x < kx; x++) for(int ic = 0; ic < cin; ic++) for(int oc = 0; oc < cout; oc++)
        B[(((size_t)y * kx + x) * cin + ic) * cout + oc] = W(oc, ic, y, x);
      w.resize(1);
      w[0].pack(B.data(), ky * kx * cin, cout);
    }
  }
};
struct PBN {
  int C = 0, act = ACTIVATION_IDENTITY;
  vector<float> scale, bias;
  void prepare(const BatchNormLayerDesc& d, const ActivationLayerDesc& a) { C = d.numChannels; scale = d.mergedScale; bias = d.mergedBias; act = a.activation; }
};
struct PMatMul {
  int cin = 0, cout = 0;
  PackedB w;
  void prepare(const MatMulLayerDesc& d) { cin = d.inChannels; cout = d.outChannels; w.pack(d.weights.data(), cin, cout); }   // (ic, oc)
};

struct PBlock {
  int kind = 0;
  PBN preBN, midBN, gpoolBN, postBN;
  PConv conv1, conv2, gpoolConv;      // ordinary / gpool: regular + final; nested: pre (1x1) + post (1x1)
  PMatMul gpoolToBias;
  vector<PBlock> inner;
};

static void prepareBlocks(const vector<pair<int, unique_ptr_void>>& blocks, vector<PBlock>& out) {
  for(const auto& kv : blocks) {
    PBlock b;
    b.kind = kv.first;
    if(kv.first == ORDINARY_BLOCK_KIND) {
      const ResidualBlockDesc* d = (const ResidualBlockDesc*)kv.second.get();
      b.preBN.prepare(d->preBN, d->preActivation); b.conv1.prepare(d->regularConv);
      b.midBN.prepare(d->midBN, d->midActivation); b.conv2.prepare(d->finalConv);
    }
    else if(kv.first == GLOBAL_POOLING_BLOCK_KIND) {
      const GlobalPoolingResidualBlockDesc* d = (const GlobalPoolingResidualBlockDesc*)kv.second.get();
      b.preBN.prepare(d->preBN, d->preActivation); b.conv1.prepare(d->regularConv); b.gpoolConv.prepare(d->gpoolConv);
      b.gpoolBN.prepare(d->gpoolBN, d->gpoolActivation); b.gpoolToBias.prepare(d->gpoolToBiasMul);
      b.midBN.prepare(d->midBN, d->midActivation); b.conv2.prepare(d->finalConv);
    }
    else if(kv.first == NESTED_BOTTLENECK_BLOCK_KIND) {
      const NestedBottleneckResidualBlockDesc* d = (const NestedBottleneckResidualBlockDesc*)kv.second.get();
      b.preBN.prepare(d->preBN, d->preActivation); b.conv1.prepare(d->preConv);
      prepareBlocks(d->blocks, b.inner);
      b.postBN.prepare(d->postBN, d->postActivation); b.conv2.prepare(d->postConv);
    }
    else throw StringError("cpubackend: unsupported block kind (transformer blocks are outside the scope of this baseline)");
    out.push_back(std::move(b));
  }
}

struct PModel {
  int version = 0, numInputChannels = 0, numInputGlobalChannels = 0, numPolicyChannels = 0, numValueChannels = 0, numScoreValueChannels = 0;
  int trunkC = 0;
  PConv initialConv; PMatMul initialMatMul;
  vector<PBlock> blocks;
  PBN tipBN;
  // policy head
  PConv p1Conv, g1Conv, p2Conv; PBN g1BN, p1BN; PMatMul gpoolToBias, gpoolToPass, gpoolToPass2; vector<float> passBias; int passAct = ACTIVATION_IDENTITY;
  // value head
  PConv v1Conv, ownershipConv; PBN v1BN; PMatMul v2Mul, v3Mul, sv3Mul; vector<float> v2Bias, v3Bias, sv3Bias; int v2Act = ACTIVATION_IDENTITY;

  explicit PModel(const ModelDesc& m) {
    version = m.modelVersion; numInputChannels = m.numInputChannels; numInputGlobalChannels = m.numInputGlobalChannels;
    numPolicyChannels = m.numPolicyChannels; numValueChannels = m.numValueChannels; numScoreValueChannels = m.numScoreValueChannels;
    if(m.numInputMetaChannels > 0 || m.metaEncoderVersion > 0) throw StringError("cpubackend: SGF metadata encoders are not supported");
    const TrunkDesc& t = m.trunk;
    trunkC = t.trunkNumChannels;
    initialConv.prepare(t.initialConv); initialMatMul.prepare(t.initialMatMul);
    prepareBlocks(t.blocks, blocks);
    tipBN.prepare(t.trunkTipBN, t.trunkTipActivation);
    const PolicyHeadDesc& p = m.policyHead;
    p1Conv.prepare(p.p1Conv); g1Conv.prepare(p.g1Conv); g1BN.prepare(p.g1BN, p.g1Activation); gpoolToBias.prepare(p.gpoolToBiasMul);
    p1BN.prepare(p.p1BN, p.p1Activation); p2Conv.prepare(p.p2Conv); gpoolToPass.prepare(p.gpoolToPassMul);
    if(version >= 15) { passBias = p.gpoolToPassBias.weights; passAct = p.passActivation.activation; gpoolToPass2.prepare(p.gpoolToPassMul2); }
    const ValueHeadDesc& v = m.valueHead;
    v1Conv.prepare(v.v1Conv); v1BN.prepare(v.v1BN, v.v1Activation); v2Mul.prepare(v.v2Mul); v2Bias = v.v2Bias.weights; v2Act = v.v2Activation.activation;
    v3Mul.prepare(v.v3Mul); v3Bias = v.v3Bias.weights; sv3Mul.prepare(v.sv3Mul); sv3Bias = v.sv3Bias.weights; ownershipConv.prepare(v.vOwnershipConv);
  }
};

// ------------------------------------------------------------------------------------------------------------------
// Forward pass (one handle = one thread)
// ------------------------------------------------------------------------------------------------------------------
struct Workspace {
  int n = 0, H = 0, W = 0;
  vector<float> V, Mx, patch, col;   // Winograd transform buffers, im2col
  vector<vector<float>> pool;        // activation buffers handed out by index
  float* buf(int idx, size_t count) {
    if((int)pool.size() <= idx) pool.resize(idx + 1);
    if(pool[idx].size() < count) pool[idx].resize(count);
    return pool[idx].data();
  }
};

// out[n][H][W][cout] (+)= conv(in[n][H][W][cin]); zero padding ("same")
static void conv(const PConv& c, const float* in, float* out, int n, int H, int W, Workspace& ws, bool accumulate) {
  const int cin = c.cin, cout = c.cout;
  if(c.ky == 1 && c.kx == 1) { gemm(n * H * W, in, cin, c.w[0], out, cout, accumulate); return; }
  if(!c.wino) {   // im2col + GEMM (5x5 first layers of old nets)
    const int K = c.ky * c.kx * cin, py = c.ky / 2, px = c.kx / 2;
    ws.col.resize((size_t)n * H * W * K);
    for(int b = 0; b < n; b++) for(int y = 0; y < H; y++) for(int x = 0; x < W; x++) {
      float* dst = ws.col.data() + (((size_t)b * H + y) * W + x) * K;
      for(int dy = 0; dy < c.ky; dy++) for(int dx = 0; dx < c.kx; dx++) {
        const int yy = y + dy - py, xx = x + dx - px;
        float* d = dst + ((size_t)dy * c.kx + dx) * cin;
        if(yy < 0 || yy >= H || xx < 0 || xx >= W) memset(d, 0, sizeof(float) * cin);
        else memcpy(d, in + (((size_t)b * H + yy) * W + xx) * cin, sizeof(float) * cin);
      }
    }
    gemm(n * H * W, ws.col.data(), K, c.w[0], out, cout, accumulate);
    return;
  }
  // Winograd F(4x4,3x3): tiles of 4x4 outputs, 6x6 inputs
  const int ty = (H + 3) / 4, tx = (W + 3) / 4, tiles = n * ty * tx;
  ws.V.resize((size_t)36 * tiles * cin);
  ws.Mx.resize((size_t)36 * tiles * cout);
  ws.patch.resize((size_t)2 * 36 * max(cin, cout));
  float* d = ws.patch.data();                       // [6][6][cin] gathered patch
  float* t = d + (size_t)36 * max(cin, cout);      // [6][6][.] after the first 1-D pass
  for(int b = 0; b < n; b++) for(int iy = 0; iy < ty; iy++) for(int ix = 0; ix < tx; ix++) {
    const int tile = (b * ty + iy) * tx + ix;
    for(int r = 0; r < 6; r++) for(int q = 0; q < 6; q++) {
      const int yy = iy * 4 + r - 1, xx = ix * 4 + q - 1;
      float* dst = d + ((size_t)r * 6 + q) * cin;
      if(yy < 0 || yy >= H || xx < 0 || xx >= W) memset(dst, 0, sizeof(float) * cin);
      else memcpy(dst, in + (((size_t)b * H + yy) * W + xx) * cin, sizeof(float) * cin);
    }
    for(int q = 0; q < 6; q++)    // columns: t[.][q] = B^T d[.][q]
      winoInput1D(d + (0 * 6 + q) * (size_t)cin, d + (1 * 6 + q) * (size_t)cin, d + (2 * 6 + q) * (size_t)cin, d + (3 * 6 + q) * (size_t)cin,
                  d + (4 * 6 + q) * (size_t)cin, d + (5 * 6 + q) * (size_t)cin, t + (0 * 6 + q) * (size_t)cin, t + (1 * 6 + q) * (size_t)cin,
                  t + (2 * 6 + q) * (size_t)cin, t + (3 * 6 + q) * (size_t)cin, t + (4 * 6 + q) * (size_t)cin, t + (5 * 6 + q) * (size_t)cin, cin);
    for(int r = 0; r < 6; r++) {  // rows: V[r][.] = t[r][.] B, scattered to the 36 GEMM operands
      float* v[6];
      for(int q = 0; q < 6; q++) v[q] = ws.V.data() + ((size_t)(r * 6 + q) * tiles + tile) * cin;
      winoInput1D(t + (r * 6 + 0) * (size_t)cin, t + (r * 6 + 1) * (size_t)cin, t + (r * 6 + 2) * (size_t)cin, t + (r * 6 + 3) * (size_t)cin,
                  t + (r * 6 + 4) * (size_t)cin, t + (r * 6 + 5) * (size_t)cin, v[0], v[1], v[2], v[3], v[4], v[5], cin);
    }
  }
  for(int xi = 0; xi < 36; xi++)
    gemm(tiles, ws.V.data() + (size_t)xi * tiles * cin, cin, c.w[xi], ws.Mx.data() + (size_t)xi * tiles * cout, cout, false);
  float* m = ws.patch.data();                      // [6][4][cout] after the first pass, then [4][4][cout]
  float* y4 = m + (size_t)36 * max(cin, cout);
  for(int b = 0; b < n; b++) for(int iy = 0; iy < ty; iy++) for(int ix = 0; ix < tx; ix++) {
    const int tile = (b * ty + iy) * tx + ix;
    auto M = [&](int r, int q) { return ws.Mx.data() + ((size_t)(r * 6 + q) * tiles + tile) * cout; };
    for(int r = 0; r < 6; r++)   // rows: m[r][0..3] = M[r][.] A
      winoOutput1D(M(r, 0), M(r, 1), M(r, 2), M(r, 3), M(r, 4), M(r, 5), m + (r * 4 + 0) * (size_t)cout, m + (r * 4 + 1) * (size_t)cout,
                   m + (r * 4 + 2) * (size_t)cout, m + (r * 4 + 3) * (size_t)cout, cout);
    for(int q = 0; q < 4; q++)   // columns: y[0..3][q] = A^T m[.][q]
      winoOutput1D(m + (0 * 4 + q) * (size_t)cout, m + (1 * 4 + q) * (size_t)cout, m + (2 * 4 + q) * (size_t)cout, m + (3 * 4 + q) * (size_t)cout,
                   m + (4 * 4 + q) * (size_t)cout, m + (5 * 4 + q) * (size_t)cout, y4 + (0 * 4 + q) * (size_t)cout, y4 + (1 * 4 + q) * (size_t)cout,
                   y4 + (2 * 4 + q) * (size_t)cout, y4 + (3 * 4 + q) * (size_t)cout, cout);
    for(int r = 0; r < 4; r++) for(int q = 0; q < 4; q++) {
      const int yy = iy * 4 + r, xx = ix * 4 + q;
      if(yy < H && xx < W) copyOrAdd(out + (((size_t)b * H + yy) * W + xx) * cout, y4 + (r * 4 + q) * (size_t)cout, cout, accumulate);
    }
  }
}

// [n][3C]: mean, mean * (sqrt(maskSum) - 14) / 10, max over on-board points (eigenbackend.cpp:152-177)
static void gpool(const float* x, const float* mask, const float* maskSum, int n, int HW, int C, float* out) {
  for(int b = 0; b < n; b++) {
    float* o = out + (size_t)b * 3 * C;
    for(int c = 0; c < C; c++) { o[c] = 0.0f; o[2 * C + c] = -1.0f; }
    for(int i = 0; i < HW; i++) {
      const float* r = x + ((size_t)b * HW + i) * C;
      const float mm = mask[(size_t)b * HW + i] - 1.0f;
      for(int c = 0; c < C; c++) { o[c] += r[c]; const float v = r[c] + mm; if(v > o[2 * C + c]) o[2 * C + c] = v; }
    }
    const float div = maskSum[b], sq = sqrtf(div);
    for(int c = 0; c < C; c++) { const float mean = o[c] / div; o[c] = mean; o[C + c] = mean * (sq - 14.0f) * 0.1f; }
  }
}
// value head pooling: mean, mean * (s - 14) / 10, mean * ((s - 14)^2 / 100 - 0.1) (eigenbackend.cpp:179-197)
static void vpool(const float* x, const float* maskSum, int n, int HW, int C, float* out) {
  for(int b = 0; b < n; b++) {
    float* o = out + (size_t)b * 3 * C;
    for(int c = 0; c < C; c++) o[c] = 0.0f;
    for(int i = 0; i < HW; i++) { const float* r = x + ((size_t)b * HW + i) * C; for(int c = 0; c < C; c++) o[c] += r[c]; }
    const float div = maskSum[b], sq = sqrtf(div);
    for(int c = 0; c < C; c++) {
      const float mean = o[c] / div;
      o[c] = mean; o[C + c] = mean * (sq - 14.0f) * 0.1f; o[2 * C + c] = mean * ((sq - 14.0f) * (sq - 14.0f) * 0.01f - 0.1f);
    }
  }
}
static void addChannelBias(float* x, const float* bias /*[n][C]*/, int n, int HW, int C) {
  for(int b = 0; b < n; b++) for(int i = 0; i < HW; i++) addRows(x + ((size_t)b * HW + i) * C, bias + (size_t)b * C, C);
}

struct Runner {
  const PModel& m;
  Workspace ws;
  int n = 0, H = 0, W = 0, HW = 0;
  const float* mask = nullptr;
  vector<float> maskSum;
  int nextBuf = 0;
  explicit Runner(const PModel& pm) : m(pm) {}
  float* alloc(size_t count) { return ws.buf(nextBuf++, count); }

  void bnact(const PBN& bn, const float* in, float* out) { bnActMask(in, out, n * HW, bn.C, bn.scale.data(), bn.bias.data(), bn.act, mask); }

  // x: [n][HW][C] residual stream, updated in place.  Scratch buffers are taken from the pool and returned (stack discipline).
  void applyBlocks(const vector<PBlock>& blocks, float* x, int C) {
    for(const PBlock& b : blocks) {
      const int mark = nextBuf;
      float* t = alloc((size_t)n * HW * C);
      bnact(b.preBN, x, t);
      if(b.kind == ORDINARY_BLOCK_KIND) {
        const int midC = b.conv1.cout;
        float* mid = alloc((size_t)n * HW * midC);
        conv(b.conv1, t, mid, n, H, W, ws, false);
        bnact(b.midBN, mid, mid);
        conv(b.conv2, mid, x, n, H, W, ws, true);
      }
      else if(b.kind == GLOBAL_POOLING_BLOCK_KIND) {
        const int regC = b.conv1.cout, gC = b.gpoolConv.cout;
        float* reg = alloc((size_t)n * HW * regC);
        float* g = alloc((size_t)n * HW * gC);
        conv(b.conv1, t, reg, n, H, W, ws, false);
        conv(b.gpoolConv, t, g, n, H, W, ws, false);
        bnact(b.gpoolBN, g, g);
        float* pooled = alloc((size_t)n * 3 * gC);
        float* bias = alloc((size_t)n * regC);
        gpool(g, mask, maskSum.data(), n, HW, gC, pooled);
        gemm(n, pooled, 3 * gC, b.gpoolToBias.w, bias, regC, false);
        addChannelBias(reg, bias, n, HW, regC);
        bnact(b.midBN, reg, reg);
        conv(b.conv2, reg, x, n, H, W, ws, true);
      }
      else {
        const int inC = b.conv1.cout;
        float* y = alloc((size_t)n * HW * inC);
        conv(b.conv1, t, y, n, H, W, ws, false);
        applyBlocks(b.inner, y, inC);
        bnact(b.postBN, y, y);
        conv(b.conv2, y, x, n, H, W, ws, true);
      }
      nextBuf = mark;
    }
  }

  // spatial: [n][HW][Cin] NHWC (symmetry already applied), global: [n][G].  Outputs in the backend's own spatial frame.
  void forward(int n_, int H_, int W_, const float* spatial, const float* global, float* policy /*[n][HW][cp]*/, float* policyPass /*[n][cp]*/,
               float* value /*[n][3]*/, float* scoreValue /*[n][numSV]*/, float* ownership /*[n][HW]*/) {
    n = n_; H = H_; W = W_; HW = H * W; nextBuf = 0;
    const int Cin = m.numInputChannels, C = m.trunkC;
    float* msk = alloc((size_t)n * HW);
    maskSum.assign(n, 0.0f);
    for(int b = 0; b < n; b++) for(int i = 0; i < HW; i++) { const float v = spatial[((size_t)b * HW + i) * Cin]; msk[(size_t)b * HW + i] = v; maskSum[b] += v; }
    mask = msk;
    float* x = alloc((size_t)n * HW * C);
    conv(m.initialConv, spatial, x, n, H, W, ws, false);
    float* gb = alloc((size_t)n * C);
    gemm(n, global, m.numInputGlobalChannels, m.initialMatMul.w, gb, C, false);
    addChannelBias(x, gb, n, HW, C);
    applyBlocks(m.blocks, x, C);
    float* tip = alloc((size_t)n * HW * C);
    bnact(m.tipBN, x, tip);
    // policy head (eigenbackend.cpp:1723-2036)
    const int p1C = m.p1Conv.cout, g1C = m.g1Conv.cout, cp = m.numPolicyChannels;
    float* p1 = alloc((size_t)n * HW * p1C);
    float* g1 = alloc((size_t)n * HW * g1C);
    conv(m.p1Conv, tip, p1, n, H, W, ws, false);
    conv(m.g1Conv, tip, g1, n, H, W, ws, false);
    bnact(m.g1BN, g1, g1);
    float* pooled = alloc((size_t)n * 3 * g1C);
    gpool(g1, mask, maskSum.data(), n, HW, g1C, pooled);
    float* pbias = alloc((size_t)n * p1C);
    gemm(n, pooled, 3 * g1C, m.gpoolToBias.w, pbias, p1C, false);
    addChannelBias(p1, pbias, n, HW, p1C);
    bnact(m.p1BN, p1, p1);
    conv(m.p2Conv, p1, policy, n, H, W, ws, false);
    if(m.version >= 15) {
      const int midP = m.gpoolToPass.cout;
      float* h = alloc((size_t)n * midP);
      gemm(n, pooled, 3 * g1C, m.gpoolToPass.w, h, midP, false);
      for(int b = 0; b < n; b++) for(int c = 0; c < midP; c++) h[(size_t)b * midP + c] = actScalar(h[(size_t)b * midP + c] + m.passBias[c], m.passAct);
      gemm(n, h, midP, m.gpoolToPass2.w, policyPass, cp, false);
    }
    else gemm(n, pooled, 3 * g1C, m.gpoolToPass.w, policyPass, cp, false);
    // value head (eigenbackend.cpp:2051-2216)
    const int v1C = m.v1Conv.cout, v2C = m.v2Mul.cout, numSV = m.sv3Mul.cout;
    float* v1 = alloc((size_t)n * HW * v1C);
    conv(m.v1Conv, tip, v1, n, H, W, ws, false);
    bnact(m.v1BN, v1, v1);
    float* vp = alloc((size_t)n * 3 * v1C);
    vpool(v1, maskSum.data(), n, HW, v1C, vp);
    float* v2 = alloc((size_t)n * v2C);
    gemm(n, vp, 3 * v1C, m.v2Mul.w, v2, v2C, false);
    for(int b = 0; b < n; b++) for(int c = 0; c < v2C; c++) v2[(size_t)b * v2C + c] = actScalar(v2[(size_t)b * v2C + c] + m.v2Bias[c], m.v2Act);
    gemm(n, v2, v2C, m.v3Mul.w, value, 3, false);
    for(int b = 0; b < n; b++) for(int c = 0; c < 3; c++) value[b * 3 + c] += m.v3Bias[c];
    gemm(n, v2, v2C, m.sv3Mul.w, scoreValue, numSV, false);
    for(int b = 0; b < n; b++) for(int c = 0; c < numSV; c++) scoreValue[(size_t)b * numSV + c] += m.sv3Bias[c];
    conv(m.ownershipConv, v1, ownership, n, H, W, ws, false);
  }
};

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// The NeuralNet interface
// ------------------------------------------------------------------------------------------------------------------
struct LoadedModel {
  ModelDesc modelDesc;
  LoadedModel(const string& file, const string& sha) { ModelDesc::loadFromFileMaybeGZipped(file, modelDesc, sha); }
};
struct ComputeContext {
  int nnXLen, nnYLen;
  std::unique_ptr<PModel> model;
};
struct ComputeHandle {
  ComputeContext* ctx;
  std::unique_ptr<Runner> runner;
  int maxBatch;
  bool inputsUseNHWC;
  vector<float> spatial, global, policy, policyPass, value, scoreValue, ownership, tmp;
};
struct InputBuffers { int maxBatchSize; };

void NeuralNet::globalInitialize() {}
void NeuralNet::globalCleanup() {}
void NeuralNet::printDevices() { cout << "CPU baseline backend (restated Eigen path), " << (haveAvx512() ? "AVX-512" : "plain-loop") << " GEMM" << endl; }

LoadedModel* NeuralNet::loadModelFile(const string& file, const string& expectedSha256) { return new LoadedModel(file, expectedSha256); }
void NeuralNet::freeLoadedModel(LoadedModel* loadedModel) { delete loadedModel; }
const ModelDesc& NeuralNet::getModelDesc(const LoadedModel* loadedModel) { return loadedModel->modelDesc; }

ComputeContext* NeuralNet::createComputeContext(const std::vector<int>& gpuIdxs, Logger* logger, int nnXLen, int nnYLen, const string& homeDataDirOverride,
                                                enabled_t useFP16Mode, const LoadedModel* loadedModel, ConfigParser& cfg) {
  (void)gpuIdxs; (void)logger; (void)homeDataDirOverride; (void)cfg;
  if(useFP16Mode == enabled_t::True) throw StringError("cpubackend: FP16 is not supported");
  ComputeContext* c = new ComputeContext();
  c->nnXLen = nnXLen; c->nnYLen = nnYLen;
  c->model.reset(new PModel(loadedModel->modelDesc));
  return c;
}
void NeuralNet::freeComputeContext(ComputeContext* computeContext) { delete computeContext; }

ComputeHandle* NeuralNet::createComputeHandle(ComputeContext* context, const LoadedModel* loadedModel, Logger* logger, int maxBatchSize, bool requireExactNNLen,
                                              bool inputsUseNHWC, int gpuIdxForThisThread, int serverThreadIdx) {
  (void)loadedModel; (void)requireExactNNLen; (void)gpuIdxForThisThread;
  ComputeHandle* h = new ComputeHandle();
  h->ctx = context; h->maxBatch = maxBatchSize; h->inputsUseNHWC = inputsUseNHWC;
  h->runner.reset(new Runner(*context->model));
  if(logger != NULL && serverThreadIdx == 0) logger->write(string("CPU baseline backend: ") + (haveAvx512() ? "AVX-512" : "plain-loop") + " GEMM, one thread per handle");
  return h;
}
void NeuralNet::freeComputeHandle(ComputeHandle* computeHandle) { delete computeHandle; }
bool NeuralNet::isUsingFP16(const ComputeHandle* computeHandle) { (void)computeHandle; return false; }
bool NeuralNet::setIsWarmup(const ComputeHandle* computeHandle, bool isWarmup) { (void)computeHandle; (void)isWarmup; return false; }

InputBuffers* NeuralNet::createInputBuffers(const LoadedModel* loadedModel, int maxBatchSize, int nnXLen, int nnYLen) {
  (void)loadedModel; (void)nnXLen; (void)nnYLen;
  InputBuffers* b = new InputBuffers();
  b->maxBatchSize = maxBatchSize;
  return b;
}
void NeuralNet::freeInputBuffers(InputBuffers* buffers) { delete buffers; }

void NeuralNet::getOutput(ComputeHandle* h, InputBuffers* buffers, int numBatchEltsFilled, NNResultBuf** inputBufs, vector<NNOutput*>& outputs) {
  const int n = numBatchEltsFilled;
  if(n <= 0 || n > h->maxBatch || n > buffers->maxBatchSize) throw StringError("cpubackend: batch size out of range");
  const PModel& m = *h->ctx->model;
  const int X = h->ctx->nnXLen, Y = h->ctx->nnYLen, HW = X * Y, Cin = m.numInputChannels, G = m.numInputGlobalChannels, cp = m.numPolicyChannels;
  const int numSV = m.sv3Mul.cout;
  h->spatial.resize((size_t)n * HW * Cin); h->global.resize((size_t)n * G); h->tmp.resize((size_t)HW * Cin);
  h->policy.resize((size_t)n * HW * cp); h->policyPass.resize((size_t)n * cp); h->value.resize((size_t)n * 3);
  h->scoreValue.resize((size_t)n * numSV); h->ownership.resize((size_t)n * HW);
  for(int b = 0; b < n; b++) {
    const NNResultBuf* rb = inputBufs[b];
    if(rb->hasRowMeta) throw StringError("cpubackend: SGF metadata inputs are not supported");
    float* dst = h->spatial.data() + (size_t)b * HW * Cin;
    if(h->inputsUseNHWC) SymmetryHelpers::copyInputsWithSymmetry(rb->rowSpatialBuf.data(), dst, 1, Y, X, Cin, true, rb->symmetry);
    else {   // NCHW rows: apply the symmetry in their own layout, then transpose to this backend's NHWC
      SymmetryHelpers::copyInputsWithSymmetry(rb->rowSpatialBuf.data(), h->tmp.data(), 1, Y, X, Cin, false, rb->symmetry);
      for(int c = 0; c < Cin; c++) for(int i = 0; i < HW; i++) dst[(size_t)i * Cin + c] = h->tmp[(size_t)c * HW + i];
    }
    std::copy(rb->rowGlobalBuf.begin(), rb->rowGlobalBuf.begin() + G, h->global.begin() + (size_t)b * G);
  }
  h->runner->forward(n, Y, X, h->spatial.data(), h->global.data(), h->policy.data(), h->policyPass.data(), h->value.data(), h->scoreValue.data(),
                     h->ownership.data());
  float policyTmp[NNPos::MAX_NN_POLICY_SIZE];
  for(int b = 0; b < n; b++) {
    NNOutput* o = outputs[b];
    const float optimism = (float)inputBufs[b]->policyOptimism;
    const float* pol = h->policy.data() + (size_t)b * HW * cp;
    const float* pass = h->policyPass.data() + (size_t)b * cp;
    if(cp == 2 || (cp == 4 && m.version >= 16)) {
      for(int i = 0; i < HW; i++) { const float p = pol[(size_t)i * cp], po = pol[(size_t)i * cp + 1]; policyTmp[i] = p + (po - p) * optimism; }
      SymmetryHelpers::copyOutputsWithSymmetry(policyTmp, o->policyProbs, 1, Y, X, inputBufs[b]->symmetry);
      o->policyProbs[HW] = pass[0] + (pass[1] - pass[0]) * optimism;
    }
    else {
      for(int i = 0; i < HW; i++) policyTmp[i] = pol[(size_t)i * cp];
      SymmetryHelpers::copyOutputsWithSymmetry(policyTmp, o->policyProbs, 1, Y, X, inputBufs[b]->symmetry);
      o->policyProbs[HW] = pass[0];
    }
    o->whiteWinProb = h->value[b * 3]; o->whiteLossProb = h->value[b * 3 + 1]; o->whiteNoResultProb = h->value[b * 3 + 2];
    const float* sv = h->scoreValue.data() + (size_t)b * numSV;
    if(m.version >= 9) { o->whiteScoreMean = sv[0]; o->whiteScoreMeanSq = sv[1]; o->whiteLead = sv[2]; o->varTimeLeft = sv[3]; o->shorttermWinlossError = sv[4]; o->shorttermScoreError = sv[5]; }
    else if(m.version >= 8) { o->whiteScoreMean = sv[0]; o->whiteScoreMeanSq = sv[1]; o->whiteLead = sv[2]; o->varTimeLeft = sv[3]; o->shorttermWinlossError = 0; o->shorttermScoreError = 0; }
    else if(m.version >= 4) { o->whiteScoreMean = sv[0]; o->whiteScoreMeanSq = sv[1]; o->whiteLead = sv[0]; o->varTimeLeft = 0; o->shorttermWinlossError = 0; o->shorttermScoreError = 0; }
    else { o->whiteScoreMean = sv[0]; o->whiteScoreMeanSq = sv[0] * sv[0]; o->whiteLead = sv[0]; o->varTimeLeft = 0; o->shorttermWinlossError = 0; o->shorttermScoreError = 0; }
    if(o->whiteOwnerMap != NULL) SymmetryHelpers::copyOutputsWithSymmetry(h->ownership.data() + (size_t)b * HW, o->whiteOwnerMap, 1, Y, X, inputBufs[b]->symmetry);
  }
}

// Layer-level test hooks: not exposed by this baseline backend.
bool NeuralNet::testEvaluateConv(const ConvLayerDesc*, int, int, int, bool, bool, const std::vector<float>&, std::vector<float>&) { return false; }
bool NeuralNet::testEvaluateBatchNorm(const BatchNormLayerDesc*, int, int, int, bool, bool, const std::vector<float>&, const std::vector<float>&, std::vector<float>&) { return false; }
bool NeuralNet::testEvaluateResidualBlock(const ResidualBlockDesc*, int, int, int, bool, bool, const std::vector<float>&, const std::vector<float>&, std::vector<float>&) { return false; }
bool NeuralNet::testEvaluateGlobalPoolingResidualBlock(const GlobalPoolingResidualBlockDesc*, int, int, int, bool, bool, const std::vector<float>&, const std::vector<float>&, std::vector<float>&) { return false; }
