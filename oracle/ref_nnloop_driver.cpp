// oracle/ref_nnloop_driver.cpp - TEST / MEASUREMENT INFRASTRUCTURE (same-box GPU competitor, SURVEY.md §8d "the GPU bar to beat").
// Times the reference's own backend boundary, NeuralNet::getOutput (neuralnet/nninterface.h:117), in a plain loop at a fixed batch:
// no search threads, no batching queue - the NN forward including the backend's own host<->device copies, which is exactly what
// NNEvaluator::serve pays per batch (neuralnet/nneval.cpp:709).  The same source is linked twice:
//   oracle/_ref/kgref_nnloop_cuda   against the reference's CUDA/cuDNN backend (cudabackend.cpp, oracle/Makefile.cuda objects)
//   oracle/_ref/kgref_nnloop_b200   against integration/b200backend.cpp -> libkgb200.so
// so both numbers come from identical caller code on the same box.
//
//   kgref_nnloop_X MODEL BATCH ITERS [fp16=1] [nhwc=1] [ownership=1] [DUMP.bin]
// DUMP.bin (optional): inputs and outputs of the first 8 rows, little-endian float32 / int32, for tests/gpu_checks/competitor_parity.py:
//   int32 rows, numSpatial, numGlobal; per row: int32 symmetry, float spatial[numSpatial*361] (as passed), float global[numGlobal],
//   float policy[362], float value[3], float score[6], float ownership[361]
#include "core/global.h"
#include "core/config_parser.h"
#include "core/logger.h"
#include "core/rand.h"
#include "game/board.h"
#include "neuralnet/nninterface.h"
#include "neuralnet/nninputs.h"
#include "neuralnet/nneval.h"
#include "neuralnet/modelversion.h"

#include <chrono>
#include <cstdio>
#include <iostream>

using namespace std;

namespace Version {  // main.cpp normally defines these (cpp/main.h)
  std::string getKataGoVersion() { return "ref_nnloop_driver"; }
  std::string getKataGoVersionForHelp() { return "ref_nnloop_driver"; }
  std::string getKataGoVersionFullInfo() { return "ref_nnloop_driver"; }
  std::string getGitRevision() { return "<none>"; }
  std::string getGitRevisionWithBackend() { return "<none>"; }
}

int main(int argc, char** argv) {
  if(argc < 4) { cerr << "usage: kgref_nnloop MODEL BATCH ITERS [fp16=1] [nhwc=1] [ownership=1]" << endl; return 1; }
  const string modelFile = argv[1];
  const int batch = atoi(argv[2]);
  const int iters = atoi(argv[3]);
  const bool fp16 = argc > 4 ? atoi(argv[4]) != 0 : true;
  const bool nhwc = argc > 5 ? atoi(argv[5]) != 0 : true;
  const bool wantOwner = argc > 6 ? atoi(argv[6]) != 0 : true;
  const string dumpFile = argc > 7 ? argv[7] : "";
  const int L = 19;
  Board::initHash();
  ScoreValue::initTables();
  NeuralNet::globalInitialize();
  Logger logger(nullptr, false, false, false, false);
  LoadedModel* model = NeuralNet::loadModelFile(modelFile, "");
  const ModelDesc& desc = NeuralNet::getModelDesc(model);
  const int numSpatial = NNModelVersion::getNumSpatialFeatures(desc.modelVersion);
  const int numGlobal = NNModelVersion::getNumGlobalFeatures(desc.modelVersion);
  ConfigParser cfg(std::map<std::string, std::string>{});
  ComputeContext* ctx = NeuralNet::createComputeContext({0}, &logger, L, L, "", fp16 ? enabled_t::True : enabled_t::False, model, cfg);
  ComputeHandle* handle = NeuralNet::createComputeHandle(ctx, model, &logger, batch, true, nhwc, 0, 0);
  InputBuffers* inputBuffers = NeuralNet::createInputBuffers(model, batch, L, L);

  // random stones on a full 19x19 board: plane 0 = on board, planes 1/2 = stones, a few binary planes and globals set at random
  Rand rand("nnloop");
  vector<NNResultBuf*> bufs(batch);
  vector<NNOutput*> outs(batch);
  for(int b = 0; b < batch; b++) {
    NNResultBuf* buf = new NNResultBuf();
    buf->includeOwnerMap = wantOwner;
    buf->boardXSizeForServer = L; buf->boardYSizeForServer = L;
    buf->rowSpatialBuf.assign((size_t)numSpatial * L * L, 0.0f);
    buf->rowGlobalBuf.assign(numGlobal, 0.0f);
    buf->hasRowMeta = false;
    buf->symmetry = (int)rand.nextUInt(8);
    buf->policyOptimism = 0.0;
    for(int pos = 0; pos < L * L; pos++) {
      for(int c = 0; c < numSpatial; c++) {
        float v = 0.0f;
        if(c == 0) v = 1.0f;
        else if(c == 1 || c == 2) v = 0.0f;
        else v = rand.nextBool(0.05) ? 1.0f : 0.0f;
        size_t idx = nhwc ? (size_t)pos * numSpatial + c : (size_t)c * L * L + pos;
        buf->rowSpatialBuf[idx] = v;
      }
      const int stone = (int)rand.nextUInt(3);
      if(stone > 0) {
        size_t idx = nhwc ? (size_t)pos * numSpatial + stone : (size_t)stone * L * L + pos;
        buf->rowSpatialBuf[idx] = 1.0f;
      }
    }
    for(int c = 0; c < numGlobal; c++) buf->rowGlobalBuf[c] = (float)(rand.nextDouble() - 0.5);
    bufs[b] = buf;
    NNOutput* o = new NNOutput();
    o->nnXLen = L; o->nnYLen = L;
    o->whiteOwnerMap = wantOwner ? new float[L * L] : nullptr;
    o->noisedPolicyProbs = nullptr;
    outs[b] = o;
  }

  for(int i = 0; i < 5; i++) NeuralNet::getOutput(handle, inputBuffers, batch, bufs.data(), outs);
  auto t0 = chrono::steady_clock::now();
  for(int i = 0; i < iters; i++) NeuralNet::getOutput(handle, inputBuffers, batch, bufs.data(), outs);
  auto t1 = chrono::steady_clock::now();
  const double ms = chrono::duration<double, milli>(t1 - t0).count() / iters;
  // A backend may return its raw outputs scaled down: the reference's CUDA backend rewrites the net with
  // ModelDesc::applyScale8ToReduceActivations (cudabackend.cpp:3168) and NNEvaluator multiplies every raw output by
  // postProcessParams.outputScaleMultiplier (= 8 there, 1 for backends that leave the net alone; nneval.cpp:962-1131,1245).
  // Checksum and dump are in the evaluator's units, i.e. after that multiplication.
  const float outScale = desc.postProcessParams.outputScaleMultiplier;
  double chk = 0.0;
  for(int b = 0; b < batch; b++) chk += (double)outScale * (outs[b]->policyProbs[b % (L * L)] + outs[b]->whiteWinProb);
  cout << "{\"model\": \"" << desc.name << "\", \"batch\": " << batch << ", \"iters\": " << iters << ", \"fp16\": " << (fp16 ? 1 : 0)
       << ", \"nhwc\": " << (nhwc ? 1 : 0) << ", \"ms_per_getOutput\": " << ms << ", \"evals_per_s\": " << (batch / ms * 1e3)
       << ", \"output_scale_multiplier\": " << outScale << ", \"checksum\": " << chk << "}" << endl;
  if(dumpFile != "") {
    FILE* f = fopen(dumpFile.c_str(), "wb");
    const int rows = batch < 8 ? batch : 8;
    int32_t hdr[3] = {rows, numSpatial, numGlobal};
    fwrite(hdr, 4, 3, f);
    for(int b = 0; b < rows; b++) {
      int32_t sym = bufs[b]->symmetry;
      fwrite(&sym, 4, 1, f);
      fwrite(bufs[b]->rowSpatialBuf.data(), 4, (size_t)numSpatial * L * L, f);
      fwrite(bufs[b]->rowGlobalBuf.data(), 4, numGlobal, f);
      vector<float> pol(outs[b]->policyProbs, outs[b]->policyProbs + L * L + 1);
      for(float& x : pol) x *= outScale;
      fwrite(pol.data(), 4, L * L + 1, f);
      float v[9] = {outs[b]->whiteWinProb, outs[b]->whiteLossProb, outs[b]->whiteNoResultProb, outs[b]->whiteScoreMean, outs[b]->whiteScoreMeanSq,
                    outs[b]->whiteLead, outs[b]->varTimeLeft, outs[b]->shorttermWinlossError, outs[b]->shorttermScoreError};
      for(float& x : v) x *= outScale;
      fwrite(v, 4, 9, f);
      vector<float> own(L * L, 0.0f);
      if(outs[b]->whiteOwnerMap) std::copy(outs[b]->whiteOwnerMap, outs[b]->whiteOwnerMap + L * L, own.begin());
      for(float& x : own) x *= outScale;
      fwrite(own.data(), 4, L * L, f);
    }
    fclose(f);
  }
  NeuralNet::freeInputBuffers(inputBuffers);
  NeuralNet::freeComputeHandle(handle);
  NeuralNet::freeComputeContext(ctx);
  NeuralNet::freeLoadedModel(model);
  NeuralNet::globalCleanup();
  return 0;
}
