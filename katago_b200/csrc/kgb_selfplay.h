// Internal interface between the C-ABI layer (kgb_api.cu) and the device self-play loop (kgb_selfplay.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>

#include "../../include/kgb200.h"

namespace kgb {

struct SelfplayImpl;

struct SelfplayNNBuffers {  // device buffers of the evaluator handle the loop writes to / reads from
  float* spatial; float* global; float* optimism; int* symmetry;
  const float* policy; const float* value; const float* score;
  const float* ownership;   // [game][XY] raw ownership logits of the last wave, original orientation, mover's perspective
  double scoreMeanMultiplier, scoreStdevMultiplier, leadMultiplier;   // ModelPostProcessParams (desc.h) of the loaded net
};

SelfplayImpl* selfplayCreate(const kgb_selfplay_config& c, int X, int Y, const SelfplayNNBuffers& nn, cudaStream_t stream);
void selfplayDestroy(SelfplayImpl* sp);
void selfplayClearNNCache(SelfplayImpl* sp, cudaStream_t s);
void selfplaySetKomi(SelfplayImpl* sp, const float* komi, bool alsoCurrent);
void selfplayReadLeafKey(SelfplayImpl* sp, int g, unsigned long long* key2);
void selfplayReadKomi(SelfplayImpl* sp, float* current, float* lastFinished);
void selfplaySetGameSetup(SelfplayImpl* sp, const int* setup, bool alsoCurrent, cudaStream_t s);
void selfplayReadGameSetup(SelfplayImpl* sp, int* current, int* lastFinished);
void selfplaySetNextSearchLimits(SelfplayImpl* sp, const int* visits, const uint8_t* plain, bool alsoCurrent, cudaStream_t s);
void selfplayReadSearchLimits(SelfplayImpl* sp, int* visits, uint8_t* plain);
void selfplayReadRootRawEntropy(SelfplayImpl* sp, double* out);
void selfplayReadSymmetries(SelfplayImpl* sp, int* out);
void selfplaySetPolicyInit(SelfplayImpl* sp, const int* moves, double temperature, bool alsoCurrent, cudaStream_t s);
void selfplayReadPolicyInit(SelfplayImpl* sp, int* movesLeft, int* count, int16_t* moves, int maxMoves);
void selfplayLaunchSelect(SelfplayImpl* sp, cudaStream_t s);
void selfplayLaunchBackup(SelfplayImpl* sp, cudaStream_t s);
void selfplayLaunchFakeNN(SelfplayImpl* sp, float* policyOut, float* valueOut, float* scoreOut, float* ownershipOut, cudaStream_t s);
void selfplayPlayMoves(SelfplayImpl* sp, const int8_t* movesXY, int numMoves, cudaStream_t s, int onlyGame = -1);
void selfplaySetSearchRand(SelfplayImpl* sp, const char* seedString);
void selfplayRandomOpenings(SelfplayImpl* sp, int maxLen, cudaStream_t s);
void selfplayReadStats(SelfplayImpl* sp, kgb_selfplay_stats* out);
void selfplayReadRootRow(SelfplayImpl* sp, int g, float* spatial, float* global);
void selfplayReadDebugCycles(SelfplayImpl* sp, long long* out, bool clear);
void selfplayReadGame(SelfplayImpl* sp, int g, uint8_t* colors, int* info);
void selfplayReadRootMoments(SelfplayImpl* sp, int g, double* childMoments, double* rootMoments);
void selfplayRelease(SelfplayImpl* sp, const uint8_t* mask);
void selfplayReadRootVisitsAll(SelfplayImpl* sp, int* out);
void selfplayReadRootExtra(SelfplayImpl* sp, int g, int* childNodeVisits, double* rootNN);
void selfplayReadLastMove(SelfplayImpl* sp, int g, int* info, float* score, uint8_t* finalColors, uint8_t* finalArea);
int selfplayReadLeafPath(SelfplayImpl* sp, int g, int* movesXY, int maxLen, int* valid);
void selfplayReadRootChildren(SelfplayImpl* sp, int g, int* visits, float* policy, double* utilSum);
void selfplayReadPlaySelection(SelfplayImpl* sp, int g, double* out);
void chooseIndexTest(const char* seedString, const double* probs, int n, double temperature, double onlyBelowProb, int count, int* out);
void historyReplay(int X, int Y, int koRule, int multiSuicide, int numGames, int maxMoves, const int8_t* moves, uint8_t* flags, uint8_t* legal,
                   uint8_t* banned);
void repBoundTest(int X, int Y, int numMoves, int bound, const int8_t* moves, uint8_t* out);
void rootNoiseTest(const char* seedString, int X, int Y, int policySize, int turnNumber, int noise, double concentration, double weight,
                   double temperature, double temperatureEarly, double halflife, const float* policyIn, float* policyOut);
void boardReplay(int X, int Y, int numBoards, int numMoves, int multiSuicide, const int8_t* moves, uint8_t* colors, int8_t* ko, int16_t* caps,
                 uint8_t* libClass, uint8_t* legalNext, uint64_t* posHash, uint8_t* area);

}  // namespace kgb
