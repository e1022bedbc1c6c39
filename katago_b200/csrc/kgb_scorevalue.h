// ScoreValue::expectedWhiteScoreValue's table (neuralnet/nninputs.cpp:98-192) - SURVEY.md §8a row a21.
#pragma once
#include <cmath>
#include <vector>

namespace kgb {

static constexpr int SV_ASSUMED_BSIZE = 19;                                   // NNPos::MAX_BOARD_LEN
static constexpr int SV_EXTRA_SCORE_DISTR_RADIUS = 60;                        // NNPos::EXTRA_SCORE_DISTR_RADIUS
static constexpr int SV_MEAN_RADIUS = SV_ASSUMED_BSIZE * SV_ASSUMED_BSIZE + SV_EXTRA_SCORE_DISTR_RADIUS;   // 421
static constexpr int SV_MEAN_LEN = SV_MEAN_RADIUS * 2;                        // 842
static constexpr int SV_STDEV_LEN = SV_ASSUMED_BSIZE * SV_ASSUMED_BSIZE + SV_EXTRA_SCORE_DISTR_RADIUS;    // 421

// expectedSVTable[meanIdx * SV_STDEV_LEN + stdevIdx]: E[ (2/pi) atan(score / 19) ] for score ~ N(mean, stdev) on the table's grid,
// by the same 101-point quadrature, in the same order of operations, as ScoreValue::initTables.
std::vector<double> makeExpectedSVTable();

// Value-weighting t-distribution CDF table (search.cpp:131-137; search/distributiontable.h) - row a20.
static constexpr int VW_TABLE_SIZE = 2000;
static constexpr double VW_MIN_Z = -50.0, VW_MAX_Z = 50.0;
std::vector<double> makeValueWeightCdfTable();

#ifdef __CUDACC__
#define KGB_HD __host__ __device__
#else
#define KGB_HD
#endif
// On the device the products and sums are spelled with the round-to-nearest intrinsics so that nvcc cannot contract them into
// FMAs: the host reference (x86-64, no FMA contraction) then gives the same doubles.
#ifdef __CUDA_ARCH__
#define KGB_SV_MUL(a, b) __dmul_rn((a), (b))
#define KGB_SV_ADD(a, b) __dadd_rn((a), (b))
#define KGB_SV_SUB(a, b) __dsub_rn((a), (b))
#else
#define KGB_SV_MUL(a, b) ((a) * (b))
#define KGB_SV_ADD(a, b) ((a) + (b))
#define KGB_SV_SUB(a, b) ((a) - (b))
#endif
// ScoreValue::expectedWhiteScoreValue (nninputs.cpp:160-192): bilinear lookup, mean index rounded, stdev index floored.
KGB_HD inline double svExpectedWhiteScoreValue(const double* table, double whiteScoreMean, double whiteScoreStdev, double center, double scale,
                                               double sqrtBoardArea) {
  const double scaleFactor = (double)SV_ASSUMED_BSIZE / KGB_SV_MUL(scale, sqrtBoardArea);
  const double meanScaled = KGB_SV_MUL(KGB_SV_SUB(whiteScoreMean, center), scaleFactor);
  const double stdevScaled = KGB_SV_MUL(whiteScoreStdev, scaleFactor);
  const double meanRounded = round(meanScaled);
  const double stdevFloored = floor(stdevScaled);
  int meanIdx0 = (int)meanRounded + SV_MEAN_RADIUS, stdevIdx0 = (int)stdevFloored;
  int meanIdx1 = meanIdx0 + 1, stdevIdx1 = stdevIdx0 + 1;
  if(meanIdx0 < 0) { meanIdx0 = 0; meanIdx1 = 0; }
  if(meanIdx1 >= SV_MEAN_LEN) { meanIdx0 = SV_MEAN_LEN - 1; meanIdx1 = SV_MEAN_LEN - 1; }
  if(stdevIdx1 >= SV_STDEV_LEN) { stdevIdx0 = SV_STDEV_LEN - 1; stdevIdx1 = SV_STDEV_LEN - 1; }
  const double lambdaMean = KGB_SV_ADD(KGB_SV_SUB(meanScaled, meanRounded), 0.5);
  const double lambdaStdev = KGB_SV_SUB(stdevScaled, stdevFloored);
  const double a00 = table[meanIdx0 * SV_STDEV_LEN + stdevIdx0], a01 = table[meanIdx0 * SV_STDEV_LEN + stdevIdx1];
  const double a10 = table[meanIdx1 * SV_STDEV_LEN + stdevIdx0], a11 = table[meanIdx1 * SV_STDEV_LEN + stdevIdx1];
  const double b0 = KGB_SV_ADD(a00, KGB_SV_MUL(lambdaStdev, KGB_SV_SUB(a01, a00)));
  const double b1 = KGB_SV_ADD(a10, KGB_SV_MUL(lambdaStdev, KGB_SV_SUB(a11, a10)));
  return KGB_SV_ADD(b0, KGB_SV_MUL(lambdaMean, KGB_SV_SUB(b1, b0)));
}
// DistributionTable::getCdf (search/distributiontable.h:58-70)
KGB_HD inline double vwCdf(const double* table, double z) {
  const double d = KGB_SV_MUL((double)(VW_TABLE_SIZE - 1), KGB_SV_SUB(z, VW_MIN_Z)) / (VW_MAX_Z - VW_MIN_Z);
  if(d <= 0) return 0.0;
  const int idx = (int)d;
  if(idx >= VW_TABLE_SIZE - 1) return 1.0;
  const double lambda = KGB_SV_SUB(d, (double)idx);
  const double y0 = table[idx], y1 = table[idx + 1];
  return KGB_SV_ADD(y0, KGB_SV_MUL(lambda, KGB_SV_SUB(y1, y0)));
}
// ScoreValue::getScoreStdev
KGB_HD inline double svScoreStdev(double scoreMean, double scoreMeanSq) {
  const double variance = KGB_SV_SUB(scoreMeanSq, KGB_SV_MUL(scoreMean, scoreMean));
  return variance <= 0.0 ? 0.0 : sqrt(variance);
}

}  // namespace kgb
