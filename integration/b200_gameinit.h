// Per-game board size, rules and komi for the C++ host: what the reference's GameInitializer draws when a game is created
// (GameInitializer::initShared / createRulesUnsynchronized / createGameSharedUnsynchronized, program/play.cpp:83-214, 470-482, 530, 596-608;
// PlayUtils::chooseExtraBlackAndKomi / setKomiWithNoise / roundAndClipKomi, program/playutils.cpp:24-106, 363-371), for the options the device
// loop has: board size from bSizes / bSizeRelProbs (allowRectangleProb: every ordered pair of edges), ko rule and multi-stone suicide uniform
// over their lists, komi = komiMean + truncated Gaussian noise (komiStdev / komiBigStdev / komiBiggerStdev, scaled by sqrt(area) / 19), rounded
// to a half-integer with linear probability, clipped, made non-integer with probability 1 - komiAllowIntegerProb.
// The stand-alone twin of katago_b200/game_initializer.py, draw for draw (same Mersenne Twister stream: PyRandom): the reference seeds its
// GameInitializer from the clock, so there is no reference stream to follow, only its distributions (tests/golden/gameinit_hist.json pins the
// Python twin to 200 000 games of the reference's createGame).
#pragma once
#include "b200_recorder.h"

namespace b200 {

// PlayUtils::roundAndClipKomi (playutils.cpp:363-371)
inline double roundAndClipKomi(double unrounded, int xSize, int ySize) {
  const double range = 20.0 + xSize * ySize;        // NNPos::KOMI_CLIP_RADIUS + area
  unrounded = std::min(std::max(unrounded, -range), range);
  return unrounded >= 0 ? 0.5 * std::floor(2.0 * unrounded + 0.5) : -0.5 * std::floor(-2.0 * unrounded + 0.5);      // C round(): halves away from zero
}

class GameInitializer {
 public:
  struct Config {
    std::vector<int> edges; std::vector<double> relProbs; double allowRectangleProb = 0.0;
    std::vector<int> koRules{0}; std::vector<int> multiStoneSuicideLegals{1};
    double komiMean = 7.5, komiStdev = 0.0, komiBigStdevProb = 0.0, komiBigStdev = 10.0, komiBiggerStdevProb = 0.0, komiBiggerStdev = 30.0, komiAllowIntegerProb = 1.0;
  };
  struct Game { int x, y, koRule, multiStoneSuicideLegal; float komi; };

  GameInitializer(const Config& c, uint64_t seed) : c_(c), rand_(seed) {
    // the board size distribution of GameInitializer::initShared (play.cpp:139-172)
    if(c.edges.empty() || c.edges.size() != c.relProbs.size()) throw std::invalid_argument("bSizeRelProbs: one entry per bSizes entry");
    double total = 0.0;
    for(double p : c.relProbs) total += p;
    if(!(total > 0)) throw std::invalid_argument("bSizeRelProbs must sum to a positive value");
    for(size_t i = 0; i < c.edges.size(); i++)
      for(size_t j = 0; j < c.edges.size(); j++) {
        if(i == j) {
          sizes_.push_back({c.edges[i], c.edges[j]});
          sizeProbs_.push_back((1.0 - c.allowRectangleProb) * c.relProbs[i] / total + c.allowRectangleProb * c.relProbs[i] * c.relProbs[j] / total / total);
        }
        else if(c.allowRectangleProb > 0.0) {
          sizes_.push_back({c.edges[i], c.edges[j]});
          sizeProbs_.push_back(c.allowRectangleProb * c.relProbs[i] * c.relProbs[j] / total / total);
        }
      }
  }
  int maxEdge() const { int m = 0; for(int e : c_.edges) m = std::max(m, e); return m; }

  // chooseExtraBlackAndKomi (no handicap) + setKomiWithNoise
  double komiMean() const { return c_.komiMean; }
  double uniform() { return rand_.random(); }        // one draw of the initializer's own stream (adjustKomiToEven's rounding draws from it)
  // mean: instead of komiMean (komiAuto: the fair komi of the empty board; NaN = komiMean)
  float drawKomi(int xSize, int ySize, double mean = std::nan("")) {
    double stdev = c_.komiStdev > 0 ? c_.komiStdev : 0.0;
    if(c_.komiBigStdev > 0 && rand_.random() < c_.komiBigStdevProb) stdev = c_.komiBigStdev;
    if(c_.komiBiggerStdev > 0 && c_.komiBiggerStdevProb > 0 && rand_.random() < c_.komiBiggerStdevProb) stdev = c_.komiBiggerStdev;
    stdev *= std::sqrt((double)(xSize * ySize)) / 19.0;       // no massive komis on small boards
    const bool allowInteger = rand_.random() < c_.komiAllowIntegerProb;
    double komi = std::isnan(mean) ? c_.komiMean : mean;
    if(stdev > 0) {
      double d = rand_.gauss(0.0, 1.0);
      while(d < -3.0 || d > 3.0) d = rand_.gauss(0.0, 1.0);   // nextGaussianTruncated(3.0)
      komi += stdev * d;
    }
    const double lower = std::floor(komi * 2.0) / 2.0, upper = std::ceil(komi * 2.0) / 2.0;          // roundKomiWithLinearProb
    komi = lower == upper ? lower : (rand_.random() < (komi - lower) / (upper - lower) ? upper : lower);
    komi = roundAndClipKomi(komi, xSize, ySize);
    if(!allowInteger && komi == (double)(long)komi) komi += rand_.random() < 0.5 ? -0.5 : 0.5;
    return (float)komi;
  }
  // numInitialMovesToPlay of PlayUtils::initializeGameUsingPolicy (playutils.cpp:243-250, gamma shape 1): floor of an exponential with mean
  // board area * policyInitAreaProp - drawn from the same stream, after the game's own draws
  int openingLength(int xSize, int ySize, double areaProp) { return (int)std::floor(rand_.expovariate(1.0) * xSize * ySize * areaProp); }
  Game draw() {
    const std::pair<int, int> size = sizes_[rand_.choiceIndex(sizeProbs_)];
    const int ko = c_.koRules[rand_.randrange((uint32_t)c_.koRules.size())];
    const int suicide = c_.multiStoneSuicideLegals[rand_.randrange((uint32_t)c_.multiStoneSuicideLegals.size())];
    return Game{size.first, size.second, ko, suicide, drawKomi(size.first, size.second)};
  }

 private:
  Config c_; PyRandom rand_;
  std::vector<std::pair<int, int>> sizes_; std::vector<double> sizeProbs_;
};

}  // namespace b200
