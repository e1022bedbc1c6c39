// Komi bisection for the C++ host: the fair komi of a position (komiAuto) and the lead target of recorded turns (estimateLeadProb) -
// PlayUtils::getNaiveEvenKomiHelper / adjustKomiToEven / computeLead (program/playutils.cpp:455-660), the stand-alone twin of
// katago_b200/komi_search.py.
//
// The reference runs the searches inline (evalKomi, :419-453); here every search is a job on a side device loop, so the algorithm must be
// able to stop at each evaluation it needs and go on when the answer is there.  The Python host uses generators for that; this file keeps
// the algorithm a plain function over an oracle `eval(komi) -> (lead, winLoss)` that answers from the evaluations known so far (the
// reference's scoreWLCache) and throws NeedKomi for the first one that is missing: the caller has that komi searched, adds the answer and
// calls the function again from the top - it is deterministic and cheap, so it arrives at the same place and goes one evaluation further.
// Parity: tests/test_komi_search.py drives computeLead with the reference's own per-komi searches (tests/golden/komitable.json.gz) and lands
// on the reference's PlayUtils::computeLead bit for bit in 472 results.
#pragma once
#include <cmath>
#include <functional>
#include <map>
#include <utility>

#include "b200_gameinit.h"

namespace b200 {

struct NeedKomi { float komi; };      // thrown by a KomiOracle: this (rounded, clipped) komi has not been searched yet

// answers from the evaluations known so far; the key is the rounded, clipped komi as a float (the reference's map<float, ...>)
class KomiOracle {
 public:
  KomiOracle(int xSize, int ySize) : x_(xSize), y_(ySize) {}
  void add(float komi, double lead, double winLoss) { known_[komi] = {lead, winLoss}; }
  size_t size() const { return known_.size(); }
  std::pair<double, double> operator()(double komi) const {
    const float k = (float)roundAndClipKomi(komi, x_, y_);
    auto it = known_.find(k);
    if(it == known_.end()) throw NeedKomi{k};
    return it->second;
  }
  double clip(double komi) const { return (double)(float)roundAndClipKomi(komi, x_, y_); }
 private:
  int x_, y_; std::map<float, std::pair<double, double>> known_;
};

// getNaiveEvenKomiHelper (playutils.cpp:455-589): the komi at which the position is even; white's perspective throughout
inline double naiveEvenKomi(double oldKomi, const KomiOracle& ev) {
  double komi = oldKomi, lastShift = 0.0, lastWinLoss = 0.0, lastLead = 0.0;
  for(int i = 0; i < 3; i++) {
    const std::pair<double, double> r = ev(komi);
    const double lead = r.first, winLoss = r.second;
    if(i > 0 && ((lastLead > 0 && lead > lastLead + 5 && winLoss < 0.75) || (lastLead < 0 && lead < lastLead - 5 && winLoss > -0.75) ||
                 (lastWinLoss > 0 && winLoss > lastWinLoss + 0.1) || (lastWinLoss < 0 && winLoss < lastWinLoss - 0.1))) {
      komi = ev.clip(komi - (double)(float)lastShift * 0.5);      // the shift made things worse: take half of it back
      break;
    }
    lastLead = lead; lastWinLoss = winLoss;
    double shift = -lead;
    if(i > 0 && std::fabs(shift) > std::fabs(lastShift)) shift = shift < 0 ? -std::fabs(lastShift) : shift > 0 ? std::fabs(lastShift) : shift;
    lastShift = shift;
    if((shift > 0 && winLoss > 0) || (shift < 0 && lead < 0)) break;        // score and win rate pull in opposite directions
    komi = ev.clip(komi + shift);
    if(std::fabs(shift) < 16.0) break;
  }
  auto winLossAt = [&](double delta) { return ev(komi + delta).second; };
  const double wl0 = winLossAt(0.0);
  double lower, upper, lowerWL, upperWL;
  if(wl0 < 0) {
    lower = 0.0; lowerWL = wl0; upper = 0.0; upperWL = 0.0;
    for(int i = 0; i < 6; i++) { upper = std::round(std::pow(2.0, i)); upperWL = winLossAt(upper); if(upperWL >= 0) break; }
  }
  else {
    upper = 0.0; upperWL = wl0; lower = 0.0; lowerWL = 0.0;
    for(int i = 0; i < 6; i++) { lower = -std::round(std::pow(2.0, i)); lowerWL = winLossAt(lower); if(lowerWL <= 0) break; }
  }
  while(upper - lower > 0.50001) {
    const double mid = 0.5 * (lower + upper), midWL = winLossAt(mid);
    if(midWL < 0) { lower = mid; lowerWL = midWL; }
    else { upper = mid; upperWL = midWL; }
  }
  double final_;
  if(lowerWL >= upperWL - 1e-30) final_ = 0.5 * (lower + upper);
  else if(upperWL <= 0) final_ = upper;
  else if(lowerWL >= 0) final_ = lower;
  else final_ = lower + (upper - lower) * (0 - lowerWL) / (upperWL - lowerWL);
  return komi + final_;
}

// PlayUtils::adjustKomiToEven (playutils.cpp:591-610): the even komi, rounded to a half-integer with linear probability.  `unit` is the one
// uniform draw the rounding takes; it is only consumed once the bisection is complete (the function does not throw after reading it).
inline float adjustKomiToEven(double oldKomi, int xSize, int ySize, const KomiOracle& ev, const std::function<double()>& unit) {
  const double newKomi = naiveEvenKomi(oldKomi, ev);
  const double lower = std::floor(newKomi * 2.0) * 0.5, upper = lower + 0.5;
  const double rounded = unit() < (newKomi - lower) / (upper - lower) ? upper : lower;
  return (float)roundAndClipKomi(rounded, xSize, ySize);
}

// PlayUtils::computeLead (playutils.cpp:612-660) under area scoring without button (coarse 2-point granularity): how many points white is
// ahead at `oldKomi` = oldKomi - the even komi, the even komi smoothed over the granularity
inline float computeLead(double oldKomi, const KomiOracle& ev) {
  const double naive = naiveEvenKomi(oldKomi, ev);
  if(naive == std::round(naive)) return (float)(oldKomi - naive);
  const double lower = std::floor(naive * 2.0) * 0.5, upper = lower + 0.5;
  const double wlUpper = ev(upper).second, wlLowerMinus = ev(lower - 0.5).second;
  const double lowerWL = 0.5 * (wlUpper + wlLowerMinus);
  const double wlUpperPlus = ev(upper + 0.5).second, wlLower = ev(lower).second;
  const double upperWL = 0.5 * (wlUpperPlus + wlLower);
  double result;
  if(lowerWL >= upperWL - 1e-30) result = 0.5 * (lower + upper);
  else {
    result = lower + (upper - lower) * (0 - lowerWL) / (upperWL - lowerWL);
    result = std::min(std::max(result, lower - 0.5), upper + 0.5);
  }
  return (float)(oldKomi - result);
}

}  // namespace b200
