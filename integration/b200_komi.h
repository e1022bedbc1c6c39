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
#include <memory>
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

// Position queries of a job (forked games, fork_play.py): "search the position after these moves at this komi" -> lead, winLoss, the net's own score of
// the root and the legal moves, or invalid when replaying the moves ended the game on the way.  Like KomiOracle the oracle answers from what is
// known and throws for the first query that is missing; since the algorithm is re-run from the top every time, its queries - and the random
// draws it makes on the way (draw()) - are remembered in the order they were made, so that a re-run sees the same draws and arrives at the same place.
struct PositionAnswer { bool valid = false; double lead = 0, winLoss = 0, nnScoreMean = 0; std::vector<uint8_t> legal; std::shared_ptr<void> payload; };
struct NeedPosition { std::vector<Move> moves; float komi; };
class PositionOracle {
 public:
  const PositionAnswer& operator()(const std::vector<Move>& moves, float komi) {
    if(queryCursor_ < answers_.size()) return answers_[queryCursor_++];
    throw NeedPosition{moves, komi};
  }
  uint32_t draw(const std::function<uint32_t()>& fresh) {
    if(drawCursor_ == draws_.size()) draws_.push_back(fresh());
    return draws_[drawCursor_++];
  }
  void restart() { queryCursor_ = drawCursor_ = 0; }
  void add(PositionAnswer a) { answers_.push_back(std::move(a)); }
 private:
  std::vector<PositionAnswer> answers_; std::vector<uint32_t> draws_; size_t queryCursor_ = 0, drawCursor_ = 0;
};

// Runs komi-search jobs on a side device loop: its own handle, a few slots in hold mode, the game's search parameters with the root noise off and
// numVisits visits (getNoiselessParams, playutils.cpp:372-387).  A job is an algorithm over a KomiOracle (above) plus the position it is about; a
// slot is loaded with the position at the komi the job asks for by ending whatever game it holds (passes: kgb_selfplay_play_moves_game restarts
// the slot with the setup and komi handed over for its next game) and replaying the position's moves from the empty board.  The twin of
// katago_b200/komi_search.py KomiSearcher, scheduling included (which job gets which slot when), so that both hosts finish jobs in the same order.
class KomiSearcher {
 public:
  using Algorithm = std::function<void(const KomiOracle&)>;     // throws NeedKomi until it can finish; then delivers its result itself
  using PositionAlgorithm = std::function<void(PositionOracle&)>;   // the same over position queries (throws NeedPosition)

  KomiSearcher(GameSlots& side, int maxVisits) : sp_(side), maxVisits_(maxVisits), n_(side.numSlots()) {
    for(int g = 0; g < n_; g++) free_.push_back(g);
    setups_.assign((size_t)n_, GameSlots::GameSetup{side.xLen(), side.yLen(), 0, 1});
    komis_.assign((size_t)n_, 7.5f);
  }
  void submit(const GameSlots::GameSetup& setup, const std::vector<Move>& moves, Algorithm algorithm) {
    queue_.push_back(std::unique_ptr<Job>(new Job{setup, moves, KomiOracle(setup.x, setup.y), std::move(algorithm), 0.0f, false, PositionOracle(), nullptr, {}}));
    dispatch();
  }
  void submitPositions(const GameSlots::GameSetup& setup, PositionAlgorithm algorithm) {
    queue_.push_back(std::unique_ptr<Job>(new Job{setup, {}, KomiOracle(setup.x, setup.y), nullptr, 0.0f, true, PositionOracle(), std::move(algorithm), {}}));
    dispatch();
  }
  // called for every valid position answer while the slot still holds the searched position: what else the job wants read from it (a side
  // position's training targets) goes into the answer's payload
  std::function<std::shared_ptr<void>(const GameSlots& loop, int slot)> readPosition;
  int pending() const { return (int)(queue_.size() + running_.size()); }
  long searches() const { return searches_; }

  // `waves` waves of the side loop; searches that have finished hand their (lead, winLoss) to their jobs.  Returns the jobs in flight.
  int step(int waves) {
    if(running_.empty() && queue_.empty()) return 0;
    sp_.runWaves(waves);
    const std::vector<int32_t> visits = sp_.rootVisitsAll();
    std::vector<int> done;
    for(const auto& r : running_) if(visits[(size_t)r.first] >= maxVisits_) done.push_back(r.first);
    for(int slot : done) {
      std::vector<double> childMoments; double root[5];
      sp_.rootValueStatsByPos(slot, childMoments, root);      // winLoss, noResult, scoreMean, scoreMeanSq, lead - white's perspective
      Job* job = nullptr;
      for(auto& r : running_) if(r.first == slot) job = r.second.get();
      if(!job->positions) job->oracle.add(job->asked, root[4], root[0]);
      else {                                   // a position query: also the net's own score of the root and the legal moves
        PositionAnswer a;
        a.valid = sp_.moveNumber(slot) == (int)job->askedMoves.size();       // else the replay ended the game on the way: no such position
        if(a.valid) {
          std::vector<int32_t> nodeVisits; double nn[5];
          sp_.rootExtraByPos(slot, nodeVisits, nn);
          a.lead = root[4]; a.winLoss = root[0]; a.nnScoreMean = nn[2];
          for(float p : sp_.rootPolicy(slot)) a.legal.push_back(p >= 0 ? 1 : 0);
          if(readPosition) a.payload = readPosition(sp_, slot);
        }
        job->posOracle.add(std::move(a));
      }
      advance(slot);
    }
    dispatch();
    return pending();
  }
  void drain(int maxSteps = 100000) {
    for(int i = 0; i < maxSteps; i++) if(step(8) == 0) return;
    throw std::runtime_error("KomiSearcher: jobs did not finish");
  }

 private:
  struct Job {
    GameSlots::GameSetup setup; std::vector<Move> moves; KomiOracle oracle; Algorithm algorithm; float asked;
    bool positions; PositionOracle posOracle; PositionAlgorithm posAlgorithm; std::vector<Move> askedMoves;
  };

  // run the job's algorithm on what is known: it finishes (slot freed) or names the next komi (slot loaded with it)
  void advance(int slot) {
    size_t at = 0;
    while(running_[at].first != slot) at++;
    Job& job = *running_[at].second;
    try {
      if(job.positions) { job.posOracle.restart(); job.posAlgorithm(job.posOracle); }
      else job.algorithm(job.oracle);
    }
    catch(const NeedKomi& need) { load(slot, job, need.komi, job.moves); return; }
    catch(const NeedPosition& need) { job.askedMoves = need.moves; load(slot, job, need.komi, need.moves); return; }
    running_.erase(running_.begin() + (long)at);
    free_.push_back(slot);
  }
  void load(int slot, Job& job, float komi, const std::vector<Move>& moves) {
    job.asked = komi;
    setups_[(size_t)slot] = job.setup; komis_[(size_t)slot] = komi;
    sp_.setGameSetups(setups_); sp_.setKomis(komis_);
    bool empty = false;
    for(int i = 0; i < 4 && !empty; i++) {         // two passes end a game (three under spight ko; one if the position's last move was a pass)
      sp_.playMoves(slot, {Move()});
      empty = sp_.moveNumber(slot) == 0;
    }
    if(!empty) throw std::runtime_error("KomiSearcher: could not end the slot's previous game");
    sp_.playMoves(slot, moves);
    searches_++;
  }
  void dispatch() {
    while(!queue_.empty() && !free_.empty()) {
      const int slot = free_.back(); free_.pop_back();
      running_.emplace_back(slot, std::move(queue_.front()));
      queue_.erase(queue_.begin());
      advance(slot);
    }
  }

  GameSlots& sp_; int maxVisits_, n_;
  std::vector<int> free_;
  std::vector<std::pair<int, std::unique_ptr<Job>>> running_;      // in the order the slots were taken (the Python dict's iteration order)
  std::vector<std::unique_ptr<Job>> queue_;
  std::vector<GameSlots::GameSetup> setups_; std::vector<float> komis_;
  long searches_ = 0;
};

}  // namespace b200
