// Forked games for the C++ host: Play::maybeForkGame and the fork pool (program/play.cpp:2356-2508, ForkData :40-80, GameInitializer's
// initialPosition branch :497-526) - the stand-alone twin of katago_b200/fork_play.py.
//
// After a finished game, with probability earlyForkGameProb (else forkGameProb) a position of it is chosen - early forks an exponentially
// distributed number of moves from the start (earlyForkGameExpectedMoveProp x board area), late forks uniformly over the game - replayed, a few
// random legal moves (forkGameMinChoices .. (early)ForkGameMaxChoices, with replacement) are each evaluated by the net after being played, and the
// one the net scores best for the player to move is made.  The position goes into a pool; a game that starts while the pool is non-empty starts
// from a random entry of it.  The evaluations are position queries on a side loop (b200_komi.h KomiSearcher::submitPositions).  The draws
// come from Python's random.Random(loop seed ^ 0x466F726B) stream like the Python host's, in the same order.
#pragma once
#include "b200_komi.h"

namespace b200 {

class ForkManager {
 public:
  struct Settings {
    double earlyForkGameProb = 0.0, earlyForkGameExpectedMoveProp = 0.0, forkGameProb = 0.0, forkCompensateKomiProb = 0.0;
    int forkGameMinChoices = 1, earlyForkGameMaxChoices = 1, forkGameMaxChoices = 1;
  };
  struct Fork { std::vector<Move> moves; GameSlots::GameSetup setup; float komi; };

  ForkManager(const Settings& s, uint64_t seed) : s_(s), rand_(seed) {
    if(enabled() && s.forkGameMinChoices > std::max(s.earlyForkGameMaxChoices, s.forkGameMaxChoices)) throw std::invalid_argument("fork game max choices < forkGameMinChoices");
  }
  bool enabled() const { return s_.earlyForkGameProb > 0 || s_.forkGameProb > 0; }
  const Settings& settings() const { return s_; }
  long forksMade() const { return forksMade_; }
  long forksUsed() const { return forksUsed_; }

  // The job for KomiSearcher::submitPositions, or nothing (returns false) when this game is not forked.  allMoves: the finished game's moves from
  // the empty board; done(moves): the forked position's moves, empty when no fork came of it.
  bool job(const std::vector<Move>& allMoves, const GameSlots::GameSetup& setup, float komi, int xFrame, int yFrame,
           KomiSearcher::PositionAlgorithm& out, std::function<void(const std::vector<Move>&)> done) {
    const bool early = rand_.random() < s_.earlyForkGameProb;
    const bool late = !early && s_.forkGameProb > 0 && rand_.random() < s_.forkGameProb;
    if(!(early || late) || allMoves.empty()) return false;
    size_t idx;
    if(early) idx = (size_t)std::floor(rand_.expovariate(1.0) * s_.earlyForkGameExpectedMoveProp * setup.x * setup.y);
    else idx = rand_.randrange((uint32_t)allMoves.size());
    idx = std::min(idx, allMoves.size() - 1);                       // prior to the last move (replayGameUpToMove)
    const int lo = s_.forkGameMinChoices, hi = early ? s_.earlyForkGameMaxChoices : s_.forkGameMaxChoices;
    const int numChoices = lo + (int)rand_.randrange((uint32_t)(hi - lo + 1));                 // randint(lo, hi)
    const std::vector<Move> prefix(allMoves.begin(), allMoves.begin() + (long)idx);
    const bool blackToMove = idx % 2 == 0;
    out = [this, prefix, komi, numChoices, blackToMove, xFrame, yFrame, done](PositionOracle& ev) {
      const PositionAnswer& root = ev(prefix, komi);
      if(!root.valid) { done({}); return; }
      std::vector<int> legal;
      for(size_t p = 0; p < root.legal.size(); p++) if(root.legal[p]) legal.push_back((int)p);
      if(legal.empty()) { done({}); return; }
      const int passPos = xFrame * yFrame;
      bool haveBest = false; Move best; double bestScore = 0.0;
      for(int i = 0; i < numChoices; i++) {                         // chooseRandomLegalMoves: with replacement, the pass included
        const int pos = legal[ev.draw([&]() { return rand_.randrange((uint32_t)legal.size()); })];
        Move mv; if(pos != passPos) { mv.x = pos % xFrame; mv.y = pos / xFrame; }
        std::vector<Move> after = prefix; after.push_back(mv);
        const PositionAnswer& a = ev(after, komi);
        if(!a.valid) continue;                                      // (that move ended the game)
        if(!haveBest || (!blackToMove && a.nnScoreMean > bestScore) || (blackToMove && a.nnScoreMean < bestScore)) { haveBest = true; best = mv; bestScore = a.nnScoreMean; }
      }
      if(!haveBest) { done({}); return; }
      std::vector<Move> forked = prefix; forked.push_back(best);
      done(forked);
    };
    return true;
  }
  void add(const std::vector<Move>& moves, const GameSlots::GameSetup& setup, float komi) { pool_.push_back(Fork{moves, setup, komi}); forksMade_++; }
  // ForkData::get: a random entry of the pool (removed)
  bool pop(Fork& out) {
    if(pool_.empty()) return false;
    const size_t i = rand_.randrange((uint32_t)pool_.size());
    std::swap(pool_[i], pool_.back());
    forksUsed_++;
    out = std::move(pool_.back()); pool_.pop_back();
    return true;
  }
  double uniform() { return rand_.random(); }

 private:
  Settings s_; PyRandom rand_; std::vector<Fork> pool_; long forksMade_ = 0, forksUsed_ = 0;
};

}  // namespace b200
