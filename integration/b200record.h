// Fills the reference's own FinishedGameData (dataio/trainingwrite.h:84-170) from the device game slots, so that the reference's
// TrainingDataWriter::writeGame / addRow / writeToZipFile run UNCHANGED on games played by libkgb200.
//
// This is reference-side glue: it includes the reference's headers and is compiled inside the reference tree next to
// b200backend.cpp.  It replaces the bookkeeping Play::runGame does around its Search (program/play.cpp:1757-2027):
//   extractSearchTargetsThisTurn (:931-948)  -> per-turn value / Q / policy targets, surprise and entropies from the slot's statistics
//   the move and the game-end block (:1938-2027) -> the device's move replayed on the reference's own Board / BoardHistory (which
//                                                   also re-checks legality, game end and the final score against the device's)
// Not filled (options the device loop does not have): side positions, cheap searches and their target weights (every turn has
// weight 1), lead estimation, reanalysis, net changes.  Rules: area scoring, no tax, no button, no handicap bonus.
#pragma once
#include <cmath>
#include <functional>
#include <memory>
#include <sstream>

#include "b200selfplay.h"
#include "dataio/trainingwrite.h"
#include "game/board.h"
#include "game/boardhistory.h"
#include "neuralnet/nninputs.h"

namespace b200 {

// ReportedSearchValues (search/reportedsearchvalues.cpp:10-51) without a Search: the clamps that turn NodeStats moments into values
struct ReportedValues { double winValue, lossValue, noResultValue, winLossValue, expectedScore; };
inline ReportedValues reportedValues(const GameSlots::ValueStats& s) {
  ReportedValues r;
  r.winLossValue = std::min(std::max(s.winLossValueAvg, -1.0), 1.0);
  r.noResultValue = std::min(std::max(s.noResultValueAvg, 0.0), 1.0 - std::fabs(r.winLossValue));
  r.winValue = std::min(std::max(0.5 * (r.winLossValue + (1.0 - r.noResultValue)), 0.0), 1.0);
  r.lossValue = std::min(std::max(0.5 * (-r.winLossValue + (1.0 - r.noResultValue)), 0.0), 1.0);
  r.expectedScore = s.scoreMeanAvg;
  return r;
}

class GameRecorder {
 public:
  using OnGame = std::function<void(int slot, FinishedGameData* data)>;   // takes ownership of data

  GameRecorder(GameSlots& slots, const Rules& rules, int maxVisits, double drawEquivalentWinsForWhite, OnGame onGame)
    : slots_(slots), rules_(rules), maxVisits_(maxVisits), drawEq_(drawEquivalentWinsForWhite), onGame_(std::move(onGame)) {
    for(int i = 0; i < slots_.numSlots(); i++) { games_.emplace_back(); startGame(i, 0); }
    slots_.runWaves(1);   // evaluates every root (its NN input row is the root's fillRowV7 row; the reference's addRow recomputes it)
  }

  int64_t gamesFinished() const { return gamesFinished_; }

  // One move of every slot: finish the searches, record this turn's targets, let the device move, replay the move here.
  void step() {
    while(!slots_.allHeld(maxVisits_)) slots_.runWaves(8);
    const int X = slots_.xLen(), Y = slots_.yLen();
    for(int i = 0; i < slots_.numSlots(); i++) {
      Slot& g = games_[i];
      FinishedGameData* d = g.data.get();
      // extractValueTargets (play.cpp:848-857)
      const ReportedValues rv = reportedValues(slots_.rootStats(i));
      ValueTargets vt;
      vt.win = (float)rv.winValue; vt.loss = (float)rv.lossValue; vt.noResult = (float)rv.noResultValue; vt.score = (float)rv.expectedScore;
      d->whiteValueTargetsByTurn.push_back(vt);
      // extractQValueTargets (play.cpp:859-888)
      QValueTargets q;
      for(const GameSlots::ChildStats& c : slots_.rootChildren(i)) {
        if(c.nodeVisits <= 0) continue;
        GameSlots::ValueStats s; s.winLossValueAvg = c.winLossValueAvg; s.noResultValueAvg = c.noResultValueAvg; s.scoreMeanAvg = c.scoreMeanAvg;
        const ReportedValues cv = reportedValues(s);
        q.targets.emplace_back(locOf(c.move, X), (float)cv.winLossValue, (float)cv.expectedScore, (int64_t)c.nodeVisits);
      }
      d->whiteQValueTargetsByTurn.push_back(q);
      // Play::extractPolicyTarget (play.cpp:810-846): play selection values scaled so that the largest is at least 10, at most 30000
      const std::vector<std::pair<Move, double>> psv = slots_.getPlaySelectionValues(i);
      double maxValue = 0.0, sumValues = 0.0;
      for(const auto& mv : psv) { maxValue = std::max(maxValue, mv.second); sumValues += mv.second; }
      double factor = (maxValue > 0.0 && maxValue < 10.0) ? 10.0 / maxValue : 1.0;
      if(maxValue * factor > 30000.0) factor = 30000.0 / maxValue;
      std::vector<PolicyTargetMove>* pt = new std::vector<PolicyTargetMove>();
      for(const auto& mv : psv) pt->emplace_back(locOf(mv.first, X), (int16_t)std::round(mv.second * factor));
      d->policyTargetsByTurn.push_back(PolicyTarget(pt, slots_.getRootVisits(i)));
      // Search::getPolicySurpriseAndEntropy (searchresults.cpp:631-695)
      const std::vector<float> policy = slots_.rootPolicy(i);
      double surprise = 0.0, searchEntropy = 0.0, policyEntropy = 0.0;
      for(const auto& mv : psv) {
        const int pos = mv.first.isPass() ? X * Y : mv.first.y * X + mv.first.x;
        const double p = std::max((double)policy[pos], 1e-100), target = mv.second / sumValues;
        if(target > 1e-100) { surprise += target * (std::log(target) - std::log(p)); searchEntropy += -target * std::log(target); }
      }
      for(float p : policy) if(p > 1e-100) policyEntropy += -(double)p * std::log((double)p);
      d->policySurpriseByTurn.push_back(std::max(surprise, 0.0));
      d->searchEntropyByTurn.push_back(std::max(searchEntropy, 0.0));
      d->policyEntropyByTurn.push_back(std::max(policyEntropy, 0.0));
      // computeNNRawStats (play.cpp:890-914) from the root's own evaluation
      const GameSlots::ValueStats nn = slots_.rootNNStats(i);
      NNRawStats raw; raw.whiteWinLoss = nn.winLossValueAvg; raw.whiteScoreMean = nn.scoreMeanAvg; raw.policyEntropy = std::max(slots_.rootRawPolicyEntropy(i), 0.0);
      d->nnRawStatsByTurn.push_back(raw);
      d->targetWeightByTurn.push_back(1.0f);
      d->targetWeightByTurnUnrounded.push_back(1.0f);
    }
    slots_.release();
    slots_.runWaves(1);
    for(int i = 0; i < slots_.numSlots(); i++) {
      Slot& g = games_[i];
      const GameSlots::LastMove lm = slots_.lastMove(i);
      const Loc loc = locOf(lm.move, X);
      if(!g.hist.isLegal(g.board, loc, g.pla)) fail(i, "the device played a move the reference's BoardHistory rejects", lm);
      g.hist.makeBoardMoveAssumeLegal(g.board, loc, g.pla, NULL);
      g.pla = getOpp(g.pla);
      if(g.hist.isGameFinished != (lm.gameOver && !lm.hitMoveLimit)) fail(i, "game end disagrees with the reference's BoardHistory", lm);
      if(lm.gameOver) finishGame(i, lm);
    }
  }

 private:
  struct Slot { std::unique_ptr<FinishedGameData> data; Board board; BoardHistory hist; Player pla = P_BLACK; };

  static Loc locOf(const Move& m, int X) { return m.isPass() ? Board::PASS_LOC : Location::getLoc(m.x, m.y, X); }

  void startGame(int i, int gameIndex) {
    Slot& g = games_[i];
    g.board = Board(slots_.xLen(), slots_.yLen());
    g.pla = P_BLACK;
    g.hist = BoardHistory(g.board, g.pla, rules_, 0, false);
    g.data.reset(new FinishedGameData());
    FinishedGameData* d = g.data.get();
    d->startBoard = g.board; d->startHist = g.hist; d->startPla = g.pla;
    // FinishedGameData::gameHash is two draws of the game's own Rand in the reference; any per-game constant serves
    d->gameHash = Hash128((uint64_t)(i + 1) * 0x9E3779B97F4A7C15ULL + (uint64_t)gameIndex, (uint64_t)(gameIndex + 1) * 0xC2B2AE3D27D4EB4FULL + (uint64_t)i);
    d->drawEquivalentWinsForWhite = drawEq_;
    d->playoutDoublingAdvantagePla = C_EMPTY; d->playoutDoublingAdvantage = 0.0;
    d->numExtraBlack = 0; d->mode = FinishedGameData::MODE_NORMAL; d->trainingWeight = 1.0;
  }

  // the game-end block of Play::runGame (play.cpp:1938-2027), with the device's final position and score checked against the reference's
  void finishGame(int i, const GameSlots::LastMove& lm) {
    Slot& g = games_[i];
    FinishedGameData* d = g.data.get();
    const int X = slots_.xLen(), Y = slots_.yLen();
    d->endHist = g.hist;
    d->hitTurnLimit = !g.hist.isGameFinished;
    d->finalFullArea = new Color[Board::MAX_ARR_SIZE];
    d->finalOwnership = new Color[Board::MAX_ARR_SIZE];
    d->finalSekiAreas = new bool[Board::MAX_ARR_SIZE];
    d->finalWhiteScoring = new float[Board::MAX_ARR_SIZE];
    std::fill(d->finalSekiAreas, d->finalSekiAreas + Board::MAX_ARR_SIZE, false);
    ValueTargets fin;
    if(g.hist.isGameFinished && g.hist.isNoResult) {
      if(!lm.noResult) fail(i, "the reference ended the game without result, the device did not", lm);
      fin.win = 0.0f; fin.loss = 0.0f; fin.noResult = 1.0f; fin.score = 0.0f;
      std::fill(d->finalFullArea, d->finalFullArea + Board::MAX_ARR_SIZE, C_EMPTY);
      std::fill(d->finalOwnership, d->finalOwnership + Board::MAX_ARR_SIZE, C_EMPTY);
    }
    else {
      if(lm.noResult) fail(i, "the device ended the game without result, the reference did not", lm);
      g.hist.endAndScoreGameNow(g.board, d->finalOwnership);
      fin.win = (float)ScoreValue::whiteWinsOfWinner(g.hist.winner, drawEq_);
      fin.loss = 1.0f - fin.win; fin.noResult = 0.0f;
      fin.score = (float)ScoreValue::whiteScoreDrawAdjust(g.hist.finalWhiteMinusBlackScore, drawEq_, g.hist);
      fin.hasLead = true; fin.lead = fin.score;
      g.board.calculateArea(d->finalFullArea, true, true, true, g.hist.suicideLegalForPassAlive());
      // the device's own final position, area and score must be the reference's
      if(g.hist.finalWhiteMinusBlackScore != lm.finalWhiteMinusBlackScore) fail(i, "final score differs from the reference's", lm);
      for(int y = 0; y < Y; y++) for(int x = 0; x < X; x++) {
        const Loc l = Location::getLoc(x, y, X);
        if(g.board.colors[l] != lm.finalColors[(size_t)y * X + x] || d->finalOwnership[l] != lm.finalArea[(size_t)y * X + x])
          fail(i, "final position or area differs from the reference's", lm);
      }
    }
    d->whiteValueTargetsByTurn.push_back(fin);
    NNInputs::fillScoring(g.board, d->finalOwnership, false, d->finalWhiteScoring);
    d->hasFullData = true;
    gamesFinished_++;
    FinishedGameData* out = g.data.release();
    startGame(i, lm.gameIndex + 1);
    onGame_(i, out);
  }

  [[noreturn]] void fail(int slot, const char* what, const GameSlots::LastMove& lm) const {
    std::ostringstream s;
    s << "b200::GameRecorder slot " << slot << " move " << lm.moveNumber << " (" << lm.move.x << "," << lm.move.y << "): " << what;
    throw std::runtime_error(s.str());
  }

  GameSlots& slots_;
  Rules rules_;
  int maxVisits_;
  double drawEq_;
  OnGame onGame_;
  std::vector<Slot> games_;
  int64_t gamesFinished_ = 0;
};

}  // namespace b200
