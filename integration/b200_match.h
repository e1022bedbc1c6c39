// Games between two nets on the device loops - the engine under `gatekeeper` (command/gatekeeper.cpp) and `match` (command/match.cpp;
// program/play.cpp:2712-2714 "two nets per game", MatchPairer :653-790): the stand-alone twin of katago_b200/match_play.py.
//
// Every bot has its own device loop (GameSlots in hold mode on its own handle: own net, own search parameters) with the same number of game
// slots; slot g of both loops holds THE SAME game.  Both loops search every position of it; the move comes from the loop of the bot whose turn
// it is - released when its search is finished, the device chooses and plays the move with that bot's temperature rules - and is mirrored into
// the other loop with kgb_selfplay_play_moves_game, which clears that loop's tree of the position and, when the move ended the game, starts the
// slot's next game there too.  A slot's bots swap colours from game to game.  Board size, rules and komi of every game come from a
// GameInitializer, identical for both loops.  Nothing here touches the evaluator or the search: it is host bookkeeping between waves.
#pragma once
#include <functional>

#include "b200_gameinit.h"

namespace b200 {

// The resignation check of Play::runGame after move `turnIndex` (0-based) by the mover (play.cpp:1903-1929): not before turn 1 + area / 5; the
// last resignConsecTurns root win/loss values (white's perspective, one per search so far) must each say that THIS player is lost.
inline bool shouldResign(const std::vector<double>& winLossValues, long turnIndex, int boardArea, bool moverIsBlack, double resignThreshold, int resignConsecTurns) {
  if((long)winLossValues.size() < resignConsecTurns || turnIndex < 1 + boardArea / 5) return false;
  for(size_t i = winLossValues.size() - (size_t)resignConsecTurns; i < winLossValues.size(); i++) {
    const double wl = winLossValues[i];
    int loserIsBlack = wl < resignThreshold ? 0 : wl > -resignThreshold ? 1 : -1;
    if(loserIsBlack < 0 || (loserIsBlack == 1) != moverIsBlack) return false;
  }
  return true;
}

class MatchPlay {
 public:
  struct Settings {
    long numGamesTotal = 0; double drawEquivalentWinsForWhite = 0.5, noResultUtilityForWhite = 0.0;
    bool allowResignation = false; double resignThreshold = -0.90; int resignConsecTurns = 5;
    int maxVisits[2] = {0, 0};
  };
  // (slot, the finished game as the record writer takes it, black's name, white's name, result text)
  using OnGame = std::function<void(int slot, const FinishedGame& game, const std::string& blackName, const std::string& whiteName, const std::string& result)>;
  struct Result { int blackBot, whiteBot; std::string text; size_t moves; };

  MatchPlay(GameSlots& a, GameSlots& b, const std::string& nameA, const std::string& nameB, const Settings& s, GameInitializer* init, OnGame onGame)
      : loops_{&a, &b}, names_{nameA, nameB}, s_(s), n_(a.numSlots()), init_(init), onGame_(std::move(onGame)) {
    if(a.numSlots() != b.numSlots() || a.xLen() != b.xLen() || a.yLen() != b.yLen())
      throw std::invalid_argument("MatchPlay: both loops need the same number of game slots and the same evaluator frame");
    if(s.allowResignation && !(s.resignThreshold <= 0)) throw std::invalid_argument("resignThreshold must not be positive");
    for(int g = 0; g < n_; g++) blackBot_.push_back(g % 2);
    toMove_ = blackBot_;
    moves_.resize((size_t)n_); winLoss_.resize((size_t)n_);
    gamesStarted_ = n_;
    live_.assign((size_t)n_, 1);
    if(s.numGamesTotal > 0 && s.numGamesTotal < n_) { for(int g = (int)s.numGamesTotal; g < n_; g++) live_[(size_t)g] = 0; gamesStarted_ = s.numGamesTotal; }
    if(init_) {
      drawAll();
      for(GameSlots* sp : loops_) { sp->setGameSetups(setups_, true); sp->setKomis(komis_, true); }
      drawAll();
      for(GameSlots* sp : loops_) { sp->setGameSetups(setups_); sp->setKomis(komis_); }
    }
    for(GameSlots* sp : loops_) sp->runWaves(1);
  }

  double winPoints(int bot) const { return winPoints_[bot]; }
  long gamesTallied() const { return gamesTallied_; }
  const std::vector<Result>& results() const { return results_; }
  bool done() const { return terminated_ || (s_.numGamesTotal > 0 && gamesTallied_ >= s_.numGamesTotal); }

  // `waves` waves for both loops, then every slot whose bot-to-move has finished its search moves once.  Returns moves made.
  int pump(int waves) {
    for(GameSlots* sp : loops_) sp->runWaves(waves);
    int made = 0;
    for(int b = 0; b < 2; b++) {
      GameSlots& sp = *loops_[b];
      const std::vector<int32_t> visits = sp.rootVisitsAll();
      std::vector<uint8_t> mine((size_t)n_, 0);
      bool any = false;
      for(int g = 0; g < n_; g++) if(visits[(size_t)g] >= s_.maxVisits[b] && toMove_[(size_t)g] == b) { mine[(size_t)g] = 1; any = true; }
      if(!any) continue;
      if(s_.allowResignation)          // historicalMctsWinLossValues: the root value of the search the move comes from
        for(int g = 0; g < n_; g++) if(mine[(size_t)g]) winLoss_[(size_t)g].push_back(sp.rootStats(g).winLossValueAvg);
      sp.release(mine);
      sp.runWaves(1);
      GameSlots& other = *loops_[1 - b];
      for(int g = 0; g < n_; g++) {
        if(!mine[(size_t)g]) continue;
        const GameSlots::LastMove last = sp.lastMove(g);
        other.playMoves(g, {last.move});
        moves_[(size_t)g].push_back({last.move.x, last.move.y});
        made++;
        const bool moverIsBlack = blackBot_[(size_t)g] == b;
        if(last.gameOver) finish(g, last, b, -1);
        else if(resigns(g, moverIsBlack)) {
          restart(g);
          GameSlots::LastMove ended = last;
          ended.gameOver = true; ended.noResult = false; ended.hitMoveLimit = false; ended.finalWhiteMinusBlackScore = 0.0f;
          finish(g, ended, b, moverIsBlack ? 1 : 0);
        }
        else toMove_[(size_t)g] = 1 - b;
      }
    }
    return made;
  }
  // play until numGamesTotal games are tallied (or stop(*this) says so - the gatekeeper's early termination)
  void run(int waves, const std::function<bool(const MatchPlay&)>& stop) {
    while(!done()) {
      pump(waves);
      if(stop && stop(*this)) terminated_ = true;
    }
  }

 private:
  void drawAll() {
    setups_.resize((size_t)n_); komis_.resize((size_t)n_);
    for(int g = 0; g < n_; g++) { const GameInitializer::Game d = init_->draw(); setups_[(size_t)g] = {d.x, d.y, d.koRule, d.multiStoneSuicideLegal}; komis_[(size_t)g] = d.komi; }
  }
  // what the reference's data-write loop tallies (gatekeeper.cpp:127-196)
  void tally(int blackBot, bool noResult, int winner) {
    double whitePoints;
    if(noResult) whitePoints = s_.drawEquivalentWinsForWhite;
    else if(winner == P_BLACK) whitePoints = 0.0;
    else if(winner == P_WHITE) whitePoints = 1.0;
    else whitePoints = 0.5 * s_.noResultUtilityForWhite + 0.5;
    winPoints_[blackBot] += 1.0 - whitePoints;
    winPoints_[1 - blackBot] += whitePoints;
    gamesTallied_++;
  }
  // end slot g's game in both loops without a result on the board (resignation): passes until the slot's next game has begun
  void restart(int g) {
    for(GameSlots* sp : loops_) {
      bool empty = false;
      for(int i = 0; i < 4 && !empty; i++) { sp->playMoves(g, {Move()}); empty = sp->moveNumber(g) == 0; }
      if(!empty) throw std::runtime_error("MatchPlay: could not end the resigned game");
    }
  }
  bool resigns(int g, bool moverIsBlack) {
    if(!s_.allowResignation) return false;
    std::vector<GameSlots::GameSetup> cur;
    loops_[0]->gameSetups(&cur, nullptr);
    return shouldResign(winLoss_[(size_t)g], (long)moves_[(size_t)g].size() - 1, cur[(size_t)g].x * cur[(size_t)g].y, moverIsBlack, s_.resignThreshold, s_.resignConsecTurns);
  }
  // resignedBlack: -1 no resignation, 1 black resigned, 0 white resigned
  void finish(int g, const GameSlots::LastMove& last, int mover, int resignedBlack) {
    GameSlots& sp = *loops_[mover];
    std::vector<GameSlots::GameSetup> lastSetups; std::vector<float> lastKomis;
    sp.gameSetups(nullptr, &lastSetups); sp.komis(nullptr, &lastKomis);           // the finished game's own board, rules and komi
    const GameSlots::GameSetup setup = lastSetups[(size_t)g];
    const int bb = blackBot_[(size_t)g];
    FinishedGame d;
    d.xSize = setup.x; d.ySize = setup.y; d.komi = lastKomis[(size_t)g];
    d.gameHash[0] = ((uint64_t)(g + 1) * 0x9E3779B97F4A7C15ULL + (uint64_t)last.gameIndex); d.gameHash[1] = ((uint64_t)last.gameIndex * 0xC2B2AE3D27D4EB4FULL + (uint64_t)g);
    d.mode = 0;
    d.endFinished = !last.hitMoveLimit; d.hitTurnLimit = last.hitMoveLimit; d.endNoResult = last.noResult;
    d.moves = moves_[(size_t)g];
    for(size_t i = 0; i < d.moves.size(); i++) d.nextPlayerByTurn.push_back(i % 2 == 0 ? P_BLACK : P_WHITE);
    static const char* KO[] = {"SIMPLE", "POSITIONAL", "SITUATIONAL", "SPIGHT"};
    d.koRule = KO[setup.koRule & 3]; d.multiStoneSuicideLegal = setup.multiStoneSuicideLegal != 0;
    bool noResult = false; int winner = 0; std::string text;
    char buf[64];
    if(resignedBlack >= 0) {                          // BoardHistory::setWinnerByResignation
      winner = resignedBlack ? P_WHITE : P_BLACK;
      d.winner = winner; d.resigned = true; d.endFinished = true; d.endNoResult = false;
      text = resignedBlack ? "W+R" : "B+R";
    }
    else if(d.endNoResult) { noResult = true; text = "Void"; }
    else {
      // a game stopped by the move limit is scored as it stands (gatekeeper.cpp:143-146 endAndScoreGameNow): the device has done that
      const float score = last.finalWhiteMinusBlackScore;
      winner = score > 0 ? P_WHITE : score < 0 ? P_BLACK : 0;
      d.winner = winner; d.finalWhiteMinusBlackScore = score;
      if(winner == P_WHITE) { std::snprintf(buf, sizeof(buf), "W+%g", (double)score); text = buf; }
      else if(winner == P_BLACK) { std::snprintf(buf, sizeof(buf), "B+%g", -(double)score); text = buf; }
      else text = "0";
      d.endFinished = true;
    }
    if(live_[(size_t)g]) {
      tally(bb, noResult, winner);
      results_.push_back(Result{bb, 1 - bb, text, d.moves.size()});
      if(onGame_) onGame_(g, d, names_[bb], names_[1 - bb], text);
    }
    // the slot's next game: colours swapped, fresh setup for the game after it
    moves_[(size_t)g].clear(); winLoss_[(size_t)g].clear();
    blackBot_[(size_t)g] = 1 - bb;
    toMove_[(size_t)g] = blackBot_[(size_t)g];
    if(s_.numGamesTotal > 0 && gamesStarted_ >= s_.numGamesTotal) live_[(size_t)g] = 0;
    else { live_[(size_t)g] = 1; gamesStarted_++; }
    if(init_) {
      const GameInitializer::Game nd = init_->draw();
      setups_[(size_t)g] = {nd.x, nd.y, nd.koRule, nd.multiStoneSuicideLegal}; komis_[(size_t)g] = nd.komi;
      for(GameSlots* lp : loops_) { lp->setGameSetups(setups_); lp->setKomis(komis_); }
    }
  }

  GameSlots* loops_[2]; std::string names_[2]; Settings s_; int n_; GameInitializer* init_; OnGame onGame_;
  std::vector<int> blackBot_, toMove_; std::vector<uint8_t> live_;
  std::vector<std::vector<std::pair<int, int>>> moves_; std::vector<std::vector<double>> winLoss_;
  std::vector<GameSlots::GameSetup> setups_; std::vector<float> komis_;
  long gamesStarted_ = 0, gamesTallied_ = 0; double winPoints_[2] = {0.0, 0.0};
  std::vector<Result> results_; bool terminated_ = false;
};

}  // namespace b200
