// Search-shaped C++ facade over the device-resident game slots of libkgb200 (boundary 2).
//
// The reference drives one `Search` object per game from a game thread (program/play.cpp:1757-1936):
//     bot->setPosition(pla, board, hist);  bot->runWholeSearch(pla);  Loc loc = bot->getChosenMoveLoc();  bot->makeMove(loc, pla);
// Here all games of a GPU advance together: `runWaves` spends playouts on every slot, each slot chooses and plays its move when its
// visit budget is used up, and restarts when its game ends.  The readers below have the names and meanings of the reference's
// (search/search.h: getRootVisits, getPlaySelectionValues; search/searchresults.cpp) so that the game-recording code above them can
// stay as it is.  Plain C++17, no CUDA headers: everything goes through the C ABI of include/kgb200.h.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/kgb200.h"

namespace b200 {

struct Move { int x = -1, y = -1; bool isPass() const { return x < 0; } };

class GameSlots {
 public:
  // handle: an evaluator handle created with inputs_nhwc = 1 (NeuralNet::createComputeHandle in b200backend.cpp); cfg: SearchParams /
  // Rules fields by their cfg names.
  GameSlots(kgb_handle* handle, const kgb_selfplay_config& cfg, int xLen, int yLen) : x_(xLen), y_(yLen), n_(cfg.num_games) {
    check(kgb_selfplay_create(handle, &cfg, &sp_));
    handle_ = handle;
  }
  ~GameSlots() { if(sp_) kgb_selfplay_free(sp_); }
  GameSlots(const GameSlots&) = delete;
  GameSlots& operator=(const GameSlots&) = delete;

  int numSlots() const { return n_; }
  // Search::setPosition + makeMove for an opening common to all slots (x, y pairs; pass = -1, -1)
  void playOpening(const std::vector<Move>& moves) {
    std::vector<int8_t> xy;
    for(const Move& m : moves) { xy.push_back((int8_t)m.x); xy.push_back((int8_t)m.y); }
    check(kgb_selfplay_play_moves(sp_, xy.data(), (int)moves.size()));
  }
  // the same for one slot (games of different board sizes; the opponent's move in match play): a move that ends the game restarts the slot
  void playMoves(int slot, const std::vector<Move>& moves) {
    std::vector<int8_t> xy;
    for(const Move& m : moves) { xy.push_back((int8_t)m.x); xy.push_back((int8_t)m.y); }
    check(kgb_selfplay_play_moves_game(sp_, slot, xy.data(), (int)moves.size()));
  }
  // GameInitializer::createGame's draws for the slots' next games (play.cpp:330-650): board X, board Y, ko rule, multi-stone suicide per slot; komi per slot
  struct GameSetup { int32_t x, y, koRule, multiStoneSuicideLegal; };
  void setGameSetups(const std::vector<GameSetup>& setups, bool alsoCurrentGames = false) {
    if((int)setups.size() != n_) throw std::invalid_argument("setGameSetups: one entry per slot");
    check(kgb_selfplay_set_game_setup(sp_, &setups[0].x, alsoCurrentGames ? 1 : 0));
  }
  std::vector<GameSetup> currentGameSetups() const { std::vector<GameSetup> v((size_t)n_); check(kgb_selfplay_get_game_setup(sp_, &v[0].x, nullptr)); return v; }
  void setKomis(const std::vector<float>& komis, bool alsoCurrentGames = false) {
    if((int)komis.size() != n_) throw std::invalid_argument("setKomis: one entry per slot");
    check(kgb_selfplay_set_komi(sp_, komis.data(), alsoCurrentGames ? 1 : 0));
  }
  // getSearchLimitsThisMove (play.cpp:1093-1223): visit budget and "plain root" of the root after each slot's next move, [slot][0] the game goes on, [slot][1] it ends
  void setNextSearchLimits(const std::vector<int32_t>& visits2, const std::vector<uint8_t>& plain2, bool alsoCurrentRoots = false) {
    if((int)visits2.size() != 2 * n_ || (!plain2.empty() && (int)plain2.size() != 2 * n_)) throw std::invalid_argument("setNextSearchLimits: two entries per slot");
    check(kgb_selfplay_set_next_search_limits(sp_, visits2.data(), plain2.empty() ? nullptr : plain2.data(), alsoCurrentRoots ? 1 : 0));
  }
  std::vector<int32_t> visitBudgets() const { std::vector<int32_t> v((size_t)n_); check(kgb_selfplay_get_search_limits(sp_, v.data(), nullptr)); return v; }
  // PlayUtils::initializeGameUsingPolicy (playutils.cpp:232-266): opening moves each slot's next game draws from the raw policy
  void setPolicyInit(const std::vector<int32_t>& numMoves, double temperature, bool alsoCurrentGames = false) {
    if((int)numMoves.size() != n_) throw std::invalid_argument("setPolicyInit: one entry per slot");
    check(kgb_selfplay_set_policy_init(sp_, numMoves.data(), temperature, alsoCurrentGames ? 1 : 0));
  }
  // moves left in each slot's opening (> 0: the slot moves on by itself and is not held) and, if asked for, the opening played so far
  std::vector<int32_t> policyInitState(std::vector<std::vector<Move>>* openings = nullptr) const {
    const int maxMoves = 512;
    std::vector<int32_t> left((size_t)n_), count((size_t)n_);
    std::vector<int16_t> mv(openings ? (size_t)n_ * maxMoves : 0);
    check(kgb_selfplay_get_policy_init(sp_, left.data(), count.data(), openings ? mv.data() : nullptr, openings ? maxMoves : 0));
    if(openings) {
      openings->assign((size_t)n_, {});
      for(int g = 0; g < n_; g++)
        for(int i = 0; i < std::min(count[(size_t)g], maxMoves); i++) {
          const int p = mv[(size_t)g * maxMoves + i];
          Move m; if(p < x_ * y_) { m.x = p % x_; m.y = p / x_; }
          (*openings)[(size_t)g].push_back(m);
        }
    }
    return left;
  }
  // the playout loop: `waves` playout waves for every slot (asynchronous), then wait
  void runWaves(int waves) { check(kgb_selfplay_run(sp_, waves)); check(kgb_handle_sync(handle_)); }

  // Search::getRootVisits
  int64_t getRootVisits(int slot) const { return gameInfo(slot)[5]; }
  int moveNumber(int slot) const { return gameInfo(slot)[0]; }
  bool blackToMove(int slot) const { return gameInfo(slot)[1] != 0; }
  // root board, row-major [y][x]: 0 empty, 1 black, 2 white
  std::vector<uint8_t> rootBoard(int slot) const {
    std::vector<uint8_t> colors((size_t)x_ * y_);
    int32_t info[6];
    check(kgb_selfplay_get_game(sp_, slot, colors.data(), info));
    return colors;
  }
  // Search::getPlaySelectionValues: (move, value) for every root child, in move-position order
  std::vector<std::pair<Move, double>> getPlaySelectionValues(int slot) const {
    std::vector<double> v((size_t)x_ * y_ + 1);
    check(kgb_selfplay_get_play_selection_values(sp_, slot, v.data()));
    std::vector<std::pair<Move, double>> out;
    for(size_t i = 0; i < v.size(); i++)
      if(v[i] >= 0.0) { Move m; if(i < (size_t)x_ * y_) { m.x = (int)(i % x_); m.y = (int)(i / x_); } out.emplace_back(m, v[i]); }
    return out;
  }
  // per root child: edge visits, policy prior (the noised one at the root), utilityAvg - what ReportedSearchValues are built from
  struct ChildStats { Move move; int visits; int nodeVisits; float prior; double utilityAvg, winLossValueAvg, noResultValueAvg, scoreMeanAvg, scoreMeanSqAvg, leadAvg; };
  std::vector<ChildStats> rootChildren(int slot) const {
    const size_t ps = (size_t)x_ * y_ + 1;
    std::vector<int32_t> visits(ps); std::vector<float> policy(ps); std::vector<double> util(ps), mom(ps * 5), rootMom(5);
    check(kgb_selfplay_get_root_children(sp_, slot, visits.data(), policy.data(), util.data()));
    check(kgb_selfplay_get_root_value_stats(sp_, slot, mom.data(), rootMom.data()));
    const std::vector<int32_t> nodeVisits = childNodeVisits(slot);
    std::vector<ChildStats> out;
    for(size_t i = 0; i < ps; i++) {
      if(visits[i] <= 0) continue;
      ChildStats c;
      if(i < ps - 1) { c.move.x = (int)(i % x_); c.move.y = (int)(i / x_); }
      c.visits = visits[i]; c.nodeVisits = nodeVisits[i]; c.prior = policy[i]; c.utilityAvg = util[i];
      c.winLossValueAvg = mom[i * 5]; c.noResultValueAvg = mom[i * 5 + 1]; c.scoreMeanAvg = mom[i * 5 + 2]; c.scoreMeanSqAvg = mom[i * 5 + 3]; c.leadAvg = mom[i * 5 + 4];
      out.push_back(c);
    }
    return out;
  }
  // with kgb_handle_commit_weights: the cached evaluations belong to the previous net
  void clearNNCache() { check(kgb_selfplay_clear_nn_cache(sp_)); }
  kgb_selfplay_stats stats() const { kgb_selfplay_stats s; check(kgb_selfplay_get_stats(sp_, &s)); return s; }

  // ---- game recording (kgb_selfplay_config.debug_hold_at_max_visits = 1): a slot whose search is finished idles until release() ----
  std::vector<int32_t> rootVisitsAll() const { std::vector<int32_t> v((size_t)n_); check(kgb_selfplay_get_root_visits(sp_, v.data())); return v; }
  bool allHeld(int maxVisits) const { for(int32_t v : rootVisitsAll()) if(v < maxVisits) return false; return true; }
  void release() { check(kgb_selfplay_release(sp_, nullptr)); }
  void release(const std::vector<uint8_t>& mask) { if((int)mask.size() != n_) throw std::invalid_argument("release: one entry per slot"); check(kgb_selfplay_release(sp_, mask.data())); }
  // NodeStats moments of the root (white's perspective): what Search::getNodeValues(rootNode) is built from; and of the root's own evaluation
  struct ValueStats { double winLossValueAvg = 0, noResultValueAvg = 0, scoreMeanAvg = 0, scoreMeanSqAvg = 0, leadAvg = 0; };
  ValueStats rootStats(int slot) const {
    std::vector<double> mom(((size_t)x_ * y_ + 1) * 5), r(5);
    check(kgb_selfplay_get_root_value_stats(sp_, slot, mom.data(), r.data()));
    return ValueStats{r[0], r[1], r[2], r[3], r[4]};
  }
  ValueStats rootNNStats(int slot) const {
    std::vector<int32_t> nv((size_t)x_ * y_ + 1); double r[5];
    check(kgb_selfplay_get_root_extra(sp_, slot, nv.data(), r));
    return ValueStats{r[0], r[1], r[2], r[3], r[4]};
  }
  // entropy of the root's policy as the net gave it, before temperature and noise (NNRawStats::policyEntropy, play.cpp:890-914)
  double rootRawPolicyEntropy(int slot) const {
    std::vector<double> e((size_t)numSlots());
    check(kgb_selfplay_get_root_raw_policy_entropy(sp_, e.data()));
    return e[(size_t)slot];
  }
  // visits of the root's child nodes by move position (0 = no child; not the edge visits under graph search)
  std::vector<int32_t> childNodeVisits(int slot) const {
    std::vector<int32_t> nv((size_t)x_ * y_ + 1); double r[5];
    check(kgb_selfplay_get_root_extra(sp_, slot, nv.data(), r));
    return nv;
  }
  // root policy by move position as searched (temperature and noise applied), -1 = illegal: NNOutput::getPolicyProbsMaybeNoised
  std::vector<float> rootPolicy(int slot) const {
    const size_t ps = (size_t)x_ * y_ + 1;
    std::vector<int32_t> visits(ps); std::vector<float> policy(ps); std::vector<double> util(ps);
    check(kgb_selfplay_get_root_children(sp_, slot, visits.data(), policy.data(), util.data()));
    return policy;
  }
  // the NN input row of the slot's last wave (NHWC [y*x][22] and 19 globals)
  void inputRow(int slot, std::vector<float>& spatial, std::vector<float>& global) const {
    spatial.resize((size_t)x_ * y_ * 22); global.resize(19);
    check(kgb_selfplay_get_nn_row(sp_, slot, spatial.data(), global.data()));
  }
  // the NN input row of the slot's current root, kept on the device since the wave that evaluated it
  void rootInputRow(int slot, std::vector<float>& spatial, std::vector<float>& global) const {
    spatial.resize((size_t)x_ * y_ * 22); global.resize(19);
    check(kgb_selfplay_get_root_row(sp_, slot, spatial.data(), global.data()));
  }
  struct LastMove {
    Move move; bool gameOver = false, noResult = false, hitMoveLimit = false; int moveNumber = 0, gameIndex = 0;
    float finalWhiteMinusBlackScore = 0;           // komi included
    std::vector<uint8_t> finalColors, finalArea;   // [y][x]: 0 none, 1 black, 2 white (valid when gameOver)
  };
  LastMove lastMove(int slot) const {
    LastMove m; int32_t info[4];
    m.finalColors.resize((size_t)x_ * y_); m.finalArea.resize((size_t)x_ * y_);
    check(kgb_selfplay_get_last_move(sp_, slot, info, &m.finalWhiteMinusBlackScore, m.finalColors.data(), m.finalArea.data()));
    if(info[0] < x_ * y_) { m.move.x = info[0] % x_; m.move.y = info[0] / x_; }
    m.gameOver = info[1] & 1; m.noResult = info[1] & 2; m.hitMoveLimit = info[1] & 4; m.moveNumber = info[2]; m.gameIndex = info[3];
    return m;
  }
  int xLen() const { return x_; }
  int yLen() const { return y_; }

  // ---- the same readers as plain arrays by move position (0 .. X*Y, pass last), for a recorder without the reference's types
  //      (integration/b200_recorder.h) ----
  struct RootPosition { std::vector<uint8_t> colors; int moveNumber = 0; bool blackToMove = true; int64_t rootVisits = 0; };
  RootPosition rootPosition(int slot) const {
    RootPosition r; r.colors.resize((size_t)x_ * y_);
    int32_t info[6];
    check(kgb_selfplay_get_game(sp_, slot, r.colors.data(), info));
    r.moveNumber = info[0]; r.blackToMove = info[1] != 0; r.rootVisits = info[5];
    return r;
  }
  std::vector<double> playSelectionValuesByPos(int slot) const {      // -1 = no child
    std::vector<double> v((size_t)x_ * y_ + 1);
    check(kgb_selfplay_get_play_selection_values(sp_, slot, v.data()));
    return v;
  }
  void rootValueStatsByPos(int slot, std::vector<double>& childMoments /* [X*Y+1][5] */, double rootMoments[5]) const {
    childMoments.resize(((size_t)x_ * y_ + 1) * 5);
    check(kgb_selfplay_get_root_value_stats(sp_, slot, childMoments.data(), rootMoments));
  }
  void rootExtraByPos(int slot, std::vector<int32_t>& childNodeVisitsOut, double rootNNMoments[5]) const {
    childNodeVisitsOut.resize((size_t)x_ * y_ + 1);
    check(kgb_selfplay_get_root_extra(sp_, slot, childNodeVisitsOut.data(), rootNNMoments));
  }
  // board, rules and komi of the games in progress and of each slot's last finished game
  void gameSetups(std::vector<GameSetup>* current, std::vector<GameSetup>* lastFinished) const {
    if(current) current->resize((size_t)n_);
    if(lastFinished) lastFinished->resize((size_t)n_);
    check(kgb_selfplay_get_game_setup(sp_, current ? &(*current)[0].x : nullptr, lastFinished ? &(*lastFinished)[0].x : nullptr));
  }
  void komis(std::vector<float>* current, std::vector<float>* lastFinished) const {
    if(current) current->resize((size_t)n_);
    if(lastFinished) lastFinished->resize((size_t)n_);
    check(kgb_selfplay_get_komi(sp_, current ? current->data() : nullptr, lastFinished ? lastFinished->data() : nullptr));
  }
  std::vector<double> rootRawPolicyEntropies() const { std::vector<double> e((size_t)n_); check(kgb_selfplay_get_root_raw_policy_entropy(sp_, e.data())); return e; }

 private:
  struct Info { int32_t v[6]; int32_t operator[](int i) const { return v[i]; } };
  Info gameInfo(int slot) const {
    std::vector<uint8_t> colors((size_t)x_ * y_);
    Info info;
    check(kgb_selfplay_get_game(sp_, slot, colors.data(), info.v));
    return info;
  }
  static void check(int rc) { if(rc != 0) throw std::runtime_error(std::string("libkgb200: ") + kgb_last_error()); }
  kgb_selfplay* sp_ = nullptr;
  kgb_handle* handle_ = nullptr;
  int x_, y_, n_;
};

}  // namespace b200
