// Game recorder of the C++ host: device game slots in hold mode -> FinishedGame (integration/b200_npz.h), without any reference header.
//
// The stand-alone twin of katago_b200/game_recorder.py (GameRecorder.pump / _record_root / _after_move / _finish_game, search_limits_this_move): what Play::runGame does around its Search (program/play.cpp:1757-2163) once search, rules and features live on the device -
//   extractSearchTargetsThisTurn (:931-948): value targets (ReportedSearchValues of the root's NodeStats), Q targets (child nodes), policy
//     target (Play::extractPolicyTarget :810-846 on the play selection values), policy surprise and entropies
//     (Search::getPolicySurpriseAndEntropy, search/searchresults.cpp:631-695), NNRawStats (:890-914);
//   the game-end block (:1964-2027): outcome value targets, ownership / area / scoring planes from the device's final area;
//   surprise weighting (:2034-2163: computeValueSurpriseByTurn, policy- and value-surprise redistribution of the turn weights) and
//     resolveWeight (:2274-2289);
//   getSearchLimitsThisMove (:1093-1223): cheap searches and reduced visits, drawn one root ahead and applied by the device.
// (integration/b200record.h is the reference-side variant of this file: it fills the reference's own FinishedGameData.)
// Parity: tests/test_cpp_host.py runs the C++ host against a CPU mock of the ABI and compares its .npz rows and .sgfs records with the Python
// recorder's on the same games, bit for bit.
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <memory>

#include "b200_npz.h"
#include "b200selfplay.h"

namespace b200 {

// ReportedSearchValues (search/reportedsearchvalues.cpp:10-51) from NodeStats moments, white's perspective
struct Reported { double win, loss, noResult, winLoss, score; };
inline Reported reportedSearchValues(const double* m) {
  Reported r;
  r.winLoss = std::min(std::max(m[0], -1.0), 1.0);
  r.noResult = std::min(std::max(m[1], 0.0), 1.0 - std::fabs(r.winLoss));
  r.win = std::min(std::max(0.5 * (r.winLoss + (1.0 - r.noResult)), 0.0), 1.0);
  r.loss = std::min(std::max(0.5 * (-r.winLoss + (1.0 - r.noResult)), 0.0), 1.0);
  r.score = m[2];
  return r;
}

// Play::extractPolicyTarget (play.cpp:810-846): scaleMaxToAtLeast = 10, cap at 30000, C round(), int16 - for every child, in position order
inline std::vector<PolicyTargetMove> policyTargetMoves(const std::vector<double>& psv, int xLen) {
  const size_t n = psv.size();
  std::vector<double> v(n);
  double mx = 0.0;
  for(size_t i = 0; i < n; i++) { v[i] = psv[i] > 0 ? psv[i] : 0.0; mx = std::max(mx, v[i]); }
  if(mx > 0 && mx < 10.0) { const double f = 10.0 / std::max(mx, 1e-300); mx = 0.0; for(double& x : v) { x *= f; mx = std::max(mx, x); } }
  if(mx > 30000.0) { const double f = 30000.0 / std::max(mx, 1e-300); for(double& x : v) x *= f; }
  std::vector<PolicyTargetMove> out;
  for(size_t pos = 0; pos < n; pos++)
    if(psv[pos] >= 0) out.push_back(PolicyTargetMove{pos == n - 1 ? -1 : (int)(pos % xLen), pos == n - 1 ? -1 : (int)(pos / xLen), (int16_t)std::floor(v[pos] + 0.5)});
  return out;
}

// Search::getPolicySurpriseAndEntropy: KL(target || policy), entropy of the target, entropy of the policy
inline void policySurpriseAndEntropy(const std::vector<double>& psv, const std::vector<float>& policy, double& surprise, double& searchEntropy, double& policyEntropy) {
  double total = 0.0;
  for(size_t i = 0; i < psv.size(); i++) if(psv[i] >= 0) total += psv[i];
  surprise = searchEntropy = policyEntropy = 0.0;
  for(size_t i = 0; i < psv.size(); i++) {
    if(psv[i] < 0) continue;
    const double p = std::max((double)policy[i], 1e-100), target = psv[i] / total;
    if(target > 1e-100) { const double lt = std::log(target); surprise += target * (lt - std::log(p)); searchEntropy += -target * lt; }
  }
  for(float pf : policy) { const double p = pf; if(p > 1e-100) policyEntropy += -p * std::log(p); }
  surprise = std::max(surprise, 0.0); searchEntropy = std::max(searchEntropy, 0.0); policyEntropy = std::max(policyEntropy, 0.0);
}

// valueSurpriseKL (play.cpp:1303-1314)
inline double valueSurpriseKL(double win, double loss, double noResult, const std::array<double, 3>& raw) {
  double s = 0.0;
  const double v[3] = {win, loss, noResult};
  for(int i = 0; i < 3; i++) if(v[i] > 1e-100) s += v[i] * (std::log(v[i]) - std::log(std::max(raw[i], 1e-100)));
  return std::min(std::max(s, 0.0), 1.0);
}

// computeValueSurpriseByTurn (play.cpp:1322-1352)
inline std::vector<double> computeValueSurpriseByTurn(const std::vector<ValueTargets>& vt, const std::vector<std::array<double, 3>>& rawNN, int boardArea, bool useSearchValueSurprise) {
  const size_t n = rawNN.size();
  std::vector<double> out(n, 0.0);
  if(useSearchValueSurprise) { for(size_t i = 0; i < n; i++) out[i] = valueSurpriseKL(vt[i].win, vt[i].loss, vt[i].noResult, rawNN[i]); return out; }
  const double now = 1.0 / (1.0 + boardArea * 0.016);
  double win = vt.back().win, loss = vt.back().loss, nores = vt.back().noResult;
  for(size_t k = n; k-- > 0;) {
    win += now * ((double)vt[k].win - win); loss += now * ((double)vt[k].loss - loss); nores += now * ((double)vt[k].noResult - nores);
    out[k] = valueSurpriseKL(win, loss, nores, rawNN[k]);
  }
  return out;
}

// The surprise weighting of Play::runGame (play.cpp:2084-2163) for games without cheap-search reanalysis
inline std::vector<float> surpriseTargetWeights(const std::vector<float>& targetWeights, const std::vector<double>& policySurprise, const std::vector<double>& valueSurprise,
                                                double policySurpriseDataWeight, double valueSurpriseDataWeight) {
  const size_t n = targetWeights.size();
  std::vector<double> w(targetWeights.begin(), targetWeights.end());
  if(!(policySurpriseDataWeight > 0 || valueSurpriseDataWeight > 0)) return targetWeights;
  double sumW = 0, sumP = 0, sumV = 0;
  for(size_t i = 0; i < n; i++) {
    if(!(w[i] >= 0.0 && w[i] <= 1.0)) throw std::runtime_error("surprise weighting expects target weights in [0, 1]");
    sumW += w[i]; sumP += policySurprise[i] * w[i]; sumV += valueSurprise[i] * w[i];
  }
  if(sumW < 1) return targetWeights;
  const double avgP = sumP / sumW, avgV = sumV / sumW;
  double vsw = valueSurpriseDataWeight;
  if(avgV < 0.010) vsw *= avgV / 0.010;
  const double threshold = avgP * 1.5;
  std::vector<double> pProp(n), vProp(n);
  double sp = 0, sv = 0;
  for(size_t i = 0; i < n; i++) {
    pProp[i] = w[i] * policySurprise[i] + (1 - w[i]) * std::max(0.0, policySurprise[i] - threshold);
    vProp[i] = w[i] * valueSurprise[i];
  }
  for(size_t i = 0; i < n; i++) { sp += pProp[i]; sv += vProp[i]; }
  sp = std::max(sp, 1e-10); sv = std::max(sv, 1e-10);
  std::vector<float> out(n);
  for(size_t i = 0; i < n; i++)
    out[i] = (float)((1.0 - policySurpriseDataWeight - vsw) * w[i] + policySurpriseDataWeight * pProp[i] * sumW / sp + vsw * vProp[i] * sumW / sv);
  return out;
}

// resolveWeight (play.cpp:2277-2283)
inline float resolveTargetWeight(float weight, RowRand& rand) {
  const double w = std::max((double)weight, 0.0), floored = std::floor(w);
  return (float)(rand.nextBool((double)((float)w - (float)floored)) ? floored + 1 : floored);
}

// FinishedGameData::gameHash: splitmix64 of (seed, slot, game index), as katago_b200/selfplay_cli.py _game_hash
inline void gameHashOf(uint64_t seed, int slot, int index, uint64_t out[2]) {
  auto mix = [](uint64_t z) { z += 0x9E3779B97F4A7C15ULL; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); };
  out[0] = mix(mix(seed) ^ ((uint64_t)slot << 32) ^ (uint64_t)index);
  out[1] = mix(out[0]);
}

// Python's random.Random(seed).random() for a non-negative integer seed (MT19937 seeded by init_by_array over the seed's 32-bit words, 53-bit
// doubles from two outputs): the stream katago_b200/game_recorder.py draws the per-move search limits from, so that both hosts make the same draws.
class PyRandom {
 public:
  explicit PyRandom(uint64_t seed) {
    uint32_t key[2] = {(uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32)};
    const int keyLength = key[1] != 0 ? 2 : 1;
    mt_[0] = 19650218u;
    for(int i = 1; i < N; i++) mt_[i] = 1812433253u * (mt_[i - 1] ^ (mt_[i - 1] >> 30)) + (uint32_t)i;
    int i = 1, j = 0;
    for(int k = N > keyLength ? N : keyLength; k > 0; k--) {
      mt_[i] = (mt_[i] ^ ((mt_[i - 1] ^ (mt_[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
      if(++i >= N) { mt_[0] = mt_[N - 1]; i = 1; }
      if(++j >= keyLength) j = 0;
    }
    for(int k = N - 1; k > 0; k--) {
      mt_[i] = (mt_[i] ^ ((mt_[i - 1] ^ (mt_[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
      if(++i >= N) { mt_[0] = mt_[N - 1]; i = 1; }
    }
    mt_[0] = 0x80000000u;
    idx_ = N;
  }
  double random() { const uint32_t a = next() >> 5, b = next() >> 6; return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0); }
  double expovariate(double lambd) { return -std::log(1.0 - random()) / lambd; }
  // randrange(n): _randbelow_with_getrandbits - k = bit length of n, k-bit draws until one is below n
  uint32_t randrange(uint32_t n) {
    int k = 0;
    for(uint32_t v = n; v; v >>= 1) k++;
    uint32_t r = next() >> (32 - k);
    while(r >= n) r = next() >> (32 - k);
    return r;
  }
  // gauss(mu, sigma): a pair of normal deviates per two uniform draws, the second one kept for the next call
  double gauss(double mu, double sigma) {
    double z;
    if(haveGaussNext_) { z = gaussNext_; haveGaussNext_ = false; }
    else {
      const double x2pi = random() * (2.0 * 3.141592653589793), g2rad = std::sqrt(-2.0 * std::log(1.0 - random()));
      z = std::cos(x2pi) * g2rad;
      gaussNext_ = std::sin(x2pi) * g2rad; haveGaussNext_ = true;
    }
    return mu + z * sigma;
  }
  // choices(population, weights)[0]: bisect_right of random() * total in the running sums, capped at the last index
  size_t choiceIndex(const std::vector<double>& weights) {
    std::vector<double> cum; double acc = 0.0;
    for(size_t i = 0; i < weights.size(); i++) { acc = i == 0 ? weights[0] : acc + weights[i]; cum.push_back(acc); }
    const double x = random() * (cum.back() + 0.0);
    size_t lo = 0, hi = weights.size() - 1;
    while(lo < hi) { const size_t mid = (lo + hi) / 2; if(x < cum[mid]) hi = mid; else lo = mid + 1; }
    return lo;
  }
 private:
  double gaussNext_ = 0.0; bool haveGaussNext_ = false;
  static constexpr int N = 624, M = 397;
  uint32_t next() {
    if(idx_ >= N) {
      for(int k = 0; k < N; k++) {
        const uint32_t y = (mt_[k] & 0x80000000u) | (mt_[(k + 1) % N] & 0x7FFFFFFFu);
        mt_[k] = mt_[(k + M) % N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
      }
      idx_ = 0;
    }
    uint32_t y = mt_[idx_++];
    y ^= y >> 11; y ^= (y << 7) & 0x9D2C5680u; y ^= (y << 15) & 0xEFC60000u; y ^= y >> 18;
    return y;
  }
  uint32_t mt_[N]; int idx_;
};

// getSearchLimitsThisMove (program/play.cpp:1093-1223) without hint moves and asymmetric playouts: the next search's visits, whether its root
// is plain (no noise / temperature: an unrecorded cheap search), the turn's target weight, whether it is a cheap search
struct PlaySettings {
  double cheapSearchProb = 0.0; int cheapSearchVisits = 0; double cheapSearchTargetWeight = 0.0;
  bool reduceVisits = false; double reduceVisitsThreshold = 100.0; int reduceVisitsThresholdLookback = 1; int reducedVisitsMin = 0; double reducedVisitsWeight = 1.0;
  bool active() const { return cheapSearchProb > 0.0 || reduceVisits; }
};
struct SearchLimits { int visits; bool plainRoot; float targetWeight; bool cheap; };
inline SearchLimits searchLimitsThisMove(int maxVisits, const PlaySettings& ps, PyRandom& rand, const std::vector<double>& historicalWinLoss) {
  SearchLimits r{maxVisits, false, 1.0f, false};
  if(ps.cheapSearchProb > 0.0 && rand.random() < ps.cheapSearchProb) {
    if(ps.cheapSearchVisits <= 0 || ps.cheapSearchVisits > maxVisits) throw std::runtime_error("cheapSearchVisits must lie in 1..maxVisits");
    r.cheap = true; r.visits = std::min(r.visits, ps.cheapSearchVisits);
    r.targetWeight = (float)((double)r.targetWeight * (double)(float)ps.cheapSearchTargetWeight);
    if(ps.cheapSearchTargetWeight <= 0.0) r.plainRoot = true;
  }
  else if(ps.reduceVisits) {
    if(ps.reducedVisitsMin <= 0 || ps.reducedVisitsMin > maxVisits) throw std::runtime_error("reducedVisitsMin must lie in 1..maxVisits");
    const size_t look = (size_t)ps.reduceVisitsThresholdLookback, h = historicalWinLoss.size();
    if(h >= look) {
      double lo = 1e20, hi = -1e20;
      for(size_t j = 0; j < look; j++) { lo = std::min(lo, historicalWinLoss[h - 1 - j]); hi = std::max(hi, historicalWinLoss[h - 1 - j]); }
      const double mostExtreme = std::min(std::max(lo, -hi), 1.0), through = mostExtreme - ps.reduceVisitsThreshold;
      if(through > 0) {
        static double (*volatile libmPow)(double, double) = std::pow;       // the same libm call as Python's `** 2`, not folded into a product
        const double prop = libmPow(through / (1.0 - ps.reduceVisitsThreshold), 2.0);
        r.visits = (int)std::floor(r.visits + prop * (ps.reducedVisitsMin - r.visits) + 0.5);
        r.targetWeight = (float)((double)r.targetWeight + prop * ((double)(float)ps.reducedVisitsWeight - (double)r.targetWeight));
        r.visits = std::max(r.visits, ps.reducedVisitsMin);
      }
    }
  }
  r.visits = std::max(2, r.visits);
  return r;
}

class HostRecorder {
 public:
  struct Settings {
    float komi = 7.5f; double drawEquivalentWinsForWhite = 0.5; int koRule = 0; bool multiStoneSuicideLegal = true; int maxVisits = 0;
    double policySurpriseDataWeight = 0.0, valueSurpriseDataWeight = 0.0; bool useSearchValueSurprise = false;
    uint64_t hashSeed = 0; std::string weightRandSeed;      // weightRandSeed empty: fractional weights go to the writer unresolved
    PlaySettings play; uint64_t limitsRandSeed = 0x4C696D69;   // cheap searches / reduced visits, drawn per move
    bool policyInit = false;       // policy-initialised openings (kgb_selfplay_set_policy_init): a slot in its opening is not held; the opening is the game's start history
    bool perGameSetups = false;    // board size, rules and komi are per game (kgb_selfplay_set_game_setup / set_komi): read them from the device
  };
  struct Turn {      // what a finished root search gave
    int nextPlayer, moveNum; std::pair<int, int> move{-1, -1};
    std::vector<uint8_t> packed; std::array<float, NUM_GLOBAL> global;
    std::vector<PolicyTargetMove> policyTarget; int64_t unreducedNumVisits;
    ValueTargets valueTargets; std::vector<QValueTarget> qTargets;
    double surprise, searchEntropy, policyEntropy; std::array<double, 3> nnRawStats, rawNNValues;
    float targetWeight = 1.0f;
  };
  using OnGame = std::function<void(int slot, const FinishedGame&)>;
  // called with the slot when its next game has begun on the device (before any of its turns is recorded): the host's draw for the game after it
  std::function<void(int slot)> onGameStart;
  // Lead targets (play.cpp:2290-2324): after a game, turns drawn with estimateLeadProb get PlayUtils::computeLead of the position before their move
  // as a job on a side loop (integration/b200_komi.h KomiSearcher); the game is handed on when the last answer is back, the slot plays on meanwhile.
  // submitLead(komi, setup, moves, done): `done(lead)` is called once with the job's result.
  using LeadDone = std::function<void(float lead)>;
  std::function<void(float komi, const GameSlots::GameSetup& setup, const std::vector<Move>& moves, LeadDone done)> submitLead;
  double estimateLeadProb = 0.0;
  int gamesWaitingForLead() const { return gamesWaiting_; }
  // Side positions (PlaySettings::sidePositionProb = cfg forkSidePositionProb, play.cpp:1846-1860, 2166-2203): with that probability per turn a forking
  // move is made off the main line (chooseRandomForkingMove: 70 % a temperature-1 policy move, 25 % temperature-2, 5 % any legal move, never the move
  // played) and the position searched on a side loop with the game's own search parameters; its row is written with the game.
  // submitSide(setup, moves, komi, done): `done(position or null)` is called once; sidePositionFrom(loop, slot) reads a searched side-loop slot.
  using SideDone = std::function<void(std::shared_ptr<SidePosition>)>;
  std::function<void(const GameSlots::GameSetup& setup, const std::vector<Move>& moves, float komi, SideDone done)> submitSide;
  double sidePositionProb = 0.0;
  static std::shared_ptr<SidePosition> sidePositionFrom(const GameSlots& loop, int slot) {
    Turn t; GameSlots::RootPosition pos; std::vector<float> policy;
    if(!extractTurn(loop, slot, loop.rootRawPolicyEntropies()[(size_t)slot], t, pos, policy)) return nullptr;
    std::shared_ptr<SidePosition> sp(new SidePosition());
    sp->nextPlayer = t.nextPlayer; sp->packedInput = std::move(t.packed); sp->globalInput = t.global; sp->policyTarget = std::move(t.policyTarget);
    sp->unreducedNumVisits = t.unreducedNumVisits; sp->whiteValueTargets = t.valueTargets; sp->whiteQValueTargets = std::move(t.qTargets);
    sp->policySurprise = t.surprise; sp->policyEntropy = t.policyEntropy; sp->searchEntropy = t.searchEntropy; sp->nnRawStats = t.nnRawStats;
    return sp;
  }
  // the slot's game that has just begun starts from `moves` (already played into the device slot by the caller): a forked game (mode 2)
  void startFrom(int slot, const std::vector<Move>& moves, int mode = 2) {
    InProgress& gm = games_[(size_t)slot];
    gm.presetMoves.clear();
    for(const Move& m : moves) gm.presetMoves.push_back({m.x, m.y});
    gm.mode = mode;
  }

  HostRecorder(GameSlots& slots, const Settings& s, OnGame onGame) : slots_(slots), s_(s), onGame_(std::move(onGame)), games_((size_t)slots.numSlots()) {
    if(!s.weightRandSeed.empty()) weightRand_.reset(new RowRand(s.weightRandSeed));
    const size_t n = (size_t)slots.numSlots();
    curLimits_.assign(n, SearchLimits{s.maxVisits, false, 1.0f, false});
    if(s_.play.active()) {
      // the limits of a root are drawn one move ahead and handed to the device for "the root after this slot's next move" - once for the game
      // going on, once for a new game (kgb_selfplay_set_next_search_limits); the very first roots get theirs here
      limitsRand_.reset(new PyRandom(s.limitsRandSeed));
      nextVisits_.resize(2 * n); nextPlain_.resize(2 * n); pending_.resize(n);
      for(size_t g = 0; g < n; g++) {
        const SearchLimits f = searchLimitsThisMove(s.maxVisits, s_.play, *limitsRand_, {});
        nextVisits_[2 * g] = nextVisits_[2 * g + 1] = f.visits; nextPlain_[2 * g] = nextPlain_[2 * g + 1] = f.plainRoot ? 1 : 0;
        curLimits_[g] = f; pending_[g] = {f, f};
      }
      slots_.setNextSearchLimits(nextVisits_, nextPlain_, true);
    }
    slots_.runWaves(1);            // evaluates every root (its input row stays on the device)
  }
  ~HostRecorder() { for(const std::weak_ptr<Waiting>& weak : waitingGames_) if(std::shared_ptr<Waiting> w = weak.lock()) w->side->waiting.reset(); }      // games still waiting at the end
  int64_t movesRecorded() const { return movesRecorded_; }
  int64_t gamesFinished() const { return gamesFinished_; }

  // `waves` playout waves for every slot; then the slots whose search is finished are read, released, and move in one more wave
  // (the others keep searching in it).  Returns the number of moves recorded.
  int pump(int waves) {
    const int n = slots_.numSlots();
    slots_.runWaves(waves);
    const std::vector<int32_t> visits = slots_.rootVisitsAll();
    const std::vector<int32_t> budgets = s_.play.active() ? slots_.visitBudgets() : std::vector<int32_t>((size_t)n, s_.maxVisits);
    std::vector<uint8_t> held((size_t)n, 0);
    std::vector<int> idx;
    const std::vector<int32_t> openingLeft = s_.policyInit ? slots_.policyInitState() : std::vector<int32_t>((size_t)n, 0);
    for(int g = 0; g < n; g++) if(visits[(size_t)g] >= budgets[(size_t)g] && openingLeft[(size_t)g] <= 0) { held[(size_t)g] = 1; idx.push_back(g); }
    if(idx.empty()) return 0;
    const std::vector<double> rawEntropy = slots_.rootRawPolicyEntropies();
    for(int g : idx) recordRoot(g, rawEntropy[(size_t)g]);
    if(s_.play.active()) slots_.setNextSearchLimits(nextVisits_, nextPlain_);
    slots_.release(held);
    slots_.runWaves(1);
    movesRecorded_ += (int64_t)idx.size();
    for(int g : idx) afterMove(g);
    return (int)idx.size();
  }

 private:
  // a finished game that waits for its lead and side-position jobs; the side positions of a game: searched ones, jobs in flight, the waiting game
  struct SideState;
  // (the two point at each other while the game waits: the side state must outlive its last job for the waiting game to see every side
  //  position, the waiting game must outlive the side state's jobs; the link is cut when the game is handed on, or by the destructor)
  struct Waiting { FinishedGame game; int slot; size_t left; std::shared_ptr<SideState> side; };
  struct SideState { std::vector<std::shared_ptr<SidePosition>> list; int pending = 0; std::shared_ptr<Waiting> waiting; };
  std::vector<std::weak_ptr<Waiting>> waitingGames_;
  void rememberWaiting(const std::shared_ptr<Waiting>& w) {
    if(waitingGames_.size() >= 1024) waitingGames_.erase(std::remove_if(waitingGames_.begin(), waitingGames_.end(), [](const std::weak_ptr<Waiting>& x) { return x.expired(); }), waitingGames_.end());
    waitingGames_.push_back(w);
  }
  void jobBack(std::shared_ptr<Waiting> w) {
    if(--w->left != 0) return;
    gamesWaiting_--;
    w->game.sidePositions = w->side->list;
    w->side->waiting.reset();
    if(onGame_) onGame_(w->slot, w->game);
  }
  struct InProgress {
    std::vector<Turn> turns; std::vector<std::vector<uint8_t>> boards; std::vector<double> winLoss;      // winLoss: historicalMctsWinLossValues
    bool haveSetup = false; GameSlots::GameSetup setup{0, 0, 0, 1};     // this game's own board and rules, read when its first turn is recorded
    std::vector<std::pair<int, int>> startMoves;                         // moves before the first recorded turn (fork prefix + policy-initialised opening): startHist
    std::vector<float> lastPolicy;                                       // the root policy of the turn being played (for the side position's forking move)
    std::shared_ptr<SideState> side{new SideState()};                    // side positions of this game
    std::vector<std::pair<int, int>> presetMoves; int mode = 0;         // moves the host played into the slot before the game's first search (a forked game); FinishedGameData::mode
  };
  // a board or area of the evaluator's frame cut down to the game's own board (its top-left corner)
  std::vector<uint8_t> crop(const std::vector<uint8_t>& frame, int bx, int by) const {
    const int X = slots_.xLen();
    std::vector<uint8_t> out((size_t)bx * by);
    for(int y = 0; y < by; y++) for(int x = 0; x < bx; x++) out[(size_t)y * bx + x] = frame[(size_t)y * X + x];
    return out;
  }

  // What extractSearchTargetsThisTurn / the side-position block of Play::runGame (play.cpp:931-948, 2178-2203) read of a finished search, from slot g
  // of loop `sp` held at its budget.  Returns false when the kept input row belongs to another position.
  static bool extractTurn(const GameSlots& sp, int g, double rawPolicyEntropy, Turn& t, GameSlots::RootPosition& pos, std::vector<float>& policy) {
    const int X = sp.xLen(), Y = sp.yLen(), A = X * Y;
    pos = sp.rootPosition(g);
    std::vector<float> spatial, global;
    sp.rootInputRow(g, spatial, global);
    policy = sp.rootPolicy(g);
    std::vector<double> childMoments; double rootMoments[5], nn[5];
    sp.rootValueStatsByPos(g, childMoments, rootMoments);
    const std::vector<double> psv = sp.playSelectionValuesByPos(g);
    std::vector<int32_t> nodeVisits;
    sp.rootExtraByPos(g, nodeVisits, nn);
    t.nextPlayer = pos.blackToMove ? P_BLACK : P_WHITE;
    t.moveNum = pos.moveNumber;
    // the kept row must be this root's: its own / opponent stone planes are the root position
    for(int p = 0; p < A; p++)
      if((spatial[(size_t)p * NUM_BIN + 1] != 0) != (pos.colors[(size_t)p] == t.nextPlayer) || (spatial[(size_t)p * NUM_BIN + 2] != 0) != (pos.colors[(size_t)p] == 3 - t.nextPlayer))
        return false;
    const int packedLen = (A + 7) / 8;
    t.packed.assign((size_t)NUM_BIN * packedLen, 0);          // packBits: 8 points per byte, first point in the high bit
    for(int p = 0; p < A; p++)
      for(int c = 0; c < NUM_BIN; c++)
        if(spatial[(size_t)p * NUM_BIN + c] != 0) t.packed[(size_t)c * packedLen + p / 8] |= (uint8_t)(0x80 >> (p % 8));
    for(int i = 0; i < NUM_GLOBAL; i++) t.global[(size_t)i] = global[(size_t)i];
    t.policyTarget = policyTargetMoves(psv, X);
    t.unreducedNumVisits = pos.rootVisits;
    const Reported rv = reportedSearchValues(rootMoments);
    t.valueTargets.win = (float)rv.win; t.valueTargets.loss = (float)rv.loss; t.valueTargets.noResult = (float)rv.noResult; t.valueTargets.score = (float)rv.score;
    for(size_t p = 0; p < nodeVisits.size(); p++) {             // extractQValueTargets (play.cpp:859-888)
      if(nodeVisits[p] <= 0) continue;
      const Reported c = reportedSearchValues(&childMoments[p * 5]);
      const bool pass = p == nodeVisits.size() - 1;
      t.qTargets.push_back(QValueTarget{pass ? -1 : (int)(p % X), pass ? -1 : (int)(p / X), (float)c.winLoss, (float)c.score, nodeVisits[p]});
    }
    policySurpriseAndEntropy(psv, policy, t.surprise, t.searchEntropy, t.policyEntropy);
    t.nnRawStats = {nn[0], nn[2], rawPolicyEntropy};
    const Reported rn = reportedSearchValues(nn);
    t.rawNNValues = {rn.win, rn.loss, rn.noResult};
    return true;
  }

  void recordRoot(int g, double rawPolicyEntropy) {
    const int X = slots_.xLen(), Y = slots_.yLen();
    Turn t; GameSlots::RootPosition pos; std::vector<float> policy;
    if(!extractTurn(slots_, g, rawPolicyEntropy, t, pos, policy))
      throw std::runtime_error("HostRecorder: slot " + std::to_string(g) + ", move " + std::to_string(pos.moveNumber) + ": the kept input row belongs to another position");
    t.targetWeight = curLimits_[(size_t)g].targetWeight;
    InProgress& gm = games_[(size_t)g];
    gm.lastPolicy = policy;
    if(!gm.haveSetup) {
      gm.setup = GameSlots::GameSetup{X, Y, s_.koRule, s_.multiStoneSuicideLegal ? 1 : 0};
      if(s_.perGameSetups) { std::vector<GameSlots::GameSetup> cur; slots_.gameSetups(&cur, nullptr); gm.setup = cur[(size_t)g]; }
      gm.haveSetup = true;
      gm.startMoves = gm.presetMoves;
      if(s_.policyInit) {
        std::vector<std::vector<Move>> openings;
        slots_.policyInitState(&openings);
        for(const Move& m : openings[(size_t)g]) gm.startMoves.push_back({m.x, m.y});
      }
      if(s_.policyInit || !gm.presetMoves.empty()) {
        if((int)gm.startMoves.size() != pos.moveNumber)
          throw std::runtime_error("HostRecorder: slot " + std::to_string(g) + ": " + std::to_string(pos.moveNumber) + " moves played before the first searched move, " +
                                   std::to_string(gm.startMoves.size()) + " opening moves kept");
      }
    }
    gm.boards.push_back(crop(pos.colors, gm.setup.x, gm.setup.y));
    gm.winLoss.push_back((double)t.valueTargets.win - (double)t.valueTargets.loss);
    if(s_.play.active()) {         // limits of the search that follows this slot's move: the game goes on / a new game starts
      const SearchLimits cont = searchLimitsThisMove(s_.maxVisits, s_.play, *limitsRand_, gm.winLoss);
      const SearchLimits fresh = searchLimitsThisMove(s_.maxVisits, s_.play, *limitsRand_, {});
      pending_[(size_t)g] = {cont, fresh};
      nextVisits_[2 * (size_t)g] = cont.visits; nextVisits_[2 * (size_t)g + 1] = fresh.visits;
      nextPlain_[2 * (size_t)g] = cont.plainRoot ? 1 : 0; nextPlain_[2 * (size_t)g + 1] = fresh.plainRoot ? 1 : 0;
    }
    gm.turns.push_back(std::move(t));
  }

  void afterMove(int g) {
    const GameSlots::LastMove last = slots_.lastMove(g);
    games_[(size_t)g].turns.back().move = {last.move.x, last.move.y};
    if(submitSide && !last.gameOver && sideRand_.random() < sidePositionProb) submitSidePosition(g, last);
    if(s_.play.active()) curLimits_[(size_t)g] = last.gameOver ? pending_[(size_t)g].second : pending_[(size_t)g].first;
    if(last.gameOver) finishGame(g, last);
  }

  // PlayUtils::chooseRandomForkingMove (play.cpp:796-808): a move position or -1.  policy by move position, -1 = illegal (here the root policy as
  // searched - the reference reads the un-noised one)
  int chooseRandomForkingMove(const std::vector<float>& policy, int banPos) {
    const double r = sideRand_.random();
    std::vector<int> legal;
    for(size_t i = 0; i < policy.size(); i++) if(policy[i] >= 0 && (int)i != banPos) legal.push_back((int)i);
    if(r >= 0.95) return legal.empty() ? -1 : legal[sideRand_.randrange((uint32_t)legal.size())];
    const double t = r < 0.70 ? 1.0 : 2.0;
    std::vector<int> cand; std::vector<double> w;
    static double (*volatile libmPow)(double, double) = std::pow;       // the same libm call as Python's `**`
    for(int i : legal) if(policy[(size_t)i] > 0) { cand.push_back(i); w.push_back(libmPow((double)policy[(size_t)i], 1.0 / t)); }
    if(cand.empty()) return -1;
    return cand[sideRand_.choiceIndex(w)];
  }
  void submitSidePosition(int g, const GameSlots::LastMove& last) {
    InProgress& gm = games_[(size_t)g];
    const int X = slots_.xLen(), A = X * slots_.yLen();
    const int pos = chooseRandomForkingMove(gm.lastPolicy, last.move.isPass() ? A : last.move.y * X + last.move.x);
    if(pos < 0 || !gm.haveSetup) return;
    std::vector<Move> moves;
    for(const auto& m : gm.startMoves) { Move mv; mv.x = m.first; mv.y = m.second; moves.push_back(mv); }
    for(size_t i = 0; i + 1 < gm.turns.size(); i++) { Move mv; mv.x = gm.turns[i].move.first; mv.y = gm.turns[i].move.second; moves.push_back(mv); }
    Move fork; if(pos != A) { fork.x = pos % X; fork.y = pos / X; }
    moves.push_back(fork);
    float komi = s_.komi;
    if(s_.perGameSetups) { std::vector<float> cur; slots_.komis(&cur, nullptr); komi = cur[(size_t)g]; }
    const int turnIdx = gm.turns.back().moveNum + 1;
    std::shared_ptr<SideState> side = gm.side;
    side->pending++;
    submitSide(gm.setup, moves, komi, [this, side, turnIdx](std::shared_ptr<SidePosition> sp) {
      if(sp) { sp->turnIdx = turnIdx; side->list.push_back(sp); }      // (null: the forking move ended the game, or the row was not the root's)
      side->pending--;
      if(side->waiting) jobBack(side->waiting);
    });
  }

  void finishGame(int g, const GameSlots::LastMove& last) {
    InProgress gm = std::move(games_[(size_t)g]);
    games_[(size_t)g] = InProgress();
    const int X = gm.setup.x, Y = gm.setup.y;            // this game's own board; rows are written inside the evaluator's frame
    FinishedGame d;
    d.xSize = X; d.ySize = Y; d.komi = s_.komi;
    if(s_.perGameSetups) { std::vector<float> last_; slots_.komis(nullptr, &last_); d.komi = last_[(size_t)g]; }      // the slot's last finished game's
    d.drawEquivalentWinsForWhite = s_.drawEquivalentWinsForWhite;
    gameHashOf(s_.hashSeed, g, last.gameIndex, d.gameHash);
    d.endFinished = !last.hitMoveLimit; d.hitTurnLimit = last.hitMoveLimit; d.endNoResult = last.noResult;
    static const char* KO[] = {"SIMPLE", "POSITIONAL", "SITUATIONAL", "SPIGHT"};
    d.koRule = KO[gm.setup.koRule & 3]; d.multiStoneSuicideLegal = gm.setup.multiStoneSuicideLegal != 0;
    d.startMoves = gm.startMoves; d.startHistMoves = (int)gm.startMoves.size(); d.mode = gm.mode;
    d.boardsByTurn = std::move(gm.boards);
    d.boardsByTurn.push_back(crop(last.finalColors, X, Y));
    std::vector<std::array<double, 3>> rawNN;
    for(Turn& t : gm.turns) {
      d.moves.push_back(t.move); d.nextPlayerByTurn.push_back(t.nextPlayer);
      d.packedInputByTurn.push_back(std::move(t.packed)); d.globalInputByTurn.push_back(t.global);
      d.targetWeightByTurn.push_back(t.targetWeight);
      d.policyTargetsByTurn.push_back(std::move(t.policyTarget)); d.unreducedNumVisitsByTurn.push_back(t.unreducedNumVisits);
      d.policySurpriseByTurn.push_back(t.surprise); d.policyEntropyByTurn.push_back(t.policyEntropy); d.searchEntropyByTurn.push_back(t.searchEntropy);
      d.whiteValueTargetsByTurn.push_back(t.valueTargets); d.whiteQValueTargetsByTurn.push_back(std::move(t.qTargets));
      d.nnRawStatsByTurn.push_back(t.nnRawStats); rawNN.push_back(t.rawNNValues);
    }
    std::vector<uint8_t> area((size_t)X * Y, 0);           // a game without a result: nobody owns anything (play.cpp:1977-1988)
    if(d.endNoResult) d.whiteValueTargetsByTurn.push_back(finalValueTargets(0, 0.0f, s_.drawEquivalentWinsForWhite, d.komi, true));
    else {
      // area scoring without tax: ownership = full area = calculateArea with every flag on (boardhistory.cpp:591-610)
      area = crop(last.finalArea, X, Y);
      const float score = last.finalWhiteMinusBlackScore;
      d.winner = score > 0 ? P_WHITE : score < 0 ? P_BLACK : 0;
      d.finalWhiteMinusBlackScore = score;
      d.whiteValueTargetsByTurn.push_back(finalValueTargets(d.winner, score, s_.drawEquivalentWinsForWhite, d.komi, false));
    }
    d.finalFullArea = area; d.finalOwnership = area;
    if(s_.policySurpriseDataWeight > 0 || s_.valueSurpriseDataWeight > 0) {        // play.cpp:2034-2163
      const std::vector<double> valueSurprise = computeValueSurpriseByTurn(d.whiteValueTargetsByTurn, rawNN, X * Y, s_.useSearchValueSurprise);
      d.targetWeightByTurn = surpriseTargetWeights(d.targetWeightByTurn, d.policySurpriseByTurn, valueSurprise, s_.policySurpriseDataWeight, s_.valueSurpriseDataWeight);
      d.targetWeightByTurnUnrounded = d.targetWeightByTurn;
    }
    if(weightRand_) {                                                               // play.cpp:2274-2289
      if(d.targetWeightByTurnUnrounded.empty()) d.targetWeightByTurnUnrounded = d.targetWeightByTurn;
      for(float& w : d.targetWeightByTurn) w = resolveTargetWeight(w, *weightRand_);
    }
    d.finalWhiteScoring.resize(area.size());                // NNInputs::fillScoring without group tax: white area +1, black area -1
    for(size_t i = 0; i < area.size(); i++) d.finalWhiteScoring[i] = area[i] == P_WHITE ? 1.0f : area[i] == P_BLACK ? -1.0f : 0.0f;
    gamesFinished_++;
    if(onGameStart) onGameStart(g);
    std::shared_ptr<SideState> side = gm.side;
    d.sidePositions = side->list;
    if(submitLead && estimateLeadProb > 0 && !d.endNoResult) {
      std::vector<size_t> turns;
      for(size_t t = 0; t < d.targetWeightByTurn.size(); t++)         // (the draw is made only for turns that qualify)
        if(d.targetWeightByTurn[t] > 0 && (double)d.whiteValueTargetsByTurn[t].noResult < 0.3 && leadRand_.random() < estimateLeadProb) turns.push_back(t);
      if(!turns.empty()) {
        std::shared_ptr<Waiting> w(new Waiting{std::move(d), g, turns.size() + (size_t)side->pending, side});
        side->waiting = w;
        rememberWaiting(w);
        gamesWaiting_++;
        for(size_t t : turns) {
          std::vector<Move> moves;
          for(const auto& m : w->game.startMoves) { Move mv; mv.x = m.first; mv.y = m.second; moves.push_back(mv); }
          for(size_t i = 0; i < t; i++) { Move mv; mv.x = w->game.moves[i].first; mv.y = w->game.moves[i].second; moves.push_back(mv); }
          submitLead(w->game.komi, gm.setup, moves, [this, w, t](float lead) {
            w->game.whiteValueTargetsByTurn[t].hasLead = true; w->game.whiteValueTargetsByTurn[t].lead = lead;      // ValueTargets::hasLead, lead
            jobBack(w);
          });
        }
        return;
      }
    }
    if(side->pending > 0) {                  // side positions of this game are still being searched
      side->waiting.reset(new Waiting{std::move(d), g, (size_t)side->pending, side});
      rememberWaiting(side->waiting);
      gamesWaiting_++;
      return;
    }
    if(onGame_) onGame_(g, d);
  }

  GameSlots& slots_; Settings s_; OnGame onGame_;
  std::vector<InProgress> games_;
  std::unique_ptr<RowRand> weightRand_;
  std::unique_ptr<PyRandom> limitsRand_;
  PyRandom leadRand_{0x4C656164}, sideRand_{0x53696465}; int gamesWaiting_ = 0;
  std::vector<SearchLimits> curLimits_; std::vector<std::pair<SearchLimits, SearchLimits>> pending_;
  std::vector<int32_t> nextVisits_; std::vector<uint8_t> nextPlain_;
  int64_t movesRecorded_ = 0, gamesFinished_ = 0;
};

}  // namespace b200
