// tests/mock/kgb200_mock.cpp - TEST INFRASTRUCTURE: a CPU stand-in for the part of libkgb200's C ABI that the recording path uses
// (include/kgb200.h: evaluator handle life cycle + kgb_selfplay_* in hold mode), so that integration/b200record.h and the reference's
// own TrainingDataWriter can be exercised without a GPU (tests/test_game_recorder.py::test_cpp_recorder_against_python_recorder_on_scripted_slots).
//
// Each slot plays uniformly random legal moves on the reference's own Board / BoardHistory (this file is compiled against the
// reference like oracle/ref_record_driver.cpp) and reports made-up, reproducible search statistics.  Everything it serves is also
// appended to the JSON-lines file named by KGB_MOCK_LOG, from which the Python side replays exactly the same slots.
#include "include/kgb200.h"

#include "game/board.h"
#include "game/boardhistory.h"
#include "neuralnet/nninputs.h"
#include "core/rand.h"

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

namespace {
struct Lcg {
  uint64_t s;
  explicit Lcg(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ULL + 777) {}
  uint32_t next() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(s >> 33); }
  double unit() { return (next() & 0xFFFFFF) / 16777216.0; }
};

struct Slot {
  Board board; BoardHistory hist; Player pla = P_BLACK;
  int moveNum = 0, gameIndex = 0;
  int32_t setup[4] = {0, 0, 0, 1}; float komi = 7.5f;      // this game's board X, Y, ko rule, multi-stone suicide; komi
  int openingLeft = 0;              // KGB_MOCK_UNEVEN: opening moves still to be played, one per wave; the slot has no root meanwhile
  int wavesLeft = 0;                // KGB_MOCK_UNEVEN: waves the search of the current root still needs before its visits reach the budget
  int initPending = 0; std::vector<int16_t> initMoves;      // policy-initialised opening: moves to play when the game starts / the ones played (frame positions)
  bool held = true;                 // searches finish instantly in the mock
  // what the "search" of the current root found
  std::vector<int32_t> edgeVisits, nodeVisits; std::vector<float> policy; std::vector<double> childStats, psv;
  double rootStats[5], rootNN[5];
  std::vector<float> rowSpatial, rowGlobal;
  Loc nextMove = Board::PASS_LOC;
  // last move
  int32_t last[4] = {0, 0, 0, 0}; float lastScore = 0; std::vector<uint8_t> finalColors, finalArea;
};
}  // namespace

struct kgb_model { std::string path; };
struct kgb_context { int x, y; };
struct kgb_handle { int x, y; bool otherArch; };
struct kgb_selfplay {
  kgb_selfplay_config cfg; int X, Y; Rules rules; Lcg rng{1}, waveRng{12345}; std::vector<Slot> slots; std::vector<uint8_t> released; std::ofstream log;
  // per-root search limits (kgb_selfplay_set_next_search_limits): the current roots' and, per slot, those of the root after its next move [goes on, new game]
  std::vector<int32_t> budget, nextBudget; std::vector<uint8_t> plain, nextPlain;
  // per-game board, rules and komi (kgb_selfplay_set_game_setup / set_komi): of each slot's next game and of its last finished one
  std::vector<int32_t> nextSetup, lastSetup; std::vector<float> nextKomi, lastKomi;
  std::vector<int32_t> nextInit;    // opening length of each slot's next game (kgb_selfplay_set_policy_init)
  bool started = false;             // the first games begin with the first run / read, after the host has handed over their setups
};

static std::string g_err;

static int unevenWaves() { static const int k = getenv("KGB_MOCK_UNEVEN") ? atoi(getenv("KGB_MOCK_UNEVEN")) : 0; return k; }

static void searchRoot(kgb_selfplay* sp, int g) {
  Slot& s = sp->slots[g];
  if(s.openingLeft > 0) return;       // still in its opening: the root comes when the last opening move has been played
  const int X = sp->X, Y = sp->Y, P = X * Y + 1;          // the evaluator's frame; the game's board is its top-left corner
  const int BX = s.setup[0], BY = s.setup[1];
  Lcg& r = sp->rng;
  s.edgeVisits.assign(P, 0); s.nodeVisits.assign(P, 0); s.policy.assign(P, -1.0f); s.childStats.assign((size_t)P * 5, 0.0); s.psv.assign(P, -1.0);
  std::vector<int> legal;
  for(int y = 0; y < BY; y++) for(int x = 0; x < BX; x++) if(s.hist.isLegal(s.board, Location::getLoc(x, y, BX), s.pla)) legal.push_back(y * X + x);
  legal.push_back(P - 1);
  double tot = 0;
  for(int pos : legal) { s.policy[pos] = (float)(0.05 + r.unit()); tot += s.policy[pos]; }
  for(int pos : legal) s.policy[pos] = (float)(s.policy[pos] / tot);
  // the move the slot will play: uniform over the legal board points, a pass once in 25 moves, answered by a second pass one time in three
  const bool lastWasPass = s.hist.moveHistory.size() > 0 && s.hist.moveHistory.back().loc == Board::PASS_LOC;
  int chosen = (legal.size() == 1 || r.next() % 25 == 0 || (lastWasPass && r.next() % 3 == 0)) ? P - 1 : legal[r.next() % (legal.size() - 1)];
  for(int pos : legal) {
    if(pos != chosen && r.next() % 3 != 0) continue;
    s.edgeVisits[pos] = 1 + (int)(r.next() % 40); s.nodeVisits[pos] = s.edgeVisits[pos] + (int)(r.next() % 3);
    s.psv[pos] = (double)s.edgeVisits[pos] * (r.next() % 7 == 0 ? 0.0 : 1.0) + (pos == chosen ? 0.5 : 0.0);
    s.childStats[(size_t)pos * 5 + 0] = r.unit() * 2 - 1; s.childStats[(size_t)pos * 5 + 1] = r.unit() * 0.05;
    s.childStats[(size_t)pos * 5 + 2] = (r.unit() - 0.5) * 40; s.childStats[(size_t)pos * 5 + 3] = 500 * r.unit(); s.childStats[(size_t)pos * 5 + 4] = (r.unit() - 0.5) * 30;
  }
  s.nextMove = chosen == P - 1 ? Board::PASS_LOC : Location::getLoc(chosen % X, chosen / X, BX);
  s.rootStats[0] = r.unit() * 2 - 1; s.rootStats[1] = r.unit() * 0.04; s.rootStats[2] = (r.unit() - 0.5) * 30; s.rootStats[3] = 400 * r.unit(); s.rootStats[4] = (r.unit() - 0.5) * 20;
  s.rootNN[0] = r.unit() * 2 - 1; s.rootNN[1] = 0.0; s.rootNN[2] = (r.unit() - 0.5) * 30; s.rootNN[3] = 300 * r.unit(); s.rootNN[4] = 0.0;
  // the root's input row: the reference's own fillRowV7 (NHWC) - what the device loop's featurizer is pinned to
  MiscNNInputParams ip; ip.drawEquivalentWinsForWhite = sp->cfg.draw_equivalent_wins_for_white;
  s.rowSpatial.assign((size_t)X * Y * 22, 0.0f); s.rowGlobal.assign(19, 0.0f);
  NNInputs::fillRowV7(s.board, s.hist, s.pla, ip, X, Y, true, s.rowSpatial.data(), s.rowGlobal.data());
  s.held = true;
  // KGB_MOCK_UNEVEN = K: searches take 1..K waves (else they finish at once), so that hosts see slots finish at different times like on the device
  const int uneven = unevenWaves();
  s.wavesLeft = uneven > 0 ? 1 + (int)(sp->waveRng.next() % (uint32_t)uneven) : 0;
  // log
  std::ofstream& o = sp->log;
  auto arr = [&](const char* name, auto& v, bool last = false) {
    o << "\"" << name << "\":[";
    for(size_t i = 0; i < v.size(); i++) {
      char b[64]; snprintf(b, sizeof b, "%.17g", (double)v[i]);
      o << (i ? "," : "") << (strcmp(b, "-0") == 0 ? "-0.0" : b);       // (JSON readers take "-0" for the integer 0 and drop the sign)
    }
    o << "]" << (last ? "" : ",");
  };
  std::vector<int> colors((size_t)X * Y, 0);
  for(int y = 0; y < BY; y++) for(int x = 0; x < BX; x++) colors[(size_t)y * X + x] = (int)s.board.colors[Location::getLoc(x, y, BX)];
  std::vector<double> rs(s.rootStats, s.rootStats + 5), rn(s.rootNN, s.rootNN + 5);
  o << "{\"ev\":\"root\",\"slot\":" << g << ",\"waves_needed\":" << s.wavesLeft << ",\"move_num\":" << s.moveNum << ",\"black_to_move\":" << (s.pla == P_BLACK ? 1 : 0) << ",";
  arr("colors", colors); arr("edge_visits", s.edgeVisits); arr("node_visits", s.nodeVisits); arr("policy", s.policy); arr("child_stats", s.childStats);
  arr("psv", s.psv); arr("root_stats", rs); arr("root_nn", rn); arr("init_moves", s.initMoves); arr("row_spatial", s.rowSpatial); arr("row_global", s.rowGlobal, true);
  o << "}\n";
}

// the opening the device would draw from the policy: here uniformly random legal board moves, never held for recording
static void playOpeningMoves(kgb_selfplay* sp, int g, int count) {
  Slot& s = sp->slots[g];
  for(int i = 0; i < count; i++) {
    std::vector<Loc> legal;
    for(int y = 0; y < s.setup[1]; y++) for(int x = 0; x < s.setup[0]; x++) { Loc l = Location::getLoc(x, y, s.setup[0]); if(s.hist.isLegal(s.board, l, s.pla)) legal.push_back(l); }
    if(legal.empty()) break;
    const Loc l = legal[sp->rng.next() % legal.size()];
    s.hist.makeBoardMoveAssumeLegal(s.board, l, s.pla, NULL);
    s.pla = getOpp(s.pla);
    s.initMoves.push_back((int16_t)(Location::getY(l, s.setup[0]) * sp->X + Location::getX(l, s.setup[0])));
    s.moveNum++;
  }
}

static void startGame(kgb_selfplay* sp, int g) {
  Slot& s = sp->slots[g];
  Rules rules = sp->rules;
  rules.koRule = s.setup[2] == 1 ? Rules::KO_POSITIONAL : s.setup[2] == 2 ? Rules::KO_SITUATIONAL : s.setup[2] == 3 ? Rules::KO_SPIGHT : Rules::KO_SIMPLE;
  rules.multiStoneSuicideLegal = s.setup[3] != 0; rules.komi = s.komi;
  s.board = Board(s.setup[0], s.setup[1]); s.pla = P_BLACK; s.hist = BoardHistory(s.board, s.pla, rules, 0, false); s.moveNum = 0;
  // the opening: played at once - or, with KGB_MOCK_UNEVEN, one move per wave like the device (the slot has no root meanwhile)
  s.initMoves.clear();
  if(unevenWaves() > 0) s.openingLeft = s.initPending;
  else playOpeningMoves(sp, g, s.initPending);
  s.initPending = 0;
}

static void advance(kgb_selfplay* sp, int g, bool searched = true, Loc given = Board::NULL_LOC) {
  Slot& s = sp->slots[g];
  const int X = sp->X, Y = sp->Y, BX = s.setup[0], BY = s.setup[1];
  const Loc loc = searched ? s.nextMove : given;       // (searched = false: a move the host plays into the slot, kgb_selfplay_play_moves_game)
  s.hist.makeBoardMoveAssumeLegal(s.board, loc, s.pla, NULL);
  s.pla = getOpp(s.pla);
  const int maxMoves = sp->cfg.max_moves > 0 ? sp->cfg.max_moves : 2 * BX * BY;
  const bool finished = s.hist.isGameFinished, over = finished || s.moveNum + 1 >= maxMoves;
  s.last[0] = loc == Board::PASS_LOC ? X * Y : Location::getY(loc, BX) * X + Location::getX(loc, BX);
  s.last[1] = over ? (1 | ((finished && s.hist.isNoResult) ? 2 : 0) | (finished ? 0 : 4)) : 0;
  s.last[2] = s.moveNum; s.last[3] = s.gameIndex;
  if(over) {
    Color area[Board::MAX_ARR_SIZE];
    BoardHistory h2 = s.hist;
    h2.endAndScoreGameNow(s.board, area);
    s.lastScore = h2.finalWhiteMinusBlackScore;
    s.finalColors.assign((size_t)X * Y, 0); s.finalArea.assign((size_t)X * Y, 0);
    for(int y = 0; y < BY; y++) for(int x = 0; x < BX; x++) { Loc l = Location::getLoc(x, y, BX); s.finalColors[(size_t)y * X + x] = s.board.colors[l]; s.finalArea[(size_t)y * X + x] = area[l]; }
  }
  std::ofstream& o = sp->log;
  if(searched) {
    o << "{\"ev\":\"move\",\"slot\":" << g << ",\"pos\":" << s.last[0] << ",\"flags\":" << s.last[1] << ",\"move_num\":" << s.last[2] << ",\"game_index\":" << s.last[3]
      << ",\"score\":" << s.lastScore << ",\"final_colors\":[";
    for(size_t i = 0; over && i < s.finalColors.size(); i++) o << (i ? "," : "") << (int)s.finalColors[i];
    o << "],\"final_area\":[";
    for(size_t i = 0; over && i < s.finalArea.size(); i++) o << (i ? "," : "") << (int)s.finalArea[i];
    o << "]}\n";
  }
  if(over) {                        // the slot's next game takes the setup and komi handed over for it
    for(int k = 0; k < 4; k++) { sp->lastSetup[4 * (size_t)g + k] = s.setup[k]; s.setup[k] = sp->nextSetup[4 * (size_t)g + k]; }
    sp->lastKomi[g] = s.komi; s.komi = sp->nextKomi[g];
    s.initPending = sp->nextInit[g];
    s.gameIndex++; startGame(sp, g);
  }
  else s.moveNum++;
  if(searched) searchRoot(sp, g);
}

static void ensureStarted(kgb_selfplay* sp) {
  if(sp->started) return;
  sp->started = true;
  for(size_t g = 0; g < sp->slots.size(); g++) { startGame(sp, (int)g); searchRoot(sp, (int)g); }
}

#define GUARD(body) try { body; return 0; } catch(const std::exception& e) { g_err = e.what(); return 1; }

extern "C" {
int kgb_global_init(void) { static bool done = false; if(!done) { Board::initHash(); done = true; } return 0; }     // (returns at once when the caller has done it)
int kgb_global_cleanup(void) { return 0; }
const char* kgb_last_error(void) { return g_err.c_str(); }
int kgb_model_load_file(const char* path, const char*, kgb_model** out) { *out = new kgb_model{path}; return 0; }
void kgb_model_free(kgb_model* m) { delete m; }
int kgb_model_get_info(const kgb_model*, kgb_model_info* out) { memset(out, 0, sizeof(*out)); strcpy(out->name, "mocknet"); return 0; }
// the reference's own Rand: the stream the library's host generator reproduces (tests/test_abi_and_loader.py)
int kgb_rand_uint32_stream(const char* seed, int n, uint32_t* out) { Rand r(seed); for(int i = 0; i < n; i++) out[i] = r.nextUInt(); return 0; }
int kgb_context_create(const int*, int, int x, int y, int, const kgb_model*, kgb_context** out) { *out = new kgb_context{x, y}; return 0; }
void kgb_context_free(kgb_context* c) { delete c; }
int kgb_handle_create(kgb_context* c, const kgb_model* m, int, int, int, int, kgb_handle** out) { *out = new kgb_handle{c->x, c->y, m->path.find("otherarch") != std::string::npos}; return 0; }
void kgb_handle_free(kgb_handle* h) { delete h; }
int kgb_handle_sync(kgb_handle*) { return 0; }
// weight hot-swap: nothing to swap in the mock; a net whose path contains "otherarch" is of another architecture than a handle built for one without
int kgb_handle_stage_weights(kgb_handle* h, const kgb_model* m) {
  if((m->path.find("otherarch") != std::string::npos) != h->otherArch) { g_err = "kgb_handle_stage_weights: the model is not of the architecture this handle was built for"; return 1; }
  return 0;
}
int kgb_handle_commit_weights(kgb_handle*) { return 0; }
int kgb_selfplay_clear_nn_cache(kgb_selfplay*) { return 0; }
int kgb_handle_wait_staged(kgb_handle*) { return 0; }
// the weight broadcast between ranks: no device memory here, the calls only have to line up
int kgb_nccl_unique_id(void* id) { for(int i = 0; i < 128; i++) ((unsigned char*)id)[i] = (unsigned char)(i * 7 + 1); return 0; }
int kgb_handle_comm_init(kgb_handle*, const void* id, int rank, int numRanks) {
  for(int i = 0; i < 128; i++) if(((const unsigned char*)id)[i] != (unsigned char)(i * 7 + 1)) { g_err = "mock: not the id rank 0 made"; return 1; }
  if(rank < 0 || rank >= numRanks) { g_err = "mock: bad rank"; return 1; }
  return 0;
}
int kgb_handle_broadcast_staged_weights(kgb_handle*, int, float* ms) { if(ms) *ms = 0.0f; return 0; }

int kgb_selfplay_create(kgb_handle* h, const kgb_selfplay_config* c, kgb_selfplay** out) {
  GUARD({
    const char* base = getenv("KGB_MOCK_LOG");
    if(!base) throw std::runtime_error("KGB_MOCK_LOG is not set");
    static int instances = 0;                         // a second / third loop of the process (side loops) logs to <path>.2 / <path>.3
    instances++;
    const std::string path = instances == 1 ? std::string(base) : std::string(base) + "." + std::to_string(instances);
    kgb_selfplay* sp = new kgb_selfplay();
    sp->cfg = *c; sp->X = h->x; sp->Y = h->y; sp->rng = Lcg(c->seed);
    sp->rules.koRule = c->ko_rule == 1 ? Rules::KO_POSITIONAL : c->ko_rule == 2 ? Rules::KO_SITUATIONAL : c->ko_rule == 3 ? Rules::KO_SPIGHT : Rules::KO_SIMPLE;
    sp->rules.scoringRule = Rules::SCORING_AREA; sp->rules.taxRule = Rules::TAX_NONE; sp->rules.multiStoneSuicideLegal = c->multi_stone_suicide_legal != 0;
    sp->rules.hasButton = false; sp->rules.whiteHandicapBonusRule = Rules::WHB_ZERO; sp->rules.friendlyPassOk = false; sp->rules.komi = c->komi;
    sp->log.open(path);
    sp->slots.resize(c->num_games); sp->released.assign(c->num_games, 0);
    sp->budget.assign(c->num_games, c->max_visits); sp->nextBudget.assign(2 * (size_t)c->num_games, c->max_visits);
    sp->plain.assign(c->num_games, 0); sp->nextPlain.assign(2 * (size_t)c->num_games, 0);
    for(int g = 0; g < c->num_games; g++) {
      Slot& s = sp->slots[g];
      s.setup[0] = sp->X; s.setup[1] = sp->Y; s.setup[2] = c->ko_rule; s.setup[3] = c->multi_stone_suicide_legal != 0; s.komi = c->komi;
      for(int k = 0; k < 4; k++) { sp->nextSetup.push_back(s.setup[k]); sp->lastSetup.push_back(s.setup[k]); }
      sp->nextKomi.push_back(c->komi); sp->lastKomi.push_back(c->komi); sp->nextInit.push_back(0);
    }
    *out = sp;
  })
}
void kgb_selfplay_free(kgb_selfplay* sp) { delete sp; }
int kgb_selfplay_run(kgb_selfplay* sp, int waves) {
  GUARD({
    ensureStarted(sp);
    // KGB_MOCK_NEW_MODEL = "<runs>:<path>": a new net appears in the models directory while the host is running (the trainer's export)
    if(const char* nm = getenv("KGB_MOCK_NEW_MODEL")) {
      static int runs = 0;
      const std::string spec = nm; const size_t colon = spec.find(':');
      if(++runs == atoi(spec.substr(0, colon).c_str())) { std::ofstream f(spec.substr(colon + 1)); f << "unused"; }
    }
    for(size_t g = 0; g < sp->slots.size(); g++) if(!sp->released[g]) {
      Slot& s = sp->slots[g];
      if(s.openingLeft > 0) {               // one opening move per wave; then the game's first root
        const int k = std::min(waves, s.openingLeft);
        playOpeningMoves(sp, (int)g, k);
        s.openingLeft -= k;
        if(s.openingLeft == 0) searchRoot(sp, (int)g);
      }
      else s.wavesLeft = std::max(0, s.wavesLeft - waves);      // searches go on
    }
    for(size_t g = 0; g < sp->slots.size(); g++) if(sp->released[g]) {
      if(sp->slots[g].wavesLeft > 0 || sp->slots[g].openingLeft > 0) throw std::runtime_error("mock: a slot was released before its search had finished");
      sp->released[g] = 0; advance(sp, (int)g);
      const size_t k = 2 * g + ((sp->slots[g].last[1] & 1) ? 1 : 0);       // the new root takes the limits handed over for it
      sp->budget[g] = sp->nextBudget[k]; sp->plain[g] = sp->nextPlain[k];
    }
    sp->log.flush();
  })
}
int kgb_selfplay_release(kgb_selfplay* sp, const uint8_t* mask) { for(size_t g = 0; g < sp->slots.size(); g++) sp->released[g] = mask ? mask[g] : 1; return 0; }
int kgb_selfplay_get_root_visits(kgb_selfplay* sp, int32_t* v) {      // the budget once the search has finished (at once without KGB_MOCK_UNEVEN), one less before
  ensureStarted(sp);
  for(size_t g = 0; g < sp->slots.size(); g++) v[g] = sp->budget[g] - (sp->slots[g].wavesLeft > 0 ? 1 : 0);
  return 0;
}
int kgb_selfplay_set_next_search_limits(kgb_selfplay* sp, const int32_t* visits, const uint8_t* plainRoot, int alsoCurrentRoots) {
  const size_t n = sp->slots.size();
  for(size_t i = 0; i < 2 * n; i++) {
    if(visits[i] < 2 || visits[i] > sp->cfg.max_visits) { g_err = "mock: search limit out of range"; return 1; }
    sp->nextBudget[i] = visits[i]; sp->nextPlain[i] = plainRoot ? plainRoot[i] : 0;
  }
  if(alsoCurrentRoots) for(size_t g = 0; g < n; g++) { sp->budget[g] = sp->nextBudget[2 * g]; sp->plain[g] = sp->nextPlain[2 * g]; }
  return 0;
}
int kgb_selfplay_set_game_setup(kgb_selfplay* sp, const int32_t* setups, int alsoCurrentGames) {
  if(alsoCurrentGames && sp->started) { g_err = "mock: the games in progress have begun"; return 1; }
  for(size_t g = 0; g < sp->slots.size(); g++) {
    const int32_t* q = setups + 4 * g;
    if(q[0] < 2 || q[0] > sp->X || q[1] < 2 || q[1] > sp->Y || q[2] < 0 || q[2] > 3) { g_err = "mock: game setup out of range"; return 1; }
    for(int k = 0; k < 4; k++) { sp->nextSetup[4 * g + k] = q[k]; if(alsoCurrentGames) sp->slots[g].setup[k] = q[k]; }
  }
  return 0;
}
int kgb_selfplay_get_game_setup(kgb_selfplay* sp, int32_t* current, int32_t* lastFinished) {
  for(size_t g = 0; g < sp->slots.size(); g++) for(int k = 0; k < 4; k++) {
    if(current) current[4 * g + k] = sp->slots[g].setup[k];
    if(lastFinished) lastFinished[4 * g + k] = sp->lastSetup[4 * g + k];
  }
  return 0;
}
int kgb_selfplay_set_komi(kgb_selfplay* sp, const float* komi, int alsoCurrentGames) {
  if(alsoCurrentGames && sp->started) { g_err = "mock: the games in progress have begun"; return 1; }
  for(size_t g = 0; g < sp->slots.size(); g++) { sp->nextKomi[g] = komi[g]; if(alsoCurrentGames) sp->slots[g].komi = komi[g]; }
  return 0;
}
int kgb_selfplay_get_komi(kgb_selfplay* sp, float* current, float* lastFinished) {
  for(size_t g = 0; g < sp->slots.size(); g++) { if(current) current[g] = sp->slots[g].komi; if(lastFinished) lastFinished[g] = sp->lastKomi[g]; }
  return 0;
}
int kgb_selfplay_set_policy_init(kgb_selfplay* sp, const int32_t* numMoves, double, int alsoCurrentGames) {
  if(alsoCurrentGames && sp->started) { g_err = "mock: the games in progress have begun"; return 1; }
  for(size_t g = 0; g < sp->slots.size(); g++) { sp->nextInit[g] = numMoves[g]; if(alsoCurrentGames) sp->slots[g].initPending = numMoves[g]; }
  return 0;
}
int kgb_selfplay_get_policy_init(kgb_selfplay* sp, int32_t* movesLeft, int32_t* count, int16_t* moves, int maxMoves) {
  ensureStarted(sp);
  for(size_t g = 0; g < sp->slots.size(); g++) {
    const std::vector<int16_t>& m = sp->slots[g].initMoves;
    if(movesLeft) movesLeft[g] = sp->slots[g].openingLeft;      // (0 unless KGB_MOCK_UNEVEN: then one opening move per wave)
    if(count) count[g] = (int32_t)m.size();
    if(moves) for(int i = 0; i < maxMoves; i++) moves[g * (size_t)maxMoves + i] = i < (int)m.size() ? m[i] : 0;
  }
  return 0;
}
int kgb_selfplay_get_search_limits(kgb_selfplay* sp, int32_t* visits, uint8_t* plainRoot) {
  for(size_t g = 0; g < sp->slots.size(); g++) { if(visits) visits[g] = sp->budget[g]; if(plainRoot) plainRoot[g] = sp->plain[g]; }
  return 0;
}
int kgb_selfplay_get_game(kgb_selfplay* sp, int g, uint8_t* colors, int32_t* info) {
  ensureStarted(sp);
  Slot& s = sp->slots[g];
  memset(colors, 0, (size_t)sp->X * sp->Y);
  for(int y = 0; y < s.setup[1]; y++) for(int x = 0; x < s.setup[0]; x++) colors[y * sp->X + x] = s.board.colors[Location::getLoc(x, y, s.setup[0])];
  info[0] = s.moveNum; info[1] = s.pla == P_BLACK; info[2] = -1; info[3] = 0; info[4] = 0; info[5] = sp->budget[g] + g;
  return 0;
}
int kgb_selfplay_get_root_children(kgb_selfplay* sp, int g, int32_t* visits, float* policy, double* util) {
  Slot& s = sp->slots[g];
  for(size_t i = 0; i < s.edgeVisits.size(); i++) { visits[i] = s.edgeVisits[i]; policy[i] = s.policy[i]; util[i] = 0.0; }
  return 0;
}
int kgb_selfplay_get_root_value_stats(kgb_selfplay* sp, int g, double* child, double* root) {
  Slot& s = sp->slots[g];
  memcpy(child, s.childStats.data(), s.childStats.size() * sizeof(double)); memcpy(root, s.rootStats, sizeof(s.rootStats));
  return 0;
}
int kgb_selfplay_get_play_selection_values(kgb_selfplay* sp, int g, double* v) { Slot& s = sp->slots[g]; memcpy(v, s.psv.data(), s.psv.size() * sizeof(double)); return 0; }
int kgb_selfplay_get_root_extra(kgb_selfplay* sp, int g, int32_t* nv, double* nn) {
  Slot& s = sp->slots[g];
  memcpy(nv, s.nodeVisits.data(), s.nodeVisits.size() * sizeof(int32_t)); memcpy(nn, s.rootNN, sizeof(s.rootNN));
  return 0;
}
int kgb_selfplay_get_root_raw_policy_entropy(kgb_selfplay* sp, double* e) {      // the mock's searches add no noise: the policy as searched
  for(size_t g = 0; g < sp->slots.size(); g++) {
    double h = 0.0;
    for(float p : sp->slots[g].policy) if(p > 1e-100) h += -(double)p * std::log((double)p);
    e[g] = h;
  }
  return 0;
}
int kgb_selfplay_get_root_row(kgb_selfplay* sp, int g, float* spatial, float* global) {
  Slot& s = sp->slots[g];
  memcpy(spatial, s.rowSpatial.data(), s.rowSpatial.size() * sizeof(float)); memcpy(global, s.rowGlobal.data(), 19 * sizeof(float));
  return 0;
}
int kgb_selfplay_get_nn_row(kgb_selfplay* sp, int g, float* spatial, float* global) {
  Slot& s = sp->slots[g];
  memcpy(spatial, s.rowSpatial.data(), s.rowSpatial.size() * sizeof(float)); memcpy(global, s.rowGlobal.data(), 19 * sizeof(float));
  return 0;
}
int kgb_selfplay_get_last_move(kgb_selfplay* sp, int g, int32_t* info, float* score, uint8_t* colors, uint8_t* area) {
  Slot& s = sp->slots[g];
  memcpy(info, s.last, sizeof(s.last)); *score = s.lastScore;
  for(size_t i = 0; i < s.finalColors.size(); i++) { colors[i] = s.finalColors[i]; area[i] = s.finalArea[i]; }
  return 0;
}
int kgb_selfplay_get_stats(kgb_selfplay*, kgb_selfplay_stats* out) { memset(out, 0, sizeof(*out)); return 0; }
int kgb_selfplay_play_moves(kgb_selfplay*, const int8_t*, int) { g_err = "mock: play_moves is not scripted"; return 1; }
// the host plays moves into one slot (a side loop's position; the opponent's move in match play): a move that ends the game restarts the slot with
// the setup and komi handed over for its next game, like the loop's own moves; the resulting root is searched (instantly) and logged
int kgb_selfplay_play_moves_game(kgb_selfplay* sp, int g, const int8_t* xy, int n) {
  try {
    ensureStarted(sp);
    Slot& s = sp->slots[g];
    if(s.openingLeft > 0) throw std::runtime_error("mock: play_moves_game on a slot that is in its opening");
    sp->log << "{\"ev\":\"playmoves\",\"slot\":" << g << ",\"moves\":[";
    for(int i = 0; i < n; i++) {
      const int x = xy[2 * i];
      const int y = xy[2 * i + 1];
      const Loc loc = x < 0 ? Board::PASS_LOC : Location::getLoc(x, y, s.setup[0]);
      if(x >= s.setup[0] || y >= s.setup[1] || !s.hist.isLegal(s.board, loc, s.pla)) throw std::runtime_error("mock: illegal move in play_moves_game");
      sp->log << (i ? "," : "") << "[" << x << "," << y << "]";
      advance(sp, g, false, loc);
    }
    sp->log << "]}\n";
    searchRoot(sp, g);
    sp->log.flush();
    return 0;
  }
  catch(const std::exception& e) { g_err = e.what(); return 1; }
}
}
