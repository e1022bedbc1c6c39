// Training-data writer of the C++ host: FinishedGame -> rows -> <16 hex>.npz, and the game record (.sgfs line).
//
// Plain C++17 (zlib for the zip container; the uint32 stream of the reference's `Rand` comes from the library, kgb_rand_uint32_stream).
// No reference headers: this is the stand-alone twin of katago_b200/npz_writer.py, function by function, and restates
//   TrainingWriteBuffers::addRow          dataio/trainingwrite.cpp:448-852   (TD value targets, lead, weights, history masks, game hash,
//                                                                             metadata, score distribution, ownership / future boards /
//                                                                             scoring planes, Q targets; stochastic rounding drawn from
//                                                                             `Rand` in the reference's order)
//   TrainingWriteBuffers::writeToZipFile  dataio/trainingwrite.cpp:854-886 + NumpyBuffer headers dataio/numpywrite.cpp:97-226
//   TrainingDataWriter::writeGame         dataio/trainingwrite.cpp:1097-1325 (main-line rows; side positions, reanalysis and net changes are
//                                                                             not produced by this host and not restated here)
//   WriteSgf::writeSgf(FinishedGameData)  dataio/sgf.cpp:1997-2226
// for the rule subset of the device loop (area scoring, no tax, no button, no handicap).  Float expressions keep the precision the
// reference computes them in (float where it uses float), so that rows are equal to the Python writer's - which is pinned bit for bit to
// the reference's own addRow / writeGame (tests/golden/addrow_*.json.gz, writegame_*.json.gz) - bit for bit:
// tests/test_cpp_host.py compares whole .npz files of both hosts on the same games.
#pragma once
#include <zlib.h>

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/kgb200.h"

namespace b200 {

constexpr int NUM_BIN = 22, NUM_GLOBAL = 19;                  // NNInputs::NUM_FEATURES_SPATIAL_V7 / GLOBAL_V7
constexpr int POLICY_TARGET_CHANNELS = 2, GLOBAL_TARGET_CHANNELS = 80, VALUE_SPATIAL_CHANNELS = 5, QVALUE_CHANNELS = 3;   // trainingwrite.cpp:276-279
constexpr int SCORE_DISTR_RADIUS = 60;                        // NNPos::EXTRA_SCORE_DISTR_RADIUS
constexpr int P_BLACK = 1, P_WHITE = 2;

// The reference's Rand as far as the writer uses it (core/rand.h:245-274): nextUInt, nextDouble (53 bits of two nextUInt, low word first), nextBool.
class RowRand {
 public:
  explicit RowRand(const std::string& seed, size_t prefetch = 4096) : seed_(seed) { grow(prefetch); }
  uint32_t nextUInt() { if(i_ >= buf_.size()) grow(buf_.size() * 2); return buf_[i_++]; }
  double nextDouble() { uint64_t lo = nextUInt(), hi = nextUInt(); return (double)((lo | (hi << 32)) & ((1ULL << 53) - 1)) / (double)(1ULL << 53); }
  bool nextBool(double prob) { return nextDouble() < prob; }
 private:
  void grow(size_t n) {        // a longer stream of the same seed has the old one as its prefix
    buf_.resize(n);
    if(kgb_rand_uint32_stream(seed_.c_str(), (int)n, buf_.data()) != 0) throw std::runtime_error(std::string("libkgb200: ") + kgb_last_error());
  }
  std::string seed_; std::vector<uint32_t> buf_; size_t i_ = 0;
};

struct ValueTargets { float win = 0, loss = 0, noResult = 0, score = 0; bool hasLead = false; float lead = 0; };
struct PolicyTargetMove { int x, y; int16_t value; };                       // x < 0: pass
struct QValueTarget { int x, y; float winLoss, score; int visits; };       // white's perspective; visits of the child NODE

// A position off the main line that was searched on its own (dataio/trainingwrite.h:62-84): it gets a row with its own policy / value / Q
// targets and none of the targets that need the game's continuation.
struct SidePosition {
  int nextPlayer = P_BLACK, turnIdx = 0;
  std::vector<uint8_t> packedInput; std::array<float, NUM_GLOBAL> globalInput{};
  std::vector<PolicyTargetMove> policyTarget; int64_t unreducedNumVisits = 0;
  ValueTargets whiteValueTargets; std::vector<QValueTarget> whiteQValueTargets;
  double policySurprise = 0, policyEntropy = 0, searchEntropy = 0; std::array<double, 3> nnRawStats{};
  float targetWeight = 1.0f;
};

// What Play::runGame hands to the writer (dataio/trainingwrite.h:84-170), for games of this host.
struct FinishedGame {
  int xSize = 19, ySize = 19; float komi = 7.5f;
  uint64_t gameHash[2] = {0, 0};
  double drawEquivalentWinsForWhite = 0.5;
  bool hitTurnLimit = false, endFinished = true, endNoResult = false;
  int mode = 0, startHistMoves = 0, initialTurnNumber = 0;
  float trainingWeight = 1.0f;
  std::vector<std::vector<uint8_t>> boardsByTurn;         // nTurns + 1 boards (row-major colours): before each move, and the final one
  std::vector<int> nextPlayerByTurn;
  std::vector<std::vector<uint8_t>> packedInputByTurn;    // [22][ceil(L*L/8)]
  std::vector<std::array<float, NUM_GLOBAL>> globalInputByTurn;
  std::vector<float> targetWeightByTurn, targetWeightByTurnUnrounded;
  std::vector<std::vector<PolicyTargetMove>> policyTargetsByTurn; std::vector<int64_t> unreducedNumVisitsByTurn;
  std::vector<double> policySurpriseByTurn, policyEntropyByTurn, searchEntropyByTurn;
  std::vector<ValueTargets> whiteValueTargetsByTurn;      // one more than turns: the outcome
  std::vector<std::vector<QValueTarget>> whiteQValueTargetsByTurn;
  std::vector<std::array<double, 3>> nnRawStatsByTurn;    // whiteWinLoss, whiteScoreMean, policyEntropy of the root's own evaluation
  std::vector<std::pair<int, int>> moves;                 // (x, y), (-1, -1) = pass
  std::vector<std::pair<int, int>> startMoves;            // the moves before the training period (startHist.moveHistory: policy-initialised opening), black first
  std::string koRule = "SIMPLE"; bool multiStoneSuicideLegal = true;
  int winner = 0; float finalWhiteMinusBlackScore = 0; bool resigned = false;      // resigned: BoardHistory::isResignation (match play)
  std::vector<uint8_t> finalFullArea, finalOwnership;     // row-major [ySize * xSize]: 0 none, 1 black, 2 white
  std::vector<float> finalWhiteScoring;
  std::vector<std::shared_ptr<SidePosition>> sidePositions;

  // BoardHistory::currentSelfKomi (game/boardhistory.cpp:570-589) without bonus points
  float selfKomi(int nextPlayer) const {
    const bool komiIsInt = (float)(int)komi == komi;
    const float adj = komiIsInt ? (float)(drawEquivalentWinsForWhite - 0.5) : 0.0f;
    const float w = komi + adj;
    return nextPlayer == P_WHITE ? w : -w;
  }
};

inline long cRound(double x) { return x >= 0 ? (long)std::floor(x + 0.5) : -(long)std::floor(-x + 0.5); }     // C round(): halves away from zero

// clampToRadius120 / clampToRadius32000 (trainingwrite.cpp:358-383): a float to an integer whose expectation is the float
inline int clampToRadius(float x, int radius, RowRand& rand) {
  const int low = (int)std::floor((double)x), high = low + 1;
  if(low < -radius) return -radius;
  if(high > radius) return radius;
  const float lam = x - (float)low;
  if(lam == 0.0f) return low;
  return rand.nextBool((double)lam) ? high : low;
}

// The outcome entry of the value targets (program/play.cpp:1977-2000)
inline ValueTargets finalValueTargets(int winner, float finalWhiteMinusBlackScore, double drawEquivalentWinsForWhite, float komi, bool noResult) {
  ValueTargets t;
  if(noResult) { t.noResult = 1.0f; return t; }
  t.win = (float)(winner == P_WHITE ? 1.0 : winner == P_BLACK ? 0.0 : drawEquivalentWinsForWhite);
  t.loss = 1.0f - t.win;
  const bool komiIsInt = (float)(int)komi == komi;
  const double adj = komiIsInt ? (double)(float)(drawEquivalentWinsForWhite - 0.5) : 0.0;      // whiteKomiAdjustmentForDraws returns float
  t.score = (float)((double)finalWhiteMinusBlackScore + adj);
  t.hasLead = true; t.lead = t.score;
  return t;
}

// NumpyBuffer's 256-byte header (dataio/numpywrite.cpp:97-226): magic, version 1.0, length 246, the dict without spaces, space padding, newline
inline std::string npyHeader(const std::string& descr, const std::vector<size_t>& shape) {
  std::string d = "{'descr':'" + descr + "','fortran_order':False,'shape':(";
  for(size_t i = 0; i < shape.size(); i++) d += (i ? "," : "") + std::to_string(shape[i]);
  d += ")}";
  if(10 + d.size() >= 256) throw std::runtime_error("numpy header too long");
  std::string h("\x93NUMPY\x01\x00", 8);
  h += (char)(246 & 0xFF); h += (char)(246 >> 8);
  h += d;
  h += std::string(256 - 11 - d.size(), ' ');
  h += '\n';
  return h;
}

// A zip archive with deflated members (what np.load and the reference's libzip output have in common)
class ZipWriter {
 public:
  explicit ZipWriter(const std::string& path) : f_(std::fopen(path.c_str(), "wb")) { if(!f_) throw std::runtime_error("cannot write " + path); }
  ~ZipWriter() { if(f_) std::fclose(f_); }
  void add(const std::string& name, const std::string& head, const void* data, size_t bytes) {
    std::vector<uint8_t> raw(head.size() + bytes);
    std::memcpy(raw.data(), head.data(), head.size());
    if(bytes) std::memcpy(raw.data() + head.size(), data, bytes);
    z_stream zs; std::memset(&zs, 0, sizeof(zs));
    if(deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("zlib: deflateInit2");
    std::vector<uint8_t> comp(deflateBound(&zs, (uLong)raw.size()));
    zs.next_in = raw.data(); zs.avail_in = (uInt)raw.size(); zs.next_out = comp.data(); zs.avail_out = (uInt)comp.size();
    if(deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); throw std::runtime_error("zlib: deflate"); }
    comp.resize(zs.total_out);
    deflateEnd(&zs);
    if(raw.size() >= 0xFFFFFFFFULL || comp.size() >= 0xFFFFFFFFULL || offset_ >= 0xFFFFFFFFULL) throw std::runtime_error("npz member too large (no zip64)");
    Entry e{name, (uint32_t)crc32(crc32(0L, Z_NULL, 0), raw.data(), (uInt)raw.size()), (uint32_t)comp.size(), (uint32_t)raw.size(), (uint32_t)offset_};
    std::string h;
    u32(h, 0x04034b50); u16(h, 20); u16(h, 0); u16(h, 8); u16(h, 0); u16(h, 0x21); u32(h, e.crc); u32(h, e.csize); u32(h, e.usize);
    u16(h, (uint16_t)name.size()); u16(h, 0); h += name;
    put(h.data(), h.size()); put(comp.data(), comp.size());
    entries_.push_back(e);
  }
  void close() {
    const uint64_t cdStart = offset_;
    for(const Entry& e : entries_) {
      std::string h;
      u32(h, 0x02014b50); u16(h, 20); u16(h, 20); u16(h, 0); u16(h, 8); u16(h, 0); u16(h, 0x21); u32(h, e.crc); u32(h, e.csize); u32(h, e.usize);
      u16(h, (uint16_t)e.name.size()); u16(h, 0); u16(h, 0); u16(h, 0); u16(h, 0); u32(h, 0); u32(h, e.offset); h += e.name;
      put(h.data(), h.size());
    }
    std::string h;
    u32(h, 0x06054b50); u16(h, 0); u16(h, 0); u16(h, (uint16_t)entries_.size()); u16(h, (uint16_t)entries_.size());
    u32(h, (uint32_t)(offset_ - cdStart)); u32(h, (uint32_t)cdStart); u16(h, 0);
    put(h.data(), h.size());
    if(std::fclose(f_) != 0) { f_ = nullptr; throw std::runtime_error("npz: write failed"); }
    f_ = nullptr;
  }
 private:
  struct Entry { std::string name; uint32_t crc, csize, usize, offset; };
  static void u16(std::string& s, uint16_t v) { s += (char)(v & 0xFF); s += (char)(v >> 8); }
  static void u32(std::string& s, uint32_t v) { for(int i = 0; i < 4; i++) s += (char)((v >> (8 * i)) & 0xFF); }
  void put(const void* p, size_t n) { if(n && std::fwrite(p, 1, n, f_) != n) throw std::runtime_error("npz: write failed"); offset_ += n; }
  FILE* f_; uint64_t offset_ = 0; std::vector<Entry> entries_;
};

// Row buffers of one output file (dataio/trainingwrite.h:245-352).  Boards are xSize * ySize inside a dataLen * dataLen frame.
class TrainingWriteBuffers {
 public:
  TrainingWriteBuffers(int maxRows, int dataLen) : L(dataLen), A(dataLen * dataLen), maxRows(maxRows), packedLen((dataLen * dataLen + 7) / 8),
      sdLen(2 * (dataLen * dataLen + SCORE_DISTR_RADIUS)) {
    binaryInput.resize((size_t)maxRows * NUM_BIN * packedLen); globalInput.resize((size_t)maxRows * NUM_GLOBAL);
    policyTargets.resize((size_t)maxRows * POLICY_TARGET_CHANNELS * (A + 1)); globalTargets.resize((size_t)maxRows * GLOBAL_TARGET_CHANNELS);
    scoreDistr.resize((size_t)maxRows * sdLen); valueTargets.resize((size_t)maxRows * VALUE_SPATIAL_CHANNELS * A);
    qValueTargets.resize((size_t)maxRows * QVALUE_CHANNELS * (A + 1));
  }
  const int L, A, maxRows, packedLen, sdLen;
  int curRows = 0;
  std::vector<uint8_t> binaryInput; std::vector<float> globalInput; std::vector<int16_t> policyTargets; std::vector<float> globalTargets;
  std::vector<int8_t> scoreDistr, valueTargets; std::vector<int16_t> qValueTargets;

  int posOf(int x, int y) const { return x < 0 ? A : y * L + x; }      // NNPos::locToPos

  // One main-line row of `game` for turn `idx` (addRow with valueTargetWeight = tdValueTargetWeight = leadTargetWeightFactor = 1, no
  // reanalysis, no side position, no net changes, no bonus points).  policyTarget1: the next turn's policy target or nullptr.
  // side != nullptr: the row of a side position of `game` instead (writeGame's second loop, trainingwrite.cpp:1258-1323): its own targets only -
  // the value targets are its own search's (one entry), no outcome, ownership, future boards or scoring; the game's ending still decides the
  // lead / finished flags.  idx is ignored then.
  void addRow(const FinishedGame& game, int idx, const std::vector<PolicyTargetMove>* policyTarget1, RowRand& rand, const SidePosition* side = nullptr) {
    if(curRows >= maxRows) throw std::runtime_error("TrainingWriteBuffers full");
    if(side) idx = 0;
    const int r = curRows, P = A + 1, xSize = game.xSize, ySize = game.ySize, nextPlayer = side ? side->nextPlayer : game.nextPlayerByTurn[idx];
    const bool white = nextPlayer == P_WHITE;
    const int opp = white ? P_BLACK : P_WHITE;
    const std::vector<uint8_t>& packed = side ? side->packedInput : game.packedInputByTurn[idx];
    if((int)packed.size() != NUM_BIN * packedLen) throw std::runtime_error("addRow: packed input of the wrong size");
    std::memcpy(&binaryInput[(size_t)r * NUM_BIN * packedLen], packed.data(), (size_t)NUM_BIN * packedLen);
    std::memcpy(&globalInput[(size_t)r * NUM_GLOBAL], side ? side->globalInput.data() : game.globalInputByTurn[idx].data(), sizeof(float) * NUM_GLOBAL);
    float* g = &globalTargets[(size_t)r * GLOBAL_TARGET_CHANNELS];
    std::fill(g, g + GLOBAL_TARGET_CHANNELS, 0.0f);
    g[25] = game.trainingWeight;
    int16_t* pol = &policyTargets[(size_t)r * POLICY_TARGET_CHANNELS * P];
    const std::vector<PolicyTargetMove>* targets[2] = {side ? &side->policyTarget : &game.policyTargetsByTurn[idx], side ? nullptr : policyTarget1};
    const int weightCol[2] = {26, 28};
    for(int ch = 0; ch < 2; ch++) {
      if(targets[ch] == nullptr) { std::fill(pol + ch * P, pol + (ch + 1) * P, (int16_t)1); g[weightCol[ch]] = 0.0f; }      // uniformPolicyTarget, weight 0
      else {
        std::fill(pol + ch * P, pol + (ch + 1) * P, (int16_t)0);
        for(const PolicyTargetMove& m : *targets[ch]) pol[ch * P + posOf(m.x, m.y)] = m.value;
        g[weightCol[ch]] = 1.0f;
      }
    }
    const int boardArea = xSize * ySize;
    const std::vector<ValueTargets> sideTargets = side ? std::vector<ValueTargets>{side->whiteValueTargets} : std::vector<ValueTargets>();
    const std::vector<ValueTargets>& vt = side ? sideTargets : game.whiteValueTargetsByTurn;
    const double nowFactors[5] = {0.0, 1.0 / (1.0 + boardArea * 0.176), 1.0 / (1.0 + boardArea * 0.056), 1.0 / (1.0 + boardArea * 0.016), 1.0};
    for(int k = 0; k < 5; k++) valueTDTargets(vt, idx, white, nowFactors[k], g + 4 * k);
    const float vtw = 1.0f, tdw = 1.0f;
    const bool noResultEnd = game.endFinished && game.endNoResult;
    const float cap = (float)(19 * 19 + SCORE_DISTR_RADIUS);         // NNPos::MAX_BOARD_AREA + EXTRA_SCORE_DISTR_RADIUS
    if(vt[idx].hasLead && !noResultEnd) {
      const float lead = white ? vt[idx].lead : -vt[idx].lead;
      g[21] = std::min(std::max(lead, -cap), cap);
      g[29] = vtw * 1.0f;
    }
    double s = 0.0;
    for(size_t i = (size_t)idx + 1; i < vt.size(); i++) {
      const double prevWL = (double)(vt[i - 1].win - vt[i - 1].loss), nextWL = (double)(vt[i].win - vt[i].loss);
      s += (double)(i - idx) * ((nextWL - prevWL) * (nextWL - prevWL));
    }
    g[22] = (float)s;
    g[24] = 1.0f - tdw;
    g[30] = (float)(side ? side->policySurprise : game.policySurpriseByTurn[idx]); g[31] = (float)(side ? side->policyEntropy : game.policyEntropyByTurn[idx]);
    g[32] = (float)(side ? side->searchEntropy : game.searchEntropyByTurn[idx]);
    g[35] = 1.0f - vtw;
    bool use = true;
    for(int k = 0; k < 5; k++) {                 // each earlier history step is kept with probability 0.98 (:628-637)
      use = use && rand.nextDouble() < 0.98;     // (no draw once a step has been dropped)
      g[36 + k] = use ? 1.0f : 0.0f;
    }
    const uint64_t h0 = game.gameHash[0], h1 = game.gameHash[1];
    g[41] = (float)(h0 & 0x3FFFFF); g[42] = (float)((h0 >> 22) & 0x3FFFFF); g[43] = (float)((h0 >> 44) & 0xFFFFF);
    g[44] = (float)(h1 & 0x3FFFFF); g[45] = (float)((h1 >> 22) & 0x3FFFFF); g[46] = (float)((h1 >> 44) & 0xFFFFF);
    g[47] = game.selfKomi(nextPlayer);
    g[48] = 1.0f;                                // area scoring
    g[51] = (float)(side ? side->turnIdx : idx + game.startHistMoves);
    g[52] = game.hitTurnLimit ? 1.0f : 0.0f;
    g[53] = (float)game.startHistMoves;
    g[55] = (float)game.mode;
    g[56] = (float)game.initialTurnNumber;
    const std::array<double, 3>& raw = side ? side->nnRawStats : game.nnRawStatsByTurn[idx];
    g[57] = (float)(white ? raw[0] : -raw[0]);
    g[58] = (float)(white ? raw[1] : -raw[1]);
    g[59] = (float)raw[2];
    g[60] = (float)(side ? side->unreducedNumVisits : game.unreducedNumVisitsByTurn[idx]);
    g[62] = (!side && game.endFinished && !game.hitTurnLimit) ? 1.0f : 0.0f;
    g[63] = 3.0f;

    int8_t* sd = &scoreDistr[(size_t)r * sdLen];
    int8_t* own = &valueTargets[(size_t)r * VALUE_SPATIAL_CHANNELS * A];
    std::fill(sd, sd + sdLen, (int8_t)0);
    std::fill(own, own + (size_t)VALUE_SPATIAL_CHANNELS * A, (int8_t)0);
    const int sdMid = A + SCORE_DISTR_RADIUS;
    auto frame = [&](int j) { return (j / xSize) * L + (j % xSize); };      // NNPos::xyToPos of the board's j-th point
    if(side || game.finalOwnership.empty() || noResultEnd) { sd[sdMid - 1] = 50; sd[sdMid] = 50; }
    else {
      g[27] = vtw;
      const float score = white ? vt.back().score : -vt.back().score;
      g[20] = score;
      for(int j = 0; j < boardArea; j++) {
        const int fo = game.finalOwnership[j], fa = game.finalFullArea[j];
        own[frame(j)] = fo == nextPlayer ? 1 : fo == opp ? -1 : 0;
        own[A + frame(j)] = (fa != 0 && fo == 0) ? (fa == nextPlayer ? 1 : -1) : 0;
      }
      const long center = cRound((double)score);
      const long lower = center + sdMid - 1, upper = center + sdMid;
      if(upper <= 0) sd[0] = 100;
      else if(lower >= sdLen - 1) sd[sdLen - 1] = 100;
      else {
        const float lam = score - ((float)center - 0.5f);
        const long up = cRound((double)(lam * 100.0f));
        sd[lower] = (int8_t)(100 - up);
        sd[upper] = (int8_t)up;
      }
    }
    if(!side) {                                  // posHistForFutureBoards: the game's own positions 8 and 32 turns ahead
      if(game.boardsByTurn.size() != vt.size()) throw std::runtime_error("addRow: one board per value target expected");
      g[33] = 1.0f;
      const int end = (int)game.boardsByTurn.size() - 1;
      const int ahead[2] = {8, 32};
      for(int c = 0; c < 2; c++) {
        const std::vector<uint8_t>& b = game.boardsByTurn[std::min(idx + ahead[c], end)];
        for(int j = 0; j < boardArea; j++) own[(size_t)(2 + c) * A + frame(j)] = b[j] == nextPlayer ? 1 : b[j] == opp ? -1 : 0;
      }
    }
    if(!side && !game.finalWhiteScoring.empty() && !noResultEnd) {
      g[34] = vtw;
      for(int j = 0; j < boardArea; j++) {       // y, x order: the order the reference draws its random numbers in
        const float v = white ? game.finalWhiteScoring[j] : -game.finalWhiteScoring[j];
        own[(size_t)4 * A + frame(j)] = (int8_t)clampToRadius(v * 120.0f, 120, rand);
      }
    }
    int16_t* q = &qValueTargets[(size_t)r * QVALUE_CHANNELS * P];
    std::fill(q, q + (size_t)QVALUE_CHANNELS * P, (int16_t)0);
    for(const QValueTarget& t : (side ? side->whiteQValueTargets : game.whiteQValueTargetsByTurn[idx])) {        // fillQValueTarget (:385-409)
      const int pos = posOf(t.x, t.y);
      const float wl = white ? t.winLoss : -t.winLoss;
      float sc = white ? t.score : -t.score;
      sc = std::min(std::max(sc, -cap), cap);
      q[pos] = (int16_t)clampToRadius(wl * 32000.0f, 32000, rand);
      q[P + pos] = (int16_t)clampToRadius(sc * 60.0f, 32000, rand);
      q[2 * P + pos] = (int16_t)std::max(0, std::min(t.visits, 32000));
    }
    curRows++;
  }

  // writeToZipFile (trainingwrite.cpp:854-886): the first curRows rows of every array, each a .npy under its bare name
  void writeToZipFile(const std::string& path) const {
    ZipWriter z(path);
    const size_t n = (size_t)curRows, P = (size_t)A + 1;
    z.add("binaryInputNCHWPacked", npyHeader("|u1", {n, (size_t)NUM_BIN, (size_t)packedLen}), binaryInput.data(), n * NUM_BIN * packedLen);
    z.add("globalInputNC", npyHeader("<f4", {n, (size_t)NUM_GLOBAL}), globalInput.data(), n * NUM_GLOBAL * sizeof(float));
    z.add("policyTargetsNCMove", npyHeader("<i2", {n, (size_t)POLICY_TARGET_CHANNELS, P}), policyTargets.data(), n * POLICY_TARGET_CHANNELS * P * sizeof(int16_t));
    z.add("globalTargetsNC", npyHeader("<f4", {n, (size_t)GLOBAL_TARGET_CHANNELS}), globalTargets.data(), n * GLOBAL_TARGET_CHANNELS * sizeof(float));
    z.add("scoreDistrN", npyHeader("|i1", {n, (size_t)sdLen}), scoreDistr.data(), n * sdLen);
    z.add("valueTargetsNCHW", npyHeader("|i1", {n, (size_t)VALUE_SPATIAL_CHANNELS, (size_t)L, (size_t)L}), valueTargets.data(), n * VALUE_SPATIAL_CHANNELS * A);
    z.add("qValueTargetsNCMove", npyHeader("<i2", {n, (size_t)QVALUE_CHANNELS, P}), qValueTargets.data(), n * QVALUE_CHANNELS * P * sizeof(int16_t));
    z.close();
  }

 private:
  // fillValueTDTargets (trainingwrite.cpp:411-446): exponentially weighted average of the value targets from this turn on
  static void valueTDTargets(const std::vector<ValueTargets>& vt, int idx, bool whiteToMove, double nowFactor, float* out) {
    double win = 0, loss = 0, noResult = 0, score = 0, weightLeft = 1.0;
    const size_t n = vt.size();
    for(size_t i = (size_t)idx; i < n; i++) {
      double weightNow;
      if(i == n - 1) { weightNow = weightLeft; weightLeft = 0.0; }
      else { weightNow = weightLeft * nowFactor; weightLeft *= (1.0 - nowFactor); }
      const double w = vt[i].win, l = vt[i].loss, nr = vt[i].noResult, sc = vt[i].score;
      win += weightNow * (whiteToMove ? w : l);
      loss += weightNow * (whiteToMove ? l : w);
      noResult += weightNow * nr;
      score += weightNow * (whiteToMove ? sc : -sc);
    }
    const double cap = 19 * 19 + SCORE_DISTR_RADIUS;
    score = std::min(std::max(score, -cap), cap);
    out[0] = (float)win; out[1] = (float)loss; out[2] = (float)noResult; out[3] = (float)score;
  }
};

// The reference's writer (dataio/trainingwrite.cpp:987-1325): rows of finished games go into a TrainingWriteBuffers that is written out as
// <16 hex digits>.npz whenever it is full; the first file is cut short at random.  One Rand serves the first-file size, the fractional target
// weights, addRow's rounding and the file names, in the reference's order.
class TrainingDataWriter {
 public:
  TrainingDataWriter(const std::string& outputDir, int maxRowsPerFile, double firstFileMinRandProp, int dataLen, const std::string& randSeed)
      : outputDir_(outputDir), rand_(randSeed), buffers_(maxRowsPerFile, dataLen) {
    if(!(firstFileMinRandProp >= 0.0 && firstFileMinRandProp <= 1.0)) throw std::invalid_argument("firstFileMinRandProp not in [0,1]");
    firstFileMaxRows_ = firstFileMinRandProp >= 1.0 ? maxRowsPerFile : maxRowsPerFile - (int)(maxRowsPerFile * (1.0 - firstFileMinRandProp) * rand_.nextDouble());
  }
  int64_t rowCount() const { return rowCount_; }
  const std::vector<std::string>& filesWritten() const { return files_; }

  std::string flushIfNonempty() {
    if(buffers_.curRows <= 0) return "";
    isFirstFile_ = false;
    const uint64_t lo = rand_.nextUInt(), hi = rand_.nextUInt();
    char name[32];
    std::snprintf(name, sizeof(name), "%016llX.npz", (unsigned long long)(lo | (hi << 32)));
    const std::string path = outputDir_ + "/" + name;
    buffers_.writeToZipFile(path + ".tmp");
    if(std::rename((path + ".tmp").c_str(), path.c_str()) != 0) throw std::runtime_error("cannot rename " + path + ".tmp");
    buffers_.curRows = 0;
    files_.push_back(path);
    return path;
  }

  // writeGame (:1097-1325), main-line rows: a turn with target weight w gives floor(w) rows plus one more with probability frac(w); policy
  // target 1 is the next turn's policy target
  void writeGame(const FinishedGame& game) {
    const size_t n = game.targetWeightByTurn.size();
    if(game.policyTargetsByTurn.size() != n || game.whiteQValueTargetsByTurn.size() != n || game.nnRawStatsByTurn.size() != n ||
       game.whiteValueTargetsByTurn.size() != n + 1 || game.boardsByTurn.size() != n + 1 || game.unreducedNumVisitsByTurn.size() != n)
      throw std::runtime_error("FinishedGame: per-turn lists disagree in length");
    if(!game.endFinished && !game.hitTurnLimit) throw std::runtime_error("FinishedGame: unfinished game that did not hit the turn limit");
    for(size_t t = 0; t < n; t++) {
      double targetWeight = (double)game.targetWeightByTurn[t];
      const std::vector<PolicyTargetMove>* policy1 = t + 1 < n ? &game.policyTargetsByTurn[t + 1] : nullptr;
      while(targetWeight > 0.0) {
        if(targetWeight >= 1.0 || rand_.nextBool(targetWeight)) {
          buffers_.addRow(game, (int)t, policy1, rand_);
          if(buffers_.curRows >= buffers_.maxRows || (isFirstFile_ && buffers_.curRows >= firstFileMaxRows_)) flushIfNonempty();
          rowCount_++;
        }
        targetWeight -= 1.0;
      }
    }
    for(const std::shared_ptr<SidePosition>& sp : game.sidePositions) {      // side rows (:1258-1323)
      double targetWeight = (double)sp->targetWeight;
      while(targetWeight > 0.0) {
        if(targetWeight >= 1.0 || rand_.nextBool(targetWeight)) {
          buffers_.addRow(game, 0, nullptr, rand_, sp.get());
          if(buffers_.curRows >= buffers_.maxRows || (isFirstFile_ && buffers_.curRows >= firstFileMaxRows_)) flushIfNonempty();
          rowCount_++;
        }
        targetWeight -= 1.0;
      }
    }
  }

 private:
  std::string outputDir_; RowRand rand_; TrainingWriteBuffers buffers_;
  bool isFirstFile_ = true; int firstFileMaxRows_ = 0; int64_t rowCount_ = 0; std::vector<std::string> files_;
};

// WriteSgf::writeSgf with a FinishedGameData (dataio/sgf.cpp:1997-2226, as called from program/selfplaymanager.cpp:377): root properties, the
// game comment (startTurnIdx, initTurnNum, gameHash, gtype) and per move the value targets, visits and target weight.
inline std::string writeSgf(const FinishedGame& d, const std::string& bName, const std::string& wName) {
  static const char* SGF_CHARS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ";
  static const char* GTYPES[] = {"normal", "cleanuptraining", "fork", "handicap", "sgfpos", "hintpos", "hintfork", "asym"};
  auto fmt = [](const char* f, double v) { char b[64]; std::snprintf(b, sizeof(b), f, v); return std::string(b); };
  auto g = [&](float v) { return fmt("%g", (double)v); };                    // ostream << float
  std::string out = "(;FF[4]GM[1]";
  out += d.xSize == d.ySize ? "SZ[" + std::to_string(d.xSize) + "]" : "SZ[" + std::to_string(d.xSize) + ":" + std::to_string(d.ySize) + "]";
  out += "PB[" + bName + "]PW[" + wName + "]HA[0]KM[" + g(d.komi) + "]";
  out += "RU[ko" + d.koRule + "scoreAREAtaxNONEsui" + (d.multiStoneSuicideLegal ? "1" : "0") + "]";
  std::string result;
  if(d.endFinished) {
    if(d.endNoResult) result = "Void";
    else if(d.resigned) result = d.winner == P_BLACK ? "B+R" : "W+R";          // WriteSgf::printGameResult
    else if(d.winner == P_BLACK) result = "B+" + g(-d.finalWhiteMinusBlackScore);
    else if(d.winner == P_WHITE) result = "W+" + g(d.finalWhiteMinusBlackScore);
    else result = "0";
    out += "RE[" + result + "]";
  }
  char hash[40];
  std::snprintf(hash, sizeof(hash), "%016llX%016llX", (unsigned long long)d.gameHash[1], (unsigned long long)d.gameHash[0]);
  out += "C[startTurnIdx=" + std::to_string(d.startHistMoves) + ",initTurnNum=" + std::to_string(d.initialTurnNumber) + ",gameHash=" + hash + ",gtype=" +
         (d.mode >= 0 && d.mode < 8 ? GTYPES[d.mode] : "other") + "]";
  const std::vector<float>& weights = d.targetWeightByTurnUnrounded.empty() ? d.targetWeightByTurn : d.targetWeightByTurnUnrounded;
  const size_t n = d.moves.size();
  for(size_t j = 0; j < d.startMoves.size(); j++) {       // endHist.moveHistory starts with startHist's moves: no comments on those
    out += std::string(";") + (j % 2 == 0 ? "B" : "W") + "[";
    if(d.startMoves[j].first >= 0) { out += SGF_CHARS[d.startMoves[j].first]; out += SGF_CHARS[d.startMoves[j].second]; }
    out += "]";
  }
  for(size_t i = 0; i < n; i++) {
    const int x = d.moves[i].first, y = d.moves[i].second;
    out += std::string(";") + (d.nextPlayerByTurn[i] == P_BLACK ? "B" : "W") + "[";
    if(x >= 0) { out += SGF_CHARS[x]; out += SGF_CHARS[y]; }
    out += "]";
    std::string parts;
    auto part = [&](const std::string& s) { parts += (parts.empty() ? "" : " ") + s; };
    if(i < d.whiteValueTargetsByTurn.size()) {
      const ValueTargets& t = d.whiteValueTargetsByTurn[i];
      part(fmt("%.2f", t.win) + " " + fmt("%.2f", t.loss) + " " + fmt("%.2f", t.noResult) + " " + fmt("%.1f", t.score));
    }
    if(i < d.unreducedNumVisitsByTurn.size()) part("v=" + std::to_string(d.unreducedNumVisitsByTurn[i]));
    if(i < weights.size()) part("weight=" + fmt("%.2f", weights[i]));
    if(d.endFinished && i + 1 == n) part("result=" + result);
    if(!parts.empty()) out += "C[" + parts + "]";
  }
  out += ")";
  return out;
}

}  // namespace b200
