// Ko rules and game end on top of the warp bitboard: the part of BoardHistory (game/boardhistory.cpp) that the main phase of an
// area-scored game needs - SURVEY.md §8a row a3.
//   makeBoardMoveAssumeLegal :932-1163   pass counting, ko-hash history, spight-like ending passes under simple ko (a pass in a
//                                        situation where the same player already passed ends the game), superko bans for the next
//                                        player (positional / situational), two passes end the game, long cycles under simple ko
//                                        are "no result" (third occurrence of a situation since the last pass)
//   isLegal :786-812                     board legality and not superko-banned
//   passWouldEndPhase :874-880
// Ko rules: simple, positional, situational, spight.  Not here: territory scoring, encore phases, button, handicap bonus.
//
// Histories are lists of 64-bit ko hashes (first half of the 128-bit Zobrist position hash, XOR a player constant unless the
// rule is positional).  A position reached inside a search extends its game's lists by a path part; HistLists carries both.
#pragma once
#include "kgb_board.cuh"

namespace kgb {

enum { KGB_KO_SIMPLE = 0, KGB_KO_POSITIONAL = 1, KGB_KO_SITUATIONAL = 2, KGB_KO_SPIGHT = 3 };
// BoardHistory::phaseHasSpightlikeEndingAndPassHistoryClearing (:856-860) in the main phase
__device__ __forceinline__ bool koRuleSpightlike(int koRule) { return koRule == KGB_KO_SIMPLE || koRule == KGB_KO_SPIGHT; }

struct HistLists {
  // game part (read-only while searching)
  const unsigned long long* gKo; int gKoLen;        // ko hashes since the game's last history-clearing pass
  const unsigned long long* gPassB; int gPassBLen;  // ko hashes of the situations in which black / white passed
  const unsigned long long* gPassW; int gPassWLen;
  // path part (appended while descending; for the game itself these ARE the game lists and the game part is empty)
  unsigned long long* pKo; int pKoLen;
  unsigned long long* pPassB; int pPassBLen;
  unsigned long long* pPassW; int pPassWLen;
  int gKoStart;                                     // game entries before this index are hidden (cleared by a pass on the path)
};

struct HistState {           // per position, warp-uniform except the two masks
  int passes;                // consecutiveEndingPasses
  bool finished, noResult;
  uint32_t everOcc;          // this lane's row of wasEverOccupiedOrPlayed
  uint32_t banned;           // this lane's row of superKoBanned for the player to move
};

__device__ __forceinline__ unsigned long long koHashOf(int koRule, unsigned long long posH0, bool plaBlack) {
  if(koRule == KGB_KO_POSITIONAL || koRule == KGB_KO_SPIGHT) return posH0;
  return posH0 ^ (plaBlack ? 0x6A09E667F3BCC908ULL : 0xBB67AE8584CAA73BULL);
}
// warp-parallel membership / count over one list
__device__ __forceinline__ int listCount(const unsigned long long* a, int from, int n, unsigned long long h) {
  int c = 0;
  for(int i = from + kgbLane(); i < n; i += 32) c += a[i] == h ? 1 : 0;
  return __reduce_add_sync(KGB_FULL, c);
}
__device__ __forceinline__ int koCount(const HistLists& L, unsigned long long h) {
  return listCount(L.gKo, L.gKoStart, L.gKoLen, h) + listCount(L.pKo, 0, L.pKoLen, h);
}
__device__ __forceinline__ bool passSeen(const HistLists& L, bool black, unsigned long long h) {
  if(black) return listCount(L.gPassB, 0, L.gPassBLen, h) + listCount(L.pPassB, 0, L.pPassBLen, h) > 0;
  return listCount(L.gPassW, 0, L.gPassWLen, h) + listCount(L.pPassW, 0, L.pPassWLen, h) > 0;
}

// superKoBanned for the player to move (boardhistory.cpp:1063-1084): an empty point is banned when playing there recreates a
// situation of the ko-hash history; points that never held a stone and are not suicide cannot, illegal moves are not marked.
__device__ uint32_t histSuperKoBanned(const WarpBoard& bd, const HistLists& L, uint32_t everOcc, bool nextBlack, int koRule, bool multiSuicide,
                                      const ZobEntry* zob) {
  const uint32_t rm = bd.rowMask;
  const uint32_t empty = ~(bd.b | bd.w) & rm;
  uint32_t l1, l2, l3;
  boardLibertyClasses(bd, l1, l2, l3);
  const uint32_t own = nextBlack ? bd.b : bd.w, opp = nextBlack ? bd.w : bd.b;
  const uint32_t legal = boardLegalMask(bd, nextBlack, multiSuicide, l1);          // empty, not the ko point, not an illegal suicide
  const uint32_t capturing = nbrs(opp & l1, rm);
  const uint32_t suicide = empty & ~(nbrs(empty, rm) | nbrs(own & ~l1, rm) | capturing);   // Board::isSuicide
  const uint32_t cand = legal & (everOcc | suicide);
  uint32_t banned = 0;
  // plain candidates: the position after the move is this one plus the stone
  uint32_t plain = cand & ~capturing & ~suicide;
  const int y = kgbLane();
  while(plain) {   // lanes work on their own rows; every lane scans the lists itself (short lists; L1/L2 resident)
    const int x = __ffs(plain) - 1;
    plain &= plain - 1;
    const unsigned long long h = koHashOf(koRule, bd.h0 ^ zob[(y * 32 + x) * 2 + (nextBlack ? 0 : 1)].h0, !nextBlack);
    bool hit = false;
    for(int i = L.gKoStart; i < L.gKoLen && !hit; i++) hit = L.gKo[i] == h;
    for(int i = 0; i < L.pKoLen && !hit; i++) hit = L.pKo[i] == h;
    if(hit) banned |= 1u << x;
  }
  // capturing or suicidal candidates: play the move on a copy (Board::getPosHashAfterMove, board.cpp:970-1049)
  uint32_t special = cand & (capturing | suicide);
  while(true) {
    const int p = firstPoint(special);
    if(p < 0) break;
    special &= ~pointMask(p);
    WarpBoard c = bd;
    boardPlay(c, p, nextBlack, zob);
    if(koCount(L, koHashOf(koRule, c.h0, !nextBlack)) > 0) banned |= pointMask(p);
  }
  return banned;
}

// BoardHistory::makeBoardMoveAssumeLegal for the rule subset.  p = y*32+x or < 0 for a pass; bd carries the Zobrist hash.
// computeBans = false leaves st.banned untouched (a search only needs the bans where it creates or evaluates a position).
__device__ void histMakeMove(WarpBoard& bd, HistState& st, HistLists& L, int p, bool black, int koRule, bool multiSuicide, const ZobEntry* zob,
                             bool computeBans = true) {
  const int lane = kgbLane();
  bool spight = false;
  if(p >= 0) st.passes = 0;
  else {
    if(koRuleSpightlike(koRule)) { L.gKoStart = L.gKoLen; L.pKoLen = 0; }   // passes clear the ko-hash history under simple / spight ko
    const unsigned long long hb = koHashOf(koRule, bd.h0, black);
    st.passes = koRule == KGB_KO_SPIGHT ? 0 : st.passes + 1;              // newConsecutiveEndingPassesAfterPass (:831-851)
    spight = koRuleSpightlike(koRule) && passSeen(L, black, hb);          // checked BEFORE this pass is recorded
    __syncwarp();
    if(lane == 0) { if(black) L.pPassB[L.pPassBLen] = hb; else L.pPassW[L.pPassWLen] = hb; }
    if(black) L.pPassBLen++; else L.pPassWLen++;
  }
  boardPlay(bd, p, black, zob);
  const unsigned long long ha = koHashOf(koRule, bd.h0, !black);
  __syncwarp();
  if(lane == 0) L.pKo[L.pKoLen] = ha;
  L.pKoLen++;
  __syncwarp();
  if(p >= 0) st.everOcc |= pointMask(p);
  if(computeBans) st.banned = koRule != KGB_KO_SIMPLE ? histSuperKoBanned(bd, L, st.everOcc, !black, koRule, multiSuicide, zob) : 0u;
  st.finished = st.passes >= 2 || spight;
  st.noResult = false;
  if(p >= 0 && koRule == KGB_KO_SIMPLE && koCount(L, ha) >= 3) { st.noResult = true; st.finished = true; }
}

// BoardHistory::passWouldEndPhase (:874-880) for the player to move
__device__ __forceinline__ bool histPassWouldEndPhase(const WarpBoard& bd, const HistState& st, const HistLists& L, bool black, int koRule) {
  if(koRule != KGB_KO_SPIGHT && st.passes + 1 >= 2) return true;
  return koRuleSpightlike(koRule) && passSeen(L, black, koHashOf(koRule, bd.h0, black));
}

// Order-independent hash of a set of points (the superko bans enter the state / cache keys, boardhistory.cpp:1238-1244).
__device__ __forceinline__ unsigned long long pointSetHash(uint32_t rowBits) {
  unsigned long long h = 0;
  const int y = kgbLane();
  while(rowBits) {
    const int x = __ffs(rowBits) - 1;
    rowBits &= rowBits - 1;
    unsigned long long z = (unsigned long long)(y * 32 + x) + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; h ^= z ^ (z >> 31);
  }
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) h ^= __shfl_xor_sync(KGB_FULL, h, o);
  return h;
}

}  // namespace kgb
