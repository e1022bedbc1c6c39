// Ladder reader on warp bitboards: Board::searchIsLadderCaptured / searchIsLadderCapturedAttackerFirst2Libs
// (game/board.cpp:1581-1850) and NNInputs' iterLadders (neuralnet/nninputs.cpp:815-866) - SURVEY.md §8a row a5.
//
// Same search as the reference: iterative alternating DFS with an explicit stack, defender = the chain's owner,
//   attacker node : chain must have exactly 2 liberties here (1 -> attacker wins, >= 3 -> loses); tries the liberties,
//                   with the reference's prunings (non-adjacent liberties with >= 3 empty neighbours) and its double-ko
//                   death heuristic (:1745-1757)
//   defender node : chain has 1 liberty (>= 2 -> escaped); a simple-ko point left by the attacker counts as escaped
//                   (:1703-1705); tries captures of adjacent chains in atari, then the liberty
//   illegal moves (ko, suicide with multi-stone suicide treated as illegal) are skipped; stack limit X*Y*3/2+1 -> "captured";
//   node budget 25 000 -> "not captured".
// Differences, by construction of the bitboard form (documented in DESIGN.md):
//   * moves are tried in bitboard order, not linked-list order: the minimax result is order-independent;
//   * the two "bound on liberties after play" shortcuts (:1719-1727) are sound prunings in the normal case and are simply
//     not taken (the full search reaches the same answer); in the corner where the defender's only liberty is also a
//     capturing move and two or more captures exist, the reference applies the bound to a list-order-dependent move.
// Undo = restore the saved bitboards of the level (the reference keeps MoveRecords).
#pragma once
#include "kgb_board.cuh"

namespace kgb {

static constexpr int LADDER_MAX_LEVELS = 19 * 19 * 3 / 2 + 3;
static constexpr int LADDER_MOVEBUF = 4096;
static constexpr int LADDER_NODE_BUDGET = 25000;

struct LadderScratch {   // per-warp global scratch
  uint32_t* sB;          // [LADDER_MAX_LEVELS][32]
  uint32_t* sW;
  int* sKo;              // [LADDER_MAX_LEVELS]
  int* listStart;        // [LADDER_MAX_LEVELS]
  int* listLen;
  int* listCur;
  int* moves;            // [LADDER_MOVEBUF]
};
__host__ __device__ inline size_t ladderScratchWordsPerWarp() { return (size_t)LADDER_MAX_LEVELS * (64 + 4) + LADDER_MOVEBUF; }
__device__ __forceinline__ LadderScratch ladderScratchAt(uint32_t* base) {
  LadderScratch s;
  s.sB = base; s.sW = base + (size_t)LADDER_MAX_LEVELS * 32;
  int* ib = reinterpret_cast<int*>(base + (size_t)LADDER_MAX_LEVELS * 64);
  s.sKo = ib; s.listStart = ib + LADDER_MAX_LEVELS; s.listLen = ib + 2 * LADDER_MAX_LEVELS; s.listCur = ib + 3 * LADDER_MAX_LEVELS;
  s.moves = ib + 4 * LADDER_MAX_LEVELS;
  return s;
}

__device__ __forceinline__ int pointCountEmptyNbrs(const WarpBoard& bd, int p) {   // Board::getNumImmediateLiberties
  const uint32_t empty = ~(bd.b | bd.w) & bd.rowMask;
  return warpCount(nbrs(pointMask(p), bd.rowMask) & empty);
}
__device__ __forceinline__ int chainLibCount(const WarpBoard& bd, uint32_t chain) {
  const uint32_t empty = ~(bd.b | bd.w) & bd.rowMask;
  return warpCount(nbrs(chain, bd.rowMask) & empty);
}
// Board::isLegal(loc, pla, isMultiStoneSuicideLegal = false) for a single point
__device__ __forceinline__ bool pointIsLegalNoSuicide(const WarpBoard& bd, int p, bool black) {
  const uint32_t rm = bd.rowMask, pt = pointMask(p);
  const uint32_t empty = ~(bd.b | bd.w) & rm;
  if(!__any_sync(KGB_FULL, (pt & empty) != 0) || p == bd.ko) return false;
  const uint32_t nb = nbrs(pt, rm);
  if(__any_sync(KGB_FULL, (nb & empty) != 0)) return true;
  const uint32_t own = black ? bd.b : bd.w, opp = black ? bd.w : bd.b;
  uint32_t adj = nb & own;
  while(true) {
    int q = firstPoint(adj);
    if(q < 0) break;
    uint32_t c = flood(pointMask(q), own, rm);
    if(chainLibCount(bd, c) > 1) return true;
    adj &= ~c;
  }
  adj = nb & opp;
  while(true) {
    int q = firstPoint(adj);
    if(q < 0) break;
    uint32_t c = flood(pointMask(q), opp, rm);
    if(chainLibCount(bd, c) == 1) return true;
    adj &= ~c;
  }
  return false;
}
// Board::wouldBeKoCapture(loc, pla): loc empty, every on-board neighbour is an opponent stone, exactly one neighbouring
// opponent chain is in atari and it is a single stone (game/board.cpp:518-542)
__device__ __forceinline__ bool pointWouldBeKoCapture(const WarpBoard& bd, int p, bool plaBlack) {
  const uint32_t rm = bd.rowMask, pt = pointMask(p);
  const uint32_t empty = ~(bd.b | bd.w) & rm;
  if(!__any_sync(KGB_FULL, (pt & empty) != 0)) return false;
  const uint32_t opp = plaBlack ? bd.w : bd.b;
  const uint32_t nb = nbrs(pt, rm);
  if(__any_sync(KGB_FULL, (nb & ~opp) != 0)) return false;
  int capturable = 0, capSize = 0;
  uint32_t adj = nb;   // all opponent stones
  while(true) {
    int q = firstPoint(adj);
    if(q < 0) break;
    // the reference tests each adjacent STONE (not chain): two neighbours of the same atari chain count twice -> not a ko
    uint32_t c = flood(pointMask(q), opp, rm);
    if(chainLibCount(bd, c) == 1) { capturable += 1; capSize = warpCount(c); }
    adj &= ~pointMask(q);
  }
  return capturable == 1 && capSize == 1;
}
// Board::getNumLibertiesAfterPlay(loc, pla, max) (game/board.cpp:344-430): liberties of the chain a stone at p would belong to
__device__ __forceinline__ int pointLibsAfterPlay(const WarpBoard& bd, int p, bool plaBlack, int maxv) {
  const uint32_t rm = bd.rowMask, pt = pointMask(p);
  uint32_t own = (plaBlack ? bd.b : bd.w) | pt;
  uint32_t opp = plaBlack ? bd.w : bd.b;
  const uint32_t emptyBefore = ~(bd.b | bd.w) & rm;
  uint32_t adj = nbrs(pt, rm) & opp, captured = 0;
  while(true) {
    int q = firstPoint(adj);
    if(q < 0) break;
    uint32_t c = flood(pointMask(q), opp, rm);
    if(warpCount(nbrs(c, rm) & emptyBefore) == 1) captured |= c;
    adj &= ~c;
  }
  opp &= ~captured;
  const uint32_t chain = flood(pt, own, rm);
  const uint32_t emptyAfter = ~(own | opp) & rm;
  int n = warpCount(nbrs(chain, rm) & emptyAfter);
  return n < maxv ? n : maxv;
}

// Board::searchIsLadderCaptured(loc, defenderFirst, buf).  `bd` is a private copy (passed by value).
__device__ bool ladderSearch(WarpBoard bd, int loc, bool defenderFirst, const LadderScratch& sc, int X, int Y) {
  const int lane = kgbLane();
  const uint32_t rm = bd.rowMask;
  const bool plaBlack = __any_sync(KGB_FULL, (pointMask(loc) & bd.b) != 0);
  {
    uint32_t chain0 = flood(pointMask(loc), plaBlack ? bd.b : bd.w, rm);
    int libs0 = chainLibCount(bd, chain0);
    if(libs0 > 2 || (defenderFirst && libs0 > 1)) return false;
  }
  if(defenderFirst) bd.ko = -1;
  const int stackSize = X * Y * 3 / 2 + 1;
  int level = 0, nodes = 0;
  bool ret = false, fromDeeper = false;
  if(lane == 0) { sc.listCur[0] = -1; sc.listStart[0] = 0; sc.listLen[0] = 0; }
  __syncwarp();
  while(true) {
    if(level < 0) return ret;
    if(level >= stackSize - 1) { ret = true; fromDeeper = true; level--; continue; }
    if(nodes >= LADDER_NODE_BUDGET) return false;
    const bool isDef = (defenderFirst && (level % 2) == 0) || (!defenderFirst && (level % 2) == 1);
    int cur = sc.listCur[level];
    if(cur == -1) {
      const uint32_t own = plaBlack ? bd.b : bd.w, opp = plaBlack ? bd.w : bd.b;
      const uint32_t chain = flood(pointMask(loc), own, rm);
      const uint32_t empty = ~(bd.b | bd.w) & rm;
      const uint32_t L = nbrs(chain, rm) & empty;
      const int libs = warpCount(L);
      if(!isDef && libs <= 1) { ret = true; fromDeeper = true; level--; continue; }
      if(!isDef && libs >= 3) { ret = false; fromDeeper = true; level--; continue; }
      if(isDef && libs >= 2) { ret = false; fromDeeper = true; level--; continue; }
      if(isDef && bd.ko >= 0) { ret = false; fromDeeper = true; level--; continue; }
      const int start = sc.listStart[level];
      int len = 0;
      if(isDef) {
        // capture moves: the liberty of every adjacent opponent chain in atari, then the chain's own liberty
        uint32_t M = 0;
        uint32_t adj = nbrs(chain, rm) & opp;
        while(true) {
          int q = firstPoint(adj);
          if(q < 0) break;
          uint32_t c = flood(pointMask(q), opp, rm);
          uint32_t cl = nbrs(c, rm) & empty;
          if(warpCount(cl) == 1) M |= cl;
          adj &= ~c;
        }
        uint32_t all = M | L;
        while(true) {
          int q = firstPoint(all);
          if(q < 0) break;
          if(lane == 0 && start + len < LADDER_MOVEBUF) sc.moves[start + len] = q;
          len++;
          all &= ~pointMask(q);
        }
      }
      else {
        const int l0 = firstPoint(L);
        const int l1 = firstPoint(L & ~pointMask(l0));
        int imm0 = pointCountEmptyNbrs(bd, l0), imm1 = pointCountEmptyNbrs(bd, l1);
        // double-ko death heuristic (game/board.cpp:1745-1757)
        if(imm0 == 0 && imm1 == 0 && pointWouldBeKoCapture(bd, l0, !plaBlack) && pointWouldBeKoCapture(bd, l1, !plaBlack)) {
          if(pointLibsAfterPlay(bd, l0, plaBlack, 3) <= 2 && pointLibsAfterPlay(bd, l1, plaBlack, 3) <= 2) {
            bool gaining = false;
            uint32_t adj = nbrs(chain, rm) & opp;
            while(!gaining) {
              int q = firstPoint(adj);
              if(q < 0) break;
              uint32_t c = flood(pointMask(q), opp, rm);
              if(chainLibCount(bd, c) == 1) gaining = true;
              adj &= ~c;
            }
            if(!gaining) { ret = true; fromDeeper = true; level--; continue; }
          }
        }
        int m0 = l0, m1 = l1;
        len = 2;
        const bool adjacent = __any_sync(KGB_FULL, (nbrs(pointMask(l0), rm) & pointMask(l1)) != 0);
        if(!adjacent) {
          if(imm0 >= 3 && imm1 >= 3) { ret = false; fromDeeper = true; level--; continue; }
          else if(imm0 >= 3) len = 1;
          else if(imm1 >= 3) { m0 = l1; len = 1; }
        }
        if(lane == 0 && start + 1 < LADDER_MOVEBUF) { sc.moves[start] = m0; sc.moves[start + 1] = m1; }
      }
      if(start + len >= LADDER_MOVEBUF) return false;   // scratch exhausted: same as the node budget
      if(lane == 0) { sc.listLen[level] = len; sc.listCur[level] = 0; }
      __syncwarp();
      cur = 0;
    }
    else {
      if(fromDeeper) {   // undo: restore the position saved before this level's move
        bd.b = sc.sB[level * 32 + lane]; bd.w = sc.sW[level * 32 + lane]; bd.ko = sc.sKo[level];
      }
      if(isDef && !ret) { fromDeeper = true; level--; continue; }
      if(!isDef && ret) { fromDeeper = true; level--; continue; }
      cur += 1;
      if(lane == 0) sc.listCur[level] = cur;
      __syncwarp();
    }
    const int len = sc.listLen[level];
    if(cur >= len) { ret = isDef; fromDeeper = true; level--; continue; }
    const int move = sc.moves[sc.listStart[level] + cur];
    const bool moverBlack = isDef ? plaBlack : !plaBlack;
    if(!pointIsLegalNoSuicide(bd, move, moverBlack)) { ret = isDef; fromDeeper = false; continue; }
    sc.sB[level * 32 + lane] = bd.b; sc.sW[level * 32 + lane] = bd.w;
    if(lane == 0) sc.sKo[level] = bd.ko;
    boardPlay(bd, move, moverBlack);
    nodes++;
    level++;
    if(lane == 0) { sc.listCur[level] = -1; sc.listStart[level] = sc.listStart[level - 1] + sc.listLen[level - 1]; sc.listLen[level] = 0; }
    __syncwarp();
  }
}

// iterLadders (nninputs.cpp:815-866): laddered = stones of chains with 1 or 2 liberties that are ladder-capturable;
// working = for 2-liberty chains the attacker's first moves that work (only filled when wantWorking).
__device__ void boardLadders(const WarpBoard& bd, const LadderScratch& sc, int X, int Y, uint32_t& laddered, uint32_t& workingOfBlackChains,
                             uint32_t& workingOfWhiteChains) {
  const uint32_t rm = bd.rowMask;
  const uint32_t empty = ~(bd.b | bd.w) & rm;
  laddered = 0; workingOfBlackChains = 0; workingOfWhiteChains = 0;
  uint32_t todoB = bd.b, todoW = bd.w;
  while(true) {
    int q = firstPoint(todoB);
    const bool isB = q >= 0;
    if(!isB) q = firstPoint(todoW);
    if(q < 0) break;
    const uint32_t chain = flood(pointMask(q), isB ? bd.b : bd.w, rm);
    if(isB) todoB &= ~chain; else todoW &= ~chain;
    const uint32_t L = nbrs(chain, rm) & empty;
    const int libs = warpCount(L);
    if(libs == 1) {
      if(ladderSearch(bd, q, true, sc, X, Y)) laddered |= chain;
    }
    else if(libs == 2) {
      // searchIsLadderCapturedAttackerFirst2Libs (game/board.cpp:1581-1626)
      const int m0 = firstPoint(L), m1 = firstPoint(L & ~pointMask(m0));
      bool any = false;
      for(int k = 0; k < 2; k++) {
        const int m = k == 0 ? m0 : m1;
        if(pointIsLegalNoSuicide(bd, m, !isB)) {
          WarpBoard c = bd;
          boardPlay(c, m, !isB);
          if(ladderSearch(c, q, true, sc, X, Y)) {
            any = true;
            if(isB) workingOfBlackChains |= pointMask(m); else workingOfWhiteChains |= pointMask(m);
          }
        }
      }
      if(any) laddered |= chain;
    }
  }
}

}  // namespace kgb
