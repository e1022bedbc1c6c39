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
static constexpr int LADDER_NODE_BUDGET = 25000;

// Per-warp global scratch.  The search stack holds, per level, the position before that level's move, the hunted chain and
// the moves not tried yet - all as lane-distributed masks (one coalesced 128-byte line each), written when the search goes
// down and read back in one round trip when it comes back up.
struct LadderScratch {
  uint32_t *sB, *sW;     // [LADDER_MAX_LEVELS][32]
  uint32_t* sC;          // [LADDER_MAX_LEVELS][32] the hunted chain at that level
  uint32_t* sR;          // [LADDER_MAX_LEVELS][32] moves of that level still to try (tried in bitboard order)
  int* sKo;              // [LADDER_MAX_LEVELS]
  // suspended-work state (a position's searches can be spread over several kernel launches, see boardLaddersResumable)
  int* st;               // [LADDER_ST_INTS]
  uint32_t *stB, *stW;   // [32] board of the suspended search
  uint32_t *accLad, *accWB, *accWW;   // [32] results of this part's finished searches
  unsigned long long* counters;   // [2] searches, nodes (statistics; may be null)
};
static constexpr int LADDER_ST_INTS = 16;
enum { LST_NEXT_ITEM = 0, LST_ACTIVE, LST_LEVEL, LST_NODES, LST_RET, LST_FRESH, LST_KO, LST_PLA_BLACK };
__host__ __device__ inline size_t ladderScratchWordsPerWarp() { return (size_t)LADDER_MAX_LEVELS * (128 + 1) + LADDER_ST_INTS + 5 * 32; }
__device__ __forceinline__ LadderScratch ladderScratchAt(uint32_t* base) {
  LadderScratch s;
  const size_t L32 = (size_t)LADDER_MAX_LEVELS * 32;
  s.sB = base; s.sW = base + L32; s.sC = base + 2 * L32; s.sR = base + 3 * L32;
  s.sKo = reinterpret_cast<int*>(base + 4 * L32);
  s.st = s.sKo + LADDER_MAX_LEVELS;
  uint32_t* ub = reinterpret_cast<uint32_t*>(s.st + LADDER_ST_INTS);
  s.stB = ub; s.stW = ub + 32; s.accLad = ub + 64; s.accWB = ub + 96; s.accWW = ub + 128;
  s.counters = nullptr;
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

// boardPlay for the search: no capture counters, no hash; `known` is a complete chain of colour knownBlack on bd (the hunted
// chain, which the search tracks anyway) so that the two flood fills that would walk its whole length start from all of it.
__device__ __forceinline__ void ladderPlay(WarpBoard& bd, int p, bool black, uint32_t known, bool knownBlack) {
  const uint32_t rm = bd.rowMask;
  const uint32_t stone = pointMask(p);
  uint32_t own = (black ? bd.b : bd.w) | stone;
  uint32_t opp = black ? bd.w : bd.b;
  uint32_t adjOpp = nbrs(stone, rm) & opp;
  int captured = 0, possibleKo = -1;
  while(true) {
    const int q = firstPoint(adjOpp);
    if(q < 0) break;
    const bool isKnown = knownBlack != black && __any_sync(KGB_FULL, (pointMask(q) & known) != 0);
    const uint32_t chain = isKnown ? known : flood(pointMask(q), opp, rm);
    const uint32_t empty = ~(own | opp) & rm;
    if(!__any_sync(KGB_FULL, (nbrs(chain, rm) & empty) != 0)) {
      captured += warpCount(chain);
      opp &= ~chain;
      possibleKo = q;
    }
    adjOpp &= ~chain;
  }
  uint32_t seed = stone;
  if(knownBlack == black && __any_sync(KGB_FULL, (nbrs(stone, rm) & known) != 0)) seed |= known;
  const uint32_t mine = flood(seed, own, rm);
  const uint32_t empty = ~(own | opp) & rm;
  const int myLibs = warpCount(nbrs(mine, rm) & empty);
  bd.ko = -1;
  if(captured == 1 && myLibs == 1 && warpCount(mine) == 1) bd.ko = possibleKo;
  if(myLibs == 0) own &= ~mine;   // suicide
  if(black) { bd.b = own; bd.w = opp; } else { bd.w = own; bd.b = opp; }
}

// Board::searchIsLadderCaptured(loc, defenderFirst, buf).  `bd` is a private copy (passed by value).
// Returns 0 / 1, or 2 = suspended: `budget` (moves this call may still play) ran out; the whole search state (board, level,
// counters; the stack is in `sc` anyway) is saved in `sc` and the same call with resume = true carries on where it stopped.
__device__ int ladderSearchImpl(WarpBoard bd, int loc, bool defenderFirst, const LadderScratch& sc, int X, int Y, int& nodes, int& budget,
                                bool resume) {
  const int lane = kgbLane();
  const uint32_t rm = bd.rowMask;
  nodes = 0;
  bool plaBlack;
  int level = 0;
  bool ret = false;
  bool fresh = true;        // this level was just entered (its move list is not built yet); false = came back up from level + 1
  uint32_t chain = 0;       // the hunted chain at the current level (also kept per level in sc.sC)
  bool chainHint = false;   // `chain` is the parent level's chain and the move from there was just played
  if(resume) {
    bd.b = sc.stB[lane]; bd.w = sc.stW[lane]; bd.ko = sc.st[LST_KO];
    level = sc.st[LST_LEVEL]; nodes = sc.st[LST_NODES]; ret = sc.st[LST_RET] != 0; fresh = sc.st[LST_FRESH] != 0;
    plaBlack = sc.st[LST_PLA_BLACK] != 0;
  }
  else {
    plaBlack = __any_sync(KGB_FULL, (pointMask(loc) & bd.b) != 0);
    chain = flood(pointMask(loc), plaBlack ? bd.b : bd.w, rm);
    const int libs0 = chainLibCount(bd, chain);
    if(libs0 > 2 || (defenderFirst && libs0 > 1)) return 0;
    if(defenderFirst) bd.ko = -1;
    chainHint = true;
  }
  const int stackSize = X * Y * 3 / 2 + 1;
  while(true) {
    if(level < 0) return ret ? 1 : 0;
    if(level >= stackSize - 1) { ret = true; fresh = false; level--; continue; }
    if(nodes >= LADDER_NODE_BUDGET) return 0;
    if(budget <= 0) {
      sc.stB[lane] = bd.b; sc.stW[lane] = bd.w;
      if(lane == 0) {
        sc.st[LST_KO] = bd.ko; sc.st[LST_LEVEL] = level; sc.st[LST_NODES] = nodes; sc.st[LST_RET] = ret ? 1 : 0;
        sc.st[LST_FRESH] = fresh ? 1 : 0; sc.st[LST_PLA_BLACK] = plaBlack ? 1 : 0;
      }
      __syncwarp();
      return 2;
    }
    const bool isDef = (defenderFirst && (level % 2) == 0) || (!defenderFirst && (level % 2) == 1);
    uint32_t rem;   // moves of this level not tried yet
    if(fresh) {
      const uint32_t own = plaBlack ? bd.b : bd.w, opp = plaBlack ? bd.w : bd.b;
      // the chain only ever grows along a line of play: start the fill from the parent's chain
      chain = flood(chainHint ? (chain | pointMask(loc)) : pointMask(loc), own, rm);
      chainHint = false;
      const uint32_t empty = ~(bd.b | bd.w) & rm;
      const uint32_t L = nbrs(chain, rm) & empty;
      const int libs = warpCount(L);
      if((!isDef && libs <= 1) || (!isDef && libs >= 3) || (isDef && libs >= 2) || (isDef && bd.ko >= 0)) {
        // attacker to move: 1 liberty = captured, 3 = escaped; defender to move: 2 liberties or a ko left by the attacker = escaped
        ret = !isDef && libs <= 1;
        fresh = false; level--; continue;
      }
      if(isDef) {
        // capture moves: the liberty of every adjacent opponent chain in atari, then the chain's own liberty.
        // Opponent chains holding a stone with two empty neighbours are ruled out for the whole board in one multi-seed fill.
        uint32_t M = 0;
        uint32_t adj = nbrs(chain, rm) & opp;
        {
          const uint32_t ea = empty << 1, eb = empty >> 1, ec = rowAbove(empty), ed = rowBelow(empty);
          const uint32_t atLeast2 = (ea & eb) | (ec & ed) | ((ea | eb) & (ec | ed));
          if(__any_sync(KGB_FULL, adj != 0)) adj &= ~flood(opp & atLeast2, opp, rm);
        }
        while(true) {
          int q = firstPoint(adj);
          if(q < 0) break;
          uint32_t c = flood(pointMask(q), opp, rm);
          uint32_t cl = nbrs(c, rm) & empty;
          if(warpCount(cl) == 1) M |= cl;
          adj &= ~c;
        }
        rem = M | L;
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
            if(!gaining) { ret = true; fresh = false; level--; continue; }
          }
        }
        rem = L;
        const bool adjacent = __any_sync(KGB_FULL, (nbrs(pointMask(l0), rm) & pointMask(l1)) != 0);
        if(!adjacent) {
          if(imm0 >= 3 && imm1 >= 3) { ret = false; fresh = false; level--; continue; }
          else if(imm0 >= 3) rem = pointMask(l0);
          else if(imm1 >= 3) rem = pointMask(l1);
        }
      }
    }
    else {
      // back from level + 1 with its verdict in `ret`: a refutation found ends this level
      if((isDef && !ret) || (!isDef && ret)) { level--; continue; }
      bd.b = sc.sB[level * 32 + lane]; bd.w = sc.sW[level * 32 + lane]; bd.ko = sc.sKo[level];
      chain = sc.sC[level * 32 + lane]; rem = sc.sR[level * 32 + lane];
    }
    // next legal move of this level (illegal ones - ko, suicide - are skipped)
    const bool moverBlack = isDef ? plaBlack : !plaBlack;
    int move = -1;
    while(true) {
      const int q = firstPoint(rem);
      if(q < 0) break;
      rem &= ~pointMask(q);
      if(pointIsLegalNoSuicide(bd, q, moverBlack)) { move = q; break; }
    }
    if(move < 0) { ret = isDef; fresh = false; level--; continue; }   // out of moves: the side to move here has failed
    sc.sB[level * 32 + lane] = bd.b; sc.sW[level * 32 + lane] = bd.w; sc.sC[level * 32 + lane] = chain; sc.sR[level * 32 + lane] = rem;
    if(lane == 0) sc.sKo[level] = bd.ko;
    __syncwarp();   // sKo is written by lane 0 and read by every lane
    ladderPlay(bd, move, moverBlack, chain, plaBlack);
    chainHint = true;
    nodes++;
    budget--;
    level++;
    fresh = true;
  }
}

__device__ __forceinline__ int ladderSearch(const WarpBoard& bd, int loc, bool defenderFirst, const LadderScratch& sc, int X, int Y, int& budget,
                                            bool resume) {
  int nodes;
  const int r = ladderSearchImpl(bd, loc, defenderFirst, sc, X, Y, nodes, budget, resume);
  if(r != 2 && sc.counters != nullptr && kgbLane() == 0) { atomicAdd(sc.counters, 1ULL); atomicAdd(sc.counters + 1, (unsigned long long)nodes); }
  return r;
}

// iterLadders (nninputs.cpp:815-866): laddered = stones of chains with 1 or 2 liberties that are ladder-capturable;
// working = for 2-liberty chains the attacker's first moves that work.
//   * `candidates` = the stones of chains with exactly 1 or 2 liberties (boardLibertyClasses' lib1 | lib2): only those chains
//     are walked.
//   * The searches of one position are independent: `part` of `nparts` takes every nparts-th search (in enumeration order);
//     the caller ORs the parts' results (sc.accLad / accWB / accWW) together.
//   * `budget` bounds the moves played in this call.  When it runs out the work is suspended in `sc` and false is returned;
//     calling again with fresh = false (same board, same part) continues.  The results do not depend on how the work was cut.
__device__ bool boardLaddersResumable(const WarpBoard& bd, uint32_t candidates, const LadderScratch& sc, int X, int Y, int part, int nparts,
                                      int& budget, bool fresh) {
  const int lane = kgbLane();
  const uint32_t rm = bd.rowMask;
  const uint32_t empty = ~(bd.b | bd.w) & rm;
  int nextItem = 0;
  bool active = false;
  uint32_t laddered = 0, workB = 0, workW = 0;
  if(!fresh) {
    nextItem = sc.st[LST_NEXT_ITEM]; active = sc.st[LST_ACTIVE] != 0;
    laddered = sc.accLad[lane]; workB = sc.accWB[lane]; workW = sc.accWW[lane];
  }
  int item = 0;
  bool suspended = false;
  uint32_t todoB = bd.b & candidates, todoW = bd.w & candidates;
  while(!suspended) {
    int q = firstPoint(todoB);
    const bool isB = q >= 0;
    if(!isB) q = firstPoint(todoW);
    if(q < 0) break;
    const uint32_t chain = flood(pointMask(q), isB ? bd.b : bd.w, rm);
    if(isB) todoB &= ~chain; else todoW &= ~chain;
    const uint32_t L = nbrs(chain, rm) & empty;
    const int libs = warpCount(L);
    if(libs == 1) {
      const int idx = item++;
      if((idx % nparts) != part || idx < nextItem) continue;
      const int r = ladderSearch(bd, q, true, sc, X, Y, budget, active && idx == nextItem);
      active = false;
      if(r == 2) { nextItem = idx; active = true; suspended = true; }
      else if(r == 1) laddered |= chain;
    }
    else if(libs == 2) {
      // searchIsLadderCapturedAttackerFirst2Libs (game/board.cpp:1581-1626)
      const int m0 = firstPoint(L), m1 = firstPoint(L & ~pointMask(m0));
      for(int k = 0; k < 2 && !suspended; k++) {
        const int m = k == 0 ? m0 : m1;
        const int idx = item++;
        if((idx % nparts) != part || idx < nextItem) continue;
        const bool resume = active && idx == nextItem;
        active = false;
        if(resume || pointIsLegalNoSuicide(bd, m, !isB)) {
          WarpBoard c = bd;
          if(!resume) boardPlay(c, m, !isB);
          const int r = ladderSearch(c, q, true, sc, X, Y, budget, resume);
          if(r == 2) { nextItem = idx; active = true; suspended = true; }
          else if(r == 1) {
            laddered |= chain;
            if(isB) workB |= pointMask(m); else workW |= pointMask(m);
          }
        }
      }
    }
  }
  sc.accLad[lane] = laddered; sc.accWB[lane] = workB; sc.accWW[lane] = workW;
  if(lane == 0) { sc.st[LST_NEXT_ITEM] = suspended ? nextItem : 0x7fffffff; sc.st[LST_ACTIVE] = (suspended && active) ? 1 : 0; }
  __syncwarp();
  return !suspended;
}

// One-shot form (no budget, no partition).
__device__ void boardLadders(const WarpBoard& bd, const LadderScratch& sc, int X, int Y, uint32_t& laddered, uint32_t& workingOfBlackChains,
                             uint32_t& workingOfWhiteChains) {
  uint32_t l1, l2, l3;
  boardLibertyClasses(bd, l1, l2, l3);
  int budget = 0x7fffffff;
  boardLaddersResumable(bd, l1 | l2, sc, X, Y, 0, 1, budget, true);
  const int lane = kgbLane();
  laddered = sc.accLad[lane]; workingOfBlackChains = sc.accWB[lane]; workingOfWhiteChains = sc.accWW[lane];
}

}  // namespace kgb
