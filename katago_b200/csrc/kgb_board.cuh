// Warp-per-board Go rules on bitboards (north_star "Board::playMove/isLegal chain-and-liberty
// updates as 361-bit bitboards, one warp per board").
//
// Representation: lane y of the warp holds row y of the board as two 32-bit masks (bit x = point (x,y)): black and
// white.  Lanes >= Y hold zero.  No chain ids, no linked lists (reference: colors[]/chain_head[]/next_in_chain[],
// game/board.h:322-340): a chain is recomputed when needed by flood fill - a handful of shuffles + bit ops per
// dilation step, everything in registers.  Results are defined to be IDENTICAL to the reference's:
//   playMove      Board::playMoveAssumeLegal   game/board.cpp:1051-1143  (captures, simple-ko point, suicide removal,
//                                                                          capture counters)
//   legality      Board::isLegal / isIllegalSuicide / isKoBanned   game/board.cpp:283-304, 431-453
//   liberties     Board::getNumLiberties  (classes 1/2/3 are NN input planes 3-5, nninputs.cpp:2337-2343)
#pragma once
#include <stdint.h>

namespace kgb {

#define KGB_FULL 0xffffffffu

struct WarpBoard {
  uint32_t b, w;     // this lane's row: black / white stones
  uint32_t rowMask;  // on-board bits of this lane's row ((1<<X)-1 for lanes < Y, else 0)
  int ko;            // warp-uniform: y*32+x of the simple-ko point, or -1
  int capB, capW;    // warp-uniform: numBlackCaptures / numWhiteCaptures (stones of that colour removed)
  unsigned long long h0, h1;  // warp-uniform: Board::pos_hash (Zobrist, game/board.cpp:151-216) when a table is supplied
};

// Zobrist table laid out for the bitboard: entry (y*32+x)*2 + colour (0 black, 1 white) = {h0, h1}  (kgb_rand.h)
struct ZobEntry { unsigned long long h0, h1; };

__device__ __forceinline__ int kgbLane() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t rowAbove(uint32_t v) {  // value held by lane-1 (row y-1)
  uint32_t t = __shfl_up_sync(KGB_FULL, v, 1);
  return kgbLane() == 0 ? 0u : t;
}
__device__ __forceinline__ uint32_t rowBelow(uint32_t v) {  // value held by lane+1 (row y+1)
  uint32_t t = __shfl_down_sync(KGB_FULL, v, 1);
  return kgbLane() == 31 ? 0u : t;
}
// 4-neighbourhood of a point set (not including the set itself unless adjacent members)
__device__ __forceinline__ uint32_t nbrs(uint32_t v, uint32_t rowMask) {
  return ((v << 1) | (v >> 1) | rowAbove(v) | rowBelow(v)) & rowMask;
}
__device__ __forceinline__ int warpCount(uint32_t v) {
  return __reduce_add_sync(KGB_FULL, __popc(v));   // one REDUX instruction instead of five dependent shuffles
}
// Connected component(s) of `seed` inside `allowed`.
__device__ __forceinline__ uint32_t flood(uint32_t seed, uint32_t allowed, uint32_t rowMask) {
  uint32_t f = seed & allowed;
  while(true) {
    // saturate horizontally inside the row first (cheap), then one vertical step
    uint32_t g = f;
    g |= (g << 1) & allowed; g |= (g >> 1) & allowed;   // extra in-row steps are free; the fixpoint test below decides
    g |= (g << 1) & allowed; g |= (g >> 1) & allowed;
    uint32_t n = (g | nbrs(g, rowMask)) & allowed;
    if(!__any_sync(KGB_FULL, n != f)) return f;
    f = n;
  }
}
// XOR of the table entries of every point in `chgB` (black) / `chgW` (white) - warp-uniform result added into bd.h0/h1.
__device__ __forceinline__ void boardHashToggle(WarpBoard& bd, const ZobEntry* zob, uint32_t chgB, uint32_t chgW) {
  unsigned long long a = 0, b = 0;
  const int y = kgbLane();
  while(chgB) { int x = __ffs(chgB) - 1; chgB &= chgB - 1; ZobEntry e = zob[(y * 32 + x) * 2]; a ^= e.h0; b ^= e.h1; }
  while(chgW) { int x = __ffs(chgW) - 1; chgW &= chgW - 1; ZobEntry e = zob[(y * 32 + x) * 2 + 1]; a ^= e.h0; b ^= e.h1; }
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) { a ^= __shfl_xor_sync(KGB_FULL, a, o); b ^= __shfl_xor_sync(KGB_FULL, b, o); }
  bd.h0 ^= a; bd.h1 ^= b;
}

// First set point of a warp-distributed set (lowest row, then lowest column): returns y*32+x or -1 (warp-uniform).
__device__ __forceinline__ int firstPoint(uint32_t v) {
  uint32_t rows = __ballot_sync(KGB_FULL, v != 0);
  if(rows == 0) return -1;
  int y = __ffs(rows) - 1;
  uint32_t rv = __shfl_sync(KGB_FULL, v, y);
  return y * 32 + (__ffs(rv) - 1);
}
__device__ __forceinline__ uint32_t pointMask(int p) {  // this lane's row bits of the single point p (y*32+x)
  return (p >= 0 && (p >> 5) == kgbLane()) ? (1u << (p & 31)) : 0u;
}

__device__ __forceinline__ void boardInit(WarpBoard& bd, int X, int Y) {
  bd.b = 0; bd.w = 0;
  bd.rowMask = kgbLane() < Y ? ((X >= 32) ? 0xffffffffu : ((1u << X) - 1u)) : 0u;
  bd.ko = -1; bd.capB = 0; bd.capW = 0;
  bd.h0 = 0; bd.h1 = 0;
}

// Board::playMoveAssumeLegal (game/board.cpp:1051-1143).  p = y*32+x, or p < 0 for a pass.
__device__ __forceinline__ void boardPlay(WarpBoard& bd, int p, bool black, const ZobEntry* zob = nullptr) {
  if(p < 0) { bd.ko = -1; return; }
  const uint32_t oldB = bd.b, oldW = bd.w;
  uint32_t stone = pointMask(p);
  uint32_t own = (black ? bd.b : bd.w) | stone;
  uint32_t opp = black ? bd.w : bd.b;
  const uint32_t rm = bd.rowMask;
  // capture adjacent opponent chains left without liberties
  uint32_t adjOpp = nbrs(stone, rm) & opp;
  int captured = 0, possibleKo = -1;
  while(true) {
    int q = firstPoint(adjOpp);
    if(q < 0) break;
    uint32_t chain = flood(pointMask(q), opp, rm);
    uint32_t empty = ~(own | opp) & rm;
    bool dead = !__any_sync(KGB_FULL, (nbrs(chain, rm) & empty) != 0);
    if(dead) {
      captured += warpCount(chain);
      opp &= ~chain;
      possibleKo = q;
    }
    adjOpp &= ~chain;
  }
  // the new stone's chain
  uint32_t mine = flood(stone, own, rm);
  uint32_t empty = ~(own | opp) & rm;
  int myLibs = warpCount(nbrs(mine, rm) & empty);
  int mySize = warpCount(mine);
  bd.ko = (captured == 1 && mySize == 1 && myLibs == 1) ? possibleKo : -1;
  if(black) bd.capW += captured; else bd.capB += captured;
  if(myLibs == 0) {  // suicide
    own &= ~mine;
    if(black) bd.capB += mySize; else bd.capW += mySize;
  }
  if(black) { bd.b = own; bd.w = opp; } else { bd.w = own; bd.b = opp; }
  if(zob != nullptr) boardHashToggle(bd, zob, oldB ^ bd.b, oldW ^ bd.w);
}

// Per-chain liberty classes for every stone: lib1/lib2/lib3 = stones whose chain has exactly 1/2/3 liberties.
__device__ __forceinline__ void boardLibertyClasses(const WarpBoard& bd, uint32_t& lib1, uint32_t& lib2, uint32_t& lib3) {
  const uint32_t rm = bd.rowMask;
  const uint32_t empty = ~(bd.b | bd.w) & rm;
  lib1 = lib2 = lib3 = 0;
  // stones with >= 4 empty neighbours of their own can be skipped only chain-wise, so just walk chains
  uint32_t todoB = bd.b, todoW = bd.w;
  while(true) {
    int q = firstPoint(todoB);
    bool isB = q >= 0;
    if(!isB) q = firstPoint(todoW);
    if(q < 0) break;
    uint32_t chain = flood(pointMask(q), isB ? bd.b : bd.w, rm);
    int libs = warpCount(nbrs(chain, rm) & empty);
    if(libs == 1) lib1 |= chain; else if(libs == 2) lib2 |= chain; else if(libs == 3) lib3 |= chain;
    if(isB) todoB &= ~chain; else todoW &= ~chain;
  }
}

// Board::isLegal for every point at once (game/board.cpp:283-304, 441-453): empty, not the ko point, and not an
// illegal suicide: some neighbour is empty, or an own chain with > 1 liberty (any own chain if multi-stone suicide is
// legal), or an opponent chain in atari.
__device__ __forceinline__ uint32_t boardLegalMask(const WarpBoard& bd, bool blackToMove, bool multiStoneSuicideLegal, uint32_t lib1) {
  const uint32_t rm = bd.rowMask;
  const uint32_t empty = ~(bd.b | bd.w) & rm;
  const uint32_t own = blackToMove ? bd.b : bd.w, opp = blackToMove ? bd.w : bd.b;
  uint32_t ok = nbrs(empty, rm) | nbrs(opp & lib1, rm) | nbrs(multiStoneSuicideLegal ? own : (own & ~lib1), rm);
  return empty & ok & ~pointMask(bd.ko);
}


// ------------------------------------------------------------------------------------------------------------
// Benson pass-alive groups + territory: Board::calculateArea (game/board.cpp:1853-2228), SURVEY.md §8a row a4.
//
// Restated as a greatest fixpoint over bitboards (the reference kills chains one by one with linked lists):
//   region  = maximal connected set of non-pla points that contains an empty point
//   a region is VITAL for a pla chain when every relevant point of it (all points under multi-stone suicide, else only its
//             empty points) is adjacent to that chain                                   (board.cpp:2014-2031, 2079-2101)
//   alive   = chains with >= 2 vital regions among regions that border no dead pla chain  (:2130-2170); iterate to fixpoint.
// Then (:2172-2228): alive chains; regions bordering only alive chains with <= 1 interior point, or (safeBigTerritories)
// without opponent stones, are pla's unconditionally (overwriting the opponent's marks); (unsafeBigTerritories) opponent-
// free regions are pla's where nothing is marked yet.
// resPla / resOpp are the result planes for this player / the other one (in/out, so the caller runs black then white).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bensonForPla(const WarpBoard& bd, bool plaIsBlack, bool multiStoneSuicideLegal, bool safeBigTerritories,
                                             bool unsafeBigTerritories, uint32_t& resPla, uint32_t& resOpp) {
  const uint32_t rm = bd.rowMask;
  const uint32_t pla = plaIsBlack ? bd.b : bd.w, opp = plaIsBlack ? bd.w : bd.b;
  if(!__any_sync(KGB_FULL, pla != 0)) return;   // atLeastOnePla == false: nothing is marked for this player
  const uint32_t nonPla = ~pla & rm, empty = ~(bd.b | bd.w) & rm;
  const uint32_t nbPla = nbrs(pla, rm);
  uint32_t alive = pla;
  while(true) {
    uint32_t v1 = 0, v2 = 0;   // stones of chains with >= 1 / >= 2 vital healthy regions
    uint32_t todo = empty;
    while(true) {
      const int q = firstPoint(todo);
      if(q < 0) break;
      const uint32_t seed = pointMask(q);
      const uint32_t R = flood(seed, nonPla, rm);
      todo &= ~R;
      if(__any_sync(KGB_FULL, (nbrs(R, rm) & pla & ~alive) != 0)) continue;   // borders a dead chain
      const uint32_t E = multiStoneSuicideLegal ? R : (R & empty);
      uint32_t cand = nbrs(seed, rm) & pla;
      while(true) {
        const int c = firstPoint(cand);
        if(c < 0) break;
        const uint32_t C = flood(pointMask(c), pla, rm);
        cand &= ~C;
        if(!__any_sync(KGB_FULL, (E & ~nbrs(C, rm)) != 0)) { v2 |= v1 & C; v1 |= C; }
      }
    }
    const uint32_t next = alive & v2;
    if(!__any_sync(KGB_FULL, next != alive)) break;
    alive = next;
  }
  resPla |= alive; resOpp &= ~alive;
  uint32_t todo = empty;
  while(true) {
    const int q = firstPoint(todo);
    if(q < 0) break;
    const uint32_t R = flood(pointMask(q), nonPla, rm);
    todo &= ~R;
    const bool healthy = !__any_sync(KGB_FULL, (nbrs(R, rm) & pla & ~alive) != 0);
    const bool containsOpp = __any_sync(KGB_FULL, (R & opp) != 0);
    bool mark = false;
    if(healthy) {
      if(safeBigTerritories && !containsOpp) mark = true;
      else mark = warpCount(R & ~nbPla) <= 1;
    }
    if(mark) { resPla |= R; resOpp &= ~R; }
    else if(unsafeBigTerritories && !containsOpp) resPla |= R & ~resOpp;
  }
}

// Board::calculateArea(result, nonPassAliveStones, safeBigTerritories, unsafeBigTerritories, isMultiStoneSuicideLegal)
__device__ __forceinline__ void boardCalculateArea(const WarpBoard& bd, bool nonPassAliveStones, bool safeBig, bool unsafeBig, bool multiStoneSuicideLegal,
                                                   uint32_t& areaB, uint32_t& areaW) {
  areaB = 0; areaW = 0;
  bensonForPla(bd, true, multiStoneSuicideLegal, safeBig, unsafeBig, areaB, areaW);
  bensonForPla(bd, false, multiStoneSuicideLegal, safeBig, unsafeBig, areaW, areaB);
  if(nonPassAliveStones) {
    const uint32_t unmarked = ~(areaB | areaW);
    areaB |= bd.b & unmarked;
    areaW |= bd.w & unmarked;
  }
}

// Area score black minus white (komi not included) as BoardHistory::endAndScoreGameNow counts it under area scoring:
// Board::calculateArea with nonPassAliveStones, safeBigTerritories, unsafeBigTerritories all on (boardhistory.cpp, countAreaScoreWhiteMinusBlack).
__device__ __forceinline__ int boardAreaScoreBlackMinusWhite(const WarpBoard& bd, bool multiStoneSuicideLegal) {
  uint32_t aB, aW;
  boardCalculateArea(bd, true, true, true, multiStoneSuicideLegal, aB, aW);
  return warpCount(aB) - warpCount(aW);
}

}  // namespace kgb
