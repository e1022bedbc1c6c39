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
};

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
  int c = __popc(v);
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(KGB_FULL, c, o);
  return c;
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
}

// Board::playMoveAssumeLegal (game/board.cpp:1051-1143).  p = y*32+x, or p < 0 for a pass.
__device__ __forceinline__ void boardPlay(WarpBoard& bd, int p, bool black) {
  if(p < 0) { bd.ko = -1; return; }
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

// Tromp-Taylor area score, black minus white, komi not included (Board::calculateArea with all flags on reduces to this
// for finished games; used only for terminal values inside the search this round).
__device__ __forceinline__ int boardAreaScoreBlackMinusWhite(const WarpBoard& bd) {
  const uint32_t rm = bd.rowMask;
  const uint32_t empty = ~(bd.b | bd.w) & rm;
  uint32_t reachB = flood(nbrs(bd.b, rm) & empty, empty, rm);
  uint32_t reachW = flood(nbrs(bd.w, rm) & empty, empty, rm);
  return warpCount(bd.b | (reachB & ~reachW)) - warpCount(bd.w | (reachW & ~reachB));
}

}  // namespace kgb
