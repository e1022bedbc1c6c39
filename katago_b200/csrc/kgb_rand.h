// Host-side restatement of the reference's seeded RNG and Zobrist tables, needed for bit-exact
// position hashes: Rand = XorShift1024* + PCG32 seeded from MD5/SHA-256 of the seed string (core/rand.cpp:260-320,
// core/rand_helpers.h:29-66, core/rand.h:149-185); Board::initHash draw order (game/board.cpp:151-216).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace kgb {

class RefRand {
 public:
  explicit RefRand(const std::string& seed) { init(seed); }
  void init(const std::string& seed);
  uint32_t nextUInt();
  uint64_t nextUInt64();
  // the generator's state, for upload to the device twin (kgb_devrand.cuh)
  void exportState(uint64_t a[16], uint64_t& aIdx, uint64_t& pcg) const { for(int i = 0; i < 16; i++) a[i] = a_[i]; aIdx = aIdx_; pcg = pcg_; }

 private:
  uint64_t a_[16];
  uint64_t aIdx_ = 0;
  uint64_t pcg_ = 0;
};

struct Hash128 { uint64_t h0 = 0, h1 = 0; };

// Zobrist data of Board for a given board size, re-indexed for the bitboard layout.
struct ZobristTables {
  std::vector<Hash128> board;    // [y*32 + x][2]  (colour 0 = black, 1 = white)  = ZOBRIST_BOARD_HASH[Location::getLoc(x,y,X)][colour+1]
  std::vector<Hash128> koLoc;    // [y*32 + x]     = ZOBRIST_KO_LOC_HASH[loc]
  Hash128 sizeHash;              // ZOBRIST_SIZE_X_HASH[X] ^ ZOBRIST_SIZE_Y_HASH[Y]  (pos_hash of the empty board, board.cpp Board::init)
  Hash128 player[4];             // ZOBRIST_PLAYER_HASH
};
ZobristTables makeZobristTables(int X, int Y);

void md5Words(const void* data, size_t len, uint32_t out[4]);
void sha256Words64(const void* data, size_t len, uint64_t out[4]);

}  // namespace kgb
