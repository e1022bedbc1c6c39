#include "kgb_rand.h"

#include <cstring>

#include "kgb_model.h"  // sha256Hex

namespace kgb {

// ---- MD5 (RFC 1321); words returned as the reference does: hash[i] = h_i (core/md5.cpp:45-141) ---------------------
void md5Words(const void* data, size_t len, uint32_t out[4]) {
  static const uint32_t r[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                                 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
  static const uint32_t k[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
    0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
    0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
    0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
    0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
    0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
  uint32_t h0 = 0x67452301, h1 = 0xefcdab89, h2 = 0x98badcfe, h3 = 0x10325476;
  size_t newLen = len + 1;
  while(newLen % 64 != 56) newLen++;
  std::vector<uint8_t> msg(newLen + 8, 0);
  memcpy(msg.data(), data, len);
  msg[len] = 0x80;
  uint64_t bits = 8ULL * len;
  memcpy(msg.data() + newLen, &bits, 8);  // little-endian host
  for(size_t off = 0; off < newLen + 8; off += 64) {
    uint32_t w[16];
    memcpy(w, msg.data() + off, 64);
    uint32_t a = h0, b = h1, c = h2, d = h3;
    for(uint32_t i = 0; i < 64; i++) {
      uint32_t f, g;
      if(i < 16) { f = (b & c) | (~b & d); g = i; }
      else if(i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) % 16; }
      else if(i < 48) { f = b ^ c ^ d; g = (3 * i + 5) % 16; }
      else { f = c ^ (b | ~d); g = (7 * i) % 16; }
      uint32_t t = d; d = c; c = b;
      uint32_t x = a + f + k[i] + w[g];
      b = b + ((x << r[i]) | (x >> (32 - r[i])));
      a = t;
    }
    h0 += a; h1 += b; h2 += c; h3 += d;
  }
  out[0] = h0; out[1] = h1; out[2] = h2; out[3] = h3;
}

// SHA-256 digest as four big-endian 64-bit words (core/sha2.cpp CONVERT_DIGEST_UINT64)
void sha256Words64(const void* data, size_t len, uint64_t out[4]) {
  std::string hex = sha256Hex(data, len);
  for(int i = 0; i < 4; i++) out[i] = std::stoull(hex.substr(16 * i, 16), nullptr, 16);
}

void RefRand::init(const std::string& seed) {  // core/rand.cpp:279-320
  std::string s;
  {
    uint32_t h[4];
    md5Words(seed.data(), seed.size(), h);
    s += "|";
    s += std::to_string(h[0]);
    s += "|";
    s += seed;
  }
  int counter = 0, nextIdx = 4;
  uint64_t hash[4];
  auto getNonzero = [&]() -> uint64_t {
    uint64_t v;
    do {
      if(nextIdx >= 4) {
        std::string tmp = std::to_string(counter) + s;
        counter += 37;
        sha256Words64(tmp.data(), tmp.size(), hash);
        nextIdx = 0;
      }
      v = hash[nextIdx++];
    } while(v == 0);
    return v;
  };
  for(int i = 0; i < 16; i++) a_[i] = getNonzero();
  aIdx_ = 0;
  pcg_ = getNonzero();
}

uint32_t RefRand::nextUInt() {  // rand.h:149-152 = PCG32 + XorShift1024*
  // PCG32 (rand_helpers.h:59-66)
  pcg_ = pcg_ * 6364136223846793005ULL + 1442695040888963407ULL;
  uint32_t x = (uint32_t)(((pcg_ >> 18) ^ pcg_) >> 27);
  int rot = (int)(pcg_ >> 59);
  uint32_t p = rot == 0 ? x : ((x >> rot) | (x << (32 - rot)));
  // XorShift1024* (rand_helpers.h:29-40)
  uint64_t a0 = a_[aIdx_];
  uint64_t a1 = a_[aIdx_ = (aIdx_ + 1) & 15];
  a1 ^= a1 << 31;
  a1 ^= a1 >> 11;
  a0 ^= a0 >> 30;
  a_[aIdx_] = a0 ^ a1;
  uint64_t res = a_[aIdx_] * 1181783497276652981ULL;
  return p + (uint32_t)(res >> 32);
}

uint64_t RefRand::nextUInt64() {  // rand.h:180-185
  uint64_t lo = nextUInt();
  uint64_t hi = (uint64_t)nextUInt() << 32;
  return lo | hi;
}

ZobristTables makeZobristTables(int X, int Y) {  // game/board.cpp:151-216 draw order, MAX_LEN = 19
  const int MAX_LEN = 19, MAX_ARR = (MAX_LEN + 1) * (MAX_LEN + 2) + 1;
  RefRand rand("Board::initHash()");
  auto next = [&rand]() { Hash128 h; h.h0 = rand.nextUInt64(); h.h1 = rand.nextUInt64(); return h; };
  ZobristTables z;
  for(int i = 0; i < 4; i++) z.player[i] = next();
  for(int i = 0; i < 3; i++) next();  // ZOBRIST_ENCORE_HASH
  std::vector<Hash128> boardHash(MAX_ARR * 4), koLoc(MAX_ARR);
  for(int i = 0; i < MAX_ARR; i++) {
    for(int j = 0; j < 4; j++) {
      if(j == 1 || j == 2) {
        boardHash[i * 4 + j] = next();
        next();  // ZOBRIST_KO_MARK_HASH[i][j]
      }
    }
    koLoc[i] = next();
  }
  rand.init("Board::initHash() for ZOBRIST_SIZE hashes");
  std::vector<Hash128> sx(MAX_LEN + 1), sy(MAX_LEN + 1);
  for(int i = 0; i < MAX_LEN + 1; i++) { sx[i] = next(); sy[i] = next(); }
  z.sizeHash.h0 = sx[X].h0 ^ sy[Y].h0;
  z.sizeHash.h1 = sx[X].h1 ^ sy[Y].h1;
  z.board.assign(32 * 32 * 2, Hash128());
  z.koLoc.assign(32 * 32, Hash128());
  for(int y = 0; y < Y; y++)
    for(int x = 0; x < X; x++) {
      int loc = (x + 1) + (y + 1) * (X + 1);  // Location::getLoc (board.h:52)
      z.board[(y * 32 + x) * 2 + 0] = boardHash[loc * 4 + 1];
      z.board[(y * 32 + x) * 2 + 1] = boardHash[loc * 4 + 2];
      z.koLoc[y * 32 + x] = koLoc[loc];
    }
  return z;
}

}  // namespace kgb
