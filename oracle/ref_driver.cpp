// oracle/ref_driver.cpp - TEST INFRASTRUCTURE.  Links against the unmodified reference (oracle/_ref/libkgref.a) and
// dumps golden fixtures / answers parity queries straight from the reference's own classes.  Never shipped or timed as
// product.  Built by oracle/Makefile.drivers into oracle/_ref/kgref_driver.
//
//   kgref_driver boardstream X Y NMOVES SEED MULTISUICIDE OUT.bin
//       random legal move stream on a reference Board (game/board.h): after every move records the board, ko, capture
//       counters, pos_hash, per-stone liberty counts and the legality mask of the player to move next.
#include "game/board.h"
#include "game/boardhistory.h"
#include "game/rules.h"
#include "neuralnet/nninputs.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

using namespace std;

namespace Version {  // main.cpp normally defines these (cpp/main.h)
  std::string getKataGoVersion() { return "ref_driver"; }
  std::string getKataGoVersionForHelp() { return "ref_driver"; }
  std::string getKataGoVersionFullInfo() { return "ref_driver"; }
  std::string getGitRevision() { return "<none>"; }
  std::string getGitRevisionWithBackend() { return "<none>"; }
}

struct Lcg {  // driver-local PRNG (the move stream only needs to be reproducible, not KataGo's Rand)
  uint64_t s;
  explicit Lcg(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ULL + 12345) {}
  uint32_t next() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(s >> 33); }
};

template <class T> static void put(ofstream& o, T v) { o.write((const char*)&v, sizeof(T)); }

static int cmdBoardStream(int argc, char** argv) {
  if(argc != 8) { cerr << "usage: boardstream X Y NMOVES SEED MULTISUICIDE OUT" << endl; return 1; }
  int X = atoi(argv[2]), Y = atoi(argv[3]), nMoves = atoi(argv[4]);
  uint64_t seed = strtoull(argv[5], NULL, 10);
  bool multi = atoi(argv[6]) != 0;
  Board::initHash();
  Board board(X, Y);
  Lcg rng(seed);
  ofstream out(argv[7], ios::binary);
  put<int32_t>(out, X); put<int32_t>(out, Y); put<int32_t>(out, nMoves); put<int32_t>(out, multi ? 1 : 0);
  Player pla = P_BLACK;
  for(int step = 0; step < nMoves; step++) {
    vector<Loc> legal;
    for(int y = 0; y < Y; y++)
      for(int x = 0; x < X; x++) {
        Loc loc = Location::getLoc(x, y, X);
        if(board.isLegal(loc, pla, multi)) legal.push_back(loc);
      }
    Loc mv;
    if(legal.empty() || rng.next() % 40 == 0) mv = Board::PASS_LOC;
    else mv = legal[rng.next() % legal.size()];
    board.playMoveAssumeLegal(mv, pla);
    int8_t mx = -1, my = -1;
    if(mv != Board::PASS_LOC) { mx = (int8_t)Location::getX(mv, X); my = (int8_t)Location::getY(mv, X); }
    put<int8_t>(out, mx); put<int8_t>(out, my); put<int8_t>(out, (int8_t)pla);
    int8_t kx = -1, ky = -1;
    if(board.ko_loc != Board::NULL_LOC) { kx = (int8_t)Location::getX(board.ko_loc, X); ky = (int8_t)Location::getY(board.ko_loc, X); }
    put<int8_t>(out, kx); put<int8_t>(out, ky);
    put<int16_t>(out, (int16_t)board.numBlackCaptures); put<int16_t>(out, (int16_t)board.numWhiteCaptures);
    put<uint64_t>(out, board.pos_hash.hash0); put<uint64_t>(out, board.pos_hash.hash1);
    Player next = getOpp(pla);
    for(int y = 0; y < Y; y++)
      for(int x = 0; x < X; x++) {
        Loc loc = Location::getLoc(x, y, X);
        put<uint8_t>(out, (uint8_t)board.colors[loc]);
      }
    for(int y = 0; y < Y; y++)
      for(int x = 0; x < X; x++) {
        Loc loc = Location::getLoc(x, y, X);
        int libs = (board.colors[loc] == P_BLACK || board.colors[loc] == P_WHITE) ? board.getNumLiberties(loc) : 0;
        put<uint8_t>(out, (uint8_t)std::min(libs, 255));
      }
    for(int y = 0; y < Y; y++)
      for(int x = 0; x < X; x++) {
        Loc loc = Location::getLoc(x, y, X);
        put<uint8_t>(out, board.isLegal(loc, next, multi) ? 1 : 0);
      }
    pla = next;
  }
  return 0;
}

int main(int argc, char** argv) {
  if(argc < 2) { cerr << "usage: kgref_driver <boardstream|...> ..." << endl; return 1; }
  string cmd = argv[1];
  if(cmd == "boardstream") return cmdBoardStream(argc, argv);
  cerr << "unknown command " << cmd << endl;
  return 1;
}
