// oracle/ref_cpu_selfplay.cpp - TEST / BASELINE INFRASTRUCTURE.  bench.py's CPU arm (`--impl reference`, `cpu_baseline`):
// the reference's own self-play hot path - the game loop of Play::runGame (program/play.cpp:1534-2354) around the unmodified Search
// (board, PUCT descent, featurisation, backup) and NNEvaluator - on the host cores, with the restated CPU NN backend (oracle/cpubackend.cpp; the
// reference's Eigen backend needs Eigen3, which this image does not have).  Threads as the reference's selfplay command arranges
// them: many game threads, each playing its own game with a single-threaded search on a tree cleared every move; the evaluator's
// server threads (one compute handle each, batch <= 2 like the Eigen build: program/setup.cpp:278-286) do the NN work.
//
//   kgref_cpu_selfplay MODEL SECONDS [gameThreads=2*cores] [serverThreads=cores] [maxVisits=600] [boardSize=19] [warmupSeconds=0] [openingMax=150]
//
// Every game starts from its own random legal play-out of 0..openingMax moves, like the games of bench.py's GPU arm (a self-play server
// holds games at all stages; all games searching the empty board would mostly hit the evaluation cache).
//
// Prints one JSON line: visits = root visits (Search::getRootVisits) accumulated by all games inside the timed window, i.e. the
// BASELINE.json metric; plus the evaluator's row and batch counts as the cross-check SURVEY.md §8d names.
#include "core/global.h"
#include "core/config_parser.h"
#include "core/logger.h"
#include "core/rand.h"
#include "core/timer.h"
#include "game/board.h"
#include "game/boardhistory.h"
#include "neuralnet/nneval.h"
#include "program/play.h"
#include "program/playsettings.h"
#include "search/search.h"
#include "search/searchparams.h"

#include <atomic>
#include <chrono>
#include <iostream>
#include <thread>

using namespace std;

namespace Version {  // main.cpp normally defines these (cpp/main.h)
  std::string getKataGoVersion() { return "ref_cpu_selfplay"; }
  std::string getKataGoVersionForHelp() { return "ref_cpu_selfplay"; }
  std::string getKataGoVersionFullInfo() { return "ref_cpu_selfplay"; }
  std::string getGitRevision() { return "<none>"; }
  std::string getGitRevisionWithBackend() { return "<none>"; }
}

int main(int argc, char** argv) {
  if(argc < 3) { cerr << "usage: kgref_cpu_selfplay MODEL SECONDS [gameThreads] [serverThreads] [maxVisits] [boardSize] [warmupSeconds]" << endl; return 1; }
  const string modelFile = argv[1];
  const double seconds = atof(argv[2]);
  int cores = (int)std::thread::hardware_concurrency();
  if(cores <= 0) cores = 8;
  const int gameThreads = argc > 3 && atoi(argv[3]) > 0 ? atoi(argv[3]) : 2 * cores;
  const int serverThreads = argc > 4 && atoi(argv[4]) > 0 ? atoi(argv[4]) : cores;
  const int maxVisits = argc > 5 ? atoi(argv[5]) : 600;
  const int L = argc > 6 ? atoi(argv[6]) : 19;
  const double warmupSeconds = argc > 7 ? atof(argv[7]) : 0.0;
  const int openingMax = argc > 8 ? atoi(argv[8]) : 150;

  Board::initHash();
  ScoreValue::initTables();
  Logger logger(nullptr, false, false, false, false);
  ConfigParser cfg(std::map<std::string, std::string>{});
  // nnMaxBatchSize 2: what the reference fixes for its CPU backend (setup.cpp:278-286); cache 2^20 entries; random symmetries as in self-play
  NNEvaluator* nnEval = new NNEvaluator("cpu", modelFile, "", &logger, 2, L, L, true, true, 20, 16, false, "", enabled_t::False, serverThreads,
                                        vector<int>(serverThreads, -1), "cpuselfplay", true, 0, true, cfg);
  nnEval->spawnServerThreads();

  // The search block of configs/training/selfplay8mainb18.cfg (what bench.py's GPU arm runs), single-threaded search per game
  SearchParams params;
  params.maxVisits = maxVisits; params.numThreads = 1;
  params.cpuctExploration = 1.05; params.cpuctExplorationLog = 0.28; params.cpuctExplorationBase = 500;
  params.fpuReductionMax = 0.2; params.rootFpuReductionMax = 0.0;
  params.valueWeightExponent = 0.5;
  params.fpuParentWeightByVisitedPolicy = true; params.fpuParentWeightByVisitedPolicyPow = 2.0;
  params.rootDesiredPerChildVisitsCoeff = 2.0;
  params.subtreeValueBiasFactor = 0.30; params.subtreeValueBiasWeightExponent = 0.8;
  params.useGraphSearch = true;
  params.rootNoiseEnabled = true; params.rootDirichletNoiseTotalConcentration = 10.83; params.rootDirichletNoiseWeight = 0.25;
  params.rootPolicyTemperature = 1.1; params.rootPolicyTemperatureEarly = 1.5;
  params.chosenMoveTemperature = 0.15; params.chosenMoveTemperatureEarly = 0.75; params.chosenMoveTemperatureHalflife = 19;
  params.useLcbForSelection = true; params.lcbStdevs = 5.0; params.minVisitPropForLCB = 0.15;
  params.staticScoreUtilityFactor = 0.05; params.dynamicScoreUtilityFactor = 0.30; params.dynamicScoreCenterZeroWeight = 0.25;
  params.dynamicScoreCenterScale = 0.50;
  params.rootNumSymmetriesToSample = 4;
  params.rootEndingBonusPoints = 0.5; params.rootPruneUselessMoves = true;
  params.drawEquivalentWinsForWhite = 0.5;

  Rules rules;
  rules.koRule = Rules::KO_SIMPLE; rules.scoringRule = Rules::SCORING_AREA; rules.taxRule = Rules::TAX_NONE; rules.multiStoneSuicideLegal = true;
  rules.hasButton = false; rules.whiteHandicapBonusRule = Rules::WHB_ZERO; rules.friendlyPassOk = false; rules.komi = 7.5f;

  // Per game thread: the game loop of Play::runGame reduced to what costs time - a fresh (cleared) single-threaded search of
  // maxVisits per move (play.cpp:1232-1233, 2694-2699), the move chosen by the search's own temperature rule, until the game ends.
  // Visits are published from inside the search thread (the search polls shouldStopEarly between playouts), so a timed window
  // counts the visits done in it even though one 600-visit search takes minutes of CPU time.
  std::atomic<bool> stop(false);
  vector<std::atomic<int64_t>> progress(gameThreads);
  for(auto& x : progress) x.store(0);
  std::atomic<int64_t> searches(0), gamesFinished(0);
  auto worker = [&](int idx) {
    Search* bot = new Search(params, nnEval, &logger, "cpuselfplay-thread" + Global::intToString(idx));
    Rand openingRand("cpuselfplay-opening" + Global::intToString(idx));
    int64_t completed = 0;
    std::function<bool()> poll = [&]() { progress[idx].store(completed + bot->getRootVisits()); return stop.load(); };
    while(!stop.load()) {
      Board board(L, L);
      Player pla = P_BLACK;
      BoardHistory hist(board, pla, rules, 0, false);
      const int openingLen = openingMax > 0 ? (int)openingRand.nextUInt((uint32_t)openingMax + 1) : 0;
      for(int i = 0; i < openingLen && !hist.isGameFinished; i++) {
        vector<Loc> legal;
        for(int y = 0; y < L; y++) for(int x = 0; x < L; x++) { const Loc l = Location::getLoc(x, y, L); if(hist.isLegal(board, l, pla)) legal.push_back(l); }
        if(legal.empty()) break;
        hist.makeBoardMoveAssumeLegal(board, legal[openingRand.nextUInt((uint32_t)legal.size())], pla, NULL);
        pla = getOpp(pla);
      }
      for(int moveNum = 0; moveNum < 1600 && !hist.isGameFinished && !stop.load(); moveNum++) {
        bot->setPosition(pla, board, hist);
        bot->runWholeSearch(pla, &poll);
        completed += bot->getRootVisits();
        progress[idx].store(completed);
        if(stop.load()) break;
        searches.fetch_add(1);
        Loc loc = bot->getChosenMoveLoc();
        if(loc == Board::NULL_LOC || !hist.isLegal(board, loc, pla)) loc = Board::PASS_LOC;
        hist.makeBoardMoveAssumeLegal(board, loc, pla, NULL);
        pla = getOpp(pla);
      }
      if(hist.isGameFinished) gamesFinished.fetch_add(1);
    }
    delete bot;
  };
  vector<std::thread> threads;
  for(int i = 0; i < gameThreads; i++) threads.emplace_back(worker, i);
  if(warmupSeconds > 0) std::this_thread::sleep_for(std::chrono::duration<double>(warmupSeconds));
  auto total = [&]() { int64_t t = 0; for(auto& x : progress) t += x.load(); return t; };
  const int64_t rows0 = (int64_t)nnEval->numRowsProcessed(), batches0 = (int64_t)nnEval->numBatchesProcessed(), visits0 = total(), searches0 = searches.load();
  const auto t0 = chrono::steady_clock::now();
  std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
  const double wall = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
  const int64_t visitsDone = total() - visits0, searchesDone = searches.load() - searches0;
  const int64_t rows = (int64_t)nnEval->numRowsProcessed() - rows0, batches = (int64_t)nnEval->numBatchesProcessed() - batches0;
  stop.store(true);
  for(auto& t : threads) t.join();
  cout << "{\"visits\": " << visitsDone << ", \"searches_finished\": " << searchesDone << ", \"seconds\": " << wall << ", \"visits_per_s\": " << (visitsDone / wall)
       << ", \"nn_rows\": " << rows << ", \"nn_rows_per_s\": " << (rows / wall) << ", \"nn_batches\": " << batches << ", \"games_finished\": " << gamesFinished.load()
       << ", \"game_threads\": " << gameThreads << ", \"nn_server_threads\": " << serverThreads << ", \"hardware_threads\": " << cores
       << ", \"max_visits\": " << maxVisits << ", \"board\": " << L << ", \"opening_max\": " << openingMax << "}" << endl;
  delete nnEval;
  return 0;
}
