// Standalone C++ host of the device-resident self-play loop: model file + .cfg in, training rows (.npz), game records (.sgfs) and a
// throughput line out.
//
// What `katago selfplay` is around its game threads (command/selfplay.cpp:34-330, program/play.cpp:1757-2163), reduced to the part a
// host has left when search, rules, features and the net all run on the GPU: read the configuration by the reference's key names
// (SearchParams: program/setup.cpp:381-760; the selfplay keys of configs/training/selfplay*.cfg), create evaluator and game slots
// through the C ABI (include/kgb200.h), run playout waves, read every finished root search into per-turn training targets
// (integration/b200_recorder.h), and write finished games as rows of <output-dir>/tdata/<16 hex>.npz and lines of
// <output-dir>/sgfs/<16 hex>.sgfs (integration/b200_npz.h) - files python/train.py's loader and shuffle.py read.
// Plain C++17 - no CUDA headers, no reference headers, no Python.
//
//   g++ -std=c++17 -O2 -I. integration/b200_selfplay_main.cpp -o b200_selfplay -Lkatago_b200 -lkgb200 -lz -Wl,-rpath,$PWD/katago_b200
//   ./b200_selfplay (-model net.bin.gz | -models-dir DIR) -config selfplay.cfg -output-dir out [-max-games-total N] [-seed S] [-override-config k=v,k=v]
//
// What it plays is what katago_b200/selfplay_cli.py plays, draw for draw (tests/test_cpp_host.py: the same files, bit for bit, on a CPU mock of the
// ABI): per-game board size / rules / komi like the reference's GameInitializer (b200_gameinit.h), komiAuto, policy-initialised openings,
// per-move search limits (cheap searches, reduced visits), per-turn targets, surprise weighting, lead targets, forked games, side positions
// (b200_recorder.h, b200_komi.h, b200_forks.h: the komi bisections, fork evaluations and side positions are jobs on side loops), model polling
// with weight hot-swap, one process per GPU (-rank / -world-size).  Options neither host has (handicap, seki forks, territory scoring ...) are
// listed as NOT BUILT and left out; -strict refuses them.  Without a CUDA device the program stops with the library's error (there is no CPU path).
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include "b200_config.h"
#include "b200_forks.h"

using namespace b200host;

namespace {

// Rendezvous of the ranks of one node through small files in <output-dir>/.b200_rendezvous (they share the output directory anyway): written
// under a temporary name and renamed, so a reader sees a whole file or none.
bool readWhole(const std::string& path, std::string& out) {
  std::ifstream f(path, std::ios::binary);
  if(!f) return false;
  std::stringstream ss; ss << f.rdbuf(); out = ss.str();
  return true;
}
void writeAtomically(const std::string& path, const std::string& bytes) {
  const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
  { std::ofstream f(tmp, std::ios::binary); if(!f) die("cannot write " + tmp); f << bytes; }
  if(std::rename(tmp.c_str(), path.c_str()) != 0) die("cannot rename " + tmp);
}
std::string waitForFile(const std::string& path, double timeoutSeconds) {
  const auto t0 = std::chrono::steady_clock::now();
  std::string bytes;
  while(!readWhole(path, bytes)) {
    if(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeoutSeconds) die("timed out waiting for " + path + " (are all ranks running with the same -output-dir, -seed and -nccl-token?)");
    usleep(20000);
  }
  return bytes;
}

// <output-dir>[/<net name>]/{tdata,sgfs}: one writer per net, as the reference keeps one per NNEvaluator (command/selfplay.cpp:178-225;
// katago_b200/selfplay_cli.py ModelOutputs).  switchTo() closes the files of the previous net and opens the new net's.
struct Outputs {
  std::string baseDir, writerSeed; bool perNet; int maxRowsPerFile; double firstFileMinRandProp; int dataLen;
  std::unique_ptr<b200::TrainingDataWriter> writer; std::ofstream sgfs; std::string netName;
  int generation = 0; long long rowsTotal = 0; size_t filesTotal = 0;

  void close() {
    if(!writer) return;
    writer->flushIfNonempty();
    rowsTotal += writer->rowCount(); filesTotal += writer->filesWritten().size();
    writer.reset();
    sgfs.close();
  }
  void switchTo(const std::string& modelPath, const std::string& nameInFile) {
    close();
    netName = perNet ? modelNameOf(modelPath) : nameInFile;
    const std::string dir = perNet ? baseDir + "/" + netName : baseDir;
    for(const std::string& d : {baseDir, dir, dir + "/tdata", dir + "/sgfs"})
      if(mkdir(d.c_str(), 0777) != 0 && errno != EEXIST) die("cannot create " + d);
    const std::string seed = generation == 0 ? writerSeed : writerSeed + ":net" + std::to_string(generation);
    writer.reset(new b200::TrainingDataWriter(dir + "/tdata", maxRowsPerFile, firstFileMinRandProp, dataLen, seed));
    b200::RowRand nameRand(seed + ":sgfs");
    const uint64_t lo = nameRand.nextUInt(), hi = nameRand.nextUInt();
    char sgfName[32];
    std::snprintf(sgfName, sizeof(sgfName), "%016llX.sgfs", (unsigned long long)(lo | (hi << 32)));
    sgfs.open(dir + "/sgfs/" + sgfName, std::ios::app);
    if(!sgfs) die("cannot write " + dir + "/sgfs/" + sgfName);
    generation++;
  }
  void addGame(const b200::FinishedGame& game) {
    writer->writeGame(game);
    sgfs << b200::writeSgf(game, netName, netName) << "\n";
  }
};

}  // namespace

int main(int argc, char** argv) {
  programName() = "b200_selfplay";
  std::string modelPath, modelsDir, cfgPath, outDir, overrides;
  long maxGamesTotal = 0, seed = 1;
  double modelPollSeconds = 20.0;
  int rank = 0, worldSize = 1, gpuIdx = -1;
  bool ncclWeights = false; std::string ncclToken;
  long restartGeneration = 0;            // internal: how often this run has re-started itself for a net of another architecture
  bool printOnly = false, strict = false;
  for(int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() { if(i + 1 >= argc) die("missing value after " + a); return std::string(argv[++i]); };
    if(a == "-model") modelPath = next();
    else if(a == "-models-dir") modelsDir = next();
    else if(a == "-config") cfgPath = next();
    else if(a == "-output-dir") outDir = next();
    else if(a == "-override-config") overrides = next();
    else if(a == "-print-config") printOnly = true;
    else if(a == "-strict") strict = true;
    else if(a == "-seed") seed = std::atol(next().c_str());
    else if(a == "-model-poll-seconds") modelPollSeconds = std::atof(next().c_str());
    else if(a == "-rank") rank = std::atoi(next().c_str());
    else if(a == "-world-size") worldSize = std::atoi(next().c_str());
    else if(a == "-gpu") gpuIdx = std::atoi(next().c_str());
    else if(a == "-nccl-weights") ncclWeights = true;
    else if(a == "-nccl-token") ncclToken = next();
    else if(a == "-restart-generation") restartGeneration = std::atol(next().c_str());
    else if(a == "-max-games-total") maxGamesTotal = std::atol(next().c_str());
    else if(a == "-help" || a == "--help") {
      std::printf("usage: %s (-model FILE | -models-dir DIR) -config FILE -output-dir DIR [-max-games-total N] [-seed S] [-model-poll-seconds T] [-rank R -world-size N] [-gpu I] [-nccl-weights [-nccl-token T]] [-override-config k=v,...] [-strict] [-print-config]\n", argv[0]);
      return 0;
    } else die("unknown argument " + a);
  }
  if(cfgPath.empty() || (!printOnly && ((modelPath.empty() == modelsDir.empty()) || outDir.empty())))
    die("-config, -output-dir and one of -model / -models-dir are required (-help)");
  // `katago selfplay -models-dir DIR`: the newest net of the directory, its files under <output-dir>/<net name>/ (command/selfplay.cpp:178-225)
  if(!printOnly && !modelsDir.empty()) modelPath = newestModel(modelsDir);
  // One process per GPU: games are independent, so ranks share nothing - each plays its share of the games on its own GPU with its own seeds
  // and writes its own files into the common directories (names come from the writers' Rand streams); katago_b200/selfplay_cli.py shard_plan
  if(worldSize < 1 || rank < 0 || rank >= worldSize) die("-rank must lie in 0 .. -world-size - 1");
  if(maxGamesTotal > 0 && restartGeneration == 0) maxGamesTotal = maxGamesTotal / worldSize + (rank < maxGamesTotal % worldSize ? 1 : 0);      // (a re-started run is handed its own remainder)
  if(gpuIdx < 0) gpuIdx = rank;
  const uint64_t loopSeed = (uint64_t)seed * 1000003ULL + (uint64_t)rank + 7919ULL * (uint64_t)restartGeneration;       // (selfplay_cli.py: loop_seed + 7919 * swaps for a rebuilt evaluator)
  Cfg cfg;
  cfg.load(cfgPath);
  cfg.overrides(overrides);

  auto check = [](int rc, const char* what) { if(rc != 0) die(std::string(what) + ": " + kgb_last_error()); };
  const int numGames = (int)cfg.num("numGameThreads", 256);          // concurrent games = the evaluator's batch
  // board size, ko / suicide rule and komi of every game: drawn on the host like the reference's GameInitializer (b200_gameinit.h), applied by the
  // device when the slot's next game starts; the evaluator's frame (= the data frame, dataBoardLen) holds the largest board
  int edge = 19;
  const b200::GameInitializer::Config gi = gameInitConfigFromCfg(cfg, &edge);
  // komi-bisection searches on side loops (b200_komi.h): komiAuto = the fair komi of the next game's empty board becomes the mean of its komi draw
  // (makeGameFairForEmptyBoard, play.cpp:1563-1575); estimateLeadProb = lead targets of recorded turns (play.cpp:2290-2324)
  const bool komiAuto = cfg.flag("komiAuto", false);
  const int compensateKomiVisits = (int)cfg.num("compensateKomiVisits", 20), estimateLeadVisits = (int)cfg.num("estimateLeadVisits", 6);
  const double estimateLeadProb = cfg.num("estimateLeadProb", 0.0);
  const double sidePositionProb = cfg.num("forkSidePositionProb", 0.0);      // PlaySettings::sidePositionProb (playsettings.cpp: cfg key forkSidePositionProb)
  // forked games (Play::maybeForkGame, play.cpp:2413-2508; b200_forks.h)
  b200::ForkManager::Settings forkSettings;
  forkSettings.earlyForkGameProb = cfg.num("earlyForkGameProb", 0.0); forkSettings.earlyForkGameExpectedMoveProp = cfg.num("earlyForkGameExpectedMoveProp", 0.0);
  forkSettings.forkGameProb = cfg.num("forkGameProb", 0.0); forkSettings.forkGameMinChoices = (int)cfg.num("forkGameMinChoices", 1);
  forkSettings.earlyForkGameMaxChoices = (int)cfg.num("earlyForkGameMaxChoices", 1); forkSettings.forkGameMaxChoices = (int)cfg.num("forkGameMaxChoices", 1);
  forkSettings.forkCompensateKomiProb = cfg.num("forkCompensateKomiProb", cfg.num("handicapCompensateKomiProb", 0.0));
  // policy-initialised openings (initializeGameUsingPolicy): the device draws the moves, the host the count per game
  const bool policyInit = cfg.flag("initGamesWithPolicy", false) && cfg.num("policyInitAreaProp", 0.04) > 0;
  const double policyInitAreaProp = cfg.num("policyInitAreaProp", 0.04), policyInitTemperature = cfg.num("policyInitAreaTemperature", 1.0);
  if(maxGamesTotal <= 0 && worldSize == 1) maxGamesTotal = (long)cfg.num("numGamesTotal", 0);
  const int wavesPerPoll = (int)cfg.num("b200WavesPerPoll", 16);
  const bool useFP16 = cfg.flag("b200UseFP16", true);          // false = the fp32-equivalent evaluator (3-term split-fp16 on the tensor pipe)
  cfg.num("numGamesTotal", 0); cfg.num("logGamesEvery", 50);
  kgb_selfplay_config sc = configFromCfg(cfg, numGames);
  if(!cfg.has("searchRandSeed")) sc.seed = loopSeed;            // every rank its own games
  const double policySurpriseDataWeight = cfg.num("policySurpriseDataWeight", 0.0), valueSurpriseDataWeight = cfg.num("valueSurpriseDataWeight", 0.0);
  const bool useSearchValueSurprise = cfg.flag("useSearchValueSurprise", false);
  cfg.num("maxRowsPerTrainFile", 20000); cfg.num("firstFileRandMinProp", 1.0);
  // search limits per move (getSearchLimitsThisMove, program/play.cpp:1093-1223): cheap searches and reduced visits
  b200::PlaySettings play;
  play.cheapSearchProb = cfg.num("cheapSearchProb", 0.0); play.cheapSearchVisits = (int)cfg.num("cheapSearchVisits", 0);
  play.cheapSearchTargetWeight = cfg.num("cheapSearchTargetWeight", 0.0); play.reduceVisits = cfg.flag("reduceVisits", false);
  play.reduceVisitsThreshold = cfg.num("reduceVisitsThreshold", 100.0); play.reduceVisitsThresholdLookback = (int)cfg.num("reduceVisitsThresholdLookback", 1);
  play.reducedVisitsMin = (int)cfg.num("reducedVisitsMin", 0); play.reducedVisitsWeight = cfg.num("reducedVisitsWeight", 1.0);
  markIrrelevantKeys(cfg);
  for(const auto& e : cfg.kv)
    if(!cfg.used.count(e.first)) cfg.notBuilt.push_back(e.first + " = " + e.second);
  for(const std::string& what : cfg.notBuilt) std::fprintf(stderr, "b200_selfplay: NOT BUILT (the loop runs WITHOUT it): %s\n", what.c_str());
  if(strict && !cfg.notBuilt.empty()) die("-strict: options that are not built (listed above)");

  if(printOnly) { printConfig(sc); return 0; }

  check(kgb_global_init(), "kgb_global_init");
  kgb_model* model = nullptr;
  check(kgb_model_load_file(modelPath.c_str(), nullptr, &model), "loading the model");
  kgb_model_info info;
  check(kgb_model_get_info(model, &info), "kgb_model_get_info");
  kgb_context* ctx = nullptr;
  const int gpu = gpuIdx;
  check(kgb_context_create(&gpu, 1, edge, edge, useFP16 ? 1 : 0, model, &ctx), "creating the evaluator context");
  kgb_handle* handle = nullptr;
  check(kgb_handle_create(ctx, model, numGames, /*require_exact_nn_len=*/0, /*inputs_nhwc=*/1, gpu, &handle), "creating the evaluator handle");   // (games may be smaller than the frame)

  int rc = 0;
  try {
    const std::string writerSeed = "selfplay" + std::to_string(seed) + ":rank" + std::to_string(rank) + "of" + std::to_string(worldSize);
    Outputs outputs;
    outputs.baseDir = outDir; outputs.writerSeed = writerSeed; outputs.perNet = !modelsDir.empty();
    outputs.maxRowsPerFile = (int)cfg.num("maxRowsPerTrainFile", 20000); outputs.firstFileMinRandProp = cfg.num("firstFileRandMinProp", 1.0); outputs.dataLen = edge;
    outputs.generation = (int)restartGeneration;
    outputs.switchTo(modelPath, info.name);
    // the reference's own log lines (command/selfplay.cpp, program/selfplaymanager.cpp:290-303), on stderr
    auto logLine = [](const std::string& msg) { std::fprintf(stderr, "%s\n", msg.c_str()); };
    const long logGamesEvery = std::max(1L, (long)cfg.num("logGamesEvery", 50));
    logLine("Found new neural net " + outputs.netName);
    logLine("Loaded latest neural net " + outputs.netName + " from: " + modelPath);

    // -nccl-weights (one process per GPU of one node): rank 0 alone polls, reads and packs a new net; the packed weight arena travels from its
    // device memory into the other ranks' by the library's own ncclBroadcast (kgb_handle_broadcast_staged_weights), every rank commits between
    // two waves.  The 128-byte NCCL id and the swap announcements go through files in the shared output directory.
    const std::string rendezvous = outDir + "/.b200_rendezvous", token = "seed" + std::to_string(seed) + (ncclToken.empty() ? "" : "." + ncclToken);
    const bool collective = ncclWeights && worldSize > 1;
    if(collective) {
      if(modelsDir.empty()) die("-nccl-weights needs -models-dir");
      if(komiAuto || estimateLeadProb > 0 || sidePositionProb > 0 || forkSettings.earlyForkGameProb > 0 || forkSettings.forkGameProb > 0)
        die("-nccl-weights: the side loops' handles are not part of the weight broadcast (komiAuto, estimateLeadProb, forks, side positions)");
      if(mkdir(rendezvous.c_str(), 0777) != 0 && errno != EEXIST) die("cannot create " + rendezvous);
      std::string id(128, '\0');
      if(rank == 0) { check(kgb_nccl_unique_id(&id[0]), "kgb_nccl_unique_id"); writeAtomically(rendezvous + "/nccl_id." + token, id); }
      else id = waitForFile(rendezvous + "/nccl_id." + token, 300.0);
      if(id.size() != 128) die("bad NCCL id file");
      check(kgb_handle_comm_init(handle, id.data(), rank, worldSize), "kgb_handle_comm_init");
    }

    b200::GameSlots slots(handle, sc, edge, edge);
    // side loops for the komi searches: own handles of the same net, a few slots, the loop's parameters without root noise (getNoiselessParams,
    // playutils.cpp:372-387), numVisits visits; as katago_b200/selfplay_cli.py make_aux
    struct SideLoop { kgb_handle* handle = nullptr; std::unique_ptr<b200::GameSlots> slots; std::unique_ptr<b200::KomiSearcher> searcher; };
    SideLoop fairLoop, leadLoop, sideLoop;
    auto makeSide = [&](SideLoop& side, int visits, bool noiseless = true) {
      kgb_selfplay_config c = sc;
      if(noiseless) {
        c.root_noise_enabled = 0; c.root_policy_temperature = 1.0; c.root_policy_temperature_early = 1.0; c.root_fpu_reduction_max = sc.fpu_reduction_max;
        c.root_fpu_loss_prop = sc.fpu_loss_prop; c.root_desired_per_child_visits_coeff = 0.0; c.root_num_symmetries_to_sample = 1;
      }
      c.max_moves = (sc.max_moves > 0 ? sc.max_moves : 2 * edge * edge) + 8;
      c.num_games = std::max(4, std::min(32, numGames / 4)); c.max_visits = noiseless ? std::max(2, visits) : visits;
      c.seed = loopSeed + (noiseless ? 104729 : 1299709); c.max_playouts_per_wave = 0;
      check(kgb_handle_create(ctx, model, c.num_games, 0, /*inputs_nhwc=*/1, gpu, &side.handle), "creating a side evaluator handle");
      side.slots.reset(new b200::GameSlots(side.handle, c, edge, edge));
      side.searcher.reset(new b200::KomiSearcher(*side.slots, c.max_visits));
    };
    b200::ForkManager forks(forkSettings, loopSeed ^ 0x466F726BULL);
    const bool forkNeedsLoop = forks.enabled() && !(komiAuto || estimateLeadProb > 0);        // the fork's evaluations need some side loop
    if(komiAuto || forkNeedsLoop) makeSide(fairLoop, compensateKomiVisits);
    if(estimateLeadProb > 0) makeSide(leadLoop, estimateLeadVisits);
    // side positions are searched like the game's own turns: the loop's parameters, noise and all, full visits
    if(sidePositionProb > 0) {
      makeSide(sideLoop, sc.max_visits, false);
      sideLoop.searcher->readPosition = [](const b200::GameSlots& loop, int slot) { return std::static_pointer_cast<void>(b200::HostRecorder::sidePositionFrom(loop, slot)); };
    }
    b200::KomiSearcher* forkSearcher = leadLoop.searcher ? leadLoop.searcher.get() : fairLoop.searcher.get();      // where fork evaluations and their komi compensation run
    // the draws become the games in progress (none has started), new ones are drawn for the games after them; a slot's draw for the game
    // after next is made when its next game begins (katago_b200/selfplay_cli.py SlotSetups)
    b200::GameInitializer init(gi, loopSeed ^ 0x47616D65ULL);
    std::vector<b200::GameSlots::GameSetup> setups((size_t)numGames); std::vector<float> komis((size_t)numGames);
    auto drawInto = [&](int g) { const b200::GameInitializer::Game d = init.draw(); setups[(size_t)g] = {d.x, d.y, d.koRule, d.multiStoneSuicideLegal}; komis[(size_t)g] = d.komi; };
    std::vector<int32_t> openings((size_t)numGames, 0);
    auto drawOpenings = [&]() { if(policyInit) for(int g = 0; g < numGames; g++) openings[(size_t)g] = init.openingLength(setups[(size_t)g].x, setups[(size_t)g].y, policyInitAreaProp); };
    for(int g = 0; g < numGames; g++) drawInto(g);
    drawOpenings();
    slots.setGameSetups(setups, true); slots.setKomis(komis, true);
    if(policyInit) slots.setPolicyInit(openings, policyInitTemperature, true);
    for(int g = 0; g < numGames; g++) drawInto(g);
    drawOpenings();
    slots.setGameSetups(setups); slots.setKomis(komis);
    if(policyInit) slots.setPolicyInit(openings, policyInitTemperature);
    // komiAuto: the komi at which the net calls the empty board of the slot's NEXT game even becomes the mean the komi noise is drawn around.
    // Searched on the side loop while the slot's current game is played; the answer replaces the komi handed over so far unless that game has
    // started meanwhile (the slot's serial has moved on).
    std::vector<long> serial((size_t)numGames, 0);
    auto askFairKomi = [&](int g) {
      const b200::GameSlots::GameSetup setup = setups[(size_t)g];
      const long mine = ++serial[(size_t)g];
      fairLoop.searcher->submit(setup, {}, [&, g, setup, mine](const b200::KomiOracle& ev) {
        const float fair = b200::adjustKomiToEven(init.komiMean(), setup.x, setup.y, ev, [&]() { return init.uniform(); });
        if(serial[(size_t)g] != mine) return;
        komis[(size_t)g] = init.drawKomi(setup.x, setup.y, (double)fair);
        slots.setKomis(komis);
      });
    };
    if(komiAuto) for(int g = 0; g < numGames; g++) askFairKomi(g);
    b200::HostRecorder::Settings rs;
    rs.perGameSetups = true; rs.policyInit = policyInit;
    rs.komi = sc.komi; rs.drawEquivalentWinsForWhite = sc.draw_equivalent_wins_for_white; rs.koRule = sc.ko_rule;
    rs.multiStoneSuicideLegal = sc.multi_stone_suicide_legal != 0; rs.maxVisits = sc.max_visits;
    rs.policySurpriseDataWeight = policySurpriseDataWeight; rs.valueSurpriseDataWeight = valueSurpriseDataWeight; rs.useSearchValueSurprise = useSearchValueSurprise;
    rs.hashSeed = loopSeed; rs.weightRandSeed = writerSeed + ":weights";
    rs.play = play; rs.limitsRandSeed = loopSeed ^ 0x4C696D69ULL;        // as selfplay_cli.py: Random(loop_seed ^ 0x4C696D69)
    long written = 0, gamesStarted = numGames, gamesFinished = 0;
    b200::HostRecorder recorder(slots, rs, [&](int, const b200::FinishedGame& game) {
      if(maxGamesTotal > 0 && written >= maxGamesTotal) return;        // games that end after the last counted one are dropped, like the Python host
      outputs.addGame(game);          // a finished game's rows go to the directory of the net in use when it ended (selfplay.cpp:276-319)
      written++; gamesFinished++;
      if(forks.enabled() && forkSearcher && !game.endNoResult) {        // Play::maybeForkGame on the finished game
        std::vector<b200::Move> all;
        for(const auto& m : game.startMoves) { b200::Move mv; mv.x = m.first; mv.y = m.second; all.push_back(mv); }
        for(const auto& m : game.moves) { b200::Move mv; mv.x = m.first; mv.y = m.second; all.push_back(mv); }
        const int ko = game.koRule == "POSITIONAL" ? 1 : game.koRule == "SITUATIONAL" ? 2 : game.koRule == "SPIGHT" ? 3 : 0;
        const b200::GameSlots::GameSetup setup{game.xSize, game.ySize, ko, game.multiStoneSuicideLegal ? 1 : 0};
        const float gameKomi = game.komi;
        b200::KomiSearcher::PositionAlgorithm job;
        const size_t cap = (size_t)(sc.max_moves > 0 ? sc.max_moves : 2 * game.xSize * game.ySize);      // (a fork at the game length cap would be over before its first search)
        if(forks.job(all, setup, gameKomi, edge, edge, job, [&forks, setup, gameKomi, cap](const std::vector<b200::Move>& moves) { if(!moves.empty() && moves.size() < cap) forks.add(moves, setup, gameKomi); }))
          forkSearcher->submitPositions(setup, job);
      }
    });
    if(estimateLeadProb > 0) {
      recorder.estimateLeadProb = estimateLeadProb;
      recorder.submitLead = [&](float komi, const b200::GameSlots::GameSetup& setup, const std::vector<b200::Move>& moves, b200::HostRecorder::LeadDone done) {
        leadLoop.searcher->submit(setup, moves, [komi, done](const b200::KomiOracle& ev) { done(b200::computeLead((double)komi, ev)); });
      };
    }
    // A slot's next game has begun on the device (empty board, the setup handed over): if it is a forked game, its position is played into the
    // slot and the recorder told; then the draw for the game after it - a pooled fork replaces the draw (the forked game's board, rules and komi,
    // komi noise redrawn around it, with probability forkCompensateKomiProb first adjusted to even at that position; no opening)
    struct PendingFork { bool have = false; long id = 0; b200::ForkManager::Fork fork; };
    std::vector<PendingFork> forkNext((size_t)numGames);
    long forkIds = 0;
    if(sidePositionProb > 0) {
      recorder.sidePositionProb = sidePositionProb;
      recorder.submitSide = [&](const b200::GameSlots::GameSetup& setup, const std::vector<b200::Move>& moves, float komi, b200::HostRecorder::SideDone done) {
        sideLoop.searcher->submitPositions(setup, [moves, komi, done](b200::PositionOracle& ev) {
          const b200::PositionAnswer& a = ev(moves, komi);
          done(a.valid ? std::static_pointer_cast<b200::SidePosition>(a.payload) : nullptr);       // (invalid: the forking move ended the game)
        });
      };
    }
    auto logStats = [&]() {
      const kgb_selfplay_stats st = slots.stats();
      const unsigned long long nnRows = st.total_visits - st.nn_cache_hits - st.instant_playouts;
      logLine("Games finished: " + std::to_string(gamesFinished));
      logLine("Moves played: " + std::to_string(st.total_moves));
      logLine("Data rows: " + std::to_string(outputs.rowsTotal + (outputs.writer ? outputs.writer->rowCount() : 0)));
      logLine("NN rows: " + std::to_string(nnRows));
      logLine("NN batches: " + std::to_string(std::max(1ULL, nnRows / (unsigned long long)std::max(1, numGames))));
      logLine("NN avg batch size: " + std::to_string((double)numGames));
      logLine("NN cache hits: " + std::to_string(st.nn_cache_hits));
    };
    recorder.onGameStart = [&](int g) {
      gamesStarted++;          // the slot's next game started on the device when the previous one ended
      if(gamesStarted % logGamesEvery == 0) logLine("Started " + std::to_string(gamesStarted) + " games with " + outputs.netName);
      if(gamesStarted % std::max(1000L, logGamesEvery * 100) == 0) logStats();
      PendingFork started = forkNext[(size_t)g];
      forkNext[(size_t)g] = PendingFork();
      if(started.have) { slots.playMoves(g, started.fork.moves); recorder.startFrom(g, started.fork.moves, 2); }
      b200::ForkManager::Fork fork;
      if(forks.enabled() && forks.pop(fork)) {
        const long mine = ++serial[(size_t)g], id = ++forkIds;
        setups[(size_t)g] = fork.setup;
        komis[(size_t)g] = init.drawKomi(fork.setup.x, fork.setup.y, (double)fork.komi);
        forkNext[(size_t)g].have = true; forkNext[(size_t)g].id = id; forkNext[(size_t)g].fork = fork;
        if(forkSearcher && init.uniform() < forkSettings.forkCompensateKomiProb) {
          const b200::GameSlots::GameSetup setup = fork.setup; const float forkKomi = fork.komi;
          forkSearcher->submit(setup, fork.moves, [&, g, setup, forkKomi, mine, id](const b200::KomiOracle& ev) {
            const float fair = b200::adjustKomiToEven((double)forkKomi, setup.x, setup.y, ev, [&]() { return init.uniform(); });
            if(serial[(size_t)g] != mine || !forkNext[(size_t)g].have || forkNext[(size_t)g].id != id) return;
            komis[(size_t)g] = init.drawKomi(setup.x, setup.y, (double)fair);
            slots.setKomis(komis);
          });
        }
        slots.setGameSetups(setups); slots.setKomis(komis);
        if(policyInit) { openings[(size_t)g] = 0; slots.setPolicyInit(openings, policyInitTemperature); }
        return;
      }
      drawInto(g);
      if(komiAuto) askFairKomi(g);
      slots.setGameSetups(setups); slots.setKomis(komis);
      if(policyInit) { openings[(size_t)g] = init.openingLength(setups[(size_t)g].x, setups[(size_t)g].y, policyInitAreaProp); slots.setPolicyInit(openings, policyInitTemperature); }
    };
    const auto t0 = std::chrono::steady_clock::now();
    // New nets (command/selfplay.cpp:336-352 modelLoadLoop: re-poll the models directory; :142-231 load the newest one): between two pumps
    // the newest file of -models-dir is compared with the net in use; a new one of the same architecture is packed into the handle's shadow
    // weight arena and committed between two waves (kgb_handle_stage_weights / commit_weights), the evaluation cache of the old net is
    // dropped, games move over mid-game (the reference's switchNetsMidGame) and the output directory follows the net.
    auto lastPoll = std::chrono::steady_clock::now();
    long pumps = 0; int swaps = 0;
    std::string ignoredModel;
    bool announcedDone = false;
    auto allRanksDone = [&]() { for(int r = 0; r < worldSize; r++) { std::string x; if(!readWhole(rendezvous + "/done." + token + "." + std::to_string(r), x)) return false; } return true; };
    for(;;) {
      const bool done = maxGamesTotal > 0 && written >= maxGamesTotal;
      if(done && !collective) break;
      if(!done) {
        recorder.pump(wavesPerPoll);
        pumps++;
        if(fairLoop.searcher) fairLoop.searcher->step(8);        // the side loops advance with the main loop
        if(leadLoop.searcher) leadLoop.searcher->step(8);
        if(sideLoop.searcher) sideLoop.searcher->step(8);
      }
      if(collective) {
        // a rank that has finished its games stays in the collective until every rank has: a swap announced meanwhile needs all of them
        if(done && !announcedDone) { writeAtomically(rendezvous + "/done." + token + "." + std::to_string(rank), "done"); announcedDone = true; }
        const std::string swapFile = rendezvous + "/swap." + token + "." + std::to_string(swaps + 1);
        std::string announced;
        if(rank == 0 && !done) {
          const auto now = std::chrono::steady_clock::now();
          if(std::chrono::duration<double>(now - lastPoll).count() >= modelPollSeconds) {
            lastPoll = now;
            const std::string newest = newestModel(modelsDir);
            kgb_model* next = nullptr;
            if(newest != modelPath && newest != ignoredModel) {
              if(kgb_model_load_file(newest.c_str(), nullptr, &next) != 0 || kgb_handle_stage_weights(handle, next) != 0) {
                std::fprintf(stderr, "b200_selfplay: %s: %s; keeping %s\n", newest.c_str(), kgb_last_error(), outputs.netName.c_str());
                if(next) { kgb_model_free(next); ignoredModel = newest; }
              }
              else {
                check(kgb_handle_wait_staged(handle), "kgb_handle_wait_staged");
                kgb_model_free(model); model = next;
                writeAtomically(swapFile, newest);         // the other ranks enter the broadcast when they see this
                announced = newest;
              }
            }
          }
        }
        else if(rank != 0) readWhole(swapFile, announced);
        if(!announced.empty()) {
          float ms = 0.0f;
          check(kgb_handle_broadcast_staged_weights(handle, 0, &ms), "kgb_handle_broadcast_staged_weights");
          check(kgb_handle_commit_weights(handle), "committing the new weights");
          slots.clearNNCache();
          modelPath = announced;
          swaps++;
          outputs.switchTo(modelPath, modelNameOf(modelPath));
          logLine("Model loading loop thread loaded new neural net " + outputs.netName);
          std::fprintf(stderr, "Game loop changing midgame to new neural net: %s (swap %d, after pump %ld, ncclBroadcast %.3f ms)\n", outputs.netName.c_str(), swaps, pumps, ms);
        }
        else if(done) { if(allRanksDone()) break; usleep(20000); }       // (an announcement is always answered before leaving)
        continue;
      }
      if(modelsDir.empty() || (maxGamesTotal > 0 && written >= maxGamesTotal)) continue;
      const auto now = std::chrono::steady_clock::now();
      if(std::chrono::duration<double>(now - lastPoll).count() < modelPollSeconds) continue;
      lastPoll = now;
      const std::string newest = newestModel(modelsDir);
      if(newest == modelPath || newest == ignoredModel) continue;
      kgb_model* next = nullptr;
      if(kgb_model_load_file(newest.c_str(), nullptr, &next) != 0) {        // e.g. a file that is still being written: look again at the next poll
        std::fprintf(stderr, "b200_selfplay: %s: %s; keeping %s\n", newest.c_str(), kgb_last_error(), outputs.netName.c_str());
        continue;
      }
      if(kgb_handle_stage_weights(handle, next) != 0) {
        const std::string why = kgb_last_error();
        kgb_model_free(next);
        if(why.find("architecture") == std::string::npos && why.find("layout") == std::string::npos && why.find("largest convolution") == std::string::npos) {
          std::fprintf(stderr, "b200_selfplay: %s: %s; keeping %s\n", newest.c_str(), why.c_str(), outputs.netName.c_str());
          ignoredModel = newest;
          continue;
        }
        // Another architecture: the reference builds a new NNEvaluator for any net; here that means a new evaluator, handles and loops, and the games
        // in flight are dropped (their finished predecessors are already written) - done by starting this program afresh on the new net, with the
        // games still to play, the next output generation and loop seeds of its own (as selfplay_cli.py rebuilds its evaluator in place).
        std::fprintf(stderr, "b200_selfplay: %s: %s; rebuilding the evaluator (games in progress are abandoned)\n", newest.c_str(), why.c_str());
        outputs.close();
        std::vector<std::string> args;
        for(int i = 0; i < argc; i++) {
          const std::string a = argv[i];
          if(a == "-max-games-total" || a == "-restart-generation") { i++; continue; }
          args.push_back(a);
        }
        if(maxGamesTotal > 0) { args.push_back("-max-games-total"); args.push_back(std::to_string(maxGamesTotal - written)); }
        args.push_back("-restart-generation"); args.push_back(std::to_string(std::max((long)outputs.generation, restartGeneration + 1)));
        std::vector<char*> raw;
        for(std::string& a : args) raw.push_back(&a[0]);
        raw.push_back(nullptr);
        std::fflush(nullptr);
        execv(argv[0], raw.data());
        die(std::string("cannot re-start ") + argv[0]);
      }
      check(kgb_handle_commit_weights(handle), "committing the new weights");
      slots.clearNNCache();
      for(SideLoop* side : {&fairLoop, &leadLoop, &sideLoop})
        if(side->handle) {
          check(kgb_handle_stage_weights(side->handle, next), "staging the new weights on a side loop");
          check(kgb_handle_commit_weights(side->handle), "committing the new weights on a side loop");
          side->slots->clearNNCache();
        }
      kgb_model_info nextInfo;
      check(kgb_model_get_info(next, &nextInfo), "kgb_model_get_info");
      kgb_model_free(model); model = next; modelPath = newest;
      swaps++;
      outputs.switchTo(modelPath, nextInfo.name);
      logLine("Model loading loop thread loaded new neural net " + outputs.netName);
      std::fprintf(stderr, "Game loop changing midgame to new neural net: %s (swap %d, after pump %ld)\n", outputs.netName.c_str(), swaps, pumps);
    }
    outputs.close();
    logStats();
    logLine("Total games: " + std::to_string(gamesStarted));
    logLine("Total selfplay runtime (seconds): " + std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()));
    logLine("All cleaned up, quitting");
    for(SideLoop* side : {&fairLoop, &leadLoop, &sideLoop}) { side->searcher.reset(); side->slots.reset(); if(side->handle) kgb_handle_free(side->handle); side->handle = nullptr; }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const kgb_selfplay_stats st = slots.stats();
    std::printf("{\"games_written\": %ld, \"rows\": %lld, \"files\": %zu, \"moves\": %lld, \"net_swaps\": %d, \"visits\": %llu, \"seconds\": %.3f, \"visits_per_second\": %.1f, \"nn_cache_hits\": %llu}\n",
                written, outputs.rowsTotal, outputs.filesTotal, (long long)recorder.movesRecorded(), swaps, (unsigned long long)st.total_visits, secs,
                st.total_visits / secs, (unsigned long long)st.nn_cache_hits);
  } catch(const std::exception& e) {
    std::fprintf(stderr, "b200_selfplay: %s\n", e.what());
    rc = 1;
  }
  kgb_handle_free(handle);
  kgb_context_free(ctx);
  kgb_model_free(model);
  kgb_global_cleanup();
  return rc;
}
