// `katago gatekeeper` (command/gatekeeper.cpp) as a C++-only host of the device loops: the stand-alone twin of katago_b200/gatekeeper_cli.py.
//
//   b200_gatekeeper -config gatekeeper.cfg -test-models-dir DIR -sgf-output-dir DIR -accepted-models-dir DIR -rejected-models-dir DIR
//                   [-selfplay-dir DIR] [-required-candidate-win-prop 0.5] [-no-autoreject-old-models] [-quit-if-no-nets-to-test]
//                   [-games-per-gpu N] [-override-config k=v,..] [-poll-seconds S] [-seed N] [-gpu I]
//
// Same arguments, same directory protocol: the newest net in the test directory is the candidate, the newest net in the accepted directory the
// baseline (gatekeeper.cpp:386-403); a candidate older than the baseline is rejected unplayed unless -no-autoreject-old-models (:404-408);
// otherwise numGamesPerGating games are played, candidate and baseline alternating colours (integration/b200_match.h: two device loops, one per
// net), stopping early once the verdict cannot change (:181-192); the candidate needs required-candidate-win-prop of the points, ties going to
// the candidate (:581), and its file or directory is moved to the accepted or the rejected directory (:225-238); for an accepted net the
// self-play directories are created first (:613-619).  One game record per line goes to <sgf-output-dir>/<candidate>/<16 hex>.sgfs.
// Plain C++17 over the C ABI (include/kgb200.h) - no CUDA headers, no reference headers, no Python.
#include <chrono>
#include <cmath>
#include <unistd.h>

#include "b200_config.h"
#include "b200_match.h"

using namespace b200host;

namespace {

struct Model { std::string name, file, dir; double mtime = -1; bool found = false; };

// LoadModel::findLatestModel (dataio/loadmodel.cpp:58-): the most recently modified name.bin.gz|.bin|.txt.gz|.txt file or
// name/model.bin.gz|model.txt.gz directory
Model findLatestModel(const std::string& modelsDir) {
  Model best;
  DIR* d = opendir(modelsDir.c_str());
  if(!d) return best;
  while(dirent* e = readdir(d)) {
    const std::string base = e->d_name, path = modelsDir + "/" + base;
    if(base == "." || base == "..") continue;
    struct stat st;
    if(stat(path.c_str(), &st) != 0) continue;
    Model cand;
    const double mtime = (double)st.st_mtim.tv_sec + 1e-9 * (double)st.st_mtim.tv_nsec;
    if(S_ISDIR(st.st_mode)) {
      for(const char* inner : {"model.bin.gz", "model.txt.gz", "model.bin", "model.txt"}) {
        struct stat si;
        if(stat((path + "/" + inner).c_str(), &si) == 0) { cand.file = path + "/" + inner; break; }
      }
      if(cand.file.empty()) continue;
      cand.name = base; cand.dir = path;
    }
    else if((endsWith(base, ".bin.gz") || endsWith(base, ".txt.gz") || endsWith(base, ".bin") || endsWith(base, ".txt")) && st.st_size > 0) {
      cand.name = base.substr(0, base.find('.')); cand.file = path;
    }
    else continue;
    cand.mtime = mtime; cand.found = true;
    if(!best.found || cand.mtime > best.mtime) best = cand;
  }
  closedir(d);
  return best;
}

void logLine(const std::string& s) { std::fprintf(stderr, "%s\n", s.c_str()); }
void makeDirs(const std::string& path) { makeDirsFor(path); }
// moveModel (gatekeeper.cpp:217-238): the model directory if there is one, else the file
void moveModel(const Model& m, const std::string& intoDir) {
  const std::string src = m.dir.empty() ? m.file : m.dir;
  const std::string dest = intoDir + "/" + src.substr(src.find_last_of('/') + 1);
  logLine("Moving " + src + " to " + dest);
  makeDirs(intoDir);
  if(std::rename(src.c_str(), dest.c_str()) != 0) die("cannot move " + src + " to " + dest);
}
// gatekeeper.cpp:181-192: +1 the candidate has already won enough, -1 it can no longer get there, 0 keep playing
int earlyVerdict(double candidatePoints, long gamesTallied, long gamesTotal, double requiredProp) {
  const long remaining = gamesTotal - gamesTallied;
  if(remaining <= 0) return 0;
  if(candidatePoints >= gamesTotal * requiredProp) return 1;
  if(candidatePoints + remaining + 1e-10 < gamesTotal * requiredProp) return -1;
  return 0;
}
// gatekeeper.cpp:581: the candidate wins ties
bool candidateIsAccepted(double candidatePoints, long gamesTallied, double requiredProp) { return !(candidatePoints + 1e-10 < requiredProp * gamesTallied); }

struct Args {
  std::string cfgPath, testDir, sgfDir, acceptedDir, rejectedDir, selfplayDir, overrides;
  double requiredProp = 0.5, pollSeconds = 4.0; bool noAutoreject = false, quitIfNone = false; int gamesPerGpu = 128, gpu = 0; long seed = 0;
};

struct MatchOutcome { double baselinePoints, candidatePoints; long games; };

// numGamesPerGating games baseline (bot 0) against candidate (bot 1) on the device
MatchOutcome playGatingMatch(const Cfg& cfg, const Args& a, const Model& baseline, const Model& candidate) {
  auto check = [](int rc, const char* what) { if(rc != 0) die(std::string(what) + ": " + kgb_last_error()); };
  const long total = (long)cfg.num("numGamesPerGating", 200);
  const int games = (int)std::max(2L, std::min(std::min((long)a.gamesPerGpu, (long)cfg.num("numGameThreads", a.gamesPerGpu)), total));
  int edge = 19;
  const b200::GameInitializer::Config gi = gameInitConfigFromCfg(cfg, &edge);
  Cfg searchCfg = cfg;
  if(!searchCfg.has("maxVisits")) searchCfg.kv["maxVisits"] = "150";
  kgb_selfplay_config sc = configFromCfg(searchCfg, games);
  b200::MatchPlay::Settings ms;
  ms.numGamesTotal = total; ms.drawEquivalentWinsForWhite = sc.draw_equivalent_wins_for_white; ms.noResultUtilityForWhite = sc.no_result_utility_for_white;
  ms.allowResignation = cfg.flag("allowResignation", false); ms.resignThreshold = cfg.num("resignThreshold", -0.90); ms.resignConsecTurns = (int)cfg.num("resignConsecTurns", 5);
  ms.maxVisits[0] = ms.maxVisits[1] = sc.max_visits;
  markIrrelevantKeys(cfg);
  for(const std::string& what : cfg.notBuilt) logLine("[config] NOT BUILT, ignored: " + what);

  kgb_model* models[2] = {nullptr, nullptr}; kgb_context* ctxs[2] = {nullptr, nullptr}; kgb_handle* handles[2] = {nullptr, nullptr};
  std::unique_ptr<b200::GameSlots> loops[2];
  const Model* nets[2] = {&baseline, &candidate};
  for(int i = 0; i < 2; i++) {
    check(kgb_model_load_file(nets[i]->file.c_str(), nullptr, &models[i]), "loading a model");
    check(kgb_context_create(&a.gpu, 1, edge, edge, 1, models[i], &ctxs[i]), "creating an evaluator context");
    check(kgb_handle_create(ctxs[i], models[i], games, 0, /*inputs_nhwc=*/1, a.gpu, &handles[i]), "creating an evaluator handle");
    kgb_selfplay_config c = sc;
    c.seed = (uint64_t)(a.seed * 7919 + 31 * i + 1);
    c.max_playouts_per_wave = 0;
    loops[i].reset(new b200::GameSlots(handles[i], c, edge, edge));
  }
  makeDirs(a.sgfDir + "/" + candidate.name);
  b200::RowRand nameRand("gatekeeper" + std::to_string(a.seed) + ":" + candidate.name);
  const uint64_t lo = nameRand.nextUInt(), hi = nameRand.nextUInt();
  char sgfName[32];
  std::snprintf(sgfName, sizeof(sgfName), "%016llX.sgfs", (unsigned long long)(lo | (hi << 32)));
  std::ofstream sgfs(a.sgfDir + "/" + candidate.name + "/" + sgfName, std::ios::app);
  if(!sgfs) die("cannot write the game records");

  b200::GameInitializer init(gi, (uint64_t)a.seed ^ 0x4761746BULL);
  MatchOutcome out{0, 0, 0};
  {
    b200::MatchPlay* mpPtr = nullptr;
    b200::MatchPlay mp(*loops[0], *loops[1], baseline.name, candidate.name, ms, &init,
                       [&](int, const b200::FinishedGame& game, const std::string& bName, const std::string& wName, const std::string& result) {
      sgfs << b200::writeSgf(game, bName, wName) << "\n";
      sgfs.flush();
      const std::string what = result == "Void" ? "noresult" : result == "0" ? "draw " + result :
                               std::string("winner ") + (result[0] == 'B' ? "black " + bName : "white " + wName) + " " + result;
      logLine("Game " + std::to_string(mpPtr->gamesTallied() - 1) + ": " + what);
    });
    mpPtr = &mp;
    mp.run((int)cfg.num("b200WavesPerPoll", 8), [&](const b200::MatchPlay& m) {
      const int v = earlyVerdict(m.winPoints(1), m.gamesTallied(), total, a.requiredProp);
      if(v > 0) logLine("Candidate has already won enough games, terminating remaning games");
      else if(v < 0) logLine("Candidate has already lost too many games, terminating remaning games");
      return v != 0;
    });
    out = MatchOutcome{mp.winPoints(0), mp.winPoints(1), mp.gamesTallied()};
  }
  for(int i = 0; i < 2; i++) { loops[i].reset(); kgb_handle_free(handles[i]); kgb_context_free(ctxs[i]); kgb_model_free(models[i]); }
  return out;
}

// One pass of the gatekeeper's main loop (gatekeeper.cpp:376-460, 560-640): "none" (nothing to test), "autorejected", "accepted" or "rejected"
std::string gateOnce(const Cfg& cfg, const Args& a) {
  const Model test = findLatestModel(a.testDir);
  if(!test.found) return "none";
  logLine("Found new candidate neural net " + test.name);
  const Model accepted = findLatestModel(a.acceptedDir);
  if(!accepted.found) { logLine("Error: No accepted model found in " + a.acceptedDir); return "none"; }
  if(test.mtime < accepted.mtime && !a.noAutoreject) {
    logLine("Rejecting " + test.name + " automatically since older than best accepted model");
    moveModel(test, a.rejectedDir);
    return "autorejected";
  }
  logLine("Loaded candidate neural net " + test.name + " from: " + test.file);
  logLine("Loaded accepted neural net " + accepted.name + " from: " + accepted.file);
  const MatchOutcome r = playGatingMatch(cfg, a, accepted, test);
  char buf[256];
  if(!candidateIsAccepted(r.candidatePoints, r.games, a.requiredProp)) {
    std::snprintf(buf, sizeof(buf), "Candidate lost match, score %.3f to %.3f in %ld games, rejecting candidate %s", r.candidatePoints, r.baselinePoints, r.games, test.name.c_str());
    logLine(buf);
    moveModel(test, a.rejectedDir);
    return "rejected";
  }
  std::snprintf(buf, sizeof(buf), "Candidate won match, score %.3f to %.3f in %ld games, accepting candidate %s", r.candidatePoints, r.baselinePoints, r.games, test.name.c_str());
  logLine(buf);
  if(!a.selfplayDir.empty()) for(const char* sub : {"", "/sgfs", "/tdata", "/vadata"}) makeDirs(a.selfplayDir + "/" + test.name + sub);
  moveModel(test, a.acceptedDir);
  return "accepted";
}

}  // namespace

int main(int argc, char** argv) {
  programName() = "b200_gatekeeper";
  Args a;
  for(int i = 1; i < argc; i++) {
    const std::string k = argv[i];
    auto next = [&]() { if(i + 1 >= argc) die("missing value after " + k); return std::string(argv[++i]); };
    if(k == "-config") a.cfgPath = next();
    else if(k == "-test-models-dir") a.testDir = next();
    else if(k == "-sgf-output-dir") a.sgfDir = next();
    else if(k == "-accepted-models-dir") a.acceptedDir = next();
    else if(k == "-rejected-models-dir") a.rejectedDir = next();
    else if(k == "-selfplay-dir") a.selfplayDir = next();
    else if(k == "-required-candidate-win-prop") a.requiredProp = std::atof(next().c_str());
    else if(k == "-no-autoreject-old-models") a.noAutoreject = true;
    else if(k == "-quit-if-no-nets-to-test") a.quitIfNone = true;
    else if(k == "-games-per-gpu") a.gamesPerGpu = std::atoi(next().c_str());
    else if(k == "-override-config") a.overrides = next();
    else if(k == "-poll-seconds") a.pollSeconds = std::atof(next().c_str());
    else if(k == "-seed") a.seed = std::atol(next().c_str());
    else if(k == "-gpu") a.gpu = std::atoi(next().c_str());
    else if(k == "-help" || k == "--help") {
      std::printf("usage: %s -config FILE -test-models-dir DIR -sgf-output-dir DIR -accepted-models-dir DIR -rejected-models-dir DIR [-selfplay-dir DIR] "
                  "[-required-candidate-win-prop P] [-no-autoreject-old-models] [-quit-if-no-nets-to-test] [-games-per-gpu N] [-override-config k=v,...] [-poll-seconds S] [-seed N] [-gpu I]\n", argv[0]);
      return 0;
    }
    else die("unknown argument " + k);
  }
  if(a.cfgPath.empty() || a.testDir.empty() || a.sgfDir.empty() || a.acceptedDir.empty() || a.rejectedDir.empty())
    die("-config, -test-models-dir, -sgf-output-dir, -accepted-models-dir and -rejected-models-dir are required (-help)");
  Cfg cfg;
  cfg.load(a.cfgPath);
  cfg.overrides(a.overrides);
  for(const std::string& d : {a.acceptedDir, a.rejectedDir, a.sgfDir}) makeDirs(d);
  if(kgb_global_init() != 0) die(std::string("kgb_global_init: ") + kgb_last_error());
  logLine("Gatekeeper Engine starting...");
  { char b[64]; std::snprintf(b, sizeof(b), "%g", a.requiredProp); logLine(std::string("Required candidate win prop: ") + b); }
  logLine("Loaded all config stuff, watching for new neural nets in " + a.testDir);
  for(;;) {
    const std::string verdict = gateOnce(cfg, a);
    if(verdict == "none") {
      if(a.quitIfNone) break;
      usleep((useconds_t)(a.pollSeconds * 1e6));
    }
    a.seed++;
  }
  logLine("All cleaned up, quitting");
  kgb_global_cleanup();
  return 0;
}
