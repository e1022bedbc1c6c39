// oracle/ref_record_driver.cpp - TEST INFRASTRUCTURE.  The integrated configuration of INTEGRATION.md §6 in one binary: games are
// played by libkgb200's device loop (through integration/b200selfplay.h), integration/b200record.h fills the reference's own
// FinishedGameData, and the UNMODIFIED reference TrainingDataWriter (dataio/trainingwrite.cpp) turns them into rows - recomputing
// every input plane with its own NNInputs::fillRowV7 from its own Board / BoardHistory.  The rows go to the writer's text sink;
// tests/test_game_recorder.py compares them with the rows katago_b200/game_recorder.py + npz_writer.py produce for the same games.
// Built by oracle/Makefile.drivers into oracle/_ref/kgref_record (needs /root/reference to build, not to run).
//
//   kgref_record MODEL LEN CONFIG.bin NUMGAMES OUT.txt      CONFIG.bin = the bytes of the kgb_selfplay_config the test uses itself
#include "integration/b200record.h"

#include <fstream>
#include <iostream>

using namespace std;

namespace Version {  // main.cpp normally defines these (cpp/main.h)
  std::string getKataGoVersion() { return "ref_record_driver"; }
  std::string getKataGoVersionForHelp() { return "ref_record_driver"; }
  std::string getKataGoVersionFullInfo() { return "ref_record_driver"; }
  std::string getGitRevision() { return "<none>"; }
  std::string getGitRevisionWithBackend() { return "<none>"; }
}

static void ck(int rc, const char* what) { if(rc != 0) { cerr << what << ": " << kgb_last_error() << endl; exit(2); } }

int main(int argc, char** argv) {
  if(argc != 6) { cerr << "usage: kgref_record MODEL LEN CONFIG.bin NUMGAMES OUT.txt   (CONFIG.bin = the bytes of a kgb_selfplay_config)" << endl; return 1; }
  const int L = atoi(argv[2]);
  const int numGames = atoi(argv[4]);
  kgb_selfplay_config c;
  {
    ifstream in(argv[3], ios::binary);
    in.read((char*)&c, sizeof(c));
    if(in.gcount() != (std::streamsize)sizeof(c) || in.peek() != EOF) { cerr << "kgref_record: CONFIG.bin is not a kgb_selfplay_config (" << sizeof(c) << " bytes)" << endl; return 1; }
  }
  if(!c.debug_hold_at_max_visits) { cerr << "kgref_record: the configuration must have debug_hold_at_max_visits = 1" << endl; return 1; }
  const int G = c.num_games, V = c.max_visits;
  Board::initHash();
  ScoreValue::initTables();

  ck(kgb_global_init(), "kgb_global_init");
  kgb_model* model = NULL; kgb_context* ctx = NULL; kgb_handle* handle = NULL;
  ck(kgb_model_load_file(argv[1], NULL, &model), "kgb_model_load_file");
  const int gpu = 0;
  ck(kgb_context_create(&gpu, 1, L, L, 1, model, &ctx), "kgb_context_create");
  ck(kgb_handle_create(ctx, model, G, 0, 1, 0, &handle), "kgb_handle_create");

  Rules rules;
  rules.koRule = c.ko_rule == 1 ? Rules::KO_POSITIONAL : c.ko_rule == 2 ? Rules::KO_SITUATIONAL : c.ko_rule == 3 ? Rules::KO_SPIGHT : Rules::KO_SIMPLE;
  rules.scoringRule = Rules::SCORING_AREA; rules.taxRule = Rules::TAX_NONE; rules.multiStoneSuicideLegal = c.multi_stone_suicide_legal != 0; rules.hasButton = false;
  rules.whiteHandicapBonusRule = Rules::WHB_ZERO; rules.friendlyPassOk = false; rules.komi = c.komi;

  int rc = 0;
  try {
    b200::GameSlots slots(handle, c, L, L);
    ofstream out(argv[5]);
    ofstream movesOut(string(argv[5]) + ".moves");
    {
      TrainingDataWriter writer(&out, 7, 4096, 1.0, L, L, 1, "recorder-test");
      int written = 0;
      b200::GameRecorder rec(slots, rules, V, c.draw_equivalent_wins_for_white, [&](int slot, FinishedGameData* d) {
        if(written < numGames) {
          writer.writeGame(*d);
          movesOut << "game slot " << slot << " moves " << d->endHist.moveHistory.size() << " hitTurnLimit " << d->hitTurnLimit << " noResult " << d->endHist.isNoResult
                   << " score " << d->endHist.finalWhiteMinusBlackScore << endl;
        }
        written++;
        delete d;
      });
      int steps = 0;
      while(written < numGames && steps < 2000) { rec.step(); steps++; }
      writer.flushIfNonempty();
      movesOut << "steps " << steps << " games " << written << endl;
    }
  }
  catch(const std::exception& e) { cerr << "kgref_record: " << e.what() << endl; rc = 3; }
  kgb_handle_free(handle); kgb_context_free(ctx); kgb_model_free(model);
  return rc;
}
