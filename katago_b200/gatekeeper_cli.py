"""`python -m katago_b200.gatekeeper_cli` - the reference's `katago gatekeeper` command (command/gatekeeper.cpp) on the device loops
(SURVEY.md §8f row 4).

    python -m katago_b200.gatekeeper_cli -config gatekeeper.cfg -test-models-dir DIR -sgf-output-dir DIR -accepted-models-dir DIR
                                         -rejected-models-dir DIR [-selfplay-dir DIR] [-required-candidate-win-prop 0.5]
                                         [-no-autoreject-old-models] [-quit-if-no-nets-to-test] [-games-per-gpu N] [-override-config k=v,..]

Same arguments, same directory protocol: the newest net in the test directory is the candidate, the newest net in the accepted directory
the baseline (gatekeeper.cpp:386-403); a candidate older than the baseline is rejected unplayed unless -no-autoreject-old-models
(:404-408); otherwise `numGamesPerGating` games are played, candidate and baseline alternating colours (katago_b200/match_play.py: two
device loops, one per net), stopping early once the verdict cannot change (:181-192); the candidate needs `required-candidate-win-prop` of
the points, ties going to the candidate (:581), and its file or directory is moved to the accepted or the rejected directory (:225-238);
for an accepted net the self-play directories are created first (:613-619).  One game record per line goes to
`<sgf-output-dir>/<candidate>/<16 hex>.sgfs`.  The search block, rules, board sizes and komi come from the reference's .cfg keys through
the same mapping as the selfplay command; resignation (`allowResignation`, `resignThreshold`, `resignConsecTurns`) is the match engine's."""
import argparse
import glob
import os
import shutil
import sys
import time


def find_latest_model(models_dir):
    """LoadModel::findLatestModel (dataio/loadmodel.cpp:58-): the most recently modified `name.bin.gz|.bin|.txt.gz|.txt` file or
    `name/model.bin.gz|model.txt.gz` directory.  (name, file, directory or None, mtime) or None."""
    best = None
    for path in glob.glob(os.path.join(models_dir, "*")):
        base = os.path.basename(path)
        if os.path.isdir(path):
            inner = [os.path.join(path, f) for f in ("model.bin.gz", "model.txt.gz", "model.bin", "model.txt") if os.path.exists(os.path.join(path, f))]
            if not inner:
                continue
            cand = (base, inner[0], path, os.path.getmtime(path))
        elif any(base.endswith(ext) for ext in (".bin.gz", ".txt.gz", ".bin", ".txt")) and os.path.getsize(path) > 0:
            cand = (base.split(".")[0], path, None, os.path.getmtime(path))
        else:
            continue
        if best is None or cand[3] > best[3]:
            best = cand
    return best


def move_model(model, into_dir, log):
    """moveModel (gatekeeper.cpp:217-238): the model directory if there is one, else the file."""
    name, path, model_dir, _ = model
    src = model_dir if model_dir is not None else path
    dest = os.path.join(into_dir, os.path.basename(src))
    log(f"Moving {src} to {dest}")
    os.makedirs(into_dir, exist_ok=True)
    shutil.move(src, dest)
    return dest


def early_verdict(candidate_points, games_tallied, games_total, required_prop):
    """gatekeeper.cpp:181-192: +1 the candidate has already won enough, -1 it can no longer get there, 0 keep playing."""
    remaining = games_total - games_tallied
    if remaining <= 0:
        return 0
    if candidate_points >= games_total * required_prop:
        return 1
    if candidate_points + remaining + 1e-10 < games_total * required_prop:
        return -1
    return 0


def candidate_is_accepted(candidate_points, games_tallied, required_prop):
    """gatekeeper.cpp:581: the candidate wins ties."""
    return not (candidate_points + 1e-10 < required_prop * games_tallied)


def play_gating_match(cfg, baseline_file, candidate_file, names, sgf_dir, games_per_gpu, required_prop, log, seed=0, gpu=0):
    """`numGamesPerGating` games baseline (bot 0) against candidate (bot 1) on the device.  Returns (baseline points, candidate points, games)."""
    from . import selfplay_cli as C
    from .game_initializer import GameInitializer
    from .match_play import MatchPlay
    from .nn_backend import NeuralNet, SelfPlay
    kw, data, report = C.selfplay_kwargs_from_cfg(cfg)
    for line in report["fixed"]:
        log("[config] " + line)
    if report["not_built"]:
        log("[config] NOT BUILT, ignored: " + "; ".join(report["not_built"]))
    total = int(cfg.get("numGamesPerGating", 200))
    games = max(2, min(games_per_gpu, int(cfg.get("numGameThreads", games_per_gpu)), total))
    L = data["board_size"]
    max_visits = kw.pop("max_visits", 150)
    loops, owned = [], []
    try:
        for i, path in enumerate((baseline_file, candidate_file)):
            lm = NeuralNet.loadModelFile(path)
            ctx = NeuralNet.createComputeContext([gpu], L, L, True, lm)
            h = NeuralNet.createComputeHandle(ctx, lm, games, False, True, gpu)
            owned += [h, ctx]
            loops.append(SelfPlay(h, games, max_visits, komi=data["komi"], seed=seed * 7919 + 31 * i + 1, debug_hold_at_max_visits=True, **kw))
        sink = C.SgfSink(sgf_dir, f"gatekeeper{seed}:{names[1]}", names[0], names[1])

        def on_game(slot, game, b_name, w_name, result):
            from .npz_writer import write_sgf
            with open(sink.path, "a") as f:
                f.write(write_sgf(game, b_name, w_name) + "\n")
            log(f"Game {mp.games_tallied - 1}: " + ("noresult" if result == "Void" else "draw " + result if result == "0" else
                                                      f"winner {'black ' + b_name if result.startswith('B') else 'white ' + w_name} {result}"))
        mp = MatchPlay(loops, names, total, GameInitializer(seed=seed ^ 0x4761746B, **data["game_init"]), on_game=on_game,
                       draw_equivalent_wins_for_white=kw.get("draw_equivalent_wins_for_white", 0.5), no_result_utility_for_white=kw.get("no_result_utility_for_white", 0.0),
                       allow_resignation=C._B(cfg.get("allowResignation", "false")), resign_threshold=float(cfg.get("resignThreshold", -0.90)),
                       resign_consec_turns=int(cfg.get("resignConsecTurns", 5)))

        def stop(m):
            v = early_verdict(m.win_points[1], m.games_tallied, total, required_prop)
            if v > 0:
                log("Candidate has already won enough games, terminating remaning games")
            elif v < 0:
                log("Candidate has already lost too many games, terminating remaning games")
            return v != 0
        mp.run(stop=stop)
        return mp.win_points[0], mp.win_points[1], mp.games_tallied
    finally:
        for sp in loops:
            sp.free()
        for o in owned:
            o.free()


def gate_once(a, cfg, log, play_match=play_gating_match):
    """One pass of the gatekeeper's main loop (gatekeeper.cpp:376-460, 560-640).  Returns "none" (nothing to test), "autorejected",
    "accepted" or "rejected"."""
    test = find_latest_model(a.test_models_dir)
    if test is None:
        return "none"
    log(f"Found new candidate neural net {test[0]}")
    accepted = find_latest_model(a.accepted_models_dir)
    if accepted is None:
        log(f"Error: No accepted model found in {a.accepted_models_dir}")
        return "none"
    if test[3] < accepted[3] and not a.no_autoreject_old_models:
        log(f"Rejecting {test[0]} automatically since older than best accepted model")
        move_model(test, a.rejected_models_dir, log)
        return "autorejected"
    log(f"Loaded candidate neural net {test[0]} from: {test[1]}")
    log(f"Loaded accepted neural net {accepted[0]} from: {accepted[1]}")
    base_pts, cand_pts, tallied = play_match(cfg, accepted[1], test[1], (accepted[0], test[0]), os.path.join(a.sgf_output_dir, test[0]), a.games_per_gpu,
                                             a.required_candidate_win_prop, log, seed=a.seed)
    if not candidate_is_accepted(cand_pts, tallied, a.required_candidate_win_prop):
        log("Candidate lost match, score %.3f to %.3f in %d games, rejecting candidate %s" % (cand_pts, base_pts, tallied, test[0]))
        move_model(test, a.rejected_models_dir, log)
        return "rejected"
    log("Candidate won match, score %.3f to %.3f in %d games, accepting candidate %s" % (cand_pts, base_pts, tallied, test[0]))
    if a.selfplay_dir:
        for sub in ("", "sgfs", "tdata", "vadata"):
            os.makedirs(os.path.join(a.selfplay_dir, test[0], sub), exist_ok=True)
    move_model(test, a.accepted_models_dir, log)
    return "accepted"


def main(argv=None):
    ap = argparse.ArgumentParser(prog="katago_b200.gatekeeper_cli", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-config", required=True)
    ap.add_argument("-test-models-dir", required=True)
    ap.add_argument("-sgf-output-dir", required=True)
    ap.add_argument("-accepted-models-dir", required=True)
    ap.add_argument("-rejected-models-dir", required=True)
    ap.add_argument("-selfplay-dir", default="")
    ap.add_argument("-required-candidate-win-prop", type=float, default=0.5)
    ap.add_argument("-no-autoreject-old-models", action="store_true")
    ap.add_argument("-quit-if-no-nets-to-test", action="store_true")
    ap.add_argument("-games-per-gpu", type=int, default=128, help="concurrent games (numGameThreads is capped by it)")
    ap.add_argument("-override-config", default="")
    ap.add_argument("-poll-seconds", type=float, default=4.0)
    ap.add_argument("-seed", type=int, default=0)
    a = ap.parse_args(argv)
    from . import selfplay_cli as C
    cfg = C.parse_cfg(a.config)
    for kv in [s for s in a.override_config.split(",") if s.strip()]:
        k, v = kv.split("=", 1)
        cfg[k.strip()] = v.strip()
    for d in (a.accepted_models_dir, a.rejected_models_dir, a.sgf_output_dir):
        os.makedirs(d, exist_ok=True)
    log = lambda s: print(s, file=sys.stderr, flush=True)
    log("Gatekeeper Engine starting...")
    log(f"Required candidate win prop: {a.required_candidate_win_prop}")
    log(f"Loaded all config stuff, watching for new neural nets in {a.test_models_dir}")
    try:
        while True:
            verdict = gate_once(a, cfg, log)
            if verdict == "none":
                if a.quit_if_no_nets_to_test:
                    break
                time.sleep(a.poll_seconds)
            a.seed += 1
    except KeyboardInterrupt:
        log("Exited cleanly after signal")
    log("All cleaned up, quitting")
    return 0


if __name__ == "__main__":
    sys.exit(main())
