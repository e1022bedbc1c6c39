"""Long run of the device loop with every feature on (small board, small net, many finished games):
   python tests/gpu_checks/soak_selfplay.py [waves] [X] [games] [visits]"""
import os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from katago_b200 import NeuralNet, SelfPlay, modelgen
waves = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
X = int(sys.argv[2]) if len(sys.argv) > 2 else 9
games = int(sys.argv[3]) if len(sys.argv) > 3 else 256
visits = int(sys.argv[4]) if len(sys.argv) > 4 else 60
p = modelgen.write_model(os.path.join(tempfile.mkdtemp(), "t.bin"), "tiny_nbt", seed=3)
lm = NeuralNet.loadModelFile(p)
ctx = NeuralNet.createComputeContext([0], X, X, True, lm)
h = NeuralNet.createComputeHandle(ctx, lm, games, False, True, 0)
sp = SelfPlay(h, games, visits, komi=7.5, seed=99, cpuct_exploration=1.05, cpuct_exploration_log=0.28, fpu_reduction_max=0.2, root_fpu_reduction_max=0.0,
              value_weight_exponent=0.5, fpu_parent_weight_by_visited_policy=True, fpu_parent_weight_by_visited_policy_pow=2.0,
              root_desired_per_child_visits_coeff=2.0, subtree_value_bias_factor=0.3, subtree_value_bias_weight_exponent=0.8, use_graph_search=True,
              root_noise_enabled=True, root_policy_temperature=1.1, root_policy_temperature_early=1.5, nn_cache_size_power_of_two=16,
              root_num_symmetries_to_sample=4, use_play_selection=True, use_lcb_for_selection=True, use_non_buggy_lcb=True, lcb_stdevs=5.0,
              min_visit_prop_for_lcb=0.15, chosen_move_temperature=0.15, chosen_move_temperature_early=0.75, static_score_utility_factor=0.05,
              dynamic_score_utility_factor=0.3, dynamic_score_center_zero_weight=0.25, dynamic_score_center_scale=0.5, ladder_nodes_per_wave=256)
sp.random_openings(30)
t0 = time.time()
done = 0
while done < waves:
    sp.run(1000); h.sync(); done += 1000
    st = sp.stats()
    print(done, {k: st[k] for k in ("total_visits", "total_moves", "games_finished", "black_wins", "stalled_waves", "instant_playouts", "nn_cache_hits")}, flush=True)
st = sp.stats()
assert st["games_finished"] > games and 0 < st["black_wins"] < st["games_finished"], st
assert st["total_moves"] * visits <= st["total_visits"] + games * visits
for g in range(0, games, 37):
    colors, info = sp.game(g)
    assert info["cap_b"] >= 0 and info["cap_w"] >= 0 and info["move_num"] <= 2 * X * X
print(f"soak ok: {waves} waves in {time.time() - t0:.1f} s, {st['games_finished']} games finished, {st['total_visits'] / (time.time() - t0):.0f} visits/s")
