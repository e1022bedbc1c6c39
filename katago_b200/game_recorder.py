"""Per-turn training targets from a finished root search, and the recorder that turns the device loop's games into
`FinishedGameData` for the training-data writer.

The reference derives a turn's targets from its `Search` right after the search ends (`extractSearchTargetsThisTurn`,
program/play.cpp:931-948).  The functions here do the same from what the device loop exposes for a root - NodeStats moments of
the root and its children, play selection values, the (possibly noised) root policy - all by move position (y * X + x, pass
last).  CPU parity: tests/test_game_recorder.py against tests/golden/searchtargets.npz (dumped from the reference `Search`).

`GameRecorder` drives a `SelfPlay` created with `debug_hold_at_max_visits=True`: a game whose search is finished idles until the
recorder has read its root (a handful of small device reads per game and move, against a search of max_visits waves) and releases
it; the device then chooses and plays the move exactly as it does without a recorder.  Finished games go to a `TrainingDataWriter`.
Target weights: every turn enters with weight 1 (no cheap searches / reduced visits on the device) and is then redistributed by policy and
value surprise like runGame does (surprise_target_weights, parity against whole reference games in tests/golden/rungame.json.gz).
Not recorded (the reference's options that the device loop does not have): side positions, lead estimation (hasLead only on the
outcome entry), reanalysis, net changes.
"""
import math

import numpy as np

from .npz_writer import FinishedGameData, P_BLACK, P_WHITE, SidePosition, final_value_targets, pack_bits, policy_target_from_play_selection, scoring_from_area

_f32 = np.float32


def reported_search_values(moments):
    """ReportedSearchValues (search/reportedsearchvalues.cpp:10-51) from NodeStats moments (winLossValueAvg, noResultValueAvg,
    scoreMeanAvg, scoreMeanSqAvg, leadAvg), white's perspective: (winValue, lossValue, noResultValue, winLossValue, expectedScore)."""
    win_loss, no_result, score_mean = float(moments[0]), float(moments[1]), float(moments[2])
    win_loss = min(max(win_loss, -1.0), 1.0)
    no_result = min(max(no_result, 0.0), 1.0 - abs(win_loss))
    win = min(max(0.5 * (win_loss + (1.0 - no_result)), 0.0), 1.0)
    loss = min(max(0.5 * (-win_loss + (1.0 - no_result)), 0.0), 1.0)
    return win, loss, no_result, win_loss, score_mean


def value_targets_from_root(root_moments):
    """extractValueTargets (play.cpp:848-857): (win, loss, noResult, score, hasLead, lead) as float32; lead is not estimated."""
    win, loss, no_result, _, score = reported_search_values(root_moments)
    return (_f32(win), _f32(loss), _f32(no_result), _f32(score), 0, _f32(0.0))


def q_targets_from_children(child_moments, child_node_visits, x_size):
    """extractQValueTargets (play.cpp:859-888): one (x, y, winLoss, score, visits) per child with visits, white's perspective;
    visits are the child NODE's visits (under graph search they can exceed the edge's)."""
    out = []
    n = len(child_node_visits)
    for pos in range(n):
        v = int(child_node_visits[pos])
        if v <= 0:
            continue
        _, _, _, win_loss, score = reported_search_values(child_moments[pos])
        x, y = (-1, -1) if pos == n - 1 else (pos % x_size, pos // x_size)
        out.append((x, y, _f32(win_loss), _f32(score), v))
    return out


def policy_surprise_and_entropy(play_selection_values, policy):
    """Search::getPolicySurpriseAndEntropy (search/searchresults.cpp:631-695): KL(target || policy), entropy of the target and of
    the policy, where target = play selection values normalised.  play_selection_values: -1 where there is no child."""
    psv = np.asarray(play_selection_values, np.float64)
    pol = np.asarray(policy, np.float32)
    idx = np.flatnonzero(psv >= 0)
    total = 0.0
    for i in idx:
        total += psv[i]
    surprise = search_entropy = 0.0
    for i in idx:
        p = max(float(pol[i]), 1e-100)
        target = psv[i] / total
        if target > 1e-100:
            lt = math.log(target)
            surprise += target * (lt - math.log(p))
            search_entropy += -target * lt
    policy_entropy = 0.0
    for p in pol:
        p = float(p)
        if p > 1e-100:
            policy_entropy += -p * math.log(p)
    return max(surprise, 0.0), max(search_entropy, 0.0), max(policy_entropy, 0.0)


def policy_target_moves(play_selection_values, x_size):
    """Play::extractPolicyTarget as the sparse list the writer takes: (x, y, int16 value) for every child."""
    psv = np.asarray(play_selection_values, np.float64)
    vals = policy_target_from_play_selection(psv)
    n = len(psv)
    return [((-1, -1) if pos == n - 1 else (pos % x_size, pos // x_size)) + (int(vals[pos]),) for pos in range(n) if psv[pos] >= 0]


def value_surprise_kl(win, loss, no_result, raw):
    """valueSurpriseKL (program/play.cpp:1303-1314): KL divergence of a win / loss / noResult distribution from the raw net's (raw =
    (win, loss, noResult)), floored at 0 and capped at 1."""
    s = 0.0
    for v, r in ((win, raw[0]), (loss, raw[1]), (no_result, raw[2])):
        if v > 1e-100:
            s += v * (math.log(v) - math.log(max(float(r), 1e-100)))
    return min(max(s, 0.0), 1.0)


def compute_value_surprise_by_turn(white_value_targets, raw_nn_values, board_area, use_search_value_surprise=False):
    """computeValueSurpriseByTurn (play.cpp:1322-1352): how surprising each turn's value result was to the raw net - by default the
    smoothed forward-looking game result (exponential average back from the outcome, nowFactor = 1 / (1 + area * 0.016)), or the
    turn's own search values.  white_value_targets has one entry more than raw_nn_values (the outcome)."""
    n = len(raw_nn_values)
    if len(white_value_targets) != n + 1:
        raise ValueError("white_value_targets must have one entry per turn plus the outcome")
    f = lambda v: float(_f32(v))
    if use_search_value_surprise:
        return [value_surprise_kl(f(t[0]), f(t[1]), f(t[2]), raw_nn_values[i]) for i, t in enumerate(white_value_targets[:n])]
    now = 1.0 / (1.0 + board_area * 0.016)
    win, loss, nores = (f(v) for v in white_value_targets[-1][:3])
    out = [0.0] * n
    for i in range(n - 1, -1, -1):
        t = white_value_targets[i]
        win = win + now * (f(t[0]) - win)
        loss = loss + now * (f(t[1]) - loss)
        nores = nores + now * (f(t[2]) - nores)
        out[i] = value_surprise_kl(win, loss, nores, raw_nn_values[i])
    return out


def surprise_target_weights(target_weights, policy_surprise, value_surprise, policy_surprise_data_weight, value_surprise_data_weight):
    """The surprise weighting of Play::runGame (play.cpp:2084-2163) for games without cheap-search reanalysis: part of every turn's
    weight is redistributed in proportion to its policy surprise (for reduced-weight turns: the surprise in excess of 1.5x the
    average) and to its value surprise, keeping the game's total weight.  Returns float32 weights."""
    w = [float(_f32(x)) for x in target_weights]
    if not (policy_surprise_data_weight > 0 or value_surprise_data_weight > 0):
        return [_f32(x) for x in w]
    n = len(w)
    sum_w = sum_p = sum_v = 0.0
    for i in range(n):
        if not (0.0 <= w[i] <= 1.0):
            raise ValueError("surprise weighting expects target weights in [0, 1]")
        sum_w += w[i]
        sum_p += policy_surprise[i] * w[i]
        sum_v += value_surprise[i] * w[i]
    if sum_w < 1:
        return [_f32(x) for x in w]
    avg_p, avg_v = sum_p / sum_w, sum_v / sum_w
    vsw = value_surprise_data_weight
    if avg_v < 0.010:
        vsw *= avg_v / 0.010
    threshold = avg_p * 1.5
    p_prop = [w[i] * policy_surprise[i] + (1 - w[i]) * max(0.0, policy_surprise[i] - threshold) for i in range(n)]
    v_prop = [w[i] * value_surprise[i] for i in range(n)]
    sp = sv = 0.0
    for i in range(n):
        sp += p_prop[i]
        sv += v_prop[i]
    sp, sv = max(sp, 1e-10), max(sv, 1e-10)
    return [_f32((1.0 - policy_surprise_data_weight - vsw) * w[i] + policy_surprise_data_weight * p_prop[i] * sum_w / sp + vsw * v_prop[i] * sum_w / sv)
            for i in range(n)]


def resolve_target_weight(weight, rand):
    """resolveWeight (play.cpp:2277-2283): a fractional weight becomes floor or floor + 1 with the matching probability."""
    w = max(float(_f32(weight)), 0.0)
    floored = math.floor(w)
    return _f32(floored + 1 if rand.next_bool(float(_f32(_f32(w) - _f32(floored)))) else floored)


def search_limits_this_move(max_visits, settings, rand, historical_win_loss):
    """getSearchLimitsThisMove (program/play.cpp:1093-1223) without hint moves and asymmetric playouts: (visits, remove root noise, target
    weight, is cheap search) of the next search.  settings: PlaySettings fields by their cfg names in snake case (cheap_search_prob,
    cheap_search_visits, cheap_search_target_weight, reduce_visits, reduce_visits_threshold, reduce_visits_threshold_lookback,
    reduced_visits_min, reduced_visits_weight); rand: random.Random (the reference draws from the game's Rand);
    historical_win_loss: the root's winLossValue (white's perspective) after every search of the game so far."""
    visits, plain, weight, cheap = int(max_visits), False, _f32(1.0), False
    p = float(settings.get("cheap_search_prob", 0.0))
    if p > 0.0 and rand.random() < p:
        cv = int(settings["cheap_search_visits"])
        if cv <= 0 or cv > max_visits:
            raise ValueError("cheapSearchVisits must lie in 1..maxVisits")
        cheap, visits = True, min(visits, cv)
        weight = _f32(float(weight) * float(_f32(settings.get("cheap_search_target_weight", 0.0))))
        if float(settings.get("cheap_search_target_weight", 0.0)) <= 0.0:
            plain = True          # not recorded: no noise / temperature at the root (the reference also keeps the tree; this loop clears it)
    elif settings.get("reduce_visits", False):
        vmin = int(settings["reduced_visits_min"])
        if vmin <= 0 or vmin > max_visits:
            raise ValueError("reducedVisitsMin must lie in 1..maxVisits")
        look, thr = int(settings.get("reduce_visits_threshold_lookback", 1)), float(settings.get("reduce_visits_threshold", 100.0))
        if len(historical_win_loss) >= look:
            recent = [historical_win_loss[len(historical_win_loss) - 1 - j] for j in range(look)]
            lo, hi = (min(recent), max(recent)) if recent else (1e20, -1e20)
            most_extreme = min(max(lo, -hi), 1.0)
            through = most_extreme - thr
            if through > 0:
                prop = (through / (1.0 - thr)) ** 2
                visits = int(math.floor(visits + prop * (vmin - visits) + 0.5))
                weight = _f32(float(weight) + prop * (float(_f32(settings.get("reduced_visits_weight", 1.0))) - float(weight)))
                visits = max(visits, vmin)
    return max(2, visits), plain, weight, cheap


def extract_root_targets(sp, g, X, Y):
    """What extractSearchTargetsThisTurn / the side-position block of Play::runGame (play.cpp:931-948, 2178-2203) read of a finished search, from
    slot g of loop `sp` held at its budget: position, input row, policy / value / Q targets, surprise and entropies, NNRawStats."""
    colors, info = sp.game(g)
    spatial, glob = sp.root_row(g)
    _, policy, _ = sp.root_children(g)
    child_stats, root_stats = sp.root_value_stats(g)
    psv = sp.play_selection_values(g)
    extra = sp.root_extra(g)
    surprise, search_entropy, policy_entropy = policy_surprise_and_entropy(psv, policy)
    nn = extra["root_nn_moments"]
    flat, own = np.asarray(colors, np.uint8).reshape(-1), (P_BLACK if info["black_to_move"] else P_WHITE)
    sp_row = np.asarray(spatial, np.float32).reshape(X * Y, 22)
    # the kept row must be this root's: its own / opponent stone planes are the root position (a row of any other leaf differs)
    row_ok = bool(np.array_equal(sp_row[:, 1] != 0, flat == own) and np.array_equal(sp_row[:, 2] != 0, flat == 3 - own))
    return dict(info=info, flat=flat, own=own, policy=np.asarray(policy, np.float32), row_ok=row_ok,
                packed=pack_bits(np.transpose(sp_row.reshape(1, X * Y, 22), (0, 2, 1)))[0], global_input=np.asarray(glob, np.float32).copy(),
                policy_target=(policy_target_moves(psv, X), int(info["root_visits"])), value_targets=value_targets_from_root(root_stats),
                q_targets=q_targets_from_children(child_stats, extra["child_node_visits"], X),
                surprise=surprise, search_entropy=search_entropy, policy_entropy=policy_entropy,
                # NNRawStats (play.cpp:890-914) from the root's own evaluation; the entropy is that of the root policy before temperature and noise
                nn_raw_stats=(float(nn[0]), float(nn[2]), float(sp.root_raw_policy_entropy()[g]) if hasattr(sp, "root_raw_policy_entropy") else policy_entropy),
                raw_nn_values=reported_search_values(nn)[:3])       # Search::getRootRawNNValues: win, loss, noResult of the root's own evaluation


def choose_random_forking_move(policy, x_frame, rand, ban_pos):
    """PlayUtils::chooseRandomForkingMove (playutils.cpp): 70 % a temperature-1 policy move, 25 % a temperature-2 policy move, 5 % a uniformly random
    legal move; never `ban_pos` (the move the game actually played); the pass is allowed.  policy: by move position, -1 = illegal (here the root
    policy as searched - the reference reads the un-noised one).  Returns a move position or None."""
    r = rand.random()
    legal = [i for i in range(len(policy)) if policy[i] >= 0 and i != ban_pos]
    if r >= 0.95:
        return legal[rand.randrange(len(legal))] if legal else None
    t = 1.0 if r < 0.70 else 2.0
    cand = [i for i in legal if policy[i] > 0]
    if not cand:
        return None
    w = [float(policy[i]) ** (1.0 / t) for i in cand]
    return rand.choices(cand, weights=w)[0]


class _GameInProgress:
    def __init__(self):
        self.turns = []       # per turn: what the finished root search gave
        self.boards = []      # position before each move
        self.setup = None     # (board X, board Y, ko rule, multi-stone suicide legal) of this game, read when its first turn is recorded
        self.win_loss = []    # the root's winLossValue after each search (historicalMctsWinLossValues of Play::runGame)
        self.start_moves = [] # moves before the first recorded turn (fork prefix + policy-initialised opening)
        self.preset_moves = []   # moves the host played into the slot before the game's first search (a forked game's position)
        self.mode = 0            # FinishedGameData::mode: 0 normal, 2 fork
        self.side = {"list": [], "pending": 0, "waiting": None}     # side positions of this game: searched ones, jobs in flight, the finished game waiting for them
        self.last_policy = None  # the root policy of the turn being played (for the side position's forking move)


class GameRecorder:
    """Records every game of a `SelfPlay` in hold mode and hands finished games to `writer.write_game`.

    sp: katago_b200.nn_backend.SelfPlay created with debug_hold_at_max_visits=True.  `step()` = one move of every game:
      1. waves until every game is held at max_visits;
      2. per game: root position, the root's input row (kept on the device since the wave that evaluated the root,
         kgb_selfplay_get_root_row - any ladder budget, any number of waves ago), root / child statistics,
         play selection values, root policy -> this turn's targets (functions above);
      3. release; the next wave lets the device choose and play each move (its own Rand, temperature, LCB - unchanged) and
         evaluates the new roots; a move that ended a game leaves the final position, its area and score readable, the recorder
         builds the FinishedGameData (game-end targets of program/play.cpp:1977-2027) and the slot has already started a new game.
    The game hash (FinishedGameData::gameHash, two 64-bit draws of the game's Rand in the reference) comes from `game_hash_fn`."""

    def __init__(self, sp, writer, komi, draw_equivalent_wins_for_white=0.5, on_game=None, game_hash_fn=None,
                 policy_surprise_data_weight=0.0, value_surprise_data_weight=0.0, use_search_value_surprise=False, weight_rand=None,
                 play_settings=None, limits_rand=None, policy_init=False, lead_estimator=None, estimate_lead_prob=0.0, lead_rand=None, on_game_start=None,
                 side_searcher=None, side_position_prob=0.0):
        """policy_surprise_data_weight / value_surprise_data_weight / use_search_value_surprise: PlaySettings of the same names - the
        finished game's target weights are redistributed by surprise (surprise_target_weights).  weight_rand (a RowRand): fractional
        weights are then resolved to integers like runGame does (resolve_target_weight); None leaves them fractional for the writer,
        which draws the extra row itself."""
        self.sp, self.writer, self.X, self.Y, self.komi = sp, writer, sp.x, sp.y, float(komi)
        self.draw_eq = draw_equivalent_wins_for_white
        self.games = [_GameInProgress() for _ in range(sp.num_games)]
        self.on_game = on_game
        self.game_hash_fn = game_hash_fn or (lambda slot, index: (((slot + 1) * 0x9E3779B97F4A7C15 + index) & (2 ** 64 - 1),
                                                                  ((index + 1) * 0xC2B2AE3D27D4EB4F + slot) & (2 ** 64 - 1)))
        self.games_written = 0
        self.moves_recorded = 0
        self.policy_surprise_data_weight, self.value_surprise_data_weight = float(policy_surprise_data_weight), float(value_surprise_data_weight)
        self.use_search_value_surprise, self.weight_rand = bool(use_search_value_surprise), weight_rand
        cfg = getattr(sp, "cfg", None)             # rules for the game record (write_sgf)
        # The root's input row is kept by the device when the net evaluates the root.  With the evaluation cache on and a single root
        # evaluation, that root is normally a cache hit (it was a child of the previous tree) and no row is produced for it.
        if cfg is not None and int(getattr(cfg, "nn_cache_size_power_of_two", 0)) > 0 and int(getattr(cfg, "root_num_symmetries_to_sample", 0)) <= 1:
            raise ValueError("GameRecorder: with nn_cache_size_power_of_two > 0 the root must be evaluated by the net itself "
                             "(root_num_symmetries_to_sample >= 2, as in the stock self-play configurations), or the cache switched off")
        self.default_setup = (self.X, self.Y, int(getattr(cfg, "ko_rule", 0)), int(bool(getattr(cfg, "multi_stone_suicide_legal", 1))))
        # search limits per move (cheap searches, reduced visits): the host draws what getSearchLimitsThisMove would, the device applies it
        # to the root after each slot's next move (SelfPlay.set_next_search_limits); cur_limits[g] = (target weight, is cheap) of slot g's root
        self.policy_init_active = bool(policy_init) and hasattr(sp, "policy_init")
        # lead targets (play.cpp:2290-2324): after a game, turns drawn with estimateLeadProb get the komi-bisection searches of computeLead as
        # jobs on a side loop (katago_b200/komi_search.py KomiSearcher); the game is written when they are back
        self.lead, self.lead_prob = lead_estimator, float(estimate_lead_prob)
        if self.lead is not None:
            import random
            self.lead_rand = lead_rand or random.Random(0x4C656164)
        self.games_waiting_for_lead = 0
        # side positions (PlaySettings::sidePositionProb = cfg forkSidePositionProb, play.cpp:1846-1860, 2166-2203): with that probability per turn a
        # forking move is made off the main line and the position searched on a side loop with the game's own search parameters; its row is
        # written with the game.  Not restated: the 25 % chance of continuing a side position by the search's reply.
        self.side_searcher, self.side_prob = side_searcher, float(side_position_prob)
        if self.side_searcher is not None:
            import random
            self.side_rand = random.Random(0x53696465)
        self.on_game_start = on_game_start         # called with the slot when its next game has begun on the device (before any of its turns is recorded)
        ps = play_settings or {}
        self.play_settings = ps if (float(ps.get("cheap_search_prob", 0.0)) > 0.0 or ps.get("reduce_visits", False)) else None
        n = sp.num_games
        self.cur_limits = [(_f32(1.0), False)] * n
        if self.play_settings is not None:
            import random
            self.limits_rand = limits_rand or random.Random(0x4C696D69)
            first = [search_limits_this_move(sp.max_visits, self.play_settings, self.limits_rand, []) for _ in range(n)]
            self.next_visits = np.array([[f[0], f[0]] for f in first], np.int32)
            self.next_plain = np.array([[f[1], f[1]] for f in first], np.uint8)
            sp.set_next_search_limits(self.next_visits, self.next_plain, also_current_roots=True)
            self.cur_limits = [(f[2], f[3]) for f in first]
            self.pending = [(f, f) for f in first]
        sp.run(1)                                    # evaluates every root (its row stays on the device)

    def _held(self):
        """Slots whose search is finished: root visits have reached the root's own budget."""
        budget = self.sp.search_limits()[0] if self.play_settings is not None else self.sp.max_visits
        held = np.asarray(self.sp.root_visits()) >= budget
        if self.policy_init_active:          # a slot in its policy-drawn opening moves on by itself: those are not recorded turns
            held &= np.asarray(self.sp.policy_init()[0]) <= 0
        return held

    def _record_root(self, g):
        """Slot g is held: read its finished search and append this turn's targets (extractSearchTargetsThisTurn)."""
        sp = self.sp
        ex = extract_root_targets(sp, g, self.X, self.Y)
        info, flat, own = ex["info"], ex["flat"], ex["own"]
        target_weight, is_cheap = self.cur_limits[g]
        if not ex["row_ok"]:
            raise RuntimeError(f"GameRecorder: slot {g}, move {info['move_num']}: the wave after the previous move did not evaluate the new root "
                               "(the kept input row belongs to another position)")
        gm = self.games[g]
        if gm.setup is None:
            # board size and rules are per game (SelfPlay.set_game_setup; GameInitializer draws them per game): X, Y below stay the
            # evaluator's frame = the data frame (dataBoardLen) the rows are written in, the game's own board is its top-left corner
            gm.setup = tuple(int(v) for v in sp.game_setups()[0][g]) if hasattr(sp, "game_setups") else self.default_setup
            gm.start_moves = list(gm.preset_moves)
            if self.policy_init_active:      # the opening the device drew from the policy: the game's start history (startHist)
                gm.start_moves = gm.start_moves + sp.policy_init(max_moves=512)[2][g]
            if self.policy_init_active or gm.preset_moves:
                if len(gm.start_moves) != info["move_num"]:
                    raise RuntimeError(f"GameRecorder: slot {g}: {info['move_num']} moves played before the first searched move, {len(gm.start_moves)} opening moves kept")
        bx, by = gm.setup[0], gm.setup[1]
        gm.boards.append(np.ascontiguousarray(flat.reshape(self.Y, self.X)[:by, :bx]).reshape(-1).copy())
        values = ex["value_targets"]
        gm.win_loss.append(float(values[0]) - float(values[1]))
        gm.last_policy = ex["policy"]
        if self.play_settings is not None:       # limits of the search that follows this slot's move: the game goes on / a new game starts
            nxt = (search_limits_this_move(sp.max_visits, self.play_settings, self.limits_rand, gm.win_loss),
                   search_limits_this_move(sp.max_visits, self.play_settings, self.limits_rand, []))
            self.pending[g] = nxt
            self.next_visits[g] = (nxt[0][0], nxt[1][0]); self.next_plain[g] = (nxt[0][1], nxt[1][1])
        turn = {k: ex[k] for k in ("packed", "global_input", "policy_target", "value_targets", "q_targets", "surprise", "search_entropy", "policy_entropy",
                                   "nn_raw_stats", "raw_nn_values")}
        turn.update(next_player=own, move_num=info["move_num"], target_weight=target_weight, is_cheap_search=is_cheap)
        gm.turns.append(turn)

    def _after_move(self, g):
        """Slot g was released and one wave has run: the device has played its move and evaluated the new root."""
        sp = self.sp
        last = sp.last_move(g)
        self.games[g].turns[-1]["move"] = last["xy"]
        if self.side_searcher is not None and not last["game_over"] and self.side_rand.random() < self.side_prob:
            self._submit_side_position(g, last)
        if self.play_settings is not None:
            cont, fresh = self.pending[g]
            self.cur_limits[g] = (fresh[2], fresh[3]) if last["game_over"] else (cont[2], cont[3])
        if last["game_over"]:
            self._finish_game(g, last)

    def step(self, max_waves=1000000):
        """One move of EVERY slot (lockstep): waves until all slots are held, record, release all, one wave."""
        sp, n = self.sp, self.sp.num_games
        waves = 0
        while not self._held().all():
            sp.run(8)
            waves += 8
            if waves > max_waves:
                raise RuntimeError("GameRecorder: games did not reach max_visits")
        for g in range(n):
            self._record_root(g)
        if self.play_settings is not None:
            sp.set_next_search_limits(self.next_visits, self.next_plain)
        sp.release()
        sp.run(1)
        self.moves_recorded += n
        for g in range(n):
            self._after_move(g)
        return waves + 1

    def pump(self, waves=8):
        """Without lockstep: `waves` waves for everybody, then only the slots that are held by now are recorded and released (the
        others keep searching during the extra wave in which the released ones move).  Returns the number of moves recorded."""
        sp = self.sp
        sp.run(waves)
        held = self._held()
        if not held.any():
            return 0
        idx = [int(g) for g in np.flatnonzero(held)]
        for g in idx:
            self._record_root(g)
        if self.play_settings is not None:
            sp.set_next_search_limits(self.next_visits, self.next_plain)
        sp.release(held.astype(np.uint8))
        sp.run(1)
        self.moves_recorded += len(idx)
        for g in idx:
            self._after_move(g)
        return len(idx)

    def _finish_game(self, g, last):
        gm = self.games[g]
        X, Y, ko_rule, multi_suicide = gm.setup if gm.setup is not None else self.default_setup     # this game's own board and rules
        crop = lambda a: np.ascontiguousarray(np.asarray(a, np.uint8).reshape(self.Y, self.X)[:Y, :X]).reshape(-1).copy()
        # komi is per game (SelfPlay.set_komi; the reference's GameInitializer draws one per game): the slot's last finished game's
        komi = float(self.sp.komi_values()[1][g]) if hasattr(self.sp, "komi_values") else self.komi
        data = FinishedGameData(X, Y, komi)
        data.draw_equivalent_wins_for_white = self.draw_eq
        data.game_hash = self.game_hash_fn(g, last["game_index"])
        data.end_finished = not last["hit_move_limit"]
        data.hit_turn_limit = bool(last["hit_move_limit"])
        data.end_no_result = bool(last["no_result"])
        data.moves = [t["move"] for t in gm.turns]
        data.start_moves = list(gm.start_moves)                 # startHist: played before the training period (policy-initialised opening)
        data.start_hist_moves = len(data.start_moves)
        data.ko_rule = ("SIMPLE", "POSITIONAL", "SITUATIONAL", "SPIGHT")[ko_rule]
        data.multi_stone_suicide_legal = bool(multi_suicide)
        data.boards_by_turn = gm.boards + [crop(last["final_colors"])]
        for t in gm.turns:
            data.next_player_by_turn.append(t["next_player"])
            data.packed_input_by_turn.append(t["packed"]); data.global_input_by_turn.append(t["global_input"])
            data.target_weight_by_turn.append(float(t.get("target_weight", 1.0)))
            data.policy_targets_by_turn.append(t["policy_target"])
            data.policy_surprise_by_turn.append(t["surprise"]); data.policy_entropy_by_turn.append(t["policy_entropy"]); data.search_entropy_by_turn.append(t["search_entropy"])
            data.white_value_targets_by_turn.append(t["value_targets"])
            data.white_q_value_targets_by_turn.append(t["q_targets"])
            data.nn_raw_stats_by_turn.append(t["nn_raw_stats"])
        if data.end_no_result:
            area = np.zeros(X * Y, np.uint8)          # "nobody owns anything" (play.cpp:1977-1988)
            data.white_value_targets_by_turn.append(final_value_targets(0, 0.0, self.draw_eq, komi, no_result=True))
        else:
            # area scoring without tax: ownership = full area = calculateArea with every flag on (boardhistory.cpp:591-610)
            area = crop(last["final_area"])
            score = float(last["final_white_minus_black_score"])
            winner = P_WHITE if score > 0 else P_BLACK if score < 0 else 0
            data.winner, data.final_white_minus_black_score = winner, score
            data.white_value_targets_by_turn.append(final_value_targets(winner, score, self.draw_eq, komi))
        data.final_full_area, data.final_ownership = area, area
        if self.policy_surprise_data_weight > 0 or self.value_surprise_data_weight > 0:      # play.cpp:2034-2163
            value_surprise = compute_value_surprise_by_turn(data.white_value_targets_by_turn, [t["raw_nn_values"] for t in gm.turns], X * Y,
                                                            self.use_search_value_surprise)
            data.value_surprise_by_turn = value_surprise
            data.target_weight_by_turn = surprise_target_weights(data.target_weight_by_turn, data.policy_surprise_by_turn, value_surprise,
                                                                 self.policy_surprise_data_weight, self.value_surprise_data_weight)
            data.target_weight_by_turn_unrounded = list(data.target_weight_by_turn)
        if self.weight_rand is not None:                                                      # play.cpp:2274-2289
            if data.target_weight_by_turn_unrounded is None:
                data.target_weight_by_turn_unrounded = list(data.target_weight_by_turn)
            data.target_weight_by_turn = [resolve_target_weight(w, self.weight_rand) for w in data.target_weight_by_turn]
        data.final_white_scoring = scoring_from_area(area)
        data.mode = gm.mode
        data.side_positions = gm.side["list"]            # (jobs still in flight append to this list)
        self.games[g] = _GameInProgress()
        if self.on_game_start is not None:
            self.on_game_start(g)
        if self.lead is not None and self.lead_prob > 0 and not data.end_no_result:
            from .komi_search import compute_lead
            turns = [t for t in range(len(gm.turns)) if float(data.target_weight_by_turn[t]) > 0 and float(data.white_value_targets_by_turn[t][2]) < 0.3 and
                     self.lead_rand.random() < self.lead_prob]
            if turns:
                waiting = {"data": data, "slot": g, "left": len(turns) + gm.side["pending"]}
                gm.side["waiting"] = waiting
                self.games_waiting_for_lead += 1
                for t in turns:
                    self.lead.submit(compute_lead(komi, X, Y), (X, Y, ko_rule, multi_suicide), list(data.start_moves) + list(data.moves[:t]),
                                     lambda lead, t=t, w=waiting: self._lead_done(w, t, lead))
                return
        if gm.side["pending"] > 0:                      # side positions of this game are still being searched
            gm.side["waiting"] = {"data": data, "slot": g, "left": gm.side["pending"]}
            self.games_waiting_for_lead += 1
            return
        self._emit(g, data)

    def _submit_side_position(self, g, last):
        """play.cpp:1846-1860: a forking move from the position just left (not the move played), the resulting position searched off line."""
        gm = self.games[g]
        turn = gm.turns[-1]
        pos = choose_random_forking_move(gm.last_policy, self.X, self.side_rand, last["pos"])
        if pos is None or gm.setup is None:
            return
        n = self.X * self.Y
        mv = (-1, -1) if pos == n else (pos % self.X, pos // self.X)
        moves = list(gm.start_moves) + [t["move"] for t in gm.turns[:-1]] + [mv]
        komi = float(self.sp.komi_values()[0][g]) if hasattr(self.sp, "komi_values") else self.komi
        side, turn_idx, X, Y = gm.side, int(turn["move_num"]) + 1, self.X, self.Y
        side["pending"] += 1

        def job():
            ans = yield {"moves": moves, "komi": komi}
            if ans is None:                       # the forking move ended the game: no side position
                return None
            ex = extract_root_targets(ans["loop"], ans["slot"], X, Y)
            if not ex["row_ok"]:
                return None
            return SidePosition(ex["own"], turn_idx, ex["packed"], ex["global_input"], ex["policy_target"][0], ex["policy_target"][1], ex["value_targets"],
                                ex["q_targets"], ex["surprise"], ex["policy_entropy"], ex["search_entropy"], ex["nn_raw_stats"])

        def done(obj):
            if obj is not None:
                side["list"].append(obj)
            side["pending"] -= 1
            w = side["waiting"]
            if w is not None:
                w["left"] -= 1
                if w["left"] == 0:
                    self.games_waiting_for_lead -= 1
                    self._emit(w["slot"], w["data"])
        self.side_searcher.submit(job(), gm.setup, [], done)

    def start_from(self, g, moves, mode=2):
        """The slot's game that has just begun starts from `moves` (already played into the device slot by the caller): a forked game."""
        self.games[g].preset_moves = [tuple(m) for m in moves]
        self.games[g].mode = mode

    def _lead_done(self, waiting, t, lead):
        data = waiting["data"]
        v = data.white_value_targets_by_turn[t]
        data.white_value_targets_by_turn[t] = (v[0], v[1], v[2], v[3], 1, _f32(lead))           # hasLead, lead (ValueTargets, trainingwrite.h)
        waiting["left"] -= 1
        if waiting["left"] == 0:
            self.games_waiting_for_lead -= 1
            self._emit(waiting["slot"], data)

    def _emit(self, g, data):
        if self.writer is not None:
            self.writer.write_game(data)
        self.games_written += 1
        if self.on_game is not None:
            self.on_game(g, data)
