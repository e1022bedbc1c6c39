"""Training-data container of the reference's self-play output (container + schema layer).

The reference writes one .npz per batch of rows (dataio/trainingwrite.cpp:854-886): seven arrays, each a NumPy v1.0 .npy whose
header is exactly 256 bytes in the reference's own compact spelling (dataio/numpywrite.cpp:97-226), stored in the zip under
the bare array name.  This module writes that container byte-compatibly (headers equal the reference's, tests/test_npz_writer.py)
and packs the binary input planes like `packBits` (trainingwrite.cpp:314-334: 8 points per byte, first point in the high bit,
46 bytes per 19x19 plane).

What fills the arrays: `rows_from_root_observations` turns what the device loop exposes for a root position - the fillRowV7 row
and the root's play selection values - into rows with the input planes, the global inputs and policy target 0
(`Play::extractPolicyTarget`, program/play.cpp:810-846: the selection values scaled so that the largest is at least 10, capped at
30000, rounded to int16; `fillPolicyTarget`, trainingwrite.cpp:346-357), every game-dependent target at weight zero.

`TrainingWriteBuffers` mirrors the reference class of that name: `add_row` restates `addRow` (trainingwrite.cpp:448-852) - TD value
targets, lead, time of arrival, weights, history masks, game hash, metadata, score distribution, ownership / future boards /
scoring planes, Q targets - including the stochastic rounding, which draws from the reference's `Rand` in the reference's order
(`RowRand`), so a row is equal to the reference's byte for byte given the same arguments and seed (tests/test_npz_writer.py
against tests/golden/addrow_*.json, dumped from the reference's own addRow).  The input planes and global inputs of a row are
what the device loop's fillRowV7 produced; this module does not recompute them.
"""
import math
import zipfile

import numpy as np

NUM_BIN, NUM_GLOBAL = 22, 19                       # NNInputs::NUM_FEATURES_SPATIAL_V7 / GLOBAL_V7
POLICY_TARGET_CHANNELS, GLOBAL_TARGET_CHANNELS, VALUE_SPATIAL_CHANNELS, QVALUE_CHANNELS = 2, 80, 5, 3    # trainingwrite.cpp:276-279
SCORE_DISTR_RADIUS = 60                            # NNPos::EXTRA_SCORE_DISTR_RADIUS
HEADER_BYTES = 256

# array name -> (descr, trailing shape for a data length L = dataXLen = dataYLen)
def schema(L=19):
    packed = (L * L + 7) // 8
    ps = L * L + 1
    return {
        "binaryInputNCHWPacked": ("|u1", (NUM_BIN, packed)),
        "globalInputNC": ("<f4", (NUM_GLOBAL,)),
        "policyTargetsNCMove": ("<i2", (POLICY_TARGET_CHANNELS, ps)),
        "globalTargetsNC": ("<f4", (GLOBAL_TARGET_CHANNELS,)),
        "scoreDistrN": ("|i1", (2 * (L * L + SCORE_DISTR_RADIUS),)),
        "valueTargetsNCHW": ("|i1", (VALUE_SPATIAL_CHANNELS, L, L)),
        "qValueTargetsNCMove": ("<i2", (QVALUE_CHANNELS, ps)),
    }


def npy_header(descr: str, shape) -> bytes:
    """NumpyBuffer's header: magic, version 1.0, length 246, the dict without spaces, space padding, newline at byte 255."""
    d = "{'descr':'%s','fortran_order':False,'shape':(%s)}" % (descr, ",".join(str(int(x)) for x in shape))
    body = d.encode("ascii")
    if 10 + len(body) >= HEADER_BYTES:
        raise ValueError("numpy header too long")
    return b"\x93NUMPY\x01\x00" + bytes([(HEADER_BYTES - 10) & 0xFF, (HEADER_BYTES - 10) >> 8]) + body + b" " * (HEADER_BYTES - 11 - len(body)) + b"\n"


def pack_bits(planes: np.ndarray) -> np.ndarray:
    """[N, C, L*L] 0/1 -> [N, C, ceil(L*L/8)] uint8, first point in the most significant bit (packBits)."""
    return np.packbits(np.asarray(planes) != 0, axis=2, bitorder="big")


def write_npz(path: str, arrays: dict, L: int = 19, compress: bool = True):
    """Write the seven arrays (all with the same number of rows) the way TrainingWriteBuffers::writeToZipFile does."""
    sch = schema(L)
    n = None
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED if compress else zipfile.ZIP_STORED) as z:
        for name, (descr, rest) in sch.items():
            a = np.ascontiguousarray(arrays[name], dtype=np.dtype(descr))
            if a.shape[1:] != tuple(rest):
                raise ValueError(f"{name}: shape {a.shape} does not end in {rest}")
            n = a.shape[0] if n is None else n
            if a.shape[0] != n:
                raise ValueError(f"{name}: {a.shape[0]} rows, expected {n}")
            z.writestr(name, npy_header(descr, a.shape) + a.tobytes())
    return n


def policy_target_from_play_selection(values):
    """Play::extractPolicyTarget (program/play.cpp:810-846) on Search::getPlaySelectionValues by move position (-1 = no child):
    scaleMaxToAtLeast = 10, cap at 30000, round to nearest (C `round`: halves away from zero), int16."""
    v = np.where(np.asarray(values, np.float64) > 0, np.asarray(values, np.float64), 0.0)
    mx = v.max(axis=-1, keepdims=True)
    v = np.where((mx > 0) & (mx < 10.0), v * (10.0 / np.maximum(mx, 1e-300)), v)
    mx = v.max(axis=-1, keepdims=True)
    v = np.where(mx > 30000.0, v * (30000.0 / np.maximum(mx, 1e-300)), v)
    return np.floor(v + 0.5).astype(np.int16)


def rows_from_root_observations(spatial_nhwc, global_in, play_selection_values, L: int = 19, target_weight: float = 1.0, turn_idx=None,
                                num_visits=None):
    """spatial_nhwc [N, L*L, 22] and global_in [N, 19] as kgb_selfplay_get_nn_row gives them for a root, play_selection_values
    [N, L*L+1] from kgb_selfplay_get_play_selection_values (-1 = no child).  Everything that needs the finished game carries zero weight."""
    sp = np.asarray(spatial_nhwc, np.float32)
    n = sp.shape[0]
    sch = schema(L)
    out = {k: np.zeros((n,) + tuple(rest), np.dtype(descr)) for k, (descr, rest) in sch.items()}
    out["binaryInputNCHWPacked"] = pack_bits(np.transpose(sp, (0, 2, 1)))
    out["globalInputNC"] = np.asarray(global_in, np.float32)
    out["policyTargetsNCMove"][:, 0, :] = policy_target_from_play_selection(play_selection_values)
    out["policyTargetsNCMove"][:, 1, :] = 1                     # uniformPolicyTarget: no next-move target (weight C28 = 0)
    gt = out["globalTargetsNC"]
    gt[:, 24] = 1.0; gt[:, 35] = 1.0          # 1 - weight of the td value targets / of the value targets
    gt[:, 25] = target_weight                 # weight of the row
    gt[:, 26] = 1.0                           # policy target present
    gt[:, 36:41] = 1.0                        # history masks: use all five previous moves
    gt[:, 48] = 1.0                           # area scoring
    gt[:, 63] = 3.0                           # data format version
    if turn_idx is not None:
        gt[:, 51] = np.asarray(turn_idx, np.float32)
    if num_visits is not None:
        gt[:, 60] = np.asarray(num_visits, np.float32)
    return out


class RowRand:
    """The reference's `Rand` (core/rand.h) as far as addRow uses it: nextDouble (:245-257, 53 bits of nextUInt64 = two
    nextUInt, low word first) and nextBool (:271-274).  The uint32 stream comes from the library's host Rand twin."""

    def __init__(self, seed_string: str, prefetch: int = 4096):
        self._seed, self._n, self._i = seed_string, 0, 0
        self._buf = np.zeros(0, np.uint32)
        self._grow(prefetch)

    def _grow(self, n):
        from .nn_backend import rand_uint32_stream
        self._buf = rand_uint32_stream(self._seed, n)     # a longer stream of the same seed has the old one as its prefix
        self._n = n

    def next_uint(self) -> int:
        if self._i >= self._n:
            self._grow(self._n * 2)
        v = int(self._buf[self._i])
        self._i += 1
        return v

    def next_double(self) -> float:
        lo = self.next_uint()
        hi = self.next_uint()
        return float((lo | (hi << 32)) & ((1 << 53) - 1)) / float(1 << 53)

    def next_bool(self, prob: float) -> bool:
        return self.next_double() < prob


_f32 = np.float32


def _c_round(x) -> int:
    """C `round`: halves away from zero."""
    x = float(x)
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def _clamp_to_radius(x, radius: int, rand: RowRand) -> int:
    """clampToRadius120 / clampToRadius32000 (trainingwrite.cpp:358-383): x (float32) to an integer whose expectation is x."""
    x = _f32(x)
    low = int(math.floor(float(x)))
    high = low + 1
    if low < -radius:
        return -radius
    if high > radius:
        return radius
    lam = _f32(x - _f32(low))
    if lam == 0.0:
        return low
    return high if rand.next_bool(float(lam)) else low


def _value_td_targets(white_value_targets, idx, white_to_move, now_factor):
    """fillValueTDTargets (trainingwrite.cpp:411-446): exponentially weighted average of the value targets from this turn on."""
    win = loss = no_result = score = 0.0
    weight_left = 1.0
    n = len(white_value_targets)
    for i in range(idx, n):
        if i == n - 1:
            weight_now, weight_left = weight_left, 0.0
        else:
            weight_now = weight_left * now_factor
            weight_left *= (1.0 - now_factor)
        t = white_value_targets[i]
        w, l, nr, sc = float(_f32(t[0])), float(_f32(t[1])), float(_f32(t[2])), float(_f32(t[3]))
        win += weight_now * (w if white_to_move else l)
        loss += weight_now * (l if white_to_move else w)
        no_result += weight_now * nr
        score += weight_now * (sc if white_to_move else -sc)
    cap = 19 * 19 + SCORE_DISTR_RADIUS          # NNPos::MAX_BOARD_AREA + EXTRA_SCORE_DISTR_RADIUS
    score = min(max(score, -cap), cap)
    return [_f32(win), _f32(loss), _f32(no_result), _f32(score)]


P_BLACK, P_WHITE = 1, 2


class TrainingWriteBuffers:
    """Row buffers of one output file, like the reference's class (dataio/trainingwrite.h:245-352).  Boards are x_size * y_size
    inside a data_len * data_len frame (pos = y * data_len + x, pass = data_len^2).  Colours: 0 empty, 1 black, 2 white."""

    def __init__(self, max_rows: int, data_len: int = 19):
        self.L, self.max_rows, self.cur_rows = data_len, max_rows, 0
        self.arrays = {k: np.zeros((max_rows,) + tuple(rest), np.dtype(descr)) for k, (descr, rest) in schema(data_len).items()}

    def _pos(self, x, y):
        return self.L * self.L if x < 0 else y * self.L + x     # NNPos::locToPos

    def add_row(self, *, x_size, y_size, next_player, packed_input, global_input, turn_idx, target_weight, unreduced_num_visits,
                policy_target0, policy_target1, policy_surprise, policy_entropy, search_entropy, white_value_targets, white_q_value_targets,
                white_value_targets_idx, value_target_weight, td_value_target_weight, lead_target_weight_factor, nn_raw_stats,
                final_full_area, final_ownership, final_white_scoring, pos_hist_for_future_boards, is_side_position,
                num_neural_nets_behind_latest, game_hash, num_changed_neural_nets, hit_turn_limit, num_extra_black, mode, rand: RowRand,
                self_komi, area_scoring_or_encore2=True, start_hist_moves=0, initial_turn_number=0, white_bonus_now=0.0, white_bonus_end=0.0,
                end_finished=True, end_no_result=False, always_pass_alive_under_suicide_rules=False, reanalysis=(False, 0.0, 0.0, 0)):
            """One row.  Arguments follow addRow's (trainingwrite.cpp:448-485); what addRow reads from its three BoardHistory
            arguments is passed as scalars (self_komi = hist.currentSelfKomi(nextPlayer, drawEquivalentWinsForWhite), ...).
            policy targets: list of (x, y, value) with x < 0 for pass, or None.  white_value_targets: per turn
            (win, loss, noResult, score, hasLead, lead).  white_q_value_targets: list of (x, y, winLoss, score, visits).
            final_* planes: row-major [y_size * x_size] or None.  pos_hist_for_future_boards: one board per value target, or None."""
            if self.cur_rows >= self.max_rows:
                raise ValueError("TrainingWriteBuffers full")
            L, A, r = self.L, self.L * self.L, self.cur_rows
            P = A + 1
            white = next_player == P_WHITE
            opp = P_BLACK if white else P_WHITE
            self.arrays["binaryInputNCHWPacked"][r] = packed_input
            self.arrays["globalInputNC"][r] = global_input
            g = self.arrays["globalTargetsNC"][r]
            g[:] = 0
            g[25] = target_weight
            pol = self.arrays["policyTargetsNCMove"][r]
            for ch, wt_col, target in ((0, 26, policy_target0), (1, 28, policy_target1)):
                if target is None:
                    pol[ch, :] = 1                       # uniformPolicyTarget, weight 0
                    g[wt_col] = 0.0
                else:
                    pol[ch, :] = 0
                    for (x, y, v) in target:
                        pol[ch, self._pos(x, y)] = v
                    g[wt_col] = 1.0
            board_area = x_size * y_size
            idx = white_value_targets_idx
            for k, now_factor in enumerate((0.0, 1.0 / (1.0 + board_area * 0.176), 1.0 / (1.0 + board_area * 0.056), 1.0 / (1.0 + board_area * 0.016), 1.0)):
                g[4 * k:4 * k + 4] = _value_td_targets(white_value_targets, idx, white, now_factor)
            vtw, tdw = _f32(value_target_weight), _f32(td_value_target_weight)
            no_result_end = bool(end_finished) and bool(end_no_result)
            this = white_value_targets[idx]
            if this[4] and not no_result_end:
                lead = _f32(this[5]) if white else -_f32(this[5])
                cap = _f32(19 * 19 + SCORE_DISTR_RADIUS)
                g[21] = min(max(lead, -cap), cap)
                g[29] = vtw * _f32(lead_target_weight_factor)
            s = 0.0
            for i in range(idx + 1, len(white_value_targets)):
                prev, cur = white_value_targets[i - 1], white_value_targets[i]
                prev_wl = float(_f32(prev[0]) - _f32(prev[1]))
                next_wl = float(_f32(cur[0]) - _f32(cur[1]))
                s += (i - idx) * ((next_wl - prev_wl) * (next_wl - prev_wl))
            g[22] = s
            g[24] = _f32(1.0) - tdw
            g[30], g[31], g[32] = policy_surprise, policy_entropy, search_entropy
            g[35] = _f32(1.0) - vtw
            use = True
            for k in range(5):                           # each earlier history step is kept with probability 0.98 (:628-637)
                use = use and rand.next_double() < 0.98
                g[36 + k] = 1.0 if use else 0.0
            h0, h1 = int(game_hash[0]), int(game_hash[1])
            g[41], g[42], g[43] = h0 & 0x3FFFFF, (h0 >> 22) & 0x3FFFFF, (h0 >> 44) & 0xFFFFF
            g[44], g[45], g[46] = h1 & 0x3FFFFF, (h1 >> 22) & 0x3FFFFF, (h1 >> 44) & 0xFFFFF
            g[47] = self_komi
            g[48] = 1.0 if area_scoring_or_encore2 else 0.0
            g[49] = 1.0 if num_changed_neural_nets > 0 else 0.0
            g[50] = num_neural_nets_behind_latest
            g[51] = turn_idx
            g[52] = 1.0 if hit_turn_limit else 0.0
            g[53] = start_hist_moves
            g[54] = num_extra_black
            g[55] = mode
            g[56] = initial_turn_number
            g[57] = nn_raw_stats[0] if white else -nn_raw_stats[0]
            g[58] = nn_raw_stats[1] if white else -nn_raw_stats[1]
            g[59] = nn_raw_stats[2]
            g[60] = unreduced_num_visits
            if not is_side_position:
                wb = _f32(white_bonus_end) - _f32(white_bonus_now)
                sb = wb if white else -wb
                g[61] = sb if sb != 0 else 0.0
            g[62] = 1.0 if (not is_side_position and end_finished and not hit_turn_limit) else 0.0
            g[63] = 3.0
            if reanalysis[0]:
                g[64], g[65], g[66], g[67] = 1.0, reanalysis[1], reanalysis[2], float(reanalysis[3])
            g[68] = 1.0 if always_pass_alive_under_suicide_rules else 0.0      # trainingwrite.cpp:702

            sd = self.arrays["scoreDistrN"][r]
            own = self.arrays["valueTargetsNCHW"][r].reshape(VALUE_SPATIAL_CHANNELS, A)
            sd[:] = 0
            own[:] = 0
            sd_len, sd_mid = 2 * A + 2 * SCORE_DISTR_RADIUS, A + SCORE_DISTR_RADIUS
            frame = (np.arange(y_size)[:, None] * L + np.arange(x_size)[None, :]).reshape(-1)    # NNPos::xyToPos of the board's points
            if final_ownership is None or no_result_end:
                sd[sd_mid - 1] = 50
                sd[sd_mid] = 50
            else:
                g[27] = vtw
                last = white_value_targets[-1]
                score = _f32(last[3]) if white else -_f32(last[3])
                g[20] = score
                fo, fa = np.asarray(final_ownership).reshape(-1), np.asarray(final_full_area).reshape(-1)
                own[0, frame] = np.where(fo == next_player, 1, np.where(fo == opp, -1, 0))
                own[1, frame] = np.where((fa != 0) & (fo == 0), np.where(fa == next_player, 1, -1), 0)
                center = _c_round(score)
                lower, upper = center + sd_mid - 1, center + sd_mid
                if upper <= 0:
                    sd[0] = 100
                elif lower >= sd_len - 1:
                    sd[sd_len - 1] = 100
                else:
                    lam = _f32(score - _f32(_f32(center) - _f32(0.5)))
                    up = _c_round(_f32(lam * _f32(100.0)))
                    sd[lower] = 100 - up
                    sd[upper] = up
            if pos_hist_for_future_boards is not None:
                boards = pos_hist_for_future_boards
                if len(boards) != len(white_value_targets):
                    raise ValueError("pos_hist_for_future_boards must hold one board per value target")
                g[33] = 1.0
                end = len(boards) - 1
                for ch, ahead in ((2, 8), (3, 32)):
                    b = np.asarray(boards[min(idx + ahead, end)]).reshape(-1)
                    own[ch, frame] = np.where(b == next_player, 1, np.where(b == opp, -1, 0))
            if final_white_scoring is not None and not no_result_end:
                g[34] = vtw
                sc = np.asarray(final_white_scoring, np.float32).reshape(-1)
                for j in range(board_area):               # y, x order: the order the reference draws its random numbers in
                    v = sc[j] if white else -sc[j]
                    own[4, frame[j]] = _clamp_to_radius(_f32(v * _f32(120.0)), 120, rand)
            q = self.arrays["qValueTargetsNCMove"][r]
            q[:] = 0
            cap = _f32(19 * 19 + SCORE_DISTR_RADIUS)
            for (x, y, wl, sc, visits) in white_q_value_targets:          # fillQValueTarget (:385-409)
                pos = self._pos(x, y)
                wl = _f32(wl) if white else -_f32(wl)
                sc = _f32(sc) if white else -_f32(sc)
                sc = min(max(sc, -cap), cap)
                q[0, pos] = _clamp_to_radius(_f32(wl * _f32(32000.0)), 32000, rand)
                q[1, pos] = _clamp_to_radius(_f32(sc * _f32(60.0)), 32000, rand)
                q[2, pos] = max(0, min(int(visits), 32000))
            self.cur_rows += 1

    def write_to_zip_file(self, path: str, compress: bool = True):
        """writeToZipFile (trainingwrite.cpp:854-886): the first cur_rows rows of every array."""
        return write_npz(path, {k: v[:self.cur_rows] for k, v in self.arrays.items()}, self.L, compress)


# ---- a finished game and the writer that turns it into rows ---------------------------------------------------------------------------

class FinishedGameData:
    """What `Play::runGame` hands to the writer (dataio/trainingwrite.h:84-170), for the rule subset of the device loop (area
    scoring, no tax, no button, no handicap bonus).  Per-turn lists have one entry per move of the training period; the
    value targets have one more (the game outcome).  `packed_input_by_turn` / `global_input_by_turn` are the fillRowV7 rows
    of the root positions as the device loop produced them when it searched them."""

    def __init__(self, x_size, y_size, komi, start_pla=P_BLACK):
        self.x_size, self.y_size, self.komi, self.start_pla = x_size, y_size, float(komi), start_pla
        self.game_hash = (0, 0)
        self.draw_equivalent_wins_for_white = 0.5
        self.hit_turn_limit = False
        self.num_extra_black = 0
        self.mode = 0
        self.training_weight = 1.0
        self.start_hist_moves = 0                  # moves before the training period (startHist.moveHistory.size())
        self.initial_turn_number = 0
        self.always_pass_alive_under_suicide_rules = False
        self.end_finished, self.end_no_result = True, False
        self.boards_by_turn = []                   # nTurns + 1 boards (colours, row-major): the position before each move and the final one
        self.next_player_by_turn = []
        self.packed_input_by_turn, self.global_input_by_turn = [], []
        self.target_weight_by_turn = []
        self.policy_targets_by_turn = []           # (list of (x, y, value), unreducedNumVisits)
        self.policy_surprise_by_turn, self.policy_entropy_by_turn, self.search_entropy_by_turn = [], [], []
        self.white_value_targets_by_turn = []
        self.white_q_value_targets_by_turn = []
        self.nn_raw_stats_by_turn = []
        self.reanalysis_by_turn = []               # empty, or per turn (wasReanalyzed, usedOutcomeTargets, polSurprise, valSurprise, origVisits, netChangesSoFar)
        self.changed_neural_net_turns = []         # turn index at which each new net took over
        self.side_positions = []                   # SidePosition objects
        self.moves = []                            # (x, y) per turn, (-1, -1) = pass
        self.start_moves = []                      # (x, y) of the moves before the training period (startHist.moveHistory), black first
        self.ko_rule, self.multi_stone_suicide_legal = "SIMPLE", True
        self.winner, self.final_white_minus_black_score = 0, 0.0      # winner: 0 draw, P_BLACK, P_WHITE (finished games with a result)
        self.changed_neural_net_names = None       # names of the nets in changed_neural_net_turns (SGF comment only)
        self.target_weight_by_turn_unrounded = None
        self.value_surprise_by_turn = None         # statistics only (FinishedGameData::valueSurpriseByTurn)
        self.final_full_area = self.final_ownership = self.final_white_scoring = None

    def self_komi(self, next_player):
        """BoardHistory::currentSelfKomi (game/boardhistory.cpp:570-589) without bonus points: komi plus the draw adjustment
        when results are integers."""
        komi_is_int = float(int(self.komi)) == self.komi
        adj = _f32(self.draw_equivalent_wins_for_white - 0.5) if komi_is_int else _f32(0.0)
        w = _f32(_f32(self.komi) + adj)
        return w if next_player == P_WHITE else -w


class SidePosition:
    """A position off the main line that was searched on its own (dataio/trainingwrite.h:62-84): it gets rows with its own
    policy / value / Q targets and none of the targets that need the game's continuation."""

    def __init__(self, next_player, turn_idx, packed_input, global_input, policy_target, unreduced_num_visits, white_value_targets,
                 white_q_value_targets, policy_surprise, policy_entropy, search_entropy, nn_raw_stats, target_weight=1.0, num_neural_net_changes_so_far=0):
        self.next_player, self.turn_idx, self.packed_input, self.global_input = next_player, turn_idx, packed_input, global_input
        self.policy_target, self.unreduced_num_visits = policy_target, unreduced_num_visits
        self.white_value_targets, self.white_q_value_targets = white_value_targets, white_q_value_targets
        self.policy_surprise, self.policy_entropy, self.search_entropy, self.nn_raw_stats = policy_surprise, policy_entropy, search_entropy, nn_raw_stats
        self.target_weight, self.num_neural_net_changes_so_far = target_weight, num_neural_net_changes_so_far


def final_value_targets(winner, final_white_minus_black_score, draw_equivalent_wins_for_white, komi, no_result=False):
    """The outcome entry of the value targets (program/play.cpp:1977-2000): win / loss from the winner (a draw counts as
    drawEquivalentWinsForWhite), the draw-adjusted score, lead = score."""
    if no_result:
        return (0.0, 0.0, 1.0, 0.0, 0, 0.0)
    win = _f32(1.0 if winner == P_WHITE else 0.0 if winner == P_BLACK else draw_equivalent_wins_for_white)
    komi_is_int = float(int(komi)) == float(komi)
    adj = float(_f32(draw_equivalent_wins_for_white - 0.5)) if komi_is_int else 0.0     # whiteKomiAdjustmentForDraws returns float
    score = _f32(float(final_white_minus_black_score) + adj)
    return (win, _f32(1.0) - win, 0.0, score, 1, score)


def scoring_from_area(area):
    """NNInputs::fillScoring without group tax (neuralnet/nninputs.cpp:204-226): white area +1, black area -1."""
    a = np.asarray(area)
    return np.where(a == P_WHITE, 1.0, np.where(a == P_BLACK, -1.0, 0.0)).astype(np.float32)


class TrainingDataWriter:
    """The reference's writer (dataio/trainingwrite.cpp:987-1325): rows of finished games go into a TrainingWriteBuffers that is
    written out as <16 hex digits>.npz whenever it is full; the first file of a writer is cut short at random so that writers
    started together do not all flush at once.  One Rand (seeded by `rand_seed`) serves the first-file size, the fractional
    target weights, addRow's rounding and the file names, in the reference's order."""

    def __init__(self, output_dir, max_rows_per_file, first_file_min_rand_prop, data_len, rand_seed, on_flush=None):
        if not (0.0 <= first_file_min_rand_prop <= 1.0):
            raise ValueError("first_file_min_rand_prop not in [0,1]")
        self.output_dir, self.L = output_dir, data_len
        self.rand = RowRand(rand_seed)
        self.buffers = TrainingWriteBuffers(max_rows_per_file, data_len)
        self.is_first_file = True
        if first_file_min_rand_prop >= 1.0:
            self.first_file_max_rows = max_rows_per_file
        else:
            self.first_file_max_rows = max_rows_per_file - int(max_rows_per_file * (1.0 - first_file_min_rand_prop) * self.rand.next_double())
        self.row_count = 0
        self.on_flush = on_flush                   # tests: called with the buffers instead of writing a file

    def is_empty(self):
        return self.buffers.cur_rows <= 0

    def flush_if_nonempty(self):
        if self.buffers.cur_rows <= 0:
            return None
        self.is_first_file = False
        if self.on_flush is not None:
            self.on_flush(self.buffers)
            name = ""
        else:
            lo = self.rand.next_uint()
            hi = self.rand.next_uint()
            name = "%s/%016X.npz" % (self.output_dir, lo | (hi << 32))
            import os
            self.buffers.write_to_zip_file(name + ".tmp")
            os.replace(name + ".tmp", name)
        self.buffers.cur_rows = 0
        return name

    def _write_and_clear_if_full(self):
        b = self.buffers
        if b.cur_rows >= b.max_rows or (self.is_first_file and b.cur_rows >= self.first_file_max_rows):
            self.flush_if_nonempty()

    def write_game(self, data: FinishedGameData):
        """writeGame (:1097-1325): a turn with target weight w gives floor(w) rows plus one more with probability frac(w); policy
        target 1 is the next turn's policy target; reanalysed turns may drop the outcome-derived targets; then the side positions."""
        n = len(data.target_weight_by_turn)
        if not (len(data.policy_targets_by_turn) == len(data.white_q_value_targets_by_turn) == len(data.nn_raw_stats_by_turn) == n
                and len(data.white_value_targets_by_turn) == n + 1 and len(data.boards_by_turn) == n + 1
                and len(data.reanalysis_by_turn) in (0, n)):
            raise ValueError("FinishedGameData: per-turn lists disagree in length")
        if not data.end_finished and not data.hit_turn_limit:
            raise ValueError("FinishedGameData: unfinished game that did not hit the turn limit")
        for t in range(n):
            target_weight = float(_f32(data.target_weight_by_turn[t]))
            turn_idx = t + data.start_hist_moves
            policy0, unreduced = data.policy_targets_by_turn[t]
            policy1 = data.policy_targets_by_turn[t + 1][0] if t + 1 < n else None
            re = data.reanalysis_by_turn[t] if t < len(data.reanalysis_by_turn) else (False, True, 0.0, 0.0, 0, 0)
            nets = data.changed_neural_net_turns
            behind = 0
            if re[0]:
                behind = len(nets) - re[5]
            else:
                for i, net_turn in enumerate(nets):
                    if net_turn > turn_idx:
                        behind = len(nets) - i
                        break
            skip_outcome = bool(re[0]) and not bool(re[1])
            pla = data.next_player_by_turn[t]
            while target_weight > 0.0:
                if target_weight >= 1.0 or self.rand.next_bool(target_weight):
                    self.buffers.add_row(
                        x_size=data.x_size, y_size=data.y_size, next_player=pla,
                        packed_input=data.packed_input_by_turn[t], global_input=data.global_input_by_turn[t],
                        turn_idx=turn_idx, target_weight=_f32(data.training_weight), unreduced_num_visits=unreduced,
                        policy_target0=policy0, policy_target1=None if skip_outcome else policy1,
                        policy_surprise=data.policy_surprise_by_turn[t], policy_entropy=data.policy_entropy_by_turn[t],
                        search_entropy=data.search_entropy_by_turn[t], white_value_targets=data.white_value_targets_by_turn,
                        white_q_value_targets=data.white_q_value_targets_by_turn[t], white_value_targets_idx=t,
                        value_target_weight=1.0, td_value_target_weight=1.0, lead_target_weight_factor=1.0,
                        nn_raw_stats=data.nn_raw_stats_by_turn[t],
                        final_full_area=None if skip_outcome else data.final_full_area,
                        final_ownership=None if skip_outcome else data.final_ownership,
                        final_white_scoring=None if skip_outcome else data.final_white_scoring,
                        pos_hist_for_future_boards=None if skip_outcome else data.boards_by_turn,
                        is_side_position=False, num_neural_nets_behind_latest=behind, game_hash=data.game_hash,
                        num_changed_neural_nets=len(nets), hit_turn_limit=data.hit_turn_limit, num_extra_black=data.num_extra_black,
                        mode=data.mode, rand=self.rand, self_komi=data.self_komi(pla), area_scoring_or_encore2=True,
                        start_hist_moves=data.start_hist_moves, initial_turn_number=data.initial_turn_number,
                        end_finished=data.end_finished, end_no_result=data.end_no_result,
                        always_pass_alive_under_suicide_rules=data.always_pass_alive_under_suicide_rules,
                        reanalysis=(bool(re[0]), re[2], re[3], re[4]))
                    self._write_and_clear_if_full()
                    self.row_count += 1
                target_weight -= 1.0
        # side rows (:1258-1323): their own targets only; the game's ending still decides the lead / finished flags
        for sp in data.side_positions:
            target_weight = float(_f32(sp.target_weight))
            while target_weight > 0.0:
                if target_weight >= 1.0 or self.rand.next_bool(target_weight):
                    self.buffers.add_row(
                        x_size=data.x_size, y_size=data.y_size, next_player=sp.next_player, packed_input=sp.packed_input, global_input=sp.global_input,
                        turn_idx=sp.turn_idx, target_weight=_f32(data.training_weight), unreduced_num_visits=sp.unreduced_num_visits,
                        policy_target0=sp.policy_target, policy_target1=None, policy_surprise=sp.policy_surprise, policy_entropy=sp.policy_entropy,
                        search_entropy=sp.search_entropy, white_value_targets=[sp.white_value_targets], white_q_value_targets=sp.white_q_value_targets,
                        white_value_targets_idx=0, value_target_weight=1.0, td_value_target_weight=1.0, lead_target_weight_factor=1.0,
                        nn_raw_stats=sp.nn_raw_stats, final_full_area=None, final_ownership=None, final_white_scoring=None,
                        pos_hist_for_future_boards=None, is_side_position=True,
                        num_neural_nets_behind_latest=len(data.changed_neural_net_turns) - sp.num_neural_net_changes_so_far, game_hash=data.game_hash,
                        num_changed_neural_nets=len(data.changed_neural_net_turns), hit_turn_limit=data.hit_turn_limit, num_extra_black=data.num_extra_black,
                        mode=data.mode, rand=self.rand, self_komi=data.self_komi(sp.next_player), area_scoring_or_encore2=True,
                        start_hist_moves=data.start_hist_moves, initial_turn_number=data.initial_turn_number,
                        end_finished=data.end_finished, end_no_result=data.end_no_result,
                        always_pass_alive_under_suicide_rules=data.always_pass_alive_under_suicide_rules)
                    self._write_and_clear_if_full()
                    self.row_count += 1
                target_weight -= 1.0


_SGF_CHARS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
_GTYPES = ("normal", "cleanuptraining", "fork", "handicap", "sgfpos", "hintpos", "hintfork", "asym")


def write_sgf(data: FinishedGameData, b_name: str, w_name: str) -> str:
    """The game record the reference's self-play writes next to its rows (WriteSgf::writeSgf with a FinishedGameData,
    dataio/sgf.cpp:1997-2226, as called from program/selfplaymanager.cpp:377): root properties, the game comment
    (startTurnIdx, initTurnNum, gameHash, gtype, net changes) and per move the value targets, visits and target weight.
    For games that start from the empty board under the rule subset of the loop (no handicap, no encore)."""
    g = lambda v: "%g" % float(_f32(v))           # ostream << float / Global::doubleToString
    out = ["(;FF[4]GM[1]"]
    out.append("SZ[%d]" % data.x_size if data.x_size == data.y_size else "SZ[%d:%d]" % (data.x_size, data.y_size))
    out.append("PB[%s]PW[%s]HA[0]KM[%s]" % (b_name, w_name, g(data.komi)))
    out.append("RU[ko%sscoreAREAtaxNONEsui%d]" % (data.ko_rule, 1 if data.multi_stone_suicide_legal else 0))
    result = ""
    if data.end_finished:
        if data.end_no_result:
            result = "Void"
        elif getattr(data, "resigned", False):               # BoardHistory::isResignation (WriteSgf::printGameResult)
            result = "B+R" if data.winner == P_BLACK else "W+R"
        elif data.winner == P_BLACK:
            result = "B+" + g(-float(_f32(data.final_white_minus_black_score)))
        elif data.winner == P_WHITE:
            result = "W+" + g(data.final_white_minus_black_score)
        else:
            result = "0"
        out.append("RE[%s]" % result)
    mode = _GTYPES[data.mode] if 0 <= data.mode < len(_GTYPES) else "other"
    comment = "startTurnIdx=%d,initTurnNum=%d,gameHash=%016X%016X,gtype=%s" % (
        data.start_hist_moves, data.initial_turn_number, int(data.game_hash[1]), int(data.game_hash[0]), mode)
    for j, turn in enumerate(data.changed_neural_net_turns):
        name = data.changed_neural_net_names[j] if data.changed_neural_net_names else "net%d" % j
        comment += ",newNeuralNetTurn%d=%s" % (turn, name)
    out.append("C[%s]" % comment)
    weights = data.target_weight_by_turn_unrounded if data.target_weight_by_turn_unrounded is not None else data.target_weight_by_turn
    n = len(data.moves)
    for j, (x, y) in enumerate(getattr(data, "start_moves", [])):       # endHist.moveHistory starts with startHist's moves: no comments on those
        out.append(";%s[%s]" % ("B" if j % 2 == 0 else "W", "" if x < 0 else _SGF_CHARS[x] + _SGF_CHARS[y]))       # (no handicap: black moves first)
    for i, (x, y) in enumerate(data.moves):
        pla = data.next_player_by_turn[i]
        out.append(";%s[%s]" % ("B" if pla == P_BLACK else "W", "" if x < 0 else _SGF_CHARS[x] + _SGF_CHARS[y]))
        parts = []
        if i < len(data.white_value_targets_by_turn):
            t = data.white_value_targets_by_turn[i]
            parts.append("%.2f %.2f %.2f %.1f" % tuple(float(_f32(v)) for v in t[:4]))
        if i < len(data.policy_targets_by_turn):
            re = data.reanalysis_by_turn[i] if i < len(data.reanalysis_by_turn) else None
            if re is not None and re[0]:
                parts.append("v=%d rv=%d" % (int(re[4]), int(data.policy_targets_by_turn[i][1])))
            else:
                parts.append("v=%d" % int(data.policy_targets_by_turn[i][1]))
        if i < len(weights):
            parts.append("weight=%.2f" % float(_f32(weights[i])))
        if data.end_finished and i + 1 == n:
            parts.append("result=" + result)
        if parts:
            out.append("C[%s]" % " ".join(parts))
    out.append(")")
    return "".join(out)
