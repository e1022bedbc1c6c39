"""Games between two nets on the device loops - the engine under `gatekeeper` and `match` (reference:
command/gatekeeper.cpp, command/match.cpp, program/play.cpp:2712-2714 "two nets per game", MatchPairer program/play.cpp:653-790).

Every bot has its own device loop (`SelfPlay` in hold mode on its own handle: own net, own search parameters, own stream) with the same
number of game slots; slot g of every loop holds THE SAME game.  Both loops search every position of it (on the GPU their waves overlap:
one loop's select / backup kernels run while the other's evaluator has the tensor pipes); the move comes from the loop of the bot whose
turn it is - released when its search is finished, the device chooses and plays the move with that bot's temperature rules - and is
mirrored into the other loop with `play_moves_game`, which clears that loop's tree of the position and, when the move ended the game,
starts the slot's next game there too (`kgb_selfplay_play_moves_game`).  A slot's bots swap colours from game to game (the reference's
MatchPairer alternates sides per pairing).  Board size, rules and komi of every game come from a `GameInitializer`, identical for both
loops.  Nothing here touches the evaluator or the search: it is host bookkeeping between waves."""
import numpy as np

from .npz_writer import FinishedGameData, P_BLACK, P_WHITE


def should_resign(win_loss_values, turn_index, board_area, mover_is_black, resign_threshold, resign_consec_turns):
    """The resignation check of Play::runGame after move `turn_index` (0-based) by the mover (play.cpp:1903-1929): not before turn 1 + area / 5; the
    last resignConsecTurns root win/loss values (white's perspective, one per search so far) must each say that THIS player is lost - below
    the (non-positive) threshold for white, above its negative for black."""
    if len(win_loss_values) < resign_consec_turns or turn_index < 1 + board_area // 5:
        return False
    for wl in win_loss_values[len(win_loss_values) - resign_consec_turns:]:
        loser_is_black = False if wl < resign_threshold else True if wl > -resign_threshold else None
        if loser_is_black is None or loser_is_black != mover_is_black:
            return False
    return True


class MatchPlay:
    def __init__(self, loops, names, num_games_total, game_initializer=None, on_game=None, draw_equivalent_wins_for_white=0.5,
                 no_result_utility_for_white=0.0, game_hash_fn=None, allow_resignation=False, resign_threshold=-0.90, resign_consec_turns=5):
        if len(loops) != 2 or len(names) != 2:
            raise ValueError("MatchPlay: exactly two bots")
        a, b = loops
        if a.num_games != b.num_games or (a.x, a.y) != (b.x, b.y):
            raise ValueError("MatchPlay: both loops need the same number of game slots and the same evaluator frame")
        self.loops, self.names, self.n, self.total = loops, list(names), a.num_games, int(num_games_total)
        self.on_game = on_game
        self.draw_eq, self.no_result_utility = float(draw_equivalent_wins_for_white), float(no_result_utility_for_white)
        self.game_hash_fn = game_hash_fn or (lambda slot, index: (((slot + 1) * 0x9E3779B97F4A7C15 + index) & (2 ** 64 - 1), (index * 0xC2B2AE3D27D4EB4F + slot) & (2 ** 64 - 1)))
        self.black_bot = np.array([g % 2 for g in range(self.n)], np.int32)      # which bot plays black in the slot's current game
        self.to_move = self.black_bot.copy()
        self.moves = [[] for _ in range(self.n)]
        # resignation (PlaySettings allowResignation / resignThreshold / resignConsecTurns, play.cpp:1903-1929): the mover's root win/loss values
        self.allow_resignation, self.resign_threshold, self.resign_consec = bool(allow_resignation), float(resign_threshold), int(resign_consec_turns)
        if self.allow_resignation and not self.resign_threshold <= 0:
            raise ValueError("resignThreshold must not be positive")
        self.win_loss = [[] for _ in range(self.n)]
        self.games_started = self.n
        self.live = np.ones(self.n, bool)                                          # slots whose current game counts towards the total
        if self.total > 0 and self.total < self.n:
            self.live[self.total:] = False
            self.games_started = self.total
        self.games_tallied, self.win_points = 0, [0.0, 0.0]
        self.results = []                   # (black bot, white bot, "B+.." / "W+.." / "0" / "Void", moves)
        self.terminated = False
        self.init = game_initializer
        if game_initializer is not None:
            self.setups, self.komis = game_initializer.draw_many(self.n)
            for sp in loops:
                sp.set_game_setup(self.setups, also_current_games=True); sp.set_komi(self.komis, also_current_games=True)
            self.setups, self.komis = game_initializer.draw_many(self.n)
            for sp in loops:
                sp.set_game_setup(self.setups); sp.set_komi(self.komis)
        for sp in loops:
            sp.run(1)

    # ---- what the reference's data-write loop tallies (gatekeeper.cpp:127-196)
    def _tally(self, black_bot, result_kind, winner):
        if result_kind == "noresult":
            white_points = self.draw_eq
        elif winner == P_BLACK:
            white_points = 0.0
        elif winner == P_WHITE:
            white_points = 1.0
        else:
            white_points = 0.5 * self.no_result_utility + 0.5
        self.win_points[black_bot] += 1.0 - white_points
        self.win_points[1 - black_bot] += white_points
        self.games_tallied += 1

    def _restart(self, g):
        """End slot g's game in both loops without a result on the board (resignation): passes until the slot's next game has begun."""
        for sp in self.loops:
            for _ in range(4):
                sp.play_moves_game(g, [None])
                if sp.game(g)[1]["move_num"] == 0:
                    break
            else:
                raise RuntimeError("MatchPlay: could not end the resigned game")

    def _should_resign(self, g, mover_is_black):
        if not self.allow_resignation:
            return False
        x, y = (int(v) for v in self.loops[0].game_setups()[0][g][:2])
        return should_resign(self.win_loss[g], len(self.moves[g]) - 1, x * y, mover_is_black, self.resign_threshold, self.resign_consec)

    def _finish(self, g, last, mover, resigned_black=None):
        sp = self.loops[mover]
        X, Y, ko_rule, multi = (int(v) for v in sp.game_setups()[1][g])          # the finished game's own board and rules
        komi = float(sp.komi_values()[1][g])
        bb = int(self.black_bot[g])
        data = FinishedGameData(X, Y, komi)
        data.game_hash = self.game_hash_fn(g, int(last["game_index"]))
        data.mode = 0
        data.end_finished = not last["hit_move_limit"]
        data.hit_turn_limit = bool(last["hit_move_limit"])
        data.end_no_result = bool(last["no_result"])
        data.moves = list(self.moves[g])
        data.next_player_by_turn = [P_BLACK if i % 2 == 0 else P_WHITE for i in range(len(data.moves))]
        data.ko_rule = ("SIMPLE", "POSITIONAL", "SITUATIONAL", "SPIGHT")[ko_rule]
        data.multi_stone_suicide_legal = bool(multi)
        if resigned_black is not None:                     # BoardHistory::setWinnerByResignation
            winner = P_WHITE if resigned_black else P_BLACK
            data.winner, data.resigned, data.end_finished, data.end_no_result = winner, True, True, False
            kind, text = "scored", ("W+R" if resigned_black else "B+R")
        elif data.end_no_result:
            kind, winner, text = "noresult", 0, "Void"
        else:
            # a game stopped by the move limit is scored as it stands (gatekeeper.cpp:143-146 endAndScoreGameNow): the device has done that
            score = float(last["final_white_minus_black_score"])
            winner = P_WHITE if score > 0 else P_BLACK if score < 0 else 0
            data.winner, data.final_white_minus_black_score = winner, score
            kind, text = "scored", ("W+%g" % score if winner == P_WHITE else "B+%g" % -score if winner == P_BLACK else "0")
            data.end_finished = True
        if self.live[g]:
            self._tally(bb, kind, winner)
            self.results.append((bb, 1 - bb, text, len(data.moves)))
            if self.on_game is not None:
                self.on_game(g, data, self.names[bb], self.names[1 - bb], text)
        # the slot's next game: colours swapped, fresh setup for the game after it
        self.moves[g] = []
        self.win_loss[g] = []
        self.black_bot[g] = 1 - bb
        self.to_move[g] = self.black_bot[g]
        if self.total > 0 and self.games_started >= self.total:
            self.live[g] = False
        else:
            self.live[g] = True
            self.games_started += 1
        if self.init is not None:
            x, y, ko, suicide, k = self.init.draw()
            self.setups[g] = (x, y, ko, suicide); self.komis[g] = k
            for lp in self.loops:
                lp.set_game_setup(self.setups); lp.set_komi(self.komis)

    def pump(self, waves=8):
        """`waves` waves for both loops, then every slot whose bot-to-move has finished its search moves once.  Returns moves made."""
        for sp in self.loops:
            sp.run(waves)
        made = 0
        for b, sp in enumerate(self.loops):
            held = np.asarray(sp.root_visits()) >= sp.max_visits
            mine = held & (self.to_move == b)
            if not mine.any():
                continue
            if self.allow_resignation:          # historicalMctsWinLossValues: the root value of the search the move comes from
                for g in (int(v) for v in np.flatnonzero(mine)):
                    self.win_loss[g].append(float(sp.root_value_stats(g)[1][0]))
            sp.release(mine.astype(np.uint8))
            sp.run(1)
            other = self.loops[1 - b]
            for g in (int(v) for v in np.flatnonzero(mine)):
                last = sp.last_move(g)
                x, y = last["xy"]
                other.play_moves_game(g, [None if x < 0 else (int(x), int(y))])
                self.moves[g].append((int(x), int(y)))
                made += 1
                if last["game_over"]:
                    self._finish(g, last, b)
                elif self._should_resign(g, mover_is_black=(self.black_bot[g] == b)):
                    black = bool(self.black_bot[g] == b)
                    self._restart(g)
                    self._finish(g, dict(last, game_over=True, no_result=False, hit_move_limit=False, final_white_minus_black_score=0.0), b, resigned_black=black)
                else:
                    self.to_move[g] = 1 - b
        return made

    def done(self):
        return self.terminated or (self.total > 0 and self.games_tallied >= self.total)

    def run(self, waves=8, max_pumps=10 ** 9, stop=None):
        """Play until `num_games_total` games are tallied (or stop(self) says so - the gatekeeper's early termination)."""
        pumps = 0
        while not self.done() and pumps < max_pumps:
            self.pump(waves)
            pumps += 1
            if stop is not None and stop(self):
                self.terminated = True
        return self.win_points
