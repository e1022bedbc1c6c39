"""Forked games: Play::maybeForkGame and the fork pool (program/play.cpp:2356-2508, ForkData program/play.cpp:40-80, GameInitializer's
initialPosition branch :497-526) for the selfplay command.

After a finished game, with probability `earlyForkGameProb` (else `forkGameProb`) a position of it is chosen - early forks an exponentially
distributed number of moves from the start (`earlyForkGameExpectedMoveProp` x board area), late forks uniformly over the game - replayed, a few
random legal moves (`forkGameMinChoices` .. `(early)ForkGameMaxChoices`, with replacement) are each evaluated by the net after being played, and the
one the net scores best for the player to move is made.  The resulting position goes into a pool; a game that starts while the pool is
non-empty starts from a random entry of it (same board size, rules and komi as the forked game, komi noise redrawn around it; with
probability `forkCompensateKomiProb` the komi is first adjusted to even at that position).  The evaluations are position queries on the side
loop (komi_search.KomiSearcher): the root's own evaluation gives the net's whiteScoreMean, the root policy the legal moves."""
import math


class ForkManager:
    def __init__(self, settings, rand):
        """settings: early_fork_game_prob, early_fork_game_expected_move_prop, fork_game_prob, fork_game_min_choices, early_fork_game_max_choices,
        fork_game_max_choices, fork_compensate_komi_prob (PlaySettings / GameInitializer keys in snake case)."""
        self.s, self.rand = dict(settings), rand
        self.pool = []                    # dicts: moves, setup, komi
        self.forks_made = self.forks_used = 0
        if self.enabled and int(self.s.get("fork_game_min_choices", 1)) > max(int(self.s.get("early_fork_game_max_choices", 1)), int(self.s.get("fork_game_max_choices", 1))):
            raise ValueError("fork game max choices < forkGameMinChoices")

    @property
    def enabled(self):
        return float(self.s.get("early_fork_game_prob", 0.0)) > 0 or float(self.s.get("fork_game_prob", 0.0)) > 0

    def job(self, all_moves, setup, komi, x_frame):
        """The generator for KomiSearcher.submit, or None when this game is not forked.  all_moves: the finished game's moves from the empty
        board ((x, y) or (-1, -1)); returns the forked position's moves (or None) through the job's on_done."""
        r, s = self.rand, self.s
        early = r.random() < float(s.get("early_fork_game_prob", 0.0))
        late = (not early) and float(s.get("fork_game_prob", 0.0)) > 0 and r.random() < float(s["fork_game_prob"])
        if not (early or late) or len(all_moves) == 0:
            return None
        x, y = int(setup[0]), int(setup[1])
        if early:
            idx = int(math.floor(r.expovariate(1.0) * float(s.get("early_fork_game_expected_move_prop", 0.0)) * x * y))
        else:
            idx = r.randrange(len(all_moves))
        idx = min(idx, len(all_moves) - 1)                                   # prior to the last move (replayGameUpToMove)
        n_choices = r.randint(int(s.get("fork_game_min_choices", 1)), int(s.get("early_fork_game_max_choices" if early else "fork_game_max_choices", 1)))
        prefix = [tuple(m) for m in all_moves[:idx]]
        black_to_move = idx % 2 == 0

        def gen():
            ans = yield {"moves": prefix, "komi": komi}
            if ans is None:
                return None
            legal = [int(p) for p in range(len(ans["legal"])) if ans["legal"][p]]
            if not legal:
                return None
            n_pos = len(ans["legal"]) - 1
            best, best_score = None, 0.0
            for _ in range(n_choices):                                       # chooseRandomLegalMoves: with replacement, the pass included
                pos = legal[r.randrange(len(legal))]
                mv = (-1, -1) if pos == n_pos else (pos % x_frame, pos // x_frame)
                a = yield {"moves": prefix + [mv], "komi": komi}
                if a is None:
                    continue                                                 # (that move ended the game)
                score = a["nn_score_mean"]
                if best is None or (not black_to_move and score > best_score) or (black_to_move and score < best_score):
                    best, best_score = mv, score
            return None if best is None else prefix + [best]
        return gen()

    def add(self, moves, setup, komi):
        self.pool.append(dict(moves=[tuple(m) for m in moves], setup=tuple(int(v) for v in setup), komi=float(komi)))
        self.forks_made += 1

    def pop(self):
        """ForkData::get: a random entry of the pool (removed), or None."""
        if not self.pool:
            return None
        i = self.rand.randrange(len(self.pool))
        self.pool[i], self.pool[-1] = self.pool[-1], self.pool[i]
        self.forks_used += 1
        return self.pool.pop()
