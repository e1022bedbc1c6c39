"""Komi bisection searches on a side loop: fair komi for a game's empty board (`komiAuto`) and the lead target of recorded turns
(`estimateLeadProb`) - PlayUtils::adjustKomiToEven / computeLead / getNaiveEvenKomiHelper / evalKomi (program/playutils.cpp:372-660) as the
selfplay command uses them (program/play.cpp:1563-1575, 2290-2324).

The reference runs these searches inline on the game's own `Search` objects.  Here they are jobs on a second device loop (`KomiSearcher`: its
own handle, a few slots, the game's search parameters with the root noise off and `numVisits` visits - getNoiselessParams): a job is a Python
generator that yields the komi it wants evaluated for its position and receives (lead, winLoss) of the finished search, exactly the
sequence of evalKomi calls the reference makes (cached per komi like its scoreWLCache).  A slot is loaded with a position by ending whatever
game it holds (two passes: `kgb_selfplay_play_moves_game` restarts the slot with the setup and komi handed over for its next game) and
replaying the position's moves from the empty board, so history-dependent features and superko state are the game's own.  The main loop
never waits: a fair komi is computed for a slot's NEXT game while the current one is played, lead targets are filled in after the game like
the reference does, and a finished game is written once its lead jobs are back."""
import math

import numpy as np

from .game_initializer import round_and_clip_komi


def naive_even_komi(old_komi, x, y):
    """getNaiveEvenKomiHelper (playutils.cpp:455-589) as a generator: yields komis (rounded and clipped), receives (lead, winLoss) - white's
    perspective - and returns the komi at which the position is even.  The caller's komi is `old_komi`; nothing is changed."""
    cache = {}

    def ev(k):
        k = float(round_and_clip_komi(k, x, y))
        if k not in cache:
            cache[k] = yield k
        return cache[k]

    komi = float(old_komi)
    last_shift = last_win_loss = last_lead = 0.0
    for i in range(3):
        lead, win_loss = yield from ev(komi)
        if i > 0 and ((last_lead > 0 and lead > last_lead + 5 and win_loss < 0.75) or (last_lead < 0 and lead < last_lead - 5 and win_loss > -0.75) or
                      (last_win_loss > 0 and win_loss > last_win_loss + 0.1) or (last_win_loss < 0 and win_loss < last_win_loss - 0.1)):
            komi = float(round_and_clip_komi(komi - float(np.float32(last_shift)) * 0.5, x, y))       # the shift made things worse: take half of it back
            break
        last_lead, last_win_loss = lead, win_loss
        shift = -lead
        if i > 0 and abs(shift) > abs(last_shift):
            shift = -abs(last_shift) if shift < 0 else abs(last_shift) if shift > 0 else shift
        last_shift = shift
        if (shift > 0 and win_loss > 0) or (shift < 0 and lead < 0):          # score and win rate pull in opposite directions
            break
        komi = float(round_and_clip_komi(komi + shift, x, y))
        if abs(shift) < 16.0:
            break

    def win_loss_at(delta):
        _, wl = yield from ev(komi + delta)
        return wl
    wl0 = yield from win_loss_at(0.0)
    if wl0 < 0:
        lower, lower_wl = 0.0, wl0
        for i in range(6):
            upper = float(round(2.0 ** i))
            upper_wl = yield from win_loss_at(upper)
            if upper_wl >= 0:
                break
    else:
        upper, upper_wl = 0.0, wl0
        for i in range(6):
            lower = -float(round(2.0 ** i))
            lower_wl = yield from win_loss_at(lower)
            if lower_wl <= 0:
                break
    while upper - lower > 0.50001:
        mid = 0.5 * (lower + upper)
        mid_wl = yield from win_loss_at(mid)
        if mid_wl < 0:
            lower, lower_wl = mid, mid_wl
        else:
            upper, upper_wl = mid, mid_wl
    if lower_wl >= upper_wl - 1e-30:
        final = 0.5 * (lower + upper)
    elif upper_wl <= 0:
        final = upper
    elif lower_wl >= 0:
        final = lower
    else:
        final = lower + (upper - lower) * (0 - lower_wl) / (upper_wl - lower_wl)
    return komi + final, ev


def adjust_komi_to_even(old_komi, x, y, rand):
    """PlayUtils::adjustKomiToEven (playutils.cpp:591-610): the even komi, rounded to a half-integer with linear probability."""
    new_komi, _ = yield from naive_even_komi(old_komi, x, y)
    lower = math.floor(new_komi * 2.0) * 0.5
    upper = lower + 0.5
    new_komi = upper if rand.random() < (new_komi - lower) / (upper - lower) else lower
    return float(round_and_clip_komi(new_komi, x, y))


def compute_lead(old_komi, x, y):
    """PlayUtils::computeLead (playutils.cpp:612-660) under area scoring without button (coarse 2-point granularity): how many points white
    is ahead at `old_komi` = old_komi - the even komi, the even komi smoothed over the granularity."""
    naive, ev = yield from naive_even_komi(old_komi, x, y)
    if naive == round(naive):
        return float(np.float32(old_komi - naive))
    lower = math.floor(naive * 2.0) * 0.5
    upper = lower + 0.5

    def wl(k):
        _, w = yield from ev(k)
        return w
    lower_wl = 0.5 * ((yield from wl(upper)) + (yield from wl(lower - 0.5)))
    upper_wl = 0.5 * ((yield from wl(upper + 0.5)) + (yield from wl(lower)))
    if lower_wl >= upper_wl - 1e-30:
        result = 0.5 * (lower + upper)
    else:
        result = lower + (upper - lower) * (0 - lower_wl) / (upper_wl - lower_wl)
        result = min(max(result, lower - 0.5), upper + 0.5)
    return float(np.float32(old_komi - result))


class KomiSearcher:
    """Runs komi-search jobs on a side `SelfPlay` loop.  sp: a loop in hold mode whose max_visits is the job's numVisits and whose root
    parameters are the noiseless ones (create it with `noiseless_kwargs`).  submit(gen, setup, moves, on_done): gen is one of the generators
    above; setup = (x, y, ko rule, multi-stone suicide); moves = the position's moves from the empty board ((x, y) or (-1, -1) / None for a
    pass, black first)."""

    def __init__(self, sp):
        self.sp, self.n = sp, sp.num_games
        self.free = list(range(self.n))
        self.running = {}                 # slot -> [generator, setup, moves, on_done]
        self.queue = []
        self.setups = np.tile(np.array([sp.x, sp.y, 0, 1], np.int32), (self.n, 1))
        self.komis = np.full(self.n, 7.5, np.float32)
        self.searches = 0

    @staticmethod
    def noiseless_kwargs(kw):
        """getNoiselessParams (playutils.cpp:372-387) on the loop's keyword arguments."""
        out = dict(kw)
        out.update(root_noise_enabled=False, root_policy_temperature=1.0, root_policy_temperature_early=1.0, root_fpu_reduction_max=kw.get("fpu_reduction_max", 0.2),
                   root_fpu_loss_prop=kw.get("fpu_loss_prop", 0.0), root_desired_per_child_visits_coeff=0.0, root_num_symmetries_to_sample=1)
        return out

    def submit(self, gen, setup, moves, on_done):
        self.queue.append([gen, tuple(int(v) for v in setup), list(moves), on_done, None, None])
        self._dispatch()

    def pending(self):
        return len(self.queue) + len(self.running)

    def _load(self, slot, job, query):
        """Slot <- the job's position at the komi it asks for - or, for a query {"moves": ..., "komi": ...}, that position (fork_play.py): end
        the slot's game (its next game takes the setup and komi set here), replay the moves."""
        komi, moves = (query["komi"], query["moves"]) if isinstance(query, dict) else (query, job[2])
        job[4:] = [query, moves]
        self.setups[slot] = job[1]
        self.komis[slot] = komi
        self.sp.set_game_setup(self.setups)
        self.sp.set_komi(self.komis)
        for _ in range(4):                # two passes end a game (three under spight ko; one if the position's last move was a pass)
            self.sp.play_moves_game(slot, [None])
            if self.sp.game(slot)[1]["move_num"] == 0:
                break
        else:
            raise RuntimeError("KomiSearcher: could not end the slot's previous game")
        self.sp.play_moves_game(slot, [None if (m is None or m[0] < 0) else (int(m[0]), int(m[1])) for m in moves])
        self.searches += 1

    def _advance(self, slot, job, answer):
        """Feed `answer` to the job's generator; load its next query into the slot or finish the job."""
        try:
            komi = job[0].send(answer)
        except StopIteration as stop:
            job[3](stop.value)
            del self.running[slot]
            self.free.append(slot)
            return
        self._load(slot, job, komi)

    def _dispatch(self):
        while self.queue and self.free:
            slot, job = self.free.pop(), self.queue.pop(0)
            self.running[slot] = job
            self._advance(slot, job, None)

    def step(self, waves=8):
        """`waves` waves of the side loop; searches that have finished hand their (lead, winLoss) to their jobs.  Returns the jobs in flight."""
        if not self.running and not self.queue:
            return 0
        self.sp.run(waves)
        done = np.asarray(self.sp.root_visits()) >= self.sp.max_visits
        for slot in [s for s in list(self.running) if done[s]]:
            job = self.running[slot]
            _, root = self.sp.root_value_stats(slot)          # winLoss, noResult, scoreMean, scoreMeanSq, lead - white's perspective
            if not isinstance(job[4], dict):
                answer = (float(root[4]), float(root[0]))
            elif self.sp.game(slot)[1]["move_num"] != len(job[5]):
                answer = None                                 # the replay ended the game on the way: no such position
            else:                                             # a position query: also the net's own score of the root and the legal moves
                answer = dict(lead=float(root[4]), win_loss=float(root[0]), nn_score_mean=float(self.sp.root_extra(slot)["root_nn_moments"][2]),
                              legal=np.asarray(self.sp.root_children(slot)[1]) >= 0, loop=self.sp, slot=slot)      # (the slot stays held while the job reads it)
            self._advance(slot, job, answer)
        self._dispatch()
        return self.pending()

    def drain(self, max_steps=100000):
        for _ in range(max_steps):
            if self.step() == 0:
                return
        raise RuntimeError("KomiSearcher: jobs did not finish")
