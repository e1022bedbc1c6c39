"""Per-game board size, rules and komi: what the reference's GameInitializer draws when a game is created.

Restates, for the options the device loop has, `GameInitializer::initShared` / `createRulesUnsynchronized` /
`createGameSharedUnsynchronized` (program/play.cpp:83-214, 470-482, 530, 596-608) and `PlayUtils::chooseExtraBlackAndKomi` /
`setKomiWithNoise` / `roundAndClipKomi` (program/playutils.cpp:24-106, 363-371):

  * board size: one of `bSizes` (edge lengths) with `bSizeRelProbs`; with `allowRectangleProb` > 0 every ordered pair of edges,
    weighted like play.cpp:152-172;
  * ko rule and multi-stone suicide: uniform over `koRules` / `multiStoneSuicideLegals`;
  * komi: `komiMean` plus truncated Gaussian noise (`komiStdev`, or with probability `komiBigStdevProb` `komiBigStdev`, or with
    `komiBiggerStdevProb` `komiBiggerStdev`), the noise scaled by sqrt(board area) / 19, rounded to a half-integer with linear
    probability, clipped to +-(20 + board area), made non-integer with probability 1 - `komiAllowIntegerProb`.

The draws go to the device loop through `SelfPlay.set_game_setup` / `set_komi` (include/kgb200.h), one per slot, and take effect when
the slot's next game starts.  The reference seeds its GameInitializer from the clock (command/selfplay.cpp:94), so there is no stream to
reproduce: this class uses Python's Mersenne Twister, the distributions are the reference's (tests/golden/gameinit_hist.json: 200 000 games
of the reference's createGame, test_game_initializer_reproduces_the_reference_distributions).
`komiAuto` replaces komiMean by the fair komi of the game's empty board (katago_b200/komi_search.py; `draw_komi(mean=...)`).
Not built (the caller reports them): handicap stones (`handicapProb`), makeGameFair for forks / handicap, start positions, scoring / tax /
button rules other than area scoring without tax and button."""
import math
import random

KO_RULES = {"SIMPLE": 0, "POSITIONAL": 1, "SITUATIONAL": 2, "SPIGHT": 3}
KOMI_CLIP_RADIUS = 20.0     # NNPos::KOMI_CLIP_RADIUS (neuralnet/nninputs.h)


def board_size_distribution(edges, rel_probs, allow_rectangle_prob=0.0):
    """[(x, y), ...], [prob, ...] like GameInitializer::initShared (play.cpp:139-172)."""
    if len(edges) != len(rel_probs):
        raise ValueError(f"bSizeRelProbs has {len(rel_probs)} entries, bSizes has {len(edges)}")
    total = float(sum(rel_probs))
    if not total > 0:
        raise ValueError("bSizeRelProbs must sum to a positive value")
    sizes, probs = [], []
    for i, x in enumerate(edges):
        for j, y in enumerate(edges):
            if i == j:
                sizes.append((x, y))
                probs.append((1.0 - allow_rectangle_prob) * rel_probs[i] / total + allow_rectangle_prob * rel_probs[i] * rel_probs[j] / total / total)
            elif allow_rectangle_prob > 0.0:
                sizes.append((x, y))
                probs.append(allow_rectangle_prob * rel_probs[i] * rel_probs[j] / total / total)
    return sizes, probs


def round_and_clip_komi(unrounded, x_size, y_size):
    """PlayUtils::roundAndClipKomi (playutils.cpp:363-371)."""
    rng = KOMI_CLIP_RADIUS + x_size * y_size
    unrounded = min(max(unrounded, -rng), rng)
    return 0.5 * math.floor(2.0 * unrounded + 0.5) if unrounded >= 0 else -0.5 * math.floor(-2.0 * unrounded + 0.5)    # C round(): halves away from zero


class GameInitializer:
    def __init__(self, sizes, size_probs, ko_rules=(0,), multi_stone_suicide_legals=(True,), komi_mean=7.5, komi_stdev=0.0,
                 komi_big_stdev_prob=0.0, komi_big_stdev=10.0, komi_bigger_stdev_prob=0.0, komi_bigger_stdev=30.0, komi_allow_integer_prob=1.0, seed=0):
        if not sizes or len(sizes) != len(size_probs):
            raise ValueError("GameInitializer: one probability per board size")
        self.sizes, self.size_probs = [tuple(s) for s in sizes], [float(p) for p in size_probs]
        self.ko_rules, self.suicides = [int(k) for k in ko_rules], [bool(s) for s in multi_stone_suicide_legals]
        self.komi_mean, self.komi_stdev = float(komi_mean), float(komi_stdev)
        self.komi_big_stdev_prob, self.komi_big_stdev = float(komi_big_stdev_prob), float(komi_big_stdev)
        self.komi_bigger_stdev_prob, self.komi_bigger_stdev = float(komi_bigger_stdev_prob), float(komi_bigger_stdev)
        self.komi_allow_integer_prob = float(komi_allow_integer_prob)
        self.rand = random.Random(seed)

    @property
    def max_edge(self):
        return max(max(s) for s in self.sizes)

    def _gaussian_truncated(self, bound):
        d = self.rand.gauss(0.0, 1.0)
        while d < -bound or d > bound:
            d = self.rand.gauss(0.0, 1.0)
        return d

    def draw_komi(self, x_size, y_size, mean=None):
        """chooseExtraBlackAndKomi (no handicap) + setKomiWithNoise.  mean: instead of komiMean (komiAuto: the fair komi of the empty board)."""
        r = self.rand
        stdev = self.komi_stdev if self.komi_stdev > 0 else 0.0
        if self.komi_big_stdev > 0 and r.random() < self.komi_big_stdev_prob:
            stdev = self.komi_big_stdev
        if self.komi_bigger_stdev > 0 and self.komi_bigger_stdev_prob > 0 and r.random() < self.komi_bigger_stdev_prob:
            stdev = self.komi_bigger_stdev
        stdev *= math.sqrt(x_size * y_size) / 19.0        # no massive komis on small boards
        allow_integer = r.random() < self.komi_allow_integer_prob
        komi = self.komi_mean if mean is None else float(mean)
        if stdev > 0:
            komi += stdev * self._gaussian_truncated(3.0)
        lower, upper = math.floor(komi * 2.0) / 2.0, math.ceil(komi * 2.0) / 2.0       # roundKomiWithLinearProb
        komi = lower if lower == upper else (upper if r.random() < (komi - lower) / (upper - lower) else lower)
        komi = round_and_clip_komi(komi, x_size, y_size)
        if not allow_integer and komi == int(komi):
            komi += -0.5 if r.random() < 0.5 else 0.5
        return float(komi)

    def draw(self):
        """One game: (board X, board Y, ko rule, multi-stone suicide legal 0/1, komi)."""
        x, y = self.rand.choices(self.sizes, weights=self.size_probs)[0]
        ko = self.ko_rules[self.rand.randrange(len(self.ko_rules))]
        suicide = self.suicides[self.rand.randrange(len(self.suicides))]
        return x, y, ko, int(suicide), self.draw_komi(x, y)

    def draw_many(self, n):
        """(setups int32 [n, 4], komis float32 [n]) for SelfPlay.set_game_setup / set_komi."""
        import numpy as np
        g = [self.draw() for _ in range(n)]
        return np.array([t[:4] for t in g], np.int32).reshape(n, 4), np.array([t[4] for t in g], np.float32)
