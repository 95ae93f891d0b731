"""Seed contract of the stub: np_random(seed) -> (RandomState(seed), seed).

The reference's only golden vector (TDBoard.py:675-676) seeds a RandomState
directly; real gym<=0.21 hashes the seed first, which cannot be reproduced
offline, so RandomState(seed) is the documented contract.

DRAW_BUDGET (None = unlimited) bounds the number of `randint` calls a generator
may serve; it implements the seed-skip rule of SURVEY.md 9.8 for the reference's
road generator, which never terminates for ~2 % of L=10 / 3-road seeds.
"""
import numpy as np

DRAW_BUDGET = None


class BudgetExceeded(RuntimeError):
    pass


class CountingRandomState(np.random.RandomState):
    """RandomState that counts randint() calls (free and consuming alike)."""

    def __init__(self, seed=None):
        super().__init__(seed)
        self.n_randint = 0
        self.budget = DRAW_BUDGET

    def randint(self, *a, **kw):
        self.n_randint += 1
        if self.budget is not None and self.n_randint > self.budget:
            raise BudgetExceeded("randint budget %d exceeded" % self.budget)
        return super().randint(*a, **kw)


def np_random(seed=None):
    rng = CountingRandomState(seed)
    return rng, seed
