"""Seeded random binary CSPs exercising every constraint kind / domain quirk the engine supports."""
from __future__ import annotations

import random
from typing import List

from dequan_b200.model import (CSP, AllDifferentConstraint, Domain, DomainType, EqualityConstraint, Op,
                               OpConstraint, OrRangeConstraint, TableConstraint)


def random_domain(rng: random.Random, lo: int, hi: int, max_size: int) -> Domain:
    kind = rng.random()
    if kind < 0.4:  # single range
        a = rng.randint(lo, hi - 1)
        b = min(hi, a + rng.randint(1, max_size))
        return Domain(DomainType.Ranges, [a, b])
    if kind < 0.6:  # two ranges (ascending, disjoint)
        a = rng.randint(lo, hi - 3)
        b = a + rng.randint(1, max(1, max_size // 2))
        c = b + rng.randint(1, 3)
        d = c + rng.randint(1, max(1, max_size // 2))
        return Domain(DomainType.Ranges, [a, b, c, d])
    if kind < 0.9:  # explicit values, unsorted, no duplicates
        n = rng.randint(1, max_size)
        vals = rng.sample(range(lo, hi + max_size), n)
        return Domain(DomainType.Values, vals)
    return Domain(DomainType.Values, [rng.randint(lo, hi)])  # fixed var


def random_model(seed: int, n_vars: int = 6, n_cons: int = 8, max_dom: int = 6, kinds: str = "all") -> CSP:
    rng = random.Random(seed)
    csp = CSP()
    for _ in range(n_vars):
        csp.AddIntVar(random_domain(rng, -3, 6, max_dom))
    for _ in range(n_cons):
        a, b = rng.sample(range(n_vars), 2)
        r = rng.random()
        if kinds == "ne":
            csp.AddConstraint(OpConstraint(a, b, Op.NotEqual, rng.choice([0, 0, 1, -1, 2])))
        elif r < 0.45:
            csp.AddConstraint(OpConstraint(a, b, Op(rng.randint(0, 5)), rng.randint(-2, 2)))
        elif r < 0.55:
            csp.AddConstraint(EqualityConstraint(a, b))
        elif r < 0.70:
            m = rng.randint(2, min(4, n_vars))
            csp.AddConstraint(AllDifferentConstraint(rng.sample(range(n_vars), m)))
        elif r < 0.80:
            lo = rng.randint(-3, 4)
            csp.AddConstraint(OrRangeConstraint(a, b, lo, lo + rng.randint(1, 5)))
        else:
            pairs = [(x, y) for x in range(-3, 13) for y in range(-3, 13) if rng.random() < 0.6]
            csp.AddConstraint(TableConstraint(a, b, pairs))
    csp.FinalizeModel()
    return csp


def model_suite(n: int = 200, seed0: int = 1000) -> List[CSP]:
    out = []
    for i in range(n):
        rng = random.Random(seed0 + i)
        nv = rng.randint(2, 9)
        nc = rng.randint(1, 14)
        out.append(random_model(seed0 + i, nv, nc, rng.randint(2, 7), "ne" if i % 5 == 0 else "all"))
    return out


def random_model_dups(seed: int, n_vars: int = 6, n_cons: int = 9, max_dom: int = 5) -> CSP:
    """random_model with Values domains that list some values more than once (SURVEY.md par. 9 Q2: iteration visits every
    copy, Domain::Exclude erases the first match only, Domain::Intersect leaves one copy)."""
    csp = random_model(seed, n_vars, n_cons, max_dom)
    rng = random.Random(seed * 7919 + 13)
    for d in csp.domains:
        if d.type == DomainType.Values and rng.random() < 0.7:
            for _ in range(rng.randint(1, 3)):
                d.values.insert(rng.randint(0, len(d.values)), rng.choice(d.values))
    return csp


def dup_suite(n: int = 150, seed0: int = 5000) -> List[CSP]:
    out = []
    for i in range(n):
        rng = random.Random(seed0 + i)
        out.append(random_model_dups(seed0 + i, rng.randint(2, 8), rng.randint(2, 12), rng.randint(2, 5)))
    return out
