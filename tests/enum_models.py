"""The fixed model set of the enumeration goldens (tests/golden/enumerate_reference.json)."""
from dequan_b200.model import CSP, Domain, DomainType, Op, OpConstraint, colouring, nqueens
from randmodels import random_model

import numpy as np


def enum_models():
    out = [(f"nqueens{n}", nqueens(n)) for n in (1, 2, 4, 5, 6, 7, 8)]
    for seed in range(7000, 7040):                      # every constraint kind and domain quirk
        out.append((f"random{seed}", random_model(seed, n_vars=3 + seed % 4, n_cons=1 + seed % 4, max_dom=3 + seed % 3)))
    for seed in range(7100, 7120):                      # not-equal models with offsets
        out.append((f"ne{seed}", random_model(seed, n_vars=4 + seed % 4, n_cons=5 + seed % 6, max_dom=3 + seed % 2, kinds="ne")))
    for seed in range(7200, 7212):                      # domains of 33..64 values (64-bit domain words on the device)
        out.append((f"wide{seed}", random_model(seed, n_vars=3 + seed % 2, n_cons=4 + seed % 3, max_dom=40 + 2 * (seed % 12))))
    edges = np.array([[0, 1], [1, 2], [2, 3], [3, 0], [0, 2], [4, 0], [4, 3], [5, 1]], dtype=np.uint8)
    out.append(("colour6_k3", colouring(6, 3, edges)))
    ordered = CSP()                                      # unsorted Values domains: enumeration follows the list order
    a = ordered.AddIntVar(Domain(DomainType.Values, [5, 1, 3]))
    b = ordered.AddIntVar(Domain(DomainType.Values, [2, 9, 4, 0]))
    c = ordered.AddIntVar(Domain(DomainType.Ranges, [0, 3, 7, 9]))
    ordered.AddConstraint(OpConstraint(a, b, Op.NotEqual, 1))
    ordered.AddConstraint(OpConstraint(c, a, Op.Inf, 0))
    ordered.FinalizeModel()
    out.append(("ordered_values", ordered))
    return out
