"""The compare order used by max3_first (kernels_banded.cuh) picks the same winner as the reference's
running 'replace on strictly greater' (hmm.pyx:2035-2042) for every combination of ties and -inf."""
import itertools
import math


def reference_winner(a):
    best, arg = a[0], 0
    for k in (1, 2):
        if a[k] > best:
            best, arg = a[k], k
    return best, arg


def kernel_winner(a):
    p2 = a[2] > a[1]                      # the two early candidates first
    t, targ = (a[2], 2) if p2 else (a[1], 1)
    p1 = t > a[0]                         # the late candidate meets their winner
    bits = (1 if p1 else 0) | (2 if p2 else 0)
    decoded = (1 + ((bits >> 1) & 1)) if bits & 1 else 0       # banded_walk
    return (t if p1 else a[0]), decoded


def test_same_winner_for_every_tie_pattern():
    values = [-math.inf, -3.5, -3.5 + 2 ** -50, -1.0, 0.0]
    for a in itertools.product(values, repeat=3):
        assert kernel_winner(a) == reference_winner(a), a
