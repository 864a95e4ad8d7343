"""Genotype call from the repeat counts the Viterbi paths give (host side, downstream of the path).

Restates the statistics ``vntr_finder.py`` applies to the observed repeat counts, so that a locus
decoded on the device ends in the same genotype tuple as the reference:

* ``find_genotype_based_on_observed_repeats``  ``vntr_finder.py:473-532`` (+ ``get_conditional_likelihood`` ``:466-484``)
* ``genotype_from_illumina_counts``             ``vntr_finder.py:850-879`` (spanning + flanking reads)
* ``dominant_copy_numbers``                      ``vntr_finder.py:566-585`` (PacBio spanning reads)
* ``identify_frameshift``                        ``vntr_finder.py:256-263``

The arithmetic order (which likelihoods are multiplied, in which order, which posterior wins a
tie) follows the reference so that ``max_prob`` is the same float, not only the same genotype.
"""
from __future__ import annotations

from collections import Counter

import numpy as np

SEQUENCING_ERROR = 0.03            # r, vntr_finder.py:498
SR_MIN_SUPPORT = 3                 # settings.ACCURACY_FILTER_SR_MIN_SUPPORT / vntr_finder.py:859


def conditional_likelihood(ck, ci, cj, r, r_e):
    """P(observing count ck | genotype (ci, cj))."""
    if ck == ci == cj:
        return 1 - r
    if cj == 0:
        return 0.5 * (1 - r)
    if ck == ci:
        return 0.5 * ((1 - r) + r_e ** abs(ck - cj))
    if ck == cj:
        return 0.5 * ((1 - r) + r_e ** abs(ck - ci))
    return 0.5 * (r_e ** abs(ck - ci) + r_e ** abs(ck - cj))


def find_genotype_based_on_observed_repeats(observed_copy_numbers, is_haploid=False):
    """-> ``((c1, c2), max_prob)``, or ``(None, 1e-20)`` when nothing was observed."""
    counts = {}
    for cn in observed_copy_numbers:
        counts[cn] = counts.get(cn, 0) + 1
    if len(counts) < 2:
        prior = 0.5
        counts[0] = 1
    else:
        prior = 1.0 / (len(counts) * (len(counts) - 1) / 2)
    ranked = sorted(counts.items(), key=lambda kv: kv[1], reverse=True)     # stable, as the reference's
    r = SEQUENCING_ERROR
    r_e = r / (2 + r)
    factors = {}
    for ck, occ in ranked:
        if ck == 0:
            continue
        for i in range(len(ranked)):
            ci = ranked[i][0]
            for j in range(i, len(ranked)):
                if is_haploid and i != j:
                    continue
                cj = ranked[j][0]
                factors.setdefault((ci, cj), []).append(conditional_likelihood(ck, ci, cj, r, r_e) ** occ)
    posteriors = {key: np.prod(np.array(vals)) * prior for key, vals in factors.items()}
    total = sum(posteriors.values())
    max_prob, result = 1e-20, None
    for key, value in posteriors.items():
        if value / total > max_prob:
            max_prob, result = value / total, key
    return result, max_prob


def _drop_unsupported(counts, min_support=SR_MIN_SUPPORT):
    kept = []
    for key, count in Counter(counts).most_common():
        if count >= min_support:
            kept.extend([key] * count)
    return kept


def genotype_from_illumina_counts(covered_repeats, flanking_repeats, accuracy_filter=False, is_haploid=False):
    """Spanning-read counts plus, when at least five flanking reads agree on the largest lower
    bound and it is not below the largest spanning count, that lower bound."""
    flanking = sorted(flanking_repeats)
    covered = list(covered_repeats)
    floor = max(covered) if covered else 0
    top = [x for x in flanking if x == max(flanking) and x >= floor]
    if len(top) < 5:
        top = []
    if accuracy_filter:
        covered = _drop_unsupported(covered)
        top = []
    return find_genotype_based_on_observed_repeats(covered + top, is_haploid)


def dominant_copy_numbers(observed_copy_numbers, accuracy_filter=False, is_haploid=False):
    """PacBio: genotype from the repeat counts of the spanning reads."""
    observed = list(observed_copy_numbers)
    if accuracy_filter:
        observed = _drop_unsupported(observed)
    return find_genotype_based_on_observed_repeats(observed, is_haploid)


def identify_frameshift(location_coverage, observed_indel_transitions, expected_indels, error_rate=0.01):
    """Binomial test: is the indel seen too often to be sequencing error?"""
    if observed_indel_transitions >= location_coverage:
        return True
    from scipy.stats import binom
    sequencing_error_prob = binom.pmf(observed_indel_transitions, location_coverage, error_rate)
    frameshift_prob = binom.pmf(observed_indel_transitions, location_coverage, expected_indels)
    return sequencing_error_prob / frameshift_prob < 0.01
