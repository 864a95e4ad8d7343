"""The call sites of the hot path, batched: every read of a locus in ONE device call.

``vntr_finder.py`` decodes one read at a time (``:738`` mapped reads, ``:239-246`` unmapped
reads on both strands).  ``LocusDecoder`` keeps the same decisions -- which strand of an
unmapped read wins, which reads are recruited, which are spanning -- but feeds all reads of
the locus (or of many loci, ``decode_many``) to the engine at once and applies the path
consumers to the returned paths.  The statistics applied to the repeat counts afterwards (genotype likelihood,
frameshift test) are in ``genotype.py``; BAM input is in ``bam_ingest.py`` (``pipeline.py`` chains it);
the flank alignment of UNALIGNED PacBio reads (pairwise2) and the DNN pre-filter stay outside
(out of scope).
"""
from __future__ import annotations

import os

from . import engine, fast_compile, genotype, path_utils, pomegranate, read_matcher


class SelectedRead(object):
    """``vntr_finder.py:45-52``."""

    def __init__(self, sequence, logp, vpath, is_mapped=True, query_name=None, mapq=None, reference_start=None):
        self.sequence, self.logp, self.vpath = sequence, logp, vpath
        self.is_mapped, self.query_name = is_mapped, query_name
        self.mapq, self.reference_start = mapq, reference_start


_COMP = str.maketrans("ACGT", "TGCA")


def reverse_complement(seq):
    return seq.translate(_COMP)[::-1]


class LocusDecoder(object):
    def __init__(self, left_flank, right_flank, repeat_segments, read_length=150, scaled_score=None,
                 error_rate=read_matcher.DEFAULT_MAX_ERROR_RATE, flank_size=None, locus_id=None,
                 trained_hmms_dir=None, aligned_segments=None):
        """``aligned_segments``: a multiple alignment of ``repeat_segments`` (equal-length strings over
        ``ACGT-``).  The reference gets it from MUSCLE when the segments differ in length
        (``profile_hmm.py:165-171``); MUSCLE is not part of this package, so a caller with such a locus
        passes the alignment (equal-length segments are their own alignment)."""
        # get_vntr_matcher_hmm builds the matcher with flanking_region_size = read_length
        # (vntr_finder.py:131-132): a 100 bp or 250 bp library gets 100 / 250-base flank models
        if flank_size is None:
            flank_size = read_length
        self.flank_size = flank_size
        self.id = locus_id
        self.left_flank, self.right_flank = left_flank, right_flank
        self.segments = list(repeat_segments)
        if aligned_segments is not None:
            aligned_segments = [a.upper() for a in aligned_segments]
            if sorted(a.replace("-", "") for a in aligned_segments) != sorted(x.upper() for x in self.segments):
                raise ValueError("aligned_segments is not an alignment of repeat_segments")
        profile_rows = aligned_segments if aligned_segments is not None else self.segments
        self.pattern = self.segments[0]
        self.read_length = read_length
        self.scaled_score = scaled_score
        self.min_repeat_bp_to_add_read = 2            # vntr_finder.py:66-69
        self.error_rate = error_rate
        copies = read_matcher.copies_for_read_length(read_length, len(self.pattern))
        # get_vntr_matcher_hmm (vntr_finder.py:116-137) with settings.USE_TRAINED_HMMS: a model stored
        # by an earlier run is reloaded (from_json bakes it with merge='All', so it is NOT the same
        # state set as a fresh build -- as in the reference), otherwise built and stored
        stored = None if trained_hmms_dir is None else \
            os.path.join(trained_hmms_dir, "%s_%s.json" % (locus_id, read_length))
        if stored is not None and os.path.isfile(stored):
            self.model = pomegranate.HiddenMarkovModel.from_json(stored)
            return
        if stored is not None:                       # the graph-level builder: only it can serialise
            self.model = read_matcher.build_vntr_matcher_hmm(left_flank, right_flank, profile_rows, copies,
                                                             flank_size=flank_size, error_rate=error_rate)
            with open(stored, "w") as outfile:
                outfile.write(self.model.to_json())
            return
        self.model = fast_compile.build_vntr_matcher_hmm(left_flank, right_flank, profile_rows, copies,
                                                         flank_size=flank_size, error_rate=error_rate)

    # vntr_finder.py:174-177
    def min_score_to_select_a_read(self, read_length=None):
        if not self.scaled_score:
            return None
        return self.scaled_score * (read_length or self.read_length)

    def _vpath(self, res, i):
        p = res.path(i)
        if p is None:
            return None
        st = self.model.states
        return [(int(k), st[k]) for k in p]

    def select_reads(self, mapped_reads, unmapped_reads=()):
        """``select_illumina_reads`` (``vntr_finder.py:701-773``) minus the file IO: mapped reads
        are decoded as given, unmapped reads on both strands keeping the better one."""
        mapped = [r.upper() for r in mapped_reads if "N" not in r.upper()]
        unmapped = [r.upper() for r in unmapped_reads if "N" not in r.upper() and len(r) >= self.read_length]
        batch = mapped + [s for r in unmapped for s in (r, reverse_complement(r))]
        res = self.model.viterbi_batch(batch)
        score = self.min_score_to_select_a_read()
        selected = []
        for i, seq in enumerate(mapped):
            vp = self._vpath(res, i)
            if vp is None:
                continue
            if path_utils.recruit_read(res.logp[i], vp, score, seq, self.left_flank, self.right_flank):
                selected.append(SelectedRead(seq, float(res.logp[i]), vp, True))
        base = len(mapped)
        for j, seq in enumerate(unmapped):
            f, r = base + 2 * j, base + 2 * j + 1
            k, s = (r, reverse_complement(seq)) if res.logp[f] < res.logp[r] else (f, seq)
            vp = self._vpath(res, k)
            if vp is None:
                continue
            repeat_bp = path_utils.get_number_of_repeat_bp_matches_in_vpath(vp)
            if path_utils.recruit_read(res.logp[k], vp, score, s, self.left_flank, self.right_flank) and \
                    repeat_bp > self.min_repeat_bp_to_add_read:
                selected.append(SelectedRead(s, float(res.logp[k]), vp, False))
        return selected

    def select_reads_from_alignment_file(self, alignment_file, chromosome, start_point, unmapped_filtered_reads=()):
        """``select_illumina_reads(alignment_file, unmapped_filtered_reads)`` (``vntr_finder.py:701-773``)
        with its file access: the region fetch and the read-level tests run in ``libadvbam``
        (``bam_ingest.select_mapped_illumina``); mapped reads keep ``mapq``, ``reference_start`` and
        ``query_name`` as the reference's ``SelectedRead`` does.  ``find_frameshift_from_alignment_file``
        (``:776-780``) is ``frameshift_candidate(select_reads_from_alignment_file(...))``."""
        from . import bam_ingest
        own = not isinstance(alignment_file, bam_ingest.AlignmentFile)
        f = bam_ingest.AlignmentFile(alignment_file) if own else alignment_file
        try:
            end = start_point + sum(len(seg) for seg in self.segments)
            m = bam_ingest.select_mapped_illumina(f, chromosome, start_point, end, self.read_length)
        finally:
            if own:
                f.close()
        mapped = ["".join("ACGT"[c] for c in m["codes"][m["off"][i]:m["off"][i + 1]]) for i in range(len(m["names"]))]
        selected = self.select_reads(mapped, unmapped_filtered_reads)
        by_seq = {}
        for i, seq in enumerate(mapped):
            by_seq.setdefault(seq, []).append(i)
        for read in selected:
            if read.is_mapped:
                i = by_seq[read.sequence].pop(0)      # select_reads keeps the order of the mapped reads
                read.query_name, read.mapq, read.reference_start = m["names"][i], int(m["mapq"][i]), int(m["reference_start"][i])
        return selected

    def observed_repeats(self, selected, accuracy_filter=False):
        """``find_repeat_count_from_alignment_file`` (``vntr_finder.py:810-850``): repeat counts
        of spanning reads and (lower bounds from) flanking reads."""
        covered, flanking = [], []
        for read in selected:
            n = path_utils.get_number_of_repeats_in_vpath(read.vpath)
            if path_utils.read_flanks_repeats_with_confidence(read.vpath, read.sequence, self.left_flank,
                                                              self.right_flank):
                covered.append(n)
            elif not accuracy_filter:
                flanking.append(n)
        return covered, sorted(flanking)

    def genotype(self, selected, accuracy_filter=False, is_haploid=False):
        """The genotype call of ``find_repeat_count_from_alignment_file`` without a coverage estimate
        (``vntr_finder.py:846-875``) -> dict with the fields of the reference's ``GenotypeResult``."""
        covered, flanking = self.observed_repeats(selected, accuracy_filter)
        copy_numbers, max_prob = genotype.genotype_from_illumina_counts(covered, flanking, accuracy_filter, is_haploid)
        if accuracy_filter:
            covered = genotype._drop_unsupported(covered)
        return {"copy_numbers": copy_numbers, "recruited_reads_count": len(selected),
                "spanning_reads_count": len(covered), "flanking_reads_count": len(flanking),
                "maximum_likelihood": max_prob}

    def updated_model(self, selected):
        """One ``--update`` step (``vntr_finder.py:667-698``): re-estimate the repeat-unit profile from
        the repeat segments of the selected reads and of the reference repeats (decoded on the
        current model), and rebuild the read matcher (``get_read_matcher_model(..., vpaths)``).

        The reference's loop compares a fitness that it computes from the ORIGINAL selection
        (``:693``), so it always stops after the first rebuild; this is that rebuild."""
        ref = [seg.upper() for seg in self.segments]
        res = self.model.viterbi_batch(ref)
        vpaths = [(r.sequence, r.vpath) for r in selected]
        vpaths += [(seq, self._vpath(res, i)) for i, seq in enumerate(ref)]
        copies = read_matcher.copies_for_read_length(self.read_length, len(self.pattern))
        flank = self.read_length                      # vntr_finder.py:681-683
        # get_read_matcher_model(..., vpaths) estimates the repeat-unit profile from the alignment of the repeat
        # segments the paths mark out (hmm_utils.py:427-429); everything else is the model of that alignment,
        # which the native compiler builds (tests: update_model.npz, made by the reference's own call)
        alignment = path_utils.get_multiple_alignment_of_repeats_from_reads(vpaths)
        return fast_compile.get_read_matcher_model(self.left_flank[-flank:], self.right_flank[:flank], alignment,
                                                   copies, error_rate=self.error_rate)

    def frameshift_candidate(self, selected):
        """Most frequent frame-shifting indel state among the selected reads and its count
        (``vntr_finder.py:265-300``); the binomial test on it is ``identify_frameshift``."""
        mutations, repeat_bp = path_utils.frameshift_mutations(selected, len(self.pattern))
        ranked = sorted(mutations.items(), key=lambda x: x[1])
        return (ranked[-1] if ranked else (None, 0)), repeat_bp


def decode_many(decoders, reads_per_locus, ctx=None, both_strands=False):
    """One device call for the reads of many loci (``advhmm_viterbi_multi``)."""
    ctx = ctx or engine.Context.default()
    models = [d.model._device_model() for d in decoders]
    groups = [[d.model._encode(r) for r in reads] for d, reads in zip(decoders, reads_per_locus)]
    return ctx.viterbi_multi(models, groups, both_strands=both_strands)


def dominant_copy_numbers_from_spanning_reads(left_flank, right_flank, repeat_segments, spanning_reads,
                                              error_rate=0.3, accuracy_filter=False, is_haploid=False):
    """PacBio call site (``vntr_finder.py:534-585``): one model with enough unrolled copies for the
    longest spanning read (100 bp flanks, ``:109``, ``:549``), every spanning read decoded in one
    device call (the long-read kernel), genotype from the repeat counts.
    -> ``(copy_numbers, max_prob, observed_copy_numbers)``."""
    if len(spanning_reads) < 1:
        return None, 0, []
    pattern = repeat_segments[0]
    longest = max(max(len(r) - 100 for r in spanning_reads), 0)
    max_copies = int(round(longest / float(len(pattern))))
    model = fast_compile.build_vntr_matcher_hmm(left_flank, right_flank, list(repeat_segments), max_copies,
                                                flank_size=100, error_rate=error_rate)
    # the repeat count of every read comes from the on-device path reducer (get_number_of_repeats_in_vpath on the
    # device): the 10-20 k state paths of PacBio reads stay there
    res = model.viterbi_batch([r.upper() for r in spanning_reads], want_path=False, want_summary=True)
    if (res.path_len < 0).any():                     # the reference subscripts the None path of an impossible read
        raise TypeError("'NoneType' object is not subscriptable")
    observed = [int(x) for x in res.summaries["repeats"]]
    copy_numbers, max_prob = genotype.dominant_copy_numbers(observed, accuracy_filter, is_haploid)
    return copy_numbers, max_prob, observed


def repeat_count_from_pacbio_alignment_file(alignment_file, chromosome, start_point, left_flank, right_flank,
                                            repeat_segments, unaligned_spanning_reads=(), error_rate=0.3,
                                            accuracy_filter=False, is_haploid=False):
    """``find_repeat_count_from_pacbio_alignment_file`` (``vntr_finder.py:640-651``) for an indexed BAM:
    the mapped reads that span the locus are cut to the modelled region by the CIGAR walk of
    ``check_if_pacbio_mapped_read_spans_vntr`` (``bam_ingest.spanning_pacbio_segments``, native), joined
    with spanning reads found among the unaligned ones (their flank alignment, pairwise2, is not part of
    this repo: pass the cut sequences), and genotyped.  -> the fields of the reference's ``GenotypeResult``."""
    from . import bam_ingest
    vntr_end = start_point + sum(len(seg) for seg in repeat_segments)            # reference_vntr.py:66
    own = not isinstance(alignment_file, bam_ingest.AlignmentFile)
    f = bam_ingest.AlignmentFile(alignment_file) if own else alignment_file
    try:
        mapped = [seq for _, seq, _ in bam_ingest.spanning_pacbio_segments(f, chromosome, start_point, vntr_end)]
    finally:
        if own:
            f.close()
    spanning = mapped + list(unaligned_spanning_reads)                           # :646
    copy_numbers, max_prob, observed = dominant_copy_numbers_from_spanning_reads(
        left_flank, right_flank, repeat_segments, spanning, error_rate, accuracy_filter, is_haploid)
    return {"copy_numbers": copy_numbers, "recruited_reads_count": len(spanning),
            "spanning_reads_count": len(spanning), "flanking_reads_count": 0, "maximum_likelihood": max_prob,
            "observed_repeats": observed}
