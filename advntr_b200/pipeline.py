"""Many loci at once: keyword pre-filter -> one batched Viterbi call -> per-locus genotype.

The shape of ``GenomeAnalyzer.find_repeat_counts_from_alignment_file`` (``genome_analyzer.py:273-297``)
without its file IO: the reference writes the keywords of every target locus, runs the
``adVNTR-Filtering`` binary over the unmapped reads, then loops over the loci one by one and over
their reads one by one (``vntr_finder.py:727-767``).  Here

1. the unmapped reads are scanned ONCE on the device against the keywords of all loci
   (``keyword_filter.KeywordFilter``, same selection and per-locus cap as the binary);
2. the mapped reads of every locus and both strands of its filtered unmapped reads go to the device
   in ONE ``advhmm_viterbi_multi_summary`` call; the backtrack kernels reduce every path on the
   device to the handful of numbers adVNTR reads off it (``advhmm_read_summary``), so no state path
   crosses the bus;
3. recruitment (``recruit_read``), the spanning test and the genotype statistics of all loci run on those
   numbers in one native call on all host threads (``advhmm_genotypes_from_summaries``).

Decisions are the ones ``LocusDecoder`` makes read by read from full paths (tests/test_pipeline.py
checks they are identical).  The caller either passes the mapped reads per locus (``genotype``) or an
indexed BAM file (``genotype_alignment_file``): then the region fetches, the read-level tests of
``select_illumina_reads`` and the unmapped-read extraction run in ``libadvbam.so`` (``bam_ingest.py``)
and the mapped reads reach the engine as code arrays without passing through Python strings.
"""
from __future__ import annotations

import numpy as np

from . import bam_ingest, engine, fast_compile, genotype, keyword_filter, path_utils
from .locus_batch import LocusDecoder, reverse_complement


class LocusSpec(object):
    """What ``ReferenceVNTR`` holds for one locus (``reference_vntr.py:7-40``)."""

    def __init__(self, locus_id, left_flank, right_flank, repeat_segments, scaled_score=None, chromosome=None,
                 start_point=None, aligned_segments=None):
        self.id = int(locus_id)
        self.chromosome, self.start_point = chromosome, start_point   # only for genotype_alignment_file
        self.left_flank, self.right_flank = left_flank, right_flank
        self.repeat_segments = list(repeat_segments)
        self.pattern = self.repeat_segments[0]
        self.scaled_score = scaled_score
        # multiple alignment of the repeat segments when they differ in length (what the reference asks
        # MUSCLE for, profile_hmm.py:165-171); None: equal-length segments are their own alignment
        self.aligned_segments = aligned_segments


class GenotypingRun(object):
    def __init__(self, loci, read_length=150, keyword_size=15):
        self.ctx = engine.Context.default()         # the context the models' device tables live on
        self.read_length = read_length
        self.loci = list(loci)
        self.decoders = [LocusDecoder(l.left_flank, l.right_flank, l.repeat_segments, read_length,
                                      scaled_score=l.scaled_score, locus_id=l.id,
                                      aligned_segments=getattr(l, "aligned_segments", None)) for l in self.loci]
        # the models of all loci in ONE native call (profiles, parameter chains, device tables, upload)
        fast_compile.attach_device_models([d.model for d in self.decoders], self.ctx)
        keywords = [(l.id, sorted(keyword_filter.get_keywords_for_filtering(
            l.left_flank, l.right_flank, l.repeat_segments, l.pattern, keyword_size=keyword_size))) for l in self.loci]
        self.filter = keyword_filter.KeywordFilter(keywords, ctx=self.ctx)

    @classmethod
    def from_alignment_file(cls, loci, alignment_file, keyword_size=15):
        """Read length = median of the first five records, as ``select_illumina_reads`` sets it
        (``vntr_finder.py:714-718``)."""
        with bam_ingest.AlignmentFile(alignment_file) as f:
            read_length = bam_ingest.median_head_read_length(f)
        return cls(loci, read_length=read_length, keyword_size=keyword_size)

    def close(self):
        self.filter.close()

    # -- step 1 ----------------------------------------------------------------------------------
    def filter_unmapped(self, names, seqs):
        """{locus id: [(name, sequence) ...]} as ``get_filtered_read_ids`` hands them to the finders
        (``genome_analyzer.py:186-197``, ``:289``): the reads the binary lists for the locus, in the
        order of its read section (sorted by name)."""
        per_locus, reads = self.filter.filter_reads(list(names), list(seqs))
        wanted = {vid: set(shown) for vid, _, shown in per_locus}
        return {vid: [(n, s) for n, s in reads if n in wanted[vid]] for vid in wanted}

    # -- steps 2 + 3 -----------------------------------------------------------------------------
    def genotype(self, mapped_reads_by_locus, unmapped_names=(), unmapped_seqs=(), accuracy_filter=False,
                 is_haploid=False):
        """-> {locus id: dict(copy_numbers, recruited_reads_count, spanning_reads_count,
        flanking_reads_count, maximum_likelihood, covered_repeats, flanking_repeats)}."""
        filtered = self.filter_unmapped(unmapped_names, unmapped_seqs) if len(unmapped_names) else {}
        L = self.read_length
        batch, goff, layout = [], [0], []
        for dec in self.decoders:
            mapped = [r.upper() for r in mapped_reads_by_locus.get(dec.id, ()) if "N" not in r.upper()]
            unm = [s.upper() for _, s in filtered.get(dec.id, ()) if "N" not in s.upper() and len(s) >= L]
            batch += mapped
            for s in unm:
                batch += [s, reverse_complement(s)]
            goff.append(len(batch))
            layout.append((len(mapped), len(unm)))
        seqs, off = engine.encode_batch(batch)
        return self._call(seqs, off, goff, layout, accuracy_filter, is_haploid)

    def genotype_alignment_file(self, alignment_file, accuracy_filter=False, is_haploid=False, threads=0):
        """``find_repeat_counts_from_alignment_file`` (``genome_analyzer.py:273-297``) for an indexed BAM:
        unmapped reads -> keyword filter, one region fetch per locus -> the read-level tests of
        ``select_illumina_reads`` (``vntr_finder.py:727-737``, ``utils.py:20-38``), everything decoded in
        one device call.  Every locus needs ``chromosome`` and ``start_point``."""
        L = self.read_length
        with bam_ingest.AlignmentFile(alignment_file) as f:
            names, useqs = bam_ingest.extract_unmapped_reads(f, threads=threads)
            filtered = self.filter_unmapped(names, useqs) if names else {}
            mapped, strands, goff, layout, n = [], [], [0], [], 0
            for dec, spec in zip(self.decoders, self.loci):
                end = spec.start_point + sum(len(seg) for seg in spec.repeat_segments)     # reference_vntr.py:66
                m = bam_ingest.select_mapped_illumina(f, spec.chromosome, spec.start_point, end, L)
                unm = [s.upper() for _, s in filtered.get(dec.id, ()) if "N" not in s.upper() and len(s) >= L]
                for s in unm:
                    strands += [s, reverse_complement(s)]
                mapped.append(m)
                n += len(m["names"]) + 2 * len(unm)
                goff.append(n)
                layout.append((len(m["names"]), len(unm)))
        u_codes, u_off = engine.encode_batch(strands)
        chunks, lens, cur = [], [], 0
        for m, (_, n_unm) in zip(mapped, layout):
            chunks.append(m["codes"])
            lens.append(np.diff(m["off"]))
            chunks.append(u_codes[u_off[cur]:u_off[cur + 2 * n_unm]])
            lens.append(np.diff(u_off[cur:cur + 2 * n_unm + 1]))
            cur += 2 * n_unm
        off = np.zeros(n + 1, dtype=np.int64)
        if n:
            np.cumsum(np.concatenate(lens), out=off[1:])
        seqs = np.ascontiguousarray(np.concatenate(chunks + [np.zeros(1, dtype=np.uint8)]))
        out = self._call(seqs, off, goff, layout, accuracy_filter, is_haploid)
        for m, spec in zip(mapped, self.loci):
            out[spec.id]["vntr_bp_in_mapped_reads"] = m["vntr_bp"]
        return out

    def find_frameshifts(self, reads_by_locus):
        """``--frameshift`` mode (``genome_analyzer.py:260``, ``vntr_finder.py:776-780``, ``:265-309``) for all
        loci: every read decoded to its full state path in ONE device call, recruited reads walked by the native
        consumer (``advhmm_frameshift_candidates``, all host threads), the binomial test per locus
        (``identify_frameshift``, scipy).  ``reads_by_locus``: {locus id: the reads ``find_frameshift_from_alignment_file``
        would select from (mapped reads of the region)}.  -> {locus id: the frame-shifting state label, e.g.
        ``'I7A'`` / ``'D12'``, or None}; ``self.frameshift_records`` keeps the per-locus records."""
        batch, goff = [], [0]
        for dec in self.decoders:
            batch += [r.upper() for r in reads_by_locus.get(dec.id, ()) if "N" not in r.upper()]
            goff.append(len(batch))
        seqs, off = engine.encode_batch(batch)
        models = [d.model._device_model() for d in self.decoders]
        goff = np.asarray(goff, dtype=np.int64)
        res = self.ctx._run(models, goff, seqs, off, False, True, False, None, want_summary=True)
        if not hasattr(self, "_fs_tables"):
            self._fs_tables = [path_utils.frameshift_state_tables([s.name for s in d.model.states]) for d in self.decoders]
        scores = [d.min_score_to_select_a_read() for d in self.decoders]
        rec = engine.frameshift_candidates(goff, [len(d.pattern) for d in self.decoders],
                                           [np.nan if s is None else s for s in scores], self._fs_tables, res.logp,
                                           res.summaries, res.path_len, res.path_off, res.paths, seqs, off)
        self.frameshift_records = rec
        out = {}
        for dec, c in zip(self.decoders, rec):
            vntr_len = sum(len(seg) for seg in dec.segments)                        # reference_vntr.get_length()
            coverage = float(c["repeat_bp"]) / vntr_len / 2
            label = path_utils.frameshift_label(c)
            # find_frameshift_from_selected_reads divides by the coverage (ZeroDivisionError without repeat bases)
            shifted = coverage > 0 and genotype.identify_frameshift(coverage, int(c["count"]), 1 / coverage)
            out[dec.id] = label if shifted else None
        return out

    def _call(self, seqs, off, goff, layout, accuracy_filter, is_haploid):
        models = [d.model._device_model() for d in self.decoders]
        res = self.ctx._run(models, np.asarray(goff, dtype=np.int64), seqs, off, False, False, False, None,
                            want_summary=True)
        calls = native_genotypes_from_summaries(res.logp, res.summaries, res.path_len, off, goff, layout,
                                                [d.min_score_to_select_a_read() for d in self.decoders],
                                                accuracy_filter, is_haploid, self.decoders[0].min_repeat_bp_to_add_read
                                                if self.decoders else 2)
        return {dec.id: c for dec, c in zip(self.decoders, calls)}


def native_genotypes_from_summaries(logp, S, path_len, seq_off, goff, layout, scores, accuracy_filter=False,
                                    is_haploid=False, min_repeat_bp=2, threads=0):
    """The step after the decode for all loci in ONE native call (``advhmm_genotypes_from_summaries``, all
    host threads): same arguments as ``genotypes_from_summaries`` except that the read offsets are passed
    instead of the lengths; same list of result dicts (tests/test_locus_calls.py compares the two)."""
    goff = np.asarray(goff, dtype=np.int64)
    score = np.array([np.nan if s is None else s for s in scores], dtype=np.float64)
    calls, cls = engine.genotypes_from_summaries(goff, [m for m, _ in layout], [u for _, u in layout], score, logp, S,
                                                 path_len, seq_off, accuracy_filter, is_haploid, min_repeat_bp, threads,
                                                 want_read_class=True)
    repeats = np.asarray(S["repeats"])
    out = []
    for g, c in enumerate(calls):
        a, b = int(goff[g]), int(goff[g + 1])
        k, r = cls[a:b], repeats[a:b]
        out.append({"copy_numbers": (int(c["c1"]), int(c["c2"])) if c["has_call"] else None,
                    "recruited_reads_count": int(c["recruited"]), "spanning_reads_count": int(c["spanning"]),
                    "flanking_reads_count": int(c["flanking"]), "maximum_likelihood": float(c["max_prob"]),
                    "covered_repeats": r[k == 1].tolist(),
                    "flanking_repeats": [] if accuracy_filter else sorted(r[k == 2].tolist())})
    return out


def genotypes_from_summaries(logp, S, path_len, lens, goff, layout, scores, accuracy_filter=False, is_haploid=False,
                             min_repeat_bp=2):
    """The literal numpy / Python form of the step after the decode (the checker of the native
    ``advhmm_genotypes_from_summaries``, which ``GenotypingRun`` uses).
    Recruitment (``recruit_read``, ``vntr_finder.py:179-190``), the better strand of every unmapped read
    (``:235-254``), the spanning test (``:311-322``) and the genotype statistics (``:846-875``) of every locus
    from the per-read device results: ``logp``, the on-device path summaries ``S`` (``advhmm_read_summary``
    records), ``path_len`` (< 0: impossible read) and the read lengths.  ``layout[g]`` = (mapped reads,
    unmapped reads) of locus g, whose reads start at ``goff[g]`` (mapped first, then both strands of every
    unmapped read); ``scores[g]`` = its minimum Viterbi score or None.  -> one result dict per locus."""
    with np.errstate(divide="ignore", invalid="ignore"):
        right = np.where(S["right_bp"] > 0, S["right_hits"] / S["right_bp"].astype(np.float64), 1.0)
        left = np.where(S["left_bp"] > 0, S["left_hits"] / S["left_bp"].astype(np.float64), 1.0)
    rate = np.minimum(right, left)
    possible = path_len >= 0
    # everything that does not depend on the locus, for all reads at once
    ok_no_score = (S["n_match"] >= 0.9 * lens) & (logp > -lens) & (rate >= 0.9) & possible
    ok_rate = (rate >= 0.9) & possible
    spanning_all = (rate >= 0.95) & (S["left_bp"] > 5) & (S["right_bp"] > 5)
    repeats, repeat_bp = S["repeats"], S["repeat_bp"]
    out = []
    for (n_mapped, n_unm), a, score in zip(layout, goff[:-1], scores):
        a = int(a)
        idx = np.arange(a, a + n_mapped)
        if n_unm:                                   # the better strand of every unmapped read
            f = a + n_mapped + 2 * np.arange(n_unm)
            best = np.where(logp[f] < logp[f + 1], f + 1, f)
            best = best[repeat_bp[best] > min_repeat_bp]
            idx = np.concatenate([idx, best])
        keep = (logp[idx] > score) & ok_rate[idx] if score is not None else ok_no_score[idx]
        sel = idx[keep]
        spanning = spanning_all[sel]
        covered = repeats[sel][spanning].tolist()
        flanking = [] if accuracy_filter else sorted(repeats[sel][~spanning].tolist())
        cn, prob = genotype.genotype_from_illumina_counts(covered, flanking, accuracy_filter, is_haploid)
        kept = genotype._drop_unsupported(covered) if accuracy_filter else covered
        out.append({"copy_numbers": cn, "recruited_reads_count": int(len(sel)),
                    "spanning_reads_count": len(kept), "flanking_reads_count": len(flanking),
                    "maximum_likelihood": prob, "covered_repeats": covered, "flanking_repeats": flanking})
    return out
