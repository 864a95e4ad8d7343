"""Read ingest (libadvbam.so) against the plain-Python restatement in oracle/bam_oracle.py.

The BAM / BAI fixtures are written by tests/bam_writer.py (no samtools / pysam in this image; see the
oracle's header: parity of this row is pinned on the published formats, not on pysam outputs).
"""
import os
import random
import re

import numpy as np
import pytest

from conftest import ROOT

import bam_oracle
import bam_writer
from advntr_b200 import bam_ingest, build

build.build_bam_library()

REFS = [("chr1", 3000000), ("chr2", 150000000), ("chrM", 16571)]
LOCI = [(0, 40000, 40090), (0, 1000020, 1000075), (1, 16383990, 16384060), (1, 134217700, 134217800), (2, 5000, 5040)]


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def rand_quals(rng, n):
    kind = rng.random()
    if kind < 0.45:
        return [rng.randint(25, 40) for _ in range(n)]
    if kind < 0.6:                                      # scattered low-quality bases around the 10 % limit
        q = [rng.randint(25, 40) for _ in range(n)]
        for i in rng.sample(range(n), min(n, rng.randint(0, max(1, n // 6)))):
            q[i] = rng.randint(0, 19)
        return q
    if kind < 0.85:                                     # one run of low-quality bases, possibly at the end
        q = [rng.randint(20, 40) for _ in range(n)]
        run = rng.randint(1, 6)
        s = rng.choice([rng.randint(0, max(0, n - run)), max(0, n - run), max(0, n - rng.randint(1, 3))])
        for i in range(s, min(n, s + run)):
            q[i] = rng.randint(0, 19)
        return q
    return [rng.randint(0, 40) for _ in range(n)]


def make_reads(seed, n_per_locus=60):
    rng = random.Random(seed)
    reads = []
    for tid, start, end in LOCI:
        for k in range(n_per_locus):
            L = rng.choice([150, 150, 150, 150, 148, 134, 120, 60, 30])
            pos = rng.randint(max(0, start - 260), end + 60)
            shape = rng.random()
            if shape < 0.5:
                cigar = "%dM" % L
            elif shape < 0.6:
                c = rng.randint(1, min(40, L - 10))
                cigar = "%dS%dM" % (c, L - c)
            elif shape < 0.7:
                a = rng.randint(10, L - 20)
                cigar = "%dM%dD%dM" % (a, rng.randint(1, 30), L - a)
            elif shape < 0.8:
                a, i = rng.randint(5, L - 20), rng.randint(1, 8)
                cigar = "%dM%dI%dM" % (a, i, L - a - i)
            elif shape < 0.85:
                a = rng.randint(10, L - 20)
                cigar = "5H%dM%dN%dM3H" % (a, rng.choice([200, 20000, 300000]), L - a)   # higher-level bins
            elif shape < 0.9:
                cigar = "%d=%dX%d=" % (L - 11, 1, 10)
            elif shape < 0.95:
                cigar = "%dM%dS" % (L - 20, 20)
            else:
                cigar = ""                              # mapped flag but no CIGAR: reference_end is None
            flag = rng.choice([0, 16, 0, 16, 99, 147, 83, 163, 1024, 1040, 4, 73, 133, 256, 2048, 2064])
            seq = rand_seq(rng, L, "ACGT" if rng.random() < 0.9 else "ACGTN")
            qual = None if rng.random() < 0.02 and (flag & 0x404 or L < 100) else rand_quals(rng, L)
            mapq = rng.choice([0, 0, 3, 20, 60, 60, 60, 60])
            reads.append(bam_writer.Read("r%d_%d_%d" % (tid, start, k), flag, tid, pos, mapq, cigar, seq, qual,
                                         tags=rng.choice([b"", b"NMC\x02", b"RGZgrp1\0NMC\x00", b"XSi\x05\0\0\0ZBBs\x02\0\0\0\x01\0\x02\0"])))
    reads.sort(key=lambda r: (r.tid, r.pos))
    for k in range(40):                                 # unplaced reads at the end of the file
        L = rng.choice([150, 151, 100])
        flag = rng.choice([4, 77, 141, 69, 133, 4 | 16, 77 | 0x100, 141 | 0x800, 4 | 0x40 | 0x80])
        reads.append(bam_writer.Read("u%d" % k, flag, -1, -1, 0, "", rand_seq(rng, L, "ACGTN" if k % 7 == 0 else "ACGT"),
                                     rand_quals(rng, L) if k % 5 else None))
    return reads


@pytest.fixture(scope="module", params=[0xff00, 1500, 211])
def sample(request, tmp_path_factory):
    d = tmp_path_factory.mktemp("bam%d" % request.param)
    path = str(d / "sample.bam")
    bam_writer.write_bam(path, REFS, make_reads(11 + request.param), block_size=request.param)
    names, lengths, records = bam_oracle.read_bam(path)
    return path, names, lengths, records


def same_record(got, want):
    assert got.query_name == want.query_name
    assert got.flag == want.flag and got.mapq == want.mapq
    assert got.reference_start == want.reference_start and got.reference_end == want.reference_end
    assert got.seq == want.seq
    assert got.query_qualities == want.query_qualities
    assert got.cigartuples == want.cigartuples
    assert got.is_unmapped == want.is_unmapped and got.is_duplicate == want.is_duplicate
    assert got.is_read2 == want.is_read2


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "advbam.h")).read()
    declared = set(re.findall(r"\b(advbam_[a-z_0-9]+)\s*\(", header))
    assert declared == set(bam_ingest.EXPORTS)
    lib = bam_ingest.load_library()
    for name in declared:
        assert hasattr(lib, name), name


def test_header_and_head(sample):
    path, names, lengths, records = sample
    with bam_ingest.AlignmentFile(path) as f:
        assert list(f.references) == names and list(f.lengths) == lengths
        assert f.get_tid("chr2") == 1 and f.get_tid("nope") == -1
        for got, want in zip(f.head(5), records[:5]):
            same_record(got, want)
        assert len(list(f.head(5))) == 5
        assert bam_ingest.median_head_read_length(f) == bam_oracle.head_read_length(records)
        assert bam_ingest.get_reference_genome_of_alignment_file(f) == "HG19"
        assert bam_ingest.chromosome_name_in(f, "chr1") == "chr1"


def test_fetch_matches_a_linear_scan(sample):
    path, names, lengths, records = sample
    rng = random.Random(5)
    regions = [(t, s, e) for t, s, e in LOCI]
    for t, s, e in LOCI:
        for _ in range(12):
            a = rng.randint(max(0, s - 400), e + 200)
            regions.append((t, a, a + rng.choice([1, 2, 50, 300, 20000, 400000])))
    regions += [(0, 0, 3000000), (1, 0, 150000000), (2, 0, 1), (1, 16384000, 16384001), (0, 2999999, 3000000)]
    total = 0
    with bam_ingest.AlignmentFile(path) as f:
        for t, a, b in regions:
            want = bam_oracle.fetch(records, t, a, b)
            got = list(f.fetch(names[t], a, b))
            assert [g.query_name for g in got] == [w.query_name for w in want], (t, a, b)
            for g, w in zip(got, want):
                same_record(g, w)
            total += len(got)
        assert total > 500
        with pytest.raises(ValueError):
            f.fetch("chrUn", 0, 10)


def test_scan_reads_every_record_with_any_thread_count(sample):
    path, names, lengths, records = sample
    with bam_ingest.AlignmentFile(path) as f:
        for threads in (1, 4):
            b = f.scan_batch(threads=threads)
            assert len(b) == len(records)
            for i in (0, 1, len(records) // 2, len(records) - 1):
                same_record(b[i], records[i])
            assert [b.name(i) for i in range(len(b))] == [r.query_name for r in records]
        dup = f.scan_batch(require_flags=0x400)
        assert [r.query_name for r in dup] == [r.query_name for r in records if r.flag & 0x400]


def test_illumina_selection_matches_the_reference_loop(sample):
    path, names, lengths, records = sample
    n_decoded = 0
    with bam_ingest.AlignmentFile(path) as f:
        for read_length in (150, 100):
            for t, s, e in LOCI:
                usable = [r for r in bam_oracle.fetch(records, t, s, e)]
                try:
                    want, want_bp = bam_oracle.select_illumina_mapped(records, t, s, e, read_length)
                except TypeError:                        # a record without qualities reaches the quality test
                    with pytest.raises(TypeError):
                        bam_ingest.select_mapped_illumina(f, names[t], s, e, read_length)
                    continue
                got = bam_ingest.select_mapped_illumina(f, names[t], s, e, read_length)
                assert got["n_fetched"] == len(usable)
                assert got["vntr_bp"] == want_bp
                assert got["names"] == [r.query_name for r, _ in want]
                assert got["mapq"].tolist() == [r.mapq for r, _ in want]
                assert got["reference_start"].tolist() == [r.reference_start for r, _ in want]
                seqs = ["".join("ACGT"[c] for c in got["codes"][got["off"][i]:got["off"][i + 1]])
                        for i in range(len(got["names"]))]
                assert seqs == [s_ for _, s_ in want]
                n_decoded += len(seqs)
    assert n_decoded > 20


def test_low_quality_rule_fuzz(tmp_path):
    """is_low_quality_read (utils.py:20-38) on many quality strings and read lengths (short reads make the
    reference's run limit collapse to 'any low-quality base')."""
    rng = random.Random(99)
    reads = []
    for k in range(1500):
        L = rng.choice([20, 39, 40, 41, 79, 80, 81, 100, 119, 120, 150, 151, 250])
        reads.append(bam_writer.Read("q%d" % k, 0, 0, 1000 + k // 50, rng.choice([0, 1, 60]), "%dM" % L,
                                     rand_seq(rng, L), rand_quals(rng, L)))
    path = str(tmp_path / "q.bam")
    bam_writer.write_bam(path, [("1", 100000)], reads)
    _, _, records = bam_oracle.read_bam(path)
    with bam_ingest.AlignmentFile(path) as f:
        assert bam_ingest.get_reference_genome_of_alignment_file(f) == "GRCh37"
        assert bam_ingest.chromosome_name_in(f, "chr1") == "1"
        b = f.fetch_batch("1", 900, 1200)
        decision, _ = b.select_illumina(900, 1200, 150, min_read_length=0)
        want = [bam_ingest.SKIP_LOW_QUALITY if bam_oracle.is_low_quality_read(r) else bam_ingest.DECODE for r in records]
        assert decision.tolist() == want
        assert 200 < sum(want) // bam_ingest.SKIP_LOW_QUALITY < 1300


def test_pacbio_spanning_segments(tmp_path):
    rng = random.Random(3)
    vntr_start, vntr_end = 50000, 50600
    reads = []
    for k in range(120):
        pos = rng.randint(vntr_start - 3000, vntr_start + 200)
        ops, ref, n_read = [], pos, 0
        if rng.random() < 0.4:
            c = rng.randint(1, 300)
            ops.append((4, c))
            n_read += c
        target = rng.choice([vntr_end + 20, vntr_end + 105, vntr_end + 3000, vntr_start + 100, vntr_start - 50])
        while ref < target or not any(op != 4 for op, _ in ops):
            m = rng.randint(1, 60)
            ops.append((rng.choice([0, 0, 0, 7, 8]), m))
            ref += m
            n_read += m
            r = rng.random()
            if r < 0.35:
                i = rng.randint(1, 12)
                ops.append((1, i))
                n_read += i
            elif r < 0.7:
                d = rng.randint(1, 40 if rng.random() < 0.9 else 400)
                ops.append((2, d))
                ref += d
        if ops[-1][0] in (1, 2):
            ops.append((0, 5))
            n_read += 5
        if rng.random() < 0.3:
            c = rng.randint(1, 50)
            ops.append((4, c))
            n_read += c
        flag = rng.choice([0, 16, 0, 16, 4, 256])
        reads.append(bam_writer.Read("p%d" % k, flag, 0, pos, 60, ops, rand_seq(rng, n_read), None,
                                     long_cigar=(k % 3 == 0), tags=b"NMC\x01" if k % 2 else b""))
    reads.sort(key=lambda r: r.pos)
    path = str(tmp_path / "p.bam")
    bam_writer.write_bam(path, [("chr7", 200000)], reads, block_size=4000)
    _, _, records = bam_oracle.read_bam(path)
    assert any(len(r.cigartuples) > 2 and r.query_name in ("p0", "p3") for r in records)    # CG tag resolved
    want = bam_oracle.pacbio_spanning_segments(records, 0, vntr_start, vntr_end)
    with bam_ingest.AlignmentFile(path) as f:
        got = bam_ingest.spanning_pacbio_segments(f, "chr7", vntr_start, vntr_end)
        for g, w in zip(f.fetch("chr7", vntr_start, vntr_end), bam_oracle.fetch(records, 0, vntr_start, vntr_end)):
            same_record(g, w)
            assert g.get_reference_positions(full_length=True) == w.get_reference_positions(full_length=True)
    assert got == want
    assert 10 < len(want) < 110


def test_unmapped_reads_as_the_samtools_pipeline_prints_them(sample):
    path, names, lengths, records = sample
    want = bam_oracle.unmapped_fasta_records(records)
    got_names, got_seqs = bam_ingest.extract_unmapped_reads(path, threads=3)
    assert list(zip(got_names, got_seqs)) == want
    assert any(n.endswith("/1") for n in got_names) and any(n.endswith("/2") for n in got_names)
    assert any("/" not in n for n in got_names)


def test_errors(tmp_path, sample):
    path = sample[0]
    with pytest.raises(IOError):
        bam_ingest.AlignmentFile(str(tmp_path / "missing.bam"))
    with pytest.raises(ValueError):
        bam_ingest.AlignmentFile(str(tmp_path / "reads.sam"))
    junk = tmp_path / "junk.bam"
    junk.write_bytes(b"this is not a BGZF file at all, not even close" * 10)
    with pytest.raises(ValueError):
        bam_ingest.AlignmentFile(str(junk))
    data = open(path, "rb").read()
    # no index next to the file: opening works, scanning works, fetching says why it cannot
    lone = tmp_path / "lone.bam"
    lone.write_bytes(data)
    with bam_ingest.AlignmentFile(str(lone)) as f:
        assert len(f.scan_batch()) == len(sample[3])
        with pytest.raises(ValueError, match="index"):
            f.fetch_batch("chr1", 0, 100)
    # an index of another file
    other = tmp_path / "other.bam"
    bam_writer.write_bam(str(other), [("chr1", 1000)], [])
    with pytest.raises(ValueError, match="references"):
        bam_ingest.AlignmentFile(str(lone), index_filename=str(other) + ".bai")
    # index found under the samtools-style name <stem>.bai
    (tmp_path / "lone.bai").write_bytes(open(path + ".bai", "rb").read())
    with bam_ingest.AlignmentFile(str(lone)) as f:
        assert len(f.fetch_batch("chr1", 40000, 40090)) == len(bam_oracle.fetch(sample[3], 0, 40000, 40090))
    # a flipped byte inside a compressed block is caught by the block checksum
    broken = bytearray(data)
    broken[len(broken) // 2] ^= 0x55
    bad = tmp_path / "broken.bam"
    bad.write_bytes(bytes(broken))
    with bam_ingest.AlignmentFile(str(bad)) as f:
        with pytest.raises(ValueError):
            f.scan_batch()
    # truncated in the middle of a block
    cut = tmp_path / "cut.bam"
    cut.write_bytes(data[:len(data) // 2])
    with bam_ingest.AlignmentFile(str(cut)) as f:
        with pytest.raises(ValueError):
            f.scan_batch()


def test_empty_file_and_empty_regions(tmp_path):
    path = str(tmp_path / "empty.bam")
    bam_writer.write_bam(path, REFS, [])
    with bam_ingest.AlignmentFile(path) as f:
        assert len(f.scan_batch()) == 0 and len(f.fetch_batch("chr1", 0, 1000)) == 0
        assert len(list(f.head(5))) == 0
        with pytest.raises(IndexError):
            bam_ingest.median_head_read_length(f)
        out = bam_ingest.select_mapped_illumina(f, "chr1", 100, 200, 150)
        assert out["names"] == [] and out["off"].tolist() == [0] and len(out["codes"]) == 0
    assert bam_ingest.extract_unmapped_reads(path) == ([], [])


def test_records_the_reference_would_crash_on_raise_the_same_exceptions(tmp_path):
    good = [30] * 150
    reads = [bam_writer.Read("ok", 0, 0, 900, 60, "150M", "ACGT" * 37 + "AC", good),
             bam_writer.Read("iupac", 0, 0, 910, 60, "150M", "ACGT" * 37 + "AR", good)]
    path = str(tmp_path / "a.bam")
    bam_writer.write_bam(path, [("chr1", 100000)], reads)
    with bam_ingest.AlignmentFile(path) as f:
        with pytest.raises(ValueError, match="iupac"):      # hmm.viterbi: symbol not in the alphabet
            bam_ingest.select_mapped_illumina(f, "chr1", 1000, 1040, 150)
    reads[1] = bam_writer.Read("noqual", 0, 0, 910, 60, "150M", "ACGT" * 37 + "AC", None)
    bam_writer.write_bam(path, [("chr1", 100000)], reads)
    with bam_ingest.AlignmentFile(path) as f:
        with pytest.raises(TypeError, match="noqual"):      # utils.py:24 enumerates None
            bam_ingest.select_mapped_illumina(f, "chr1", 1000, 1040, 150)
        b = f.fetch_batch("chr1", 1000, 1040)
        assert b[1].query_qualities is None and b[0].query_qualities == good
    _, _, records = bam_oracle.read_bam(path)
    with pytest.raises(TypeError):
        bam_oracle.select_illumina_mapped(records, 0, 1000, 1040, 150)
    # a read with mapping quality 0 is dropped before its missing qualities are looked at (utils.py:21)
    reads[1] = bam_writer.Read("noqual", 0, 0, 910, 0, "150M", "ACGT" * 37 + "AC", None)
    bam_writer.write_bam(path, [("chr1", 100000)], reads)
    with bam_ingest.AlignmentFile(path) as f:
        assert bam_ingest.select_mapped_illumina(f, "chr1", 1000, 1040, 150)["names"] == ["ok"]


@pytest.mark.gpu
def test_genotypes_from_an_alignment_file_equal_genotypes_from_read_lists(tmp_path):
    """BAM -> libadvbam -> code arrays -> device equals oracle ingest -> strings -> device, and the
    simulated alleles come out."""
    from test_pipeline import _sample
    from advntr_b200 import pipeline, synth
    loci, mapped, names, seqs, truth = _sample(n_loci=10, seed=8)
    rng = random.Random(4)
    specs, reads, genome_len = [], [], 0
    for lid, left, right, segs in loci:
        base = 10000 * lid
        start = base + len(left)
        specs.append(pipeline.LocusSpec(lid, left, right, segs, chromosome="chr1", start_point=start))
        for k, read in enumerate(mapped[lid]):
            # where the read starts does not matter for the decode, only for the region test: spread them
            pos = rng.randint(start - 200, start + sum(map(len, segs)) + 40)
            kind = rng.random()
            flag, mapq, qual, seq = rng.choice([0, 16, 99, 147]), 60, [rng.randint(28, 40) for _ in read], read
            if kind < 0.05:
                flag |= 0x400
            elif kind < 0.1:
                mapq = 0
            elif kind < 0.15:
                qual = [rng.randint(2, 15) if 40 <= i < 46 else q for i, q in enumerate(qual)]
            elif kind < 0.2:
                seq = seq[:70] + "N" + seq[71:]
            reads.append(bam_writer.Read("m%d_%d" % (lid, k), flag, 0, pos, mapq, "%dM" % len(seq), seq, qual))
        genome_len = base + 20000
    reads.sort(key=lambda r: r.pos)
    for k, (n, s) in enumerate(zip(names, seqs)):
        flag = rng.choice([4, 77, 141, 4 | 16])
        stored = synth.revcomp(s) if flag & 16 else s           # bam2fq turns reverse-flagged records back
        reads.append(bam_writer.Read(n, flag, -1, -1, 0, "", stored, [30] * len(s)))
    path = str(tmp_path / "sample.bam")
    bam_writer.write_bam(path, [("chr1", genome_len)], reads, block_size=20000)

    run = pipeline.GenotypingRun.from_alignment_file(specs, path)
    assert run.read_length == 150
    got = run.genotype_alignment_file(path)

    _, _, records = bam_oracle.read_bam(path)
    want_mapped, bp = {}, {}
    for spec in specs:
        end = spec.start_point + sum(map(len, spec.repeat_segments))
        sel, bp[spec.id] = bam_oracle.select_illumina_mapped(records, 0, spec.start_point, end, 150)
        want_mapped[spec.id] = [s for _, s in sel]
    un = bam_oracle.unmapped_fasta_records(records)
    want = run.genotype(want_mapped, [n for n, _ in un], [s for _, s in un])
    dropped = sum(len(mapped[l]) - len(want_mapped[l]) for l in mapped)
    assert dropped > 20                                         # the read-level tests did reject reads
    for lid in want:
        assert got[lid].pop("vntr_bp_in_mapped_reads") == bp[lid]
        assert got[lid] == want[lid], lid
    right_calls = sum(1 for lid in truth if got[lid]["copy_numbers"] is not None and
                      sorted(got[lid]["copy_numbers"]) == truth[lid])
    assert right_calls >= len(truth) - 3
    run.close()


@pytest.mark.gpu
def test_pacbio_genotype_from_an_alignment_file(tmp_path):
    """Long reads aligned with indels: BAM -> CIGAR walk (native) -> long-read kernel -> genotype, equal to
    the oracle's cut of the same records fed as read lists, and the simulated alleles come out."""
    from advntr_b200 import locus_batch, synth
    rng = random.Random(12)
    R, nref = 30, 8
    ru = synth.rand_dna(rng, R)
    left, right = synth.rand_dna(rng, 600), synth.rand_dna(rng, 600)
    base = 20000
    vntr_start = base + len(left)
    alleles = (6, 11)
    reads = []
    for k in range(24):
        copies = alleles[k % 2]
        # the read covers the whole allele; its alignment to the reference (nref copies) puts the copy
        # number difference into one insertion or deletion inside the repeat
        seq = left + ru * copies + right
        noisy = synth.sequencing_errors(rng, seq, 0.01, 0.0, 0.0)
        if len(noisy) != len(seq):
            continue
        half = len(left) + R * min(copies, nref) // 2
        if copies >= nref:
            ops = [(0, half), (1, R * (copies - nref)), (0, len(seq) - half - R * (copies - nref))]
        else:
            ops = [(0, half), (2, R * (nref - copies)), (0, len(seq) - half)]
        ops = [(op, n) for op, n in ops if n]
        reads.append(bam_writer.Read("pb%d" % k, rng.choice([0, 16]), 0, base, 60, ops, noisy, None))
    path = str(tmp_path / "pb.bam")
    bam_writer.write_bam(path, [("chr1", 100000)], reads)
    got = locus_batch.repeat_count_from_pacbio_alignment_file(path, "chr1", vntr_start, left, right, [ru] * nref,
                                                              error_rate=0.05)
    _, _, records = bam_oracle.read_bam(path)
    cut = [s for _, s, _ in bam_oracle.pacbio_spanning_segments(records, 0, vntr_start, vntr_start + R * nref)]
    assert len(cut) == len(reads) and all(len(s) > 200 for s in cut)
    want = locus_batch.dominant_copy_numbers_from_spanning_reads(left, right, [ru] * nref, cut, error_rate=0.05)
    assert (got["copy_numbers"], got["maximum_likelihood"], got["observed_repeats"]) == want
    assert got["spanning_reads_count"] == len(reads)
    assert sorted(got["copy_numbers"]) == sorted(alleles)


def test_damaged_records_are_refused_not_crashed_on(tmp_path):
    """Random damage INSIDE valid BGZF blocks (checksums recomputed, so only the record parser can notice):
    every call either works or raises ValueError; the process survives."""
    import struct
    import zlib
    reads = make_reads(5, n_per_locus=12)
    good = str(tmp_path / "good.bam")
    bam_writer.write_bam(good, REFS, reads, block_size=3000)
    data = open(good, "rb").read()
    index = open(good + ".bai", "rb").read()
    blocks, q = [], 0
    while q < len(data):
        size = struct.unpack_from("<H", data, q + 16)[0] + 1
        blocks.append(zlib.decompressobj(-15).decompress(data[q + 18:q + size - 8]))
        q += size
    rng = random.Random(77)
    outcomes = {"ok": 0, "refused": 0}
    for trial in range(60):
        hurt = [bytearray(b) for b in blocks]
        for _ in range(rng.randint(1, 4)):
            b = rng.randrange(1, len(hurt) - 1)                  # keep the header block and the end marker
            if not hurt[b]:
                continue
            k = rng.randrange(len(hurt[b]))
            hurt[b][k] = rng.choice([0, 0xff, hurt[b][k] ^ (1 << rng.randrange(8)), rng.randrange(256)])
        path = str(tmp_path / ("hurt%d.bam" % trial))
        with open(path, "wb") as fh:
            for b in hurt:
                co = zlib.compressobj(6, zlib.DEFLATED, -15)
                comp = co.compress(bytes(b)) + co.flush()
                fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(comp) + 25) + comp +
                         struct.pack("<II", zlib.crc32(bytes(b)), len(b)))
        with open(path + ".bai", "wb") as fh:
            fh.write(index)                                      # offsets stay valid only if sizes did; fine either way
        try:
            with bam_ingest.AlignmentFile(path) as f:
                for threads in (1, 3):
                    b = f.scan_batch(threads=threads)
                    [r.query_name for r in b]
                for t, s, e in LOCI:
                    fb = f.fetch_batch(REFS[t][0], s - 300, e + 300)
                    fb.spanning_segments(s, e)
                    try:
                        fb.select_illumina(s, e, 150)
                    except TypeError:
                        pass
            outcomes["ok"] += 1
        except ValueError:
            outcomes["refused"] += 1
        except UnicodeDecodeError:
            outcomes["ok"] += 1                                  # a damaged name: parsed, just not text
    assert outcomes["ok"] + outcomes["refused"] == 60 and outcomes["refused"] > 0


def test_committed_bam_fixture():
    """tests/golden/tiny.bam (+ .bai) with the answers the plain-Python reader gave when it was made
    (tests/golden/make_golden_bam.py): the byte-level format is pinned independently of today's writer."""
    import json
    from conftest import GOLDEN
    want = json.load(open(os.path.join(GOLDEN, "tiny_bam_expected.json")))
    with bam_ingest.AlignmentFile(os.path.join(GOLDEN, "tiny.bam")) as f:
        assert list(f.references) == want["references"] and list(f.lengths) == want["lengths"]
        assert len(f.scan_batch()) == want["n_records"]
        got_head = [[r.query_name, r.flag, r.reference_start, r.reference_end, r.seq, r.mapq] for r in f.head(5)]
        assert got_head == want["head"]
        for case in want["fetch"]:
            t, s, e = case["region"]
            assert [r.query_name for r in f.fetch(want["references"][t], s, e)] == case["names"]
        for case in want["select"]:
            t, s, e = case["region"]
            got = bam_ingest.select_mapped_illumina(f, want["references"][t], s, e, 150)
            assert got["names"] == case["names"] and got["vntr_bp"] == case["vntr_bp"]
            seqs = ["".join("ACGT"[c] for c in got["codes"][got["off"][i]:got["off"][i + 1]]) for i in range(len(got["names"]))]
            assert seqs == case["sequences"]
        names, seqs = bam_ingest.extract_unmapped_reads(f)
        assert [list(x) for x in zip(names, seqs)] == want["unmapped"]


@pytest.mark.gpu
def test_one_locus_selection_and_frameshift_from_an_alignment_file(tmp_path):
    """LocusDecoder.select_reads_from_alignment_file == select_reads on the reads the oracle's loop hands
    over, with the mapped reads' names / mapq / positions attached."""
    from test_pipeline import _sample
    from advntr_b200 import locus_batch
    loci, mapped, names, seqs, truth = _sample(n_loci=2, seed=21)
    lid, left, right, segs = loci[0]
    rng = random.Random(6)
    start = 5000 + len(left)
    reads = []
    for k, read in enumerate(mapped[lid]):
        pos = rng.randint(start - 140, start + sum(map(len, segs)) - 5)
        reads.append(bam_writer.Read("m%d" % k, rng.choice([0, 16]), 0, pos, rng.choice([60, 60, 60, 0]), "150M", read,
                                     [rng.randint(25, 40) for _ in read]))
    reads.sort(key=lambda r: r.pos)
    path = str(tmp_path / "one.bam")
    bam_writer.write_bam(path, [("chr1", 50000)], reads)
    dec = locus_batch.LocusDecoder(left, right, segs, read_length=150, locus_id=lid)
    unm = [s for s in seqs[:40]]
    got = dec.select_reads_from_alignment_file(path, "chr1", start, unm)
    _, _, records = bam_oracle.read_bam(path)
    sel, _ = bam_oracle.select_illumina_mapped(records, 0, start, start + sum(map(len, segs)), 150)
    want = dec.select_reads([s for _, s in sel], unm)
    assert [(r.sequence, r.logp, r.is_mapped) for r in got] == [(r.sequence, r.logp, r.is_mapped) for r in want]
    by_name = {r.query_name: r for r, _ in sel}
    n_mapped = 0
    for r in got:
        if r.is_mapped:
            n_mapped += 1
            src = by_name[r.query_name]
            assert src.seq == r.sequence and src.mapq == r.mapq and src.reference_start == r.reference_start
            assert r.mapq > 0
    assert n_mapped > 10
    assert dec.frameshift_candidate(got) == dec.frameshift_candidate(want)
