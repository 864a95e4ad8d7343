"""End-to-end demo on synthetic data: what `advntr genotype` does for Illumina reads, minus BAM IO.

    python tools/demo_genotype.py [n_loci=24] [coverage=30] [--bam]

For every locus a diploid sample is simulated (two alleles with their own repeat counts, 150 bp reads
with sequencing errors; a fifth of the reads is "unmapped" and arrives on either strand, mixed with
random decoys).  Then, as genome_analyzer.py:273-297 does:

  1. the unmapped reads are filtered against the keywords of all loci (device keyword filter);
  2. mapped + filtered reads of all loci are decoded in ONE device call (banded Viterbi + on-device
     path reducers);
  3. every locus is genotyped from the repeat counts of its recruited reads.

Prints one line per locus (id, pattern length, truth, call, likelihood, reads used) and a summary.

With ``--bam`` the sample is first written as a coordinate-sorted BAM + BAI (tests/bam_writer.py: mapped
reads at their loci on a synthetic chromosome, unmapped reads at the end of the file) and genotyped
from that file: region fetches, read-level tests and unmapped-read extraction in libadvbam
(``GenotypingRun.genotype_alignment_file``).
"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from advntr_b200 import pipeline, synth

use_bam = "--bam" in sys.argv
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
n_loci = int(argv[0]) if len(argv) > 0 else 24
coverage = int(argv[1]) if len(argv) > 1 else 30
rng = random.Random(2026)

loci, mapped, names, seqs, truth = [], {}, [], [], {}
for lid in range(1, n_loci + 1):
    R = rng.choice((6, 9, 12, 17, 24, 31, 40, 52))
    ru = synth.rand_dna(rng, R)
    left, right = synth.rand_dna(rng, 300), synth.rand_dna(rng, 300)
    top = max(2, 110 // R)
    a = rng.randint(2, top)
    b = a if rng.random() < 0.4 else rng.randint(2, top)
    truth[lid] = tuple(sorted((a, b)))
    loci.append(pipeline.LocusSpec(lid, left, right, [ru] * max(2, 100 // R), chromosome="chr1",
                                   start_point=10000 * lid + len(left)))
    mapped[lid] = []
    for copies in (a, b):
        allele = left + ru * copies + right
        for _ in range(int(round((R * copies + 150) * coverage / 2 / 150.0))):
            s = rng.randrange(300 - 149, 300 + R * copies - 1)
            read = synth.sequencing_errors(rng, allele[s:s + 158], 0.004, 0.0003, 0.0003)[:150]
            if len(read) < 150:
                continue
            if rng.random() < 0.2:
                names.append("u%06d" % len(names))
                seqs.append(synth.revcomp(read) if rng.random() < 0.5 else read)
            else:
                mapped[lid].append(read)
for _ in range(20 * n_loci):
    names.append("u%06d" % len(names))
    seqs.append(synth.rand_dna(rng, 150))

bam_path = None
if use_bam:
    import tempfile
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import bam_writer
    records = []
    for spec in loci:
        end = spec.start_point + sum(map(len, spec.repeat_segments))
        for k, read in enumerate(mapped[spec.id]):
            pos = rng.randint(spec.start_point - 140, end - 5)       # placement only matters for the region test
            records.append(bam_writer.Read("m%d_%d" % (spec.id, k), rng.choice([0, 16]), 0, pos, 60, "%dM" % len(read),
                                           read, [rng.randint(25, 40) for _ in read]))
    records.sort(key=lambda r: r.pos)
    records += [bam_writer.Read(n, 4, -1, -1, 0, "", s, [30] * len(s)) for n, s in zip(names, seqs)]
    bam_path = os.path.join(tempfile.mkdtemp(), "sample.bam")
    bam_writer.write_bam(bam_path, [("chr1", 10000 * (n_loci + 2))], records)
    print("wrote %s (%d records, %.1f MB)" % (bam_path, len(records), os.path.getsize(bam_path) / 1e6))

t0 = time.time()
run = pipeline.GenotypingRun.from_alignment_file(loci, bam_path) if use_bam else pipeline.GenotypingRun(loci)
t1 = time.time()
calls = run.genotype_alignment_file(bam_path) if use_bam else run.genotype(mapped, names, seqs)
t2 = time.time()
right_calls = 0
for spec in loci:
    c = calls[spec.id]
    call = tuple(sorted(c["copy_numbers"])) if c["copy_numbers"] is not None else None
    right_calls += call == truth[spec.id]
    print("locus %3d  RU %2d bp  truth %-8s call %-8s p=%.6f  recruited %3d (spanning %2d, flanking %2d)%s" % (
        spec.id, len(spec.pattern), "%d/%d" % truth[spec.id], "None" if call is None else "%d/%d" % call,
        c["maximum_likelihood"], c["recruited_reads_count"], c["spanning_reads_count"], c["flanking_reads_count"],
        "" if call == truth[spec.id] else "   <-- differs"))
n_reads = sum(len(v) for v in mapped.values()) + len(seqs)
print("%d loci, %d reads: models + keyword tables %.2f s, filter + decode + genotype %.3f s; %d / %d calls equal the "
      "simulated genotype" % (n_loci, n_reads, t1 - t0, t2 - t1, right_calls, n_loci))
run.close()
