"""Ingest throughput of libadvbam.so on a synthetic BAM (written by tests/bam_writer.py).

    python tools/kbench_bam.py [n_reads=300000]

Prints whole-file scan rates for 1 thread and for all cores, the unmapped-read extraction, and region
fetch + read-level selection per locus.
"""
import os
import random
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bam_writer  # noqa: E402
from advntr_b200 import bam_ingest, build  # noqa: E402

build.build_bam_library()
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
rng = random.Random(1)
genome = 50000000
t0 = time.time()
pool = ["".join(rng.choice("ACGT") for _ in range(150)) for _ in range(512)]
quals = [[rng.randint(25, 40) for _ in range(150)] for _ in range(64)]
pos = sorted(rng.randrange(0, genome - 200) for _ in range(n_reads))
reads = [bam_writer.Read("read%09d" % i, rng.choice([99, 147, 83, 163]), 0, p, 60, "150M", pool[i & 511], quals[i & 63],
                         tags=b"NMC\x01RGZgrp1\0") for i, p in enumerate(pos)]
reads += [bam_writer.Read("un%09d" % i, rng.choice([77, 141]), -1, -1, 0, "", pool[i & 511], quals[i & 63])
          for i in range(n_reads // 20)]
path = os.path.join(tempfile.mkdtemp(), "bench.bam")
bam_writer.write_bam(path, [("chr1", genome)], reads)
size = os.path.getsize(path)
print("wrote %d records, %.1f MB compressed, in %.1f s (python writer)" % (len(reads), size / 1e6, time.time() - t0))

f = bam_ingest.AlignmentFile(path)
for threads in (1, 0):
    best = 1e9
    for _ in range(3):
        t = time.time()
        b = f.scan_batch(threads=threads)
        best = min(best, time.time() - t)
        n, raw = len(b), int(b.seq_off[-1])
        b.close()
    print("scan, %s: %.3f s  %.2f M records/s  %.0f MB/s compressed" % (
        "1 thread" if threads else "%d threads" % (os.cpu_count() or 1), best, n / best / 1e6, size / best / 1e6))
t = time.time()
names, seqs = bam_ingest.extract_unmapped_reads(f)
print("unmapped extraction (scan -f4 -F0x900 + fastq orientation + python strings): %.3f s for %d reads" % (time.time() - t, len(names)))
loci = [rng.randrange(1000, genome - 1000) for _ in range(2000)]
t = time.time()
total = 0
for s in loci:
    total += len(bam_ingest.select_mapped_illumina(f, "chr1", s, s + 60, 150)["names"])
dt = time.time() - t
print("region fetch + selection: %d loci in %.3f s (%.0f loci/s, %d reads to decode)" % (len(loci), dt, len(loci) / dt, total))
