#!/usr/bin/env python
"""A small BAM + BAI fixture with the answers of the plain-Python reader (oracle/bam_oracle.py).

    python tests/golden/make_golden_bam.py

Writes tests/golden/tiny.bam, tiny.bam.bai and tiny_bam_expected.json.  The file is produced by
tests/bam_writer.py (no samtools / pysam in this image); committing it pins the byte-level format the
library reads, so that writer and reader cannot drift together unnoticed.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import bam_oracle   # noqa: E402
import bam_writer   # noqa: E402
import test_bam_ingest as T   # noqa: E402


def main():
    path = os.path.join(HERE, "tiny.bam")
    reads = T.make_reads(2026, n_per_locus=14)
    bam_writer.write_bam(path, T.REFS, reads, block_size=2500)
    names, lengths, records = bam_oracle.read_bam(path)
    rng = random.Random(1)
    regions = [list(l) for l in T.LOCI] + [[0, 0, 3000000], [1, 16380000, 16390000]]
    expected = {"references": names, "lengths": lengths, "n_records": len(records),
                "head": [[r.query_name, r.flag, r.reference_start, r.reference_end, r.seq, r.mapq] for r in records[:5]],
                "fetch": [], "select": [], "unmapped": bam_oracle.unmapped_fasta_records(records)}
    for t, s, e in regions:
        expected["fetch"].append({"region": [t, s, e], "names": [r.query_name for r in bam_oracle.fetch(records, t, s, e)]})
    for t, s, e in T.LOCI:
        sel, bp = bam_oracle.select_illumina_mapped(records, t, s, e, 150)
        expected["select"].append({"region": [t, s, e], "names": [r.query_name for r, _ in sel],
                                   "sequences": [q for _, q in sel], "vntr_bp": bp})
    with open(os.path.join(HERE, "tiny_bam_expected.json"), "w") as fh:
        json.dump(expected, fh, indent=1)
    print("%d records, %d bytes; %d regions" % (len(records), os.path.getsize(path), len(regions)))


if __name__ == "__main__":
    main()
