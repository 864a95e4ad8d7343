"""TEST INFRASTRUCTURE -- plain-Python restatement of the read-ingest path.

Only tests/ may import this module.  The product path is ``advntr_b200/csrc/bam_ingest.cpp``
(``libadvbam.so``); nothing in the package falls back to this code.

PARITY UNPINNED against pysam / samtools: neither is installed in this image and the reference ships
no alignment fixtures, so this oracle is anchored on the published formats and documented semantics
instead of on outputs of the reference's own dependencies:

* BGZF / BAM records: hts-specs SAMv1 section 4 and 4.1 (every block decoded with Python's ``zlib``,
  the whole file read linearly -- the index is NOT used here, so index handling of the product is
  checked against a plain scan);
* ``fetch(reference, start, end)``: records with ``pos < end`` and ``bam_endpos > start`` (htslib
  ``bam_endpos``: unmapped or CIGAR-less records span one base);
* ``AlignedSegment`` attributes as pysam documents them: ``reference_end`` is None for unmapped /
  CIGAR-less records, ``query_qualities`` None when the record stores 0xff, ``seq`` None when empty,
  ``get_reference_positions(full_length)``;
* the loops of the reference restated literally on those records: ``select_illumina_reads``
  (``/root/reference/advntr/vntr_finder.py:714-753``), ``is_low_quality_read``
  (``/root/reference/advntr/utils.py:20-38``), ``check_if_pacbio_mapped_read_spans_vntr``
  (``vntr_finder.py:373-420``), and the shell pipeline of
  ``extract_unmapped_reads_to_fasta_file`` (``sam_utils.py:9-23``: ``samtools view -f4``, then
  ``samtools bam2fq`` with its default ``-F 0x900`` and ``/1`` ``/2`` name suffixes, reverse-strand
  records reverse-complemented).
"""
import struct
import zlib

NIBBLE = "=ACMGRSVTWYHKDBN"
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "M": "K", "K": "M", "R": "Y", "Y": "R", "S": "S", "W": "W",
         "V": "B", "B": "V", "H": "D", "D": "H", "N": "N", "=": "="}

MAPQ_CUTOFF = 0                       # settings.py:26
QUALITY_SCORE_CUTOFF = 20             # settings.py:24
LOW_QUALITY_BP_TO_DISCARD_READ = 0.10  # settings.py:25


class Record(object):
    """The attributes of ``pysam.AlignedSegment`` the reference touches."""

    def __init__(self, raw):
        tid, pos, l_name, mapq, _bin, n_cigar, flag, l_seq, _ntid, _npos, _tlen = struct.unpack_from("<iiBBHHHiiii", raw, 0)
        q = 32
        self.query_name = self.qname = raw[q:q + l_name].split(b"\0")[0].decode()
        q += l_name
        cigar = [struct.unpack_from("<I", raw, q + 4 * i)[0] for i in range(n_cigar)]
        q += 4 * n_cigar
        packed = raw[q:q + (l_seq + 1) // 2]
        q += (l_seq + 1) // 2
        seq = "".join(NIBBLE[(packed[i >> 1] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        qual = raw[q:q + l_seq]
        q += l_seq
        self.tags = raw[q:]
        if n_cigar == 2 and cigar[0] & 15 == 4 and cigar[0] >> 4 == l_seq and cigar[1] & 15 == 3:
            long_cigar = self._long_cigar()
            if long_cigar is not None:
                cigar = long_cigar
        self.cigartuples = [(c & 15, c >> 4) for c in cigar]
        self.flag, self.tid, self.reference_start, self.pos = flag, tid, pos, pos
        self.mapq = self.mapping_quality = mapq
        self.seq = self.query_sequence = seq if l_seq else None
        self.query_qualities = None if (l_seq == 0 or qual[0] == 0xff) else list(qual)
        self.is_unmapped = bool(flag & 0x4)
        self.is_reverse = bool(flag & 0x10)
        self.is_read1 = bool(flag & 0x40)
        self.is_read2 = bool(flag & 0x80)
        self.is_secondary = bool(flag & 0x100)
        self.is_duplicate = bool(flag & 0x400)
        self.is_supplementary = bool(flag & 0x800)
        rlen = sum(n for op, n in self.cigartuples if op in (0, 2, 3, 7, 8))
        self.reference_end = None if (self.is_unmapped or not self.cigartuples) else pos + (rlen or 1)
        self.endpos = pos + ((0 if self.is_unmapped else rlen) or 1)      # htslib bam_endpos

    def _long_cigar(self):
        t, q = self.tags, 0
        while q + 3 <= len(t):
            tag, ty = t[q:q + 2], chr(t[q + 2])
            q += 3
            if ty in "AcC":
                q += 1
            elif ty in "sS":
                q += 2
            elif ty in "iIf":
                q += 4
            elif ty in "ZH":
                q = t.index(b"\0", q) + 1
            elif ty == "B":
                sub, cnt = chr(t[q]), struct.unpack_from("<I", t, q + 1)[0]
                w = 1 if sub in "cC" else 2 if sub in "sS" else 4
                if tag == b"CG" and sub == "I":
                    return list(struct.unpack_from("<%dI" % cnt, t, q + 5))
                q += 5 + cnt * w
            else:
                return None
        return None

    def get_reference_positions(self, full_length=False):
        out, ref = [], self.reference_start
        if self.is_unmapped:
            return out
        for op, n in self.cigartuples:
            if op in (0, 7, 8):
                out.extend(range(ref, ref + n))
                ref += n
            elif op in (1, 4):
                if full_length:
                    out.extend([None] * n)
            elif op in (2, 3):
                ref += n
        return out


def read_bam(path):
    """-> (reference names, reference lengths, [Record ...]) by a linear pass over every BGZF block."""
    data = open(path, "rb").read()
    plain, q = bytearray(), 0
    while q < len(data):
        assert data[q:q + 4] == b"\x1f\x8b\x08\x04", "not a BGZF block"
        xlen = struct.unpack_from("<H", data, q + 10)[0]
        x, bsize = q + 12, None
        while x < q + 12 + xlen:
            si1, si2, slen = struct.unpack_from("<BBH", data, x)
            if (si1, si2, slen) == (66, 67, 2):
                bsize = struct.unpack_from("<H", data, x + 4)[0] + 1
            x += 4 + slen
        body = zlib.decompressobj(-15).decompress(data[q + 12 + xlen:q + bsize - 8])
        crc, isize = struct.unpack_from("<II", data, q + bsize - 8)
        assert len(body) == isize and zlib.crc32(body) == crc
        plain += body
        q += bsize
    plain = bytes(plain)
    assert plain[:4] == b"BAM\1"
    l_text = struct.unpack_from("<i", plain, 4)[0]
    q = 8 + l_text
    n_ref = struct.unpack_from("<i", plain, q)[0]
    q += 4
    names, lengths = [], []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", plain, q)[0]
        names.append(plain[q + 4:q + 4 + l_name - 1].decode())
        lengths.append(struct.unpack_from("<i", plain, q + 4 + l_name)[0])
        q += 8 + l_name
    records = []
    while q < len(plain):
        size = struct.unpack_from("<i", plain, q)[0]
        records.append(Record(plain[q + 4:q + 4 + size]))
        q += 4 + size
    return names, lengths, records


def fetch(records, tid, start, end):
    return [r for r in records if r.tid == tid and r.pos < end and r.endpos > start]


# utils.py:20-38, literally
def is_low_quality_read(read):
    if read.mapq <= MAPQ_CUTOFF:
        return True
    low_quality_base_pairs = [i for i, q in enumerate(read.query_qualities) if q < QUALITY_SCORE_CUTOFF]
    if len(low_quality_base_pairs) >= LOW_QUALITY_BP_TO_DISCARD_READ * len(read.query_qualities):
        return True
    maximum_low_quality_run = int(LOW_QUALITY_BP_TO_DISCARD_READ * len(read.query_qualities) / 4)
    for i in low_quality_base_pairs:
        passed = False
        for j in range(i + 1, i + maximum_low_quality_run):
            if j not in low_quality_base_pairs:
                passed = True
                break
        if not passed:
            return True
    return False


def head_read_length(records):
    """vntr_finder.py:714-718."""
    lengths = [len(r.seq) for r in records[:5]]
    return sorted(lengths)[len(lengths) // 2]


def select_illumina_mapped(records, tid, vntr_start, vntr_end, read_length, min_read_length=None):
    """The mapped-read loop of ``select_illumina_reads`` (``vntr_finder.py:727-753``) up to the Viterbi
    call: -> ([(record, sequence) that the reference decodes and does not drop for quality], vntr_bp)."""
    if min_read_length is None:
        min_read_length = int(read_length * 0.9)
    out, vntr_bp = [], 0
    for read in fetch(records, tid, vntr_start, vntr_end):
        if read.is_unmapped or read.is_duplicate:
            continue
        if len(read.seq) < min_read_length:
            continue
        read_end = read.reference_end if read.reference_end else read.reference_start + len(read.seq)
        if vntr_start - read_length < read.reference_start < vntr_end or vntr_start < read_end < vntr_end:
            if read.seq.count('N') <= 0:
                sequence = str(read.seq).upper()
                if not is_low_quality_read(read):
                    out.append((read, sequence))
            end = min(read_end, vntr_end)
            start = max(read.reference_start, vntr_start)
            vntr_bp += end - start
    return out, vntr_bp


def pacbio_spanning_segments(records, tid, vntr_start, vntr_end):
    """``get_spanning_reads_of_aligned_pacbio_reads`` + ``check_if_pacbio_mapped_read_spans_vntr``
    (``vntr_finder.py:441-420``): -> [(query_name, sequence, length-distribution entry)]."""
    hmm_flanking_region_size = 100
    min_flanking_bp = 10
    region_start = vntr_start - hmm_flanking_region_size
    out = []
    for read in fetch(records, tid, vntr_start, vntr_end):
        if len(read.get_reference_positions()) == 0:
            continue
        first_aligned_position = read.get_reference_positions()[0]
        last_aligned_position = read.get_reference_positions()[-1]
        if first_aligned_position <= vntr_start - min_flanking_bp and vntr_end + min_flanking_bp < last_aligned_position:
            read_region_start = None
            read_region_end = None
            left_flanking_bp = 0
            right_flanking_bp = 0
            for read_pos, ref_pos in enumerate(read.get_reference_positions(full_length=True)):
                if ref_pos is None:
                    continue
                if ref_pos > vntr_end + hmm_flanking_region_size:
                    break
                if region_start <= ref_pos < vntr_end + hmm_flanking_region_size:
                    if region_start <= ref_pos < vntr_start:
                        if read_region_start is None:
                            read_region_start = read_pos
                        left_flanking_bp += 1
                    elif vntr_start <= ref_pos < vntr_end:
                        pass
                    else:
                        if read_region_end is None:
                            read_region_end = read_pos
                        right_flanking_bp += 1
            if left_flanking_bp < min_flanking_bp or right_flanking_bp < min_flanking_bp:
                continue
            if read_region_start is not None and read_region_end is not None and read.seq is not None:
                result_seq = read.seq[read_region_start: read_region_end + right_flanking_bp]
                out.append((read.query_name, result_seq, len(result_seq) - left_flanking_bp - right_flanking_bp))
    return out


def unmapped_fasta_records(records):
    """What ``extract_unmapped_reads_to_fasta_file`` leaves in the fasta file, as (name, sequence)."""
    out = []
    for r in records:
        if not r.flag & 0x4 or r.flag & 0x900:
            continue
        name = r.query_name
        if r.is_read1 != r.is_read2:
            name += "/1" if r.is_read1 else "/2"
        seq = r.seq or ""
        if r.is_reverse:
            seq = "".join(_COMP[c] for c in reversed(seq))
        out.append((name, seq))
    return out
