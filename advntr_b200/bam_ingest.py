"""Read ingest: alignment file -> the read batches the Viterbi engine takes.

Host-side mirror of the pysam surface adVNTR uses to feed its hot path, over ``libadvbam.so``
(``csrc/bam_ingest.cpp``, ``include/advbam.h``):

=====================================================  ==========================================
reference                                              here
=====================================================  ==========================================
``pysam.AlignmentFile(path, 'rb')``                    ``AlignmentFile(path)``
``samfile.references``, ``.head(5)``, ``.fetch(...)``  same names (records expose the pysam attributes
                                                       the reference reads); ``fetch_batch`` keeps columns
``select_illumina_reads`` mapped loop + quality test   ``select_mapped_illumina`` (decisions in native code)
(``vntr_finder.py:714-753``, ``utils.py:20-38``)
``get_spanning_reads_of_aligned_pacbio_reads``         ``spanning_pacbio_segments``
(``vntr_finder.py:441-470``, ``:373-420``)
``extract_unmapped_reads_to_fasta_file``               ``extract_unmapped_reads`` (no temp files, no samtools)
(``sam_utils.py:9-23``)
=====================================================  ==========================================

BAM only (``.sam`` text and ``.cram`` raise ``ValueError``).  The library is required: there is no
Python fallback reader in the package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("ADVBAM_LIB") or os.path.join(_PKG, "libadvbam.so")
_lib = None

EXPORTS = ("advbam_last_error", "advbam_open", "advbam_close", "advbam_n_references", "advbam_reference_name",
           "advbam_reference_length", "advbam_reference_id", "advbam_head", "advbam_fetch", "advbam_scan",
           "advbam_reads_to_fastq_orientation", "advbam_reads_free", "advbam_reads_view", "advbam_select_illumina",
           "advbam_gather_codes", "advbam_spanning_segments")

DECODE, SKIP_FLAGS, SKIP_SHORT, SKIP_REGION, SKIP_N, SKIP_LOW_QUALITY, BAD_SYMBOL, NO_QUALITIES = range(8)

MAPQ_CUTOFF = 0                        # settings.py:26
QUALITY_SCORE_CUTOFF = 20              # settings.py:24
LOW_QUALITY_BP_TO_DISCARD_READ = 0.10  # settings.py:25


class _View(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(k, C.c_void_p) for k in (
        "flag", "mapq", "tid", "pos", "ref_end", "has_qual", "seq_off", "seq", "qual", "name_off", "names",
        "cigar_off", "cigar")]


class _IlluminaParams(C.Structure):
    _fields_ = [("vntr_start", C.c_int64), ("vntr_end", C.c_int64), ("read_length", C.c_int32),
                ("min_read_length", C.c_int32), ("mapq_cutoff", C.c_int32), ("quality_cutoff", C.c_int32),
                ("low_quality_fraction", C.c_double)]


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError("libadvbam.so is not built (python -m advntr_b200.build); there is no Python fallback reader")
        lib = C.CDLL(_LIB_PATH)
        lib.advbam_last_error.restype = C.c_char_p
        lib.advbam_reference_name.restype = C.c_char_p
        lib.advbam_reference_name.argtypes = [C.c_void_p, C.c_int32]
        lib.advbam_reference_length.restype = C.c_int64
        lib.advbam_reference_length.argtypes = [C.c_void_p, C.c_int32]
        lib.advbam_reference_id.argtypes = [C.c_void_p, C.c_char_p]
        lib.advbam_n_references.argtypes = [C.c_void_p]
        lib.advbam_open.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
        lib.advbam_close.argtypes = [C.c_void_p]
        lib.advbam_close.restype = None
        lib.advbam_head.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        lib.advbam_fetch.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]
        lib.advbam_scan.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(C.c_void_p)]
        lib.advbam_reads_to_fastq_orientation.argtypes = [C.c_void_p]
        lib.advbam_reads_free.argtypes = [C.c_void_p]
        lib.advbam_reads_free.restype = None
        lib.advbam_reads_view.argtypes = [C.c_void_p, C.POINTER(_View)]
        lib.advbam_select_illumina.argtypes = [C.c_void_p, C.POINTER(_IlluminaParams), C.c_void_p, C.POINTER(C.c_int64)]
        lib.advbam_gather_codes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        lib.advbam_spanning_segments.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def _check(rc):
    if rc != 0:
        msg = load_library().advbam_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise IOError(msg)
        raise ValueError(msg)


def _array(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,))


class AlignedRead(object):
    """One record with the ``pysam.AlignedSegment`` attributes adVNTR reads."""

    __slots__ = ("_b", "_i")

    def __init__(self, batch, i):
        self._b, self._i = batch, i

    @property
    def query_name(self):
        b = self._b
        return b.names[b.name_off[self._i]:b.name_off[self._i + 1]].tobytes().decode()

    qname = query_name

    @property
    def flag(self):
        return int(self._b.flag[self._i])

    @property
    def seq(self):
        b = self._b
        a, e = b.seq_off[self._i], b.seq_off[self._i + 1]
        return b.seq[a:e].tobytes().decode() if e > a else None

    query_sequence = query = seq

    @property
    def query_qualities(self):
        b = self._b
        if not b.has_qual[self._i]:
            return None
        return b.qual[b.seq_off[self._i]:b.seq_off[self._i + 1]].tolist()

    @property
    def mapq(self):
        return int(self._b.mapq[self._i])

    mapping_quality = mapq

    @property
    def reference_start(self):
        return int(self._b.pos[self._i])

    pos = reference_start

    @property
    def reference_end(self):
        e = int(self._b.ref_end[self._i])
        return None if e < 0 else e

    aend = reference_end

    @property
    def reference_id(self):
        return int(self._b.tid[self._i])

    tid = reference_id

    @property
    def cigartuples(self):
        b = self._b
        return [(int(c) & 15, int(c) >> 4) for c in b.cigar[b.cigar_off[self._i]:b.cigar_off[self._i + 1]]]

    is_unmapped = property(lambda self: bool(self.flag & 0x4))
    is_reverse = property(lambda self: bool(self.flag & 0x10))
    is_read1 = property(lambda self: bool(self.flag & 0x40))
    is_read2 = property(lambda self: bool(self.flag & 0x80))
    is_secondary = property(lambda self: bool(self.flag & 0x100))
    is_duplicate = property(lambda self: bool(self.flag & 0x400))
    is_supplementary = property(lambda self: bool(self.flag & 0x800))

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


class ReadBatch(object):
    """Columns of the records of one fetch / scan (numpy views of the library's buffers)."""

    def __init__(self, handle):
        self._h = handle
        lib = load_library()
        v = _View()
        _check(lib.advbam_reads_view(handle, C.byref(v)))
        n = self.n = int(v.n)
        self.flag = _array(v.flag, n, np.uint16)
        self.mapq = _array(v.mapq, n, np.uint8)
        self.tid = _array(v.tid, n, np.int32)
        self.pos = _array(v.pos, n, np.int32)
        self.ref_end = _array(v.ref_end, n, np.int32)
        self.has_qual = _array(v.has_qual, n, np.uint8)
        self.seq_off = _array(v.seq_off, n + 1, np.int64)
        self.name_off = _array(v.name_off, n + 1, np.int64)
        self.cigar_off = _array(v.cigar_off, n + 1, np.int64)
        self.seq = _array(v.seq, int(self.seq_off[-1]), np.uint8)
        self.qual = _array(v.qual, int(self.seq_off[-1]), np.uint8)
        self.names = _array(v.names, int(self.name_off[-1]), np.uint8)
        self.cigar = _array(v.cigar, int(self.cigar_off[-1]), np.uint32)

    def __len__(self):
        return self.n

    def __iter__(self):
        return (AlignedRead(self, i) for i in range(self.n))

    def __getitem__(self, i):
        return AlignedRead(self, range(self.n)[i])

    def name(self, i):
        return self.names[self.name_off[i]:self.name_off[i + 1]].tobytes().decode()

    def sequence(self, i):
        return self.seq[self.seq_off[i]:self.seq_off[i + 1]].tobytes().decode()

    def close(self):
        if self._h:
            load_library().advbam_reads_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the reference's per-read tests, on the whole batch ------------------------------------
    def select_illumina(self, vntr_start, vntr_end, read_length, min_read_length=None, mapq_cutoff=MAPQ_CUTOFF,
                        quality_cutoff=QUALITY_SCORE_CUTOFF, low_quality_fraction=LOW_QUALITY_BP_TO_DISCARD_READ):
        """-> (decision per record, vntr_bp_in_mapped_reads); raises what the reference raises for a
        record with non-ACGTN bases (``ValueError`` from ``hmm.viterbi``) or without qualities
        (``TypeError``, ``utils.py:24``)."""
        if min_read_length is None:
            min_read_length = int(read_length * 0.9)    # vntr_finder.py:719
        p = _IlluminaParams(int(vntr_start), int(vntr_end), int(read_length), int(min_read_length), int(mapq_cutoff),
                            int(quality_cutoff), float(low_quality_fraction))
        decision = np.zeros(max(self.n, 1), dtype=np.uint8)
        bp = C.c_int64()
        _check(load_library().advbam_select_illumina(self._h, C.byref(p), decision.ctypes.data, C.byref(bp)))
        decision = decision[:self.n]
        bad = np.flatnonzero(decision == BAD_SYMBOL)
        if len(bad):
            raise ValueError("read %s holds a symbol that is not defined in the model" % self.name(int(bad[0])))
        bad = np.flatnonzero(decision == NO_QUALITIES)
        if len(bad):
            raise TypeError("read %s has no base qualities ('NoneType' object is not iterable)" % self.name(int(bad[0])))
        return decision, int(bp.value)

    def codes(self, decision=None):
        """-> (codes u8 0..3, offsets int64, record indices) of the records to decode."""
        lib = load_library()
        ns, nc = C.c_int64(), C.c_int64()
        d = None if decision is None else np.ascontiguousarray(decision, dtype=np.uint8)
        dp = None if d is None else d.ctypes.data
        _check(lib.advbam_gather_codes(self._h, dp, None, None, None, C.byref(ns), C.byref(nc)))
        codes = np.empty(max(nc.value, 1), dtype=np.uint8)
        off = np.zeros(ns.value + 1, dtype=np.int64)
        index = np.zeros(max(ns.value, 1), dtype=np.int64)
        _check(lib.advbam_gather_codes(self._h, dp, codes.ctypes.data, off.ctypes.data, index.ctypes.data,
                                       C.byref(ns), C.byref(nc)))
        return codes[:nc.value], off, index[:ns.value]

    def spanning_segments(self, vntr_start, vntr_end, hmm_flank=100, min_flank_bp=10):
        n = max(self.n, 1)
        s, e = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
        lb, rb = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        _check(load_library().advbam_spanning_segments(self._h, int(vntr_start), int(vntr_end), hmm_flank, min_flank_bp,
                                                       s.ctypes.data, e.ctypes.data, lb.ctypes.data, rb.ctypes.data))
        return s[:self.n], e[:self.n], lb[:self.n], rb[:self.n]


class AlignmentFile(object):
    def __init__(self, filename, mode="rb", index_filename=None, reference_filename=None):
        if mode != "rb" or not str(filename).endswith(".bam"):
            raise ValueError("only BAM files are read here (mode 'rb'); convert SAM / CRAM input first")
        self._h = C.c_void_p()
        idx = None if index_filename is None else os.fsencode(index_filename)
        _check(load_library().advbam_open(os.fsencode(filename), idx, C.byref(self._h)))
        lib = load_library()
        n = lib.advbam_n_references(self._h)
        self.references = tuple(lib.advbam_reference_name(self._h, i).decode() for i in range(n))
        self.lengths = tuple(int(lib.advbam_reference_length(self._h, i)) for i in range(n))
        self.filename = filename

    def close(self):
        if self._h:
            load_library().advbam_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_tid(self, reference):
        return int(load_library().advbam_reference_id(self._h, reference.encode()))

    def head_batch(self, n):
        h = C.c_void_p()
        _check(load_library().advbam_head(self._h, n, C.byref(h)))
        return ReadBatch(h)

    def head(self, n):
        return iter(self.head_batch(n))

    def fetch_batch(self, reference, start, end):
        tid = self.get_tid(reference)
        if tid < 0:
            raise ValueError("invalid contig `%s`" % reference)      # pysam's message
        h = C.c_void_p()
        _check(load_library().advbam_fetch(self._h, tid, int(start), int(end), C.byref(h)))
        return ReadBatch(h)

    def scan_batch(self, require_flags=0, exclude_flags=0, threads=0):
        h = C.c_void_p()
        _check(load_library().advbam_scan(self._h, require_flags, exclude_flags, threads, C.byref(h)))
        return ReadBatch(h)

    def fetch(self, reference=None, start=None, end=None, until_eof=False):
        if reference is None:
            return iter(self.scan_batch())
        tid = self.get_tid(reference)
        if tid < 0:
            raise ValueError("invalid contig `%s`" % reference)
        start = 0 if start is None else start
        end = self.lengths[tid] if end is None else end
        return iter(self.fetch_batch(reference, start, end))


def get_reference_genome_of_alignment_file(samfile):
    """``sam_utils.py:32-39``."""
    result = None
    if '1' in samfile.references:
        result = 'GRCh37'
    for reference in samfile.references:
        if reference.startswith('chr'):
            result = 'HG19'
    return result


def chromosome_name_in(samfile, chromosome):
    """``vntr_finder.py:711``: hg19-style names are kept, otherwise the ``chr`` prefix is dropped."""
    return chromosome if get_reference_genome_of_alignment_file(samfile) == 'HG19' else chromosome[3:]


def median_head_read_length(samfile, n=5):
    """``vntr_finder.py:714-718``."""
    b = samfile.head_batch(n)
    lengths = sorted(int(x) for x in (b.seq_off[1:] - b.seq_off[:-1]))
    b.close()
    if not lengths:
        raise IndexError("list index out of range")     # what the reference hits on an empty file
    return lengths[len(lengths) // 2]


def select_mapped_illumina(samfile, chromosome, vntr_start, vntr_end, read_length, min_read_length=None):
    """The mapped-read loop of ``select_illumina_reads`` up to the Viterbi call (``vntr_finder.py:727-753``)
    for one locus: -> dict(codes, off, names, mapq, reference_start, vntr_bp); ``codes`` / ``off`` is the
    batch layout of ``advhmm_viterbi_multi``.  Reads the reference decodes but then drops for low quality
    (``:739``) are not decoded at all; the recruited set is the same."""
    b = samfile.fetch_batch(chromosome_name_in(samfile, chromosome), vntr_start, vntr_end)
    decision, bp = b.select_illumina(vntr_start, vntr_end, read_length, min_read_length)
    codes, off, index = b.codes(decision)
    out = {"codes": codes.copy(), "off": off, "names": [b.name(int(i)) for i in index],
           "mapq": b.mapq[index].copy(), "reference_start": b.pos[index].copy(), "vntr_bp": bp,
           "n_fetched": b.n, "decision": decision.copy()}
    b.close()
    return out


def spanning_pacbio_segments(samfile, chromosome, vntr_start, vntr_end):
    """``get_spanning_reads_of_aligned_pacbio_reads`` (``vntr_finder.py:441-470``): -> [(query_name,
    sequence, length-distribution entry)] of the mapped reads that span the locus, cut to the part the
    HMM models (100 bp flanks)."""
    b = samfile.fetch_batch(chromosome_name_in(samfile, chromosome), vntr_start, vntr_end)
    s, e, lb, rb = b.spanning_segments(vntr_start, vntr_end)
    out = []
    for i in np.flatnonzero(s >= 0):
        a = int(b.seq_off[i])
        seq = b.seq[a + int(s[i]):a + int(e[i])].tobytes().decode()
        out.append((b.name(int(i)), seq, len(seq) - int(lb[i]) - int(rb[i])))
    b.close()
    return out


def extract_unmapped_reads(alignment_file, threads=0):
    """``extract_unmapped_reads_to_fasta_file`` (``sam_utils.py:9-23``) without the temporary files:
    -> (names, sequences) of the records ``samtools view -f4 | samtools bam2fq`` prints."""
    own = not isinstance(alignment_file, AlignmentFile)
    f = AlignmentFile(alignment_file) if own else alignment_file
    b = f.scan_batch(require_flags=0x4, exclude_flags=0x900, threads=threads)
    _check(load_library().advbam_reads_to_fastq_orientation(b._h))
    b = _rebind(b)
    names = [b.name(i) for i in range(b.n)]
    seqs = [b.sequence(i) for i in range(b.n)]
    b.close()
    if own:
        f.close()
    return names, seqs


def _rebind(batch):
    h, batch._h = batch._h, None
    return ReadBatch(h)
