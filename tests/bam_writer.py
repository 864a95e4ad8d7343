"""Test helper: write small coordinate-sorted BAM files and their BAI index (hts-specs SAMv1 4, 4.1, 5.2).

There is no samtools / pysam in this image, so the fixtures of tests/test_bam_ingest.py are produced
here.  ``block_size`` controls how many uncompressed bytes go into one BGZF block, which lets a test
force records to straddle blocks.
"""
import struct
import zlib

NIBBLE = "=ACMGRSVTWYHKDBN"
CIGAR_OPS = "MIDNSHP=X"


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def parse_cigar(text):
    out, num = [], ""
    for ch in text:
        if ch.isdigit():
            num += ch
        else:
            out.append((CIGAR_OPS.index(ch), int(num)))
            num = ""
    return out


def reference_length(cigar):
    return sum(n for op, n in cigar if op in (0, 2, 3, 7, 8))


class Read(object):
    def __init__(self, name, flag, tid, pos, mapq, cigar, seq, qual=None, tags=b"", long_cigar=False):
        self.name, self.flag, self.tid, self.pos, self.mapq = name, flag, tid, pos, mapq
        self.cigar = parse_cigar(cigar) if isinstance(cigar, str) else list(cigar)
        self.seq, self.qual, self.tags, self.long_cigar = seq, qual, tags, long_cigar

    def end(self):
        rlen = 0 if self.flag & 4 else reference_length(self.cigar)
        return self.pos + (rlen or 1)

    def encode(self):
        cigar, tags = self.cigar, self.tags
        if self.long_cigar:                             # real CIGAR in CG:B,I, placeholder in the record
            tags = tags + b"CGBI" + struct.pack("<I%dI" % len(cigar), len(cigar), *[n << 4 | op for op, n in cigar])
            cigar = [(4, len(self.seq)), (3, reference_length(self.cigar))]
        name = self.name.encode() + b"\0"
        l_seq = len(self.seq)
        packed = bytearray((l_seq + 1) // 2)
        for i, c in enumerate(self.seq):
            packed[i >> 1] |= NIBBLE.index(c) << (4 if i % 2 == 0 else 0)
        qual = bytes([0xff] * l_seq) if self.qual is None else bytes(self.qual)
        body = struct.pack("<iiBBHHHiiii", self.tid, self.pos, len(name), self.mapq,
                           reg2bin(self.pos, self.end()) if self.tid >= 0 else 4680, len(cigar), self.flag, l_seq, -1, -1, 0)
        body += name + b"".join(struct.pack("<I", n << 4 | op) for op, n in cigar) + bytes(packed) + qual + tags
        return struct.pack("<i", len(body)) + body


class _Bgzf(object):
    def __init__(self, fh, block_size):
        self.fh, self.block_size, self.buf, self.coff = fh, block_size, bytearray(), 0

    def tell(self):
        return self.coff << 16 | len(self.buf)

    def flush(self):
        if not self.buf and self.coff:
            return
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = co.compress(bytes(self.buf)) + co.flush()
        block = (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(comp) + 25) + comp +
                 struct.pack("<II", zlib.crc32(bytes(self.buf)), len(self.buf)))
        self.fh.write(block)
        self.coff += len(block)
        self.buf = bytearray()

    def write(self, data):
        data = memoryview(data)
        while len(data):
            k = min(len(data), self.block_size - len(self.buf))
            self.buf += data[:k]
            data = data[k:]
            if len(self.buf) == self.block_size:
                self.flush()

    def close(self):
        if self.buf:
            self.flush()
        self.fh.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))   # EOF marker block


def write_bam(path, references, reads, block_size=0xff00, header_text=None, index_path=None, with_metadata=True):
    """``references``: [(name, length)], ``reads``: Read objects sorted by (tid, pos), unplaced (tid -1) last."""
    if header_text is None:
        header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in references)
    with open(path, "wb") as fh:
        z = _Bgzf(fh, block_size)
        text = header_text.encode()
        z.write(b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(references)))
        for name, length in references:
            z.write(struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", length))
        z.flush()                                       # samtools starts the alignments on a fresh block
        bins = [dict() for _ in references]
        linear = [dict() for _ in references]
        meta = [[None, None, 0, 0] for _ in references]
        n_no_coor = 0
        for r in reads:
            beg = z.tell()
            z.write(r.encode())
            end = z.tell()
            if r.tid < 0:
                n_no_coor += 1
                continue
            chunks = bins[r.tid].setdefault(reg2bin(r.pos, r.end()), [])
            if chunks and chunks[-1][1] == beg:
                chunks[-1][1] = end
            else:
                chunks.append([beg, end])
            for w in range(r.pos >> 14, ((r.end() - 1) >> 14) + 1):
                linear[r.tid].setdefault(w, beg)
            m = meta[r.tid]
            m[0] = beg if m[0] is None else m[0]
            m[1] = end
            m[3 if r.flag & 4 else 2] += 1
        z.close()
    with open(index_path or path + ".bai", "wb") as fh:
        fh.write(b"BAI\1" + struct.pack("<i", len(references)))
        for tid in range(len(references)):
            b = bins[tid]
            extra = 1 if (with_metadata and b) else 0
            fh.write(struct.pack("<i", len(b) + extra))
            for bin_id in sorted(b):
                fh.write(struct.pack("<Ii", bin_id, len(b[bin_id])))
                for beg, end in b[bin_id]:
                    fh.write(struct.pack("<QQ", beg, end))
            if extra:
                m = meta[tid]
                fh.write(struct.pack("<Ii", 37450, 2) + struct.pack("<QQQQ", m[0], m[1], m[2], m[3]))
            n_intv = max(linear[tid]) + 1 if linear[tid] else 0
            fh.write(struct.pack("<i", n_intv))
            prev = 0
            for w in range(n_intv):
                prev = linear[tid].get(w, prev)
                fh.write(struct.pack("<Q", prev))
        fh.write(struct.pack("<Q", n_no_coor))
    return path
