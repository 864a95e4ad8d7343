"""Keyword pre-filter throughput (device-resident reads, CUDA events on the launching stream) next to
the reference's own `adVNTR-Filtering` binary (oracle/_ref, one host core: it is single-threaded)
on a sample of the same reads.  Workload: the 15-mer keywords of the first N config-2 loci
(genome_analyzer.py:181) against `reads` unmapped 150 bp reads: 2 % locus-derived (either strand,
Illumina-like errors), the rest random decoys.

    python tools/kbench_filter.py [n_loci=6719] [n_reads=4000000] [cpu_sample=200000]

Prints one JSON line (also the input of profiles/r1_kfilter.md).
"""
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import torch

from advntr_b200 import engine, keyword_filter, synth

n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 6719
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 4000000
cpu_sample = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
L = 150

t0 = time.time()
loci = [synth.config2_locus(i) for i in range(1, n_loci + 1)]
kw = [(l.id, sorted(keyword_filter.get_keywords_for_filtering(l.left, l.right, l.segments, l.pattern, keyword_size=15)))
      for l in loci]
rng = np.random.Generator(np.random.PCG64(5))
text = rng.integers(0, 4, size=(n_reads, L), dtype=np.uint8)
n_true = n_reads // 50
which = rng.integers(0, n_loci, size=n_true)
for i, li in enumerate(which):                      # every 50th read comes from a locus
    l = loci[li]
    seq = synth.encode(l.sequence)
    s = int(rng.integers(350, 500 + sum(len(x) for x in l.segments) - 20))
    r = seq[s:s + L]
    if len(r) < L:
        continue
    r = np.where(rng.random(L) < 0.01, rng.integers(0, 4, size=L, dtype=np.uint8), r).astype(np.uint8)
    if rng.random() < 0.5:
        r = synth.revcomp_codes(r)
    text[i * 50] = r
ascii_reads = np.frombuffer(b"ACGT", dtype=np.uint8)[text].reshape(-1)
off = np.arange(n_reads + 1, dtype=np.int64) * L
setup_s = time.time() - t0

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = engine.Context(0, stream.cuda_stream)
kf = keyword_filter.KeywordFilter(kw, ctx=ctx)
d_seqs = torch.from_numpy(ascii_reads).cuda()
cap = max(1 << 20, n_reads // 4)
d_r, d_l, d_c = (torch.empty(cap, dtype=torch.int32, device="cuda") for _ in range(3))
d_n = torch.zeros(1, dtype=torch.int64, device="cuda")


d_off = torch.from_numpy(off).cuda()


def step():
    kf.filter.scan_device(d_seqs.data_ptr(), d_off.data_ptr(), n_reads, keyword_filter.MIN_MATCHES, d_r.data_ptr(),
                          d_l.data_ptr(), d_c.data_ptr(), cap, d_n.data_ptr())


for _ in range(3):
    step()
torch.cuda.synchronize()
steps = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ctx.profile(True)
ctx.profile_read()
e0.record(stream)
for _ in range(steps):
    step()
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
scan_ms = ctx.profile_read()[0] / steps
ctx.profile(False)
hits = int(d_n.item())

# end to end through the host API (pageable host buffers in, triples out)
t = time.time()
hr, hl, hc = kf.filter.scan(ascii_reads, off, keyword_filter.MIN_MATCHES)
e2e_s = time.time() - t
assert len(hr) == hits

out = {"tool": "kbench_filter", "n_loci": n_loci, "n_keywords": sum(len(w) for _, w in kw), "n_reads": n_reads,
       "read_length": L, "text_bytes": int(ascii_reads.nbytes), "hits": hits,
       "gpu_ms_per_scan": ms, "scan_kernel_ms": scan_ms, "scan_kernel_text_gb_per_s": ascii_reads.nbytes / scan_ms / 1e6, "gpu_reads_per_s": n_reads / ms * 1e3, "gpu_text_gb_per_s": ascii_reads.nbytes / ms / 1e6,
       "e2e_host_api_s": e2e_s, "e2e_reads_per_s": n_reads / e2e_s, "setup_s": setup_s}

# the reference binary on a sample of the same reads, and parity of the selected pairs on it
fbin = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "adVNTR-Filtering")
if cpu_sample and os.path.exists(fbin):
    n = min(cpu_sample, n_reads)
    with tempfile.TemporaryDirectory() as d:
        fa, kwf = os.path.join(d, "reads.fa"), os.path.join(d, "kw.txt")
        rows = ascii_reads[:n * L].reshape(n, L)
        with open(fa, "wb") as fh:
            for i in range(n):
                fh.write(b">r%d\n" % i)
                fh.write(rows[i].tobytes())
                fh.write(b"\n")
        with open(kwf, "w") as fh:
            for vid, words in kw:
                fh.write("%d %s\n" % (vid, " ".join(words)))
        with open(kwf) as stdin:                      # machine construction alone (no reads)
            t = time.time()
            subprocess.run([fbin, os.devnull], stdin=stdin, capture_output=True, check=True)
            build_s = time.time() - t
        with open(kwf) as stdin:
            t = time.time()
            res = subprocess.run([fbin, fa], stdin=stdin, capture_output=True, text=True, check=True)
            total_s = time.time() - t
    want = set()
    for line in res.stdout.split("\n"):
        tok = line.split()
        if len(tok) >= 2 and tok[0].isdigit() and tok[1].isdigit():
            want.update((int(name[1:]), int(tok[0])) for name in tok[2:])
    sel = hr < n
    got = set(zip(hr[sel].tolist(), hl[sel].tolist()))
    out.update({"cpu_sample_reads": n, "cpu_build_s": build_s, "cpu_total_s": total_s,
                "cpu_scan_reads_per_s": n / max(total_s - build_s, 1e-9), "cpu_cores": 1,
                "pairs_equal_on_sample": got == want, "pairs_on_sample": len(want)})
print(json.dumps(out))
