"""Side workloads of bench.py (not the driver's default line): BASELINE config 3 (PacBio-like long
reads against the 100-copy model) and config 4 (--frameshift mode: full state paths, indel calls).

    python bench.py --workload config3 [--long-reads N] [--steps K]
    python bench.py --workload frameshift [--loci N] [--steps K]

Each prints ONE JSON line in the layout of the main bench (metric, value, e2e, roofline, cpu_baseline).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    return oracle


# ------------------------------------------------------------------------------------ config 3
def config3_reads(loc, n_reads, seed=31):
    """PacBio-like reads: left flank + k copies of the repeat unit + right flank, k drawn so that the
    reads are 10-20 kb, CLR-like errors (2 % substitutions, 5 % insertions, 5 % deletions)."""
    from advntr_b200 import synth
    rng = random.Random(seed)
    R = len(loc.pattern)
    lo, hi = (10000 - 200) // R + 1, (20000 - 200) // R - 12
    out = []
    for _ in range(n_reads):
        k = rng.randint(lo, hi)
        out.append(synth.encode(synth.sequencing_errors(rng, loc.left + loc.pattern * k + loc.right, 0.02, 0.05, 0.05)[:20000]))
    return out


def _cpu_long_worker(job):
    baked, codes = job
    oracle = _oracle()
    om = oracle.OracleModel(baked)
    t0 = time.perf_counter()
    om.viterbi([codes])
    return len(codes) * om.n_states, time.perf_counter() - t0


def run_config3(args):
    import bench
    from advntr_b200 import engine, fast_compile, synth
    D = bench.Dist()
    torch = D.torch
    ctx, stream = bench.make_context(D)
    lib = engine.load_library()
    loc = synth.config3_locus()
    t0 = time.time()
    model = fast_compile.compile_many([(loc.left, loc.right, loc.segments, loc.copies, 0.3)], ctx)[0]
    t_model = time.time() - t0
    dm = model._device_model()
    m = dm.info.n_states
    reads = config3_reads(loc, args.long_reads, seed=31 + D.rank)
    seqs, off = engine.pack_reads(reads)
    R = len(reads)
    lens = np.diff(off)
    cells = float(lens.sum()) * m
    t = dm.tables()
    deg = np.diff(t["in_off"])
    e_emit, e_sil = int(deg[:t["silent_start"]].sum()), int(deg[t["silent_start"]:].sum())
    ops = float(lens.sum()) * (2 * e_emit + e_sil + e_emit + e_sil)
    handles = (C.c_void_p * 1)(dm._h)
    goff = np.asarray([0, R], dtype=np.int64)
    path_cap = int(off[-1]) + R * 512
    d_seqs = torch.from_numpy(seqs).cuda()
    d_logp = torch.empty(R, dtype=torch.float64, device="cuda")
    d_plen = torch.empty(R, dtype=torch.int32, device="cuda")
    d_poff = torch.empty(R, dtype=torch.int64, device="cuda")
    d_path = torch.empty(path_cap, dtype=torch.int32, device="cuda")
    d_total = torch.zeros(1, dtype=torch.int64, device="cuda")
    h_seqs = torch.from_numpy(seqs).pin_memory()
    h_logp = torch.empty(R, dtype=torch.float64).pin_memory()
    h_plen = torch.empty(R, dtype=torch.int32).pin_memory()
    h_poff = torch.empty(R, dtype=torch.int64).pin_memory()
    h_path = torch.empty(path_cap, dtype=torch.int32).pin_memory()
    h_total = C.c_int64(0)

    def step_device():
        engine._check(lib.advhmm_viterbi_multi(ctx._h, handles, 1, goff.ctypes.data, d_seqs.data_ptr(), off.ctypes.data, R,
                                               engine.WANT_PATH | engine.DEVICE_BUFFERS, d_logp.data_ptr(), d_plen.data_ptr(),
                                               d_poff.data_ptr(), d_path.data_ptr(), path_cap, d_total.data_ptr()))

    def step_host():
        engine._check(lib.advhmm_viterbi_multi(ctx._h, handles, 1, goff.ctypes.data, h_seqs.data_ptr(), off.ctypes.data, R,
                                               engine.WANT_PATH, h_logp.data_ptr(), h_plen.data_ptr(), h_poff.data_ptr(),
                                               h_path.data_ptr(), path_cap, C.byref(h_total)))

    sampler = bench.ClockSampler(D.local)
    for _ in range(max(args.warmup, 3)):
        step_device()
    D.barrier()
    n_paths = int(d_total.item())
    assert n_paths <= path_cap and int((d_plen < 0).sum().item()) == 0
    # self-check outside the timed region: the two shortest reads against the CPU oracle, every path re-scored
    verified = None
    if D.rank == 0:
        from advntr_b200 import path_utils
        oracle = _oracle()
        order = np.argsort(lens)[:2]
        lp, paths = oracle.OracleModel(t).viterbi([reads[i] for i in order])
        h_lp, h_pl, h_po = d_logp.cpu().numpy(), d_plen.cpu().numpy(), d_poff.cpu().numpy()
        h_pa = d_path[:n_paths].cpu().numpy()
        same = all(lp[k] == h_lp[i] and np.array_equal(paths[k], h_pa[h_po[i]:h_po[i] + h_pl[i]]) for k, i in enumerate(order))
        sub = list(range(0, R, max(1, R // 16)))[:16]
        sc = path_utils.rescore_paths(t, [reads[i] for i in sub], [h_pa[h_po[i]:h_po[i] + h_pl[i]] for i in sub])
        ok = bool(np.array_equal(sc.view(np.int64), h_lp[sub].view(np.int64)))
        verified = {"oracle_reads": [int(lens[i]) for i in order], "equal_to_cpu_oracle": bool(same),
                    "paths_rescored": len(sub), "paths_rescored_bit_exact": ok}
        if not (same and ok):
            raise SystemExit("config3 self-check failed")
    ctx.profile(True); ctx.profile_read()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    t0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    D.barrier()
    t1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    fill_ms, fill_n, bt_ms, bt_n = ctx.profile_read()
    ctx.profile(False)
    clocks = sampler.window(t0, t1)
    sampler.stop()
    step_host()
    D.barrier()
    te = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - te) * 1e3
    assert torch.equal(h_logp.view(torch.int64), d_logp.cpu().view(torch.int64))
    ms_all, e2e_all = D.reduce([ms, e2e_ms], "max")
    reads_all, cells_all = D.reduce([float(R), cells], "sum")
    if D.rank == 0:
        K = args.steps
        fp64_peak = ctx.fp64_add_peak()
        fill_s = fill_ms * 1e-3
        achieved = ops * K / fill_s / 1e9
        line = {"metric": "viterbi_reads_per_s", "value": reads_all * K / (ms_all * 1e-3), "unit": "reads/s",
                "gcups": cells_all * K / (ms_all * 1e-3) / 1e9, "kernel_gcups": cells * K / fill_s / 1e9,
                "n_gpus": D.world, "steps": K, "warmup": args.warmup, "ms_per_step": ms_all / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "config3: PacBio-like 10-20 kb reads (2 % sub, 5 % ins, 5 % del) vs the 60 bp x 100 copies "
                                       "model (18,918 states, 6,412 columns, error rate 0.3), full state paths",
                           "reads_per_gpu": R, "read_length_min": int(lens.min()), "read_length_max": int(lens.max()),
                           "read_length_mean": float(lens.mean()), "n_states": int(m), "want_path": True,
                           "l2": "1.3 MB of tables per model, 60-100 MB of traceback per read: exceeds L2"},
                "gpu_launches": int(launches), "clocks": clocks, "model_compile_s": round(t_model, 3),
                "e2e": {"value": reads_all * K / (e2e_all * 1e-3), "unit": "reads/s",
                        "gcups": cells_all * K / (e2e_all * 1e-3) / 1e9, "h2d_bytes_per_step": int(off[-1]),
                        "d2h_bytes_per_step": R * 20 + n_paths * 4, "ms_per_step": e2e_all / K},
                "verified": verified,
                "roofline": {"kernel": "banded_long_kernel", "bound": "fp64_issue", "achieved": achieved, "peak": fp64_peak,
                             "unit": "Gop/s", "frac": achieved / fp64_peak, "launches": int(fill_n),
                             "avg_launch_ms": fill_ms / max(fill_n, 1), "share_of_step": fill_ms / ms,
                             "backtrack_share_of_step": bt_ms / ms,
                             "hbm_achieved_gbs": cells * K / fill_s / 1e9, "traffic": None,
                             "note": "hbm_achieved_gbs = 1 algorithmic traceback byte per DP cell (SURVEY 8d); the kernel "
                                     "writes 0.8 B per cell (6 bits per cell, 5 rows per 32-bit word)"}}
        if not args.no_cpu_baseline and D.world == 1:
            import multiprocessing as mp
            procs = bench.host_cores()
            sample = [reads[i] for i in np.argsort(lens)[:procs]]
            with mp.get_context("fork").Pool(procs) as pool:
                res = pool.map(_cpu_long_worker, [(t, c) for c in sample])
            busy = max(r[1] for r in res)
            line["cpu_baseline"] = {"value": len(sample) / busy, "unit": "reads/s", "gcups": sum(r[0] for r in res) / busy / 1e9,
                                    "cores": procs, "kind": "port",
                                    "sample": "the %d shortest reads (%d-%d bases), one per process, %.1f s; oracle/hmm_oracle.c "
                                              "(the compiled reference needs > 1 min per read of this size)" %
                                              (len(sample), len(sample[0]), len(sample[-1]), busy)}
        print(json.dumps(line))
    model._release_engine()
    ctx.close()
    D.close()


# ---------------------------------------------------------------------------------- frameshift
def frameshift_locus_reads(lid, coverage=30):
    """A coding-VNTR-like config-2 locus whose sample carries a 1 bp indel in one repeat unit on one
    haplotype: about half of the reads that cover that unit show it."""
    from advntr_b200 import synth
    loc = synth.config2_locus(lid)
    rng = random.Random(99991 * lid + 5)
    R = len(loc.pattern)
    unit = rng.randrange(len(loc.segments))
    pos = rng.randrange(1, R - 1)
    segs = list(loc.segments)
    if rng.random() < 0.5:
        segs[unit] = segs[unit][:pos] + segs[unit][pos + 1:]                       # deletion
    else:
        segs[unit] = segs[unit][:pos] + rng.choice("ACGT") + segs[unit][pos:]      # insertion
    alleles = [loc.left + "".join(loc.segments) + loc.right, loc.left + "".join(segs) + loc.right]
    vntr_len = sum(len(s) for s in loc.segments)
    L = loc.read_length
    n = max(1, int(round((vntr_len + L) * coverage / float(L))))
    reads = []
    for _ in range(n):
        a = alleles[rng.random() < 0.5]
        s = rng.randint(len(loc.left) - L + 20, len(loc.left) + vntr_len - 20)
        s = max(0, min(s, len(a) - L - 8))
        reads.append(synth.sequencing_errors(rng, a[s:s + L + 8], 0.004, 0.0003, 0.0003)[:L])
    return loc, reads


def run_frameshift(args):
    """--frameshift mode (genome_analyzer.py:260, vntr_finder.py:776-780, :265-309): every read of a locus
    decoded to its FULL state path, reads recruited, the frame-shifting indel states counted.  Timed two
    ways: the pinned C-ABI route (paths in one flat array) and the pageable Python route a drop-in
    caller gets (LocusDecoder.select_reads + frameshift_candidate: Python objects per read).  The
    frameshift call of every locus is compared between the device paths and CPU-oracle paths through
    the same host logic."""
    import bench
    from advntr_b200 import engine, fast_compile, locus_batch, path_utils, synth
    D = bench.Dist()
    torch = D.torch
    ctx, stream = bench.make_context(D)
    lib = engine.load_library()
    n_loci = min(args.loci, 512)
    ids = list(range(D.rank * n_loci + 1, (D.rank + 1) * n_loci + 1))
    loci, reads = [], []
    for lid in ids:
        loc, rs = frameshift_locus_reads(lid, args.coverage)
        loci.append(loc); reads.append(rs)
    decoders = [locus_batch.LocusDecoder(l.left, l.right, l.segments, read_length=150, locus_id=l.id) for l in loci]
    fast_compile.attach_device_models([d.model for d in decoders], ctx)
    models = [d.model._device_model() for d in decoders]
    flat = [engine.encode_acgt(r)[0] for rs in reads for r in rs]
    seqs, off = engine.pack_reads(flat)
    goff = np.zeros(len(loci) + 1, dtype=np.int64)
    np.cumsum([len(rs) for rs in reads], out=goff[1:])
    R = len(flat)
    cells = float(sum(len(rs) * 150 * m.info.n_states for rs, m in zip(reads, models)))
    handles = (C.c_void_p * len(models))(*[m._h for m in models])
    path_cap = int(off[-1]) + R * 96
    h_seqs = torch.from_numpy(seqs).pin_memory()
    h_logp = torch.empty(R, dtype=torch.float64).pin_memory()
    h_plen = torch.empty(R, dtype=torch.int32).pin_memory()
    h_poff = torch.empty(R, dtype=torch.int64).pin_memory()
    h_path = torch.empty(path_cap, dtype=torch.int32).pin_memory()
    h_total = C.c_int64(0)
    h_summ = torch.zeros((R, 8), dtype=torch.int32).pin_memory()

    def step_host():
        engine._check(lib.advhmm_viterbi_multi_summary(
            ctx._h, handles, len(models), goff.ctypes.data, h_seqs.data_ptr(), off.ctypes.data, R,
            engine.WANT_PATH | engine.WANT_SUMMARY, h_logp.data_ptr(), h_plen.data_ptr(), h_poff.data_ptr(), h_path.data_ptr(),
            path_cap, C.byref(h_total), h_summ.data_ptr()))

    # what advhmm_frameshift_candidates reads instead of state names, per model (made once, like the models)
    state_tables = [path_utils.frameshift_state_tables([s.name for s in d.model.states]) for d in decoders]
    pattern_len = [len(d.pattern) for d in decoders]

    def calls_native(threads=0):
        """The same calls from the library: recruit_read + the path walk of find_frameshift_from_selected_reads for
        every locus on all host threads."""
        rec = engine.frameshift_candidates(goff, pattern_len, None, state_tables, h_logp.numpy(),
                                           h_summ.numpy().view(engine.SUMMARY_DTYPE).reshape(-1), h_plen.numpy(), h_poff.numpy(),
                                           h_path.numpy(), seqs, off, threads=threads)
        return [(path_utils.frameshift_label(c), int(c["count"])) for c in rec]

    def calls_from_flat():
        """The frameshift call of every locus from the flat path array of the C-ABI route."""
        lp, pl, po, pa = h_logp.numpy(), h_plen.numpy(), h_poff.numpy(), h_path.numpy()
        out = []
        for g, (dec, rs) in enumerate(zip(decoders, reads)):
            st = dec.model.states
            sel = []
            for i, r in enumerate(rs):
                k = int(goff[g]) + i
                vp = [(int(x), st[x]) for x in pa[po[k]:po[k] + pl[k]]]
                if path_utils.recruit_read(lp[k], vp, None, r, dec.left_flank, dec.right_flank):
                    sel.append(locus_batch.SelectedRead(r, float(lp[k]), vp))
            out.append(dec.frameshift_candidate(sel)[0])
        return out

    sampler = bench.ClockSampler(D.local)
    for _ in range(max(args.warmup, 3)):
        step_host()
    D.barrier()
    t0 = time.time()
    te = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - te) * 1e3
    t1 = time.time()
    clocks = sampler.window(t0, t1)
    sampler.stop()
    tc = time.perf_counter()
    calls = calls_from_flat()
    consumers_s = time.perf_counter() - tc
    calls_native()
    tc = time.perf_counter()
    for _ in range(5):
        nat_calls = calls_native()
    native_s = (time.perf_counter() - tc) / 5
    tc = time.perf_counter()
    calls_native(threads=1)
    native1_s = time.perf_counter() - tc
    # one step carried on to the calls: decode + native consumers, wall clock
    tc = time.perf_counter()
    step_host()
    calls_native()
    to_calls_s = time.perf_counter() - tc
    # pageable Python route, per locus, as a drop-in caller of vntr_finder's loop would run it
    tp = time.perf_counter()
    py_calls = []
    for dec, rs in zip(decoders, reads):
        selected = dec.select_reads(rs)
        py_calls.append(dec.frameshift_candidate(selected)[0])
    py_s = time.perf_counter() - tp
    # the same host logic on CPU-oracle paths, first loci (outside every timed region)
    n_or = min(3 * args.oracle_loci if args.oracle_loci > 0 else 0, len(loci))     # ~0.5 s of CPU oracle per locus
    oracle_same = None
    if n_or and D.rank == 0:
        oracle = _oracle()
        oracle_same = True
        for g in range(n_or):
            dec, rs = decoders[g], reads[g]
            lp, paths = oracle.OracleModel(dec.model.baked).viterbi([oracle.encode(r) for r in rs])
            st = dec.model.states
            sel = [locus_batch.SelectedRead(r, float(lp[i]), [(int(x), st[x]) for x in paths[i]]) for i, r in enumerate(rs)
                   if path_utils.recruit_read(lp[i], [(int(x), st[x]) for x in paths[i]], None, r, dec.left_flank, dec.right_flank)]
            oracle_same = oracle_same and dec.frameshift_candidate(sel)[0] == calls[g]
    agree = calls == py_calls
    native_agree = nat_calls == calls
    with_indel = sum(1 for c in calls if c[0] is not None and c[1] >= 3)
    e2e_all = D.reduce([e2e_ms], "max")[0]
    reads_all, cells_all = D.reduce([float(R), cells], "sum")
    if D.rank == 0:
        K = args.steps
        if not agree or not native_agree or oracle_same is False:
            raise SystemExit("frameshift workload: calls differ between routes / from the CPU oracle")
        line = {"metric": "viterbi_reads_per_s", "value": reads_all * K / (e2e_all * 1e-3), "unit": "reads/s",
                "gcups": cells_all * K / (e2e_all * 1e-3) / 1e9, "n_gpus": D.world, "steps": K, "warmup": max(args.warmup, 3),
                "ms_per_step": e2e_all / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "config4 (--frameshift): config-2 loci, 30x reads, a 1 bp indel in one repeat unit on one "
                                       "haplotype, full Viterbi paths to the host for every read",
                           "loci_per_gpu": len(loci), "reads_per_gpu": R, "want_path": True},
                "note": "value = e2e: pinned host reads in, logp + full state paths out through advhmm_viterbi_multi",
                "e2e": {"value": reads_all * K / (e2e_all * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": int(off[-1]),
                        "d2h_bytes_per_step": R * 52 + int(h_total.value) * 4},
                "frameshift": {"loci": len(loci), "loci_with_an_indel_call_of_3_or_more_reads": with_indel,
                               "calls_equal_between_c_abi_and_python_routes": agree,
                               "calls_equal_to_cpu_oracle_paths": oracle_same, "oracle_loci": n_or,
                               "native_consumers_reads_per_s": R / native_s,
                               "native_consumers_one_thread_reads_per_s": R / native1_s,
                               "native_calls_equal_to_python_consumers": native_agree,
                               "reads_to_calls_reads_per_s": R / to_calls_s,
                               "python_consumers_reads_per_s": R / consumers_s,
                               "note": "native_consumers = advhmm_frameshift_candidates (recruit_read + the path walk of "
                                       "find_frameshift_from_selected_reads for every locus, all host threads) on the flat path "
                                       "array; reads_to_calls = one decode step + native consumers, wall clock; python_consumers "
                                       "= the same logic in host Python, one process: what a Python caller is bound by"},
                "pageable_python_route": {"value": R / py_s, "unit": "reads/s",
                                          "note": "LocusDecoder.select_reads + frameshift_candidate per locus: viterbi_batch(list "
                                                  "of str), (idx, State) lists, Python consumers; one process"},
                "clocks": clocks, "gpu_launches": int(ctx.launch_count)}
        print(json.dumps(line))
    for d in decoders:
        d.model._release_engine()
    ctx.close()
    D.close()
