#!/usr/bin/env python
"""Benchmark of the hot path: batched Viterbi decoding of Illumina reads against per-locus
VNTR read-matcher HMMs (BASELINE.json: "Viterbi GCUPS & reads/s (150bp Illumina) ...").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU engine

Workload (``config.workload``): BASELINE config 2 -- synthetic loci shaped like the recommended
hg19 Illumina set (RU 6..70 bp, 150 bp flanks in the model), per locus the 30x mapped reads that
overlap the VNTR plus 50 decoy unmapped reads on both strands, 150 bp reads.  One *step* = one
pass of the hot path over every read of every locus of the rank: 2-bit packing, banded Viterbi
fill, device backtrack to full state paths.  Weak scaling: every rank decodes its own
``--loci`` loci (a disjoint slice of the locus id space), no collective on the data path.

  value   reads/s over all ranks with inputs resident in HBM (CUDA events, max over ranks)
  e2e     the same through the host-buffer C-ABI call: pinned host reads in, logp + paths out
  roofline / cpu_baseline: see DESIGN.md section "Measurement"
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = int(os.environ.get("ADVNTR_BENCH_READ_LEN", "150"))   # 150 = BASELINE config 2; other lengths for side experiments only


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--loci", type=int, default=6719, help="loci per GPU (config 2: 6,719)")
    ap.add_argument("--coverage", type=int, default=30)
    ap.add_argument("--decoys", type=int, default=50)
    ap.add_argument("--cpu-sample-loci", type=int, default=0, help="loci in the CPU sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--verify-loci", type=int, default=48,
                    help="re-score the state paths of the first N loci on the host (0: skip)")
    return ap.parse_args()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------ workload
def locus_ids(rank, n_loci):
    return range(rank * n_loci + 1, (rank + 1) * n_loci + 1)


def build_workload(rank, n_loci, coverage, decoys):
    """Models (host tables) + reads of this rank's loci."""
    from advntr_b200 import fast_compile, synth
    baked, flats, lens, goff, cells, relax = [], [], [], [0], 0, 0
    n_states = []
    for lid in locus_ids(rank, n_loci):
        loc = synth.config2_locus(lid, READ_LEN)
        model = fast_compile.build_vntr_matcher_hmm(loc.left, loc.right, loc.segments, loc.copies,
                                                    flank_size=loc.flank, error_rate=loc.error_rate)
        flat, ln = synth.config2_read_codes(loc, coverage, decoys)
        baked.append(model.baked)
        flats.append(flat)
        lens.append(ln)
        goff.append(goff[-1] + len(ln))
        m = model.baked["n_states"]
        n_states.append(m)
        cells += int(ln.sum()) * m
        relax += int(ln.sum()) * len(model.baked["in_src"])      # edge relaxations = sum n * E
    lens = np.concatenate(lens)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return {"baked": baked, "seqs": np.concatenate(flats), "seq_off": off,
            "group_off": np.asarray(goff, dtype=np.int64), "cells": cells, "relaxations": relax, "n_reads": len(lens),
            "n_states": np.asarray(n_states), "edges": [len(b["in_src"]) for b in baked]}


# ------------------------------------------------------------------------------- clock sampling
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ------------------------------------------------------------------------- reference / CPU arm
def _cpu_worker(args):
    """Build the models of a slice of loci ON THE REFERENCE ENGINE (oracle/_ref: the unmodified
    vendored pomegranate, compiled) and time model.viterbi(read) over their reads, one read at a
    time exactly as vntr_finder.py:727-767 does.  Falls back to the C oracle port if the compiled
    reference is unavailable."""
    ids, coverage, decoys, kind = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from advntr_b200 import read_matcher, synth
    reads_done, cells, seconds = 0, 0, 0.0
    if kind == "reference":
        import refenv
        pom = refenv.reference_pomegranate()
    else:
        import oracle
    alphabet = np.array(list("ACGT"))
    for lid in ids:
        loc = synth.config2_locus(lid, READ_LEN)
        flat, ln = synth.config2_read_codes(loc, coverage, decoys)
        off = np.concatenate([[0], np.cumsum(ln)])
        if kind == "reference":
            model = read_matcher.build_vntr_matcher_hmm(loc.left, loc.right, loc.segments, loc.copies,
                                                        flank_size=loc.flank, error_rate=loc.error_rate, pom=pom)
            m = len(model.states)
            strs = ["".join(alphabet[flat[off[i]:off[i + 1]]]) for i in range(len(ln))]
            t0 = time.perf_counter()
            for s in strs:
                model.viterbi(s)
            seconds += time.perf_counter() - t0
        else:
            model = read_matcher.build_vntr_matcher_hmm(loc.left, loc.right, loc.segments, loc.copies,
                                                        flank_size=loc.flank, error_rate=loc.error_rate)
            om = oracle.OracleModel(model.baked)
            m = om.n_states
            codes = [flat[off[i]:off[i + 1]] for i in range(len(ln))]
            t0 = time.perf_counter()
            om.viterbi(codes)
            seconds += time.perf_counter() - t0
        reads_done += len(ln)
        cells += int(ln.sum()) * m
    return reads_done, cells, seconds


def cpu_reference_pass(n_sample_loci, coverage, decoys, procs):
    """One bounded pass of the reference CPU path over `n_sample_loci` loci with `procs` processes
    (processes, not threads: viterbi() holds the GIL, hmm.pyx:1958).  Returns reads/s etc."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refenv
    kind = "reference" if refenv.have_reference_engine() else "port"
    ids = list(range(1, n_sample_loci + 1))
    slices = [ids[i::procs] for i in range(procs) if ids[i::procs]]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(slices)) as pool:
        res = pool.map(_cpu_worker, [(s, coverage, decoys, kind) for s in slices])
    wall = time.perf_counter() - t0
    reads = sum(r[0] for r in res)
    cells = sum(r[1] for r in res)
    busy = max(r[2] for r in res)          # slowest worker's decode time = the pass's duration
    return {"kind": kind, "reads": reads, "cells": cells, "seconds": busy, "wall": wall,
            "procs": len(slices), "reads_per_s": reads / busy, "gcups": cells / busy / 1e9}


def auto_sample_loci(procs):
    # ~155 reads/locus at ~200 reads/s/core: one locus per worker is ~0.8 s of decoding; aim at
    # roughly 15 s of CPU work per pass
    return max(procs * 16, 16)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = host_cores()
    n_loci = args.cpu_sample_loci or auto_sample_loci(procs)
    times, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_pass(n_loci, args.coverage, args.decoys, procs)
        if i >= args.warmup:
            times.append(last)
    reads_s = sum(t["reads"] for t in times) / sum(t["seconds"] for t in times)
    gcups = sum(t["cells"] for t in times) / sum(t["seconds"] for t in times) / 1e9
    sample = "%d config-2 loci (%d reads) per step, %d processes" % (n_loci, last["reads"], last["procs"])
    line = {"impl": "reference", "metric": "viterbi_reads_per_s", "value": reads_s, "unit": "reads/s",
            "gcups": gcups, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean([t["seconds"] for t in times])),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args, n_loci),
            "cpu_baseline": {"value": reads_s, "unit": "reads/s", "cores": last["procs"], "kind": last["kind"],
                             "sample": sample},
            "e2e": {"value": reads_s, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, n_loci):
    return {"workload": "config2: Illumina 150 bp reads vs synthetic hg19-shaped VNTR loci "
                        "(RU 6-70 bp, 150 bp flanks, 30x mapped + 50 decoys x 2 strands per locus)",
            "loci_per_gpu": n_loci, "read_length": READ_LEN, "coverage": args.coverage,
            "decoys_per_locus": args.decoys, "want_path": True,
            "l2": "inputs+traceback workspace exceed L2 (no explicit flush needed)"}


# ------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from advntr_b200 import engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback of the hot path")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner to stdout; the contract is ONE JSON line on stdout
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_build = time.time()
    wl = build_workload(rank, args.loci, args.coverage, args.decoys)
    # a dedicated (non-default) torch stream: the library launches on it, and the torch events that
    # time the region are recorded on the very same stream
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    ctx = engine.Context(device=local, stream=stream.cuda_stream)
    assert stream.cuda_stream != 0 and ctx.stream == stream.cuda_stream
    models = [engine.DeviceModel(ctx, b) for b in wl["baked"]]
    t_build = time.time() - t_build
    lib = engine.load_library()
    handles = (C.c_void_p * len(models))(*[m._h for m in models])
    R = wl["n_reads"]
    goff, off = wl["group_off"], wl["seq_off"]
    flags_dev = engine.WANT_PATH | engine.DEVICE_BUFFERS
    path_cap = int(off[-1]) + R * 96          # read length + ~silent states on a typical path

    # device-resident buffers (value) and pinned host buffers (e2e)
    d_seqs = torch.from_numpy(wl["seqs"]).cuda()
    d_logp = torch.empty(R, dtype=torch.float64, device="cuda")
    d_plen = torch.empty(R, dtype=torch.int32, device="cuda")
    d_poff = torch.empty(R, dtype=torch.int64, device="cuda")
    d_path = torch.empty(path_cap, dtype=torch.int32, device="cuda")
    d_total = torch.zeros(1, dtype=torch.int64, device="cuda")

    def step_device():
        rc = lib.advhmm_viterbi_multi(ctx._h, handles, len(models), goff.ctypes.data, d_seqs.data_ptr(),
                                      off.ctypes.data, R, flags_dev, d_logp.data_ptr(), d_plen.data_ptr(),
                                      d_poff.data_ptr(), d_path.data_ptr(), path_cap, d_total.data_ptr())
        engine._check(rc)

    h_seqs = torch.from_numpy(wl["seqs"]).pin_memory()
    h_logp = torch.empty(R, dtype=torch.float64).pin_memory()
    h_plen = torch.empty(R, dtype=torch.int32).pin_memory()
    h_poff = torch.empty(R, dtype=torch.int64).pin_memory()
    h_path = torch.empty(path_cap, dtype=torch.int32).pin_memory()
    h_total = C.c_int64(0)

    def step_host():
        rc = lib.advhmm_viterbi_multi(ctx._h, handles, len(models), goff.ctypes.data, h_seqs.data_ptr(),
                                      off.ctypes.data, R, engine.WANT_PATH, h_logp.data_ptr(), h_plen.data_ptr(),
                                      h_poff.data_ptr(), h_path.data_ptr(), path_cap, C.byref(h_total))
        engine._check(rc)

    sampler = ClockSampler(local)
    # ---- warm-up ------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    n_paths = int(d_total.item())
    if n_paths > path_cap or int((d_plen < 0).sum().item()) != 0:
        raise SystemExit("path buffer too small or impossible reads in the synthetic workload")

    # ---- self-check (outside every timed region): the state paths of the first loci, re-scored with
    #      the models' own tables, must give the returned log-probabilities bit for bit -------------
    verified = None
    if rank == 0 and args.verify_loci > 0:
        from advntr_b200 import path_utils
        nv = min(args.verify_loci, len(models))
        r_hi = int(goff[nv])
        h_lp = d_logp[:r_hi].cpu().numpy()
        h_pl = d_plen[:r_hi].cpu().numpy()
        h_po = d_poff[:r_hi].cpu().numpy()
        h_pa = d_path[:n_paths].cpu().numpy()
        ok = True
        for g in range(nv):
            a, b = int(goff[g]), int(goff[g + 1])
            codes = [wl["seqs"][off[r]:off[r + 1]] for r in range(a, b)]
            paths = [h_pa[h_po[r]:h_po[r] + h_pl[r]] for r in range(a, b)]
            sc = path_utils.rescore_paths(wl["baked"][g], codes, paths)
            ok = ok and bool(np.array_equal(sc.view(np.int64), h_lp[a:b].view(np.int64)))
        verified = {"loci": nv, "reads": r_hi, "paths_rescored_bit_exact": ok}
        if not ok:
            raise SystemExit("self-check failed: re-scored paths do not reproduce the log-probabilities")

    # ---- value: inputs resident in HBM, CUDA events on the launching stream ---------------------
    ctx.profile(True)
    ctx.profile_read()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    t1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    fill_ms, fill_n, bt_ms, bt_n = ctx.profile_read()
    ctx.profile(False)
    clocks = sampler.window(t0, t1)

    # ---- e2e: host buffers through the public C-ABI call, copies inside the timed region --------
    e2e_ms = None
    if not args.no_e2e:
        step_host()
        barrier()
        te0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - te0) * 1e3
        assert h_total.value == n_paths
        # the host-buffer call returns what the device-resident call computed, bit for bit
        assert torch.equal(h_logp.view(torch.int64), d_logp.cpu().view(torch.int64))
        assert torch.equal(h_plen, d_plen.cpu())
    # ---- extra (reported, not the headline): on-device reducers instead of paths; fp32 mode ----
    extra = {}
    if not args.no_e2e:
        summ = np.zeros(R, dtype=engine.SUMMARY_DTYPE)
        h_summ = torch.from_numpy(summ.view(np.int32).reshape(R, 8)).pin_memory()

        def step_summary():
            rc = lib.advhmm_viterbi_multi_summary(ctx._h, handles, len(models), goff.ctypes.data, h_seqs.data_ptr(),
                                                  off.ctypes.data, R, engine.WANT_SUMMARY, h_logp.data_ptr(),
                                                  h_plen.data_ptr(), h_poff.data_ptr(), None, 0, None,
                                                  h_summ.data_ptr())
            engine._check(rc)
        step_summary()
        torch.cuda.synchronize()
        ts = time.perf_counter()
        for _ in range(args.steps):
            step_summary()
        torch.cuda.synchronize()
        extra["e2e_summary_ms"] = (time.perf_counter() - ts) * 1e3 / args.steps

        def step_fp32():
            rc = lib.advhmm_viterbi_multi(ctx._h, handles, len(models), goff.ctypes.data, d_seqs.data_ptr(),
                                          off.ctypes.data, R, flags_dev | engine.FP32, d_logp.data_ptr(),
                                          d_plen.data_ptr(), d_poff.data_ptr(), d_path.data_ptr(), path_cap,
                                          d_total.data_ptr())
            engine._check(rc)
        step_fp32()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            step_fp32()
        f1.record(stream)
        torch.cuda.synchronize()
        extra["fp32_ms"] = f0.elapsed_time(f1) / args.steps
    sampler.stop()

    # ---- max over ranks -------------------------------------------------------------------------
    t = torch.tensor([ms, e2e_ms or 0.0, float(R), float(wl["cells"]), fill_ms, float(wl["relaxations"])],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    else:
        tmax = tsum = t
    ms_all, e2e_all = float(tmax[0]), float(tmax[1])
    reads_all, cells_all = float(tsum[2]), float(tsum[3])

    if rank == 0:
        K = args.steps
        value = reads_all * K / (ms_all * 1e-3)
        gcups = cells_all * K / (ms_all * 1e-3) / 1e9
        line = {"metric": "viterbi_reads_per_s", "value": value, "unit": "reads/s", "gcups": gcups,
                "edge_relaxations_per_s": float(tsum[5]) * K / (ms_all * 1e-3),
                "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_all / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config(args, args.loci),
                "reads_per_step": reads_all, "gpu_launches": int(launches), "clocks": clocks,
                "setup_s": round(t_build, 1)}
        if e2e_ms is not None:
            h2d = int(off[-1])
            d2h = R * (8 + 4 + 8) + n_paths * 4
            line["e2e"] = {"value": reads_all * K / (e2e_all * 1e-3), "unit": "reads/s",
                           "gcups": cells_all * K / (e2e_all * 1e-3) / 1e9,
                           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_all / K}
        if extra:
            line["extras"] = {
                "e2e_summary_only": {"value": R * world / (extra["e2e_summary_ms"] * 1e-3), "unit": "reads/s",
                                     "note": "host buffers, on-device path reducers (32 B/read) instead of "
                                             "full state paths; rank-0 time x n_gpus",
                                     "d2h_bytes_per_step": R * (8 + 4 + 32)},
                "fp32_mode": {"value": R * world / (extra["fp32_ms"] * 1e-3), "unit": "reads/s",
                              "note": "optional ADVHMM_FP32 mode, device-resident, full paths; tolerance "
                                      "and RU-count concordance in tests/test_gpu_parity.py::test_fp32_mode"}}
        if verified:
            line["verified"] = verified
        line["roofline"] = roofline(ctx, wl, fill_ms, fill_n, bt_ms, bt_n, ms, K)
        if world == 1 and not args.no_cpu_baseline:
            procs = host_cores()
            cb = cpu_reference_pass(args.cpu_sample_loci or auto_sample_loci(procs), args.coverage, args.decoys, procs)
            line["cpu_baseline"] = {"value": cb["reads_per_s"], "unit": "reads/s", "gcups": cb["gcups"],
                                    "cores": cb["procs"], "kind": cb["kind"],
                                    "sample": "%d reads of the first %d config-2 loci, %.1f s" %
                                              (cb["reads"], cb["reads"] // 155 or 1, cb["seconds"])}
        print(json.dumps(line))
    for m in models:
        m.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def roofline(ctx, wl, fill_ms, fill_n, bt_ms, bt_n, total_ms, steps):
    """Dominant kernel = banded_fill_kernel.  HBM view per the contract (algorithmic bytes per
    SURVEY.md section 8d: 1 traceback byte per DP cell + packed read + outputs) and, because the
    kernel is bound by the fp64 add/compare pipe rather than by HBM, the fp64 view as well."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, which = 6650.0, "fallback"
    if os.path.exists(peaks_path):
        try:
            hbm_peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    lens = np.diff(wl["seq_off"])
    m_per_read = np.repeat(wl["n_states"], np.diff(wl["group_off"]))
    # traceback 1 B per (position, state) + 2-bit packed read + logp, per read
    bytes_per_step = float((lens * m_per_read).sum()) + float(((lens + 3) // 4).sum()) + len(lens) * 8.0
    # fp64 pipe operations per read: 2 adds per edge into an emitting state, 1 per edge into a
    # silent state, 1 compare per edge; emitting share from the models
    ops = 0.0
    for b, r0, r1 in zip(wl["baked"], wl["group_off"][:-1], wl["group_off"][1:]):
        deg = np.diff(b["in_off"])
        e_emit = int(deg[:b["silent_start"]].sum())
        e_sil = int(deg[b["silent_start"]:].sum())
        ops += float(lens[r0:r1].sum()) * (2 * e_emit + e_sil + e_emit + e_sil)
    fill_s = fill_ms * 1e-3
    achieved = bytes_per_step * steps / fill_s / 1e9 if fill_s > 0 else None
    fp64_peak = ctx.fp64_add_peak()
    # measured DRAM bytes per read of this kernel from the committed ncu --set full capture
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_fill_traffic.json")
    if os.path.exists(tpath) and fill_n:
        try:
            traffic = json.load(open(tpath))["dram_bytes_per_read"] * len(lens) * steps / fill_n
        except Exception:
            traffic = None
    out = {"kernel": "banded_fill_kernel", "bound": "hbm",
           "achieved": achieved, "peak": hbm_peak,
           "peak_source": which, "unit": "GB/s", "frac": achieved / hbm_peak if achieved else None,
           "traffic": traffic, "algorithmic_bytes_per_launch": bytes_per_step * steps / max(fill_n, 1),
           "launches": int(fill_n), "avg_launch_ms": fill_ms / max(fill_n, 1),
           "share_of_step": fill_ms / total_ms if total_ms else None,
           "backtrack_share_of_step": bt_ms / total_ms if total_ms else None,
           "fp64": {"achieved_gops": ops * steps / fill_s / 1e9 if fill_s > 0 else None,
                    "peak_gops": fp64_peak, "peak_source": "measured (DADD microbenchmark, this run)",
                    "frac": (ops * steps / fill_s / 1e9) / fp64_peak if fill_s > 0 and fp64_peak else None,
                    "note": "the kernel's binding resource: fp64 add + compare issue, no FMA/tensor work"}}
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
