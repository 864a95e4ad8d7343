"""BASELINE config 3 on the device: PacBio-like reads vs a 60 bp x 100-copy model (18,918 states).
Checks two reads against the oracle and times the striped long-read kernel."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle
from advntr_b200 import engine, synth
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 64
loc = synth.config3_locus(copies=100)
t = time.time(); model = loc.build_model(); t_build = time.time() - t
rng = random.Random(9)
ru = loc.pattern
reads = [synth.sequencing_errors(rng, loc.left + ru * rng.randint(40, 100) + loc.right, 0.02, 0.05, 0.05) for _ in range(n_reads)]
codes = [synth.encode(r) for r in reads]
ctx = engine.Context(0)
dm = engine.DeviceModel(ctx, model.baked)
print("model: %d states, %d edges, %d columns, built in %.2f s; kind %s, image %d KB" % (
    model.baked["n_states"], len(model.baked["in_src"]), dm.info.n_columns, t_build, dm.kind, dm.info.smem_bytes // 1024))
dm.viterbi(codes[:4])
ctx.profile(True); ctx.profile_read()
t = time.time(); res = dm.viterbi(codes); dt = time.time() - t
fm, fn, bm, bn = ctx.profile_read()
cells = sum(len(c) for c in codes) * model.baked["n_states"]
print("%d reads (%.0f bp mean): %.3f s wall, fill %.1f ms (%d launches), backtrack %.1f ms -> %.1f reads/s, %.1f GCUPS (kernel: %.1f GCUPS)" % (
    n_reads, np.mean([len(c) for c in codes]), dt, fm, fn, bm, n_reads / dt, cells / dt / 1e9, cells / (fm * 1e-3) / 1e9))
t = time.time(); lp, paths = oracle.OracleModel(model.baked).viterbi(codes[:2]); t_cpu = time.time() - t
ok = all(res.logp[i] == lp[i] and np.array_equal(res.path(i), paths[i]) for i in range(2))
print("oracle check on 2 reads: %s (CPU port %.1f s/read = %.4f GCUPS)" % ("bit-exact" if ok else "MISMATCH", t_cpu / 2,
      sum(len(c) for c in codes[:2]) * model.baked["n_states"] / t_cpu / 1e9))
