"""Forward (log_probability) throughput on config 1 through the host-buffer C-ABI call: banded
forward kernel vs the generic CSR forward kernel (wall clock around the call, H2D/D2H included)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from advntr_b200 import engine, synth
R = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
loc = synth.config1_locus()
model = loc.build_model()
base = [engine.encode_acgt(r)[0] for r in synth.config1_reads(1000)]
codes = [base[i % 1000] for i in range(R)]
ctx = engine.Context(device=0)
dm = engine.DeviceModel(ctx, model.baked)
cells = sum(len(c) for c in codes) * model.baked["n_states"]
for force, n in ((False, R), (True, min(R, 4000))):
    sub = codes[:n]
    dm.log_probability(sub[:512], force_generic=force)
    t = time.time(); lp = dm.log_probability(sub, force_generic=force); dt = time.time() - t
    c = cells * n / R
    print("forward %s: %d reads in %.3f s -> %.0f reads/s, %.2f GCUPS, checksum %r" % (
        "generic" if force else "banded ", n, dt, n / dt, c / dt / 1e9, float(lp[:1000].sum())))
