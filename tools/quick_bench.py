"""Scratch timing of the host-buffer API on config 1 (wall clock; not the judged bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from advntr_b200 import engine, synth
R = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
loc = synth.config1_locus()
model = loc.build_model()
base = synth.config1_reads(1000)
codes1k = [engine.encode_acgt(r)[0] for r in base]
codes = [codes1k[i % 1000] for i in range(R)]
ctx = engine.Context(device=0)
dm = engine.DeviceModel(ctx, model.baked)
print("kind", dm.kind, "cols", dm.info.n_columns, "smem", dm.info.smem_bytes)
for want_path in (True, False):
    for force in (False, True):
        if force and R > 4000:
            sub = codes[:4000]
        else:
            sub = codes
        dm.viterbi(sub[:256], want_path=want_path, force_generic=force)
        t = time.time(); res = dm.viterbi(sub, want_path=want_path, force_generic=force); dt = time.time() - t
        cells = sum(len(c) for c in sub) * model.baked["n_states"]
        print("want_path=%s generic=%s: %d reads in %.3fs -> %.0f reads/s, %.2f GCUPS" % (want_path, force, len(sub), dt, len(sub)/dt, cells/dt/1e9))
