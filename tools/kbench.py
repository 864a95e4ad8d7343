"""Kernel-variant timing on the config-2 workload (device-resident buffers, CUDA events around
every fill launch).  ADVHMM_LIB selects a library built with other -D variant flags; the checksums
(scores, path lengths, path entries) must be equal across variants."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from advntr_b200 import engine
if os.environ.get('ADVHMM_LIB'):
    engine.LIB_PATH = os.path.abspath(os.environ['ADVHMM_LIB'])
n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = bench.build_workload(range(1, n_loci + 1), 30, 50, "config2", 8)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = engine.Context(0, stream.cuda_stream)
models = ctx.compile_loci(wl["cols"])
stats = bench.model_stats(models, wl)
lib = engine.load_library()
handles = (C.c_void_p * len(models))(*[m._h for m in models])
R = wl["n_reads"]; goff, off = wl["group_off"], wl["seq_off"]
cap = int(off[-1]) + R * 96
d_seqs = torch.from_numpy(wl["seqs"]).cuda()
d_logp = torch.empty(R, dtype=torch.float64, device="cuda"); d_plen = torch.empty(R, dtype=torch.int32, device="cuda")
d_poff = torch.empty(R, dtype=torch.int64, device="cuda"); d_path = torch.empty(cap, dtype=torch.int32, device="cuda")
d_total = torch.zeros(1, dtype=torch.int64, device="cuda")
def step(flags):
    engine._check(lib.advhmm_viterbi_multi(ctx._h, handles, len(models), goff.ctypes.data, d_seqs.data_ptr(), off.ctypes.data, R,
        flags, d_logp.data_ptr(), d_plen.data_ptr(), d_poff.data_ptr(), d_path.data_ptr(), cap, d_total.data_ptr()))
F = engine.WANT_PATH | engine.DEVICE_BUFFERS | (engine.FP32 if os.environ.get('ADVHMM_PRECISION') == 'fp32' else 0)
for _ in range(2): step(F)
torch.cuda.synchronize()
ctx.profile(True); ctx.profile_read()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(steps): step(F)
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
fm, fn, bm, bn = ctx.profile_read()
n_paths = int(d_total.item())
print("lib=%s prec=%s loci=%d reads=%d: step %.2f ms (%.2f Mreads/s, %.0f GCUPS) fill %.2f ms/step (%d launches) backtrack %.2f ms/step | "
      "checksums logp %r path_len %d paths %d" % (
    os.path.basename(engine.LIB_PATH), os.environ.get("ADVHMM_PRECISION", "fp64"), n_loci, R, ms, R / ms / 1e3, stats["cells"] / ms / 1e6,
    fm / steps, fn // steps, bm / steps, float(d_logp.sum().item()), int(d_plen.sum().item()),
    int(d_path[:n_paths].to(torch.int64).sum().item())))
