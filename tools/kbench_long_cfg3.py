"""Kernel-variant timing on config 3 as stated (10-20 kb reads vs the 18,918-state model), device-resident
buffers, CUDA events around every launch of the long-read kernel; no oracle (bench.py --workload config3 and the
GPU suite check parity).  ADVHMM_LIB selects a library built with other flags; checksums must be equal."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, bench_workloads
from advntr_b200 import engine, fast_compile, synth
if os.environ.get('ADVHMM_LIB'):
    engine.LIB_PATH = os.path.abspath(os.environ['ADVHMM_LIB'])
R = int(sys.argv[1]) if len(sys.argv) > 1 else 592
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = engine.Context(0, stream.cuda_stream)
lib = engine.load_library()
loc = synth.config3_locus()
model = fast_compile.compile_many([(loc.left, loc.right, loc.segments, loc.copies, 0.3)], ctx)[0]
dm = model._device_model()
reads = bench_workloads.config3_reads(loc, R, seed=31)
seqs, off = engine.pack_reads(reads)
cells = float(np.diff(off).sum()) * dm.info.n_states
handles = (C.c_void_p * 1)(dm._h)
goff = np.asarray([0, R], dtype=np.int64)
cap = int(off[-1]) + R * 512
d_seqs = torch.from_numpy(seqs).cuda()
d_logp = torch.empty(R, dtype=torch.float64, device="cuda"); d_plen = torch.empty(R, dtype=torch.int32, device="cuda")
d_poff = torch.empty(R, dtype=torch.int64, device="cuda"); d_path = torch.empty(cap, dtype=torch.int32, device="cuda")
d_total = torch.zeros(1, dtype=torch.int64, device="cuda")
def step():
    engine._check(lib.advhmm_viterbi_multi(ctx._h, handles, 1, goff.ctypes.data, d_seqs.data_ptr(), off.ctypes.data, R,
        engine.WANT_PATH | engine.DEVICE_BUFFERS, d_logp.data_ptr(), d_plen.data_ptr(), d_poff.data_ptr(), d_path.data_ptr(), cap,
        d_total.data_ptr()))
step(); torch.cuda.synchronize()
ctx.profile(True); ctx.profile_read()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(steps): step()
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
fm, fn, bm, bn = ctx.profile_read()
n_paths = int(d_total.item())
print("lib=%s config3 reads=%d: step %.2f ms (%.0f GCUPS) fill %.2f ms/step = %.0f GCUPS (%d launches) backtrack %.2f ms/step | "
      "checksums logp %r path_len %d paths %d" % (os.path.basename(engine.LIB_PATH), R, ms, cells / ms / 1e6, fm / steps,
      cells / (fm / steps) / 1e6, fn // steps, bm / steps, float(d_logp.sum().item()), int(d_plen.sum().item()),
      int(d_path[:n_paths].to(torch.int64).sum().item())))
