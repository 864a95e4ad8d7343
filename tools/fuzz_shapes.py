"""One-off extended run of tests/test_gpu_parity.py::test_random_loci_random_reads_vs_oracle over many seeds.

    python tools/fuzz_shapes.py [first_seed=1000] [n_seeds=40]      # 8 model shapes x 40 reads per seed
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import conftest  # noqa: F401,E402  (puts oracle/ on the path, builds the C restatement)
import test_gpu_parity as T  # noqa: E402
from advntr_b200 import engine  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ctx = engine.Context(device=0)
t0 = time.time()
class _Env(object):                      # stands in for pytest's monkeypatch
    def setenv(self, k, v):
        os.environ[k] = v


for seed in range(first, first + n):
    T.test_random_loci_random_reads_vs_oracle(ctx, seed)
    T.test_random_shapes_native_models_and_long_reads_vs_oracle(ctx, seed, _Env())
os.environ.pop("ADVHMM_LONG_WPR", None)
print("seeds %d..%d: %d model shapes (%d of them through the native compiler, with reads of up to 2,600 bases on the "
      "long-read kernel), %d reads, every routing bit-exact against the oracle (%.0f s)" % (
          first, first + n - 1, 14 * n, 6 * n, (8 * 40 + 6 * 10) * n, time.time() - t0))
