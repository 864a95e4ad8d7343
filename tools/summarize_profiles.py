"""Copy the round's bench lines from gpurun_out/ into profiles/ and write profiles/r2_scaling.md from them
(so that every number in the table is the number of a committed JSON line)."""
import json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def load(name):
    path = os.path.join(G, name + ".json")
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
    if not lines:
        return None
    open(os.path.join(P, name + ".json"), "w").write(lines[-1] + "\n")     # the JSON line only (older runs had an NCCL banner before it)
    return json.loads(lines[-1])


rows, strong, c5 = [], [], []
for N in (1, 2, 4, 8):
    d = load("r2_bench_n%d" % N)
    if d:
        rows.append((N, d))
        if d.get("strong"):
            strong.append((N, d["strong"]))
    c = load("r2_config5_n%d" % N)
    if c:
        c5.append((N, c))
for extra in ("r2_bench_reference_n1", "r2_config3_n1", "r2_frameshift_n1"):
    load(extra)
out = ["# Round 2 — scaling on one 8 x B200 box (bench.py lines in profiles/r2_bench_n*.json, r2_config5_n*.json)", "",
       "## Weak scaling: 6,719 config-2 loci PER GPU (the bench's default line)", "",
       "| GPUs | value (reads/s, device-resident) | x | e2e (reads/s, host buffers) | x | pipeline cold (loci/s, compile on the clock) | pipeline warm |",
       "|---|---|---|---|---|---|---|"]
base = rows[0][1] if rows else None
for N, d in rows:
    p = d.get("pipeline") or {}
    out.append("| %d | %.2f M | %.2f | %.2f M | %.2f | %s | %s |" % (
        N, d["value"] / 1e6, d["value"] / base["value"], d["e2e"]["value"] / 1e6, d["e2e"]["value"] / base["e2e"]["value"],
        "%.1f k" % (p["cold"]["loci_per_s"] / 1e3) if p else "-", "%.1f k" % (p["warm"]["loci_per_s"] / 1e3) if p else "-"))
out += ["", "## Strong scaling: the SAME 6,719 loci (1.04 M reads) split by LPT, results gathered on rank 0 inside the timed region", "",
        "| GPUs | ms per pass | reads/s | speed-up | efficiency | per-rank busy ms (min .. max) | busy spread | gathered table = single-rank decode |",
        "|---|---|---|---|---|---|---|---|"]
b = strong[0][1]["ms_per_step"] if strong else None
for N, s in strong:
    out.append("| %d | %.2f | %.2f M | %.2f | %.3f | %.2f .. %.2f | %.1f %% | %s |" % (
        N, s["ms_per_step"], s["value"] / 1e6, b / s["ms_per_step"], b / s["ms_per_step"] / N,
        min(s["rank_busy_ms"]), max(s["rank_busy_ms"]), 100 * s["rank_busy_spread"], s["gathered_equals_single_rank_decode"]))
out += ["", "## Config 5: 158,522 genic-like loci x 30x reads (37.3 M reads), sharded by locus (LPT), model compilation inside the timed region", "",
        "| GPUs | s per pass | reads/s | loci/s | speed-up | efficiency | per-rank busy s (max) | per-rank compile s (host, overlapped) |",
        "|---|---|---|---|---|---|---|---|"]
b5 = c5[0][1]["strong"]["ms_per_step"] if c5 else None
for N, c in c5:
    s = c["strong"]
    out.append("| %d | %.3f | %.2f M | %.1f k | %.2f | %.3f | %.3f | %.3f |" % (
        N, s["ms_per_step"] / 1e3, s["value"] / 1e6, s["loci_per_s"] / 1e3, b5 / s["ms_per_step"], b5 / s["ms_per_step"] / N,
        max(s["rank_busy_ms"]) / 1e3, max(s["rank_compile_ms"]) / 1e3))
open(os.path.join(P, "r2_scaling.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
