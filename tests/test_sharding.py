"""N > 1 host logic on CPU: world_size-2 gloo run of the locus sharding + host gather.  Each rank
"decodes" its loci with the oracle (this is a test: the product path needs a GPU) and the gathered
result must equal the single-process result."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from advntr_b200 import sharding


def test_lpt_is_a_balanced_partition():
    costs = [9, 7, 6, 5, 5, 4, 3, 1]
    owner, load = sharding.lpt_assign(costs, 3)
    assert sorted(sum((sharding.my_units(owner, r) for r in range(3)), [])) == list(range(len(costs)))
    assert max(load) - min(load) <= max(costs)
    assert sharding.lpt_assign(costs, 1)[0] == [0] * len(costs)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import oracle
    from conftest import Golden
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cases = [Golden(n) for n in ("small_a", "small_b", "divergent", "config1")]
    costs = [sharding.locus_cost(len(c.baked["in_src"]), [len(r) for r in c.reads]) for c in cases]
    owner, _ = sharding.lpt_assign(costs, world)
    local = {}
    for i in sharding.my_units(owner, rank):
        logp, paths = oracle.OracleModel(cases[i].baked).viterbi(cases[i].codes())
        local[i] = (logp, [None if p is None else p.tolist() for p in paths])
    full = sharding.gather_results(local, owner, rank, world)
    ok = all(np.array_equal(full[i][0].view(np.int64), c.logp.view(np.int64)) for i, c in enumerate(cases))
    ok = ok and all(full[i][1][j] == (None if c.path(j) is None else c.path(j).tolist())
                    for i, c in enumerate(cases) for j in range(len(c.reads)))
    ok = ok and len(set(owner)) == world
    open(os.path.join(out_dir, "rank%d.ok" % rank), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(os.path.join(str(tmp_path), "rank%d.ok" % r)).read() == "1"


def test_cost_estimate_matches_the_generators_and_balances_config2():
    """bench.py's strong-scaling legs balance the loci of a run from locus_cost_estimate alone (every
    rank computes the same assignment without generating anything): it must predict the read count the
    generator really makes, and LPT on it must split config 2 evenly."""
    import numpy as np
    from advntr_b200 import synth
    for gen, make in (("config2", synth.config2_locus), ("config5", synth.config5_locus)):
        for lid in (1, 2, 17, 444, 6719, 158522):
            loc = make(lid)
            _, lens = synth.config2_read_codes(loc, 30, 50)
            n, cells = synth.locus_cost_estimate(lid, gen, 150, 30, 50)
            assert n == len(lens), (gen, lid)
            m = 3 * 150 + 3 * 150 + loc.copies * (3 * len(loc.pattern) + 3) + 18
            assert cells == n * 150 * m
    est = [synth.locus_cost_estimate(i, "config2")[1] for i in range(1, 6720)]
    for world in (2, 4, 8):
        owner, load = sharding.lpt_assign(est, world)
        assert (max(load) - min(load)) / max(load) < 5e-3
        assert sorted(np.bincount(owner, minlength=world).tolist())[0] > 6719 // world - 60


def _gather_worker(rank, world, port, out_dir):
    import numpy as np
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from advntr_b200 import synth
    n_loci = 300
    est = [synth.locus_cost_estimate(i, "config2") for i in range(1, n_loci + 1)]
    owner, _ = sharding.lpt_assign([e[1] for e in est], world)      # every rank computes the same assignment
    owner = np.asarray(owner)
    reads = np.asarray([e[0] for e in est], dtype=np.int64)
    counts = [int(reads[owner == r].sum()) for r in range(world)]
    rows_max = max(counts)
    # this rank's result rows: (locus id, read index within the locus), its loci in ascending order
    mine = np.nonzero(owner == rank)[0]
    rows = torch.zeros((rows_max, 2), dtype=torch.int64)
    k = 0
    for u in mine:
        for j in range(int(reads[u])):
            rows[k, 0], rows[k, 1] = int(u), j
            k += 1
    table = torch.empty((world * rows_max, 2), dtype=torch.int64)
    dist.all_gather_into_tensor(table, rows)
    if rank == 0:
        idx = sharding.gathered_row_of_every_read(owner, reads, rows_max)
        got = table.numpy()[idx]
        want = np.array([(u, j) for u in range(n_loci) for j in range(int(reads[u]))], dtype=np.int64)
        np.save(os.path.join(out_dir, "ok.npy"), np.array([int(np.array_equal(got, want)), len(want), len(set(owner.tolist()))]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_of_result_rows(tmp_path):
    """The strong-scaling legs of bench.py gather one result row per read with ONE all_gather of the
    per-rank tables (NCCL on the GPUs; gloo here) and put them back into locus order with
    sharding.gathered_row_of_every_read: two ranks, 300 config-2 loci, every read's row must land where the
    single-rank order has it."""
    import numpy as np
    port = _free_port()
    mp.spawn(_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok, n_rows, ranks_used = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok == 1 and n_rows > 40000 and ranks_used == 2


def _calls_worker(rank, world, port, out_dir):
    import sys
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import test_locus_calls as T
    from advntr_b200 import engine
    # the same per-read results on every rank (seeded); each rank calls only the loci it owns
    logp, S, plen, off, goff, layout, scores = T._random_results(np.random.default_rng(77), 240, 0.9)
    reads = np.diff(goff)
    owner, _ = sharding.lpt_assign(reads.tolist(), world)
    score = np.array([np.nan if s is None else s for s in scores])
    local = {}
    for g in sharding.my_units(owner, rank):
        a, b = int(goff[g]), int(goff[g + 1])
        rec, _ = engine.genotypes_from_summaries([0, b - a], [layout[g][0]], [layout[g][1]], score[g:g + 1], logp[a:b], S[a:b],
                                                 plen[a:b], off[a:b + 1], threads=1)
        local[g] = rec.tobytes()
    full = sharding.gather_results(local, owner, rank, world)
    whole, _ = engine.genotypes_from_summaries(goff, [m for m, _ in layout], [u for _, u in layout], score, logp, S, plen, off)
    ok = b"".join(full) == whole.tobytes() and int(whole["has_call"].sum()) > 100 and len(set(owner)) == world
    open(os.path.join(out_dir, "calls%d.ok" % rank), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_two_rank_gloo_locus_calls_gather(tmp_path):
    """The step after the decode shards like the decode: every rank calls its own loci with the native stage,
    the gathered records equal one call over all loci."""
    port = _free_port()
    mp.spawn(_calls_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(os.path.join(str(tmp_path), "calls%d.ok" % r)).read() == "1"
