"""Host-side logic that needs no GPU: synthetic workload generators, sharding, the import shim."""
import numpy as np
import pytest

from oracle import vecode_oracle as po


def test_splitmix_matches_python_oracle(vo):
    w = vo.workloads
    idx = np.array([0, 1, 2, 12345, 2**40 + 7], dtype=np.uint64)
    assert [int(v) for v in w.splitmix64(42, idx)] == [po.splitmix64(42, int(i)) for i in idx]
    assert list(w.uniform01(7, idx)) == [po.uniform01(7, int(i)) for i in idx]


def test_lorenz_x0_is_shard_invariant(vo):
    w = vo.workloads
    full = w.lorenz_x0(1000)
    parts = []
    for r in range(3):
        lo, hi = w.shard_range(1000, r, 3)
        parts.append(w.lorenz_x0(hi - lo, first=lo))
    assert np.array_equal(np.concatenate(parts), full)
    assert np.all(np.abs(full - 1.0) <= 1e-3)


def test_vdp_mu_sweep_and_shards(vo):
    w = vo.workloads
    mu = w.vdp_mu(1001)
    assert mu[0] == 0.5 and mu[-1] == 20.0
    lo, hi = w.shard_range(1001, 1, 4)
    assert np.array_equal(w.vdp_mu(1001, hi - lo, lo), mu[lo:hi])


def test_shard_range_covers_everything(vo):
    for n, g in [(10**6, 8), (7, 8), (100001, 4), (1, 1)]:
        spans = [vo.workloads.shard_range(n, r, g) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(g - 1))


def test_schrodinger_system_is_hermitian(vo):
    H0, H1 = vo.workloads.schrodinger_system(64)
    assert np.array_equal(H0, H0.conj().T) and np.allclose(H1, H1.conj().T, atol=0)
    gp = vo.workloads.schrodinger_drive(100)
    assert gp.shape == (100, 1, 3) and np.all(gp[:, 0, 0] >= 0.5) and np.all(gp[:, 0, 1] < 3.0)


def test_chunk_range_partitions_and_tapers(vo):
    """pipeline.chunk_range: the chunks of a pipelined solve tile [0, n) exactly, in order, and get smaller towards the end (the last
    chunk's transfer is the one nothing overlaps)."""
    from vecode_b200.pipeline import chunk_range
    for n in (1_000_000, 125_000, 40_003, 1001, 7, 3, 1):
        for parts in (1, 2, 3, 4, 8):
            r = [chunk_range(n, q, parts) for q in range(parts)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(parts - 1)), (n, parts, r)
            sizes = [b - a for a, b in r]
            assert all(x >= 0 for x in sizes)
            if n >= 1000 and parts > 1:
                assert sizes == sorted(sizes, reverse=True) and sizes[-1] < sizes[0]
    assert [b - a for a, b in (chunk_range(1_000_000, q, 4) for q in range(4))] == [320000, 280000, 240000, 160000]


def test_bench_taylor_plan_matches_the_oracles_plan():
    """bench.py's vectorised Taylor plan (the algorithmic term count of the roofline of config 5) against the pure-Python oracle's
    plan, which the kernels are tested against."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from oracle import exp_oracle as eo
    import numpy as np
    thetas = np.array([1e-9, 1e-3, 0.05, 0.3, 0.477, 0.999, 1.0, 1.0001, 1.7, 2.0, 3.3, 6.02, 11.5])
    want = [np.prod(eo.BasisSplit.plan(float(t))) for t in thetas]
    assert list(bench._plan_terms(thetas)) == [float(w) for w in want]


def test_odestep_mirror_and_check_step():
    """ODEStep<T> (src/base/ode.rs:41-77) and check_step (ode.rs:389-399) on the host: the enum's helpers behave as the reference's, and
    check_step agrees with the oracle's restatement on its edge cases (zero remainder, remainder below / above dt, negative micro-step)."""
    import math
    import vecode_b200 as vo
    from oracle import vecode_oracle as po
    s = vo.ODEStep.Step(0.25)
    assert s.kind == "Step" and s.unwrap_dt() == 0.25 and s.unwrap_dt_or(1.0) == 0.25
    seen = []
    assert s.map_dt(seen.append) is s and seen == [0.25]                       # Ok(_) keeps the step

    def fail(dt):
        raise vo.ODEError("rhs failed")
    e = s.map_dt(fail)
    assert e.kind == "Err" and e.err.msg == "rhs failed"                       # Err(e) => ODEStep::Err(e)
    for other in (vo.ODEStep.Chkpt, vo.ODEStep.Reject, vo.ODEStep.End):
        assert other.map_dt(fail) is other and other.unwrap_dt_or(3.0) == 3.0  # every other variant passes through
        try:
            other.unwrap_dt()
            assert False
        except RuntimeError as ex:
            assert "expected Step(T)" in str(ex)
    eps = 2.220446049250313e-16
    cases = [(0.0, 1.0, 0.1), (0.95, 1.0, 0.1), (1.0, 1.0, 0.1), (1.0 - eps / 2, 1.0, 0.1), (1.0 + 1e-12, 1.0, 0.1), (0.0, 1e-300, 0.1),
             (3.0, 3.0 + 4 * eps, 1.0), (math.pi, 10.0, 7.0)]
    for t0, tf, dt in cases:
        assert vo.check_step(t0, tf, dt) == po.check_step(t0, tf, dt), (t0, tf, dt)


@pytest.mark.parametrize("interleave", [False, True])
def test_sharded_chunk_plan_tiles_the_ensemble(vo, monkeypatch, interleave):
    """pipeline.ShardedChunkedSolve's plan (which local trajectories form chunk q on rank r, and where the root places them) for ragged
    ensemble sizes, every rank of 1..8 and 1..5 chunks, with the device objects replaced by stand-ins: every rank derives the SAME gather
    arguments for a chunk, the per-rank counts match what the gather entry points expect (vo_group_gather_placed: rows[rank];
    vo_group_gather_interleaved: ceil((tot - rank) / G)), and the chunks tile [0, n_total) exactly once."""
    from vecode_b200 import pipeline
    from vecode_b200.workloads import shard_range

    class FakeCtx:
        device = 0

        def __init__(self, *a, **k):
            pass

    class FakeEns:
        def __init__(self, ctx, d, n):
            self.d, self.n = d, n

    class FakeGroup:
        def __init__(self, rank, world):
            self.ctxs, self.ranks, self.world = [FakeCtx()], [rank], world

    monkeypatch.setattr(pipeline, "Context", FakeCtx)
    monkeypatch.setattr(pipeline, "Ensemble", FakeEns)
    for n_total in (1, 7, 64, 1000, 1001, 12345):
        for G in (1, 2, 3, 4, 8):
            for parts in (1, 2, 4, 5):
                made = {}
                plans = []
                for r in range(G):
                    sc = pipeline.ShardedChunkedSolve(FakeGroup(r, G), n_total, 2, lambda ctx, lo, hi, x0, r=r: made.setdefault((r, lo, hi), x0.n), parts=parts,
                                                      interleave=interleave)
                    plans.append(sc)
                    sc.pool.shutdown()
                    # this rank's chunks tile its shard in order
                    edges = [(lo, hi) for lo, hi, _ in sc.plan]
                    full = [e for e in edges if e[1] > e[0]]  # (an empty chunk of the round-robin plan carries no meaningful offset)
                    assert (not full and sc.n_local == 0) or (full[0][0] == 0 and full[-1][1] == sc.n_local and all(a[1] == b[0] for a, b in zip(full, full[1:]))), \
                        (n_total, G, parts, r, edges)
                    assert sum(hi - lo for lo, hi in edges) == sc.n_local
                    for (lo, hi), ch in zip(edges, sc.chunks):
                        assert (ch is None) == (hi <= lo) and (ch is None or (ch[0], ch[1], ch[3].n) == (lo, hi, hi - lo))
                covered = np.zeros(n_total, dtype=np.int64)
                for q in range(parts):
                    args = [p.plan[q][2] for p in plans]
                    assert all(a == args[0] for a in args)  # a collective: same arguments on every rank
                    if interleave:
                        tot, row0 = args[0]
                        assert 0 <= row0 and row0 + tot <= n_total
                        covered[row0:row0 + tot] += 1
                        for r, p in enumerate(plans):
                            lo, hi, _ = p.plan[q]
                            assert hi - lo == max(0, -(-(tot - r) // G))           # what vo_group_gather_interleaved expects from rank r
                            assert tot == 0 or hi == lo or row0 == lo * G           # local index i of rank r is trajectory r + G i
                    else:
                        rows, off = args[0]
                        for r, p in enumerate(plans):
                            lo, hi, _ = p.plan[q]
                            slo, shi = shard_range(n_total, r, G)
                            assert rows[r] == hi - lo and (rows[r] == 0 or off[r] == slo + lo) and off[r] + rows[r] <= n_total
                            covered[off[r]:off[r] + rows[r]] += 1
                assert np.all(covered == 1), (n_total, G, parts, interleave)
                assert sum(p.n_local for p in plans) == n_total
