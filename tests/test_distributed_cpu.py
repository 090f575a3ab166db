"""world_size-2 ``gloo`` tests (CPU) of the multi-GPU host logic: row sharding, record reduction with
first-row tie-breaks, ragged gathers and candidate ordering.  Each rank derives the records its GPU
kernels would produce from its shard of a golden fixture's Q with NumPy, combines them through
``safeopt_b200.distributed`` and must reproduce the single-process reference result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _local_safe_record(Q, S, row0):
    rows = np.flatnonzero(S)
    if rows.size == 0:
        return 0, -np.inf, -1, -np.inf, -1
    l, u = Q[rows, 0], Q[rows, 1]
    return rows.size, l.max(), row0 + rows[np.argmax(l)], u.max(), row0 + rows[np.argmax(u)]


def _local_max_record(Q, S, max_l0, scaling, row0):
    M = S & (Q[:, 1] >= max_l0)
    rows = np.flatnonzero(M)
    if rows.size == 0:
        return M, (0, -np.inf, -np.inf, -1)
    w = (Q[rows, 1::2] - Q[rows, ::2])
    val = np.max(w / scaling, axis=1)
    return M, (rows.size, (Q[rows, 1] - Q[rows, 0]).max(), val.max(), row0 + rows[np.argmax(val)])


def _worker(rank, world, port, fixture, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import load_golden, unpack_mask
        from safeopt_b200 import distributed as D
        g = load_golden(fixture)
        n_rows = int(g["n_rows"])
        Q = g["Q"]
        fmin = g["fmin"]
        G = Q.shape[1] // 2
        scaling = np.full(G, np.sqrt(float(g["variance"])))
        comm = D.Comm()
        assert comm.world == world and comm.rank == rank and comm.active
        r0, r1 = D.shard_bounds(n_rows, world, rank)
        Ql = Q[r0:r1]
        Sl = np.all(Ql[:, ::2] > fmin, axis=1)
        from safeopt_b200.engine import MAX_REC_DTYPE, SAFE_REC_DTYPE
        rec = np.zeros(1, dtype=SAFE_REC_DTYPE)
        rec["n_safe"], rec["max_l0"], rec["argmax_l0"], rec["max_u0"], rec["argmax_u0"] = _local_safe_record(Ql, Sl, r0)
        # the record travels as 64 raw bytes, exactly like the device-resident one
        safe = D.reduce_safe_records(comm.gather_records(torch.from_numpy(rec.view(np.uint8).copy()), SAFE_REC_DTYPE))
        S_ref = unpack_mask(g["S"], n_rows)
        assert safe["n_safe"] == int(S_ref.sum())
        rows_ref = np.flatnonzero(S_ref)
        assert safe["max_l0"] == Q[S_ref, 0].max() and safe["argmax_l0"] == rows_ref[np.argmax(Q[S_ref, 0])]
        assert safe["argmax_u0"] == rows_ref[np.argmax(Q[S_ref, 1])] == int(g["row_ucb"])
        Ml, rec = _local_max_record(Ql, Sl, safe["max_l0"], scaling, r0)
        mrec = np.zeros(1, dtype=MAX_REC_DTYPE)
        mrec["n_max"], mrec["max_width0"], mrec["best_value"], mrec["best_row"] = rec
        mx = D.reduce_max_records(comm.gather_records(torch.from_numpy(mrec.view(np.uint8).copy()), MAX_REC_DTYPE), scaling[0])
        M_ref = unpack_mask(g["M"], n_rows)
        assert mx["n_max"] == int(M_ref.sum())
        assert mx["max_var"] == np.max(Q[M_ref, 1] - Q[M_ref, 0]) / scaling[0]
        # reassemble sharded masks
        assert np.array_equal(D.gather_row_blocks(comm, Sl.astype(np.uint8), n_rows).astype(bool), S_ref)
        assert np.array_equal(D.gather_row_blocks(comm, Ml.astype(np.uint8), n_rows).astype(bool), M_ref)
        assert np.array_equal(D.gather_row_blocks(comm, Ql, n_rows), Q)
        # candidates: s = S & ~M & wide enough; global order = key desc, row desc on ties (stable argsort reversed)
        beta, thr = float(g["beta"]), float(g["threshold"])
        w = Ql[:, 1::2] - Ql[:, ::2]
        s = Sl & ~Ml & (np.max(w / scaling, axis=1) > mx["max_var"]) & np.any(w > thr * beta, axis=1)
        rows = np.flatnonzero(s) + r0
        keys = np.max(w[s], axis=1)
        o = np.lexsort((-rows, -keys))
        order = D.order_candidates(comm, rows[o], keys[o])
        wf = Q[:, 1::2] - Q[:, ::2]
        sf = S_ref & ~M_ref & (np.max(wf / scaling, axis=1) > mx["max_var"]) & np.any(wf > thr * beta, axis=1)
        rf, kf = np.flatnonzero(sf), np.max(wf[sf], axis=1)
        assert np.array_equal(order, rf[np.argsort(kf, kind="stable")[::-1]])
        # the no-expander query point: best scaled width over M, first row on ties
        G_ref = unpack_mask(g["G"], n_rows)
        if not G_ref.any():
            assert mx["best_row"] == int(g["row_next"])
        # flags OR across ranks
        f = np.zeros(8, dtype=np.uint8)
        f[rank] = 1
        assert np.array_equal(comm.any_flags(f)[:world], np.ones(world, dtype=np.uint8))
        assert np.array_equal(D.gather_ragged(comm, np.arange(rank + 1, dtype=np.int64)), np.concatenate([np.arange(r + 1) for r in range(world)]))
        open(os.path.join(out_dir, "ok_%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fixture", ["config_C3_n80", "expander_g2", "doctest_1d"])
def test_sharded_record_reduction_gloo(fixture, tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, fixture, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok_%d" % r)) for r in range(world))


def _swarm_worker(rank, world, port, out_dir):
    """Host logic of the sharded swarm: random-block slicing, padded row gather (positions + keep flags of the
    safe-set insertion) and the best-record exchange, with CPU tensors over gloo."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from safeopt_b200 import _lib
        from safeopt_b200 import distributed as D
        comm = D.Comm()
        P, d = 11, 3                                     # odd size: the last block is short
        p0, p1 = D.shard_bounds(P, world, rank)
        np.random.seed(9)                                # every rank draws the same stream
        full = np.random.rand(2 * P, d)
        mine = D.shard_stacked_blocks(full, 2, P, p0, p1)
        assert mine.shape == (2 * (p1 - p0), d)
        assert np.array_equal(mine[:p1 - p0], full[p0:p1]) and np.array_equal(mine[p1 - p0:], full[P + p0:P + p1])
        pos = np.arange(P * d, dtype=float).reshape(P, d)
        keep = (np.arange(P) % 3 != 0).astype(np.uint8)
        got_pos = D.gather_padded_rows(comm, torch.from_numpy(pos[p0:p1].copy()), P).numpy()
        got_keep = D.gather_padded_rows(comm, torch.from_numpy(keep[p0:p1].copy()), P).numpy()
        assert np.array_equal(got_pos, pos) and np.array_equal(got_keep, keep)
        # best-record exchange: {value, global index, position}; equal values -> lowest global index wins
        rec = torch.zeros(_lib.SWARM_REC_DOUBLES, dtype=torch.float64)
        rec[0], rec[1] = 2.5, float(p0 + 1)
        rec[2:2 + d] = torch.from_numpy(pos[p0 + 1])
        recs = torch.zeros((world, _lib.SWARM_REC_DOUBLES), dtype=torch.float64)
        comm.all_gather_into(recs, rec)
        r = recs.numpy()
        assert np.array_equal(r[:, 1], [D.shard_bounds(P, world, k)[0] + 1 for k in range(world)])
        v, idx = D.combine_max_first(r[:, 0], r[:, 1].astype(np.int64))
        assert (v, idx) == (2.5, 1) and np.array_equal(r[0, 2:2 + d], pos[1])
        open(os.path.join(out_dir, "ok_%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_swarm_host_logic_gloo(tmp_path):
    world = 2
    mp.spawn(_swarm_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok_%d" % r)) for r in range(world))


def _orchestration_worker(rank, world, port, fixture, out_dir):
    """The whole sharded optimize() of the product's host code on CPU: every rank drives the NumPy stand-in engine
    (tests/fake_engine.py) on its row block, records and flags cross the ranks over gloo."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import golden_lipschitz, golden_problem, load_golden, unpack_mask
        from fake_engine import FakeEngine
        import safeopt_b200 as sb
        from safeopt_b200 import gp_opt
        gp_opt.DeviceEngine = FakeEngine
        g = load_golden(fixture)
        gps, grid, fmin = golden_problem(g, "gpu")
        n_rows = int(g["n_rows"])
        opt = sb.SafeOpt(gps if len(gps) > 1 else gps[0], grid, fmin if len(gps) > 1 else fmin[0], lipschitz=golden_lipschitz(g),
                         beta=float(g["beta"]), threshold=float(g["threshold"]))
        assert opt._comm.world == world and opt._row1 - opt._row0 < n_rows
        x = opt.optimize()
        ok = opt.last_query_row == int(g["row_next"]) and np.array_equal(x, g["x_next"])
        ok = ok and np.abs(opt.Q - g["Q"]).max() < 1e-9
        ok = ok and np.array_equal(opt.S, unpack_mask(g["S"], n_rows)) and np.array_equal(opt.M, unpack_mask(g["M"], n_rows))
        ok = ok and np.array_equal(opt.G, unpack_mask(g["G"], n_rows))
        mx = opt.get_maximum()
        ok = ok and np.array_equal(mx[0], g["max_x"]) and abs(mx[1] - float(g["max_val"])) < 1e-9
        open(os.path.join(out_dir, "ok_%d" % rank), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fixture,world", [("expander_g2", 2), ("expander_tight", 3), ("lipschitz_g2", 2), ("matern32_3d", 2)])
def test_sharded_optimize_host_orchestration_gloo(fixture, world, tmp_path):
    mp.spawn(_orchestration_worker, args=(world, _free_port(), fixture, str(tmp_path)), nprocs=world, join=True)
    assert all(open(os.path.join(str(tmp_path), "ok_%d" % r)).read() == "1" for r in range(world))


def _swarm_orchestration_worker(rank, world, port, fixture, out_dir):
    """SafeOptSwarm with the device swarm back end sharded over the ranks (stand-in engine, gloo): particle blocks, the
    per-iteration best-record exchange and the gathered safe-set insertion must follow the reference's trajectory."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import load_golden
        from fake_engine import FakeEngine
        import safeopt_b200 as sb
        from safeopt_b200 import gp_opt
        from safeopt_b200 import distributed as D
        gp_opt.DeviceEngine = FakeEngine
        g = load_golden(fixture)
        X, Y = g["X"], g["Y"]
        d = X.shape[1]
        cls = {0: sb.RBF, 1: sb.Matern32, 2: sb.Matern52}[int(g["kind"])]
        gps = [sb.GPRegression(X, Y[:, [i]], kernel=cls(d, variance=float(g["variance"]), lengthscale=g["lengthscale"], ARD=True),
                               noise_var=float(g["noise_var"])) for i in range(Y.shape[1])]
        opt = sb.SafeOptSwarm(gps, list(g["fmin"]), bounds=[tuple(b) for b in g["bounds"]], beta=float(g["beta"]),
                              swarm_size=int(g["swarm_size"]), swarm_backend="device", rng="host")
        opt.max_iters = int(g["max_iters"])
        np.random.seed(int(g["seed"]))
        x = opt.optimize()
        sw = opt.swarms["expanders"]
        ok = sw.comm.world == world and sw.p1 - sw.p0 < int(g["swarm_size"])
        ok = ok and np.abs(x - g["x_next"]).max() < 1e-9 and opt.S.shape == g["S_final"].shape and np.abs(opt.S - g["S_final"]).max() < 1e-9
        full = D.gather_padded_rows(sw.comm, sw.best_positions, int(g["swarm_size"])).numpy()
        ok = ok and np.abs(full - g["expanders_best_positions"]).max() < 1e-9
        open(os.path.join(out_dir, "ok_%d" % rank), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fixture,world", [("swarm_query_2d", 2), ("swarm_query_2d_mat32", 3)])
def test_sharded_swarm_host_orchestration_gloo(fixture, world, tmp_path):
    mp.spawn(_swarm_orchestration_worker, args=(world, _free_port(), fixture, str(tmp_path)), nprocs=world, join=True)
    assert all(open(os.path.join(str(tmp_path), "ok_%d" % r)).read() == "1" for r in range(world))


def test_single_rank_comm_is_identity():
    from safeopt_b200 import distributed as D
    comm = D.Comm()
    assert not comm.active and comm.world == 1
    a = np.arange(6.0).reshape(2, 3)
    assert np.array_equal(comm.all_gather(a)[0], a)
    assert np.array_equal(D.gather_row_blocks(comm, a, 2), a)
    from safeopt_b200.engine import SAFE_REC_DTYPE
    rec = np.zeros(1, dtype=SAFE_REC_DTYPE)
    rec["n_safe"], rec["max_l0"], rec["argmax_l0"], rec["max_u0"], rec["argmax_u0"] = 3, 1.5, 7, 2.5, 9
    got = D.reduce_safe_records(comm.gather_records(torch.from_numpy(rec.view(np.uint8).copy()), SAFE_REC_DTYPE))
    assert got == dict(n_safe=3, max_l0=1.5, argmax_l0=7, max_u0=2.5, argmax_u0=9)
