"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: candidate sharding + winner
all-gather with the deterministic tie-break, and rating-block sharding + all-reduce with the
prior term counted once.  The per-shard numerics come from the oracle here (no GPU needed);
the same combine code runs under NCCL in bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from active_matrix_factorization_b200 import parallel as P


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pmf_oracle as O
        from active_matrix_factorization_b200 import _native as N
        rng = np.random.RandomState(0)                       # same problem on every rank
        n, m, d, nnz, ncand = 40, 30, 4, 600, 1001
        R = np.column_stack((rng.randint(0, n, nnz), rng.randint(0, m, nnz), rng.normal(size=nnz)))
        U, V = rng.normal(size=(n, d)), rng.normal(size=(m, d))
        ci, cj = rng.randint(0, n, ncand), rng.randint(0, m, ncand)
        ci[700], cj[700] = ci[3], cj[3]                      # an exact tie across shards

        # ---- gradient: rating blocks, prior once -------------------------------------------
        lo, hi = P.shard_bounds(nnz, world, rank)
        prm = P.prior_once_params(N.PmfParams(.7, 5., 9., 0.), rank)
        gu, gv = O.gradient(R[lo:hi], U, V, sigma_sq=prm.sigma_sq, sigma_u_sq=prm.sigma_u_sq,
                            sigma_v_sq=prm.sigma_v_sq)
        resid = R[lo:hi, 2] - O.predictions(R[lo:hi], U, V)
        sums = torch.tensor([float(resid @ resid), float((U * U).sum()), float((V * V).sum())],
                            dtype=torch.float64)
        dU, dV = torch.from_numpy(gu.copy()), torch.from_numpy(gv.copy())
        P.combine_loss_grad(dU, dV, sums, world, rank)
        fu, fv = O.gradient(R, U, V, sigma_sq=.7, sigma_u_sq=5., sigma_v_sq=9.)
        full = R[:, 2] - O.predictions(R, U, V)
        assert np.allclose(dU.numpy(), fu, rtol=1e-12, atol=1e-12)
        assert np.allclose(dV.numpy(), fv, rtol=1e-12, atol=1e-12)
        assert np.allclose(sums.numpy(), [full @ full, (U * U).sum(), (V * V).sum()], rtol=1e-12)
        # same with both gradients in one allocation: a single collective carries dU and dV
        fU, fV, flat = P.alloc_grads(torch.from_numpy(U), torch.from_numpy(V))
        fU.copy_(torch.from_numpy(gu)); fV.copy_(torch.from_numpy(gv))
        s2 = torch.tensor([float(resid @ resid), float((U * U).sum()), float((V * V).sum())],
                          dtype=torch.float64)
        P.combine_loss_grad(fU, fV, s2, world, rank, grads_flat=flat)
        assert np.allclose(fU.numpy(), fu, rtol=1e-12, atol=1e-12)
        assert np.allclose(fV.numpy(), fv, rtol=1e-12, atol=1e-12)
        assert torch.equal(s2, sums)

        # ---- gradient, sharded by USER RANGE: own rows of U and dU stay local, only dV travels -
        u_lo, u_hi = P.shard_bounds(n, world, rank)
        mine = (R[:, 0] >= u_lo) & (R[:, 0] < u_hi)
        R_loc = R[mine].copy()
        R_loc[:, 0] -= u_lo                                  # local user ids
        U_loc = U[u_lo:u_hi]
        prm = P.user_range_params(N.PmfParams(.7, 5., 9., 0.), rank)
        assert prm.sigma_u_sq == 5. and (prm.sigma_v_sq == 9. if rank == 0 else np.isinf(prm.sigma_v_sq))
        # every local row must exist for the oracle's shape inference: pass the tables explicitly
        gu, gv = O.gradient(R_loc, U_loc, V, sigma_sq=prm.sigma_sq, sigma_u_sq=prm.sigma_u_sq,
                            sigma_v_sq=prm.sigma_v_sq)
        resid = R_loc[:, 2] - O.predictions(R_loc, U_loc, V)
        s3 = torch.tensor([float(resid @ resid), float((U_loc * U_loc).sum()), float((V * V).sum())],
                          dtype=torch.float64)
        dV = torch.from_numpy(gv.copy())
        P.combine_loss_grad_user_range(dV, s3, world, rank)
        assert np.allclose(gu, fu[u_lo:u_hi], rtol=1e-12, atol=1e-12)      # no collective for dU
        assert np.allclose(dV.numpy(), fv, rtol=1e-12, atol=1e-12)
        assert np.allclose(s3.numpy(), [full @ full, (U * U).sum(), (V * V).sum()], rtol=1e-12)

        # ---- scoring: candidate shards, winner all-gather --------------------------------------
        lo, hi = P.shard_bounds(ncand, world, rank)
        for maximize in (True, False):
            vals = np.einsum("nd,nd->n", U[ci[lo:hi]], V[cj[lo:hi]])
            loc = int(np.argmax(vals) if maximize else np.argmin(vals))
            best = torch.empty(2, dtype=torch.int64)
            best[0] = torch.tensor([vals[loc]], dtype=torch.float64).view(torch.int64)[0]
            best[1] = lo + loc
            rec = P.gather_winner(best, world)
            v, idx = P.winner_from_records(rec, maximize)
            allv = np.einsum("nd,nd->n", U[ci], V[cj])
            want = int(np.argmax(allv) if maximize else np.argmin(allv))
            assert idx == want and v == allv[want]
        # tie across shards -> the lower global index wins on every rank
        tie = torch.empty(2, dtype=torch.int64)
        tie[0] = torch.tensor([1.5], dtype=torch.float64).view(torch.int64)[0]
        tie[1] = 700 if rank == 0 else 3
        v, idx = P.winner_from_records(P.gather_winner(tie, world), True)
        assert (v, idx) == (1.5, 3)
        # an empty shard (index -1) never wins
        emp = torch.empty(2, dtype=torch.int64)
        emp[0] = torch.tensor([99.0 if rank == 0 else 1.0], dtype=torch.float64).view(torch.int64)[0]
        emp[1] = -1 if rank == 0 else 12
        assert P.winner_from_records(P.gather_winner(emp, world), True) == (1.0, 12)
        # ---- class-level helpers: sharded criterion evaluation over a pool --------------------
        class FakeModel:                     # stands in for ActivePMF: scores = i*100 + j
            unrated = {(1, 2), (3, 4)}
            def _get_key_vals(self, pool, key, procs, worker_pool):
                return [100.0 * i + j for i, j in pool]
            def pred_variance(self, ij):
                pass
        FakeModel.pred_variance.chooser = max
        pool = [(i, j) for i in range(7) for j in range(5)]
        vals = P.sharded_key_vals(FakeModel(), pool, FakeModel.pred_variance)
        assert vals == [100.0 * i + j for i, j in pool]
        assert P.sharded_pick_query_point(FakeModel(), pool, FakeModel.pred_variance) == (6, 4)
        rows = torch.zeros((11, 3), dtype=torch.float64)
        lo, hi = P.shard_bounds(11, world, rank)
        rows[lo:hi] = torch.arange(lo, hi, dtype=torch.float64)[:, None] + 1
        got = P.all_gather_rows(rows, 11, world)
        assert torch.equal(got[:, 0], torch.arange(1, 12, dtype=torch.float64))
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert list(out) == [1, 1]


def test_shard_bounds_and_reduce():
    assert [P.shard_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert P.shard_bounds(0, 2, 1) == (0, 0)
    v = torch.tensor([1., 3., 3., float('nan')], dtype=torch.float64)
    i = torch.tensor([5, 9, 7, 1])
    assert P.reduce_winners(v, i, True) == (3.0, 7)
    assert P.reduce_winners(v, i, False) == (1.0, 5)
    v, idx = P.reduce_winners(torch.tensor([float('nan')], dtype=torch.float64), torch.tensor([-1]))
    assert idx == -1 and np.isnan(v)


def test_peer_winner_exchange_is_off_for_one_rank_and_by_request(monkeypatch):
    """the NVLink peer-memory exchange needs several ranks; world = 1, more than 32 ranks or
    AMF_PEER_EXCHANGE=0 keep the plain path (no device, no process group touched)"""
    from active_matrix_factorization_b200 import parallel as P
    assert P.PeerWinnerExchange.create(1, 0) is None
    assert P.PeerWinnerExchange.create(64, 3) is None
    monkeypatch.setenv("AMF_PEER_EXCHANGE", "0")
    assert P.PeerWinnerExchange.create(2, 0) is None
