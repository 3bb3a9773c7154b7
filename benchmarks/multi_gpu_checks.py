"""torchrun checks of the multi-GPU paths that bench.py's step does not exercise (SURVEY.md 8e
rows 3 and 4), each with a parity assert against the single-GPU result computed in the same run:

  * Gibbs: rows of every half-sweep split over the ranks (BayesianPMF.shard_group), rank 0's
    draws broadcast, new rows all-gathered -- the chain must equal the unsharded chain;
  * lookahead / variance criteria: `parallel.sharded_key_vals` and `sharded_pick_query_point` over
    the drugbank configuration (scalable mode) -- scores and pick equal to one rank's.

    torchrun --nproc-per-node N benchmarks/multi_gpu_checks.py      -> one JSON line on rank 0
"""
import json
import os
import sys
import time
from itertools import islice

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def golden(name):
    with np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from active_matrix_factorization_b200 import build
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    from active_matrix_factorization_b200 import active_pmf as A, bayes_pmf as Bm, parallel as P
    out = {"n_gpus": world}

    # ---- Gibbs, movielens-100k split, rank 15 ------------------------------------------------
    g = golden("c4_movielens_bayes")
    R = g["ratings"].astype(float)

    def chain(shard, n_samples, seed):
        np.random.seed(0)
        b = Bm.BayesianPMF(R, 15, subtract_mean=True, knowable=())
        b.shard_group = True if shard else None
        np.random.seed(seed + (rank if shard else 0))      # the ranks' host streams DIFFER on purpose
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s = list(islice(b.samples(num_gibbs=2), n_samples))
        torch.cuda.synchronize()
        return s, time.perf_counter() - t0

    chain(world > 1, 2, 7)
    sharded, t_sh = chain(world > 1, 20, 7)
    single, t_1 = chain(False, 20, 7) if rank == 0 else (None, None)
    if rank == 0:
        err = max(np.abs(a[0] - b[0]).max() for a, b in zip(sharded, single))
        out["gibbs"] = {"samples": 20, "seconds_sharded": t_sh, "seconds_one_gpu": t_1,
                        "max_abs_diff_vs_one_gpu_chain": float(err)}
        assert err < 1e-9, "sharded chain left the single-GPU chain"
        # (that the unsharded chain is the reference's own is tests/test_gpu_configs.py::test_c4)
    if world > 1:
        dist.barrier()

    # ---- criteria over the drugbank pool, sharded -------------------------------------------------
    g = golden("c2_drugbank")
    R = g["ratings"].astype(float)
    a = A.ActivePMF(R, 5, rating_values={-1, 1}, discrete_expectations=True)
    a.users, a.items = g["users"].copy(), g["items"].copy()
    a.blocks_tol, a.blocks_max_sweeps = 1e-12, 2000
    a.initialize_approx()
    a.fit_normal()
    known = np.zeros((94, 425), bool)
    known[R[:, 0].astype(int), R[:, 1].astype(int)] = True
    ii, jj = np.nonzero(~known)
    pool = list(zip(ii.tolist(), jj.tolist()))
    res = {}
    for key, want in ((A.ActivePMF.exp_approx_entropy, g["b_uv_entropy"]),):
        P.sharded_key_vals(a, pool, key)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vals = np.array(P.sharded_key_vals(a, pool, key))
        pick = P.sharded_pick_query_point(a, pool, key)
        dt = time.perf_counter() - t0
        np.testing.assert_allclose(vals, want, rtol=1e-9)
        assert pick == pool[int(np.argmin(want))]
        res[key.__name__] = {"seconds_scores_plus_pick": dt, "max_rel_vs_fixture": float(np.abs(vals / want - 1).max()),
                             "same_pick_on_every_rank": True}
    pv = np.array(P.sharded_key_vals(a, pool, A.ActivePMF.pred_variance))
    np.testing.assert_allclose(pv[g["sub"]], g["b_pred_var_sub"], rtol=1e-9)
    res["pred_variance"] = {"max_rel_vs_fixture": float(np.abs(pv[g["sub"]] / g["b_pred_var_sub"] - 1).max())}
    out["sharded_criteria_c2"] = res

    # ---- winner exchange over NVLink peer memory against the NCCL all-gather + amf_best_reduce ----
    if world > 1:
        from active_matrix_factorization_b200 import _native as N, device as D
        lib = N.require_device()
        peer = P.PeerWinnerExchange.create(world, rank)
        ex = {"available": peer is not None}
        if peer is not None:
            rng = np.random.RandomState(1234)              # the same table of records on every rank
            trials = 200
            vals = rng.normal(size=(trials, world))
            vals[::7] = np.round(vals[::7])                  # ties: the lowest index must win
            idxs = rng.randint(0, 10**9, size=(trials, world)).astype(np.int64)
            idxs[::11, 0] = -1                               # a rank without a candidate
            vals[5, :] = np.nan                              # NaN never wins
            t_peer = t_nccl = 0.0
            for mx in (1, 0):
                for t in range(trials):
                    mine = torch.tensor([np.float64(vals[t, rank]).view(np.int64), idxs[t, rank]],
                                        dtype=torch.int64, device="cuda")
                    ref = mine.clone()
                    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                    e0.record()
                    peer.reduce(mine, bool(mx))
                    e1.record()
                    rec = P.gather_winner(ref, world)
                    N.check(lib.amf_best_reduce(D.ptr(rec), world, mx, D.ptr(ref), D.stream_ptr()))
                    e2.record()
                    torch.cuda.synchronize()
                    t_peer += e0.elapsed_time(e1)
                    t_nccl += e1.elapsed_time(e2)
                    a_, b_ = mine.cpu().numpy(), ref.cpu().numpy()
                    assert a_[1] == b_[1] and (a_[0] == b_[0] or a_[1] < 0), (t, mx, a_, b_)
            # the exchange fused into the pool scoring kernel (amf_pool_score_pred_peer): every rank
            # scores its own random shard; the winner over all shards must equal the NCCL path's
            from active_matrix_factorization_b200 import scoring as S
            rs = np.random.RandomState(100 + rank)
            nu, ni, dd, nc = 500, 3000, 32, 200_000
            Uh = np.random.RandomState(7).normal(size=(nu, dd))
            Vh = np.random.RandomState(8).normal(size=(ni, dd))
            ii, jj = rs.randint(0, nu, nc), rs.randint(0, ni, nc)
            for name in ("f32", "f64"):
                pool = S.Pool(ii, jj, nu, ni, name, dd, tile_bytes=64 * 1024)
                Ut, Vt = pool.pad(Uh), pool.pad(Vh)
                for mx in (True, False):
                    for rep in range(3):
                        _, fused = pool.score_pred(Ut, Vt, maximize=mx, index_base=rank * nc, peer=peer)
                        _, local = pool.score_pred(Ut, Vt, maximize=mx, index_base=rank * nc)
                        rec = P.gather_winner(local, world)
                        ref = torch.empty(2, dtype=torch.int64, device="cuda")
                        N.check(lib.amf_best_reduce(D.ptr(rec), world, 1 if mx else 0, D.ptr(ref), D.stream_ptr()))
                        torch.cuda.synchronize()
                        assert (fused.cpu().numpy() == ref.cpu().numpy()).all(), (name, mx, fused, ref)
                pool.close()
            ex["fused_into_pool_kernel_equal_to_nccl_path"] = True
            ex.update({"trials": 2 * trials, "equal_to_nccl_path": True,
                       "peer_kernel_us": 1e3 * t_peer / (2 * trials), "nccl_path_us": 1e3 * t_nccl / (2 * trials)})
            peer.close()
        out["peer_winner_exchange"] = ex
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
