"""One process per GPU: sharding of the two data-parallel axes of the path.

* candidate scoring (active_pmf.py:739-770): the pool is split into contiguous shards, one per
  rank; each rank scores its shard with the fused arg-best kernel, then a 16-byte all-gather of
  (value, global index) and the same deterministic reduction on every rank (best value, lowest
  global index on ties).  No score ever crosses NVLink unless the full matrix is requested.
* PMF loss+gradient (pmf_cy.pyx:170-223): ratings are split by blocks; every rank evaluates
  the data term on its block, rank 0 alone adds the prior term and the norms, and dU, dV and
  the three objective sums are all-reduced (NCCL over NVLink; gloo in the CPU tests).

The collectives go through torch.distributed so the same code runs under NCCL and gloo.
"""
import math
import os

import numpy as np
import torch

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


def world_rank():
    """(world_size, rank) of the default process group, (1, 0) when not initialised"""
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def all_gather_rows(local_full, rows, world, group=None):
    """`local_full` is a (rows, d) tensor in which this rank filled rows shard_bounds(rows, world,
    rank); returns the tensor with every rank's rows (all-gather of padded equal-size shards)."""
    rank = dist.get_rank(group)
    per = (rows + world - 1) // world + 1
    lo, hi = shard_bounds(rows, world, rank)
    send = torch.zeros((per,) + tuple(local_full.shape[1:]), dtype=local_full.dtype,
                       device=local_full.device)
    send[:hi - lo] = local_full[lo:hi]
    recv = torch.empty((world * per,) + tuple(local_full.shape[1:]), dtype=local_full.dtype,
                       device=local_full.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    out = torch.empty_like(local_full)
    for r in range(world):
        a, b = shard_bounds(rows, world, r)
        out[a:b] = recv[r * per:r * per + (b - a)]
    return out


def sharded_key_vals(model, pool, key, device=None, group=None):
    """Multi-GPU `_get_key_vals` (active_pmf.py:739-770): the pool is cut into contiguous shards,
    every rank evaluates its shard with the batched GPU path, the scores are all-gathered and
    every rank returns the full list aligned with `pool`.  Works for every criterion, including
    the lookahead ones (each (candidate, value) problem is independent)."""
    pool = list(pool)
    world, rank = world_rank()
    if world == 1:
        return model._get_key_vals(pool, key, None, None)
    lo, hi = shard_bounds(len(pool), world, rank)
    local = model._get_key_vals(pool[lo:hi], key, None, None)
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) \
            if dist.get_backend(group) == 'nccl' else torch.device('cpu')
    full = torch.zeros((len(pool), 1), dtype=torch.float64, device=device)
    if hi > lo:
        full[lo:hi, 0] = torch.tensor(local, dtype=torch.float64, device=device)
    return all_gather_rows(full, len(pool), world, group)[:, 0].cpu().tolist()


def sharded_pick_query_point(model, pool=None, key=None, group=None):
    """Multi-GPU `pick_query_point` (active_pmf.py:709-737): same answer on every rank."""
    import operator
    if pool is None:
        pool = model.unrated
    pool = list(pool)
    if key is None:
        key = type(model).pred_variance
    if len(pool) == 0:
        raise ValueError("can't pick a query point from an empty pool")
    if len(pool) == 1:
        return pool[0]
    vals = sharded_key_vals(model, pool, key, group=group)
    return getattr(key, 'chooser', max)(zip(pool, vals), key=operator.itemgetter(1))[0]


def shard_bounds(total, world, rank):
    """Contiguous, balanced [lo, hi) of `total` items for `rank` of `world`."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_winners(values, indices, maximize=True):
    """values (W,) float64, indices (W,) int64 (-1 = shard had no valid candidate).
    Returns (value, index) of the best; ties go to the lowest index; (nan, -1) if none."""
    valid = indices >= 0
    valid &= ~torch.isnan(values)
    if not bool(valid.any()):
        return float('nan'), -1
    fill = -math.inf if maximize else math.inf
    v = torch.where(valid, values, torch.full_like(values, fill))
    top = v.max() if maximize else v.min()
    tie = valid & (v == top)
    big = torch.iinfo(torch.int64).max
    idx = torch.where(tie, indices, torch.full_like(indices, big)).min()
    return float(top), int(idx)


def gather_winner(best, world, group=None):
    """best: tensor of 2 int64 words holding {float64 value, int64 index} (amf_best_t).
    All-gathers the records and returns the reduced (value, index) -- identical on every rank."""
    if world == 1:
        rec = best.view(1, 2)
    else:
        rec = torch.empty((world, 2), dtype=torch.int64, device=best.device)
        dist.all_gather_into_tensor(rec, best.view(1, 2).contiguous(), group=group)
    return rec


class PeerWinnerExchange:
    """The winner all-gather + reduction as ONE kernel per rank over NVLink peer memory
    (csrc/peer.cu): every rank stores its 16-byte record into every peer's mailbox, raises a flag
    there, waits for all flags in its own and reduces.  One process per GPU on one node; mailboxes
    are opened in the peers through CUDA IPC.  `create` returns None when that is not possible
    (the caller then keeps the NCCL all-gather)."""

    def __init__(self, handle, world):
        self._h, self.world = handle, world

    @classmethod
    def create(cls, world, rank, group=None):
        import ctypes as C
        from . import _native as N
        if world <= 1 or world > 32 or os.environ.get("AMF_PEER_EXCHANGE", "1") == "0":
            return None
        lib = N.require_device()
        h = C.c_void_p()
        mine = (C.c_ubyte * 64)()
        ok = lib.amf_peer_create(C.byref(h), world, rank, mine) == 0
        # every rank takes part in the handle exchange whether or not its own setup worked
        send = torch.tensor(list(bytes(mine)) + [1 if ok else 0], dtype=torch.uint8, device='cuda')
        allh = torch.empty((world, 65), dtype=torch.uint8, device='cuda')
        dist.all_gather_into_tensor(allh, send.view(1, 65), group=group)
        allh = allh.cpu().numpy()
        good = ok and bool(allh[:, 64].all())
        if good:
            buf = np.ascontiguousarray(allh[:, :64]).tobytes()
            good = lib.amf_peer_connect(h, buf) == 0
        flag = torch.tensor([1 if good else 0], dtype=torch.int32, device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # all ranks or none
        if int(flag.item()) != 1:
            if ok:
                lib.amf_peer_destroy(h)
            return None
        return cls(h, world)

    def reduce(self, best, maximize=True):
        """best: this rank's amf_best_t (2 int64 words) -> the winner over all ranks, in place"""
        from . import _native as N
        from . import device as D
        N.check(N.load().amf_peer_best_reduce(self._h, D.ptr(best), 1 if maximize else 0, D.ptr(best),
                                              D.stream_ptr()))

    def close(self):
        if self._h:
            from . import _native as N
            torch.cuda.synchronize()
            N.load().amf_peer_destroy(self._h)
            self._h = None


def winner_from_records(rec, maximize=True):
    vals = rec[:, 0].contiguous().view(torch.float64)
    return reduce_winners(vals, rec[:, 1].contiguous(), maximize)


def prior_once_params(params, rank):
    """Only rank 0 contributes the prior term -U/sigma_u^2 (and V's); the others pass an
    infinite prior variance so that the all-reduced gradient counts it once."""
    if rank == 0:
        return params
    return type(params)(params.sigma_sq, math.inf, math.inf, params.mean_offset)


def combine_loss_grad(dU, dV, sums, world, rank, group=None, grads_flat=None):
    """All-reduce of the per-shard data terms; |U|^2, |V|^2 are counted once (rank 0).
    grads_flat: one tensor whose storage holds dU followed by dV (see alloc_grads): the two
    gradients then travel in one collective instead of two."""
    if world == 1:
        return
    if rank != 0:
        sums[1:].zero_()
    if dU is not None:
        if grads_flat is not None:
            dist.all_reduce(grads_flat, group=group)
        else:
            dist.all_reduce(dU, group=group)
            dist.all_reduce(dV, group=group)
    dist.all_reduce(sums, group=group)


def user_range_params(params, rank):
    """Ratings sharded by USER RANGE: a rank owns its user rows outright, so
    their prior term is its own; only V's prior has to be counted once."""
    if rank == 0:
        return params
    return type(params)(params.sigma_sq, params.sigma_u_sq, math.inf, params.mean_offset)


def combine_loss_grad_user_range(dV, sums, world, rank, group=None):
    """User-range sharding: dU rows are rank-private and complete after the local pass; only dV
    (M x d instead of (N + M) x d) and the three sums travel.  The squared error and |U|^2
    (disjoint rows) add up over the ranks, |V|^2 is counted once (rank 0)."""
    if world == 1:
        return
    if rank != 0:
        sums[2:].zero_()
    if dV is not None:
        dist.all_reduce(dV, group=group)
    dist.all_reduce(sums, group=group)


def alloc_grads(U, V):
    """(dU, dV, flat): gradient buffers shaped like U and V that share one allocation, so that a
    sharded step can all-reduce both with one collective."""
    flat = torch.empty(U.numel() + V.numel(), dtype=U.dtype, device=U.device)
    return flat[:U.numel()].view_as(U), flat[U.numel():].view_as(V), flat


class ShardedStep:
    """The benchmarked step: fused loss+gradient over this rank's rating block, then scoring
    of this rank's candidate shard with a fused arg-best, each followed by its collective."""

    def __init__(self, rat, d, name, world=1, rank=0, grad_shard='ratings'):
        self.rat, self.d, self.name, self.world, self.rank = rat, d, name, world, rank
        # 'ratings': every rank holds a block of the rating list over ALL users, dU and dV are
        # all-reduced (north_star's scheme).  'users': a rank holds the ratings of its own user
        # range and its own rows of U; dU needs no collective, only dV is all-reduced.
        if grad_shard not in ('ratings', 'users'):
            raise ValueError("grad_shard must be 'ratings' or 'users'")
        self.grad_shard = grad_shard
        self.index_base = 0
        self._rec = None
        self.pool = None          # optional scoring.Pool (tiled layout) for the pred criterion
        # winners of the shards: one kernel over NVLink peer memory when the ranks can map each
        # other's mailboxes (PeerWinnerExchange), else a 16-byte NCCL all-gather + a reduction launch
        self.peer = None
        self._peer_tried = False
        # kernels launched by one step: 2 prior + 2 side passes; scoring + winner reduction
        # (+ the reduction of the gathered winners on multi-GPU runs)
        self.launches_per_step = 6 + (1 if world > 1 else 0)   # (5 with the fused winner exchange)

    def set_candidate_offset(self, ncand_local):
        """global index = offset of this rank's shard + local index"""
        if self.world == 1:
            self.index_base = 0
            return
        cnt = torch.tensor([ncand_local], dtype=torch.int64, device='cuda')
        allc = torch.empty(self.world, dtype=torch.int64, device='cuda')
        dist.all_gather_into_tensor(allc, cnt)
        self.index_base = int(allc[:self.rank].sum().item())

    # SMs left to NCCL while the second gradient pass runs (its CTAs otherwise fill every SM)
    SMS_FOR_COLLECTIVE = 8

    def loss_grad(self, U, V, params, dU, dV, sums, grads_flat=None, overlap=False):
        """Fused loss+gradient over this rank's rating block, all-reduced.  overlap=True runs
        the pass that completes dU first, starts its all-reduce (and that of the sums)
        asynchronously and lets it travel while the second pass computes dV on all but a few
        SMs; measured on 2 and 8 B200s it does not pay (0.764 vs 0.780 ms and 1.412 vs 1.386 ms,
        benchmarks/check_sharded_grad.py), so one all-reduce after both passes is the default."""
        from . import device as D
        if self.grad_shard == 'users' and self.world > 1:
            D.loss_grad(self.rat, self.d, U, V, user_range_params(params, self.rank), dU, dV, sums)
            combine_loss_grad_user_range(dV, sums, self.world, self.rank)
            return
        prm = prior_once_params(params, self.rank)
        if self.world == 1:
            D.loss_grad(self.rat, self.d, U, V, prm, dU, dV, sums)
            return
        if not overlap:
            D.loss_grad(self.rat, self.d, U, V, prm, dU, dV, sums)
            combine_loss_grad(dU, dV, sums, self.world, self.rank, grads_flat=grads_flat)
            return
        n_sms = torch.cuda.get_device_properties(U.device).multi_processor_count
        D.loss_grad_part(self.rat, self.d, U, V, prm, dU, dV, sums, 0)
        if self.rank != 0:
            sums[1:].zero_()
        w_u = dist.all_reduce(dU, async_op=True)
        w_s = dist.all_reduce(sums, async_op=True)
        D.loss_grad_part(self.rat, self.d, U, V, prm, dU, dV, sums, 1,
                         max_ctas=max(1, n_sms - self.SMS_FOR_COLLECTIVE))
        dist.all_reduce(dV)
        w_u.wait()
        w_s.wait()

    def select(self, criterion, ci, cj, U, V, view, cutoff, maximize, best):
        from . import device as D
        from . import _native as N
        import ctypes as C
        if self._rec is None:
            self.set_candidate_offset(int(ci.numel()))
            self._rec = True
        lib = N.require_device()
        if self.world > 1 and not self._peer_tried:
            self._peer_tried = True
            self.peer = PeerWinnerExchange.create(self.world, self.rank)
        if self.pool is not None and criterion == N.CRIT_PRED:
            # with a peer exchange the scoring kernel itself ends with the cross-GPU winner
            self.pool.score_pred(U, V, False, maximize, self.index_base, best, peer=self.peer)
            # 2 prior + 2 side passes + the scoring kernel (which ends with the winner reduction and,
            # with a peer exchange, the cross-GPU winner); + 1 for the NCCL fallback's reduction
            self.launches_per_step = 5 if (self.world == 1 or self.peer is not None) else 6
            if self.peer is not None:
                return
        else:
            N.check(lib.amf_score_candidates(
                criterion, D.code(self.name), int(ci.numel()), D.ptr(ci), D.ptr(cj), self.d,
                U.shape[1] if U is not None else 0, D.ptr(U), D.ptr(V),
                C.byref(view) if view is not None else None, float(cutoff), None,
                1 if maximize else 0, self.index_base, D.ptr(best), D.stream_ptr()))
        if self.world > 1:
            if self.peer is not None:
                # one launch: records exchanged with remote stores over NVLink, same tie-break
                self.peer.reduce(best, maximize)
                return
            # 16-byte all-gather of the per-rank winners, then one launch applies the same
            # tie-break on every rank; no host sync inside the step
            rec = gather_winner(best, self.world)
            N.check(lib.amf_best_reduce(D.ptr(rec), self.world, 1 if maximize else 0, D.ptr(best),
                                        D.stream_ptr()))

    def kernel_times(self, U, V, params, dU, dV, sums, ci, cj, best, reps=5):
        """Average device time of the two dominant kernels, timed alone on the current stream
        with CUDA events (inputs are far larger than L2, so every launch streams from HBM)."""
        from . import device as D
        from . import _native as N
        lib = N.require_device()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.loss_grad(self.rat, self.d, U, V, params, dU, dV, sums)   # builds the tiled copy on first use
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            D.loss_grad(self.rat, self.d, U, V, params, dU, dV, sums)
        e1.record()
        torch.cuda.synchronize()
        side = e0.elapsed_time(e1) / reps
        e0.record()
        for _ in range(reps):
            N.check(lib.amf_score_candidates(N.CRIT_PRED, D.code(self.name), int(ci.numel()),
                                             D.ptr(ci), D.ptr(cj), self.d, U.shape[1], D.ptr(U),
                                             D.ptr(V), None, 0.0, None, 1, 0, D.ptr(best),
                                             D.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        out = {"side_pass_ms": side, "score_flat_ms": e0.elapsed_time(e1) / reps}
        out["score_ms"] = out["score_flat_ms"]
        if self.pool is not None:
            e0.record()
            for _ in range(reps):
                self.pool.score_pred(U, V, False, True, 0, best)
            e1.record()
            torch.cuda.synchronize()
            out["score_tiled_ms"] = e0.elapsed_time(e1) / reps
            out["score_ms"] = out["score_tiled_ms"]
        return out
