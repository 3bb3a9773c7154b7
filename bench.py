#!/usr/bin/env python3
"""Benchmark of the python-pmf hot path on B200 (BASELINE.json metric):
candidate entries scored/sec and PMF ratings/sec/iter, vs the reference's CPU path.

One "step" = one active-learning inner step on the C5 workload: one fused PMF
loss+gradient evaluation over the observed-rating list (pmf_cy.pyx:170-223) followed by one
scoring pass with fused arg-max over the candidate pool (active_pmf.py:709-770, criterion
`pred`).  `value` is candidates scored per second over the scoring phase; the gradient phase
is reported beside it as ratings/sec/iter (`phases`).  `ms_per_step` covers both phases.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "candidates_scored_per_sec"
UNIT = "candidates/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--users", type=int, default=200_000)
    ap.add_argument("--items", type=int, default=50_000)
    ap.add_argument("--latent-d", type=int, default=32)
    ap.add_argument("--nnz", type=int, default=50_000_000, help="observed ratings per GPU")
    ap.add_argument("--ncand", type=int, default=100_000_000, help="candidates per GPU")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--pool", default="tiled", choices=["tiled", "flat"],
                    help="candidate pool layout for the device-resident scoring phase")
    ap.add_argument("--layout", default="auto", choices=["auto", "rows", "tiled"],
                    help="rating-list copy the fused loss+gradient runs on (amf_ratings_set_layout)")
    ap.add_argument("--grad-shard", default="ratings", choices=["ratings", "users"],
                    help="multi-GPU gradient: 'ratings' = rating blocks over all users, all-reduce of "
                         "dU and dV (north_star's scheme, default); 'users' = each GPU owns the ratings "
                         "and the U rows of its own user range (the matrix has --users x N rows), only "
                         "dV is all-reduced")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the C1-C4 / S2 / S3 / B1 entries (`configs`) and the fp64 rooflines")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    return ap.parse_args()


def workload_name(a):
    return "C5 synthetic %dx%d, rank %d, %d ratings + %d candidates per GPU" % (
        a.users, a.items, a.latent_d, a.nnz, a.ncand)


def measured_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant
    kernels, from the committed `ncu --set full` capture of this same command
    (profiles/r03_main_ncu_full_raw.csv, recipe benchmarks/profile_step.sh); {} if the summary file is
    missing."""
    try:
        with open(os.path.join(ROOT, "profiles", "r03_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, count=1):
        self.index = index if count == 1 else ",".join(str(i) for i in range(count))
        self.count = count
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        out = {"sm_mhz": float(np.median(sm)) if sm else None,
               "sm_max_mhz": float(max(mx)) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if self.count > 1 and sm:
            # nvidia-smi prints the GPUs round-robin: per-GPU medians expose a slow device
            out["sm_mhz_per_gpu"] = [float(np.median(sm[g::self.count])) for g in range(self.count)]
            out["sm_mhz"] = float(min(out["sm_mhz_per_gpu"]))
        return out


# ------------------------------------------------------------------------------------------
# synthetic C5 data, generated on the device
# ------------------------------------------------------------------------------------------
def make_problem(a, rank, torch):
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    n, m, d = a.users, a.items, a.latent_d
    tdt = torch.float32 if a.dtype == "f32" else torch.float64
    # observed cells: uniform, every row and column hit at least once, duplicates removed
    i = torch.randint(0, n, (a.nnz,), generator=g, device=dev, dtype=torch.int64)
    j = torch.randint(0, m, (a.nnz,), generator=g, device=dev, dtype=torch.int64)
    i = torch.cat([i, torch.arange(n, device=dev), torch.randint(0, n, (m,), generator=g, device=dev)])
    j = torch.cat([j, torch.randint(0, m, (n,), generator=g, device=dev), torch.arange(m, device=dev)])
    keys = torch.unique(i * m + j)
    del i, j
    perm = torch.randperm(keys.numel(), generator=g, device=dev)
    keys_shuf = keys[perm]                 # the rating LIST is unordered, like the reference's
    del perm
    ri = (keys_shuf // m).to(torch.int32)
    rj = (keys_shuf % m).to(torch.int32)
    del keys_shuf
    g0 = torch.Generator(device=dev)
    g0.manual_seed(99)                     # the model is the same on every rank
    scale = 1.0 / np.sqrt(np.sqrt(d))
    Ut = torch.randn((n, d), generator=g0, device=dev, dtype=tdt) * scale
    Vt = torch.randn((m, d), generator=g0, device=dev, dtype=tdt) * scale
    r = (Ut[ri.long()] * Vt[rj.long()]).sum(1) + 0.25 * torch.randn(ri.numel(), generator=g, device=dev, dtype=tdt)
    U = (Ut + 0.1 * torch.randn((n, d), generator=g0, device=dev, dtype=tdt)).contiguous()
    V = (Vt + 0.1 * torch.randn((m, d), generator=g0, device=dev, dtype=tdt)).contiguous()
    del Ut, Vt
    # candidates: distinct unobserved cells, sorted by (i, j)
    ck = torch.randint(0, n * m, (int(a.ncand * 1.02) + 1024,), generator=g, device=dev, dtype=torch.int64)
    ck = torch.unique(ck)
    pos = torch.searchsorted(keys, ck).clamp_(max=keys.numel() - 1)
    ck = ck[keys[pos] != ck]
    del pos, keys
    if ck.numel() > a.ncand:
        # drop a random subset but keep the order
        keep = torch.randperm(ck.numel(), generator=g, device=dev)[:a.ncand].sort().values
        ck = ck[keep]
        del keep
    ci = (ck // m).to(torch.int32)
    cj = (ck % m).to(torch.int32)
    del ck
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return dict(ri=ri, rj=rj, r=r.contiguous(), U=U, V=V, ci=ci, cj=cj)


# ------------------------------------------------------------------------------------------
# CPU reference legs (the only places that execute oracle/)
# ------------------------------------------------------------------------------------------
def cpu_reference_sample(a, prob_host, sample, fanout=False):
    """Times the reference's own CPU path (oracle/_ref when built, else the numpy port) on a
    bounded sample of the same workload.  Returns the cpu_baseline dict."""
    from oracle import build_ref
    n, m, d = a.users, a.items, a.latent_d
    ri, rj, r, U, V, ci, cj = prob_host
    s_r = min(sample, len(ri))
    s_c = min(sample, len(ci))
    R = np.column_stack((ri[:s_r], rj[:s_r], r[:s_r])).astype(float)
    R[0, 0], R[1, 1] = n - 1, m - 1        # pin the model shape (reference infers it from max id)
    pool = list(zip(ci[:s_c].tolist(), cj[:s_c].tolist()))
    out = {"cores": 1, "unit": UNIT}
    ref = None
    if build_ref.built():
        try:
            from oracle import ref_loader
            ref = ref_loader.load()
        except Exception as exc:          # e.g. a box whose Python cannot load the built modules
            out["reference_unavailable"] = repr(exc)
    if ref is not None:
        apmf = ref.active_pmf.ActivePMF(R, d, knowable=())
        apmf.users, apmf.items = U.astype(float), V.astype(float)
        t0 = time.perf_counter()
        apmf.gradient()
        apmf.log_likelihood()
        t_grad = time.perf_counter() - t0
        t0 = time.perf_counter()
        vals = apmf._get_key_vals(pool, ref.active_pmf.ActivePMF.pred, 1, None)
        best = max(zip(pool, vals), key=lambda t: t[1])[0]
        t_score = time.perf_counter() - t0
        out["kind"] = "reference"
        how = "oracle/_ref (reference Cython) ActivePMF.gradient()+log_likelihood() and _get_key_vals(pool, pred, procs=1)"
        # the reference's own all-core path: multiprocessing.Pool fan-out of the same call
        # (active_pmf.py:765-770, the model is pickled to the workers); the faster of the two
        # is reported.  The gradient has no multi-core path in the reference.
        if fanout:
            try:
                t0 = time.perf_counter()
                vals_mp = apmf._get_key_vals(pool, ref.active_pmf.ActivePMF.pred, None, None)
                max(zip(pool, vals_mp), key=lambda t: t[1])
                t_mp = time.perf_counter() - t0
                out["fanout"] = {"procs": os.cpu_count(), "seconds": t_mp, "single_process_seconds": t_score}
                if t_mp < t_score:
                    t_score = t_mp
                    out["cores"] = os.cpu_count()
                    how += "; scoring through the reference's multiprocessing.Pool fan-out (procs=None)"
            except Exception as exc:
                out["fanout"] = {"failed": repr(exc)[:200]}
        else:
            # our arm does not fork() from a process that holds a CUDA context: the all-core
            # fan-out is timed by `bench.py --impl reference` (same box, same run of the driver)
            out["fanout"] = None
            how += "; single process (the Pool fan-out is timed by --impl reference)"
    else:
        from oracle import pmf_oracle as O
        t0 = time.perf_counter()
        O.gradient(R, U.astype(float), V.astype(float))
        O.log_likelihood(R, U.astype(float), V.astype(float))
        t_grad = time.perf_counter() - t0
        t0 = time.perf_counter()
        vals = np.einsum("nd,nd->n", U[ci[:s_c]].astype(float), V[cj[:s_c]].astype(float))
        best = pool[int(np.argmax(vals))]
        t_score = time.perf_counter() - t0
        out["kind"] = "port"
        how = "oracle/pmf_oracle.py (numpy port) gradient+log_likelihood and vectorised pred"
    out["value"] = s_c / t_score
    out["pmf_ratings_per_sec_iter"] = s_r / t_grad
    out["sample"] = "%d of the ratings and %d of the candidates of rank 0's shard, fp64, %s; " \
                    "per-item cost is size independent (SURVEY.md 6)" % (s_r, s_c, how)
    out["seconds"] = {"grad_plus_ll": t_grad, "score": t_score}
    out["cpu_model"] = cpu_model()
    out["host_cores"] = os.cpu_count()
    out["_best"] = best
    return out


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(a):
    """--impl reference: the reference's CPU implementation on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rng = np.random.RandomState(1234)
    n, m, d = a.users, a.items, a.latent_d
    # bounded sample: keep the whole run (warm-up + steps) within ~3 minutes of host time
    # (the reference spends ~10 us per rating on gradient+LL+construction)
    budget_s = 180.0 / max(1, a.steps + a.warmup)
    s = int(min(a.cpu_sample, max(20_000, budget_s * 1e5)))
    ri, rj = rng.randint(0, n, s), rng.randint(0, m, s)
    U = (rng.standard_normal((n, d)) / np.sqrt(np.sqrt(d))).astype(np.float32)
    V = (rng.standard_normal((m, d)) / np.sqrt(np.sqrt(d))).astype(np.float32)
    r = (U[ri] * V[rj]).sum(1) + .25 * rng.standard_normal(s)
    ck = np.unique(rng.randint(0, n * m, s))
    ci, cj = ck // m, ck % m
    host = (ri, rj, r, U, V, ci, cj)
    times = []
    res = None
    for it in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        res = cpu_reference_sample(a, host, s, fanout=True)
        if it >= a.warmup:
            times.append((time.perf_counter() - t0, res["seconds"]["score"], res["seconds"]["grad_plus_ll"]))
    t_score = float(np.mean([t[1] for t in times]))
    t_grad = float(np.mean([t[2] for t in times]))
    value = len(ci) / t_score
    res.pop("_best", None)
    res["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * float(np.mean([t[0] for t in times])),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(a),
                                        "note": "each step is a bounded sample (%d ratings, %d candidates) "
                                                "of the workload on the host CPU" % (s, len(ci))},
        "phases": {"pmf_loss_grad": {"ratings_per_sec_iter": s / t_grad},
                   "score_pred": {"candidates_per_sec": value}},
        "cpu_baseline": res,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: native libraries that print there (NCCL's version
    # banner) are sent to stderr until the line is written
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from active_matrix_factorization_b200 import build as B
    if rank == 0:
        B.build()
    if world > 1:
        dist.barrier()
    from active_matrix_factorization_b200 import _native as N
    from active_matrix_factorization_b200 import device as D
    from active_matrix_factorization_b200 import parallel as P
    lib = N.require_device()

    n, m, d = a.users, a.items, a.latent_d
    name = a.dtype
    es = 4 if name == "f32" else 8
    prob = make_problem(a, rank, torch)
    rat = D.Ratings(n, m, prob["ri"], prob["rj"], prob["r"], name)
    rat.set_layout(a.layout)
    nnz, ncand = rat.nnz, int(prob["ci"].numel())
    tiled_grad = a.layout == "tiled" or (a.layout == "auto" and nnz >= (1 << 20) and d * es in (64, 128, 256))
    U, V, ci, cj = prob["U"], prob["V"], prob["ci"], prob["cj"]
    ld = D.padded_ld(d, name)
    assert ld == d, "bench uses an unpadded rank"
    dU, dV, grads_flat = P.alloc_grads(U, V)       # one allocation: one all-reduce for both
    sums = torch.zeros(3, dtype=torch.float64, device=U.device)
    best = torch.zeros(2, dtype=torch.int64, device=U.device)
    params = D.pmf_params(1.0, 10.0, 10.0, 0.0)
    step = P.ShardedStep(rat, d, name, world, rank, grad_shard=a.grad_shard)
    pool_ms = None
    if a.pool == "tiled":
        from active_matrix_factorization_b200 import scoring as S
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step.pool = S.Pool(ci, cj, n, m, name, d)
        torch.cuda.synchronize()
        pool_ms = 1e3 * (time.perf_counter() - t0)

    def one_step(ev=None):
        if ev: ev[0].record()
        step.loss_grad(U, V, params, dU, dV, sums, grads_flat)
        if ev: ev[1].record()
        step.select(N.CRIT_PRED, ci, cj, U, V, None, 0.0, True, best)
        if ev: ev[2].record()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        one_step()
    sync()
    sampler = ClockSampler(local, world)
    if rank == 0:
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    t_begin.record()
    for k in range(a.steps):
        one_step(evs[k])
    t_end.record()
    sync()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_begin.elapsed_time(t_end)
    grad_ms = sum(e[0].elapsed_time(e[1]) for e in evs)
    score_ms = total_ms - grad_ms          # scoring + selection (+ its collective) up to step end
    # the step's selection, read BEFORE anything else writes `best`: on N GPUs this is the
    # all-gathered, reduced winner (global index = shard offset + local index)
    raw = best.cpu().numpy()
    bv, bi = float(raw[:1].view(np.float64)[0]), int(raw[1])
    # independent check of the sharded selection: every rank recomputes its local winner with the
    # flat (unbucketed) kernel, the records are all-gathered with a plain collective and reduced in
    # Python with the same rule (best value, lowest global index)
    chk = torch.zeros(2, dtype=torch.int64, device=U.device)
    N.check(lib.amf_score_candidates(N.CRIT_PRED, D.code(name), ncand, D.ptr(ci), D.ptr(cj), d, ld,
                                     D.ptr(U), D.ptr(V), None, 0.0, None, 1, step.index_base,
                                     D.ptr(chk), D.stream_ptr()))
    recs = P.gather_winner(chk, world)
    want_v, want_i = P.winner_from_records(recs, True)
    # same candidate; the value agrees to fp32 rounding (the two kernels sum the 32 products in
    # different orders)
    assert want_i == bi and abs(want_v - bv) <= 1e-5 * max(1.0, abs(bv)), \
        "sharded selection %r differs from the recomputed winner %r" % ((bv, bi), (want_v, want_i))
    selection_check = {"ranks": world, "recomputed_with": "score_pred_kernel per rank + all_gather + host reduction",
                       "equal": True, "local_winner_indices": recs[:, 1].cpu().tolist()}
    # isolated timing of the two dominant kernels for the roofline (same stream, CUDA events)
    kt = step.kernel_times(U, V, params, dU, dV, sums, ci, cj, best, reps=max(5, a.steps // 2))
    tm = torch.tensor([total_ms, grad_ms, score_ms, kt["side_pass_ms"], kt["score_ms"], kt["score_flat_ms"]],
                      dtype=torch.float64, device=U.device)
    per_rank = None
    if world > 1:
        # every rank's own numbers, so that a straggling GPU shows as such in the line
        allr = torch.empty((world, tm.numel()), dtype=torch.float64, device=U.device)
        dist.all_gather_into_tensor(allr, tm.view(1, -1).contiguous())
        per_rank = {"grad_phase_ms": (allr[:, 1] / a.steps).tolist(), "score_phase_ms": (allr[:, 2] / a.steps).tolist(),
                    "grad_kernel_ms": allr[:, 3].tolist(), "score_kernel_ms": allr[:, 4].tolist()}
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms, grad_ms, score_ms, side_ms, scorek_ms, score_flat_ms = tm.tolist()

    # ---- end to end through the host-buffer C ABI (PCIe copies inside the timed region) ------
    U_h = torch.empty((n, d), dtype=U.dtype).pin_memory(); U_h.copy_(U)
    V_h = torch.empty((m, d), dtype=V.dtype).pin_memory(); V_h.copy_(V)
    dU_h = torch.empty((n, d), dtype=U.dtype).pin_memory()
    dV_h = torch.empty((m, d), dtype=V.dtype).pin_memory()
    ci_h = torch.empty(ncand, dtype=torch.int32).pin_memory(); ci_h.copy_(ci)
    cj_h = torch.empty(ncand, dtype=torch.int32).pin_memory(); cj_h.copy_(cj)
    sums_h = np.zeros(3)
    best_h = N.Best()

    # the pool is sorted by user: the host call takes it as row offsets + item ids (CSR), which
    # halves the PCIe bytes of the (i, j) pair form, and with at most 65536 items the ids travel
    # as 16-bit words (amf_score_pred_host_csr16), which halves them again; all forms are timed
    ptr_h = torch.zeros(n + 1, dtype=torch.int64).pin_memory()
    ptr_h[1:].copy_(torch.cumsum(torch.bincount(ci.long(), minlength=n), 0))
    narrow = m <= 65536
    if narrow:
        cj16_h = torch.empty(ncand, dtype=torch.int16).pin_memory()   # same 16 bits as uint16
        cj16_h.copy_(cj.to(torch.int16))

    def e2e_step(csr=True, ids16=False):
        t0 = time.perf_counter()
        N.check(lib.amf_pmf_loss_grad_host(rat.handle, D.code(name), d, C.c_void_p(U_h.data_ptr()),
                                           C.c_void_p(V_h.data_ptr()), C.byref(params),
                                           C.c_void_p(dU_h.data_ptr()), C.c_void_p(dV_h.data_ptr()),
                                           N.host_ptr(sums_h)))
        t1 = time.perf_counter()
        if csr and ids16:
            N.check(lib.amf_score_pred_host_csr16(D.code(name), C.c_void_p(ptr_h.data_ptr()),
                                                  C.c_void_p(cj16_h.data_ptr()), n, m, d,
                                                  C.c_void_p(U_h.data_ptr()), C.c_void_p(V_h.data_ptr()),
                                                  None, 1, C.byref(best_h)))
        elif csr:
            N.check(lib.amf_score_pred_host_csr(D.code(name), C.c_void_p(ptr_h.data_ptr()),
                                                C.c_void_p(cj_h.data_ptr()), n, m, d,
                                                C.c_void_p(U_h.data_ptr()), C.c_void_p(V_h.data_ptr()),
                                                None, 1, C.byref(best_h)))
        else:
            N.check(lib.amf_score_pred_host(D.code(name), ncand, C.c_void_p(ci_h.data_ptr()),
                                            C.c_void_p(cj_h.data_ptr()), n, m, d,
                                            C.c_void_p(U_h.data_ptr()), C.c_void_p(V_h.data_ptr()),
                                            None, 1, C.byref(best_h)))
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    e2e_step(csr=False)
    sync()
    e2e_pairs_s = float(np.mean([e2e_step(csr=False)[1] for _ in range(a.e2e_steps)]))
    e2e_csr32_s = None
    if narrow:
        e2e_step()
        sync()
        e2e_csr32_s = float(np.mean([e2e_step()[1] for _ in range(a.e2e_steps)]))
    e2e_step(ids16=narrow)
    sync()
    e2e = [e2e_step(ids16=narrow) for _ in range(a.e2e_steps)]
    e2e_t = torch.tensor([float(np.mean([t[0] for t in e2e])), float(np.mean([t[1] for t in e2e]))],
                         dtype=torch.float64, device=U.device)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_grad_s, e2e_score_s = e2e_t.tolist()
    # the two paths must agree on the winner
    if world == 1:
        assert best_h.index == bi, "device-resident and end-to-end paths picked different candidates"

    cnt = torch.tensor([nnz, ncand], dtype=torch.int64, device=U.device)
    if world > 1:
        dist.all_reduce(cnt)
    nnz_all, ncand_all = cnt.tolist()

    # ---- strong scaling as config 5 words it: ONE pool of `ncand` candidates cut over the N GPUs
    # (rank r scores the r-th contiguous slice of rank 0's... of its own shard: same size, same
    # statistics), winner all-gathered; the weak-scaling `value` above keeps `ncand` per GPU
    strong = None
    if world > 1 and a.pool == "tiled":
        from active_matrix_factorization_b200 import scoring as S
        lo, hi = P.shard_bounds(ncand, world, rank)
        sl_i, sl_j = ci[lo:hi].contiguous(), cj[lo:hi].contiguous()
        sl_pool = S.Pool(sl_i, sl_j, n, m, name, d)
        sl_best = torch.zeros(2, dtype=torch.int64, device=U.device)

        def strong_step():
            sl_pool.score_pred(U, V, False, True, lo, sl_best, peer=step.peer)
            if step.peer is not None:
                return
            rec = P.gather_winner(sl_best, world)
            N.check(lib.amf_best_reduce(D.ptr(rec), world, 1, D.ptr(sl_best), D.stream_ptr()))
        for _ in range(3):
            strong_step()
        sync()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(a.steps):
            strong_step()
        s1.record()
        sync()
        st = torch.tensor([s0.elapsed_time(s1) / a.steps], dtype=torch.float64, device=U.device)
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        strong = {"ncand_total": ncand, "ncand_per_gpu": hi - lo, "ms": float(st.item()),
                  "candidates_per_sec": ncand / (float(st.item()) * 1e-3),
                  "note": "fixed total pool cut into N contiguous slices, pool kernel + 16-byte winner all-gather + reduction"}
        sl_pool.close()
        del sl_i, sl_j

    # ---- parity mode (fp64, the drop-in classes' default dtype) on the same workload ---------
    f64 = None
    if world == 1 and name == "f32" and not a.no_configs:
        from active_matrix_factorization_b200 import scoring as S
        rat64 = D.Ratings(n, m, prob["ri"], prob["rj"], prob["r"].double(), "f64")
        U64, V64 = U.double().contiguous(), V.double().contiguous()
        dU64, dV64 = torch.empty_like(U64), torch.empty_like(V64)
        step64 = P.ShardedStep(rat64, d, "f64", 1, 0)
        step64.pool = S.Pool(ci, cj, n, m, "f64", d)
        kt64 = step64.kernel_times(U64, V64, params, dU64, dV64, sums, ci, cj, best, reps=5)
        hbm_peak64, _ = peaks()
        tables64 = (n + m) * d * 8
        sb, gb = ncand * 8 + tables64, nnz * 16 + 2 * tables64
        f64 = {"roofline_f64": {"kernel": "pool_pred_kernel<double>", "bound": "hbm", "kernel_ms": kt64["score_ms"],
                                "algorithmic_bytes": sb, "achieved": sb / (kt64["score_ms"] * 1e-3) / 1e9,
                                "peak": hbm_peak64, "unit": "GB/s",
                                "frac": sb / (kt64["score_ms"] * 1e-3) / 1e9 / hbm_peak64, "traffic": None},
               "roofline_gradient_f64": {"kernel": "tiled_side_kernel<double> x2 + prior_kernel x2", "bound": "hbm",
                                         "kernel_ms": kt64["side_pass_ms"], "algorithmic_bytes": gb,
                                         "achieved": gb / (kt64["side_pass_ms"] * 1e-3) / 1e9, "peak": hbm_peak64,
                                         "unit": "GB/s", "frac": gb / (kt64["side_pass_ms"] * 1e-3) / 1e9 / hbm_peak64,
                                         "traffic": None,
                                         "note": "fp64 ratings: 16 algorithmic bytes per rating (i, j, r)"}}
        step64.pool.close()
        rat64.close()
        del U64, V64, dU64, dV64, step64, rat64
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_kind = peaks()
    traffic = measured_traffic() if (a.users, a.items, a.latent_d, a.nnz, a.ncand, a.dtype) == (
        200_000, 50_000, 32, 50_000_000, 100_000_000, "f32") else {}
    tables = (n + m) * d * es
    score_bytes = ncand * 8 + tables
    grad_bytes = nnz * 12 + 2 * tables
    score_gbs = score_bytes / (scorek_ms * 1e-3) / 1e9
    grad_gbs = grad_bytes / (side_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": ncand_all / (score_ms / a.steps * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": name, "data": "synthetic",
        "config": {"workload": workload_name(a), "criterion": "pred (MAP prediction) with fused arg-max, winner only", "pool_layout": a.pool,
                   "parallelism": ("single GPU" if world == 1 else
                                   "candidates and ratings sharded per GPU (weak), NCCL all-reduce of dU/dV/sums + all-gather of winners"
                                   if a.grad_shard == "ratings" else
                                   "candidates sharded per GPU, ratings and U rows sharded by user range (%d users per GPU), "
                                   "NCCL all-reduce of dV/sums + all-gather of winners" % n),
                   "l2": "inputs larger than the 126 MB L2: the kernels stream >= %.0f MB of rating entries per side (6 bytes each, bundled-runs layout) and >= %.0f MB of candidate indices per GPU (2 bytes each) every step" % (nnz * 6 / 1e6, ncand * 2 / 1e6),
                   "rating_layout": "tiled" if tiled_grad else "rows",
                   "value_is": "candidates / scoring-phase time; ms_per_step covers gradient + scoring"},
        "phases": {
            "pmf_loss_grad": {"ms": grad_ms / a.steps, "ratings_per_sec_iter": nnz_all / (grad_ms / a.steps * 1e-3), "nnz_total": nnz_all},
            "score_pred": {"ms": score_ms / a.steps, "candidates_per_sec": ncand_all / (score_ms / a.steps * 1e-3), "ncand_total": ncand_all},
        },
        "roofline": {"kernel": "pool_pred_kernel (bundled runs: V tile in shared memory via TMA, one lane per (user, tile) run, whole U row in registers)" if a.pool == "tiled" else "score_pred_kernel", "bound": "hbm", "achieved": score_gbs, "peak": hbm_peak,
                     "peak_source": peak_kind, "unit": "GB/s", "frac": score_gbs / hbm_peak,
                     "algorithmic_bytes": score_bytes, "kernel_ms": scorek_ms,
                     "traffic": traffic.get("pool_pred_kernel" if a.pool == "tiled" else "score_pred_kernel"),
                     "flat_kernel_ms": score_flat_ms, "pool_build_ms": pool_ms},
        "roofline_gradient": {"kernel": ("tiled_side_kernel x2 (bundled runs: own row + accumulator in registers, other side's tile in shared memory)" if tiled_grad else "side_pass_kernel x2") + " + prior_kernel x2", "bound": "hbm", "achieved": grad_gbs,
                              "peak": hbm_peak, "peak_source": peak_kind, "unit": "GB/s", "frac": grad_gbs / hbm_peak,
                              "algorithmic_bytes": grad_bytes, "kernel_ms": side_ms,
                              "traffic": traffic.get("tiled_side_kernel_x2" if tiled_grad else "side_pass_kernel_x2")},
        "e2e": {"value": ncand_all / e2e_score_s, "unit": UNIT,
                "h2d_bytes_per_step": int(ncand * (2 if narrow else 4) + (n + 1) * 8 + 2 * tables), "d2h_bytes_per_step": int(tables + 24 + 16),
                "pmf_ratings_per_sec_iter": nnz_all / e2e_grad_s,
                "pairs_form_value": ncand / e2e_pairs_s if world == 1 else None,
                "csr32_form_value": ncand / e2e_csr32_s if (world == 1 and e2e_csr32_s) else None,
                "note": "amf_pmf_loss_grad_host + %s (pool as row offsets + %s item ids) with pinned host buffers; "
                        "rating list resident; csr32_form_value = amf_score_pred_host_csr with 32-bit item ids (4 bytes per candidate), "
                        "pairs_form_value = amf_score_pred_host with (i, j) arrays (8 bytes per candidate)"
                        % (("amf_score_pred_host_csr16", "16-bit") if narrow else ("amf_score_pred_host_csr", "32-bit"))},
        "gpu_launches": a.steps * step.launches_per_step,
        "clocks": clocks,
        "selected": {"value": float(bv), "index": bi, "check": selection_check},
    }
    # the resource that binds both kernels: one random factor row per candidate
    # (and per rating and pass) through the SM's 128 B/clk load-store data pipe, on `sms` SMs
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    mhz = float(clocks["sm_mhz"]) if clocks and clocks.get("sm_mhz") else 1965.0
    row_clk = d * es / 128.0
    floor_score = ncand * row_clk / (sms * mhz * 1e3)
    floor_grad = 2 * nnz * row_clk / (sms * mhz * 1e3)
    line["binding_roof"] = {"resource": "SM load/store data pipe, 128 B/clk/SM: one %d-byte row per candidate, and per rating and pass" % (d * es),
                            "sms": sms, "sm_mhz": mhz,
                            "scoring_floor_ms": floor_score, "scoring_frac": floor_score / scorek_ms,
                            "gradient_floor_ms": floor_grad, "gradient_frac": floor_grad / side_ms}
    if world > 1:
        line["selection_collective"] = ("fused into the scoring kernel: its last CTA exchanges the winners over NVLink peer memory (CUDA IPC mailboxes, csrc/peer.cuh)"
                                        if step.peer is not None else "NCCL all-gather of 16-byte records + amf_best_reduce")
    if per_rank is not None:
        line["per_rank"] = per_rank
    if strong is not None:
        line["strong_scaling"] = strong
    if f64 is not None:
        line.update(f64)
    if not a.no_configs and world == 1:
        from benchmarks import config_lines as CL
        line["configs"] = CL.all_configs(hbm_peak)
        try:
            line["configs"]["c5_extra"] = CL.c5_extra(rat, n, m, d, ci, cj, hbm_peak, name)
        except Exception as exc:
            line["configs"]["c5_extra"] = {"failed": repr(exc)[:300]}
    if world == 1 and not a.no_cpu_baseline:
        host = (prob["ri"][:a.cpu_sample].cpu().numpy(), prob["rj"][:a.cpu_sample].cpu().numpy(),
                prob["r"][:a.cpu_sample].double().cpu().numpy(), U.cpu().numpy(), V.cpu().numpy(),
                ci[:a.cpu_sample].cpu().numpy(), cj[:a.cpu_sample].cpu().numpy())
        cb = cpu_reference_sample(a, host, a.cpu_sample)
        # parity spot check of the sample against the device path
        cb.pop("_best", None)
        line["cpu_baseline"] = cb
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
