"""torchrun check of the sharded fused loss+gradient (parallel.ShardedStep.loss_grad): every rank
holds a block of the ratings; the all-reduced dU / dV / sums must equal a single-GPU pass over
all ratings, with and without the overlapped schedule.  Prints timings of both schedules.

    torchrun --nproc-per-node 2 benchmarks/check_sharded_grad.py"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from active_matrix_factorization_b200 import device as D, parallel as P

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
a = types.SimpleNamespace(users=200_000, items=50_000, latent_d=32, nnz=int(os.environ.get("CHECK_NNZ", "20000000")), ncand=1000, dtype="f32")
p = bench.make_problem(a, rank, torch)                       # factors are the same on every rank
n, m, d = a.users, a.items, a.latent_d
U, V = p["U"], p["V"]
rat = D.Ratings(n, m, p["ri"], p["rj"], p["r"], "f32")
step = P.ShardedStep(rat, d, "f32", world, rank)
params = D.pmf_params(1.0, 10.0, 10.0, 0.0)
out = {}
for overlap in (False, True):
    dU, dV, flat = P.alloc_grads(U, V)
    sums = torch.zeros(3, dtype=torch.float64, device="cuda")
    for _ in range(3):
        step.loss_grad(U, V, params, dU, dV, sums, flat, overlap=overlap)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step.loss_grad(U, V, params, dU, dV, sums, flat, overlap=overlap)
    e1.record(); torch.cuda.synchronize()
    out[overlap] = (dU.clone(), dV.clone(), sums.clone(), e0.elapsed_time(e1) / 20)
# reference: gather every rank's ratings on rank 0 and run one pass there
parts = [torch.empty_like(p["ri"]) for _ in range(world)]
sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
dist.all_gather(sizes, torch.tensor([p["ri"].numel()], device="cuda"))
mx = int(max(s.item() for s in sizes))
def gather(t):
    pad = torch.zeros(mx, dtype=t.dtype, device="cuda"); pad[:t.numel()] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:int(s.item())] for b, s in zip(bufs, sizes)])
ri, rj, r = gather(p["ri"]), gather(p["rj"]), gather(p["r"])
if rank == 0:
    full = D.Ratings(n, m, ri, rj, r, "f32")
    dU, dV = torch.empty_like(U), torch.empty_like(V)
    sums = D.loss_grad(full, d, U, V, params, dU, dV)
    for overlap in (False, True):
        gu, gv, sm, ms = out[overlap]
        eu = ((gu - dU).abs().max() / dU.abs().max()).item()
        ev = ((gv - dV).abs().max() / dV.abs().max()).item()
        es = ((sm - sums).abs() / sums.abs()).max().item()
        print("overlap=%s: %.3f ms per sharded pass; rel. error dU %.2e dV %.2e sums %.2e" % (overlap, ms, eu, ev, es))
        assert eu < 2e-5 and ev < 2e-5 and es < 1e-6
dist.destroy_process_group()
