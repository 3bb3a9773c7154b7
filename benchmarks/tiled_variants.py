"""Times the tiled fused loss+gradient (tiled_side_kernel x2 + prior x2) on the C5 rating list
for the tile sizes (KB of shared memory, AMF_TILED_KB) given on the command line, each in a
child process; 'rows' times the row-sorted kernels."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import ctypes as C, types
    import torch
    import bench
    from active_matrix_factorization_b200 import _native as N, device as D
    a = types.SimpleNamespace(users=int(os.environ.get("TV_USERS", 200_000)), items=int(os.environ.get("TV_ITEMS", 50_000)),
                              latent_d=32, nnz=50_000_000, ncand=1_000_000, dtype=os.environ.get("TV_DTYPE", "f32"))
    torch.cuda.set_device(0)
    p = bench.make_problem(a, 0, torch)
    rat = D.Ratings(a.users, a.items, p["ri"], p["rj"], p["r"], a.dtype)
    rat.set_layout(sys.argv[2])
    lib = N.require_device()
    U, V = p["U"], p["V"]
    dU, dV = torch.empty_like(U), torch.empty_like(V)
    sums = torch.zeros(3, dtype=torch.float64, device="cuda")
    params = D.pmf_params(1.0, 10.0, 10.0, 0.0)
    def run():
        N.check(lib.amf_pmf_loss_grad(rat.handle, D.code(a.dtype), 32, 32, D.ptr(U), D.ptr(V), C.byref(params),
                                      D.ptr(dU), D.ptr(dV), D.ptr(sums), D.stream_ptr()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    print("%s tile_kb=%s: %.4f ms  sums=%s |dU|=%.6e |dV|=%.6e" % (
        sys.argv[2], os.environ.get("AMF_TILED_KB"), e0.elapsed_time(e1) / 20,
        sums.cpu().numpy(), dU.double().norm().item(), dV.double().norm().item()))
else:
    for spec in sys.argv[1:] or ["224"]:
        if spec == "rows":
            subprocess.run([sys.executable, __file__, "--child", "rows"], check=True)
            continue
        env = dict(os.environ, AMF_TILED_KB=spec)
        subprocess.run([sys.executable, __file__, "--child", "tiled"], env=env, check=True)
