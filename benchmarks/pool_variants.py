"""Times pool_pred_kernel on the C5 candidate pool for the item-tile sizes (KB of shared memory)
given on the command line, each in a child process."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import types
    import torch
    import bench
    from active_matrix_factorization_b200 import scoring as S
    tile_kb = int(sys.argv[2])
    a = types.SimpleNamespace(users=200_000, items=50_000, latent_d=32, nnz=1_000_000, ncand=100_000_000, dtype="f32")
    torch.cuda.set_device(0)
    p = bench.make_problem(a, 0, torch)
    pool = S.Pool(p["ci"], p["cj"], a.users, a.items, "f32", 32, tile_bytes=tile_kb * 1024)
    best = torch.zeros(2, dtype=torch.int64, device="cuda")
    for _ in range(3):
        pool.score_pred(p["U"], p["V"], best=best)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        pool.score_pred(p["U"], p["V"], best=best)
    e1.record(); torch.cuda.synchronize()
    print("tile=%d KB (%d rows): %.4f ms  best=%s" % (tile_kb, pool.tile_rows, e0.elapsed_time(e1) / 20,
                                                  S.unpack_best(best)))
else:
    for tile in sys.argv[1:] or ["224"]:
        subprocess.run([sys.executable, __file__, "--child", tile], check=True)
