#!/bin/bash
# Multi-GPU gradient, the two sharding schemes side by side:
#   ratings: rating blocks over all users, dU and dV all-reduced (32 MB at C5)      -- default
#   users:   each GPU owns the ratings and the U rows of its user range, only dV travels (6.4 MB)
# usage: benchmarks/grad_shard_users.sh [N_GPUS]      (needs N B200s; prints ms of the gradient phase)
N=${1:-8}
for mode in ratings users; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port $((29600 + RANDOM % 200)) bench.py --gpus "$N" --steps 50 --warmup 3 \
      --no-cpu-baseline --e2e-steps 1 --grad-shard "$mode" 2>/dev/null |
    python -c "
import json, sys
l = json.loads(sys.stdin.read())
p = l['phases']['pmf_loss_grad']
print('$mode: gradient phase %.3f ms, %.3e ratings/s/iter on %d GPUs' % (p['ms'], p['ratings_per_sec_iter'], l['n_gpus']))"
done
