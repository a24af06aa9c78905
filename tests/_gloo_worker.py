"""world_size-2 worker for tests/test_multirank.py: the N>1 aggregation of bench.py on the gloo backend."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist

import bench

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
# rank r "measured" 100 + 50 r ms for 20 steps, and 300 - 10 r ms end to end
mx = bench.max_over_ranks([100.0 + 50.0 * rank, 300.0 - 10.0 * rank], world)
dist.barrier()
if rank == 0:
    print(json.dumps({"max_ms": mx, "value": bench.aggregate_value(world, 20, mx[0]), "world": world}))
dist.destroy_process_group()
