"""Per-class device time of the slab multigrid iteration over NCCL (run under torchrun, 2+ GPUs):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/run_slab_prof.py 4096"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import sayram2d_b200 as sy
from sayram2d_b200 import fields
from sayram2d_b200.shard import slab_range

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", rank=rank, world_size=world)
torch.cuda.set_device(rank)
ids = [sy.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
xe, ye = fields.uniform_edges(n, n)
lo, hi = slab_range(n, rank, world)
eng = sy.Engine(xe, ye, 0.002, device=rank, slab=(rank, world, ids[0]))
eng.set_options(precond=2, check_every=1)
Dxx, Dxy, Dyy, inv_tau = fields.synthetic_tensor(xe, ye, rows=(lo, hi))
eng.set_coeffs(fields.ay_G(xe, ye, rows=(lo, hi)), Dxx, Dxy, Dyy, inv_tau)
f0, bct, lines = fields.ay_init_and_bc(xe, ye, rows=(lo, hi))
eng.set_bc(bct, *lines)
eng.set_f(f0)
eng.step(2)
torch.cuda.synchronize(); dist.barrier()
st_graph = eng.step(2)
eng.set_profiling(True)
st = eng.step(1)
prof = eng.profile()
eng.set_profiling(False)
if rank == 0:
    print(json.dumps({"n": n, "world": world, "ms_per_step_graph": 1e3 * st_graph["seconds_device"] / 2, "iters": st_graph["iters_total"] / 2,
                      "profiled_step_ms": 1e3 * st["seconds_device"], "iters_profiled": st["iters_total"],
                      "classes": {k: {"ms": round(v["ms"], 2), "launches": v["launches"]} for k, v in prof.items() if v["launches"]}}))
eng.close()
dist.barrier()
dist.destroy_process_group()
