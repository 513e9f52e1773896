"""torchrun --nproc-per-node N scripts/multigpu_check.py: the population sharded over N GPUs (block partition, NCCL gather of
the device-resident records) gives, record for record and bit for bit, what ONE GPU gives for the whole population --
configs 2 (bipeds) and 4 (mixed shapes: the launch plan of a shard differs from the whole batch's, the results must not)."""
import importlib, os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
H = importlib.import_module("2dhmsr_b200")
sharding = importlib.import_module("2dhmsr_b200.sharding")
import workloads

rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok = True
for name, (loco, robots) in (("config 2", workloads.config2(8192 + 3, seed=0)), ("config 4", workloads.config4(6000 + 1, seed=2, final_t=5.0))):
    n = len(robots)
    b, e = sharding.shard_range(n, rank, ws)
    with H.DeviceBatch(loco.compile(robots[b:e])) as db:
        db.upload(lr).run()
        full = sharding.gather_results(db.results(), n, device=torch.device("cuda", lr))   # ragged shards: padded all_gather
    if rank == 0:
        with H.DeviceBatch(loco.compile(robots)) as db:
            db.upload(lr).run()
            one = db.results()
        same = full.tobytes() == one.tobytes()
        ok = ok and same
        print(f"{name}: {n} robots over {ws} GPUs == one GPU: {same}; status != 0: {int((one['status'] != 0).sum())}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
