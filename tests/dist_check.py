"""torchrun --nproc-per-node N tests/dist_check.py : the real multi-GPU path (both exchange
modes) checked against a gathered host sort.  Run under `gpurun --gpus N`."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200sort import datagen  # noqa: E402
from b200sort.dist import DistSorter  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bad = 0
    for exchange in ("p2p", "nccl"):
        for dist_name, n in (("uniform", 1 << 22), ("skewed90", 1 << 20), ("ascending", 300000), ("all_equal", 5000)):
            keys = datagen.make(dist_name, n, seed=70 + rank)
            sorter = DistSorter(n, exchange=exchange, headroom=float(world) + 0.5)
            d = torch.from_numpy(keys).cuda()
            for _ in range(2):
                out, m = sorter.sort(d)
            torch.cuda.synchronize()
            mine = out.cpu().numpy().copy()
            gathered_in = [None] * world
            gathered_out = [None] * world
            dist.all_gather_object(gathered_in, keys)
            dist.all_gather_object(gathered_out, mine)
            if rank == 0:
                ok = np.concatenate(gathered_out).tobytes() == np.sort(np.concatenate(gathered_in)).tobytes()
                print(f"{exchange:5s} {dist_name:10s} n/rank={n:8d} sizes={[len(x) for x in gathered_out]} {'ok' if ok else 'MISMATCH'}", flush=True)
                bad += (not ok)
            sorter.close()
    if rank == 0:
        print("DIST CHECK", "PASSED" if bad == 0 else f"FAILED ({bad})", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
