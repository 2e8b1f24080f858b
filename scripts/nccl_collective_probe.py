"""Collective bandwidth on this box, stand-alone (torchrun --nproc-per-node N scripts/nccl_collective_probe.py):
all-reduce and reduce-scatter of 256 MiB / 1 GiB in bf16 and fp32; bus bandwidth = algorithm bytes x factor / time."""
import os
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
for dt in (torch.bfloat16, torch.float32):
    for mib in (256, 1024):
        n = mib * (1 << 20) // dt.itemsize
        x = torch.ones(n, dtype=dt, device=dev)
        shard = torch.empty(n // world, dtype=dt, device=dev)
        for name, fn, factor in (("all_reduce", lambda: dist.all_reduce(x), 2 * (world - 1) / world),
                                 ("reduce_scatter", lambda: dist.reduce_scatter_tensor(shard, x), (world - 1) / world)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            if rank == 0:
                print(f"{name:15s} {str(dt):15s} {mib:5d} MiB  {ms:8.3f} ms  algbw {mib / 1024 * 1.0737 / (ms * 1e-3):7.1f} GB/s"
                      f"  busbw {mib / 1024 * 1.0737 * factor / (ms * 1e-3):7.1f} GB/s", flush=True)
dist.destroy_process_group()
