"""Raw host<->device copy ceiling of the box, all GPUs at once: no kernels, no library -- just pinned buffers and
cudaMemcpyAsync on two streams per GPU (is the flat 8-GPU end-to-end curve the host's ceiling or the
3-slot pipeline's?).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        profiles/pcie_ceiling.py > gpurun_out/pcie_ceiling_N.json

One process per GPU, each with its own pinned 1 GiB source and 1 GiB destination; every mode is started behind a
barrier so that all ranks copy concurrently; wall-clock of the slowest rank.  Modes: H2D alone, D2H alone, both
directions at once (what fwd_host + inv_host need), whole-buffer copies and 32 MiB chunks (the pipeline's granularity).
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        saved = os.dup(1); os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier(); torch.cuda.synchronize()
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    GiB = 1 << 30
    words = GiB // 4
    h_src = torch.empty(words, dtype=torch.int32).pin_memory()
    h_dst = torch.empty(words, dtype=torch.int32).pin_memory()
    h_src.fill_(rank + 1)
    d_in = torch.empty(words, dtype=torch.int32, device=dev)
    d_out = torch.ones(words, dtype=torch.int32, device=dev)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()

    def run(mode, chunk_words, reps=4):
        def once():
            for off in range(0, words, chunk_words):
                sl = slice(off, off + chunk_words)
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s_up):
                        d_in[sl].copy_(h_src[sl], non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s_dn):
                        h_dst[sl].copy_(d_out[sl], non_blocking=True)
        once(); torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_dir = GiB * reps / float(t) / 1e9
        return per_dir

    out = {"gpus": world, "buffer_GiB_per_direction_per_gpu": 1, "host_cpus": len(os.sched_getaffinity(0)), "modes": {}}
    for name, chunk in (("whole", words), ("32MiB_chunks", (32 << 20) // 4)):
        for mode in ("h2d", "d2h", "both"):
            g = run(mode, chunk)
            out["modes"][f"{mode}_{name}"] = {"GBps_per_direction_per_gpu": round(g, 2),
                                              "GBps_per_direction_aggregate": round(g * world, 2),
                                              "GBps_combined_aggregate": round(g * world * (2 if mode == "both" else 1), 2)}
    if rank == 0:
        try:
            out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
        except Exception:
            out["numa_nodes"] = None
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
