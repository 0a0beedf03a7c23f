#!/usr/bin/env python
"""Raw pinned host->device bandwidth with N ranks uploading at once -- the ceiling of bench.py's host-fed (e2e) numbers.

    python tools/h2d_probe.py                                   # one rank
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py

Every rank copies the bench step's frame bytes (2 x 64 x 1080p uint8 = 265 MB) from pinned memory to its GPU in 8 chunks on one
stream, all ranks starting together (barrier), CUDA-event timed, max over ranks.  Variants: default pinned memory and
write-combined pinned memory (cudaHostAllocWriteCombined).  Rank 0 prints one JSON line."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 64 * 2 * 1080 * 1920
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    rt = ctypes.CDLL("libcudart.so.12")          # the runtime torch has already loaded
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    out = {"ranks": world, "bytes_per_rank": n}
    for name, flags in (("pinned", 0), ("write_combined", 4)):          # cudaHostAllocWriteCombined = 0x04
        if flags:
            p = ctypes.c_void_p()
            if rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags)) != 0:
                continue
            buf = (ctypes.c_ubyte * n).from_address(p.value)
            ctypes.memset(p.value, 7, n)
            src_ptr = p.value
        else:
            h = torch.empty(n, dtype=torch.uint8).pin_memory()
            h.fill_(7)
            src_ptr = h.data_ptr()
        stream = torch.cuda.current_stream().cuda_stream
        chunks, reps = 8, 12

        def run(reps_):
            for _ in range(reps_):
                for c in range(chunks):
                    a, b = c * n // chunks, (c + 1) * n // chunks
                    rc = rt.cudaMemcpyAsync(d.data_ptr() + a, src_ptr + a, b - a, 1, stream)
                    assert rc == 0, rc
        run(2)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(reps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        out[name] = {"ms_per_265MB": round(ms, 3), "gbps_per_rank": round(n / ms / 1e6, 2), "gbps_aggregate": round(world * n / ms / 1e6, 2)}
        if flags:
            rt.cudaFreeHost(p)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
