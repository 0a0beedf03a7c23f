"""Multi-GPU sharding of independent tracking units (frame pairs or whole sequences): SURVEY 8(e).

One process per GPU.  Units are independent, so there is NO collective on the data path: rank r processes the units
shard_range(n, r, world) on its own GPU; only the small per-unit feature lists (x, y, val: 20 bytes per feature) are
gathered at the end (torch.distributed, NCCL on GPUs / gloo on CPU)."""
import numpy as np


def shard_range(n_units, rank, world):
    """Contiguous block partition: the first (n_units % world) ranks get one extra unit."""
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_features(local, n_units, n_features, dist=None, device=None):
    """local: {unit index: (x[n], y[n], val[n])} for this rank's units -> on every rank the full
    (x[n_units, n], y[n_units, n], val[n_units, n]) arrays.  Fixed-size tensors, one all_reduce per array
    (every slot is written by exactly one rank, the others contribute zeros)."""
    x = np.zeros((n_units, n_features))
    y = np.zeros((n_units, n_features))
    v = np.zeros((n_units, n_features), np.int32)
    for u, (ux, uy, uv) in local.items():
        x[u], y[u], v[u] = ux, uy, uv
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return x, y, v
    import torch
    out = []
    for a in (x, y, v):
        t = torch.from_numpy(a)
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out.append(t.cpu().numpy())
    return out[0], out[1], out[2].astype(np.int32)


def bind_host_to_gpu(device_index):
    """Pin this process to the CPUs that are local to its GPU (NVML's ideal CPU affinity = the GPU's NUMA node), so that
    the pinned staging buffers it allocates afterwards and the threads that fill them sit on the memory controller next
    to the GPU's PCIe root.  With one process per GPU this keeps N uploads from crowding one socket's memory.
    Returns the number of CPUs bound to, or 0 when NVML / the affinity call is unavailable (nothing changes then)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0
