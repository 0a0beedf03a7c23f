"""Driver of tests/test_multi_rank_cpu.py::test_bench_collectives_are_rank_symmetric (not a test module itself).

Runs bench.py's run_b200() as one rank of a WORLD_SIZE-rank job ON THE CPU: the torch.distributed collectives are real (gloo), the
GPU library is replaced by a stub whose calls do nothing and whose timers return constants.  What is exercised is the control
flow of the script around its barriers and all_reduces: if any rank issues a collective the others do not (for example after
ranks != 0 have left), the job deadlocks and the test's timeout fires."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeLib(object):
    def __getattr__(self, name):
        return lambda *a, **k: 0


class FakeCtx(object):
    handle = C.c_void_p(1)
    device = 0

    def __init__(self, *a, **k):
        self._launches = 0

    def pinned_array(self, shape, dtype):
        return np.zeros(shape, dtype)

    def device_alloc(self, nbytes):
        return 4096

    def timer_elapsed_ms(self):
        return 1.0

    def launch_count(self):
        self._launches += 7
        return self._launches

    def profile_read(self):
        return {"stream_level01": dict(ms=1.0, launches=6, bytes=6e9), "stream_down2": dict(ms=0.3, launches=6, bytes=1e9),
                "lk_windowed": dict(ms=0.7, launches=3, bytes=1e9)}

    def traffic_mix_probe(self, kind, total_bytes, reps=10):
        return 5500.0, 0.1

    def __getattr__(self, name):            # sync, memcpy, check, profile, profile_reset, timer_start, timer_stop, device_free, ...
        return lambda *a, **k: None


class FakePyramid(object):
    handle = C.c_void_p(2)

    def __init__(self, *a, **k):
        pass

    def nbytes(self):
        return 1 << 20

    def __getattr__(self, name):            # build_u8, close, ...
        return lambda *a, **k: None


class FakeSequence(object):
    def __init__(self, ctx, params, taps, w, h, n_sequences, n_features, precision, select_mode):
        self.S, self.n = n_sequences, n_features

    def features(self):
        z = np.zeros((self.S, self.n))
        return z, z, np.zeros((self.S, self.n), np.int32), np.zeros((self.S, self.n), np.int32)

    def select_stats(self):
        return np.zeros((self.S, 4))

    def sync(self):
        return 0

    def uses_graph(self):
        return True

    def __getattr__(self, name):            # start, step, close
        return lambda *a, **k: None


class FakeFeature(object):
    x, y, val = 10.0, 10.0, 0


def main():
    import torch
    import torch.distributed as dist
    real_init, real_tensor = dist.init_process_group, torch.tensor
    dist.init_process_group = lambda backend, **k: real_init("gloo")
    # every collective this rank issues, in order: the test compares the ranks' lists (a collective one rank issues and another
    # does not is a deadlock under NCCL; gloo may turn it into an exception that the script's try/except swallows)
    log = []
    real_barrier, real_all_reduce = dist.barrier, dist.all_reduce

    def barrier(*a, **k):
        log.append("barrier")
        return real_barrier(*a, **k)

    def all_reduce(t, *a, **k):
        log.append("all_reduce[%d]" % t.numel())
        return real_all_reduce(t, *a, **k)
    dist.barrier, dist.all_reduce = barrier, all_reduce
    torch.tensor = lambda data, **k: real_tensor(data, **{a: b for a, b in k.items() if a != "device"})
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.synchronize = lambda *a, **k: None
    from pyfeaturetrack_b200 import _capi, selectGoodFeatures as sgf, shard
    fake_ctx = FakeCtx()
    _capi.set_device = lambda *a, **k: None
    _capi.default_ctx = lambda: fake_ctx
    _capi.lib = lambda: FakeLib()
    _capi.Context = FakeCtx
    _capi.Pyramid = FakePyramid
    _capi.Sequence = FakeSequence
    sgf.KLTSelectGoodFeatures = lambda tc, img, n: [FakeFeature() for _ in range(n)]
    real_gather = shard.gather_features
    shard.gather_features = lambda local, n_units, n, d=None, device=None: real_gather(local, n_units, n, d, None)
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.environ.get("KLT_BENCH_PATH", os.path.join(ROOT, "bench.py")))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    bench.motion_cycle = lambda H, W, seed, period=100: np.zeros((period, H, W), np.uint8)
    bench.WORKLOADS["B"] = dict(bench.WORKLOADS["B"], H=48, W=64, n=10)      # the sequence leg hard-codes workload B
    sys.argv = ["bench.py", "--gpus", os.environ["WORLD_SIZE"], "--steps", "4", "--warmup", "3", "--workload", "A", "--pairs", "2",
                "--distinct", "1", "--seqs", "1", "--seq-frames", "3", "--sustain-s", "0.001", "--no-cpu-baseline", "--api-pairs", "0"]
    try:
        bench.main()
    finally:
        with open(os.path.join(os.environ["KLT_FLOW_OUT"], "collectives_%s.txt" % os.environ["RANK"]), "w") as f:
            f.write("\n".join(log) + "\n")


if __name__ == "__main__":
    main()
