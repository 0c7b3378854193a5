"""TEST INFRASTRUCTURE: the sliver of torch that bench.py's single-GPU flow touches, so that the flow itself (replays, timed arms,
ring timing of one kernel, JSON assembly) can be dry-run on the kernel-logic emulation without a GPU
(tests/test_bench_dry_run.py: PYTHONPATH=tests/emu/fake_torch python bench.py --lib tests/emu/_build/libhikari_emu.so ...).
Times are host wall-clock; nothing measured this way is a bench number."""
import ctypes
import time

import numpy as np

uint8 = np.uint8
float16 = np.float16
int32 = np.int32
float64 = np.float64


class _Tensor:
    def __init__(self, arr=None, n=0):
        self.arr = arr if arr is not None else np.zeros(n, np.uint8)
        self.device = "cuda:0"

    def pin_memory(self):
        return self

    def data_ptr(self):
        return self.arr.ctypes.data

    def numel(self):
        return self.arr.size


def empty(n, dtype=np.uint8, pin_memory=False, device=None):
    return _Tensor(np.zeros(n, dtype))


zeros = empty


def as_tensor(obj, device=None):
    return _Tensor(n=1)


def device(kind, index=0):
    return f"{kind}:{index}"


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3

    def synchronize(self):
        pass


class _Stream:
    def __init__(self, device=None):
        self.cuda_stream = None      # NULL: the context creates its own (emulated) stream

    def wait_event(self, e):
        pass


class cuda:
    Event = _Event
    Stream = _Stream

    @staticmethod
    def is_available():
        return True

    @staticmethod
    def set_device(i):
        pass

    @staticmethod
    def set_stream(s):
        pass

    @staticmethod
    def synchronize():
        pass
