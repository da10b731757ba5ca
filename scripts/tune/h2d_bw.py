"""Host->device copy bandwidth of this box: regular pinned vs write-combined pinned memory, one large copy and the e2e chunk sizes (dev tool)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch


def wc_pinned(nbytes):
    rt = ctypes.CDLL("libcudart.so")
    ptr = ctypes.c_void_p()
    err = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))  # cudaHostAllocWriteCombined
    assert err == 0, err
    buf = (ctypes.c_byte * nbytes).from_address(ptr.value)
    return torch.frombuffer(buf, dtype=torch.float32)


def bw(host, dev, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 0.0
    for _ in range(reps):
        e0.record()
        dev.copy_(host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, host.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


torch.cuda.init()
for mb in (16, 256, 1024):
    n = mb * (1 << 20) // 4
    dev = torch.empty(n, device="cuda")
    reg = torch.empty(n).pin_memory()
    reg.fill_(1.0)
    wc = wc_pinned(n * 4)
    wc.fill_(1.0)
    print("H2D %5d MB: pinned %.1f GB/s (is_pinned %s) | write-combined %.1f GB/s (is_pinned %s)" % (mb, bw(reg, dev), reg.is_pinned(), bw(wc, dev), wc.is_pinned()), flush=True)
    back = torch.empty(n).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); back.copy_(dev, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print("D2H %5d MB: pinned %.1f GB/s" % (mb, n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9), flush=True)
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
os.system("nvidia-smi topo -m 2>/dev/null | head -8; cat /sys/devices/system/node/online 2>/dev/null")
