"""dev: HeadEngine (fused BAN heads from neck features) device-resident and end-to-end rates for a few chunk sizes."""
import json
import sys
import os
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from hdn_b200 import _lib, head_engine as he  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "256/512"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
chunks = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2, 4, 8]
IT = int(os.environ.get("HEAD_BENCH_ITERS", "10"))
if "HEAD_BENCH_SHIFT" in os.environ:  # 0: conv_search through the generic implicit GEMM (tensor-memory-operand kernel for large launches)
    from hdn_b200 import ops as _ops
    _ops.set_conv_shift(int(os.environ["HEAD_BENCH_SHIFT"]))
dev = torch.device("cuda", 0)
host = he.make_inputs(wl, B, seed=1, pin=True, u8_crop=True)
up = lambda v: [t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)  # noqa: E731
dev_in = {k: up(host[k]) for k in he.FRAME_KEYS}
zf, zf_lp = up(host["zf"]), up(host["zf_lp"])
for c in chunks:
    eng = he.HeadEngine(wl, B, dev, chunk=c)
    eng.set_template(zf, zf_lp)
    eng.bind(dev_in)
    for _ in range(3):
        eng.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    e0.record()
    for _ in range(IT):
        eng.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / IT
    launches = (_lib.launch_count() - n0) // IT
    h2d, d2h = eng.alloc_host_io(host)
    for _ in range(3):
        eng.run_host(host)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(IT):
        eng.run_host(host, wait=False)
    eng.finish()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / IT
    ms2 = e0.elapsed_time(e1) / IT
    print(json.dumps({"workload": wl, "B": B, "chunk": c, "resident_ms": ms, "resident_fps": B / ms * 1e3, "launches": launches, "e2e_ms": ms2, "e2e_wall_ms": wall * 1e3,
                      "e2e_fps": B / ms2 * 1e3, "h2d_MB": h2d / 1e6, "d2h_MB": d2h / 1e6, "h2d_GBs": h2d / ms2 / 1e6}), flush=True)
    del eng
