import ctypes as C, sys, torch
sys.path.insert(0, ".")
from crab_b200 import ops, lib
ops.init(0)
from crab_b200 import build as _b
L = C.CDLL(str(_b.build_diag()))
L.crab_last_error.restype = C.c_char_p
dev = torch.device("cuda:0")
nbuf, size = 6, 192 << 20
bufs = [torch.randint(0, 255, (size,), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
sink = torch.zeros(2, dtype=torch.int64, device=dev)
def run(chunk, stages, ctas):
    def fn():
        for b in bufs:
            assert 0 == (L.crab_debug_stream(C.c_void_p(b.data_ptr()), C.c_int64(size), C.c_int(chunk), C.c_int(stages), C.c_int(ctas),
                                          C.c_void_p(sink.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / nbuf
    print(f"chunk={chunk//1024:3d}KB stages={stages:2d} ctas={ctas:3d}: {us:7.1f} us/launch  {size/us/1e3:7.1f} GB/s", flush=True)
for chunk, stages in [(16384, 5), (16384, 10), (32768, 3), (32768, 6), (65536, 3)]:
    for ctas in (148, 296):
        if ctas == 296 and stages * chunk > 110 * 1024:
            continue
        run(chunk, stages, ctas)
