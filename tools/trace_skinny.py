import ctypes as C, os, sys, numpy as np, torch
os.environ["CRAB_SK_TRACE"] = "1"
sys.path.insert(0, ".")
from crab_b200 import ops, lib
ops.init(0)
dev = torch.device("cuda:0")
L = lib.load()
for name, N, K, act, resid in [("qkv", 12288, 4192, 0, False), ("o", 4096, 4128, 0, True), ("gateup", 22016, 4160, 3, False)]:
    Ws = [ops.pack_skinny_weight(torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02) for _ in range(4)]
    x = torch.randn(32, K, device=dev, dtype=torch.bfloat16)
    out = torch.zeros(32, N // 2 if act else N, device=dev, dtype=torch.bfloat16)
    for i in range(8):
        ops.gemm_skinny(x, Ws[i % 4], act=act, out=out, residual=out if resid else None)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (148 * 8))()
    lib.check(L.crab_debug_skinny_trace(buf, C.c_int(148)))
    t = np.array(buf, dtype=np.uint64).reshape(148, 8).astype(np.int64)
    ok = t[:, 6] > 0
    base = t[ok, 1]
    names = ["wait_tfull(start->full)", "tmem+ws_stores", "bar1", "ticket+bar2", "ws_loads+sum", "epilogue_stores", ]
    print(name, "CTAs traced:", int(ok.sum()))
    for k in range(6):
        a, b = t[ok, k], t[ok, k + 1]
        v = (b - a)[(a > 0) & (b > 0)] / 1e3
        if len(v):
            print(f"   {names[k]:26s} mean {v.mean():7.2f} us   p50 {np.median(v):7.2f}   max {v.max():7.2f}   n={len(v)}")
    end = t[ok, 6]
    print(f"   spread of CTA finish times: {(end.max() - end.min()) / 1e3:.2f} us; full->end mean {((t[ok,6]-t[ok,1]).mean())/1e3:.2f} us")
