import importlib, os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sp = importlib.import_module("stark-prover_b200")
synth = importlib.import_module("stark-prover_b200.synthetic")
P = sp.P_DEFAULT
ctx = sp.Context()
L = sp.lib()
c = synth.synthetic_poly_exact_degree(43, 1 << 21)
ch = sp.Channel(P)
pr = sp.fri_commit(ctx, ctx.upload(c), sp.CosetFri(ctx, 5, 24), ch)
out = np.zeros(1 << 16, dtype=np.uint8)
n = C.c_size_t(0)
idx = np.array([12345], dtype=np.uint64)
def open_once(first_layer):
    L.stark_fri_open_layers(pr.h, first_layer, C.c_void_p(idx.ctypes.data), 1, C.c_void_p(out.ctypes.data), out.size, C.byref(n))
for fl in (0, 11, 21):
    for _ in range(20): open_once(fl)
    t0 = time.perf_counter()
    for i in range(200):
        idx[0] = (i * 7919) % (1 << 24)
        open_once(fl)
    dt = (time.perf_counter() - t0) / 200
    print(f"open of layers >= {fl:2d} ({n.value} bytes): {dt * 1e6:.1f} us per call")
t0 = time.perf_counter()
for i in range(200): ctx.sync()
print(f"ctx.sync on an idle stream: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")
