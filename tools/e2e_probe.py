"""Where does the end-to-end call lose time against the device-resident forward?  python tools/e2e_probe.py"""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clair_b200 import synth, weights as W, _lib
from clair_b200.model import Clair, pinned_empty
n = 75000
m = Clair(max_sites=n, batch_sites=1000)
m.set_weights(W.random_weights(seed=1234))
X = pinned_empty((n, 33, 8, 4), np.float32)
Xs = synth.synthetic_tensors(5000, seed=1)
for i in range(0, n, 5000):
    X[i:i + 5000] = Xs
out_pin = pinned_empty((n, 90), np.float32)
out_page = np.empty((n, 90), np.float32)
out_page[...] = 0
xd = torch.from_numpy(X).cuda()
od = torch.empty((n, 90), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream()
def timeit(f, reps=8):
    f(); f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def dev():
    m.predict_device(xd.data_ptr(), _lib.DTYPE_F32, n, od.data_ptr(), st.cuda_stream)
def call(dst):
    return lambda: m._lib.clairb_predict(m._h, X.ctypes.data_as(ctypes.c_void_p), _lib.DTYPE_F32, n, dst.ctypes.data_as(ctypes.c_void_p))
for name, f in (("device-resident", dev), ("clairb_predict -> pinned out", call(out_pin)), ("clairb_predict -> reused pageable out", call(out_page)),
                ("predict_packed (fresh array per call)", lambda: m.predict_packed(X))):
    ms = timeit(f)
    print("%-40s %.3f ms  %.2f M sites/s" % (name, ms, n / ms / 1e3), flush=True)
t0 = time.perf_counter(); xd2 = torch.from_numpy(X).cuda(); torch.cuda.synchronize(); print("H2D 317MB pinned: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
m.close()
