"""Per-kernel timing of one forward (no parity check): python tools/kbench.py [sites]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clair_b200 import synth, weights as W, _lib
from clair_b200.model import Clair
n = int(sys.argv[1]) if len(sys.argv) > 1 else 37000
m = Clair(max_sites=n, batch_sites=1000)
m.set_weights(W.random_weights(seed=1234))
X = synth.synthetic_tensors(min(n, 4000), seed=1)
X = np.concatenate([X] * (n // len(X) + 1))[:n]
xd = torch.from_numpy(X).cuda()
od = torch.empty((n, 90), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream()
for _ in range(3):
    m.predict_device(xd.data_ptr(), _lib.DTYPE_F32, n, od.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
m.set_profiling(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
R = 5
for _ in range(R):
    m.predict_device(xd.data_ptr(), _lib.DTYPE_F32, n, od.data_ptr(), st.cuda_stream)
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / R
print("sites %d  %.3f ms  %.2f M sites/s  (%.1f ns/site)" % (n, ms, n / ms / 1e3, ms * 1e6 / n), " ".join(
    "%s=%.3f" % (p["kernel"], p["ms"] / p["launches"]) for p in m.read_profile()), flush=True)
m.close()
