"""Clock / power under a sustained loop of the device-resident forward: python tools/power_probe.py [sites] [seconds]
Prints per-kernel ms, median SM clock and power - the forward runs under the 1000 W cap, so cycles and time part ways."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clair_b200 import synth, weights as W, _lib
from clair_b200.model import Clair
n = int(sys.argv[1]) if len(sys.argv) > 1 else 151552
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
m = Clair(max_sites=n, batch_sites=1000)
m.set_weights(W.random_weights(seed=1234))
X = synth.synthetic_tensors(min(n, 4000), seed=1)
X = np.concatenate([X] * (n // len(X) + 1))[:n]
xd = torch.from_numpy(X).cuda()
od = torch.empty((n, 90), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream()
lines = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                     stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append((time.time(), l)) for l in p.stdout], daemon=True).start()
for _ in range(3):
    m.predict_device(xd.data_ptr(), _lib.DTYPE_F32, n, od.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
time.sleep(0.5)
m.set_profiling(True)
t0 = time.time()
reps = 0
while time.time() - t0 < secs:
    for _ in range(4):
        m.predict_device(xd.data_ptr(), _lib.DTYPE_F32, n, od.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    reps += 4
t1 = time.time()
time.sleep(0.1)
p.terminate()
rows = [tuple(float(v) for v in l.split(",")) for ts, l in lines if t0 + 0.3 <= ts <= t1]
clk = sorted(r[0] for r in rows); pw = sorted(r[1] for r in rows)
prof = m.read_profile()
print("sites/s %.2f M  clock median %s MHz  power median %s W  (%d samples)  " % (
    n * reps / (t1 - t0) / 1e6, clk[len(clk) // 2] if clk else None, pw[len(pw) // 2] if pw else None, len(rows)) +
    " ".join("%s=%.3f" % (q["kernel"], q["ms"] / q["launches"]) for q in prof), flush=True)
m.close()
