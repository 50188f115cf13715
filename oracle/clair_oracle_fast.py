"""Multi-threaded fp32 CPU port of the forward (torch CPU ops).  TEST/BENCH INFRASTRUCTURE ONLY.

Same restatement as oracle/clair_oracle.py (which stays the fp64 ground truth; see its header for
the reference file:line map and the PARITY UNPINNED note), arranged the way a CPU framework would
run it - one big input-projection GEMM per direction, a [n,128]x[128,512] GEMM per time step, a
batched matmul for the 256 slice-dense units - so that ``bench.py``'s cpu_baseline /
``--impl reference`` leg is a fair "reference CPU path on all host threads" and not a strawman.
tests/test_oracle.py pins it to the numpy oracle.
"""
import numpy as np
import torch

from . import clair_oracle as O


class FastOracle(object):
    def __init__(self, weights, threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        self.threads = torch.get_num_threads()
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        self.lstm = {}
        for layer, fin in ((1, O.F), (2, 2 * O.H)):
            for d in ("fw", "bw"):
                k = t(weights[O.LSTM_NAME.format(layer="LSTM%d" % layer, d=d, v="kernel")])
                b = t(weights[O.LSTM_NAME.format(layer="LSTM%d" % layer, d=d, v="bias")])
                self.lstm[(layer, d)] = (k[:fin].contiguous(), k[fin:].contiguous(), b)
        k3, b3 = O.stack_l3(weights)
        self.k3 = t(k3)                       # [256,33,30]
        self.b3 = t(b3)                       # [256,30]
        self.w4, self.b4 = t(weights["L4/kernel"]), t(weights["L4/bias"])
        self.w5 = [t(weights["L5_%d/kernel" % (k + 1)]) for k in range(4)]
        self.b5 = [t(weights["L5_%d/bias" % (k + 1)]) for k in range(4)]
        self.wh = [t(weights["Prediction/%s/kernel" % n]) for n in O.HEAD_NAMES]
        self.bh = [t(weights["Prediction/%s/bias" % n]) for n in O.HEAD_NAMES]
        self.selu = torch.nn.SELU()           # same constants as clair/selu.py:28-29

    def _direction(self, x_tm, key, reverse):
        wx, wh, b = self.lstm[key]
        Tn, n, fin = x_tm.shape
        xs = torch.flip(x_tm, [0]) if reverse else x_tm
        pre = torch.addmm(b, xs.reshape(Tn * n, fin), wx).reshape(Tn, n, 4 * O.H)
        h = torch.zeros(n, O.H)
        c = torch.zeros(n, O.H)
        out = torch.empty(Tn, n, O.H)
        for t in range(Tn):
            z = torch.addmm(pre[t], h, wh)
            i, g, f, o = z.chunk(4, dim=1)                      # LSTMBlockCell order i, c, f, o
            c = torch.tanh(g) * torch.sigmoid(i) + c * torch.sigmoid(f)
            h = torch.tanh(c) * torch.sigmoid(o)
            out[t] = h
        return torch.flip(out, [0]) if reverse else out

    def _bilstm(self, x_tm, layer):
        return torch.cat([self._direction(x_tm, (layer, "fw"), False),
                          self._direction(x_tm, (layer, "bw"), True)], dim=2)

    @torch.no_grad()
    def forward_packed(self, X):
        """X [n,33,8,4] -> [n,90] float32 probabilities (21|3|33|33)."""
        x = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32))
        n = x.shape[0]
        x_tm = x.reshape(n, O.T, O.F).transpose(0, 1).contiguous()
        l2 = self._bilstm(self._bilstm(x_tm, 1), 2)             # [33,n,256]
        l3 = torch.baddbmm(self.b3[:, None, :], l2.permute(2, 1, 0), self.k3)     # [256,n,30]
        l3 = self.selu(l3).permute(1, 2, 0).reshape(n, O.L3_UNITS * 2 * O.H)      # index o*256+c
        l4 = self.selu(torch.addmm(self.b4, l3, self.w4))
        outs = []
        for k in range(4):
            a = self.selu(torch.addmm(self.b5[k], l4, self.w5[k]))
            z = self.selu(torch.addmm(self.bh[k], a, self.wh[k]))
            outs.append(torch.softmax(z, dim=1))
        return torch.cat(outs, dim=1).numpy()
