"""Third restatement of the forward: torch.nn.LSTM (cuDNN's fp32 LSTM when run on a GPU) + torch dense trunk.
TEST/BENCH INFRASTRUCTURE ONLY (see oracle/clair_oracle.py for the reference file:line map and the PARITY UNPINNED note).

Why it exists: the arithmetic of the reference lives in TensorFlow 1.13 (`CudnnCompatibleLSTMCell`,
clair/model.py:282-312), which cannot be installed here, so the numpy oracle cannot be pinned on the reference's own
outputs.  `CudnnCompatibleLSTMCell` is by construction the cell whose weights load into cuDNN's LSTM and give the same
results (that is what the released models, trained with `CudnnLSTM`, rely on: clair/model.py:281-296).  Running cuDNN's own
LSTM on the same weights is therefore the closest independent implementation of the reference's recurrence that exists
on the GPU box: a third implementation, written by neither the reference's authors nor this repository, that must agree
with the numpy oracle (tests/test_gpu_parity.py::test_cudnn_restatement_agrees_with_the_oracle, bench.py `oracle_pinning`).

Weight mapping: TF kernel [(in+128), 512] has rows [x; h] and gate columns i, c(candidate), f, o (oracle/clair_oracle.py);
torch wants weight_ih [512, in] / weight_hh [512, 128] with gate rows i, f, g(candidate), o and two bias vectors.
TF32 is switched off for the duration of a call: the comparison is about fp32 arithmetic.
"""
import numpy as np
import torch

from . import clair_oracle as O

_TORCH_GATES = (0, 2, 1, 3)      # torch gate block k (i, f, g, o) <- TF gate block (i, c, f, o)[_TORCH_GATES[k]]


def _regate(a):
    """[..., 512] in TF gate order -> torch gate order along the last axis."""
    blocks = np.split(np.asarray(a), 4, axis=-1)
    return np.concatenate([blocks[g] for g in _TORCH_GATES], axis=-1)


class CudnnOracle(object):
    def __init__(self, weights, device="cuda", dtype=torch.float32):
        self.device, self.dtype = torch.device(device), dtype
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device, dtype)
        self.lstm = []
        for layer, fin in ((1, O.F), (2, 2 * O.H)):
            rnn = torch.nn.LSTM(input_size=fin, hidden_size=O.H, num_layers=1, bidirectional=True).to(self.device, dtype)
            with torch.no_grad():
                for d, suffix in (("fw", ""), ("bw", "_reverse")):
                    k = _regate(weights[O.LSTM_NAME.format(layer="LSTM%d" % layer, d=d, v="kernel")])
                    b = _regate(weights[O.LSTM_NAME.format(layer="LSTM%d" % layer, d=d, v="bias")])
                    getattr(rnn, "weight_ih_l0" + suffix).copy_(t(k[:fin].T))
                    getattr(rnn, "weight_hh_l0" + suffix).copy_(t(k[fin:].T))
                    getattr(rnn, "bias_ih_l0" + suffix).copy_(t(b))
                    getattr(rnn, "bias_hh_l0" + suffix).zero_()
            rnn.flatten_parameters()
            self.lstm.append(rnn.eval())
        k3, b3 = O.stack_l3(weights)
        self.k3, self.b3 = t(k3), t(b3)
        self.w4, self.b4 = t(weights["L4/kernel"]), t(weights["L4/bias"])
        self.w5 = [t(weights["L5_%d/kernel" % (k + 1)]) for k in range(4)]
        self.b5 = [t(weights["L5_%d/bias" % (k + 1)]) for k in range(4)]
        self.wh = [t(weights["Prediction/%s/kernel" % n]) for n in O.HEAD_NAMES]
        self.bh = [t(weights["Prediction/%s/bias" % n]) for n in O.HEAD_NAMES]
        self.selu = torch.nn.SELU()

    @torch.no_grad()
    def forward(self, X):
        """X [n,33,8,4] -> (probabilities [n,90], post-SELU logits [n,90]) as numpy arrays of the working precision."""
        tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            x = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(self.device, self.dtype)
            n = x.shape[0]
            h = x.reshape(n, O.T, O.F).transpose(0, 1).contiguous()                     # time-major [33,n,32]
            for rnn in self.lstm:
                h, _ = rnn(h)                                                            # [33,n,256] = fw | bw
            l3 = torch.baddbmm(self.b3[:, None, :], h.permute(2, 1, 0), self.k3)         # [256,n,30]
            l3 = self.selu(l3).permute(1, 2, 0).reshape(n, O.L3_UNITS * 2 * O.H)         # index o*256+c
            l4 = self.selu(torch.addmm(self.b4, l3, self.w4))
            probs, logits = [], []
            for k in range(4):
                a = self.selu(torch.addmm(self.b5[k], l4, self.w5[k]))
                z = self.selu(torch.addmm(self.bh[k], a, self.wh[k]))
                logits.append(z)
                probs.append(torch.softmax(z, dim=1))
            return torch.cat(probs, dim=1).cpu().numpy(), torch.cat(logits, dim=1).cpu().numpy()
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
