"""The oracle against its golden vectors and against an independent restatement (CPU only)."""
import hashlib

import numpy as np
import pytest

from clair_b200 import synth, weights as W
from oracle import clair_oracle as O


def test_param_count_matches_reference_graph():
    # 2,377,818 trainable parameters on the forward path
    assert W.n_params() == 2377818
    shapes = W.weight_shapes()
    assert shapes[W.lstm_name(1, "fw", "kernel")] == (160, 512)
    assert shapes[W.lstm_name(2, "bw", "kernel")] == (384, 512)
    assert shapes["L4/kernel"] == (7680, 192)


def test_weights_are_reproducible(weights1234, golden_forward):
    h = hashlib.sha256()
    for k in sorted(weights1234):
        h.update(k.encode())
        h.update(np.ascontiguousarray(weights1234[k]).tobytes())
    assert h.hexdigest() == str(golden_forward["weights_sha256"])


def test_oracle_matches_golden_fp64(weights1234, golden_forward):
    g = golden_forward
    probs, im = O.forward(g["X"], weights1234, np.float64, intermediates=True)
    np.testing.assert_allclose(np.concatenate(probs, 1), g["probs"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(np.concatenate(im["logits"], 1), g["logits"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(im["l4"], g["l4"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(im["lstm1"], g["lstm1"], rtol=0, atol=1e-6)   # stored as float32
    np.testing.assert_allclose(im["lstm2"], g["lstm2"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(im["l3"], g["l3"], rtol=0, atol=1e-6)


def test_golden_inputs_cover_edge_cases(golden_forward):
    g = golden_forward
    np.testing.assert_array_equal(O.subtract_channel0(g["counts"]), g["X"])
    assert not g["X"][0].any()                       # zero-depth site
    assert g["counts"][1].max() >= 250               # deep pile-up
    assert g["X"].min() < 0                          # negative post-subtraction evidence
    assert np.isfinite(g["probs"]).all()
    np.testing.assert_allclose(g["probs"][:, :21].sum(1), 1.0, atol=1e-12)


def test_fp32_oracle_tracks_fp64(weights1234):
    X = synth.synthetic_tensors(64, seed=11)
    p64, i64 = O.forward(X, weights1234, np.float64, True)
    p32, i32 = O.forward(X, weights1234, np.float32, True)
    for k in range(4):
        assert np.abs(i64["logits"][k] - i32["logits"][k]).max() < 5e-6
        assert (p64[k].argmax(1) == p32[k].argmax(1)).all()


def _torch_bilstm(x_tm, w, layer, fin):
    """Independent restatement through torch.nn.LSTM (gate order i,f,g,o; rows [x;h] split)."""
    import torch
    m = torch.nn.LSTM(fin, 128, bidirectional=True).double()

    def reorder(t):
        i, c, f, o = t.chunk(4, dim=-1)               # TF LSTMBlockCell column blocks i, c, f, o
        return torch.cat([i, f, c, o], dim=-1)

    for d, suf in (("fw", ""), ("bw", "_reverse")):
        k = torch.tensor(w[W.lstm_name(layer, d, "kernel")], dtype=torch.float64)
        b = torch.tensor(w[W.lstm_name(layer, d, "bias")], dtype=torch.float64)
        getattr(m, "weight_ih_l0" + suf).data = reorder(k[:fin]).t().contiguous()
        getattr(m, "weight_hh_l0" + suf).data = reorder(k[fin:]).t().contiguous()
        getattr(m, "bias_ih_l0" + suf).data = reorder(b)
        getattr(m, "bias_hh_l0" + suf).data = torch.zeros(512, dtype=torch.float64)
    with torch.no_grad():
        return m(x_tm)[0]


def test_lstm_layers_match_torch_restatement(weights1234):
    import torch
    X = synth.synthetic_tensors(16, seed=3)
    _, im = O.forward(X, weights1234, np.float64, True)
    x_tm = torch.tensor(X, dtype=torch.float64).reshape(-1, 33, 32).transpose(0, 1).contiguous()
    l1 = _torch_bilstm(x_tm, weights1234, 1, 32)
    l2 = _torch_bilstm(l1, weights1234, 2, 256)
    assert np.abs(l1.numpy() - im["lstm1"]).max() < 1e-12
    assert np.abs(l2.numpy() - im["lstm2"]).max() < 1e-12


def test_dense_trunk_matches_torch_restatement(weights1234):
    import torch
    X = synth.synthetic_tensors(8, seed=4)
    probs, im = O.forward(X, weights1234, np.float64, True)
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    selu = torch.nn.SELU()
    lstm2 = t(im["lstm2"]).permute(1, 0, 2)                      # [n,33,256]
    cols = []
    for c in range(256):                                          # the 256 independent dense(33->30)
        cols.append(selu(lstm2[:, :, c] @ t(weights1234["L3/Unit_%d/kernel" % c]) + t(weights1234["L3/Unit_%d/bias" % c])))
    l3 = torch.stack(cols, dim=2)                                 # [n,30,256]
    assert np.abs(l3.numpy() - im["l3"]).max() < 1e-12
    l4 = selu(l3.reshape(8, -1) @ t(weights1234["L4/kernel"]) + t(weights1234["L4/bias"]))
    assert np.abs(l4.numpy() - im["l4"]).max() < 1e-12
    for k in range(4):
        a = selu(l4 @ t(weights1234["L5_%d/kernel" % (k + 1)]) + t(weights1234["L5_%d/bias" % (k + 1)]))
        z = selu(a @ t(weights1234["Prediction/%s/kernel" % O.HEAD_NAMES[k]]) + t(weights1234["Prediction/%s/bias" % O.HEAD_NAMES[k]]))
        p = torch.softmax(z, dim=1)
        assert np.abs(p.numpy() - probs[k]).max() < 1e-12


def test_selu_constants_and_shape():
    x = np.array([-3.0, -1e-3, 0.0, 1e-3, 2.5])
    y = O.selu(x)
    assert y[2] == 0.0 and y[4] == pytest.approx(2.5 * 1.0507009873554805)
    assert y[0] == pytest.approx(1.0507009873554805 * 1.6732632423543772 * (np.exp(-3.0) - 1))


def test_sites_are_independent(weights1234):
    # every op is per-site: permuting the batch permutes the result
    X = synth.synthetic_tensors(12, seed=8)
    perm = np.random.default_rng(0).permutation(12)
    a = O.forward_packed(X, weights1234, np.float64)
    b = O.forward_packed(X[perm], weights1234, np.float64)
    np.testing.assert_allclose(a[perm], b, rtol=0, atol=1e-13)


def test_fast_cpu_port_matches_oracle(weights1234):
    from oracle.clair_oracle_fast import FastOracle
    X = synth.synthetic_tensors(40, seed=6)
    fast = FastOracle(weights1234).forward_packed(X)
    ref = O.forward_packed(X, weights1234, np.float64)
    assert fast.shape == (40, 90) and fast.dtype == np.float32
    assert np.abs(fast - ref).max() < 5e-6
    for a, b in ((0, 21), (21, 24), (24, 57), (57, 90)):
        np.testing.assert_array_equal(fast[:, a:b].argmax(1), ref[:, a:b].argmax(1))


def test_torch_lstm_restatement_agrees_with_the_numpy_oracle(weights1234):
    # the cuDNN oracle's weight mapping (gate order i,c,f,o -> i,f,g,o; rows [x;h] -> weight_ih / weight_hh), checked
    # here on torch's CPU LSTM kernels: in float64 it must reproduce the numpy fp64 oracle to rounding
    import torch
    from clair_b200 import synth
    from oracle.clair_oracle_cudnn import CudnnOracle
    X = synth.synthetic_tensors(24, seed=5)
    ref_probs, im = O.forward(X, weights1234, np.float64, intermediates=True)
    p, l = CudnnOracle(weights1234, device="cpu", dtype=torch.float64).forward(X)
    assert np.abs(p - np.concatenate(ref_probs, axis=1)).max() <= 1e-12
    assert np.abs(l - np.concatenate(im["logits"], axis=1)).max() <= 1e-11
    p32, _ = CudnnOracle(weights1234, device="cpu", dtype=torch.float32).forward(X)
    assert np.abs(p32 - np.concatenate(ref_probs, axis=1)).max() <= 1e-5


def test_oracle_reproduces_the_reference_model_code_run_over_the_tf_standin(weights1234):
    # tests/golden/reference_model_forward.npz: the REFERENCE's unmodified clair/model.py (Clair() -> init -> restore_parameters
    # -> predict) executed over the numpy stand-in for TensorFlow (oracle/tf_standin, oracle/gen_golden_reference_model.py).
    # Everything above the TF ops - reshape / transpose order, slice-dense axes, flatten order, activations, variable
    # scopes, output order - is the reference's own code there; the restatement must land on the same numbers.
    import os
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_model_forward.npz")) as z:
        X, want32, want64, names = z["X"].astype(np.float32), z["probs_f32"], z["probs_f64"], z["variable_names"].tolist()
    assert names == sorted(W.weight_shapes())                      # the names the reference's scopes produce = ours
    got64 = O.forward_packed(X, weights1234, np.float64)
    assert np.abs(got64 - want64).max() <= 1e-12
    got32 = O.forward_packed(X, weights1234, np.float32)
    assert np.abs(got32 - want32).max() <= 2e-6
    for a, b in ((0, 21), (21, 24), (24, 57), (57, 90)):
        np.testing.assert_array_equal(got64[:, a:b].argmax(1), want64[:, a:b].argmax(1))
