"""Parity of the CUDA path against the oracle, through the C-ABI (python -m pytest -m gpu).

Bar: |d| <= 1e-4 * max(1, |ref|) on the 90 post-SELU logits against the
fp64 oracle AND identical arg-max on each of the four heads.
"""
import ctypes

import numpy as np
import pytest

from clair_b200 import _lib, synth
from oracle import clair_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


def assert_parity(m, X, w, check_layers=False):
    n = X.shape[0]
    probs = m.predict(X)
    ref_probs, im = O.forward(X.astype(np.float32), w, np.float64, intermediates=True)
    logits = m.get_layer(_lib.LAYER_LOGITS, n)
    ref_logits = np.concatenate(im["logits"], axis=1)
    err = np.abs(logits - ref_logits) / np.maximum(1.0, np.abs(ref_logits))
    assert err.max() <= TOL, "scaled logit error %g" % err.max()
    assert [p.shape for p in probs] == [(n, 21), (n, 3), (n, 33), (n, 33)]
    for k in range(4):
        assert probs[k].dtype == np.float32
        np.testing.assert_array_equal(probs[k].argmax(1), ref_probs[k].argmax(1))
        assert np.abs(probs[k] - ref_probs[k]).max() <= TOL
        np.testing.assert_allclose(probs[k].sum(1), 1.0, atol=1e-5)
    if check_layers:
        for layer, key in ((_lib.LAYER_LSTM1, "lstm1"), (_lib.LAYER_LSTM2, "lstm2"), (_lib.LAYER_L3, "l3"),
                           (_lib.LAYER_L4, "l4")):
            got = m.get_layer(layer, n)
            ref = im[key]
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() <= TOL, key
    return err.max()


def test_golden_vectors(gpu_model, weights1234, golden_forward):
    g = golden_forward
    probs = gpu_model.predict(g["X"])
    logits = gpu_model.get_layer(_lib.LAYER_LOGITS, 8)
    err = np.abs(logits - g["logits"]) / np.maximum(1.0, np.abs(g["logits"]))
    assert err.max() <= TOL
    packed = np.concatenate(probs, axis=1)
    assert np.abs(packed - g["probs"]).max() <= TOL
    for a, b in ((0, 21), (21, 24), (24, 57), (57, 90)):
        np.testing.assert_array_equal(packed[:, a:b].argmax(1), g["probs"][:, a:b].argmax(1))
    for layer, key in ((_lib.LAYER_LSTM1, "lstm1"), (_lib.LAYER_LSTM2, "lstm2"), (_lib.LAYER_L3, "l3"), (_lib.LAYER_L4, "l4")):
        assert np.abs(gpu_model.get_layer(layer, 8) - g[key]).max() <= TOL, key


def test_config1_256_sites_every_layer(gpu_model, weights1234):
    X = synth.synthetic_tensors(256, seed=20240607)
    assert_parity(gpu_model, X, weights1234, check_layers=True)


@pytest.mark.parametrize("n", [1, 2, 127, 128, 129, 1000, 1001, 2500])
def test_ragged_batches(gpu_model, weights1234, n):
    # n=1: smallest batch; 1001/2500: several predict-batches with a ragged last one (utils.py:105)
    X = synth.synthetic_tensors(n, seed=1000 + n)
    assert_parity(gpu_model, X, weights1234)


def test_int16_transport_is_bit_identical(gpu_model):
    counts = synth.synthetic_counts(300, seed=5).astype(np.int32)
    counts[..., 1:] -= counts[..., 0:1]
    xi = counts.astype(np.int16)
    a = gpu_model.predict_packed(xi)
    b = gpu_model.predict_packed(xi.astype(np.float32))
    np.testing.assert_array_equal(a, b)


def test_input_forms_and_fresh_outputs(gpu_model):
    X = synth.synthetic_tensors(64, seed=9)
    a = gpu_model.predict(X)
    assert gpu_model.prediction is a
    b = gpu_model.predict(np.asfortranarray(X.astype(np.float64)))      # non-contiguous float64
    c = gpu_model.predict(X.reshape(64, -1))                             # flat rows
    for k in range(4):
        np.testing.assert_array_equal(a[k], b[k])
        np.testing.assert_array_equal(a[k], c[k])
        assert a[k] is not b[k] and not np.shares_memory(a[k], b[k])      # caller keeps the old ones
    with pytest.raises(ValueError):
        gpu_model.predict(np.zeros((4, 33, 8, 3), np.float32))
    with pytest.raises(ValueError):
        gpu_model.predict(np.zeros((0, 33, 8, 4), np.float32))


def test_deterministic_and_site_independent(gpu_model):
    X = synth.synthetic_tensors(1500, seed=21)
    a = gpu_model.predict_packed(X)
    b = gpu_model.predict_packed(X)
    np.testing.assert_array_equal(a, b)
    perm = np.random.default_rng(1).permutation(1500)
    c = gpu_model.predict_packed(X[perm])
    np.testing.assert_array_equal(a[perm], c)          # per-site: position in the batch is irrelevant


def test_multi_chunk_equals_single_calls(gpu_model):
    # 8192 sites = 9 predict-batches through the chunked, copy-overlapped host path
    X = synth.synthetic_tensors(8192, seed=33)
    full = gpu_model.predict_packed(X)
    for s in (0, 3000, 8000):
        part = gpu_model.predict_packed(X[s:s + 192])
        np.testing.assert_array_equal(full[s:s + 192], part)
    assert np.isfinite(full).all()


def test_chunk_pipeline_pageable_and_pinned_outputs(weights1234, monkeypatch):
    # small chunks force the copy-overlapped multi-chunk pipeline (3 chunks + a ragged 4th); results must not depend on
    # chunking nor on whether the destination is pageable (staged through pinned memory) or pinned (written directly)
    from clair_b200.model import Clair, pinned_empty, pinned_free
    monkeypatch.setenv("CLAIRB_CHUNK_SITES", "1024")
    m = Clair(max_sites=4096, batch_sites=1000)
    monkeypatch.delenv("CLAIRB_CHUNK_SITES")
    m.set_weights(weights1234)
    X = synth.synthetic_tensors(3300, seed=77)
    a = m.predict_packed(X)                                   # pageable destination
    ref = O.forward_packed(X[::13], weights1234, np.float64)
    assert np.abs(a[::13] - ref).max() <= TOL
    out = pinned_empty((3300, 90), np.float32)
    rc = m._lib.clairb_predict(m._h, X.ctypes.data_as(ctypes.c_void_p), _lib.DTYPE_F32, 3300,
                               out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    np.testing.assert_array_equal(a, out)
    one = Clair(max_sites=4096, batch_sites=1000)             # default chunking: a single chunk
    one.set_weights(weights1234)
    np.testing.assert_array_equal(a, one.predict_packed(X))
    pinned_free(out)
    m.close()
    one.close()


@pytest.mark.parametrize("env", [{"CLAIRB_ENGINE": "simt"}, {"CLAIRB_FUSED_TAIL": "0"}, {"CLAIRB_L2_STREAM": "0"},
                                 {"CLAIRB_L2_STREAM": "0", "CLAIRB_FUSED_TAIL": "0"}])
def test_cross_check_engines(weights1234, monkeypatch, env):
    # the CUDA-core fp32 engine and the mixed path (tensor-core LSTMs into the CUDA-core slice-dense / L4 / heads) are the
    # on-device cross-checks of the production tensor-core path: all three must meet the oracle and each other
    from clair_b200.model import Clair
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    with pytest.raises(ValueError, match="cross-check"):       # the product library carries one engine and says so
        Clair(max_sites=1024, batch_sites=1000)
    alt = Clair(max_sites=1024, batch_sites=1000, library=_lib.XCHECK_PATH)
    assert b"cross-check" in alt._lib.clairb_version()
    for k in env:
        monkeypatch.delenv(k)
    alt.set_weights(weights1234)
    X = synth.synthetic_tensors(300, seed=91)
    assert_parity(alt, X, weights1234, check_layers=True)
    prod = Clair(max_sites=1024, batch_sites=1000)
    prod.set_weights(weights1234)
    a, b = alt.predict_packed(X), prod.predict_packed(X)
    assert np.abs(a - b).max() <= 2e-5
    alt.close()
    prod.close()


def test_full_machine_stress_is_exact_and_repeatable(weights1234):
    # 40,000 sites fill every SM for several waves, which is where inter-CTA ordering bugs show (a relaxed mbarrier
    # arrive between the two CTAs of a pair once produced rare ~3e-3 errors that no small test saw): results must be
    # bit-identical between calls, identical to the same sites predicted alone, and meet the oracle on a sample
    from clair_b200.model import Clair
    m = Clair(max_sites=75000, batch_sites=1000)
    m.set_weights(weights1234)
    X = synth.synthetic_tensors(40000, seed=20240608)
    a = m.predict_packed(X)
    for _ in range(2):
        np.testing.assert_array_equal(a, m.predict_packed(X))
    idx = np.arange(0, 40000, 211)
    ref = O.forward_packed(X[idx], weights1234, np.float64)
    assert np.abs(a[idx] - ref).max() <= TOL
    for h, (lo, hi) in enumerate(((0, 21), (21, 24), (24, 57), (57, 90))):
        np.testing.assert_array_equal(a[idx, lo:hi].argmax(1), ref[:, lo:hi].argmax(1))
    np.testing.assert_array_equal(a[idx], m.predict_packed(X[idx]))     # position / chunk independence at scale
    m.close()


def test_predict_from_worker_thread(gpu_model):
    import threading
    X = synth.synthetic_tensors(100, seed=2)
    ref = gpu_model.predict_packed(X)
    t = threading.Thread(target=gpu_model.predict, kwargs={"batchX": X})     # call_var.py:1343
    t.start()
    t.join()
    np.testing.assert_array_equal(np.concatenate(gpu_model.prediction, 1), ref)


def test_extreme_inputs_stay_finite(gpu_model, weights1234):
    X = np.zeros((6, 33, 8, 4), np.float32)
    X[1] = 2000.0
    X[2] = -2000.0
    X[3, :, :, 0] = 30000.0
    X[4] = np.random.default_rng(0).normal(0, 3, X[4].shape)                 # non-integer input
    X[5] = 0.125
    assert_parity(gpu_model, X, weights1234)


def test_zero_bias_reference_initialisation():
    from clair_b200.model import Clair
    from clair_b200 import weights as W
    w = W.random_weights(seed=7, bias_std=0.0)                                # TF's own init has zero biases
    m = Clair(max_sites=512)
    m.set_weights(w)
    X = synth.synthetic_tensors(130, seed=8)
    assert_parity(m, X, w)
    m.close()


def test_pinned_staging_buffers(gpu_model):
    from clair_b200.model import pinned_empty, pinned_free
    X = synth.synthetic_tensors(500, seed=12)
    buf = pinned_empty(X.shape, np.float32)
    buf[...] = X
    np.testing.assert_array_equal(gpu_model.predict_packed(buf), gpu_model.predict_packed(X))
    pinned_free(buf)


def test_device_resident_entry_point(gpu_model):
    import torch
    X = synth.synthetic_tensors(2100, seed=14)
    ref = gpu_model.predict_packed(X)
    xd = torch.from_numpy(X).cuda()
    od = torch.empty((2100, 90), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream()
    before = gpu_model.kernel_launches()
    gpu_model.predict_device(xd.data_ptr(), _lib.DTYPE_F32, 2100, od.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert gpu_model.kernel_launches() > before
    np.testing.assert_array_equal(od.cpu().numpy(), ref)


def test_predict_before_weights_is_an_error():
    from clair_b200.model import Clair
    m = Clair(max_sites=256)
    with pytest.raises(RuntimeError):
        m.predict(np.zeros((1, 33, 8, 4), np.float32))
    lib = _lib.load()
    out = np.zeros((1, 90), np.float32)
    x = np.zeros((1, 33, 8, 4), np.float32)
    rc = lib.clairb_predict(m._h, x.ctypes.data_as(ctypes.c_void_p), 0, 1, out.ctypes.data_as(ctypes.c_void_p))
    assert rc == _lib.EINVAL and "finalize" in _lib.last_error(m._h)
    m.close()


def test_drop_in_batch_loop(gpu_model, weights1234):
    from clair_b200 import call_var
    sizes = [1000, 1000, 333]
    Xs = [synth.synthetic_tensors(s, seed=50 + i) for i, s in enumerate(sizes)]
    gen = ((X, [["c", str(j), "A" * 33] for j in range(len(X))]) for X in Xs)
    got = []
    call_var.run_batches(gpu_model, gen, lambda mb, Y: got.append((mb[0], [y.copy() for y in Y])))
    assert len(got) == 3
    for (X, Y), Xref in zip(got, Xs):
        assert X is Xref
        ref = O.forward(Xref, weights1234, np.float32)
        for k in range(4):
            np.testing.assert_array_equal(Y[k].argmax(1), ref[k].argmax(1))


@pytest.mark.parametrize("batch,n", [(4096, 2 * 4096 + 896), (8192, 8192 + 117)])
def test_configs_3_and_4_batch_sizes(weights1234, batch, n):
    # BASELINE.json configs[2] / [3]: predict-batches of 4096 (PacBio-CCS-scale run, ragged last batch of 896 as in
    # 30,000,000 = 7324 x 4096 + 896) and 8192 (Illumina-scale run); same architecture, only the batching differs
    from clair_b200.model import Clair
    m = Clair(max_sites=4 * batch, batch_sites=batch)
    m.set_weights(weights1234)
    X = synth.synthetic_tensors(n, seed=batch)
    out = m.predict_packed(X)
    idx = np.arange(0, n, 37)
    ref = O.forward_packed(X[idx], weights1234, np.float64)
    assert np.abs(out[idx] - ref).max() <= TOL
    for lo, hi in ((0, 21), (21, 24), (24, 57), (57, 90)):
        np.testing.assert_array_equal(out[idx, lo:hi].argmax(1), ref[:, lo:hi].argmax(1))
    # the same sites through single-batch calls give the same bits (tiles run through batch boundaries)
    np.testing.assert_array_equal(out[batch - 5:batch + 5], m.predict_packed(X[batch - 5:batch + 5]))
    m.close()


def test_full_size_one_million_sites_property(weights1234):
    # BASELINE.json configs[1] at full size: 1,000,000 sites in one call.  The sites are 20 shuffled copies of 50,000
    # distinct ones, so every distinct site is computed 20 times at unrelated positions (different chunks, waves, CTA
    # pairs, TMEM lanes): all copies must agree bit for bit, and a sample must meet the oracle.
    from clair_b200.model import Clair
    distinct, copies = 50000, 20
    base = synth.synthetic_tensors(distinct, seed=20240609)
    rng = np.random.default_rng(7)
    order = np.concatenate([rng.permutation(distinct) for _ in range(copies)])
    m = Clair(max_sites=distinct * copies, batch_sites=1000)
    m.set_weights(weights1234)
    out = m.predict_packed(base[order])
    assert out.shape == (distinct * copies, 90) and np.isfinite(out).all()
    first = np.empty((distinct, 90), np.float32)
    first[order[:distinct]] = out[:distinct]
    for k in range(1, copies):
        sl = slice(k * distinct, (k + 1) * distinct)
        np.testing.assert_array_equal(out[sl], first[order[sl]])
    idx = np.arange(0, distinct, 499)
    ref = O.forward_packed(base[idx], weights1234, np.float64)
    assert np.abs(first[idx] - ref).max() <= TOL
    np.testing.assert_allclose(out[:, :21].sum(1), 1.0, atol=1e-5)
    m.close()


def test_predict_split_layout_equals_packed(weights1234, monkeypatch):
    # predict() (four arrays, written head-major by the heads kernel) and predict_packed() ([n,90] rows) are the same
    # numbers, across several chunks, for a ragged size, and on the cross-check engines (host-side scatter)
    from clair_b200.model import Clair
    monkeypatch.setenv("CLAIRB_CHUNK_SITES", "1024")
    m = Clair(max_sites=4096, batch_sites=1000)
    monkeypatch.setenv("CLAIRB_FUSED_TAIL", "0")
    alt = Clair(max_sites=4096, batch_sites=1000, library=_lib.XCHECK_PATH)
    monkeypatch.delenv("CLAIRB_FUSED_TAIL")
    monkeypatch.delenv("CLAIRB_CHUNK_SITES")
    for eng in (m, alt):
        eng.set_weights(weights1234)
        for n in (1, 777, 3300):
            X = synth.synthetic_tensors(n, seed=600 + n)
            packed = eng.predict_packed(X)
            four = eng.predict(X)
            assert [a.shape for a in four] == [(n, 21), (n, 3), (n, 33), (n, 33)]
            assert all(a.flags["C_CONTIGUOUS"] and a.dtype == np.float32 for a in four)
            np.testing.assert_array_equal(np.concatenate(four, axis=1), packed)
        eng.close()


def test_cudnn_restatement_agrees_with_the_oracle(gpu_model, weights1234):
    # the reference's arithmetic lives in TensorFlow 1.13 (not installable), whose
    # CudnnCompatibleLSTMCell is defined to be weight-compatible with cuDNN's LSTM (clair/model.py:281-312).  cuDNN's own
    # fp32 LSTM (torch.nn.LSTM on cuda) + a torch dense trunk is a third implementation, written by nobody involved here:
    # it must agree with the fp64 numpy oracle to 1e-5 on config 1, and the product must meet both.
    from oracle.clair_oracle_cudnn import CudnnOracle
    X = synth.synthetic_tensors(256, seed=20240607)
    ref_probs, im = O.forward(X, weights1234, np.float64, intermediates=True)
    ref_logits = np.concatenate(im["logits"], axis=1)
    cp, cl = CudnnOracle(weights1234, device="cuda").forward(X)
    scaled = lambda a, b: (np.abs(a - b) / np.maximum(1.0, np.abs(b))).max()
    assert scaled(cl, ref_logits) <= 1e-5
    assert np.abs(cp - np.concatenate(ref_probs, axis=1)).max() <= 1e-5
    gpu_model.predict(X)
    ours = gpu_model.get_layer(_lib.LAYER_LOGITS, 256)
    assert scaled(ours, cl) <= TOL and scaled(ours, ref_logits) <= TOL
    for a, b in ((0, 21), (21, 24), (24, 57), (57, 90)):
        np.testing.assert_array_equal(cp[:, a:b].argmax(1), np.concatenate(ref_probs, axis=1)[:, a:b].argmax(1))
    try:
        import tensorflow  # noqa: F401
        print("tensorflow is importable on this box: the reference itself could be run")
    except Exception as exc:
        print("tensorflow not importable on the GPU box (%s): the oracle stays pinned on cuDNN + the two CPU restatements" % type(exc).__name__)


def test_inputs_that_are_not_exact_in_fp16_keep_their_low_parts(gpu_model, weights1234):
    # the layer-1 kernel skips the x_lo . W_hi tensor-core products when prep_tiles48 saw only fp16-exact inputs (the
    # generator's integer counts); fractional or large inputs must switch them back on, per chunk
    rng = np.random.default_rng(11)
    exact = synth.synthetic_tensors(300, seed=12)
    frac = (exact + rng.normal(0, 0.37, exact.shape)).astype(np.float32)         # 24-bit mantissas
    big = exact.copy()
    big[7, 16, 2, 0] = 4099.0                                                      # an integer fp16 cannot hold
    for X in (frac, big, exact, frac):                                             # the flag must not stick either way
        got = gpu_model.predict_packed(X)
        logits = gpu_model.get_layer(_lib.LAYER_LOGITS, len(X))
        ref_probs, im = O.forward(X, weights1234, np.float64, intermediates=True)
        ref_logits = np.concatenate(im["logits"], axis=1)
        assert (np.abs(logits - ref_logits) / np.maximum(1.0, np.abs(ref_logits))).max() <= TOL
        assert np.abs(got - np.concatenate(ref_probs, axis=1)).max() <= TOL
    i16 = np.clip(exact, -32768, 32767).astype(np.int16)
    i16[3, 10, 1, 0] = 30001                                                       # int16 transport, beyond fp16's integers
    got = gpu_model.predict_packed(i16)
    ref = O.forward_packed(i16.astype(np.float32), weights1234, np.float64)
    assert np.abs(got - ref).max() <= TOL


def test_device_forward_against_the_reference_model_code(gpu_model):
    # the reference's own clair/model.py run over the TensorFlow stand-in (tests/golden/reference_model_forward.npz, see
    # tests/test_oracle.py): the device's probabilities against its float64 evaluation
    import os
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_model_forward.npz")) as z:
        X, want64 = z["X"], z["probs_f64"]
    for Xin in (X.astype(np.float32), X):                         # float32 like the reference's generator, and the int16 transport
        got = gpu_model.predict_packed(Xin)
        assert np.abs(got - want64).max() <= TOL
        for a, b in ((0, 21), (21, 24), (24, 57), (57, 90)):
            np.testing.assert_array_equal(got[:, a:b].argmax(1), want64[:, a:b].argmax(1))
