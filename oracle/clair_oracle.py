"""CPU oracle for Clair's batched variant-calling forward path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU *restatement* (numpy) of what the reference computes when
``Clair.predict`` runs ``session.run(self.Y)`` (reference clair/model.py:946-966) on the
"2BiLSTM" graph (clair/model.py:400-622).  It exists to check the CUDA path; nothing in the
product package ``clair_b200`` imports it.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` leg may import it.

PARITY: pinned on the reference's own model code, NOT on TensorFlow's kernels.  The arithmetic of the
reference lives in TensorFlow 1.13.2 (pinned only in prose, reference README.md:127), which is neither
vendored under /root/reference nor installable here (no wheel for Python 3.12, no network), and the
reference ships no tests, golden vectors, fixtures or checkpoints.  What pins this restatement:
  (1) tests/golden/reference_model_forward.npz - the reference's UNMODIFIED clair/model.py (Clair() ->
      init -> restore_parameters -> predict) executed over a numpy stand-in for the ~60 TensorFlow
      symbols it touches (oracle/tf_standin/, oracle/gen_golden_reference_model.py).  Graph wiring,
      axis conventions, activations, flatten order, variable scopes / names and the output order are
      the reference's own code there; this file reproduces its float64 evaluation to 1e-12
      (tests/test_oracle.py).  The stand-in's op semantics - notably the LSTMBlockCell arithmetic
      behind CudnnCompatibleLSTMCell - are TensorFlow's documented ones, restated; so
  (2) on the GPU box, cuDNN's own fp32 LSTM (the kernel CudnnCompatibleLSTMCell is defined to be
      weight-compatible with) + a torch dense trunk must agree with this file to 1e-5
      (oracle/clair_oracle_cudnn.py, tests/test_gpu_parity.py, bench.py `oracle_pinning`), and
  (3) an independent torch.nn.LSTM CPU restatement agrees to 1e-6 (oracle/clair_oracle_fast.py).
"Parity unpinned" in the strict sense - no output of TensorFlow itself has ever been compared - still
applies to the op kernels; DESIGN.md section 5 says so.

Semantics followed (reference file:line -> what it means here)
  clair/utils.py:96-98      channels 1..3 -= channel 0 happens in the generator, NOT in predict
  clair/model.py:403-418    [B,33,8,4] -> [B,33,32] (feature = row*4+channel) -> time-major
  clair/model.py:299-312    CudnnCompatibleLSTMCell(128) == LSTMBlockCell(forget_bias=0,
                            no clip, no peephole): kernel [(in+128), 512], rows = [x ; h],
                            column blocks in gate order i, c(candidate), f, o; bias [512];
                            zero initial state; c' = tanh(c)*sig(i) + c_prev*sig(f);
                            h = tanh(c')*sig(o)
  clair/model.py:306-312    stack_bidirectional_dynamic_rnn: bw consumes t=32..0 and its
                            outputs are re-reversed; out[t] = concat(h_fw[t], h_bw[t])
  clair/model.py:434-459    dropouts are identity at inference (training=False)
  clair/model.py:461        transpose back to [B,33,256]
  clair/model.py:225-244,464-471   slice_dense_layer over axis 2: 256 independent dense(33->30)+SELU
  clair/model.py:474-478    row-major flatten [B,30,256] -> [B,7680], index = o*256 + c
  clair/model.py:482-488    L4 dense 7680->192 + SELU ; selu.dropout_selu identity (selu.py:72-74)
  clair/model.py:507-578    L5_1..4 dense 192->96 + SELU
  clair/model.py:581-622    heads: dense(96->n_k) + SELU, then softmax over the SELU'd values
  clair/selu.py:26-30       selu(x) = scale * where(x>=0, x, alpha*(exp(x)-1))
  clair/task/main.py:10-29, clair/task/variant_length.py:6-12   head sizes 21 / 3 / 33 / 33
"""
import numpy as np

# clair/selu.py:28-29 (constants to full printed precision)
SELU_ALPHA = 1.6732632423543772848170429916717
SELU_SCALE = 1.0507009873554804934193349852946

T = 33            # shared/param.py:9   2*flankingBaseNum+1
ROWS = 8          # shared/param.py:10  matrixRow
CHANNELS = 4      # shared/param.py:11  matrixNum
F = ROWS * CHANNELS
H = 128           # clair/model.py:92-93 LSTM{1,2}_num_units
L3_UNITS = 30     # clair/model.py:80 (L2_num_units is what slice_dense is called with, :466)
L4_UNITS = 192    # clair/model.py:81
L5_UNITS = 96     # clair/model.py:83-90
HEADS = (21, 3, 33, 33)   # clair/task/main.py:10-29

LSTM_NAME = "{layer}/stack_bidirectional_rnn/cell_0/bidirectional_rnn/{d}/cudnn_compatible_lstm_cell/{v}"
HEAD_NAMES = ("Y_base_change_logits", "Y_genotype_logits",
              "Y_indel_length_logits_1", "Y_indel_length_logits_2")


def selu(x):
    # clair/selu.py:26-30 ; tf.nn.elu(x) = exp(x)-1 for x<0
    neg = SELU_ALPHA * np.expm1(np.minimum(x, 0))
    return (SELU_SCALE * np.where(x >= 0.0, x, neg)).astype(x.dtype)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def softmax(x):
    # tf.nn.softmax over the last axis (max-subtracted)
    z = x - x.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


def lstm_direction(x_tm, kernel, bias, reverse):
    """One direction of one bidirectional layer.  x_tm: [T,B,Fin] time-major.

    LSTMBlockCell semantics (reference clair/model.py:299-300); returns [T,B,H] aligned with
    the input time axis (backward outputs re-reversed, clair/model.py:306-312).
    """
    Tn, B, Fin = x_tm.shape
    Hn = kernel.shape[1] // 4
    Wx, Wh = kernel[:Fin], kernel[Fin:]
    xs = x_tm[::-1] if reverse else x_tm
    pre = xs.reshape(Tn * B, Fin) @ Wx + bias
    pre = pre.reshape(Tn, B, 4 * Hn)
    h = np.zeros((B, Hn), dtype=x_tm.dtype)
    c = np.zeros((B, Hn), dtype=x_tm.dtype)
    out = np.empty((Tn, B, Hn), dtype=x_tm.dtype)
    for t in range(Tn):
        z = pre[t] + h @ Wh
        i, g, f, o = z[:, :Hn], z[:, Hn:2 * Hn], z[:, 2 * Hn:3 * Hn], z[:, 3 * Hn:]
        c = np.tanh(g) * sigmoid(i) + c * sigmoid(f)
        h = np.tanh(c) * sigmoid(o)
        out[t] = h
    return out[::-1] if reverse else out


def bilstm(x_tm, w, layer):
    fw = lstm_direction(x_tm, w[LSTM_NAME.format(layer=layer, d="fw", v="kernel")],
                        w[LSTM_NAME.format(layer=layer, d="fw", v="bias")], False)
    bw = lstm_direction(x_tm, w[LSTM_NAME.format(layer=layer, d="bw", v="kernel")],
                        w[LSTM_NAME.format(layer=layer, d="bw", v="bias")], True)
    return np.concatenate([fw, bw], axis=2)


def stack_l3(w):
    """L3/Unit_c/{kernel[33,30],bias[30]} for c=0..255 -> ([256,33,30], [256,30])."""
    k = np.stack([w["L3/Unit_%d/kernel" % c] for c in range(2 * H)])
    b = np.stack([w["L3/Unit_%d/bias" % c] for c in range(2 * H)])
    return k, b


def forward(X, weights, dtype=np.float64, intermediates=False):
    """X: [n,33,8,4] (already channel-subtracted).  Returns list of 4 probability arrays.

    With intermediates=True returns (probs, dict of every layer output in `dtype`).
    """
    w = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
    X = np.asarray(X, dtype=dtype)
    n = X.shape[0]
    x2d = X.reshape(n, T, F)                                  # model.py:403-411
    x_tm = np.ascontiguousarray(x2d.transpose(1, 0, 2))       # model.py:416-418
    lstm1 = bilstm(x_tm, w, "LSTM1")                          # model.py:423-430
    lstm2 = bilstm(lstm1, w, "LSTM2")                         # model.py:443-450
    lstm2_bt = lstm2.transpose(1, 0, 2)                       # model.py:461  [n,33,256]
    k3, b3 = stack_l3(w)
    l3 = selu(np.einsum("btc,cto->boc", lstm2_bt, k3) + b3.T[None])   # model.py:464-471 [n,30,256]
    l3_flat = l3.reshape(n, L3_UNITS * 2 * H)                 # model.py:474-478
    l4 = selu(l3_flat @ w["L4/kernel"] + w["L4/bias"])        # model.py:482-488
    l5, logits, probs = [], [], []
    for k in range(4):
        a = selu(l4 @ w["L5_%d/kernel" % (k + 1)] + w["L5_%d/bias" % (k + 1)])   # model.py:507-578
        z = selu(a @ w["Prediction/%s/kernel" % HEAD_NAMES[k]] +
                 w["Prediction/%s/bias" % HEAD_NAMES[k]])                         # model.py:581-618
        l5.append(a)
        logits.append(z)
        probs.append(softmax(z))                                                  # model.py:589-619
    if not intermediates:
        return probs
    return probs, dict(lstm1=lstm1, lstm2=lstm2, l3=l3, l4=l4, l5=l5, logits=logits)


def forward_packed(X, weights, dtype=np.float32):
    """[n,90] = concat of the four heads (21+3+33+33), the layout the C-ABI returns."""
    return np.concatenate(forward(X, weights, dtype=dtype), axis=1)


def subtract_channel0(X):
    """The generator-side transform, clair/utils.py:96-98 (in place on a copy)."""
    X = np.array(X, dtype=np.float32, copy=True)
    for i in range(1, CHANNELS):
        X[:, :, :, i] -= X[:, :, :, 0]
    return X
