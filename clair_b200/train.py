"""Training step on the device: the training-side members of the reference's model object.

Mirrors what `Clair.train(batchX, batchY)` does per call (reference clair/model.py:913-945 with the graph of :400-740):
forward in training phase (dropout behind LSTM2, alpha-dropout behind L4 / L5_k), focal loss of the four heads + L2,
gradients of every trainable variable, clip_by_global_norm(5.0), Adam.  Underneath: `clairb_trainer_*` of the C-ABI
library (csrc/train_kernels.cuh: fp32 results; sequence kernels on thread-block clusters, 3xTF32 tensor-core GEMMs).  `Trainer`
drives one GPU; `DataParallelTrainer` is the config-5 shape (one process per GPU, NCCL all-reduce of the flat gradient buffer in
two pieces that overlap the backward pass, ordered stream to stream).
The training loop around it (learning-rate schedule, validation, checkpoints: clair/train.py:78-263) stays with the caller.
"""
import ctypes

import numpy as np

from . import _lib, param, weights as _weights

DROPOUTS = ("lstm2", "l4", "l5_1", "l5_2", "l5_3", "l5_4")
DEFAULT_RATES = (0.5, 0.5, 0.2, 0.2, 0.2, 0.2)                    # clair/model.py:83-97 (LSTM1_dropout_rate is 0)
CLIP_NORM = 5.0                                                    # clair/model.py:727


class Trainer(object):
    def __init__(self, device=0, max_batch=512, learning_rate=param.initialLearningRate, l2_lambda=param.l2RegularizationLambda,
                 dropout_rates=DEFAULT_RATES, seed=0, library=None):
        self._lib = _lib.load(library)
        handle = ctypes.c_void_p()
        rc = self._lib.clairb_trainer_create(int(device), int(max_batch), ctypes.byref(handle))
        self._t = None
        self._check(rc, "clairb_trainer_create")
        self._t = handle
        self.device, self.max_batch = int(device), int(max_batch)
        self.learning_rate_value, self.l2_regularization_lambda_value = float(learning_rate), float(l2_lambda)
        self.seed, self.step = int(seed), 0
        self.num_params = int(self._lib.clairb_trainer_num_params(self._t))
        self.dense_offset = int(self._lib.clairb_trainer_dense_offset(self._t))
        self.set_dropout_rates(dropout_rates)
        self.training_loss_on_one_batch = None
        self.loss_parts = None

    def _check(self, rc, what):
        if rc:
            msg = self._lib.clairb_trainer_last_error(self._t)
            msg = "%s: %s" % (what, msg.decode("utf-8", "replace") if msg else "")
            raise (ValueError if rc in (_lib.EINVAL, _lib.EWEIGHTS) else MemoryError if rc == _lib.ENOMEM else RuntimeError)(msg)

    # ---- parameters ------------------------------------------------------------------------------------------------
    def set_weights(self, weights):
        _weights.check_weights(weights)
        for name in _weights.weight_shapes():
            arr = np.ascontiguousarray(weights[name], dtype=np.float32)
            shape = (ctypes.c_int64 * arr.ndim)(*arr.shape)
            self._check(self._lib.clairb_trainer_set_weight(self._t, name.encode(), arr.ctypes.data_as(ctypes.c_void_p), shape, arr.ndim),
                        "clairb_trainer_set_weight(%s)" % name)

    def init(self):
        """Reference: run the TF initialisers (clair/model.py:807-813)."""
        self.set_weights(_weights.random_weights(seed=self.seed, bias_std=0.0))

    def _get(self, which):
        out = {}
        for name, shape in _weights.weight_shapes().items():
            arr = np.empty(shape, np.float32)
            self._check(self._lib.clairb_trainer_get(self._t, which, name.encode(), arr.ctypes.data_as(ctypes.c_void_p), arr.size),
                        "clairb_trainer_get(%s)" % name)
            out[name] = arr
        return out

    def get_weights(self):
        """{TF variable name: float32 array} - what save_parameters would write (clair/model.py:1010-1014); feed it to
        clair_b200.weights.save_blob or straight to Clair.set_weights for inference."""
        return self._get(0)

    def gradients(self):
        return self._get(1)

    def set_dropout_rates(self, rates):
        r = (ctypes.c_float * 6)(*[float(x) for x in rates])
        self._check(self._lib.clairb_trainer_set_dropout_rates(self._t, r), "clairb_trainer_set_dropout_rates")
        self.dropout_rates = tuple(float(x) for x in rates)

    # ---- one step, in its three parts ------------------------------------------------------------------------------------
    def _inputs(self, batchX, batchY, masks):
        X = np.asarray(batchX)
        if X.ndim == 2:
            X = X.reshape(-1, 33, 8, 4)
        n = X.shape[0]
        if X.shape[1:] != (33, 8, 4) or not 1 <= n <= self.max_batch:
            raise ValueError("batchX must be [n,33,8,4] with 1 <= n <= %d" % self.max_batch)
        if X.dtype == np.int16:
            X, dtype = np.ascontiguousarray(X), _lib.DTYPE_I16
        else:
            X, dtype = np.ascontiguousarray(X, dtype=np.float32), _lib.DTYPE_F32
        Y = np.ascontiguousarray(batchY, dtype=np.float32)
        if Y.shape != (n, _lib.N_OUT):
            raise ValueError("batchY must be [n,90]")
        ptrs = None
        keep = [X, Y]
        if masks is not None:
            shapes = {"lstm2": (33, n, 256), "l4": (n, 192), "l5_1": (n, 96), "l5_2": (n, 96), "l5_3": (n, 96), "l5_4": (n, 96)}
            arr = (ctypes.c_void_p * 6)()
            for i, name in enumerate(DROPOUTS):
                m = np.ascontiguousarray(masks[name], dtype=np.uint8)
                if m.shape != shapes[name]:
                    raise ValueError("mask %s must have shape %s" % (name, shapes[name]))
                keep.append(m)
                arr[i] = m.ctypes.data
            ptrs = arr
        self.step += 1
        self._n = n
        seed = ctypes.c_uint64((self.seed * 1000003 + self.step) & (2 ** 64 - 1))
        return X.ctypes.data_as(ctypes.c_void_p), dtype, Y.ctypes.data_as(ctypes.c_void_p), n, ptrs, seed, keep

    def forward_backward(self, batchX, batchY, masks=None):
        """Forward, loss and the backward pass through the dense layers -> [gt21, genotype, length 1, length 2] focal-loss sums
        and the L2 sum without lambda.  masks: {name in DROPOUTS: uint8 keep-mask} (parity tests) or None (drawn on the device)."""
        x, dtype, y, n, ptrs, seed, keep = self._inputs(batchX, batchY, masks)
        losses = (ctypes.c_double * 5)()
        self._check(self._lib.clairb_trainer_forward_backward(self._t, x, dtype, y, n, ptrs, seed, losses), "clairb_trainer_forward_backward")
        self.loss_parts = [float(v) for v in losses]
        return self.loss_parts

    def backward_lstm(self):
        self._check(self._lib.clairb_trainer_backward_lstm(self._t), "clairb_trainer_backward_lstm")

    def apply(self):
        """L2 gradient, clip, Adam -> the global gradient norm before clipping."""
        norm = ctypes.c_double()
        rc = self._lib.clairb_trainer_apply(self._t, self.learning_rate_value, self.l2_regularization_lambda_value, CLIP_NORM, self.step,
                                            ctypes.byref(norm))
        self._check(rc, "clairb_trainer_apply")
        return norm.value

    def total_loss(self, parts=None):
        """total_loss of clair/model.py:696-709 with the default task weights (all 1)."""
        p = self.loss_parts if parts is None else parts
        return p[0] + p[1] + p[2] + p[3] + self.l2_regularization_lambda_value * p[4]

    def train(self, batchX, batchY, masks=None):
        """Reference clair/model.py:913-945: one optimisation step -> the training loss of the batch.  One library call
        (clairb_trainer_step: forward_backward, backward_lstm and apply with a single synchronisation)."""
        x, dtype, y, n, ptrs, seed, keep = self._inputs(batchX, batchY, masks)
        losses, norm = (ctypes.c_double * 5)(), ctypes.c_double()
        rc = self._lib.clairb_trainer_step(self._t, x, dtype, y, n, ptrs, seed, self.learning_rate_value, self.l2_regularization_lambda_value,
                                           CLIP_NORM, self.step, losses, ctypes.byref(norm))
        self._check(rc, "clairb_trainer_step")
        self.loss_parts = [float(v) for v in losses]
        self.grad_norm = norm.value
        self.training_loss_on_one_batch = self.total_loss()
        return self.training_loss_on_one_batch

    def probabilities(self):
        out = np.empty((self._n, _lib.N_OUT), np.float32)
        self._check(self._lib.clairb_trainer_get_probabilities(self._t, out.ctypes.data_as(ctypes.c_void_p), self._n), "clairb_trainer_get_probabilities")
        return out

    def kernel_launches(self):
        return int(self._lib.clairb_trainer_kernel_launches(self._t))

    def close(self):
        t, self._t = getattr(self, "_t", None), None
        if t:
            self._lib.clairb_trainer_destroy(t)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GradientExchange(object):
    """The all-reduce of one flat gradient buffer in two pieces: `dense()` starts the tail [dense_offset:] - complete once the
    backward pass has left the dense layers -, `head()` the LSTM part [:dense_offset] after BPTT, `wait()` returns when both
    sums are in place on every rank.  SUM, not mean: the reference's loss is a sum over the batch (clair/model.py:696-709, 783-805),
    so the gradient of the global batch is the sum of the ranks' gradients.  Works on any torch.distributed backend (the CPU
    tests run it over gloo).  On CUDA, `compute` is the stream the gradients are produced on and `comm` the communication stream:
    a piece starts when `compute` has reached this point, and wait() makes `compute` (not the host) wait for the sums."""

    def __init__(self, grad, dense_offset, dist, compute=None, comm=None):
        self.grad, self.dense_offset, self.dist, self.compute, self.comm = grad, int(dense_offset), dist, compute, comm
        self._work = []

    def _start(self, piece):
        if self.comm is not None:
            import torch
            self.comm.wait_stream(self.compute)
            with torch.cuda.stream(self.comm):
                self._work.append(self.dist.all_reduce(piece, op=self.dist.ReduceOp.SUM, async_op=True))
        else:
            self._work.append(self.dist.all_reduce(piece, op=self.dist.ReduceOp.SUM, async_op=True))

    def dense(self):
        self._start(self.grad[self.dense_offset:])

    def head(self):
        self._start(self.grad[:self.dense_offset])

    def wait(self):
        if self.comm is not None:
            import torch
            with torch.cuda.stream(self.compute):
                for w in self._work:
                    w.wait()                                         # NCCL: blocks the current STREAM, not the host
        else:
            for w in self._work:
                w.wait()
        self._work = []


class DataParallelTrainer(Trainer):
    """One process per GPU (torch.distributed, NCCL): every rank runs the step on its own batch.  The flat gradient buffer is a
    torch tensor with four extra floats behind it that receive the rank's focal-loss sums; its dense tail (and the loss sums) is
    all-reduced while the LSTM backward runs, the LSTM head after it (GradientExchange), then every rank applies the same update.
    The parts of the step are only enqueued (deferred mode): the library's stream waits for the collectives, the host
    synchronises once, in apply().  Weights must start identical on all ranks (set_weights with the same blob / seed)."""

    def __init__(self, **kw):
        import torch
        import torch.distributed as dist
        self._dist, self._torch = dist, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        Trainer.__init__(self, **kw)
        self._buf = torch.zeros(self.num_params + 4, dtype=torch.float32, device="cuda:%d" % self.device)
        self._grad = self._buf[:self.num_params]
        self._check(self._lib.clairb_trainer_set_grad_buffer(self._t, ctypes.c_void_p(self._buf.data_ptr())), "clairb_trainer_set_grad_buffer")
        tail = ctypes.c_void_p(self._buf.data_ptr() + 4 * self.num_params)
        self._check(self._lib.clairb_trainer_set_deferred(self._t, 1, tail), "clairb_trainer_set_deferred")
        compute = torch.cuda.ExternalStream(self._lib.clairb_trainer_stream(self._t), device=self.device)
        self._exchange = GradientExchange(self._buf, self.dense_offset, dist, compute, torch.cuda.Stream(device=self.device))

    def train(self, batchX, batchY, masks=None):
        x, dtype, y, n, ptrs, seed, keep = self._inputs(batchX, batchY, masks)
        self._check(self._lib.clairb_trainer_forward_backward(self._t, x, dtype, y, n, ptrs, seed, None), "clairb_trainer_forward_backward")
        self._exchange.dense()                                        # 8.3 MB + the loss sums travel while BPTT runs
        self.backward_lstm()
        self._exchange.head()
        self._exchange.wait()
        self.grad_norm = self.apply()                                 # the one host synchronisation of the step
        own = (ctypes.c_double * 5)()
        self._check(self._lib.clairb_trainer_read_losses(self._t, own), "clairb_trainer_read_losses")
        self.loss_parts = [float(v) for v in self._buf[self.num_params:].tolist()] + [float(own[4])]
        self.training_loss_on_one_batch = self.total_loss()
        return self.training_loss_on_one_batch
