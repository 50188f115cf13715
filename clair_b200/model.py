"""Drop-in for the reference's model object on the forward path.

Mirrors ``clair.model.Clair`` (reference clair/model.py): ``Clair(**kwargs)`` (:58),
``init()`` (:807), ``restore_parameters(file_name)`` (:1016), ``predict(batchX)`` and the
``.prediction`` side effect (:946-966), ``close()`` (:872), attributes ``input_shape`` and
``output_label_split`` (:119,173-178).  Everything underneath is the C-ABI library
(include/clair_b200.h) running hand-written sm_100a CUDA; training-side members
(train/validate/save_parameters, loss, optimiser) are out of scope.
"""
import ctypes
import threading

import numpy as np

from . import _lib, param, weights as _weights

GT21_LABELS, GENOTYPE_LABELS, VARLEN_LABELS = 21, 3, 33   # clair/task/main.py:10-29


class Clair(object):
    def __init__(self, **kwargs):
        params = dict(                                         # clair/model.py:61-105 (forward-path keys)
            input_shape=(2 * param.flankingBaseNum + 1, param.matrixRow, param.matrixNum),
            structure="2BiLSTM",
            output_gt21_shape=GT21_LABELS,
            output_genotype_shape=GENOTYPE_LABELS,
            output_indel_length_shape_1=VARLEN_LABELS,
            output_indel_length_shape_2=VARLEN_LABELS,
            tensor_transform_function=lambda X, Y, phase: (X, Y),
            # B200-side options (safe defaults; not in the reference)
            device=0,
            max_sites=32 * param.predictBatchSize,
            batch_sites=param.predictBatchSize,
            seed=None,
        )
        params.update(param.get_model_parameters())            # clair/model.py:108-109
        for key, value in kwargs.items():                      # clair/model.py:112-116
            if key in params:
                params[key] = value
            else:
                print("Info: the parameter %s, with value %s is not supported" % (key, value))
        if params["structure"] != "2BiLSTM":                    # clair/model.py:400 is the only branch
            raise ValueError("structure %r is not supported (reference only builds '2BiLSTM')" % params["structure"])
        self.input_shape = tuple(params["input_shape"])
        if self.input_shape != (33, 8, 4):
            raise ValueError("input_shape must be (33, 8, 4)")
        self.tensor_transform_function = params["tensor_transform_function"]
        self.output_gt21_shape = params["output_gt21_shape"]
        self.output_genotype_shape = params["output_genotype_shape"]
        self.output_indel_length_shape_1 = params["output_indel_length_shape_1"]
        self.output_indel_length_shape_2 = params["output_indel_length_shape_2"]
        self.output_label_split = [                            # clair/model.py:173-178
            self.output_gt21_shape, self.output_genotype_shape,
            self.output_indel_length_shape_1, self.output_indel_length_shape_2,
        ]
        if self.output_label_split != [21, 3, 33, 33]:
            raise ValueError("output shapes must be 21/3/33/33")
        self.structure = params["structure"]
        self.device = int(params["device"])
        self.max_sites = int(params["max_sites"])
        self.batch_sites = int(params["batch_sites"])
        self._seed = params["seed"]
        self.prediction = None
        self._lock = threading.Lock()
        self._lib = _lib.load()                                # raises if the extension is missing
        handle = ctypes.c_void_p()
        rc = self._lib.clairb_create(self.device, self.max_sites, self.batch_sites, ctypes.byref(handle))
        _lib.check(rc, None, "clairb_create")
        self._h = handle
        self._has_weights = False

    # ---- weights ---------------------------------------------------------------------------
    def init(self):
        """Reference: run the TF initialisers (clair/model.py:807-813)."""
        self.set_weights(_weights.random_weights(seed=self._seed, bias_std=0.0))

    def restore_parameters(self, file_name):
        """Reference: tf.train.Saver.restore (clair/model.py:1016-1020).  `file_name` is either the prefix of a
        TensorFlow V2 checkpoint bundle (<prefix>.index + <prefix>.data-*, read without TensorFlow by
        clair_b200/checkpoint.py) or the .npz weight blob keyed by TF variable name (clair_b200/weights.py)."""
        from . import checkpoint as _checkpoint
        if _checkpoint.is_checkpoint_prefix(file_name):
            self.set_weights(_checkpoint.load_checkpoint(file_name))
        else:
            self.set_weights(_weights.load_blob(file_name))

    def set_weights(self, weights):
        _weights.check_weights(weights)
        for name in _weights.weight_shapes():
            arr = np.ascontiguousarray(weights[name], dtype=np.float32)
            shape = (ctypes.c_int64 * arr.ndim)(*arr.shape)
            rc = self._lib.clairb_set_weight(self._h, name.encode(), arr.ctypes.data_as(ctypes.c_void_p),
                                             shape, arr.ndim)
            _lib.check(rc, self._h, "clairb_set_weight(%s)" % name)
        _lib.check(self._lib.clairb_finalize_weights(self._h), self._h, "clairb_finalize_weights")
        self._has_weights = True

    # ---- forward ---------------------------------------------------------------------------
    def _as_input(self, batchX):
        X = np.asarray(batchX)
        if X.ndim == 2 and X.shape[1] == _lib.SITE_ELEMS:
            X = X.reshape((-1,) + self.input_shape)
        if X.ndim != 4 or tuple(X.shape[1:]) != self.input_shape:
            raise ValueError("Inconsistent shape: expected [n,33,8,4], got %s" % (X.shape,))
        if X.shape[0] < 1:
            raise ValueError("empty batch")
        if X.dtype == np.int16:
            return np.ascontiguousarray(X), _lib.DTYPE_I16
        return np.ascontiguousarray(X, dtype=np.float32), _lib.DTYPE_F32

    def predict_packed(self, batchX):
        """[n,90] float32: the four heads' probabilities side by side (21|3|33|33)."""
        if not self._has_weights:
            raise RuntimeError("predict() before init()/restore_parameters()")
        X, _ = self.tensor_transform_function(batchX, None, "predict")      # clair/model.py:953
        X, dtype = self._as_input(X)
        n = X.shape[0]
        out = np.empty((n, _lib.N_OUT), dtype=np.float32)                    # fresh every call
        with self._lock:
            for s in range(0, n, self.max_sites):
                m = min(self.max_sites, n - s)
                rc = self._lib.clairb_predict(self._h, X[s:s + m].ctypes.data_as(ctypes.c_void_p), dtype, m,
                                              out[s:s + m].ctypes.data_as(ctypes.c_void_p))
                _lib.check(rc, self._h, "clairb_predict")
        return out

    def predict(self, batchX):
        """Reference clair/model.py:946-966: list of 4 float32 arrays, also stored in .prediction."""
        if not self._has_weights:
            raise RuntimeError("predict() before init()/restore_parameters()")
        from .create_tensor import DeviceTensors
        if isinstance(batchX, DeviceTensors):
            # tensors that clairb_create_tensors left on the device: no host copy of X exists
            packed = batchX.block.predict(batchX.index)
            bounds = np.cumsum([0] + self.output_label_split)
            prediction = [np.ascontiguousarray(packed[:, a:b]) for a, b in zip(bounds[:-1], bounds[1:])]
            self.prediction = prediction
            return prediction
        X, _ = self.tensor_transform_function(batchX, None, "predict")      # clair/model.py:953
        X, dtype = self._as_input(X)
        n = X.shape[0]
        prediction = [np.empty((n, k), dtype=np.float32) for k in self.output_label_split]    # fresh every call
        with self._lock:
            for s in range(0, n, self.max_sites):
                m = min(self.max_sites, n - s)
                rc = self._lib.clairb_predict_split(
                    self._h, X[s:s + m].ctypes.data_as(ctypes.c_void_p), dtype, m,
                    *[a[s:s + m].ctypes.data_as(ctypes.c_void_p) for a in prediction])
                _lib.check(rc, self._h, "clairb_predict_split")
        self.prediction = prediction
        return prediction

    def predict_and_decide(self, batchX, ref_bases):
        """predict() plus the first-choice variant decision of every site in the same device pass: returns
        (prediction, Decision) with `prediction` the list of four arrays predict() returns (also stored in
        .prediction).  See predict_and_decide_packed."""
        from . import decision as _decision
        if not self._has_weights:
            raise RuntimeError("predict() before init()/restore_parameters()")
        X, _ = self.tensor_transform_function(batchX, None, "predict")
        X, dtype = self._as_input(X)
        n = X.shape[0]
        ref = np.ascontiguousarray(ref_bases, dtype=np.uint8).reshape(-1)
        if ref.shape[0] != n or (ref > 3).any():
            raise ValueError("ref_bases must be %d codes in 0..3" % n)
        prediction = [np.empty((n, k), dtype=np.float32) for k in self.output_label_split]
        rec = np.empty((n, _lib.DECISION_WORDS), dtype=np.int32)
        with self._lock:
            for s in range(0, n, self.max_sites):
                m = min(self.max_sites, n - s)
                rc = self._lib.clairb_predict_split_decide(
                    self._h, X[s:s + m].ctypes.data_as(ctypes.c_void_p), dtype, m, ref[s:s + m].ctypes.data_as(ctypes.c_void_p),
                    *([a[s:s + m].ctypes.data_as(ctypes.c_void_p) for a in prediction] + [rec[s:s + m].ctypes.data_as(ctypes.c_void_p)]))
                _lib.check(rc, self._h, "clairb_predict_split_decide")
        self.prediction = prediction
        return prediction, _decision.unpack(rec)

    def predict_and_decide_packed(self, batchX, ref_bases):
        """predict_packed() plus the first-choice variant decision of every site in the same device pass.

        ref_bases: [n] uint8 codes 0..3 (clair_b200.decision.ref_base_codes).  Returns ([n,90] probabilities,
        clair_b200.decision.Decision of arrays) - what possible_outcome_probabilites_from + the first pass of
        output_from's loop would select (clair/call_var.py:589-690, 732-760)."""
        from . import decision as _decision
        if not self._has_weights:
            raise RuntimeError("predict() before init()/restore_parameters()")
        X, _ = self.tensor_transform_function(batchX, None, "predict")
        X, dtype = self._as_input(X)
        n = X.shape[0]
        ref = np.ascontiguousarray(ref_bases, dtype=np.uint8).reshape(-1)
        if ref.shape[0] != n or (ref > 3).any():
            raise ValueError("ref_bases must be %d codes in 0..3" % n)
        out = np.empty((n, _lib.N_OUT), dtype=np.float32)
        rec = np.empty((n, _lib.DECISION_WORDS), dtype=np.int32)
        with self._lock:
            for s in range(0, n, self.max_sites):
                m = min(self.max_sites, n - s)
                rc = self._lib.clairb_predict_decide(
                    self._h, X[s:s + m].ctypes.data_as(ctypes.c_void_p), dtype, m,
                    ref[s:s + m].ctypes.data_as(ctypes.c_void_p), out[s:s + m].ctypes.data_as(ctypes.c_void_p),
                    rec[s:s + m].ctypes.data_as(ctypes.c_void_p))
                _lib.check(rc, self._h, "clairb_predict_decide")
        return out, _decision.unpack(rec)

    def decide(self, probs, ref_bases, batchX=None):
        """Decision alone from [n,90] probabilities the caller holds (e.g. an ensemble average)."""
        from . import decision as _decision
        P = np.ascontiguousarray(probs, dtype=np.float32)
        if P.ndim != 2 or P.shape[1] != _lib.N_OUT or P.shape[0] < 1:
            raise ValueError("probs must be [n,90]")
        n = P.shape[0]
        ref = np.ascontiguousarray(ref_bases, dtype=np.uint8).reshape(-1)
        if ref.shape[0] != n or (ref > 3).any():
            raise ValueError("ref_bases must be %d codes in 0..3" % n)
        xp, dtype = None, _lib.DTYPE_F32
        if batchX is not None:
            X, dtype = self._as_input(batchX)
            if X.shape[0] != n:
                raise ValueError("batchX and probs disagree on n")
            xp = X.ctypes.data_as(ctypes.c_void_p)
        rec = np.empty((n, _lib.DECISION_WORDS), dtype=np.int32)
        with self._lock:
            rc = self._lib.clairb_decide(self._h, P.ctypes.data_as(ctypes.c_void_p), ref.ctypes.data_as(ctypes.c_void_p),
                                         xp, dtype, n, rec.ctypes.data_as(ctypes.c_void_p))
            _lib.check(rc, self._h, "clairb_decide")
        return _decision.unpack(rec)

    def get_layer(self, layer, n):
        """Parity hook: activations of the last (single-chunk) predict at one graph stage."""
        shapes = {_lib.LAYER_LSTM1: (33, n, 256), _lib.LAYER_LSTM2: (33, n, 256), _lib.LAYER_L3: (n, 30, 256),
                  _lib.LAYER_L4: (n, 192), _lib.LAYER_LOGITS: (n, 90)}
        out = np.empty(shapes[layer], dtype=np.float32)
        rc = self._lib.clairb_get_layer(self._h, layer, out.ctypes.data_as(ctypes.c_void_p), n)
        _lib.check(rc, self._h, "clairb_get_layer")
        return out

    def predict_device(self, x_ptr, dtype, n, out_ptr, stream=0):
        """Device-resident forward on `stream` (raw pointers; used by bench.py / shard path)."""
        rc = self._lib.clairb_predict_device(self._h, ctypes.c_void_p(x_ptr), dtype, n, ctypes.c_void_p(out_ptr),
                                             ctypes.c_void_p(stream))
        _lib.check(rc, self._h, "clairb_predict_device")

    def kernel_launches(self):
        return int(self._lib.clairb_kernel_launches(self._h))

    def set_profiling(self, enabled):
        _lib.check(self._lib.clairb_set_profiling(self._h, int(bool(enabled))), self._h, "clairb_set_profiling")

    def read_profile(self):
        """[{kernel, launches, ms}] since profiling was enabled / last read."""
        import json
        buf = ctypes.create_string_buffer(1 << 16)
        _lib.check(self._lib.clairb_read_profile(self._h, buf, len(buf)), self._h, "clairb_read_profile")
        return json.loads(buf.value.decode())

    # ---- lifetime --------------------------------------------------------------------------
    def close(self):
        """Reference clair/model.py:872-876."""
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.clairb_destroy(h)

    def __del__(self):                                         # clair/model.py:1149
        try:
            self.close()
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float32):
    """numpy array over page-locked host memory (async staging buffer for predict())."""
    lib = _lib.load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    ptr = ctypes.c_void_p()
    _lib.check(lib.clairb_host_alloc(ctypes.byref(ptr), max(nbytes, 1)), None, "clairb_host_alloc")
    buf = (ctypes.c_char * nbytes).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    _PINNED[arr.ctypes.data] = ptr
    return arr


def pinned_free(arr):
    ptr = _PINNED.pop(arr.ctypes.data, None)
    if ptr is not None:
        _lib.load().clairb_host_free(ptr)


_PINNED = {}
