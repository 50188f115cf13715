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
            devices=None,      # e.g. [0, 1, ..., 7]: one engine (weights replicated) per GPU behind the same predict()
            library=None,      # another build of the C-ABI library (tests: the cross-check build, _lib.XCHECK_PATH)
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
        devices = params["devices"]
        self.devices = [int(d) for d in devices] if devices is not None else [int(params["device"])]
        if not self.devices or len(set(self.devices)) != len(self.devices):
            raise ValueError("devices must be a non-empty list of distinct CUDA ordinals")
        self.device = self.devices[0]
        self.max_sites = int(params["max_sites"])
        self.batch_sites = int(params["batch_sites"])
        self._seed = params["seed"]
        self.prediction = None
        self._lock = threading.Lock()
        self._lib = _lib.load(params["library"])               # raises if the extension is missing
        self._engines = []                                     # one handle per GPU; every op is per-site
        self._h = None
        for dev in self.devices:
            handle = ctypes.c_void_p()
            rc = self._lib.clairb_create(dev, self.max_sites, self.batch_sites, ctypes.byref(handle))
            if rc:
                self.close()
            _lib.check(rc, None, "clairb_create(device %d)" % dev, self._lib)
            self._engines.append(handle)
        self._h = self._engines[0]
        self._next_engine = 0
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
        for h in self._engines:                                # replicated: 9.5 MB per GPU
            for name in _weights.weight_shapes():
                arr = np.ascontiguousarray(weights[name], dtype=np.float32)
                shape = (ctypes.c_int64 * arr.ndim)(*arr.shape)
                rc = self._lib.clairb_set_weight(h, name.encode(), arr.ctypes.data_as(ctypes.c_void_p), shape, arr.ndim)
                _lib.check(rc, h, "clairb_set_weight(%s)" % name, self._lib)
            _lib.check(self._lib.clairb_finalize_weights(h), h, "clairb_finalize_weights", self._lib)
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
                _lib.check(rc, self._h, "clairb_predict", self._lib)
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
        if len(self._engines) > 1:
            return self.predict_async(batchX, shard=True).result()
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
                _lib.check(rc, self._h, "clairb_predict_split", self._lib)
        self.prediction = prediction
        return prediction

    def predict_async(self, batchX, ref_bases=None, shard=None):
        """Queue predict(batchX) and return a PredictTicket at once; ``ticket.result()`` blocks until the four arrays
        are written, stores them in ``.prediction`` like predict() and returns them (with ref_bases:
        ``(prediction, Decision)`` like predict_and_decide).  Many calls may be in flight; the library packs the sites
        of consecutive calls into full device chunks (clairb_predict_async), so a loop of 1000-site batches
        (clair/call_var.py:1340-1344) runs at the throughput of one large call.  Results are bit-identical to
        predict().  With several engines (devices=[...]) whole batches go round-robin to the GPUs, or - shard=True,
        the default for calls of at least 4096 sites per GPU - one call is cut into contiguous slices
        (shard.shard_bounds), one per GPU."""
        if not self._has_weights:
            raise RuntimeError("predict() before init()/restore_parameters()")
        X, _ = self.tensor_transform_function(batchX, None, "predict")
        X, dtype = self._as_input(X)
        n = X.shape[0]
        ref = rec = None
        if ref_bases is not None:
            ref = np.ascontiguousarray(ref_bases, dtype=np.uint8).reshape(-1)
            if ref.shape[0] != n or (ref > 3).any():
                raise ValueError("ref_bases must be %d codes in 0..3" % n)
            rec = np.empty((n, _lib.DECISION_WORDS), dtype=np.int32)
        prediction = [np.empty((n, k), dtype=np.float32) for k in self.output_label_split]    # fresh every call
        g = len(self._engines)
        if shard is None:
            shard = g > 1 and n >= 4096 * g
        if g > 1 and shard:
            from .shard import shard_bounds
            parts = [(self._engines[r],) + shard_bounds(n, g, r) for r in range(g)]
        else:
            parts = [(self._engines[self._next_engine % g], 0, n)]
            self._next_engine += 1
        waits = []
        for h, lo, hi in parts:
            if hi <= lo:
                continue
            t = ctypes.c_int64()
            rc = self._lib.clairb_predict_async(
                h, X[lo:hi].ctypes.data_as(ctypes.c_void_p), dtype, hi - lo,
                *([a[lo:hi].ctypes.data_as(ctypes.c_void_p) for a in prediction] +
                  [ref[lo:hi].ctypes.data_as(ctypes.c_void_p) if ref is not None else None,
                   rec[lo:hi].ctypes.data_as(ctypes.c_void_p) if rec is not None else None, ctypes.byref(t)]))
            if rc:
                for hw, tw in waits:                           # do not leave earlier slices in flight behind an error
                    self._lib.clairb_predict_wait(hw, tw)
            _lib.check(rc, h, "clairb_predict_async", self._lib)
            waits.append((h, t.value))
        return PredictTicket(self, waits, (X, ref), prediction, rec)

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
                _lib.check(rc, self._h, "clairb_predict_split_decide", self._lib)
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
                _lib.check(rc, self._h, "clairb_predict_decide", self._lib)
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
            _lib.check(rc, self._h, "clairb_decide", self._lib)
        return _decision.unpack(rec)

    def get_layer(self, layer, n):
        """Parity hook: activations of the last (single-chunk) predict at one graph stage."""
        shapes = {_lib.LAYER_LSTM1: (33, n, 256), _lib.LAYER_LSTM2: (33, n, 256), _lib.LAYER_L3: (n, 30, 256),
                  _lib.LAYER_L4: (n, 192), _lib.LAYER_LOGITS: (n, 90)}
        out = np.empty(shapes[layer], dtype=np.float32)
        rc = self._lib.clairb_get_layer(self._h, layer, out.ctypes.data_as(ctypes.c_void_p), n)
        _lib.check(rc, self._h, "clairb_get_layer", self._lib)
        return out

    def predict_device(self, x_ptr, dtype, n, out_ptr, stream=0):
        """Device-resident forward on `stream` (raw pointers; used by bench.py / shard path)."""
        rc = self._lib.clairb_predict_device(self._h, ctypes.c_void_p(x_ptr), dtype, n, ctypes.c_void_p(out_ptr),
                                             ctypes.c_void_p(stream))
        _lib.check(rc, self._h, "clairb_predict_device", self._lib)

    def kernel_launches(self):
        return int(self._lib.clairb_kernel_launches(self._h))

    def set_profiling(self, enabled):
        _lib.check(self._lib.clairb_set_profiling(self._h, int(bool(enabled))), self._h, "clairb_set_profiling", self._lib)

    def read_profile(self):
        """[{kernel, launches, ms}] since profiling was enabled / last read."""
        import json
        buf = ctypes.create_string_buffer(1 << 16)
        _lib.check(self._lib.clairb_read_profile(self._h, buf, len(buf)), self._h, "clairb_read_profile", self._lib)
        return json.loads(buf.value.decode())

    # ---- lifetime --------------------------------------------------------------------------
    def close(self):
        """Reference clair/model.py:872-876."""
        engines, self._engines, self._h = getattr(self, "_engines", []), [], None
        for h in engines:
            self._lib.clairb_destroy(h)

    def __del__(self):                                         # clair/model.py:1149
        try:
            self.close()
        except Exception:
            pass


class PredictTicket(object):
    """One predict_async call in flight.  Holds the input and output arrays alive until result() has returned."""

    def __init__(self, model, waits, inputs, prediction, rec):
        self._model, self._waits, self._inputs, self._prediction, self._rec = model, waits, inputs, prediction, rec
        self._error = None

    def result(self):
        m = self._model
        while self._waits:
            h, t = self._waits.pop(0)
            rc = m._lib.clairb_predict_wait(h, t)
            if rc and self._error is None:
                try:
                    _lib.check(rc, h, "clairb_predict_wait", m._lib)
                except Exception as exc:                       # keep waiting for the other slices, then raise
                    self._error = exc
        self._inputs = None
        if self._error is not None:
            raise self._error
        m.prediction = self._prediction
        if self._rec is None:
            return self._prediction
        from . import decision as _decision
        return self._prediction, _decision.unpack(self._rec)

    def __del__(self):
        try:
            while self._waits:                                 # never let the library write into freed arrays
                h, t = self._waits.pop(0)
                self._model._lib.clairb_predict_wait(h, t)
        except Exception:
            pass


class PinnedPool(object):
    """A fixed set of page-locked batch buffers handed out by take() and returned by give(): the staging ring of the
    batch loop (replaces the pageable np.empty per batch of clair/utils.py:85).  take() has the signature of
    tensor_generator_from's `alloc`; it blocks while every buffer is in use, which is what bounds the loader."""

    def __init__(self, count, shape, dtype=np.float32):
        import queue
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self._slab = pinned_empty((count,) + self.shape, self.dtype)
        self._free = queue.Queue()
        self._index = {}
        for i in range(count):
            self._index[self._slab[i].ctypes.data] = i
            self._free.put(i)

    def take(self, shape=None, dtype=None):
        if shape is not None and (int(np.prod(shape)) != int(np.prod(self.shape)) or np.dtype(dtype or self.dtype) != self.dtype):
            raise ValueError("PinnedPool holds %s %s buffers" % (self.shape, self.dtype))
        buf = self._slab[self._free.get()]
        return buf if shape is None else buf.reshape(shape)

    def give(self, arr):
        """Return the buffer `arr` is a view of (any leading-rows view of what take() handed out)."""
        i = self._index.get(arr.ctypes.data)
        if i is not None:
            self._free.put(i)

    def close(self):
        slab, self._slab = self._slab, None
        if slab is not None:
            pinned_free(slab)


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
