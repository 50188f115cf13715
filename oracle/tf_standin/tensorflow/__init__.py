"""A numpy stand-in for the ~60 TensorFlow 1.13 symbols HKU-BAL/Clair's `clair/model.py` and `clair/selu.py` touch.
TEST INFRASTRUCTURE ONLY (oracle/gen_golden_reference_model.py puts this directory on sys.path so that the REFERENCE's
unmodified model code imports it as `tensorflow`; nothing in the product or in the tests' run time imports it).

Why: the reference's arithmetic lives in TensorFlow 1.13.2, which has no wheel for this container's Python and cannot be
installed (no network), so the reference's `Clair.predict` cannot run here as it is.  What CAN run is the reference's own
graph-building code - reshape / transpose order, which tensor feeds which layer, the slice-dense unstack / stack axes,
the flatten order, which activation sits where, the variable scopes, the output list - over a lazy numpy interpreter of
the TF ops it calls.  The ops' semantics are TensorFlow's documented ones:
    tf.layers.dense            act(x @ kernel + bias), kernel [in, out], variables "<scope>/<name>/{kernel,bias}"
    tf.reshape / transpose / unstack / stack / split     row-major, as numpy
    tf.nn.softmax              along the last axis        tf.nn.elu   exp(x) - 1 for x < 0
    tf.contrib.cudnn_rnn.CudnnCompatibleLSTMCell(n)  =  LSTMBlockCell(n, forget_bias=0, cell_clip=None, use_peephole=False):
        [i, ci, f, o] = split([x, h] @ kernel + bias, 4);  c' = sigmoid(f) * c + sigmoid(i) * tanh(ci);  h' = sigmoid(o) * tanh(c')
        (tensorflow/contrib/rnn/python/ops/lstm_ops.py: "icfo" gate order), variables ".../cudnn_compatible_lstm_cell/{kernel,bias}"
    tf.contrib.rnn.stack_bidirectional_dynamic_rnn    per layer: fw over t = 0..T-1, bw over t = T-1..0 re-reversed,
        outputs concatenated [fw, bw] on the last axis; scopes "stack_bidirectional_rnn/cell_<k>/bidirectional_rnn/{fw,bw}/"
Training-side calls (optimisers, summaries, savers) are accepted and ignored; evaluating a dropout branch in training
mode raises.  The cell arithmetic above is therefore still a restatement of TensorFlow's kernels (cross-checked by cuDNN's
LSTM on the GPU box, oracle/clair_oracle_cudnn.py); everything ABOVE the ops is the reference's own code.
"""
import contextlib
import os

import numpy as np

float32, float64, bool, uint8 = np.float32, np.float64, np.bool_, np.uint8
COMPUTE_DTYPE = np.float64 if os.environ.get("TF_STANDIN_FLOAT64") else None       # None: the placeholder's own dtype


# ---- lazy tensors -----------------------------------------------------------------------------------------------------
class Tensor(object):
    def __init__(self, fn, name=None, dtype=None, static=None):
        self._fn, self.name, self.dtype = fn, name, dtype
        self.static = static                         # static shape where the stand-in can know it (None entries = unknown)

    def eval(self, feed, cache):
        key = id(self)
        if key not in cache:
            cache[key] = feed[self] if self in feed else self._fn(feed, cache)
        return cache[key]

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def get_shape(self):
        return _AnyShape()

    def set_shape(self, shape):
        pass

    def _binary(self, other, op, swap=False):
        o = convert_to_tensor(other)
        return Tensor(lambda f, c: op(o.eval(f, c), self.eval(f, c)) if swap else op(self.eval(f, c), o.eval(f, c)))

    __add__ = lambda s, o: s._binary(o, np.add)
    __radd__ = lambda s, o: s._binary(o, np.add, True)
    __sub__ = lambda s, o: s._binary(o, np.subtract)
    __rsub__ = lambda s, o: s._binary(o, np.subtract, True)
    __mul__ = lambda s, o: s._binary(o, np.multiply)
    __rmul__ = lambda s, o: s._binary(o, np.multiply, True)
    __truediv__ = lambda s, o: s._binary(o, np.divide)
    __rtruediv__ = lambda s, o: s._binary(o, np.divide, True)
    __pow__ = lambda s, o: s._binary(o, np.power)
    __gt__ = lambda s, o: s._binary(o, np.greater)
    __ge__ = lambda s, o: s._binary(o, np.greater_equal)
    __lt__ = lambda s, o: s._binary(o, np.less)
    __neg__ = lambda s: Tensor(lambda f, c: -s.eval(f, c))

    def __getitem__(self, index):
        return Tensor(lambda f, c: self.eval(f, c)[index])


class _AnyShape(object):
    def assert_is_compatible_with(self, other):
        pass

    def __iter__(self):
        return iter(())


class Variable(Tensor):
    def __init__(self, name, shape, dtype):
        Tensor.__init__(self, lambda f, c: self.value.astype(COMPUTE_DTYPE) if COMPUTE_DTYPE else self.value, name + ":0", dtype)
        self.var_name, self.shape = name, tuple(shape)
        self.value = np.zeros(self.shape, dtype)
        self.op = type("Op", (), {"name": name})()


def convert_to_tensor(value, dtype=None, name=None):
    if isinstance(value, Tensor):
        return value
    # Python scalars stay Python scalars: TF gives a constant the dtype of the tensor it meets, numpy does the same for a
    # Python float ("weak" scalar) but would promote a float32 array to float64 against a 0-d float64 array
    const = value if (isinstance(value, (int, float)) and dtype is None) else np.asarray(value, dtype=dtype)
    return Tensor(lambda f, c: const)


constant = convert_to_tensor


# ---- graph / scopes / session --------------------------------------------------------------------------------------------
class Graph(object):
    _stack = []

    def __init__(self):
        self.variables = {}                          # name -> Variable, in creation order
        self.scope = []

    @contextlib.contextmanager
    def as_default(self):
        Graph._stack.append(self)
        try:
            yield self
        finally:
            Graph._stack.pop()


_DEFAULT = Graph()


def _graph():
    return Graph._stack[-1] if Graph._stack else _DEFAULT


class _Scope(object):
    def __init__(self, name):
        self.name, self.reuse = name, False


@contextlib.contextmanager
def variable_scope(name, *a, **kw):
    g = _graph()
    g.scope.append(name)
    try:
        yield _Scope("/".join(g.scope))
    finally:
        g.scope.pop()


@contextlib.contextmanager
def name_scope(name, default_name=None, values=None):
    yield (name or default_name or "") + "/"         # op names only: variables are not scoped by name_scope


def get_variable_scope():
    return _Scope("/".join(_graph().scope))


def get_variable(name, shape, dtype=float32):
    g = _graph()
    full = "/".join(g.scope + [name])
    if full in g.variables:
        raise ValueError("Variable %s already exists" % full)
    g.variables[full] = Variable(full, shape, dtype)
    return g.variables[full]


def trainable_variables():
    return list(_graph().variables.values())


def global_variables_initializer():
    return Tensor(lambda f, c: None, "init")


def set_random_seed(seed):
    pass


class ConfigProto(object):
    pass


class Dimension(object):
    def __init__(self, value):
        self.value = value


class Session(object):
    def __init__(self, graph=None, config=None):
        self.graph = graph

    def run(self, fetches, feed_dict=None):
        feed = {}
        for k, v in (feed_dict or {}).items():
            dt = COMPUTE_DTYPE if (COMPUTE_DTYPE and k.dtype in (float32, float64)) else k.dtype
            feed[k] = np.asarray(v, dtype=dt)
        cache = {}
        one = lambda t: t.eval(feed, cache) if isinstance(t, Tensor) else t
        if isinstance(fetches, (list, tuple)):
            return [one(t) for t in fetches]
        return one(fetches)

    def close(self):
        pass


def placeholder(dtype, shape=None, name=None):
    def missing(f, c, name=name):
        raise KeyError("placeholder %s was not fed" % name)
    return Tensor(missing, name, dtype, static=tuple(shape) if shape is not None else None)


# ---- ops ---------------------------------------------------------------------------------------------------------------
def _unary(fn):
    return lambda x, name=None, **kw: Tensor(lambda f, c: fn(convert_to_tensor(x).eval(f, c)))


log = _unary(np.log)
floor = _unary(np.floor)
sqrt = _unary(np.sqrt)


def shape(x, name=None):
    return Tensor(lambda f, c: np.array(convert_to_tensor(x).eval(f, c).shape))


def reshape(tensor, shape, name=None):
    dims = [convert_to_tensor(d) for d in shape]
    static = tuple(d if isinstance(d, (int, np.integer)) else None for d in shape)
    return Tensor(lambda f, c: np.reshape(tensor.eval(f, c), [int(d.eval(f, c)) for d in dims]), static=static)   # row-major like TF


def transpose(a, perm=None, name=None):
    static = tuple(a.static[p] for p in perm) if (a.static is not None and perm is not None) else None
    return Tensor(lambda f, c: np.transpose(a.eval(f, c), perm), static=static)


def unstack(value, axis=0, name=None, num=None):
    # TF needs the static length of the axis; so does this (Clair unstacks axis 2 of [B, 33, 2 * LSTM2_num_units])
    if num is None:
        if value.static is None or value.static[axis] is None:
            raise ValueError("Cannot infer num from shape %s" % (value.static,))
        num = value.static[axis]
    return [Tensor(lambda f, c, i=i: np.take(value.eval(f, c), i, axis=axis)) for i in range(num)]


def stack(values, axis=0, name=None):
    vals = [convert_to_tensor(v) for v in values]
    return Tensor(lambda f, c: np.stack([v.eval(f, c) for v in vals], axis=axis))


def split(value, num_or_size_splits, axis=0, name=None):
    bounds = np.cumsum(num_or_size_splits)[:-1]
    return [Tensor(lambda f, c, i=i: np.split(value.eval(f, c), bounds, axis=axis)[i]) for i in range(len(num_or_size_splits))]


def where(condition, x, y, name=None):
    cnd, a, b = convert_to_tensor(condition), convert_to_tensor(x), convert_to_tensor(y)
    return Tensor(lambda f, c: np.where(cnd.eval(f, c), a.eval(f, c), b.eval(f, c)))


def reduce_sum(x, axis=None, name=None, reduction_indices=None, **kw):
    ax = axis if axis is not None else reduction_indices
    return Tensor(lambda f, c: np.sum(x.eval(f, c), axis=tuple(ax) if isinstance(ax, (list, tuple)) else ax))


def multiply(x, y, name=None):
    return convert_to_tensor(x) * y


def clip_by_value(t, lo, hi, name=None):
    return Tensor(lambda f, c: np.clip(convert_to_tensor(t).eval(f, c), lo, hi))


def add_n(inputs, name=None):
    return Tensor(lambda f, c: sum(t.eval(f, c) for t in inputs))


def clip_by_global_norm(t_list, clip_norm, name=None):
    return list(t_list), Tensor(lambda f, c: None)


def py_func(*a, **kw):
    return Tensor(lambda f, c: None)


# ---- tf.nn / tf.layers / tf.contrib / tf.train / tf.summary ----------------------------------------------------------------
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _softmax(x):
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


class _NN(object):
    softmax = staticmethod(lambda logits, name=None, **kw: Tensor(lambda f, c: _softmax(logits.eval(f, c))))
    elu = staticmethod(lambda x, name=None: Tensor(lambda f, c: (lambda v: np.where(v < 0, np.expm1(np.minimum(v, 0)), v))(x.eval(f, c))))
    l2_loss = staticmethod(lambda t, name=None: Tensor(lambda f, c: np.sum(np.square(t.eval(f, c))) / 2))

    class rnn_cell(object):
        @staticmethod
        def MultiRNNCell(cells):
            raise NotImplementedError("unidirectional LSTM stacks are not on Clair's forward path")

    @staticmethod
    def dynamic_rnn(*a, **kw):
        raise NotImplementedError("unidirectional LSTM stacks are not on Clair's forward path")


nn = _NN()


class _Layers(object):
    @staticmethod
    def dense(inputs, units, activation=None, name=None, kernel_initializer=None, **kw):
        # tf.layers.dense builds its variables under variable_scope(name): "<scope>/<name>/kernel" [in, units], "/bias" [units];
        # the input width is only known at run time here, so the variables are declared lazily at first evaluation - except that
        # trainable_variables() / Saver must see them at build time: their shapes are filled in by Saver.restore
        with variable_scope(name):
            kernel, bias = get_variable("kernel", (0, units)), get_variable("bias", (units,))

        def run(f, c):
            x = inputs.eval(f, c)
            return np.matmul(x, kernel.value.astype(x.dtype)) + bias.value.astype(x.dtype)
        out = Tensor(run)
        return activation(out) if activation is not None else out

    @staticmethod
    def dropout(inputs, rate=0.5, noise_shape=None, seed=None, training=False, name=None):
        def run(f, c):
            if np.asarray(convert_to_tensor(training).eval(f, c)).item() and rate:
                raise NotImplementedError("training-mode dropout is outside the forward path")
            return inputs.eval(f, c)                     # tf.layers.dropout is the identity when training is False
        return Tensor(run, static=inputs.static)


layers = _Layers()


class CudnnCompatibleLSTMCell(object):
    """= LSTMBlockCell(num_units, forget_bias=0, cell_clip=None, use_peephole=False) with the scope name
    "cudnn_compatible_lstm_cell" (tensorflow/contrib/cudnn_rnn/python/ops/cudnn_rnn_ops.py)."""
    scope_name = "cudnn_compatible_lstm_cell"

    def __init__(self, num_units, reuse=None):
        self.num_units = num_units

    def build(self):
        with variable_scope(self.scope_name):
            self.kernel, self.bias = get_variable("kernel", (0, 4 * self.num_units)), get_variable("bias", (4 * self.num_units,))

    def run(self, xs):
        """xs [T, B, in] -> hs [T, B, units]; zero initial state."""
        T, B = xs.shape[0], xs.shape[1]
        n = self.num_units
        k, b = self.kernel.value.astype(xs.dtype), self.bias.value.astype(xs.dtype)
        h = np.zeros((B, n), xs.dtype)
        cs = np.zeros((B, n), xs.dtype)
        out = np.empty((T, B, n), xs.dtype)
        for t in range(T):
            z = np.matmul(np.concatenate([xs[t], h], axis=1), k) + b
            i, ci, fg, o = np.split(z, 4, axis=1)                                  # "icfo"
            cs = _sigmoid(fg) * cs + _sigmoid(i) * np.tanh(ci)                    # forget_bias = 0
            h = _sigmoid(o) * np.tanh(cs)
            out[t] = h
        return out


def stack_bidirectional_dynamic_rnn(cells_fw, cells_bw, inputs, dtype=None, time_major=False, **kw):
    if not time_major:
        raise NotImplementedError("Clair feeds time-major inputs")
    with variable_scope("stack_bidirectional_rnn"):
        for k, (cf, cb) in enumerate(zip(cells_fw, cells_bw)):
            with variable_scope("cell_%d" % k):
                with variable_scope("bidirectional_rnn"):
                    with variable_scope("fw"):
                        cf.build()
                    with variable_scope("bw"):
                        cb.build()

    def run(f, c):
        x = inputs.eval(f, c)
        for cf, cb in zip(cells_fw, cells_bw):
            fw = cf.run(x)
            bw = cb.run(x[::-1])[::-1]                   # array_ops.reverse_sequence in, and the outputs reversed back
            x = np.concatenate([fw, bw], axis=2)
        return x
    static = (inputs.static[0], inputs.static[1], 2 * cells_fw[-1].num_units) if inputs.static is not None else None
    out = Tensor(run, static=static)
    return out, Tensor(lambda f, c: None), Tensor(lambda f, c: None)


class _Optimizer(object):
    def __init__(self, *a, **kw):
        pass

    def compute_gradients(self, loss):
        return [(Tensor(lambda f, c: None), v) for v in trainable_variables()]

    def apply_gradients(self, grads_and_vars):
        return Tensor(lambda f, c: None)

    def minimize(self, loss):
        return Tensor(lambda f, c: None)


class _Saver(object):
    """restore(session, path): `path` is an .npz keyed by variable name (clair_b200.weights.save_blob)."""

    def __init__(self, *a, **kw):
        pass

    def restore(self, session, path):
        variables = session.graph.variables
        with np.load(path if str(path).endswith(".npz") else str(path) + ".npz") as z:
            names = set(z.files)
            if names != set(variables):
                raise KeyError("checkpoint / graph variable names differ: only in checkpoint %s, only in graph %s" % (
                    sorted(names - set(variables))[:3], sorted(set(variables) - names)[:3]))
            for n in z.files:
                v = variables[n]
                a = z[n]
                if a.shape[1:] != v.shape[1:] or a.ndim != len(v.shape):
                    raise ValueError("variable %s: checkpoint shape %s, graph shape %s" % (n, a.shape, v.shape))
                v.value, v.shape = a, a.shape

    def save(self, session, path):
        np.savez(path, **{n: v.value for n, v in session.graph.variables.items()})


class _Train(object):
    AdamOptimizer = MomentumOptimizer = _Optimizer
    Saver = _Saver


train = _Train()


class _Summary(object):
    scalar = histogram = image = staticmethod(lambda *a, **kw: Tensor(lambda f, c: None))
    merge = staticmethod(lambda inputs, **kw: Tensor(lambda f, c: None))
    merge_all = staticmethod(lambda **kw: Tensor(lambda f, c: None))

    class FileWriter(object):
        def __init__(self, *a, **kw):
            pass

        def add_summary(self, *a, **kw):
            pass

        def close(self):
            pass


summary = _Summary()


from . import contrib  # noqa: E402,F401  (tensorflow.contrib.{layers,cudnn_rnn,rnn}: real subpackages, `from tensorflow.contrib import layers`)
