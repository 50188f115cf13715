def random_uniform(*a, **kw):
    raise NotImplementedError("training-mode dropout is outside the forward path")
