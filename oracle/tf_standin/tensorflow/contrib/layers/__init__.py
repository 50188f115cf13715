def variance_scaling_initializer(**kw):
    return None                                  # weights are restored, never initialised, in the stand-in
