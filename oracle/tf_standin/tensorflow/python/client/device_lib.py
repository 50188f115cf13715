def list_local_devices():
    return []                                    # no GPU: Clair builds the CudnnCompatibleLSTMCell branch (clair/model.py:298-312)
