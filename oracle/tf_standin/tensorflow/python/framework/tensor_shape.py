def scalar():
    return ()
