def set_loglevel(*a, **k):
    pass
