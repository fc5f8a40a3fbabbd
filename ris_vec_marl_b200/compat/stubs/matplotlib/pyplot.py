def __getattr__(name):
    def _noop(*a, **k):
        return None

    return _noop
