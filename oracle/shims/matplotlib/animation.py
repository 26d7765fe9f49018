class FuncAnimation:
    def __init__(self, *a, **k):
        raise NotImplementedError("matplotlib stand-in: no rendering")
