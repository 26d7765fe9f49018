def get_cmap(name):
    raise NotImplementedError("matplotlib stand-in: no rendering")
