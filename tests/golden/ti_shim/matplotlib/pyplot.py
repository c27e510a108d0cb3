def axis(*a, **k):
    raise NotImplementedError("matplotlib stand-in of the Taichi shim")
