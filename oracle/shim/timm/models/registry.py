_REG = {}


def register_model(fn):
    _REG[fn.__name__] = fn
    return fn
