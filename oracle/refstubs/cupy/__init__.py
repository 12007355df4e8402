def memoize(for_each_device=False):
    def deco(f):
        return f
    return deco
