import functools


def configurable(init_func):
    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        if len(args) == 1 and not kwargs and hasattr(args[0], "keys") and hasattr(type(self), "from_config"):
            init_func(self, **type(self).from_config(args[0]))
        else:
            init_func(self, *args, **kwargs)
    return wrapped
