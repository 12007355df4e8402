class CfgNode(dict):
    def __init__(self, init=None, **kw):
        super().__init__()
        for k, v in dict(init or {}, **kw).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v
