import torch
TORCH_VERSION = tuple(int(x) for x in torch.__version__.split(".")[:2])
