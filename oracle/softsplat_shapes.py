"""Shapes the reference's splat kernel is expanded for (oracle/build_softsplat_ref.py) and the GPU parity test replays
(tests/test_gpu_ops.py::test_softsplat_vs_reference_kernel).  (B, C, H, W) of the splat INPUT before the 'softmax' packing
adds the normalisation channel: the reference kernel sees C + 1 channels."""
SHAPES = [
    (1, 4, 12, 20),      # 2 samples + 2 costs (update_past_cost), small
    (2, 4, 48, 156),     # the same at the KITTI 384x1248 1/8 scale (BASELINE config C3), B=2
    (2, 3, 24, 40),      # local map, 3 channels
    (1, 1, 12, 20),      # first local-map frame, 1 channel
    (1, 4, 9, 13),       # odd sizes
]
