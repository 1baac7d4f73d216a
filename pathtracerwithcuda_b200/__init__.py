"""pathtracerwithcuda_b200 — B200-native drop-in for PathTracerWithCuda's render hot path.

The package is a thin host layer over the C-ABI library (include/ptb200.h, built in-tree into
libptb200.so from csrc/): `Renderer` wraps the handle API, `PathTracer` mirrors the reference's
`path_tracer` class.  No CPU / PyTorch fallback exists.
"""
from .api import (Camera, PathTracer, PtbError, Renderer, Stats, default_camera, device_count, last_error,  # noqa: F401
                  load_library, write_png, write_pfm, MultiRenderer, dist_unique_id, nccl_version, decode_image, set_jpeg_decode, MATERIAL_DTYPE, SPHERE_DTYPE, CONFIG_DTYPE)
