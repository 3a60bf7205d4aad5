"""`import dream2real_b200.pyngp as ngp` -- the names of the reference's pybind11 module that the
Dream2Real path uses (reference reconstruction/instant-ngp/src/python_api.cu:262-292,382-566)."""
from .testbed import (Depth, Nerf, RenderMode, Shade, Testbed, TestbedMode,  # noqa: F401
                      free_temporary_memory)
