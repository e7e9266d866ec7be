"""Import shim: put `<repo>/gaussianavatars_b200/compat` on sys.path and the reference's
`from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`
(gaussian_renderer/__init__.py:15) resolves to the B200-native operator, unmodified."""
import os
import sys

_root = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
if _root not in sys.path:
    sys.path.insert(0, _root)

from gaussianavatars_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: E402,F401
                                             rasterize_gaussians)
