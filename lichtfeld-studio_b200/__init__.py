"""lichtfeld-studio_b200 -- B200-native (sm_100a) 3D Gaussian Splatting training rasterizer.

Drop-in for the training hot path of MrNeRF/LichtFeld-Studio behind its own operator surface:
  ops      -- gsplat/Ops.h mirror (projection_ut_3dgs_fused, spherical_harmonics_fwd/bwd, intersect_tile,
              intersect_offset, rasterize_to_pixels_from_world_3dgs_fwd/bwd) + fastgs adam_step
  trainer  -- fused per-view step + multi-tensor Adam + view-sharded data parallel
  scene    -- seeded synthetic scenes for the BASELINE.json configurations
All compute goes through liblfs_b200.so (C ABI: include/lfs_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import LfsError, LfsUnsupported, UTParams, build, load  # noqa: F401

__all__ = ["_lib", "ops", "trainer", "scene", "dp", "LfsError", "LfsUnsupported", "UTParams", "build", "load"]
