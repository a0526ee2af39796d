"""ctypes loader for liblfs_b200.so (the C ABI declared in include/lfs_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, an exception is
raised.  Nothing here imports, loads or executes anything under oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "liblfs_b200.so")
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")

ABI_VERSION = 2  # LFS_ABI_VERSION of include/lfs_b200.h
LFS_OK = 0
LFS_ERR_UNSUPPORTED = -2
LFS_ERR_CAPACITY = -5
LFS_TAG_SCRATCH, LFS_TAG_ISECT_IDS, LFS_TAG_FLATTEN_IDS = 0, 1, 2
LFS_TAG_FG_PER_PRIMITIVE, LFS_TAG_FG_PER_TILE, LFS_TAG_FG_PER_INSTANCE, LFS_TAG_FG_PER_BUCKET = 10, 11, 12, 13

PINHOLE, ORTHO, FISHEYE = 0, 1, 2
SHUTTER_GLOBAL = 4
IMG_U8_HWC, IMG_F32_HWC, IMG_F32_CHW = 0, 1, 2


class LfsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"lfs_b200 error {code}: {msg}")
        self.code = code


class LfsUnsupported(LfsError):
    pass


class UTParams(C.Structure):
    """gsplat/Cameras.h:27-44 UnscentedTransformParameters"""

    _fields_ = [
        ("alpha", C.c_float),
        ("beta", C.c_float),
        ("kappa", C.c_float),
        ("in_image_margin_factor", C.c_float),
        ("require_all_sigma_points_valid", C.c_int32),
    ]

    @staticmethod
    def default() -> "UTParams":
        return UTParams(0.1, 2.0, 0.0, 0.1, 1)


class TrainerDesc(C.Structure):
    _fields_ = [
        ("n_gaussians", C.c_uint32),
        ("sh_degree_max", C.c_uint32),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("eps2d", C.c_float),
        ("near_plane", C.c_float),
        ("far_plane", C.c_float),
        ("radius_clip", C.c_float),
        ("ut", UTParams),
        ("instance_capacity", C.c_uint64),
    ]


class AdamReg(C.Structure):
    """lfs_adam_reg: regulariser gradients folded into the Adam step (include/lfs_b200.h)."""

    _fields_ = [("kind", C.c_int * 16), ("coef", C.c_float * 16), ("plane_elems", C.c_int64), ("n_valid", C.c_int64)]


ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int, C.c_size_t)

_vp, _u32, _i32, _i64, _f, _sz = C.c_void_p, C.c_uint32, C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/lfs_b200.h one to one (checked by tests/test_abi.py)
SIGNATURES = {
    "lfs_last_error": (C.c_char_p, []),
    "lfs_abi_version": (C.c_int, []),
    "lfs_launch_count": (C.c_uint64, []),
    "lfs_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "lfs_projection_ut_3dgs_fused": (
        C.c_int,
        [_vp] * 7 + [_u32, _u32, _u32, _u32, _f, _f, _f, _f, C.c_int, C.POINTER(UTParams), C.c_int] + [_vp] * 3
        + [_vp] * 5 + [_vp],
    ),
    "lfs_spherical_harmonics_fwd": (C.c_int, [_u32, _vp, _vp, _vp, _u32, _u32, _vp, _vp]),
    "lfs_spherical_harmonics_bwd": (C.c_int, [_u32, _u32, _vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp]),
    "lfs_intersect_tile": (
        C.c_int,
        [_vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, C.c_int, _vp, ALLOC_FN, _vp, C.POINTER(_vp), C.POINTER(_vp),
         C.POINTER(_i64), _vp],
    ),
    "lfs_intersect_tile_packed": (
        C.c_int,
        [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, C.c_int, _vp, ALLOC_FN, _vp, C.POINTER(_vp), C.POINTER(_vp),
         C.POINTER(_i64), _vp],
    ),
    "lfs_fastgs_forward": (
        C.c_int,
        [_vp] * 8 + [_u32] + [C.c_int] * 4 + [C.c_float] * 6 + [_vp, _vp, ALLOC_FN, _vp] + [C.POINTER(C.c_int)] * 4 + [_vp],
    ),
    "lfs_fastgs_backward": (
        C.c_int,
        [_vp] * 8 + [_vp] * 4 + [_vp] * 8 + [_u32] + [C.c_int] * 7 + [C.c_float] * 4 + [ALLOC_FN, _vp, _vp],
    ),
    "lfs_intersect_offset": (C.c_int, [_vp, _i64, _u32, _u32, _u32, _vp, _vp]),
    "lfs_rasterize_to_pixels_from_world_3dgs_fwd": (
        C.c_int,
        [_vp] * 7 + [_u32] * 6 + [_vp] * 3 + [C.c_int, C.POINTER(UTParams), C.c_int] + [_vp] * 3 + [_vp, _vp, _i64]
        + [ALLOC_FN, _vp] + [_vp] * 3 + [_vp],
    ),
    "lfs_rasterize_to_pixels_from_world_3dgs_bwd": (
        C.c_int,
        [_vp] * 7 + [_u32] * 6 + [_vp] * 3 + [C.c_int, C.POINTER(UTParams), C.c_int] + [_vp] * 3 + [_vp, _vp, _i64]
        + [_vp] * 4 + [ALLOC_FN, _vp] + [_vp] * 5 + [_vp],
    ),
    "lfs_quat_scale_to_covar_preci_fwd": (C.c_int, [_vp, _vp, _u32, C.c_int, _vp, _vp, _vp]),
    "lfs_quat_scale_to_covar_preci_bwd": (C.c_int, [_vp, _vp, _u32, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "lfs_projection_ewa_3dgs_fused_fwd": (
        C.c_int, [_vp] * 7 + [_u32] * 4 + [_f] * 4 + [C.c_int] + [_vp] * 5 + [_vp]),
    "lfs_rasterize_to_pixels_3dgs_fwd": (C.c_int, [_vp] * 6 + [_u32] * 6 + [_vp, _vp, _i64] + [_vp] * 3 + [_vp]),
    "lfs_rasterize_to_pixels_3dgs_bwd": (
        C.c_int, [_vp] * 6 + [_u32] * 6 + [_vp, _vp, _i64] + [_vp] * 4 + [_vp] * 5 + [_vp]),
    "lfs_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _f, _vp]),
    "lfs_adam_step_multi": (
        C.c_int,
        [_vp, _vp, _vp, _vp, C.c_int, C.POINTER(_i64), C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), _f, _f, _f,
         C.c_int, C.POINTER(AdamReg), _vp],
    ),
    "lfs_adam_step_multi_p2p": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp,
                                          _f, _f, _f, C.POINTER(AdamReg), _vp]),
    "lfs_quats_to_rotmats": (C.c_int, [_vp, _u32, _vp, _vp]),
    "lfs_relocation": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _u32, _vp, _vp, _vp]),
    "lfs_add_noise": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _f, _u32, _vp]),
    "lfs_trainer_arena_floats": (C.c_uint64, [C.POINTER(TrainerDesc)]),
    "lfs_trainer_create": (_vp, [C.POINTER(TrainerDesc)]),
    "lfs_trainer_destroy": (None, [_vp]),
    "lfs_trainer_scratch_bytes": (C.c_uint64, [_vp]),
    "lfs_trainer_instance_capacity": (C.c_uint64, [_vp]),
    "lfs_trainer_pack": (C.c_int, [_vp] * 9),
    "lfs_trainer_unpack": (C.c_int, [_vp] * 9),
    "lfs_trainer_view_forward": (C.c_int, [_vp, _vp, C.POINTER(_f), C.POINTER(_f), _u32, C.POINTER(_f), _vp, _vp, _vp]),
    "lfs_trainer_view_loss_l1": (C.c_int, [_vp, _vp, C.c_int, _f, _vp, _vp]),
    "lfs_trainer_view_loss_ssim_l1": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_float, _vp, _vp]),
    "lfs_trainer_view_set_grad": (C.c_int, [_vp, _vp, _vp, _vp]),
    "lfs_trainer_view_backward": (C.c_int, [_vp, _vp, _vp, _vp]),
    "lfs_trainer_view_backward_blend": (C.c_int, [_vp, _vp]),
    "lfs_trainer_view_backward_params": (C.c_int, [_vp, _vp, _vp, _vp]),
    "lfs_trainer_set_profile": (C.c_int, [_vp, C.c_int]),
    "lfs_trainer_get_profile": (C.c_int, [_vp, C.POINTER(_f), C.POINTER(C.c_int)]),
    "lfs_trainer_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _vp]),
    "lfs_adam_p2p_owner": (C.c_int, [C.c_int64, C.c_int]),
    "lfs_adam_p2p_owned_chunks": (C.c_int, [C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64),
                                            C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "lfs_arena_gather": (C.c_int, [_vp] * 8 + [_u32, _u32, _u32, C.c_uint64, C.c_uint64, _vp]),
    "lfs_arena_set_rows": (C.c_int, [_vp, C.c_uint64, _u32, _u32, _vp, _vp, _u32, _u32, _vp]),
    "lfs_arena_get_rows": (C.c_int, [_vp, C.c_uint64, _u32, _u32, _vp, _vp, _u32, _u32, _vp]),
    "lfs_adam_p2p_zero_unowned": (C.c_int, [_vp, _i64, C.c_int, C.c_int, _vp]),
    "lfs_trainer_poll_capacity": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "lfs_trainer_debug_copy": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, C.POINTER(C.c_uint64), _vp]),
}

_lib = None
_lock = threading.Lock()


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into liblfs_b200.so (in-tree)."""
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building liblfs_b200.so failed")
    return LIB_PATH


def load() -> C.CDLL:
    """Load the library (once). Raises if it has not been built -- there is no CPU fallback."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the product path has no fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.lfs_abi_version() != ABI_VERSION:
            raise RuntimeError("lfs_b200 ABI version mismatch")
        # tuning switches for A/B measurements, e.g. LFS_OPTIONS="fwd_variant=1,bwd_variant=1" (see lfs_set_option)
        for item in filter(None, os.environ.get("LFS_OPTIONS", "").split(",")):
            key, _, val = item.partition("=")
            if lib.lfs_set_option(key.strip().encode(), int(val or "1")) != LFS_OK:
                raise ValueError(f"LFS_OPTIONS: {lib.lfs_last_error().decode()}")
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc == LFS_OK:
        return
    msg = load().lfs_last_error().decode("utf-8", "replace")
    if rc == LFS_ERR_UNSUPPORTED:
        raise LfsUnsupported(rc, msg)
    raise LfsError(rc, msg)
