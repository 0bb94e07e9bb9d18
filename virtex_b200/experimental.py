"""Opt-in experimental kernels (libvirtex_b200_x.so, include/virtex_b200_x.h).

Nothing here is on the default path.  A feature is used only when the environment variable `VTX_EXPERIMENTAL` names it
(comma separated, or `all`); its GPU tests are skipped otherwise.  Features:

  stem_s2d   the 7x7/2 stem conv as a 4-tap implicit GEMM over a space-to-depth view of the image (vtx_gemm_x
             conv_mode 5 / 6) instead of im2col + GEMM: ~2 GB less HBM traffic per step and no 1 GB im2col buffer.

  pdl        programmatic dependent launch: the library compiled with -DVTX_PDL (libvirtex_b200_pdl.so) is loaded
             instead of libvirtex_b200.so (virtex_b200/lib.py); every kernel triggers its dependents at entry and the
             GEMM -- launched with the programmatic-serialisation attribute -- overlaps its prologue with the
             previous kernel's tail, then `griddepcontrol.wait`s.

  head_x     register-accumulating LayerNorm / embedding backward kernels (head.cu compiled with -DVTX_HEAD_X): the
             entry points vtx_ln_bwd / vtx_embed_bwd / vtx_colsum (bias gradients: row lanes reduced in
             shared memory, 4-40x fewer atomics) / vtx_cross_entropy (row held in registers: one read pass) are routed to
             libvirtex_b200_x.so.

  gemm_x     every vtx_gemm call goes to vtx_gemm_x (gemm_tc.cu compiled with -DVTX_GEMM_X): BN-statistics pass of
             64- / 128-wide tiles spread over all 256 epilogue threads (4 / 8 rows each instead of 16 rows on a
             quarter / half of the threads).

  backbone_x shared-memory tiled stem max-pool (backbone.cu compiled with -DVTX_BACKBONE_X).  Backward: pooled gradients
             and argmax slots are staged once per CTA instead of being gathered from L2 up to nine times; forward:
             BN + ReLU applied once per input element into shared memory, pooled from there.  im2col3x3 /
             col2im3x3 / subsample / upsample_add with 32-bit index arithmetic (the validated kernels are bound by 64-bit integer division).

Validation procedure on a B200: `VTX_EXPERIMENTAL=all python -m pytest tests -m gpu -q` and
`VTX_EXPERIMENTAL=all python bench.py`; then move the kernels into the main library.
"""
import ctypes
import os

import torch

from . import lib as L
from . import ops

FEATURES = ("stem_s2d", "pdl", "head_x", "gemm_x", "backbone_x")
# entry points of the MAIN ABI that libvirtex_b200_x.so re-implements (same signature); routed there by ops._get when
# the feature is enabled
ROUTED = {"vtx_ln_bwd": "head_x", "vtx_embed_bwd": "head_x", "vtx_colsum": "head_x", "vtx_cross_entropy": "head_x",
          "vtx_gemm": "gemm_x", "vtx_maxpool_bwd": "backbone_x", "vtx_bn_relu_maxpool": "backbone_x",
          "vtx_im2col3x3": "backbone_x", "vtx_col2im3x3": "backbone_x", "vtx_subsample": "backbone_x",
          "vtx_upsample_add": "backbone_x"}
_ROUTED_SYMBOL = {"vtx_gemm": "vtx_gemm_x"}  # where the name differs in the experimental library
_P, _I = ctypes.c_void_p, ctypes.c_int
_PROTOS = {
    "vtx_gemm_x": [_P, _P],
    "vtx_x_stem_s2d": [_P, _P, _I, _I, _I, _P],
    "vtx_x_stem_w_pack": [_P, _P, _I, _P],
    "vtx_x_stem_w_unpack_add": [_P, _P, _I, _P],
}
_lib = None


def enabled(feature: str) -> bool:
    assert feature in FEATURES, feature
    v = os.environ.get("VTX_EXPERIMENTAL", "")
    return v == "all" or feature in [t.strip() for t in v.split(",")]


def any_enabled() -> bool:
    return any(enabled(f) for f in FEATURES)


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvirtex_b200_x.so")


def load():
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise L.VtxError(f"{path} is missing: run `python -m virtex_b200.build`")
        lib = ctypes.CDLL(path)
        for name, argtypes in _PROTOS.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = argtypes, ctypes.c_int
        lib.vtx_last_error.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_PROTOS)


def routed_lib(name):
    """The experimental library if entry point `name` is re-implemented there and its feature is enabled, else None."""
    feature = ROUTED.get(name)
    if feature is None or not enabled(feature):
        return None
    return load()


def routed_symbol(name) -> str:
    return _ROUTED_SYMBOL.get(name, name)


def last_error() -> str:
    return load().vtx_last_error().decode("utf-8", "replace")


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise L.VtxError(f"{name} failed ({rc}): {lib.vtx_last_error().decode()}")
    ops.launch_count += 1


_gemm_struct = L.VtxGemm()


def gemm(A, B, D, M, N, K, *, lda, ldb, ldd=None, stats=None, atomic=False, split_k=1, conv=None, conv_mode=0,
         out_f32=None):
    """vtx_gemm_x: same contract as ops.gemm plus conv_mode 5 / 6 (stem conv over the space-to-depth view)."""
    g = _gemm_struct
    g.A, g.B, g.D = A.data_ptr(), B.data_ptr(), D.data_ptr()
    g.bias, g.residual, g.stats = 0, 0, (0 if stats is None else stats.data_ptr())
    g.lda, g.ldb = lda, ldb
    g.ldd = D.stride(0) if ldd is None else ldd
    g.ldr = 0
    g.M, g.N, g.K = M, N, K
    g.a_mn = g.b_mn = 0
    g.out_f32 = int(D.dtype == torch.float32) if out_f32 is None else int(out_f32)
    g.atomic, g.act, g.split_k, g.tile_n = int(atomic), 0, split_k, 0
    g.alpha = 1.0
    g.conv_n, g.conv_h, g.conv_w, g.conv_c = conv if conv is not None else (0, 0, 0, 0)
    g.conv_mode = conv_mode
    call("vtx_gemm_x", ctypes.addressof(g), ops._stream())
