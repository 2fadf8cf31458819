"""ctypes binding of include/ugaitnet_b200.h.  Tensors cross by DLPack.

A torch CUDA tensor is exported once with ``tensor.__dlpack__()``; the capsule's
``DLManagedTensor*`` starts with the ``DLTensor`` the C ABI expects (``ugn_tensor`` is
layout-identical), so we pass that pointer straight through.  The capsule is kept alive by
the :class:`TRef` wrapper; the library never calls the deleter.

There is no CPU fallback: if the shared library is missing the import fails loudly, and a
non-CUDA tensor is rejected by the library with UGN_ERR_DEVICE.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_void_p

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libugaitnet_b200.so")


class UgnError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m ugaitnet_b200.build` "
            "(ugaitnet_b200 has no CPU / PyTorch fallback path)")
    return ctypes.CDLL(LIB_PATH)


lib = _load()

_T = c_void_p  # const ugn_tensor*

_PROTOS = {
    "ugn_abi_version": (c_int, []),
    "ugn_last_error": (c_char_p, []),
    "ugn_ctx_create": (c_int, [c_int, POINTER(c_void_p)]),
    "ugn_ctx_destroy": (c_int, [c_void_p]),
    "ugn_ctx_check": (c_int, [c_void_p]),
    "ugn_ctx_has_tcgen05": (c_int, [c_void_p]),
    "ugn_launch_count": (c_int64, [c_void_p]),
    "ugn_pack_input": (c_int, [c_void_p, _T, _T, c_void_p]),
    "ugn_pack_input_expand": (c_int, [c_void_p, _T, _T, _T, _T, c_float, _T, c_void_p]),
    "ugn_pack_input_augment": (c_int, [c_void_p, _T, _T, _T, _T, _T, _T, c_float, c_float, c_float, c_float, _T, c_void_p]),
    "ugn_conv3d_fwd": (c_int, [c_void_p, _T, _T, _T, _T, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "ugn_conv3d_wgrad": (c_int, [c_void_p, _T, _T, _T, _T, c_int, c_int, c_int, c_void_p]),
    "ugn_conv3d_dgrad": (c_int, [c_void_p, _T, _T, _T, c_int, c_int, c_int, c_void_p]),
    "ugn_decode_samples": (c_int, [c_void_p, _T, c_float, c_float, c_float, c_float, c_float, _T, c_void_p]),
    "ugn_pack_weight": (c_int, [c_void_p, _T, _T, c_void_p]),
    "ugn_split_bf16": (c_int, [c_void_p, _T, _T, c_void_p]),
    "ugn_conv2d_fwd": (c_int, [c_void_p, _T, _T, _T, _T, _T, c_int, c_float, c_int, c_void_p]),
    "ugn_conv2d_bwd_act": (c_int, [c_void_p, _T, _T, _T, _T, _T, c_int, c_float, c_int, c_void_p]),
    "ugn_conv2d_dgrad": (c_int, [c_void_p, _T, _T, _T, c_void_p]),
    "ugn_conv2d_wgrad": (c_int, [c_void_p, _T, _T, _T, _T, c_void_p]),
    "ugn_flatten_chw": (c_int, [c_void_p, _T, _T, c_void_p]),
    "ugn_unflatten_chw": (c_int, [c_void_p, _T, _T, c_void_p]),
    "ugn_linear_fwd": (c_int, [c_void_p, _T, _T, _T, _T, _T, _T, c_int, c_float, c_void_p]),
    "ugn_act_mask_bwd": (c_int, [c_void_p, _T, _T, _T, _T, _T, c_int, c_float, c_void_p]),
    "ugn_linear_bwd": (c_int, [c_void_p, _T, _T, _T, _T, _T, _T, c_void_p]),
    "ugn_fuse_fc1_fwd": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), _T, _T, _T, _T, c_int, c_int,
                                 _T, _T, _T, _T, _T, c_int, c_float, c_void_p]),
    "ugn_fuse_fwd": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), _T, _T, _T, _T,
                             c_int, c_int, c_void_p]),
    "ugn_fuse_bwd": (c_int, [c_void_p, c_int, _T, _T, _T, _T, POINTER(c_void_p), POINTER(c_void_p),
                             c_int, c_int, c_void_p]),
    "ugn_softmax_ce": (c_int, [c_void_p, _T, _T, _T, _T, c_float, c_void_p]),
    "ugn_softmax_ce_ls": (c_int, [c_void_p, _T, _T, _T, _T, c_float, c_float, c_void_p]),
    "ugn_triplet_workspace_bytes": (c_int64, [c_int, c_int]),
    "ugn_triplet_all": (c_int, [c_void_p, _T, _T, c_float, c_float, _T, _T, _T, c_void_p]),
    "ugn_triplet_all_tc": (c_int, [c_void_p, _T, _T, _T, c_float, c_float, _T, _T, _T, c_void_p]),
    "ugn_triplet_hard": (c_int, [c_void_p, _T, _T, _T, c_float, c_float, _T, _T, _T, c_void_p]),
    "ugn_pair_verif_loss": (c_int, [c_void_p, _T, _T, c_float, c_float, _T, _T, _T, c_void_p]),
    "ugn_adam_step": (c_int, [c_void_p, _T, _T, _T, _T, _T, _T, c_float, c_float, c_float, c_float,
                              c_float, _T, _T, _T, c_int, c_int, c_void_p]),
    "ugn_adam_step_ex": (c_int, [c_void_p, _T, _T, _T, _T, _T, c_float, _T, _T, c_float, c_float, c_float, c_float,
                                 c_float, _T, _T, _T, c_int, c_int, c_void_p]),
    "ugn_dp_optim_step": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int64), POINTER(c_int64), c_int64, c_int64,
                                  _T, _T, _T, _T, _T,
                                  c_float, _T, _T, c_float, c_float, c_float, _T, POINTER(c_int64), _T,
                                  _T, POINTER(c_int64), c_int, _T, c_int, c_int, POINTER(c_int64), c_int64, c_void_p]),
    "ugn_sgd_step": (c_int, [c_void_p, _T, _T, _T, _T, _T, c_float, c_float, c_float, _T, _T, _T, c_int, c_int,
                             c_void_p]),
    "ugn_knn_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64, c_int]),
    "ugn_knn_gallery_norms": (c_int, [c_void_p, _T, _T, _T, c_void_p]),
    "ugn_knn_pack": (c_int, [c_void_p, _T, _T, c_void_p]),
    "ugn_knn_topk_tc": (c_int, [c_void_p, _T, _T, _T, _T, _T, _T, _T, c_int, c_int64, _T, _T, _T, _T, _T, c_void_p]),
    "ugn_knn_topk": (c_int, [c_void_p, _T, _T, _T, _T, c_int, c_int64, _T, _T, _T, _T, c_void_p]),
    "ugn_knn_merge_vote": (c_int, [c_void_p, _T, _T, _T, c_int, _T, _T, _T, _T, c_void_p]),
    "ugn_knn_merge_vote_packed": (c_int, [c_void_p, _T, ctypes.c_longlong, c_int, _T, _T, _T, _T, c_void_p]),
    "ugn_segment_pool": (c_int, [c_void_p, _T, _T, _T, c_int, _T, c_void_p]),
    "ugn_segment_mode": (c_int, [c_void_p, _T, _T, _T, c_int, _T, c_void_p]),
    "ugn_gs_conv1_fwd": (c_int, [c_void_p, _T, _T, _T, c_float, c_void_p]),
    "ugn_gs_conv1_wgrad": (c_int, [c_void_p, _T, _T, _T, _T, c_float, c_void_p]),
    "ugn_pad_hw": (c_int, [c_void_p, _T, _T, c_void_p]),
    "ugn_crop_hw": (c_int, [c_void_p, _T, _T, c_int, c_void_p]),
    "ugn_setmax_fwd": (c_int, [c_void_p, _T, c_int, _T, _T, _T, c_void_p]),
    "ugn_setmax_bwd": (c_int, [c_void_p, _T, _T, _T, c_int, _T, c_int, c_void_p]),
    "ugn_hpp_fwd": (c_int, [c_void_p, _T, c_int, _T, c_void_p]),
    "ugn_hpp_bwd": (c_int, [c_void_p, _T, _T, c_int, _T, c_int, c_void_p]),
    "ugn_bmm_f32": (c_int, [c_void_p, _T, c_int, _T, c_int, _T, c_void_p]),
    "ugn_fuse3_fwd": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), _T, _T, _T, c_int, c_void_p]),
    "ugn_fuse3_bwd": (c_int, [c_void_p, c_int, _T, _T, _T, _T, POINTER(c_void_p), POINTER(c_void_p), c_int, c_void_p]),
    "ugn_permute102": (c_int, [c_void_p, _T, _T, c_void_p]),
    "ugn_gemm_bf16": (c_int, [c_void_p, _T, c_int, _T, c_int, _T, c_int, c_void_p]),
    "ugn_grad_scale_update": (c_int, [c_void_p, _T, c_float, c_void_p]),
    "ugn_grad_scale_set": (c_int, [c_void_p, c_float, c_void_p]),
    "ugn_set_fwd_passes": (c_int, [c_void_p, c_int, c_int]),
    "ugn_colsum": (c_int, [c_void_p, _T, _T, c_void_p]),
    "ugn_linear_bwd_ex": (c_int, [c_void_p, _T, _T, _T, _T, _T, _T, _T, _T, _T, c_void_p]),
    "ugn_dropout_advance": (c_int, [c_void_p, _T, c_void_p]),
    "ugn_dropout_mask": (c_int, [c_void_p, _T, c_int, c_float, _T, c_void_p]),
    "ugn_linear_fwd_philox": (c_int, [c_void_p, _T, _T, _T, _T, c_int, c_float, _T, _T, c_int, c_float, c_void_p]),
    "ugn_linear_bwd_philox": (c_int, [c_void_p, _T, _T, _T, _T, _T, c_int, c_float, _T, _T, _T, _T, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)

for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(lib, _name)      # AttributeError here == header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args

ctypes.pythonapi.PyCapsule_GetPointer.restype = c_void_p
ctypes.pythonapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, c_char_p]


class TRef:
    """Keeps a DLPack export of a torch tensor alive and exposes its DLTensor* address."""
    __slots__ = ("t", "cap", "ptr")

    def __init__(self, t: torch.Tensor):
        if not isinstance(t, torch.Tensor):
            raise TypeError("TRef expects a torch.Tensor")
        if not t.is_contiguous():
            raise ValueError("ugaitnet_b200 tensors must be contiguous")
        self.t = t
        self.cap = t.__dlpack__()
        self.ptr = ctypes.pythonapi.PyCapsule_GetPointer(self.cap, b"dltensor")


def ref(t):
    """torch.Tensor | TRef | None -> DLTensor* (c_void_p value)."""
    if t is None:
        return None
    if isinstance(t, TRef):
        return t.ptr
    return TRef(t)


def _p(x):
    if x is None:
        return None
    if isinstance(x, TRef):
        return x.ptr
    raise TypeError("expected TRef or None")


def check(rc: int):
    if rc != 0:
        msg = lib.ugn_last_error()
        raise UgnError(f"ugaitnet_b200 error {rc}: {msg.decode() if msg else '?'}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class Ctx:
    """One ugn_ctx per GPU per process."""

    def __init__(self, device: int = 0):
        h = c_void_p()
        check(lib.ugn_ctx_create(int(device), ctypes.byref(h)))
        self.h = h
        self.device = int(device)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.ugn_ctx_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def check(self):
        """Synchronise and raise if any kernel reported an asynchronous failure."""
        check(lib.ugn_ctx_check(self.h))

    @property
    def launches(self) -> int:
        return int(lib.ugn_launch_count(self.h))

    @property
    def has_tcgen05(self) -> bool:
        return bool(lib.ugn_ctx_has_tcgen05(self.h))


def ptr_array(refs):
    arr = (c_void_p * len(refs))()
    for i, r in enumerate(refs):
        arr[i] = r.ptr
    return arr
