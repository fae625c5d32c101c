"""ctypes binding of libstarkrings_cuda.so (include/stark_rings_cuda.h).

The library is the product: there is no CPU path behind it.  Import of this module succeeds
without a GPU (so that the symbol table can be checked), but creating a Context without a
usable B200 raises, and a missing shared library raises at import.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("STARK_RINGS_LIB") or os.path.join(HERE, "libstarkrings_cuda.so")  # override: tuning builds

SR_OK, SR_ERR_BAD_LENGTH, SR_ERR_CUDA, SR_ERR_INVALID, SR_ERR_NOMEM = 0, 1, 2, 3, 4
SR_GOLDILOCKS, SR_BABYBEAR, SR_STARK = 0, 1, 2
SR_HOST, SR_DEVICE = 0, 1

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "stark_rings_b200: %s is missing -- build it with `make -C stark_rings_b200/csrc` "
        "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

_vp, _sz, _int = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
_pp = ctypes.POINTER(ctypes.c_void_p)


def _sig(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


_sig("sr_init", _int, _int, _pp)
_sig("sr_destroy", _int, _vp)
_sig("sr_last_error", ctypes.c_char_p, _vp)
_sig("sr_version", ctypes.c_char_p)
_sig("sr_set_stream", _int, _vp, _vp)
_sig("sr_reset_stream", _int, _vp)
_sig("sr_sync", _int, _vp)
_sig("sr_set_pipelined", _int, _vp, _int)
_sig("sr_elem_limbs", _sz, _int)
_sig("sr_kernel_launches", ctypes.c_uint64, _vp)
_sig("sr_dev_alloc", _int, _vp, _sz, _pp)
_sig("sr_dev_free", _int, _vp, _vp)
_sig("sr_host_alloc", _int, _vp, _sz, _pp)
_sig("sr_host_free", _int, _vp, _vp)
_sig("sr_h2d", _int, _vp, _vp, _vp, _sz)
_sig("sr_d2h", _int, _vp, _vp, _vp, _sz)
_sig("sr_timer_start", _int, _vp)
_sig("sr_timer_stop", _int, _vp, ctypes.POINTER(ctypes.c_float))
_sig("sr_imad_peak", _int, _vp, ctypes.POINTER(ctypes.c_double))
_sig("sr_crt_batch", _int, _vp, _int, _vp, _sz, _int)
_sig("sr_icrt_batch", _int, _vp, _int, _vp, _sz, _int)
_sig("sr_ntt_mul_batch", _int, _vp, _int, _vp, _vp, _sz, _int)
_sig("sr_ring_mul_batch", _int, _vp, _int, _vp, _vp, _vp, _sz, _int)
_sig("sr_add_batch", _int, _vp, _int, _vp, _vp, _sz, _int)
_sig("sr_sub_batch", _int, _vp, _int, _vp, _vp, _sz, _int)
_sig("sr_neg_batch", _int, _vp, _int, _vp, _sz, _int)
_sig("sr_sum_batch", _int, _vp, _int, _vp, _sz, _vp, _int)
_sig("sr_matvec", _int, _vp, _int, _pp, _sz, _sz, _vp, _sz, _vp, _int)
_sig("sr_matvec_partial", _int, _vp, _int, _pp, _sz, _sz, _vp, _sz, _vp, _int)
_sig("sr_modsum_partials", _int, _vp, _int, _vp, _sz, _sz, _vp, _int)
_sig("sr_reduce_batch", _int, _vp, _int, _vp, _sz, _sz, _vp, _int)
_sig("sr_rot_batch", _int, _vp, _int, _vp, _vp, _sz, _int)
_sig("sr_gadget_decompose", _int, _vp, _int, _vp, _sz, ctypes.c_uint64, ctypes.c_uint64, _sz, _vp, _int)
_sig("sr_gadget_recompose", _int, _vp, _int, _vp, _sz, ctypes.c_uint64, ctypes.c_uint64, _sz, _vp, _int)
_sig("sr_sparse_matvec", _int, _vp, _int, _sz, _sz, _vp, _vp, _vp, _vp, _sz, _vp, _int)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_szp = ctypes.POINTER(ctypes.c_size_t)
_sig("sr_sparse_matmat_symbolic", _int, _sz, _vp, _vp, _sz, _sz, _vp, _vp, _szp, _szp, _vp, _vp, _vp, _vp, _vp)
_sig("sr_sparse_matmat_values", _int, _vp, _int, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _int)
_sig("sr_matmat", _int, _vp, _int, _pp, _sz, _sz, _pp, _sz, _sz, _pp, _int)
_sig("sr_ntt_scale_batch", _int, _vp, _int, _vp, _sz, _vp, _int)
_sig("sr_mailbox_create", _int, _vp, _int, _sz, _int, _pp, ctypes.c_char_p)
_sig("sr_mailbox_open", _int, _vp, _int, _sz, _int, ctypes.c_char_p, _pp)
_sig("sr_mailbox_destroy", _int, _vp, _vp)
_sig("sr_mailbox_error", _int, _vp, _vp, ctypes.POINTER(ctypes.c_int))
_sig("sr_commit_send", _int, _vp, _int, _pp, _sz, _sz, _vp, _sz, _vp, _int, ctypes.c_uint64)
_sig("sr_commit_reduce", _int, _vp, _int, _vp, _sz, ctypes.c_uint64, _vp)
_sig("sr_commit_root", _int, _vp, _int, _pp, _sz, _sz, _vp, _sz, _vp, _int, ctypes.c_uint64, _vp)
_sig("sr_mailbox_set_timeout", _int, _vp, _vp, ctypes.c_uint64)
_sig("sr_serialized_bytes", _sz, _int, _sz)
_sig("sr_serialize_batch", _int, _vp, _int, _vp, _sz, _vp, _int)
_sig("sr_deserialize_batch", _int, _vp, _int, _vp, _sz, _vp, _int)
SR_IPC_HANDLE_BYTES = 64
for _tag in ("gl", "bb", "sp"):
    _sig("sr_%s_crt_batch" % _tag, _int, _vp, _vp, _sz, _int)
    _sig("sr_%s_icrt_batch" % _tag, _int, _vp, _vp, _sz, _int)
    _sig("sr_%s_ntt_mul_batch" % _tag, _int, _vp, _vp, _vp, _sz, _int)
    _sig("sr_%s_ring_mul_batch" % _tag, _int, _vp, _vp, _vp, _vp, _sz, _int)
    _sig("sr_%s_matvec" % _tag, _int, _vp, _pp, _sz, _sz, _vp, _sz, _vp, _int)

# every symbol include/stark_rings_cuda.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "sr_init", "sr_destroy", "sr_last_error", "sr_version", "sr_set_stream", "sr_reset_stream", "sr_sync", "sr_set_pipelined", "sr_elem_limbs",
    "sr_kernel_launches", "sr_dev_alloc", "sr_dev_free", "sr_host_alloc", "sr_host_free", "sr_h2d", "sr_d2h",
    "sr_timer_start", "sr_timer_stop", "sr_imad_peak", "sr_crt_batch", "sr_icrt_batch", "sr_ntt_mul_batch", "sr_ring_mul_batch",
    "sr_add_batch", "sr_sub_batch", "sr_neg_batch", "sr_sum_batch",
    "sr_matvec", "sr_matvec_partial", "sr_modsum_partials", "sr_reduce_batch", "sr_rot_batch",
    "sr_gadget_decompose", "sr_gadget_recompose", "sr_sparse_matvec", "sr_sparse_matmat_symbolic", "sr_sparse_matmat_values", "sr_matmat", "sr_ntt_scale_batch",
    "sr_mailbox_create", "sr_mailbox_open", "sr_mailbox_destroy", "sr_mailbox_error", "sr_mailbox_set_timeout", "sr_commit_send", "sr_commit_root",
    "sr_commit_reduce",
    "sr_serialized_bytes", "sr_serialize_batch", "sr_deserialize_batch",
] + ["sr_%s_%s" % (t, f) for t in ("gl", "bb", "sp")
     for f in ("crt_batch", "icrt_batch", "ntt_mul_batch", "ring_mul_batch", "matvec")]
