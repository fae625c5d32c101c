"""Host-side mirror of the reference's operator surface for the hot path, over the C ABI.

Reference item (crates/ring/src/...)                         -> here
  cyclotomic_ring/ring_config.rs:11-35  CyclotomicConfig      -> RingConfig.{crt_in_place, icrt_in_place, crt, icrt}
  cyclotomic_ring/crt.rs:6-50           CRT / ICRT traits     -> RqPoly.crt / RqNTT.icrt, CRT.elementwise_crt, ICRT.elementwise_icrt
  cyclotomic_ring/coeff_form.rs:31-33   RqPoly ([Fp; D])      -> RqPoly (batch of n >= 1 elements over one flat limb buffer)
  cyclotomic_ring/ntt_form.rs:25-27     RqNTT                 -> RqNTT
  ntt_form.rs:159-189,521-550           Mul / MulUnchecked    -> RqNTT.__mul__ / mul_unchecked (slot-wise)
  coeff_form.rs:54-67,250-258           Mul (poly_mul+reduce) -> RqPoly.__mul__ (fused CRT -> mul -> ICRT kernel)
  cyclotomic_ring/flatten.rs:10-34      Flatten               -> flatten_to_coeffs / promote_from_coeffs
  models/{goldilocks,babybear,stark_prime}/mod.rs             -> GoldilocksRingConfig, BabyBearRingConfig, StarkRingConfig

Buffers are the reference's raw memory: little-endian u64 Montgomery limbs.  A buffer is either a
host numpy.uint64 array or a CUDA torch tensor (int64 / uint64 bit patterns); host buffers are
copied through the library's pipelined H2D/D2H path, device buffers are transformed in place.
Everything is computed by libstarkrings_cuda.so; there is no CPU implementation here.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib as L
from .errors import LengthPanic, StarkRingsError

try:  # torch is plumbing only (device memory, streams); host-only use works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


class Context:
    """Owns an sr_ctx (one CUDA device).  Raises if no B200 is usable: there is no fallback."""

    def __init__(self, device: int = 0):
        h = ctypes.c_void_p()
        rc = L.lib.sr_init(device, ctypes.byref(h))
        if rc != L.SR_OK:
            raise StarkRingsError("sr_init(device=%d) failed with status %d: no usable sm_100 CUDA device "
                                  "(libstarkrings_cuda.so has no CPU fallback)" % (device, rc))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            L.lib.sr_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, what=""):
        if rc == L.SR_OK:
            return
        msg = L.lib.sr_last_error(self.h).decode()
        if rc == L.SR_ERR_BAD_LENGTH:
            raise LengthPanic("%s: %s" % (what, msg))
        raise StarkRingsError("%s failed (status %d): %s" % (what, rc, msg))

    def sync(self):
        self.check(L.lib.sr_sync(self.h), "sr_sync")

    def use_torch_stream(self):
        """Enqueue device-resident calls on torch's current stream: they are then ordered with the
        surrounding torch operations and visible to torch.cuda.Event.  Called automatically whenever
        a CUDA tensor is handed to this package."""
        s = torch.cuda.current_stream(self.device).cuda_stream
        if s != getattr(self, "_stream", -1):
            self.check(L.lib.sr_set_stream(self.h, ctypes.c_void_p(s)), "sr_set_stream")
            self._stream = s

    def set_pipelined(self, on: bool = True):
        """Successive mat-vec / commitment kernels of this context may overlap (programmatic dependent launch): the
        column loop of a product starts while the previous one is still in its tail.  For resident inputs."""
        self.check(L.lib.sr_set_pipelined(self.h, 1 if on else 0), "sr_set_pipelined")

    @property
    def kernel_launches(self) -> int:
        return int(L.lib.sr_kernel_launches(self.h))

    def imad_peak(self):
        """Measured integer multiply-add peaks of this device in 10^12 op/s: (IMAD, IMAD.WIDE.U32, IMAD.WIDE.U32.X)."""
        t = (ctypes.c_double * 3)()
        self.check(L.lib.sr_imad_peak(self.h, t), "sr_imad_peak")
        return float(t[0]), float(t[1]), float(t[2])

    def timer_start(self):
        self.check(L.lib.sr_timer_start(self.h), "sr_timer_start")

    def timer_stop(self) -> float:
        ms = ctypes.c_float()
        self.check(L.lib.sr_timer_stop(self.h, ctypes.byref(ms)), "sr_timer_stop")
        return float(ms.value)


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _ptr_loc(buf):
    """(pointer, n_limbs, loc, device) of a flat limb buffer."""
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.uint64 or not buf.flags["C_CONTIGUOUS"] or buf.ndim != 1:
            raise TypeError("host buffers must be 1-D contiguous numpy.uint64 arrays")
        return ctypes.c_void_p(buf.ctypes.data), buf.size, L.SR_HOST, None
    if torch is not None and isinstance(buf, torch.Tensor):
        if buf.dim() != 1 or not buf.is_contiguous() or buf.element_size() != 8:
            raise TypeError("tensor buffers must be 1-D contiguous 64-bit tensors")
        if buf.is_cuda:
            return ctypes.c_void_p(buf.data_ptr()), buf.numel(), L.SR_DEVICE, buf.device.index
        return ctypes.c_void_p(buf.data_ptr()), buf.numel(), L.SR_HOST, None
    raise TypeError("unsupported buffer type %r" % type(buf))


class RingConfig:
    """CyclotomicConfig<N> (ring_config.rs:11-35) for one prime."""

    def __init__(self, name, tag, ring_id, D, N, crt_ext_degree, modulus):
        self.name, self.tag, self.ring_id = name, tag, ring_id
        self.D, self.N = D, N
        self.CRT_FIELD_EXTENSION_DEGREE = crt_ext_degree
        self.modulus = modulus
        self.limbs = D * N  # u64 limbs per ring element

    def __repr__(self):
        return "<RingConfig %s D=%d N=%d>" % (self.name, self.D, self.N)

    def _ctx(self, device, ctx):
        c = ctx or default_context(0 if device is None else device)
        if device is not None:
            c.use_torch_stream()
        return c

    def _unary(self, fn_name, buf, single, ctx):
        p, n, loc, dev = _ptr_loc(buf)
        if single and n != self.limbs:  # assert_eq!(coefficients.len(), D)
            raise LengthPanic("%s: slice of %d limbs, expected %d" % (fn_name, n, self.limbs))
        c = self._ctx(dev, ctx)
        fn = getattr(L.lib, "sr_%s_%s" % (self.tag, fn_name))
        c.check(fn(c.h, p, n, loc), fn_name)
        return buf

    # -- CyclotomicConfig ------------------------------------------------------------------
    def crt_in_place(self, coefficients, ctx=None):
        """ring_config.rs:27: one element (D field elements) in place; wrong length panics."""
        return self._unary("crt_batch", coefficients, True, ctx)

    def icrt_in_place(self, evaluations, ctx=None):
        """ring_config.rs:34."""
        return self._unary("icrt_batch", evaluations, True, ctx)

    def crt(self, coefficients, ctx=None):
        out = coefficients.copy() if isinstance(coefficients, np.ndarray) else coefficients.clone()
        return self.crt_in_place(out, ctx)

    def icrt(self, evaluations, ctx=None):
        out = evaluations.copy() if isinstance(evaluations, np.ndarray) else evaluations.clone()
        return self.icrt_in_place(out, ctx)

    # -- batch forms (crt.rs:10-25, 34-49) ------------------------------------------------
    def crt_batch(self, buf, ctx=None):
        return self._unary("crt_batch", buf, False, ctx)

    def icrt_batch(self, buf, ctx=None):
        return self._unary("icrt_batch", buf, False, ctx)

    def ntt_mul_batch(self, a_inout, b, ctx=None):
        pa, na, loc, dev = _ptr_loc(a_inout)
        pb, nb, locb, _ = _ptr_loc(b)
        if na != nb or loc != locb:
            raise LengthPanic("ntt_mul: operands differ in length or location")
        c = self._ctx(dev, ctx)
        fn = getattr(L.lib, "sr_%s_ntt_mul_batch" % self.tag)
        c.check(fn(c.h, pa, pb, na, loc), "ntt_mul_batch")
        return a_inout

    def _addsub(self, which, a_inout, b, ctx=None):
        pa, na, loc, dev = _ptr_loc(a_inout)
        c = self._ctx(dev, ctx)
        if which == "neg":
            c.check(L.lib.sr_neg_batch(c.h, self.ring_id, pa, na, loc), "sr_neg_batch")
            return a_inout
        pb, nb, locb, _ = _ptr_loc(b)
        if na != nb or loc != locb:
            raise LengthPanic("%s: operands differ in length or location" % which)
        fn = L.lib.sr_add_batch if which == "add" else L.lib.sr_sub_batch
        c.check(fn(c.h, self.ring_id, pa, pb, na, loc), "sr_%s_batch" % which)
        return a_inout

    def add_batch(self, a_inout, b, ctx=None):
        """a[i] += b[i] (ntt_form.rs:588-601 / coeff_form.rs Add: field element by field element in either form)."""
        return self._addsub("add", a_inout, b, ctx)

    def sub_batch(self, a_inout, b, ctx=None):
        return self._addsub("sub", a_inout, b, ctx)

    def neg_batch(self, a_inout, ctx=None):
        return self._addsub("neg", a_inout, None, ctx)

    def sum_batch(self, buf, ctx=None):
        """Sum of a slice of elements (ntt_form.rs:640-654: fold from ZERO with Add) -> one element."""
        p, n, loc, dev = _ptr_loc(buf)
        out = (np.empty(self.limbs, dtype=np.uint64) if isinstance(buf, np.ndarray)
               else torch.empty(self.limbs, dtype=buf.dtype, device=buf.device))
        c = self._ctx(dev, ctx)
        c.check(L.lib.sr_sum_batch(c.h, self.ring_id, p, n, _ptr_loc(out)[0], loc), "sr_sum_batch")
        return out

    def reduce_batch(self, polys, coeffs_per_poly, ctx=None):
        """CyclotomicConfig::reduce_in_place on a batch (ring_config.rs:23): polynomials of coeffs_per_poly field
        elements (D <= coeffs_per_poly <= 2D) -> coefficient-form ring elements (new buffer)."""
        p, n, loc, dev = _ptr_loc(polys)
        per = coeffs_per_poly * self.N
        if coeffs_per_poly < self.D or coeffs_per_poly > 2 * self.D or n % per:
            raise LengthPanic("reduce: %d limbs is not a batch of polynomials of %d coefficients" % (n, coeffs_per_poly))
        cnt = n // per
        out = (np.empty(cnt * self.limbs, dtype=np.uint64) if isinstance(polys, np.ndarray)
               else torch.empty(cnt * self.limbs, dtype=polys.dtype, device=polys.device))
        c = self._ctx(dev, ctx)
        c.check(L.lib.sr_reduce_batch(c.h, self.ring_id, p, n, coeffs_per_poly, _ptr_loc(out)[0], loc), "reduce_batch")
        return out

    def rot_batch(self, a, ctx=None):
        """Cyclotomic::rot on a batch: every element multiplied by X (new buffer)."""
        p, n, loc, dev = _ptr_loc(a)
        if n % self.limbs:
            raise LengthPanic("rot: buffer is not a whole number of ring elements")
        out = np.empty_like(a) if isinstance(a, np.ndarray) else torch.empty_like(a)
        c = self._ctx(dev, ctx)
        c.check(L.lib.sr_rot_batch(c.h, self.ring_id, p, _ptr_loc(out)[0], n, loc), "rot_batch")
        return out

    def _gadget(self, fn, buf, b, padding_size, grow, ctx):
        p, n, loc, dev = _ptr_loc(buf)
        per = self.limbs if grow else self.limbs * padding_size
        if padding_size < 1 or n % per:
            raise LengthPanic("gadget (de/re)composition: buffer is not a whole number of elements")
        cnt = (n // per) * (padding_size if grow else 1) * self.limbs
        out = (np.empty(cnt, dtype=np.uint64) if isinstance(buf, np.ndarray)
               else torch.empty(cnt, dtype=buf.dtype, device=buf.device))
        c = self._ctx(dev, ctx)
        c.check(fn(c.h, self.ring_id, p, n, b & 0xFFFFFFFFFFFFFFFF, b >> 64, padding_size, _ptr_loc(out)[0], loc),
                "gadget_decompose" if grow else "gadget_recompose")
        return out

    def gadget_decompose(self, buf, b, padding_size, ctx=None):
        """GadgetDecompose for &[RqPoly] (balanced_decomposition/mod.rs:163-175): n coefficient-form elements ->
        n * padding_size digit elements (digits in [-b/2, b/2]).  Too small a padding_size raises LengthPanic
        (the reference panics)."""
        return self._gadget(L.lib.sr_gadget_decompose, buf, b, padding_size, True, ctx)

    def gadget_recompose(self, buf, b, padding_size, ctx=None):
        """GadgetRecompose for &[R] (mod.rs:177-190)."""
        return self._gadget(L.lib.sr_gadget_recompose, buf, b, padding_size, False, ctx)

    def ring_mul_batch(self, a, b, out=None, ctx=None):
        pa, na, loc, dev = _ptr_loc(a)
        pb, nb, locb, _ = _ptr_loc(b)
        if out is None:
            out = np.empty_like(a) if isinstance(a, np.ndarray) else torch.empty_like(a)
        po, no, loco, _ = _ptr_loc(out)
        if not (na == nb == no) or not (loc == locb == loco):
            raise LengthPanic("ring_mul: operands differ in length or location")
        c = self._ctx(dev, ctx)
        fn = getattr(L.lib, "sr_%s_ring_mul_batch" % self.tag)
        c.check(fn(c.h, pa, pb, po, na, loc), "ring_mul_batch")
        return out


Fq_GOLDILOCKS = 18446744069414584321
Fq_BABYBEAR = 2013265921
Fq_STARK = 3618502788666131213697322783095070105623107215331596699973092056135872020481

GoldilocksRingConfig = RingConfig("goldilocks", "gl", L.SR_GOLDILOCKS, 24, 1, 3, Fq_GOLDILOCKS)
BabyBearRingConfig = RingConfig("babybear", "bb", L.SR_BABYBEAR, 72, 1, 9, Fq_BABYBEAR)
StarkRingConfig = RingConfig("stark_prime", "sp", L.SR_STARK, 16, 4, 1, Fq_STARK)
CONFIGS = {c.name: c for c in (GoldilocksRingConfig, BabyBearRingConfig, StarkRingConfig)}
CONFIGS.update({c.tag: c for c in list(CONFIGS.values())})


class _RqBase:
    """A batch of n ring elements over one flat limb buffer (n = 1: a single element)."""

    FORM = None

    def __init__(self, config: RingConfig, data, ctx=None):
        _, n, _, _ = _ptr_loc(data)
        if n % config.limbs:
            raise LengthPanic("buffer of %d limbs is not a whole number of %s elements (%d limbs each)"
                              % (n, config.name, config.limbs))
        self.config, self.data, self.ctx = config, data, ctx

    def __len__(self):
        return _ptr_loc(self.data)[1] // self.config.limbs

    def dimension(self) -> int:
        """PolyRing::dimension (coeff_form.rs:539-566, ntt_form.rs:672-697): the const generic D of the type, i.e. the
        number of base-field coefficients per element; here a property of the element's ring configuration."""
        return self.config.D

    # -- Add / Sub / Neg / Sum (ntt_form.rs:588-626, 640-654 and the same operators of coeff_form.rs) -----------------
    def __add__(self, rhs):
        self._same(rhs)
        out = self.clone()
        self.config.add_batch(out.data, rhs.data, self.ctx)
        return out

    def __iadd__(self, rhs):
        self._same(rhs)
        self.config.add_batch(self.data, rhs.data, self.ctx)
        return self

    def __sub__(self, rhs):
        self._same(rhs)
        out = self.clone()
        self.config.sub_batch(out.data, rhs.data, self.ctx)
        return out

    def __isub__(self, rhs):
        self._same(rhs)
        self.config.sub_batch(self.data, rhs.data, self.ctx)
        return self

    def __neg__(self):
        out = self.clone()
        self.config.neg_batch(out.data, self.ctx)
        return out

    def sum(self):
        """Sum over the batch -> a batch of one element."""
        return type(self)(self.config, self.config.sum_batch(self.data, self.ctx), self.ctx)

    def clone(self):
        d = self.data.copy() if isinstance(self.data, np.ndarray) else self.data.clone()
        return type(self)(self.config, d, self.ctx)

    def flatten_to_coeffs(self):
        """flatten.rs:10-18: the same allocation viewed as base-field limbs."""
        return self.data

    @classmethod
    def promote_from_coeffs(cls, config, flat, ctx=None):
        """flatten.rs:20-34: None when the length is not a multiple of the dimension."""
        if _ptr_loc(flat)[1] % config.limbs:
            return None
        return cls(config, flat, ctx)

    def _same(self, other):
        if type(other) is not type(self) or other.config is not self.config:
            raise TypeError("operands must be the same ring type")

    # -- CanonicalSerialize / CanonicalDeserialize (coeff_form.rs:154-189, ntt_form.rs:24), batched on the device ------
    def serialized_size(self) -> int:
        """serialized_size (coeff_form.rs:165-167) summed over the batch (no Vec length prefix)."""
        return int(L.lib.sr_serialized_bytes(self.config.ring_id, len(self)))

    def serialize(self):
        """The batch as ark-serialize bytes: standard-form little-endian integers, 8 / 4 / 32 bytes per field element,
        elements back to back.  Returns a uint8 array (numpy for host buffers, CUDA tensor for device buffers)."""
        p, n, loc, dev = _ptr_loc(self.data)
        nbytes = self.serialized_size()
        if loc == L.SR_DEVICE:
            out = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.data.device)[:nbytes]
            po = ctypes.c_void_p(out.data_ptr())
        else:
            out = np.empty(nbytes, dtype=np.uint8)
            po = ctypes.c_void_p(out.ctypes.data)
        c = self.config._ctx(dev, self.ctx)
        c.check(L.lib.sr_serialize_batch(c.h, self.config.ring_id, p, n, po, loc), "sr_serialize_batch")
        return out

    @classmethod
    def deserialize(cls, config, data, ctx=None):
        """Inverse of serialize; raises StarkRingsError where ark-serialize returns SerializationError::InvalidData
        (an integer not below the modulus) and LengthPanic on a truncated buffer."""
        if isinstance(data, np.ndarray):
            if data.dtype != np.uint8 or not data.flags["C_CONTIGUOUS"]:
                raise TypeError("host bytes must be a contiguous numpy.uint8 array")
            pi, nb, loc, dev = ctypes.c_void_p(data.ctypes.data), data.size, L.SR_HOST, None
        else:
            if data.dtype != torch.uint8 or not data.is_contiguous():
                raise TypeError("device bytes must be a contiguous torch.uint8 tensor")
            pi, nb = ctypes.c_void_p(data.data_ptr()), data.numel()
            loc, dev = (L.SR_DEVICE, data.device.index) if data.is_cuda else (L.SR_HOST, None)
        per = int(L.lib.sr_serialized_bytes(config.ring_id, 1))
        n = nb // per
        if loc == L.SR_DEVICE:
            out = torch.empty(n * config.limbs, dtype=torch.int64, device=data.device)
        else:
            out = np.empty(n * config.limbs, dtype=np.uint64)
        c = config._ctx(dev, ctx)
        c.check(L.lib.sr_deserialize_batch(c.h, config.ring_id, pi, nb, _ptr_loc(out)[0] if n else None, loc),
                "sr_deserialize_batch")
        return cls(config, out, ctx)


class RqPoly(_RqBase):
    """CyclotomicPolyRingGeneral (coeff_form.rs:31-33), batched."""

    FORM = "coeff"

    def crt(self) -> "RqNTT":
        """CRT::crt / elementwise_crt (crt.rs:9-25): consumes self, reuses the allocation."""
        self.config.crt_batch(self.data, self.ctx)
        out = RqNTT(self.config, self.data, self.ctx)
        self.data = None
        return out

    def rot(self) -> "RqPoly":
        """Cyclotomic::rot (models/*/mod.rs): self * X."""
        return RqPoly(self.config, self.config.rot_batch(self.data, self.ctx), self.ctx)

    @classmethod
    def from_coeffs_vec(cls, config, polys, coeffs_per_poly, ctx=None) -> "RqPoly":
        """coeff_form.rs:36-40 / From<Vec<Fp>> (:568-578): reduce mod Phi."""
        return cls(config, config.reduce_batch(polys, coeffs_per_poly, ctx), ctx)

    def __mul__(self, rhs: "RqPoly") -> "RqPoly":
        """coeff_form.rs:250-258 (poly_mul + reduce) == icrt(crt(a) * crt(b)), one fused kernel."""
        self._same(rhs)
        return RqPoly(self.config, self.config.ring_mul_batch(self.data, rhs.data, None, self.ctx), self.ctx)


class RqNTT(_RqBase):
    """CyclotomicPolyRingNTTGeneral (ntt_form.rs:25-27), batched."""

    FORM = "ntt"

    def icrt(self) -> RqPoly:
        """ICRT::icrt / elementwise_icrt (crt.rs:33-49)."""
        self.config.icrt_batch(self.data, self.ctx)
        out = RqPoly(self.config, self.data, self.ctx)
        self.data = None
        return out

    def __mul__(self, rhs: "RqNTT") -> "RqNTT":
        """ntt_form.rs:159-175 (by value: lhs buffer is reused)."""
        self._same(rhs)
        self.config.ntt_mul_batch(self.data, rhs.data, self.ctx)
        out = RqNTT(self.config, self.data, self.ctx)
        self.data = None
        return out

    def mul_unchecked(self, rhs: "RqNTT") -> "RqNTT":
        """ntt_form.rs:177-189: identical on the GPU (no zero short-circuit exists there)."""
        return self.__mul__(rhs)

    def __imul__(self, rhs: "RqNTT"):
        self._same(rhs)
        self.config.ntt_mul_batch(self.data, rhs.data, self.ctx)
        return self


class CRT:
    """crt.rs:6-26."""

    @staticmethod
    def elementwise_crt(vec: RqPoly) -> RqNTT:
        return vec.crt()


class ICRT:
    """crt.rs:30-50."""

    @staticmethod
    def elementwise_icrt(vec: RqNTT) -> RqPoly:
        return vec.icrt()
