"""Host-side mirror of the reference's MSM operator surface over the C ABI.

  PreparedMsm            <-> rust_kzg_blst_sppark::prepare_multi_scalar_mult / multi_scalar_mult_prepared
                             (blst-sppark/src/lib.rs:8-38), the handle kept in SpparkPrecomputation
                             (kzg/src/msm/sppark.rs:5-22)
  g1_lincomb             <-> G1LinComb::g1_lincomb(points, scalars, len, precomputation) (kzg/src/lib.rs:142-159)
                             via blst::kzg_proofs::g1_linear_combination (blst/src/kzg_proofs.rs:25-72)
Arrays are numpy uint64 in blst layouts: scalars (n,4) Montgomery Fr, affine points (n,12), Jacobian (n,18).
"""
import ctypes as C
import numpy as np

from . import _lib

__all__ = ["PreparedMsm", "g1_lincomb", "mult_pippenger", "microbench_int", "microbench_mul", "g1_sum_device", "msm_plan"]


def _L():
    from . import lib
    return lib()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a, width):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, width)
    return a


def msm_plan(npoints, fixed=True):
    """window plan of an MSM over `npoints` bases (host-only): {"c", "c0", "W", "fold_bits"}"""
    c, c0, w, f = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _L().b200_msm_plan(npoints, int(fixed), C.byref(c), C.byref(c0), C.byref(w), C.byref(f))
    return {"c": c.value, "c0": c0.value, "W": w.value, "fold_bits": f.value}


class PreparedMsm:
    """Fixed-base MSM context resident on the GPU (prepare_msm)."""

    def __init__(self, affine_points, _borrowed_handle=None):
        self._owned = _borrowed_handle is None
        if not self._owned:                      # view of a handle owned by someone else (ShardedMsm.local_handle)
            self.h, self.npoints = _borrowed_handle, 0
            return
        pts = _u64(affine_points, 12)
        self.npoints = pts.shape[0]
        self.h = _L().prepare_msm(_p(pts), self.npoints)
        if not self.h:
            raise _lib.B200Error("prepare_msm failed (no CUDA device or out of memory); no CPU fallback")

    def info(self):
        c, w, l = C.c_int(), C.c_int(), C.c_int()
        tb = C.c_size_t()
        _L().b200_msm_info(self.h, C.byref(c), C.byref(w), C.byref(tb), C.byref(l))
        aff = bool(_L().b200_msm_last_affine(self.h))
        return {"c": c.value, "W": w.value, "table_bytes": tb.value, "launches": l.value,
                # path of the LAST run: batch-affine accumulation (6 M per addition + ~0.6 M of shared inversion work) or XYZZ
                "accumulate": "affine" if aff else "xyzz", "accumulate_kernel": "k_accumulate_affine" if aff else "k_accumulate",
                "fp_mul_per_add": 6.6 if aff else 10.0, "randomized": bool(_L().b200_msm_randomized(self.h)),
                # window width of the direct-lookup table full-length calls use instead of the bucket pipeline (0: none)
                "direct_bits": int(_L().b200_msm_direct_bits(self.h))}

    def last_counts(self):
        """(entries, tasks) of the last run: the accumulate kernel did entries - tasks mixed additions."""
        e, t = C.c_size_t(), C.c_size_t()
        _lib.check(_L().b200_msm_last_counts(self.h, C.byref(e), C.byref(t)))
        return e.value, t.value

    def last_stats(self):
        """work counters of the last run (b200_msm_last_stats) and the additions they imply, by kernel"""
        st = np.zeros(8, np.uint64)
        _lib.check(_L().b200_msm_last_stats(self.h, _p(st)))
        e, t, ne, keys, kf, nbr, D, groups = (int(v) for v in st)
        adds = {"accumulate": e - t,                      # a task's first point is a load
                "bucket_combine": t - ne,                  # partials of buckets that were cut into several tasks
                "segment_fold": (keys + ne) if kf else 0,  # per bucket: acc += run, and run += B when the bucket is non-empty
                "marginal_reduce": groups * (D + (1 if kf else 0)) * nbr}   # every (segment) sum enters D (+1) marginal trees
        return {"entries": e, "tasks": t, "nonempty_buckets": ne, "bucket_keys": keys, "fold_bits": kf, "adds": adds,
                "adds_total": sum(adds.values())}

    def mult(self, scalars):
        """multi_scalar_mult_prepared: -> Jacobian point (18 u64)."""
        sc = _u64(scalars, 4)
        out = np.zeros(18, np.uint64)
        _lib.check(_L().mult_pippenger_prepared(self.h, _p(out), sc.shape[0], _p(sc)))
        return out

    def mult_batch(self, scalars, batch):
        sc = _u64(scalars, 4)
        n = sc.shape[0] // batch
        out = np.zeros((batch, 18), np.uint64)
        _lib.check(_L().b200_msm_prepared_batch(self.h, _p(out), n, _p(sc), batch))
        return out

    def mult_device(self, out_ptr, npoints, scalars_ptr, batch=1, stream=0):
        """Device-pointer variant (asynchronous on `stream`)."""
        _lib.check(_L().b200_msm_prepared_device(self.h, C.c_void_p(out_ptr), npoints, C.c_void_p(scalars_ptr), batch,
                                                 C.c_void_p(stream)))

    def set_profiling(self, on=True):
        _L().b200_msm_set_profiling(self.h, int(on))

    def profile_read(self):
        ms, runs = C.c_double(), C.c_int()
        _lib.check(_L().b200_msm_profile_read(self.h, C.byref(ms), C.byref(runs)))
        return ms.value, runs.value

    def close(self):
        if self.h and self._owned:
            _L().b200_free_msm(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mult_pippenger(affine_points, scalars):
    """multi_scalar_mult (blst-sppark/src/lib.rs:40-62): variable-base, bases travel with the call."""
    pts, sc = _u64(affine_points, 12), _u64(scalars, 4)
    if pts.shape[0] != sc.shape[0]:
        raise ValueError("length mismatch")
    out = np.zeros(18, np.uint64)
    _lib.check(_L().mult_pippenger(_p(out), _p(pts), pts.shape[0], _p(sc)))
    return out


def g1_lincomb(affine_points, scalars, length=None, precomputation: PreparedMsm = None):
    """G1LinComb::g1_lincomb.  Unlike the reference's sppark branch (blst/src/kzg_proofs.rs:37-45) small lengths
    (< 8) also run on the device -- same group element."""
    sc = _u64(scalars, 4)
    if length is None:
        length = sc.shape[0]
    if precomputation is not None:
        return precomputation.mult(sc[:length])
    return mult_pippenger(_u64(affine_points, 12)[:length], sc[:length])


def g1_sum_device(out_ptr, points_ptr, n, stream=0):
    """sum of n Jacobian points resident on the device (multi-GPU combine step)"""
    _lib.check(_L().b200_g1_sum_device(C.c_void_p(out_ptr), C.c_void_p(points_ptr), n, C.c_void_p(stream)))


def microbench_mul(field=0, mode=0):
    """field multiplications per second of a multiplier variant (0 = Fp / 1 = Fr; mode 0 = CIOS, 5 = Karatsuba)"""
    a = C.c_double()
    _lib.check(_L().b200_microbench_mul(field, mode, C.byref(a)))
    return a.value


def microbench_int():
    a, b = C.c_double(), C.c_double()
    _lib.check(_L().b200_microbench_int(C.byref(a), C.byref(b)))
    return {"imad_per_s": a.value, "fpmul_per_s": b.value}
