"""ctypes binding of include/evosops.h (libevosops_b200.so).  No CPU fallback: if the library is
missing it is rebuilt with nvcc, and without a CUDA device `Context()` raises."""
import ctypes as C
import os

import numpy as np

from . import _build

AGG, SEP, COAT, LOCO = 0, 1, 2, 3
RNG_VERIFY, RNG_FAST = 0, 1
F_DUMP, F_INIT_ONLY, F_THEORY_TABLE, F_FORCE_WARP, F_FORCE_CTA = 1, 2, 4, 8, 16
F_ROUNDS_SERIAL, F_ROUNDS_PARALLEL, F_ROUNDS_ADAPTIVE = 32, 64, 128
GENOME_LEN = {AGG: 48, SEP: 600, COAT: 600, LOCO: 144}
GENOME_SHAPE = {AGG: (4, 3, 4), SEP: (10, 6, 10), COAT: (10, 6, 10), LOCO: (3, 4, 3, 4)}
E_ARG, E_GENE, E_SIZE, E_UNSUPPORTED, E_CUDA, E_STATE, E_INIT_REJECT = -1, -2, -3, -4, -5, -6, -7

# every symbol include/evosops.h declares
SYMBOLS = ["sops_ctx_create", "sops_ctx_destroy", "sops_last_error", "sops_eval_batch", "sops_batch_prepare",
           "sops_batch_launch", "sops_batch_collect", "sops_batch_sync", "sops_batch_counters", "sops_dump_state",
           "sops_instance_shape", "sops_coat_distance_grid", "sops_ctx_set_trial_steps"]

SIZE_DTYPE = np.dtype([("arena", "<u2"), ("layers", "<u2"), ("coat", "<u2")])
RESULT_DTYPE = np.dtype([("i0", "<u4"), ("i1", "<u4"), ("i2", "<u4"), ("f", "<f4"), ("norm", "<f8")], align=True)
assert SIZE_DTYPE.itemsize == 6 and RESULT_DTYPE.itemsize == 24


class SopsSize(C.Structure):
    _fields_ = [("arena", C.c_uint16), ("layers", C.c_uint16), ("coat", C.c_uint16)]


class SopsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("evosops_b200 error %d: %s" % (code, msg))
        self.code = code


_LIB = None


def lib():
    """Load (building if needed) the shared library.  Fails loudly if it cannot be built/loaded."""
    global _LIB
    if _LIB is None:
        path = _build.build()
        L = C.CDLL(path)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.sops_ctx_create.argtypes = [C.POINTER(vp), vp, i32]
        L.sops_ctx_destroy.argtypes = [vp]
        L.sops_ctx_destroy.restype = None
        L.sops_last_error.argtypes = [vp]
        L.sops_last_error.restype = C.c_char_p
        batch = [vp, i32, vp, u32, vp, vp, u32, C.c_uint8, C.c_float, C.c_float, i32, vp, vp, u64, u32]
        L.sops_eval_batch.argtypes = batch + [vp]
        L.sops_batch_prepare.argtypes = batch
        L.sops_batch_launch.argtypes = [vp]
        L.sops_batch_collect.argtypes = [vp, vp]
        L.sops_batch_sync.argtypes = [vp, C.POINTER(C.c_double)]
        L.sops_batch_counters.argtypes = [vp, vp]
        L.sops_dump_state.argtypes = [vp, u32, vp, vp]
        L.sops_instance_shape.argtypes = [i32, SopsSize, C.POINTER(u32), C.POINTER(u32), C.POINTER(u64)]
        L.sops_coat_distance_grid.argtypes = [SopsSize, vp]
        L.sops_ctx_set_trial_steps.argtypes = [vp, vp, u32]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _mk_size(s):
    s = tuple(int(v) for v in s) + (0,) * (3 - len(s))
    return SopsSize(s[0], s[1], s[2])


def instance_shape(behavior, size):
    """(grid_side, n_particles, sim_duration) of init_sops_env(size) — host arithmetic only."""
    G, n, d = C.c_uint32(), C.c_uint32(), C.c_uint64()
    rc = lib().sops_instance_shape(behavior, _mk_size(size), C.byref(G), C.byref(n), C.byref(d))
    if rc:
        raise SopsError(rc, "invalid size %r" % (size,))
    return G.value, n.value, d.value


def coat_distance_grid(size):
    G, _, _ = instance_shape(COAT, size)
    out = np.zeros((G, G), dtype=np.uint16)
    rc = lib().sops_coat_distance_grid(_mk_size(size), _ptr(out))
    if rc:
        raise SopsError(rc, "invalid size %r" % (size,))
    return out


class Context:
    """sops_ctx: owns the streams and device buffers of one or more B200s."""

    def __init__(self, device_ids=None, n_dev=None):
        L = lib()
        if device_ids is None:
            device_ids = list(range(n_dev or 1))
        ids = np.asarray(device_ids, dtype=np.int32)
        h = C.c_void_p()
        rc = L.sops_ctx_create(C.byref(h), _ptr(ids), len(ids))
        if rc:
            raise SopsError(rc, "sops_ctx_create failed (no CUDA device?)")
        self.h = h
        self.device_ids = list(int(i) for i in ids)
        self._shape = None

    def close(self):
        if getattr(self, "h", None):
            lib().sops_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise SopsError(rc, lib().sops_last_error(self.h).decode())

    def _pack(self, behavior, genomes, sizes, seeds, move_seeds, orient):
        g = np.ascontiguousarray(np.asarray(genomes, dtype=np.uint8)).reshape(-1, GENOME_LEN[behavior])
        P, T = g.shape[0], len(seeds)
        assert len(sizes) == T, "sizes and seeds are the flattened trial list (equal length)"
        sz = np.zeros(T, dtype=SIZE_DTYPE)
        for t, s in enumerate(sizes):
            s = tuple(s) + (0,) * (3 - len(s))
            sz[t] = s
        sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
        ms = None if move_seeds is None else np.ascontiguousarray(np.asarray(move_seeds, dtype=np.uint64).reshape(P * T))
        orr = None if orient is None else np.ascontiguousarray(np.asarray(orient, dtype=np.uint8).reshape(P * T))
        return g, sz, sd, ms, orr, P, T

    def prepare(self, behavior, genomes, sizes, seeds, gran=10, w1=0.65, w2=0.35, rng_mode=RNG_VERIFY,
                move_seeds=None, orient=None, steps_override=0, flags=0):
        g, sz, sd, ms, orr, P, T = self._pack(behavior, genomes, sizes, seeds, move_seeds, orient)
        self._check(lib().sops_batch_prepare(self.h, behavior, _ptr(g), P, _ptr(sz), _ptr(sd), T, gran, w1, w2,
                                             rng_mode, _ptr(ms), _ptr(orr), steps_override, flags))
        self._shape = (P, T)

    def set_trial_steps(self, steps):
        """Per-trial chain lengths for the next prepare / eval_batch (None clears)."""
        a = None if steps is None else np.ascontiguousarray(np.asarray(steps, dtype=np.uint64))
        self._check(lib().sops_ctx_set_trial_steps(self.h, _ptr(a), 0 if a is None else a.size))

    def launch(self):
        self._check(lib().sops_batch_launch(self.h))

    def sync(self):
        ms = C.c_double()
        self._check(lib().sops_batch_sync(self.h, C.byref(ms)))
        return ms.value

    def collect(self):
        P, T = self._shape
        out = np.zeros((P, T), dtype=RESULT_DTYPE)
        self._check(lib().sops_batch_collect(self.h, _ptr(out)))
        return out

    def eval_batch(self, behavior, genomes, sizes, seeds, gran=10, w1=0.65, w2=0.35, rng_mode=RNG_VERIFY,
                   move_seeds=None, orient=None, steps_override=0, flags=0):
        """One call = one GA generation's nested par_iter (base_ga.rs:393-411).  Returns results[P, T]."""
        g, sz, sd, ms, orr, P, T = self._pack(behavior, genomes, sizes, seeds, move_seeds, orient)
        out = np.zeros((P, T), dtype=RESULT_DTYPE)
        self._shape = (P, T)
        self._check(lib().sops_eval_batch(self.h, behavior, _ptr(g), P, _ptr(sz), _ptr(sd), T, gran, w1, w2,
                                          rng_mode, _ptr(ms), _ptr(orr), steps_override, flags, _ptr(out)))
        return out

    def counters(self):
        """dict(attempts, possible, committed, launches, h2d_bytes, d2h_bytes) of the last collected batch."""
        c = np.zeros(6, dtype=np.uint64)
        self._check(lib().sops_batch_counters(self.h, _ptr(c)))
        keys = ("attempts", "possible", "committed", "launches", "h2d_bytes", "d2h_bytes")
        return {k: int(v) for k, v in zip(keys, c)}

    def dump_state(self, instance, G, n):
        grid = np.zeros((G, G), dtype=np.uint8)
        parts = np.zeros((n, 3), dtype=np.uint8)
        self._check(lib().sops_dump_state(self.h, instance, _ptr(grid), _ptr(parts)))
        return grid, parts
