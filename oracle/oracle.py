"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes binding of oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (evosops_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

AGG, SEP, COAT, LOCO = 0, 1, 2, 3
GENOME_LEN = {AGG: 48, SEP: 600, COAT: 600, LOCO: 144}


class Result(C.Structure):
    _fields_ = [("i0", C.c_uint32), ("i1", C.c_uint32), ("i2", C.c_uint32), ("f", C.c_float), ("norm", C.c_double)]


RESULT_DTYPE = np.dtype([("i0", "<u4"), ("i1", "<u4"), ("i2", "<u4"), ("f", "<f4"), ("norm", "<f8")], align=True)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("sops_oracle.cpp", "rng.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_env_create.restype = C.c_void_p
        L.oracle_env_create.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_float,
                                        C.c_float, C.c_int, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.oracle_env_destroy.argtypes = [C.c_void_p]
        L.oracle_env_grid_size.argtypes = [C.c_void_p]
        L.oracle_env_particle_count.argtypes = [C.c_void_p]
        L.oracle_env_particle_count.restype = C.c_uint32
        L.oracle_env_sim_duration.argtypes = [C.c_void_p]
        L.oracle_env_sim_duration.restype = C.c_uint64
        L.oracle_env_step.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_env_fitness.argtypes = [C.c_void_p, C.POINTER(Result)]
        L.oracle_env_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_env_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_env_load_grid.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_env_dist_grid.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_env_light.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_stdrng_key.argtypes = [C.c_uint64, C.c_void_p]
        L.oracle_stdrng_u64.argtypes = [C.c_uint64, C.c_int, C.c_void_p]
        L.oracle_chacha_block.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        L.oracle_philox2x32_10.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.oracle_wyrand_u64.argtypes = [C.c_uint64, C.c_int, C.c_void_p]
        L.oracle_move_draws.argtypes = [C.c_uint64, C.c_int, C.c_uint32, C.c_void_p]
        L.oracle_eval_batch.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                        C.c_uint32, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_lemire_redraws.restype = C.c_ulonglong
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Env:
    """One oracle environment: init_sops_env → step / simulate → evaluate_fitness → dump."""

    def __init__(self, behavior, genome, arena, layers, coat=0, seed=1, w1=0.65, w2=0.35, rng_mode=0,
                 move_seed=None, orientation=1, prob_table=0):
        L = lib()
        g = np.ascontiguousarray(np.asarray(genome, dtype=np.uint8).reshape(-1))
        assert g.size == GENOME_LEN[behavior], (g.size, behavior)
        err = C.c_int(0)
        ms = seed if move_seed is None else move_seed
        self.h = L.oracle_env_create(behavior, _ptr(g), arena, layers, coat, seed, w1, w2, rng_mode, ms,
                                     orientation, prob_table, C.byref(err))
        if not self.h:
            raise ValueError("oracle init failed: rc=%d" % err.value)
        self.behavior = behavior
        self.G = L.oracle_env_grid_size(self.h)
        self.n = L.oracle_env_particle_count(self.h)
        self.sim_duration = L.oracle_env_sim_duration(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_env_destroy(self.h)
            self.h = None

    def step(self, count=1):
        lib().oracle_env_step(self.h, int(count))

    def simulate(self, steps=None):
        self.step(self.sim_duration if steps is None else steps)
        return self.fitness()

    def fitness(self):
        r = Result()
        lib().oracle_env_fitness(self.h, C.byref(r))
        return r

    def counters(self):
        out = np.zeros(3, dtype=np.uint64)
        lib().oracle_env_counters(self.h, _ptr(out))
        return out

    def grid(self):
        g = np.zeros((self.G, self.G), dtype=np.uint8)
        lib().oracle_env_dump(self.h, _ptr(g), None)
        return g

    def particles(self):
        p = np.zeros((self.n, 3), dtype=np.uint8)
        lib().oracle_env_dump(self.h, None, _ptr(p))
        return p

    def load_grid(self, grid):
        g = np.ascontiguousarray(np.asarray(grid, dtype=np.uint8))
        assert g.shape == (self.G, self.G)
        self.n = lib().oracle_env_load_grid(self.h, _ptr(g))
        return self.n

    def dist_grid(self):
        d = np.zeros((self.G, self.G), dtype=np.uint16)
        lib().oracle_env_dist_grid(self.h, _ptr(d))
        return d

    def light(self):
        out = np.zeros((self.G, 2), dtype=np.uint8)
        lib().oracle_env_light(self.h, _ptr(out))
        return out


def stdrng_key(seed):
    k = np.zeros(8, dtype=np.uint32)
    lib().oracle_stdrng_key(seed, _ptr(k))
    return k


def stdrng_u64(seed, count):
    out = np.zeros(count, dtype=np.uint64)
    lib().oracle_stdrng_u64(seed, count, _ptr(out))
    return out


def chacha_block(key, counter, stream, double_rounds):
    k = np.ascontiguousarray(np.asarray(key, dtype=np.uint32))
    out = np.zeros(16, dtype=np.uint32)
    lib().oracle_chacha_block(_ptr(k), counter, stream, double_rounds, _ptr(out))
    return out


def philox2x32_10(c0, c1, key):
    out = np.zeros(2, dtype=np.uint32)
    lib().oracle_philox2x32_10(c0, c1, key, _ptr(out))
    return out


def wyrand_u64(seed, count):
    out = np.zeros(count, dtype=np.uint64)
    lib().oracle_wyrand_u64(seed, count, _ptr(out))
    return out


def move_draws(move_seed, count, n):
    out = np.zeros((count, 3), dtype=np.uint32)
    lib().oracle_move_draws(move_seed, count, n, _ptr(out))
    return out


def eval_batch(behavior, genomes, sizes, seeds, w1=0.65, w2=0.35, rng_mode=0, move_seeds=None, orient=None,
               steps_override=0, prob_table=0, n_threads=0):
    """genomes [P, L] u8; sizes list of (arena, layers[, coat]) and seeds of equal length T (the
    flattened trial list).  Returns (results[P, T] structured array, counters[3])."""
    g = np.ascontiguousarray(np.asarray(genomes, dtype=np.uint8))
    g = g.reshape(-1, GENOME_LEN[behavior])
    P = g.shape[0]
    T = len(seeds)
    sz = np.zeros((T, 3), dtype=np.uint16)
    for t, s in enumerate(sizes):
        sz[t, :len(s)] = s
    sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
    ms = None if move_seeds is None else np.ascontiguousarray(np.asarray(move_seeds, dtype=np.uint64).reshape(P * T))
    orr = None if orient is None else np.ascontiguousarray(np.asarray(orient, dtype=np.uint8).reshape(P * T))
    out = np.zeros((P, T), dtype=RESULT_DTYPE)
    counters = np.zeros(3, dtype=np.uint64)
    rc = lib().oracle_eval_batch(behavior, _ptr(g), P, g.shape[1], _ptr(sz), _ptr(sd), T, w1, w2, rng_mode,
                                 _ptr(ms), _ptr(orr), steps_override, prob_table, n_threads, _ptr(out),
                                 _ptr(counters))
    if rc != 0:
        raise ValueError("oracle_eval_batch rc=%d" % rc)
    return out, counters


def lemire_redraws():
    """How many times the oracle's WyRand Lemire loops redrew so far (process-wide test instrumentation)."""
    return int(lib().oracle_lemire_redraws())
