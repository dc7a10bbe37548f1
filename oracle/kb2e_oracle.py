"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle and for the compiled reference.

* ``Oracle``    -> oracle/_build/libkb2e_oracle.so (our plain-C restatement, kb2e_oracle.c)
* ``Reference`` -> oracle/_ref/libkb2e_ref.so (the UNMODIFIED reference compiled from
  /root/reference by oracle/Makefile, behind oracle/ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package kb2e_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libkb2e_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libkb2e_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "bin")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    """Compile the oracle (always) and the reference harness (when /root/reference exists)."""
    if force or not os.path.exists(ORACLE_SO) or os.path.exists("/root/reference/Makefile"):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)


def _d(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def _i(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_ip)


def f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        L = self.lib = C.CDLL(ORACLE_SO)
        L.orc_energy.restype = C.c_double
        L.orc_vec_len.restype = C.c_double
        L.orc_train_batch_ref.restype = C.c_double
        L.orc_train_batch_dfr.restype = C.c_double
        L.orc_sampler_create.restype = C.c_void_p
        L.orc_sampler_pr.restype = _dp

    # ---- reference semantics ---------------------------------------------------------------
    def energy(self, model, distance, ent, rel, w, h, t, r):
        ent, rel, w = f64(ent), f64(rel), f64(w)
        h, t, r = i32(h), i32(t), i32(r)
        out = np.empty(len(h), dtype=np.float64)
        self.lib.orc_energy_many(model, distance, ent.shape[1], _d(ent), _d(rel), _d(w),
                                 C.c_long(len(h)), _i(h), _i(t), _i(r), _d(out))
        return out

    def norm(self, a, ignore_short=True):
        a = f64(a).copy()
        self.lib.orc_norm(_d(a), len(a), int(ignore_short))
        return a

    def norm2(self, a, b, rate):
        a, b = f64(a).copy(), f64(b).copy()
        self.lib.orc_norm2(_d(a), _d(b), len(a), C.c_double(rate))
        return a, b

    def grad(self, model, distance, lr, ent, rel, w, h, t, r, corrupted):
        ent, rel, w = f64(ent), f64(rel), f64(w)
        en, rn = ent.copy(), rel.copy()
        wn = None if w is None else w.copy()
        self.lib.orc_grad(model, distance, ent.shape[1], C.c_double(lr), _d(ent), _d(rel), _d(w),
                          _d(en), _d(rn), _d(wn), int(h), int(t), int(r), int(corrupted))
        return en, rn, wn

    def grad_raw(self, model, distance, lr, ent, rel, w, h, t, r, corrupted, tail=False):
        """next tables after the accumulation part of one gradientUpdate (no normalisation); with
        tail=True the normalisation tail is applied afterwards (== grad, bitwise)."""
        ent, rel, w = f64(ent), f64(rel), f64(w)
        en, rn = ent.copy(), rel.copy()
        wn = None if w is None else w.copy()
        self.lib.orc_grad_raw(model, distance, ent.shape[1], C.c_double(lr), _d(ent), _d(rel), _d(w),
                              _d(en), _d(rn), _d(wn), int(h), int(t), int(r), int(corrupted))
        if tail:
            self.lib.orc_grad_tail(model, ent.shape[1], C.c_double(lr), _d(en), _d(rn), _d(wn), int(h), int(t), int(r))
        return en, rn, wn

    def train_batch_ref(self, model, distance, lr, margin, ent, rel, w, pairs):
        ent, rel, w = f64(ent), f64(rel), f64(w)
        pairs = i32(pairs).reshape(-1, 6)
        en, rn = np.empty_like(ent), np.empty_like(rel)
        wn = None if w is None else np.empty_like(w)
        losses = np.empty(len(pairs), dtype=np.float64)
        total = self.lib.orc_train_batch_ref(model, distance, ent.shape[1], ent.shape[0], rel.shape[0],
                                             C.c_double(lr), C.c_double(margin), _d(ent), _d(rel), _d(w),
                                             C.c_long(len(pairs)), _i(pairs), _d(en), _d(rn), _d(wn), _d(losses))
        return en, rn, wn, losses, total

    def bern(self, h, t, r, nR):
        h, t, r = i32(h), i32(t), i32(r)
        hm, tm = np.empty(nR), np.empty(nR)
        self.lib.orc_bern(C.c_long(len(h)), _i(h), _i(t), _i(r), nR, _d(hm), _d(tm))
        return hm, tm

    def rank(self, model, distance, ent, rel, w, test, filt):
        """test, filt: (n,3) int arrays of (h, t, r).  Returns raw_lo, raw_hi, filt_lo, filt_hi (2*nTest)."""
        ent, rel, w = f64(ent), f64(rel), f64(w)
        test = i32(test).reshape(-1, 3)
        filt = i32(filt).reshape(-1, 3)
        th, tt, tr = (i32(test[:, k]) for k in range(3))
        fh, ft, fr = (i32(filt[:, k]) for k in range(3))
        outs = [np.empty(2 * len(test), dtype=np.int32) for _ in range(4)]
        self.lib.orc_rank(model, distance, ent.shape[1], ent.shape[0], rel.shape[0], _d(ent), _d(rel), _d(w),
                          C.c_long(len(test)), _i(th), _i(tt), _i(tr),
                          C.c_long(len(filt)), _i(fh), _i(ft), _i(fr), *[_i(o) for o in outs])
        return outs

    # ---- deferred / counter-RNG semantics ---------------------------------------------------
    def philox(self, c, k):
        out = (C.c_uint32 * 4)()
        self.lib.orc_philox(*[C.c_uint32(int(x)) for x in c], C.c_uint32(int(k[0])), C.c_uint32(int(k[1])), out)
        return np.array(list(out), dtype=np.uint32)

    def sampler(self, train, nE, nR, method):
        return Sampler(self, train, nE, nR, method)

    def randmax_draws(self, seed, x, n):
        out = np.empty(n, dtype=np.int32)
        self.lib.orc_randmax_draws(C.c_uint64(seed), int(x), C.c_long(n), _i(out))
        return out

    def train_batch_dfr(self, model, distance, lr, margin, ent, rel, w, carry, pairs):
        """In place on ent/rel/w/carry (float64, contiguous).  Returns (loss, n_active)."""
        pairs = i32(pairs).reshape(-1, 6)
        active = C.c_long(0)
        loss = self.lib.orc_train_batch_dfr(model, distance, ent.shape[1], ent.shape[0], rel.shape[0],
                                            C.c_double(lr), C.c_double(margin), _d(ent), _d(rel), _d(w), _d(carry),
                                            C.c_long(len(pairs)), _i(pairs), C.byref(active))
        return loss, active.value


class Sampler:
    def __init__(self, oracle, train, nE, nR, method):
        self.o = oracle
        train = i32(train).reshape(-1, 3)
        self.n = len(train)
        self.nE, self.nR = nE, nR
        h, t, r = (i32(train[:, k]) for k in range(3))
        self.ptr = C.c_void_p(oracle.lib.orc_sampler_create(C.c_long(self.n), _i(h), _i(t), _i(r), nE, nR, method))

    def __del__(self):
        try:
            self.o.lib.orc_sampler_destroy(self.ptr)
        except Exception:
            pass

    def set_mode(self, mode):
        """0: uniform indices (default); 1: the index distribution of the reference's randMax."""
        self.o.lib.orc_sampler_set_mode(self.ptr, int(mode))
        return self

    def pr(self):
        p = self.o.lib.orc_sampler_pr(self.ptr)
        return np.ctypeslib.as_array(p, shape=(self.nR,)).copy()

    def sample_batch(self, seed, global_batch, count):
        out = np.empty((count, 6), dtype=np.int32)
        self.o.lib.orc_sample_batch(self.ptr, C.c_uint64(seed), C.c_uint32(global_batch), C.c_long(count), _i(out))
        return out

    def train_epochs_ref(self, model, distance, lr, margin, batches, first_epoch, epochs, seed, ent, rel, w):
        """Reference sequential batch semantics + the uniform counter sampler; in place on ent/rel/w."""
        loss = np.empty(epochs, dtype=np.float64)
        self.o.lib.orc_train_epochs_ref(self.ptr, model, distance, ent.shape[1], ent.shape[0], rel.shape[0],
                                        C.c_double(lr), C.c_double(margin), batches, first_epoch, epochs,
                                        C.c_uint64(seed), _d(ent), _d(rel), _d(w), _d(loss))
        return loss

    def train_epochs_dfr(self, model, distance, lr, margin, batches, first_epoch, epochs, seed, ent, rel, w):
        loss = np.empty(epochs, dtype=np.float64)
        self.o.lib.orc_train_epochs_dfr(self.ptr, model, distance, ent.shape[1], ent.shape[0], rel.shape[0],
                                        C.c_double(lr), C.c_double(margin), batches, first_epoch, epochs,
                                        C.c_uint64(seed), _d(ent), _d(rel), _d(w), _d(loss))
        return loss


class Reference:
    """The compiled, unmodified reference.  ``Reference.available()`` is False where oracle/_ref
    was not built (no /root/reference at build time and nothing shipped)."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        L = self.lib = C.CDLL(REF_SO)
        L.ref_vec_len.restype = C.c_double
        L.ref_train_files.restype = C.c_double
        L.ref_trainer_create.restype = C.c_void_p
        L.ref_trainer_epochs.restype = C.c_double
        L.ref_trainer_epochs.argtypes = [C.c_void_p, C.c_int]
        L.ref_trainer_write.argtypes = [C.c_void_p]
        L.ref_trainer_destroy.argtypes = [C.c_void_p]

    def energy(self, model, distance, ent, rel, w, h, t, r, zero_work=True):
        ent, rel, w = f64(ent), f64(rel), f64(w)
        h, t, r = i32(h), i32(t), i32(r)
        out = np.empty(len(h), dtype=np.float64)
        rc = self.lib.ref_energy(model, distance, ent.shape[1], ent.shape[0], rel.shape[0], _d(ent), _d(rel), _d(w),
                                 C.c_long(len(h)), _i(h), _i(t), _i(r), int(zero_work), _d(out))
        assert rc == 0
        return out

    def grad(self, model, distance, lr, ent, rel, w, h, t, r, corrupted):
        ent, rel, w = f64(ent), f64(rel), f64(w)
        en, rn = np.empty_like(ent), np.empty_like(rel)
        wn = None if w is None else np.empty_like(w)
        rc = self.lib.ref_grad(model, distance, ent.shape[1], ent.shape[0], rel.shape[0], C.c_double(lr),
                               _d(ent), _d(rel), _d(w), int(h), int(t), int(r), int(corrupted), _d(en), _d(rn), _d(wn))
        assert rc == 0
        return en, rn, wn

    def train_batch(self, model, distance, lr, margin, ent, rel, w, pairs, zero_work=True):
        ent, rel, w = f64(ent), f64(rel), f64(w)
        pairs = i32(pairs).reshape(-1, 6)
        en, rn = np.empty_like(ent), np.empty_like(rel)
        wn = None if w is None else np.empty_like(w)
        losses = np.empty(len(pairs), dtype=np.float64)
        rc = self.lib.ref_train_batch(model, distance, ent.shape[1], ent.shape[0], rel.shape[0], C.c_double(lr),
                                      C.c_double(margin), _d(ent), _d(rel), _d(w), C.c_long(len(pairs)), _i(pairs),
                                      int(zero_work), _d(en), _d(rn), _d(wn), _d(losses))
        assert rc == 0
        return en, rn, wn, losses

    def randmax(self, seed, x, n):
        """n draws of the reference's own randMax(x) after srand(seed)."""
        out = np.empty(n, dtype=np.int32)
        self.lib.ref_randmax(C.c_uint(seed), int(x), C.c_long(n), _i(out))
        return out

    def norm(self, a, ignore_short=True):
        a = f64(a).copy()
        self.lib.ref_norm(_d(a), len(a), int(ignore_short))
        return a

    def norm2(self, a, b, rate):
        a, b = f64(a).copy(), f64(b).copy()
        self.lib.ref_norm2(_d(a), _d(b), len(a), C.c_double(rate))
        return a, b

    def rank(self, model, distance, ent, rel, w, test, filt, zero_work=True):
        ent, rel, w = f64(ent), f64(rel), f64(w)
        test = i32(test).reshape(-1, 3)
        filt = i32(filt).reshape(-1, 3)
        th, tt, tr = (i32(test[:, k]) for k in range(3))
        fh, ft, fr = (i32(filt[:, k]) for k in range(3))
        raw = np.empty(2 * len(test), dtype=np.int32)
        flt = np.empty(2 * len(test), dtype=np.int32)
        rc = self.lib.ref_rank(model, distance, ent.shape[1], ent.shape[0], rel.shape[0], _d(ent), _d(rel), _d(w),
                               C.c_long(len(test)), _i(th), _i(tt), _i(tr), C.c_long(len(filt)), _i(fh), _i(ft), _i(fr),
                               int(zero_work), _i(raw), _i(flt))
        assert rc == 0
        return raw, flt

    def bern(self, datadir, nR):
        hm, tm = np.empty(nR), np.empty(nR)
        self.lib.ref_bern(datadir.encode(), nR, _d(hm), _d(tm))
        return hm, tm

    def train_files(self, model, datadir, outdir, D, lr, margin, method, distance, batches, epochs, seed,
                    seeddir=".", seedmethod=0, zero_work=True, write=True):
        """Returns seconds spent in the reference's bfgs() loop."""
        return self.lib.ref_train_files(model, datadir.encode(), outdir.encode(), D, C.c_double(lr), C.c_double(margin),
                                        method, distance, batches, epochs, C.c_uint(seed), seeddir.encode(),
                                        seedmethod, int(zero_work), int(write))


class ReferenceTrainer:
    """A live reference trainer (loadFiles + prepTrain done once); ``epochs(n)`` times n more epochs of
    the reference's own bfgs() loop.  Used by bench.py --impl reference and the statistical-parity runs."""

    def __init__(self, ref, model, datadir, outdir, D, lr, margin, method, distance, batches, seed,
                 seeddir=".", seedmethod=0, zero_work=True):
        self.ref = ref
        self.h = C.c_void_p(ref.lib.ref_trainer_create(model, datadir.encode(), outdir.encode(), D, C.c_double(lr),
                                                        C.c_double(margin), method, distance, batches, C.c_uint(seed),
                                                        seeddir.encode(), seedmethod, int(zero_work)))

    def epochs(self, n):
        return self.ref.lib.ref_trainer_epochs(self.h, int(n))

    def write(self):
        self.ref.lib.ref_trainer_write(self.h)

    def close(self):
        if self.h:
            self.ref.lib.ref_trainer_destroy(self.h)
            self.h = None
