"""ctypes mirror of include/kb2e_b200.h.  Every method maps 1:1 onto a C-ABI entry point."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libkb2e_b200.so")

MODELS = {"transe": 0, "transh": 1, "transr": 2}
TABLE_ENTITY, TABLE_RELATION, TABLE_WEIGHTS = 0, 1, 2
FLAG_RANK_EXACT_ONLY = 1
FLAG_TRANSR_NO_QUIRK = 2
FLAG_SAMPLER_RANDMAX = 4
FLAG_DETERMINISTIC = 8

SYMBOLS = [
    "kb2e_create", "kb2e_destroy", "kb2e_last_error", "kb2e_stream", "kb2e_set_train_triples", "kb2e_set_bern",
    "kb2e_init_embeddings", "kb2e_upload", "kb2e_download", "kb2e_train_epochs", "kb2e_get_train_stats",
    "kb2e_score", "kb2e_set_test_triples", "kb2e_add_filter_triples", "kb2e_rank", "kb2e_get_rank_stats",
    "kb2e_sample_batch", "kb2e_train_batch_pairs", "kb2e_train_batch_deltas", "kb2e_debug_transr_projection", "kb2e_set_replicas", "kb2e_select_replica",
    "kb2e_dist_setup", "kb2e_dist_connect", "kb2e_dist_init_embeddings", "kb2e_dist_upload", "kb2e_dist_download",
    "kb2e_dist_train_epochs", "kb2e_dist_teardown",
]


class Kb2eError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("model", C.c_int32), ("dim", C.c_int32), ("method", C.c_int32), ("distance", C.c_int32),
                ("batches", C.c_int32), ("device", C.c_int32), ("num_entities", C.c_int64), ("num_relations", C.c_int64),
                ("rate", C.c_double), ("margin", C.c_double), ("seed", C.c_uint64), ("flags", C.c_uint32),
                ("reserved", C.c_uint32)]


class TrainStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("active", C.c_uint64), ("touched_ent", C.c_uint64),
                ("touched_rel", C.c_uint64), ("launches", C.c_uint64), ("kernel_ms", C.c_double)]


class RankStats(C.Structure):
    _fields_ = [("queries", C.c_uint64), ("rechecked", C.c_uint64), ("launches", C.c_uint64),
                ("kernel_ms", C.c_double), ("main_kernel_ms", C.c_double), ("project_ms", C.c_double),
                ("project_kernel_ms", C.c_double)]


_lib = None


def load_library():
    """Load libkb2e_b200.so; fails loudly when it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Kb2eError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(or `make -C kb2e_b200/csrc`); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.kb2e_last_error.restype = C.c_char_p
    lib.kb2e_last_error.argtypes = [C.c_void_p]
    lib.kb2e_stream.restype = C.c_void_p
    lib.kb2e_stream.argtypes = [C.c_void_p]
    lib.kb2e_destroy.restype = None
    lib.kb2e_destroy.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _columns(triples):
    """(n, 3) array of (head, tail, relation) -> three contiguous int32 columns; a tuple/list of three
    int32 arrays is passed through untouched (so callers can hand over pinned host buffers)."""
    if isinstance(triples, (tuple, list)) and len(triples) == 3 and all(isinstance(x, np.ndarray) for x in triples):
        h, t, r = (_i32(x) for x in triples)
        return h, t, r, len(h)
    tr = _i32(triples).reshape(-1, 3)
    return _i32(tr[:, 0]), _i32(tr[:, 1]), _i32(tr[:, 2]), len(tr)


class Context:
    """One kb2e_ctx: one model instance on one GPU."""

    def __init__(self, model, dim, num_entities, num_relations, *, method=1, distance=0, batches=100,
                 rate=0.001, margin=1.0, seed=0, device=0, flags=0):
        self.lib = load_library()
        self.model = MODELS[model] if isinstance(model, str) else int(model)
        self.dim, self.nE, self.nR = int(dim), int(num_entities), int(num_relations)
        cfg = Config(self.model, self.dim, int(method), int(distance), int(batches), int(device), self.nE, self.nR,
                     float(rate), float(margin), int(seed), int(flags), 0)
        self.ptr = C.c_void_p()
        rc = self.lib.kb2e_create(C.byref(cfg), C.byref(self.ptr))
        if rc != 0:
            msg = self.lib.kb2e_last_error(None).decode()
            self.ptr = None
            raise Kb2eError(f"kb2e_create failed ({rc}): {msg}")

    def close(self):
        if getattr(self, "ptr", None):
            self.lib.kb2e_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise Kb2eError(f"{what} failed ({rc}): {self.lib.kb2e_last_error(self.ptr).decode()}")

    @property
    def stream(self):
        return self.lib.kb2e_stream(self.ptr)

    def table_shape(self, table):
        if table == TABLE_ENTITY:
            return (self.nE, self.dim)
        if table == TABLE_RELATION:
            return (self.nR, self.dim)
        return (self.nR, self.dim) if self.model == 1 else (self.nR * self.dim, self.dim)

    # ---- training ----
    def set_train_triples(self, triples):
        """triples: (n, 3) ints (head, tail, relation), or a tuple of three int32 arrays."""
        h, t, r, n = _columns(triples)
        self._check(self.lib.kb2e_set_train_triples(self.ptr, _p(h, C.c_int32), _p(t, C.c_int32), _p(r, C.c_int32),
                                                    C.c_int64(n)), "kb2e_set_train_triples")

    def set_bern(self, head_mean=None, tail_mean=None):
        hm = None if head_mean is None else np.ascontiguousarray(head_mean, dtype=np.float64)
        tm = None if tail_mean is None else np.ascontiguousarray(tail_mean, dtype=np.float64)
        self._check(self.lib.kb2e_set_bern(self.ptr, _p(hm, C.c_double), _p(tm, C.c_double)), "kb2e_set_bern")

    def init_embeddings(self):
        self._check(self.lib.kb2e_init_embeddings(self.ptr), "kb2e_init_embeddings")

    def upload(self, table, array):
        a = np.ascontiguousarray(array, dtype=np.float64)
        rows, cols = self.table_shape(table)
        a = a.reshape(rows, cols)
        self._check(self.lib.kb2e_upload(self.ptr, int(table), _p(a, C.c_double), C.c_int64(rows), C.c_int64(cols)),
                    "kb2e_upload")

    def download(self, table, out=None):
        rows, cols = self.table_shape(table)
        if out is None:
            out = np.empty((rows, cols), dtype=np.float64)
        assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"] and out.size == rows * cols
        self._check(self.lib.kb2e_download(self.ptr, int(table), _p(out, C.c_double), C.c_int64(rows), C.c_int64(cols)),
                    "kb2e_download")
        return out

    def set_replicas(self, n_models, rates=None, margins=None, seeds=None):
        """Batched training: n_models stacked models (call right after construction)."""
        r = None if rates is None else np.ascontiguousarray(rates, dtype=np.float64)
        m = None if margins is None else np.ascontiguousarray(margins, dtype=np.float64)
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
        self._check(self.lib.kb2e_set_replicas(self.ptr, int(n_models), _p(r, C.c_double), _p(m, C.c_double), _p(s, C.c_uint64)),
                    "kb2e_set_replicas")
        self.n_models = int(n_models)

    def select_replica(self, index):
        self._check(self.lib.kb2e_select_replica(self.ptr, int(index)), "kb2e_select_replica")

    def train_epochs(self, first_epoch, n_epochs):
        """Per-epoch losses; after set_replicas(K > 1) an array of shape (K, n_epochs)."""
        k = getattr(self, "n_models", 1)
        loss = np.zeros(max(k * n_epochs, 1), dtype=np.float64)
        self._check(self.lib.kb2e_train_epochs(self.ptr, int(first_epoch), int(n_epochs), _p(loss, C.c_double)),
                    "kb2e_train_epochs")
        return loss[:n_epochs] if k == 1 else loss[:k * n_epochs].reshape(k, n_epochs)

    def train_stats(self):
        s = TrainStats()
        self._check(self.lib.kb2e_get_train_stats(self.ptr, C.byref(s)), "kb2e_get_train_stats")
        return {k: getattr(s, k) for k, _ in TrainStats._fields_}

    # ---- scoring / ranking ----
    def score(self, triples, precision=1):
        h, t, r, n = _columns(triples)
        out = np.empty(n, dtype=np.float64)
        self._check(self.lib.kb2e_score(self.ptr, _p(h, C.c_int32), _p(t, C.c_int32), _p(r, C.c_int32), C.c_int64(n),
                                        int(precision), _p(out, C.c_double)), "kb2e_score")
        return out

    def set_test_triples(self, triples):
        h, t, r, n = _columns(triples)
        self.n_test = n
        self._check(self.lib.kb2e_set_test_triples(self.ptr, _p(h, C.c_int32), _p(t, C.c_int32), _p(r, C.c_int32),
                                                   C.c_int64(n)), "kb2e_set_test_triples")

    def add_filter_triples(self, triples):
        h, t, r, n = _columns(triples)
        if n == 0:
            return
        self._check(self.lib.kb2e_add_filter_triples(self.ptr, _p(h, C.c_int32), _p(t, C.c_int32), _p(r, C.c_int32),
                                                     C.c_int64(n)), "kb2e_add_filter_triples")

    def clear_filter_triples(self):
        self._check(self.lib.kb2e_add_filter_triples(self.ptr, None, None, None, C.c_int64(0)), "kb2e_add_filter_triples")

    def rank(self, first=0, count=None, want_ranks=True, out=None):
        """Returns dict(raw, filt, raw_ties, filt_ties, sums) for test triples [first, first+count).
        out: optional list of four int32 arrays (2*count each) to receive the per-query results."""
        if count is None:
            count = self.n_test - first
        outs = out if out is not None else [np.empty(2 * count, dtype=np.int32) if want_ranks else None for _ in range(4)]
        sums = np.zeros(4, dtype=np.int64)
        self._check(self.lib.kb2e_rank(self.ptr, C.c_int64(first), C.c_int64(count), *[_p(o, C.c_int32) for o in outs],
                                       _p(sums, C.c_int64)), "kb2e_rank")
        return {"raw": outs[0], "filt": outs[1], "raw_ties": outs[2], "filt_ties": outs[3], "sums": sums}

    def rank_stats(self):
        s = RankStats()
        self._check(self.lib.kb2e_get_rank_stats(self.ptr, C.byref(s)), "kb2e_get_rank_stats")
        return {k: getattr(s, k) for k, _ in RankStats._fields_}

    # ---- test hooks ----
    def sample_batch(self, epoch, batch, count):
        out = np.empty((count, 6), dtype=np.int32)
        self._check(self.lib.kb2e_sample_batch(self.ptr, int(epoch), int(batch), C.c_int64(count), _p(out, C.c_int32)),
                    "kb2e_sample_batch")
        return out

    def train_batch_pairs(self, pairs):
        p = _i32(pairs).reshape(-1, 6)
        loss = C.c_double(0)
        active = C.c_int64(0)
        self._check(self.lib.kb2e_train_batch_pairs(self.ptr, _p(p, C.c_int32), C.c_int64(len(p)), C.byref(loss),
                                                    C.byref(active)), "kb2e_train_batch_pairs")
        return loss.value, active.value

    def debug_transr_projection(self, relation):
        """(tensor-core projection of every entity under `relation` [nE][dim] float32, relative error bound)."""
        out = np.empty((self.nE, self.dim), dtype=np.float32)
        eps = C.c_double(0)
        self._check(self.lib.kb2e_debug_transr_projection(self.ptr, int(relation), _p(out, C.c_float), C.byref(eps)),
                    "kb2e_debug_transr_projection")
        return out, eps.value

    def train_batch_deltas(self, pairs):
        """The summed pre-normalisation updates of the pairs (tables untouched): (d_ent, d_rel, d_w, loss, n_active)."""
        p = _i32(pairs).reshape(-1, 6)
        de = np.zeros((self.nE, self.dim), dtype=np.float64)
        dr = np.zeros((self.nR, self.dim), dtype=np.float64)
        dw = np.zeros(self.table_shape(TABLE_WEIGHTS), dtype=np.float64) if self.model != 0 else None
        loss = C.c_double(0)
        active = C.c_int64(0)
        self._check(self.lib.kb2e_train_batch_deltas(self.ptr, _p(p, C.c_int32), C.c_int64(len(p)), _p(de, C.c_double),
                                                     _p(dr, C.c_double), _p(dw, C.c_double), C.byref(loss), C.byref(active)),
                    "kb2e_train_batch_deltas")
        return de, dr, dw, loss.value, active.value

    # ---- entity-partitioned multi-GPU training (one Context per process per GPU) ----
    def dist_setup(self, rank, world):
        """Allocates this rank's arena; returns its 64-byte CUDA IPC handle."""
        buf = C.create_string_buffer(64)
        self._check(self.lib.kb2e_dist_setup(self.ptr, int(rank), int(world), buf), "kb2e_dist_setup")
        self.dist_rank, self.dist_world = int(rank), int(world)
        self.rows_local = (self.nE - rank + world - 1) // world
        return buf.raw

    def dist_connect(self, handles):
        """handles: the concatenated 64-byte handles of all ranks, in rank order."""
        blob = b"".join(handles) if not isinstance(handles, (bytes, bytearray)) else bytes(handles)
        assert len(blob) == 64 * self.dist_world
        self._check(self.lib.kb2e_dist_connect(self.ptr, C.c_char_p(blob)), "kb2e_dist_connect")

    def dist_init_embeddings(self):
        self._check(self.lib.kb2e_dist_init_embeddings(self.ptr), "kb2e_dist_init_embeddings")

    def dist_upload(self, table, array):
        a = np.ascontiguousarray(array, dtype=np.float64)
        rows = self.rows_local if table == TABLE_ENTITY else self.nR
        a = a.reshape(rows, self.dim)
        self._check(self.lib.kb2e_dist_upload(self.ptr, int(table), _p(a, C.c_double), C.c_int64(rows), C.c_int64(self.dim)),
                    "kb2e_dist_upload")

    def dist_download(self, table):
        rows = self.rows_local if table == TABLE_ENTITY else self.nR
        out = np.empty((rows, self.dim), dtype=np.float64)
        self._check(self.lib.kb2e_dist_download(self.ptr, int(table), _p(out, C.c_double), C.c_int64(rows), C.c_int64(self.dim)),
                    "kb2e_dist_download")
        return out

    def dist_train_epochs(self, first_epoch, n_epochs):
        """Collective: every rank calls it with the same arguments.  Returns this rank's share of the loss."""
        loss = np.zeros(max(n_epochs, 1), dtype=np.float64)
        self._check(self.lib.kb2e_dist_train_epochs(self.ptr, int(first_epoch), int(n_epochs), _p(loss, C.c_double)),
                    "kb2e_dist_train_epochs")
        return loss[:n_epochs]
