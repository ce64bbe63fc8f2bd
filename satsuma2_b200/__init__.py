"""satsuma2_b200 -- Satsuma2's chunk-pair cross-correlation hot path on NVIDIA B200.

The product is ``libsatsuma_b200.so`` (hand-written sm_100a CUDA kernels behind the C ABI in
``include/satsuma_xcorr.h``).  This module is the thin ctypes binding used by the tests, the
benchmark and Python callers; it mirrors the reference's slave-side interface
(``HomologyByXCorr::align_target`` over ``t_pair`` blocks, analysis/HomologyByXCorrSlave.cc:270-300)
and never computes anything itself: without the compiled library or a CUDA device it raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsatsuma_b200.so")

SX_OK, SX_ERR_ARG, SX_ERR_CUDA, SX_ERR_NOMEM, SX_ERR_CAPACITY, SX_ERR_STATE = 0, -1, -2, -3, -4, -5
ABI_VERSION = 1

# t_result (analysis/WorkQueue.h:23-33), 72 bytes
RESULT_DTYPE = np.dtype(
    [("query_id", "<u8"), ("target_id", "<u8"), ("query_size", "<u8"), ("qstart", "<u8"), ("tstart", "<u8"),
     ("len", "<u8"), ("reverse", "u1"), ("pad", "u1", (7,)), ("prob", "<f8"), ("ident", "<f8")]
)
# t_pair (analysis/WorkQueue.h:17-22), 28 bytes
PAIR_DTYPE = np.dtype(
    [("target_from", "<i4"), ("target_to", "<i4"), ("query_from", "<i4"), ("query_to", "<i4"), ("fast", "u1"),
     ("pad0", "u1", (3,)), ("slave_id", "<i4"), ("status", "u1"), ("pad1", "u1", (3,))]
)
SEGMENT_DTYPE = np.dtype([("start_target", "<i4"), ("start_query", "<i4"), ("len", "<i4")])
assert RESULT_DTYPE.itemsize == 72 and PAIR_DTYPE.itemsize == 28


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("t_chunk", C.c_int32), ("q_chunk", C.c_int32),
        ("cutoff", C.c_double), ("cutoff_fast", C.c_double), ("min_len", C.c_int32), ("use_prob_table", C.c_int32),
        ("min_prob", C.c_double), ("prob_table_value", C.c_double), ("target_total", C.c_double),
        ("rc_coord_mode", C.c_int32), ("max_batch_pairs", C.c_int32), ("spectra_cache_bytes", C.c_int64),
        ("sort_results", C.c_int32), ("debug_small_pools", C.c_int32), ("async_upload", C.c_int32),
        ("debug_flags", C.c_int32), ("fuse_pairs", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("chunk_pairs", C.c_int64), ("strand_pairs", C.c_int64), ("signals", C.c_int64), ("candidates", C.c_int64),
        ("segments", C.c_int64), ("matches", C.c_int64), ("kernel_launches", C.c_int64), ("batches", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("retries", C.c_int64), ("ms_encode_fft", C.c_double),
        ("ms_xcorr", C.c_double), ("ms_scan_score", C.c_double), ("ms_total", C.c_double), ("positions", C.c_int64),
        ("fused_pairs", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SatsumaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


_lib = None


def load_library():
    """Loads libsatsuma_b200.so; fails loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m satsuma2_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.sx_default_config.argtypes = [C.POINTER(Config)]
    L.sx_default_config.restype = None
    L.sx_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.sx_destroy.argtypes = [vp]
    L.sx_destroy.restype = None
    L.sx_last_error.restype = C.c_char_p
    for name in ("sx_set_targets", "sx_set_queries"):
        getattr(L, name).argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, i32]
    L.sx_invalidate_spectra.argtypes = [vp]
    L.sx_align_blocks.argtypes = [vp, vp, i32, vp, i64, C.POINTER(i64)]
    L.sx_align_pairs.argtypes = [vp, vp, i64, i32, vp, i64, C.POINTER(i64)]
    L.sx_tap_signal.argtypes = [vp, i32, i32, i32, vp]
    L.sx_tap_xcorr.argtypes = [vp, i32, i32, i32, vp]
    L.sx_tap_candidates.argtypes = [vp, i32, i32, i32, i32, vp, i32, C.POINTER(i32)]
    L.sx_tap_segments.argtypes = [vp, i32, i32, i32, i32, vp, i32, C.POINTER(i32)]
    L.sx_tap_matchup.argtypes = [vp, i32, i32, dbl, vp, vp, i32, C.POINTER(i32)]
    L.sx_set_profiling.argtypes = [vp, i32]
    L.sx_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.sx_reset_stats.argtypes = [vp]
    L.sx_stream.argtypes = [vp, C.POINTER(vp)]
    L.sx_build_prob_table.argtypes = [dbl, vp]
    L.sx_set_prob_table.argtypes = [vp, vp]
    L.sx_multi_create.argtypes = [C.POINTER(Config), vp, i32, i32, i32, C.POINTER(vp)]
    L.sx_multi_destroy.argtypes = [vp]
    L.sx_multi_destroy.restype = None
    L.sx_multi_last_error.restype = C.c_char_p
    L.sx_multi_device_count.argtypes = [vp]
    L.sx_multi_target_range.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    for name in ("sx_multi_set_targets", "sx_multi_set_queries"):
        getattr(L, name).argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, i32]
    L.sx_multi_set_prob_table.argtypes = [vp, vp]
    L.sx_multi_invalidate_spectra.argtypes = [vp]
    L.sx_multi_align_blocks.argtypes = [vp, vp, i32, vp, i64, C.POINTER(i64)]
    L.sx_multi_get_stats.argtypes = [vp, i32, C.POINTER(Stats)]
    L.sx_multi_reset_stats.argtypes = [vp]
    L.sx_multi_stream.argtypes = [vp, i32, C.POINTER(vp)]
    if L.sx_abi_version() != ABI_VERSION:
        raise ImportError("libsatsuma_b200.so ABI version mismatch")
    _lib = L
    return L


def default_config(**overrides) -> Config:
    L = load_library()
    cfg = Config()
    L.sx_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def _check(rc: int):
    if rc != SX_OK:
        raise SatsumaError(rc, load_library().sx_last_error().decode(errors="replace"))


def build_prob_table(target_total: float) -> np.ndarray:
    """Host-side ProbTable::Setup (analysis/ProbTable.cc:15-56)."""
    tab = np.zeros((512, 2048), dtype=np.float64)
    _check(load_library().sx_build_prob_table(float(target_total), tab.ctypes.data))
    return tab


def make_blocks(blocks) -> np.ndarray:
    """(target_from, target_to, query_from, query_to[, fast]) tuples, inclusive ranges -> t_pair array (PAIR_DTYPE)."""
    if isinstance(blocks, np.ndarray) and blocks.dtype == PAIR_DTYPE:
        return np.ascontiguousarray(blocks)
    arr = np.zeros(len(blocks), dtype=PAIR_DTYPE)
    for i, b in enumerate(blocks):
        arr[i]["target_from"], arr[i]["target_to"], arr[i]["query_from"], arr[i]["query_to"] = b[:4]
        arr[i]["fast"] = 1 if (len(b) > 4 and b[4]) else 0
    return arr


class ChunkSet:
    """Flat chunk list as the slave holds it (vecDNAVector + vector<SeqChunk> + ChunkManager sizes).

    bases: one uint8 blob; chunk i = bases[offsets[i] : offsets[i] + lens[i]] (chunks may overlap in the
    blob, as target chunks overlap in a genome)."""

    def __init__(self, bases: np.ndarray, offsets, lens, starts=None, seq_ids=None, seq_sizes=None):
        self.bases = bases if isinstance(bases, np.ndarray) else np.frombuffer(bytes(bases), dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.lens = np.ascontiguousarray(lens, dtype=np.int32)
        n = len(self.lens)
        self.starts = np.ascontiguousarray(starts if starts is not None else np.zeros(n), dtype=np.int32)
        self.seq_ids = np.ascontiguousarray(seq_ids if seq_ids is not None else np.zeros(n), dtype=np.int32)
        if seq_sizes is None:
            seq_sizes = [int(self.lens.max()) if n else 0]
        self.seq_sizes = np.ascontiguousarray(seq_sizes, dtype=np.int32)

    def __len__(self):
        return len(self.lens)

    @classmethod
    def from_list(cls, chunks: Sequence[tuple]):
        """chunks: (bases, start, seq_id, seq_size) tuples; seq sizes are gathered per seq_id."""
        blobs = [bytes(c[0]) if not isinstance(c[0], str) else c[0].encode() for c in chunks]
        lens = np.array([len(b) for b in blobs], dtype=np.int32)
        offsets = np.zeros(len(blobs), dtype=np.int64)
        if len(blobs) > 1:
            offsets[1:] = np.cumsum(lens[:-1])
        nseq = max((c[2] for c in chunks), default=-1) + 1
        sizes = np.zeros(max(nseq, 1), dtype=np.int32)
        for c in chunks:
            sizes[c[2]] = c[3]
        bases = np.frombuffer(b"".join(blobs) or b"\0", dtype=np.uint8)
        return cls(bases, offsets, lens, [c[1] for c in chunks], [c[2] for c in chunks], sizes)

    @classmethod
    def independent(cls, bases2d: np.ndarray):
        """n chunks of equal length, each its own sequence (synthetic independent pairs)."""
        n, length = bases2d.shape
        return cls(bases2d.reshape(-1), np.arange(n, dtype=np.int64) * length, np.full(n, length, np.int32),
                   np.zeros(n, np.int32), np.arange(n, dtype=np.int32), np.full(n, length, np.int32))


class XCorrEngine:
    """One GPU context.  Mirrors the slave: load chunks once, then align blocks (t_pair) or pairs."""

    def __init__(self, cfg: Optional[Config] = None, **overrides):
        self._L = load_library()
        self.cfg = cfg if cfg is not None else default_config(**overrides)
        if cfg is not None:
            for k, v in overrides.items():
                setattr(self.cfg, k, v)
        h = C.c_void_p()
        _check(self._L.sx_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        self.N = 2 * self.cfg.t_chunk
        self._keep = [None, None]

    def close(self):
        if getattr(self, "_h", None):
            self._L.sx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- loading
    def _set(self, fn, cs: ChunkSet, slot: int = 0):
        bases = np.ascontiguousarray(cs.bases, dtype=np.uint8)
        # async_upload: the library reads the blob until the next sx_align_* call returns -- keep the (possibly
        # temporary) array and the chunk set alive until then
        self._keep[slot] = (bases, cs)
        _check(fn(self._h, bases.ctypes.data, cs.offsets.ctypes.data, cs.lens.ctypes.data, cs.starts.ctypes.data,
                  cs.seq_ids.ctypes.data, len(cs), cs.seq_sizes.ctypes.data, len(cs.seq_sizes)))

    def set_targets(self, cs: ChunkSet):
        self._set(self._L.sx_set_targets, cs)

    def set_queries(self, cs: ChunkSet):
        self._set(self._L.sx_set_queries, cs, 1)

    def set_targets_raw(self, bases_ptr: int, cs: ChunkSet):
        """Same as set_targets but reads the blob from an explicit host address (e.g. pinned memory)."""
        _check(self._L.sx_set_targets(self._h, bases_ptr, cs.offsets.ctypes.data, cs.lens.ctypes.data,
                                      cs.starts.ctypes.data, cs.seq_ids.ctypes.data, len(cs),
                                      cs.seq_sizes.ctypes.data, len(cs.seq_sizes)))

    def set_queries_raw(self, bases_ptr: int, cs: ChunkSet):
        _check(self._L.sx_set_queries(self._h, bases_ptr, cs.offsets.ctypes.data, cs.lens.ctypes.data,
                                      cs.starts.ctypes.data, cs.seq_ids.ctypes.data, len(cs),
                                      cs.seq_sizes.ctypes.data, len(cs.seq_sizes)))

    def invalidate_spectra(self):
        _check(self._L.sx_invalidate_spectra(self._h))

    def set_prob_table(self, table: np.ndarray):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.shape == (512, 2048)
        _check(self._L.sx_set_prob_table(self._h, t.ctypes.data))

    # ---- hot path
    def _collect(self, call, guess: int, out: np.ndarray = None) -> np.ndarray:
        """Runs an align call into `out` (a caller-owned RESULT_DTYPE array that is reused across calls) or into
        a fresh array of `guess` records; returns the filled prefix (a view)."""
        if out is not None:
            assert out.dtype == RESULT_DTYPE and out.flags["C_CONTIGUOUS"]
            cap = len(out)
        else:
            cap = max(guess, 1024)
            out = np.zeros(cap, dtype=RESULT_DTYPE)
        n = C.c_int64(0)
        rc = call(out, cap, n)
        if rc == SX_ERR_CAPACITY:  # never truncated: redo with the size the library asked for
            cap = int(n.value)
            out = np.zeros(cap, dtype=RESULT_DTYPE)
            rc = call(out, cap, n)
        _check(rc)
        return out[: n.value]

    def align_blocks(self, blocks: Iterable[tuple], cap_hint: int = 0, out: np.ndarray = None) -> np.ndarray:
        """blocks: (target_from, target_to, query_from, query_to, fast) with inclusive ranges (t_pair)."""
        arr = make_blocks(blocks)
        return self._collect(
            lambda o, cap, n: self._L.sx_align_blocks(self._h, arr.ctypes.data, len(arr), o.ctypes.data, cap,
                                                      C.byref(n)), cap_hint or 1 << 16, out)

    def align_pairs(self, pairs, fast: bool = False, cap_hint: int = 0, out: np.ndarray = None) -> np.ndarray:
        p = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        return self._collect(
            lambda o, cap, n: self._L.sx_align_pairs(self._h, p.ctypes.data, len(p), int(fast), o.ctypes.data,
                                                     cap, C.byref(n)), cap_hint or max(1 << 16, 4 * len(p)), out)

    # ---- taps
    def tap_signal(self, is_target: bool, chunk: int, strand: int = 0) -> np.ndarray:
        out = np.zeros((5, self.N), dtype=np.float32)
        _check(self._L.sx_tap_signal(self._h, int(is_target), chunk, strand, out.ctypes.data))
        return out

    def tap_xcorr(self, target: int, query: int, strand: int = 0) -> np.ndarray:
        out = np.zeros(self.N, dtype=np.float32)
        _check(self._L.sx_tap_xcorr(self._h, target, query, strand, out.ctypes.data))
        return out

    def tap_candidates(self, target: int, query: int, strand: int = 0, fast: bool = False) -> np.ndarray:
        out = np.zeros(self.N, dtype=np.int32)
        n = C.c_int32(0)
        _check(self._L.sx_tap_candidates(self._h, target, query, strand, int(fast), out.ctypes.data, self.N,
                                         C.byref(n)))
        return out[: n.value].copy()

    def tap_segments(self, target: int, query: int, strand: int = 0, fast: bool = False) -> np.ndarray:
        cap = 1 << 20
        out = np.zeros(cap, dtype=SEGMENT_DTYPE)
        n = C.c_int32(0)
        _check(self._L.sx_tap_segments(self._h, target, query, strand, int(fast), out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].copy()

    # ---- measurement
    def set_profiling(self, on: bool):
        _check(self._L.sx_set_profiling(self._h, int(on)))

    def stream_handle(self) -> int:
        """cudaStream_t of this context (wrap with torch.cuda.ExternalStream to record events on it)."""
        p = C.c_void_p()
        _check(self._L.sx_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def stats(self) -> dict:
        s = Stats()
        _check(self._L.sx_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        _check(self._L.sx_reset_stats(self._h))


class MultiEngine:
    """Several GPUs behind one handle (sx_multi_*): the target chunk list is cut into shard_world * n_devices
    contiguous ranges, this handle owns n_devices of them; every GPU is sent only its target range and the query
    chunks its blocks touch; records of all GPUs are gathered into one array.  devices=None: every visible GPU."""

    def __init__(self, devices=None, shard_rank: int = 0, shard_world: int = 1, cfg: Optional[Config] = None, **overrides):
        self._L = load_library()
        self.cfg = cfg if cfg is not None else default_config(**overrides)
        if cfg is not None:
            for k, v in overrides.items():
                setattr(self.cfg, k, v)
        dev = np.ascontiguousarray(devices if devices is not None else [], dtype=np.int32)
        h = C.c_void_p()
        self._check(self._L.sx_multi_create(C.byref(self.cfg), dev.ctypes.data if len(dev) else None, len(dev), shard_rank,
                                            shard_world, C.byref(h)))
        self._h = h
        self._keep = [None, None]

    def _check(self, rc: int):
        if rc != SX_OK:
            raise SatsumaError(rc, self._L.sx_multi_last_error().decode(errors="replace"))

    def close(self):
        if getattr(self, "_h", None):
            self._L.sx_multi_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_devices(self) -> int:
        return int(self._L.sx_multi_device_count(self._h))

    def target_range(self, shard: int):
        lo, hi = C.c_int32(), C.c_int32()
        self._check(self._L.sx_multi_target_range(self._h, shard, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def _set(self, fn, cs: ChunkSet, slot: int, ptr: int = 0):
        bases = np.ascontiguousarray(cs.bases, dtype=np.uint8)
        self._keep[slot] = (bases, cs)  # the handle fetches query ranges from this blob on demand
        self._check(fn(self._h, ptr or bases.ctypes.data, cs.offsets.ctypes.data, cs.lens.ctypes.data, cs.starts.ctypes.data,
                       cs.seq_ids.ctypes.data, len(cs), cs.seq_sizes.ctypes.data, len(cs.seq_sizes)))

    def set_targets(self, cs: ChunkSet, ptr: int = 0):
        self._set(self._L.sx_multi_set_targets, cs, 0, ptr)

    def set_queries(self, cs: ChunkSet, ptr: int = 0):
        self._set(self._L.sx_multi_set_queries, cs, 1, ptr)

    def set_prob_table(self, table: np.ndarray):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.shape == (512, 2048)
        self._check(self._L.sx_multi_set_prob_table(self._h, t.ctypes.data))

    def invalidate_spectra(self):
        self._check(self._L.sx_multi_invalidate_spectra(self._h))

    def align_blocks(self, blocks, out: np.ndarray = None) -> np.ndarray:
        arr = make_blocks(blocks)
        if out is None:
            out = np.zeros(1 << 16, dtype=RESULT_DTYPE)
        n = C.c_int64(0)
        rc = self._L.sx_multi_align_blocks(self._h, arr.ctypes.data, len(arr), out.ctypes.data, len(out), C.byref(n))
        if rc == SX_ERR_CAPACITY:
            out = np.zeros(int(n.value), dtype=RESULT_DTYPE)
            rc = self._L.sx_multi_align_blocks(self._h, arr.ctypes.data, len(arr), out.ctypes.data, len(out), C.byref(n))
        self._check(rc)
        return out[: n.value]

    def stats(self, shard: int = -1) -> dict:
        s = Stats()
        self._check(self._L.sx_multi_get_stats(self._h, shard, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._check(self._L.sx_multi_reset_stats(self._h))

    def stream_handle(self, shard: int = 0) -> int:
        p = C.c_void_p()
        self._check(self._L.sx_multi_stream(self._h, shard, C.byref(p)))
        return int(p.value or 0)
