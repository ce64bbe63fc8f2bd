"""TEST INFRASTRUCTURE ONLY.

ctypes bindings for
  * ``liboracle.so``            -- the C restatement of the hot path (oracle/sx_oracle.c)
  * ``_ref/libsatsuma_ref.so``  -- the UNMODIFIED reference compiled from /root/reference
                                   (oracle/ref_harness.cc); optional, present when built.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this package.  The product (``satsuma2_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsatsuma_ref.so")
REF_TOOL = os.path.join(HERE, "_ref", "HomologyByXCorr_ref")  # the reference's standalone tool, unmodified
REF_KMATCH = os.path.join(HERE, "_ref", "KMatch_ref")  # the reference's k-mer seeding program, unmodified
REFERENCE_ROOT = "/root/reference"


def build(ref: bool = True) -> None:
    """Compile the C oracle (always) and the reference harness (when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", "reftool", "refkmatch", "-j8"])
        # the reference's own slave with libsatsuma_b200 bound in; needs the product
        # library, so it is (re)built only when that exists and is newer
        lib = os.path.join(os.path.dirname(HERE), "satsuma2_b200", "libsatsuma_b200.so")
        exe = os.path.join(HERE, "_ref", "HomologyByXCorrSlave_b200bind")
        bind = os.path.join(os.path.dirname(HERE), "satsuma2_b200", "host", "binding")
        deps = [lib, os.path.join(HERE, "make_refslave_b200.py"), os.path.join(HERE, "shim_check.cc")] + [
            os.path.join(bind, f) for f in sorted(os.listdir(bind))]
        if os.path.exists(lib) and (not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps)):
            subprocess.check_call(["make", "-s", "-C", HERE, "refslave_b200"])


RESULT_DTYPE = np.dtype(
    [
        ("query_id", "<u8"),
        ("target_id", "<u8"),
        ("query_size", "<u8"),
        ("qstart", "<u8"),
        ("tstart", "<u8"),
        ("len", "<u8"),
        ("reverse", "u1"),
        ("pad", "u1", (7,)),
        ("prob", "<f8"),
        ("ident", "<f8"),
    ]
)
assert RESULT_DTYPE.itemsize == 72
SEG_DTYPE = np.dtype([("start_target", "<i4"), ("start_query", "<i4"), ("len", "<i4")])


class _Params(C.Structure):
    _fields_ = [
        ("t_chunk", C.c_int32),
        ("q_chunk", C.c_int32),
        ("cutoff", C.c_double),
        ("cutoff_fast", C.c_double),
        ("min_len", C.c_int32),
        ("use_prob_table", C.c_int32),
        ("min_prob", C.c_double),
        ("table_value", C.c_double),
        ("target_total", C.c_double),
        ("prob_table", C.c_void_p),
    ]


class _Chunk(C.Structure):
    _fields_ = [
        ("bases", C.c_void_p),
        ("len", C.c_int32),
        ("start", C.c_int32),
        ("seq_id", C.c_int32),
        ("seq_size", C.c_int32),
    ]


def _b(x) -> bytes:
    if isinstance(x, (bytes, bytearray)):
        return bytes(x)
    if isinstance(x, str):
        return x.encode()
    return np.ascontiguousarray(x, dtype=np.uint8).tobytes()


def _ptr(a: np.ndarray, t=C.c_void_p):
    return a.ctypes.data_as(t)


class Oracle:
    """C restatement (oracle/sx_oracle.c)."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = self.lib = C.CDLL(ORACLE_SO)
        L.sxo_rc_base.restype = C.c_char
        L.sxo_equal.restype = C.c_double
        L.sxo_match_prob.restype = C.c_double
        L.sxo_prob_table_lookup.restype = C.c_double
        L.sxo_align_pair.restype = C.c_long
        L.sxo_align_pairs_mt.restype = C.c_long
        L.sxo_match_prob.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.sxo_prob_table_lookup.argtypes = [
            C.c_void_p, C.c_double, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.sxo_prob_table_build.argtypes = [C.c_double, C.c_void_p]
        L.sxo_findtop.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
        L.sxo_matchup.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                  C.c_void_p, C.c_int]
        L.sxo_diag.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int]

    # -- codec
    def codec_tables(self):
        acgt = np.zeros((256, 4))
        rc = np.zeros(256, dtype=np.uint8)
        for i in range(256):
            v = (C.c_double * 4)()
            self.lib.sxo_codec(i, v)
            acgt[i] = list(v)
            rc[i] = ord(self.lib.sxo_rc_base(i))
        return acgt, rc

    def equal(self, a: int, b: int) -> float:
        return self.lib.sxo_equal(a, b)

    def score(self, a: int, b: int) -> int:
        return self.lib.sxo_score(a, b)

    def revcomp(self, seq) -> bytes:
        s = _b(seq)
        out = C.create_string_buffer(len(s) + 1)
        self.lib.sxo_revcomp(s, len(s), out)
        return out.raw[: len(s)]

    # -- stages
    def encode(self, seq, N: int) -> np.ndarray:
        s = _b(seq)
        out = np.zeros((5, N), dtype=np.float32)
        self.lib.sxo_encode(s, len(s), N, _ptr(out))
        return out

    def xcorr_signals(self, tsig4: np.ndarray, qsig4: np.ndarray) -> np.ndarray:
        t = np.ascontiguousarray(tsig4, dtype=np.float32)
        q = np.ascontiguousarray(qsig4, dtype=np.float32)
        N = t.shape[1]
        out = np.zeros(N, dtype=np.float32)
        self.lib.sxo_xcorr(_ptr(t), _ptr(q), N, _ptr(out))
        return out

    def xcorr(self, tseq, qseq, N: int) -> np.ndarray:
        return self.xcorr_signals(self.encode(tseq, N)[1:], self.encode(qseq, N)[1:])

    def findtop(self, xc: np.ndarray, cutoff: float, with_env: bool = False):
        xc = np.ascontiguousarray(xc, dtype=np.float32)
        N = xc.shape[0]
        idx = np.zeros(N, dtype=np.int32)
        env = np.zeros(max(N // 256, 1), dtype=np.float64)
        n = self.lib.sxo_findtop(_ptr(xc), N, cutoff, _ptr(idx), N, _ptr(env))
        return (idx[:n].copy(), env) if with_env else idx[:n].copy()

    def diag(self, qseq, tseq, shift: int) -> np.ndarray:
        q, t = _b(qseq), _b(tseq)
        cap = 8192
        out = np.zeros(cap, dtype=SEG_DTYPE)
        n = self.lib.sxo_diag(q, len(q), t, len(t), shift, _ptr(out), cap)
        assert n <= cap
        return out[:n].copy()

    def matchup(self, qseq, tseq, xc: np.ndarray, cutoff: float) -> np.ndarray:
        q, t = _b(qseq), _b(tseq)
        xc = np.ascontiguousarray(xc, dtype=np.float32)
        cap = 1 << 16
        while True:
            out = np.zeros(cap, dtype=SEG_DTYPE)
            n = self.lib.sxo_matchup(q, len(q), t, len(t), _ptr(xc), xc.shape[0], cutoff, _ptr(out), cap)
            if n <= cap:
                return out[:n].copy()
            cap = n

    def match_prob(self, tseq, qseq, startT: int, startQ: int, length: int, target_size: float):
        ident = C.c_double()
        p = self.lib.sxo_match_prob(_b(tseq), _b(qseq), startT, startQ, length, target_size, C.byref(ident))
        return p, ident.value

    def prob_table(self, target_size: float) -> np.ndarray:
        tab = np.zeros((512, 2048), dtype=np.float64)
        self.lib.sxo_prob_table_build(target_size, _ptr(tab))
        return tab

    def prob_table_lookup(self, table, table_value, tseq, qseq, startT, startQ, length):
        ident = C.c_double()
        p = self.lib.sxo_prob_table_lookup(_ptr(table), table_value, _b(tseq), _b(qseq), startT, startQ, length,
                                           C.byref(ident))
        return p, ident.value

    # -- full path
    @staticmethod
    def make_params(t_chunk=4096, q_chunk=4096, cutoff=1.8, cutoff_fast=2.9, min_len=0, min_prob=0.99,
                    target_total=0.0, prob_table=None, table_value=0.9999):
        p = _Params()
        p.t_chunk, p.q_chunk = t_chunk, q_chunk
        p.cutoff, p.cutoff_fast = cutoff, cutoff_fast
        p.min_len = min_len
        p.use_prob_table = 1 if prob_table is not None else 0
        p.min_prob = min_prob
        p.table_value = table_value
        p.target_total = target_total
        p._keep = prob_table
        p.prob_table = prob_table.ctypes.data if prob_table is not None else None
        return p

    @staticmethod
    def _chunks(chunks):
        """chunks: list of (bases, start, seq_id, seq_size)."""
        arr = (_Chunk * len(chunks))()
        keep = []
        for i, (bases, start, seq_id, seq_size) in enumerate(chunks):
            buf = C.create_string_buffer(_b(bases), len(_b(bases)) + 1)
            keep.append(buf)
            arr[i].bases = C.cast(buf, C.c_void_p)
            arr[i].len = len(_b(bases))
            arr[i].start, arr[i].seq_id, arr[i].seq_size = start, seq_id, seq_size
        return arr, keep

    def align_pairs(self, params, targets, queries, pairs, fast=False, threads=1) -> np.ndarray:
        tarr, tk = self._chunks(targets)
        qarr, qk = self._chunks(queries)
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        cap = max(1024, 64 * len(pairs))
        while True:
            out = np.zeros(cap, dtype=RESULT_DTYPE)
            n = self.lib.sxo_align_pairs_mt(C.byref(params), tarr, qarr, _ptr(pairs), C.c_long(len(pairs)),
                                            int(fast), threads, _ptr(out), C.c_long(cap))
            if n <= cap:
                return out[:n].copy()
            cap = n


class Reference:
    """The unmodified reference (oracle/_ref/libsatsuma_ref.so). Raises if not built."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        L = self.lib = C.CDLL(REF_SO)
        L.ref_configure.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double]
        L.ref_target_total.restype = C.c_double
        L.ref_set_target_total.argtypes = [C.c_double]
        L.ref_align_block.restype = C.c_long
        L.ref_align_block.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long]
        L.ref_align_pairs_mt.restype = C.c_long
        L.ref_align_pairs_mt.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_long, C.c_void_p]
        L.ref_match_prob.restype = C.c_double
        L.ref_match_prob.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_double, C.c_void_p]
        L.ref_ident.restype = C.c_double
        L.ref_ident.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_diag.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.ref_findtop.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int]
        L.ref_matchup.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                  C.c_void_p, C.c_int]
        L.ref_prob_table.argtypes = [C.c_double, C.c_double, C.c_void_p]
        L.ref_set_chunks.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p]
        assert L.ref_sizeof_t_result() == 72 and L.ref_sizeof_t_pair() == 28

    def configure(self, t_chunk=4096, q_chunk=4096, cutoff=1.8, cutoff_fast=2.9, min_len=0,
                  use_prob_table=False, min_prob_flag=0.9999):
        self.lib.ref_configure(t_chunk, q_chunk, cutoff, cutoff_fast, min_len, int(use_prob_table), min_prob_flag)

    def build_prob_table(self):
        self.lib.ref_build_prob_table()

    def load_fasta(self, target_fasta: str, query_fasta: str):
        self.lib.ref_load_fasta(target_fasta.encode(), query_fasta.encode())

    def target_total(self) -> float:
        return self.lib.ref_target_total()

    def chunks(self, is_target: bool):
        """-> list of (bases, start, seq_id, seq_size) as the reference chunked them."""
        out = []
        t = int(is_target)
        for i in range(self.lib.ref_num_chunks(t)):
            n = self.lib.ref_chunk_len(t, i)
            buf = C.create_string_buffer(n + 1)
            self.lib.ref_chunk_bases(t, i, buf)
            sid = self.lib.ref_chunk_seq(t, i)
            out.append((buf.raw[:n], self.lib.ref_chunk_start(t, i), sid, self.lib.ref_seq_size(t, sid)))
        return out

    def seq_sizes(self, is_target: bool):
        t = int(is_target)
        return [self.lib.ref_seq_size(t, i) for i in range(self.lib.ref_num_seqs(t))]

    def set_chunks(self, is_target: bool, chunks, seq_sizes):
        """chunks: list of (bases, start, seq_id, seq_size)."""
        blob = b"".join(_b(c[0]) for c in chunks)
        lens = np.array([len(_b(c[0])) for c in chunks], dtype=np.int32)
        offs = np.zeros(len(chunks), dtype=np.int64)
        if len(chunks) > 1:
            offs[1:] = np.cumsum(lens[:-1])
        ids = np.array([c[2] for c in chunks], dtype=np.int32)
        starts = np.array([c[1] for c in chunks], dtype=np.int32)
        ss = np.array(seq_sizes, dtype=np.int32)
        self.lib.ref_set_chunks(int(is_target), len(chunks), blob, _ptr(offs), _ptr(lens), _ptr(ids), _ptr(starts),
                                len(ss), _ptr(ss))

    def align_block(self, tFrom, tTo, qFrom, qTo, fast=False) -> np.ndarray:
        cap = 1 << 16
        while True:
            out = np.zeros(cap, dtype=RESULT_DTYPE)
            n = self.lib.ref_align_block(tFrom, tTo, qFrom, qTo, int(fast), _ptr(out), cap)
            if n <= cap:
                return out[:n].copy()
            cap = n  # deterministic: redo with room

    def align_pairs_mt(self, tpairs, threads: int):
        """tpairs: n x 5 int32 (tFrom,tTo,qFrom,qTo,fast). -> (results, seconds)"""
        tp = np.ascontiguousarray(tpairs, dtype=np.int32).reshape(-1, 5)
        cap = max(1 << 16, 256 * len(tp))
        out = np.zeros(cap, dtype=RESULT_DTYPE)
        secs = C.c_double()
        n = self.lib.ref_align_pairs_mt(_ptr(tp), len(tp), threads, _ptr(out), cap, C.byref(secs))
        return out[: min(n, cap)].copy(), secs.value, n

    def signal(self, seq, N: int) -> np.ndarray:
        s = _b(seq)
        out = np.zeros((5, N), dtype=np.float32)
        self.lib.ref_signal(s, len(s), N, _ptr(out))
        return out

    def revcomp(self, seq) -> bytes:
        s = _b(seq)
        out = C.create_string_buffer(len(s) + 1)
        self.lib.ref_revcomp(s, len(s), out)
        return out.raw[: len(s)]

    def xcorr(self, tseq, qseq, N: int) -> np.ndarray:
        t, q = _b(tseq), _b(qseq)
        out = np.zeros(N, dtype=np.float32)
        self.lib.ref_xcorr(t, len(t), q, len(q), N, _ptr(out))
        return out

    def fft(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        f = np.zeros_like(x)
        self.lib.ref_fft(_ptr(x), x.shape[0], _ptr(f))
        return f

    def findtop(self, xc: np.ndarray, cutoff: float) -> np.ndarray:
        xc = np.ascontiguousarray(xc, dtype=np.float32)
        idx = np.zeros(xc.shape[0], dtype=np.int32)
        n = self.lib.ref_findtop(_ptr(xc), xc.shape[0], cutoff, _ptr(idx), xc.shape[0])
        return idx[:n].copy()

    def matchup(self, qseq, tseq, xc: np.ndarray, cutoff: float) -> np.ndarray:
        q, t = _b(qseq), _b(tseq)
        xc = np.ascontiguousarray(xc, dtype=np.float32)
        cap = 1 << 16
        while True:
            out = np.zeros(cap, dtype=SEG_DTYPE)
            n = self.lib.ref_matchup(q, len(q), t, len(t), _ptr(xc), xc.shape[0], cutoff, _ptr(out), cap)
            if n <= cap:
                return out[:n].copy()
            cap = n

    def diag(self, qseq, tseq, shift: int) -> np.ndarray:
        q, t = _b(qseq), _b(tseq)
        out = np.zeros(8192, dtype=SEG_DTYPE)
        n = self.lib.ref_diag(q, len(q), t, len(t), shift, _ptr(out), 8192)
        return out[:n].copy()

    def match_prob(self, tseq, qseq, startT, startQ, length, target_size):
        t, q = _b(tseq), _b(qseq)
        ident = C.c_double()
        p = self.lib.ref_match_prob(t, len(t), q, len(q), startT, startQ, length, target_size, C.byref(ident))
        return p, ident.value

    def ident(self, qseq, tseq, startT, startQ, length) -> float:
        q, t = _b(qseq), _b(tseq)
        return self.lib.ref_ident(q, len(q), t, len(t), startT, startQ, length)

    def prob_table(self, target_size: float, cutoff: float) -> np.ndarray:
        tab = np.zeros((512, 2048), dtype=np.float64)
        self.lib.ref_prob_table(target_size, cutoff, _ptr(tab))
        return tab

    def read_match_file(self, path: str):
        """MultiMatches::Read of the reference -> (n x 10 array, n_targets, n_queries)."""
        self.lib.ref_read_match_file.restype = C.c_long
        self.lib.ref_read_match_file.argtypes = [C.c_char_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
        cap = 1 << 20
        out = np.zeros((cap, 10), dtype=np.float64)
        nt, nq = C.c_int(), C.c_int()
        n = self.lib.ref_read_match_file(path.encode(), _ptr(out), cap, C.byref(nt), C.byref(nq))
        return out[: max(n, 0)].copy(), nt.value, nq.value

    def sort_collapse(self, recs: np.ndarray, collapse: bool = True) -> np.ndarray:
        """MultiMatches::Sort (+ Collapse) of the reference on n x 10 records (layout of read_match_file)."""
        self.lib.ref_sort_collapse.restype = C.c_long
        self.lib.ref_sort_collapse.argtypes = [C.c_void_p, C.c_long, C.c_int]
        io = np.ascontiguousarray(recs, dtype=np.float64).copy()
        k = self.lib.ref_sort_collapse(_ptr(io), len(io), int(collapse))
        return io[:k].copy()

    def chain(self, recs: np.ndarray, target_sizes, query_sizes, dups: bool = False) -> np.ndarray:
        """RunMatchDynProg (dups: RunMatchDynProgMult) of the reference on sorted + collapsed n x 10 records."""
        self.lib.ref_chain.restype = C.c_long
        self.lib.ref_chain.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        io = np.ascontiguousarray(recs, dtype=np.float64).copy()
        ts = np.ascontiguousarray(target_sizes, dtype=np.int32)
        qs = np.ascontiguousarray(query_sizes, dtype=np.int32)
        k = self.lib.ref_chain(_ptr(io), len(io), len(ts), len(qs), _ptr(ts), _ptr(qs), int(dups))
        return io[:k].copy()

    def codec(self):
        acgt = np.zeros((128, 4))
        rc = np.zeros(128, dtype=np.uint8)
        eq = np.zeros((128, 128))
        amb = np.zeros((128, 128))
        self.lib.ref_codec(_ptr(acgt), _ptr(rc), _ptr(eq), _ptr(amb))
        return acgt, rc, eq, amb


def have_reference() -> bool:
    return os.path.exists(REF_SO)
