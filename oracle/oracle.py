"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

Python face of the CPU oracle for the k-mer hot path.  Three independent statements of the same
algorithm live here so that they can be checked against each other and against the reference's
recorded answers:

* ``Ref``      ctypes binding of ``oracle/_ref/libkmer_ref.so`` = the reference's own ``kmer.c``
               compiled unmodified against ``oracle/pgshim`` (built by ``oracle/Makefile``).
* ``COracle``  ctypes binding of ``oracle/libkmer_oracle.so`` = the plain-C restatement
               ``oracle/kmer_oracle.c``.
* ``np_*``     numpy restatement (vectorised; used for the larger parity sizes).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module.  Nothing here is ever on the product path.

Conventions (shared with include/kmer_cuda.h): a=0 c=1 g=2 t=3, first base in the most
significant used bit pair of a uint64 code.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SO = HERE / "_ref" / "libkmer_ref.so"
ORACLE_SO = HERE / "libkmer_oracle.so"
REFERENCE_DIR = Path(os.environ.get("KMER_REFERENCE_DIR", "/root/reference"))

# error identities (the reference's SQLSTATE + message, kmer.c:33-36,117-119,151-153,179-181,311-313)
SQLSTATE_22P02 = 0x22503
SQLSTATE_22001 = 0x22001
SQLSTATE_22023 = 0x22023

ORC_OK, ORC_INVALID_DNA, ORC_KMER_TOO_LONG, ORC_INVALID_QKMER, ORC_INVALID_K, ORC_QKMER_TOO_LONG = range(6)

OP_EQUALS, OP_STARTS_WITH, OP_STARTS_WITH_OP, OP_CONTAINS, OP_CONTAINING = range(5)


def build(ref: bool = True) -> None:
    """Compile the C oracle and (when the reference tree is present) oracle/_ref."""
    subprocess.run(["make", "-s", "-C", str(HERE), "oracle"], check=True)
    if ref and (REFERENCE_DIR / "kmer.c").exists():
        subprocess.run(["make", "-s", "-C", str(HERE), "ref", f"REFERENCE={REFERENCE_DIR}"], check=True)


class RefSqlError(Exception):
    """ereport(ERROR) raised inside the reference code."""

    def __init__(self, sqlstate: int, message: str, detail: str, row: int):
        super().__init__(f"ERROR:  {message}" + (f"\nDETAIL:  {detail}" if detail else ""))
        self.sqlstate, self.message, self.detail, self.row = sqlstate, message, detail, row


class _RefError(C.Structure):
    _fields_ = [("sqlstate", C.c_int), ("row", C.c_int64), ("message", C.c_char * 128), ("detail", C.c_char * 128)]


class _RefCounts(C.Structure):
    _fields_ = [("n_distinct", C.c_uint64), ("n_kmers", C.c_uint64), ("k", C.c_int),
                ("keys", C.POINTER(C.c_char)), ("counts", C.POINTER(C.c_uint64))]


def rows_to_flat(rows) -> tuple[bytes, np.ndarray]:
    """list of str/bytes -> (flat bytes, uint64 offsets[n+1])."""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in rows]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    return b"".join(bs), off


def _as_flat(flat) -> np.ndarray:
    if isinstance(flat, (bytes, bytearray)):
        return np.frombuffer(flat, dtype=np.uint8)
    return np.ascontiguousarray(flat, dtype=np.uint8)


class Ref:
    """The reference's own C functions (kmer.c, unmodified) behind an executor stand-in."""

    def __init__(self, path: Path = REF_SO):
        if not Path(path).exists():
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        L = self.lib = C.CDLL(str(path))
        L.ref_type_roundtrip.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(_RefError)]
        L.ref_generate_kmers.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(_RefError)]
        L.ref_predicate.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(_RefError)]
        L.ref_predicate_column.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_char_p, C.c_void_p, C.POINTER(_RefError)]
        L.ref_count.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(_RefCounts), C.POINTER(_RefError)]
        L.ref_counts_free.argtypes = [C.POINTER(_RefCounts)]
        L.ref_generate_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(_RefError)]

    @staticmethod
    def _raise(e: _RefError):
        raise RefSqlError(e.sqlstate, e.message.decode(), e.detail.decode(), e.row)

    def cast(self, typ: str, text: str) -> str:
        """'text'::dna|kmer|qkmer, printed back through the type's output function."""
        which = {"dna": 0, "kmer": 1, "qkmer": 2}[typ]
        out = C.create_string_buffer(len(text) + 8)
        e = _RefError()
        if self.lib.ref_type_roundtrip(which, text.encode(), out, len(out), C.byref(e)):
            self._raise(e)
        return out.value.decode()

    def generate_kmers(self, dna: str, k: int) -> list[str]:
        cap = max(len(dna), 1)
        kk = max(k, 1)
        buf = C.create_string_buffer(cap * kk + 1)
        n = C.c_uint64()
        e = _RefError()
        if self.lib.ref_generate_kmers(dna.encode(), k, buf, cap, C.byref(n), C.byref(e)):
            self._raise(e)
        raw = buf.raw
        return [raw[i * k:(i + 1) * k].decode() for i in range(n.value)]

    def predicate(self, op: int, a: str, b: str) -> bool:
        r = C.c_int()
        e = _RefError()
        if self.lib.ref_predicate(op, a.encode(), b.encode(), C.byref(r), C.byref(e)):
            self._raise(e)
        return bool(r.value)

    def predicate_column(self, op: int, col_ascii: np.ndarray, k: int, const: str) -> np.ndarray:
        col = np.ascontiguousarray(col_ascii, dtype=np.uint8)
        m = col.size // k if k else 0
        out = np.zeros(m, dtype=np.uint8)
        e = _RefError()
        rc = self.lib.ref_predicate_column(op, col.ctypes.data, m, k, const.encode(), out.ctypes.data, C.byref(e))
        if rc:
            self._raise(e)
        return out

    def count(self, flat, off: np.ndarray, k: int, threads: int = 1):
        """GROUP BY kmer/count(*) over generate_kmers(row, k): returns (ascii keys [D,k] uint8, counts[D], n_kmers)."""
        f = _as_flat(flat)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        rc_out = _RefCounts()
        e = _RefError()
        if self.lib.ref_count(f.ctypes.data, off.ctypes.data, len(off) - 1, k, threads, C.byref(rc_out), C.byref(e)):
            self._raise(e)
        d = rc_out.n_distinct
        keys = np.frombuffer(C.string_at(rc_out.keys, d * k), dtype=np.uint8).reshape(d, k).copy() if d else np.zeros((0, max(k, 0)), np.uint8)
        counts = np.ctypeslib.as_array(rc_out.counts, shape=(d,)).copy() if d else np.zeros(0, np.uint64)
        n = rc_out.n_kmers
        self.lib.ref_counts_free(C.byref(rc_out))
        return keys, counts, n

    def generate_rows(self, flat, off: np.ndarray, k: int):
        f = _as_flat(flat)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n, cs = C.c_uint64(), C.c_uint64()
        e = _RefError()
        if self.lib.ref_generate_rows(f.ctypes.data, off.ctypes.data, len(off) - 1, k, C.byref(n), C.byref(cs), C.byref(e)):
            self._raise(e)
        return n.value, cs.value


class OracleError(Exception):
    def __init__(self, code: int, row: int = -1):
        super().__init__(f"oracle error {code} at row {row}")
        self.code, self.row = code, row


class COracle:
    """oracle/kmer_oracle.c through ctypes."""

    def __init__(self, path: Path = ORACLE_SO):
        if not Path(path).exists():
            build(ref=False)
        L = self.lib = C.CDLL(str(path))
        L.orc_validate_dna.restype = C.c_int64
        L.orc_validate_dna.argtypes = [C.c_char_p, C.c_uint64]
        L.orc_kmer_encode.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.orc_kmer_decode.argtypes = [C.c_uint64, C.c_int, C.c_char_p]
        L.orc_qkmer_parse.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p]
        L.orc_match_column.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_char_p, C.c_void_p]
        L.orc_generate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
        L.orc_count.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
        L.orc_multiset_checksum.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64),
                                            C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]

    def multiset_checksum(self, flat, off: np.ndarray, k: int) -> tuple[int, int]:
        """(sum of mix64(code) over all windows mod 2^64, number of windows) -- see orc_multiset_checksum."""
        f = _as_flat(flat)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        s, n, bad = C.c_uint64(), C.c_uint64(), C.c_int64()
        rc = self.lib.orc_multiset_checksum(f.ctypes.data, off.ctypes.data, len(off) - 1, k, C.byref(s), C.byref(n), C.byref(bad))
        if rc:
            raise OracleError(rc, bad.value)
        return int(s.value), int(n.value)

    def kmer_encode(self, text: str) -> tuple[int, int]:
        v = C.c_uint64()
        rc = self.lib.orc_kmer_encode(text.encode(), len(text), C.byref(v))
        if rc:
            raise OracleError(rc)
        return v.value, len(text)

    def kmer_decode(self, code: int, k: int) -> str:
        buf = C.create_string_buffer(k + 1)
        self.lib.orc_kmer_decode(code, k, buf)
        return buf.raw[:k].decode()

    def qkmer_parse(self, text: str) -> str:
        buf = C.create_string_buffer(len(text) + 1)
        rc = self.lib.orc_qkmer_parse(text.encode(), len(text), buf)
        if rc:
            raise OracleError(rc)
        return buf.raw[:len(text)].decode()

    def generate(self, flat, off: np.ndarray, k: int) -> np.ndarray:
        f = _as_flat(flat)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n_rows = len(off) - 1
        cap = max(int(off[-1]) if n_rows else 0, 1)
        codes = np.zeros(cap, dtype=np.uint64)
        n, bad = C.c_uint64(), C.c_int64()
        rc = self.lib.orc_generate(f.ctypes.data, off.ctypes.data, n_rows, k, codes.ctypes.data, C.byref(n), C.byref(bad))
        if rc:
            raise OracleError(rc, bad.value)
        return codes[: n.value].copy()

    def count(self, flat, off: np.ndarray, k: int):
        f = _as_flat(flat)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n_rows = len(off) - 1
        cap = max(int(off[-1]) if n_rows else 0, 1)
        keys = np.zeros(cap, dtype=np.uint64)
        counts = np.zeros(cap, dtype=np.uint64)
        d, n, bad = C.c_uint64(), C.c_uint64(), C.c_int64()
        rc = self.lib.orc_count(f.ctypes.data, off.ctypes.data, n_rows, k, keys.ctypes.data, counts.ctypes.data,
                                C.byref(d), C.byref(n), C.byref(bad))
        if rc:
            raise OracleError(rc, bad.value)
        return keys[: d.value].copy(), counts[: d.value].copy(), n.value

    def match_column(self, op: int, codes: np.ndarray, k: int, const: str, lens: np.ndarray | None = None) -> np.ndarray:
        """op: 0 equals, 1 starts_with(const prefix, col), 2 contains(const qkmer, col)."""
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        out = np.zeros(codes.size, dtype=np.uint8)
        lp = None
        if lens is not None:
            lens = np.ascontiguousarray(lens, dtype=np.uint8)
            lp = lens.ctypes.data
        rc = self.lib.orc_match_column(op, codes.ctypes.data, lp, codes.size, k, const.encode(), out.ctypes.data)
        if rc:
            raise OracleError(rc)
        return out


# --------------------------------------------------------------------------- numpy restatement

_CODE_LUT = np.full(256, 255, dtype=np.uint8)
for _i, _ch in enumerate("acgt"):
    _CODE_LUT[ord(_ch)] = _i
    _CODE_LUT[ord(_ch.upper())] = _i

# match(), kmer.h:21-53 as 4-bit sets over (a=1, c=2, g=4, t=8); 'u' names no set.
IUPAC_MASK = {"a": 1, "c": 2, "g": 4, "t": 8, "r": 5, "y": 10, "k": 12, "m": 3, "s": 6, "w": 9,
              "b": 14, "d": 13, "h": 11, "v": 7, "n": 15, "u": 0}


def np_generate(flat, off: np.ndarray, k: int) -> np.ndarray:
    """generate_kmers over all rows (kmer.c:289-351) -> uint64 codes, row-major position order."""
    f = _as_flat(flat)
    off = np.asarray(off, dtype=np.int64)
    n_rows = len(off) - 1
    lens = np.diff(off)
    codes2 = _CODE_LUT[f]
    bad_char = np.nonzero(codes2 == 255)[0]
    bad_char_row = int(np.searchsorted(off, bad_char[0], side="right") - 1) if bad_char.size else None
    short = np.nonzero(lens < k)[0] if (1 <= k <= 32) else np.arange(n_rows)
    short_row = int(short[0]) if short.size else None
    if bad_char_row is not None and (short_row is None or bad_char_row <= short_row):
        raise OracleError(ORC_INVALID_DNA, bad_char_row)
    if short_row is not None:
        raise OracleError(ORC_INVALID_K, short_row)
    n_total = int(off[-1]) if n_rows else 0
    if n_total < k or n_rows == 0:
        return np.zeros(0, dtype=np.uint64)
    acc = np.zeros(n_total - k + 1, dtype=np.uint64)
    for j in range(k):
        acc = (acc << np.uint64(2)) | codes2[j:n_total - k + 1 + j].astype(np.uint64)
    pos = np.arange(n_total - k + 1, dtype=np.int64)
    row = np.searchsorted(off, pos, side="right") - 1
    valid = pos + k <= off[row + 1]
    return acc[valid]


def np_count(flat, off: np.ndarray, k: int):
    """GROUP BY kmer / count(*): (sorted distinct codes, counts, n_kmers)."""
    codes = np_generate(flat, off, k)
    keys, counts = np.unique(codes, return_counts=True)
    return keys.astype(np.uint64), counts.astype(np.uint64), int(codes.size)


def np_mix64(x: np.ndarray) -> np.ndarray:
    """murmur3 finaliser on uint64 arrays (wraps mod 2^64), the hash of the multiset checksum."""
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
    return x


def np_table_checksum(keys: np.ndarray, counts: np.ndarray) -> int:
    """sum over groups of count * mix64(code) mod 2^64: equals COracle.multiset_checksum of the rows the table counts."""
    with np.errstate(over="ignore"):
        return int((np_mix64(keys) * np.asarray(counts, dtype=np.uint64)).sum(dtype=np.uint64))


def np_decode(codes: np.ndarray, k: int) -> np.ndarray:
    """codes -> [n, k] uint8 lower-case ASCII."""
    codes = np.asarray(codes, dtype=np.uint64)
    out = np.zeros((codes.size, k), dtype=np.uint8)
    lut = np.frombuffer(b"acgt", dtype=np.uint8)
    for j in range(k):
        out[:, j] = lut[((codes >> np.uint64(2 * (k - 1 - j))) & np.uint64(3)).astype(np.int64)]
    return out


def np_encode_ascii(kmers_ascii: np.ndarray) -> np.ndarray:
    """[n, k] ASCII -> uint64 codes."""
    a = _CODE_LUT[np.asarray(kmers_ascii, dtype=np.uint8)]
    assert (a != 255).all()
    acc = np.zeros(a.shape[0], dtype=np.uint64)
    for j in range(a.shape[1]):
        acc = (acc << np.uint64(2)) | a[:, j].astype(np.uint64)
    return acc


def np_match(op: int, codes: np.ndarray, k: int, const: str, lens: np.ndarray | None = None) -> np.ndarray:
    """op 0 equals(col,const); 1 starts_with(const,col); 2 contains(const::qkmer, col). -> uint8[m]."""
    codes = np.asarray(codes, dtype=np.uint64)
    lk = np.full(codes.size, k, dtype=np.int64) if lens is None else np.asarray(lens, dtype=np.int64)
    c = const.lower()
    lc = len(c)
    if op in (0, 1):
        ccode = np.uint64(0)
        for ch in c:
            ccode = (ccode << np.uint64(2)) | np.uint64(_CODE_LUT[ord(ch)])
        if op == 0:
            return ((lk == lc) & (codes == ccode)).astype(np.uint8)
        sh = (2 * np.maximum(lk - lc, 0)).astype(np.uint64)
        shifted = np.where(sh >= 64, np.uint64(0), codes >> np.minimum(sh, np.uint64(63)))
        return ((lk >= lc) & (shifted == ccode)).astype(np.uint8)
    ok = lk == lc
    for j, ch in enumerate(c):
        base = ((codes >> np.uint64(2 * max(lc - 1 - j, 0))) & np.uint64(3)).astype(np.int64)
        ok &= ((IUPAC_MASK[ch] >> base) & 1).astype(bool)
    return ok.astype(np.uint8)
