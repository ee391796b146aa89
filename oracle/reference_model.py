"""CPU restatement of PasaLab/marlin's dense multiply / transpose / add path (numpy + a small C core).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package; the product (marlin_b200/) never does.

It restates the *algorithm* of the reference, RDD semantics included, on plain Python containers:
an RDD[(BlockID, SubMatrix)] is a list of ((row, col), ndarray) pairs (Fortran-ordered float64,
Breeze's column-major DenseMatrix), an RDD[(Long, BDV)] is a list of (index, 1-D ndarray) pairs.
Every function cites the reference lines it follows (paths relative to
/root/reference/src/main/scala/edu/nju/pasalab/marlin/).

Arithmetic backends (`gemm=`):
  "f2j"  — oracle/marlin_oracle.c: reference-BLAS dgemm loop order, no FMA (netlib-java's default
           pure-Java F2J backend, README.md:31).  Use for small cases.
  "blas" — numpy matmul (OpenBLAS dgemm, all host threads): netlib-java with native BLAS installed.
           Use for large cases and for the timed CPU baseline.
Parity status: pinned against the reference's own golden vectors (DistributedMatrixSuite.scala) in
tests/test_oracle_golden.py — all exact small integers.  For non-integer fp64 data and for the random
generator the reference holds no golden vectors => "parity unpinned" for those.
"""
from __future__ import annotations

import ctypes as C
import math
import re
import subprocess
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libmarlin_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """gcc the C core (see oracle/Makefile). -ffp-contract=off: the JVM never fuses multiply-add."""
    src = _HERE / "marlin_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-s", "-C", str(_HERE)], check=True)
    return _SO


def clib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(_SO))
        dp = C.POINTER(C.c_double)
        lib.mo_dgemm_f2j.restype = C.c_int
        lib.mo_dgemm_f2j.argtypes = [C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_double, dp, C.c_long, C.c_int,
                                     dp, C.c_long, C.c_int, C.c_double, dp, C.c_long, C.c_int]
        lib.mo_dgemv_f2j.restype = C.c_int
        lib.mo_dgemv_f2j.argtypes = [C.c_char, C.c_int, C.c_int, C.c_double, dp, C.c_long, C.c_int, dp, C.c_double, dp]
        lib.mo_ddot_f2j.restype = C.c_double
        lib.mo_ddot_f2j.argtypes = [C.c_long, dp, dp]
        lib.mo_binary.argtypes = [C.c_int, C.c_long, dp, dp, dp]
        lib.mo_transpose_copy.argtypes = [C.c_int, C.c_int, dp, C.c_long, dp]
        lib.mo_hash_seed.restype = C.c_int64
        lib.mo_hash_seed.argtypes = [C.c_int64]
        lib.mo_uniform_fill.argtypes = [C.c_int64, C.c_long, C.c_long, C.c_double, C.c_double, dp]
        lib.mo_uniform_fill_from_state.argtypes = [C.c_uint64, C.c_long, C.c_double, C.c_double, dp]
        lib.mo_java_random_longs.argtypes = [C.c_int64, C.c_int, C.POINTER(C.c_int64)]
        lib.mo_split_method.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_int)]
        _lib = lib
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f(a) -> np.ndarray:
    return np.asfortranarray(a, dtype=np.float64)


# --------------------------------------------------------------------------------------------
# Local block kernels (L2 in SURVEY.md): SubMatrix.multiply / add / ... on Breeze BDM[Double]
# --------------------------------------------------------------------------------------------
def dgemm_f2j(transa: str, transb: str, m: int, n: int, k: int, alpha: float, a: np.ndarray, a_off: int, lda: int,
              b: np.ndarray, b_off: int, ldb: int, beta: float, c: np.ndarray, c_off: int, ldc: int) -> None:
    """netlib BLAS.dgemm on flat double arrays (the third-party seam of SURVEY §8b-1)."""
    info = clib().mo_dgemm_f2j(transa.encode(), transb.encode(), m, n, k, alpha, _dp(a), a_off, lda, _dp(b), b_off, ldb,
                               beta, _dp(c), c_off, ldc)
    if info:
        raise ValueError(f"dgemm: illegal argument {info}")


def block_multiply(a: np.ndarray, b: np.ndarray, gemm: str = "f2j") -> np.ndarray:
    """SubMatrix.multiply (matrix/SubMatrix.scala:87-91): `denseBlock * other.denseBlock` -> Breeze ->
    dgemm(transString(a), transString(b), a.rows, b.cols, a.cols, 1.0, a.data, a.offset, a.majorStride, ...,
    0.0, c.data, 0, c.rows).  A numpy array that is C-contiguous plays the role of an isTranspose view."""
    if a.shape[1] != b.shape[0]:
        raise ValueError(f"Dimension mismatch: {a.shape[1]} vs {b.shape[0]}")
    m, k = a.shape
    n = b.shape[1]
    if gemm == "blas":
        return _f(np.matmul(a, b))
    if gemm != "f2j":
        raise ValueError(gemm)

    def operand(x):
        if x.flags.f_contiguous:
            return "N", x, max(1, x.shape[0])
        if x.flags.c_contiguous:          # Breeze `.t` view: data is the column-major array of x^T
            return "T", x, max(1, x.shape[1])
        xf = _f(x)                         # Breeze copies non-contiguous slices before calling BLAS
        return "N", xf, max(1, xf.shape[0])

    ta, abuf, lda = operand(a)
    tb, bbuf, ldb = operand(b)
    c = np.zeros((m, n), order="F")
    dgemm_f2j(ta, tb, m, n, k, 1.0, abuf.reshape(-1, order="A"), 0, lda, bbuf.reshape(-1, order="A"), 0, ldb, 0.0,
              c.reshape(-1, order="F"), 0, max(1, m))
    return c


def block_multiply_vector(a: np.ndarray, x: np.ndarray) -> np.ndarray:
    """SubMatrix.multiply(v: Vector) (matrix/SubMatrix.scala:131-139): `denseBlock * v.inner.get` -> Breeze ->
    dgemv(transString(a), storedRows, storedCols, 1.0, a.data, a.offset, a.majorStride, x, 0.0, y).  A C-contiguous
    array plays the isTranspose view, as in block_multiply."""
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    if a.shape[1] != x.shape[0]:
        raise ValueError(f"Dimension mismatch: {a.shape[1]} vs {x.shape[0]}")
    y = np.zeros(a.shape[0])
    if a.size == 0:
        return y
    if a.flags.f_contiguous:
        info = clib().mo_dgemv_f2j(b"N", a.shape[0], a.shape[1], 1.0, _dp(a), 0, max(1, a.shape[0]), _dp(x), 0.0, _dp(y))
    elif a.flags.c_contiguous:
        info = clib().mo_dgemv_f2j(b"T", a.shape[1], a.shape[0], 1.0, _dp(a), 0, max(1, a.shape[1]), _dp(x), 0.0, _dp(y))
    else:
        af = _f(a)
        info = clib().mo_dgemv_f2j(b"N", af.shape[0], af.shape[1], 1.0, _dp(af), 0, max(1, af.shape[0]), _dp(x), 0.0, _dp(y))
    if info:
        raise ValueError(f"dgemv: illegal argument {info}")
    return y


def vector_dot(x: np.ndarray, y: np.ndarray) -> float:
    """Breeze `v.t * w` -> ddot (matrix/DistributedVector.scala:167)."""
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    if x.shape != y.shape:
        raise ValueError("the length of these two vectors are not the same")
    return float(clib().mo_ddot_f2j(x.shape[0], _dp(x), _dp(y)))


def vector_outer(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Breeze `v * w.t` (matrix/DistributedVector.scala:157): the column vector as an n x 1 matrix times the 1 x n row,
    i.e. dgemm with k = 1."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
    y = np.asarray(y, dtype=np.float64).reshape(1, -1)
    return block_multiply(_f(x), _f(y))


def block_add(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """SubMatrix.add (matrix/SubMatrix.scala:41-45): `this.denseBlock + other.denseBlock` (new matrix)."""
    if a.shape != b.shape:
        raise ValueError("matrix dimension mismatch")
    return _f(a + b)


def block_subtract(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """matrix/SubMatrix.scala:60-64"""
    if a.shape != b.shape:
        raise ValueError("matrix dimension mismatch")
    return _f(a - b)


def block_transpose(a: np.ndarray) -> np.ndarray:
    """`x._2.denseBlock.t.copy` (matrix/BlockMatrix.scala:517): materialised column-major transpose."""
    a = _f(a)
    out = np.empty((a.shape[1], a.shape[0]), order="F")
    if a.size:
        clib().mo_transpose_copy(a.shape[0], a.shape[1], _dp(a), max(1, a.shape[0]), _dp(out))
    return out


# --------------------------------------------------------------------------------------------
# Integer host logic
# --------------------------------------------------------------------------------------------
def split_method(m: int, k: int, n: int, cores: int) -> Tuple[int, int, int]:
    """MTUtils.splitMethod (utils/MTUtils.scala:150-175)."""
    out = (C.c_int * 3)()
    clib().mo_split_method(m, k, n, cores, out)
    return out[0], out[1], out[2]


def regrid_split_method(old_range: Sequence[Tuple[int, int]], new_sub_blk: int):
    """MTUtils.splitMethod(oldRange, newSubBlk) (utils/MTUtils.scala:182-202)."""
    status = []
    for (start, end) in old_range:
        start_id = start // new_sub_blk
        end_id = end // new_sub_blk
        num = end_id - start_id + 1
        buf = []
        tmp = 0
        for j in range(num):
            tmp_end = min((j + start_id + 1) * new_sub_blk - 1 - start, end - start)
            buf.append((j + start_id, (tmp, tmp_end), ((tmp + start) % new_sub_blk, (tmp_end + start) % new_sub_blk)))
            tmp = tmp_end + 1
        status.append(buf)
    return status


def mult_seq(i: int, j: int, kk: int, m: int, k: int, n: int) -> int:
    """seq of matrix/BlockMatrix.scala:163,168 = the partition of rdd/MatrixMultPartitioner.scala:12-22."""
    return i * n * k + j * k + kk


def hash_seed(seed: int) -> int:
    return int(clib().mo_hash_seed(seed))


def java_random_longs(seed: int, n: int) -> List[int]:
    out = (C.c_int64 * n)()
    clib().mo_java_random_longs(seed, n, out)
    return [int(v) for v in out]


def uniform_stream(partition_seed: int, first: int, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """n successive UniformGenerator(lo,hi).nextValue() after setSeed(partition_seed), skipping `first`."""
    out = np.empty(n)
    if n:
        clib().mo_uniform_fill(partition_seed, first, n, lo, hi, _dp(out))
    return out


_M64 = (1 << 64) - 1


def _xs_step(s: int) -> int:
    """XORShiftRandom.next's state update (utils/RandomDataGenerator.scala:121-127) on a Python int."""
    s ^= (s << 21) & _M64
    s ^= s >> 35
    s ^= (s << 4) & _M64
    return s


def xorshift_jump(state: int, steps: int) -> int:
    """state after `steps` updates, by square-and-multiply on the 64x64 GF(2) matrix of the (linear) update —
    an independent check of the device kernel's jump-ahead tables."""
    def apply(cols, v):
        out, b = 0, 0
        while v:
            if v & 1:
                out ^= cols[b]
            v >>= 1
            b += 1
        return out
    power = [_xs_step(1 << b) for b in range(64)]           # columns of T
    while steps:
        if steps & 1:
            state = apply(power, state)
        steps >>= 1
        if steps:
            power = [apply(power, c) for c in power]         # T^(2k) = T^k . T^k
    return state


def uniform_stream_far(partition_seed: int, first: int, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """uniform_stream for offsets too far to step through: jump 2*first updates, then generate sequentially."""
    state = xorshift_jump(hash_seed(partition_seed) & _M64, 2 * first)
    out = np.empty(n)
    if n:
        clib().mo_uniform_fill_from_state(state, n, lo, hi, _dp(out))
    return out


def _ceil_div_d(total: int, parts: int) -> int:
    return int(math.ceil(float(total) / float(parts)))


# --------------------------------------------------------------------------------------------
# Distributed matrices (L4): lists stand in for RDDs
# --------------------------------------------------------------------------------------------
class BlockMatrix:
    """matrix/BlockMatrix.scala:28-67.  blocks: list of ((row, col), ndarray)."""

    def __init__(self, blocks, n_rows: int = 0, n_cols: int = 0, blks_by_row: int = 0, blks_by_col: int = 0):
        self.blocks = [((int(r), int(c)), _f(b)) for (r, c), b in blocks]
        self._n_rows, self._n_cols, self._by_row, self._by_col = n_rows, n_cols, blks_by_row, blks_by_col

    # :36-65 — lazily derived dims; `reduce` on an empty RDD throws (DistributedMatrixSuite "empty rows")
    def num_rows(self) -> int:
        if self._n_rows <= 0:
            vals = [b.shape[0] for (r, c), b in self.blocks if c == 0]
            if not vals:
                raise RuntimeError("empty collection")
            self._n_rows = sum(vals)
        return self._n_rows

    def num_cols(self) -> int:
        if self._n_cols <= 0:
            vals = [b.shape[1] for (r, c), b in self.blocks if r == 0]
            if not vals:
                raise RuntimeError("empty collection")
            self._n_cols = sum(vals)
        return self._n_cols

    def num_blks_by_row(self) -> int:
        if self._by_row <= 0:
            self._by_row = sum(1 for (r, c), _ in self.blocks if c == 0)
        return self._by_row

    def num_blks_by_col(self) -> int:
        if self._by_col <= 0:
            self._by_col = sum(1 for (r, c), _ in self.blocks if r == 0)
        return self._by_col

    def to_breeze(self) -> np.ndarray:
        """:70-85 — placement uses the ceil block size for every block."""
        m, n = self.num_rows(), self.num_cols()
        rl = _ceil_div_d(m, self.num_blks_by_row())
        cl = _ceil_div_d(n, self.num_blks_by_col())
        mat = np.zeros((m, n), order="F")
        for (r, c), b in self.blocks:
            mat[r * rl:r * rl + b.shape[0], c * cl:c * cl + b.shape[1]] = b
        return mat

    def multiply(self, other: "BlockMatrix", gemm: str = "f2j", reduce_order: str = "ascending") -> "BlockMatrix":
        """:149-220.  Replicate A blocks n times / B blocks m times with seq keys, join inside the
        m*k*n partitions, one dgemm per partition, reduceByKey over kk.  Spark fixes no order for the
        k-way sum; `reduce_order` picks ascending / descending kk so tests can bound the spread."""
        if self.num_cols() != other.num_rows():
            raise ValueError(f"Dimension mismatch during matrix-matrix multiplication: {self.num_cols()} vs {other.num_rows()}")
        if self.num_blks_by_col() == other.num_blks_by_row():
            m, k, n = self.num_blks_by_row(), self.num_blks_by_col(), other.num_blks_by_col()
            parts: Dict[int, dict] = {}
            for (r, c), blk in self.blocks:          # :161-165
                for j in range(n):
                    parts.setdefault(r * n * k + j * k + c, {})["a"] = ((r, j), blk)
            for (r, c), blk in other.blocks:         # :166-171
                for i in range(m):
                    parts.setdefault(i * n * k + c * k + r, {})["b"] = ((i, c), blk)
            partial: Dict[Tuple[int, int], List[Tuple[int, np.ndarray]]] = {}
            for seq in sorted(parts):                # join: partitions holding both sides
                p = parts[seq]
                if "a" in p and "b" in p:
                    key = p["a"][0]
                    partial.setdefault(key, []).append((seq, block_multiply(p["a"][1], p["b"][1], gemm)))
            result = []
            for key, lst in partial.items():          # :177 reduceByKey((a, b) => a.add(b))
                lst.sort(key=lambda t: t[0], reverse=(reduce_order == "descending"))
                acc = lst[0][1]
                for _, blk in lst[1:]:
                    acc = block_add(acc, blk)
                result.append((key, acc))
            return BlockMatrix(result, self.num_rows(), other.num_cols(), m, n)
        if self.num_blks_by_col() % other.num_blks_by_row() == 0:      # :187-201
            self._check_even_cols()
            ratio = self.num_blks_by_col() // other.num_blks_by_row()
            blks = []
            for (r, c), mat in other.blocks:
                for i in range(ratio):
                    blks.append(((r * ratio + i, c), mat[i * mat.shape[0] // ratio:(i + 1) * mat.shape[0] // ratio, :]))
            return self.multiply(BlockMatrix(blks), gemm, reduce_order)
        if other.num_blks_by_row() % self.num_blks_by_col() == 0:      # :202-216 (slices rows of `this`, as written)
            self._check_even_cols()
            ratio = other.num_blks_by_row() // self.num_blks_by_col()
            blks = []
            for (r, c), mat in self.blocks:
                for i in range(ratio):
                    blks.append(((r * ratio + i, c), mat[i * mat.shape[0] // ratio:(i + 1) * mat.shape[0] // ratio, :]))
            return BlockMatrix(blks).multiply(other, gemm, reduce_order)
        raise ValueError("currently not supported for the two dimension of matrices")

    def _check_even_cols(self):
        if self.num_cols() % self.num_blks_by_col() != 0:
            raise ValueError("only supported BlockMatrix which all the sub-matrices have the same cols")
        if (self.num_cols() // self.num_blks_by_col()) % 2 != 0:
            raise ValueError("only supported sub-matrices with even number cols")

    def multiply_split(self, other, split_mode: Tuple[int, int, int], gemm: str = "f2j") -> "BlockMatrix":
        """:131-147"""
        if self.num_cols() != other.num_rows():
            raise ValueError("Dimension mismatch during matrix-matrix multiplication")
        m, k, n = split_mode
        return self.to_block_matrix(m, k).multiply(other.to_block_matrix(k, n), gemm)

    def multiply_local(self, B: np.ndarray, gemm: str = "f2j") -> "BlockMatrix":
        """multiply(B: BDM) :280-303 — broadcast B; reduce over column blocks when there are several."""
        if self.num_cols() != B.shape[0]:
            raise ValueError(f"Dimension mismatch during matrix-matrix multiplication: {self.num_cols()} vs {B.shape[0]}")
        B = _f(B)
        if self.num_blks_by_col() == 1:
            res = [((r, c), block_multiply(blk, B, gemm)) for (r, c), blk in self.blocks]
            return BlockMatrix(res, self.num_rows(), B.shape[1], self.num_blks_by_row(), self.num_blks_by_col())
        col_blk = _ceil_div_d(self.num_cols(), self.num_blks_by_col())
        acc: Dict[Tuple[int, int], np.ndarray] = {}
        for (r, c), blk in sorted(self.blocks, key=lambda t: t[0]):
            start = c * col_blk
            end = self.num_cols() if (c + 1) * col_blk > self.num_cols() else (c + 1) * col_blk
            p = block_multiply(blk, B[start:end, :], gemm)
            acc[(r, 0)] = block_add(acc[(r, 0)], p) if (r, 0) in acc else p
        # the reference reports numBlksByCol() here although all keys have column 0 (:301)
        return BlockMatrix(list(acc.items()), self.num_rows(), B.shape[1], self.num_blks_by_row(), self.num_blks_by_col())

    def multiply_dist_vector(self, v: "DistributedVector") -> "DistributedVector":
        """multiply(v: DistributedVector) :240-259 — every vector piece goes to the blocks of its block column, block x
        piece, reduceByKey(add) over the block row (ascending column here).  The result is labelled with v's length
        and split count (:252,257), as the reference does."""
        if self.num_cols() != v.length():
            raise ValueError(f"Dimension mismatch during matrix-matrix multiplication {self.num_cols()} v.s {v.length()}")
        if self.num_blks_by_col() != v.split_num():
            raise ValueError("not supported matrix or vector")
        pieces = dict(v.vectors)
        acc: Dict[int, np.ndarray] = {}
        for (r, c), blk in sorted(self.blocks, key=lambda t: t[0]):
            p = block_multiply_vector(blk, pieces[c])
            acc[r] = acc[r] + p if r in acc else p
        return DistributedVector(sorted(acc.items()), v.length(), v.split_num())

    def multiply_vector(self, v: np.ndarray) -> "DistributedVector":
        """multiply(v: BDV[Double]) :265-274 — broadcast vector, one block column only."""
        v = np.asarray(v, dtype=np.float64).reshape(-1)
        if self.num_cols() != v.shape[0]:
            raise ValueError(f"matrix columns size {self.num_cols()} not support vector length {v.shape[0]}")
        if self.num_blks_by_col() != 1:
            raise ValueError("should not split the matrix by column")
        res = [(r, block_multiply_vector(blk, v)) for (r, c), blk in sorted(self.blocks, key=lambda t: t[0])]
        return DistributedVector(res, self.num_rows(), self.num_blks_by_row())

    def multiply_auto(self, other, cores: int, broadcast_threshold: int = 300, gemm: str = "f2j"):
        """multiply(other, cores, broadcastThreshold) :87-122"""
        if self.num_cols() != other.num_rows():
            raise ValueError(f"Dimension mismatch during matrix-matrix multiplication: {self.num_cols()} vs {other.num_rows()}")
        bsize = _jvm_int(broadcast_threshold * 1024 * 1024) // 8
        if other.num_rows() * other.num_cols() <= bsize:
            return self.multiply_local(other.to_breeze(), gemm)
        if self.num_rows() * self.num_cols() <= bsize:
            if isinstance(other, DenseVecMatrix):
                return other.multiply_local(self.to_breeze(), gemm)       # :97-98 (operand-order quirk, kept)
            raise NotImplementedError("multiplyBy (BlockMatrix.scala:309-335) is outside the hot-path table")
        if isinstance(other, DenseVecMatrix) and _squareish(self.num_rows(), self.num_cols(), other.num_cols()):
            s = int(math.floor(math.pow(3 * cores, 1.0 / 3.0)))
            return self.multiply_split(other, (s, s, s), gemm)
        return self.multiply_split(other, split_method(self.num_rows(), self.num_cols(), other.num_cols(), cores), gemm)

    def scalar(self, op: str, b: float) -> "BlockMatrix":
        """add(b) :368-371, subtract(b) :404-407, multiply(b) :229-232, divide(b) :432-435,
        subtractBy :414-424, divideBy :442-452."""
        f = {"add": lambda x: x + b, "subtract": lambda x: x - b, "multiply": lambda x: x * b, "divide": lambda x: x / b,
             "subtractBy": lambda x: b - x, "divideBy": lambda x: b / x}[op]
        return BlockMatrix([(k, _f(f(v))) for k, v in self.blocks], self.num_rows(), self.num_cols(),
                           self.num_blks_by_row(), self.num_blks_by_col())

    def add(self, other, subtract: bool = False):
        """add :344-360 / subtract :380-396"""
        if self.num_rows() != other.num_rows() or self.num_cols() != other.num_cols():
            raise ValueError("matrix dimension mismatch")
        op = block_subtract if subtract else block_add
        if isinstance(other, DenseVecMatrix):
            return self.to_dense_vec_matrix().add(other, subtract)
        if self.num_blks_by_row() != other.num_blks_by_row() or self.num_blks_by_col() != other.num_blks_by_col():
            return self.to_dense_vec_matrix().add(other.to_dense_vec_matrix(), subtract)
        mine = dict(self.blocks)
        res = [(key, op(mine[key], blk)) for key, blk in other.blocks if key in mine]      # join
        return BlockMatrix(res, self.num_rows(), self.num_cols(), self.num_blks_by_row(), self.num_blks_by_col())

    def dot_product(self, other):
        """:486-507 (element-wise product; same-grid branch)"""
        if self.num_rows() != other.num_rows() or self.num_cols() != other.num_cols():
            raise ValueError("dimension mismatch")
        if isinstance(other, DenseVecMatrix):
            return self.to_dense_vec_matrix().dot_product(other)
        mine = dict(self.blocks)
        res = [(key, _f(mine[key] * blk)) for key, blk in other.blocks if key in mine]
        return BlockMatrix(res, self.num_rows(), self.num_cols(), self.num_blks_by_row(), self.num_blks_by_col())

    def sum(self) -> float:
        """:467-472 — per block `data.sum` (sequential left fold), then reduce(_ + _)."""
        total = None
        for _, blk in self.blocks:
            s = 0.0
            for v in blk.reshape(-1, order="F"):
                s += float(v)
            total = s if total is None else total + s
        if total is None:
            raise RuntimeError("empty collection")
        return total

    def transpose(self) -> "BlockMatrix":
        """:514-523"""
        res = [((c, r), block_transpose(blk)) for (r, c), blk in self.blocks]
        return BlockMatrix(res, self.num_cols(), self.num_rows(), self.num_blks_by_col(), self.num_blks_by_row())

    def to_dense_vec_matrix(self) -> "DenseVecMatrix":
        """:575-594"""
        rl = _ceil_div_d(self.num_rows(), self.num_blks_by_row())
        cl = _ceil_div_d(self.num_cols(), self.num_blks_by_col())
        rows: Dict[int, np.ndarray] = {}
        for (r, c), blk in self.blocks:
            for i in range(blk.shape[0]):
                vec = rows.setdefault(r * rl + i, np.zeros(self.num_cols()))
                vec[cl * c:cl * c + blk.shape[1]] = blk[i, :]
        return DenseVecMatrix(list(rows.items()))

    def to_block_matrix(self, new_by_row: int, new_by_col: int) -> "BlockMatrix":
        """:610-665 — re-grid via MTUtils.splitMethod(ranges, newLen)."""
        if self._by_row == new_by_row and self._by_col == new_by_col:
            return self
        nr, nc = self.num_rows(), self.num_cols()
        rl, cl = _ceil_div_d(nr, self.num_blks_by_row()), _ceil_div_d(nc, self.num_blks_by_col())
        nrl, ncl = _ceil_div_d(nr, new_by_row), _ceil_div_d(nc, new_by_col)
        new_br, new_bc = int(math.ceil(nr / nrl)), int(math.ceil(nc / ncl))
        split_col = [(cl * i, min(cl * (i + 1) - 1, nc - 1)) for i in range(self.num_blks_by_col())]
        split_row = [(rl * i, min(rl * (i + 1) - 1, nr - 1)) for i in range(self.num_blks_by_row())]
        st_col = regrid_split_method(split_col, ncl)
        st_row = regrid_split_method(split_row, nrl)
        pieces: Dict[Tuple[int, int], list] = {}
        for (r, c), blk in self.blocks:
            for (rid, (or1, or2), (nr1, nr2)) in st_row[r]:
                for (cid, (oc1, oc2), (nc1, nc2)) in st_col[c]:
                    pieces.setdefault((rid, cid), []).append((nr1, nr2, nc1, nc2, blk[or1:or2 + 1, oc1:oc2 + 1].copy()))
        res = []
        for (rid, cid), lst in pieces.items():
            row_len = nr - rid * nrl if (rid + 1) * nrl > nr else nrl
            col_len = nc - cid * ncl if (cid + 1) * ncl > nc else ncl
            mat = np.zeros((row_len, col_len), order="F")
            for (r1, r2, c1, c2, piece) in lst:
                mat[r1:r2 + 1, c1:c2 + 1] = piece
            res.append(((rid, cid), mat))
        return BlockMatrix(res, nr, nc, new_br, new_bc)

    def save_block_lines(self) -> List[str]:
        """saveToFileSystem(path, "blockmatrix") :550-555: `row-col-rows-cols:v,v,...` column-major."""
        return [f"{r}-{c}-{b.shape[0]}-{b.shape[1]}:" + ",".join(_jdouble(v) for v in b.reshape(-1, order="F"))
                for (r, c), b in self.blocks]


class DenseVecMatrix:
    """matrix/DenseVecMatrix.scala:41-69.  rows: list of (index, 1-D ndarray), any order."""

    def __init__(self, rows, n_rows: int = 0, n_cols: int = 0):
        self.rows = [(int(i), np.asarray(v, dtype=np.float64)) for i, v in rows]
        self._n_rows, self._n_cols = n_rows, n_cols

    def num_cols(self) -> int:
        if self._n_cols <= 0:
            if not self.rows:
                raise RuntimeError("empty collection")
            self._n_cols = self.rows[0][1].size
        return self._n_cols

    def num_rows(self) -> int:
        if self._n_rows <= 0:
            if not self.rows:
                raise RuntimeError("empty collection")
            self._n_rows = max(i for i, _ in self.rows) + 1
        return self._n_rows

    def to_breeze(self) -> np.ndarray:
        """:74-84"""
        mat = np.zeros((self.num_rows(), self.num_cols()), order="F")
        for i, v in self.rows:
            mat[i, :] = v
        return mat

    def to_block_matrix(self, num_by_row: int, num_by_col: int) -> BlockMatrix:
        """:1259-1328"""
        m_rows, m_cols = self.num_rows(), self.num_cols()
        brs, bcs = _ceil_div_d(m_rows, num_by_row), _ceil_div_d(m_cols, num_by_col)
        by_row, by_col = int(math.ceil(m_rows / brs)), int(math.ceil(m_cols / bcs))
        groups: Dict[Tuple[int, int], list] = {}
        for idx, vec in self.rows:
            for i in range(by_col):
                start = i * bcs
                end = min(start + bcs, m_cols)
                groups.setdefault((idx // brs, i), []).append((idx, vec[start:end].copy()))
        res = []
        for (br, bc), lst in groups.items():
            row_base, col_base = br * brs, bc * bcs
            sm_rows = m_rows - row_base if row_base + brs - 1 >= m_rows else brs
            sm_cols = m_cols - col_base if col_base + bcs - 1 >= m_cols else bcs
            mat = np.zeros((sm_rows, sm_cols), order="F")
            for idx, v in lst:
                mat[idx - row_base, :] = v
            res.append(((br, bc), mat))
        return BlockMatrix(res, m_rows, m_cols, by_row, by_col)

    def to_blocks(self, m: int, k: int, n: int, mode: str):
        """:1084-1223 — row->block conversion fused with the seq replication.
        Returns list of ((row, col, seq), ndarray)."""
        mode = mode.lower()
        if mode not in ("right", "left"):
            raise ValueError(f"only 'right' mode or 'left' mode is supported, you should change mode {mode}")
        if not (m > 0 and k > 0 and n > 0):
            raise ValueError(f"not supported (m, k, n): ({m}, {k}, {n})")
        m_rows, m_cols = self.num_rows(), self.num_cols()
        if mode == "right":
            blkmat = self.to_block_matrix(m, k)
            out = []
            for (r, c), mat in blkmat.blocks:
                for i in range(n):
                    out.append(((r, i, r * n * k + i * k + c), mat))      # :1115,:1151
            return out
        blkmat = self.to_block_matrix(k, n)
        out = []
        for (r, c), mat in blkmat.blocks:
            for i in range(m):
                out.append(((i, c, i * n * k + c * k + r), mat))          # :1182,:1217
        return out

    def multiply_split(self, other, split_mode: Tuple[int, int, int], gemm: str = "f2j",
                       reduce_order: str = "ascending") -> BlockMatrix:
        """multiply(other, splitMode) :109-141"""
        if self.num_cols() != other.num_rows():
            raise ValueError(f"Dimension mismatch during matrix-matrix multiplication: {self.num_cols()} vs {other.num_rows()}")
        m, k, n = split_mode
        if isinstance(other, BlockMatrix):
            # :136-139 re-grids `that` with (m, k) — the reference's quirk, reproduced
            return self.to_block_matrix(m, k).multiply(other.to_block_matrix(m, k), gemm, reduce_order)
        left = {key[2]: (key, mat) for key, mat in self.to_blocks(m, k, n, "right")}
        right = {key[2]: (key, mat) for key, mat in other.to_blocks(m, k, n, "left")}
        partial: Dict[Tuple[int, int], list] = {}
        for seq in sorted(left):
            if seq in right:
                key = left[seq][0]
                partial.setdefault((key[0], key[1]), []).append((seq, block_multiply(left[seq][1], right[seq][1], gemm)))
        res = []
        for key, lst in partial.items():
            lst.sort(key=lambda t: t[0], reverse=(reduce_order == "descending"))
            acc = lst[0][1]
            for _, blk in lst[1:]:
                acc = block_add(acc, blk)
            res.append((key, acc))
        return BlockMatrix(res, self.num_rows(), other.num_cols(), m, n)

    def multiply_dist_vector(self, v: "DistributedVector", split_mode: Tuple[int, int]) -> "DistributedVector":
        """multiply(vector: DistributedVector, splitMode) matrix/DenseVecMatrix.scala:149-154"""
        if self.num_cols() != v.length():
            raise ValueError(f"Dimension mismatch during matrix-matrix multiplication: {self.num_cols()} vs {v.length()}")
        m, k = split_mode
        return self.to_block_matrix(m, k).multiply_dist_vector(v)

    def multiply_vector(self, v: np.ndarray, split_mode: Optional[int] = None):
        """multiply(vector: BDV, splitMode) :162-165 -> DistributedVector; multiply(vector: BDV) :171-184 -> the local
        vector of row . vector products (`v.t * vec` = ddot per row)."""
        v = np.asarray(v, dtype=np.float64).reshape(-1)
        if split_mode is not None:
            return self.to_block_matrix(split_mode, 1).multiply_vector(v)
        out = np.zeros(len(self.rows))
        for i, row in self.rows:
            out[int(i)] = vector_dot(row, v)
        return out

    def multiply_local(self, B: np.ndarray, gemm: str = "f2j", partitions: int = 2) -> "DenseVecMatrix":
        """multiply(B: BDM) :1660-1680 — per partition: rowsMat(::, i) := row_i (K x rowsInPart),
        matrix = B^T.copy * rowsMat, emit column i as row i."""
        if self.num_cols() != B.shape[0]:
            raise ValueError(f"Dimension mismatch during matrix-matrix multiplication: {self.num_cols()} vs {B.shape[0]}")
        bt = _f(np.asarray(B, dtype=np.float64).T)         # B.t.copy
        out = []
        chunks = _partition(self.rows, partitions)
        for part in chunks:
            if not part:
                continue
            rows_mat = np.zeros((self.num_cols(), len(part)), order="F")
            for i, (_, v) in enumerate(part):
                rows_mat[:, i] = v
            res = block_multiply(bt, rows_mat, gemm)
            for i, (idx, _) in enumerate(part):
                out.append((idx, res[:, i].copy()))
        return DenseVecMatrix(out, 0, B.shape[1])

    def multiply_auto(self, other, cores: int, broadcast_threshold: int = 300, gemm: str = "f2j"):
        """multiply(other, cores, broadcastThreshold) :196-231"""
        if self.num_cols() != other.num_rows():
            raise ValueError(f"Dimension mismatch during matrix-matrix multiplication: {self.num_cols()} vs {other.num_rows()}")
        bsize = _jvm_int(broadcast_threshold * 1024 * 1024) // 8
        if other.num_rows() * other.num_cols() <= bsize:
            return self.multiply_local(other.to_breeze(), gemm)
        if self.num_rows() * self.num_cols() <= bsize:
            if isinstance(other, DenseVecMatrix):
                return other.multiply_local(self.to_breeze(), gemm)       # :206-207 (operand-order quirk, kept)
            raise NotImplementedError("multiplyBy (BlockMatrix.scala:309-335) is outside the hot-path table")
        if isinstance(other, DenseVecMatrix) and _squareish(self.num_rows(), self.num_cols(), other.num_cols()):
            s = int(math.floor(math.pow(3 * cores, 1.0 / 3.0)))
            return self.multiply_split(other, (s, s, s), gemm)
        return self.multiply_split(other, split_method(self.num_rows(), self.num_cols(), other.num_cols(), cores), gemm)

    # ---- factorizations (matrix/DenseVecMatrix.scala:283-764).  Breeze's brzLU / brzCholesky / brzInv / `\` are LAPACK
    #      calls (dgetrf, dpotrf, dgetrf+dgetri, dgesv); scipy / numpy call the same routines and stand in for them. ----
    @staticmethod
    def _mode_is_dist(mode: str, n: int) -> bool:
        """:284-299 — "auto": distributed above 6000 rows, else local Breeze."""
        if mode == "auto":
            return n > 6000
        if mode == "breeze":
            return False
        if mode == "dist":
            return True
        raise ValueError(f"Do not support mode {mode}.")

    @staticmethod
    def _perm_from_ipiv(ipiv) -> list:
        """:303-308 — pArray: apply LAPACK's interchanges (0-based here) to 0..n-1."""
        p = list(range(len(ipiv)))
        for i, q in enumerate(ipiv):
            p[i], p[q] = p[q], p[i]
        return p

    def lu_decompose(self, mode: str = "auto", base: int = 1000, keep_unfactored_diagonal: bool = True):
        """luDecompose :283-466 -> (BlockMatrix with L (unit lower) and U packed, permutation array).
        keep_unfactored_diagonal reproduces the reference as written: for every diagonal block except the last the
        result holds the ORIGINAL block (`scatterRdds(i)(0) = matFirst.cache()`, :355) instead of its factors."""
        import scipy.linalg as sla
        n = self.num_rows()
        if n != self.num_cols():
            raise ValueError(f"LU decompose only support square matrix: {n} v.s {self.num_cols()}")
        if not self._mode_is_dist(mode, n):
            lu, piv = sla.lu_factor(self.to_breeze())
            return BlockMatrix([((0, 0), lu)], n, n, 1, 1), self._perm_from_ipiv(piv)
        nb = int(math.ceil(n / float(base)))
        sub = int(math.ceil(n / float(nb)))
        cur = dict(self.to_block_matrix(nb, nb).blocks)
        p_array = [0] * n
        scatter = {}
        for i in range(nb):
            first = cur[(i, i)]
            mat, piv = sla.lu_factor(first)
            perm = self._perm_from_ipiv(piv)
            for j, pj in enumerate(perm):
                p_array[i * sub + j] = i * sub + pj
            if i == nb - 1:
                cur = {(i, i): mat}
                continue
            second = {k: v for k, v in cur.items() if k[0] == i and k[1] > i}
            third = {k: v for k, v in cur.items() if k[0] > i and k[1] == i}
            forth = {k: v for k, v in cur.items() if k[0] > i and k[1] > i}
            l = np.tril(mat, -1) + np.eye(mat.shape[0])
            u = np.triu(mat)
            pm = np.zeros((len(perm), len(perm)))
            for j, pj in enumerate(perm):
                pm[j, pj] = 1.0
            scatter[(i, 0)] = {(i, i): first if keep_unfactored_diagonal else mat}
            scatter[(i, 1)] = {k: np.linalg.solve(l, pm) @ v for k, v in second.items()}          # (l \ permutation) * block
            scatter[(i, 2)] = {k: v @ np.linalg.inv(u) for k, v in third.items()}                  # block * brzInv(u)
            mult = {(r, c): third[(r, i)] @ np.linalg.solve(first, second[(i, c)]) for (r, c) in forth}   # blk1 * (bdata \ blk2)
            cur = {k: forth[k] - mult[k] for k in forth}
        for part in scatter.values():
            cur.update(part)
        out = []
        for (r, c), blk in cur.items():
            if r > c:                                       # :440-455 — rows of a block below the diagonal follow its block row's pivots
                arr = p_array[sub * r: (n if r == nb - 1 else sub * r + sub)]
                pm = np.zeros((len(arr), len(arr)))
                for j, a in enumerate(arr):
                    pm[j, a - sub * r] = 1.0
                blk = pm @ blk
            out.append(((r, c), blk))
        return BlockMatrix(out, n, n, nb, nb), p_array

    def cholesky_decompose(self, mode: str = "auto", base: int = 1000) -> BlockMatrix:
        """choleskyDecompose :475-561 -> lower-triangular blocks of L with L L^T = this (blocks above the diagonal are
        absent from the result, as in the reference).  The reference builds the result with `new BlockMatrix(blkMat)`,
        whose lazily derived numCols / numBlksByCol would then see one block in block-row 0; dims are given here."""
        n = self.num_rows()
        if n != self.num_cols():
            raise ValueError(f"LU decompose only support square matrix: {n} v.s {self.num_cols()}")
        if not self._mode_is_dist(mode, n):
            return BlockMatrix([((0, 0), np.linalg.cholesky(self.to_breeze()))], n, n, 1, 1)
        nb = int(math.ceil(n / float(base)))
        cur = dict(self.to_block_matrix(nb, nb).blocks)
        scatter = {}
        for i in range(nb):
            if i == nb - 1:
                cur = {k: np.linalg.cholesky(v) for k, v in cur.items()}
                continue
            third = {k: v for k, v in cur.items() if k[0] > i and k[1] == i}
            forth = {k: v for k, v in cur.items() if k[1] > i and k[0] >= k[1]}
            mat = cur[(i, i)].copy()
            for j in range(mat.shape[0]):                   # :521-523 — the lower triangle is overwritten from the upper one
                for kk in range(j):
                    mat[j, kk] = mat[kk, j]
            l = np.linalg.cholesky(mat)
            lt_inv = np.linalg.inv(l.T)
            scatter[(i, 0)] = {(i, i): l}
            scatter[(i, 1)] = {k: v @ lt_inv for k, v in third.items()}
            mult = {(r, c): scatter[(i, 1)][(r, i)] @ scatter[(i, 1)][(c, i)].T for (r, c) in forth}
            cur = {k: forth[k] - mult[k] for k in forth}
        for part in scatter.values():
            cur.update(part)
        return BlockMatrix(list(cur.items()), n, n, nb, nb)

    def inverse(self, mode: str = "auto", base: int = 1000) -> BlockMatrix:
        """inverse :568-764 — local: brzInv; distributed: block elimination (inverse of the diagonal block, scaled row and
        column panels, Schur complement), then the back substitution that assembles the inverse from the last block up."""
        n = self.num_rows()
        if n != self.num_cols():
            raise ValueError(f"Inversion only support square matrix: {n} v.s {self.num_cols()}")
        if not self._mode_is_dist(mode, n):
            return BlockMatrix([((0, 0), np.linalg.inv(self.to_breeze()))], n, n, 1, 1)
        nb = int(math.ceil(n / float(base)))
        cur = dict(self.to_block_matrix(nb, nb).blocks)
        sc = {}
        for i in range(nb):
            if i == nb - 1:
                cur = {k: np.linalg.inv(v) for k, v in cur.items()}
                continue
            second = {k: v for k, v in cur.items() if k[0] == i and k[1] > i}
            third = {k: v for k, v in cur.items() if k[0] > i and k[1] == i}
            forth = {k: v for k, v in cur.items() if k[0] > i and k[1] > i}
            inv = np.linalg.inv(cur[(i, i)])
            sc[(i, 0)] = {(i, i): inv}
            sc[(i, 1)] = {k: -inv @ v for k, v in second.items()}
            sc[(i, 2)] = {k: -v @ inv for k, v in third.items()}
            mult = {(r, c): (third[(r, i)] @ inv) @ second[(i, c)] for (r, c) in forth}
            cur = {k: forth[k] - mult[k] for k in forth}
        for i in range(nb - 2, -1, -1):
            second_mat, third_mat = sc[(i, 1)], sc[(i, 2)]
            rng_ = range(i + 1, nb)
            mult_third = {(r, i): sum(cur[(r, c)] @ third_mat[(c, i)] for c in rng_) for r in rng_}        # :690-716
            mult_second = {(i, c): sum(second_mat[(i, r)] @ cur[(r, c)] for r in rng_) for c in rng_}       # :718-742
            first = sum(second_mat[(i, c)] @ mult_third[(c, i)] for c in rng_) + sc[(i, 0)][(i, i)]         # :744-754
            cur.update(mult_second)
            cur.update(mult_third)
            cur[(i, i)] = first
        return BlockMatrix(list(cur.items()), n, n, nb, nb)

    def add(self, other, subtract: bool = False) -> "DenseVecMatrix":
        """add :771-788 / subtract :795-810 — rows.join(that.rows), v1 + v2"""
        if isinstance(other, BlockMatrix):
            other = other.to_dense_vec_matrix()
        if self.num_rows() != other.num_rows() or self.num_cols() != other.num_cols():
            raise ValueError("Dimension mismatch")
        theirs = dict(other.rows)
        res = [(i, (v - theirs[i]) if subtract else (v + theirs[i])) for i, v in self.rows if i in theirs]
        return DenseVecMatrix(res, self.num_rows(), self.num_cols())

    def scalar(self, op: str, b: float) -> "DenseVecMatrix":
        """:817-871"""
        f = {"add": lambda x: x + b, "subtract": lambda x: x - b, "multiply": lambda x: x * b, "divide": lambda x: x / b,
             "subtractBy": lambda x: b - x, "divideBy": lambda x: b / x}[op]
        return DenseVecMatrix([(i, f(v)) for i, v in self.rows], self.num_rows(), self.num_cols())

    def dot_product(self, other) -> "DenseVecMatrix":
        if isinstance(other, BlockMatrix):
            other = other.to_dense_vec_matrix()
        theirs = dict(other.rows)
        return DenseVecMatrix([(i, v * theirs[i]) for i, v in self.rows if i in theirs], self.num_rows(), self.num_cols())

    def sum(self) -> float:
        total = None
        for _, v in self.rows:
            s = 0.0
            for x in v:
                s += float(x)
            total = s if total is None else total + s
        if total is None:
            raise RuntimeError("empty collection")
        return total

    def transpose(self, num_blocks: int = 2) -> BlockMatrix:
        """:1420-1436 — toBlockMatrix(min(parallelism, rows/2), 1).transpose(); local[2] => 2"""
        return self.to_block_matrix(min(num_blocks, self.num_rows() // 2), 1).transpose()

    def save_lines(self) -> List[str]:
        """saveToFileSystem :1042-1046: `t._1 + ":" + t._2.toString`, and Breeze's DenseVector.toString is
        `DenseVector(v0, v1, ...)` (valuesIterator.mkString("DenseVector(", ", ", ")"))."""
        return [f"{i}:DenseVector(" + ", ".join(_jdouble(x) for x in v) + ")" for i, v in self.rows]

    def description(self) -> str:
        """saveWithDescription :1055-1064, content of `_description`."""
        return f"MatrixName\tN/A\nMatrixSize\t{self.num_rows()} {self.num_cols()}"


class DistributedVector:
    """matrix/DistributedVector.scala:16-186.  vectors: list of (id, 1-D ndarray)."""

    def __init__(self, vectors, length: int = 0, splits: int = 0):
        self.vectors = [(int(i), np.array(v, dtype=np.float64).reshape(-1)) for i, v in vectors]
        self._len, self._splits = int(length), int(splits)
        self.column_major = True

    def split_num(self) -> int:                     # :31-36
        if self._splits <= 0:
            self._splits = len(self.vectors)
        return self._splits

    def length(self) -> int:                        # :38-43
        if self._len <= 0:
            self._len = sum(v.shape[0] for _, v in self.vectors)
        return self._len

    def substract(self, other: "DistributedVector") -> "DistributedVector":          # :45-49 (sic)
        if self.length() != other.length():
            raise ValueError(f"unsupported vector length: {self.length()} v.s {other.length()}")
        theirs = dict(other.vectors)
        return DistributedVector([(i, v - theirs[i]) for i, v in self.vectors if i in theirs], other.length(), self.split_num())

    def transpose(self) -> "DistributedVector":      # :56-60
        out = DistributedVector(self.vectors, self.length(), self.split_num())
        out.column_major = False
        return out

    def to_breeze(self) -> np.ndarray:               # :65-73 — piece id starts at id * (length / pieces), integer division
        out = np.zeros(self.length())
        offset = self.length() // len(self.vectors)
        for i, v in self.vectors:
            if i * offset + v.shape[0] > out.shape[0]:
                raise IndexError("slice out of bounds")
            out[i * offset:i * offset + v.shape[0]] = v
        return out

    def to_dis_vector(self, split_status_by_row, split_num: int) -> "DistributedVector":     # :84-107
        most = _ceil_div_d(self.length(), split_num)
        groups: Dict[int, list] = {}
        for pid, (_, vec) in enumerate(self.vectors):         # one vector per partition (iter.next())
            for vec_id, (old_start, old_end), (new_start, new_end) in split_status_by_row[pid]:
                groups.setdefault(vec_id, []).append((new_start, new_end, vec[old_start:old_end + 1]))
        out = []
        for vec_id, parts in sorted(groups.items()):
            vlen = self.length() - vec_id * most if (vec_id + 1) * most > self.length() else most
            v = np.zeros(vlen)
            for s0, s1, piece in parts:
                v[s0:s1 + 1] = piece
            out.append((vec_id, v))
        return DistributedVector(out)

    def multiply(self, other: "DistributedVector", mode: str = "dist"):                      # :146-180
        """column x row -> BlockMatrix of outer products; row x column -> Double."""
        if self.length() != other.length():
            raise ValueError("the length of these two vectors are not the same")
        if self.split_num() != other.split_num():
            raise ValueError("currently, only support two vectors with the same splits")
        if self.column_major and not other.column_major:
            theirs = dict(other.vectors)
            blocks = [((i, j), vector_outer(v, theirs[j])) for i, v in self.vectors for j in range(self.split_num()) if j in theirs]
            return BlockMatrix(blocks, self.length(), self.length(), self.split_num(), self.split_num())
        if not self.column_major and other.column_major:
            if mode.lower() == "dist":
                theirs = dict(other.vectors)
                total = None
                for i, v in sorted(self.vectors):
                    if i in theirs:
                        d = vector_dot(v, theirs[i])
                        total = d if total is None else total + d
                if total is None:
                    raise RuntimeError("empty collection")
                return total
            if mode.lower() == "local":
                return vector_dot(self.to_breeze(), other.to_breeze())
            raise ValueError("unrecognized mode")
        raise ValueError("the columnMajor status of the two distributed vectors are the same")

    @staticmethod
    def from_vector(vector: np.ndarray, num_splits: int) -> "DistributedVector":             # :184-190
        vector = np.asarray(vector, dtype=np.float64).reshape(-1)
        vlen = _ceil_div_d(vector.shape[0], num_splits)
        pieces = [(i, vector[i * vlen:min((i + 1) * vlen, vector.shape[0])]) for i in range(num_splits)]
        return DistributedVector(pieces, vector.shape[0], num_splits)


def _partition(seq: list, parts: int) -> List[list]:
    """sc.parallelize(seq, parts) slicing (Spark ParallelCollectionRDD.slice positions)."""
    n = len(seq)
    return [seq[(i * n) // parts:((i + 1) * n) // parts] for i in range(parts)]


def _jvm_int(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def _squareish(m: int, k: int, n: int) -> bool:
    """matrix/DenseVecMatrix.scala:208-211 (numRows()/numCols() is Long division)"""
    ratio = float(m * n) / float(k * k)
    q = m // k
    return 0.8 < ratio < 1.2 and q < 1.2 and q > 0.8


def _jdouble(v: float) -> str:
    """java.lang.Double.toString: shortest digits that round-trip, plain decimal for 1e-3 <= |v| < 1e7 and computerized
    scientific notation (d.dddE[-]n) outside that range, always at least one digit after the point.  (JDKs before 19
    print a few values with one digit more than the shortest, JDK-4511638; the shortest form is used here.)"""
    v = float(v)
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "Infinity" if v > 0 else "-Infinity"
    if v == 0.0:
        return "-0.0" if math.copysign(1.0, v) < 0 else "0.0"
    from decimal import Decimal
    sign, digs, exp = Decimal(repr(abs(v))).as_tuple()
    digits = "".join(str(d) for d in digs).lstrip("0")
    exp += len(digits) - len(digits.rstrip("0"))
    digits = digits.rstrip("0") or "0"
    lead = len(digits) + exp - 1                        # decimal exponent of the leading digit
    neg = "-" if v < 0 else ""
    if 1e-3 <= abs(v) < 1e7:
        if lead >= 0:
            whole = digits[:lead + 1].ljust(lead + 1, "0")
            frac = digits[lead + 1:] or "0"
        else:
            whole, frac = "0", "0" * (-lead - 1) + digits
        return f"{neg}{whole}.{frac}"
    return f"{neg}{digits[0]}.{digits[1:] or '0'}E{lead}"


# --------------------------------------------------------------------------------------------
# Loaders / generators (L3)
# --------------------------------------------------------------------------------------------
_SEP = re.compile(r",\s?|\s+")


def load_matrix_file(path: str) -> DenseVecMatrix:
    """MTUtils.loadMatrixFile (utils/MTUtils.scala:286-300): `rowIndex:v,v,...`, separators `,\\s?|\\s+`."""
    rows = []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if not line:
                continue
            head, body = line.split(":")
            rows.append((int(head), np.array([float(t) for t in _SEP.split(body) if t != ""])))
    return DenseVecMatrix(rows)


def load_block_matrix_lines(lines: Iterable[str]) -> BlockMatrix:
    """MTUtils.loadBlockMatrixFile (utils/MTUtils.scala:324-340): `r-c-rows-cols:colmajor,...`"""
    blocks = []
    for line in lines:
        line = line.strip()
        if not line:
            continue
        head, body = line.split(":")
        r, c, nr, nc = (int(t) for t in head.split("-"))
        arr = np.array([float(t) for t in _SEP.split(body) if t != ""])
        blocks.append(((r, c), arr.reshape((nr, nc), order="F")))
    return BlockMatrix(blocks)


def random_den_vec_matrix(n_rows: int, n_cols: int, num_partitions: int, seed: int, lo: float = 0.0,
                          hi: float = 1.0) -> DenseVecMatrix:
    """MTUtils.randomDenVecMatrix (utils/MTUtils.scala:63-73) -> RandomDenVecRDD (rdd/RandomRDD.scala:161-182):
    partition p holds rows [p*N/P, (p+1)*N/P) (:38-41), seeded with the p-th nextLong of Random(seed);
    each row is Array.fill(cols)(nextValue()) (:70-79)."""
    seeds = java_random_longs(seed, num_partitions)
    rows = []
    start = 0
    for p in range(num_partitions):
        end = ((p + 1) * n_rows) // num_partitions
        cnt = end - start
        vals = uniform_stream(seeds[p], 0, cnt * n_cols, lo, hi).reshape(cnt, n_cols)
        rows.extend((start + i, vals[i]) for i in range(cnt))
        start = end
    return DenseVecMatrix(rows, n_rows, n_cols)


def random_block_matrix(n_rows: int, n_cols: int, num_by_row: int, num_by_col: int, seed: int, lo: float = 0.0,
                        hi: float = 1.0) -> BlockMatrix:
    """MTUtils.randomBlockMatrix (utils/MTUtils.scala:34-50) -> RandomBlockRDD (rdd/RandomRDD.scala:184-223):
    one partition per block in row-major BlockID order, BDM.create(rows, cols, Array.fill(..)) column-major."""
    brs, bcs = _ceil_div_d(n_rows, num_by_row), _ceil_div_d(n_cols, num_by_col)
    by_row, by_col = int(math.ceil(n_rows / brs)), int(math.ceil(n_cols / bcs))
    seeds = java_random_longs(seed, by_row * by_col)
    blocks = []
    for idx in range(by_row * by_col):
        i, j = divmod(idx, by_col)
        rows = brs
        if idx >= (by_row - 1) * by_col and brs * by_row > n_rows:
            rows = n_rows - brs * (by_row - 1)
        cols = bcs
        if (idx + 1) % by_col == 0 and bcs * by_col > n_cols:
            cols = n_cols - bcs * (by_col - 1)
        vals = uniform_stream(seeds[idx], 0, rows * cols, lo, hi)
        blocks.append(((i, j), vals.reshape((rows, cols), order="F")))
    return BlockMatrix(blocks, n_rows, n_cols, by_row, by_col)
