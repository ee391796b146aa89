"""Partitioners of the multiply path.  In this engine a partition id maps onto a GPU (rank)."""
from __future__ import annotations

from ..matrix.block import BlockID


class MatrixMultPartitioner:
    """rdd/MatrixMultPartitioner.scala:6-33 — one block product per partition, partition = seq."""

    def __init__(self, mSplitNum: int, kSplitNum: int, nSplitNum: int):
        self.mSplitNum, self.kSplitNum, self.nSplitNum = mSplitNum, kSplitNum, nSplitNum

    @property
    def numPartitions(self) -> int:
        return self.mSplitNum * self.kSplitNum * self.nSplitNum

    def getPartition(self, key) -> int:
        if not isinstance(key, BlockID):
            raise ValueError(f"Unrecognized key: {key}")
        return key.seq

    def seq(self, i: int, j: int, kk: int) -> int:
        """matrix/BlockMatrix.scala:163,168"""
        return i * self.nSplitNum * self.kSplitNum + j * self.kSplitNum + kk

    def __eq__(self, other):
        return (isinstance(other, MatrixMultPartitioner) and
                (self.mSplitNum, self.kSplitNum, self.nSplitNum) == (other.mSplitNum, other.kSplitNum, other.nSplitNum))


class MatrixElemOpPartitioner:
    """rdd/MatrixElemOpPartitioner.scala:7-31 — partition = row * numBlksByCol + column."""

    def __init__(self, numBlksByRow: int, numBlksByCol: int):
        self.numBlksByRow, self.numBlksByCol = numBlksByRow, numBlksByCol

    @property
    def numPartitions(self) -> int:
        return self.numBlksByRow * self.numBlksByCol

    def getPartition(self, key) -> int:
        if not isinstance(key, BlockID):
            raise ValueError(f"Unrecognized key: {key}")
        return key.row * self.numBlksByCol + key.column

    def __eq__(self, other):
        return (isinstance(other, MatrixElemOpPartitioner) and
                (self.numBlksByRow, self.numBlksByCol) == (other.numBlksByRow, other.numBlksByCol))
