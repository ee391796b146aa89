"""BlockID — matrix/Block.scala:37-48."""
from __future__ import annotations

from dataclasses import dataclass


def _jint(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


@dataclass(frozen=True, eq=False)
class BlockID:
    """`case class BlockID(row: Int, column: Int, seq: Int = 0)`; seq is the target partition of a
    multiply (rdd/MatrixMultPartitioner.scala:18) and takes part in equality (Block.scala:39-43)."""
    row: int
    column: int
    seq: int = 0

    def __eq__(self, other):
        return isinstance(other, BlockID) and (self.row, self.column, self.seq) == (other.row, other.column, other.seq)

    def __hash__(self):
        return _jint(self.row * 31 + self.column + self.seq)      # Block.scala:45-47

    def hashCode(self) -> int:
        return self.__hash__()

    def __iter__(self):
        yield self.row
        yield self.column
