from .partitioners import MatrixElemOpPartitioner, MatrixMultPartitioner

__all__ = ["MatrixElemOpPartitioner", "MatrixMultPartitioner"]
