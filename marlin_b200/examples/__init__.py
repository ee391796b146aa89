"""Drivers of the multiply path with the reference's command lines (examples/MatrixMultiply.scala, BLAS3.scala,
RMMcompare.scala).  `python -m marlin_b200.examples.MatrixMultiply 16384 16384 16384 8` on one GPU, or under
`torchrun --nproc-per-node N` with one process per GPU."""
