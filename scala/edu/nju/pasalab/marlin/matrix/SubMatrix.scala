package edu.nju.pasalab.marlin.matrix

import breeze.linalg.{DenseMatrix => BDM}

/**
 * Device-resident SubMatrix: the per-block value type of BlockMatrix with the public shape of the reference's class
 * (matrix/SubMatrix.scala:6-139 — constructors from a Breeze matrix, rows / cols / isSparse, add / subtract / divide /
 * multiply overloads), holding a handle to an `mb_block` in HBM instead of a `BDM[Double]` on the heap.
 * BlockMatrix and DenseVecMatrix above it are unchanged: `block1.multiply(block2)` (BlockMatrix.scala:175) is now one
 * DMMA launch on the B200, `a.add(b)` (:177) one HBM-bound kernel.  Data crosses PCIe only when a block is built from a
 * Breeze matrix or read back through `denseBlock` (toBreeze, save).  Dense blocks only: the sparse branches of the
 * reference (SparseMatrix / LibMatrixMult) are outside this library's scope and keep their JVM implementation.
 */
class SubMatrix private[marlin] (private[marlin] val handle: Long, val rows: Int, val cols: Int)
  extends Serializable {

  /** `new SubMatrix(denseMatrix = m)`: the Breeze record goes to HBM as it is — offset, majorStride and isTranspose are
   *  honoured by the upload, nothing is repacked on the heap. */
  def this(denseMatrix: BDM[Double]) = this(
    Native.upload(Ctx.get, denseMatrix.data, denseMatrix.offset, denseMatrix.rows, denseMatrix.cols,
      denseMatrix.majorStride, denseMatrix.isTranspose, Native.F64),
    denseMatrix.rows, denseMatrix.cols)

  def isSparse: Boolean = false

  private def fresh(r: Int, c: Int): Long = Native.alloc(Ctx.get, r, c, Native.F64)

  /** toBreeze / save only: reads the block back as a packed column-major Breeze matrix. */
  private[marlin] def denseBlock: BDM[Double] = {
    val out = new Array[Double](rows * cols)
    Native.download(Ctx.get, handle, out, math.max(1, rows))
    new BDM[Double](rows, cols, out)
  }

  /** Breeze `.t`: a transposed view, no copy (used by BlockMatrix.transpose before `.copy`). */
  private[marlin] def t: SubMatrix = new SubMatrix(Native.viewT(Ctx.get, handle), cols, rows)

  def add(other: SubMatrix): SubMatrix = {
    val out = fresh(rows, cols); Native.add(Ctx.get, handle, other.handle, out); new SubMatrix(out, rows, cols)
  }

  def add(b: Double): SubMatrix = {
    val out = fresh(rows, cols); Native.axpb(Ctx.get, handle, 1.0, b, out); new SubMatrix(out, rows, cols)
  }

  def subtract(other: SubMatrix): SubMatrix = {
    val out = fresh(rows, cols); Native.sub(Ctx.get, handle, other.handle, out); new SubMatrix(out, rows, cols)
  }

  def subtract(b: Double): SubMatrix = {
    val out = fresh(rows, cols); Native.axpb(Ctx.get, handle, 1.0, -b, out); new SubMatrix(out, rows, cols)
  }

  def divide(b: Double): SubMatrix = {
    val out = fresh(rows, cols); Native.div(Ctx.get, handle, b, false, out); new SubMatrix(out, rows, cols)
  }

  /** The kernel seam: was `this.denseBlock * other.denseBlock` -> Breeze -> netlib dgemm. */
  def multiply(other: SubMatrix): SubMatrix = {
    val out = fresh(rows, other.cols)
    Native.gemm(Ctx.get, handle, other.handle, out, false)
    new SubMatrix(out, rows, other.cols)
  }

  /** `multiply(other: BDM[Double])`: the broadcast matrix is uploaded once per call site and multiplied on the device. */
  def multiply(other: BDM[Double]): SubMatrix = {
    val b = new SubMatrix(other)
    try multiply(b) finally b.release()
  }

  def multiply(b: Double): SubMatrix = {
    val out = fresh(rows, cols); Native.axpb(Ctx.get, handle, b, 0.0, out); new SubMatrix(out, rows, cols)
  }

  /** Matrix x vector (was Breeze `BDM * BDV` -> dgemv): a vector is a block with one column. */
  def multiply(v: SubMatrix): SubMatrix = {
    require(v.cols == 1, s"expected a column vector, got ${v.rows} x ${v.cols}")
    val y = fresh(rows, 1); Native.gemv(Ctx.get, handle, v.handle, y, false); new SubMatrix(y, rows, 1)
  }

  /** `denseBlock.t.copy` of BlockMatrix.transpose (matrix/BlockMatrix.scala:514-523). */
  private[marlin] def transposeCopy(): SubMatrix = {
    val out = fresh(cols, rows); Native.transpose(Ctx.get, handle, out); new SubMatrix(out, cols, rows)
  }

  private[marlin] def sum(): Double = Native.sum(Ctx.get, handle)

  /** Breeze `.copy`: a packed device copy (the in-place factorizations work on one). */
  private[marlin] def copy(): SubMatrix = {
    val out = fresh(rows, cols); Native.copy(Ctx.get, handle, out); new SubMatrix(out, rows, cols)
  }

  private[marlin] def release(): Unit = Native.free(Ctx.get, handle)

  override protected def finalize(): Unit = release()
}
