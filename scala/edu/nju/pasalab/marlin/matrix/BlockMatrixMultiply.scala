package edu.nju.pasalab.marlin.matrix

import org.apache.spark.rdd.RDD

/**
 * The body `BlockMatrix.multiply(other: BlockMatrix)` (matrix/BlockMatrix.scala:149-186) takes when the executors of a
 * job are one process per GPU of an NVSwitch box.  Instead of replicating blocks through two shuffles, joining and
 * reducing by key, every executor hands the blocks it holds to ONE collective native call; the MatrixMultPartitioner
 * mapping (seq = i*n*k + j*k + kk), the tile movement over NVLink, the block products and the k-way sum all happen
 * inside `Native.matmulBlockedDist`.  The result is the same RDD of (BlockID(i, j), SubMatrix) the reference returns,
 * one block per key, living on the executor `Native.distPlan` names as its owner.
 *
 * `Executors` is the deployment's registry: rank of this executor, world size, and the communicator created once per
 * job with a session string the driver broadcast (`Native.commInit`).
 */
object BlockMatrixMultiply {

  trait Executors extends Serializable {
    def rank: Int
    def world: Int
    def comm: Long
    /** which executor holds block (row, column) of a matrix with `blksByCol` column blocks (MatrixElemOpPartitioner order) */
    def ownerOf(row: Int, column: Int, blksByCol: Int): Int = Native.elemPartition(row, column, blksByCol) % world
  }

  private def ceilLen(total: Long, parts: Int): Array[Int] = {
    val Array(len, actual) = Native.blockLen(total, parts)
    Array.tabulate(actual)(p => math.min(len.toLong, total - p.toLong * len).toInt)
  }

  def multiply(self: BlockMatrix, other: BlockMatrix, ex: Executors): BlockMatrix = {
    require(self.numCols() == other.numRows(), s"Dimension mismatch " +
      s"during matrix-matrix multiplication: ${self.numCols()} vs ${other.numRows()}")
    require(self.numBlksByCol() == other.numBlksByRow(), "currently not supported for the two dimension of matrices")
    val (m, k, n) = (self.numBlksByRow(), self.numBlksByCol(), other.numBlksByCol())
    val rowLen = ceilLen(self.numRows(), m)
    val kLen = ceilLen(self.numCols(), k)
    val colLen = ceilLen(other.numCols(), n)
    val aOwner = Array.tabulate(m * k)(t => ex.ownerOf(t / k, t % k, k))
    val bOwner = Array.tabulate(k * n)(t => ex.ownerOf(t / n, t % n, n))
    val cOwner = Native.distPlan(m, k, n, ex.world).drop(m * k * n)

    // one task per executor: zip the two block RDDs partition-wise (both are placed by MatrixElemOpPartitioner)
    val result: RDD[(BlockID, SubMatrix)] = self.getBlocks.zipPartitions(other.getBlocks, preservesPartitioning = false) {
      (mine, theirs) =>
        val aTiles = new Array[Long](m * k)
        val bTiles = new Array[Long](k * n)
        mine.foreach { case (id, blk) => aTiles(id.row * k + id.column) = blk.handle }
        theirs.foreach { case (id, blk) => bTiles(id.row * n + id.column) = blk.handle }
        val cTiles = new Array[Long](m * n)
        val owned = (0 until m * n).filter(c => cOwner(c) == ex.rank)
        owned.foreach(c => cTiles(c) = Native.alloc(Ctx.get, rowLen(c / n), colLen(c % n), Native.F64))
        Native.matmulBlockedDist(ex.comm, aTiles, aOwner, bTiles, bOwner, m, k, n, rowLen, kLen, colLen, Native.F64, cTiles)
        owned.iterator.map(c => (BlockID(c / n, c % n), new SubMatrix(cTiles(c), rowLen(c / n), colLen(c % n))))
    }
    new BlockMatrix(result, self.numRows(), other.numCols(), m, n)
  }
}
