package edu.nju.pasalab.marlin.utils

import edu.nju.pasalab.marlin.matrix.Native

/**
 * Driver-side integer logic served by the library, for callers that want the very same decisions the native engine
 * makes: `MTUtils.splitMethod(m, k, n, cores)` (utils/MTUtils.scala:150-175), the strategy chooser of
 * `DenseVecMatrix.multiply(other, cores, broadcastThreshold)` (matrix/DenseVecMatrix.scala:196-231) and
 * `MTUtils.hashSeed` (utils/MTUtils.scala:18-21).  tests/test_cabi_host.py pins these against the oracle's line-by-line
 * restatement of the Scala.
 */
object NativeSplit {
  def splitMethod(m: Long, k: Long, n: Long, cores: Int): (Int, Int, Int) = {
    val Array(a, b, c) = Native.chooseSplit(m, k, n, cores)
    (a, b, c)
  }

  sealed trait Strategy
  case object BroadcastOther extends Strategy
  case object BroadcastThis extends Strategy
  case class Shuffle(split: (Int, Int, Int)) extends Strategy

  def chooseStrategy(aRows: Long, aCols: Long, bCols: Long, cores: Int, broadcastThreshold: Int = 300,
                     otherIsBlock: Boolean = false): Strategy = {
    val Array(s, m, k, n) = Native.chooseStrategy(aRows, aCols, bCols, cores, broadcastThreshold, otherIsBlock)
    s match {
      case 0 => BroadcastOther
      case 1 => BroadcastThis
      case _ => Shuffle((m, k, n))
    }
  }

  def hashSeed(seed: Long): Long = Native.hashSeed(seed)
}
