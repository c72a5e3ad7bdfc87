/*
 * Drop-in GingrAlgorithm whose `update` runs on a B200 through libgingr_cuda.so.
 * Mirrors registration/config/CPD.scala:117-160 (CpdRegistration) and ICP.scala:84-110 (IcpRegistration).
 * NOT COMPILED HERE (no JVM in the build image) -- see INTEGRATION.md.
 *
 * Why `update` and `generatorCombined` are the override points: `computePosterior` / `cashedPosterior` are private
 * in GingrAlgorithm (api/GingrAlgorithm.scala:68, :281), so a subclass cannot swap the posterior alone; and
 * CpdRegistrationState cannot be reused because its constructor eagerly builds the M x N matrix P
 * (registration/config/CPD.scala:54-75).  CudaRegistrationState is its sibling with the same trait surface.
 */
package gingr.api.registration.cuda

import java.lang.foreign.*
import java.lang.foreign.ValueLayout.*

import breeze.linalg.DenseVector
import gingr.api.*
import gingr.api.registration.config.{CpdConfiguration, IcpConfiguration}
import gingr.api.sampling.generators.GeneratorWrapperDeterministic
import scalismo.common.PointId
import scalismo.geometry.{_3D, EuclideanVector, Point}
import scalismo.mesh.TriangleMesh
import scalismo.sampling.{ProposalGenerator, TransitionProbability}
import scalismo.statisticalmodel.MultivariateNormalDistribution
import scalismo.transformations.*
import scalismo.utils.Random

case class CudaRegistrationState[C <: GingrConfig](general: GeneralRegistrationState, config: C)
    extends GingrRegistrationState[CudaRegistrationState[C]] {
  override def updateGeneral(update: GeneralRegistrationState): CudaRegistrationState[C] = this.copy(general = update)
}

/** Owns the native handles of one (model, target, config) triple.  Single-threaded like GingrAlgorithm (:70). */
final class CudaSession(general: GeneralRegistrationState, cfg: MemorySegment => Unit, device: Int = 0) extends AutoCloseable {
  import GingrCudaNative.*
  private val arena = Arena.ofConfined()
  private def out(): MemorySegment = arena.allocate(ADDRESS)
  private def doubles(a: Array[Double]): MemorySegment = arena.allocateFrom(JAVA_DOUBLE, a*)
  private def ints(a: Array[Int]): MemorySegment = arena.allocateFrom(JAVA_INT, a*)
  private def flat(ps: IndexedSeq[Point[_3D]]): Array[Double] = ps.iterator.flatMap(p => Iterator(p.x, p.y, p.z)).toArray

  val model  = general.model
  val M: Int = model.reference.pointSet.numberOfPoints
  val rank: Int = model.rank
  val ctx: MemorySegment = { val p = out(); check(ctxCreate.invoke(device, p).asInstanceOf[Int], MemorySegment.NULL); p.get(ADDRESS, 0) }

  private val hModel = {
    val p   = out()
    val gp  = model.gp
    val tri = model.reference.triangulation.triangles.iterator.flatMap(t => Iterator(t.ptId1.id, t.ptId2.id, t.ptId3.id)).toArray
    // Breeze DenseMatrix is column-major with leading dimension majorStride: handed over without a copy
    check(modelUpload.invoke(ctx, M, rank, doubles(flat(model.reference.pointSet.points.toIndexedSeq)),
      doubles(gp.meanVector.toArray), doubles(gp.basisMatrix.data), gp.basisMatrix.majorStride.toLong,
      doubles(gp.variance.toArray), ints(tri), tri.length / 3, p).asInstanceOf[Int], ctx)
    p.get(ADDRESS, 0)
  }
  private val hTarget = {
    val p   = out()
    val tri = general.target.triangulation.triangles.iterator.flatMap(t => Iterator(t.ptId1.id, t.ptId2.id, t.ptId3.id)).toArray
    check(targetUpload.invoke(ctx, general.target.pointSet.numberOfPoints, doubles(flat(general.target.pointSet.points.toIndexedSeq)),
      ints(tri), tri.length / 3, p).asInstanceOf[Int], ctx)
    p.get(ADDRESS, 0)
  }
  val reg: MemorySegment = {
    val c = arena.allocate(CONFIG); cfg(c)
    val p = out()
    check(regCreate.invoke(ctx, hModel, hTarget, c, p).asInstanceOf[Int], ctx)
    val r  = p.get(ADDRESS, 0)
    val lm = general.landmarkCorrespondences // GeneralRegistrationState.scala:43-62
    if (lm.nonEmpty)
      check(regLandmarks.invoke(r, lm.length, ints(lm.map(_._1.id).toArray), doubles(flat(lm.map(_._2))),
        doubles(lm.flatMap(_._3.cov.t.toArray).toArray)).asInstanceOf[Int], ctx)
    r
  }

  private val stIn  = arena.allocate(STATE)
  private val stOut = arena.allocate(STATE)
  private val aIn   = arena.allocate(JAVA_DOUBLE, rank.toLong)
  private val aOut  = arena.allocate(JAVA_DOUBLE, rank.toLong)
  private val fit   = arena.allocate(JAVA_DOUBLE, 3L * M)

  private def write(g: GeneralRegistrationState): Unit = writeTo(g, stIn, aIn)

  private def writeTo(g: GeneralRegistrationState, st: MemorySegment, alpha: MemorySegment): Unit = {
    val p = g.modelParameters
    st.set(JAVA_DOUBLE, 0, p.scale.s)
    val t = p.pose.translation; val e = p.pose.rotation.angles; val c = p.pose.rotation.center
    Seq(t.x, t.y, t.z, e.phi, e.theta, e.psi, c.x, c.y, c.z).zipWithIndex.foreach { case (v, i) => st.set(JAVA_DOUBLE, 8L * (1 + i), v) }
    st.set(JAVA_DOUBLE, 80, g.sigma2); st.set(JAVA_DOUBLE, 88, g.stepLength)
    st.set(JAVA_INT, 96, g.globalTransformation match { case SimilarityTransforms => 0; case RigidTransforms => 1; case NoTransforms => 2 })
    st.set(JAVA_INT, 100, g.iteration); st.set(JAVA_INT, 104, g.status.id); st.set(JAVA_INT, 108, rank)
    MemorySegment.copy(p.shape.parameters.toArray, 0, alpha, JAVA_DOUBLE, 0, rank)
  }

  private def read(g: GeneralRegistrationState, st: MemorySegment, alpha: MemorySegment): GeneralRegistrationState = {
    def d(i: Int) = st.get(JAVA_DOUBLE, 8L * i)
    val pts = fit.toArray(JAVA_DOUBLE).grouped(3).map(a => Point(a(0), a(1), a(2))).toIndexedSeq
    g.updateScaling(ScaleParameter(d(0)))
      .updateTranslation(EuclideanVector(d(1), d(2), d(3)))
      .updateRotation(EulerRotation(EulerAngles(d(4), d(5), d(6)), Point(d(7), d(8), d(9))))
      .updateShapeParameters(ShapeParameters(DenseVector(alpha.toArray(JAVA_DOUBLE))))
      .updateSigma2(d(10))
      .updateStatus(FittingStatuses(st.get(JAVA_INT, 104)))
      .updateFit(TriangleMesh3D(pts, g.model.reference.triangulation)) // the refreshed fit of GingrGeneratorWrapper.propose
  }

  def initialize(g: GeneralRegistrationState): GeneralRegistrationState = {
    write(g)
    check(initializeState.invoke(reg, stIn, aIn, fit).asInstanceOf[Int], ctx)
    read(g, stIn, aIn)
  }

  /** GingrAlgorithm.update (api/GingrAlgorithm.scala:192-254) on the device. */
  def update(g: GeneralRegistrationState, probabilistic: Boolean, seed: Long): GeneralRegistrationState = {
    write(g)
    check(update.invoke(reg, stIn, aIn, if (probabilistic) 1 else 0, seed, stOut, aOut, fit).asInstanceOf[Int], ctx)
    read(g, stOut, aOut)
  }

  /** GeneratorWrapperStochastic.logTransitionProbability (GeneratorWrapperStochastic.scala:42-63) on the device. */
  def logTransitionProbability(from: GeneralRegistrationState, to: GeneralRegistrationState): Double = {
    write(from)                      // stIn / aIn
    writeTo(to, stOut, aOut)         // the second state travels in the output segments
    val out = arena.allocate(JAVA_DOUBLE)
    check(logTransition.invoke(reg, stIn, aIn, stOut, aOut, out).asInstanceOf[Int], ctx)
    out.get(JAVA_DOUBLE, 0)
  }

  /** ProbabilisticSettings -> gingr_mcmc_configure; then `steps` MH steps on the device and the best sample
    * (GingrAlgorithm.run with probabilisticSettings, :115-175, without a host round trip per step). */
  def runChain(g: GeneralRegistrationState, settings: MemorySegment, steps: Int, seed: Long): GeneralRegistrationState = {
    check(mcmcConfigure.invoke(reg, settings, MemorySegment.NULL, 0, MemorySegment.NULL, 0).asInstanceOf[Int], ctx)
    write(g)
    check(initializeState.invoke(reg, stIn, aIn, fit).asInstanceOf[Int], ctx)
    check(mcmcChain.invoke(reg, steps, seed).asInstanceOf[Int], ctx)
    check(mcmcBest.invoke(reg, stOut, aOut, fit).asInstanceOf[Int], ctx)
    read(g, stOut, aOut)
  }

  override def close(): Unit = { regDestroy.invoke(reg); targetDestroy.invoke(hTarget); modelDestroy.invoke(hModel); ctxDestroy.invoke(ctx); arena.close() }
}

abstract class GingrAlgorithmCuda[C <: GingrConfig] extends GingrAlgorithm[CudaRegistrationState[C], C] {
  protected def fillConfig(c: C)(seg: MemorySegment): Unit
  private var session: Option[CudaSession] = None
  private def sessionFor(s: CudaRegistrationState[C]): CudaSession =
    session.getOrElse { val n = new CudaSession(s.general, fillConfig(s.config)); session = Some(n); n }

  // The plugin functions of the trait are never called: the device computes correspondence and uncertainty.
  override val getCorrespondence: CudaRegistrationState[C] => CorrespondencePairs = _ => CorrespondencePairs.empty()
  override val getUncertainty: (PointId, CudaRegistrationState[C]) => MultivariateNormalDistribution =
    (_, _) => throw new UnsupportedOperationException("uncertainty is evaluated on the device")

  override def initializeState(general: GeneralRegistrationState, config: C): CudaRegistrationState[C] = {
    val s = CudaRegistrationState(general, config)
    s.updateGeneral(sessionFor(s).initialize(general))
  }

  override def update(current: CudaRegistrationState[C], probabilistic: Boolean)(implicit rnd: Random): CudaRegistrationState[C] =
    current.updateGeneral(sessionFor(current).update(current.general, probabilistic, rnd.scalaRandom.nextLong()))

  // sigma2 is already updated inside gingr_update (CPD.scala:133-147 / ICP.scala:96-99)
  override def updateSigma2(current: CudaRegistrationState[C]): Double = current.general.sigma2

  override def generatorCombined(
    probabilisticSettings: Option[ProbabilisticSettings[CudaRegistrationState[C]]],
    mixing: Option[ProposalGenerator[CudaRegistrationState[C]] with TransitionProbability[CudaRegistrationState[C]]]
  )(implicit rnd: Random): ProposalGenerator[CudaRegistrationState[C]] with TransitionProbability[CudaRegistrationState[C]] =
    probabilisticSettings match {
      case Some(setting) =>
        // the informed generator of :186 without the private CPU posterior: proposal and transition density on the device
        val informed = new GingrGeneratorWrapper[CudaRegistrationState[C]] {
          override def gingrPropose(current: CudaRegistrationState[C]): CudaRegistrationState[C] = {
            val n = update(current, true)
            n.updateGeneral(n.general.updateGeneratedBy(name))
          }
          override def logTransitionProbability(from: CudaRegistrationState[C], to: CudaRegistrationState[C]): Double =
            sessionFor(from).logTransitionProbability(from.general, to.general)
        }
        val mix = mixing.getOrElse(new Generator[CudaRegistrationState[C]]().DefaultRandom())
        MixtureProposal(setting.randomMixture *: mix + (1.0 - setting.randomMixture) *: informed)
      case _ => GeneratorWrapperDeterministic(update, name)
    }
}

class CpdRegistrationCuda extends GingrAlgorithmCuda[CpdConfiguration] {
  def name = "CPD-CUDA"
  import GingrCudaNative.CONFIG
  protected def fillConfig(c: CpdConfiguration)(s: MemorySegment): Unit = {
    s.set(JAVA_INT, 0, 0); s.set(JAVA_INT, 4, c.maxIterations); s.set(JAVA_DOUBLE, 8, c.threshold)
    s.set(JAVA_INT, 16, if (c.useLandmarkCorrespondence) 1 else 0); s.set(JAVA_INT, 20, if (c.initialSigma.isDefined) 1 else 0)
    s.set(JAVA_DOUBLE, 24, c.initialSigma.getOrElse(0.0)); s.set(JAVA_DOUBLE, 32, c.w); s.set(JAVA_DOUBLE, 40, c.lambda)
  }
}

class IcpRegistrationCuda extends GingrAlgorithmCuda[IcpConfiguration] {
  def name = "ICP-CUDA"
  protected def fillConfig(c: IcpConfiguration)(s: MemorySegment): Unit = {
    s.set(JAVA_INT, 0, 1); s.set(JAVA_INT, 4, c.maxIterations); s.set(JAVA_DOUBLE, 8, c.threshold)
    s.set(JAVA_INT, 16, if (c.useLandmarkCorrespondence) 1 else 0); s.set(JAVA_INT, 20, 1)
    s.set(JAVA_DOUBLE, 24, c.initialSigma); s.set(JAVA_DOUBLE, 48, c.endSigma)
    s.set(JAVA_INT, 56, if (c.reverseCorrespondenceDirection) 1 else 0); s.set(JAVA_INT, 60, c.correspondenceMethod.id)
  }
}
