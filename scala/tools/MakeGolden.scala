//> using scala "3.3"
//> using dep "ch.unibas.cs.gravis::scalismo:1.0-RC1"
//> using dep "ch.unibas.cs.gravis::gingr:1.0-RC1"
//
// MakeGolden -- the hook that PINS gingr-b200's parity to the real reference.
//
// The reference ships no test that pins a number (src/test/scala/DummyTest.scala.scala:1-3) and the build image of
// gingr-b200 has no JVM, so its CPU oracle (oracle/oracle.py) is a restatement that nothing produced by the Scala code
// has ever been compared with.  This script closes that gap from the maintainer's side: it runs the reference's own
//   CpdRegistration().update / IcpRegistration().update   (api/GingrAlgorithm.scala:192-254)
// followed by the fit refresh of GingrGeneratorWrapper.propose (sampling/generators/GingrGeneratorWrapper.scala:28-39)
// on the femur example (examples/DemoCPD.scala:10-33, examples/DemoICP.scala:11-34; model as in
// DemoDatasetLoader.scala:107-123) and dumps INPUTS (model arrays, target) and per-iteration OUTPUTS as JSON.
//
//   cd <GiNGR checkout>/examples && scala-cli run /path/to/gingr-b200/scala/tools/MakeGolden.scala -- data/femur /path/to/gingr-b200/tests/golden
//
// writes tests/golden/reference_cpd_femur.json and tests/golden/reference_icp_femur.json; with those files present
//   python -m pytest tests/test_reference_golden.py            (oracle vs reference, CPU)
//   python -m pytest tests/test_reference_golden.py -m gpu     (libgingr_cuda vs reference, B200)
// compare every iteration at 1e-6 (coefficients relative to max |alpha|, vertices relative to the bounding-box diagonal,
// sigma2 relative) -- the tolerances of the north star.  Because the model arrays travel inside the file, neither
// scalismo's mesh decimation nor its pivoted Cholesky has to be reproduced on the other side.
//
// NOT COMPILED in the gingr-b200 build image (no JVM there); written against the reference sources cited above and
// scalismo 1.0-RC1's public API (PointDistributionModel.gp.{meanVector, basisMatrix, variance}, MeshIO, TriangleMesh3D).

import java.io.{File, PrintWriter}

import gingr.api.{GeneralRegistrationState, ModelFittingParameters, NoTransforms, RigidTransforms, GlobalTranformationType}
import gingr.api.registration.config.{CpdConfiguration, CpdRegistration, CpdRegistrationState, IcpConfiguration, IcpRegistration,
  IcpRegistrationState}
import gingr.simple.{GaussKernel, SimpleTriangleModels3D}
import scalismo.geometry.{Point, _3D}
import scalismo.io.MeshIO
import scalismo.mesh.TriangleMesh
import scalismo.statisticalmodel.PointDistributionModel
import scalismo.utils.Random.implicits._

object Json:
  def num(x: Double): String = if x.isNaN || x.isInfinite then "null" else java.lang.Double.toString(x)
  def arr(xs: Iterable[Double]): String = xs.iterator.map(num).mkString("[", ",", "]")
  def pts(ps: Iterable[Point[_3D]]): String = ps.iterator.map(p => s"[${num(p.x)},${num(p.y)},${num(p.z)}]").mkString("[", ",", "]")
  def mesh(m: TriangleMesh[_3D]): String =
    val tri = m.triangulation.triangles.iterator.map(t => s"[${t.ptId1.id},${t.ptId2.id},${t.ptId3.id}]").mkString("[", ",", "]")
    s"""{"points":${pts(m.pointSet.points.toIndexedSeq)},"triangles":$tri}"""
  def model(m: PointDistributionModel[_3D, TriangleMesh]): String =
    val gp = m.gp
    val B = gp.basisMatrix
    val rows = (0 until B.rows).iterator.map(i => arr((0 until B.cols).map(k => B(i, k)))).mkString("[", ",", "]")
    s"""{"reference":${mesh(m.reference)},"mean":${arr(gp.meanVector.toArray)},"variance":${arr(gp.variance.toArray)},"basis":$rows}"""
  def state(g: GeneralRegistrationState): String =
    val p = g.modelParameters
    val a = p.pose.rotation.angles
    val t = p.pose.translation
    s"""{"iteration":${g.iteration},"status":"${g.status}","sigma2":${num(g.sigma2)},"scale":${num(p.scale.s)},""" +
      s""""translation":[${num(t.x)},${num(t.y)},${num(t.z)}],"euler":[${num(a.phi)},${num(a.theta)},${num(a.psi)}],""" +
      s""""alpha":${arr(p.shape.parameters.toArray)},"fit":${pts(g.fit.pointSet.points.toIndexedSeq)}}"""

@main def MakeGolden(dataDir: String, outDir: String): Unit =
  scalismo.initialize()
  val iterations = 10
  val reference = MeshIO.readMesh(new File(dataDir, "femur.stl")).get.operations.decimate(100)
  val target    = MeshIO.readMesh(new File(dataDir, "femur_target.stl")).get.operations.decimate(100)
  // DemoDatasetLoader.femur: Gauss(scaling 50, sigma 70), relativeTolerance 0.01 (DemoDatasetLoader.scala:22, :113-114)
  val model = SimpleTriangleModels3D.create(reference, GaussKernel(50.0, 70.0), relativeTolerance = 0.01)

  def dump(name: String, algorithm: String, config: String, transform: String, states: Seq[GeneralRegistrationState]): Unit =
    val out = new PrintWriter(new File(outDir, s"reference_${name}.json"))
    out.write(s"""{"schema":"gingr-b200 reference golden v1","generator":"scala/tools/MakeGolden.scala","algorithm":"$algorithm",""" +
      s""""config":$config,"globalTransformation":"$transform","model":${Json.model(model)},"target":${Json.mesh(target)},""" +
      s""""states":${states.map(Json.state).mkString("[", ",", "]")}}""")
    out.close()
    println(s"wrote ${outDir}/reference_${name}.json (${states.length - 1} iterations)")

  // one GiNGR iteration exactly as the MH loop performs it for a deterministic run: update, then the fit refresh and the
  // iteration counter of GingrGeneratorWrapper.propose
  def refresh(g: GeneralRegistrationState): GeneralRegistrationState =
    g.updateFit(ModelFittingParameters.modelInstanceShapePoseScale(g.model, g.modelParameters)).updateIteration()

  { // ---- CPD: sigma2_0 = 1 as in DemoCPD.scala:21, w = 0, lambda = 1, RigidTransforms (exercises Procrustes + Euler round trip)
    val cfg  = CpdConfiguration(maxIterations = 100, initialSigma = Option(1.0))
    val algo = new CpdRegistration()
    var st: CpdRegistrationState = algo.initializeState(GeneralRegistrationState(model, target, RigidTransforms, None), cfg)
    val states = scala.collection.mutable.ArrayBuffer(st.general)
    for _ <- 0 until iterations do
      val next = algo.update(st, probabilistic = false)
      st = next.updateGeneral(refresh(next.general))
      states += st.general
    dump("cpd_femur", "CPD", """{"initialSigma":1.0,"w":0.0,"lambda":1.0,"maxIterations":100}""", "RigidTransforms", states.toSeq)
  }
  { // ---- ICP: sigma2 1 -> 1 as in DemoICP.scala:22, TriangularClosestPoint (the default), NoTransforms as in DemoICP.scala:24
    val cfg  = IcpConfiguration(maxIterations = 100, initialSigma = 1.0, endSigma = 1.0)
    val algo = new IcpRegistration()
    var st: IcpRegistrationState = algo.initializeState(GeneralRegistrationState(model, target, NoTransforms, None), cfg)
    val states = scala.collection.mutable.ArrayBuffer(st.general)
    for _ <- 0 until iterations do
      val next = algo.update(st, probabilistic = false)
      st = next.updateGeneral(refresh(next.general))
      states += st.general
    dump("icp_femur", "ICP", """{"initialSigma":1.0,"endSigma":1.0,"maxIterations":100,"reverse":false,"method":"TriangularClosestPoint"}""",
      "NoTransforms", states.toSeq)
  }
