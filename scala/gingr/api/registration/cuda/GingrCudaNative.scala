/*
 * Panama FFM (JDK 22+) binding of libgingr_cuda.so -- the C ABI declared in include/gingr_cuda.h.
 * NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no JVM / sbt / scalismo.  It is the reference-side
 * stub a GiNGR maintainer adds (see INTEGRATION.md).  Lives under package gingr.api so that it may call the
 * private[api] updaters of GeneralRegistrationState (api/GeneralRegistrationState.scala:75-102).
 */
package gingr.api.registration.cuda

import java.lang.foreign.*
import java.lang.foreign.ValueLayout.*
import java.lang.invoke.MethodHandle

object GingrCudaNative {
  private val linker = Linker.nativeLinker()
  private val arena  = Arena.global()
  private val lib    = SymbolLookup.libraryLookup(sys.props.getOrElse("gingr.cuda.lib", "libgingr_cuda.so"), arena)

  private def fn(name: String, ret: MemoryLayout, args: MemoryLayout*): MethodHandle =
    linker.downcallHandle(lib.find(name).orElseThrow(), FunctionDescriptor.of(ret, args*))

  /** struct gingr_state (include/gingr_cuda.h): 12 doubles then 4 int32 = 112 bytes. */
  val STATE: StructLayout = MemoryLayout.structLayout(
    JAVA_DOUBLE.withName("scale"),
    MemoryLayout.sequenceLayout(3, JAVA_DOUBLE).withName("translation"),
    MemoryLayout.sequenceLayout(3, JAVA_DOUBLE).withName("euler"),
    MemoryLayout.sequenceLayout(3, JAVA_DOUBLE).withName("center"),
    JAVA_DOUBLE.withName("sigma2"),
    JAVA_DOUBLE.withName("step_length"),
    JAVA_INT.withName("global_transformation"),
    JAVA_INT.withName("iteration"),
    JAVA_INT.withName("status"),
    JAVA_INT.withName("rank")
  )

  /** struct gingr_config: int32 x2, double, int32 x2, double x4, int32 x2 (natural alignment, 64 bytes). */
  val CONFIG: StructLayout = MemoryLayout.structLayout(
    JAVA_INT.withName("algorithm"),
    JAVA_INT.withName("max_iterations"),
    JAVA_DOUBLE.withName("threshold"),
    JAVA_INT.withName("use_landmark_correspondence"),
    JAVA_INT.withName("has_initial_sigma"),
    JAVA_DOUBLE.withName("initial_sigma"),
    JAVA_DOUBLE.withName("w"),
    JAVA_DOUBLE.withName("lambda"),
    JAVA_DOUBLE.withName("end_sigma"),
    JAVA_INT.withName("reverse_correspondence_direction"),
    JAVA_INT.withName("correspondence_method")
  )

  val ctxCreate      = fn("gingr_ctx_create", JAVA_INT, JAVA_INT, ADDRESS)
  val ctxDestroy     = fn("gingr_ctx_destroy", JAVA_INT, ADDRESS)
  val lastError      = fn("gingr_last_error", ADDRESS, ADDRESS)
  val modelUpload    = fn("gingr_model_upload", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG,
                          ADDRESS, ADDRESS, JAVA_INT, ADDRESS)
  val modelDestroy   = fn("gingr_model_destroy", JAVA_INT, ADDRESS)
  val targetUpload   = fn("gingr_target_upload", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS)
  val targetDestroy  = fn("gingr_target_destroy", JAVA_INT, ADDRESS)
  val regCreate      = fn("gingr_registration_create", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS)
  val regDestroy     = fn("gingr_registration_destroy", JAVA_INT, ADDRESS)
  val regLandmarks   = fn("gingr_registration_set_landmarks", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS)
  val initializeState = fn("gingr_initialize_state", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS)
  val update         = fn("gingr_update", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS)
  val cpdEstep       = fn("gingr_cpd_estep", JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_DOUBLE, JAVA_DOUBLE, ADDRESS,
                          ADDRESS, ADDRESS)
  val icpClosest     = fn("gingr_icp_closest", JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT,
                          ADDRESS, ADDRESS, ADDRESS, ADDRESS)

  // probabilistic registration (SURVEY.md 8f-1) and multi-resolution hand-over (8f-2)
  /** struct gingr_mcmc_settings: double x2, int32 x2, double x9 = 96 bytes. */
  val MCMC_SETTINGS: StructLayout = MemoryLayout.structLayout(
    JAVA_DOUBLE.withName("random_mixture"),
    JAVA_DOUBLE.withName("uncertainty"),
    JAVA_INT.withName("evaluation_mode"),
    JAVA_INT.withName("reserved"),
    MemoryLayout.sequenceLayout(3, JAVA_DOUBLE).withName("rot_sdev"),
    MemoryLayout.sequenceLayout(3, JAVA_DOUBLE).withName("trans_sdev"),
    MemoryLayout.sequenceLayout(3, JAVA_DOUBLE).withName("shape_sdev")
  )
  val mcmcConfigure  = fn("gingr_mcmc_configure", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT)
  val evaluateLogValue = fn("gingr_evaluate_log_value", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS)
  val logTransition  = fn("gingr_log_transition_probability", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS)
  val mcmcChain      = fn("gingr_mcmc_chain", JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG)
  val mcmcBatch      = fn("gingr_mcmc_batch", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_LONG)
  val mcmcStats      = fn("gingr_mcmc_stats", JAVA_INT, ADDRESS, ADDRESS, ADDRESS)
  val mcmcBest       = fn("gingr_mcmc_best", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS)
  val stateDownload  = fn("gingr_state_download", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS)
  val modelNewReference = fn("gingr_model_new_reference", JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS)
  // GPMM construction on the device (api/gpmm/GPMMHelper.scala:99-129) and the way back into scalismo's storage
  val gpmmGaussianMixture = fn("gingr_gpmm_gaussian_mixture", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT,
                               ADDRESS, ADDRESS, JAVA_DOUBLE, JAVA_INT, ADDRESS, ADDRESS)
  val modelDownload  = fn("gingr_model_download", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS)

  def check(code: Int, ctx: MemorySegment): Int = {
    if (code < 0) {
      val msg = lastError.invoke(ctx).asInstanceOf[MemorySegment].reinterpret(4096).getString(0)
      throw new RuntimeException(s"libgingr_cuda error $code: $msg")
    }
    code
  }
}
