"""gingr_b200 -- the B200-native hot path of GiNGR (`GingrAlgorithm.update`) behind a C ABI, and its host-side mirror.

  csrc/, lib/      the CUDA library libgingr_cuda.so (sm_100a) and its build (python -m gingr_b200.build)
  _native, api     ctypes binding and the mirror of the reference's classes (configurations, state, CpdRegistration,
                   IcpRegistration, SimpleRegistrator, GingrInterface, ProbabilisticSettings, Model, Target)
  template         the reference's TemplateRegistration extension point (user closures, regression on the device)
  io, helper       wire / disk formats, state-log consumers;  comparison: mesh distances;  decimate: mesh decimation
  synthetic, rotation   seeded synthetic workloads, Euler-angle conventions

Nothing here computes on the CPU what the library computes: without the built library and a B200 every device call raises.
The CPU restatement used as test oracle lives outside the package and is imported by tests, smoke() and bench.py's CPU
baseline only."""
