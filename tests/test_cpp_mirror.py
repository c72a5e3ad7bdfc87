"""include/gingr.hpp, the C++ host mirror of the reference's interface: compiles warning-free as C++17, links against the
built library, has the reference's defaults (compared with the Python mirror, which tests/test_api_mirror.py checks against
the Scala sources), and fails loudly without a CUDA device.  No GPU needed (on a GPU box the example simply runs)."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "gingr_b200", "lib")
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")


def _build(src, out, tmp_path):
    exe = str(tmp_path / out)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
           "-L" + LIBDIR, "-lgingr_cuda", "-Wl,-rpath," + LIBDIR]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return exe


@pytest.fixture(scope="module")
def built_library():
    if not os.path.exists(os.path.join(LIBDIR, "libgingr_cuda.so")):
        from gingr_b200 import build
        build.build()


def test_defaults_and_pod_conversion_match_the_python_mirror(built_library, tmp_path):
    src = tmp_path / "defaults.cpp"
    src.write_text(r'''
#include <cstdio>
#include "gingr.hpp"
// explicit instantiation: every member of both algorithm classes must compile, not only the ones used below
template class gingr::GingrAlgorithm<gingr::CpdConfiguration>;
template class gingr::GingrAlgorithm<gingr::IcpConfiguration>;
int main() {
  gingr::CpdConfiguration c; gingr::IcpConfiguration k; gingr::GeneralRegistrationState s;
  s.modelParameters.shape = {0.5, -1.5, 2.0}; s.modelParameters.euler = {0.1, 0.2, 0.3}; s.iteration = 7; s.sigma2 = 2.5;
  const gingr_state p = s.toPod(); const gingr_config cp = c.toPod(); const gingr_config kp = k.toPod();
  gingr::GeneralRegistrationState back; back.fromPod(p);
  std::printf("{\"cpd\": [%d, %.17g, %d, %d, %.17g, %.17g], \"icp\": [%d, %.17g, %d, %.17g, %.17g, %d, %d, %.17g],"
              " \"state\": [%.17g, %d, %.17g, %d, %d], \"pod\": [%d, %d, %.17g, %.17g, %d, %d], \"cfg\": [%d, %d, %d, %d, %.17g],"
              " \"names\": [\"%s\", \"%s\"], \"sizes\": [%zu, %zu]}\n",
              c.maxIterations, c.threshold, (int)c.useLandmarkCorrespondence, (int)c.initialSigma.has_value(), c.w, c.lambda,
              k.maxIterations, k.threshold, (int)k.useLandmarkCorrespondence, k.initialSigma, k.endSigma,
              (int)k.reverseCorrespondenceDirection, (int)k.correspondenceMethod, k.sigmaStep(),
              s.stepLength, (int)s.globalTransformation, gingr::GeneralRegistrationState().sigma2, (int)gingr::GeneralRegistrationState().status,
              gingr::GeneralRegistrationState().iteration,
              p.rank, p.iteration, p.sigma2, p.euler[2], back.iteration, (int)back.globalTransformation,
              cp.algorithm, cp.has_initial_sigma, kp.algorithm, kp.has_initial_sigma, kp.initial_sigma,
              gingr::CpdConfiguration::name, gingr::IcpConfiguration::name, sizeof(gingr_state), sizeof(gingr_config));
  gingr::ProbabilisticSettings ps; const gingr_mcmc_settings mp = ps.toPod();
  std::fprintf(stderr, "%.17g %.17g %d %.17g %.17g %.17g %.17g %zu\n", mp.random_mixture, mp.uncertainty, mp.evaluation_mode,
               mp.rot_sdev[0], mp.trans_sdev[2], mp.shape_sdev[0], mp.shape_sdev[2], sizeof(gingr_mcmc_settings));
  gingr::CpdConfiguration c2; c2.initialSigma = 4.0;
  return (c2.toPod().has_initial_sigma == 1 && c2.toPod().initial_sigma == 4.0 && !k.converged(s, s, 1.0) &&
          c.converged(s, back, 1e-10) ) ? 0 : 3;
}
''')
    exe = _build(str(src), "defaults", tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    got = json.loads(p.stdout)
    import ctypes
    from gingr_b200 import api, _native as nat
    c, k = api.CpdConfiguration(), api.IcpConfiguration()
    assert got["cpd"] == [c.maxIterations, c.threshold, int(c.useLandmarkCorrespondence), int(c.initialSigma is not None), c.w, c.lambda_]
    assert got["icp"] == [k.maxIterations, k.threshold, int(k.useLandmarkCorrespondence), k.initialSigma, k.endSigma,
                          int(k.reverseCorrespondenceDirection), k.correspondenceMethod, k.sigmaStep]
    assert got["state"] == [1.0, api.RIGID_TRANSFORMS, 1.0, api.STATUS_NONE, 0]
    assert got["pod"] == [3, 7, 2.5, 0.3, 7, api.RIGID_TRANSFORMS]
    assert got["cfg"] == [api.ALGO_CPD, 0, api.ALGO_ICP, 1, 100.0]
    assert got["names"] == ["CPD", "ICP"]
    assert got["sizes"] == [ctypes.sizeof(nat.GingrState), ctypes.sizeof(nat.GingrConfig)]
    ps = api.ProbabilisticSettings()
    vals = p.stderr.split()
    assert [float(v) for v in vals[:2]] == [ps.randomMixture, ps.uncertainty] and int(vals[2]) == ps.mode
    assert [float(v) for v in vals[3:7]] == [ps.rotationSdev[0], ps.translationSdev[2], ps.shapeSteps[0], ps.shapeSteps[2]]
    assert int(vals[7]) == ctypes.sizeof(nat.GingrMcmcSettings)


def test_cpp_demo_links_and_fails_loudly_without_a_device(built_library, tmp_path):
    exe = _build(os.path.join(ROOT, "examples", "demo_cpd.cpp"), "demo_cpd", tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if p.returncode == 2:
        assert "no CUDA device" in p.stderr and "no CPU fallback" in p.stderr
    else:                                                   # a box with a B200: the registration itself must succeed
        assert p.returncode == 0, (p.stdout, p.stderr)
        assert "status" in p.stdout
