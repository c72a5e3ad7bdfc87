"""The drop-in boundary: libgingr_cuda.so loads, exports every symbol include/gingr_cuda.h declares, the
ctypes table covers the header, and the product fails loudly (no CPU fallback) without a GPU."""
import pathlib
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = pathlib.Path(os.path.join(ROOT, "include", "gingr_cuda.h")).read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"GINGR_API\s+[\w\s\*]+?\b(gingr_\w+)\s*\(", src)


def test_header_declares_functions():
    names = header_functions()
    assert len(names) >= 25
    assert "gingr_update" in names and "gingr_cpd_estep" in names and "gingr_icp_closest" in names


def test_library_exports_every_declared_symbol():
    from gingr_b200 import _native as nat
    assert os.path.exists(nat.LIB_PATH), "libgingr_cuda.so not built: python -m gingr_b200.build"
    lib = ctypes.CDLL(nat.LIB_PATH)
    missing = [n for n in header_functions() if not hasattr(lib, n)]
    assert not missing, f"symbols declared in include/gingr_cuda.h but not exported: {missing}"


def test_ctypes_table_matches_header():
    from gingr_b200 import _native as nat
    assert sorted(nat.SIGNATURES) == sorted(header_functions())


def test_pod_layout_matches_header():
    from gingr_b200 import _native as nat
    # gingr_state: 12 doubles + 4 int32 ; gingr_config: see header
    assert ctypes.sizeof(nat.GingrState) == 12 * 8 + 4 * 4
    assert nat.GingrState.sigma2.offset == 80 and nat.GingrState.rank.offset == 108
    assert ctypes.sizeof(nat.GingrConfig) == 64
    assert nat.GingrConfig.threshold.offset == 8 and nat.GingrConfig.correspondence_method.offset == 60


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product must refuse to run, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gingr_b200 import api, _native as nat
    with pytest.raises(nat.GingrError) as e:
        api.Context(0)
    assert e.value.code == nat.GINGR_ERR_CUDA


def test_product_does_not_import_oracle():
    """Nothing under gingr_b200/ may import, link or execute oracle/."""
    bad = []
    for root, _, files in os.walk(os.path.join(ROOT, "gingr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = pathlib.Path(os.path.join(root, f)).read_text(errors="ignore")
                if re.search(r"^\s*(from|import)\s+oracle|libgingr_oracle|oracle/", txt, flags=re.M):
                    bad.append(os.path.join(root, f))
    assert not bad, bad


def test_header_is_plain_c99():
    """The boundary is a C ABI: the header must compile as strict C99 (no C++-isms), and the struct sizes the C compiler
    sees are the ones the ctypes mirrors use."""
    import shutil
    import subprocess
    import tempfile
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from gingr_b200 import _native as nat
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        with open(src, "w") as f:
            f.write('#include <stdio.h>\n#include "gingr_cuda.h"\nint main(void) { printf("%u %u %u\\n", (unsigned)sizeof(gingr_state), '
                    '(unsigned)sizeof(gingr_config), (unsigned)sizeof(gingr_mcmc_settings)); return 0; }\n')
        exe = os.path.join(d, "t")
        p = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                            src, "-o", exe], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(nat.GingrState), ctypes.sizeof(nat.GingrConfig), ctypes.sizeof(nat.GingrMcmcSettings)]
