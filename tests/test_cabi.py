"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/bluest_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "bluest_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(blu_[a-z0-9_A-Z]+)\s*\(", txt)))


def test_header_declares_the_reference_surface():
    syms = _header_symbols()
    # Level 1: one entry point per routine of `_cmisc_bluest` (cmisc.cpp:99-110)
    for name in ("blu_assemble_psi_c", "blu_objectiveK_c", "blu_objectiveK_c_i64", "blu_cleanupK_c", "blu_gradK_c", "blu_hessKQ_c"):
        assert name in syms
    # Level 2: the SAP closures (sap.py:131-143)
    for name in ("blu_ctx_create", "blu_ctx_set_covariance", "blu_ctx_set_invcovs", "blu_get_phi", "blu_variance",
                 "blu_variance_GH", "blu_cleanup_matrix", "blu_ctx_assemble_psi", "blu_pilot_covariance"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    from bluest_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(L, name), "libbluest_b200.so does not export %s" % name
    # and the Python binding covers the whole header
    assert set(_header_symbols()) <= set(_lib.SIGNATURES)
    _lib.lib()


def test_no_torch_types_in_the_abi():
    txt = open(os.path.join(ROOT, "include", "bluest_b200.h")).read()
    assert "torch" not in txt.lower() and "at::" not in txt and "std::" not in txt
    assert 'extern "C"' in txt


def test_fails_loudly_without_gpu():
    import bluest_b200 as blu
    if blu.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(blu.BluError) as ei:
        blu.SAP(np.eye(3), 3, blu.enumerate_groups(3), np.ones(7), verbose=False)
    assert ei.value.code == 5 and "no CPU fallback" in str(ei.value)
    with pytest.raises(blu.BluError):
        blu.pilot_covariance(np.zeros((4, 2)))
    with pytest.raises(blu.BluError):
        blu.cmisc.gradK_c(np.zeros(3), 1, 3, np.arange(3), np.ones(3), np.ones(3))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "bluest_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "liboracle" not in txt and "ref_shim" not in txt, f


def test_level1_type_checks_happen_before_the_device():
    import bluest_b200 as blu
    with pytest.raises(TypeError):
        blu.cmisc.assemble_psi_c(np.zeros(9, dtype=np.float32), 3, 1, 1, np.zeros(1, dtype=np.int64), np.ones(1))
    with pytest.raises(TypeError):
        blu.cmisc.gradK_c([0.0, 0.0], 1, 2, np.zeros(2, dtype=np.int64), np.ones(2), np.ones(3))


def test_install_patches_the_reference_in_place():
    """bluest_b200.install(): the reference's SAP keeps its solvers, takes setup + closures from the
    B200 class; the Level-1 names in bluest.misc are rebound.  Needs the reference checkout."""
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference checkout not present")
    import bluest_b200 as blu
    ns = ref_shim.load()
    ref_sap = ns.sap.SAP
    ref_grad = ns.misc.gradK_c
    try:
        hybrid = blu.install()
        assert ns.sap.SAP is hybrid and ns.mosap.SAP is hybrid
        assert issubclass(hybrid, ref_sap)
        assert hybrid.__init__ is blu.SAP.__init__ and hybrid.get_variance_functions is blu.SAP.get_variance_functions
        assert hybrid.cvxopt_solve is ref_sap.cvxopt_solve and hybrid.solve is ref_sap.solve
        assert hybrid.integer_projection is blu.SAP.integer_projection and hybrid.compute_BLUE_estimator is blu.SAP.compute_BLUE_estimator
        assert ns.misc.gradK_c is blu.cmisc.gradK_c and ns.misc.hessKQ_c is blu.cmisc.hessKQ_c
        if blu.device_count() <= 0:
            with pytest.raises(blu.BluError):          # no silent CPU fallback through the patched class
                hybrid(np.eye(3), 3, blu.enumerate_groups(3), np.ones(7), verbose=False)
    finally:
        blu.uninstall()
    assert ns.sap.SAP is ref_sap and ns.mosap.SAP is ref_sap and ns.misc.gradK_c is ref_grad


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/bluest_b200.h must compile as C99 (no C++-isms) and a plain C
    program must link against the library and get the loud no-device error on a CPU box."""
    import shutil
    import subprocess
    from bluest_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdint.h>
#include "bluest_b200.h"
int main(void) {
    blu_ctx *ctx = NULL;
    int64_t sizes[2] = {2, 1};
    int64_t groups[4] = {0, 1, 0, 1};
    int rc;
    printf("%s devices=%d\n", blu_version(), blu_device_count());
    rc = blu_ctx_create(0, 2, 2, sizes, groups, &ctx);
    printf("rc=%d err=%s\n", rc, blu_last_error());
    if (rc == BLU_OK) blu_ctx_destroy(ctx);
    return (rc == BLU_ERR_NODEVICE || rc == BLU_OK) ? 0 : 1;
}
''')
    exe = tmp_path / "t"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src),
           _lib.LIB_PATH, "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert "bluest_b200" in out and "rc=" in out
