"""The reference-side plugin (mp2p_icp_b200/host/mrpt_plugin.cpp) can only be BUILT where MRPT and mp2p_icp
are installed. Here it is type-checked against shape stubs (tests/stubs/) that restate, with file:line
citations, every MRPT / mp2p_icp declaration it touches — Matcher_Points_Base.h:119-128, Solver.h:80-101,
QualityEvaluator.h:42-60, Pairings.h:84-115, the DECLARE_PARAMETER_* / DEFINE_MRPT_OBJECT /
MRPT_INITIALIZER macros — and compiled to an object. A drift of the plugin away from those signatures
(an `override` that overrides nothing, a wrong member name, a record layout static_assert) fails here."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "mp2p_icp_b200", "host")
SRC = os.path.join(HOST, "mrpt_plugin.cpp")
FLAGS = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-DMP2P_B200_WITH_MRPT", f"-I{ROOT}/tests/stubs", f"-I{ROOT}/include"]


def test_plugin_compiles_against_the_shape_stubs(tmp_path):
    for extra in ([], ["-DMP2P_B200_BITFIELD_FAST_PATH"]):
        obj = tmp_path / "plugin.o"
        r = subprocess.run(FLAGS + extra + ["-c", SRC, "-o", str(obj)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        syms = subprocess.run(["nm", "-C", str(obj)], capture_output=True, text=True).stdout
        for cls in ("Matcher_Points_DistanceThreshold_B200", "Matcher_Points_InlierRatio_B200", "Matcher_Point2Plane_B200", "Matcher_Point2Line_B200",
                    "Matcher_Adaptive_B200", "Solver_Horn_B200", "Solver_GaussNewton_B200", "QualityEvaluator_PairedRatio_B200", "FilterDecimateVoxels_B200"):
            if cls == "FilterDecimateVoxels_B200" and cls not in open(SRC).read():
                continue
            assert f"mp2p_icp::{cls}::" in syms or f"mp2p_icp_filters::{cls}::" in syms, cls
        # the registration function sits at global scope like the reference's (register.cpp:43)
        assert "register_mp2p_icp_b200()" in syms and "mp2p_icp::register_mp2p_icp_b200" not in syms


def test_a_signature_drift_is_caught(tmp_path):
    """Negative control: the same source with one virtual signature changed must NOT compile."""
    txt = open(SRC).read()
    bad = txt.replace("const layer_name_t& globalName, const layer_name_t& localName,\n                           Pairings& out) const override", "const layer_name_t& globalName, Pairings& out) const override", 1)
    assert bad != txt
    f = tmp_path / "drift.cpp"
    f.write_text(bad)
    r = subprocess.run(FLAGS + ["-fsyntax-only", str(f)], capture_output=True, text=True)
    assert r.returncode != 0 and "override" in r.stderr


def test_safe_defaults_are_in_the_source():
    txt = open(SRC).read()
    assert "bool assumeUnmodifiedPairings = false;" in txt  # device-copy shortcut is opt-in
    assert "mp2p_b200_map_cached" in txt and "mp2p_b200_cloud_cached" in txt and "B200InvalidateLayer" in txt  # in-place edits of a layer
    assert "ctx_create(0," not in txt and "MCP_LOAD_OPT(params, device);" in txt  # no hard-coded device 0
