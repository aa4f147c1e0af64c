"""INTEGRATION.md's C++ bindings are real code: the snippets it shows are the marked regions of
tests/cpp/integration_snippets.cpp, which compiles against the shim with stand-ins that carry the reference's own member
names (kps_features_multiscale, kps_indices_multiscale, mv_corrs_fixed_level, ...)."""
import os
import re
import subprocess
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "integration_snippets.cpp")


def test_snippets_compile_against_the_shim():
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), SRC], check=True)


def test_integration_md_shows_the_compiled_text():
    src = open(SRC).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = re.findall(r"// snippet-begin (\w+)", src)
    assert sorted(names) == ["local", "narrow", "wide"]
    for n in names:
        body = re.search(r"// snippet-begin %s\n(.*?)\n\s*// snippet-end %s" % (n, n), src, re.S).group(1)
        assert textwrap.dedent(body) in doc, n
    for ident in ("kps_features_multiscale[idx_query]", "mv_corrs_fixed_level", "kps_indices_multiscale"):
        assert ident in doc
    for stale in ("st_query.features[i]", "mv_correspondences_ij"):
        assert stale not in doc
