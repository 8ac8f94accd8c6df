"""The Fortran shim (d3q19-single-phase_b200/fortran/collision_b200.f90) against the C-ABI it binds.

There is no Fortran compiler in the image, and a Fortran compiler would not catch the errors that matter here anyway: a
`bind(c)` interface is taken on trust by the linker.  oracle/shim2c.py reads the shim's ISO_C_BINDING declarations and
compares them with include/d3q19_b200.h: the `d3q19_config` mirror field by field, every interface argument by argument
(count, by value / by reference, integer / real / pointer class, width), result types, mirrored constants.  The mutation
cases show that the comparison bites.  (The EXECUTABLE part of the shim is translated to C and run under the reference's
own main program in tests/test_reference_driver.py.)"""
import os
import re

import pytest

from oracle import shim2c

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_declarations_match_the_header():
    assert shim2c.lint() == []


def test_every_subroutine_of_collision_f90_that_the_driver_calls_is_defined():
    s = shim2c.Shim()
    ext = {n for n, sub in s.subs.items() if not sub["contained"]}
    assert ext == {"collision_mrt", "macrovar", "rhoupdat", "avedensity", "forcing", "forcingp"}     # collision.f90:24-602
    assert all(not s.subs[n]["dummies"] for n in ext)                                                # argument-less, as there
    # every C function the executable part calls has an interface block
    text = open(shim2c.SHIM).read().lower()
    called = set(re.findall(r"\b(d3q19_[a-z_0-9]+)\s*\(", text)) - set(s.subs) - {"d3q19_config"}
    assert called <= set(s.iface), called - set(s.iface)


MUTATIONS = [
    # (what is replaced in the shim, by what, a word the complaint must contain)
    ("integer(c_int32_t) :: nccl_max_ctas, pf_blocks,", "integer(c_int32_t) :: pf_blocks, nccl_max_ctas,", "nccl_max_ctas"),
    ("real(c_double) :: reserved_d(5)", "real(c_double) :: reserved_d(4)", "reserved_d"),
    ("        integer(c_int32_t) :: overlap\n", "", "fields"),
    ("integer(c_int32_t), value :: has_isnodes, ndiag, nflowout, nsteps, istep0, ntime, maxiter",
     "integer(c_int32_t) :: has_isnodes, ndiag, nflowout, nsteps, istep0, ntime, maxiter", "has_isnodes"),
    ("real(c_double), value :: force_in_y, force_mag", "real(c_float), value :: force_in_y, force_mag", None),
    ("function d3q19_shim_set_schedule(h, ndiag, nflowout, nsteps, istep0)", "function d3q19_shim_set_schedule(h, ndiag, nflowout, nsteps)",
     None),
    ("bind(c, name='d3q19_shim_rhoupdat')", "bind(c, name='d3q19_shim_rhoupdate')", "no such entry point"),
    ("D3Q19_SCHEME_AUTO = 2", "D3Q19_SCHEME_AUTO = 3", "D3Q19_SCHEME_AUTO"),
    ("integer(c_int32_t), value :: istep\n          real(c_double), value :: force_in_y",
     "integer(c_int64_t), value :: istep\n          real(c_double), value :: force_in_y", "istep"),
]


@pytest.mark.parametrize("k", range(len(MUTATIONS)))
def test_lint_catches_a_broken_binding(k, tmp_path):
    old, new, word = MUTATIONS[k]
    text = open(shim2c.SHIM).read()
    assert text.count(old) >= 1, old
    p = tmp_path / "collision_b200.f90"
    p.write_text(text.replace(old, new, 1))
    try:
        bad = shim2c.lint(shim2c.Shim(str(p)))
    except (SyntaxError, KeyError) as ex:                     # an interface the reader itself rejects is caught as well
        bad = [str(ex)]
    assert bad, "mutation %d went unnoticed" % k
    if word:
        assert any(word.lower() in b.lower() for b in bad), bad


def test_free_form_source_limits():
    # what a compiler checks and the translator does not: free-form lines of at most 132 characters, continuation lines that
    # end in `&`, no tab characters
    for no, line in enumerate(open(shim2c.SHIM).read().splitlines(), 1):
        assert len(line) <= 132, (no, len(line))
        assert "\t" not in line, no


def test_no_local_name_shadows_a_module_entity():
    # a local declaration of a use-associated name is an error in Fortran; needs the reference's var_inc.f90, so only where
    # it is mounted (the build box) -- shim2c.generate() refuses to translate on a clash
    ref = "/root/reference/Channel-Flow"
    if not os.path.isdir(ref):
        pytest.skip("reference sources not mounted")
    text = shim2c.ShimTranslator(ref).generate()
    assert "void ref_collision_mrt(ref_state *S)" in text and "d3q19_shim_bind_arrays(" in text
