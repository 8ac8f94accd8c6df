"""The reference's own PROGRAM main (main.f90:19-236) driving the library through the Fortran shim.

`make -C oracle ref` (where /root/reference is mounted) builds two libraries from the reference's sources,
machine-translated to C (oracle/f90toc.py): `_ref/libref.so`, all of it, and `_ref/libref_b200.so`, in which the subroutines
of collision.f90 are left out and THIS repository's fortran/collision_b200.f90 -- translated by oracle/shim2c.py -- is linked
in their place, which is the link line INTEGRATION.md gives a maintainer of the reference.  tests/refdriver_worker.py runs
`main` in both (new run: initvel, FORCING, initpop, the pre-relaxation loop with its host-side max|rho - rhop|,
saveinitflow, statistc, the time loop with diag every ndiag steps and outputflow every nflowout steps, probe) and compares everything the driver reads or writes:
with STRICT arithmetic bit for bit, with the production arithmetic within the tolerances below.

On the build box the d3q19 library is the tests' host-sim build (1, 2 and 3 ranks: NCCL id through MPI_BCAST, cudaIpc
handles through MPI_ALLGATHER, copy-engine faces); on the GPU box the same worker runs against libd3q19b200.so.  Nothing
here reads /root/reference at run time."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

from oracle import ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "refdriver_worker.py")
needs_ref = pytest.mark.skipif(not (ref.available() and ref.available(dropin=True)),
                               reason="oracle/_ref not built (needs /root/reference at build time)")
# production arithmetic (FMA contraction): populations to 1e-12 of their maximum (BASELINE.json's bound for one step holds
# after the ~30 here); rho and u are density FLUCTUATIONS and velocities of order 1e-5..1e-2, compared with their own maximum
TOL = dict(f=1e-12, unit9010=1e-12, unit26=1e-9, unit17=1e-9, unit20=1e-9, unit60=1e-9, rho=1e-9, ux=1e-9, uy=1e-9, uz=1e-9, unit27=1e-9, unit58=1e-9)


def run_worker(lib, *args, end=None):
    res = subprocess.run([sys.executable, WORKER, "--lib", lib] + [str(a) for a in args], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=900, cwd=ROOT)
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert line, res.stdout[-3000:]
    out = json.loads(line[-1])
    assert res.returncode == 0 and not out["bad"], (out["bad"], res.stdout[-2000:])
    # the time loop ran to its end (a DO variable ends one past), or left where the caller says it must
    assert out["istep_end"] == [end if end is not None else out["nsteps"] + 1] * 2
    assert {17, 20, 26, 27, 60, 9010} <= set(out["units"])    # outputuy, outputpress, diag, statistc, probe, saveinitflow wrote
    return out


def check_fast(out):
    for k, tol in TOL.items():
        if k in out["maxrel"]:
            assert out["maxrel"][k] < tol, (k, out["maxrel"][k])
    assert out["maxrel"]["f"] > 0.0                           # (it IS another arithmetic: the comparison is not vacuous)


@pytest.fixture(scope="module")
def hostsim_lib():
    spec = importlib.util.spec_from_file_location("make_hostsim", os.path.join(ROOT, "tests", "host", "make_hostsim.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.build()


@needs_ref
@pytest.mark.parametrize("ranks,scheme", [(1, "auto"), (1, "aa"), (2, "ab"), (3, "aa")])
def test_reference_main_through_the_shim_is_bit_identical_on_the_host_sim(hostsim_lib, ranks, scheme):
    run_worker(hostsim_lib, "--ranks", ranks, "--scheme", scheme, "--math", "strict")


@needs_ref
def test_reference_main_laminar_odd_sizes_on_the_host_sim(hostsim_lib):
    run_worker(hostsim_lib, "--laminar", "--size", "9x4x5", "--nsteps", 7, "--ndiag", 3, "--scheme", "aa")


@needs_ref
def test_reference_main_with_ipart_reaches_avedensity_on_the_host_sim(hostsim_lib):
    # ipart = .true. with no particle present: solid-node branches compiled in, isnodes bound, and at step 100 the
    # all-reduced avedensity of main.f90:163-167 (then one more step on the corrected field)
    run_worker(hostsim_lib, "--ipart", "--ranks", 2, "--scheme", "aa", "--nsteps", 101, "--ndiag", 50)


@needs_ref
@pytest.mark.parametrize("ranks,scheme", [(1, "auto"), (2, "aa"), (3, "ab")])
def test_reference_continued_run_through_the_shim_on_the_host_sim(hostsim_lib, ranks, scheme):
    # after the new run: savecntdflow in both builds (the records are compared), then main again with newrun = .false.
    # (main.f90:118-121): loadcntdflow fills the host f from the checkpoint BEFORE the shim's first upload, the time loop
    # continues from istep0 = 12; each build restarts from its own checkpoint
    out = run_worker(hostsim_lib, "--ranks", ranks, "--scheme", scheme, "--restart", 9)
    assert out["restart_istep_end"] == [22, 22]


@needs_ref
def test_reference_new_run_from_a_saved_prerelaxed_flow_on_the_host_sim(hostsim_lib):
    # newinitflow = .false. (main.f90:111-118): loadinitflow reads the records the first run's saveinitflow wrote, statistc
    # runs before any macrovar, then FORCING, macrovar and the loop from step 1
    out = run_worker(hostsim_lib, "--ranks", 2, "--scheme", "ab", "--restart", 9, "--from-initflow")
    assert out["restart_istep_end"] == [10, 10]


@needs_ref
def test_reference_main_leaves_on_its_wall_clock_budget_on_the_host_sim(hostsim_lib):
    # main.f90:197-207: every ntime steps the ranks all-reduce their elapsed time and leave when it exceeds time_bond; probe
    # then reads the host arrays.  MPI_WTIME ticks once per call here: with ntime = 4 the third check (istep 12) exceeds 2.5
    run_worker(hostsim_lib, "--ranks", 2, "--nsteps", 30, "--ntime", 4, "--ndiag", 7, "--nflowout", 9, "--time-bond", 2.5, end=12)


@needs_ref
def test_reference_main_with_production_arithmetic_on_the_host_sim(hostsim_lib):
    check_fast(run_worker(hostsim_lib, "--ranks", 2, "--math", "fast"))


# ---- the same on a B200, against the product library -----------------------------------------------------------------
def product_lib():
    import __graft_entry__ as entry
    return entry.load_package().capi.LIB_PATH


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("scheme", ["aa", "ab"])
def test_reference_main_through_the_shim_is_bit_identical_on_the_gpu(scheme):
    run_worker(product_lib(), "--scheme", scheme, "--math", "strict", "--size", "24x10x12", "--nsteps", 23, "--ndiag", 5)


@pytest.mark.gpu
@needs_ref
def test_reference_main_with_production_arithmetic_on_the_gpu():
    # the shim as written: SCHEME_AUTO, MATH_FAST
    check_fast(run_worker(product_lib(), "--math", "fast", "--size", "24x10x12", "--nsteps", 23, "--ndiag", 5))


@pytest.mark.gpu
@needs_ref
def test_reference_main_with_ipart_reaches_avedensity_on_the_gpu():
    run_worker(product_lib(), "--ipart", "--scheme", "ab", "--nsteps", 101, "--ndiag", 50, "--size", "24x10x12")
