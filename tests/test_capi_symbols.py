"""CPU-side checks of the boundary: the library builds, loads, exports every symbol the
header declares, mirrors the struct sizes, and refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    entry.build()
    return entry.load_package()


def test_every_declared_symbol_is_exported(pkg):
    hdr = open(os.path.join(ROOT, "include", "d3q19_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(d3q19_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    L = pkg.capi.load()
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(pkg.capi.SYMBOLS) == declared


def test_config_struct_layout(pkg):
    # 20 int32 + 16 double + 128 bytes, no padding surprises
    assert C.sizeof(pkg.capi.Config) == 20 * 4 + 16 * 8 + 128
    assert pkg.capi.Config.s1.offset == 80
    assert pkg.capi.Config.nccl_id.offset == 80 + 16 * 8


def test_no_cpu_fallback(pkg):
    if pkg.capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.capi.D3Q19Error, match="no CPU fallback"):
        pkg.ChannelFlow(16, 4, 4)


def test_slab_partition_matches_reference_rule(pkg):
    # para.f90:240-244 / :259-261 with nprocZ = nranks
    for nz, n in [(256, 8), (200, 3), (7, 4), (944 * 8, 8), (5, 5)]:
        parts = [pkg.slab(nz, n, r) for r in range(n)]
        assert sum(lz for lz, _ in parts) == nz
        off = 0
        for r, (lz, gz) in enumerate(parts):
            assert gz == off
            expect = (nz - nz % n) // n + (1 if r < nz - n * (nz // n) else 0)
            assert lz == expect
            off += lz


def test_varinc_matches_oracle_para(pkg, oracle):
    for laminar, nx in [(True, 64), (False, 512), (False, 199)]:
        v = pkg.VarInc(nx, 32, 32, laminar)
        p = oracle.make_para(nx, 32, 32, laminar=laminar)
        for k in ("visc", "ustar", "force_in_y", "ystar", "tau", "s1", "s2", "s4", "s9", "s10", "s13", "s16",
                  "omegepsl", "omegepslj", "omegxx"):
            assert getattr(v, k) == getattr(p, k), k


def test_host_initvel_initpop_match_oracle(pkg, oracle):
    # the host-side (run-once) initial.f90 mirror against the oracle, without touching the GPU
    nx, ny, nz = 40, 6, 5
    v = pkg.VarInc(nx, ny, nz, laminar=False)

    class Fake(pkg.ChannelFlow):
        def __init__(self):
            self.v = v
            self.lx, self.ly, self.lz, self.globalz = nx, ny, nz, 0
            shp = (nz, ny, nx)
            self.f = np.zeros(shp + (19,))
            self.rho, self.ux, self.uy, self.uz = (np.zeros(shp) for _ in range(4))

            class L:
                @staticmethod
                def d3q19_shim_sync_f_to_device(h):
                    return 0
            self.L, self.h = L, None

    s = Fake()
    for A9 in (0.0, 0.3):
        w, p = oracle.make_initial_state(nx, ny, nz, laminar=False, A9=A9, noise=False)
        s.initvel(A9)
        s.initpop()
        for k in ("ux", "uy", "uz"):
            assert np.allclose(getattr(s, k), w.get(k), rtol=1e-14, atol=1e-18), k
        assert np.allclose(s.f, w.get_f(), rtol=1e-13, atol=1e-18)


def test_header_is_plain_c(tmp_path):
    # the boundary is a C ABI: the header must compile as C99 (the Fortran shim and any C host bind to exactly this)
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "d3q19_b200.h"\nint main(void) { d3q19_config c; d3q19_particle_params p; (void)c; (void)p; return 0; }\n')
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                          str(src), "-o", str(tmp_path / "hdr.o")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
