"""Checkpoint / restart through the GPU (saveload.f90:196-231 savecntdflow, :296-332 loadcntdflow) and the coherence of the
host and device copies of `f` when raw C-ABI calls and the driver-facing shim calls are mixed.

* run 10 steps -> savecntdflow -> a NEW handle (also in the other storage scheme) -> loadcntdflow -> run 10 more steps must
  equal 20 uninterrupted steps BIT FOR BIT, in production arithmetic too (the step is deterministic);
* bench.py's sequence initpop(); upload_f(); run_device(n) followed by sync_f_to_host() / savecntdflow must see the
  stepped populations, and a following collision_MRT must continue from the device state (the raw entry
  points used to leave the shim's flags stale, so a checkpoint written after them held the INITIAL field).
The 2-slab restart runs in tests/mgpu_worker.py (needs 2 GPUs) and on the host-sim worker."""
import numpy as np
import pytest

import __graft_entry__ as entry

pytestmark = pytest.mark.gpu

pkg = entry.load_package()
capi = pkg.capi
saveload = pkg.saveload
NX, NY, NZ = 40, 12, 10
U = dict(ustar=0.0025, ystar=0.0036 / 0.0025, force_in_y=2.0 * 0.0025 * 0.0025 / NX)
SCHEMES = {"aa": capi.SCHEME_AA, "ab": capi.SCHEME_AB}


def fresh(scheme, math_mode):
    sim = pkg.ChannelFlow(NX, NY, NZ, laminar=False, scheme=SCHEMES[scheme], math_mode=math_mode, **U)
    sim.initvel(A9=0.3)
    sim.add_hash_noise(1e-3 * sim.v.ustar, seed=4711)
    sim.FORCING()
    sim.initpop()
    return sim


@pytest.mark.parametrize("math_mode", [capi.MATH_STRICT, capi.MATH_FAST])
@pytest.mark.parametrize("first,second", [("aa", "aa"), ("ab", "ab"), ("aa", "ab"), ("ab", "aa")])
def test_restart_equals_the_uninterrupted_run(tmp_path, first, second, math_mode):
    whole = fresh(first, math_mode)
    whole.macrovar()
    whole.run(20)
    want = whole.sync_f_to_host().copy()
    whole.close()

    a = fresh(first, math_mode)
    a.macrovar()
    a.run(9)                                   # an odd count: the in-place scheme is checkpointed in its swapped phase
    saveload.savecntdflow(a, str(tmp_path))
    a.close()

    b = pkg.ChannelFlow(NX, NY, NZ, laminar=False, scheme=SCHEMES[second], math_mode=math_mode, **U)
    b.FORCING()
    istep0, _, _ = saveload.loadcntdflow(b, str(tmp_path), 9)
    assert istep0 == 9 and b.v.istep0 == 9
    b.macrovar()
    assert b.run(11) == 20                     # the loop counts on from istep0 (main.f90:142)
    got = b.sync_f_to_host()
    assert np.array_equal(got, want)
    b.close()


@pytest.mark.parametrize("scheme", ["aa", "ab"])
def test_raw_and_shim_calls_keep_host_and_device_coherent(oracle, tmp_path, scheme):
    w, p = oracle.make_initial_state(NX, NY, NZ, laminar=False, noise=True, **U)
    sim = pkg.ChannelFlow(NX, NY, NZ, laminar=False, scheme=SCHEMES[scheme], math_mode=capi.MATH_STRICT, **U)
    sim.f[...] = w.get_f()
    sim.host_f_changed()
    sim.FORCING()
    sim.upload_f()                             # raw
    sim.run_device(5)                          # raw: the device is 5 steps ahead of the host copy
    w.macrovar()
    for _ in range(5):
        w.collision_MRT(); w.macrovar()
    assert np.array_equal(sim.sync_f_to_host(), w.get_f())          # shim: must download, not trust the stale host f
    sim.v.nsteps, sim.v.istep0 = 5, 0
    path = saveload.savecntdflow(sim, str(tmp_path))
    chk = pkg.ChannelFlow(NX, NY, NZ, laminar=False, scheme=SCHEMES[scheme], math_mode=capi.MATH_STRICT, **U)
    saveload.loadcntdflow(chk, str(tmp_path), 5)
    assert np.array_equal(chk.f, w.get_f()), path                   # the file holds step 5, not step 0
    chk.close()
    sim.collision_MRT()                        # shim: continues from the device state, no re-upload of an old host f
    w.collision_MRT(); w.macrovar()
    out = np.empty_like(sim.f)
    sim.download_f(out)                        # raw download into a foreign buffer: the host copy stays stale ...
    assert np.array_equal(out, w.get_f())
    sim.collide_stream()                       # ... and a raw step after it as well
    w.collision_MRT(); w.macrovar()
    assert np.array_equal(sim.sync_f_to_host(), w.get_f())
    sim.device_macrovar()                      # raw macrovar: the shim's avedensity must not recompute or reuse stale moments
    sim.avedensity()
    mean_ref, _ = w.avedensity()
    assert np.max(np.abs(sim.rho - w.get("rho"))) <= 1e-13 * np.mean(np.abs(w.get("rho")) + 1e-30) + 1e-18
    sim.close(); w.close()
