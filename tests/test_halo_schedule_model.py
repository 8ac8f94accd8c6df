"""Stream/event schedules of the multi-GPU step (d3q19_api.cu step_impl) as a small model: every launch is an
operation with the set of (array, plane, population slot) addresses it reads and writes, every stream order and
every cudaStreamWaitEvent is a happens-before edge, and the test checks that ANY two operations of a rank that
touch a common address -- at least one of them writing -- are ordered.  The schedule of the NCCL and the copy-engine
transports:

  sc: B(k) I(k) B(k+1) ...   sx: X(k) after B(k);  B(k+1) after X(k)

(A second schedule, "boundary stream" -- B(k) on its own stream next to I(k) -- passed this model in round 1 and FAILED
bit-exactness on two real GPUs in round 2 (profiles/r02b_pytest_gpu_2gpu.log): the model orders launches of ONE rank and
takes a message for an edge; it is a necessary check, not a proof.  That schedule was removed.)

B = the two boundary planes, I = the interior planes, X = pack + send/recv + unpack of the faces.  The access sets
restate kernels.cuh (AB pull / AA even / AA odd) and exchange_after_step; the bit-exact behaviour of those kernels
is what tests/test_kernels_host.py checks -- this file checks that the ORDER in which the host enqueues them leaves
no race, which no single-threaded run can show.  Cross-rank ordering is the message itself (a receive completes
after the matching send), so one rank with its own arrays is the whole model.
"""
import itertools

import pytest

CZ = [0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1]
OPP = [0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15]
PZ, MZ = (5, 11, 12, 15, 16), (6, 13, 14, 17, 18)


def step_kind(scheme, k):
    """kind of step k = 1, 2, ... (an in-place run starts with the even step)"""
    return "ab" if scheme == "ab" else ("even" if k % 2 == 1 else "odd")


def arrays(scheme, k):
    """(array read, array written) by step k"""
    if scheme == "aa":
        return "A", "A"
    return ("A", "B") if k % 2 == 1 else ("B", "A")


def node_access(kind, src, dst, z):
    """addresses one node of plane z reads / writes (kernels.cuh header comment)"""
    if kind == "ab":
        return {(src, z - CZ[i], i) for i in range(19)}, {(dst, z, i) for i in range(19)}
    if kind == "even":
        a = {(src, z, i) for i in range(19)}
        return a, set(a)
    a = {(src, z - CZ[i], OPP[i]) for i in range(19)}       # odd: f_i = A[opp i][n - c_i], written back in place
    return a, set(a)


def exchange_access(kind, arr, lz):
    """exchange_after_step: what the pack reads and the unpack writes"""
    if kind == "ab":
        rd = {(arr, lz, s) for s in PZ} | {(arr, 1, s) for s in MZ}
        wr = {(arr, 0, s) for s in PZ} | {(arr, lz + 1, s) for s in MZ}
    elif kind == "even":
        rd = {(arr, lz, s) for s in MZ} | {(arr, 1, s) for s in PZ}
        wr = {(arr, 0, s) for s in MZ} | {(arr, lz + 1, s) for s in PZ}
    else:       # odd pushed across the faces into the ghosts: ghost -> the neighbour's real plane
        rd = {(arr, lz + 1, s) for s in PZ} | {(arr, 0, s) for s in MZ}
        wr = {(arr, 1, s) for s in PZ} | {(arr, lz, s) for s in MZ}
    return rd, wr


def build(scheme, lz, nsteps, schedule):
    """-> ops {name: (reads, writes)}, edges [(before, after)]"""
    ops, edges = {}, []
    for k in range(1, nsteps + 1):
        kind = step_kind(scheme, k)
        src, dst = arrays(scheme, k)
        for name, planes in (("B", (1, lz)), ("I", range(2, lz))):
            rd, wr = set(), set()
            for z in planes:
                r, w = node_access(kind, src, dst, z)
                rd |= r; wr |= w
            ops[(name, k)] = (rd, wr)
        ops[("X", k)] = exchange_access(kind, dst, lz)
    for k in range(1, nsteps + 1):
        edges.append((("B", k), ("I", k)))                          # sc: B(k), I(k)
        if k > 1:
            edges.append((("I", k - 1), ("B", k)))                  # sc: ..., I(k-1), B(k)
            edges.append((("X", k - 1), ("B", k)))                  # wait_exchange
        edges.append((("B", k), ("X", k)))                          # sx waits for evB
        if k > 1:
            edges.append((("X", k - 1), ("X", k)))                  # sx in order
    return ops, edges


def closure(ops, edges):
    after = {o: set() for o in ops}
    for a, b in edges:
        after[a].add(b)
    changed = True
    while changed:
        changed = False
        for a in ops:
            new = set()
            for b in after[a]:
                new |= after[b]
            if not new <= after[a]:
                after[a] |= new
                changed = True
    return after


def races(scheme, lz, nsteps, schedule):
    ops, edges = build(scheme, lz, nsteps, schedule)
    after = closure(ops, edges)
    bad = []
    for a, b in itertools.combinations(ops, 2):
        (ra, wa), (rb, wb) = ops[a], ops[b]
        if (wa & (rb | wb)) or (wb & ra):
            if b not in after[a] and a not in after[b]:
                bad.append((a, b, sorted((wa & (rb | wb)) | (wb & ra))[:3]))
    return bad


@pytest.mark.parametrize("scheme", ["ab", "aa"])
@pytest.mark.parametrize("lz", [3, 4, 5, 8])
def test_no_unordered_conflicts(scheme, lz):
    assert races(scheme, lz, 7, "inorder") == []


def test_the_model_sees_a_missing_wait():
    # drop the edge that makes the boundary planes of step k wait for the exchange of step k-1
    for scheme in ("ab", "aa"):
        ops, edges = build(scheme, 5, 5, "inorder")
        edges = [e for e in edges if not (e[0][0] == "X" and e[1][0] == "B")]
        after = closure(ops, edges)
        found = False
        for a, b in itertools.combinations(ops, 2):
            (ra, wa), (rb, wb) = ops[a], ops[b]
            if ((wa & (rb | wb)) or (wb & ra)) and b not in after[a] and a not in after[b]:
                found = True
        assert found, scheme


def test_boundary_and_interior_of_one_step_never_share_an_address():
    # in the in-place odd step every address belongs to exactly one node
    for scheme in ("ab", "aa"):
        for lz in (3, 4, 6):
            ops, _ = build(scheme, lz, 4, "inorder")
            for k in range(1, 5):
                (rb, wb), (ri, wi) = ops[("B", k)], ops[("I", k)]
                assert not (wb & (ri | wi)) and not (wi & rb), (scheme, lz, k)
