"""CPU model check of the device-side ordering of the p2p transport (transports.cu).

On the GPU every rank publishes `64 * execute + stages completed` to its peers and a one-CTA kernel in front of a
stage spins until the progress values of exchange_waits (core.h / planner.cpp) are reached.  Here the very lists the
library computes (pfftb200_describe_exchange_ordering) drive a small interpreter of ALL ranks of a mesh under random
schedules -- any rank whose waits are satisfied may run its next stage, ranks may run far ahead of each other, several
executes follow one another without any other synchronisation -- and a memory model of the receive areas checks that
  * a stage never reads a boundary before every storing rank has completed the stage that fills it,
  * a rank never stores into a receive area whose previous content its owner has not consumed yet,
  * nobody deadlocks.
(Reference counterpart: MPI_Alltoall's implicit synchronisation inside the FFTW-MPI transposes, kernel/transpose.c:224-316.)"""
import random

import pytest

import pfft_b200 as pf

T_IN, T_OUT = 1, 2

CASES = [
    dict(kind="c2c", n=[16, 16, 16], np_=[2, 1], flags=T_OUT),
    dict(kind="c2c", n=[16, 16, 16], np_=[2, 1], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[16, 16, 16], np_=[2, 2], flags=T_OUT),
    dict(kind="c2c", n=[16, 16, 16], np_=[2, 2], flags=0),                 # TRANSPOSED_NONE: return trip, 4 exchanges
    dict(kind="c2c", n=[16, 16, 16], np_=[2, 4], flags=T_OUT),
    dict(kind="c2c", n=[16, 16, 16], np_=[4, 2], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[12, 10, 9], np_=[3, 2], flags=0),
    dict(kind="c2c", n=[16, 16, 16], np_=[8], flags=T_OUT),
    dict(kind="c2c", n=[8, 8, 8, 8], np_=[2, 2, 2], flags=T_OUT),          # three exchanges: area 0 is reused inside one execute
    dict(kind="c2c", n=[8, 8, 8, 8], np_=[2, 2, 2], flags=0),
    dict(kind="c2c", n=[16, 16, 16], np_=[2, 2, 2], flags=T_OUT),          # 3-D data on a 3-D mesh: remap sub-groups
    dict(kind="r2c", n=[16, 16, 16], np_=[2, 4], flags=T_OUT),
    dict(kind="c2r", n=[16, 16, 16], np_=[2, 4], flags=T_IN, sign=+1),
    dict(kind="r2c", n=[24, 24, 24], ni=[16, 16, 16], no=[24, 24, 24], np_=[2, 2], flags=T_OUT),
]


def _id(c):
    return "%s-%s-np%s-f%d" % (c["kind"], "x".join(map(str, c["n"])), "x".join(map(str, c["np_"])), c.get("flags", 0))


class Area:
    """one receive area of one rank: which boundary it holds, who has stored into it, whether its owner consumed it"""
    def __init__(self):
        self.tag, self.filled, self.read_done = None, set(), True


def run_model(infos, executes, rng, greedy):
    P = len(infos)
    nst = infos[0]["nstages"]
    assert all(i["nstages"] == nst for i in infos)
    progress = [63] * P                       # transports.cu: every rank starts as "execute 0 complete"
    epoch = [1] * P                           # execute being run (xch_begin_kernel increments before the first stage)
    nxt = [0] * P                             # next stage
    areas = [[Area(), Area()] for _ in range(P)]
    finished = [False] * P

    def ready(r):
        if finished[r]:
            return False
        for q, back, code in infos[r]["waits"][nxt[r]]:
            if progress[q] - ((epoch[r] - back) * 64 + code) < 0:
                return False
        return True

    steps = 0
    while not all(finished):
        cand = [r for r in range(P) if ready(r)]
        assert cand, "deadlock: epochs %r next stages %r progress %r" % (epoch, nxt, progress)
        r = max(cand, key=lambda x: (epoch[x], nxt[x])) if greedy and rng.random() < 0.7 else rng.choice(cand)
        i, e, info = nxt[r], epoch[r], infos[r]
        if i > 0:                             # consume boundary i - 1
            a = areas[r][info["buffer"][i - 1]]
            assert a.tag == (e, i - 1), ("rank %d stage %d execute %d reads an area holding %r" % (r, i, e, a.tag))
            assert a.filled == set(info["writers"][i - 1]), ("rank %d stage %d execute %d: only %r of %r have stored"
                                                             % (r, i, e, sorted(a.filled), info["writers"][i - 1]))
            a.read_done = True
        if i < nst - 1:                       # store boundary i into the group's areas (own area included)
            for q in info["writers"][i]:
                a = areas[q][infos[q]["buffer"][i]]
                if a.tag == (e, i):
                    assert r not in a.filled
                    a.filled.add(r)
                else:
                    assert a.read_done, ("rank %d stage %d execute %d overwrites rank %d's area still holding unread %r"
                                         % (r, i, e, q, a.tag))
                    a.tag, a.filled, a.read_done = (e, i), {r}, False
        progress[r] = e * 64 + i + 1
        nxt[r] += 1
        if nxt[r] == nst:
            nxt[r] = 0
            epoch[r] += 1
            if epoch[r] > executes:
                finished[r] = True
        steps += 1
    return steps


@pytest.mark.parametrize("case", CASES, ids=_id)
def test_device_side_ordering_is_hazard_free(built_lib, case):
    P = 1
    for m in case["np_"]:
        P *= m
    infos = [pf.describe_exchange_ordering(pid=pid, **case) for pid in range(P)]
    assert all("error" not in i for i in infos), infos
    # the lists are symmetric: whoever stores into my area is a rank I store to
    for pid, info in enumerate(infos):
        for b, ws in enumerate(info["writers"]):
            for q in ws:
                assert pid in infos[q]["writers"][b]
    rng = random.Random(1234)
    for trial in range(40):
        run_model(infos, executes=4, rng=rng, greedy=trial % 2 == 1)


def test_the_model_catches_a_missing_wait(built_lib):
    """sanity of the checker itself: drop the waits and the memory model must object"""
    case = dict(kind="c2c", n=[16, 16, 16], np_=[2, 2], flags=T_OUT)
    infos = [pf.describe_exchange_ordering(pid=pid, **case) for pid in range(4)]
    for info in infos:
        info["waits"] = [[] for _ in info["waits"]]
    rng = random.Random(7)
    with pytest.raises(AssertionError):
        for _ in range(50):
            run_model(infos, executes=3, rng=rng, greedy=True)
