"""CPU tests of the planner: every rank's stage list is interpreted in numpy
(tests/schedule_sim.py) and the gathered result must equal the oracle's global
transform.  This covers the N>1 host logic (chunking, exchanges, ragged and empty
blocks) without a GPU."""
import numpy as np
import pytest

import cases
import pfft_b200 as pf
import pfft_oracle as po
import schedule_sim as ss

T_IN, T_OUT, PAD = po.TRANSPOSED_IN, po.TRANSPOSED_OUT, po.PADDED_R2C
S_IN, S_OUT = po.SHIFTED_IN, po.SHIFTED_OUT


def run_virtual(case, seed=0):
    np_ = case["np"]
    P = int(np.prod(np_))
    r = len(np_)
    scheds = [pf.describe_schedule(case["kind"], case["n"], np_, pid, case.get("flags", 0), case.get("ni"),
                                   case.get("no"), case.get("howmany", 1), case.get("iblock"), case.get("oblock"),
                                   case.get("sign", -1), case.get("kinds"), case.get("skip")) for pid in range(P)]
    for s in scheds:
        assert s["error"] == "", s["error"]
    xg = cases.make_global_input(case, seed)
    user_in = [cases.local_input(case, xg, s["local_ni"], s["local_i_start"], r) for s in scheds]
    outs = ss.simulate(scheds, user_in)
    want = cases.oracle_output(case, xg)
    scale = max(1e-300, float(np.abs(want).max()))
    err = max(cases.compare_local_output(case, want, outs[pid], s["local_no"], s["local_o_start"], r)
              for pid, s in enumerate(scheds))
    return err / scale, scheds


CASES = [
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2]),                       # BASELINE config 1
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2], flags=T_OUT),
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[8, 6, 4], np=[1, 1]),
    dict(kind="c2c", n=[8, 6, 4], np=[1, 1], flags=T_OUT),
    dict(kind="c2c", n=[8, 6, 4], np=[1, 1], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[16, 12, 10], np=[4]),                          # slab
    dict(kind="c2c", n=[16, 12, 10], np=[4], flags=T_OUT),
    dict(kind="c2c", n=[16, 12, 10], np=[3], flags=T_IN),
    dict(kind="c2c", n=[13, 14, 19, 17], np=[2, 2, 2], flags=T_OUT),   # 4-D on a 3-D mesh
    dict(kind="c2c", n=[13, 14, 19, 17], np=[2, 2, 2]),
    dict(kind="c2c", n=[13, 14, 19, 17], np=[2, 2, 2], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[13, 14, 19, 17], np=[3, 2], flags=T_OUT),
    dict(kind="c2c", n=[5, 4, 3], np=[3, 2]),                          # ragged + empty ranks
    dict(kind="c2c", n=[4, 4, 4], np=[3, 3], flags=T_OUT),
    dict(kind="c2c", n=[6, 5], np=[2]),                                # 2-D on a 1-D mesh
    dict(kind="c2c", n=[8, 6, 4], np=[2, 2], howmany=3),
    dict(kind="c2c", n=[8, 6, 4], np=[2, 2], howmany=2, flags=T_OUT),
    dict(kind="c2c", n=[16, 12, 10], np=[2, 2], iblock=[9, 7], oblock=[10, 8]),
    dict(kind="c2c", n=[16, 12, 10], np=[2, 2], oblock=[7, 6], flags=T_OUT),
    dict(kind="c2c", n=[8, 6, 4], np=[2, 2], skip=[0, 1, 0]),
    dict(kind="c2c", n=[8, 6, 4], np=[2, 2], skip=[1, 0, 1], flags=T_OUT),
    dict(kind="r2c", n=[29, 27, 31], np=[2, 2]),
    dict(kind="r2c", n=[29, 27, 31], np=[2, 2], flags=T_OUT),
    dict(kind="r2c", n=[16, 12, 10], np=[2, 2], flags=T_OUT | PAD),
    dict(kind="r2c", n=[16, 12, 10], np=[2, 2], flags=T_OUT, sign=+1),
    dict(kind="c2r", n=[29, 27, 31], np=[2, 2], sign=+1),
    dict(kind="c2r", n=[29, 27, 30], np=[2, 2], flags=T_IN, sign=+1),
    dict(kind="c2r", n=[16, 12, 10], np=[2, 2], flags=T_IN | PAD, sign=+1),
    dict(kind="c2r", n=[16, 12, 10], np=[2, 2], flags=T_IN, sign=-1),
    dict(kind="r2c", n=[8, 6, 10], np=[1, 1]),
    dict(kind="c2r", n=[8, 6, 10], np=[1, 1], sign=+1),
    # pruned / oversampled (reference tests/simple_check_ousam_*.c use 16^3 -> 29x27x31)
    dict(kind="c2c", n=[12, 10, 9], ni=[6, 5, 4], no=[12, 10, 9], np=[2, 2]),
    dict(kind="c2c", n=[12, 10, 9], ni=[12, 10, 9], no=[5, 7, 3], np=[2, 2], flags=T_OUT),
    dict(kind="c2c", n=[12, 10, 9], ni=[6, 5, 4], no=[5, 7, 3], np=[2, 2], flags=T_IN, sign=+1),
    dict(kind="r2c", n=[29, 27, 31], ni=[16, 16, 16], no=[29, 27, 31], np=[2, 2], flags=T_OUT),
    dict(kind="c2r", n=[29, 27, 31], ni=[29, 27, 31], no=[16, 16, 16], np=[2, 2], flags=T_IN, sign=+1),
    # index shifts
    dict(kind="c2c", n=[8, 6, 4], np=[2, 2], flags=S_IN | S_OUT),
    dict(kind="c2c", n=[8, 6, 4], np=[2, 2], flags=S_IN | S_OUT | T_OUT),
    dict(kind="c2c", n=[8, 6, 4], np=[2, 2], flags=S_IN),
    dict(kind="c2c", n=[8, 6, 4], np=[2, 2], flags=S_OUT),
    dict(kind="c2c", n=[16, 12, 8], ni=[8, 6, 4], no=[16, 12, 8], np=[2, 2], flags=S_IN | S_OUT),
    # 3-D data on a 3-D mesh (3dto2d remap, reference tests/simple_check_*_3d_on_3d*.c)
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2, 2]),
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2, 2], flags=T_OUT),
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2, 2], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[9, 8, 7], np=[1, 1, 4]),
    dict(kind="c2c", n=[9, 8, 7], np=[2, 1, 3], flags=T_OUT),
    dict(kind="c2c", n=[5, 4, 3], np=[1, 2, 6]),
    dict(kind="r2c", n=[29, 27, 31], np=[2, 2, 2]),
    dict(kind="r2c", n=[29, 27, 31], np=[2, 2, 2], flags=T_OUT),
    dict(kind="c2r", n=[29, 27, 31], np=[2, 2, 2], sign=+1),
    dict(kind="c2r", n=[29, 27, 31], np=[2, 2, 2], flags=T_IN, sign=+1),
    dict(kind="r2r", n=[13, 11, 9], np=[2, 2, 2], kinds=[po.REDFT00, po.REDFT01, po.REDFT10]),
    dict(kind="r2r", n=[13, 11, 9], np=[2, 2, 2], kinds=[po.RODFT00, po.RODFT10, po.REDFT11], flags=T_OUT),
    dict(kind="c2c", n=[8, 8, 8], np=[1, 1, 1]),
    # chains of register-resident power-of-two stages: micro-blocked intermediate layouts, 2-D tiles
    dict(kind="c2c", n=[64, 64, 64], np=[1, 1], flags=T_OUT),
    dict(kind="c2c", n=[64, 64, 64], np=[1, 1], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[64, 64, 64], np=[2, 2], flags=T_OUT),
    dict(kind="c2c", n=[64, 64, 64], np=[2, 4], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[64, 64, 64], np=[4, 1], flags=T_OUT),
    dict(kind="c2c", n=[128, 128, 128], np=[1, 2], flags=T_OUT),
    dict(kind="c2c", n=[64, 64, 64, 64], np=[2, 1, 2], flags=T_OUT),
    dict(kind="c2c", n=[64, 64], np=[2], flags=T_OUT),
    # r2r (reference tests/simple_check_r2r*.c)
    dict(kind="r2r", n=[13, 11, 9], np=[2, 2], kinds=[po.REDFT00, po.REDFT01, po.REDFT10]),
    dict(kind="r2r", n=[13, 11, 9], np=[2, 2], kinds=[po.RODFT00, po.RODFT10, po.REDFT11], flags=T_OUT),
    dict(kind="r2r", n=[13, 11, 9], np=[2, 2], kinds=[po.RODFT01, po.RODFT11, po.REDFT00], flags=T_IN),
    dict(kind="r2r", n=[7, 6, 5, 4], np=[2, 2, 2], kinds=[po.REDFT10, po.RODFT00, po.REDFT00, po.RODFT01], flags=T_OUT),
    dict(kind="r2r", n=[12, 10, 9], ni=[6, 5, 4], no=[12, 10, 9], np=[2, 2], kinds=[po.REDFT00, po.RODFT00, po.REDFT01]),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s-np%s-f%d" % (
    c["kind"], "x".join(map(str, c["n"])), "x".join(map(str, c["np"])), c.get("flags", 0)))
def test_schedule_reproduces_oracle(built_lib, case):
    err, _ = run_virtual(case)
    assert err < 1e-12, err


def test_local_sizes_in_schedule_match_golden_values(built_lib):
    s = pf.describe_schedule("c2c", [29, 27, 31], [2, 2], 3, T_OUT)
    assert (s["local_ni"], s["local_i_start"], s["local_no"], s["local_o_start"]) == \
        ([14, 13, 31], [15, 14, 0], [29, 13, 15], [0, 14, 16])


def test_transposed_out_has_no_extra_passes(built_lib):
    """One read + one write of the array per transformed dimension: 3 stages, 2 exchanges."""
    s = pf.describe_schedule("c2c", [1024, 1024, 1024], [2, 4], 7, T_OUT)
    assert len(s["stages"]) == 3 and len(s["exchanges"]) == 2
    assert [g["dim"] for g in s["stages"]] == [2, 1, 0]
    # headline config: every chunk of the first exchange is 512 MiB of complex doubles
    assert s["exchanges"][0]["send_cnt"] == [512 * 256 * 256] * 4
    assert s["exchanges"][1]["send_cnt"] == [512 * 256 * 512] * 2


def test_power_of_two_chains_use_micro_blocked_layouts(built_lib):
    """Headline config: 8-line tiles as 4 x 2 lines; producer and consumer tiles share 512- / 256-byte blocks."""
    for flags in (T_OUT, T_IN):
        s = pf.describe_schedule("c2c", [1024, 1024, 1024], [2, 4], 5, flags)
        assert s["error"] == ""
        g0, g1, g2 = s["stages"]
        assert [g["ntile"] for g in s["stages"]] == [8, 8, 8]
        assert (g0["oblk2"], g1["iblk2"], g1["oblk2"], g2["iblk2"]) == (4, 4, 2, 2)
        assert g0["iblk2"] == 1 and g2["oblk2"] == 1 and g0["istride"] == 1 and g2["ostride"] == 1
        assert g0["tile_ooff"] == list(range(8)) and g1["tile_ioff"] == [4 * l for l in range(8)]
        assert g1["tile_ooff"] == list(range(8)) and g2["tile_ioff"] == [2 * l for l in range(8)]
        # exchanges still move exactly the reference's chunk sizes
        assert s["exchanges"][0]["send_cnt"] == ([512 * 256 * 256] * 4 if flags == T_OUT else [512 * 256 * 512] * 2)
    # ragged or non-power-of-two sizes keep the plain layouts
    assert all(g["ntile"] == 0 for g in pf.describe_schedule("c2c", [1024, 1000, 1024], [2, 4], 0, T_OUT)["stages"])
    assert all(g["ntile"] == 0 for g in pf.describe_schedule("c2c", [1024, 1024, 1024], [3, 2], 0, T_OUT)["stages"])


def test_illegal_combinations_are_refused(built_lib):
    assert pf.describe_schedule("c2c", [8, 8, 8], [2, 2], 0, T_IN | T_OUT)["error"].startswith("illegal")
    assert pf.describe_schedule("r2c", [8, 8, 8], [2, 2], 0, T_IN)["error"].startswith("illegal")
    assert pf.describe_schedule("c2r", [8, 8, 8], [2, 2], 0, T_OUT)["error"].startswith("illegal")
    assert pf.describe_schedule("c2c", [9, 8, 8], [2, 2], 0, S_IN)["error"].startswith("illegal")
    assert pf.describe_schedule("c2c", [8, 8], [2, 2], 0, 0)["error"].startswith("illegal")


def _random_cases(count, seed):
    """Legal random configurations (sizes, pruning, meshes, flags, kinds) -- the planner must either produce a
    schedule that reproduces the oracle or say that the configuration is unsupported."""
    import random
    rnd = random.Random(seed)
    meshes = [[1], [2], [3], [1, 1], [2, 1], [1, 2], [2, 2], [3, 2], [2, 3], [2, 2, 2], [1, 2, 2]]
    out = []
    while len(out) < count:
        np_ = rnd.choice(meshes)
        r = len(np_)
        d = rnd.choice([x for x in (2, 3, 4) if x > r or (x == r == 3)])
        kind = rnd.choice(["c2c", "c2c", "r2c", "c2r", "r2r"])
        shifted = rnd.random() < 0.25 and kind in ("c2c",)
        n = [rnd.choice([4, 6, 8, 10, 12] if shifted else [3, 4, 5, 6, 7, 8, 9, 12]) for _ in range(d)]
        case = dict(kind=kind, n=n, np=np_)
        flags = 0
        tr = rnd.choice([0, T_OUT, T_IN])
        if kind == "r2c" and tr == T_IN:
            tr = T_OUT
        if kind == "c2r" and tr == T_OUT:
            tr = T_IN
        flags |= tr
        if d == r:          # 3-D data on a 3-D mesh: no padded real rows
            pass
        elif kind in ("r2c", "c2r") and rnd.random() < 0.3:
            flags |= PAD
        if shifted:
            flags |= rnd.choice([S_IN, S_OUT, S_IN | S_OUT])
        if rnd.random() < 0.3 and d > r:
            even = 2 if shifted else 1
            ni = [max(2, (x - rnd.randint(0, x // 2)) // even * even) for x in n]
            no = [max(2, (x - rnd.randint(0, x // 2)) // even * even) for x in n]
            if kind == "r2c":
                case["ni"], case["no"] = ni, n
            elif kind == "c2r":
                case["ni"], case["no"] = n, no
            else:
                case["ni"], case["no"] = ni, no
        if kind == "c2c":
            case["sign"] = rnd.choice([-1, +1])
            if rnd.random() < 0.2 and d > r:
                case["howmany"] = rnd.choice([2, 3])
        if kind == "c2r":
            case["sign"] = +1
        if kind == "r2r":
            case["kinds"] = [rnd.choice([po.REDFT00, po.REDFT01, po.REDFT10, po.REDFT11, po.RODFT00, po.RODFT01,
                                         po.RODFT10, po.RODFT11]) for _ in range(d)]
            case["n"] = [max(x, 3) for x in n]
        case["flags"] = flags
        out.append(case)
    return out


@pytest.mark.parametrize("case", _random_cases(60, 20261017), ids=lambda c: "%s-%s-np%s-f%d%s" % (
    c["kind"], "x".join(map(str, c["n"])), "x".join(map(str, c["np"])), c.get("flags", 0), "-pruned" if "ni" in c else ""))
def test_random_configurations_reproduce_oracle(built_lib, case):
    P = int(np.prod(case["np"]))
    s0 = pf.describe_schedule(case["kind"], case["n"], case["np"], 0, case.get("flags", 0), case.get("ni"), case.get("no"),
                              case.get("howmany", 1), None, None, case.get("sign", -1), case.get("kinds"), None)
    if s0["error"].startswith("illegal") or "not supported" in s0["error"]:
        pytest.skip(s0["error"])
    err, scheds = run_virtual(case)
    assert len(scheds) == P
    assert err < 1e-12, err
