"""CPU tests of the host-side kernel planning (no device): which kernel family serves every stage of the BASELINE
configs, and that the tile geometry the planner picks fits the kernels' resource classes (threads per CTA of the
compiled variant, 227 KB of shared memory).  Uses pfftb200_describe_kernels (include/pfft_b200.h)."""
import itertools

import pytest

import pfft_b200 as pf

T_IN, T_OUT, PAD = 1, 2, 1 << 11
SMEM_LIMIT = 227 * 1024


def fams(**kw):
    return [k["family"] for k in pf.describe_kernels(**kw)]


def test_baseline_configs_take_the_register_resident_kernels(built_lib):
    # config 1 / 2 / 4: chains of power-of-two c2c stages, micro-blocked
    for n, mesh in (([1024] * 3, [1, 1]), ([512] * 3, [2, 4]), ([128] * 4, [2, 2, 2])):
        ks = pf.describe_kernels(kind="c2c", n=n, np_=mesh, pid=0, flags=T_OUT)
        assert [k["family"] for k in ks] == ["pow2"] * len(n)
        assert all(k["micro_blocked"] for k in ks), ks
    # config 3: real lines as packed half-length transforms, fp32 tiles of 16 lines
    fwd = pf.describe_kernels(kind="r2c", n=[1024] * 3, np_=[1, 1], pid=0, flags=T_OUT | PAD, precision="single")
    bwd = pf.describe_kernels(kind="c2r", n=[1024] * 3, np_=[1, 1], pid=0, flags=T_IN | PAD, sign=+1, precision="single")
    assert [k["family"] for k in fwd] == ["reg", "pow2", "pow2"] and [k["family"] for k in bwd] == ["pow2", "pow2", "reg"]
    assert fwd[0]["complex_points"] == 512 and fwd[0]["tl"] == 16 and fwd[0]["threads"] == 1024
    assert fwd[1]["tl"] == 16          # the strided c2c stage: 128-byte runs in fp32
    # config 5: 768-point lines (3 * 256) and their 384-point packed real lines, pruned third not read
    fwd = pf.describe_kernels(kind="r2c", n=[768] * 3, ni=[512] * 3, no=[768] * 3, np_=[2, 4], pid=5, flags=T_OUT)
    bwd = pf.describe_kernels(kind="c2r", n=[768] * 3, ni=[768] * 3, no=[512] * 3, np_=[2, 4], pid=5, flags=T_IN, sign=+1)
    assert [k["family"] for k in fwd] == ["reg"] * 3 and [k["family"] for k in bwd] == ["reg"] * 3
    assert [k["complex_points"] for k in fwd] == [384, 768, 768] and all(k["third_zero"] for k in fwd)
    assert not any(k["third_zero"] for k in bwd)
    # the reference test's odd sizes and a large prime: the any-length kernel (Bluestein for 29, 31, 8191)
    odd = pf.describe_kernels(kind="c2c", n=[29, 27, 31], np_=[2, 2], pid=1)
    assert all(k["family"] == "mixed" for k in odd)
    prime = pf.describe_kernels(kind="c2c", n=[4, 4, 8191], np_=[1, 1], pid=0)[0]
    assert prime["bluestein"] == 1 and prime["transform_length"] == 16384 and prime["global_workspace"] == 1


LENGTHS = [128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096]


@pytest.mark.parametrize("precision", ["double", "single"])
def test_tile_geometry_fits_the_kernel_classes(built_lib, precision):
    """Every stage of a sweep over real / complex lines, meshes, ranks, pruning and layouts: threads within the
    compiled class, whole warps, shared memory within 227 KB."""
    seen = set()
    for L, (kind, flags, sign), mesh in itertools.product(
            LENGTHS, [("c2c", T_OUT, -1), ("c2c", T_IN, +1), ("c2c", 0, -1), ("r2c", T_OUT, -1), ("r2c", T_OUT | PAD, -1),
                      ("c2r", T_IN, +1), ("c2r", 0, +1)], [[1, 1], [2, 2], [3, 2], [4]]):
        for n in ([8, 12, L], [L, 12, 8], [12, L, L]):
            if n[-1] < 32 and kind != "c2c":
                continue
            variants = [dict()]
            if L % 3 == 0:
                pr = [v * 2 // 3 if v == L else v for v in n]
                variants.append(dict(ni=pr, no=n) if kind != "c2r" else dict(ni=n, no=pr))
            for extra in variants:
                P = 1
                for m in mesh:
                    P *= m
                for pid in {0, P - 1}:
                    ks = pf.describe_kernels(kind=kind, n=n, np_=mesh, pid=pid, flags=flags, sign=sign, precision=precision, **extra)
                    assert isinstance(ks, list), (ks, n, mesh, kind)
                    for k in ks:
                        assert "error" not in k, (k, n, mesh, kind, extra)
                        seen.add(k["family"])
                        if k["family"] == "reg":
                            assert 0 < k["threads"] <= k["class_threads"] and k["smem_bytes"] <= SMEM_LIMIT, (k, n, mesh, kind)
                            assert k["class_threads"] in (512, 768, 1024)
                        elif k["family"] == "pow2":
                            assert 0 < k["threads"] <= 1024, k
                        elif not k["global_workspace"]:
                            assert k["smem_bytes"] <= SMEM_LIMIT, k
    assert seen == {"pow2", "reg", "mixed"}
