"""Host-side model of the per-layer opacity constant k_stack_prep computes (movis_b200/csrc/
stack_frames.cuh) and of the fma the fused kernel applies (walk_layer, stack_float.cuh): for every
layer the reference's 256-entry table  fa[a] = trunc(float64(a) * opacity)  (imgproc.py:159) or
a * round(255 opacity) // 255  (imgproc.py:206-208) must equal floor(a * c) for the chosen
c = m / 2^18, or the layer must be flagged for the table fallback.  CPU only."""
import numpy as np

A = np.arange(256, dtype=np.int64)
MAGIC = 12582912  # 1.5 * 2^23


def table(opacity: float, pil: bool) -> np.ndarray:
    if not opacity < 1.0:
        return A.copy()
    if pil:
        return (A * int(round(opacity * 255))) // 255
    return np.trunc(A.astype(np.float64) * opacity).astype(np.int64)


def constant(k: np.ndarray):
    """(m, feasible) exactly as k_stack_prep derives them in integer arithmetic."""
    need = np.zeros(256, dtype=np.int64)
    need[1:] = ((k[1:] << 18) + A[1:] - 1) // A[1:]
    m = int(need.max())
    fits = m <= (1 << 18) and k[0] == 0 and bool(np.all(m * A[1:] < ((k[1:] + 1) << 18)))
    return m, fits


def kernel_fa(m: int) -> np.ndarray:
    """fma.rm(magic + a, c, magic - magic c) - magic with c = m / 2^18, evaluated exactly; also checks
    that every constant involved is representable in fp32 (the kernel's exactness argument)."""
    c = np.float32(m) * np.float32(1.0 / 262144.0)
    assert float(c) * 262144 == m
    mc = MAGIC * m // (1 << 18)
    assert MAGIC * m % (1 << 18) == 0 and mc == 48 * m
    assert float(np.float32(MAGIC - mc)) == MAGIC - mc and MAGIC - mc >= 0
    # exact value of the fma before rounding is magic + a m / 2^18; rounding toward -inf at ulp 1
    return (A * m) >> 18


def test_constant_reproduces_the_reference_tables():
    rng = np.random.default_rng(3)
    ops = list(rng.random(20000)) + [0.0, 1.0, 0.5, 0.25, 0.2, 0.1, 0.7, 1 / 3, 2 / 3, 1e-12, 1 - 1e-12]
    ops += [i / 255 for i in range(256)] + [i / 256 for i in range(257)]
    flagged = 0
    for op in ops:
        for pil in (False, True):
            k = table(float(op), pil)
            m, fits = constant(k)
            if fits:
                assert np.array_equal(kernel_fa(m), k), (op, pil)
            else:
                flagged += 1
                assert not pil, "the PIL table a * p // 255 always has a constant"
    assert flagged < len(ops) // 100


def test_awkward_opacities_are_flagged_for_the_table_fallback():
    # a few ulps below a fraction of small integers: float64 rounding makes trunc(a * op) jump early for
    # some a but not for its multiples, so no single c reproduces the table
    for op in (0.19999999999999998, 0.39999999999999997, 0.7999999999999999, 0.09999999999999999):
        k = table(op, False)
        m, fits = constant(k)
        assert not fits, op
    near = 0
    for a in range(1, 256):
        for kk in range(0, a + 1):
            v = np.nextafter(kk / a, 0.0)
            k = table(float(v), False)
            m, fits = constant(k)
            if fits:
                assert np.array_equal(kernel_fa(m), k)
            else:
                near += 1
    assert near > 0


def test_blend_tail_floor_division_by_255_is_exact_for_both_signs():
    """blend_chan (stack_float.cuh): b + floor(y / 255) is taken as fma.rm(y + 0.5, fl32(1/255), magic + b)
    and PIL's b + rint(y / 255) as fma.rm(y + 127.5, ...), for every integer y = fa (T - b) in
    [-65025, 65025].  The fma is exact before its single rounding towards -inf, so the claim is
    floor((y + h) * c) == floor((y + h) / 255) with c = 0x3B808081 = 8421505 / 2^31, checked here in
    integer arithmetic."""
    import struct
    c_num = 8421505
    assert struct.unpack("<f", struct.pack("<I", 0x3B808081))[0] == c_num / 2.0 ** 31 == float(np.float32(1.0) / np.float32(255.0))
    y = np.arange(-65025, 65026, dtype=np.int64)
    # h = 0.5: (2 y + 1) c_num / 2^32
    assert np.array_equal(((2 * y + 1) * c_num) >> 32, y // 255)
    # h = 127.5: (2 y + 255) c_num / 2^32  ==  floor(y / 255 + 0.5)
    assert np.array_equal(((2 * y + 255) * c_num) >> 32, (2 * y + 255) // 510)
    # and the biased sum stays inside [2^23, 2^24), where the fp32 grid is the integers
    assert MAGIC - 256 >= 2 ** 23 and MAGIC + 255 + 256 < 2 ** 24
