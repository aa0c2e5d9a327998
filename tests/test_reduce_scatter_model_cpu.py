"""The warp shuffle networks of the blend backward (csrc/blend.cu: rs_step / reduce_scatter<4> and the opt-in value-scatter
variant), modelled lane by lane in numpy: every lane that is supposed to own a total must hold exactly the 32-lane sum.  This pins
the wiring (which lane keeps which operand at which xor distance) on CPU; the arithmetic on the GPU is the same additions."""
import numpy as np

LANES = np.arange(32)
NT, RB = 9, 4


def shfl_xor(v, m):
    return v[LANES ^ m]


def rs_step(lo, hi, up, m):      # blend.cu: rs_step — `up` lanes keep hi and send lo, the others keep lo and send hi
    send = np.where(up, lo, hi)
    keep = np.where(up, hi, lo)
    return keep + shfl_xor(send, m)


def bits():
    return [(LANES & b) != 0 for b in (16, 8, 4, 2, 1)]


def slot4():
    return ((LANES >> 4) & 1) * 2 + ((LANES >> 3) & 1)       # blend.cu: rs_slot<4>


def test_reduce_scatter_rb4():
    part = np.random.default_rng(0).standard_normal((32, RB))
    b16, b8, _, _, _ = bits()
    s = rs_step(rs_step(part[:, 0], part[:, 2], b16, 16), rs_step(part[:, 1], part[:, 3], b16, 16), b8, 8)
    for m in (4, 2, 1):
        s = s + shfl_xor(s, m)
    tot = part.sum(axis=0)
    assert np.allclose(s, tot[slot4()])           # EVERY lane of slot u's group ends with the total of entry u


def test_value_scatter_variant():
    part = np.random.default_rng(1).standard_normal((32, NT, RB))
    b16, b8, b4, b2, b1 = bits()
    sv = [rs_step(rs_step(part[:, k, 0], part[:, k, 2], b16, 16), rs_step(part[:, k, 1], part[:, k, 3], b16, 16), b8, 8) for k in range(NT)]
    r0, r1 = rs_step(sv[1], sv[5], b4, 4), rs_step(sv[2], sv[6], b4, 4)
    r3, r4 = rs_step(sv[3], sv[7], b4, 4), rs_step(sv[4], sv[8], b4, 4)
    r2 = sv[0] + shfl_xor(sv[0], 4)
    u0, u1 = rs_step(r0, r3, b2, 2), rs_step(r1, r4, b2, 2)
    u2 = r2 + shfl_xor(r2, 2)
    w = rs_step(u0, u1, b1, 1)
    t2 = u1 + shfl_xor(u1, 1)
    go = u2 + shfl_xor(u2, 1)
    tot = part.sum(axis=0)                        # [value, entry]
    owner_value = {0: 1, 2: 3, 3: 4, 4: 5, 5: 6, 6: 7, 7: 8}      # lane & 7 -> index of the total it holds in `w`
    slot = slot4()
    for lane in range(32):
        l8, u = lane & 7, slot[lane]
        if l8 == 1:
            assert np.isclose(go[lane], tot[0, u])                 # dL/dopacity
        else:
            assert np.isclose(w[lane], tot[owner_value[l8], u])
        if l8 == 0:
            assert np.isclose(t2[lane], tot[2, u])                 # the coupled pair lives on one lane


def test_bit_matrix_transpose32():
    """csrc/binning.cu: transpose32 — five xor-shuffle stages; lane t must end with bit l = bit t of lane l's input."""
    rng = np.random.default_rng(2)
    x = rng.integers(0, 2 ** 32, size=32, dtype=np.uint64).astype(np.uint32)
    orig = x.copy()
    for j, m in ((16, 0x0000FFFF), (8, 0x00FF00FF), (4, 0x0F0F0F0F), (2, 0x33333333), (1, 0x55555555)):
        m = np.uint32(m)
        y = x[LANES ^ j]
        hi = (LANES & j) != 0
        x = np.where(hi, (x & ~m) | ((y >> np.uint32(j)) & m), (x & m) | ((y << np.uint32(j)) & ~m)).astype(np.uint32)
    for t in range(32):
        for l in range(32):
            assert ((int(x[t]) >> l) & 1) == ((int(orig[l]) >> t) & 1)
