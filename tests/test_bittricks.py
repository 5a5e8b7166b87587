"""CPU: the SIMD-in-register tricks of csrc/sketch_scan.cuh, emulated bit for bit in Python and checked
exhaustively / on random words: PRMT-LUT classification, multiply-gather packing, the gather-and-reverse skip
mask, the branch-free squeeze, the prefilter bit layout with shared shifts, and the revcomp helper."""
import numpy as np

M32 = 0xFFFFFFFF


def prmt(x, y, sel):
    src = [(x >> (8 * i)) & 0xFF for i in range(4)] + [(y >> (8 * i)) & 0xFF for i in range(4)]
    r = 0
    for i in range(4):
        nib = (sel >> (4 * i)) & 0xF
        assert nib < 8          # no sign-replicate mode in our selectors
        r |= src[nib] << (8 * i)
    return r


def classify4(w):
    """mirror of classify4(): returns (diff, t3, m)"""
    s1 = w >> 1
    s2 = w >> 2
    t3 = s1 & 0x07070707                      # bits 0-1 = (b >> 1) & 3, bit 2 = bit 3 of b (set for \n and \r only)
    u = w & ~(s1 & 0x20202020) & M32
    a = (t3 | (t3 >> 4)) & M32
    sel = prmt(a, 0, 0x4420) & 0xFFFF
    e = prmt(0x47544341, 0xFF0D0AFF, sel)
    t2 = (s1 ^ s2) & 0x03030303               # A0 C1 G2 T3 for letters (their bit 3 is clear)
    return u ^ e, t3, (t2 * 0x40100401) & M32


REF = {ord(c): i for i, c in enumerate("ACGT")}
REF.update({ord(c): i for i, c in enumerate("acgt")})


def test_classification_is_exact_for_every_byte_in_every_position():
    for b in range(256):
        for pos in range(4):
            w = (0x41414141 & ~(0xFF << (8 * pos))) | (b << (8 * pos))
            diff, t3, m = classify4(w)
            clean = b in REF or b in (10, 13)
            assert ((diff >> (8 * pos)) & 0xFF == 0) == clean, (hex(b), pos)
            for o in range(4):
                if o != pos:
                    assert (diff >> (8 * o)) & 0xFF == 0
            if clean:
                assert bool((t3 >> (8 * pos)) & 4) == (b in (10, 13))       # skip flag = bit 3 of the byte


def test_multiply_gather_packs_codes_oldest_first():
    rng = np.random.default_rng(0)
    letters = np.frombuffer(b"ACGTacgt", dtype=np.uint8)
    for _ in range(5000):
        bs = rng.choice(letters, 4)
        w = int(bs[0]) | int(bs[1]) << 8 | int(bs[2]) << 16 | int(bs[3]) << 24
        _, _, m = classify4(w)
        exp = REF[int(bs[0])] << 6 | REF[int(bs[1])] << 4 | REF[int(bs[2])] << 2 | REF[int(bs[3])]
        assert m >> 24 == exp
    for _ in range(500):            # the PRMT gather of four products' top bytes
        m = [int(x) for x in rng.integers(0, 1 << 32, 4)]
        codes = prmt(prmt(m[3], m[2], 0x0073), prmt(m[1], m[0], 0x0073), 0x5410)
        assert codes == (m[0] >> 24) << 24 | (m[1] >> 24) << 16 | (m[2] >> 24) << 8 | (m[3] >> 24)


def rev_flags8(t_old, t_new):
    g = (((t_new >> 2) | (t_old << 2)) & 0x11111111) & M32
    return ((g * 0x08040201) & M32) >> 24


def test_gather_and_reverse_skip_mask():
    rng = np.random.default_rng(1)
    for _ in range(20000):
        fl = rng.integers(0, 2, 16)
        words = []
        for i in range(4):
            w = 0
            for j in range(4):
                w |= (int(fl[4 * i + j]) << 2 | int(rng.integers(0, 4))) << (8 * j)
            words.append(w)
        rsk = (rev_flags8(words[0], words[1]) << 8) | rev_flags8(words[2], words[3])
        assert rsk == sum(1 << (15 - b) for b in range(16) if fl[b])


def squeeze(c, rsk):
    while True:
        iso = rsk & -rsk
        low = (iso * iso - 1) & M32
        c = (((c >> 2) & ~low) | (c & low)) & M32
        rsk = (rsk ^ iso) >> 1
        if rsk == 0:
            return c


def test_branch_free_squeeze():
    rng = np.random.default_rng(2)
    for trial in range(20000):
        codes = rng.integers(0, 4, 16)
        fl = rng.integers(0, 2, 16) if trial % 3 else (rng.integers(0, 16, 16) == 0).astype(int)
        c = sum(int(codes[b]) << (30 - 2 * b) for b in range(16))
        rsk = sum(1 << (15 - b) for b in range(16) if fl[b])
        exp = 0
        for b in range(16):
            if not fl[b]:
                exp = (exp << 2) | int(codes[b])
        assert squeeze(c, rsk) == exp


def test_prefilter_layout_shares_shifts():
    """word offset of window d = t[d-1] & (0x7fff << 2), shift amount = t[d+8] (low 5 bits), t[e] = X >> 2e."""
    rng = np.random.default_rng(3)
    for _ in range(2000):
        X = int(rng.integers(0, 1 << 62)) << 34 | int(rng.integers(0, 1 << 34))      # 96 random bits
        t = lambda e: ((X << 2) if e < 0 else (X >> (2 * e))) & M32
        for d in range(32):
            v = (X >> (2 * d)) & M32
            assert t(d - 1) & (0x7FFF << 2) == (v & 0x7FFF) << 2
            assert t(d + 8) & 31 == (v >> 16) & 31


def revcomp2(x, nb):
    out = 0
    for i in range(nb):
        out |= (3 - ((x >> (2 * i)) & 3)) << (2 * (nb - 1 - i))
    return out


def test_revcomp_matches_rolling_definition():
    """reference iseq2comem.c:218: crvstuple = (crvstuple >> 2) + ((b ^ 3) << (4k - 2))"""
    rng = np.random.default_rng(4)
    for k in (8, 10, 11, 16):
        TL = 2 * k
        for _ in range(200):
            bases = rng.integers(0, 4, TL)
            fwd = rc = 0
            for b in bases:
                fwd = ((fwd << 2) | int(b)) & ((1 << (4 * k)) - 1)
                rc = (rc >> 2) + ((int(b) ^ 3) << (4 * k - 2))
            assert revcomp2(fwd, TL) == rc
