"""CPU: the lane arithmetic of the clean path of csrc/sketch_scan3.cuh, emulated bit for bit (multiply-gather
classification, squeeze, history hand-over, Y / X assembly, block probes, the per-block table of window members, k-mer
extraction, byte-offset recovery, the resolver's change of representation), against a direct per-base walk of the
reference's rules.  The kernel itself is checked on the GPU against the oracle; this pins the formulas on a CPU box."""
import random, sys
M32 = 0xFFFFFFFF
M64 = (1 << 64) - 1

def prmt(x, y, sel):
    src = [(x >> (8 * i)) & 0xFF for i in range(4)] + [(y >> (8 * i)) & 0xFF for i in range(4)]
    r = 0
    for i in range(4):
        r |= src[(sel >> (4 * i)) & 7] << (8 * i)
    return r

def fsl(lo, hi, s):   # __funnelshift_l wrap
    s &= 31
    return (((hi << 32) | lo) << s >> 32) & M32
def fslc(lo, hi, s):
    s = min(s, 32)
    return (((hi << 32) | lo) << s >> 32) & M32
def fsr(lo, hi, s):
    s &= 31
    return (((hi << 32) | lo) >> s) & M32
def fsrc(lo, hi, s):
    s = min(s, 32)
    return (((hi << 32) | lo) >> s) & M32

def classify_lazy8(w0, w1):
    f0 = w0 & 0x08080808; f1 = w1 & 0x08080808
    c0 = ((w0 & 0x06060606) * 0x00820820) & M32
    c1 = ((w1 & 0x06060606) * 0x00820820) & M32
    d = (w0 & ((f0 * 0x1e) & M32)) | (w1 & ((f1 * 0x1e) & M32))
    g = (f0 * 0x00204081 + f1 * 0x02040810) & M32
    return d, c0, c1, g

def top_bytes4(a, b, c, d):
    return prmt(prmt(a, b, 0x0073), prmt(c, d, 0x0073), 0x5410)

def rev_groups64(x, nb):
    r = 0
    for i in range(nb):
        r |= ((x >> (2 * i)) & 3) << (2 * (nb - 1 - i))
    return r
def to_scan_repr(x, nb):
    r = rev_groups64(x, nb)
    return r ^ ((r >> 1) & 0x5555555555555555)
def revcomp_ref(x, nb):
    r = 0
    for i in range(nb):
        r |= (3 - ((x >> (2 * i)) & 3)) << (2 * (nb - 1 - i))
    return r
def gtab_ext(win, r):
    x = (win & ((1 << (2 * r)) - 1)) | ((win >> (2 * r + 20)) << (2 * r))
    return (x ^ (x >> 4)) & 15
def pf3b_index(win):
    return ((win * 0x9E3779B1) & M32) >> (32 - 18)
def low_mask(b):
    return M32 if b >= 32 else (1 << b) - 1

REF = {c: i for i, c in enumerate(b"ACGT")}
REF.update({c: i for i, c in enumerate(b"acgt")})

def run(k, s, L, text, S_set, ST):
    """S_set: set of sampled inner 2s-mers (reference repr). returns list of (ord, kmer_ref) found by emulation"""
    TL = 2 * k; out = k - s; BIG = 2 * (TL - 1) >= 32
    hsh = (2 * (TL - 1)) & 31
    innermask = (1 << (4 * s)) - 1; tupmask = (1 << (4 * k)) - 1
    wbits = 4 * s
    # bitmaps
    pf = [0] * (1 << 15); gtab = {}
    reps3 = 1 << (20 - wbits) if wbits < 20 else 1
    for d in S_set:
        for x in (d, revcomp_ref(d, 2 * s)):
            y = to_scan_repr(x, 2 * s)
            for r in range(ST):
                for hi in range(reps3):
                    b = ((y >> (2 * r)) | (hi << wbits)) & 0xfffff
                    pf[b & 0x7fff] |= 0x80000000 >> (b >> 15)
                    if ST == 3: gtab[b] = gtab.get(b, 0) | (1 << (16 * r + gtab_ext(y, r)))
    NPROBE = 12 if ST == 3 else 32
    assert len(text) % 1024 == 0
    res = []
    cw0 = cw1 = 0
    since_break = 0
    for it in range(len(text) // 1024):
        S_prev = None
        lanes = []
        for lane in range(32):
            bs = text[it * 1024 + 32 * lane: it * 1024 + 32 * lane + 32]
            w = [int.from_bytes(bs[4 * i:4 * i + 4], 'little') for i in range(8)]
            dacc = 0; c = [0] * 8; g = [0] * 4
            for p in range(4):
                d, c[2 * p], c[2 * p + 1], g[p] = classify_lazy8(w[2 * p], w[2 * p + 1])
                dacc |= d
            F = top_bytes4(*g)
            nA = 16 - bin(F & 0xffff).count("1"); n = 32 - bin(F).count("1")
            assert dacc == 0 and n >= TL - 1, "emulator only covers clean steady iterations"
            PA = top_bytes4(c[0], c[1], c[2], c[3]); PB = top_bytes4(c[4], c[5], c[6], c[7])
            fa = F & 0xffff; fb = F >> 16
            while True:
                ia = fa & (-fa & M32); ib = fb & (-fb & M32)
                la = (ia * ia - 1) & M32; lb = (ib * ib - 1) & M32
                PA = ((PA >> 2) & ~la & M32) | (PA & la)
                PB = ((PB >> 2) & ~lb & M32) | (PB & lb)
                fa = (fa ^ ia) >> 1; fb = (fb ^ ib) >> 1
                if not (fa | fb): break      # (kernel: warp-wide any; extra rounds are no-ops)
            Q0 = PA | fslc(0, PB, 2 * nA); Q1 = fslc(PB, 0, 2 * nA)
            d = 2 * (n - (TL - 1))
            if BIG: S0 = fsrc(Q0, Q1, d); S1 = fsrc(Q1, 0, d)
            else: S0 = (((Q1 << 32) | Q0) >> d) & M32; S1 = 0
            lanes.append((F, n, Q0, Q1, S0, S1))
        N = sum(l[1] for l in lanes)
        o_l = 0
        for lane in range(32):
            F, n, Q0, Q1, S0, S1 = lanes[lane]
            if lane == 0: H0, H1 = cw0, cw1
            else: H0, H1 = lanes[lane - 1][4], (lanes[lane - 1][5] if BIG else 0)
            if BIG:
                Y0 = H0; Y1 = H1 | ((Q0 << hsh) & M32); Y2 = fsl(Q0, Q1, hsh); Y3 = fslc(Q1, 0, hsh)
            else:
                Y0 = H0 | ((Q0 << hsh) & M32); Y1 = fsl(Q0, Q1, hsh); Y2 = fslc(Q1, 0, hsh); Y3 = 0
            X0 = fsr(Y0, Y1, 2 * out); X1 = fsr(Y1, Y2, 2 * out); X2 = fsr(Y2, Y3, 2 * out)
            def xsh(kk):
                if kk < 0: return (X0 << (-kk)) & M32
                if kk < 32: return fsr(X0, X1, kk)
                if kk < 64: return fsr(X1, X2, kk - 32)
                return X2 >> (kk - 64)
            cand = 0
            for i in range(NPROBE - 1, -1, -1):
                word = pf[(xsh(2 * ST * i - 2) & 0x1fffc) >> 2]
                cand = fsl(fsl(0, word, xsh(2 * ST * i + 15)), cand, 1)
            wm = low_mask(n)
            need = TL - 1 - since_break - o_l
            if need > 0: wm &= ~low_mask(min(need, 32)) & M32
            o_l += n
            if cand:
                Y = [Y0, Y1, Y2, Y3]
                hitsj = []
                if ST == 1:
                    c2 = cand & wm
                    while c2:
                        j = (c2 & -c2).bit_length() - 1; c2 &= c2 - 1
                        hitsj.append(j)
                else:
                    c2 = cand
                    while c2:
                        i = (c2 & -c2).bit_length() - 1; c2 &= c2 - 1
                        o = 2 * (out + 3 * i) - 4
                        if o >= 0:
                            a = o >> 5
                            S = fsr(Y[a], Y[a + 1] if a < 3 else 0, o)
                        else:
                            S = (Y[0] << (-o)) & M32
                        mm = gtab.get((S >> 4) & 0xfffff, 0)
                        for r in range(3):
                            j = 3 * i - r
                            ee = gtab_ext((S >> (2 * (2 - r))) & innermask, r)
                            if 0 <= j < 32 and (wm >> j) & 1 and (mm >> (16 * r + ee)) & 1: hitsj.append(j)
                for j in hitsj:
                    a = (2 * j) >> 5
                    kmer = ((fsr(Y[a + 1], Y[a + 2], 2 * j) << 32) | fsr(Y[a], Y[a + 1], 2 * j)) & tupmask
                    cnt = 0; sub = None
                    for b in range(32):
                        if not (F >> b) & 1:
                            cnt += 1
                            if cnt == j + 1: sub = b; break
                    # resolver
                    yf = kmer ^ ((kmer >> 1) & 0x5555555555555555)
                    fwd = rev_groups64(yf, TL); rc = ~yf & tupmask
                    u = min(fwd, rc)
                    inner = (u >> (2 * out)) & innermask
                    if inner in S_set:
                        res.append((it * 1024 + 32 * lane + sub, fwd))
        since_break = min(since_break + N, 64)
        cw0, cw1 = lanes[31][4], (lanes[31][5] if BIG else 0)
    return res

def direct(k, s, L, text, S_set):
    TL = 2 * k; out = k - s
    innermask = (1 << (4 * s)) - 1; tupmask = (1 << (4 * k)) - 1
    fwd = 0; run_ = 0; res = []
    for p, ch in enumerate(text):
        if ch in REF:
            fwd = ((fwd << 2) | REF[ch]) & tupmask; run_ += 1
            if run_ >= TL:
                rc = revcomp_ref(fwd, TL)
                u = min(fwd, rc)
                if ((u >> (2 * out)) & innermask) in S_set: res.append((p, fwd))
        elif ch in (10, 13): pass
        else: run_ = 0
    return res

def make_text(nbytes, width, seed, crlf=False):
    rnd = random.Random(seed)
    out = bytearray()
    col = 0
    while len(out) < nbytes:
        out.append(rnd.choice(b"ACGTacgt"))
        col += 1
        if col == width:
            if crlf: out += b"\r\n"
            else: out += b"\n"
            col = 0
    return bytes(out[:nbytes])

def run_configs():
    ok = True
    for (k, s, L, rate) in [(10, 6, 3, 1 / 256), (8, 5, 2, 1 / 64), (11, 6, 3, 1 / 256), (10, 7, 4, 1 / 1024), (16, 6, 3, 1 / 256), (9, 6, 3, 1 / 256), (12, 6, 3, 1/256)]:
        rnd = random.Random(k * 100 + s)
        text = make_text(8192, 80 if k != 9 else 61, k, crlf=(k == 11))
        # sampled set: take inner windows actually present so hits exist, plus random ones
        TL = 2 * k; out = k - s
        innermask = (1 << (4 * s)) - 1
        S_set = set()
        allw = [x for _, x in direct(k, s, L, text, set(range(0)) ) ]
        # collect all inner canon windows
        fwd = 0; run_ = 0; inners = []
        for ch in text:
            if ch in REF:
                fwd = ((fwd << 2) | REF[ch]) & ((1 << (4 * k)) - 1); run_ += 1
                if run_ >= TL:
                    u = min(fwd, revcomp_ref(fwd, TL)); inners.append((u >> (2 * out)) & innermask)
        for x in inners:
            if rnd.random() < 0.02: S_set.add(x)
        for _ in range(2000): S_set.add(rnd.getrandbits(4 * s))
        ST = 3 if 2 * s >= 12 else 1
        a = run(k, s, L, text, S_set, ST)
        b = direct(k, s, L, text, S_set)
        same = sorted(a) == sorted(b)
        print(k, s, L, "ST", ST, len(a), len(b), "OK" if same else "MISMATCH")
        if not same:
            sa, sb = set(a), set(b)
            print("  only emu", sorted(sa - sb)[:5], " only direct", sorted(sb - sa)[:5])
            ok = False
    return ok


def test_clean_path_arithmetic_matches_direct_walk():
    assert run_configs()


if __name__ == "__main__":
    sys.exit(0 if run_configs() else 1)
