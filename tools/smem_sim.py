"""Shared-memory wavefront model for the ND hex kernel's access patterns (design aid, not shipped code).
Model: a warp request of W-byte accesses is split into groups of 128/W... lanes (64-bit: half-warps, 128-bit:
quarter-warps, 32-bit: full warp); wavefronts of a group = max over the 32 4-byte banks of the number of distinct
4-byte words touched in that bank."""
import itertools, sys
from collections import defaultdict


def wavefronts(addrs_bytes, width, model="group"):
    """addrs_bytes: list of (lane, byte address) for active lanes."""
    lanes_per_group = {4: 32, 8: 16, 16: 8}[width]
    total = 0
    groups = defaultdict(list)
    for lane, a in addrs_bytes:
        groups[lane // lanes_per_group if model == "group" else 0].append(a)
    for g, al in groups.items():
        bank_words = defaultdict(set)
        for a in al:
            for wd in range(width // 4):
                word = a // 4 + wd
                bank_words[word % 32].add(word)
        total += max(len(s) for s in bank_words.values())
    return total


def ideal(addrs_bytes, width):
    words = set()
    for lane, a in addrs_bytes:
        for wd in range(width // 4):
            words.add(a // 4 + wd)
    return (len(words) + 31) // 32


class Tally:
    def __init__(self):
        self.d = defaultdict(lambda: [0, 0, 0])

    def add(self, name, addrs, width):
        if not addrs:
            return
        t = self.d[name]
        t[0] += wavefronts(addrs, width, "group")
        t[1] += wavefronts(addrs, width, "any")
        t[2] += ideal(addrs, width)

    def report(self):
        tot = [0, 0, 0]
        for k, v in self.d.items():
            print(f"{k:28s} group {v[0]:5d}  any {v[1]:5d}  ideal {v[2]:5d}")
            for i in range(3):
                tot[i] += v[i]
        print(f"{'TOTAL':28s} group {tot[0]:5d}  any {tot[1]:5d}  ideal {tot[2]:5d}")
        return tot


def current_kernel(p=3, q=4, KIND=2):
    n = p + 1
    MASS, CURL = KIND in (1, 2), KIND in (0, 2)
    Q = q ** 3
    P = 3 * p * n * n
    D3 = p * n * n
    PS = (P + 3) & ~3
    ZXA = 0; ZXB = ZXA + p * n * q; ZYA = ZXB + p * n * q; ZYB = ZYA + n * p * q; ZZA = ZYB + n * p * q; ZSZ = ZZA + n * n * q
    YX1 = ZSZ; YX2 = YX1 + p * q * q; YX3 = YX2 + p * q * q; YY1 = YX3 + p * q * q; YY2 = YY1 + n * q * q
    YZ1 = YY2 + n * q * q; YZ3 = YZ1 + n * q * q; YEND = YZ3 + n * q * q
    ES = (YEND + 1) & ~1
    NEW = 1 if q * q >= 32 else 32 // (q * q)
    QQ = q * q
    GE = 10 * Q
    T = Tally()
    W0 = 0  # sW base (doubles)
    # ---- phase Z
    IX, IZZ = NEW * p * n, NEW * n * n
    R = (max(IX, IZZ) + 31) // 32
    for r in range(R):
        lanes = range(32)
        def wx(l):
            w = l + 32 * r
            return w if w < IX else 0
        def wz(l):
            w = l + 32 * r
            return w if w < IZZ else 0
        for k in range(n):
            T.add("Z ld U(x)", [(l, 8 * ((wx(l) // (p * n)) * PS + wx(l) % (p * n) + p * n * k)) for l in lanes], 8)
            T.add("Z ld I(x)", [(l, 4 * ((wx(l) // (p * n)) * PS + wx(l) % (p * n) + p * n * k)) for l in lanes], 4)
            T.add("Z ld U(y)", [(l, 8 * ((wx(l) // (p * n)) * PS + D3 + wx(l) % (p * n) + p * n * k)) for l in lanes], 8)
            T.add("Z ld I(y)", [(l, 4 * ((wx(l) // (p * n)) * PS + D3 + wx(l) % (p * n) + p * n * k)) for l in lanes], 4)
        for k in range(p):
            T.add("Z ld U(z)", [(l, 8 * ((wz(l) // (n * n)) * PS + 2 * D3 + wz(l) % (n * n) + n * n * k)) for l in lanes], 8)
            T.add("Z ld I(z)", [(l, 4 * ((wz(l) // (n * n)) * PS + 2 * D3 + wz(l) % (n * n) + n * n * k)) for l in lanes], 4)
        ax = [l for l in lanes if l + 32 * r < IX]
        az = [l for l in lanes if l + 32 * r < IZZ]
        for qz in range(q):
            for off, on in ((ZXA, True), (ZXB, CURL), (ZYA, True), (ZYB, CURL)):
                if on:
                    T.add("Z st Z(xy)", [(l, 8 * ((wx(l) // (p * n)) * ES + off + q * (wx(l) % (p * n)) + qz)) for l in ax], 8)
            T.add("Z st Z(z)", [(l, 8 * ((wz(l) // (n * n)) * ES + ZZA + q * (wz(l) % (n * n)) + qz)) for l in az], 8)
    # ---- phase Y
    IX, IN = NEW * p * q, NEW * n * q
    R = (IN + 31) // 32
    for r in range(R):
        lanes = range(32)
        def wx(l):
            w = l + 32 * r
            return w if w < IX else 0
        def wn(l):
            w = l + 32 * r
            return w if w < IN else 0
        def dec(w, m):
            e, t = w // (m * q), w % (m * q)
            return e, t, t % q, t // q
        for j in range(n):
            for off, on in ((ZXA, True), (ZXB, CURL)):
                if on:
                    T.add("Y ld Z(x)", [(l, 8 * (dec(wx(l), p)[0] * ES + off + dec(wx(l), p)[2] + q * dec(wx(l), p)[3] + q * p * j)) for l in lanes], 8)
            T.add("Y ld Z(z)", [(l, 8 * (dec(wn(l), n)[0] * ES + ZZA + dec(wn(l), n)[2] + q * dec(wn(l), n)[3] + q * n * j)) for l in lanes], 8)
        for j in range(p):
            for off, on in ((ZYA, True), (ZYB, CURL)):
                if on:
                    T.add("Y ld Z(y)", [(l, 8 * (dec(wn(l), n)[0] * ES + off + dec(wn(l), n)[2] + q * dec(wn(l), n)[3] + q * n * j)) for l in lanes], 8)
        ax = [l for l in lanes if l + 32 * r < IX]
        an = [l for l in lanes if l + 32 * r < IN]
        for qy in range(q):
            for off, on in ((YX1, MASS), (YX2, CURL), (YX3, CURL)):
                if on:
                    T.add("Y st Y(x)", [(l, 8 * (dec(wx(l), p)[0] * ES + off + q * dec(wx(l), p)[1] + qy)) for l in ax], 8)
            for off, on in ((YY1, True), (YY2, CURL), (YZ1, True), (YZ3, CURL)):
                if on:
                    T.add("Y st Y(yz)", [(l, 8 * (dec(wn(l), n)[0] * ES + off + q * dec(wn(l), n)[1] + qy)) for l in an], 8)
    # ---- XDX
    for w0 in range(0, NEW * QQ, 32):
        lanes = [l for l in range(32) if w0 + l < NEW * QQ]
        def es(l):
            w = w0 + l
            return w // QQ, w % QQ
        for i in range(p):
            for off, on in ((YX1, MASS), (YX2, CURL), (YX3, CURL)):
                if on:
                    T.add("XDX ld Y", [(l, 8 * (es(l)[0] * ES + es(l)[1] + off + QQ * i)) for l in lanes], 8)
                    T.add("XDX st Y", [(l, 8 * (es(l)[0] * ES + es(l)[1] + off + QQ * i)) for l in lanes], 8)
        for i in range(n):
            for off, on in ((YY1, True), (YY2, CURL), (YZ1, True), (YZ3, CURL)):
                if on:
                    T.add("XDX ld Y", [(l, 8 * (es(l)[0] * ES + es(l)[1] + off + QQ * i)) for l in lanes], 8)
                    T.add("XDX st Y", [(l, 8 * (es(l)[0] * ES + es(l)[1] + off + QQ * i)) for l in lanes], 8)
        for qx in range(q):
            for c in range(10):
                T.add("XDX ld G", [(l, 8 * (es(l)[0] * GE + es(l)[1] + QQ * qx + c * Q)) for l in lanes], 8)
            for c in (0, 9):
                T.add("XDX ld C", [(l, 8 * (es(l)[0] * 18 + c)) for l in lanes], 8)
    # ---- Yt : loads are thread-contiguous vectors (as Y stores), stores as Y loads
    IX, IN = NEW * p * q, NEW * n * q
    for r in range((IN + 31) // 32):
        lanes = range(32)
        def wx(l):
            w = l + 32 * r
            return w if w < IX else 0
        def wn(l):
            w = l + 32 * r
            return w if w < IN else 0
        def dec(w, m):
            e, t = w // (m * q), w % (m * q)
            return e, t, t % q, t // q
        ax = [l for l in lanes if l + 32 * r < IX]
        an = [l for l in lanes if l + 32 * r < IN]
        for qy in range(q):
            for off, on in ((YX1, MASS), (YX2, CURL), (YX3, CURL)):
                if on:
                    T.add("Yt ld Y(x)", [(l, 8 * (dec(wx(l), p)[0] * ES + off + q * dec(wx(l), p)[1] + qy)) for l in lanes], 8)
            for off, on in ((YY1, True), (YY2, CURL), (YZ1, True), (YZ3, CURL)):
                if on:
                    T.add("Yt ld Y(yz)", [(l, 8 * (dec(wn(l), n)[0] * ES + off + q * dec(wn(l), n)[1] + qy)) for l in lanes], 8)
        for j in range(n):
            for off, on in ((ZXA, True), (ZXB, CURL)):
                if on:
                    T.add("Yt st Z(x)", [(l, 8 * (dec(wx(l), p)[0] * ES + off + dec(wx(l), p)[2] + q * dec(wx(l), p)[3] + q * p * j)) for l in ax], 8)
            T.add("Yt st Z(z)", [(l, 8 * (dec(wn(l), n)[0] * ES + ZZA + dec(wn(l), n)[2] + q * dec(wn(l), n)[3] + q * n * j)) for l in an], 8)
        for j in range(p):
            for off, on in ((ZYA, True), (ZYB, CURL)):
                if on:
                    T.add("Yt st Z(y)", [(l, 8 * (dec(wn(l), n)[0] * ES + off + dec(wn(l), n)[2] + q * dec(wn(l), n)[3] + q * n * j)) for l in an], 8)
    # ---- Zt
    IX, IZZ = NEW * p * n, NEW * n * n
    for r in range((max(IX, IZZ) + 31) // 32):
        lanes = range(32)
        def wx(l):
            w = l + 32 * r
            return w if w < IX else 0
        def wz(l):
            w = l + 32 * r
            return w if w < IZZ else 0
        for qz in range(q):
            for off, on in ((ZXA, True), (ZXB, CURL), (ZYA, True), (ZYB, CURL)):
                if on:
                    T.add("Zt ld Z(xy)", [(l, 8 * ((wx(l) // (p * n)) * ES + off + q * (wx(l) % (p * n)) + qz)) for l in lanes], 8)
            T.add("Zt ld Z(z)", [(l, 8 * ((wz(l) // (n * n)) * ES + ZZA + q * (wz(l) % (n * n)) + qz)) for l in lanes], 8)
        for k in range(n):
            T.add("Zt ld I", [(l, 4 * ((wx(l) // (p * n)) * PS + wx(l) % (p * n) + p * n * k)) for l in lanes], 4)
            T.add("Zt ld I", [(l, 4 * ((wx(l) // (p * n)) * PS + D3 + wx(l) % (p * n) + p * n * k)) for l in lanes], 4)
        for k in range(p):
            T.add("Zt ld I", [(l, 4 * ((wz(l) // (n * n)) * PS + 2 * D3 + wz(l) % (n * n) + n * n * k)) for l in lanes], 4)
    return T


if __name__ == "__main__" and len(sys.argv) <= 3:
    p = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    q = int(sys.argv[2]) if len(sys.argv) > 2 else p + 1
    T = current_kernel(p, q)
    T.report()


def pow2ceil(x):
    r = 1
    while r < x:
        r *= 2
    return r


def lane_stride(items, NEW):
    """Per-element lane stride for the Z / Zt phases: pad to a power of two when that costs no extra round."""
    pc = pow2ceil(items)
    rounds = (NEW * items + 31) // 32
    return pc if NEW * pc <= 32 * rounds else items


def new_layout(p, q, KIND=2, PADA=0, PADB=0, PADY=0, lane_pad=True, detail=False):
    n = p + 1
    MASS, CURL = KIND in (1, 2), KIND in (0, 2)
    QQ = q * q
    NEW = 1 if QQ >= 32 else 32 // QQ
    nXA, nNA = p * q, n * q
    A_XA = 0; A_XB = A_XA + NEW * nXA; A_ZA = A_XB + (NEW * nXA if CURL else 0); LA = A_ZA + NEW * nNA; RSA = LA + PADA
    B_YA = 0; B_YB = NEW * nNA; LB = B_YB + (NEW * nNA if CURL else 0); RSB = LB + PADB
    ZA0 = 0; ZB0 = n * RSA; ZSZ = ZB0 + p * RSB
    Y_X1 = 0; Y_X2 = Y_X1 + (NEW * nXA if MASS else 0); Y_X3 = Y_X2 + (NEW * nXA if CURL else 0)
    Y_Y1 = Y_X3 + (NEW * nXA if CURL else 0); Y_Y2 = Y_Y1 + NEW * nNA; Y_Z1 = Y_Y2 + (NEW * nNA if CURL else 0)
    Y_Z3 = Y_Z1 + NEW * nNA; LY = Y_Z3 + (NEW * nNA if CURL else 0); RSY = LY + PADY
    Y0 = ZSZ; WTOT = Y0 + q * RSY
    T = Tally()
    # phase Z stores (== Zt loads)
    ipx, ipz = p * n, n * n
    lsx = lane_stride(ipx, NEW) if lane_pad else ipx
    lsz = lane_stride(ipz, NEW) if lane_pad else ipz
    def zlanes(ls, items):
        out = []
        for r in range((NEW * ls + 31) // 32):
            rr = []
            for l in range(32):
                w = l + 32 * r
                e, t = w // ls, w % ls
                if e < NEW and t < items:
                    rr.append((l, e, t))
            out.append(rr)
        return out
    for rr in zlanes(lsx, ipx):
        for qz in range(q):
            for off, on in ((A_XA, True), (A_XB, CURL)):
                if on:
                    T.add("Z st XA/XB", [(l, 8 * (ZA0 + (t // p) * RSA + off + e * nXA + qz + q * (t % p))) for l, e, t in rr], 8)
            for off, on in ((B_YA, True), (B_YB, CURL)):
                if on:
                    T.add("Z st YA/YB", [(l, 8 * (ZB0 + (t // n) * RSB + off + e * nNA + qz + q * (t % n))) for l, e, t in rr], 8)
    for rr in zlanes(lsz, ipz):
        for qz in range(q):
            T.add("Z st ZA", [(l, 8 * (ZA0 + (t // n) * RSA + A_ZA + e * nNA + qz + q * (t % n))) for l, e, t in rr], 8)
    # phase Y loads/stores: contiguous in w -> ideal by construction (count them)
    IX, IN = NEW * p * q, NEW * n * q
    for r in range((IN + 31) // 32):
        ax = [(l, l + 32 * r) for l in range(32) if l + 32 * r < IX]
        an = [(l, l + 32 * r) for l in range(32) if l + 32 * r < IN]
        for j in range(n):
            for off, on in ((A_XA, True), (A_XB, CURL)):
                if on:
                    T.add("Y ld Z", [(l, 8 * (ZA0 + j * RSA + off + w)) for l, w in ax], 8)
            T.add("Y ld Z", [(l, 8 * (ZA0 + j * RSA + A_ZA + w)) for l, w in an], 8)
        for j in range(p):
            for off, on in ((B_YA, True), (B_YB, CURL)):
                if on:
                    T.add("Y ld Z", [(l, 8 * (ZB0 + j * RSB + off + w)) for l, w in an], 8)
        for qy in range(q):
            for off, on in ((Y_X1, MASS), (Y_X2, CURL), (Y_X3, CURL)):
                if on:
                    T.add("Y st Y", [(l, 8 * (Y0 + qy * RSY + off + w)) for l, w in ax], 8)
            for off, on in ((Y_Y1, True), (Y_Y2, CURL), (Y_Z1, True), (Y_Z3, CURL)):
                if on:
                    T.add("Y st Y", [(l, 8 * (Y0 + qy * RSY + off + w)) for l, w in an], 8)
    # XDX loads (== stores)
    for w0 in range(0, NEW * QQ, 32):
        lanes = [(l, (w0 + l) // QQ, (w0 + l) % QQ) for l in range(32) if w0 + l < NEW * QQ]
        for i in range(p):
            for off, on in ((Y_X1, MASS), (Y_X2, CURL), (Y_X3, CURL)):
                if on:
                    T.add("XDX ld Y", [(l, 8 * (Y0 + (s % q) * RSY + off + e * nXA + s // q + q * i)) for l, e, s in lanes], 8)
        for i in range(n):
            for off, on in ((Y_Y1, True), (Y_Y2, CURL), (Y_Z1, True), (Y_Z3, CURL)):
                if on:
                    T.add("XDX ld Y", [(l, 8 * (Y0 + (s % q) * RSY + off + e * nNA + s // q + q * i)) for l, e, s in lanes], 8)
    if detail:
        T.report()
        print("WTOT doubles", WTOT, "RSA", RSA, "RSB", RSB, "RSY", RSY)
    tot = [0, 0, 0]
    for v in T.d.values():
        for k in range(3):
            tot[k] += v[k]
    return tot, WTOT


def search(p, q, KIND=2):
    best = None
    for pa in range(16):
        for pb in range(16):
            za = None
            t, w = new_layout(p, q, KIND, pa, pb, 0)
            key = (t[0], pa + pb)
            if best is None or key < best[0]:
                best = (key, pa, pb)
    _, pa, pb = best
    besty = None
    for py in range(16):
        t, w = new_layout(p, q, KIND, pa, pb, py)
        key = (t[0], py)
        if besty is None or key < besty[0]:
            besty = (key, py, t, w)
    return pa, pb, besty[1], besty[2], besty[3]


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "search":
    for (pp, qq) in [(1, 2), (1, 3), (1, 4), (1, 5), (1, 6), (1, 7), (2, 3), (2, 4), (2, 5), (2, 6), (2, 7), (3, 4), (3, 5), (3, 6), (3, 7),
                     (4, 5), (4, 6), (4, 7), (5, 6), (5, 7), (6, 7)]:
        for kind in (0, 1, 2):
            pa, pb, py, t, w = search(pp, qq, kind)
            print(f"p={pp} q={qq} kind={kind}: PADA={pa} PADB={pb} PADY={py} wavefronts group {t[0]} any {t[1]} ideal {t[2]}  W={w}")


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "emit":
    print("// Generated by tools/smem_sim.py <p> <q> emit: row paddings (doubles) that make the shared work arrays of")
    print("// nd_hex_apply4_kernel bank-conflict free under the half-warp wavefront model. {p, q, kind, PADA, PADB, PADY}")
    for (pp, qq) in [(1, 2), (1, 3), (1, 4), (1, 5), (1, 6), (1, 7), (2, 3), (2, 4), (2, 5), (2, 6), (2, 7), (3, 4), (3, 5), (3, 6), (3, 7),
                     (4, 5), (4, 6), (4, 7), (5, 6), (5, 7), (6, 7)]:
        for kind in (0, 1, 2):
            pa, pb, py, t, w = search(pp, qq, kind)
            print(f"  {{{pp}, {qq}, {kind}, {pa}, {pb}, {py}}},  // wavefronts {t[0]} (ideal {t[2]})")
