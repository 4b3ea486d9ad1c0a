#!/usr/bin/env python3
"""Generator + limb-level model of the dedicated Montgomery SQUARING for the Pasta base fields (vimz_b200/csrc/fp_sqr.cuh).

One description of the carry chains drives both (a) an integer model that is checked against x*x*R^-1 mod p on random and
edge inputs (run here, on the CPU) and (b) the emitted inline-PTX (every chain one asm statement: PTX has a single carry
flag).  Layout of the computation (p = 2^254 + t, t < 2^128, p[0] = 1, -p^-1 = -1 mod 2^32):
  1. off-diagonal triangle sum_{i<j} a_i a_j B^(i+j) in an even-aligned and an odd-aligned set of 64-bit lanes (28 wide products)
  2. U = E + O*B, doubled by funnel shifts, plus the 8 diagonal squares as one chain of wide multiply-adds  -> T (16 limbs)
  3. 8 reduction-only Montgomery rows on T's low half (3 wide products each: m*p[1..3]; m*p[7] is two shifts), plus T's high half
  => 60 wide products instead of the 88 of fp_mul(a, a).
usage: python tools/gen_fp_sqr.py [--emit]   (no flag: run the model's self-test)"""
import random, sys

M32 = 0xffffffff
P_PALLAS = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001
P_VESTA = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001


def program():
    """List of statements: ('wide', lo, hi, x, y) plain 64-bit product; ('cxx', text, fn) plain C++ with its model; ('chain', ops).
    Chain ops: (kind, dst, x, y, addend, cin, cout) with kind in mad_lo / mad_hi / add (add: dst = x + y)."""
    S = []
    A = [f"a{i}" for i in range(8)]
    e = [f"e{k}" for k in range(16)]
    o = [f"o{k}" for k in range(16)]
    fresh = set()

    def wide(arr, k, i, j):
        S.append(("wide", arr[k], arr[k + 1], A[i], A[j])); fresh.update((arr[k], arr[k + 1]))

    def chain(arr, start, pairs, tail):
        ops, k, first = [], start, True
        for (i, j) in pairs:
            for kind, d in (("mad_lo", arr[k]), ("mad_hi", arr[k + 1])):
                addend = d if d in fresh else "0"
                last = (i, j) == pairs[-1] and kind == "mad_hi" and tail == "none"
                ops.append((kind, d, A[i], A[j], addend, not first, not last))
                fresh.add(d); first = False
            k += 2
        if tail == "carry":
            ops.append(("add", arr[k], "0", "0", None, True, False)); fresh.add(arr[k])
        S.append(("chain", ops))

    wide(o, 0, 0, 1); wide(o, 2, 0, 3); wide(o, 4, 0, 5); wide(o, 6, 0, 7)
    wide(e, 2, 0, 2); wide(e, 4, 0, 4); wide(e, 6, 0, 6)
    chain(o, 2, [(1, 2), (1, 4), (1, 6)], "carry")
    chain(e, 4, [(1, 3), (1, 5), (1, 7)], "none")
    chain(o, 4, [(2, 3), (2, 5), (2, 7)], "none")
    chain(e, 6, [(2, 4), (2, 6)], "carry")
    chain(o, 6, [(3, 4), (3, 6)], "carry")
    chain(e, 8, [(3, 5), (3, 7)], "none")
    chain(o, 8, [(4, 5), (4, 7)], "none")
    chain(e, 10, [(4, 6)], "carry")
    chain(o, 10, [(5, 6)], "carry")
    chain(e, 12, [(5, 7)], "none")
    chain(o, 12, [(6, 7)], "none")
    # U = E + O*B in place in e[2..14] (e[14] is fresh: = o13 + carry); u0 = 0, u1 = o0
    ops = []
    for k in range(2, 15):
        x = e[k] if e[k] in fresh else "0"
        ops.append(("add", e[k], x, o[k - 1], None, k > 2, k < 14))
    S.append(("chain", ops))
    u = ["0", o[0]] + e[2:15] + ["0"]
    # t' = 2U
    t = [f"t{k}" for k in range(16)]
    for k in range(16):
        lo = u[k - 1] if k else "0"
        hi = u[k]
        if hi == "0" and lo == "0":
            S.append(("cxx", f"uint32_t {t[k]} = 0u;", (t[k], lambda env, lo=lo, hi=hi: 0)))
        elif hi == "0":
            S.append(("cxx", f"uint32_t {t[k]} = {lo} >> 31;", (t[k], lambda env, lo=lo: env[lo] >> 31)))
        elif lo == "0":
            S.append(("cxx", f"uint32_t {t[k]} = {hi} << 1;", (t[k], lambda env, hi=hi: (env[hi] << 1) & M32)))
        else:
            S.append(("cxx", f"uint32_t {t[k]} = __funnelshift_l({lo}, {hi}, 1);",
                      (t[k], lambda env, lo=lo, hi=hi: ((env[hi] << 1) | (env[lo] >> 31)) & M32)))
    # T = t' + diagonal squares
    ops = []
    for k in range(8):
        ops.append(("mad_lo", t[2 * k], A[k], A[k], t[2 * k], k > 0, True))
        ops.append(("mad_hi", t[2 * k + 1], A[k], A[k], t[2 * k + 1], True, k < 7))
    S.append(("chain", ops))
    S.append(("check_square", t))
    # reduction-only rows on c = T_low
    c = t[:8]
    for i in range(8):
        m, mlo, mhi, z, c8 = f"m{i}", f"ml{i}", f"mh{i}", f"z{i}", f"c8_{i}"
        S.append(("cxx", f"const uint32_t {m} = 0u - {c[0]}, {mlo} = {m} << SH30, {mhi} = {m} >> SH2;", None))
        S.append(("model", lambda env, m=m, mlo=mlo, mhi=mhi, c0=c[0]: env.update({m: (-env[c0]) & M32, mlo: ((-env[c0]) << 30) & M32,
                                                                                    mhi: ((-env[c0]) & M32) >> 2})))
        ops = [("add", z, c[0], m, None, False, True),
               ("mad_lo", c[1], m, "P1", c[1], True, True), ("mad_hi", c[2], m, "P1", c[2], True, True),
               ("mad_lo", c[3], m, "P3", c[3], True, True), ("mad_hi", c[4], m, "P3", c[4], True, True),
               ("add", c[5], c[5], "0", None, True, True), ("add", c[6], c[6], "0", None, True, True),
               ("add", c[7], c[7], mlo, None, True, True), ("add", c8, mhi, "0", None, True, False)]
        S.append(("chain", ops))
        ops = [("mad_lo", c[2], m, "P2", c[2], False, True), ("mad_hi", c[3], m, "P2", c[3], True, True)]
        ops += [("add", c[k], c[k], "0", None, True, True) for k in range(4, 8)]
        ops += [("add", c8, c8, "0", None, True, False)]
        S.append(("chain", ops))
        c = c[1:] + [c8]
    ops = [("add", c[k], c[k], t[8 + k], None, k > 0, k < 7) for k in range(8)]
    S.append(("chain", ops))
    return S, c


def model(a_limbs, p):
    pl = [(p >> (32 * i)) & M32 for i in range(8)]
    env = {f"a{i}": a_limbs[i] for i in range(8)}
    env.update({"0": 0, "P1": pl[1], "P2": pl[2], "P3": pl[3]})
    S, res = program()
    for st in S:
        if st[0] == "wide":
            pr = env[st[3]] * env[st[4]]
            env[st[1]], env[st[2]] = pr & M32, pr >> 32
        elif st[0] == "cxx":
            if st[2] is not None:
                env[st[2][0]] = st[2][1](env)
        elif st[0] == "model":
            st[1](env)
        elif st[0] == "check_square":
            A = sum(a_limbs[i] << (32 * i) for i in range(8))
            assert sum(env[n] << (32 * k) for k, n in enumerate(st[1])) == A * A
        else:
            carry = 0
            for (kind, d, x, y, addend, cin, cout) in st[1]:
                if kind == "add":
                    v = env[x] + env[y]
                else:
                    pr = env[x] * env[y]
                    v = ((pr & M32) if kind == "mad_lo" else (pr >> 32)) + env[addend]
                v += carry if cin else 0
                if cout:
                    carry = v >> 32
                else:
                    assert v >> 32 == 0, ("carry lost", kind, d)
                    carry = 0
                env[d] = v & M32
    R = sum(env[n] << (32 * k) for k, n in enumerate(res))
    assert R < 2 * p
    return R - p if R >= p else R


def selftest(n=4000):
    rng = random.Random(7)
    for p in (P_PALLAS, P_VESTA):
        rinv = pow(1 << 256, -1, p)
        cases = [0, 1, 2, p - 1, p - 2, 0xffffffff, (1 << 254) - 1, 1 << 253, (1 << 254), p >> 1, (p >> 1) + 1,
                 int.from_bytes(bytes([0xff] * 31 + [0x3f]), "little"), sum(0xffffffff << (64 * i) for i in range(4)) & ((1 << 254) - 1)]
        cases += [rng.randrange(p) for _ in range(n)]
        cases += [rng.randrange(p) | (((1 << 224) - 1) << 16) & ((1 << 254) - 1) for _ in range(200)]
        for x in cases:
            x %= p
            got = model([(x >> (32 * i)) & M32 for i in range(8)], p)
            assert got == x * x * rinv % p, hex(x)
    return True


def emit():
    S, res = program()
    out = []
    w = out.append
    w("// fp_sqr.cuh -- GENERATED by tools/gen_fp_sqr.py (do not edit; the generator also holds the limb-level model that is checked")
    w("// against x*x*R^-1 mod p on the CPU).  Dedicated Montgomery squaring for the Pasta base fields: 28 off-diagonal + 8 diagonal")
    w("// + 24 reduction wide products instead of the 88 of fp_mul(a, a); every carry chain is one asm statement.")
    w("#pragma once")
    w("namespace vimz {")
    w("template <class F>")
    w("__device__ __forceinline__ void fp_sqr_pasta_limbs(uint32_t (&r)[8], const uint32_t (&a)[8]) {")
    w("  const uint32_t " + ", ".join(f"a{i} = a[{i}]" for i in range(8)) + ";")
    w("  const uint32_t P1 = F::p(1), P2 = F::p(2), P3 = F::p(3);")
    w("#if VIMZ_OPAQUE_P  // shift counts ptxas cannot see: (m << 30, m >> 2) stays two shifts instead of a wide multiply by 2^30 (fp.cuh)")
    w("  const uint32_t SH30 = red_const<F>[3], SH2 = red_const<F>[4];")
    w("#else")
    w("  const uint32_t SH30 = 30u, SH2 = 2u;")
    w("#endif")
    declared = set(f"a{i}" for i in range(8)) | {"P1", "P2", "P3", "0", "SH30", "SH2"}
    for st in S:
        if st[0] == "wide":
            _, lo, hi, x, y = st
            w(f"  uint32_t {lo}, {hi};")
            w(f"  {{ const uint64_t pr = (uint64_t){x} * {y}; {lo} = (uint32_t)pr; {hi} = (uint32_t)(pr >> 32); }}")
            declared.update((lo, hi))
        elif st[0] == "cxx":
            w("  " + st[1])
            if st[2] is not None:
                declared.add(st[2][0])
            else:  # the m / mlo / mhi line
                for tok in st[1].replace(",", " ").replace("=", " ").split():
                    if tok[0] in "m" and tok[1:].lstrip("lh").isdigit():
                        declared.add(tok)
        elif st[0] in ("model", "check_square"):
            continue
        else:
            ops = st[1]
            # operand table: read-write (in place), write-only, read-only
            dsts = [op[1] for op in ops]
            rw, wo = [], []
            seen_written = set()
            for (kind, d, x, y, addend, cin, cout) in ops:
                srcs = [x, y] + ([addend] if addend is not None else [])
                for s_ in srcs:
                    if s_ in dsts and s_ not in seen_written and s_ not in rw and s_ != "0":
                        rw.append(s_)       # read before (or when) written in this chain: must be an in/out operand
                seen_written.add(d)
            for d in dsts:
                if d not in rw and d not in wo:
                    wo.append(d)
            ro = []
            for (kind, d, x, y, addend, cin, cout) in ops:
                for s_ in [x, y] + ([addend] if addend is not None else []):
                    if s_ != "0" and s_ not in rw and s_ not in wo and s_ not in ro:
                        ro.append(s_)
            # a write-only destination that is ALSO read later in the same chain is fine (it is read after its write) -- but the
            # compiler may give a write-only operand the register of an input that is still needed: use early-clobber
            order = rw + wo + ro
            num = {n: i for i, n in enumerate(order)}
            def opnd(n):
                return "0" if n == "0" else f"%{num[n]}"
            lines = []
            for (kind, d, x, y, addend, cin, cout) in ops:
                if kind == "add":
                    mn = {(False, True): "add.cc", (True, True): "addc.cc", (True, False): "addc", (False, False): "add"}[(cin, cout)]
                    lines.append(f"{mn}.u32 {opnd(d)}, {opnd(x)}, {opnd(y)};")
                else:
                    half = "lo" if kind == "mad_lo" else "hi"
                    mn = {(False, True): f"mad.{half}.cc", (True, True): f"madc.{half}.cc", (True, False): f"madc.{half}", (False, False): f"mad.{half}"}[(cin, cout)]
                    lines.append(f"{mn}.u32 {opnd(d)}, {opnd(x)}, {opnd(y)}, {opnd(addend)};")
            for d in wo:
                if d not in declared:
                    w(f"  uint32_t {d};")
                    declared.add(d)
            for d in rw:
                assert d in declared, d
            assert len(order) <= 30, len(order)
            body = "\n      ".join('"' + l + ('\\n\\t"' if i + 1 < len(lines) else '"') for i, l in enumerate(lines))
            outs = ", ".join([f'"+r"({n})' for n in rw] + [f'"=&r"({n})' for n in wo])
            ins = ", ".join(f'"r"({n})' for n in ro)
            w(f"  asm({body}\n      : {outs}\n      : {ins});")
    for k, n in enumerate(res):
        w(f"  r[{k}] = {n};")
    w("}")
    w("}  // namespace vimz")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    assert selftest(), "model"
    if "--emit" in sys.argv:
        sys.stdout.write(emit())
    else:
        print("model ok (Pallas, Vesta)")
