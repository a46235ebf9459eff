#!/usr/bin/env python
"""Generates canon_eval.inc: the 27 regular monotone boolean functions of 4 ordered variables as
minimal LOP3 sequences behind one PTX jump table (brx.idx).

A function is *regular* in (z0, z1, z2, z3) when it is monotone non-decreasing in every variable
and z_i dominates z_j for i < j (moving a 1 from z_j to z_i never turns the function off).  Every
threshold function sum_k a_k (2 z_k - 1) > theta with a_0 >= a_1 >= a_2 >= a_3 >= 0 is regular, and
there are exactly 27 regular functions of 4 variables.  Truth table H: bit p, p = z0|z1<<1|z2<<2|z3<<3.

For each function the generator searches the cheapest form
    1 LOP3   f depends on at most 3 variables
    2 LOP3   f = A(B(x, y, w), u, v)
    3 LOP3   f = f|z3=0  |  z3 & f|z3=1           (Shannon; always possible)
lop3.b32 d, a, b, c, T computes bit (T >> (a<<2 | b<<1 | c)) & 1 lane by lane.

Usage: python gen_canon.py > canon_eval.inc      (the output is committed; rerun only on change)
"""
import itertools


def monotone(h):
    return all(not ((h >> p) & 1) or (h >> (p | 1 << n)) & 1 for p in range(16) for n in range(4))


def regular(h):
    for p in range(16):
        for i in range(4):
            for j in range(i + 1, 4):
                if not (p >> i) & 1 and (p >> j) & 1:
                    q = (p | 1 << i) & ~(1 << j)
                    if (h >> p) & 1 and not (h >> q) & 1:
                        return False
    return True


CANON = [h for h in range(1 << 16) if monotone(h) and regular(h)]
assert len(CANON) == 27


def bit(p, k):
    return (p >> k) & 1


def support(h):
    return [k for k in range(4) if any(((h >> p) & 1) != ((h >> (p ^ (1 << k))) & 1) for p in range(16))]


def lut_of(h, vars3):
    """LOP3 table of f as a function of (a, b, c) = vars3 (f must not depend on the 4th variable)."""
    t = 0
    for idx in range(8):
        p = 0
        for pos, k in enumerate(vars3):
            if (idx >> (2 - pos)) & 1:
                p |= 1 << k
        if (h >> p) & 1:
            t |= 1 << idx
    return t


def plan(h):
    """list of (dst, a, b, c, T); operands are 'z0'..'z3', 't0', 't1'; the last dst is 'f'."""
    if h == 0:
        return [("const", 0)]
    if h == 0xFFFF:
        return [("const", 0xFFFFFFFF)]
    sup = support(h)
    if len(sup) <= 3:
        v = (sup + [k for k in range(4) if k not in sup])[:3]
        return [("f", "z%d" % v[0], "z%d" % v[1], "z%d" % v[2], lut_of(h, v))]
    # two LOP3: f = A(B(x,y,w), u, v)
    for trip in itertools.combinations(range(4), 3):
        for B in range(1, 255):
            for u, v in itertools.combinations(range(4), 2):
                table = {}
                ok = True
                for p in range(16):
                    bidx = bit(p, trip[0]) << 2 | bit(p, trip[1]) << 1 | bit(p, trip[2])
                    key = ((B >> bidx) & 1) << 2 | bit(p, u) << 1 | bit(p, v)
                    val = (h >> p) & 1
                    if table.setdefault(key, val) != val:
                        ok = False
                        break
                if ok:
                    A = sum(val << key for key, val in table.items())
                    return [("t0", "z%d" % trip[0], "z%d" % trip[1], "z%d" % trip[2], B),
                            ("f", "t0", "z%d" % u, "z%d" % v, A)]
    t0, t1 = h & 0xFF, h >> 8
    return [("t0", "z2", "z1", "z0", t0), ("t1", "z2", "z1", "z0", t1), ("f", "t0", "z3", "t1", 0xF8)]


def check(h, steps):
    """simulate the plan on all 16 patterns"""
    for p in range(16):
        env = {"z%d" % k: bit(p, k) for k in range(4)}
        if steps[0][0] == "const":
            val = steps[0][1] & 1
        else:
            for d, a, b, c, T in steps:
                env[d] = (T >> (env[a] << 2 | env[b] << 1 | env[c])) & 1
            val = env["f"]
        assert val == (h >> p) & 1, (hex(h), steps)


def emit_asm(w, macro, ngroups):
    """one jump table evaluating the function for `ngroups` independent 32-lane groups: outputs
    %0..%(g-1), function id %g, then z0..z3 of group j at %(g+1+4j)..%(g+4+4j)"""
    g = ngroups
    w("// operands: %%0..%%%d out (one per 32-lane group); %%%d fid; z0..z3 of group j at %%(%d+4j)..%%(%d+4j)"
      % (g - 1, g, g + 1, g + 4))
    w("#define %s \\" % macro)
    w('    "{\\n\\t" \\')
    w('    ".reg .b32 t0, t1;\\n\\t" \\')
    w('    "ts: .branchtargets ' + ", ".join("C%d" % k for k in range(27)) + ';\\n\\t" \\')
    w('    "brx.idx %%%d, ts;\\n\\t" \\' % g)
    nl = {}
    for k, h in enumerate(CANON):
        steps = plan(h)
        nl[k] = 0 if steps[0][0] == "const" else len(steps)
        w('    "C%d:\\n\\t" \\' % k)
        for grp in range(g):
            dst, base = "%%%d" % grp, g + 1 + 4 * grp

            def opnd(s):
                if s[0] == "z":
                    return "%%%d" % (base + int(s[1]))
                return {"t0": "t0", "t1": "t1", "f": dst}[s]
            for st in steps:
                if st[0] == "const":
                    w('    "mov.b32 %s, 0x%08x;\\n\\t" \\' % (dst, st[1]))
                else:
                    d, a, b, c, T = st
                    w('    "lop3.b32 %s, %s, %s, %s, 0x%02x;\\n\\t" \\' % (opnd(d), opnd(a), opnd(b), opnd(c), T))
        w('    "bra.uni DONE;\\n\\t" \\')
    w('    "DONE:\\n\\t" \\')
    w('    "}"')
    return nl


def emit():
    for h in CANON:
        check(h, plan(h))
    out = []
    w = out.append
    w("// GENERATED by gen_canon.py -- do not edit.  The 27 regular functions of 4 ordered variables,")
    w("// one PTX jump table (brx.idx) per macro, minimal LOP3 sequences per 32-lane group.")
    w("#define PIQMC_NCANON 27")
    w("#define PIQMC_CANON_TABLES " + ", ".join("0x%04xu" % h for h in CANON))
    w("// one 64-lane word: groups = (low half, high half)")
    nl = emit_asm(w, "PIQMC_CANON_EVAL_ASM", 2)
    w("// two 64-lane words (two rows per thread): groups = (row 0 low, row 0 high, row 1 low, row 1 high)")
    emit_asm(w, "PIQMC_CANON_EVAL_ASM2", 4)
    w("// LOP3 per 32 lanes: " + ", ".join("%04x:%d" % (h, nl[k]) for k, h in enumerate(CANON)))
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    import sys
    sys.stdout.write(emit())
