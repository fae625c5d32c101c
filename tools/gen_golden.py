#!/usr/bin/env python3
"""Extract the reference's own test vectors into tests/golden/*.json.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box).  It never copies reference *code*: it reads the numeric literals of the reference's
inline #[test] functions (the known-answer vectors SURVEY.md section 8c lists) and, for the slot
isomorphisms that no KAT pins, *interprets* the straight-line assignment statements of
`nonresidue_to_*` / `homogenize_*` in the reference source on random vectors, recording
input -> output pairs.  The oracle (oracle/ref_py.py) is then checked against these files by
tests/test_oracle.py.

Usage: python tools/gen_golden.py [--ref /root/reference] [--out tests/golden]
"""
import argparse
import json
import os
import random
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MODELS_DIR = "crates/ring/src/cyclotomic_ring/models"


def fn_body(src, name):
    """Body text of `fn name(...) { ... }` (brace matched)."""
    m = re.search(r"fn\s+" + re.escape(name) + r"\s*\([^)]*\)[^{]*\{", src)
    if not m:
        raise KeyError(name)
    i = m.end()
    depth = 1
    while depth:
        ch = src[i]
        depth += ch == "{"
        depth -= ch == "}"
        i += 1
    return src[m.end():i - 1]


ITEM = re.compile(r'MontFp!\("(\d+)"\)|Fq::from\((\d+)\)|Fq::(zero)\(\)|Fq::(one)\(\)')


def vec_literal(body, var):
    """Values of `let [mut] var: Vec<Fq> = vec![ ... ];`"""
    m = re.search(r"let\s+(?:mut\s+)?" + var + r"\s*:\s*Vec<Fq>\s*=\s*vec!\[(.*?)\];", body, re.S)
    if not m:
        raise KeyError(var)
    out = []
    for a, b, z, o in ITEM.findall(m.group(1)):
        out.append(int(a) if a else int(b) if b else 0 if z else 1)
    return out


def indexed_literal(body, var, n):
    out = [0] * n
    for i, v in re.findall(var + r'\[(\d+)\]\s*=\s*MontFp!\("(\d+)"\);', body):
        out[int(i)] = int(v)
    return out


def const_table(src, name):
    m = re.search(r"const\s+" + name + r"\s*:\s*&\[Fq\]\s*=\s*&\[(.*?)\];", src, re.S)
    body = m.group(1)
    vals = re.findall(r"BigInt\(\[(\d+)u64\]\)", body)
    if not vals:
        vals = re.findall(r'MontFp!\("(\d+)"\)', body)
    return [int(v) for v in vals]


def const_scalar(src, name):
    m = re.search(r"const\s+" + name + r"\s*:\s*Fq\s*=\s*(.*?);", src, re.S)
    v = re.search(r"(\d{3,})", m.group(1))
    return int(v.group(1))


# ---- mini interpreter for the straight-line slot maps --------------------------------
class SlotInterp:
    def __init__(self, src, p, roots, table_name, swaps):
        self.src, self.p, self.W, self.table, self.swaps = src, p, roots, table_name, swaps

    def term(self, t, c, env):
        t = t.strip()
        neg = t.startswith("-")
        if neg:
            t = t[1:].strip()
        parts = [x.strip() for x in t.split("*")]
        val = 1
        for q in parts:
            m = re.fullmatch(r"c\[(\d+)\]", q)
            if m:
                val = val * c[int(m.group(1))] % self.p
                continue
            m = re.fullmatch(self.table + r"\[(\d+)\]", q)
            if m:
                val = val * self.W[int(m.group(1))] % self.p
                continue
            if q in env:
                val = val * env[q] % self.p
                continue
            raise ValueError("cannot interpret term: " + q)
        return (-val) % self.p if neg else val

    def run_fn(self, name, c):
        body = fn_body(self.src, name)
        env = {}
        for stmt in body.split(";"):
            s = re.sub(r"//.*", "", stmt).strip()
            if not s:
                continue
            m = re.fullmatch(r"let\s+(\w+)\s*=\s*(.+)", s, re.S)
            if m:
                env[m.group(1)] = self.term(m.group(2), c, env)
                continue
            m = re.fullmatch(r"c\[(\d+)\]\s*\*=\s*(.+)", s, re.S)
            if m:
                i = int(m.group(1))
                c[i] = c[i] * self.term(m.group(2), c, env) % self.p
                continue
            m = re.fullmatch(r"c\[(\d+)\]\s*=\s*(.+)", s, re.S)
            if m:
                c[int(m.group(1))] = self.term(m.group(2), c, env)
                continue
            if re.fullmatch(r"permute_to_fq9_of_fq3\(c\)", s):
                for i, j in self.swaps:
                    c[i], c[j] = c[j], c[i]
                continue
            raise ValueError("cannot interpret statement in %s: %r" % (name, s))
        return c

    def run_dispatch(self, name, c):
        """homogenize_* / dehomogenize_*: a list of calls on sub-slices."""
        body = fn_body(self.src, name)
        c = list(c)
        for fn, lo, hi in re.findall(r"(\w+)\(&mut c\[(\d+)\.\.(\d+)\]\);", body):
            lo, hi = int(lo), int(hi)
            sub = c[lo:hi]
            if fn == "permute_to_fq9_of_fq3":
                for i, j in self.swaps:
                    sub[i], sub[j] = sub[j], sub[i]
            else:
                sub = self.run_fn(fn, sub)
            c[lo:hi] = sub
        return c


def gen_phi3(ref, sub, D, homog_fn, dehomog_fn, swaps, rng):
    src = open(os.path.join(ref, MODELS_DIR, sub, "ntt.rs")).read()
    modsrc = open(os.path.join(ref, MODELS_DIR, sub, "mod.rs")).read()
    p = int(re.search(r'#\[modulus = "(\d+)"\]', modsrc).group(1))
    nonres = int(re.search(r'NONRESIDUE: Self::Fp = MontFp!\("(\d+)"\)', modsrc).group(1))
    roots = const_table(src, "ROOTS_OF_UNITY_24")
    g = {
        "model": sub, "source": "%s/%s/ntt.rs" % (MODELS_DIR, sub), "p": str(p), "D": D,
        "nonresidue": str(nonres),
        "roots": [str(x) for x in roots],
        "KAPPA": str(const_scalar(src, "KAPPA")),
        "EIGHT_INV": str(const_scalar(src, "EIGHT_INV")),
        "FOUR_INV": str(const_scalar(src, "FOUR_INV")),
        "crt_kats": [], "homogenize": [], "dehomogenize": [],
    }
    if sub == "goldilocks":
        for t in ("test_crt", "test_crt2"):
            body = fn_body(src, t)
            inp = vec_literal(body, "test_poly")
            inp += [0] * (D - len(inp))
            g["crt_kats"].append({"test": t, "coeffs": [str(x) for x in inp],
                                  "slot_remainders": [str(x) for x in vec_literal(body, "expected")]})
        # test_icrt / test_icrt_2 hold the same vectors in the inverse direction; check that
        for t, k in (("test_icrt", 0), ("test_icrt_2", 1)):
            body = fn_body(src, t)
            ev = vec_literal(body, "evaluations")
            ex = vec_literal(body, "expected")
            ex += [0] * (D - len(ex))
            assert [str(x) for x in ev] == g["crt_kats"][k]["slot_remainders"]
            assert [str(x) for x in ex] == g["crt_kats"][k]["coeffs"]
    else:
        body = fn_body(src, "test_babybear_icrt_hardcoded")
        g["crt_kats"].append({"test": "test_babybear_icrt_hardcoded",
                              "coeffs": [str(x) for x in indexed_literal(body, "expected", D)],
                              "slot_remainders": [str(x) for x in indexed_literal(body, "initial_ntt", D)]})
    it = SlotInterp(src, p, roots, "ROOTS_OF_UNITY_24", swaps)
    for _ in range(8):
        x = [rng.randrange(p) for _ in range(D)]
        g["homogenize"].append({"in": [str(v) for v in x],
                                "out": [str(v) for v in it.run_dispatch(homog_fn, x)]})
        g["dehomogenize"].append({"in": [str(v) for v in x],
                                  "out": [str(v) for v in it.run_dispatch(dehomog_fn, x)]})
    # unit vectors pin every entry of the (monomial) maps
    for i in range(D):
        x = [0] * D
        x[i] = 1
        g["homogenize"].append({"in": [str(v) for v in x],
                                "out": [str(v) for v in it.run_dispatch(homog_fn, x)]})
    return g


def gen_stark(ref):
    sub = "stark_prime"
    src = open(os.path.join(ref, MODELS_DIR, sub, "ntt.rs")).read()
    modsrc = open(os.path.join(ref, MODELS_DIR, sub, "mod.rs")).read()
    p = int(re.search(r'#\[modulus = "(\d+)"\]', modsrc).group(1))
    g = {"model": sub, "source": "%s/%s/ntt.rs" % (MODELS_DIR, sub), "p": str(p), "D": 16,
         "roots": [str(x) for x in const_table(src, "ROOTS_OF_UNITY_32")],
         "SIXTEEN_INV": str(const_scalar(src, "SIXTEEN_INV")),
         "SIXTEEN_INV_TIMES_ROOT_OF_UNITY_32_24": str(const_scalar(src, "SIXTEEN_INV_TIMES_ROOT_OF_UNITY_32_24")),
         "crt_kats": []}
    for t in ("test_crt", "test_crt2"):
        body = fn_body(src, t)
        inp = vec_literal(body, "test_poly")
        inp += [0] * (16 - len(inp))
        g["crt_kats"].append({"test": t, "coeffs": [str(x) for x in inp],
                              "evaluations": [str(x) for x in vec_literal(body, "expected")]})
    for t, k in (("test_icrt", 0), ("test_icrt_2", 1)):
        body = fn_body(src, t)
        ev = vec_literal(body, "evaluations")
        ex = vec_literal(body, "expected")
        ex += [0] * (16 - len(ex))
        assert [str(x) for x in ev] == g["crt_kats"][k]["evaluations"]
        assert [str(x) for x in ex] == g["crt_kats"][k]["coeffs"]
    return g


def gen_derived(out_dir, rng):
    """Vectors computed by the (pinned) oracle: full CRT layout, ring mul, raw Montgomery limbs,
    mat-vec.  Marked 'derived': they pin the C restatement and the CUDA path to the oracle,
    not the oracle to the reference."""
    from oracle import ref_py as O
    for key in ("goldilocks", "babybear", "stark_prime"):
        M = O.MODELS[key]
        cases = []
        for _ in range(4):
            a = [rng.randrange(M.p) for _ in range(M.D)]
            b = [rng.randrange(M.p) for _ in range(M.D)]
            ca, cb = M.crt(a), M.crt(b)
            prod = O.ring_mul(M, a, b)
            assert prod == O.poly_mul(M, a, b)
            cases.append({
                "a": [str(x) for x in a], "b": [str(x) for x in b],
                "crt_a": [str(x) for x in ca], "crt_b": [str(x) for x in cb],
                "ntt_mul": [str(x) for x in M.ntt_mul(ca, cb)],
                "ring_mul": [str(x) for x in prod],
                "a_raw": [str(x) for x in O.to_raw(M, a)],
                "crt_a_raw": [str(x) for x in O.to_raw(M, ca)],
                "ring_mul_raw": [str(x) for x in O.to_raw(M, prod)],
            })
        kappa, m = 3, 5
        rows = [[M.crt([rng.randrange(M.p) for _ in range(M.D)]) for _ in range(m)] for _ in range(kappa)]
        v = [M.crt([rng.randrange(M.p) for _ in range(M.D)]) for _ in range(m)]
        y = O.matvec(M, rows, v)
        mv = {"kappa": kappa, "m": m,
              "rows_raw": [[[str(x) for x in O.to_raw(M, e)] for e in row] for row in rows],
              "v_raw": [[str(x) for x in O.to_raw(M, e)] for e in v],
              "y_raw": [[str(x) for x in O.to_raw(M, e)] for e in y]}
        json.dump({"model": key, "kind": "derived-from-oracle", "cases": cases, "matvec": mv},
                  open(os.path.join(out_dir, key + "_derived.json"), "w"), indent=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    rng = random.Random(0x5EED)
    g = gen_phi3(args.ref, "goldilocks", 24, "homogenize_fq3", "dehomogenize_fq3", [], rng)
    json.dump(g, open(os.path.join(args.out, "goldilocks.json"), "w"), indent=0)
    bb_src = open(os.path.join(args.ref, MODELS_DIR, "babybear", "ntt.rs")).read()
    swaps_txt = re.search(r"const SWAPS[^=]*=\s*\[(.*?)\];", bb_src, re.S).group(1)
    swaps = [(int(i), int(j)) for i, j in re.findall(r"\((\d+),\s*(\d+)\)", swaps_txt)]
    g = gen_phi3(args.ref, "babybear", 72, "homogenize_fq9", "dehomogenize_fq9", swaps, rng)
    g["swaps"] = swaps
    json.dump(g, open(os.path.join(args.out, "babybear.json"), "w"), indent=0)
    json.dump(gen_stark(args.ref), open(os.path.join(args.out, "stark_prime.json"), "w"), indent=0)
    gen_derived(args.out, rng)
    print("golden vectors written to", args.out)


if __name__ == "__main__":
    main()
