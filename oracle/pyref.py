"""ORACLE (Python twin) — TEST INFRASTRUCTURE ONLY.

A second, independent restatement of the reference verifier in plain Python (big ints, lists), written from the
Haskell sources and reading the JSON wire format directly with the `json` module — it shares no code with the C++
oracle (oracle/*.hpp) nor with the product's parser.  tests/test_pyref.py runs both restatements on the bundled
fixtures and on tampered copies and requires identical challenges, combined constraints and verdicts.

Cited files are under /root/reference/src/.  Pure-Python speed (~600 permutations/s) limits it to the small shapes.
"""
import json
import os
import re

P = 0xFFFFFFFF00000001          # Algebra/Goldilocks.hs:125-126
MUL_GEN = 0xc65c18b67785d900    # :135
TWO_ADIC = 0x64fdd1a46201e246   # :54-55


# ---- Algebra/Goldilocks.hs, GoldilocksExt.hs, FFT.hs, Poly.hs ------------------------------------------------
def inv(x):
    return pow(x, P - 2, P)     # inv 0 = 0 (:155)


def fpow(x, e):                 # negative exponents invert first (:169)
    return pow(inv(x), -e, P) if e < 0 else pow(x, e, P)


def subgroup_generator(k):      # rootsOfUnity ! k (:68-74)
    x = TWO_ADIC
    for _ in range(32 - k):
        x = x * x % P
    return x


class E:
    """F[X]/(X^2 - 7), GoldilocksExt.hs:28-83."""
    __slots__ = ("a", "b")

    def __init__(self, a, b=0):
        self.a, self.b = a % P, b % P

    def __add__(s, o): return E(s.a + o.a, s.b + o.b)
    def __sub__(s, o): return E(s.a - o.a, s.b - o.b)
    def __mul__(s, o): return E(s.a * o.a + 7 * s.b * o.b, s.a * o.b + o.a * s.b)
    def __eq__(s, o): return s.a == o.a and s.b == o.b
    def scale(s, k): return E(k * s.a, k * s.b)

    def inv(s):
        d = inv((s.a * s.a - 7 * s.b * s.b) % P)
        return E(s.a * d, -s.b * d)

    def __truediv__(s, o): return s * o.inv()

    def pow(s, e):
        if e < 0:
            return s.inv().pow(-e)
        acc, sq = E(1), s
        while e:
            if e & 1:
                acc = acc * sq
            sq = sq * sq
            e >>= 1
        return acc

    def pair(s): return (s.a, s.b)


E0, E1 = E(0), E(1)


def reverse_bits(n, w):         # FFT.hs:20-25
    return int(format(w, "0%db" % n)[::-1], 2) if n else 0


def reduce_with_powers(alpha, xs):  # Goldilocks.hs:180-183
    acc = E0
    for x in reversed(xs):
        acc = x + alpha * acc
    return acc


# ---- Hash/Constants.hs (tables parsed from the generated data header), Poseidon.hs, Sponge.hs, Merkle.hs ------
def _tables():
    txt = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "poseidon_constants.h")).read()
    out = {}
    for m in re.finditer(r"#define (\w+) \{(.*?)\}", txt, re.S):
        out[m.group(1)] = [int(x[:-3], 16) for x in re.findall(r"0x[0-9a-f]+ULL", m.group(2))]
    return out


T = _tables()
CIRC, DIAG, RC = T["P2V_MDS_CIRC"], T["P2V_MDS_DIAG"], T["P2V_ALL_ROUND_CONSTANTS"]
FIRST_RC, PARTIAL_RC = T["P2V_FAST_PARTIAL_FIRST_RC"], T["P2V_FAST_PARTIAL_RCS"]
VS, W_HATS, INIT_MAT = T["P2V_FAST_PARTIAL_VS"], T["P2V_FAST_PARTIAL_W_HATS"], T["P2V_FAST_PARTIAL_INIT_MAT"]


def mds_coeff(i, j):            # Constants.hs:24-25
    return CIRC[(j - i) % 12] + (DIAG[i] if i == j else 0)


MDS = [[mds_coeff(i, j) for j in range(12)] for i in range(12)]
PERM_COUNT = [0]


def permutation(s):             # Poseidon.hs:42-101
    PERM_COUNT[0] += 1
    s = list(s)
    for r in range(30):
        rc = RC[12 * r: 12 * r + 12]
        if r < 4 or r >= 26:
            s = [pow(x + c, 7, P) for x, c in zip(s, rc)]
        else:
            s = [pow(s[0] + rc[0], 7, P)] + [(x + c) % P for x, c in zip(s[1:], rc[1:])]
        s = [sum(MDS[i][j] * s[j] for j in range(12)) % P for i in range(12)]
    return s


def sponge(xs):                 # Sponge.hs:26-31
    st = [0] * 12
    for k in range(0, len(xs), 8):
        chunk = xs[k:k + 8]
        st = permutation(chunk + st[len(chunk):])
    return st[:4]


def compress(x, y):             # Merkle.hs:21-23
    return permutation(list(x) + list(y) + [0, 0, 0, 0])[:4]


def check_merkle_proof(cap, idx, leaf, siblings):  # Merkle.hs:27-42
    node = sponge(leaf)
    for sib in siblings:
        node = compress(node, sib) if idx % 2 == 0 else compress(sib, node)
        idx >>= 1
    return cap[idx] == node     # IndexError = the (!!) exception


# ---- Challenge/Pure.hs -------------------------------------------------------------------------------------------
class Duplex:
    def __init__(self):
        self.state, self.absorbing, self.buf = [0] * 12, True, []

    def absorb(self, x):        # absorbFelt :50-58
        if not self.absorbing:
            self.absorbing, self.buf = True, []
        if len(self.buf) == 8:
            self.state = permutation(self.buf + self.state[8:])
            self.buf = []
        self.buf = self.buf + [x % P]

    def absorb_all(self, xs):
        for x in xs:
            self.absorb(x)

    def squeeze(self):          # squeezeFelt :60-69
        if self.absorbing:
            self.state = permutation(self.buf + self.state[len(self.buf):])
            self.absorbing, self.buf = False, list(reversed(self.state[:8]))
        elif not self.buf:
            self.state = permutation(self.state)
            self.buf = list(reversed(self.state[:8]))
        return self.buf.pop(0)

    def squeeze_ext(self):
        a = self.squeeze()
        return E(a, self.squeeze())


# ---- JSON (Types.hs) ---------------------------------------------------------------------------------------------
def dg(d): return [x % P for x in d["elements"]]
def ex(xs): return [E(a, b) for a, b in xs]


def parse_gate(s):              # Gate/Parser.hs:107-240 (only the shapes the fixtures contain)
    m = re.match(r"^(\w+)", s)
    name = m.group(1)
    nums = lambda key: int(re.search(key + r": (\d+)", s).group(1))
    if name in ("ArithmeticGate", "ArithmeticExtensionGate", "MulExtensionGate"): return (name, nums("num_ops"))
    if name == "BaseSumGate": return (name, nums("num_limbs"), nums("Base"))
    if name == "ConstantGate": return (name, nums("num_consts"))
    if name == "ExponentiationGate": return (name, nums("num_power_bits"))
    if name in ("ReducingGate", "ReducingExtensionGate"): return (name, nums("num_coeffs"))
    if name == "RandomAccessGate": return (name, nums("bits"), nums("num_copies"), nums("num_extra_constants"))
    if name == "CosetInterpolationGate":
        w = [int(x) % P for x in re.search(r"barycentric_weights: \[([0-9, ]*)\]", s).group(1).split(",")]
        return (name, nums("subgroup_bits"), nums("degree"), w)
    if name in ("PoseidonGate", "PoseidonMdsGate"): return (name, int(re.search(r"WIDTH=(\d+)", s).group(1)))
    if name in ("NoopGate", "PublicInputGate", "LookupGate", "LookupTableGate"): return (name,)
    return ("UnknownGate", s)


def expand_strategy(degree_bits, strat):   # Plonk/FRI.hs:337-354
    if "Fixed" in strat:
        return list(strat["Fixed"])
    a, f = strat["ConstantArityBits"]
    out, logn = [], degree_bits
    while logn > f:
        out.append(a)
        logn -= a
    return out


# ---- Gates (Gate/Constraints.hs, Gate/Custom/*.hs); ext-of-ext values are pairs (E, E) ---------------------------
def ee_mul(x, y): return (x[0] * y[0] + E(7) * x[1] * y[1], x[0] * y[1] + y[0] * x[1])
def ee_add(x, y): return (x[0] + y[0], x[1] + y[1])
def ee_sub(x, y): return (x[0] - y[0], x[1] - y[1])
def ee_scale(s, x): return (s * x[0], s * x[1])


def gate_constraints(g, w, c, pih):
    out = []
    wE = lambda i: (w[i], w[i + 1])
    sb = lambda x: (x * x * x) * ((x * x) * (x * x))
    name = g[0]
    if name == "ArithmeticGate":
        for i in range(g[1]):
            j = 4 * i
            out.append(w[j + 3] - c[0] * w[j] * w[j + 1] - c[1] * w[j + 2])
    elif name == "ArithmeticExtensionGate":
        for i in range(g[1]):
            j = 8 * i
            t = ee_sub(ee_sub(wE(j + 6), ee_mul(ee_scale(c[0], wE(j)), wE(j + 2))), ee_scale(c[1], wE(j + 4)))
            out += [t[0], t[1]]
    elif name == "MulExtensionGate":
        for i in range(g[1]):
            j = 6 * i
            t = ee_sub(wE(j + 4), ee_mul(ee_scale(c[0], wE(j)), wE(j + 2)))
            out += [t[0], t[1]]
    elif name == "BaseSumGate":
        L, B = g[1], g[2]
        h = w[L]
        for k in range(L - 2, -1, -1):
            h = w[k + 1] + E(B) * h
        out.append(h - w[0])
        for i in range(L):
            pr = E1
            for k in range(B):
                pr = pr * (w[i + 1] - E(k))
            out.append(pr)
    elif name == "ConstantGate":
        out += [c[i] - w[i] for i in range(g[1])]
    elif name == "PublicInputGate":
        out += [w[i] - E(pih[i]) for i in range(4)]
    elif name == "ExponentiationGate":
        n = g[1]
        for i in range(n):
            prev = E1 if i == 0 else w[n + 2 + i - 1] * w[n + 2 + i - 1]
            cur = w[1 + (n - 1 - i)]
            out.append(prev * (cur * w[0] + (E1 - cur)) - w[n + 2 + i])
        out.append(w[n + 1] - w[n + 2 + n - 1])
    elif name in ("ReducingGate", "ReducingExtensionGate"):
        n, ext = g[1], name == "ReducingExtensionGate"
        output, alpha, prev = wE(0), wE(2), wE(4)
        for i in range(n):
            acc = wE(6 + (2 * n if ext else n) + 2 * i) if i < n - 1 else output
            coeff = wE(6 + 2 * i) if ext else (w[6 + i], E0)
            t = ee_sub(ee_add(ee_mul(prev, alpha), coeff), acc)
            out += [t[0], t[1]]
            prev = acc
    elif name == "RandomAccessGate":
        nb, copies, extra = g[1], g[2], g[3]
        width = 2 + (1 << nb)
        start = width * copies + extra
        for k in range(copies):
            bits = [w[start + k * nb + j] for j in range(nb)]
            out += [b * (b - E1) for b in bits]
            rec = E0
            for b in reversed(bits):
                rec = E(2) * rec + b
            out.append(rec - w[k * width])
            vals = [w[k * width + 2 + i] for i in range(1 << nb)]
            for b in bits:
                vals = [vals[i] + b * (vals[i + 1] - vals[i]) for i in range(0, len(vals), 2)]
            out.append(vals[0] - w[k * width + 1])
        out += [c[j] - w[copies * width + j] for j in range(extra)]
    elif name == "CosetInterpolationGate":
        bits, degree, weights = g[1], g[2], g[3]
        npts = 1 << bits
        nint = (npts - 2) // (degree - 1)
        gen = subgroup_generator(bits)
        domain = [pow(gen, k, P) for k in range(npts)]
        base = 1 + 2 * (npts + 2)
        shifted = wE(base + 4 * nint)
        t = ee_sub(wE(1 + 2 * npts), ee_scale(w[0], shifted))
        out += [t[0], t[1]]
        chunks = [list(range(0, degree))] + [list(range(s, min(s + degree - 1, npts))) for s in range(degree, npts, degree - 1)]
        inits = [((E0, E0), (E1, E0))] + [(wE(base + 2 * i), wE(base + 2 * (nint + i))) for i in range(nint)]
        stuff = []
        for (ev, pr), chunk in zip(inits, chunks):
            for idx in chunk:
                val = ee_scale(E(weights[idx]), wE(1 + 2 * idx))
                term = (shifted[0] - E(domain[idx]), shifted[1])
                ev, pr = ee_add(ee_mul(term, ev), ee_mul(val, pr)), ee_mul(term, pr)
            stuff.append((ev, pr))
        for i, (ev, pr) in enumerate(stuff[:-1]):
            a, b = ee_sub(wE(base + 2 * i), ev), ee_sub(wE(base + 2 * (nint + i)), pr)
            out += [a[0], a[1], b[0], b[1]]
        t = ee_sub(wE(1 + 2 * npts + 2), stuff[-1][0])
        out += [t[0], t[1]]
    elif name == "PoseidonMdsGate":
        for i in range(12):
            acc = (E0, E0)
            for j in range(12):
                acc = ee_add(acc, ee_scale(E(MDS[i][j]), wE(2 * j)))
            t = ee_sub(wE(2 * (i + 12)), acc)
            out += [t[0], t[1]]
    elif name == "PoseidonGate":
        swap = w[24]
        out.append(swap * (swap - E1))
        out += [swap * (w[i + 4] - w[i]) - w[25 + i] for i in range(4)]
        st = [w[i] + w[25 + i] for i in range(4)] + [w[i] - w[25 + i - 4] for i in range(4, 8)] + [w[i] for i in range(8, 12)]
        mds = lambda s: [sum((s[j].scale(MDS[i][j]) for j in range(12)), E0) for i in range(12)]
        for r in range(4):
            st = [x + E(RC[12 * r + i]) for i, x in enumerate(st)]
            if r:
                sin = [w[29 + 12 * (r - 1) + i] for i in range(12)]
                out += [st[i] - sin[i] for i in range(12)]
                st = sin
            st = mds([sb(x) for x in st])
        st = [x + E(FIRST_RC[i]) for i, x in enumerate(st)]
        st = [st[0]] + [sum((st[j + 1].scale(INIT_MAT[11 * j + i]) for j in range(11)), E0) for i in range(11)]
        for r in range(22):
            sin = w[65 + r]
            out.append(st[0] - sin)
            y = sb(sin) + E(PARTIAL_RC[r] if r < 21 else 0)
            d = y.scale(MDS[0][0])
            for i in range(11):
                d = d + st[i + 1].scale(W_HATS[11 * r + i])
            st = [d] + [st[i + 1] + y.scale(VS[11 * r + i]) for i in range(11)]
        for r in range(4):
            st = [x + E(RC[12 * (26 + r) + i]) for i, x in enumerate(st)]
            sin = [w[87 + 12 * r + i] for i in range(12)]
            out += [st[i] - sin[i] for i in range(12)]
            st = mds([sb(x) for x in sin])
        out += [st[i] - w[12 + i] for i in range(12)]
    elif name in ("NoopGate", "LookupGate", "LookupTableGate"):
        pass
    else:
        raise ValueError("gateConstraints: unknown gate")
    return out


# ---- the verifier ------------------------------------------------------------------------------------------------------
def verify(common, vkey, proof, trace=None):
    """-> status word as in include/p2v.h (0 accept).  `trace` (dict) receives challenges / combined."""
    cfg, fc = common["config"], common["config"]["fri_config"]
    r, nwires, routed = cfg["num_challenges"], cfg["num_wires"], cfg["num_routed_wires"]
    nbits, rate_bits, cap_h = common["fri_params"]["degree_bits"], fc["rate_bits"], fc["cap_height"]
    lde_bits, N = nbits + rate_bits, 1 << nbits
    qdf, npp_c, nlp = common["quotient_degree_factor"], common["num_partial_products"], common["num_lookup_polys"]
    gates = [parse_gate(s) for s in common["gates"]]
    sel_idx, groups = common["selectors_info"]["selector_indices"], common["selectors_info"]["groups"]
    luts = [[(a % P, b % P) for a, b in t] for t in common["luts"]]
    pr, pis = proof["proof"], [x % P for x in proof["public_inputs"]]
    op = {k: ex(v) for k, v in pr["openings"].items()}
    fri = pr["opening_proof"]
    caps = {k: [dg(d) for d in pr[k]] for k in ("wires_cap", "plonk_zs_partial_products_cap", "quotient_polys_cap")}
    vk_cap, digest = [dg(d) for d in vkey["constants_sigmas_cap"]], dg(vkey["circuit_digest"])
    commit_caps = [[dg(d) for d in c] for c in fri["commit_phase_merkle_caps"]]
    final_poly = ex(fri["final_poly"]["coeffs"])
    # ---- proofChallenges, Challenge/Verifier.hs:58-103 + Challenge/FRI.hs:65-104
    pih = sponge(pis)
    dx = Duplex()
    dx.absorb_all(digest); dx.absorb_all(pih)
    for d in caps["wires_cap"]: dx.absorb_all(d)
    betas = [dx.squeeze() for _ in range(r)]
    gammas = [dx.squeeze() for _ in range(r)]
    deltas = []
    if nlp > 0:
        new = [dx.squeeze() for _ in range(2 * r)]
        flat = betas + gammas + new
        deltas = [flat[i:i + 4] for i in range(0, len(flat), 4)]
    for d in caps["plonk_zs_partial_products_cap"]: dx.absorb_all(d)
    alphas = [dx.squeeze() for _ in range(r)]
    for d in caps["quotient_polys_cap"]: dx.absorb_all(d)
    zeta = dx.squeeze_ext()
    batch_this = op["constants"] + op["plonk_sigmas"] + op["wires"] + op["plonk_zs"] + op["partial_products"] + op["quotient_polys"] + op["lookup_zs"]
    batch_next = op["plonk_zs_next"] + op["lookup_zs_next"]
    for e in batch_this + batch_next: dx.absorb_all(e.pair())
    fri_alpha = dx.squeeze_ext()
    fri_betas = []
    for cap in commit_caps:
        for d in cap: dx.absorb_all(d)
        fri_betas.append(dx.squeeze_ext())
    for e in final_poly: dx.absorb_all(e.pair())
    dx.absorb(fri["pow_witness"])
    pow_response = dx.squeeze()
    indices = [dx.squeeze() % (1 << lde_bits) for _ in range(fc["num_query_rounds"])]
    if trace is not None:
        trace["challenges"] = betas + gammas + alphas + [x for d in deltas for x in d] + list(zeta.pair()) + list(fri_alpha.pair()) + \
            [x for b in fri_betas for x in b.pair()] + [pow_response] + indices
    # ---- evalAllPlonkConstraints, Plonk/Vanishing.hs:60-111
    ngs, nls = len(groups), common["num_lookup_selectors"]
    consts = op["constants"]
    gsel, lsel, gconst = consts[:ngs], consts[ngs:ngs + nls], consts[ngs + nls:]
    assert len(gconst) == cfg["num_constants"]
    w = op["wires"]
    zeta_n = zeta.pow(N)
    l0 = E1 if zeta == E1 else (zeta_n - E1) / (E(N) * (zeta - E1))
    terms = [l0 * (z - E1) for z in op["plonk_zs"]]
    pp_chunks = [op["partial_products"][i:i + npp_c] for i in range(0, len(op["partial_products"]), npp_c)] if npp_c else []
    k_is = [k % P for k in common["k_is"]]
    for z, zn, b, g_, chunk in zip(op["plonk_zs"], op["plonk_zs_next"], betas, gammas, pp_chunks):
        num = [wi + zeta.scale(b * k % P) + E(g_) for k, wi in zip(k_is, w)]
        den = [wi + s.scale(b) + E(g_) for s, wi in zip(op["plonk_sigmas"], w)]
        cur = [z] + chunk + [zn]
        for t, (pv, nx) in enumerate(zip(cur, cur[1:])):
            nch, dch = num[t * qdf:(t + 1) * qdf], den[t * qdf:(t + 1) * qdf]
            if not nch or not dch:
                break
            pn = pd = E1
            for x in nch: pn = pn * x
            for x in dch: pd = pd * x
            terms.append(pv * pn - nx * pd)
    if luts:   # Plonk/Lookups.hs:45-132
        lu_slots, lut_slots, nsldc = routed // 2, routed // 3, nlp - 1
        lu_deg, lut_deg = qdf - 1, -(-lut_slots // nsldc)
        cols = list(zip(op["lookup_zs"], op["lookup_zs_next"]))
        for rd in range(len(deltas)):
            A, B, al, de = deltas[rd]
            rcols = cols[rd * nlp:(rd + 1) * nlp]
            (re_, re_next), sl = rcols[0], rcols[1:]
            sldc, sldc_next = [a for a, _ in sl], [b for _, b in sl]
            pairs2 = [w[i:i + 2] for i in range(0, len(w), 2)][:lu_slots]
            triples = [w[i:i + 3] for i in range(0, len(w), 3)][:lut_slots]
            lu = [p_[0] + p_[1].scale(A) for p_ in pairs2 if len(p_) == 2]
            lutA = [t_[0] + t_[1].scale(A) for t_ in triples if len(t_) == 3]
            lutB = [t_[0] + t_[1].scale(B) for t_ in triples if len(t_) == 3]
            mults = [w[3 * i + 2] for i in range(lut_slots)]
            terms += [lsel[3] * sldc[-1], lsel[2] * sldc[0], lsel[2] * re_]
            for k, tab in enumerate(luts):
                rows = -(-len(tab) // lut_slots)
                padded = (tab + [tab[0]] * (rows * lut_slots))[:rows * lut_slots]
                cur = 0
                for a, b in padded:
                    cur = (de * cur + a + B * b) % P
                terms.append(lsel[4 + k] * (re_ - E(cur)))
            cs = re_next
            for e in lutB:
                cs = cs.scale(de) + e
            terms.append(lsel[0] * (re_ - cs))
            seq = [sldc_next[-1]] + sldc
            ch = lambda xs, d: [xs[i:i + d] for i in range(0, len(xs), d)]
            for (pv, th), luc, lutc, mc in zip(zip(seq, seq[1:]), ch(lu, lu_deg), ch(lutA, lut_deg), ch(mults, lut_deg)):
                prod = lambda xs, skip=-1: eprod([E(al) - x for i, x in enumerate(xs) if i != skip])
                lu_sum = sum((prod(luc, i) for i in range(len(luc))), E0)
                lut_sum = sum((m * prod(lutc, i) for i, m in zip(range(len(lutc)), mc)), E0)
                terms.append(lsel[0] * (prod(lutc) * (th - pv) - lut_sum))
                terms.append(lsel[1] * (prod(luc) * (th - pv) + lu_sum))
    gate_cols = []
    for k, g in enumerate(gates):
        grp = sel_idx[k]
        x = gsel[grp]
        f = (E(0xFFFFFFFF) - x) if len(groups) > 1 else E1
        for j in range(groups[grp]["start"], groups[grp]["end"]):
            if j != k:
                f = f * (E(j) - x)
        for i, cns in enumerate(gate_constraints(g, w, gconst, pih)):
            while len(gate_cols) <= i:
                gate_cols.append(E0)
            gate_cols[i] = gate_cols[i] + f * cns
    terms += gate_cols
    combined = []
    for a in alphas:
        acc = E0
        for t in reversed(terms):
            acc = t + acc.scale(a)
        combined.append(acc)
    if trace is not None:
        trace["combined"] = [x for c in combined for x in c.pair()]
        trace["pih"] = list(pih)
    # ---- checkCombinedPlonkEquations', Plonk/Verifier.hs:35-52
    mask = 0
    q = op["quotient_polys"]
    for i, c in enumerate(combined):
        acc = E0
        for x in reversed(q[i * qdf:(i + 1) * qdf]):
            acc = x + zeta_n * acc
        if not (acc * (zeta_n - E1) == c):
            mask |= 1 << i
    if mask:
        return 1 | (mask << 16)
    # ---- checkFRIProof, Plonk/FRI.hs:358-407
    pbits = fc["proof_of_work_bits"]
    if pbits and (pow_response >> (64 - pbits)) != 0:
        return 2
    y0, y1 = reduce_with_powers(fri_alpha, batch_this), reduce_with_powers(fri_alpha, batch_next)
    arities = expand_strategy(nbits, fc["reduction_strategy"])
    npp = -(-routed // qdf)
    omega, eta = subgroup_generator(nbits), subgroup_generator(lde_bits)
    all_caps = [vk_cap, caps["wires_cap"], caps["plonk_zs_partial_products_cap"], caps["quotient_polys_cap"]]
    for qi, (idx, rd) in enumerate(zip(indices, fri["query_round_proofs"])):
        eps = rd["initial_trees_proof"]["evals_proofs"]
        bad = 0
        for o in range(4):
            leaf, path = [x % P for x in eps[o][0]], [dg(d) for d in eps[o][1]["siblings"]]
            if not check_merkle_proof(all_caps[o], idx, leaf, path):
                bad |= 1 << o
        if bad:
            return 16 | (qi << 8) | (bad << 16)
        oc, ow, opl, oq = [[x % P for x in eps[o][0]] for o in range(4)]
        opp, olk = opl[:r * npp], opl[r * npp:]
        first = [E(x) for x in oc + ow + opp + oq + olk]
        second = [E(x) for x in opp[:r] + olk]
        g0, g1 = reduce_with_powers(fri_alpha, first), reduce_with_powers(fri_alpha, second)
        px = E(MUL_GEN * fpow(eta, reverse_bits(lde_bits, idx)) % P)
        ev = fri_alpha.pow(len(second)) * ((g0 - y0) / (px - zeta)) + (g1 - y1) / (px - zeta.scale(omega))
        shift, size, qx = MUL_GEN, lde_bits, idx
        for s, (a, beta, cap, step) in enumerate(zip(arities, fri_betas, commit_caps, rd["steps"])):
            A = 1 << a
            evals = ex(step["evals"])
            flat = [v for e in evals for v in e.pair()]
            if not check_merkle_proof(cap, qx >> a, flat, [dg(d) for d in step["merkle_proof"]["siblings"]]):
                return 17 | (qi << 8) | (s << 16)
            if not (evals[qx % A] == ev):
                return 18 | (qi << 8) | (s << 16)
            # prepareCoset + foldCosetWith, literally (:248-279)
            eb = subgroup_generator(size)
            ofs = shift * fpow(eb, reverse_bits(size, (qx >> a) << a)) % P
            vals = [None] * A
            for i, v in enumerate(evals):
                vals[reverse_bits(a, i)] = v
            om = subgroup_generator(a)
            ys = []
            for k in range(A):
                acc = E0
                for j in range(A):
                    acc = acc + vals[j].scale(fpow(ofs * pow(om, j, P) % P, -k))
                ys.append(acc)
            acc, bp = E0, E1
            for y in ys:
                acc, bp = acc + bp * y, beta * bp
            ev = acc.scale(inv(A))
            shift, size, qx = pow(shift, A, P), size - a, qx >> a
        xf = E(shift * fpow(subgroup_generator(size), reverse_bits(size, qx)) % P)
        acc, xp = E0, E1
        for co in final_poly:
            acc, xp = acc + co * xp, xf * xp
        if not (acc == ev):
            return 3 | (qi << 8)
    return 0


def eprod(xs):
    acc = E1
    for x in xs:
        acc = acc * x
    return acc


def load_fixture(golden_dir, name, common_name=None):
    rd = lambda n, k: json.load(open(os.path.join(golden_dir, "%s_%s.json" % (n, k))))
    return rd(common_name or name, "common"), rd(name, "vkey"), rd(name, "proof")


def testmain_text(common, vkey, proof):
    """Exactly what the reference's driver prints for this proof (src/testmain.hs:40-63) with the reference's Show
    instances: `MkDigest a b c d` (derived Show, Hash/Digest.hs:36-38, over Goldilocks' decimal `show`,
    Algebra/Goldilocks.hs:90-91), `(re + X*im)` (Algebra/GoldilocksExt.hs:37-38), Haskell list / Bool syntax.  A proof that
    hits one of the reference's `error` sites ends with `testmain: <message>` (what GHC prints on stderr before aborting);
    the messages are the ones at Plonk/FRI.hs:108, :310, :311."""
    tr = {}
    st = verify(common, vkey, proof, tr)
    op = proof["proof"]["openings"]
    out = ["public inputs hash = MkDigest %d %d %d %d" % tuple(tr["pih"][:4])]
    for k in ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products", "quotient_polys", "lookup_zs", "lookup_zs_next"):
        out.append("%-26s = %d" % ("# opening_" + k, len(op[k])))
    comb = tr["combined"]
    out.append("[" + ",".join("(%d + X*%d)" % (comb[2 * i], comb[2 * i + 1]) for i in range(len(comb) // 2)) + "]")
    r = common["config"]["num_challenges"]
    mask = (st >> 16) if (st & 0xFF) == 1 else 0
    out.append("[" + ",".join("False" if (mask >> i) & 1 else "True" for i in range(r)) + "]")
    code = st & 0xFF
    verdict = {0: "True", 1: "False", 2: "False", 3: "False",
               16: "testmain: checkInitialTreeProofs: at least one Merkle proof failed",
               17: "testmain: folding step Merkle proof does not check out",
               18: "testmain: folding step evaluation does not match the opening"}.get(code, "testmain: error site %d" % code)
    out.append("proof verification result = " + verdict)
    return "\n".join(out) + "\n"

