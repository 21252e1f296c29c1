"""CPU suite: pin oracle/symmshe.py and oracle/coeffwise.py through the reference's own SymmSHE test properties
(lol-apps/Crypto/Lol/Applications/Tests/SHETests.hs): prop_encDec (:172-177), prop_ctmul (:148-158), prop_ksQuad (:200-209), prop_modSwPT (:179-189),
prop_ksLin (:191-198), prop_ctembed / prop_cttwace (:211-226), prop_ringTunnel (:228-248), plus the correctness of modSwitch (SymmSHE.hs:236-248), which the reference's benchmarks exercise but its tests do not.

The scheme around the restated steps -- encrypt, ksQuadCircHint, toMSD / toLSD, decryptUnrestricted -- is restated here
from lol-apps/Crypto/Lol/Applications/SymmSHE.hs (:131-141, :199-206, :222-232, :259-287, :346-372) for m = m', with
every ring transform (crt, crtInv, l, lInv, divGDec) done by the COMPILED reference.  If the restated decomposition,
knapsack, ciphertext product, lift or limb drop disagreed with the reference's semantics, decryption would not return
pt1 * pt2: the gadget identity and the noise bound both have to hold.
"""
import numpy as np
import pytest

from oracle import coeffwise as W
from oracle import extension as X
from oracle import symmshe as S
from oracle import tables as T
from test_oracle_extension import Ring

M, QS, P = 21, [19393921, 18869761], 5          # 21 | q - 1 for both primes; rad_odd(21) is a unit modulo 5
BIGQ = 2148249601                               # 21 | BIGQ - 1: exact integer ring products through its CRT


class Scheme:
    def __init__(self, reference, rng, p=P, m=M, s_dec=None):
        self.ref, self.rng, self.p, self.m = reference, rng, p, m
        self.R = Ring(reference, m, QS)
        self.n, self.k = self.R.n, len(QS)
        self.pe = T.pe_array(m)
        self.g = T.g_crt_vectors(m, QS)[0]
        self.tables = (self.pe, self.R.ru, self.R.rui, self.R.mh, self.g)
        self.q = np.asarray(QS, dtype=np.int64)
        # genSK (SymmSHE.hs:118-121): errorRounded -> small integers in the decoding basis
        self.s_dec = rng.integers(-1, 2, size=self.n) if s_dec is None else np.asarray(s_dec, dtype=np.int64)
        self.sq_crt = self.R.crt(self.dec_to_pow(self.s_dec))

    # -- helpers over R_q (ABI arrays [n, k])
    def dec_to_pow(self, z_dec):
        """reduce an integer Dec-basis element into R_q and convert to the powerful basis (l: Dec -> Pow)."""
        return self.R.l(S.reduce_digit(z_dec, QS))

    def mul(self, a, b): return S._mulmod(a, b, QS)
    def add(self, a, b): return S._addmod(a, b, QS)
    def neg(self, a): return (self.q - a) % self.q
    def uniform(self): return np.stack([self.rng.integers(0, q, size=self.n) for q in QS], axis=-1).astype(np.int64)

    def small_error(self):
        """errorRounded (UCyc.hs:427-429): round of a Gaussian in the decoding basis."""
        return W.round_coset(self.rng.standard_normal((self.n, 1)) * 1.5, None, [self.p])[:, 0]

    # -- SymmSHE.hs:131-141
    def encrypt(self, pt_dec):
        gauss = self.rng.standard_normal((self.n, 1)) * 1.5 * self.p                     # tGaussian (svar p^2), UCyc.hs:443
        e_dec = W.round_coset(gauss, np.asarray(pt_dec, dtype=np.int64).reshape(-1, 1), [self.p])[:, 0]   # roundCoset <$> c <*> err
        assert np.array_equal(e_dec % self.p, np.asarray(pt_dec) % self.p)
        c1 = self.uniform()
        c0 = self.add(self.R.crt(self.dec_to_pow(e_dec)), self.neg(self.mul(c1, self.sq_crt)))
        return {"enc": "LSD", "k": 0, "l": 1, "c": [c0, c1]}                       # CRT-basis components

    # -- SymmSHE.hs:222-232
    def to_msd(self, ct):
        if ct["enc"] == "MSD":
            return ct
        qprod = int(np.prod([int(q) for q in QS], dtype=object))       # zpScale = -q mod p (ZqBasic.hs:135-137; Prelude.hs:311-315)
        return {"enc": "MSD", "k": ct["k"], "l": ct["l"] * (-qprod) % self.p, "c": [self.scal_inv_p(c) for c in ct["c"]]}

    def scal_inv_p(self, a):
        w = np.asarray([pow(self.p, -1, q) for q in QS], dtype=np.int64)      # zqScale = recip (reduce p), limb by limb
        return S._mulmod(a, np.broadcast_to(w, a.shape), QS)

    # -- SymmSHE.hs:259-287: ksHint sk (s*s) with lweSample; hint [ell, 2, n, k] in the CRT basis
    def ks_quad_circ_hint(self, base):
        val = self.mul(self.sq_crt, self.sq_crt)
        hint = []
        for gad in S.gadget(QS, base):
            c1 = self.uniform()
            b = self.add(self.mul(c1, self.neg(self.sq_crt)), self.R.crt(self.dec_to_pow(self.small_error())))
            valgad = self.mul(val, np.broadcast_to(np.asarray(gad, dtype=np.int64), val.shape))
            hint.append([self.add(valgad, b), c1])
        return np.asarray(hint, dtype=np.int64)

    # -- SymmSHE.hs:199-206 (decryptUnrestricted), for a polynomial of any degree over the limbs qs
    def decrypt(self, ct, qs=None, p=None):
        qs = QS if qs is None else qs
        P = self.p if p is None else p                                               # plaintext modulus of this ciphertext
        R = self.R if qs == QS else Ring(self.ref, self.m, qs)
        sq = self.sq_crt if qs == QS else R.crt(R.l(S.reduce_digit(self.s_dec, qs)))
        l = ct["l"]
        comps = ct["c"]
        if ct["enc"] == "MSD":                                                       # toLSD
            qprod = int(np.prod([int(q) for q in qs], dtype=object))
            l = l * pow((-qprod) % P, -1, P) % P
            w = np.asarray([P % q for q in qs], dtype=np.int64)
            comps = [S._mulmod(c, np.broadcast_to(w, c.shape), qs) for c in comps]
        ev, spow = np.zeros_like(comps[0]), None
        for i, c in enumerate(comps):                                                # evaluate c sq
            spow = sq if i == 1 else (S._mulmod(spow, sq, qs) if i > 1 else None)
            ev = S._addmod(ev, c if i == 0 else S._mulmod(c, spow, qs), qs)
        dec = R.l_inv(R.crt_inv(ev))                                                 # uncycDec
        Q = int(np.prod([int(q) for q in qs], dtype=object))
        e = []
        for row in dec:                                                              # lift over the product modulus, reduce mod p
            x = sum(int(r) * (Q // q) * pow(Q // q, -1, q) for r, q in zip(row, qs)) % Q
            e.append((x if 2 * x < Q else x - Q) % P)
        e = np.asarray(e, dtype=np.int64).reshape(-1, 1)
        for _ in range(ct["k"]):                                                     # iterate divG' e !! k
            e, ok = self.ref.tensorGInvDecRq(e, self.pe, [P])
            assert ok == 1
            e = e.reshape(-1, 1)
        return (e[:, 0] * l) % P

    def plain_product(self, a_dec, b_dec):
        """pt1 * pt2 in R_p, Dec-basis coefficients: exact integer product through the CRT of a large prime."""
        B = Ring(self.ref, self.m, [BIGQ])
        to_pow = lambda z: self.ref.tensorLR(np.asarray(z, dtype=np.int64).reshape(-1, 1), self.pe).reshape(-1, 1)
        prod = B.crt_inv(S._mulmod(B.crt(to_pow(a_dec) % BIGQ), B.crt(to_pow(b_dec) % BIGQ), [BIGQ]))
        prod = W.lift(prod, [BIGQ])
        return self.ref.tensorLInvR(prod, self.pe).reshape(-1) % self.p


@pytest.fixture()
def scheme(reference):
    return Scheme(reference, np.random.default_rng(2024))


def test_prop_encDec(scheme):
    for _ in range(3):
        pt = scheme.rng.integers(0, P, size=scheme.n)
        assert np.array_equal(scheme.decrypt(scheme.encrypt(pt)), pt)
        assert np.array_equal(scheme.decrypt(scheme.to_msd(scheme.encrypt(pt))), pt)


def test_prop_ctmul_quadratic_ciphertext(scheme):
    """(*) on CT (SymmSHE.hs:443-449) = oracle/symmshe.ct_mul_crt: the degree-2 product decrypts to pt1 * pt2 with k = 1."""
    pt1, pt2 = scheme.rng.integers(0, P, size=scheme.n), scheme.rng.integers(0, P, size=scheme.n)
    ct1, ct2 = scheme.encrypt(pt1), scheme.encrypt(pt2)
    prod = {"enc": "LSD", "k": 1, "l": 1, "c": S.ct_mul_crt(ct1["c"], ct2["c"], scheme.g, QS)}
    assert np.array_equal(scheme.decrypt(prod), scheme.plain_product(pt1, pt2))


@pytest.mark.parametrize("base", [0, 2, 16], ids=["TrivGad", "BaseBGad2", "BaseBGad16"])
def test_prop_ksQuad(scheme, base):
    """keySwitchQuadCirc hint (ct1 * ct2) through oracle/symmshe.mul_and_switch decrypts to pt1 * pt2."""
    pt1, pt2 = scheme.rng.integers(0, P, size=scheme.n), scheme.rng.integers(0, P, size=scheme.n)
    ct1, ct2 = scheme.encrypt(pt1), scheme.to_msd(scheme.encrypt(pt2))             # LSD * MSD -> MSD: toMSD inside the switch is id
    hint = scheme.ks_quad_circ_hint(base)
    assert hint.shape == (S.gadget_length(QS, base), 2, scheme.n, scheme.k)
    pow_ = lambda ct: [scheme.R.crt_inv(c) for c in ct["c"]]
    out = S.mul_and_switch(scheme.ref, pow_(ct1), pow_(ct2), hint, scheme.tables, QS, base)
    lin = {"enc": "MSD", "k": ct1["k"] + ct2["k"] + 1, "l": ct1["l"] * ct2["l"] % P, "c": out}
    assert np.array_equal(scheme.decrypt(lin), scheme.plain_product(pt1, pt2))
    # a wrong hint must NOT decrypt (the test has teeth)
    bad = hint.copy()
    bad[0, 0] = scheme.uniform()
    out_bad = S.mul_and_switch(scheme.ref, pow_(ct1), pow_(ct2), bad, scheme.tables, QS, base)
    assert not np.array_equal(scheme.decrypt({**lin, "c": out_bad}), scheme.plain_product(pt1, pt2))


@pytest.mark.parametrize("drop", [0, 1])
def test_mod_switch_keeps_the_plaintext(scheme, drop):
    """modSwitch (SymmSHE.hs:236-248): rescaleDec on c0, rescalePow on c1 with the limb-drop instance
    (oracle/coeffwise.rescale_drop); the ciphertext over the remaining modulus decrypts to the same plaintext."""
    pt = scheme.rng.integers(0, P, size=scheme.n)
    ct = scheme.to_msd(scheme.encrypt(pt))
    keep = [q for t, q in enumerate(QS) if t != drop]
    Rk = Ring(scheme.ref, M, keep)
    c0_dec = scheme.R.l_inv(scheme.R.crt_inv(ct["c"][0]))
    c1_pow = scheme.R.crt_inv(ct["c"][1])
    c0 = Rk.crt(Rk.l(W.rescale_drop(c0_dec, QS, drop)))                             # rescaleDec c0
    c1 = Rk.crt(W.rescale_drop(c1_pow, QS, drop))                                   # rescalePow c1
    assert np.array_equal(scheme.decrypt({"enc": "MSD", "k": 0, "l": ct["l"], "c": [c0, c1]}, keep), pt)


def test_prop_modSwPT(reference):
    """SHETests.hs:179-189: z = (p/p') * ct; modSwitchPT z decrypts (mod p') to rescaleCyc Dec (decrypt z) -- pins
    oracle/coeffwise.rescale_mod (the `Rescale (ZqBasic p) (ZqBasic p')` instance = rescaleMod, Prelude.hs:143-153)."""
    p, p2 = 25, 5
    sch = Scheme(reference, np.random.default_rng(7), p)
    for _ in range(3):
        pt = sch.rng.integers(0, p, size=sch.n)
        y = sch.encrypt(pt)
        # (fromIntegral (p `div` p')) * y : a CT product with the constant ciphertext CT LSD 0 one [c] (SymmSHE.hs:436-449)
        w = np.asarray([(p // p2) % q for q in QS], dtype=np.int64)
        comps = [S._mulmod(S._mulmod(c, np.broadcast_to(w, c.shape), QS), sch.g, QS) for c in y["c"]]
        z = {"enc": "LSD", "k": 1, "l": 1, "c": comps}
        x = sch.decrypt(z)
        assert np.array_equal(x, (p // p2) * pt % p)
        zm = sch.to_msd(z)
        l2 = int(W.reduce(W.lift(np.asarray([[zm["l"]]]), [p]), [p2])[0, 0])        # modSwitchPT: reduce (lift l)
        x2 = sch.decrypt({**zm, "l": l2}, p=p2)
        assert np.array_equal(x2, W.rescale_mod(x.reshape(-1, 1), [p], [p2])[:, 0])


@pytest.mark.parametrize("m,m2", [(3, 21), (7, 21), (1, 21)])
def test_prop_cttwace_and_prop_ctembed(reference, m, m2):
    """SHETests.hs:211-226 with r = r' = m', s = s' = m: twaceCT (SymmSHE.hs:497-503) of a ciphertext under embedSK sk decrypts
    under sk to twace pt, and embedCT (:471-479) of a ciphertext under sk decrypts under embedSK sk to embed pt -- the
    restated twaceCRT / embedCRT / embedDec / twacePowDec (oracle/extension.py) inside the scheme, compiled reference
    transforms in both rings."""
    rng = np.random.default_rng(100 * m + m2)
    info = X.ExtInfo(m, m2)
    small = Scheme(reference, rng, P, m)                                             # sk in O_m
    big = Scheme(reference, rng, P, m2, s_dec=X.embed_dec(info, small.s_dec.reshape(-1, 1))[:, 0])    # embedSK: same key in O_m'
    for _ in range(2):
        pt = rng.integers(0, P, size=big.n)
        ct = big.encrypt(pt)
        tw = {"enc": "LSD", "k": 0, "l": 1, "c": [X.twace_crt_zq(info, c, QS) for c in ct["c"]]}
        assert np.array_equal(small.decrypt(tw), X.twace_powdec(info, pt.reshape(-1, 1))[:, 0])
        pt = rng.integers(0, P, size=small.n)
        ct = small.encrypt(pt)
        em = {"enc": "LSD", "k": 0, "l": 1, "c": [X.embed_crt(info, c) for c in ct["c"]]}
        assert np.array_equal(big.decrypt(em), X.embed_dec(info, pt.reshape(-1, 1), [P])[:, 0])


@pytest.mark.parametrize("base", [0, 4], ids=["TrivGad", "BaseBGad4"])
def test_prop_ksLin(reference, base):
    """SHETests.hs:191-198: keySwitchLinear (SymmSHE.hs:333-344) from s_in to s_out with ksHint s_out s_in, built from the
    restated decompose / reduce / knapsack: c0 + switch hint c1 decrypts under s_out to the same plaintext."""
    rng = np.random.default_rng(77 + base)
    sin, sout = Scheme(reference, rng), Scheme(reference, rng)
    pt = rng.integers(0, P, size=sin.n)
    ct = sin.to_msd(sin.encrypt(pt))
    # ksHint skout sin: gadget * s_in + LWE samples under s_out
    hint = []
    for gad in S.gadget(QS, base):
        c1 = sout.uniform()
        b = sout.add(sout.mul(c1, sout.neg(sout.sq_crt)), sout.R.crt(sout.dec_to_pow(sout.small_error())))
        hint.append([sout.add(sout.mul(sin.sq_crt, np.broadcast_to(np.asarray(gad, dtype=np.int64), c1.shape)), b), c1])
    hint = np.asarray(hint, dtype=np.int64)
    digits = S.decompose_reduced(sin.R.crt_inv(ct["c"][1]), QS, base)                 # fmap reduce <$> decompose c1 (Pow basis)
    digits_crt = np.stack([sin.R.crt(np.ascontiguousarray(d)) for d in digits])
    out = S.knapsack(hint, digits_crt, ct["c"][0], np.zeros_like(ct["c"][0]), QS)     # P.const c0 + switch hint c1
    assert np.array_equal(sout.decrypt({**ct, "c": out}), pt)
    assert not np.array_equal(sin.decrypt({**ct, "c": out}), pt)                      # and no longer under s_in


def test_prop_ringTunnel(reference):
    """SHETests.hs:228-248 with r = r' = 21, s = s' = 15, e = e' = 3: tunnel (SymmSHE.hs:548-572) applies a random E-linear
    function R -> S homomorphically.  Restated from SymmSHE.hs:520-572 and Linear.hs:65-79, 113-119 on top of
    oracle/extension.py (coeffs in the Dec and Pow bases, embedDec, embedPow, the powerful extension basis) and
    oracle/symmshe.py (decompose, knapsack); every transform and ring product by the compiled reference."""
    p, r, s, e = 2, 21, 15, 3
    rng = np.random.default_rng(99)
    skin, skout = Scheme(reference, rng, p, r), Scheme(reference, rng, p, s)
    ext_r, ext_s = X.ExtInfo(e, r), X.ExtInfo(e, s)
    nb = ext_r.rel                                                                   # |basis of R/E| = 6
    bs_p = [rng.integers(0, p, size=skout.n) for _ in range(nb)]                     # linearDec bs: images in S_p (Dec basis)

    # -- plaintext side: expected = evalLin f x = sum_i bs_i * embed (coeffsDec x)_i   (Linear.hs:75-79)
    x = rng.integers(0, p, size=skin.n)
    cs = X.coeffs_powdec(ext_r, x.reshape(-1, 1))                                    # [nb, phi(e), 1]
    expected = np.zeros(skout.n, dtype=np.int64)
    for i in range(nb):
        expected = (expected + skout.plain_product(bs_p[i], X.embed_dec(ext_s, cs[i], [p])[:, 0])) % p

    # -- f'q = reduce (extendLin (lift f)): the images mod q, CRT basis of S'
    Rr, Rs = skin.R, skout.R
    bs_q = [Rs.crt(skout.dec_to_pow(W.lift(b.reshape(-1, 1), [p])[:, 0])) for b in bs_p]

    def eval_lin_q(c_crt):
        """evalLin f'q on an element of R'_q given in the CRT basis -> element of S'_q in the CRT basis."""
        parts = X.coeffs_powdec(ext_r, Rr.l_inv(Rr.crt_inv(c_crt)))                  # coeffsDec: [nb, phi(e), k]
        acc = np.zeros((skout.n, len(QS)), dtype=np.int64)
        for i in range(nb):
            emb = Rs.crt(Rs.l(X.embed_dec(ext_s, parts[i], QS)))
            acc = S._addmod(acc, S._mulmod(bs_q[i], emb, QS), QS)
        return acc

    # -- tunnelHint (SymmSHE.hs:520-533): ksHint skout (evalLin f' (sin * p_i)), p_i the powerful basis of R'/E'
    base, hints = 0, []
    for i in range(nb):
        p_i = np.zeros((skin.n, len(QS)), dtype=np.int64)
        p_i[(ext_r.base_pow_j0 == i) & (ext_r.base_pow_j1 == 0)] = 1
        val = eval_lin_q(S._mulmod(skin.sq_crt, Rr.crt(p_i), QS))
        hint = []
        for gad in S.gadget(QS, base):
            c1 = skout.uniform()
            b = skout.add(skout.mul(c1, skout.neg(skout.sq_crt)), Rs.crt(skout.dec_to_pow(skout.small_error())))
            hint.append([skout.add(skout.mul(val, np.broadcast_to(np.asarray(gad, dtype=np.int64), val.shape)), b), c1])
        hints.append(np.asarray(hint, dtype=np.int64))

    # -- tunnel (SymmSHE.hs:548-572)
    ct = skin.to_msd(skin.encrypt(x))
    c0, c1 = ct["c"]
    out0 = eval_lin_q(c0)
    out1 = np.zeros_like(out0)
    c1s = X.coeffs_powdec(ext_r, Rr.crt_inv(c1))                                     # coeffsPow c1 :: [E'_q]
    for i in range(nb):
        emb_pow = X.embed_pow(ext_s, c1s[i])                                         # embed <$> c1s (Pow basis of S')
        digits = S.decompose_reduced(emb_pow, QS, base)
        digits_crt = np.stack([Rs.crt(np.ascontiguousarray(d)) for d in digits])
        out0, out1 = S.knapsack(hints[i], digits_crt, out0, out1, QS)                # zipWith switch hints, summed
    actual = skout.decrypt({"enc": "MSD", "k": 0, "l": ct["l"], "c": [out0, out1]})
    assert np.array_equal(actual, expected)
    assert expected.any()                                                            # not the trivial function
