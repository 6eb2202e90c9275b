"""Pin the CPU oracle: against SURVEY.md App. A known answers, the big-int definition, the golden outputs
produced by the reference's own code (tests/golden/make_golden.py), and -- where oracle/_ref exists -- the
reference's code run live.  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as O

Q0, Q1, Q2 = O.SEAL_PRIMES_30

# (q, n) -> (psi, NTT(ramp)[0:4], sha16(NTT(ramp)), NTT(sm42)[0:2], sha16(NTT(sm42)))
KAT = {
    (Q0, 1024): (27772, (377994127, 713461820, 966212584, 891832897), "dadf656e0db1fcec", (30194933, 559184699), "87deb95a0431f1a7"),
    (Q0, 2048): (59576, (817346130, 496550137, 441992372, 360269859), "ef15ec72a1b8f75a", (878816017, 846135057), "4c0990d3c733f583"),
    (Q0, 4096): (503422, (56743001, 187342116, 276515956, 692783084), "f06769648ff96d35", (432347685, 879299372), "36a1101e7932a140"),
    (Q1, 1024): (57806, (463925991, 117852635, 601173197, 498028427), "f4a32865a5c55bd4", (162778881, 158521124), "ddd9d5aed8314f36"),
    (Q1, 2048): (859143, (600368160, 228072640, 110861058, 1045439711), "add679a1500a8c8e", (801691217, 114847704), "9a0d9a9ff159beef"),
    (Q1, 4096): (16768, (897254567, 555156310, 745884611, 433811507), "43aee98a3e463189", (195094614, 435882889), "c01082f426ec6f16"),
    (Q2, 1024): (1047695, (20992226, 621941223, 417253198, 968432625), "de6e87421e79aed7", (294032930, 100657508), "a04b585d08730040"),
    (Q2, 2048): (842745, (953337782, 934426826, 88020632, 204444244), "d80ea0fd5ba2706e", (492727231, 478551259), "8f486de34535d858"),
    (Q2, 4096): (7305, (2635272, 749013975, 884736591, 875608941), "9b587721c09a3c7a", (738261613, 246020003), "84177fb9278b7649"),
}


@pytest.mark.parametrize("q,n", sorted(KAT))
def test_survey_kats(q, n):
    psi, ramp4, ramp_sha, sm2, sm_sha = KAT[(q, n)]
    assert O.min_psi(n, q) == psi
    P = O.Plan(n, [q])
    for variant in ("shoup", "barrett"):
        ramp = np.arange(n, dtype=np.uint32).reshape(1, 1, n).copy()
        y = P.fwd(ramp.copy(), variant=variant)
        assert tuple(int(v) for v in y.ravel()[:4]) == ramp4 and O.sha16(y) == ramp_sha
        sm = P.synthetic(1, seed=42)
        assert int(sm[0, 0, 5]) == O.splitmix64(42 + 5) % q
        z = P.fwd(sm.copy(), variant=variant)
        assert tuple(int(v) for v in z.ravel()[:2]) == sm2 and O.sha16(z) == sm_sha
        assert (P.inv(z.copy(), variant=variant) == sm).all()


def test_min_psi_independent():
    assert O.py_min_psi(1024, Q0) == 27772 == O.min_psi(1024, Q0)
    assert O.py_min_psi(32, Q1) == O.min_psi(32, Q1)


@pytest.mark.parametrize("n", [32, 256, 1024])
def test_definition(n):
    q = Q2
    psi = O.min_psi(n, q)
    P = O.Plan(n, [q])
    x = P.synthetic(1, seed=3)
    y = P.fwd(x.copy())
    assert (O.textbook_fwd(x.ravel(), q, psi) == y.ravel()).all()
    if n <= 256:
        assert (O.textbook_inv(y.ravel(), q, psi) == x.ravel()).all()
    # simple vectors (SEAL-Embedded-style checks): zero, delta, constant*delta, monomial X
    z = np.zeros((1, 1, n), dtype=np.uint32)
    assert not P.fwd(z.copy()).any()
    d = z.copy(); d[0, 0, 0] = 7
    assert (P.fwd(d.copy()) == 7).all()
    m = z.copy(); m[0, 0, 1] = 1
    logn = n.bit_length() - 1
    expect = [pow(psi, 2 * O.bitrev(k, logn) + 1, q) for k in range(n)]
    assert P.fwd(m.copy()).ravel().tolist() == expect


def test_u64_restatement_equals_u32_values():
    n, q = 1024, Q0
    r64, p64 = O.tables_u64(n, q)
    r32, _ = O.tables_u32(n, q)
    assert (r64 == r32).all()
    P = O.Plan(n, [q])
    x = P.synthetic(2, seed=11)
    y64 = O.ref_fwd_u64(x.ravel().astype(np.uint64), x.ravel().astype(np.uint64), q, r64, p64, 2)
    assert (y64 == P.fwd(x.copy()).ravel()).all()


def test_polymul_kat_and_schoolbook():
    n, q = 2048, Q0
    a = np.arange(1, n + 1, dtype=np.uint32)
    b = (2 * np.arange(n) + 1).astype(np.uint32)
    c = O.polymul_schoolbook(a, b, q)
    assert c[:4].tolist() == [291855365, 287667213, 283483167, 279303231]
    assert c[-2:].tolist() == [745190398, 757771262] and O.sha16(c) == "cc989f874ebe7af7"
    assert (O.np_polymul_exact(a, b, q) == c).all()
    P = O.Plan(n, [q])
    assert (P.polymul(a.reshape(1, 1, n).copy(), b.reshape(1, 1, n).copy()).ravel() == c).all()
    # X^(n-1) * X = -1
    e1 = np.zeros(n, dtype=np.uint32); e1[n - 1] = 1
    e2 = np.zeros(n, dtype=np.uint32); e2[1] = 1
    r = P.polymul(e1.reshape(1, 1, n).copy(), e2.reshape(1, 1, n).copy()).ravel()
    assert r[0] == q - 1 and not r[1:].any()


def test_polymul_bigint_small():
    n, q = 64, Q1
    P = O.Plan(n, [q])
    a = P.synthetic(1, seed=1).ravel(); b = P.synthetic(1, seed=2).ravel()
    assert (O.textbook_polymul(a, b, q) == O.polymul_schoolbook(a, b, q)).all()


def test_rns_layout_and_threads():
    n = 4096
    P = O.Plan(n, O.SEAL_PRIMES_30)
    x = P.synthetic(4, seed=42)
    y1 = P.fwd(x.copy(), threads=1)
    y2 = P.fwd(x.copy(), threads=4, variant="barrett")
    assert (y1 == y2).all()
    for l, q in enumerate(O.SEAL_PRIMES_30):
        Pl = O.Plan(n, [q])
        assert (Pl.fwd(np.ascontiguousarray(x[:, l:l + 1, :])) == y1[:, l:l + 1, :]).all()
        assert (y1[:, l, :] < q).all()
    assert (P.inv(y1.copy(), threads=2) == x).all()
    parts = O.checksum_u32(x[:2]) + O.checksum_u32(x[2:], first_index=x[:2].size)
    assert parts % 2**64 == O.checksum_u32(x)  # checksum of shards == checksum of the whole


def _golden_inputs(kind, N, q, frames, seed):
    g = np.arange(N * frames, dtype=np.uint64)
    if kind == "ramp":
        return g % np.uint64(q)
    if kind == "splitmix":
        return np.array([O.splitmix64(seed + int(i)) % q for i in g], dtype=np.uint64)
    if kind == "lazy4q":
        return np.array([O.splitmix64(seed + int(i)) % (4 * q) for i in g], dtype=np.uint64)
    x = np.zeros(N * frames, dtype=np.uint64); x[::N] = 1
    return x


def test_golden_from_reference_code(golden):
    """Outputs recorded from the reference's own ntt.cpp (tests/golden/make_golden.py) == C restatement."""
    g, meta = golden
    n_cases = 0
    for name, m in meta.items():
        if name.startswith("main_dummy"):
            continue
        N, q, psi, kind, frames, seed = int(m[0]), int(m[1]), int(m[2]), m[3], int(m[4]), int(m[5])
        assert O.min_psi(N, q) == psi
        r, p = O.tables_u64(N, q, psi)
        x = _golden_inputs(kind, N, q, frames, seed)
        y = O.ref_fwd_u64(x, x, q, r, p, frames)
        assert (y == g[name]).all(), name
        if kind != "lazy4q":  # the u32 datapath gives the same residues
            P = O.Plan(N, [q])
            assert (P.fwd(x.astype(np.uint32).reshape(frames, 1, N).copy()).ravel() == g[name]).all(), name
        n_cases += 1
    assert n_cases >= 10


def u64_case_io(c):
    mod = 4 * c["q"] if c["lazy"] else c["q"]
    n = c["N"] * c["frames"]
    return O.synthetic_u64(n, c["seed_in"], mod), O.synthetic_u64(n, c["seed_in2"], mod)


def test_golden_u64_large_primes(golden_u64):
    """The moduli the reference datapath is written for (64-bit, ntt.cpp:147-148, 344-363): 50-, 60- and 62-bit NTT
    primes, lazy inputs, several frames, in2 != in -- restatement == outputs recorded from the reference's own code;
    and a 63-bit prime, where [0,4q) no longer fits and the arithmetic wraps mod 2^64 exactly like the reference's."""
    g, cases = golden_u64
    assert len(cases) >= 10
    for name, c in cases.items():
        assert O.is_prime(c["q"]) and O.min_psi(c["N"], c["q"]) == c["psi"]
        r, p = O.tables_u64(c["N"], c["q"], c["psi"])
        x, x2 = u64_case_io(c)
        y = O.ref_fwd_u64(x, x2, c["q"], r, p, c["frames"])
        assert hashlib.sha256(y.tobytes()).hexdigest() == c["sha256"], name
        if name in g.files:
            assert (y == g[name]).all(), name
        if c["q"] < 2**62:
            assert (y < c["q"]).all()
            # and it IS the transform: frame 0 against an independent big-int evaluation at a few output points
            logn = c["N"].bit_length() - 1
            f0 = [int(v) for v in x[:c["N"] // 2]] + [int(v) for v in x2[c["N"] // 2:c["N"]]]
            for k in (0, 1, c["N"] - 1):
                w = pow(c["psi"], 2 * O.bitrev(k, logn) + 1, c["q"])
                acc, pw = 0, 1
                for v in f0:
                    acc += v * pw
                    pw = pw * w % c["q"]
                assert acc % c["q"] == int(y[k]), (name, k)


def test_main_dummy_data_kat(golden):
    """main.cpp:49-55 dummy data through the u64 path; hashes from SURVEY.md App. A and from the reference run."""
    g, meta = golden
    N = 16384
    i = np.arange(N, dtype=np.uint64)
    y = O.ref_fwd_u64(i, i + 1, 65537, i + 2, i + 3)
    assert y[:4].tolist() == [15752083817248508221, 15907608836333597093, 16669572192778225120, 5382384308376256666]
    assert int(y[-1]) == 14647912012985937668
    assert hashlib.sha256(y.tobytes()).hexdigest() == meta["main_dummy_sha256_le64"][0] == \
        "79f01363d7e86876e7c91319724484ce4acfdd266ca8dfaf53dd394faff75aed"
    txt = "".join("%d\n" % int(v) for v in y)
    assert hashlib.sha256(txt.encode()).hexdigest() == meta["main_dummy_sha256_lines"][0] == \
        "68db50a07a87d4e18387a721d88b42291198aba2f9e9138e2e52d45ad6537c5c"
    assert (y == g["main_dummy_u64"]).all()


@pytest.mark.parametrize("N", O.REF_SIZES)
def test_live_reference_code(N):
    """Where oracle/_ref was built (build container), run the reference's own kernel live against the restatement."""
    if not O.ref_available(N):
        pytest.skip("oracle/_ref not built here")
    q = Q1
    r, p = O.tables_u64(N, q)
    frames = 2 if N <= 8192 else 1
    x = np.array([O.splitmix64(1000 + i) % (4 * q) for i in range(N * frames)], dtype=np.uint64)
    x2 = np.array([O.splitmix64(5000 + i) % q for i in range(N * frames)], dtype=np.uint64)
    a = O.reference_fwd_u64(N, x, x2, q, r, p, frames)
    b = O.ref_fwd_u64(x, x2, q, r, p, frames)
    assert (a == b).all() and (a < q).all()
    for bits in (50, 60, 63):          # the 64-bit moduli the kernel is written for; 63 bits wraps mod 2^64
        q = O.U64_PRIMES[bits]
        r, p = O.tables_u64(N, q)
        mod = 4 * q if bits < 62 else q
        x, x2 = O.synthetic_u64(N, 77 + bits, mod), O.synthetic_u64(N, 78 + bits, mod)
        assert (O.reference_fwd_u64(N, x, x2, q, r, p, 1) == O.ref_fwd_u64(x, x2, q, r, p, 1)).all(), bits
