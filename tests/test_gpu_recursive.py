"""GPU parity of the two-curve `RecursiveSNARK` mirror (vimz_b200/recursive.py): the prove_step call order of
nova-snark 0.23.0 on both curves of the cycle, against a CPU chain driven through oracle/nova_cpu.c with the same
transcript, and -- at the full grayscale_step_HD size -- a whole 720-step HD proof ending in the reference's own
acceptance checks (RecursiveSNARK::verify = is_sat_relaxed x2 + is_sat, /root/reference/vimz/src/nova_snark_backend/folding.rs:46-56;
720 steps = one per image row, /root/reference/vimz/src/transformation.rs:93-101)."""
import numpy as np
import pytest

import vimz_b200
from vimz_b200 import (CommitmentKey, PublicParams, R1CSShape, RecursiveSNARK, TranscriptRO, fold_input, verify_folded_proof)
from vimz_b200.field import CURVES, affine_to_mont, ints_to_mont, mont_to_affine, mont_to_ints
from vimz_host import synthetic as S
from oracle import pyref as P

pytestmark = pytest.mark.gpu
K0, DK = 77, 1234577


def _key(eng, c, coracle, n, host_copy=True):
    import torch
    d = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    vimz_b200._lib.check(vimz_b200.lib.vimz_gen_bases_dev(eng._h, K0, DK, n, d.data_ptr()))
    ck = CommitmentKey.from_device(eng, d.data_ptr(), n)
    Bm = None
    if host_copy:
        Bm = coracle.gen_bases(c.curve_id, affine_to_mont([P.generator(c)], c.p)[0], K0, DK, n)
        assert np.array_equal(d.cpu().numpy().view(np.uint64).reshape(-1, 8), Bm)
    return ck, Bm


class CpuCurve:
    """One curve of the CPU chain: running relaxed pair + NIFS::prove through the C oracle, same transcript as the mirror."""

    def __init__(self, coracle, c, sh, Bm, digest):
        self.o, self.c, self.sh, self.Bm, self.digest = coracle, c, sh, Bm, digest
        m, n, io = sh.num_cons, sh.num_vars, sh.num_io
        self.W = np.zeros((n, 4), np.uint64); self.E = np.zeros((m, 4), np.uint64)
        self.u = np.zeros((1, 4), np.uint64); self.X = np.zeros((io, 4), np.uint64)
        self.cW = np.zeros(12, np.uint64); self.cE = np.zeros(12, np.uint64)
        self.one = ints_to_mont([1], c.q)

    def aff(self, jac):
        return mont_to_affine(self.o.to_affine(self.c.curve_id, jac), self.c.p)[0]

    def commit(self, v):
        return self.o.msm(self.c.curve_id, v, self.Bm, 4)

    def start_from(self, W, X):   # RelaxedR1CS*::from_r1cs_*
        self.W, self.X, self.u = W.copy(), X.copy(), self.one.copy()
        self.cW = self.commit(W)

    def nifs(self, W2, X2, comm_W2):
        o, c, sh, q = self.o, self.c, self.sh, self.c.q
        T = o.commit_T(c.curve_id, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C, self.W, self.u, self.X, W2, X2, self.one, nthreads=2)
        comm_T = self.commit(T)
        ro = TranscriptRO()
        ro.absorb_ints(self.digest)
        for pt in (self.cW, self.cE):
            a = self.aff(pt); ro.absorb_ints(*(a if a else (0, 0)))
        ro.absorb_ints(*mont_to_ints(self.u, q), *mont_to_ints(self.X, q))
        a = self.aff(comm_W2); ro.absorb_ints(*(a if a else (0, 0)))
        ro.absorb_ints(*mont_to_ints(X2, q))
        a = self.aff(comm_T); ro.absorb_ints(*(a if a else (0, 0)))
        r = ints_to_mont([ro.squeeze()], q)
        self.W = o.axpy(c.curve_id, self.W, W2, r); self.E = o.axpy(c.curve_id, self.E, T, r)
        tail = o.axpy(c.curve_id, np.concatenate([self.u, self.X]), np.concatenate([self.one, X2]), r)
        self.u, self.X = tail[:1], tail[1:]
        self.cW = o.point_scale_add(c.curve_id, self.cW, r, comm_W2); self.cE = o.point_scale_add(c.curve_id, self.cE, r, comm_T)
        return comm_T


def _witnesses(sh, q, count, seed):
    out = []
    for k in range(count):
        Wi, Xi = S.synthetic_witness(sh, seed + k)
        out.append((ints_to_mont(Wi, q), ints_to_mont(Xi, q)))
    return out


@pytest.mark.parametrize("cycle", [("pallas", "vesta"), ("bn254", "grumpkin")])
@pytest.mark.parametrize("overlap", [False, True])
def test_recursive_snark_chain_both_curves(cycle, overlap, coracle):
    """72 prove_step calls on both curves (strict order and with the deferred secondary commit): the first 8 steps are
    compared with the CPU chain value by value (comm_T of both NIFS instances, then the folded pairs), the whole chain ends
    in RecursiveSNARK::verify's three checks, and a tampered fresh secondary witness is rejected."""
    c1, c2 = P.CURVES[cycle[0]], P.CURVES[cycle[1]]
    e1, e2 = vimz_b200.Engine(cycle[0], 0), vimz_b200.Engine(cycle[1], 0)
    sh1 = S.synthetic_shape(CURVES[cycle[0]], "grayscale", seed=61, scale=0.012)
    sh2 = S.synthetic_shape(CURVES[cycle[1]], "secondary", seed=62, scale=0.08)
    ck1, B1 = _key(e1, c1, coracle, max(sh1.num_cons, sh1.num_vars))
    ck2, B2 = _key(e2, c2, coracle, max(sh2.num_cons, sh2.num_vars))
    pp = PublicParams(R1CSShape(e1, sh1.num_cons, sh1.num_vars, sh1.num_io, sh1.A, sh1.B, sh1.C), ck1,
                      R1CSShape(e2, sh2.num_cons, sh2.num_vars, sh2.num_io, sh2.A, sh2.B, sh2.C), ck2, digest=0xD16E57)
    w1, w2 = _witnesses(sh1, c1.q, 6, 700), _witnesses(sh2, c2.q, 5, 800)
    steps = 72
    seq = [(w1[k % 6], w2[k % 5]) for k in range(steps)]
    # CPU chain for the first 8 steps, in prove_step order
    cpu1, cpu2 = CpuCurve(coracle, c1, sh1, B1, pp.digest), CpuCurve(coracle, c2, sh2, B2, pp.digest)
    cpu1.start_from(*seq[0][0])
    l_w2 = seq[0][1]
    l_cw2 = cpu2.commit(l_w2[0])
    snark = RecursiveSNARK(pp, seq[0][0], seq[0][1], overlap_secondary=overlap)
    snark.prove_step(*seq[0])
    assert snark.i == 1
    for k in range(1, steps):
        snark.prove_step(*seq[k])
        if k <= 8:
            ct2 = cpu2.nifs(l_w2[0], l_w2[1], l_cw2)                      # (1)
            cw1 = cpu1.commit(seq[k][0][0])                               # (3)
            ct1 = cpu1.nifs(seq[k][0][0], seq[k][0][1], cw1)              # (4)
            l_w2 = seq[k][1]; l_cw2 = cpu2.commit(l_w2[0])               # (6)
            assert e2.to_affine_ints(snark.nifs_secondary) == cpu2.aff(ct2), (k, "comm_T secondary")
            assert e1.to_affine_ints(snark.nifs_primary) == cpu1.aff(ct1), (k, "comm_T primary")
            if not overlap:
                assert e2.to_affine_ints(snark.l_u_secondary.comm_W) == cpu2.aff(l_cw2)
        if k == 8:
            for acc, cpu, eng in ((snark.acc_primary, cpu1, e1), (snark.acc_secondary, cpu2, e2)):
                U, W = acc.download()
                assert np.array_equal(W.W, cpu.W) and np.array_equal(W.E, cpu.E) and np.array_equal(U.u, cpu.u) and np.array_equal(U.X, cpu.X)
                assert eng.to_affine_ints(U.comm_W) == cpu.aff(cpu.cW) and eng.to_affine_ints(U.comm_E) == cpu.aff(cpu.cE)
    assert snark.i == steps
    verify_folded_proof(snark, steps)
    with pytest.raises(vimz_b200.UnSat):
        verify_folded_proof(snark, steps + 1)                              # wrong step count
    # a fresh secondary pair that does not satisfy the circuit must fail the third check (is_sat on l_u2 / l_w2)
    bad = seq[3][1][0].copy()
    bad[sh2.num_vars - 1] = ints_to_mont([12345], c2.q)[0]
    snark._set_fresh_secondary((bad, seq[3][1][1]))
    with pytest.raises(vimz_b200.UnSat):
        snark.verify()
    snark.close()
    for h in (pp.shape_primary, pp.shape_secondary, ck1, ck2):
        h.close()
    e1.close(); e2.close()


def test_full_hd_grayscale_proof_720_steps(coracle):
    """The north-star stand-in: a whole HD grayscale proof -- 720 fold steps (one per image row) on Pallas at the full
    grayscale_step_HD size and on Vesta at the secondary size -- driven through fold_input, accepted by
    verify_folded_proof; the last primary commitment of the cross term is re-derived from the definition on the CPU."""
    c1, c2 = P.PALLAS, P.VESTA
    e1, e2 = vimz_b200.Engine("pallas", 0), vimz_b200.Engine("vesta", 0)
    sh1 = S.synthetic_shape(CURVES["pallas"], "grayscale")
    sh2 = S.synthetic_shape(CURVES["vesta"], "secondary")
    ck1, B1 = _key(e1, c1, coracle, 1 << 17)
    ck2, _ = _key(e2, c2, coracle, 1 << 14, host_copy=False)
    pp = PublicParams(R1CSShape(e1, sh1.num_cons, sh1.num_vars, sh1.num_io, sh1.A, sh1.B, sh1.C), ck1,
                      R1CSShape(e2, sh2.num_cons, sh2.num_vars, sh2.num_io, sh2.A, sh2.B, sh2.C), ck2, digest=1)
    w1, w2 = _witnesses(sh1, c1.q, 8, 7000), _witnesses(sh2, c2.q, 8, 8000)
    steps = 720
    snark = fold_input(pp, ((w1[k % 8], w2[(3 * k) % 8]) for k in range(steps - 1)), overlap_secondary=True)
    # last step by hand so the running primary pair just before it can be handed to the CPU oracle
    U, W = snark.acc_primary.download()
    snark.prove_step(w1[5], w2[2])
    assert snark.i == steps
    one = ints_to_mont([1], c1.q)
    T = coracle.commit_T(c1.curve_id, sh1.num_cons, sh1.num_vars, sh1.num_io, sh1.A, sh1.B, sh1.C, W.W, U.u, U.X, w1[5][0], w1[5][1], one, nthreads=4)
    assert np.array_equal(snark.acc_primary.last_T(), T)
    exp = mont_to_affine(coracle.to_affine(c1.curve_id, coracle.msm(c1.curve_id, T, B1, 8)), c1.p)[0]
    assert e1.to_affine_ints(snark.nifs_primary) == exp
    verify_folded_proof(snark, steps)
    snark.close()
    for h in (pp.shape_primary, pp.shape_secondary, ck1, ck2):
        h.close()
    e1.close(); e2.close()


def test_sonobe_call_sequence_bn254_grumpkin(coracle):
    """the fold step of the Sonobe backend (/root/reference/vimz/src/sonobe_backend/folding.rs:52-65 -- Nova + CycleFold
    over BN254 / Grumpkin) in its own call order: NIFS on G1 with the pair committed by the previous step, two CycleFold folds
    on G2, commit of the next fresh witness.  Checked against the CPU chain in the same order and by Nova::verify's
    satisfiability checks."""
    from vimz_b200 import SonobeNova
    c1, c2 = P.BN254, P.GRUMPKIN
    e1, e2 = vimz_b200.Engine("bn254", 0), vimz_b200.Engine("grumpkin", 0)
    sh1 = S.synthetic_shape(CURVES["bn254"], "grayscale", seed=91, scale=0.012)
    sh2 = S.synthetic_shape(CURVES["grumpkin"], "secondary", seed=92, scale=0.14)      # ~1.5 k rows: the size of a CycleFold circuit
    ck1, B1 = _key(e1, c1, coracle, max(sh1.num_cons, sh1.num_vars))
    ck2, B2 = _key(e2, c2, coracle, max(sh2.num_cons, sh2.num_vars))
    s1 = R1CSShape(e1, sh1.num_cons, sh1.num_vars, sh1.num_io, sh1.A, sh1.B, sh1.C)
    s2 = R1CSShape(e2, sh2.num_cons, sh2.num_vars, sh2.num_io, sh2.A, sh2.B, sh2.C)
    w1, w2 = _witnesses(sh1, c1.q, 5, 1700), _witnesses(sh2, c2.q, 7, 1800)
    digest = 0x50B0
    nova = SonobeNova(s1, ck1, s2, ck2, w1[0], digest=digest)
    cpu1, cpu2 = CpuCurve(coracle, c1, sh1, B1, digest), CpuCurve(coracle, c2, sh2, B2, digest)
    fresh = w1[0]
    fresh_cw = cpu1.commit(fresh[0])
    steps = 12
    for k in range(steps):
        nxt, cfW, cfE = w1[(k + 1) % 5], w2[(2 * k) % 7], w2[(2 * k + 1) % 7]
        nova.prove_step(nxt, cfW, cfE)
        ct1 = cpu1.nifs(fresh[0], fresh[1], fresh_cw)                                   # (a)
        ctW = cpu2.nifs(cfW[0], cfW[1], cpu2.commit(cfW[0]))                            # (b)
        ctE = cpu2.nifs(cfE[0], cfE[1], cpu2.commit(cfE[0]))                            # (c)
        fresh, fresh_cw = nxt, cpu1.commit(nxt[0])                                      # (d)
        assert e1.to_affine_ints(nova.cmT) == cpu1.aff(ct1), k
        assert e2.to_affine_ints(nova.cf_cmT[0]) == cpu2.aff(ctW) and e2.to_affine_ints(nova.cf_cmT[1]) == cpu2.aff(ctE), k
        assert e1.to_affine_ints(nova.u_i.comm_W) == cpu1.aff(fresh_cw)
    for acc, cpu, eng in ((nova.acc, cpu1, e1), (nova.cf_acc, cpu2, e2)):
        U, W = acc.download()
        assert np.array_equal(W.W, cpu.W) and np.array_equal(W.E, cpu.E) and np.array_equal(U.u, cpu.u) and np.array_equal(U.X, cpu.X)
        assert eng.to_affine_ints(U.comm_W) == cpu.aff(cpu.cW) and eng.to_affine_ints(U.comm_E) == cpu.aff(cpu.cE)
    assert nova.i == steps
    nova.verify()
    nova.close()
    for h in (s1, s2, ck1, ck2):
        h.close()
    e1.close(); e2.close()
