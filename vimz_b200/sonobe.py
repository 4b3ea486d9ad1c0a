"""Call-sequence mirror of the Sonobe backend's fold step on the same resident accumulators.

zero-savvy/vimz's second backend folds with Sonobe's Nova + CycleFold over BN254 / Grumpkin
(/root/reference/vimz/src/sonobe_backend/folding.rs:22 `Nova<G1, G2, Circuit, KZG<Bn254>, Pedersen<G2>, false>`,
step loop at folding.rs:52-65 `folding.prove_step(rng, input, None)`).  Sonobe is a git dependency that is not in the tree
([EXT sonobe@d312916] folding/nova/{mod,nifs,cyclefold}.rs), so the sequence below is restated from its published source:

    prove_step, i >= 1:
      (a) NIFS.prove on the PRIMARY curve (BN254 G1): T = cross term of the running (U_i, W_i) and the fresh (u_i, w_i) that the
          previous step left behind; cmT = commit(T) -- a KZG commitment without blinding is the same MSM over the SRS
          points as a Pedersen commitment over its generators; r from the transcript; fold.
      (b) CycleFold, instance "cfW": the tiny Grumpkin circuit that checks cmW_{i+1} = cmW_i + r * cmw_i; its witness is
          committed (Pedersen over G2) and folded into the running CycleFold pair.
      (c) CycleFold, instance "cfE": the same for cmE_{i+1} = cmE_i + r * cmT, folded into the result of (b).
      (d) the augmented F circuit is synthesised for the next step (untouched host code), w_{i+1} = its witness,
          cmw_{i+1} = commit(w_{i+1}) on G1.

The three-part computation per fold (mat-vecs, cross term, MSMs, witness / instance folds) is the one the nova-snark path
runs, over the BN254 / Grumpkin instantiations of the same kernels; only the order differs: on the primary curve the fresh
commitment is made at the END of a step and folded at the START of the next one (vimz_acc_commit_fresh / vimz_acc_cross_begin),
and the secondary curve folds TWO small instances per step.  Witness synthesis, the Poseidon transcript, the decider
(Groth16 + KZG) and the Solidity verifier are out of scope (BASELINE.json north_star): witnesses are inputs, the challenge
comes from the SHAKE stand-in `TranscriptRO` with NIFS's absorb order.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .field import as_fr
from .nova import (CommitmentKey, FoldAccumulator, R1CSInstance, R1CSShape, R1CSWitness, is_sat, is_sat_relaxed)
from .recursive import Witness, nifs_challenge


class SonobeNova:
    """[EXT sonobe] folding/nova/mod.rs `Nova{W_i, U_i, w_i, u_i, cf_W_i, cf_U_i, i, ..}` with both running pairs in HBM."""

    def __init__(self, shape: R1CSShape, srs: CommitmentKey, cf_shape: R1CSShape, cf_ck: CommitmentKey, w0: Witness, digest: int = 0):
        """`Nova::init`: dummy (all-zero) running pairs on both curves; the first augmented-circuit witness `w0` is committed
        and stays fresh until the first prove_step folds it."""
        self.shape, self.srs, self.cf_shape, self.cf_ck, self.digest = shape, srs, cf_shape, cf_ck, digest
        self.acc = FoldAccumulator(shape, srs)            # (U_i, W_i) over G1
        self.cf_acc = FoldAccumulator(cf_shape, cf_ck)    # (cf_U_i, cf_W_i) over G2
        self.i = 0
        self.u_i: Optional[R1CSInstance] = None
        self.cmT: Optional[np.ndarray] = None
        self.cf_cmT: Tuple[Optional[np.ndarray], Optional[np.ndarray]] = (None, None)
        self._commit_fresh(w0)

    def _commit_fresh(self, w: Witness) -> None:   # (d): cmw = commit(w) on G1; (w, x) staged for the next step's NIFS
        W, X = as_fr(w[0], self.shape.num_vars), as_fr(w[1], self.shape.num_io)
        self.u_i = R1CSInstance(self.acc.commit_fresh(W, X), X)

    def _fold_cyclefold(self, w_cf: Witness) -> np.ndarray:   # (b) / (c): commit the CycleFold witness, fold it
        eng = self.cf_shape.engine
        W, X = as_fr(w_cf[0], self.cf_shape.num_vars), as_fr(w_cf[1], self.cf_shape.num_io)
        U_run = self.cf_acc.instance()
        comm_W, comm_T = self.cf_acc.step_begin(W, X)
        r = nifs_challenge(eng, self.digest, U_run, R1CSInstance(comm_W, X), comm_T)
        self.cf_acc.step_end(eng.scalars([r]))
        return comm_T

    def prove_step(self, w_next: Witness, w_cfW: Witness, w_cfE: Witness) -> None:
        """One image row.  `w_next` is the augmented-circuit witness of the NEXT step (what (d) synthesises), `w_cfW` / `w_cfE`
        the witnesses of the two CycleFold instances of this step."""
        eng = self.shape.engine
        # (a) NIFS.prove on G1 with the pair staged by the previous step
        U_run = self.acc.instance()
        self.cmT = self.acc.cross_begin()
        r = nifs_challenge(eng, self.digest, U_run, self.u_i, self.cmT)
        self.acc.step_end(eng.scalars([r]))
        # (b), (c) the two CycleFold folds on G2
        self.cf_cmT = (self._fold_cyclefold(w_cfW), self._fold_cyclefold(w_cfE))
        # (d) the next fresh pair
        self._commit_fresh(w_next)
        self.i += 1

    def verify(self) -> None:
        """The satisfiability checks of `Nova::verify` ([EXT sonobe] folding/nova/mod.rs): the running pairs satisfy the relaxed
        relations on both curves, the fresh pair the plain one.  (The u_i.x = H(i, z_0, z_i, U_i) checks need the Poseidon
        transcript and the augmented circuit's public IO and are not part of this mirror.)"""
        U, W = self.acc.download()
        is_sat_relaxed(self.shape, self.srs, U, W)
        cU, cW = self.cf_acc.download()
        is_sat_relaxed(self.cf_shape, self.cf_ck, cU, cW)
        lW, lX = self.acc.fresh_witness()
        is_sat(self.shape, self.srs, R1CSInstance(self.u_i.comm_W, lX), R1CSWitness(lW))

    def close(self) -> None:
        self.acc.close()
        self.cf_acc.close()
