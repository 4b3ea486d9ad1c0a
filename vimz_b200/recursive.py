"""Two-curve mirror of nova-snark's `RecursiveSNARK` ([EXT nova-snark 0.23.0] src/lib.rs) for the data-parallel part of
`prove_step`, on top of the resident accumulators of libvimz_gpu.so.

What zero-savvy/vimz drives (/root/reference/vimz/src/nova_snark_backend/folding.rs:28-56):

    fold_input            -> create_recursive_circuit [EXT nova-scotia]: RecursiveSNARK::new, then prove_step per image row
    verify_folded_proof   -> RecursiveSNARK::verify: is_sat_relaxed (primary), is_sat_relaxed (secondary), is_sat (secondary, fresh)

`prove_step` runs, per step:
    (1) NIFS::prove on the SECONDARY curve, folding the fresh pair (l_u2, l_w2) left by the previous step
    (2) synthesis of the primary augmented circuit            -- untouched host code: the caller hands the witness in
    (3) comm_W = commit(ck1, W)                               -- r1cs_instance_and_witness
    (4) NIFS::prove on the PRIMARY curve
    (5) synthesis of the secondary circuit                     -- untouched host code
    (6) comm_W = commit(ck2, W)                               -- becomes (l_u2, l_w2) of the next step
This class holds both accumulators (r_U / r_W of both curves stay in HBM), the l_u2 / l_w2 hand-off and the step
counter, and issues the GPU work in exactly that order.  The witnesses of (2) and (5) are inputs: bellperson synthesis,
the Circom witness generator and the Poseidon RO are out of scope (BASELINE.json north_star); the challenge comes from
`TranscriptRO`, a SHAKE-256 stand-in absorbing the same values in the same order as NIFS::prove.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

from .field import as_fr
from .nova import (NUM_CHALLENGE_BITS, CommitmentEngine, CommitmentKey, FoldAccumulator, R1CSInstance, R1CSShape, R1CSWitness,
                   RelaxedR1CSInstance, RelaxedR1CSWitness, TranscriptRO, is_sat, is_sat_relaxed)

Witness = Tuple[np.ndarray, np.ndarray]   # (W, X): aux assignment and public IO of one synthesised step circuit


@dataclass
class PublicParams:
    """[EXT nova-snark] src/lib.rs `PublicParams`: the two augmented-circuit shapes and their commitment keys
    (`prepare_folding`, folding.rs:21-25, builds this once per proof)."""
    shape_primary: R1CSShape
    ck_primary: CommitmentKey
    shape_secondary: R1CSShape
    ck_secondary: CommitmentKey
    digest: int = 0   # stand-in for pp.digest() absorbed by every RO instance


def _absorb_point(eng, ro: TranscriptRO, pt: np.ndarray) -> None:
    a = eng.to_affine_ints(pt)
    ro.absorb_ints(*(a if a else (0, 0)))


def nifs_challenge(eng, digest: int, U1: RelaxedR1CSInstance, U2: R1CSInstance, comm_T: np.ndarray) -> int:
    """The absorb order of NIFS::prove ([EXT nova-snark] src/nifs.rs): pp digest, U1 (comm_W, comm_E, u, X), U2 (comm_W, X),
    comm_T; squeeze NUM_CHALLENGE_BITS.  Canonical affine coordinates / canonical scalars go in."""
    ro = TranscriptRO()
    ro.absorb_ints(digest)
    _absorb_point(eng, ro, U1.comm_W)
    _absorb_point(eng, ro, U1.comm_E)
    ro.absorb_ints(*eng.scalar_ints(U1.u), *eng.scalar_ints(U1.X))
    _absorb_point(eng, ro, U2.comm_W)
    ro.absorb_ints(*eng.scalar_ints(U2.X))
    _absorb_point(eng, ro, comm_T)
    return ro.squeeze(NUM_CHALLENGE_BITS)


class RecursiveSNARK:
    """[EXT nova-snark] src/lib.rs `RecursiveSNARK{r_W_primary, r_U_primary, r_W_secondary, r_U_secondary, l_w_secondary,
    l_u_secondary, i, ..}` with the running pairs resident on the GPU.

    overlap_secondary = False (default) keeps the strict call order of prove_step: (6) is a synchronous commit at the end of
    the step and (1) only runs the cross term and commit(T) (`vimz_acc_commit_fresh` / `vimz_acc_cross_begin`).
    overlap_secondary = True defers (6) into the next step's (1) -- one `vimz_acc_step_begin` that runs commit(W) beside
    cross term + commit(T) on two stream lanes.  l_u_secondary.comm_W is then produced lazily (a reader -- verify, the next
    RO -- forces it); nothing else in prove_step reads it earlier.  Both orders give identical values."""

    def __init__(self, pp: PublicParams, w_primary: Witness, w_secondary: Witness, overlap_secondary: bool = False):
        self.pp = pp
        self.overlap_secondary = overlap_secondary
        self.acc_primary = FoldAccumulator(pp.shape_primary, pp.ck_primary)
        self.acc_secondary = FoldAccumulator(pp.shape_secondary, pp.ck_secondary)
        # base case of RecursiveSNARK::new: the first primary instance becomes the running one
        # (RelaxedR1CSInstance::from_r1cs_instance: comm_E = identity, u = 1; from_r1cs_witness: E = 0) ...
        Wp, Xp = as_fr(w_primary[0], pp.shape_primary.num_vars), as_fr(w_primary[1], pp.shape_primary.num_io)
        eng1 = pp.shape_primary.engine
        comm_W = CommitmentEngine.commit(pp.ck_primary, Wp)
        U = RelaxedR1CSInstance(comm_W, np.zeros(12, np.uint64), Xp, eng1.scalars([1]))
        self.acc_primary.load(U, RelaxedR1CSWitness(Wp, np.zeros((pp.shape_primary.num_cons, 4), np.uint64)))
        # ... the secondary running pair starts from the default instance and the first secondary instance stays fresh
        self._pending_secondary: Optional[Witness] = None
        self.l_u_secondary: Optional[R1CSInstance] = None
        self._set_fresh_secondary(w_secondary)
        self.i = 0
        self.nifs_primary: Optional[np.ndarray] = None     # comm_T of the last NIFS on each curve (NIFS{comm_T})
        self.nifs_secondary: Optional[np.ndarray] = None

    # -- (6): the fresh secondary pair ------------------------------------------------------------------------------
    def _set_fresh_secondary(self, w: Witness) -> None:
        s2 = self.pp.shape_secondary
        W, X = as_fr(w[0], s2.num_vars), as_fr(w[1], s2.num_io)
        if self.overlap_secondary:
            self._pending_secondary, self.l_u_secondary = (W, X), None
        else:
            self.l_u_secondary = R1CSInstance(self.acc_secondary.commit_fresh(W, X), X)

    def _flush_secondary(self) -> None:
        if self._pending_secondary is not None:
            W, X = self._pending_secondary
            self.l_u_secondary = R1CSInstance(self.acc_secondary.commit_fresh(W, X), X)
            self._pending_secondary = None

    # -- prove_step ---------------------------------------------------------------------------------------------------
    def prove_step(self, w_primary: Witness, w_secondary: Witness) -> None:
        """One image row.  `w_primary` / `w_secondary` are the witnesses the untouched host code synthesised at (2) / (5).
        Like nova-snark 0.23.0, the first call only advances the counter: RecursiveSNARK::new already consumed step 0."""
        if self.i == 0:
            self.i = 1
            return
        pp = self.pp
        eng1, eng2 = pp.shape_primary.engine, pp.shape_secondary.engine
        # (1) NIFS::prove on the secondary curve with the pair left by the previous step
        U2_run = self.acc_secondary.instance()
        if self._pending_secondary is not None:      # deferred (6): commit(W) beside cross term + commit(T)
            W, X = self._pending_secondary
            comm_W, comm_T = self.acc_secondary.step_begin(W, X)
            self.l_u_secondary, self._pending_secondary = R1CSInstance(comm_W, X), None
        else:
            comm_T = self.acc_secondary.cross_begin()
        r = nifs_challenge(eng2, pp.digest, U2_run, self.l_u_secondary, comm_T)
        self.acc_secondary.step_end(eng2.scalars([r]))
        self.nifs_secondary = comm_T
        # (3) + (4): commit the fresh primary witness and NIFS::prove on the primary curve (adjacent in prove_step: one call)
        Wp, Xp = as_fr(w_primary[0], pp.shape_primary.num_vars), as_fr(w_primary[1], pp.shape_primary.num_io)
        U1_run = self.acc_primary.instance()
        comm_W, comm_T = self.acc_primary.step_begin(Wp, Xp)
        r = nifs_challenge(eng1, pp.digest, U1_run, R1CSInstance(comm_W, Xp), comm_T)
        self.acc_primary.step_end(eng1.scalars([r]))
        self.nifs_primary = comm_T
        # (6) the fresh secondary pair of the next step
        self._set_fresh_secondary(w_secondary)
        self.i += 1

    # -- verify -------------------------------------------------------------------------------------------------------
    def verify(self) -> None:
        """The three satisfiability checks of RecursiveSNARK::verify (folding.rs:53-55); raises UnSat.  The state-hash
        checks of the reference need the Poseidon RO and the augmented circuit's IO and are not part of this mirror."""
        pp = self.pp
        self._flush_secondary()
        U1, W1 = self.acc_primary.download()
        is_sat_relaxed(pp.shape_primary, pp.ck_primary, U1, W1)
        U2, W2 = self.acc_secondary.download()
        is_sat_relaxed(pp.shape_secondary, pp.ck_secondary, U2, W2)
        lW, lX = self.acc_secondary.fresh_witness()
        is_sat(pp.shape_secondary, pp.ck_secondary, R1CSInstance(self.l_u_secondary.comm_W, lX), R1CSWitness(lW))

    def close(self) -> None:
        self.acc_primary.close()
        self.acc_secondary.close()


def fold_input(pp: PublicParams, witnesses: Iterable[Tuple[Witness, Witness]], overlap_secondary: bool = False) -> RecursiveSNARK:
    """Mirror of `fold_input` (folding.rs:28-43) -> `create_recursive_circuit` [EXT nova-scotia 0.5.0]: RecursiveSNARK::new on
    the first step's circuits, then prove_step once per step (the first call is the counter-only one).  `witnesses` yields,
    per image row, the (primary, secondary) witnesses the Circom witness generator + bellperson synthesis produce."""
    it = iter(witnesses)
    first = next(it)
    snark = RecursiveSNARK(pp, first[0], first[1], overlap_secondary)
    snark.prove_step(first[0], first[1])
    for wp, ws in it:
        snark.prove_step(wp, ws)
    return snark


def verify_folded_proof(snark: RecursiveSNARK, num_steps: int) -> None:
    """Mirror of `verify_folded_proof` (folding.rs:46-56): `proof.verify(pp, iteration_count, ..)`."""
    if snark.i != num_steps:
        from .nova import UnSat
        raise UnSat(f"ProofVerifyError: proof has {snark.i} steps, expected {num_steps}")
    snark.verify()
