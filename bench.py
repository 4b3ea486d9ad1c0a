#!/usr/bin/env python3
"""bench.py -- Nova fold steps/s (HD grayscale step circuit) and Pallas MSM Mpts/s on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of nova-snark's path (oracle/)

One "step" = the data-parallel body of RecursiveSNARK::prove_step for one image row
(/root/reference/vimz/src/nova_snark_backend/folding.rs:35 -> [EXT nova-snark] NIFS::prove):
  primary curve (Pallas, grayscale_step_HD shape m=130 864 n=128 307):
      comm_W2 = commit(ck, W2); T = cross-term(A,B,C; z1, z2); comm_T = commit(ck, T); r = RO(comm_T);
      W1 += r W2; E1 += r T; (u, X) += r (1, X2); comm_W1 += r comm_W2; comm_E1 += r comm_T
  secondary curve (Vesta, ~10.5k-row shape): the same.
Witness generation, bellperson synthesis and the Poseidon RO are untouched host code in the reference
and are not part of the step (a SHAKE-256 stand-in derives r).  Inputs are synthetic (vimz_b200/synthetic.py):
the real .r1cs / witnesses cannot be produced in this image.

N > 1: one process per GPU, each rank folds an independent transformation (weak scaling, no data-path
collective -- SURVEY.md section 8e "replicas"); the MSM sweep additionally shards one MSM by point range
and combines the per-rank partial sums with an NCCL all-gather.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0xB200
K0, DK = 77, 1234577           # bases (K0 + i*DK) * G
NUM_WITNESSES = 32             # distinct fresh witnesses (one image row each) cycled through the steps
# Untimed folds before measuring.  The step cost depends (mildly) on how far the proof is: the cross-term scalars T are
# ~(128 + log2(folds))-bit values, and once they pass 135 bits (fold ~130) they spill into one more 15-bit window -- +10 %
# bucket insertions and a few giant buckets: 0.89 ms/step before, 0.95 ms/step after (measured).  An HD
# proof is 720 steps, so the timed window is placed around its median step (folds 263..463 with the default --steps 200),
# not in the cheaper first 130 folds.
SECONDARY_DIRECT_C = 14
PREFOLD = int(os.environ.get("VIMZ_BENCH_PREFOLD", "260"))   # the ncu scripts under tools/ use 32 to keep their captures short
IMAD_PER_MODMUL = 272          # 8-limb CIOS: 2*8^2 + 8 products x 2 IMAD
MODMUL_PER_MADD = 10           # XYZZ mixed add 8M + 2S


def load_peaks():
    """Roofline denominators: HBM copy bandwidth from MEASURED_PEAKS.json (driver-written), integer multiply from
    profiles/int_peak.json (tools/int_peak.cu: register-only IMAD loop with the NVML clock / power / event reasons it ran
    under recorded per test); nominal fallbacks when a file is missing."""
    peaks = {"hbm_gbs": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)", "imad_tops": 148 * 64 * 1.965e9 / 1e12,
             "imad_src": "nominal 148 SM x 64 IMAD/clk x 1.965 GHz (SURVEY.md section 8d)", "modmul_peak": None}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            peaks["hbm_gbs"] = float(json.load(open(p))["hbm_gbs"])
            peaks["hbm_src"] = "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    p = os.path.join(ROOT, "profiles", "int_peak.json")
    if os.path.exists(p):
        try:
            pk = json.load(open(p))
            tests = {t["name"]: t for t in pk["tests"]}
            t = tests["imad32"]
            peaks["imad_tops"] = float(t["Gops_per_s"]) / 1e3
            peaks["imad_src"] = ("measured 32-bit IMAD rate of tools/int_peak.cu (profiles/int_peak.json: %.0f IMAD/clk/SM at %.0f MHz%s)"
                                 % (t["ops_per_clk_per_sm"], t["eff_clock_mhz"],
                                    (", NVML %s MHz / %s W / reasons %s" % (t.get("nvml_sm_mhz"), t.get("nvml_power_w"), t.get("nvml_reasons")))
                                    if "nvml_sm_mhz" in t else ""))
            if "fp_mul_pallas_base" in tests:
                peaks["modmul_peak"] = float(tests["fp_mul_pallas_base"]["Gops_per_s"]) * 1e9
        except Exception:
            pass
    return peaks


class ClockSampler(threading.Thread):
    """SM clock / throttle-reason sampler for the timed region (B200_PROFILING.md recipe).  NVML is polled
    in-process every ~2 ms (an nvidia-smi subprocess takes longer than a short timed region); nvidia-smi is the
    fallback when pynvml is unavailable."""

    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self._stop_evt = threading.Event()
        self.max_mhz = None
        self.source = "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES-less single-node layout: NVML index == CUDA index here
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None
            self.source = "nvidia-smi"

    def _poll_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.samples.append(float(out[0]))
        self.max_mhz = float(out[1])
        for n, v in zip(names, out[2:]):
            if v.strip().lower().startswith("active"):
                self.reasons.add(n)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nv is not None:
                    nv = self._nv
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for n, bit in self.BAD.items():
                        if mask & bit:
                            self.reasons.add(n)
                    self._stop_evt.wait(0.002)
                else:
                    self._poll_smi()
                    self._stop_evt.wait(0.05)
            except Exception:
                self._stop_evt.wait(0.01)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "source": self.source}


def challenge_from(comm_T_bytes: bytes, step: int) -> int:
    """Stand-in for the Poseidon RO squeeze: 128-bit r from the commitment bytes."""
    return int.from_bytes(hashlib.shake_256(comm_T_bytes + step.to_bytes(4, "little")).digest(16), "little")


def build_problem(curve_name: str, circuit: str, seed: int, num_witnesses: int = 0):
    """Host-side synthetic shape + NUM_WITNESSES satisfying witnesses (Montgomery arrays)."""
    from vimz_host import synthetic as S       # neutral host package: does not load libvimz_gpu.so
    from vimz_host.field import CURVES, ints_to_mont
    cv = CURVES[curve_name]
    sh = S.synthetic_shape(cv, circuit, seed=seed)
    wits = []
    for k in range(num_witnesses or NUM_WITNESSES):
        Wi, Xi = S.synthetic_witness(sh, seed + 1000 + k)
        wits.append((ints_to_mont(Wi, cv.scalar_modulus), ints_to_mont(Xi, cv.scalar_modulus)))
    return cv, sh, wits


# ------------------------------------------------------------------------------------------------
# reference arm: CPU restatement of nova-snark's path (oracle/), all host cores
# ------------------------------------------------------------------------------------------------
class CpuFold:
    def __init__(self, curve_name, circuit, seed, threads, num_witnesses=0, problem=None):
        from oracle import c as oracle_c
        from vimz_host.field import affine_to_mont, ints_to_mont
        self.o = oracle_c()
        self.cv, self.sh, self.wits = problem or build_problem(curve_name, circuit, seed, num_witnesses)
        self.cid = self.cv.curve_id
        self.threads = threads
        gens = {"pallas": (self.cv.base_modulus - 1, 2), "vesta": (self.cv.base_modulus - 1, 2), "bn254": (1, 2),
                "grumpkin": (1, 17631683881184975370165255887551781615748388533673675138860)}
        g = affine_to_mont([gens[curve_name]], self.cv.base_modulus)[0]
        nck = max(self.sh.num_cons, self.sh.num_vars)
        self.bases = self.o.gen_bases(self.cid, g, K0, DK, nck)
        q = self.cv.scalar_modulus
        self.one = ints_to_mont([1], q)
        m, n, io = self.sh.num_cons, self.sh.num_vars, self.sh.num_io
        self.W1 = np.zeros((n, 4), np.uint64); self.E1 = np.zeros((m, 4), np.uint64)
        self.u1 = np.zeros((1, 4), np.uint64); self.X1 = np.zeros((io, 4), np.uint64)
        self.cW = np.zeros(12, np.uint64); self.cE = np.zeros(12, np.uint64)
        self.q = q

    def step(self, k: int):
        from vimz_host.field import ints_to_mont
        o, cid, sh, t = self.o, self.cid, self.sh, self.threads
        W2, X2 = self.wits[k % len(self.wits)]
        comm_W2 = o.msm(cid, W2, self.bases, t)
        T = o.commit_T(cid, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C, self.W1, self.u1, self.X1, W2, X2, self.one, nthreads=t)
        comm_T = o.msm(cid, T, self.bases, t)
        # r from the CANONICAL (affine) bytes of comm_T, as the reference's RO absorbs coordinates: both arms of the
        # parity replay then derive the same challenge whatever Jacobian representative their MSM returned
        r = ints_to_mont([challenge_from(self.affine(comm_T), k)], self.q)
        self.W1 = o.axpy(cid, self.W1, W2, r, t)
        self.E1 = o.axpy(cid, self.E1, T, r, t)
        tail = o.axpy(cid, np.concatenate([self.u1, self.X1]), np.concatenate([self.one, X2]), r, 1)
        self.u1, self.X1 = tail[:1], tail[1:]
        self.cW = o.point_scale_add(cid, self.cW, r, comm_W2)
        self.cE = o.point_scale_add(cid, self.cE, r, comm_T)
        self.last = (comm_W2, comm_T)

    def affine(self, jac):
        return self.o.to_affine(self.cid, jac).tobytes()


CYCLES = {"pasta": ("pallas", "vesta"), "bn254": ("bn254", "grumpkin")}


def run_cpu_steps(steps: int, warmup: int, threads: int, cycle: str = "pasta", circuit: str = "grayscale", problems=None,
                  budget_s: float = 150.0, record: bool = False):
    """`warmup` + `steps` fold steps of the CPU restatement (oracle/nova_cpu.c) from the default relaxed instance.  The
    timed steps are cut short if they would exceed `budget_s` (returned as the number actually run).  record = True also
    returns what the parity replay compares: per step the canonical bytes of (comm_W2, comm_T) of both curves, and the
    final folded pairs."""
    prim = CpuFold(CYCLES[cycle][0], circuit, SEED, threads, problem=problems[0] if problems else None)
    sec = CpuFold(CYCLES[cycle][1], "secondary", SEED + 1, threads, problem=problems[1] if problems else None)
    recs = []

    def one(k):
        sec.step(k); prim.step(k)
        if record:
            recs.append(tuple(f.affine(pt) for f in (sec, prim) for pt in f.last))

    t_w = time.perf_counter()
    for k in range(warmup):
        one(k)
    per = (time.perf_counter() - t_w) / max(warmup, 1)
    run = steps if per <= 0 else max(1, min(steps, int(budget_s / per)))
    t0 = time.perf_counter()
    for k in range(warmup, warmup + run):
        one(k)
    dt = time.perf_counter() - t0
    final = None
    if record:
        final = [(f.W1, f.E1, f.u1, f.X1, f.affine(f.cW), f.affine(f.cE)) for f in (sec, prim)]
    return run / dt, dt, prim.sh, run, recs, final


def workload_config(sh, extra=None, cycle="pasta"):
    """`sh` = the primary step shape actually folded (vimz_b200.synthetic.SyntheticShape)."""
    nck = 1 << (max(sh.num_cons, sh.num_vars) - 1).bit_length()
    prim, sec = CYCLES[cycle]
    table_mb = nck * 64 * 17 // (1 << 20)
    cfg = {"workload": ("" if cycle == "pasta" else "[BN254/Grumpkin cycle] ") +
                       f"{sh.name}_step_HD fold step: primary {prim.capitalize()} relaxed-R1CS m={sh.num_cons} n={sh.num_vars} nnz={sh.nnz} "
                       f"(ck 2^{nck.bit_length() - 1} points) + secondary {sec.capitalize()} m=n=10500; "
                       "synthetic satisfying witnesses (86% 0/1 for the pixel circuits), SHAKE-256 stand-in for the RO",
           "circuit": sh.name,
           "curve_cycle": "/".join(CYCLES[cycle]),
           "l2": f"no explicit flush: a step touches the ~{table_mb} MB window table + {sh.nnz * 36 // (1 << 20)} MB CSR + "
                 f"{(2 * sh.num_vars + 2 * sh.num_cons) * 32 // (1 << 20)} MB of vectors "
                 f"({'more' if table_mb + sh.nnz * 36 // (1 << 20) > 126 else 'LESS'} than the 126 MB L2) and every step "
                 "folds a different witness (32 distinct rows cycled)"}
    if extra:
        cfg.update(extra)
    return cfg


def main_reference(args, rank, world):
    """--impl reference: the CPU restatement of nova-snark's path on all host cores, same workload / metric / unit.  Imports
    only oracle/ and vimz_host/ (the product library is never loaded by this arm).  Rank 0 alone runs under torchrun."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    warmup = max(1, args.warmup)
    sps, dt, sh, run, _, _ = run_cpu_steps(args.steps, warmup, threads, args.cycle, args.circuit)
    sample = f"{run} full {args.circuit}_HD fold steps (primary+secondary) after {warmup} warm-up, {dt:.1f} s"
    if run != args.steps:
        sample += f" (cut from the requested {args.steps}: the CPU arm is bounded to ~150 s of timed work)"
    line = {"metric": "nova_fold_steps_per_sec", "value": sps, "unit": "steps/s", "n_gpus": args.gpus, "steps": run, "warmup": warmup,
            "ms_per_step": 1e3 / sps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (255-bit Montgomery)",
            "data": "synthetic", "impl": "reference",
            "config": workload_config(sh, cycle=args.cycle, extra={"note": "CPU restatement of nova-snark 0.23.0 (oracle/nova_cpu.c): the Rust crate cannot be built here (no cargo/rustc)"}),
            "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "loaded_native": sorted({os.path.basename(l.split()[-1]) for l in open("/proc/self/maps") if "/libvimz_gpu" in l or "/liboracle" in l})}
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class GpuFold:
    def __init__(self, curve_name, circuit, seed, device, torch, num_witnesses=0):
        import vimz_b200
        from vimz_b200 import CommitmentKey, FoldAccumulator, R1CSShape
        self.torch = torch
        self.cv, self.sh, self.wits = build_problem(curve_name, circuit, seed, num_witnesses)
        self.eng = vimz_b200.Engine(curve_name, device)
        win = os.environ.get("VIMZ_WINDOW_" + curve_name.upper())
        if win:
            self.eng.set_option("msm_window", int(win))
        seg = os.environ.get("VIMZ_SEG_MIN_" + curve_name.upper())
        if seg:
            self.eng.set_option("msm_seg_min", int(seg))
        if os.environ.get("VIMZ_CROSS_CACHE"):
            self.eng.set_option("cross_cache", int(os.environ["VIMZ_CROSS_CACHE"]))
        if os.environ.get("VIMZ_DIRECT_C"):
            self.eng.set_option("msm_direct_c", int(os.environ["VIMZ_DIRECT_C"]))
        elif circuit == "secondary":
            # the secondary curve's 10.5 k-point key keeps ALL multiples of 14-bit digits resident (19 windows instead of 26 at the
            # library's default c = 10): 105 GB of the 180 GB that the prover has no other use for, -28 us per step (measured A/B)
            self.eng.set_option("msm_direct_c", SECONDARY_DIRECT_C)
        if os.environ.get("VIMZ_DIRECT_BPS"):
            self.eng.set_option("msm_direct_bps", int(os.environ["VIMZ_DIRECT_BPS"]))
        if os.environ.get("VIMZ_DIRECT_MAX"):
            self.eng.set_option("msm_direct_max", int(os.environ["VIMZ_DIRECT_MAX"]))
        if os.environ.get("VIMZ_SPIN_WAIT"):
            self.eng.set_option("spin_wait", int(os.environ["VIMZ_SPIN_WAIT"]))
        if os.environ.get("VIMZ_ACC_BLOCKS"):
            self.eng.set_option("msm_acc_blocks", int(os.environ["VIMZ_ACC_BLOCKS"]))
        opts = os.environ.get("VIMZ_OPTS", "") + "," + os.environ.get("VIMZ_OPTS_" + ("SECONDARY" if circuit == "secondary" else "PRIMARY"), "")
        for kv in filter(None, opts.split(",")):   # A/B experiments: VIMZ_OPTS=bitrow_fold=0,graph=1 (VIMZ_OPTS_PRIMARY / _SECONDARY: one curve)
            key, val = kv.split("=")
            self.eng.set_option(key, int(val))
        sh = self.sh
        self.shape = R1CSShape(self.eng, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C)
        nck = max(sh.num_cons, sh.num_vars)
        if circuit != "secondary":   # (nova-snark sizes ck to the next power of two; bases past max(m, n) are never read, and the
            nck = 1 << (nck - 1).bit_length()   # secondary key's multiples table is 10 MB per point, so that one holds exactly what is used)
        d_bases = torch.empty(nck * 8, dtype=torch.int64, device=f"cuda:{device}")
        vimz_b200._lib.check(vimz_b200.lib.vimz_gen_bases_dev(self.eng._h, K0, DK, nck, d_bases.data_ptr()))
        self.ck = CommitmentKey.from_device(self.eng, d_bases.data_ptr(), nck)
        del d_bases
        self.acc = FoldAccumulator(self.shape, self.ck)
        # resident copies (for `value`) and pinned host copies (for `e2e`) of the fresh witnesses
        self.dev_W = [torch.from_numpy(w.view(np.int64)).to(f"cuda:{device}") for w, _ in self.wits]
        # (page-locked buffers from the library's own allocator, vimz_host_alloc: what a host keeps its witness vectors in)
        from vimz_b200.nova import pinned_fr
        self.pin_keep = [pinned_fr(w) for w, _ in self.wits]
        self.pin_W = [torch.from_numpy(a.view(np.int64).reshape(-1)) for a in self.pin_keep]
        self.q = self.cv.scalar_modulus
        # host-side constants of the loop, prepared once: device addresses, host views of the pinned witnesses, X2 as raw bytes
        self.dev_ptr = [t.data_ptr() for t in self.dev_W]
        self.pin_np = [t.numpy().view(np.uint64).reshape(-1, 4) for t in self.pin_W]
        self.X2_bytes = [np.ascontiguousarray(x, dtype=np.uint64).tobytes() for _, x in self.wits]

    def step(self, k: int, resident, acc=None):
        """resident: True = W2 already in HBM (`value`); False = pinned host buffer (`e2e`); "pageable" = plain numpy memory,
        what a Rust Vec<Scalar> is (`e2e.pageable_value`)."""
        i = k % len(self.wits)
        X2 = self.X2_bytes[i]
        acc = acc or self.acc
        if resident is True:
            cw, ct = acc.step_begin_dev(self.dev_ptr[i], X2)
        elif resident == "pageable":
            cw, ct = acc.step_begin(self.wits[i][0], X2)
        else:
            cw, ct = acc.step_begin(self.pin_np[i], X2)
        # r = RO(comm_T) as a Montgomery-form scalar, 32 little-endian bytes (what the Rust side hands over)
        r = ((challenge_from(ct.tobytes(), k) << 256) % self.q).to_bytes(32, "little")
        acc.step_end(r)

    # -- pipelined end-to-end path: the fold-independent part of the witness is handed over early -------------------------------
    def staged_split(self):
        """Rows [0, n_early) of W2 are the Circom step circuit's variables: they depend on the image row and the running hash only,
        not on the previous fold, so the host can produce and upload them ahead; the last NOVA_AUGMENTED (~10 k) variables are those of
        Nova's augmented circuit (hashes of the folded instances) and exist only once the previous step has finished."""
        from vimz_host.synthetic import NOVA_AUGMENTED
        n = self.sh.num_vars
        return max(0, n - NOVA_AUGMENTED) if n > 2 * NOVA_AUGMENTED else 0

    def stage(self, k: int, resident=False):
        i = k % len(self.wits)
        self.acc.stage_fresh(self.dev_ptr[i] if resident else self.pin_np[i], 0, self.staged_split())

    def step_staged(self, k: int, resident=False, wait=True):
        i = k % len(self.wits)
        e = self.staged_split()
        out = self.acc.step_begin_staged(self.dev_ptr[i] if resident else self.pin_np[i], e, self.sh.num_vars - e, self.X2_bytes[i], wait=wait)
        if not wait:
            return
        cw, ct = out
        self.acc.step_end(((challenge_from(ct.tobytes(), k) << 256) % self.q).to_bytes(32, "little"))

    def begin_async(self, k: int, resident=False):
        i = k % len(self.wits)
        self.acc.step_begin_async(self.dev_ptr[i] if resident else self.pin_np[i], self.X2_bytes[i])

    def finish_async(self, k: int):
        cw, ct = self.acc.step_wait()
        self.acc.step_end(((challenge_from(ct.tobytes(), k) << 256) % self.q).to_bytes(32, "little"))

    def replay_from_zero(self, nsteps: int):
        """Parity replay: `nsteps` folds from the default relaxed instance on a FRESH accumulator over the same shape / key,
        host witnesses through the C ABI, challenge from the canonical bytes of comm_T (as CpuFold does).  Returns the per-step
        canonical (comm_W2, comm_T) bytes and the final folded pair."""
        from vimz_b200 import FoldAccumulator
        acc = FoldAccumulator(self.shape, self.ck)
        recs = []
        for k in range(nsteps):
            i = k % len(self.wits)
            cw, ct = acc.step_begin(self.wits[i][0], self.X2_bytes[i])
            aw, at = self.eng.to_affine(cw).tobytes(), self.eng.to_affine(ct).tobytes()
            acc.step_end(((challenge_from(at, k) << 256) % self.q).to_bytes(32, "little"))
            recs.append((aw, at))
        U, W = acc.download()
        final = (W.W, W.E, U.u, U.X, self.eng.to_affine(U.comm_W).tobytes(), self.eng.to_affine(U.comm_E).tobytes())
        acc.close()
        return recs, final

    def close(self):
        self.dev_W = self.pin_W = self.pin_np = self.dev_ptr = self.pin_keep = None
        self.acc.close(); self.shape.close(); self.ck.close(); self.eng.close()

    def h2d_bytes(self):
        return (self.sh.num_vars + self.sh.num_io + 1 + 1) * 32

    def cross_term_bytes(self):
        """Algorithmic HBM bytes of the cross term of one step (k_matvec_stream + k_cross_finish): (col, value-index) per
        non-zero, three row pointer arrays, one pass over z2, (Az2, Bz2, Cz2) written then read, the cached (Az1, Bz1, Cz1)
        read, T written, and the recoded digit array of T written for the MSM that follows."""
        sh = self.sh
        return (sh.nnz * 8 + 3 * (sh.num_cons + 1) * 4 + (sh.num_vars + 1 + sh.num_io) * 32 + 9 * sh.num_cons * 32 + sh.num_cons * 32
                + sh.num_cons * self.ck.num_windows * 4)


class GpuFoldSharded:
    """ONE transformation folded by all ranks together (strong scaling): the primary curve's
    constraint rows, E / T and both commitment-key ranges are split across the ranks (vimz_b200.sharding.FoldShard);
    the step's only collective is the NCCL all-gather of the two partial commitments.  Every rank builds the same
    problem (same seed) and keeps W replicated."""

    def __init__(self, curve_name, circuit, seed, device, torch, dist, rank, world, comm=None, num_witnesses=0):
        import vimz_b200
        from vimz_b200 import CommitmentKey
        from vimz_b200.sharding import Comm, FoldShard, ShardedFoldAccumulator
        self.torch, self.dist, self.rank = torch, dist, rank
        self.cv, self.sh, self.wits = build_problem(curve_name, circuit, seed, num_witnesses)
        self.eng = eng = vimz_b200.Engine(curve_name, device)
        self.dev = f"cuda:{device}"
        sh = self.sh

        def make_ck(first, count):
            d = torch.empty(max(count, 1) * 8, dtype=torch.int64, device=self.dev)
            vimz_b200._lib.check(vimz_b200.lib.vimz_gen_bases_dev(eng._h, K0 + first * DK, DK, count, d.data_ptr()))
            return CommitmentKey.from_device(eng, d.data_ptr(), count)

        self.shard = FoldShard(eng, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C, make_ck, rank, world)
        # the exchange runs inside libvimz_gpu.so: its own NCCL communicator, collectives on the context's stream
        self.comm = comm if comm is not None else Comm.from_torch_dist(dist, device)
        self.acc = ShardedFoldAccumulator(dist, self.shard, eng.point_sum, device=self.dev, comm=self.comm)
        self.dev_W = [torch.from_numpy(w.view(np.int64)).to(self.dev) for w, _ in self.wits]
        self.pin_np = None
        if rank == 0:
            self.pin_W = [torch.from_numpy(w.view(np.int64).copy()).pin_memory() for w, _ in self.wits]
            self.pin_np = [t.numpy().view(np.uint64).reshape(-1, 4) for t in self.pin_W]
        self.X2_bytes = [np.ascontiguousarray(x, dtype=np.uint64).tobytes() for _, x in self.wits]
        self.q = self.cv.scalar_modulus

    def step(self, k: int, resident: bool):
        i = k % len(self.wits)
        X2 = self.X2_bytes[i]
        if resident:
            cw, ct = self.acc.step_begin_dev(self.dev_W[i].data_ptr(), X2)
        else:  # the fresh witness exists on rank 0's host only: H2D there + NCCL broadcast + step, all inside the library call
            cw, ct = self.acc.step_begin_root(self.pin_np[i] if self.rank == 0 else None, X2, root=0)
        self.acc.step_end(((challenge_from(ct.tobytes(), k) << 256) % self.q).to_bytes(32, "little"))
        return ct


def sharded_step_bench(args, torch, dist, rank, world, local_rank, steps, warmup, circuit=None, prefold=None, num_witnesses=0, expect=None,
                       comm=None):
    """Strong-scaling leg of an N > 1 run: the same grayscale (or `circuit`) proof folded by all ranks together."""
    circuit = circuit or args.circuit
    prefold = PREFOLD if prefold is None else prefold
    prim = GpuFoldSharded(CYCLES[args.cycle][0], circuit, SEED, local_rank, torch, dist, rank, world, comm=comm, num_witnesses=num_witnesses)
    if os.environ.get("VIMZ_SHARD_SECONDARY", "1") == "1":   # the secondary curve's 10.5k rows are sharded the same way
        sec = GpuFoldSharded(CYCLES[args.cycle][1], "secondary", SEED + 1, local_rank, torch, dist, rank, world, comm=prim.comm,
                             num_witnesses=num_witnesses)
    else:
        sec = GpuFold(CYCLES[args.cycle][1], "secondary", SEED + 1, local_rank, torch)
    engines = [prim.eng, sec.eng]
    last = {}

    def step_resident(k):
        sec.step(k, True); last["ct"] = prim.step(k, True)

    def step_e2e(k):
        sec.step(k, False); last["ct"] = prim.step(k, False)

    for k in range(prefold + warmup):
        step_resident(k)
    for k in range(2):
        step_e2e(k)
    ms, _ = timed_region(torch, engines, lambda k: step_resident(prefold + warmup + k), steps, dist)
    ms_e2e, _ = timed_region(torch, engines, lambda k: step_e2e(prefold + warmup + steps + k), steps, dist)
    # every rank must have derived the same transcript: compare the last comm_T across ranks
    t = torch.from_numpy(last["ct"].view(np.int64).copy()).to(f"cuda:{local_rank}")
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    same = all(bool((p == parts[0]).all()) for p in parts)
    # parity of the sharded path: the first folds of the proof again, from the default instance, with the challenge derived
    # from canonical coordinates -- every full commitment must equal the one-GPU fold of the same witnesses (`expect`, rank 0's
    # replica replayed through the plain entry points, itself checked against the CPU oracle at N = 1)
    equals_unsharded = None
    if expect is not None and isinstance(sec, GpuFoldSharded):
        ok = True
        for f, exp in ((sec, expect[0]), (prim, expect[1])):
            f.shard.acc.reset()
            for k, (ew, et) in enumerate(exp):
                i = k % len(f.wits)
                cw, ct = f.acc.step_begin_dev(f.dev_W[i].data_ptr(), f.X2_bytes[i])
                aw, at = f.eng.to_affine(cw).tobytes(), f.eng.to_affine(ct).tobytes()
                ok = ok and aw == ew and at == et
                f.acc.step_end(((challenge_from(at, k) << 256) % f.q).to_bytes(32, "little"))
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=f"cuda:{local_rank}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        equals_unsharded = bool(int(flag[0]))
    res = {"circuit": circuit, "equals_one_gpu_fold": equals_unsharded, "steps_per_s": steps / (ms * 1e-3), "ms_per_step": ms / steps, "e2e_steps_per_s": steps / (ms_e2e * 1e-3),
           "e2e_ms_per_step": ms_e2e / steps, "scaling": "strong", "ranks_agree_on_comm_T": same,
           "rows_per_rank": prim.shard.m_local, "vars_per_rank": prim.shard.var_count,
           "exchange": "inside libvimz_gpu.so (vimz_acc_step_begin_sharded*): NCCL all-gather of 2 partial commitments (192 B per rank) on the "
                       "context stream + k_point_sum_batch, one host wait per curve; e2e adds rank 0's H2D copy and the NCCL broadcast of W2 "
                       f"({prim.sh.num_vars * 32} B) in the same call",
           "parallelism": f"rows / E / T / ck of {'both curves' if isinstance(sec, GpuFoldSharded) else 'the primary curve (secondary replicated)'} "
                          f"sharded x{world} by constraint-row range, W replicated"}
    prim.shard.close()
    if isinstance(sec, GpuFoldSharded):
        sec.shard.close()
    return res


CONFIG_CIRCUITS = ["brightness", "contrast", "resize", "crop", "blur4k", "sharpness4k"]   # BASELINE.json configs 2-4 at N = 1
CONFIG_SHARDED = ["resize", "crop", "blur4k", "sharpness4k"]                                # configs 3-4: one proof over N GPUs
CONFIG_MIXED = ["grayscale", "brightness", "contrast", "resize", "crop", "blur", "sharpness", "hash"]  # config 5: one transformation per GPU
# (few distinct witnesses make T degenerate -- with 3 of them every bit row of T takes one of 4 values and the buckets collapse into
# giants: 540 instead of 760 steps/s for brightness -- so the running instance is mixed from 24 rows first)
CONFIG_STEPS, CONFIG_PREFOLD, CONFIG_WITNESSES = 10, 30, 24


def run_configs(args, torch, dist, rank, world, local_rank, sec, peaks, comm):
    """The remaining BASELINE.json configurations, bounded (10 timed steps each after 30 pre-folds over 24 distinct witnesses): sizes from
    /root/reference/circuits/nova_snark/circuit_parameters.csv:2-9 (+ Nova's augmented circuit); blur4k / sharpness4k are the x3-width
    estimates of SURVEY.md section 8 (the 4K circuits do not exist in the reference).  N = 1: every step circuit on one GPU plus the
    MSM sweep.  N > 1: resize / crop / blur4k / sharpness4k as ONE proof sharded over the N GPUs, the 2^24 MSM sharded by point
    range, and N DIFFERENT transformations folded concurrently, one per GPU (the analogue of /root/reference/benchmark.sh:25-58)."""
    prim_curve = CYCLES[args.cycle][0]
    out = {"steps_each": CONFIG_STEPS, "prefold": CONFIG_PREFOLD}
    iters = 3
    if world == 1:
        circuits = []
        for circ in CONFIG_CIRCUITS:
            f = GpuFold(prim_curve, circ, SEED, local_rank, torch, num_witnesses=CONFIG_WITNESSES)
            engines = [f.eng, sec.eng]

            def sr(k, f=f):
                sec.step(k, True); f.step(k, True)

            def se(k, f=f):
                sec.step(k, False); f.step(k, False)

            for k in range(CONFIG_PREFOLD + 3):
                sr(k)
            se(0)
            ms, _ = timed_region(torch, engines, lambda k: sr(100 + k), CONFIG_STEPS, None)
            ms_e, _ = timed_region(torch, engines, lambda k: se(200 + k), CONFIG_STEPS, None)
            circuits.append({"circuit": circ, "num_cons": f.sh.num_cons, "num_vars": f.sh.num_vars, "nnz": f.sh.nnz, "ck_log2": (len(f.ck) - 1).bit_length(),
                             "window_bits": f.ck.window_bits, "steps_per_s": CONFIG_STEPS / (ms * 1e-3), "ms_per_step": ms / CONFIG_STEPS,
                             "e2e_steps_per_s": CONFIG_STEPS / (ms_e * 1e-3), "proof_steps": {"resize": 240, "blur4k": 2160, "sharpness4k": 2160}.get(circ, 720)})
            f.close()
        out["circuits"] = circuits
        out["msm"] = [msm_bench(torch, local_rank, rank, world, dist, lg, iters, peaks, "uniform") for lg in (16, 24)]
        return out
    sharded = []
    for circ in CONFIG_SHARDED:
        sharded.append(sharded_step_bench(args, torch, dist, rank, world, local_rank, CONFIG_STEPS, 3, circuit=circ, prefold=CONFIG_PREFOLD,
                                          num_witnesses=CONFIG_WITNESSES, comm=comm))
    out["sharded_step"] = sharded
    out["msm"] = [msm_bench(torch, local_rank, rank, world, dist, 24, iters, peaks, "uniform", comm=comm)]
    # N different transformations, one per GPU, concurrently (no collective on the data path)
    circ = CONFIG_MIXED[rank % len(CONFIG_MIXED)]
    f = GpuFold(prim_curve, circ, SEED + 100 * rank, local_rank, torch, num_witnesses=CONFIG_WITNESSES)

    def sr(k):
        sec.step(k, True); f.step(k, True)

    for k in range(CONFIG_PREFOLD + 3):
        sr(k)
    stream = torch.cuda.ExternalStream(f.eng.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f.eng.sync(); sec.eng.sync(); torch.cuda.synchronize()
    dist.barrier()
    e0.record(stream)
    for k in range(CONFIG_STEPS):
        sr(100 + k)
    f.eng.sync(); sec.eng.sync()
    e1.record(stream)
    torch.cuda.synchronize()
    mine = {"rank": rank, "circuit": circ, "num_cons": f.sh.num_cons, "steps_per_s": CONFIG_STEPS / (e0.elapsed_time(e1) * 1e-3)}
    f.close()
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    out["mixed_transformations"] = {"per_gpu": allr, "aggregate_steps_per_s": sum(r["steps_per_s"] for r in allr),
                                    "note": "N different step circuits folded concurrently, one prover context per GPU (benchmark.sh:25-58)"}
    return out


class Overlapped:
    """Host sequencing of one two-curve fold step.  The host needs each curve's challenge r to produce the OTHER curve's next
    witness, not that curve's folded accumulator, so a curve's step_end (W += r W2 ... on its own streams) is issued while the
    other curve's step_begin is already running instead of in front of it: the ~25 us of launches per step leave the critical
    path.  Same calls, same order per accumulator (begin, end, begin, ...), same results."""

    def __init__(self, prim, sec):
        self.prim, self.sec, self.pend = prim, sec, None

    def step(self, k: int, resident: bool, staged: bool):
        prim, sec = self.prim, self.sec
        sec.begin_async(k, resident)
        if self.pend is not None:
            prim.acc.step_end(self.pend)
            self.pend = None
        if staged:   # the fold-independent rows of the primary witness travel while the secondary curve is folded
            prim.stage(k, resident)
        cw, ct = sec.acc.step_wait()
        r_s = ((challenge_from(ct.tobytes(), k) << 256) % sec.q).to_bytes(32, "little")
        if staged:
            prim.step_staged(k, resident, wait=False)
        else:
            prim.begin_async(k, resident)
        sec.acc.step_end(r_s)
        cw, ct = prim.acc.step_wait()
        self.pend = ((challenge_from(ct.tobytes(), k) << 256) % prim.q).to_bytes(32, "little")
        return ct

    def flush(self):
        if self.pend is not None:
            self.prim.acc.step_end(self.pend)
            self.pend = None


def timed_region(torch, engines, fn, steps, dist, finish=None):
    """barrier + sync; CUDA events on the primary context's stream around `steps` calls of fn (+ `finish`, the last deferred call of
    an overlapped sequence); max over ranks."""
    stream = torch.cuda.ExternalStream(engines[0].stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for e in engines:
        e.sync()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0.record(stream)
    t0 = time.perf_counter()
    for k in range(steps):
        fn(k)
    if finish is not None:
        finish()
    for e in engines:
        e.sync()
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, wall = float(t[0]), float(t[1]) / 1e3
    return ms, wall


def msm_bench(torch, device, rank, world, dist, log2n, iters, peaks, scalar_dist="uniform", check=True, comm=None):
    """Pallas MSM over 2^log2n resident points, uniform full-width scalars resident in HBM.  world > 1: point-range
    shards (the window is chosen per rank-local key length); the per-rank partial sums are all-gathered by NCCL on the
    context's stream and added on the GPU inside ONE library call (vimz_msm_sharded_dev), one host wait per MSM."""
    import vimz_b200
    from vimz_b200 import CommitmentEngine, CommitmentKey
    from vimz_host import synthetic as S
    eng = vimz_b200.Engine("pallas", device)
    if os.environ.get("VIMZ_WINDOW_MSM"):
        eng.set_option("msm_window", int(os.environ["VIMZ_WINDOW_MSM"]))
    if os.environ.get("VIMZ_ACC_BLOCKS"):   # A/B experiments only
        eng.set_option("msm_acc_blocks", int(os.environ["VIMZ_ACC_BLOCKS"]))
    from vimz_b200.sharding import shard_range
    n = 1 << log2n
    first, per = shard_range(n, rank, world)   # point-range shard of this rank
    d_bases = torch.empty(per * 8, dtype=torch.int64, device=f"cuda:{device}")
    vimz_b200._lib.check(vimz_b200.lib.vimz_gen_bases_dev(eng._h, K0 + first * DK, DK, per, d_bases.data_ptr()))
    ck = CommitmentKey.from_device(eng, d_bases.data_ptr(), per)
    del d_bases
    q = eng.curve.scalar_modulus

    def gen(j):  # SURVEY.md section 8d distributions: U uniform (T / E-like), B witness-like (93 % 0/1), Z edge sets
        if scalar_dist == "witness":
            return S.witness_like_scalars_mont(per, q, SEED + 7 * rank + j)
        if scalar_dist == "edge":  # thirds of zero / one / q - 1: every scalar lands in bucket 1 or nowhere (giant-bucket path)
            from vimz_host.field import ints_to_mont
            pat = ints_to_mont([0, 1, q - 1], q)
            return np.ascontiguousarray(pat[(np.arange(per) + j) % 3])
        return S.uniform_scalars_mont(per, q, SEED + 7 * rank + j)

    sc = [torch.from_numpy(gen(j).view(np.int64)).to(f"cuda:{device}") for j in range(2)]
    d_out = torch.zeros(12, dtype=torch.int64, device=f"cuda:{device}")
    stream = torch.cuda.ExternalStream(eng.stream)

    def one(j):
        if world > 1:
            return comm.commit_dev(ck, sc[j % 2].data_ptr(), per, 0)
        CommitmentEngine.commit_async_dev(ck, sc[j % 2].data_ptr(), per, d_out.data_ptr())
        return None

    for j in range(3):
        one(j)
    eng.sync()
    eng.set_option("profile", 1)
    eng.profile(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0.record(stream)
    t0 = time.perf_counter()
    for j in range(iters):
        one(j)
    eng.sync()
    e1.record(stream)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    prof = eng.profile(reset=True)
    eng.set_option("profile", 0)
    # parity of what was just timed: the commitment must be (sum_i s_i k_i) * G for the bases k_i * G -- one scalar
    # multiplication by the python big-integer oracle (the checker, not the thing measured), independent of any MSM code
    verified = None
    if check:
        from oracle import pyref as P
        from vimz_host.synthetic import closed_form_log
        c = P.CURVES["pallas"]
        local = CommitmentEngine.commit_dev(ck, sc[0].data_ptr(), per)
        exp = P.scalar_mul(c, closed_form_log(sc[0].cpu().numpy().view(np.uint64).reshape(-1, 4), K0, DK, q, first), P.generator(c))
        verified = bool(eng.to_affine_ints(local) == exp)
        if world > 1:   # every rank's shard must check out, and so must the gathered sum: the logs of all shards add up
            full = comm.commit_dev(ck, sc[0].data_ptr(), per, 0)
            logs = [None] * world
            dist.all_gather_object(logs, closed_form_log(sc[0].cpu().numpy().view(np.uint64).reshape(-1, 4), K0, DK, q, first))
            verified = verified and bool(eng.to_affine_ints(full) == P.scalar_mul(c, sum(logs) % q, P.generator(c)))
            t = torch.tensor([1 if verified else 0], dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            verified = bool(int(t[0]))
    acc_ms, acc_calls = prof["msm_accumulate_kernel"]
    entries = prof["msm_entries"][1]
    res = {"log2_points": log2n, "mpts_per_s": n * iters / ms / 1e3, "ms_per_msm": ms / iters, "window_bits": ck.window_bits,
           "result_equals_closed_form": verified,
           "windows": ck.num_windows,
           "scalars": {"uniform": "uniform 255-bit", "witness": "witness-like: 93% 0/1, 2% bytes, 5% uniform", "edge": "edge: 0 / 1 / q-1 in thirds"}[scalar_dist], "sharding": f"point-range x{world}" if world > 1 else "none"}
    if acc_ms > 0:
        imad = entries * MODMUL_PER_MADD * IMAD_PER_MODMUL
        res["accumulate_ms"] = acc_ms / max(acc_calls, 1)
        res["sort_ms"] = prof["msm_sort"][0] / max(prof["msm_sort"][1], 1)
        res["reduce_ms"] = prof["msm_reduce"][0] / max(prof["msm_reduce"][1], 1)
        res["accumulate_timad_per_s"] = imad / (acc_ms * 1e-3) / 1e12
        res["accumulate_frac_of_imad_peak"] = res["accumulate_timad_per_s"] / peaks["imad_tops"]
        res["whole_msm_timad_per_s"] = (n * iters * ck.num_windows * MODMUL_PER_MADD * IMAD_PER_MODMUL) / (ms * 1e-3) / 1e12 / world * world
    ck.close()
    eng.close()
    return res


def main_gpu(args, rank, world, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this arm has no CPU fallback (use --impl reference for the CPU restatement)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as d
        d.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist = d
    peaks = load_peaks()
    steps, warmup = args.steps, max(args.warmup, 3)
    if args.msm_only:
        msm = [msm_bench(torch, local_rank, rank, world, dist, lg, max(3, min(steps, 10)), peaks, args.msm_dist) for lg in args.msm_log2]
        if rank == 0:
            emit({"metric": "pallas_msm_mpts_per_sec", "msm": msm})
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # each rank folds its own transformation (different witnesses), same circuit
    prim = GpuFold(CYCLES[args.cycle][0], args.circuit, SEED + 100 * rank, local_rank, torch)
    sec = GpuFold(CYCLES[args.cycle][1], "secondary", SEED + 100 * rank + 1, local_rank, torch)
    engines = [prim.eng, sec.eng]

    def step_resident(k):
        sec.step(k, True); prim.step(k, True)

    def step_e2e(k):
        sec.step(k, False); prim.step(k, False)

    if os.environ.get("VIMZ_BENCH_TRACE"):   # per-50-fold wall time of the pre-folds (how the step cost moves along a proof)
        t_blk = time.perf_counter()
        for k in range(PREFOLD):
            step_resident(k)
            if (k + 1) % 50 == 0:
                for e in engines:
                    e.sync()
                now = time.perf_counter()
                st = prim.eng.lane_stats()
                print(f"[trace] folds {k - 48:4d}..{k + 1:4d}: {(now - t_blk) * 1e3 / 50:.4f} ms/step  lane0 entries {st.get('lane0_entries')} "
                      f"giants {st.get('lane0_ngiant')} chunks {st.get('lane0_nchunk')} mids {st.get('lane0_nmid')}", file=sys.stderr, flush=True)
                t_blk = now
    else:
        for k in range(PREFOLD):          # untimed: mix the running instance
            step_resident(k)
    for k in range(warmup):
        step_resident(PREFOLD + k)
    for k in range(2):
        step_e2e(k)
    ov = Overlapped(prim, sec)
    for k in range(3):
        ov.step(k, True, False)
    ov.flush()
    launches0 = sum(e.launch_count for e in engines)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, wall = timed_region(torch, engines, lambda k: ov.step(PREFOLD + warmup + k, True, False), steps, dist, finish=ov.flush)
    clocks = sampler.stop()
    launches = sum(e.launch_count for e in engines) - launches0
    ms_serial, _ = timed_region(torch, engines, lambda k: step_resident(PREFOLD + warmup + 5 * steps + k), steps, dist)
    ms_e2e, wall_e2e = timed_region(torch, engines, lambda k: step_e2e(PREFOLD + warmup + steps + k), steps, dist)

    def step_staged(k):     # pipelined e2e: the Circom part of the primary witness travels while the secondary curve is folded
        # (H2D copies share a copy engine in issue order: the secondary's own witness is enqueued first, then the staged rows)
        i = k % len(sec.wits)
        sec.acc.step_begin_async(sec.pin_np[i], sec.X2_bytes[i])
        prim.stage(k)
        cw, ct = sec.acc.step_wait()
        sec.acc.step_end(((challenge_from(ct.tobytes(), k) << 256) % sec.q).to_bytes(32, "little"))
        prim.step_staged(k)

    for k in range(2):
        step_staged(k)
    ms_staged_serial, _ = timed_region(torch, engines, lambda k: step_staged(PREFOLD + warmup + 4 * steps + k), steps, dist)
    for k in range(3):
        ov.step(k, False, True)
    ov.flush()
    ms_staged, _ = timed_region(torch, engines, lambda k: ov.step(PREFOLD + warmup + 6 * steps + k, False, True), steps, dist, finish=ov.flush)

    def step_pageable(k):   # W2 in plain (pageable) host memory, as a Rust Vec<Scalar> would be
        sec.step(k, "pageable"); prim.step(k, "pageable")

    for k in range(2):
        step_pageable(k)
    ms_page, _ = timed_region(torch, engines, lambda k: step_pageable(PREFOLD + warmup + 3 * steps + k), steps, dist)
    # Same K steps once more with the library's per-phase CUDA-event timers on (this pass launches the kernels
    # one by one instead of replaying the captured graph, so its events can sit between kernels).
    for e in engines:
        e.set_option("profile", 1)
        e.set_option("aux_lane", 0)   # one lane: kernel times not inflated by the concurrent commit(W2)
        e.profile(reset=True)
    ms_prof, _ = timed_region(torch, engines, lambda k: step_resident(PREFOLD + warmup + 2 * steps + k), steps, dist)
    prof = prim.eng.profile(reset=True)
    prof_sec = sec.eng.profile(reset=True)
    for e in engines:
        e.set_option("profile", 0)
        e.set_option("aux_lane", 1)

    value = world * steps / (ms * 1e-3)
    e2e_value = world * steps / (ms_e2e * 1e-3)
    # what the e2e path pays for: the pinned-host -> HBM copy of one fresh primary witness, timed alone
    probe_dst = torch.empty_like(prim.dev_W[0]).reshape(-1)   # pin_W is flat (vimz_host_alloc buffer), dev_W is (n, 4)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe_dst.copy_(prim.pin_W[0], non_blocking=True)
    torch.cuda.synchronize()
    pe0.record()
    for j in range(10):
        probe_dst.copy_(prim.pin_W[j % len(prim.pin_W)], non_blocking=True)
    pe1.record()
    torch.cuda.synchronize()
    h2d_us = pe0.elapsed_time(pe1) * 1e3 / 10
    h2d_gbs = prim.pin_W[0].numel() * 8 / (h2d_us * 1e-6) / 1e9
    del probe_dst

    # roofline of the dominant kernel (primary-curve bucket accumulation)
    acc_ms, acc_calls = prof["msm_accumulate_kernel"]
    entries = prof["msm_entries"][1]
    imad = entries * MODMUL_PER_MADD * IMAD_PER_MODMUL
    achieved = imad / (acc_ms * 1e-3) / 1e12 if acc_ms > 0 else None
    traffic, traffic_src = None, None
    for tname in ("r2h_traffic.json", "r2f_traffic.json", "r2_traffic.json", "r1s3_traffic.json"):   # dram__bytes of one ncu --set full capture of this kernel, per launch
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("k_msm_accumulate_fold_T")
                traffic_src = f"profiles/{tname} (ncu --set full capture of commit(T)'s launch; not measured by this run)"
                break
            except Exception:
                traffic = None
    modmul_per_s = entries * MODMUL_PER_MADD / (acc_ms * 1e-3) if acc_ms > 0 else None
    # the two launch classes of a step separately: commit(T) (1.3 M insertions, fills the GPU) and commit(W2) (0.29 M insertions of a
    # 93 % 0/1 vector: 4 per thread, latency-bound whatever the kernel does)
    accT_ms, accT_calls = prof.get("msm_accumulate_kernel_T", (0.0, 0))
    entries_T = prof.get("msm_entries_T", (0, 0))[1]
    per_class = None
    if accT_ms > 0 and acc_calls > accT_calls:
        tT = entries_T * MODMUL_PER_MADD * IMAD_PER_MODMUL / (accT_ms * 1e-3) / 1e12
        tW = (entries - entries_T) * MODMUL_PER_MADD * IMAD_PER_MODMUL / ((acc_ms - accT_ms) * 1e-3) / 1e12
        per_class = {"commit_T": {"insertions_per_launch": entries_T / accT_calls, "launch_us": accT_ms * 1e3 / accT_calls, "timad_per_s": tT,
                                  "frac": tT / peaks["imad_tops"]},
                     "commit_W2": {"insertions_per_launch": (entries - entries_T) / (acc_calls - accT_calls),
                                   "launch_us": (acc_ms - accT_ms) * 1e3 / (acc_calls - accT_calls), "timad_per_s": tW, "frac": tW / peaks["imad_tops"]}}
    roofline = {"kernel": "k_msm_accumulate<%s>" % prim.cv.name, "bound": "imad", "achieved": achieved, "peak": peaks["imad_tops"],
                "unit": "TIMAD/s", "frac": (achieved / peaks["imad_tops"]) if achieved else None, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peaks["imad_src"],
                "per_launch_class": per_class,
                "modmul_per_s": modmul_per_s, "modmul_peak_per_s": peaks.get("modmul_peak"),
                "modmul_frac": (modmul_per_s / peaks["modmul_peak"]) if modmul_per_s and peaks.get("modmul_peak") else None,
                "algorithmic": f"{entries} bucket insertions x {MODMUL_PER_MADD} modmul x {IMAD_PER_MODMUL} IMAD over {acc_calls} launches "
                               f"(commit(W2) and commit(T) of every step)",
                "launch_us_avg": (acc_ms * 1e3 / acc_calls) if acc_calls else None,
                "insertions_per_launch": (entries / acc_calls) if acc_calls else None,
                "share_of_step": acc_ms / ms_prof if ms_prof > 0 else None,
                "timing": "library CUDA-event pairs around the kernel over a profiled pass of the same K steps (stream launches); "
                          "`value` is timed separately with the step's launch sequence replayed as a CUDA graph",
                "note": "integer-multiply bound, not hbm/tensor: nothing on this path is a dense contraction (BASELINE.json north_star). "
                        "ONE denominator: the 32-bit IMAD rate measured by tools/int_peak.cu on this pool's B200 with its clock / power / "
                        "event reasons recorded beside it; modmul_frac is the same kernel against the measured field-multiplication peak "
                        "(a Pasta product executes fewer than the algorithmic 272 IMAD)"}
    ct_ms, ct_calls = prof["cross_term"]
    ct_bytes = prim.cross_term_bytes()
    ct_gbs = ct_bytes * ct_calls / (ct_ms * 1e-3) / 1e9 if ct_ms > 0 else None
    roofline_hbm = {"kernel": "k_matvec_stream + k_cross_finish (3 mat-vecs with z2: TMA bulk copy of the chunk's index stream, cp.async z gathers; then T "
                              "from the cached products of z1 + digit recoding of T; scalar field of %s)" % prim.cv.name,
                    "bound": "hbm", "achieved": ct_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": (ct_gbs / peaks["hbm_gbs"]) if ct_gbs else None, "traffic": None, "peak_source": peaks["hbm_src"],
                    "us_per_step": (ct_ms * 1e3 / ct_calls) if ct_calls else None,
                    "algorithmic": f"{ct_bytes} B per step = nnz*8 (col + coefficient index) + 3(m+1)*4 + (n+3)*32 (z2) + 9*m*32 (Az2,Bz2,Cz2 out and in, cached Az1,Bz1,Cz1 in) + m*32 (T) + m*W*4 (digits of T)",
                    "note": "gather-latency / occupancy bound, not bandwidth bound: z2 (4 MB) and the products live in the 126 MB L2, the two kernels "
                            "run ~45 us against ~7 us of pure HBM time (profiles/r2_timeline_fold_step.txt, DESIGN.md section 5)"}
    phases = {k: {"ms_per_step": v[0] / steps, "calls": v[1]} for k, v in prof.items() if not k.startswith("msm_entries")}
    phases_sec = {k: {"ms_per_step": v[0] / steps, "calls": v[1]} for k, v in prof_sec.items() if not k.startswith("msm_entries")}

    # Pallas MSM throughput (second half of the metric)
    comm = None
    if world > 1:
        from vimz_b200.sharding import Comm
        comm = Comm.from_torch_dist(dist, local_rank)   # NCCL communicator inside libvimz_gpu.so (collectives on the context streams)
    msm = []
    for lg in args.msm_log2:
        msm.append(msm_bench(torch, local_rank, rank, world, dist, lg, max(3, min(steps, 10)), peaks, args.msm_dist, comm=comm))

    sharded = None
    if world > 1 and not args.no_sharded_step:
        # what the sharded fold must reproduce: rank 0's replica (same seed as the sharded problem) replayed on ONE GPU
        expect = [None]
        if rank == 0:
            expect = [(sec.replay_from_zero(3)[0], prim.replay_from_zero(3)[0])]
        dist.broadcast_object_list(expect, src=0)
        sharded = sharded_step_bench(args, torch, dist, rank, world, local_rank, min(steps, 100), warmup, expect=expect[0], comm=comm)

    # CPU restatement on the host cores: the reported baseline AND the checker of the parity replay -- the same K = 3 (+1
    # warm-up) steps from the default relaxed instance are folded by both arms on the same witnesses and compared value by value
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        K = 3
        sps, dt, _, run, cpu_recs, cpu_final = run_cpu_steps(K, 1, threads, args.cycle, args.circuit, record=True,
                                                             problems=((prim.cv, prim.sh, prim.wits), (sec.cv, sec.sh, sec.wits)))
        cpu_baseline = {"value": sps, "unit": "steps/s", "cores": threads, "kind": "port",
                        "sample": f"{run} full {args.circuit}_HD fold steps (primary+secondary) of oracle/nova_cpu.c after 1 warm-up, {dt:.1f} s"}
        g_sec, f_sec = sec.replay_from_zero(1 + run)
        g_prim, f_prim = prim.replay_from_zero(1 + run)
        commit_eq = all(c[0] == gs[0] and c[1] == gs[1] and c[2] == gp[0] and c[3] == gp[1] for c, gs, gp in zip(cpu_recs, g_sec, g_prim))
        final_eq = all(all(np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b for a, b in zip(cf, gf))
                       for cf, gf in zip(cpu_final, (f_sec, f_prim)))
        parity = {"steps": run, "warmup_steps_also_compared": 1, "equal": bool(commit_eq and final_eq and len(cpu_recs) == 1 + run),
                  "compared": "canonical affine (comm_W2, comm_T) of every step on both curves; final W, E, u, X limb for limb and affine "
                              "comm_W, comm_E on both curves; GPU through the host-pointer C ABI vs oracle/nova_cpu.c",
                  "commitments_equal": bool(commit_eq), "folded_instances_equal": bool(final_eq)}

    configs = None
    if not args.no_configs:
        configs = run_configs(args, torch, dist, rank, world, local_rank, sec, peaks, comm)

    if rank == 0:
        bad = [r for r in clocks["reasons"] if r != "sw_power_cap"]
        line = {"metric": "nova_fold_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32x8 (255-bit Montgomery)", "data": "synthetic",
                "config": workload_config(prim.sh, cycle=args.cycle, extra={
                    "parallelism": f"replicas x{world} (one transformation per GPU)",
                    "msm_window_bits": prim.ck.window_bits, "msm_windows": prim.ck.num_windows,
                    "secondary_key": {"points": sec.ck.n if hasattr(sec.ck, "n") else None, "direct_digit_bits": sec.ck.window_bits, "windows": sec.ck.num_windows,
                                      "note": "all multiples of every digit resident (msm_direct_c): 64 B << (bits - 1) per point and window"},
                    "host_sequencing": "each curve's step_end is issued while the other curve's step_begin runs (class Overlapped); "
                                       "value_serial_calls / e2e.serial_calls_value are the same steps with begin -> end -> begin strictly in turn"}),
                "value_serial_calls": world * steps / (ms_serial * 1e-3),
                "e2e": {"value": world * steps / (ms_staged * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": prim.h2d_bytes() + sec.h2d_bytes(),
                        "d2h_bytes_per_step": 4 * 96, "ms_per_step": ms_staged / steps,
                        "api": "vimz_acc_stage_fresh + vimz_acc_step_begin_staged / _wait (primary), vimz_acc_step_begin_async / _wait (secondary), vimz_acc_step_end of "
                               "each curve issued while the other curve's step runs (class Overlapped): every "
                               "step copies its whole fresh witness host -> device inside the timed region; the fold-independent rows of the "
                               f"primary witness ({prim.staged_split()} of {prim.sh.num_vars}: the Circom step circuit's variables) are enqueued before the "
                               "secondary curve's step so the copy overlaps it, the augmented circuit's ~10 k variables go up inside step_begin",
                        "serial_calls_value": world * steps / (ms_staged_serial * 1e-3),
                        "plain_call_value": e2e_value, "plain_call_ms_per_step": ms_e2e / steps,
                        "plain_call_note": "same steps through vimz_acc_step_begin alone (whole W2 copied inside the call, nothing overlapped)",
                        "host_buffers": "pinned (cudaHostAlloc)", "pageable_value": world * steps / (ms_page * 1e-3),
                        "pageable_ms_per_step": ms_page / steps,
                        "pageable_note": "same call with W2 in ordinary pageable memory (what a Rust Vec<Scalar> is): the driver stages the copy",
                        "h2d_primary_witness_us": h2d_us, "h2d_gbs": h2d_gbs},
                "gpu_launches": int(launches),
                "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu_baseline, "parity_check": parity,
                "clocks": clocks, "clock_verdict": "rejected: " + ",".join(bad) if bad else "ok",
                "phases_primary": phases, "phases_secondary": phases_sec,
                "wall_ms_per_step": wall * 1e3 / steps, "profiled_pass_ms_per_step": ms_prof / steps,
                "msm": msm, "sharded_step": sharded, "configs": configs, "published_reference": "README-derived >= 2.99 steps/s end-to-end on a Ryzen 9 (BASELINE.md section 1), other hardware"}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries print to stdout too (NCCL's "NCCL version ..." banner under torchrun): keep fd 1 for the JSON line only.
    global _REAL_STDOUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--msm-log2", type=int, nargs="*", default=[20])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--msm-dist", default="uniform", choices=["uniform", "witness", "edge"],
                    help="scalar distribution of the MSM sweep (SURVEY.md section 8d: U / B / Z)")
    ap.add_argument("--no-sharded-step", action="store_true", help="N > 1: skip the strong-scaling leg (one proof folded by all ranks)")
    ap.add_argument("--no-configs", action="store_true", help="skip the bounded block of the other BASELINE configurations (other circuits, MSM sweep, sharded / mixed legs)")
    ap.add_argument("--msm-only", action="store_true", help="skip the fold-step measurement (window sweeps)")
    ap.add_argument("--circuit", default="grayscale", choices=["grayscale", "brightness", "contrast", "resize", "crop", "blur", "sharpness", "hash", "blur4k", "sharpness4k"],
                    help="step circuit whose published size the primary shape takes (BASELINE metric: grayscale; the others are the "
                         "remaining BASELINE configs, /root/reference/circuits/nova_snark/circuit_parameters.csv)")
    ap.add_argument("--cycle", default="pasta", choices=["pasta", "bn254"],
                    help="curve cycle: pasta = Pallas/Vesta (BASELINE metric), bn254 = BN254/Grumpkin (what the mounted vimz master instantiates)")
    args = ap.parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, world)
    else:
        main_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
