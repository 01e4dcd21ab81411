#!/usr/bin/env python3
"""bench.py — BN254 G1 MSM throughput on B200 (BASELINE.json metric: M scalar-mults/s on a 2^24-term MSM; proofs verified/s;
1/2/4/8 GPUs).

One "step" = one full pass of the hot path over the 2^24-term workload: every rank runs the Pippenger pipeline on its
contiguous chunk of the terms (util/msm.rs:322-332 shape), the 96-byte Jacobian partials are all-gathered over NCCL, and
every rank folds them and normalises (util/msm.rs:333-335 + native.rs:70).  Total work is fixed at 2^24 terms for every N
("scaling": "strong"), because that is the configuration the metric is quoted on.

  value      whole-job throughput with operands already resident in HBM (CUDA events, max over ranks)
  e2e        the same job through the C-ABI host entry points with pinned HOST buffers: H2D of scalars+points and D2H of the
             result inside the timed region
  roofline   dominant kernel (msm_bucket_accumulate[_affine]): algorithmic bytes (96 B/term) / its live CUDA-event duration vs the
             measured HBM peak — plus the integer-pipe view, because this kernel is IMAD-bound, not HBM-bound
  aux        the metric's second half and BASELINE configs 3-5 at THIS N, sharded over the ranks: proofs verified/s (4096 GWC19
             proofs, sharded by proof, one RLC + one pairing per rank), independent KZG decisions/s (4096 and 2^20 checks, check
             i -> rank i mod N), RLC-fused decide_all, config-4 aggregation jobs/s (replicas), MSM size sweep 2^10..2^26
  cpu_baseline / --impl reference
             the oracle's restatement of the reference's chunk-parallel Pippenger (util/msm.rs:308-343) on all host cores, on the
             SAME 2^24-term workload (one full-size MSM per step)
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N_DEFAULT = 24
SEED = 2024
ALG_BYTES_PER_TERM = 96            # 64 B affine point + 32 B scalar, each read once
MADD_MULMODS = 10                  # XYZZ mixed addition: 8M + 2S (csrc/g1.cuh); the 2S are SQR blocks and y3 one MUL2ADD block, i.e.
                                   # ~9.1 multiplications' worth of partial products (1160 instead of 1280) — the count below stays the formula's
AFFINE_MULMODS = 6                 # batched-affine addition: 3 for the shared inversion + 1M + 1S + 1M (csrc/bucket_affine.cuh)
IMAD_PER_MULMOD = 136              # FMA-pipe instructions per Montgomery multiplication (cuobjdump: 120 IMAD.WIDE + 8 IMAD + 8 IMAD.HI;
                                   # in the batched-affine kernel ptxas adds ~13 IMAD.X per call)
TRAFFIC_JSON = os.path.join("profiles", "r02_traffic.json")
KERNEL_SOURCES = ("msm.cu", "bucket_affine.cuh", "sort.cuh", "g1.cuh", "fp.cuh", "fp_ptx.inc")


_saved_affinity = None


def gpu_local_affinity(device_index):
    """device_index = int: bind this process to the CPU cores NVML reports as local to that GPU (its NUMA node) and return a note for
    the bench line; None: restore the affinity saved by the previous call.  Best effort: any failure leaves the process as it was."""
    global _saved_affinity
    try:
        if device_index is None:
            if _saved_affinity is not None:
                os.sched_setaffinity(0, _saved_affinity)
                _saved_affinity = None
            return None
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, v in enumerate(words) for b in range(64) if (int(v) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return "no GPU-local cores reported"
        _saved_affinity = allowed
        os.sched_setaffinity(0, cpus)
        return "pinned buffers first-touched on %d GPU-local cores of %d" % (len(cpus), len(allowed))
    except Exception as e:  # noqa: BLE001 — NVML / affinity not available: measure as is
        return "unavailable (%s)" % type(e).__name__


def workload_name(log_n):
    """The ONE description of the workload both arms (--impl cuda / reference) print in config.workload."""
    return "BN254 G1 MSM, 2^%d uniformly random scalars x points [t_i]G (seed %d)" % (log_n, SEED)


def kernel_source_hash():
    """sha256 over the sources the MSM kernels are built from: an ncu traffic capture is only quoted for the code it was taken on."""
    h = hashlib.sha256()
    for fn in KERNEL_SOURCES:
        p = os.path.join(ROOT, "snark_verifier_b200", "csrc", fn)
        if os.path.exists(p):
            with open(p, "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]


def ncu_capture(kernel, log_n, c_bits):
    """DRAM traffic and FMA-heavy pipe utilisation of `kernel` from the committed `ncu --set full` capture (tools/ncu_traffic.py
    writes profiles/r02_traffic.json) — only if it was taken on this very configuration AND on the kernel sources of this tree.
    -> (record | None, note)"""
    try:
        with open(os.path.join(ROOT, TRAFFIC_JSON)) as f:
            rec = json.load(f)[kernel]
    except Exception:
        return None, "no ncu capture committed for %s" % kernel
    if rec.get("log_n") != log_n or rec.get("window_bits") != c_bits:
        return None, "ncu capture is for another configuration (2^%s terms, c=%s)" % (rec.get("log_n"), rec.get("window_bits"))
    if rec.get("source_hash") != kernel_source_hash():
        print("bench.py: %s is STALE (kernel sources changed since the capture): roofline.traffic omitted; re-run tools/ncu_traffic.py"
              % TRAFFIC_JSON, file=sys.stderr, flush=True)
        return None, "ncu capture is older than the kernel sources (stale): omitted"
    return rec, "ncu --set full capture of this kernel on this configuration and these sources (%s)" % rec.get("source", TRAFFIC_JSON)


def per_kernel_hbm(stages, n, windows, hbm_peak, cap, acc_ms):
    """HBM view of every pipeline kernel that streams data: algorithmic bytes / live stage time vs the measured
    copy peak.  For the accumulation kernel the ALGORITHMIC bytes are the 96 B/term of the headline roofline; its `traffic_frac` is
    the DRAM traffic ncu measured per launch over the same live time (what the memory system actually sustains)."""
    rows = []

    def row(kernel, stage, bytes_per_term, what):
        ms = stages.get(stage)
        if ms:
            gbs = bytes_per_term * n / (ms / 1e3) / 1e9
            rows.append({"kernel": kernel, "ms": ms, "algorithmic_bytes_per_term": bytes_per_term, "achieved_gbs": gbs, "frac": gbs / hbm_peak, "what": what})
    row("k_digits", "msm_digits_count", 32 + 4 * windows, "small inputs: scalar read once, one 4-byte signed digit per window written (+ histogram atomics in L2)")
    row("k_scatter", "msm_digits_scatter", 8 * windows, "small inputs: digits read, 4-byte term references written at random inside one window's L2-resident region")
    row("k_digits<part>", "msm_digits", 32 + 4 * windows, "scalar read once, one 4-byte signed digit per window written; coarse-partition histogram in shared memory")
    row("k_partition", "msm_sort_partition", 10 * windows, "4-byte digits read, one 6-byte (term reference, low digit bits) record per window written into its coarse partition")
    row("k_sort_buckets", "msm_sort_buckets", 10 * windows, "6-byte records read, 4-byte term references written into their bucket runs (+ bucket counts)")
    row("k_points_prepare", "msm_points_prepare", 128, "canonical affine points read, Montgomery copy written")
    if cap:
        gbs = cap["dram_bytes_per_launch"] / (acc_ms / 1e3) / 1e9
        rows.append({"kernel": "k_bucket_accumulate*", "ms": acc_ms, "ncu_dram_bytes_per_launch": cap["dram_bytes_per_launch"], "traffic_gbs": gbs,
                     "traffic_frac": gbs / hbm_peak, "what": "measured DRAM traffic (gathers, tree levels, inversion prefixes), not algorithmic bytes"})
    return rows


def canonical_affine(b, fmt):
    """64-byte affine result in the loader's format -> canonical little-endian x || y (for the printed line only)."""
    if fmt != "montgomery":
        return b
    p = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
    rinv = pow(1 << 256, -1, p)
    return b"".join((int.from_bytes(b[i:i + 32], "little") * rinv % p).to_bytes(32, "little") for i in (0, 32))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_native_extras(threads):
    """The other two CPU legs SURVEY §8(d) names, on bounded samples: (1) the LITERAL NativeLoader MSM — the fold of
    base * scalar over the pairs, loader/native.rs:61-71, single-threaded like the reference's verifier (linear in n, so a
    2^11-term sample extrapolates); (2) `decide` per accumulator (decider.rs:70-93, G2Prepared recomputed per call) on all cores."""
    import oracle
    out = {}
    n = 1 << 11
    s, p = oracle.synth_scalars(SEED, 0, n), oracle.synth_points(SEED, 0, n, threads)
    t0 = time.perf_counter()
    oracle.msm_native(s, p, n)
    dt = time.perf_counter() - t0
    out["native_fold"] = {"value": n / dt / 1e6, "unit": "Mscalar-mults/s", "cores": 1, "kind": "port",
                          "sample": "2^11-term NativeLoader::multi_scalar_multiplication fold (native.rs:61-71) in %.2f s; O(n), no cross-term reuse" % dt}
    g2 = oracle.g2_generator()
    nchk = 64 * threads
    pts = oracle.synth_points(SEED + 1, 0, nchk, threads)
    t0 = time.perf_counter()
    oracle.kzg_decide_batch(pts, pts, nchk, g2, g2, threads)
    dt = time.perf_counter() - t0
    out["decide"] = {"value": nchk / dt, "unit": "checks/s", "cores": threads, "kind": "port",
                     "sample": "%d independent KzgAs::decide calls (decider.rs:84-93 loop) in %.2f s" % (nchk, dt)}
    return out


G2_GEN = bytes.fromhex(
    "edf692d95cbdde46ddda5ef7d422436779445c5e66006a42761e1f12efde0018c212f3aeb785e49712e7a9353349aaf1255dfb31b7bf60723a480d9293938e19"
    "aa7dfa6601cce64c7bd3430c69e7d1e38f40cb8d8071ab4aeb6d8cdba55ec8125b9722d1dcdaac55f38eb37033314bbc95330c69ad999eec75f05f58d0890609")
G1_GEN = (1).to_bytes(32, "little") + (2).to_bytes(32, "little")


class Aux:
    """Secondary measurements at N ranks.  Every leg: all ranks enter together (barrier), each works on ITS shard, local time is
    taken (CUDA events on the launch stream for device-resident legs, wall clock around the synchronous C-ABI call for host-buffer
    legs) after at least one untimed warm-up call, and the reported time is the max over ranks; a leg that throws on a rank reports
    the error instead of a number but never skips a barrier."""

    def __init__(self, sv, torch, dist, L, stream, dev, rank, world):
        self.sv, self.torch, self.dist, self.L, self.stream, self.dev, self.rank, self.world = sv, torch, dist, L, stream, dev, rank, world
        self.out = {}

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, ms, ok, extra_sum=None):
        """-> (max ms over ranks, all ok, sum of extra over ranks)"""
        torch = self.torch
        t = torch.tensor([ms if ms is not None else float("nan"), 1.0 if ok else 0.0, extra_sum or 0.0], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            mx = t.clone(); self.dist.all_reduce(mx, op=self.dist.ReduceOp.MAX)
            mn = t.clone(); self.dist.all_reduce(mn, op=self.dist.ReduceOp.MIN)
            sm = t.clone(); self.dist.all_reduce(sm, op=self.dist.ReduceOp.SUM)
            return float(mx[0]), bool(mn[1] > 0.5), float(sm[2])
        return float(t[0]), bool(t[1] > 0.5), float(t[2])

    def timed_events(self, fn, reps, warm=2):
        torch, stream = self.torch, self.stream
        with torch.cuda.stream(stream):
            for _ in range(warm):
                fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1) / reps

    def timed_wall(self, fn, reps, warm=2):
        for _ in range(warm):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        self.torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    def leg(self, name, body):
        """body() -> (local ms | None, ok, dict of fields, units for the rate, rate key, extra)"""
        try:
            self.barrier()
            res = body()
        except Exception as e:   # the headline metric must still print; every rank still reaches the reduction below
            res = (None, False, {"error": repr(e)}, 0, None)
        ms, ok, fields, units, key = res
        ms_max, ok_all, _ = self.reduce(ms, ok)
        rec = dict(fields)
        if ms is not None and ms_max == ms_max:
            rec["ms"] = ms_max
            if key:
                rec[key] = units / ms_max * 1e3
        rec["ok"] = ok_all
        self.out[name] = rec

    # ---- legs ---------------------------------------------------------------------------------------------------------------
    def run(self, do_pairing_2p20=True):
        sv, torch, L, stream, dev, rank, world = self.sv, self.torch, self.L, self.stream, self.dev, self.rank, self.world
        import numpy as np
        kz = sv.KzgAs(L, sv.KzgDecidingKey(G1_GEN, G2_GEN, G2_GEN))     # key with s = 1: accumulators (aG, aG) are valid
        n = 4096
        mine = list(range(rank, n, world))                               # check i -> rank i mod N  
        with torch.cuda.stream(stream):
            pts_all = torch.empty(n * 64, dtype=torch.uint8, device=dev)
            L.synth_points_device(SEED + 1, 0, n, pts_all.data_ptr())
            pts = pts_all.view(n, 64)[torch.tensor(mine, device=dev)].contiguous().view(-1)
            acc = torch.zeros(len(mine), dtype=torch.uint8, device=dev)
        stream.synchronize()
        nl = len(mine)

        def decide_independent():
            ms = self.timed_events(lambda: kz.decide_batch_device(pts.data_ptr(), pts.data_ptr(), nl, acc.data_ptr()), 3)
            return ms, bool(acc.min().item() == 1), {"checks": n, "what": "4096 separate 2-pair pairing checks (decider.rs:84-93), check i -> rank i mod N, "
                                                                           "operands resident in HBM"}, n, "checks_per_s"
        self.leg("decide_independent", decide_independent)

        def decide_single():
            ms = self.timed_events(lambda: kz.decide_batch_device(pts.data_ptr(), pts.data_ptr(), 1, acc.data_ptr()), 5)
            return ms, bool(acc[0].item() == 1), {"what": "one KzgAs::decide (decider.rs:70-82), operands resident; latency, not sharded"}, 1, None
        self.leg("decide_single_latency", decide_single)

        if do_pairing_2p20:
            def decide_2p20():
                N = 1 << 20
                loc = N // world
                with torch.cuda.stream(stream):
                    big = torch.empty(loc * 64, dtype=torch.uint8, device=dev)
                    L.synth_points_device(SEED + 5, rank * loc, loc, big.data_ptr())
                    rhs = big.clone()
                    v_l, v_r = big.view(loc, 64), rhs.view(loc, 64)
                    bad = torch.arange(0, loc - 1, 7, device=dev)          # every 7th check of the shard is made invalid
                    v_r[bad] = v_l[bad + 1]
                    a = torch.zeros(loc, dtype=torch.uint8, device=dev)
                ms = self.timed_events(lambda: kz.decide_batch_device(big.data_ptr(), rhs.data_ptr(), loc, a.data_ptr()), 1, warm=1)
                exp = torch.ones(loc, dtype=torch.uint8, device=dev); exp[bad] = 0
                return ms, bool((a == exp).all().item()), {"checks": N, "what": "BASELINE config 5: 2^20 independent 2-pair checks (every 7th invalid), sharded "
                                                                                "over the ranks in contiguous blocks, operands resident; accept vector verified"}, N, "checks_per_s"
            self.leg("decide_batch_2p20", decide_2p20)

        host_pts = pts.cpu().numpy()
        rho = (0x123456789ABCDEF0FEDCBA987654321).to_bytes(32, "little")

        def decide_all_fused():
            res = {}
            def call():
                res["ok"], _ = kz.decide_all_fused(host_pts, host_pts, nl, rho)
            ms = self.timed_wall(call, 3)
            return ms, bool(res["ok"]), {"accumulators": n, "what": "RLC decide_all (decider.rs:146-185): every rank fuses ITS accumulators (powers of rho + two "
                                                                    "MSMs + one pairing), verdict = AND over ranks; host buffers in, wall clock incl. H2D"}, n, "proofs_per_s"
        self.leg("decide_all_fused", decide_all_fused)

        # BASELINE config 3 with the REAL multi-open structure: 4096 GWC19 proofs of a StandardPlonk-shaped protocol (17 committed
        # polynomials opened at 3 rotations => 21-term lhs / 3-term rhs per proof) under the SRS secret s = 1 of the key
        # above, sharded by proof.  Honest proofs are built backwards from random discrete logs (commitments and opening proofs as
        # single-term MSMs on the device, outside the timed region); timed: per-proof MSM scalars by the device program compiled from
        # Gwc19::verify, one fused MSM per side (powers of rho), one pairing — per rank, on its shard.  Host buffers in, wall clock.
        m_proofs = 4096
        my_proofs = list(range(rank, m_proofs, world))

        def batch_verify_gwc19():
            import random as _random
            from snark_verifier_b200 import pcs, plonk_eval as pe
            R_MOD = sv.R_MODULUS
            npoly = 17
            omega = pe.root_of_unity(12)
            shifts = [1, omega, pow(omega, -1, R_MOD)]
            structure = [(j, shifts[j % 3]) for j in range(npoly)]
            bv = pcs.Gwc19BatchVerifier(L, kz, G1_GEN, structure, npoly)
            cpd = bv.compiled
            le32 = lambda v: (v % R_MOD).to_bytes(32, "little")
            dl, rows = [], []                                           # per proof: discrete logs of its 17 commitments + 3 W's
            for j in my_proofs:
                rnd = _random.Random((SEED + 4) * 1000003 + j)
                z, v, u = (rnd.randrange(1, R_MOD) for _ in range(3))
                c = [rnd.randrange(R_MOD) for _ in range(npoly)]
                e = [rnd.randrange(R_MOD) for _ in range(npoly)]
                w = []
                for r in range(3):
                    num = sum(pow(v, i, R_MOD) * (c[p] - e[p]) for i, p in enumerate(range(r, npoly, 3))) % R_MOD
                    w.append(num * pow((1 - shifts[r] * z) % R_MOD, -1, R_MOD) % R_MOD)      # (f(s) - eval) / (s - shift z), s = 1
                dl.append(c + w)
                rows.append(b"".join(le32(x) for x in [z, v, u] + e))
            flat = [x for d in dl for x in d]
            gp = L.msm_batch(b"".join(le32(x) for x in flat), G1_GEN * len(flat), list(range(len(flat) + 1)))
            per = npoly + 3
            slot_ix = lambda sl: None if sl == ("g",) else (sl[1] if sl[0] == "c" else npoly + sl[1])
            mp = len(my_proofs)
            pack = lambda slots: b"".join(G1_GEN if slot_ix(sl) is None else gp[j * per + slot_ix(sl)] for j in range(mp) for sl in slots)
            rows_b, lhs_pb, rhs_pb = b"".join(rows), pack(cpd.lhs_slots), pack(cpd.rhs_slots)
            rho_i = int.from_bytes(rho, "little")
            ms = self.timed_wall(lambda: kz.decide(bv.accumulate_packed(rows_b, lhs_pb, rhs_pb, mp, rho_i)), 3)   # raises unless it accepts
            return ms, True, {"proofs": m_proofs, "lhs_terms_per_proof": len(cpd.lhs_slots), "rhs_terms_per_proof": len(cpd.rhs_slots),
                              "what": "4096 honest GWC19 proofs (17 polynomials, 3 rotations), sharded by proof: per rank the MSM scalars by the device "
                                      "program compiled from Gwc19::verify + two fused MSMs (powers of rho, every point validated on the device) + one "
                                      "pairing; verdict = AND over ranks; wall clock incl. H2D"}, m_proofs, "proofs_per_s"
        self.leg("batch_verify_gwc19", batch_verify_gwc19)

        # BASELINE config 3 on REAL proofs: genuine proofs of a small hand-built PLONK protocol (tests/golden/plonk_proofs.json, made by the
        # from-scratch prover tests/plonk_toy.py) replicated to a 4096-proof batch and verified end to end from proof BYTES through
        # snark_verifier_b200.plonk.PlonkBatchVerifier: Keccak transcript challenges (device) -> per-proof program compiled from the
        # PlonkProtocol: common polynomials, instance + quotient evaluation, commitments, multi-open scalars (device) -> one fused MSM per
        # side with every proof point validated (device) -> one pairing.  Sharded by proof; verdict = AND over ranks.
        def batch_verify_plonk(scheme, transcript="evm"):
            def body():
                from snark_verifier_b200 import plonk
                with open(os.path.join(ROOT, "tests", "golden", "plonk_proofs.json")) as f:
                    fx = json.load(f)
                Hx = bytes.fromhex
                kz2 = sv.KzgAs(L, sv.KzgDecidingKey(Hx(fx["svk_g"]), Hx(fx["g2"]), Hx(fx["s_g2"])))
                try:
                    protocol = plonk.simple_plonk_protocol(fx["k"], [Hx(p) for p in fx["preprocessed"]], fx["num_public"], None, fx["initial_state"])
                    bv = plonk.PlonkBatchVerifier(L, kz2, Hx(fx["svk_g"]), protocol, scheme, transcript=transcript)
                    key = scheme if transcript == "evm" else scheme + "_" + transcript
                    good = [e for e in fx[key] if e["valid"]]
                    bad = [e for e in fx[key] if not e["valid"]][0]
                    pick = lambda e: ([[int(v) for v in col] for col in e["instances"]], Hx(e["proof"]))
                    mine = [pick(good[j % len(good)]) for j in my_proofs]
                    insts, proofs = [x[0] for x in mine], [x[1] for x in mine]
                    order = "big" if transcript == "evm" else "little"
                    if mine:
                        # the wire format of a batch: proofs and instances as contiguous byte arrays (32-byte words in the transcript's byte order)
                        proofs = np.frombuffer(b"".join(proofs), dtype=np.uint8).reshape(len(mine), -1)
                        insts = np.frombuffer(b"".join(v.to_bytes(32, order) for inst in insts for col in inst for v in col),
                                              dtype=np.uint8).reshape(len(mine), -1, 32)
                    rho_i = int.from_bytes(rho, "little")
                    res = {}

                    def call():
                        res["ok"] = bv.verify_batch(insts, proofs, rho_i)
                    ms = self.timed_wall(call, 3)
                    ok = res["ok"] is True
                    if mine:                                                # one tampered proof in the batch must flip the verdict
                        bi, bp = pick(bad)
                        if isinstance(proofs, np.ndarray):
                            p2, i2 = proofs.copy(), insts.copy()
                            p2[-1] = np.frombuffer(bp, dtype=np.uint8)
                            i2[-1] = np.frombuffer(b"".join(v.to_bytes(32, order) for col in bi for v in col), dtype=np.uint8).reshape(-1, 32)
                            ok = ok and bv.verify_batch(i2, p2, rho_i) is False
                        else:
                            ok = ok and bv.verify_batch(insts[:-1] + [bi], proofs[:-1] + [bp], rho_i) is False
                finally:
                    kz.__init__(L, kz.dk)                                   # restore the bench's deciding key on this context
                prog = bv.compiled.msm.program
                bv.close()
                return ms, ok, {"proofs": m_proofs, "scheme": scheme, "transcript": transcript, "proof_bytes": len(proofs[0]) if len(proofs) else 0,
                                "program_instructions": len(prog.instrs), "lhs_terms_per_proof": len(bv.compiled.msm.lhs_slots),
                                "what": "4096 REAL proofs (8 distinct, replicated) of a hand-built PLONK protocol, from proof bytes: device Keccak / Poseidon "
                                        "transcript + protocol-compiled scalar program + two fused MSMs (points validated on the device) + one "
                                        "pairing, sharded by proof; a batch with one tampered proof is rejected; wall clock incl. H2D and host "
                                        "packing"}, m_proofs, "proofs_per_s"
            return body
        self.leg("batch_verify_plonk_gwc19", batch_verify_plonk("gwc19"))
        self.leg("batch_verify_plonk_shplonk", batch_verify_plonk("bdfg21"))
        # the SDK's configuration: SHPLONK over the Poseidon transcript (compressed points parsed + validated on the device, Poseidon challenges on the device)
        self.leg("batch_verify_plonk_shplonk_poseidon", batch_verify_plonk("bdfg21", "poseidon"))

        # BASELINE config 3, "independent mode": every proof keeps its own accumulator — 4096 separate 21-term / 3-term MSMs
        # (snarkv_g1_msm_batch: the literal per-proof `Msm::evaluate`, util/msm.rs:81-98 -> native.rs:61-71) + 4096 separate pairing
        # checks (decider.rs:84-93).  Terms are synthetic and consistent with the key s = 1: lhs_j and rhs_j evaluate to the same point.
        def batch_verify_independent():
            mp = len(my_proofs)
            with torch.cuda.stream(stream):
                sc = torch.empty(mp * 12 * 32, dtype=torch.uint8, device=dev)
                pt = torch.empty(mp * 12 * 64, dtype=torch.uint8, device=dev)
                L.synth_scalars_device(SEED + 2, rank * mp * 12, mp * 12, sc.data_ptr())
                L.synth_points_device(SEED + 2, rank * mp * 12, mp * 12, pt.data_ptr())
            stream.synchronize()
            sc_h = sc.cpu().numpy().reshape(mp, 12, 32)
            pt_h = pt.cpu().numpy().reshape(mp, 12, 64)
            R_MOD = sv.R_MODULUS
            neg = np.empty((mp, 9, 32), dtype=np.uint8)
            for j in range(mp):                                    # r - s for the 9 cancelling pairs (host-side test-data prep)
                for k in range(9):
                    v = int.from_bytes(sc_h[j, 3 + k].tobytes(), "little")
                    neg[j, k] = np.frombuffer(((R_MOD - v) % R_MOD).to_bytes(32, "little"), dtype=np.uint8)
            lhs_s = np.concatenate([sc_h[:, :3], sc_h[:, 3:], neg], axis=1).reshape(-1)           # 3 + 9 + 9 = 21 terms
            lhs_p = np.concatenate([pt_h[:, :3], pt_h[:, 3:], pt_h[:, 3:]], axis=1).reshape(-1)
            rhs_s = np.ascontiguousarray(sc_h[:, :3]).reshape(-1)
            rhs_p = np.ascontiguousarray(pt_h[:, :3]).reshape(-1)
            lhs_off = [21 * j for j in range(mp + 1)]
            rhs_off = [3 * j for j in range(mp + 1)]
            res = {}

            def call():
                a = b"".join(L.msm_batch(lhs_s, lhs_p, lhs_off))
                b = b"".join(L.msm_batch(rhs_s, rhs_p, rhs_off))
                res["acc"], _ = kz.decide_batch(a, b, mp)
            ms = self.timed_wall(call, 2, warm=1)
            return ms, res["acc"] == b"\x01" * mp, {"proofs": m_proofs, "what": "BASELINE config 3, independent mode: 4096 separate (21-term lhs + 3-term rhs) "
                    "MSMs (one accumulator per proof) + 4096 separate pairing checks, sharded by proof; host buffers in, wall clock incl. H2D"}, m_proofs, "proofs_per_s"
        self.leg("batch_verify_independent", batch_verify_independent)

        # BASELINE config 4: one aggregation job = KzgAs::verify over 256 accumulators (accumulation.rs:41-63: two 256-term MSMs with
        # the powers of r computed on the device) + one decide (decider.rs:70-82); host buffers in, wall clock.  A single job does not
        # shard (latency-bound): REPLICAS — every rank runs its own jobs, jobs/s is the sum over ranks.
        def aggregate_256():
            n4 = 256
            reps = 5
            allp = pts_all[: n4 * 64].cpu().numpy()
            accs = [sv.KzgAccumulator(allp[64 * i:64 * i + 64].tobytes(), allp[64 * i:64 * i + 64].tobytes()) for i in range(n4)]
            ms = self.timed_wall(lambda: kz.decide(kz.verify(accs, rho)), reps)
            return ms, True, {"what": "BASELINE config 4: KzgAs::verify over 256 accumulators + decide per job; replicas only (one independent job stream per "
                                      "rank); ms = single-job latency, jobs_per_s = sum over ranks"}, world, "jobs_per_s"
        self.leg("aggregate_256_then_decide", aggregate_256)

        def plonk_scalar_eval():
            from snark_verifier_b200 import plonk_eval as pe
            proto = pe.standard_plonk_like_protocol(12, num_instance=1)
            prog = pe.compile_quotient_evaluation(proto)
            tot = proto.input_layout()["total"]
            mp = len(my_proofs)
            with torch.cuda.stream(stream):
                d_in = torch.empty(mp * tot * 32, dtype=torch.uint8, device=dev)
                d_out = torch.zeros(mp * len(prog.outputs) * 32, dtype=torch.uint8, device=dev)
                L.synth_scalars_device(SEED + 3, rank * mp * tot, mp * tot, d_in.data_ptr())
            ms = self.timed_events(lambda: L.fr_program_eval(prog, None, mp, d_inputs=d_in.data_ptr(), d_outputs=d_out.data_ptr()), 3)
            return ms, True, {"proofs": m_proofs, "instructions": len(prog.instrs),
                              "what": "StandardPlonk-shaped quotient evaluation (protocol.rs:211-283, 336-392; proof.rs:298-349) as one straight-line Fr "
                                      "program, one thread per proof, sharded by proof, operands resident in HBM"}, m_proofs, "proofs_per_s"
        self.leg("plonk_scalar_eval", plonk_scalar_eval)

        # the IPA decider over Pallas (pcs/ipa/decider.rs:47-70) — h_coeffs on the device + a 2^20-term Pallas MSM against the
        # resident committing key + comparison with u.  A single decision does not shard: replicas (every rank decides its own accumulators).
        def ipa_decide():
            from snark_verifier_b200 import pasta
            k = 20
            n = 1 << k
            PL = pasta.PallasLoader(L)
            with torch.cuda.stream(stream):
                dp = torch.empty(n * 64, dtype=torch.uint8, device=dev)
                PL.synth_points_device(SEED + 9, 0, n, dp.data_ptr())
            stream.synchronize()
            g = dp.cpu().numpy().tobytes()
            ipa = pasta.IpaAs(L, pasta.IpaDecidingKey(g))                     # uploads + validates the key once
            xi = [((7919 * (i + 1) + rank) % 65521 + 2).to_bytes(32, "little") for i in range(k)]
            u = PL.msm(PL.h_coeffs(b"".join(xi), k), g, n)                     # an accumulator that decides (outside the timed region)
            good = pasta.IpaAccumulator(xi, u)
            bad = pasta.IpaAccumulator(xi, pasta.PALLAS_GENERATOR)
            ok = ipa.decide_batch([good, bad]) == b"\x01\x00"
            ms = self.timed_wall(lambda: ipa.decide(good), 5)
            return ms, ok, {"k": k, "what": "IpaAs::decide over Pallas at k = 20 (h_coeffs + 2^20-term MSM against the resident key + compare); replicas: "
                                            "every rank decides its own accumulator; ms = latency of one decision incl. H2D of (xi, u)"}, world, "decisions_per_s"
        self.leg("ipa_decide_pallas_k20", ipa_decide)
        return self.out


def size_sweep(sv, torch, dist, Lm, stream, dev, rank, world, lo, hi):
    """BASELINE config 5: MSM size sweep 2^lo .. 2^hi at this N with the library's own plan; chunk partition + NCCL all-gather of the
    96-byte partials + fold, operands resident, best of 3 after 2 warm-ups (CUDA events, max over ranks)."""
    from snark_verifier_b200.sharding import chunk_bounds
    rows = []
    nmax = 1 << hi
    lo_max, cnt_max = chunk_bounds(nmax, world, rank)
    with torch.cuda.stream(stream):
        ds = torch.empty(max(cnt_max, 1) * 32, dtype=torch.uint8, device=dev)
        dp = torch.empty(max(cnt_max, 1) * 64, dtype=torch.uint8, device=dev)
        part = torch.zeros(96, dtype=torch.uint8, device=dev)
        parts = torch.zeros(96 * world, dtype=torch.uint8, device=dev)
        out = torch.zeros(64, dtype=torch.uint8, device=dev)
        ident = torch.zeros(96, dtype=torch.uint8, device=dev); ident[32] = 1
        if cnt_max:   # generated once: rank r's chunk of an n-term input is the first ceil(n / N) terms of its slice of the largest one
            Lm.synth_scalars_device(SEED + 7, lo_max, cnt_max, ds.data_ptr())
            Lm.synth_points_device(SEED + 7, lo_max, cnt_max, dp.data_ptr())
    for lg in range(lo, hi + 1):
        n = 1 << lg
        _, cnt = chunk_bounds(n, world, rank)
        ts = []
        for rep in range(5):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                if cnt:
                    Lm.msm_device(ds.data_ptr(), dp.data_ptr(), cnt, d_out_jacobian=part.data_ptr())
                else:
                    part.copy_(ident)
                if world > 1:
                    dist.all_gather_into_tensor(parts, part)
                    Lm.fold_partials_device(parts.data_ptr(), world, out.data_ptr())
                else:
                    Lm.fold_partials_device(part.data_ptr(), 1, out.data_ptr())
                e1.record(stream)
            stream.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep >= 2:
                ts.append(float(t.item()))
        pl = Lm.msm_plan(max(cnt, 1))
        rows.append({"log_n": lg, "ms": min(ts), "mterms_per_s": n / min(ts) / 1e3, "window_bits_per_rank": pl["window_bits"]})
    return rows


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for the path (util/msm.rs:308-343 with the `parallel` feature, restated in
    oracle/oracle.cpp) on all host threads, on the SAME workload as the CUDA arm: every step is one full 2^log_n-term MSM."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    threads = os.cpu_count() or 1
    log_n = args.ref_log_n if args.ref_log_n else args.log_n
    n = 1 << log_n
    t0 = time.perf_counter()
    s = oracle.to_mont_batch(1, oracle.synth_scalars(SEED, 0, n), n, threads)              # halo2curves in-memory layout,
    p = oracle.to_mont_batch(0, oracle.synth_points(SEED, 0, n, threads), 2 * n, threads)  # prepared outside the timed region
    gen_s = time.perf_counter() - t0
    # one full-size step costs ~10 s of all cores: the warm-up is capped at one step so that the run stays within minutes
    warm_run = min(args.warmup, 1)
    for _ in range(warm_run):
        oracle.msm_pippenger_raw(s, p, n, threads)
    t0 = time.perf_counter()
    res = None
    for _ in range(args.steps):
        res = oracle.msm_pippenger_raw(s, p, n, threads)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt / 1e6
    same = log_n == args.log_n
    sample = ("one full 2^%d-term MSM per step (the whole workload)" % log_n) if same else \
             ("2^%d-term slice of the synthetic 2^%d workload per step" % (log_n, args.log_n))
    line = {
        "impl": "reference", "metric": "BN254 G1 MSM throughput", "value": val, "unit": "Mscalar-mults/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u256 (4x64-bit Montgomery limbs)", "data": "synthetic",
        "config": {"workload": workload_name(args.log_n), "terms": 1 << args.log_n, "terms_per_step": n, "same_config_as_cuda_arm": same,
                   "byte_format": "montgomery (halo2curves in-memory layout)", "result_affine_le_hex": res.hex() if res else None,
                   "warmup_steps_run": warm_run, "input_generation_s": gen_s},
        "cpu_baseline": {"value": val, "unit": "Mscalar-mults/s", "cores": threads, "kind": "port", "sample": sample,
                         "what": "oracle restatement of util::msm::multi_scalar_multiplication with the `parallel` feature "
                                 "(util/msm.rs:308-343); the Rust reference itself cannot be built here (no cargo, halo2curves not vendored)"},
        "e2e": {"value": val, "unit": "Mscalar-mults/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--log-n", type=int, default=LOG_N_DEFAULT, help="log2 of the MSM size (metric is quoted at 24)")
    ap.add_argument("--ref-log-n", type=int, default=0, help="log2 of the per-step CPU MSM for --impl reference / cpu_baseline (0 = the full workload)")
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--format", default="montgomery", choices=["montgomery", "canonical"],
                    help="byte layout of scalars / points at the C ABI: halo2curves' in-memory Montgomery limbs (what the Rust glue passes, "
                         "zero-copy from &[Fr] / &[G1Affine]; default) or canonical little-endian `to_repr` bytes")
    ap.add_argument("--no-aux", action="store_true", help="skip the secondary KZG / scalar-evaluation measurements (profiling runs)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 2^10..2^26 MSM size sweep of the aux block")
    ap.add_argument("--sweep", default="10,26", help="lo,hi (log2) of the size sweep")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import snark_verifier_b200 as sv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    n_total = 1 << args.log_n
    from snark_verifier_b200.sharding import chunk_bounds
    lo, n_local = chunk_bounds(n_total, world, rank)   # util/msm.rs:322 chunk_size = ceil(n / threads)

    L = sv.CudaLoader(local_rank, fmt=sv.MONTGOMERY if args.format == "montgomery" else sv.CANONICAL)
    if args.window_bits:
        L.set_window_bits(args.window_bits)
    stream = torch.cuda.Stream(device=dev)
    L.set_stream(stream.cuda_stream)

    with torch.cuda.stream(stream):
        d_s = torch.empty(n_local * 32, dtype=torch.uint8, device=dev)
        d_p = torch.empty(n_local * 64, dtype=torch.uint8, device=dev)
        L.synth_scalars_device(SEED, lo, n_local, d_s.data_ptr())
        L.synth_points_device(SEED, lo, n_local, d_p.data_ptr())
        part = torch.zeros(96, dtype=torch.uint8, device=dev)
        parts = torch.zeros(96 * world, dtype=torch.uint8, device=dev)
        result = torch.zeros(64, dtype=torch.uint8, device=dev)
    stream.synchronize()

    def step_device():
        """hot path, operands resident in HBM"""
        L.msm_device(d_s.data_ptr(), d_p.data_ptr(), n_local, d_out_jacobian=part.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(parts, part)
            L.fold_partials_device(parts.data_ptr(), world, result.data_ptr())
        else:
            L.fold_partials_device(part.data_ptr(), 1, result.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident: warm-up, then exactly K timed steps with clocks sampled during the region ---------------------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
    launches0 = L.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = (L.launch_count - launches0)
    value = n_total * args.steps / (ms_total / 1e3) / 1e6
    res_dev = bytes(result.cpu().numpy())

    # ---- per-stage durations (CUDA events on the launch stream, inside the library), averaged over K more steps -------
    L.profile(True)
    stage_acc = {}
    with torch.cuda.stream(stream):
        for _ in range(args.steps):
            L.msm_device(d_s.data_ptr(), d_p.data_ptr(), n_local, d_out_jacobian=part.data_ptr())
            for name, ms, k in L.stage_times():
                a = stage_acc.setdefault(name, [0.0, 0])
                a[0] += ms; a[1] += k
    L.profile(False)
    stages = {k: v[0] / args.steps for k, v in stage_acc.items()}
    # the dominant kernel is whichever bucket-accumulation kernel the library chose for this size (batched affine for long
    # bucket lists, XYZZ otherwise — msm.cu msm_accumulate_phase)
    if "msm_bucket_accumulate_affine" in stages:
        acc_kernel, acc_ms, acc_mulmods = "k_bucket_accumulate_affine", stages["msm_bucket_accumulate_affine"], AFFINE_MULMODS
    else:
        acc_kernel, acc_ms, acc_mulmods = "k_bucket_accumulate", stages.get("msm_bucket_accumulate", float("nan")), MADD_MULMODS

    # ---- end to end: pinned host buffers through the C-ABI host entry points ------------------------------------------
    # The pinned buffers are allocated (first-touched) from the CPU cores next to this rank's GPU, so that with several ranks per
    # box every H2D copy reads the memory of its own socket; the process affinity is restored right after.
    numa_note = gpu_local_affinity(local_rank)
    h_s = torch.empty(n_local * 32, dtype=torch.uint8).pin_memory()
    h_p = torch.empty(n_local * 64, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(64, dtype=torch.uint8).pin_memory()
    gpu_local_affinity(None)
    h_s.copy_(d_s); h_p.copy_(d_p)
    torch.cuda.synchronize()

    def step_e2e():
        if world == 1:
            out = L.msm(h_s.numpy(), h_p.numpy(), n_local)          # snarkv_g1_msm: H2D + pipeline + D2H, synchronous
            h_out.numpy()[:] = np.frombuffer(out, dtype=np.uint8)
        else:
            L.msm_partial(h_s.numpy(), h_p.numpy(), n_local, part.data_ptr())   # H2D + pipeline, partial stays on device
            dist.all_gather_into_tensor(parts, part)
            L.fold_partials_device(parts.data_ptr(), world, result.data_ptr())
            h_out.copy_(result, non_blocking=True)
            stream.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(2):
            step_e2e()
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for _ in range(args.steps):
            step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_total * args.steps / float(e2e_s.item()) / 1e6
    res_e2e = bytes(h_out.numpy())

    # ---- the metric's second half ("proofs verified/s", configs 3-5) at THIS N, sharded over the ranks ------------------------
    aux = None
    if not args.no_aux:
        try:
            La = sv.CudaLoader(local_rank)            # the secondary measurements feed canonical constants: their own canonical context
            La.set_stream(stream.cuda_stream)
            aux = Aux(sv, torch, dist, La, stream, dev, rank, world).run()
            La.close()
        except Exception as e:
            aux = {"error": repr(e)}
        if not args.no_sweep:
            try:
                s_lo, s_hi = (int(x) for x in args.sweep.split(","))
                del d_s, d_p
                torch.cuda.empty_cache()
                aux["msm_size_sweep"] = {"what": "BASELINE config 5: MSM sizes 2^%d..2^%d at this N, chunk partition + NCCL all-gather of 96-byte partials + fold, "
                                                 "operands resident, best of 3 (CUDA events, max over ranks)" % (s_lo, s_hi),
                                         "rows": size_sweep(sv, torch, dist, L, stream, dev, rank, world, s_lo, s_hi)}
            except Exception as e:
                aux["msm_size_sweep"] = {"error": repr(e)}

    if rank == 0:
        hbm_peak, peak_src = peaks()
        alg_bytes = ALG_BYTES_PER_TERM * n_local
        achieved = alg_bytes / (acc_ms / 1e3) / 1e9
        plan = L.msm_plan(n_local)
        c_bits, windows = plan["window_bits"], plan["windows"]
        cap, cap_note = ncu_capture(acc_kernel, args.log_n if world == 1 else -1, c_bits)
        mulmods_per_s = n_local * windows * acc_mulmods / (acc_ms / 1e3)
        line = {
            "metric": "BN254 G1 MSM throughput", "value": value, "unit": "Mscalar-mults/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u256 (8x32-bit Montgomery limbs, integer only)", "data": "synthetic",
            "config": {"workload": workload_name(args.log_n), "terms": n_total,
                       "partition": "chunk-partitioned over %d GPU(s), one NCCL all-gather of 96-byte Jacobian partials + fold" % world,
                       "terms_per_gpu": n_local, "window_bits": c_bits, "parallelism": "chunk%d" % world,
                       "l2_policy": "inputs (%.2f GB/GPU) exceed the 126 MB L2; no flush needed" % (n_local * 96 / 1e9),
                       "byte_format": args.format + (" (halo2curves in-memory layout, as the reference's CPU arm is fed)" if args.format == "montgomery" else ""),
                       "result_affine_le_hex": canonical_affine(res_dev, args.format).hex()},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mscalar-mults/s", "h2d_bytes_per_step": n_local * 96 * world, "d2h_bytes_per_step": 64 * world,
                    "api": "snarkv_g1_msm (N=1) / snarkv_g1_msm_partial + all_gather + fold (N>1), pinned host buffers",
                    "host_numa": numa_note,
                    "result_matches_device_path": res_e2e == res_dev},
            "gpu_launches": int(launches),
            "roofline": {"kernel": acc_kernel, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": cap["dram_bytes_per_launch"] if cap else None, "traffic_source": cap_note,
                         "peak_source": peak_src, "kernel_ms": acc_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "this kernel is bound by the integer multiplier (FMA-heavy pipe), not by HBM: %d point additions x %d "
                                 "Montgomery multiplications x ~%d FMA-pipe instructions per term; 'traffic' is each 64-byte point gathered once "
                                 "per window (plus, for the batched-affine kernel, the intermediate tree levels and the inversion prefixes)"
                                 % (windows, acc_mulmods, IMAD_PER_MULMOD),
                         "per_kernel_hbm": per_kernel_hbm(stages, n_local, windows, hbm_peak, cap, acc_ms),
                         "alu": {"mulmods_per_s": mulmods_per_s,
                                 "fmaheavy_pipe_pct_of_peak_ncu": cap["fmaheavy_pct"] if cap else None,
                                 "source": cap_note}},
            "stages_ms": stages,
            "aux": aux,
        }
        if not args.no_cpu_baseline and world == 1:
            # the SAME workload on the host cores: the pinned e2e buffers hold exactly the bytes the GPU arm consumed (Montgomery =
            # halo2curves' in-memory layout, which is what the Rust reference holds; canonical bytes are converted outside the timing)
            import oracle
            threads = os.cpu_count() or 1
            try:
                log_c = args.ref_log_n if args.ref_log_n else args.log_n
                n_c = 1 << log_c
                hs, hp = h_s.numpy()[: n_c * 32], h_p.numpy()[: n_c * 64]
                if args.format != "montgomery":
                    hs = np.frombuffer(oracle.to_mont_batch(1, hs, n_c, threads), dtype=np.uint8)
                    hp = np.frombuffer(oracle.to_mont_batch(0, hp, 2 * n_c, threads), dtype=np.uint8)
                t0 = time.perf_counter()
                res_cpu = oracle.msm_pippenger_raw(hs, hp, n_c, threads)
                secs = time.perf_counter() - t0
                line["cpu_baseline"] = {"value": n_c / secs / 1e6, "unit": "Mscalar-mults/s", "cores": threads, "kind": "port",
                                        "sample": "one %s chunk-parallel Pippenger (util/msm.rs:308-343 restated) over 2^%d terms in %.2f s"
                                                  % ("FULL-SIZE" if log_c == args.log_n else "reduced", log_c, secs),
                                        "result_matches_gpu": (res_cpu == canonical_affine(res_dev, args.format)) if log_c == args.log_n else None}
                line["cpu_baseline"]["also"] = cpu_native_extras(threads)
            except Exception as e:
                line["cpu_baseline"] = {"error": repr(e)}
        elif world > 1:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    L.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
