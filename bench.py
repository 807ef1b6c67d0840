#!/usr/bin/env python3
"""Benchmark of the census Groth16 proving path (BASELINE.json metric: census Groth16 proofs/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one process per GPU
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

A "step" is one pass of the hot path (witness -> H -> 5 MSMs -> proof) over one batch of synthetic census
inputs (BASELINE.json configs[1]: census.circom at nLevels=160, 1,024 proofs per GPU, 1,024-voter synthetic
census, seed 0xC0FFEE).  Multi-GPU runs shard independent proofs across ranks (no data-path collective): every
rank proves its own 1,024-proof batch, so the scaling is "weak"; `value` = proofs of all ranks / max-over-ranks
time.

Printed JSON line (rank 0): value = device-resident throughput (inputs already in HBM), e2e = the same metric
through the reference-shaped call (inputs.json strings in host memory -> proof.json strings in host memory,
H2D/D2H inside the timed region), roofline for the dominant kernel, cpu_baseline = the CPU oracle on this box's
host cores on a bounded sample.  Only the cpu_baseline leg and --impl reference touch oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
ART = os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160")
METRIC = "census Groth16 proofs/sec"
UNIT = "proofs/s"
WORKLOAD = "census.circom nLevels=160 (82,754 wires, domain 2^17), batch of 1024 proofs per GPU, synthetic 1024-voter census seed 0xC0FFEE"
MODMUL_PER_MADD_G1 = 10      # XYZZ mixed add: 8M + 2S
MODMUL_PER_AFFINE_ADD = 6    # batched-affine add: prefix product, 2 for the shared inverse, lambda, lambda^2, y3
MODMUL_PER_INVERSION = 256 + 110   # Fermat: square-and-multiply over the 256-bit exponent q - 2 (110 set bits)


# ---- helpers shared with the CPU (gloo) tests ---------------------------------------------------------
def shard_range(total, rank, world):
    """Contiguous proof range of `rank` (proofs dealt contiguously per GPU)."""
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def _reduce(x, device, op):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(x, device):
    import torch.distributed as dist
    return _reduce(x, device, dist.ReduceOp.MAX)


def sum_over_ranks(x, device):
    import torch.distributed as dist
    return _reduce(x, device, dist.ReduceOp.SUM)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)) if os.path.exists(p) else {}


# ---- the CPU arm (oracle) ------------------------------------------------------------------------------
DENSE_MODMUL_PER_PROOF = 1.14e8   # G1 6.44e7 + G2 4.25e7 + NTT 7.2e6 + witness 2.1e5 (what snarkjs / rapidsnark execute)


def _cpu_witness_one(inp):
    """(worker process) witness of one voter by the reference circuit.wasm compiled to native code"""
    import ref_witness as RW
    code, w = RW.witness(inp)
    return code, (w.tobytes() if code == 0 else b"")


def probe_reference_toolchains():
    """time snarkjs / go test if the toolchains ever exist on the box.  Returns what was found."""
    import shutil
    found = {k: shutil.which(k) for k in ("node", "snarkjs", "go", "rapidsnark", "circom")}
    return {k: v for k, v in found.items() if v}


def reference_toolchain_run(sample):
    """If node + snarkjs are on the box: `snarkjs groth16 fullprove` on the reference's own fixture, wall time per proof
    (ts_inputs/src/example.ts:358-365 times the same call).  None when the toolchain is absent (this image: absent)."""
    import shutil
    if not (shutil.which("node") and shutil.which("snarkjs")):
        return None
    inp = os.path.join(ROOT, "tests", "golden", "inputs_example.json")
    out = os.path.join("/tmp", "zkb_snarkjs_proof.json"), os.path.join("/tmp", "zkb_snarkjs_public.json")
    t0 = time.perf_counter()
    for _ in range(sample):
        subprocess.check_call(["snarkjs", "groth16", "fullprove", inp, os.path.join(ART, "circuit.wasm"),
                               os.path.join(ART, "proving_key.zkey"), out[0], out[1]], stdout=subprocess.DEVNULL)
    return sample / (time.perf_counter() - t0)


def cpu_prove_sample(n_proofs, inputs=None, shortcut=False):
    """Times the CPU oracle (reference wasm witness when oracle/_ref is present + restated snarkjs groth16.prove,
    OpenMP over all host threads) on n_proofs of the workload.  The witnesses of the sample run in parallel worker
    processes (one wasm instance each), the proofs one after another with every thread inside each proof.
    shortcut = True: the four witness MSMs use the same proof-independent-wire shortcut as the GPU path (template =
    the witness of another voter of the same census, its sums cached outside the timing).
    Returns (proofs/s, cores, kind, description)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import multiprocessing as mp
    import numpy as np
    import oracle_lib as O
    import ref_witness as RW
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host thread it can
    ncpu = len(os.sched_getaffinity(0))
    O.lib().orc_set_threads(ncpu)
    zk = O.ZKeyRef(open(os.path.join(ART, "proving_key.zkey"), "rb").read())
    if inputs is None:
        inputs = [json.load(open(os.path.join(ROOT, "tests", "golden", "inputs_example.json")))]
    use_wasm = RW.available()
    if not use_wasm:
        import census_model as M
        import wasm_tools as W
        mod = W.Module(open(os.path.join(ART, "circuit.wasm"), "rb").read())
        tables, wmap = W.poseidon_tables(mod), W.witness_map(mod)[0]
    zk.prove(np.zeros((zk.n_vars, 32), dtype=np.uint8), 1, 1)       # parse + cache the key outside the timing
    sample = [inputs[i % len(inputs)] for i in range(n_proofs)]
    tmpl = None
    if shortcut:
        code, tmpl = RW.witness(inputs[-1]) if use_wasm else (0, None)
        if tmpl is None:
            raise RuntimeError("shortcut leg needs the compiled reference wasm")
        zk.prove_shortcut(tmpl, tmpl, 1, 1)                          # template sums cached outside the timing
    workers = min(ncpu, n_proofs) if use_wasm else 1
    pool = mp.get_context("fork").Pool(workers) if workers > 1 else None
    try:
        if pool:
            pool.map(_cpu_witness_one, sample[:workers])             # start the workers, load the wasm library
        t0 = time.perf_counter()
        if use_wasm:
            res = pool.map(_cpu_witness_one, sample) if pool else [_cpu_witness_one(x) for x in sample]
            assert all(c == 0 for c, _ in res)
            ws = [np.frombuffer(b, dtype=np.uint8).reshape(-1, 32) for _, b in res]
        else:
            ws = []
            for inp in sample:
                vals = M.witness(tables, wmap, inp)
                ws.append(np.frombuffer(b"".join(x.to_bytes(32, "little") for x in vals), dtype=np.uint8).reshape(-1, 32))
        t_w = time.perf_counter() - t0
        for i, w in enumerate(ws):
            if shortcut:
                zk.prove_shortcut(w, tmpl, 1234567 + i, 7654321 + i)
            else:
                zk.prove(w, 1234567 + i, 7654321 + i)
        dt = time.perf_counter() - t0
    finally:
        if pool:
            pool.terminate()
    cpu_prove_sample.last_seconds = dt
    cpu_prove_sample.last_witness_seconds = t_w
    cores = O.lib().orc_threads()
    desc = (f"{n_proofs} proofs of the workload (voters of the same synthetic census): witness by the reference "
            f"circuit.wasm compiled to native code ({workers} worker processes, {t_w:.2f} s of the {dt:.2f} s), Groth16 by "
            f"the C++ restatement of snarkjs groth16.prove (Pippenger + radix-2 NTT, OpenMP {cores} threads)"
            + (", four witness MSMs over (w - template) only [same shortcut as the GPU path]" if shortcut else
               ", dense (all 82,754 wires in A/B1/B2/C, as snarkjs / rapidsnark do)")) if use_wasm else \
        f"{n_proofs} proofs, python witness model + C++ Groth16 restatement"
    return n_proofs / dt, cores, "port", desc


def reference_voters(n):
    """The first n voters of the benchmark's 1,024-voter synthetic census (seed 0xC0FFEE), generated on the host by the
    oracle's generator - identical to what the GPU arm's rank 0 proves (tests/test_gpu_prover.py checks the two
    generators against each other)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    return list(H.voters(1024))


def run_reference(args, rank):
    if rank != 0:
        return
    sample = int(os.environ.get("ZKB_REF_SAMPLE", "4"))
    voters = reference_voters(1024)
    for _ in range(args.warmup):
        cpu_prove_sample(1, voters[:1])
    rate = cores = kind = desc = None
    dt = 0.0
    for k in range(args.steps):
        lo = (k * sample) % len(voters)
        rate, cores, kind, desc = cpu_prove_sample(sample, (voters[lo:] + voters[:lo])[:sample])
        dt += cpu_prove_sample.last_seconds          # proving only: the key is parsed once outside the timing
    value = args.steps * sample / dt
    probe = probe_reference_toolchains()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u256 (4x64-bit Montgomery, host)", "data": "synthetic",
            "config": bench_config(voters, args.batch),
            "step": f"{sample} proofs per step (bounded sample of the 1024-proof batch: voters of the same census)",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "toolchain_probe": probe or "none of node / snarkjs / go / rapidsnark / circom on this box",
            "note": "snarkjs / rapidsnark are not runnable here (no node, go or wasm runtime in the image): this is the "
                    "repo's CPU restatement of the same algorithm, cpu_restatement (not snarkjs)"}
    snark = reference_toolchain_run(1)
    if snark:
        line["snarkjs_proofs_per_s"] = snark
    print(json.dumps(line), flush=True)


# ---- the GPU arm -----------------------------------------------------------------------------------------
def tree_depth(inputs):
    """last non-zero sibling + 1, maximum over the census and SIK trees (what SMTLevIns computes in the circuit)"""
    d = 0
    for k in ("censusSiblings", "sikSiblings"):
        nz = [i for i, x in enumerate(inputs[k]) if int(x) != 0]
        d = max(d, nz[-1] + 1 if nz else 0)
    return d


def bench_config(voters, batch):
    """`config` of the JSON line - the same for both arms (the reference arm proves voters of this census)."""
    d = sorted(tree_depth(v) for v in voters)
    return {"workload": WORKLOAD, "proofs_per_gpu_per_step": batch,
            "tree_depth": {"min": d[0], "max": d[-1], "median": d[len(d) // 2],
                           "what": "last non-zero sibling + 1 (census and SIK trees) over the 1024 voters; the circuit "
                                   "hashes all 161 levels, the levels below the leaf are proof-independent (SURVEY 8a W7)"},
            "l2": "per-step working set (witnesses 2.6 MB + A/B/C vectors 12 MB per proof, >10 GB per step) "
                  "exceeds the 126 MB L2; no flush needed"}


def timed_resident(c, packed, n, steps, stream, torch):
    """device-resident proofs/s of circuit c over n proofs (inputs replicated from `packed`), CUDA events"""
    import numpy as np
    reps = (n + packed.shape[0] - 1) // packed.shape[0]
    c.set_inputs(np.concatenate([packed] * reps)[:n])
    c.prove_resident(n)                                                   # warm-up (workspace, caches)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        c.prove_resident(n)
    e1.record(stream)
    torch.cuda.synchronize()
    _, _, status = c.get_results(n)
    assert (status == 0).all(), "a synthetic proof failed its circuit asserts"
    return n * steps / (e0.elapsed_time(e1) * 1e-3)


def multi_gpu_selfcheck(rank, world, local_rank, dist, torch):
    """N > 1, outside every timed region, no oracle: the three P2P paths of SURVEY.md 8e run across all N ranks and are
    compared with committed known answers / their single-GPU results.
      (a) BASELINE configs[3] shape: the 600-link chain circuit (artifacts/chain600, generated by build()) proved
          through zkb_load_circuit_shard on all ranks, proof == tests/golden/chain600_kat.json
      (b) one 2^20-point G1 MSM split by point range over the ranks (zkb_msm_session) == the same MSM on one rank
      (c) one 2^24-point 4-step NTT over the ranks (zkb_ntt_dist) == the single-GPU plan
    plus timings of the config-4-shaped proof and of the 2^26 MSM / NTT (BASELINE configs[4])."""
    import hashlib
    import numpy as np
    from zk_franchise_proof_circuit_b200 import prover, raw
    out = {}

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def bcast(obj, src=0):
        box = [obj]
        dist.broadcast_object_list(box, src=src)
        return box[0]

    def gather(obj):
        got = [None] * world
        dist.all_gather_object(got, obj)
        return got

    def rmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- (a) sharded proof of the chain circuit --------------------------------------------------------
    cdir = os.path.join(ROOT, "artifacts", "chain600")
    kat_path = os.path.join(ROOT, "tests", "golden", "chain600_kat.json")
    try:
        kat = json.load(open(kat_path))
        zkey = open(os.path.join(cdir, "proving_key.zkey"), "rb").read()
        wtns = open(os.path.join(cdir, "witness.wtns"), "rb").read()
        vkey = open(os.path.join(cdir, "verification_key.json"), "rb").read()
        assert hashlib.sha256(zkey).hexdigest() == kat["zkey_sha256"], "chain600 zkey differs from the committed KAT's"
        assert hashlib.sha256(wtns).hexdigest() == kat["wtns_sha256"], "chain600 witness differs from the committed KAT's"
        c = prover.load_shard(zkey, rank, world, device=local_rank)
        h = bcast(c.shard_export() if rank == 0 else None)
        if rank != 0:
            c.shard_attach(h)
        whs = gather(c.shard_export_witness())               # witness slices: each rank uploads 1 / N of the .wtns
        for g in range(world):
            if g != rank:
                c.shard_attach_witness(g, whs[g])
        c.set_blinding(kat["r"], kat["s"])
        barrier()
        c.prove_wtns(wtns)                                   # warm-up: allocates the workspace
        walls, res = [], []
        pj = sj = None
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            pj, sj = c.prove_wtns(wtns)
            walls.append((time.perf_counter() - t0) * 1e3)
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            c.prove_wtns(None)
            res.append((time.perf_counter() - t0) * 1e3)
        wall, resident = rmax(sorted(walls)[1]), rmax(sorted(res)[1])
        if rank == 0:
            same = json.loads(pj) == kat["proof"] and json.loads(sj) == kat["public"]
            prover.verify(vkey, sj, pj)
            out["chain600_sharded_proof"] = {"ok": bool(same), "verified": True, "ranks": world,
                                             "wires": c.n_vars, "domain_log2": c.domain.bit_length() - 1,
                                             "proof_wall_ms_p50": wall, "proof_resident_wall_ms_p50": resident,
                                             "kat": "tests/golden/chain600_kat.json (proof with pinned r, s == CPU oracle)"}
        barrier()
        c.close()
    except Exception as e:  # noqa: BLE001
        out["chain600_sharded_proof"] = {"ok": False, "error": f"{type(e).__name__}: {e}"}
    # ---- (b) MSM split by point range ------------------------------------------------------------------
    for logn, key in ((20, "msm_2p20_split"), (26, "msm_2p26_split")):
        try:
            s = raw.MsmSession(logn, rank, world, device=local_rank, seed=1, window=16)
            h = bcast(s.export_handle() if rank == 0 else None)
            if rank != 0:
                s.attach(h)
            point, times = None, []
            for it in range(3):
                barrier()
                if rank == 0:
                    s.run()
                    point, ms = s.combine(world)
                else:
                    ms = s.run(wait=True)
                barrier()
                if it:
                    times.append(rmax(ms))
            s.close()
            barrier()
            ent = {"ranks": world, "ms": min(times), "points_per_s": (1 << logn) / (min(times) * 1e-3)}
            if logn == 20 and rank == 0:                     # the same 2^20 points on one rank (same global index space)
                one = raw.MsmSession(logn, 0, 1, device=local_rank, seed=1, window=16)
                one.run()
                p1, ms1 = one.combine(1)
                one.close()
                ent.update({"ok": bool(point == p1 and point != bytes(64)), "single_gpu_ms": ms1,
                            "result_x_hex": point[:32][::-1].hex()})
            barrier()
            if rank == 0:
                out[key] = ent
        except Exception as e:  # noqa: BLE001
            out[key] = {"ok": False, "error": f"{type(e).__name__}: {e}"}
    # ---- (c) 4-step NTT over the ranks -----------------------------------------------------------------
    for logn, key in ((24, "ntt_2p24_4step"), (26, "ntt_2p26_4step")):
        try:
            s = raw.NttDist(logn, rank, world, device=local_rank)
            hs = gather(s.export_handle())
            for g in range(world):
                if g != rank:
                    s.attach(g, hs[g])
            rows = []
            for it in range(3):
                s.fill()
                barrier()
                s.run()
                ms = s.sync()
                barrier()
                if it:
                    rows.append([rmax(x) for x in ms])
            ent = {"ranks": world, "ms": min(r[0] for r in rows), "columns_ms": rows[-1][1],
                   "wait_plus_p2p_exchange_ms": rows[-1][2], "rows_ms": rows[-1][3]}
            if logn == 24:
                digests = gather(hashlib.sha256(s.read(1).tobytes()).hexdigest())
                s.close()
                barrier()
                if rank == 0:                                # the same input on one rank: the single-GPU plan
                    one = raw.NttDist(logn, 0, 1, device=local_rank)
                    one.fill()
                    one.run()
                    one.sync()
                    full = one.read(1)
                    one.close()
                    per = full.shape[0] // world
                    want = [hashlib.sha256(full[g * per:(g + 1) * per].tobytes()).hexdigest() for g in range(world)]
                    ent["ok"] = bool(want == digests)
            else:
                s.close()
            barrier()
            if rank == 0:
                out[key] = ent
        except Exception as e:  # noqa: BLE001
            out[key] = {"ok": False, "error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        checks = [out[k].get("ok") for k in ("chain600_sharded_proof", "msm_2p20_split", "ntt_2p24_4step")]
        out["all_ok"] = bool(all(checks))
    return out


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from zk_franchise_proof_circuit_b200 import prover, census_tree, raw

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    # stdout carries exactly ONE JSON line: native libraries that print to fd 1 (NCCL's version banner under
    # NCCL_DEBUG=VERSION/INFO) are sent to stderr until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def restore_stdout():
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    batch = args.batch
    zkey = open(os.path.join(ART, "proving_key.zkey"), "rb").read()
    wasm = open(os.path.join(ART, "circuit.wasm"), "rb").read()
    vkey = open(os.path.join(ART, "verification_key.json"), "rb").read()
    t_load = time.perf_counter()
    c = prover.load(zkey, wasm, device=local_rank)
    t_load = time.perf_counter() - t_load
    # synthetic census: 1,024 voters (seed 0xC0FFEE + rank so that ranks prove different censuses)
    voters = census_tree.gen_census(c, min(batch, 1024), seed=0xC0FFEE + rank)
    docs = [json.dumps(voters[i % len(voters)]).encode() for i in range(batch)]
    packed = np.stack([prover.pack_inputs(voters[i % len(voters)]) for i in range(batch)])
    stream = torch.cuda.ExternalStream(c.ctx.stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    c.set_inputs(packed)
    for _ in range(args.warmup):
        c.prove_resident(batch)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = prover.launch_count()
    ev0.record(stream)
    for _ in range(args.steps):
        c.prove_resident(batch)
    ev1.record(stream)
    barrier()
    launches = prover.launch_count() - launches0
    # per-kernel durations: the same K steps once more, instrumented with CUDA events between the stages and run
    # serially on one lane (in the timed loop above chunks overlap on four streams, which would smear the brackets)
    stages = np.zeros(8, dtype=np.float64)
    for _ in range(args.steps):
        stages += c.prove_resident(batch, stages=True)
    clk = clocks.stop()
    dev_ms = ev0.elapsed_time(ev1)
    dev_ms = max_over_ranks(dev_ms, dev)
    total = sum_over_ranks(float(batch * args.steps), dev)
    value = total / (dev_ms * 1e-3)
    proofs_bin, pubs_bin, status = c.get_results(batch)
    assert (status == 0).all(), "a synthetic proof failed its circuit asserts"
    # every proof of the last timed step is verified on the GPU under the key's verification key (outside the timing)
    verified_resident = int(prover.verify_batch_bin(vkey, pubs_bin, proofs_bin).sum())
    work = c.work_counters()                 # executed mixed adds per proof in the last chunk (exact, counted on device)

    # ---- end to end through the reference-shaped call (JSON in host memory -> JSON in host memory) ----
    c.fullprove_batch(docs[:min(batch, 64)])
    barrier()
    t0 = time.perf_counter()
    proofs = pubs = None
    for _ in range(args.steps):
        proofs, pubs, st = c.fullprove_batch(docs)
        assert all(s == 0 for s in st)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
    e2e = total / e2e_s
    verified = int(sum(prover.verify_batch(vkey, pubs, proofs)))       # the JSON documents a caller receives
    verified_all = int(sum_over_ranks(float(min(verified, verified_resident)), dev))

    # ---- single-proof latency: p50 of 30 reference-shaped calls, key tables resident ----
    lat = []
    if rank == 0:
        one = docs[0]
        for i in range(33):
            t1 = time.perf_counter()
            c.fullprove(one)
            if i >= 3:
                lat.append((time.perf_counter() - t1) * 1e3)
        lat.sort()

    # ---- multi-GPU paths of SURVEY.md 8e: self-check against committed KATs / single-GPU results (N > 1) ----
    selfcheck = None
    if world > 1 and not os.environ.get("ZKB_SKIP_SELFCHECK"):
        c.close()
        prover._circuits.clear()
        selfcheck = multi_gpu_selfcheck(rank, world, local_rank, dist, torch)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel: G1 bucket accumulation (integer pipe) ----
    stages /= args.steps                     # ms per step: witness, abc, ntt+join, sort, acc_g1, acc_g2, reduce, finalize
    # peak of the integer-multiply pipe: the best of three runs of the dependent-product benchmark, and never less than
    # the issue limit it converges to (32 multiplier instructions per clock and SM, 137 per Montgomery product) - a
    # benchmark run that comes out low must not inflate `frac`
    peak_bench = max(raw.bench_modmul("fq", 4096, 8)[0] for _ in range(3))
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    issue_limit = 32.0 * sms * float(clk.get("sm_max_mhz") or 0.0) * 1e6 / 137.0
    peak_modmul = max(peak_bench, issue_limit)
    # executed field products of the G1 accumulation stage (XYZZ mixed adds + the H MSM's affine pair tree)
    g1_modmul_per_proof = (work["g1_madds_per_proof"] * MODMUL_PER_MADD_G1 +
                           work["g1_affine_adds_per_proof"] * MODMUL_PER_AFFINE_ADD +
                           work["g1_inversions_per_proof"] * MODMUL_PER_INVERSION)
    acc_g1_s = stages[4] * 1e-3
    achieved = g1_modmul_per_proof * batch / acc_g1_s if acc_g1_s > 0 else 0.0
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (tools/ncu_summary.py --raw)
    traffic = None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes")
            break
    # executed modmuls per proof (what this pipeline runs; SURVEY 8d: never divide dense work by shortcut time)
    logd = c.domain.bit_length() - 1
    executed = (g1_modmul_per_proof + work["g2_madds_per_proof"] * 28                 # bucket accumulation
                + 6 * (c.domain // 2) * logd + 3 * c.domain                          # 6 transforms + coset scale
                + 462889 + 2 * c.domain                                              # buildABC (nnz) + join
                + 2 * 32768 * 14 + 3 * 2 * 4096 * 14 + 2 * 4096 * 14 * 3             # bucket reductions: H, A/B1/C, B2 (x3 Fq)
                + 2.0e4 + 8.0e3)                                                     # witness (levels above the leaf), assembly
    pair_tree = work["g1_affine_adds_per_proof"] > 0
    roofline = {"bound": "imad", "kernel": ("G1 bucket accumulation: k_affine_level (H MSM pair tree, batched-affine adds, ZKB_AFFINE=1) + "
                                           "k_accumulate<Fq> / k_accumulate_pts (XYZZ mixed adds)") if pair_tree else
                                          "k_accumulate<Fq> (G1 bucket accumulation, XYZZ mixed adds)",
                "executed_per_proof": {"xyzz_madds": work["g1_madds_per_proof"], "affine_adds": work["g1_affine_adds_per_proof"],
                                       "inversions": work["g1_inversions_per_proof"], "modmul": g1_modmul_per_proof,
                                       "modmul_if_all_xyzz": (work["g1_madds_per_proof"] + work["g1_affine_adds_per_proof"]) * 10},
                "achieved": achieved / 1e9, "peak": peak_modmul / 1e9, "unit": "Gmodmul/s",
                "frac": achieved / peak_modmul if peak_modmul else None, "traffic": traffic,
                "peak_source": "max(zkb_bench_modmul measured in this run: dependent 254-bit Montgomery products, 137 multiplier "
                               "instructions each, best of 3; issue limit 32 per clock and SM x SMs x max SM clock / 137); "
                               "MEASURED_PEAKS.json has no integer-pipe figure",
                "peak_benchmark": peak_bench / 1e9, "peak_issue_limit": issue_limit / 1e9,
                "share_of_step": float(stages[4] / stages.sum()) if stages.sum() else None,
                "measured_in": "instrumented serial pass (1 lane) over the same K steps, CUDA events on the launching "
                               "stream; the timed loop overlaps chunks on 4 streams",
                "traffic_note": "dram__bytes_read+write of the largest k_accumulate<Fq> launch (H MSM of one 128-proof chunk: "
                                "128 x 2.1 M gathered 64-byte table points = 17.2 GB algorithmic) from profiles/",
                "executed_modmul_per_proof": executed,
                "dense_equiv_modmul_per_proof": DENSE_MODMUL_PER_PROOF,
                "whole_step_frac_of_peak": executed * value / world / peak_modmul if peak_modmul else None,
                "hbm_gbs_measured": measured_peaks().get("hbm_gbs")}
    # ---- W7 measurement rule: throughput as a function of the tree depth, and with the shortcut switched off ----
    shortcut = generic = None
    if world == 1 and not os.environ.get("ZKB_SKIP_DEPTHS"):
        nd = int(os.environ.get("ZKB_DEPTH_BATCH", "256"))
        pts = []
        for depth in (4, 20, 40, 160):
            vs = census_tree.gen_census_depth(c, 64, depth, seed=0xC0FFEE)
            pk = np.stack([prover.pack_inputs(v) for v in vs])
            rate = timed_resident(c, pk, nd, 1, stream, torch)
            w = c.work_counters()
            pts.append({"tree_depth": depth, "proofs_per_s": rate, "g1_madds_per_proof": w["g1_madds_per_proof"],
                        "g2_madds_per_proof": w["g2_madds_per_proof"],
                        "witness_digit_entries_per_proof": w["witness_digit_entries_per_proof"]})
        c.close()                                   # the dense circuit needs the HBM the default one's workspace holds
        prover._circuits.clear()
        cd = prover.load(zkey, wasm, device=local_rank, dense=True)
        rate_dense = timed_resident(cd, packed[:64], nd, 1, stream, torch)
        wd = cd.work_counters()
        pd, qd, _ = cd.get_results(nd)
        dense_ok = int(prover.verify_batch_bin(vkey, qd, pd).sum())
        cd.close()
        prover._circuits.clear()
        # the same key with the witness program extracted from the wasm: what any other circom circuit gets
        t_gen = time.perf_counter()
        cg = prover.load(zkey, wasm, device=local_rank, generic=True)
        t_gen = time.perf_counter() - t_gen
        docs_g = docs[:nd]
        cg.fullprove_batch(docs_g)                  # warm-up at the timed size: workspace and value slots are allocated here
        t0 = time.perf_counter()
        pg, qg, sg = cg.fullprove_batch(docs_g)
        rate_generic = nd / (time.perf_counter() - t0)
        generic = {"proofs_per_s": rate_generic, "load_s": round(t_gen, 2), "status_ok": int(sum(1 for x in sg if x == 0)),
                   "verified": int(sum(prover.verify_batch(vkey, qg, pg))), "proofs": nd,
                   "what": "ZKB_LOAD_GENERIC_WITNESS: witness by the straight-line program extracted from circuit.wasm at load "
                           "time (one warp per proof, whole group in one launch); witness MSMs over the difference to the "
                           "program's all-zero-input witness (16-bit windows); zkb_fullprove_batch end to end"}
        cg.close()
        prover._circuits.clear()
        shortcut = {"what": "SURVEY 8a W7: ~91 % of the wires are identical in every proof below the leaf's level; the four "
                            "witness MSMs run over (w - template) and levels that hash (0,0) are not recomputed.  H MSM "
                            "and NTTs are unaffected.  Points below: device-resident proofs/s of " + str(nd) + " proofs",
                    "by_tree_depth": pts,
                    "shortcut_off": {"proofs_per_s": rate_dense, "g1_madds_per_proof": wd["g1_madds_per_proof"],
                                     "g2_madds_per_proof": wd["g2_madds_per_proof"], "verified": dense_ok,
                                     "what": "ZKB_LOAD_DENSE: all 161 levels hashed, A/B1/B2/C over all 82,754 wires "
                                             "(c = 16), same 1024-voter census"}}
    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ----
    cpu = None
    if world == 1:
        try:
            ns = int(os.environ.get("ZKB_CPU_SAMPLE", "8"))
            rate, cores, kind, desc = cpu_prove_sample(ns, voters[:ns])
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc,
                   "witness_seconds": cpu_prove_sample.last_witness_seconds}
            rate2, _, _, desc2 = cpu_prove_sample(ns, voters[:ns], shortcut=True)
            cpu["with_same_shortcut"] = {"value": rate2, "unit": UNIT, "sample": desc2}
            if shortcut:
                shortcut["speedup_split"] = {
                    "gpu_e2e_over_cpu_dense": e2e / rate, "gpu_e2e_over_cpu_with_shortcut": e2e / rate2,
                    "gpu_shortcut_off_over_cpu_dense": shortcut["shortcut_off"]["proofs_per_s"] / rate,
                    "algorithmic_gain_on_gpu": value / shortcut["shortcut_off"]["proofs_per_s"],
                    "algorithmic_gain_on_cpu": rate2 / rate}
        except Exception as e:  # the oracle is optional at bench time
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"unavailable: {e}"}
    n_in = c.n_inputs
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u256 (8x32-bit limb Montgomery, IMAD)", "data": "synthetic",
            "config": bench_config(voters, batch),
            "run": {"chunk": work["chunk"], "key_load_s": round(t_load, 2)},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": batch * (n_in * 32 + 64),
                    "d2h_bytes_per_step": batch * (256 + 32 * c.n_public + 4),
                    "path": "zkb_fullprove_batch: inputs.json strings -> proof.json/public.json strings"},
            "verified": verified_all, "verified_of": int(total / args.steps),
            "verified_what": "every proof of the last timed step, device-resident results (zkb_verify_batch_bin) and the JSON "
                             "documents of the last end-to-end step (zkb_verify_batch), all ranks; min of the two counts",
            "latency_ms": {"p50": lat[len(lat) // 2], "min": lat[0], "max": lat[-1], "calls": len(lat),
                           "what": "one zkb_fullprove call (inputs.json -> proof.json), key resident"},
            "gpu_launches": int(launches),
            "stage_ms_per_step": {k: float(v) for k, v in zip(
                ("witness", "build_abc", "ntt_join", "msm_sort", "msm_acc_g1", "msm_acc_g2", "msm_reduce", "finalize"), stages)},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clk}
    if shortcut:
        line["shortcut_w7"] = shortcut
        line["generic_witness"] = generic
    if selfcheck is not None:
        line["multi_gpu_selfcheck"] = selfcheck
    restore_stdout()
    print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("ZKB_BENCH_BATCH", "1024")))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
