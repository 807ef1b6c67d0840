#!/usr/bin/env python3
"""BASELINE.json config 5: raw BN254 G1 MSM and Fr NTT sweep (2^16 .. 2^26) at 1/2/4/8 GPUs against the roofline.

    python tools/sweep_raw.py --msm 16,18,20,22,24,26 --ntt 16,18,20,22,24,26 --out profiles/r01_sweep_1gpu.jsonl
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/sweep_raw.py ...

One process per GPU.  MSM: every rank owns the point range [rank N/G, (rank+1) N/G), cut into sub-MSMs of <= 2^17
points; its last kernel writes the partial sum into rank 0's exchange buffer over NVLink (CUDA IPC mapping) and rank 0
adds the G partial sums.  The host-side process group (gloo) only carries the 64-byte IPC handle and the barrier /
max-over-ranks timing; there is no data-path collective.  Fixed-base window tables (c = 16) are key material and are
built once per size; their build time is reported separately.  NTT: per-GPU resident transforms (the census path's
batched shape); `nvec` vectors per GPU, weak scaling.
Roofline: MSM = executed mixed adds x 10 modmul / time vs the integer-pipe peak measured in the same run
(zkb_bench_modmul); NTT = modmul/s vs the same peak and algorithmic bytes (2 x 32 x N per transform) / time vs
MEASURED_PEAKS.json hbm_gbs."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--msm", default="16,18,20,22,24")
    ap.add_argument("--ntt", default="16,18,20,22,24")
    ap.add_argument("--ntt-dist", default="", help="sizes (log2, >= 22) for the 4-step NTT split over all ranks")
    ap.add_argument("--nccl-compare", action="store_true",
                    help="also time the NCCL collectives the P2P exchanges replace (all_gather of one 128-byte partial "
                         "sum; all_to_all of the NTT columns), same GPUs, CUDA events")
    ap.add_argument("--variable-base", action="store_true",
                    help="also run every MSM size without a window table (bases used as they are, one bucket set per "
                         "window, Horner combination): the column for bases that are not key material")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    from zk_franchise_proof_circuit_b200 import raw
    if world > 1:
        dist.init_process_group("gloo")

    def barrier():
        if world > 1:
            dist.barrier()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    torch.cuda.set_device(local)
    peak, _ = raw.bench_modmul("fq", 4096, 8)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    lines = []

    def emit(d):
        if rank == 0:
            print(json.dumps(d), flush=True)
            lines.append(d)

    for logn, vb in [(int(x), v) for x in args.msm.split(",") if x for v in ((False, True) if args.variable_base else (False,))]:
        s = raw.MsmSession(logn, rank, world, device=local, seed=1, window=16, variable_base=vb)
        if world > 1:
            box = [s.export_handle() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                s.attach(box[0])
        point = None
        times, own = [], []
        for it in range(args.steps + 1):          # first iteration = warm-up
            barrier()
            if rank == 0:
                s.run()
                point, ms = s.combine(world)
            else:
                ms = s.run(wait=True)
            barrier()
            if it:
                times.append(reduce_max(ms))
        madds = reduce_sum(float(s.madds()))
        t = sorted(times)[len(times) // 2] * 1e-3
        emit({"kind": "msm_g1_variable_base" if vb else "msm_g1", "logn": logn, "n_gpus": world, "ms": t * 1e3, "points_per_s": (1 << logn) / t,
              "madds": madds, "gmodmul_per_s": madds * 10 / t / 1e9, "peak_gmodmul_per_s": peak * world / 1e9,
              "frac_of_imad_peak": madds * 10 / t / (peak * world), "window_bits": 16, "sub_msm_points": s.sub_size,
              "sub_msms_per_gpu": s.subs, "table_build_ms_rank0": s.table_ms, "gen_ms_rank0": s.gen_ms,
              "result_x_hex": point[:32][::-1].hex() if point else None,
              "combine": "P2P push into rank 0 + flag (no NCCL)" if world > 1 else "local"})
        s.close()
        barrier()

    for logn in [int(x) for x in args.ntt.split(",") if x]:
        n = 1 << logn
        nvec = max(1, min(384, (1 << 26) // n))    # <= 2 GiB of vectors per GPU; 384 = one census chunk (3 x 128)
        barrier()
        dif, dit = raw.ntt_bench(logn, nvec=nvec, iters=args.steps, device=local)
        dif, dit = reduce_max(dif), reduce_max(dit)
        per = (dif + dit) / 2 / nvec * 1e-3            # seconds per transform
        mm = n / 2 * logn + n / 2                      # butterflies + half of the coset-scale multiplies (avg of the two)
        emit({"kind": "ntt_fr", "logn": logn, "n_gpus": world, "nvec_per_gpu": nvec, "dif_ms": dif, "dit_ms": dit,
              "us_per_transform": per * 1e6, "transforms_per_s": world / per,
              "gmodmul_per_s": mm / per * world / 1e9, "frac_of_imad_peak": mm / per / peak,
              "algorithmic_gbs_per_gpu": 2 * 32 * n / per / 1e9,
              "frac_of_hbm_peak": (2 * 32 * n / per / 1e9) / peaks.get("hbm_gbs", 6464.6),
              "scaling": "weak (independent vectors per GPU)"})
    for logn in [int(x) for x in args.ntt_dist.split(",") if x]:
        s = raw.NttDist(logn, rank, world, device=local)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, s.export_handle())
            for g in range(world):
                if g != rank:
                    s.attach(g, handles[g])
        rows = []
        for it in range(args.steps + 1):
            s.fill()
            barrier()
            s.run()
            ms = s.sync()
            barrier()
            if it:
                rows.append([reduce_max(x) for x in ms])
        rows.sort()
        tot, colms, exms, rowms = rows[len(rows) // 2]
        n = 1 << logn
        mm = n / 2 * logn + 2 * n                      # butterflies + 2 modmul per element for the fused twiddle
        emit({"kind": "ntt_fr_4step", "logn": logn, "n_gpus": world, "ms": tot, "columns_ms": colms,
              "wait_plus_p2p_exchange_ms": exms, "rows_ms": rowms, "gmodmul_per_s": mm / (tot * 1e-3) / 1e9,
              "frac_of_imad_peak": mm / (tot * 1e-3) / (peak * world),
              "exchange_bytes_per_gpu": (n // world) * 32 * (world - 1) // world,
              "exchange": "fused into the transposing load of the row pass: P2P reads of peer buffers (CUDA IPC), "
                          "no NCCL" if world > 1 else "local transpose", "scaling": "strong"})
        s.close()
        barrier()
    if args.nccl_compare and world > 1:
        # What NCCL would cost for the same exchanges (north_star: "no NCCL unless measured to help")
        pg = dist.new_group(backend="nccl")
        dev = torch.device("cuda", local)
        part = torch.zeros(32, dtype=torch.int32, device=dev)                  # one XYZZ partial sum = 128 bytes
        outl = [torch.zeros_like(part) for _ in range(world)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(20):
            dist.all_gather(outl, part, group=pg)
        torch.cuda.synchronize()
        barrier()
        e0.record()
        for _ in range(200):
            dist.all_gather(outl, part, group=pg)
        e1.record()
        torch.cuda.synchronize()
        ag_us = reduce_max(e0.elapsed_time(e1) / 200 * 1e3)
        emit({"kind": "nccl_compare", "what": "all_gather of one 128-byte partial sum (the MSM combine)", "n_gpus": world,
              "nccl_us_per_call": ag_us, "note": "back-to-back calls on one stream, CUDA events; the P2P push is one 128-byte "
              "store + flag inside the rank's last kernel"})
        for logn in [int(x) for x in args.ntt_dist.split(",") if x]:
            per = (1 << logn) // world
            src = torch.zeros(per * 8, dtype=torch.int32, device=dev)
            dst = torch.zeros_like(src)
            for _ in range(3):
                dist.all_to_all_single(dst, src, group=pg)
            torch.cuda.synchronize()
            barrier()
            e0.record()
            for _ in range(10):
                dist.all_to_all_single(dst, src, group=pg)
            e1.record()
            torch.cuda.synchronize()
            a2a = reduce_max(e0.elapsed_time(e1) / 10)
            emit({"kind": "nccl_compare", "what": "all_to_all_single of the 4-step NTT columns (no transpose, no twiddles)",
                  "logn": logn, "n_gpus": world, "nccl_ms": a2a, "bytes_per_gpu": per * 32})
    if rank == 0 and args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
