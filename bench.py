#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Hades engine (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload W]

Workload (default `merkle4`, BASELINE.json configs[1]): one step = one batch of 2^20 independent
`Hash::digest(Domain::Merkle4, 4 scalars)` per GPU = 2^20 width-5 Hades permutations per GPU, on
synthetic uniform BlsScalars.  `value` = permutations/s of the whole job (all ranks), inputs resident
in HBM, timed with CUDA events on the launching stream, max over ranks.  `e2e` = the same metric through
the public API (`Hash.digest_batch`) with pinned HOST buffers: H2D + kernel + D2H inside the timed
region.  `--impl reference` times the reference's CPU algorithm (oracle/hades_ref.c, the faithful C
port: the Rust crate cannot be built here) on all host cores.  N > 1 (torchrun): every rank hashes its own
shard, no collective on the data path (weak scaling).

Other workloads (not the driver's headline; used for profiles/ and DESIGN.md numbers):
  --workload encrypt   2^20 x encrypt(L=2)  (configs[2])        --workload permute  raw 2^20 x 5 states
  --workload sweep     Domain::Other, in_len 1..256 at 2^18 items (configs[4]), subsampled lengths
  --workload tree      arity-4 Merkle tree over 4^11 (1 GPU) or 4^12 leaves per job with one NCCL all-gather
                       per level (configs[3] shape, scaled to the GPUs present)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hades_permutations_per_sec"
UNIT = "perm/s"
LOG2_BATCH = 20
BYTES_PER_PERM = 160          # Merkle4 digest: 4 x 32 B in + 32 B out


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons streamed (-lms) DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)                      # let the first samples arrive before the timed region
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        rows = []
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            for ln in out.splitlines():
                parts = [p.strip() for p in ln.split(",")]
                if len(parts) >= 7:
                    try:
                        float(parts[0])
                        rows.append(parts)
                    except ValueError:
                        pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(r[0]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": float(rows[0][1]),
                "reasons": reasons, "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows)}


def usable_cores():
    """Host threads this process may really use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_summary():
    """The committed ncu summary of the dominant kernel (profiles/r1_ncu_summary.json), if present."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_ncu_summary.json")) as f:
            return json.load(f).get("merkle4_2p20", {})
    except Exception:
        return {}


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture."""
    return ncu_summary().get("dram_bytes_per_launch")


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle C port; test infrastructure)
# ---------------------------------------------------------------------------------------------------------
def cpu_rate(n_items, threads, in_len=4):
    """Merkle4-shaped digests/s of the dense reference algorithm with `threads` host threads."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    import hades_oracle as o
    from poseidon252_b200.scalar import random_limbs_fast, to_mont
    rng = np.random.default_rng(123)
    data = random_limbs_fast(rng, (n_items, in_len))
    tag = to_mont(o.hash_to_scalar(o.tag_input([o.Absorb(in_len), o.Squeeze(1)], o.Domain.Merkle4)))
    c_oracle.digest(tag, data[:64], in_len, 1)
    t0 = time.perf_counter()
    c_oracle.digest(tag, data, in_len, 1, threads=threads)
    dt = time.perf_counter() - t0
    return n_items / dt, dt


def cpu_baseline_block(target_seconds=10.0):
    threads = usable_cores()
    probe_n = 4096 * threads
    cpu_rate(probe_n, threads)
    rate, _ = cpu_rate(probe_n, threads)
    n = int(min(1 << 22, max(probe_n, rate * target_seconds)))
    rate, dt = cpu_rate(n, threads)
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d Merkle4 digests (1 permutation each) of the dense reference algorithm "
                      "(oracle/hades_ref.c, 4x64-bit Montgomery), %d pthreads, %.1f s" % (n, threads, dt)}


def run_reference_arm(args, rank, world, emit):
    if rank != 0:
        return
    threads = usable_cores()
    probe_n = 4096 * threads
    cpu_rate(probe_n, threads)
    rate, _ = cpu_rate(probe_n, threads)
    per_step = int(max(probe_n, min(1 << 20, rate * 2.0)))        # ~2 s of host work per step
    for _ in range(args.warmup):
        cpu_rate(per_step, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_rate(per_step, threads)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 limbs (255-bit modular integer)", "data": "synthetic",
            "config": {"workload": "batch 2^20 Domain::Merkle4 digests (bounded sample: %d digests per step)" % per_step,
                       "algorithm": "reference dense Hades (src/hades/permutation/scalar.rs:39-64), C port"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d digests per step x %d steps, %d pthreads" % (per_step, args.steps, threads)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="merkle4", choices=["merkle4", "encrypt", "decrypt", "permute", "sweep", "tree", "convert"])
    ap.add_argument("--log2-batch", type=int, default=LOG2_BATCH)
    ap.add_argument("--log4-leaves", type=int, default=0, help="tree workload: 4^k leaves in the whole job (14 = BASELINE configs[3])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    # Exactly ONE line may reach stdout (the JSON); native libraries (e.g. NCCL's version banner) write to
    # fd 1 too, so fd 1 is pointed at stderr for the duration of the run and the JSON goes to the saved fd.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    if args.impl == "reference":
        run_reference_arm(args, rank, world, emit)
        return

    import numpy as np
    import torch
    import poseidon252_b200 as pb
    from poseidon252_b200.scalar import random_limbs_fast

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    stream = torch.cuda.Stream()
    eng = pb.Engine(local, stream=stream.cuda_stream)
    n = 1 << args.log2_batch
    rng = np.random.default_rng(0xC10D + rank)          # benches/hash.rs:53 seed, per-rank stream

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    extra = {}
    # ---- workload set-up: step(i) enqueues one batch on `stream`; returns perms per step ----------------
    if args.workload == "merkle4":
        nbuf = 4                                          # rotate over 4 x 128 MiB inputs (> 126 MB L2)
        with torch.cuda.stream(stream):
            ins = [torch.from_numpy(random_limbs_fast(rng, (n, 4)).view(np.int64)).cuda() for _ in range(nbuf)]
            out = torch.empty((n, 1, 4), dtype=torch.int64, device="cuda")
        perms_per_step, bytes_per_step = n, n * BYTES_PER_PERM
        step = lambda i: pb.Hash.digest_batch(pb.Domain.Merkle4, ins[i % nbuf], engine=eng, out=out, async_=True)
        workload = "batch 2^%d Domain::Merkle4 digests (4 scalars -> 1) per GPU" % args.log2_batch
        l2_note = "inputs rotate over %d distinct %d MiB device buffers (each > L2)" % (nbuf, n * 128 >> 20)
    elif args.workload == "permute":
        with torch.cuda.stream(stream):
            st = torch.from_numpy(random_limbs_fast(rng, (n, 5)).view(np.int64)).cuda()
        perms_per_step, bytes_per_step = n, n * 320
        step = lambda i: eng.permute_batch_inplace(st, async_=True)
        workload, l2_note = "raw permute_batch of 2^%d x 5 states in place" % args.log2_batch, "160 MiB state array > L2"
    elif args.workload == "encrypt":
        with torch.cuda.stream(stream):
            msg = torch.from_numpy(random_limbs_fast(rng, (n, 2)).view(np.int64)).cuda()
            sec = torch.from_numpy(random_limbs_fast(rng, (n, 2)).view(np.int64)).cuda()
            non = torch.from_numpy(random_limbs_fast(rng, (n,)).view(np.int64)).cuda()
            cip = torch.empty((n, 3, 4), dtype=torch.int64, device="cuda")
        perms_per_step, bytes_per_step = 2 * n, n * 256
        step = lambda i: pb.encrypt_batch(msg, sec, non, engine=eng, out=cip, async_=True)
        workload, l2_note = "encrypt_batch 2^%d messages, L=2 (benches/encrypt.rs:17)" % args.log2_batch, "256 MiB touched per step > L2"
    elif args.workload == "decrypt":
        with torch.cuda.stream(stream):
            msg = torch.from_numpy(random_limbs_fast(rng, (n, 2)).view(np.int64)).cuda()
            sec = torch.from_numpy(random_limbs_fast(rng, (n, 2)).view(np.int64)).cuda()
            non = torch.from_numpy(random_limbs_fast(rng, (n,)).view(np.int64)).cuda()
        cip = pb.encrypt_batch(msg, sec, non, engine=eng)
        eng.sync()
        perms_per_step, bytes_per_step = 2 * n, n * (192 + 64 + 1)
        ok_holder = {}

        def step(i):
            ok_holder["m"], ok_holder["ok"] = pb.decrypt_batch(cip, sec, non, engine=eng, async_=True)
        workload, l2_note = "decrypt_batch 2^%d ciphers, L=2 (benches/decrypt.rs:17)" % args.log2_batch, "257 MiB touched per step > L2"
    elif args.workload == "sweep":
        n = 1 << 18
        lens = [1, 2, 3, 4, 5, 8, 16, 32, 64, 128, 256]
        with torch.cuda.stream(stream):
            bufs = {L: torch.from_numpy(random_limbs_fast(rng, (n, L)).view(np.int64)).cuda() for L in lens}
            out = torch.empty((n, 1, 4), dtype=torch.int64, device="cuda")
        perms_per_step = sum(n * ((L + 3) // 4) for L in lens)
        bytes_per_step = sum(n * (32 * L + 32) for L in lens)

        def step(i):
            for L in lens:
                pb.Hash.digest_batch(pb.Domain.Other, bufs[L], engine=eng, out=out, async_=True)
        workload = "sponge sweep Domain::Other, in_len in %s, batch 2^18 per length" % lens
        l2_note = "each length's input is its own buffer; total %d MiB per step" % (bytes_per_step >> 20)
    elif args.workload == "convert":
        n = 1 << 25                                       # 1 GiB of scalars in, 1 GiB of bytes out
        with torch.cuda.stream(stream):
            sc = torch.from_numpy(random_limbs_fast(rng, (n,)).view(np.int64)).cuda()
            ob = torch.empty((n, 4), dtype=torch.int64, device="cuda")
        perms_per_step, bytes_per_step = n, n * 64       # "perms" here = scalars converted (no permutation)
        lib, ctx = eng._lib, eng._ctx
        step = lambda i: eng._check(lib.p252_scalars_to_bytes(ctx, sc.data_ptr(), n, ob.data_ptr(), 3))
        workload, l2_note = "to_bytes of 2^25 scalars (wire-format kernel, the one HBM-bound kernel); value = scalars/s", "1 GiB in + 1 GiB out per step"
    else:  # tree
        k = args.log4_leaves or (11 if world == 1 else 12)
        n_leaves = 4 ** k
        shard = n_leaves // world
        uid = eng.dist_unique_id() if rank == 0 else bytes(128)
        if dist is not None:
            box = [uid]
            dist.broadcast_object_list(box, src=0)
            uid = box[0]
            eng.dist_init(uid, rank, world)
        n_internal, n_levels = eng.tree_nodes(n_leaves)
        with torch.cuda.stream(stream):
            leaves = torch.from_numpy(random_limbs_fast(rng, (shard,)).view(np.int64)).cuda()
            nodes = torch.empty((n_internal, 4), dtype=torch.int64, device="cuda")
        perms_per_step, bytes_per_step = n_internal, n_internal * BYTES_PER_PERM
        step = lambda i: eng.merkle4_build_dist(leaves, n_leaves, out=nodes, async_=True)
        workload = "arity-4 Merkle tree, 4^%d leaves total, %d levels, one NCCL all-gather per level" % (k, n_levels)
        l2_note = "leaf shard %d MiB" % (shard * 32 >> 20)
        extra["tree_nodes"] = n_internal

    stream.synchronize()
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(args.steps):
            step(i)
            ev[i + 1].record(stream)
    stream.synchronize()
    barrier()
    clocks = sampler.stop()
    launches = eng.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    job_perms = perms_per_step * args.steps * (world if args.workload != "tree" else 1)
    value = job_perms / (total_ms * 1e-3)

    # ---- e2e: public API, pinned host buffers, H2D + kernel + D2H inside the timed region ---------------
    e2e = None
    if args.workload == "merkle4":
        h_in = torch.from_numpy(random_limbs_fast(rng, (n, 4)).view(np.int64)).pin_memory()
        h_out = torch.empty((n, 1, 4), dtype=torch.int64).pin_memory()
        a_in, a_out = h_in.numpy().view(np.uint64), h_out.numpy().view(np.uint64)
        e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            pb.Hash.digest_batch(pb.Domain.Merkle4, a_in, engine=eng, out=a_out)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(e_steps):
            pb.Hash.digest_batch(pb.Domain.Merkle4, a_in, engine=eng, out=a_out)      # synchronous HOST call
        e1.record(stream)
        stream.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e_ms = max(e0.elapsed_time(e1), wall_ms)
        if dist is not None:
            t = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        e2e = {"value": n * e_steps * world / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n * 128,
               "d2h_bytes_per_step": n * 32, "steps": e_steps, "ms_per_step": e_ms / e_steps,
               "api": "poseidon252_b200.Hash.digest_batch(Domain.Merkle4, pinned host array) -> p252_hash_batch(P252_MEM_HOST)"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    launch_ms = statistics.mean(per_launch_ms) / max(1, launches // args.steps)
    achieved = bytes_per_step / max(1, launches // args.steps) / (launch_ms * 1e-3) / 1e9
    sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "peak_source": peak_src,
                "kernel": "k_sponge_digest" if args.workload in ("merkle4", "sweep", "tree") else
                          ("k_crypt<false>" if args.workload == "encrypt" else "k_crypt<true>" if args.workload == "decrypt" else
                           ("k_convert<false>" if args.workload == "convert" else "k_permute<false>")),
                "algorithmic_bytes_per_launch": bytes_per_step // max(1, launches // args.steps),
                "launch_ms": launch_ms,
                "note": "the path is integer-issue bound (~10^3 integer ops per byte), not HBM bound; see int_pipe",
                "int_pipe": {"perm_per_s_per_sm_clock": value / world / sm_clock,
                             "ncu_pipe_fmaheavy_active_pct": ncu_summary().get("pipe_fmaheavy_active_pct"),
                             "ncu_issue_active_pct": ncu_summary().get("issue_active_pct"),
                             "ncu_imad_wide_per_permutation": (ncu_summary().get("warp_instructions_per_warp") or {}).get("IMAD.WIDE"),
                             "note": "from the committed ncu capture in profiles/ (the limiting resource is the IMAD pipe)"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "tree" else "weak", "vs_baseline": None,
            "dtype": "u32 limbs (255-bit modular integer, IMAD.WIDE carry chains)", "data": "synthetic",
            "config": {"workload": workload, "per_gpu_batch": perms_per_step, "l2": l2_note, "parallelism": "dp%d, no collective on the data path" % world
                       if args.workload != "tree" else "leaf shards, NCCL all-gather per level"},
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline, "target_perm_per_s_1gpu": 1e8}
    line.update(extra)
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_block()
        except Exception as exc:  # the oracle is only a reported baseline; never hide the GPU number
            line["cpu_baseline"] = {"error": repr(exc)}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
