#!/usr/bin/env python
"""bench.py -- spin-flip attempts/s of the PIQMC Metropolis-sweep hot path on B200.

Workload (BASELINE.json configs[4]): synthetic 256x256 Gaussian torus (J ~ N(0,1) float32,
RandomState(2024)), P = 64 Trotter slices, R = 4096 replicas in total sharded over the ranks
(strong scaling, R/N per GPU, no collective on the sweep path), T = 0.01, Gamma linspace(1.5, 1e-8, K).
One "step" = one schedule step = one Metropolis sweep (mcsteps = 1) over all replicas:
R*P*N = 1.72e10 attempts.  The production kernel runs the natural-order level colouring, i.e. the
sweep of the reference's per-spin-reset variant qmc.QuantumAnneal_parallel, whose residual-energy
statistics it reproduces (tests/test_gpu_colour.py).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one rank per GPU)
  python bench.py --impl reference [...]                         the reference's own Cython on host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pathintegral-qmc_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

L, P, R_TOTAL, TEMP = 256, 64, 4096, float(os.environ.get("PIQMC_BENCH_TEMP", "0.01"))   # (env: experiments only)
GAMMA0, GAMMA1 = 1.5, 1e-8
SEED = 2024
B_ALG_HBM = 0.25        # bytes/attempt: each 64-slice word read once + written once per sweep
B_ALG_SMEM = 1.0        # bytes/attempt touched on chip: own r+w, 4 neighbours, 2 Trotter bits
# from the ncu captures under profiles/: measured DRAM traffic relative to the algorithmic bytes, and executed
# warp-instructions per attempt, by rows per GPU (the kernel's per-unit overhead weighs more with few rows)
NCU_CAPTURES = {4096: {"traffic_over_algorithmic": 1.006, "winst_per_attempt": 0.1569,
                       "source": "profiles/r2_colour_sweep_fast_ncu.md (4096 rows per GPU)"},
                512: {"traffic_over_algorithmic": 1.002, "winst_per_attempt": 0.2985,
                      "source": "profiles/r2_colour_sweep_fast_ncu_512rows.md (512 rows per GPU)"}}


def ncu_capture(rows_per_gpu):
    """the capture whose rows per GPU are closest (log scale) to this run's"""
    return NCU_CAPTURES[min(NCU_CAPTURES, key=lambda r: abs(np.log(r) - np.log(max(rows_per_gpu, 1))))]
METRIC = "spin-flip attempts/sec"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ----------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU baseline: the reference's own compiled Cython (oracle/_ref) on the host cores
# ----------------------------------------------------------------------------------------------
_G = {}


def _cpu_init(nbs):
    from oracle import oracle as O
    ref = O.ref()
    if ref is not None:
        from piqmc_ref import qmc
        _G["qmc"] = qmc
    _G["O"] = O
    _G["nbs"] = nbs


def _cpu_job(args):
    """One replica of config 5 for `nsteps` sweeps on one core; returns seconds."""
    r, nsteps, kind = args
    import ctypes
    n = L * L
    rng = np.random.RandomState(r)
    ctypes.CDLL("libc.so.6").srand(r)
    sv = (2 * rng.randint(2, size=n) - 1).astype(np.float64)
    confs = np.tile(sv, (P, 1)).T.copy()
    sched = np.linspace(GAMMA0, GAMMA1, nsteps)
    t0 = time.perf_counter()
    if "qmc" in _G:
        if kind == "qa":
            _G["qmc"].QuantumAnneal(sched, 1, P, TEMP, n, confs, _G["nbs"], rng)
        elif kind.startswith("qa_omp"):                      # the reference's OpenMP variant, all threads in one process
            _G["qmc"].QuantumAnneal_parallel(sched, 1, P, TEMP, n, confs, _G["nbs"], int(kind[6:]))
        elif kind == "sa":                                   # classical pre-anneal on one slice-sized vector, P sweeps
            from piqmc_ref import sa as ref_sa
            ref_sa.Anneal(np.linspace(3.0, 0.01, nsteps * P), 1, sv, _G["nbs"], rng)
        else:
            _G["qmc"].QuantumAnneal_parallel(sched, 1, P, TEMP, n, confs, _G["nbs"], 1)
    else:
        O = _G["O"]
        if kind == "qa":
            O.qa_reference(sched, 1, P, TEMP, n, confs, _G["nbs"], O.make_perms(rng, n, nsteps))
        elif kind == "sa":
            ssched = np.linspace(3.0, 0.01, nsteps * P)
            O.sa_reference(ssched, 1, sv, _G["nbs"], O.make_perms(rng, n, ssched.size))
        else:
            O.qa_parallel1(sched, 1, P, TEMP, n, confs, _G["nbs"])
    return time.perf_counter() - t0


def cpu_rate(nbs, nsteps, kind, cores):
    """attempts/s of `cores` processes each running one replica for nsteps sweeps (the pattern of
    examples/spinglass32_mpi.py: independent replicas, one per core)."""
    from oracle import oracle as O
    have_ref = O.ref() is not None
    with mp.Pool(cores, initializer=_cpu_init, initargs=(nbs,)) as pool:
        pool.map(_cpu_job, [(r, 1, kind) for r in range(cores)])            # warm caches / imports
        t0 = time.perf_counter()
        pool.map(_cpu_job, [(r, nsteps, kind) for r in range(cores)])
        wall = time.perf_counter() - t0
    return cores * P * L * L * nsteps / wall, wall, ("reference" if have_ref else "port")


def run_reference(args):
    """--impl reference: the reference's own Cython (oracle/_ref; the C port if that build is absent)
    on all host cores, one replica per core (the pattern of examples/spinglass32_mpi.py), K sweeps of
    config 5 each.  The variant timed is the reference's fastest one for this path on this host:
    qmc.QuantumAnneal_parallel with nthreads=1 per process (per-spin energy difference, natural order
    -- the semantics the GPU production kernel reproduces); the as-shipped qmc.QuantumAnneal is ~2.5x
    slower per core (Python iterator over the permutation) and is reported in our arm's cpu_baseline."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import piqmc.tools as tools
    nbs, _ = tools.GaussianTorusNeighbors(L, SEED)
    cores = os.cpu_count() or 1
    if args.warmup > 0:
        cpu_rate(nbs, min(args.warmup, 2), "qa_par", cores)
    rate, wall, kind = cpu_rate(nbs, args.steps, "qa_par", cores)
    sample = ("%d replicas (one per core) x %d sweeps of the 256x256 P=64 workload, "
              "qmc.QuantumAnneal_parallel(nthreads=1) per process" % (cores, args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "attempts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args),                       # the workload named; what this arm ran of it: cpu_baseline
        "cpu_baseline": {"value": rate, "unit": "attempts/s", "cores": cores, "kind": kind, "sample": sample,
                         "replicas_run": cores},
        "e2e": {"value": rate, "unit": "attempts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def config_dict(args):
    state_gb = args.replicas * L * L * 8 / 1e9
    return {"workload": "synthetic %dx%d Gaussian 2D Ising torus, P=%d slices, R=%d replicas total, "
                        "T=%g, Gamma 1.5->1e-8 over K steps, mcsteps=1 (BASELINE.json configs[4]%s)"
                        % (L, L, P, args.replicas, TEMP, "" if args.replicas == 4096 else ", replica count changed"),
            "nspins": L * L, "slices": P, "replicas_total": args.replicas, "order": args.order,
            "state": "bit-packed uint64 word per (spin, replica), %.2f GB total" % state_gb,
            "l2_policy": "inputs larger than L2: per-GPU state %.0f MB at %d GPUs vs 126 MB L2"
                         % (1e3 * state_gb / max(args.gpus, 1), max(args.gpus, 1))}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import piqmc.qmc as qmc
    import piqmc.tools as tools
    from piqmc import device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    from piqmc.shard import bind_to_gpu_numa, gather_energies, shard_replicas
    numa_cores = bind_to_gpu_numa(local) if (world > 1 and not os.environ.get("PIQMC_NO_NUMA_BIND")) else 0
    replica0, R = shard_replicas(R_TOTAL, world, rank)
    n = L * L
    K, W = args.steps, args.warmup

    nbs, checker = tools.GaussianTorusNeighbors(L, SEED)
    color = tools.TorusNaturalLevels(L) if args.order == "natural" else checker
    dev = device.Device(local)
    dev.set_variant(args.variant)
    dev.set_graph(nbs, color)
    dev.state_alloc(R, P)
    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local))
    sched = np.linspace(GAMMA0, GAMMA1, K)

    def barrier():
        dev.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- warm-up (untimed), then restart from the initial state
    dev.state_init_random(SEED, replica0, tile=True)
    if W > 0:
        dev.qa_colour(np.full(W, GAMMA0), 1, TEMP, SEED, replica0=replica0, sweep0=1 << 30)
    dev.state_init_random(SEED, replica0, tile=True)
    barrier()

    # ---- timed region: K schedule steps, device-resident state, CUDA events on the launch stream
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = dev.launch_count
    barrier()
    e0.record(stream)
    dev.qa_colour(sched, 1, TEMP, SEED, replica0=replica0)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = dev.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- final energies on device + the single gather (outside the timed region, reported)
    tg0 = time.perf_counter()
    if dist is not None:
        en = gather_energies(dev, R_TOTAL).cpu().numpy()      # NCCL all-gather of float64[R, P]
    else:
        en = dev.energy()
    gather_ms = 1e3 * (time.perf_counter() - tg0)

    # ---- end to end through the public API with HOST buffers: H2D of the initial spins (pinned),
    #      the anneal, the device energy reduction, D2H of packed configurations + energies
    spins0 = device.pinned_empty((R, n), np.int8)
    spins0[:] = (2 * np.random.RandomState(SEED + rank).randint(2, size=(R, n)) - 1).astype(np.int8)
    words_out = device.pinned_empty((n, R), np.uint64)
    # one untimed one-step call so that the library's grow-only staging buffers exist (warm-up)
    qmc.QuantumAnnealReplicas(sched[:1], 1, P, TEMP, n, spins0, nbs, SEED, color=color, replica0=replica0,
                              device=dev, energies=True, download=True, words_out=words_out)
    barrier()
    pipe0 = dev.pipelined_runs
    t0 = time.perf_counter()
    out = qmc.QuantumAnnealReplicas(sched, 1, P, TEMP, n, spins0, nbs, SEED, color=color, replica0=replica0,
                                    device=dev, energies=True, download=True, words_out=words_out)
    dev.synchronize()
    e2e_s = time.perf_counter() - t0
    pipelined = dev.pipelined_runs - pipe0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = spins0.nbytes + nbs.size * 4 * 3
    d2h = out["words"].nbytes + out["energies"].nbytes

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    attempts = float(R_TOTAL) * P * n * K
    value = attempts / (ms * 1e-3)
    pk, pk_kind = peaks()
    per_launch_attempts = float(R) * P * n * K / launches
    avg_launch_s = ms * 1e-3 / launches
    achieved = B_ALG_HBM * per_launch_attempts / avg_launch_s / 1e9
    sm_mhz = (clocks or {}).get("sm_mhz") or pk.get("sm_max_mhz", 1965.0)
    nsm = torch.cuda.get_device_properties(local).multi_processor_count
    smem_peak = 128.0 * nsm * sm_mhz * 1e6 / 1e9
    cap = ncu_capture(R)
    winst = cap["winst_per_attempt"]
    line = {
        "metric": METRIC, "value": value, "unit": "attempts/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(args),
        "clocks": clocks,
        "e2e": {"value": attempts / e2e_s, "unit": "attempts/s", "h2d_bytes_per_step": h2d / K,
                "d2h_bytes_per_step": d2h / K, "seconds": e2e_s, "breakdown_s": out["seconds"],
                "overlapped_download": bool(pipelined),
                "call": "qmc.QuantumAnnealReplicas(host int8 spins) -> host packed words + float64 energies; with "
                        "overlapped_download the row chunks of the one sweep launch finish staggered and are "
                        "downloaded while the others still sweep (its breakdown then has no separate download figure)"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "colour_sweep_fast", "achieved": achieved, "peak": pk["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                     "traffic": cap["traffic_over_algorithmic"] * B_ALG_HBM * per_launch_attempts,
                     "traffic_source": cap["source"],
                     "peak_source": pk_kind, "bytes_per_attempt": B_ALG_HBM,
                     "note": "instruction-issue-bound kernel (~6 thread-instructions per attempt, no contraction): "
                             "the HBM fraction is low by construction, traffic (ncu, bytes per launch) equals the "
                             "algorithmic bytes; see roofline_issue and DESIGN.md section 4"},
        "roofline_issue": {"bound": "issue_slots", "achieved": winst * value / world / 1e9,
                           "peak": nsm * 4 * sm_mhz * 1e6 / 1e9, "unit": "G warp-inst/s",
                           "frac": winst * value / world / (nsm * 4 * sm_mhz * 1e6),
                           "peak_source": "%d SMs x 4 schedulers x 1 warp-inst/clk x %.0f MHz" % (nsm, sm_mhz),
                           "warp_inst_per_attempt": winst, "warp_inst_source": cap["source"],
                           "note": "what binds this kernel with many rows per GPU (ncu: 75% of the issue slots used at "
                                   "4096 rows, ALU pipe 63%); with 512 rows per GPU the dependency chain of the "
                                   "natural-order wavefront binds it (53% of the slots)"},
        "roofline_smem": {"achieved": B_ALG_SMEM * value / world / 1e9, "peak": smem_peak, "unit": "GB/s",
                          "frac": B_ALG_SMEM * value / world / 1e9 / smem_peak,
                          "bytes_per_attempt": B_ALG_SMEM, "peak_source": "128 B/clk/SM x %d SMs x %.0f MHz" % (nsm, sm_mhz)},
        "gather_ms": gather_ms, "host_cores_bound_per_rank": numa_cores,
        "energy_per_spin": {"mean_over_slices": float(en.mean() / n), "best_slice_mean": float(en.min(axis=1).mean() / n)},
    }
    # ---- the deterministic (bit-exact replay) path: one drop-in call and a batch, reported next to the headline
    if world == 1 and not args.no_cpu:
        try:
            line["det_path"] = det_path_rates(dev)
        except Exception as e:                       # an extra: never loses the line
            line["det_path"] = {"error": str(e)[:200]}
    # ---- CPU baseline on this box's host cores (bounded sample), rank 0, N=1 only
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        r_qa, w_qa, kind = cpu_rate(nbs, 3, "qa", cores)
        r_par, w_par, _ = cpu_rate(nbs, 3, "qa_par", cores)
        r_1, w_1, _ = cpu_rate(nbs, 3, "qa", 1)
        r_omp, w_omp, _ = cpu_rate(nbs, 2, "qa_omp%d" % cores, 1)      # OpenMP variant: one process, all threads
        r_sa, w_sa, _ = cpu_rate(nbs, 3, "sa", cores)                  # sa.Anneal, same number of attempts per job
        line["cpu_baseline"] = {
            "value": max(r_qa, r_par), "unit": "attempts/s", "cores": cores, "kind": kind,
            "sample": "%d replicas (one per core) x 3 sweeps of the same 256x256 P=64 workload" % cores,
            "qmc.QuantumAnneal": r_qa, "qmc.QuantumAnneal_parallel(nthreads=1)": r_par,
            "qmc.QuantumAnneal_1core": r_1,
            "qmc.QuantumAnneal_parallel(nthreads=%d) OpenMP, 1 process" % cores: r_omp,
            "sa.Anneal (%d processes)" % cores: r_sa}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def det_path_rates(dev):
    """Throughput of the bit-exact replay kernels (one warp per replica, the replica in shared memory: the reference's sequential
    algorithm with its own random streams) on BASELINE.json configs[1]: inst_0_32x32, P = 20, T = 0.01, 100
    schedule steps = 2.05e6 attempts per replica."""
    import piqmc.qmc as qmc
    from piqmc import device
    nbs = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))["nbs_inst_0_32x32"]
    n, P2, T2 = 1024, 20, 0.01
    sched = np.linspace(1.5, 1e-8, 100)

    def start(r):
        rng = np.random.RandomState(r)
        sv = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
        return np.tile(sv, (P2, 1)).T.copy(), rng

    confs, rng = start(0)
    t0 = time.perf_counter()
    qmc.QuantumAnneal(sched, 1, P2, T2, n, confs, nbs, rng, device=dev)
    single = time.perf_counter() - t0
    R = 512
    starts = [start(r % 32) for r in range(R)]
    spins = np.ascontiguousarray(np.array([c for c, _ in starts]), dtype=np.int8)
    perms = np.stack([qmc._draw_perms(g, n, sched.size) for _, g in starts])
    st = device.rand_states(list(range(R)))
    dev.set_graph(nbs)
    t0 = time.perf_counter()
    dev.qa_det(sched, 1, P2, T2, spins, perms, rstates=st)
    batch = time.perf_counter() - t0
    per = float(n) * P2 * sched.size
    return {"workload": "qmc.QuantumAnneal replay, inst_0_32x32, P=20, 100 steps (BASELINE.json configs[1])",
            "single_call_s": single, "single_call_attempts_per_s": per / single,
            "batch_replicas": R, "batch_device_call_s": batch, "batch_attempts_per_s": per * R / batch}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--order", default="natural", choices=["natural", "checkerboard"])
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--replicas", type=int, default=4096, help="total replicas over all ranks (default 4096)")
    args = ap.parse_args()
    globals()["R_TOTAL"] = args.replicas
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
