#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: forward + discrete-adjoint gradients per second.

Workload (BASELINE.json metric / configs[3]): Generalized Lotka-Volterra, N = 64 species, 2^20 parameter sets,
controlled Cash-Karp 5(4) with rtol = atol = 1e-8, x0 = 0.1, t in [0, 10], dt0 = 1e-3, objective J = sum_i x_i(tf),
one gradient (dJ/dx0 [64] and dJ/dalpha [4160]) per parameter set; summed-objective mode reduces dJ/dalpha over the
batch (per-CTA register accumulation, deterministic reduction, one NCCL all-reduce when N > 1).

  python bench.py [--gpus N] [--steps K] [--warmup W]           one JSON line (this repo's CUDA path)
  python bench.py --impl reference [...]                         the reference's own CPU implementation, same metric
  torchrun --nproc-per-node N bench.py --gpus N ...              one rank per GPU; the batch is sharded, no data-path
                                                                 collective except the final 33 KB all-reduce

A "step" is one pass of the hot path over the whole batch (all 2^20 parameter sets, sharded over the ranks).
`value` times the pass with inputs resident in HBM (35 GB of parameters, far larger than L2); `e2e` times the same pass
through the C-ABI with HOST buffers (pinned), host<->device copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fwd+adjoint gradients/sec, GLV N=64 batch 1M"
N_SPECIES = 64
NPAR = N_SPECIES * N_SPECIES + N_SPECIES
TOL = 1e-8
TI, TF, DT0 = 0.0, 10.0, 1e-3
SEED = 1234
STAGES = 6

# Other BASELINE.json configs, measured with the same harness (`--workload`); the default (glv64) is the contract line.
#   name: (system, n_state, stepper, adaptive, tol, ti, tf, dt0, max_steps, objective, stages, description)
WORKLOADS = {
    "glv64": (2, 64, 2, True, 1e-8, 0.0, 10.0, 1e-3, 0, 1, 6, None),
    "glv16": (2, 16, 2, True, 1e-8, 0.0, 10.0, 1e-3, 0, 1, 6,
              "GLV N=16 (Npar=272), 2^20 parameter sets, cash_karp54 controlled rtol=atol=1e-8, t=[0,10], full r and A gradient"),
    "glv256": (2, 256, 2, True, 1e-8, 0.0, 10.0, 1e-3, 0, 1, 6,
               "GLV N=256 (Npar=65792), cash_karp54 controlled rtol=atol=1e-8, t=[0,10]; cluster-pair kernel (va_glv_pair.cu: matrix on chip in two SMs; VA_GLV_NO_PAIR=1: ring-streamed kernel va_glv_ring.cu); default batch 8192"),
    "glv256long": (2, 256, 2, True, 1e-12, 0.0, 1000.0, 1e-3, 512, 1, 6,
                   "GLV N=256 long horizon: cash_karp54 controlled rtol=atol=1e-12, t=[0,1000] (about 250 accepted steps per trajectory; "
                   "store-stages slabs of 148 x 513 x 36.9 KB = 2.8 GB live in HBM, not L2); cluster-pair kernel; default batch 2048"),
    "vdp": (1, 2, 3, True, 1e-8, 0.0, 0.5, 1e-3, 1024, 1, 7,
            "Van der Pol, mu swept over [1,1024), 2^20 parameter sets, dopri5 controlled rtol=atol=1e-8, t=[0,0.5], dt0=1e-3"),
    "harmonic": (0, 2, 1, False, 0.0, 0.0, 10.0, 0.01, 1024, 2, 4,
                 "damped harmonic oscillator, 2^20 parameter sets, fixed-step RK4 dt=0.01, t=[0,10] (1000 steps), J=|r(tf)|^2/2"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="total parameter sets (all ranks)")
    ap.add_argument("--reduce", default="sum", choices=["sum", "none"])
    ap.add_argument("--workload", default="glv64", choices=sorted(WORKLOADS), help="glv64 = the headline contract line")
    ap.add_argument("--species", type=int, default=0, help="with --workload glv256: any other species count (same tolerances; not a BASELINE config)")
    ap.add_argument("--ckpt-policy", default="auto", choices=["auto", "recompute", "store"], help="side workloads: checkpoint policy of the engine")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="parameter sets per reference-arm step (0 = auto)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation (oracle/_ref if it compiled, else the C port)
# ---------------------------------------------------------------------------------------------------------------------

def cpu_run(sample, threads):
    """One pass of the reference CPU path over `sample` parameter sets of the bench workload. Returns seconds."""
    import oracle
    p = oracle.synth_params(oracle.SYS_GLV, N_SPECIES, SEED, 0, sample)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N_SPECIES, p)
    if oracle.reference_available():
        kind = "reference"
        t0 = time.perf_counter()
        oracle.reference_forward_adjoint(oracle.SYS_GLV, N_SPECIES, oracle.RK_CK54, TOL, TOL, x0, p, TI, TF, DT0,
                                         objective=oracle.OBJ_SUM, threads=threads)
        dt = time.perf_counter() - t0
    else:
        kind = "port"
        t0 = time.perf_counter()
        oracle.forward_adjoint(oracle.SYS_GLV, N_SPECIES, oracle.RK_CK54, True, TOL, TOL, x0, p, TI, TF, DT0, objective=oracle.OBJ_SUM,
                               threads=threads)
        dt = time.perf_counter() - t0
    return dt, kind


def _calibrate(cores):
    """gradients/s of the CPU path from two short runs; the difference cancels the one-off cost (AADC records and
    JIT-compiles the RHS once per thread, ~0.2 s at N=64)."""
    t1, _ = cpu_run(2 * cores, cores)
    t2, _ = cpu_run(6 * cores, cores)
    return 4 * cores / max(t2 - t1, 1e-3)


def cpu_baseline(sample=0):
    cores = os.cpu_count() or 1
    if sample <= 0:
        rate = _calibrate(cores)
        sample = int(max(8 * cores, min(65536, rate * 15)))  # ~15 s of CPU work
        sample -= sample % cores
    t, kind = cpu_run(sample, cores)
    out = dict(value=sample / t, unit="gradients/s", cores=cores, kind=kind,
               sample=f"first {sample} of the 2^20 seeded parameter sets (same generator, seed {SEED}), one pass, "
                      f"{cores} host threads, one reference Driver per thread, Nout=1; {t:.2f} s")
    # The reference's design point: its four AVX lanes carry four COST FUNCTIONS of one trajectory
    # (lib/include/detail/backpropagation.hpp:289-321). With Nout = 4 all lanes do useful work; reported as objective-gradients/s.
    try:
        import numpy as np
        import oracle
        if oracle.reference_available():
            s4 = max(cores, (sample // 8) - (sample // 8) % cores)  # ~1/4 of the main sample's time
            p = oracle.synth_params(oracle.SYS_GLV, N_SPECIES, SEED, 0, s4)
            x0 = oracle.synth_x0(oracle.SYS_GLV, N_SPECIES, p)
            seeds = np.tile(np.eye(4, N_SPECIES), (s4, 1, 1))
            t0 = time.perf_counter()
            oracle.reference_forward_adjoint(oracle.SYS_GLV, N_SPECIES, oracle.RK_CK54, TOL, TOL, x0, p, TI, TF, DT0,
                                             objective=oracle.OBJ_SEED, seeds=seeds, nout=4, threads=cores)
            t4 = time.perf_counter() - t0
            out["lanes_full"] = {"value": 4 * s4 / t4, "unit": "objective-gradients/s", "trajectories_per_s": s4 / t4,
                                 "sample": f"{s4} parameter sets x Nout=4 cost functions (all four AVX lanes of the reference busy); {t4:.2f} s"}
    except Exception as ex:  # the extra figure must never break the contract line
        out["lanes_full"] = {"value": None, "error": repr(ex)[:200]}
    return out, sample, t


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample
    if sample <= 0:
        rate = _calibrate(cores)
        sample = int(max(8 * cores, min(32768, rate * 8)))  # ~8 s per step
        sample -= sample % cores
    for _ in range(args.warmup):
        cpu_run(max(cores, sample // 8), cores)
    t0 = time.perf_counter()
    kind = "port"
    for _ in range(args.steps):
        _, kind = cpu_run(sample, cores)
    dt = time.perf_counter() - t0
    value = args.steps * sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "gradients/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, extra={"reference_step": f"{sample} parameter sets per step (bounded sample of the workload)"}),
        "cpu_baseline": {"value": value, "unit": "gradients/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} parameter sets per step x {args.steps} steps, {cores} host threads"},
        "e2e": {"value": value, "unit": "gradients/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per full-size launch, from the committed
    ncu --set full capture (profiles/r01_traffic_t8.json); None when no capture has been recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic_t8.json")) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch"), d.get("launch_parameter_sets", 1 << 20)
    except Exception:
        return None, None


def workload_config(args, extra=None):
    cfg = {"workload": "GLV N=64 (Npar=4160), 2^20 parameter sets, cash_karp54 controlled rtol=atol=1e-8, t=[0,10], dt0=1e-3, "
                       "J=sum x_i(tf), gradient wrt x0 and all 4160 parameters",
           "batch_total": args.batch, "reduce": args.reduce, "parallelism": f"batch sharded over {args.gpus} GPU(s)",
           "l2": "inputs (33 KB per parameter set, 35 GB per pass) are far larger than L2; no flush needed"}
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.path = f"/tmp/va_clocks_{os.getpid()}.csv"
        self.f = open(self.path, "w")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                   "power_w_max": max(power) if power else None}
        return out


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------

def side_workload(args):
    """The other BASELINE configs (device-resident inputs): same timing rules, one JSON line, not the contract line. Under torchrun
    the batch is sharded over the ranks exactly like the headline workload (no data-path collective; one all_reduce of the summed
    gradient in summed mode), time = max over ranks."""
    import torch
    import torch.distributed as dist
    import vectorizedadjoint_b200 as va
    system, n, stepper, adaptive, tol, ti, tf, dt0, max_steps, objective, stages, desc = WORKLOADS[args.workload]
    if args.species and args.workload == "glv256":
        n = args.species
        desc = f"GLV N={n} (Npar={n * n + n}), cash_karp54 controlled rtol=atol=1e-8, t=[0,10]; not a BASELINE config"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Btot = args.batch if (not args.workload.startswith("glv256") or args.batch != 1 << 20) else (8192 if args.workload == "glv256" else 2048)
    b0, B = va.shard_range(Btot, rank, world)
    npar = va.npar_of(system, n)
    red = va.REDUCE_SUM if args.reduce == "sum" else va.REDUCE_NONE
    f64 = dict(dtype=torch.float64, device=dev)
    params, x0 = torch.empty(B, npar, **f64), torch.empty(B, n, **f64)
    va.synth_batch_device(system, n, SEED, b0, B, params, x0)
    x_final, lam = torch.empty(B, n, **f64), torch.empty(B, 1, n, **f64)
    mu = torch.empty((1, npar) if red == va.REDUCE_SUM else (B, 1, npar), **f64)
    n_acc, n_rej, status = (torch.empty(B, dtype=torch.int32, device=dev) for _ in range(3))
    eng = va.Engine(system, n, stepper, adaptive, tol, tol, device=local, max_steps=max_steps,
                    ckpt_policy={"auto": va.CKPT_AUTO, "recompute": va.CKPT_RECOMPUTE, "store": va.CKPT_STORE_STAGES}[args.ckpt_policy])
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)

    def step():
        eng.call("va_forward_adjoint_batch", B, x0, params, ti, tf, dt0, x_final, lam, mu, objective, red, n_acc, n_rej, status,
                 stream=side.cuda_stream)
        if world > 1 and red == va.REDUCE_SUM:
            dist.all_reduce(mu)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.info()["kernel_launches"]
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item()) / args.steps
    clk = clocks.stop() if clocks else None
    T, R = int(n_acc.sum(dtype=torch.int64)), int(n_rej.sum(dtype=torch.int64))
    info = eng.info()
    line = {"metric": f"fwd+adjoint gradients/sec, {args.workload} batch {Btot}", "value": Btot / (ms * 1e-3), "unit": "gradients/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "batch_total": Btot, "reduce": args.reduce, "parallelism": f"batch sharded over {world} GPU(s)",
                       "ckpt_policy": {va.CKPT_RECOMPUTE: "recompute", va.CKPT_STORE_STAGES: "store_stages"}.get(eng.info()["ckpt_policy"], "auto")},
            "gpu_launches": info["kernel_launches"] - l0,
            "mean_accepted_steps": T / max(B, 1), "mean_rejected": R / max(B, 1), "max_accepted_steps": int(n_acc.max()) if B else 0,
            "failed_trajectories": int((status != 0).sum()), "clocks": clk}
    if system == va.SYS_GLV:
        f_rhs, f_vjp = 2 * n * n + 2 * n, 4 * n * n + 3 * n
        flops = (stages * T + (stages - 1) * R) * f_rhs + stages * T * f_vjp  # this rank's shard
        peak = va.measure_fp64_peak(local)
        line["roofline"] = {"bound": "fp64", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                            "frac": flops / (ms * 1e-3) / 1e12 / peak, "traffic": None, "scope": "rank 0's shard on its GPU"}
        line["kernel"] = info.get("kernel_name")
        if n == 256 and info.get("kernel_name") == "k_glv_ring":
            # ring-streamed kernel: every matrix-vector product re-reads the non-cached rows of the 512 KB matrix from L2/HBM
            cached = 0 if (int(os.environ.get("VA_RING_FLAGS", "2")) & 4) else 64
            products = (stages * T + (stages - 1) * R) + stages * T + B
            sbytes = products * (n - cached) * n * 8
            hbm = None
            try:
                hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            except Exception:
                pass
            line["roofline"]["matrix_stream"] = {"products": products, "bytes_per_product": (n - cached) * n * 8,
                                                 "achieved_gbs": sbytes / (ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm,
                                                 "note": "matrix chunks streamed by TMA per product; source is L2 when the matrices of the "
                                                         "resident CTAs (148 x 512 KB) stay resident, HBM otherwise"}
    else:
        # thread-per-trajectory family: compulsory HBM bytes are the checkpoints, written by the forward kernel and read back
        # by the reverse kernel: 2 * 8 * (N+1) * (T + B) bytes (+ inputs/outputs)
        byts = 2 * 8 * (n + 1) * (T + B) + B * 8 * (npar + 3 * n + npar)
        hbm = None
        try:
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        ach = byts / (ms * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm if hbm else None, "traffic": None,
                            "scope": "rank 0's shard on its GPU",
                            "note": "checkpoint arena traffic 2*8*(N+1) B per accepted step and trajectory"}
    eng.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    if args.workload != "glv64":
        side_workload(args)
        return
    import torch
    import torch.distributed as dist
    import vectorizedadjoint_b200 as va

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"

    N, B = N_SPECIES, args.batch
    b0, Bl = va.shard_range(B, rank, world)
    red = va.REDUCE_SUM if args.reduce == "sum" else va.REDUCE_NONE
    f64 = dict(dtype=torch.float64, device=dev)

    params = torch.empty(Bl, NPAR, **f64)
    x0 = torch.empty(Bl, N, **f64)
    va.synth_batch_device(va.SYS_GLV, N, SEED, b0, Bl, params, x0)
    x_final = torch.empty(Bl, N, **f64)
    lam = torch.empty(Bl, 1, N, **f64)
    mu = torch.empty((1, NPAR) if red == va.REDUCE_SUM else (Bl, 1, NPAR), **f64)
    n_acc = torch.empty(Bl, dtype=torch.int32, device=dev)
    n_rej = torch.empty(Bl, dtype=torch.int32, device=dev)
    status = torch.empty(Bl, dtype=torch.int32, device=dev)
    eng = va.Engine(va.SYS_GLV, N, va.RK_CK54, True, TOL, TOL, device=local)
    info = eng.info()
    side = torch.cuda.Stream(device=dev)  # a real (non-default) stream: kernels, events and NCCL ordering all live on it
    torch.cuda.set_stream(side)
    stream = side.cuda_stream

    def step():
        eng.call("va_forward_adjoint_batch", Bl, x0, params, TI, TF, DT0, x_final, lam, mu, va.OBJ_SUM, red, n_acc, n_rej, status,
                 stream=stream)
        if world > 1 and red == va.REDUCE_SUM:
            dist.all_reduce(mu)  # 33 KB, the only collective on the path

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    launches0 = eng.info()["kernel_launches"]
    clocks = ClockSampler(local) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    clk = clocks.stop() if clocks else None
    launches = eng.info()["kernel_launches"] - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = B * args.steps / (total_ms * 1e-3)

    # work actually done (for the roofline): accepted / rejected steps of this rank's shard
    T_sum = int(n_acc.sum(dtype=torch.int64).item())
    R_sum = int(n_rej.sum(dtype=torch.int64).item())
    bad = int((status != 0).sum().item())
    f_rhs, f_vjp = 2 * N * N + 2 * N, 4 * N * N + 3 * N
    # executed: store-stages policy, no stage recompute in the reverse sweep; K0 is not re-evaluated after a rejection
    flops_exec = (STAGES * T_sum + (STAGES - 1) * R_sum) * f_rhs + STAGES * T_sum * f_vjp
    # the reference's policy (recompute the stages in the reverse sweep), SURVEY.md section 8d formula
    flops_refpolicy = (T_sum + R_sum) * STAGES * f_rhs + T_sum * STAGES * (f_rhs + f_vjp)
    kernel_ms = sorted(step_ms)[len(step_ms) // 2]

    line = {
        "metric": METRIC, "value": value, "unit": "gradients/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args), "gpu_launches": launches,
    }
    if rank == 0:
        peak_tf = va.measure_fp64_peak(local)
        hbm_meas = None
        try:
            hbm_meas = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        ach = flops_exec / (kernel_ms * 1e-3) / 1e12
        alg_bytes = Bl * (8 * NPAR + 8 * N * 3) + (0 if red == va.REDUCE_SUM else Bl * 8 * NPAR)
        line["roofline"] = {
            "bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None,
            "traffic": (lambda tb: int(tb[0] * Bl / tb[1]) if tb[0] else None)(_ncu_traffic()),
            "peak_source": "DFMA microbenchmark run live on this GPU (va_measure_fp64_peak); FP64 is not in MEASURED_PEAKS.json",
            "kernel": "k_glv_t8<TabCK54,adaptive,exact> (va_glv_t8.cu; one launch per step per GPU, %d CTAs x %d threads, 64 threads per "
                      "trajectory; + one row-reduction kernel over the per-slot partial sums)" % (info["sm_count"] * info["ctas_per_sm"], info["threads_per_cta"]),
            "kernel_ms": kernel_ms, "flops_per_launch": flops_exec, "flops_counting": "executed algorithmic FP64 flops of rank 0: "
            "(6T+5R)(2N^2+2N) forward + 6T(4N^2+3N) reverse, store-stages policy (no stage recompute)",
            "achieved_reference_policy_formula": flops_refpolicy / (kernel_ms * 1e-3) / 1e12,
            "mean_accepted_steps": T_sum / max(Bl, 1), "mean_rejected": R_sum / max(Bl, 1), "failed_trajectories": bad,
            "hbm": {"achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_meas,
                    "frac": (alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_meas) if hbm_meas else None,
                    "note": "compulsory bytes only (parameters in, x(tf), dJ/dx0 out); measured DRAM traffic incl. checkpoint blocks spilled from L2 is roofline.traffic"},
        }
        line["clocks"] = clk

    # ---- end to end: host buffers through the C-ABI, copies inside the timed region ------------------------------------
    if not args.no_e2e:
        try:
            hp = torch.empty(Bl, NPAR, dtype=torch.float64, pin_memory=True)
            hx0 = torch.empty(Bl, N, dtype=torch.float64, pin_memory=True)
            hp.copy_(params)
            hx0.copy_(x0)
            hxf = torch.empty(Bl, N, dtype=torch.float64, pin_memory=True)
            hlam = torch.empty(Bl, 1, N, dtype=torch.float64, pin_memory=True)
            hmu = torch.empty((1, NPAR) if red == va.REDUCE_SUM else (Bl, 1, NPAR), dtype=torch.float64, pin_memory=True)
            e2e_steps = max(1, min(args.steps, 3))

            def e2e_step():
                eng.call("va_forward_adjoint_batch", Bl, hx0.numpy(), hp.numpy(), TI, TF, DT0, hxf.numpy(), hlam.numpy(), hmu.numpy(),
                         va.OBJ_SUM, red)
                if world > 1 and red == va.REDUCE_SUM:
                    m = hmu.to(dev, non_blocking=True)
                    dist.all_reduce(m)
                    hmu.copy_(m)

            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            barrier()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            h2d = Bl * 8 * (NPAR + N)
            d2h = Bl * 8 * 2 * N + hmu.numel() * 8
            line["e2e"] = {"value": B * e2e_steps / dt, "unit": "gradients/s", "h2d_bytes_per_step": h2d * world,
                           "d2h_bytes_per_step": d2h * world, "steps": e2e_steps,
                           "note": "va_forward_adjoint_batch with pinned HOST buffers; chunked 3-stream pipeline inside the call"}
            if rank == 0:
                same = bool(torch.equal(hxf.to(dev), x_final))
                line["e2e"]["matches_device_resident_run"] = same
            del hp, hx0, hxf, hlam, hmu
        except Exception as ex:  # pinned memory not available etc.: report, do not hide
            line["e2e"] = {"value": None, "unit": "gradients/s", "error": repr(ex)[:300]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb, _, _ = cpu_baseline(args.cpu_sample)
            line["cpu_baseline"] = cb
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "error": repr(ex)[:300]}
    eng.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
