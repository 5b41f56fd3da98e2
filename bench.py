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
    ap.add_argument("--ckpt-policy", default="auto", choices=["auto", "recompute", "store", "sparse"], help="side workloads: checkpoint policy of the engine")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the side block (the other BASELINE configs, N=1 only)")
    ap.add_argument("--no-check", action="store_true", help="N>1: skip the comparison with a one-GPU run of the whole batch")
    ap.add_argument("--no-parity-sample", action="store_true")
    ap.add_argument("--parity-sample", type=int, default=16384, help="parameter sets compared with the CPU oracle for flip_rate (N=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-traffic-probe", action="store_true", help="do not run the ncu DRAM-counter probe for roofline.traffic")
    ap.add_argument("--traffic-probe", action="store_true", help=argparse.SUPPRESS)  # child mode of the probe (runs under ncu)
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
    """gradients/s of the CPU path from one run of about a second (128 parameter sets per thread). It includes the one-off cost
    (AADC records and JIT-compiles the RHS once per thread, ~0.2 s at N=64), so it errs low and the sample sized from it runs
    10-15 s. (Round 2: two runs of 2 and 6 sets per thread and their difference were too short to time -- 9632 sets, 4.4 s.)"""
    t, _ = cpu_run(128 * cores, cores)
    return 128 * cores / max(t, 1e-3)


def cpu_baseline(sample=0):
    cores = os.cpu_count() or 1
    if sample <= 0:
        rate = _calibrate(cores)
        sample = int(max(8 * cores, min(65536, rate * 30)))  # 10-30 s of CPU work: the one-second calibration errs low by up to 2.5x
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
                    ckpt_policy={"auto": va.CKPT_AUTO, "recompute": va.CKPT_RECOMPUTE, "store": va.CKPT_STORE_STAGES, "sparse": va.CKPT_SPARSE}[args.ckpt_policy])
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
                       "ckpt_policy": {va.CKPT_RECOMPUTE: "recompute", va.CKPT_STORE_STAGES: "store_stages", va.CKPT_SPARSE: "sparse"}.get(eng.info()["ckpt_policy"], "auto")},
            "gpu_launches": info["kernel_launches"] - l0,
            "mean_accepted_steps": T / max(B, 1), "mean_rejected": R / max(B, 1), "max_accepted_steps": int(n_acc.max()) if B else 0,
            "failed_trajectories": int((status != 0).sum()), "clocks": clk}
    if system == va.SYS_GLV:
        f_rhs, f_vjp = 2 * n * n + 2 * n, 4 * n * n + 3 * n
        recompute = eng.info()["ckpt_policy"] in (va.CKPT_RECOMPUTE, va.CKPT_SPARSE)  # the stages are evaluated again in the reverse sweep (executed work)
        flops = (stages * T + (stages - 1) * R + (stages * T if recompute else 0)) * f_rhs + stages * T * f_vjp  # this rank's shard
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


# ---------------------------------------------------------------------------------------------------------------------
# live counters: NVML GPM (DRAM bandwidth / FP64 pipe utilisation over an interval), the bench's own regression tripwire
# ---------------------------------------------------------------------------------------------------------------------

class GpmSampler:
    """Two GPM samples around the timed region -> average DRAM bandwidth utilisation and FP64 pipe utilisation in between
    (hardware counters sampled by the driver, no replay, no serialisation: valid next to a timed run)."""
    DRAM_PEAK_BPS = 3996e6 * 2 * 8192 / 8  # HBM3e: 3996 MHz double data rate, 8192-bit bus = 8.18 TB/s, what DRAM_BW_UTIL is relative to

    def __init__(self, index):
        self.ok = False
        try:
            import pynvml as nv
            self.nv = nv
            nv.nvmlInit()
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            if not nv.nvmlGpmQueryDeviceSupport(self.h).isSupportedDevice:
                return
            self.s0, self.s1 = nv.nvmlGpmSampleAlloc(), nv.nvmlGpmSampleAlloc()
            self.ok = True
        except Exception:
            self.ok = False

    def start(self):
        if self.ok:
            try:
                self.nv.nvmlGpmSampleGet(self.h, self.s0)
            except Exception:
                self.ok = False

    def stop(self):
        if not self.ok:
            return None
        nv = self.nv
        try:
            nv.nvmlGpmSampleGet(self.h, self.s1)
            mg = nv.c_nvmlGpmMetricsGet_t()
            mg.version = nv.NVML_GPM_METRICS_GET_VERSION
            ids = [nv.NVML_GPM_METRIC_DRAM_BW_UTIL, nv.NVML_GPM_METRIC_FP64_UTIL, nv.NVML_GPM_METRIC_SM_UTIL]
            mg.numMetrics = len(ids)
            mg.sample1, mg.sample2 = self.s0, self.s1
            for k, i in enumerate(ids):
                mg.metrics[k].metricId = i
            nv.nvmlGpmMetricsGet(mg)
            vals = [mg.metrics[k].value if mg.metrics[k].nvmlReturn == 0 else None for k in range(len(ids))]
            return {"dram_bw_util_pct": vals[0], "fp64_util_pct": vals[1], "sm_util_pct": vals[2]}
        except Exception as ex:
            return {"error": repr(ex)[:120]}


PROBE_SETS = 1 << 17  # parameter sets of the traffic probe launch (1/8 of the workload: same slots in flight, same per-set traffic)


def traffic_probe_child():
    """Child of live_traffic(): two launches of the headline kernel on PROBE_SETS device-resident parameter sets."""
    import torch
    import vectorizedadjoint_b200 as va
    sh = DeviceShard(torch, va, 0, 0, PROBE_SETS, va.REDUCE_SUM)
    with va.Engine(va.SYS_GLV, N_SPECIES, va.RK_CK54, True, TOL, TOL, device=0) as eng:
        for _ in range(2):
            a = sh.args(va, va.REDUCE_SUM)
            Bn = a.pop("B")
            eng.call("va_forward_adjoint_batch", Bn, a.pop("x0"), a.pop("params"), a.pop("ti"), a.pop("tf"), a.pop("dt0"), a.pop("x_final"),
                     a.pop("lam"), a.pop("mu"), **a)
        torch.cuda.synchronize()


def live_traffic(device):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the headline kernel, measured now with ncu's DRAM counters in a
    separate process (second launch of the probe, after a warm-up launch; nothing timed runs under the profiler). Returns
    (bytes per parameter set, description) or (None, reason)."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=(vis.split(",")[device] if vis else str(device)))  # the child's device 0 = this GPU
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:k_glv_t8", "--launch-skip", "1",
           "--launch-count", "1", "--csv", sys.executable, os.path.abspath(__file__), "--traffic-probe"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env).stdout
    except Exception as ex:
        return None, repr(ex)[:120]
    import csv
    import io
    total = 0.0
    found = 0
    for row in csv.reader(io.StringIO(out)):
        if len(row) > 3 and row[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            unit, val = row[-2].lower(), float(row[-1].replace(",", ""))
            total += val * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit, 1.0)
            found += 1
    if found != 2:
        return None, "ncu gave no DRAM counters: " + out[-200:].replace("\n", " ")
    return total / PROBE_SETS, (f"live: ncu dram__bytes_read.sum + dram__bytes_write.sum of one k_glv_t8 launch over {PROBE_SETS} parameter sets, taken in a "
                                "separate process right after the timed run, scaled to this launch's parameter sets (per-set traffic does not depend "
                                "on the batch once it exceeds the 592 resident slots)")


def _committed_traffic():
    """DRAM bytes per full-size launch from the newest committed ncu --set full capture (fallback when GPM is unavailable)."""
    best = (None, None, None)
    for name in ("r02_traffic_t8.json", "r01_traffic_t8.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
            return d.get("dram_bytes_per_launch"), d.get("launch_parameter_sets", 1 << 20), name
        except Exception:
            continue
    return best


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm, headline workload
# ---------------------------------------------------------------------------------------------------------------------

class DeviceShard:
    """One GPU's part of the batch, resident in its HBM."""

    def __init__(self, torch, va, device, b0, count, red):
        self.device, self.b0, self.B = device, b0, count
        dev = torch.device("cuda", device)
        f64 = dict(dtype=torch.float64, device=dev)
        self.params = torch.empty(count, NPAR, **f64)
        self.x0 = torch.empty(count, N_SPECIES, **f64)
        with torch.cuda.device(dev):
            va.synth_batch_device(va.SYS_GLV, N_SPECIES, SEED, b0, count, self.params, self.x0)
            torch.cuda.synchronize()
        self.x_final = torch.empty(count, N_SPECIES, **f64)
        self.lam = torch.empty(count, 1, N_SPECIES, **f64)
        self.mu = torch.empty((1, NPAR) if red == va.REDUCE_SUM else (count, 1, NPAR), **f64)
        self.n_acc, self.n_rej, self.status = (torch.empty(count, dtype=torch.int32, device=dev) for _ in range(3))
        self.stream = torch.cuda.Stream(device=dev)  # a real (non-default) stream: kernels, events and NCCL ordering live on it

    def args(self, va, red):
        return dict(B=self.B, x0=self.x0, params=self.params, ti=TI, tf=TF, dt0=DT0, x_final=self.x_final, lam=self.lam, mu=self.mu,
                    objective=va.OBJ_SUM, reduce=red, n_accept=self.n_acc, n_reject=self.n_rej, status=self.status,
                    stream=self.stream.cuda_stream)


def parity_sample(va, device, sample=16384):
    """Accept/reject flips and deviations against the CPU oracle on the first `sample` seeded parameter sets (checker leg, not
    timed): north_star allows 'documented rounding-induced flips'; this reports how many there are and how far a flipped
    trajectory ends up from the oracle's."""
    import numpy as np
    import oracle
    p = oracle.synth_params(oracle.SYS_GLV, N_SPECIES, SEED, 0, sample)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N_SPECIES, p)
    with va.Engine(va.SYS_GLV, N_SPECIES, va.RK_CK54, True, TOL, TOL, device=device) as e:
        g = e.forward_adjoint(x0, p, TI, TF, DT0, objective=va.OBJ_SUM)
    o = oracle.forward_adjoint(oracle.SYS_GLV, N_SPECIES, oracle.RK_CK54, True, TOL, TOL, x0, p, TI, TF, DT0, objective=oracle.OBJ_SUM,
                               threads=os.cpu_count() or 1)
    flipped = (g["n_accept"] != o["n_accept"]) | (g["n_reject"] != o["n_reject"])

    def rel(a, b, m):
        if not m.any():
            return 0.0
        a, b = a[m].reshape(int(m.sum()), -1), b[m].reshape(int(m.sum()), -1)
        return float((np.abs(a - b).max(axis=1) / np.abs(b).max(axis=1)).max())
    out = {"sets": sample, "flips": int(flipped.sum()), "flip_rate": float(flipped.mean()), "oracle": "oracle/va_oracle.c (C port, pinned to reference fixtures)"}
    for k, gk, ok in (("x_final", g["x_final"], o["x_final"]), ("dJ_dx0", g["lam"][:, 0], o["lam"]), ("dJ_dalpha", g["mu"][:, 0], o["mu"])):
        out[f"max_rel_err_{k}"] = rel(gk, ok, ~flipped)
        out[f"max_rel_err_{k}_flipped"] = rel(gk, ok, flipped)
    return out


def side_block(va, torch, device):
    """The other BASELINE configs on this GPU, a few hundred ms each (device-resident, same timing rules): value + roofline fraction."""
    out = {}
    for name, batch, steps in (("vdp", 1 << 20, 5), ("glv16", 1 << 20, 5), ("glv256", 8192, 3), ("glv256long", 1024, 2)):
        try:
            out[name] = side_measure(va, torch, device, name, batch, steps)
        except Exception as ex:
            out[name] = {"value": None, "error": repr(ex)[:200]}
    return out


def side_measure(va, torch, device, name, Btot, steps, warmup=2):
    system, n, stepper, adaptive, tol, ti, tf, dt0, max_steps, objective, stages, desc = WORKLOADS[name]
    dev = torch.device("cuda", device)
    npar = va.npar_of(system, n)
    f64 = dict(dtype=torch.float64, device=dev)
    params, x0 = torch.empty(Btot, npar, **f64), torch.empty(Btot, n, **f64)
    va.synth_batch_device(system, n, SEED, 0, Btot, params, x0)
    x_final, lam, mu = torch.empty(Btot, n, **f64), torch.empty(Btot, 1, n, **f64), torch.empty(1, npar, **f64)
    n_acc, n_rej, status = (torch.empty(Btot, dtype=torch.int32, device=dev) for _ in range(3))
    st = torch.cuda.current_stream(dev)
    with va.Engine(system, n, stepper, adaptive, tol, tol, device=device, max_steps=max_steps) as eng:
        def step():
            eng.call("va_forward_adjoint_batch", Btot, x0, params, ti, tf, dt0, x_final, lam, mu, objective, va.REDUCE_SUM, n_acc, n_rej, status,
                     stream=st.cuda_stream)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            step()
        e1.record(st)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        kernel = eng.info()["kernel_name"]
        recompute = eng.info()["ckpt_policy"] in (va.CKPT_RECOMPUTE, va.CKPT_SPARSE)
    T, R = int(n_acc.sum(dtype=torch.int64)), int(n_rej.sum(dtype=torch.int64))
    res = {"value": Btot / (ms * 1e-3), "unit": "gradients/s", "batch": Btot, "ms_per_step": ms, "steps": steps, "kernel": kernel,
           "mean_accepted_steps": T / Btot, "failed_trajectories": int((status != 0).sum())}
    if system == va.SYS_GLV:
        # executed algorithmic flops: forward (6T+5R) products, reverse 6T vector-Jacobian products, and -- under the recompute
        # policy (the reference's) -- the 6T stage products evaluated again in the reverse sweep
        flops = (stages * T + (stages - 1) * R + (stages * T if recompute else 0)) * (2 * n * n + 2 * n) + stages * T * (4 * n * n + 3 * n)
        res.update(bound="fp64", achieved_tflops=flops / (ms * 1e-3) / 1e12, ckpt_policy="recompute" if recompute else "store_stages",
                   useful_tflops_store_stages_formula=((stages * T + (stages - 1) * R) * (2 * n * n + 2 * n) + stages * T * (4 * n * n + 3 * n)) / (ms * 1e-3) / 1e12)
    else:
        byts = 2 * 8 * (n + 1) * (T + Btot) + Btot * 8 * (npar + 3 * n + npar)
        res.update(bound="hbm", achieved_gbs=byts / (ms * 1e-3) / 1e9)
    return res


def main():
    args = parse()
    if args.traffic_probe:
        traffic_probe_child()
        return
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
    # Two ways to use N GPUs, both through the C-ABI (include/va_engine.h "Several GPUs"):
    #   torchrun, one rank per GPU  -> single-device engine + va_engine_comm_init; the all-reduce happens inside the call
    #   one process, --gpus N       -> multi-device engine (va_engine_desc.devices), caller-sharded device-resident call
    if world > 1:
        assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"
        mode, devices, total = "ranks", [local], world
    elif args.gpus > 1:
        assert torch.cuda.device_count() >= args.gpus, f"--gpus {args.gpus} but {torch.cuda.device_count()} visible"
        mode, devices, total = "devices", list(range(args.gpus)), args.gpus
    else:
        mode, devices, total = "single", [local], 1
    torch.cuda.set_device(devices[0])
    dev0 = torch.device("cuda", devices[0])
    if world > 1:
        dist.init_process_group("nccl", device_id=dev0)  # plumbing: barrier, max-over-ranks of the timings, id broadcast

    N, B = N_SPECIES, args.batch
    red = va.REDUCE_SUM if args.reduce == "sum" else va.REDUCE_NONE
    shards = []
    for k, d in enumerate(devices):
        b0, cnt = va.shard_range(B, rank if mode == "ranks" else k, total)
        shards.append(DeviceShard(torch, va, d, b0, cnt, red))
    if mode == "devices":
        eng = va.Engine(va.SYS_GLV, N, va.RK_CK54, True, TOL, TOL, devices=devices)
    else:
        eng = va.Engine(va.SYS_GLV, N, va.RK_CK54, True, TOL, TOL, device=devices[0])
        if mode == "ranks":
            idt = torch.zeros(va.COMM_ID_BYTES, dtype=torch.uint8, device=dev0)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(va.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            eng.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    info = eng.info()
    sh_args = [s.args(va, red) for s in shards]

    def step():
        # ONE C-ABI call per step; with several GPUs the summed gradient is all-reduced inside it (ncclAllReduce, 33 KB)
        if mode == "devices":
            eng.call_sharded(sh_args)
        else:
            a = dict(sh_args[0])
            Bn = a.pop("B")
            eng.call("va_forward_adjoint_batch", Bn, a.pop("x0"), a.pop("params"), a.pop("ti"), a.pop("tf"), a.pop("dt0"), a.pop("x_final"),
                     a.pop("lam"), a.pop("mu"), **a)

    def barrier():
        for s in shards:
            torch.cuda.synchronize(s.device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    launches0, coll0 = eng.info()["kernel_launches"], eng.info()["collectives"]
    clocks = ClockSampler(devices[0]) if rank == 0 else None
    gpm = GpmSampler(devices[0]) if rank == 0 else None
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)] for _ in shards]
    barrier()
    if gpm:
        gpm.start()
    for s, e in zip(shards, ev):
        e[0].record(s.stream)
    for k in range(args.steps):
        step()
        for s, e in zip(shards, ev):
            e[k + 1].record(s.stream)
    barrier()
    gpm_vals = gpm.stop() if gpm else None
    total_ms = max(e[0].elapsed_time(e[-1]) for e in ev)  # slowest GPU of this process
    step_ms = [ev[0][k].elapsed_time(ev[0][k + 1]) for k in range(args.steps)]
    clk = clocks.stop() if clocks else None
    launches = eng.info()["kernel_launches"] - launches0
    collectives = eng.info()["collectives"] - coll0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev0)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = B * args.steps / (total_ms * 1e-3)

    # work actually done (for the roofline): accepted / rejected steps of the first shard of this process
    s0 = shards[0]
    Bl = s0.B
    T_sum = int(s0.n_acc.sum(dtype=torch.int64).item())
    R_sum = int(s0.n_rej.sum(dtype=torch.int64).item())
    bad = int((s0.status != 0).sum().item())
    f_rhs, f_vjp = 2 * N * N + 2 * N, 4 * N * N + 3 * N
    # executed: store-stages policy, no stage recompute in the reverse sweep; K0 is not re-evaluated after a rejection
    flops_exec = (STAGES * T_sum + (STAGES - 1) * R_sum) * f_rhs + STAGES * T_sum * f_vjp
    # the reference's policy (recompute the stages in the reverse sweep), SURVEY.md section 8d formula
    flops_refpolicy = (T_sum + R_sum) * STAGES * f_rhs + T_sum * STAGES * (f_rhs + f_vjp)
    kernel_ms = sorted(step_ms)[len(step_ms) // 2]

    line = {
        "metric": METRIC, "value": value, "unit": "gradients/s", "n_gpus": total, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args), "gpu_launches": launches * (world if mode == "ranks" else 1),
    }
    line["config"]["multi_gpu"] = {"single": "one GPU", "ranks": f"{world} processes, one single-device engine each, attached with va_engine_comm_init",
                                   "devices": f"one process, one multi-device engine over {total} GPUs (va_engine_desc.devices)"}[mode]
    if total > 1:
        line["collective"] = {"where": "inside va_forward_adjoint_batch (libva_engine.so: ncclAllReduce, ncclDouble, ncclSum, 33 KB)",
                              "calls_in_timed_region_per_gpu": collectives // max(len(shards), 1), "nccl_version": eng.info()["nccl_version"]}
    if rank == 0:
        peak_tf = va.measure_fp64_peak(devices[0])
        hbm_meas = None
        try:
            hbm_meas = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        ach = flops_exec / (kernel_ms * 1e-3) / 1e12
        alg_bytes = Bl * (8 * NPAR + 8 * N * 3) + (0 if red == va.REDUCE_SUM else Bl * 8 * NPAR)
        traffic, traffic_source = None, None
        if not args.no_traffic_probe and total == 1:
            per_set, why = live_traffic(devices[0])
            if per_set:
                traffic, traffic_source = int(per_set * Bl), why
            else:
                traffic_source = why
        if traffic is not None:
            pass
        elif gpm_vals and gpm_vals.get("dram_bw_util_pct") is not None:
            traffic = int(gpm_vals["dram_bw_util_pct"] / 100.0 * GpmSampler.DRAM_PEAK_BPS * kernel_ms * 1e-3)
            traffic_source = ("live: NVML GPM DRAM_BW_UTIL averaged over the timed region x 8.18 TB/s (3996 MHz x 2 x 8192 bit) x the launch "
                              "duration; cross-checked against the ncu --set full capture in profiles/")
        else:
            tb = _committed_traffic()
            if tb[0]:
                traffic, traffic_source = int(tb[0] * Bl / tb[1]), (f"committed ncu --set full capture profiles/{tb[2]} scaled to this shard "
                                                                   f"(live probe: {traffic_source or 'not run at N > 1'})")
        line["roofline"] = {
            "bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None,
            "traffic": traffic, "traffic_source": traffic_source, "algorithmic_bytes": alg_bytes, "gpm": gpm_vals,
            "peak_source": "DFMA microbenchmark run live on this GPU (va_measure_fp64_peak); FP64 is not in MEASURED_PEAKS.json",
            "kernel": "%s<TabCK54,adaptive,exact%s> (one launch per step per GPU, %d CTAs x %d threads, 64 threads per "
                      "trajectory; + one row-reduction kernel over the per-slot partial sums)" % (
                          info["kernel_name"], "" if info["kernel_name"] != "k_glv_t8" or os.environ.get("VA_T8_NO_STAGE", "0") not in ("", "0")
                          else ",parameters staged in shared memory by TMA", info["sm_count"] * info["ctas_per_sm"], info["threads_per_cta"]),
            "kernel_ms": kernel_ms, "flops_per_launch": flops_exec, "flops_counting": "executed algorithmic FP64 flops of rank 0: "
            "(6T+5R)(2N^2+2N) forward + 6T(4N^2+3N) reverse, store-stages policy (no stage recompute)",
            "achieved_reference_policy_formula": flops_refpolicy / (kernel_ms * 1e-3) / 1e12,
            "mean_accepted_steps": T_sum / max(Bl, 1), "mean_rejected": R_sum / max(Bl, 1), "failed_trajectories": bad,
            "hbm": {"achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_meas,
                    "frac": (alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_meas) if hbm_meas else None,
                    "note": "compulsory bytes only (parameters in, x(tf), dJ/dx0 out); measured DRAM traffic incl. checkpoint blocks spilled from L2 is roofline.traffic"},
        }
        line["clocks"] = clk

    # ---- N > 1: is the sharded result the one-GPU result? (outside the timed region) ----------------------------------
    if total > 1 and red == va.REDUCE_SUM and not args.no_check:
        try:
            acc_tot = torch.stack([s.n_acc.sum(dtype=torch.int64).to(dev0) for s in shards]).sum()
            xf_tot = torch.stack([s.x_final.sum().to(dev0) for s in shards]).sum()
            if world > 1:
                dist.all_reduce(acc_tot)
                dist.all_reduce(xf_tot)
            if rank == 0:
                chk = DeviceShard(torch, va, devices[0], 0, B, red)  # the whole batch on ONE GPU, its own engine, no communicator
                with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, TOL, TOL, device=devices[0]) as e1:
                    a = chk.args(va, red)
                    Bn = a.pop("B")
                    e1.call("va_forward_adjoint_batch", Bn, a.pop("x0"), a.pop("params"), a.pop("ti"), a.pop("tf"), a.pop("dt0"), a.pop("x_final"),
                            a.pop("lam"), a.pop("mu"), **a)
                    torch.cuda.synchronize(devices[0])
                scale = float(chk.mu.abs().max())
                line["multi_gpu_check"] = {
                    "what": f"all-reduced summed gradient of the {total}-GPU run vs the same 2^20 parameter sets on one GPU (separate engine, no communicator)",
                    "mu_max_rel_diff": float((s0.mu - chk.mu).abs().max()) / scale,
                    "first_shard_x_final_bit_identical": bool(torch.equal(s0.x_final, chk.x_final[s0.b0:s0.b0 + s0.B])),
                    "first_shard_dJ_dx0_bit_identical": bool(torch.equal(s0.lam, chk.lam[s0.b0:s0.b0 + s0.B])),
                    "accepted_steps_total_equal": int(acc_tot.item()) == int(chk.n_acc.sum(dtype=torch.int64).item()),
                    "x_final_checksum_rel_diff": abs(float(xf_tot.item()) - float(chk.x_final.sum().item())) / abs(float(chk.x_final.sum().item())),
                }
                del chk
                torch.cuda.empty_cache()
        except Exception as ex:
            line["multi_gpu_check"] = {"error": repr(ex)[:300]}

    # ---- end to end: host buffers through the C-ABI, copies inside the timed region ------------------------------------
    if not args.no_e2e:
        try:
            line["e2e"] = e2e_leg(args, torch, dist, va, eng, shards, mode, world, rank, total, red, barrier)
        except Exception as ex:  # pinned memory not available etc.: report, do not hide
            line["e2e"] = {"value": None, "unit": "gradients/s", "error": repr(ex)[:300]}
    eng.close()
    del shards, sh_args, s0
    torch.cuda.empty_cache()

    if rank == 0 and total == 1:
        if not args.no_parity_sample:
            try:
                line["parity_sample"] = parity_sample(va, devices[0], args.parity_sample)
                line["flip_rate"] = line["parity_sample"]["flip_rate"]
            except Exception as ex:
                line["parity_sample"] = {"error": repr(ex)[:300]}
        if not args.no_side:
            line["side"] = side_block(va, torch, devices[0])
            pk = line.get("roofline", {}).get("peak")
            hb = line.get("roofline", {}).get("hbm", {}).get("peak_gbs")
            for v in line["side"].values():
                if v.get("bound") == "fp64" and pk:
                    v["frac"] = v["achieved_tflops"] / pk
                elif v.get("bound") == "hbm" and hb:
                    v["frac"] = v["achieved_gbs"] / hb
        if not args.no_cpu_baseline:
            try:
                cb, _, _ = cpu_baseline(args.cpu_sample)
                line["cpu_baseline"] = cb
            except Exception as ex:
                line["cpu_baseline"] = {"value": None, "error": repr(ex)[:300]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def e2e_leg(args, torch, dist, va, eng, shards, mode, world, rank, total, red, barrier):
    """The same pass through the reference-facing C-ABI call with HOST buffers (page-locked, from va_host_alloc): host->device
    copies of the parameters and device->host copies of the results are inside the timed region, every step."""
    import numpy as np
    N = N_SPECIES
    Bl = sum(s.B for s in shards)  # this process's parameter sets (all of them in one-process modes)
    flags = int(os.environ.get("VA_BENCH_HOST_FLAGS", str(va.HOST_NUMA_LOCAL)))
    dev0 = shards[0].device
    hp = va.host_alloc((Bl, NPAR), flags=flags, device=dev0)
    hx0 = va.host_alloc((Bl, N), flags=flags, device=dev0)
    off = 0
    for s in shards:  # fill the host inputs from the device-resident ones (not timed)
        torch.from_numpy(hp.array[off:off + s.B]).copy_(s.params)
        torch.from_numpy(hx0.array[off:off + s.B]).copy_(s.x0)
        off += s.B
    hxf = va.host_alloc((Bl, N), device=dev0)
    hlam = va.host_alloc((Bl, 1, N), device=dev0)
    hmu = va.host_alloc((1, NPAR) if red == va.REDUCE_SUM else (Bl, 1, NPAR), device=dev0)
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():  # one C-ABI call; a multi-device engine splits the host batch itself, ranks all-reduce inside the call
        eng.call("va_forward_adjoint_batch", Bl, hx0.array, hp.array, TI, TF, DT0, hxf.array, hlam.array, hmu.array, va.OBJ_SUM, red)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    dt = time.perf_counter() - t0
    dev = torch.device("cuda", dev0)
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    ranks = world if mode == "ranks" else 1
    h2d = Bl * 8 * (NPAR + N) * ranks
    d2h = (Bl * 8 * 2 * N + hmu.array.size * 8) * ranks
    out = {"value": args.batch * e2e_steps / dt, "unit": "gradients/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
           "note": "va_forward_adjoint_batch with page-locked HOST buffers (va_host_alloc); chunked 3-stream pipeline per GPU inside the call"}
    if rank == 0:
        out["matches_device_resident_run"] = bool(np.array_equal(hxf.array[:shards[0].B], shards[0].x_final.cpu().numpy()))
    # the ceiling: what the box delivers host->device with all GPUs of the run copying CONCURRENTLY from page-locked memory of the
    # same kind. One process (rank 0) drives all of them at once so that the copies really overlap (every other rank waits at the
    # barrier); the figure is total bytes / wall time of the slowest copy, best of 3 -- not a sum of per-GPU bests.
    try:
        barrier()
        if rank == 0:
            run_devices = list(range(world)) if mode == "ranks" else [s.device for s in shards]
            per, agg = va.measure_h2d_copy(run_devices, nbytes=1 << 30, reps=3, flags=flags)
            ach = h2d / (dt / e2e_steps) / 1e9
            out["roofline"] = {"bound": "h2d", "achieved": ach, "peak": agg, "unit": "GB/s", "frac": ach / agg,
                               "peak_source": f"va_measure_h2d_copy: 1 GiB per GPU from page-locked host memory to all {total} GPU(s) of this run "
                                              "concurrently (one process driving them all), total bytes / wall time, best of 3; measured in this run",
                               "per_gpu_gbs": per}
        barrier()
    except Exception as ex:
        out["roofline"] = {"error": repr(ex)[:200]}
    for h in (hp, hx0, hxf, hlam, hmu):
        h.free()
    return out


if __name__ == "__main__":
    main()
