#!/usr/bin/env python
"""bench.py -- throughput of the per-step hot path in particle-steps/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n PARTICLES]

Workload (config.workload): C5 of SURVEY.md 8d -- 3-D monodisperse pseudo-hard-sphere fluid, phi = 0.47, NVE,
dt = 1e-3, N = 2^24 (the configuration the north-star target is quoted on; it fits one B200), started from a jittered
lattice and melted by --melt NVT steps before anything is timed.  A "step" is one velocity-Verlet step over all N
particles (kick-drift-wrap, conditional neighbour rebuild, pair forces + second kick + thermo reduction).
State arrays (1.5 GiB) are far larger than L2 (126 MB), so no explicit L2 flush is needed between timed steps.

One JSON line on stdout (rank 0).  `value` = device-resident throughput (CUDA events on the engine's stream, max over
ranks); `e2e` = the same metric through the public host-buffer API (upload from pinned host memory + K steps + download
of the final state and thermo rows inside the timed region); `roofline` = dominant kernel vs the measured HBM copy
bandwidth; `cpu_baseline` / `--impl reference` = the reference-shaped OpenMP port of the Julia CPU path (oracle/), the
only thing the reference can be represented by here (no Julia in the image).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

import numpy as np  # noqa: E402

DT = 1e-3
KT = 1.4737
PHI = 0.47
CUTOFF = 1.5
# algorithmic bytes (DESIGN.md "Measurement"): every state array read once and written once per step
BYTES_STEP_3D = 176.0
# dominant kernel = K4 pair forces with fused second kick: reads x 24 + sigma 8 + v 24, writes f 24 + v 24
BYTES_FORCE_KERNEL_3D = 104.0
# fused NVE step (default in list mode): the same kernel also does the next step's kick-drift-wrap:
# reads x 24 + sigma 8 + v 24, writes f 24 + v 24 + x 24 + sigma 8 (image counters only on a crossing)
BYTES_FORCE_FUSED_3D = 136.0
# K5 kick-drift-wrap: reads x 24 + sigma 8 (same 32 B record) + v 24 + f 24 + img 12, writes x 24 (+8) + v 24 + img 12
BYTES_KICK_KERNEL_3D = 144.0


def traffic_file():
    """the newest profiles/rNN_traffic.json (ncu-derived numbers of the dominant kernels, one file per round)"""
    d = os.path.join(ROOT, "profiles")
    try:
        names = sorted(f for f in os.listdir(d) if f.startswith("r") and f.endswith("_traffic.json"))
    except OSError:
        names = []
    return os.path.join(d, names[-1]) if names else None


def measured_traffic(kernel, n, build=None):
    """(bytes, note): dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
    ncu --set full capture, scaled by particle count when the run is not at the captured size.  The capture names the
    build it was taken on (registers, stack bytes, kernel variant of the fused force kernel); numbers captured on a
    DIFFERENT build than the loaded library are refused (None), not quoted."""
    p = traffic_file()
    if not p:
        return None, "no profiles/rNN_traffic.json"
    try:
        with open(p) as fh:
            doc = json.load(fh)
        t = doc[kernel]
    except Exception as exc:
        return None, "%s: %s" % (os.path.basename(p), exc)
    want = t.get("build")
    if want is not None and build is not None:
        diff = {k: (want[k], build.get(k)) for k in want if build.get(k) != want[k]}
        if diff:
            return None, "%s was captured on another build of the kernel (capture vs loaded: %s): stale, not quoted" % (os.path.basename(p), diff)
    return (t["dram_bytes_read"] + t["dram_bytes_write"]) * n / t["n_particles"], os.path.basename(p)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.

    NVML is opened ONCE, at program start (`open()`: nvmlInit and the device handles cost tens of milliseconds and take
    driver-wide locks, so they must never fall inside or next to a timed region), and polled from a thread while the
    timed region runs: first sample 0.5 ms after `start()`, then every `period` seconds.  Multi-GPU runs poll from rank 0
    only, for every GPU of the job (one process talking to NVML instead of N: round 1's 8-GPU line lost 36 ms to eight
    processes polling every 2 ms).  nvidia-smi is the fallback when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, devices, period=0.005, enabled=True):
        self.devices = list(devices)
        self.period = period
        self.enabled = enabled
        self.nv = None
        self.handles = []
        self.proc = None
        self.path = None
        self.thread = None
        self.samples = []   # (sm_mhz, reason bitmask)
        self.sm_max = None

    def open(self):
        if not self.enabled:
            return self
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.handles = [nv.nvmlDeviceGetHandleByIndex(d) for d in self.devices]
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(self.handles[0], nv.NVML_CLOCK_SM))
            for h in self.handles:   # first query of a handle is slower than the following ones
                nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.nv = nv
        except Exception:
            self.nv = None
        return self

    def start(self):
        self.samples = []
        if not self.enabled:
            return
        if self.nv is not None:
            nv, handles = self.nv, self.handles
            self._stop = threading.Event()

            def loop():
                if self._stop.wait(0.0005):
                    return
                while True:
                    for h in handles:
                        try:
                            self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                                 int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))))
                        except Exception:
                            pass
                    if self._stop.wait(self.period):
                        break
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        try:
            fd, self.path = tempfile.mkstemp(prefix="mdb_clocks_", suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(d) for d in self.devices), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.enabled:
            out["source"] = "sampled on rank 0 for every GPU of the job"
            return out
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            self.thread = None
            nv = self.nv
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            if self.samples:
                mask = 0
                for _, m in self.samples:
                    mask |= m
                out.update(sm_mhz=statistics.median([c for c, _ in self.samples]), sm_max_mhz=self.sm_max,
                           reasons=sorted(k for k, b in bits.items() if mask & b), samples=len(self.samples), source="nvml",
                           gpus_sampled=len(self.handles), period_ms=1e3 * self.period)
            return out
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as fh:
            for line in fh:
                c = [t.strip() for t in line.split(",")]
                if len(c) < 8:
                    continue
                try:
                    sm.append(float(c[0]))
                    mx.append(float(c[1]))
                except ValueError:
                    continue
                for nm, val in zip(names, c[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")
        return out


def nvml_indices(cuda_ordinals):
    """NVML index of each CUDA ordinal (CUDA_VISIBLE_DEVICES remaps the latter)"""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            table = [int(t) for t in vis.split(",") if t.strip() != ""]
            return [table[d] for d in cuda_ordinals]
        except (ValueError, IndexError):
            pass
    return list(cuda_ordinals)


def make_workload(n, seed_shift=0):
    from mdjl_b200 import workloads
    cfg = workloads.phs_fluid(n, phi=PHI, dim=3, seed=workloads.BASE_SEED + seed_shift)
    v0 = workloads.velocities(n, 3, KT, seed=workloads.BASE_SEED + seed_shift)
    return cfg, v0


def cpu_port_rate(n_sample, steps, warmup, x=None, v=None, box=None):
    """reference-shaped OpenMP port (oracle/md_oracle.c orc_run_timing) on the host cores: particle-steps/s"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mdoracle as orc
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1 for its workers; in the reference arm the
    # other ranks exit at once, so rank 0 has the box to itself)
    try:
        orc.set_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        orc.set_threads(os.cpu_count() or 1)
    if x is None:
        cfg, v = make_workload(n_sample)
        x, box = cfg["x"], cfg["box"]
        # melt a little so pairs interact (the port's cost depends weakly on it)
        pre = 20
    else:
        pre = 0
    n = x.shape[0]
    x, v = np.array(x), np.array(v)
    f = np.zeros_like(x)
    img = np.zeros((n, 3), np.int32)
    diam = np.ones(n)
    if pre + warmup > 0:
        orc.run_timing(orc.NVE, x, v, f, img, diam, box, CUTOFF, orc.POT_PHS, (), DT, pre + warmup)
    t0 = time.perf_counter()
    orc.run_timing(orc.NVE, x, v, f, img, diam, box, CUTOFF, orc.POT_PHS, (), DT, steps)
    t = time.perf_counter() - t0
    return n * steps / t, t, orc.threads()


def run_reference(args, rank, world):
    """--impl reference: the Julia package cannot run here (no julia binary, CellListMap/StaticArrays/... absent, no
    network), so the reference arm is the reference-shaped OpenMP port in oracle/ on all host threads."""
    if rank != 0:
        return
    # the same configuration as our arm (N = args.n, default 2^24): one step of the port takes ~0.8 s on 16 threads, so the
    # driver's --steps 20 --warmup 5 run ends within a minute; --ref-sample bounds it for very long step counts
    n_sample = min(args.n, args.ref_sample) if args.ref_sample > 0 else args.n
    rate, t, threads = cpu_port_rate(n_sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "particle-steps/s", "value": rate, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C5 3-D pseudo-hard-sphere NVE N=%d phi=%.2f dt=%g cutoff=%g" % (n_sample, PHI, DT, CUTOFF) +
                               ("" if n_sample == args.n else " (bounded sample of the N=%d workload)" % args.n),
                   "n_particles": n_sample, "cutoff": CUTOFF},
        "cpu_baseline": {"value": rate, "unit": "particle-steps/s", "cores": threads, "kind": "port",
                         "sample": "N=%d particles x %d steps, cell list rebuilt every step at cell=cutoff=1.5, OpenMP" % (n_sample, args.steps)},
        "e2e": {"value": rate, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Julia reference not executable in this image; reference-shaped C/OpenMP port (oracle/md_oracle.c)",
    }
    print(json.dumps(line), flush=True)


def extra_configs(md, device):
    """BASELINE.json configs 1-4 on this GPU (C5 is the headline workload): device-timed particle-steps/s of the production
    step (graph replay / persistent small-system kernel), so that the driver's record carries every config, not just C5.
    Synthetic inputs of SURVEY 8d, melted before timing; each entry says what ran."""
    from mdjl_b200 import workloads
    out = []

    def case(name, dim, cfg, v0, tag, params, dt, kt, ens, steps, melt):
        n = cfg["x"].shape[0]
        e = md.Engine(dim, n, cfg["box"], CUTOFF, tag, params, seed=20261018, device=device)
        e.upload(cfg["x"], cfg["diam"], velocities=v0)
        if melt:
            mdt = dt if ens != "brownian" else 1e-3
            e.run_nvt(melt, mdt, kt, 100 * mdt, thermo=False)
        run = {"nve": lambda k: e.run_nve(k, dt, thermo=False), "nvt": lambda k: e.run_nvt(k, dt, kt, 100 * dt, thermo=False),
               "brownian": lambda k: e.run_brownian(k, dt, kt, thermo=False)}[ens]
        run(max(20, steps // 10))
        run(steps)
        st = e.stats()
        ms = st["last_run_ms"] / steps
        bytes_step = {("nve", 3): 176.0, ("nvt", 3): 176.0, ("brownian", 3): 80.0, ("nve", 2): 120.0}[(ens, dim)]
        hbm, _ = peaks()
        out.append({"config": name, "n_particles": n, "ensemble": ens, "steps": steps, "us_per_step": 1e3 * ms,
                    "particle_steps_per_s": n / ms * 1e3, "mode": {1: "cells", 2: "list", 3: "small-system persistent kernel"}[st["mode"]],
                    "roofline_step_frac": bytes_step * n / (ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes_per_particle_step": bytes_step})
        e.close()

    kt = workloads.KT_README
    c1 = workloads.phs_fluid(1024)
    case("C1 3-D PseudoHS N=1024 NVT (README)", 3, c1, workloads.velocities(1024, 3, kt), md._capi.POT_PSEUDOHS, (), DT, kt, "nvt", 5000, 2000)
    case("C1 3-D PseudoHS N=1024 NVE (README)", 3, c1, workloads.velocities(1024, 3, kt), md._capi.POT_PSEUDOHS, (), DT, kt, "nve", 5000, 2000)
    c2 = workloads.poly2d(1200)
    e = md.Engine(2, 1200, c2["box"], CUTOFF, md._capi.POT_POLY, (1.25, 0.2), seed=1, device=device)
    e.upload(c2["x"], c2["diam"])
    e.fire_minimize(max_steps=3000, tol=1e-3, dt_initial=1e-4, dt_max=5e-3)   # the lattice start of the mixture overlaps
    c2["x"] = e.download()[0]
    e.close()
    case("C2 2-D polydisperse N=1200 NVE", 2, c2, workloads.velocities(1200, 2, 0.11), md._capi.POT_POLY, (1.25, 0.2), 5e-3, 0.11, "nve", 5000, 2000)
    c3 = workloads.phs_fluid(1 << 20)
    v3 = workloads.velocities(1 << 20, 3, kt)
    case("C3 3-D PseudoHS N=2^20 NVT", 3, c3, v3, md._capi.POT_PSEUDOHS, (), DT, kt, "nvt", 500, 1500)
    case("C4 3-D PseudoHS N=2^20 Brownian", 3, c3, v3, md._capi.POT_PSEUDOHS, (), 1e-5, kt, "brownian", 500, 1500)
    return out


def slab_parity(md, dist, torch, rank, world, local_rank, uid, tr, n=65536, steps=(150, 150)):
    """The multi-process slab path against the single-domain engine on the same input, BEFORE anything is timed: per-step
    interacting-pair counts exact, thermo rows and final positions to rounding (what tests/mp_slab_worker.py asserts; the
    driver's 1-GPU test box cannot run that test, so the record travels on the bench line).  NVT then NVE, across list
    rebuilds and migrations; the ring uses its own communicator id derived from the main one."""
    from mdjl_b200 import slabs, workloads
    uid2 = slabs.broadcast_unique_id(dist, rank, md.unique_id, device=torch.device("cuda", local_rank))
    cfg = workloads.phs_fluid(n)
    v0 = workloads.velocities(n, 3, KT)
    ring = md.SlabRing.nccl(rank, world, uid2, 3, n, cfg["box"], CUTOFF, md._capi.POT_PSEUDOHS, seed=77, device=local_rank,
                            slab_transport=tr)
    ring.upload(cfg["x"], cfg["diam"], velocities=v0)
    t1 = ring.run_nvt(steps[0], DT, KT, 100 * DT)
    t2 = ring.run_nve(steps[1], DT)
    st = ring.lead.stats()
    ids, x, v, f, img = ring.download_local()
    parts = [None] * world
    dist.all_gather_object(parts, (ids, x, v))
    ring.close()
    rec = None
    if rank == 0:
        X, V = np.empty((n, 3)), np.empty((n, 3))
        seen = np.zeros(n, dtype=np.int64)
        for (i_, x_, v_) in parts:
            X[i_], V[i_] = x_, v_
            seen[i_] += 1
        single = md.Engine(3, n, cfg["box"], CUTOFF, md._capi.POT_PSEUDOHS, seed=77, device=local_rank, mode=md._capi.MODE_LIST)
        single.upload(cfg["x"], cfg["diam"], velocities=v0)
        s1 = single.run_nvt(steps[0], DT, KT, 100 * DT)
        s2 = single.run_nve(steps[1], DT)
        xs, vs, _, _ = single.download()
        single.close()
        rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
        rec = {"n_particles": n, "steps_nvt": steps[0], "steps_nve": steps[1], "ownership_is_partition": bool(np.all(seen == 1)),
               "pair_counts_equal_every_step": bool(np.array_equal(t1[:, 3], s1[:, 3]) and np.array_equal(t2[:, 3], s2[:, 3])),
               "thermo_max_rel_diff": max(rel(t1[:, :3], s1[:, :3]), rel(t2[:, :3], s2[:, :3])),
               "positions_max_abs_diff": float(np.max(np.abs(X - xs))), "velocities_max_abs_diff": float(np.max(np.abs(V - vs))),
               "rebuilds": int(st["rebuilds"]), "slab_transport": int(st["slab_transport"]), "slab_graph": int(st["slab_graph"]),
               "against": "single-domain engine (list mode) on rank 0, same input and RNG stream"}
        rec["ok"] = bool(rec["ownership_is_partition"] and rec["pair_counts_equal_every_step"] and rec["thermo_max_rel_diff"] < 1e-8 and
                         rec["positions_max_abs_diff"] < 1e-7)
    return rec


def main_slabs(args, rank, world, local_rank):
    """N > 1: the same N-particle workload cut into x-slabs, one process per GPU (strong scaling).  Ghost columns
    travel with ncclSend/ncclRecv every step, migration at neighbour rebuilds, ncclAllReduce for the rebuild consensus
    and the global thermo scalars.  Timing: barrier + synchronize on both sides, CUDA events on each rank's stream,
    MAX over ranks."""
    import torch
    import torch.distributed as dist
    import mdjl_b200 as md
    from mdjl_b200 import slabs
    dev = torch.device("cuda", local_rank)
    # clocks: NVML opened now, far from any timed region; rank 0 samples every GPU of the job
    sampler = ClockSampler(nvml_indices(range(world)), enabled=(rank == 0)).open()
    # NCCL prints its version banner on stdout when NCCL_DEBUG asks for it; stdout carries exactly one JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    uid = slabs.broadcast_unique_id(dist, rank, md.unique_id, device=dev)
    n = args.n
    cfg, v0 = make_workload(n)
    box = cfg["box"]
    tr = {"auto": 0, "nccl": 1, "peer": 2}[args.transport]
    parity = None if args.no_parity else slab_parity(md, dist, torch, rank, world, local_rank, uid, tr)
    ring = md.SlabRing.nccl(rank, world, uid, 3, n, box, CUTOFF, md._capi.POT_PSEUDOHS, seed=20261018, device=local_rank,
                            skin=args.skin, slab_transport=tr)
    ring.upload(cfg["x"], cfg["diam"], velocities=v0)
    ring.compute_forces()   # first collective: communicator warm-up (and its banner) happen here
    torch.cuda.synchronize()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)

    def run(k, thermo=False):
        if args.ensemble == "nve":
            return ring.run_nve(k, DT, thermo=thermo)
        if args.ensemble == "nvt":
            return ring.run_nvt(k, DT, KT, 100 * DT, thermo=thermo)
        return ring.run_brownian(k, 1e-5, KT, thermo=thermo)

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    if args.melt > 0:
        ring.run_nvt(args.melt, DT, KT, 100 * DT, thermo=False)
    run(args.warmup)
    sync()
    st0 = ring.lead.stats()
    sampler.start()
    t_thermo = run(args.steps, thermo=True)
    sync()
    st1 = ring.lead.stats()
    clocks = sampler.stop()
    t = torch.tensor([st1["last_run_ms"]], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = torch.tensor([st1["kernel_launches"] - st0["kernel_launches"]], dtype=torch.int64, device=dev)
    dist.all_reduce(launches)
    value = n * args.steps / (ms * 1e-3)
    hbm, peak_src = peaks()

    # where a slab step spends its time: a SEPARATE pass of the same number of steps after the timed region, launched eagerly with
    # CUDA events around the phases and one host sync per phase (MDB200_SLAB_PROF; the timed region above replays a graph and
    # cannot be instrumented).  Shares of the eager step, not of the timed one.
    phases = None
    if not args.no_profile:
        os.environ["MDB200_SLAB_PROF"] = "1"
        try:
            run(args.steps)
            torch.cuda.synchronize()
            st2 = ring.lead.stats()
            if st2["prof_steps"]:
                k = st2["prof_steps"]
                phases = {"head_pack_flags_wait_decision": st2["prof_kick_ms"] / k, "tail_forces_thermo": st2["prof_force_ms"] / k,
                          "rebuild_per_step": st2["prof_rebuild_ms"] / k, "eager_ms_per_step": st2["last_run_ms"] / args.steps, "steps": int(k),
                          "how": "separate eager pass after the timed region, CUDA events per phase, one host sync per phase (rank 0)"}
        except Exception as exc:
            print("slab phase pass failed on rank %d: %s" % (rank, exc), file=sys.stderr)
        finally:
            os.environ.pop("MDB200_SLAB_PROF", None)
        sync()

    e2e = None
    if not args.no_e2e:
        # steady-state round trip of a rank: hand back the rows it owns (mdb_upload_owned), step, read them again
        # (mdb_download_owned); the global arrays were only needed once, to plan the slabs.  Both directions move through
        # pinned host buffers that are allocated once and reused (capacity 1.25 x the owned count: migration changes it
        # by a thin layer per rebuild).
        cap = int(st1["n_owned"] * 1.25) + 4096
        keep = []

        def pinned(shape, dtype):
            t = torch.empty(shape, dtype={np.float64: torch.float64, np.int32: torch.int32}[dtype], pin_memory=True)
            keep.append(t)
            return t.numpy()

        def bufset():
            return {"ids": pinned((cap,), np.int32), "x": pinned((cap, 3), np.float64), "v": pinned((cap, 3), np.float64),
                    "f": pinned((cap, 3), np.float64), "img": pinned((cap, 3), np.int32), "diam": pinned((cap,), np.float64)}
        A, B = bufset(), bufset()

        def down(buf):
            k = ring.download_local_into(buf["ids"], buf["x"], buf["v"], buf["f"], buf["img"])
            return k

        parts = {}

        def round_trip(src, k_src, dst):
            ta = time.perf_counter()
            ring.upload_owned((src["ids"][:k_src], src["x"][:k_src], src["diam"][:k_src], src["v"][:k_src], src["f"][:k_src],
                               src["img"][:k_src]))
            tb = time.perf_counter()
            th = run(args.steps, thermo=True)
            tc = time.perf_counter()
            k = down(dst)
            parts.update(upload_s=tb - ta, run_s=tc - tb, download_s=time.perf_counter() - tc)
            return th, k

        ok, th, k_in, k_out, t_e2e = 1, None, 0, 0, 0.0
        trips = []
        try:
            k_in = down(A)
            A["diam"][:k_in] = cfg["diam"][A["ids"][:k_in]]
            th, k_mid = round_trip(A, k_in, B)           # untimed warm-up of exactly the timed call (buffers, first touch)
            B["diam"][:k_mid] = cfg["diam"][B["ids"][:k_mid]]
        except Exception as exc:   # the bench line must still be printed; the end-to-end figure is then absent
            print("slab e2e failed on rank %d: %s" % (rank, exc), file=sys.stderr)
            ok = 0
        # three timed round trips, each behind a barrier every rank reaches whatever happened before, MAX over ranks per
        # trip, MEDIAN over the trips: one trip is a single shot of ~50 ms through shared host memory, and one late rank
        # (a page fault, a descheduled process) shows as a 4x outlier in one trip out of a few
        src, k_src, dst = B, (k_mid if ok else 0), A
        for _ in range(3):
            sync()
            if not ok:
                continue
            try:
                t0 = time.perf_counter()
                th, k_dst = round_trip(src, k_src, dst)
                torch.cuda.synchronize()
                trips.append(time.perf_counter() - t0)
                if os.environ.get("MDB200_BENCH_E2E_TRACE"):
                    print("e2e trip rank %d: %.4f s %s" % (rank, trips[-1], {k: round(v, 4) for k, v in parts.items()}),
                          file=sys.stderr)
                k_in, k_out = k_src, k_dst
                dst["diam"][:k_dst] = cfg["diam"][dst["ids"][:k_dst]]
                src, k_src, dst = dst, k_dst, src
            except Exception as exc:
                print("slab e2e failed on rank %d: %s" % (rank, exc), file=sys.stderr)
                ok = 0
        # collectives only out here, where every rank arrives whatever happened above
        sync()
        tt = torch.tensor((trips + [0.0, 0.0, 0.0])[:3] + [-float(ok)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        trips = [float(v) for v in tt[:3].tolist()]
        t_e2e = sorted(trips)[1]
        if tt[3].item() < -0.5:
            e2e = {"value": n * args.steps / t_e2e, "unit": "particle-steps/s",
                   "h2d_bytes_per_step": k_in * (3 * 24 + 8 + 12 + 4) / args.steps,
                   "d2h_bytes_per_step": (k_out * (3 * 24 + 12 + 4) + th.nbytes) / args.steps, "seconds": t_e2e,
                   "breakdown_s_rank0": parts, "trips_s": trips,
                   "what": "per rank (max over ranks): mdb_upload_owned(own rows, pinned host) + %d steps + mdb_download_owned "
                           "into pinned host buffers; one untimed warm-up round trip, then the median of 3 timed ones" % args.steps}

    nf = 3 * (n - 1.0)
    E = t_thermo[:, 0] + t_thermo[:, 2]
    step_gbs = BYTES_STEP_3D * n * args.steps / (ms * 1e-3) / 1e9
    line = {
        "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C5 3-D pseudo-hard-sphere %s N=%d phi=%.2f dt=%g cutoff=%g" % (args.ensemble.upper(), n, PHI, DT, CUTOFF),
                   "n_particles": n, "mode": "list", "skin": args.skin or "default",
                   "l2": "per-rank state (%.2f GiB) exceeds L2" % (n * 96 / 2 ** 30 / world),
                   "melt_steps": args.melt,
                   "parallelism": "x-slabs x%d, %s" % (world, {3: "peer-memory mailboxes over NVLink (own kernels write ghosts/migrants/"
                                                                   "reductions into cudaIpc-mapped peer memory), step replayed as one CUDA graph"
                                                                   if st1["slab_graph"] == 1 else "peer-memory mailboxes over NVLink, eager launches",
                                                               2: "NCCL send/recv ghosts + allreduce, eager launches"}.get(st1["slab_transport"], "?"))},
        "slab_transport": {3: "peer", 2: "nccl"}.get(st1["slab_transport"]), "slab_graph": bool(st1["slab_graph"] == 1),
        "slab_parity": parity,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches.item()),
        "roofline": None,
        "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": hbm * world, "unit": "GB/s", "frac": step_gbs / (hbm * world),
                          "algorithmic_bytes_per_particle_step": BYTES_STEP_3D},
        "cpu_baseline": None,
        "owned_per_rank": int(st1["n_owned"]),
        "slab_phases_ms": phases if phases is not None else (
            {"head_pack_flags_wait_decision": st1["prof_kick_ms"] / max(st1["prof_steps"], 1),
             "tail_forces_thermo": st1["prof_force_ms"] / max(st1["prof_steps"], 1),
             "rebuild_per_step": st1["prof_rebuild_ms"] / max(st1["prof_steps"], 1)} if st1["prof_steps"] else None),
        "rebuilds_in_timed_region": int(st1["rebuilds"] - st0["rebuilds"]),
        "physics": {"T_mean": float(np.mean(2 * t_thermo[:, 2] / nf)), "U_per_particle": float(np.mean(t_thermo[:, 0]) / n),
                    "E_drift_rel": float((E.max() - E.min()) / abs(E[0])) if args.ensemble == "nve" else None,
                    "pairs_last": int(t_thermo[-1, 3])},
    }
    ring.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--particles", dest="n", type=int, default=1 << 24, help="under torchrun use --particles (--n is ambiguous to its parser)")
    ap.add_argument("--melt", type=int, default=1500)
    ap.add_argument("--mode", default="auto", choices=["auto", "cells", "list"])
    ap.add_argument("--skin", type=float, default=0.0)
    ap.add_argument("--ensemble", default="nve", choices=["nve", "nvt", "brownian"])
    ap.add_argument("--eager", action="store_true", help="no CUDA graph (ncu cannot profile kernel nodes of graphs with conditional nodes)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1-C4 lines (extra_configs)")
    ap.add_argument("--cpu-sample", type=int, default=1 << 20)
    ap.add_argument("--transport", default="auto", choices=["auto", "nccl", "peer"],
                    help="multi-GPU: auto = peer-memory mailboxes over NVLink (NCCL send/recv if cudaIpc is unavailable)")
    ap.add_argument("--no-parity", action="store_true", help="multi-GPU: skip the slab-vs-single-domain parity record")
    ap.add_argument("--ref-sample", type=int, default=0, help="--impl reference: cap on the particle count (0 = the full --n workload)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import mdjl_b200 as md
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        return main_slabs(args, rank, world, local_rank)
    sampler = ClockSampler(nvml_indices([local_rank])).open()

    n = args.n
    cfg, v0 = make_workload(n)
    box = cfg["box"]
    modes = {"auto": md._capi.MODE_AUTO, "cells": md._capi.MODE_CELLS, "list": md._capi.MODE_LIST}
    eng = md.Engine(3, n, box, CUTOFF, md._capi.POT_PSEUDOHS, seed=20261018, device=local_rank, mode=modes[args.mode],
                    skin=args.skin, use_graph=not args.eager)
    eng.upload(cfg["x"], cfg["diam"], velocities=v0)
    del v0

    def run(e, k, thermo=False):
        if args.ensemble == "nve":
            return e.run_nve(k, DT, thermo=thermo)
        if args.ensemble == "nvt":
            return e.run_nvt(k, DT, KT, 100 * DT, thermo=thermo)
        return e.run_brownian(k, 1e-5, KT, thermo=thermo)

    # melt the lattice (untimed) so the timed fluid has interacting pairs, then warm up
    if args.melt > 0:
        eng.run_nvt(args.melt, DT, KT, 100 * DT, thermo=False)
    run(eng, args.warmup)
    torch.cuda.synchronize()
    st0 = eng.stats()
    sampler.start()
    t_thermo = run(eng, args.steps, thermo=True)
    st1 = eng.stats()
    clocks = sampler.stop()
    ms = st1["last_run_ms"]           # CUDA events on the engine's stream around exactly K steps
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    rebuilds = st1["rebuilds"] - st0["rebuilds"]
    value = n * args.steps / (ms * 1e-3)
    hbm, peak_src = peaks()

    # per-kernel durations: same steps in eager (profiling) mode on a second handle sharing the device
    roofline = None
    prof = {}
    x_now, v_now, f_now, img_now = eng.download()
    if not args.no_profile:
        e2 = md.Engine(3, n, box, CUTOFF, md._capi.POT_PSEUDOHS, seed=20261018, device=local_rank, mode=modes[args.mode],
                       skin=args.skin, use_graph=False)
        e2.upload(x_now, cfg["diam"], velocities=v_now, forces=f_now, images=img_now)
        ksteps = min(args.steps, 100)
        run(e2, 5)
        run(e2, ksteps)
        s2 = e2.stats()
        e2.close()
        kick = s2["prof_kick_ms"] / ksteps
        force = s2["prof_force_ms"] / ksteps
        rebuild_per_step = s2["prof_rebuild_ms"] / ksteps
        prof = {"kick_drift_ms": kick, "pair_force_ms": force, "rebuild_ms_per_step": rebuild_per_step,
                "eager_steps": ksteps}
        fused = st1["mode"] == 2 and args.ensemble == "nve" and not os.environ.get("MDB200_NO_FUSE")
        if force >= kick and fused:
            name, dur, bts = "k_force_list<KICK2=2>(pair forces + second kick + next step's kick-drift-wrap)", force, BYTES_FORCE_FUSED_3D
        elif force >= kick:
            name, dur, bts = "k_force_list(pair forces + second kick)" if st1["mode"] == 2 else "k_force_cells", force, BYTES_FORCE_KERNEL_3D
        else:
            name, dur, bts = "k_kick_drift", kick, BYTES_KICK_KERNEL_3D
        achieved = bts * n / (dur * 1e-3) / 1e9
        kinfo = eng.force_kernel_info()
        traffic, traffic_note = measured_traffic("k_kick_drift" if name == "k_kick_drift" else ("k_force_list_fused" if bts == BYTES_FORCE_FUSED_3D else "k_force_list"), n,
                                                 build=None if name == "k_kick_drift" else kinfo)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                    "traffic_source": "ncu --set full capture, dram read+write per launch: %s" % traffic_note, "kernel_build": kinfo,
                    "kernel": name, "kernel_ms": dur, "algorithmic_bytes_per_particle": bts, "peak_source": peak_src}
    step_gbs = BYTES_STEP_3D * n * args.steps / (ms * 1e-3) / 1e9
    # FP64 pipe: DFMA peak measured live on this device; the dominant kernel's pipe utilisation comes from the committed
    # ncu --set full capture (sm__pipe_fp64_cycles_active), like roofline.traffic
    fp64 = None
    try:
        with open(traffic_file()) as fh:
            cap = json.load(fh).get("k_force_list_fused", {})
        stale = cap.get("build") is not None and any(eng.force_kernel_info().get(k) != v for k, v in cap["build"].items())
        fp64 = {"peak_tflops_measured": eng.measure_fp64_peak(),
                "dominant_kernel_pipe_active_pct_ncu": None if stale else cap.get("fp64_pipe_active_pct"),
                "source": "mdb_measure_fp64_peak (DFMA chains, best of 5) / %s%s" % (os.path.basename(traffic_file()), " (stale capture: refused)" if stale else "")}
    except Exception:
        pass

    # end-to-end through the public host-buffer API: upload (pinned host) + K steps + download + thermo
    e2e = None
    if not args.no_e2e:
        def pinned(a):
            t = torch.empty(a.shape, dtype=torch.float64 if a.dtype == np.float64 else torch.int32, pin_memory=True)
            arr = t.numpy()
            arr[...] = a
            return t, arr
        keep = []
        hx, hv, hf, hi, hd = [pinned(a) for a in (x_now, v_now, f_now, img_now, cfg["diam"])]
        keep = [hx, hv, hf, hi, hd]
        ox, ov, of_, oi = [pinned(np.empty_like(a)) for a in (x_now, v_now, f_now, img_now)]
        keep += [ox, ov, of_, oi]
        e3 = md.Engine(3, n, box, CUTOFF, md._capi.POT_PSEUDOHS, seed=20261018, device=local_rank, mode=modes[args.mode],
                       skin=args.skin, use_graph=True)

        parts = {}

        def e2e_call():
            ta = time.perf_counter()
            e3.upload(hx[1], hd[1], velocities=hv[1], forces=hf[1], images=hi[1])
            tb = time.perf_counter()
            th = run(e3, args.steps, thermo=True)
            tc = time.perf_counter()
            e3.download_into(ox[1], ov[1], of_[1], oi[1])
            parts.update(upload_s=tb - ta, run_s=tc - tb, download_s=time.perf_counter() - tc)
            return th
        e2e_call()  # warm-up: allocations, graph capture
        trips = []
        for _ in range(3):   # median of three round trips (a single shot through host memory is noisy)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            th = e2e_call()
            trips.append(time.perf_counter() - t0)
        t_e2e = sorted(trips)[1]
        h2d = sum(a[1].nbytes for a in (hx, hv, hf, hi, hd))
        d2h = sum(a[1].nbytes for a in (ox, ov, of_, oi)) + th.nbytes
        e2e = {"value": n * args.steps / t_e2e, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d / args.steps,
               "d2h_bytes_per_step": d2h / args.steps, "seconds": t_e2e, "breakdown_s": parts, "trips_s": trips,
               "what": "mdb_upload(pinned host) + mdb_run_%s(%d steps) + mdb_download + thermo rows" % (args.ensemble, args.steps)}
        e3.close()

    cpu = None
    if not args.no_cpu and rank == 0:
        ns = min(args.cpu_sample, n)
        rate, t, threads = cpu_port_rate(ns, 10, 2)
        cpu = {"value": rate, "unit": "particle-steps/s", "cores": threads, "kind": "port",
               "sample": "N=%d particles x 10 steps of the same fluid (after 22 untimed), reference-shaped OpenMP port, %.1f s" % (ns, t)}

    extra = None
    if not args.no_extra:
        try:
            extra = extra_configs(md, local_rank)
        except Exception as exc:   # never lose the headline line over a side measurement
            extra = [{"error": str(exc)}]
    nf = 3 * (n - 1.0)
    E = t_thermo[:, 0] + t_thermo[:, 2]
    line = {
        "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C5 3-D pseudo-hard-sphere %s N=%d phi=%.2f dt=%g cutoff=%g" % (args.ensemble.upper(), n, PHI, DT, CUTOFF),
                   "n_particles": n, "mode": {1: "cells", 2: "list"}[st1["mode"]], "skin": args.skin or "default",
                   "l2": "state arrays (%.2f GiB) exceed L2; no flush needed" % (n * 96 / 2 ** 30),
                   "melt_steps": args.melt, "parallelism": "1 GPU" if world == 1 else "x-slabs x%d" % world},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": hbm, "unit": "GB/s", "frac": step_gbs / hbm,
                          "algorithmic_bytes_per_particle_step": BYTES_STEP_3D},
        "cpu_baseline": cpu,
        "fp64": fp64,
        "kernels": prof,
        "extra_configs": extra,
        "rebuilds_in_timed_region": int(rebuilds),
        "physics": {"T_mean": float(np.mean(2 * t_thermo[:, 2] / nf)), "U_per_particle": float(np.mean(t_thermo[:, 0]) / n),
                    "E_drift_rel": float((E.max() - E.min()) / abs(E[0])) if args.ensemble == "nve" else None,
                    "pairs_last": int(t_thermo[-1, 3])},
    }
    eng.close()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
