#!/usr/bin/env python
"""Benchmark of the thermal-fin batched forward map (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo's CUDA path)
    python bench.py --impl reference --gpus N ...             # reference arm: CPU port on the host cores

One "step" = one pass of the hot path over one batch of synthetic conductivity samples per GPU:
  * FOM leg (headline `value`): BASELINE config "five-param batched FOM solves ... 10^5 samples on the reference
    mesh, QoI = subfin averages": batched sparse-direct (frontal Cholesky) solve fused with B_obs; the Jacobi-PCG kernel
    of round 1 is timed next to it (`fom_pcg`)
  * ROM leg (`rom`): BASELINE config "nine-param batched ROM solves, 10^6 samples" (n_r = 81)
  * `fom_unstructured`: the same FOM workload on a reference-like unstructured mesh (~1446 dofs, 6-7 nnz per row)
  * `fom_nodal`: nodal Gaussian-field conductivities; `fom_refined`: the 99 945-dof mesh (direct solver and streaming PCG)
Samples are sharded over ranks (weak scaling: the per-GPU batch is fixed); with N > 1 the step ends with the NCCL
all-gather of the observables.  Outside the timed regions the first outputs of every leg are compared with the CPU
oracle (`parity`), and at N > 1 every rank re-solves rows of its neighbour's shard and checks the gathered array
bit for bit (`gather_order_ok`).  Prints ONE JSON line on rank 0; exits non-zero if a parity bound is exceeded.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "thermal-fin forward solves/sec (FOM & ROM) at 1/2/4/8 B200; % HBM roofline"
RESOLUTION = 40                 # reference's get_space(40); structured m = 3 -> n = 1597
FOM_SEED, ROM_SEED, REF_SEED, NODAL_SEED = 1, 0, 2, 3
REFINED_M = 26                  # n = 99 945
TOL = 1e-12
PARITY_BOUND = {"fom": 1e-10, "fom_pcg": 1e-10, "fom_unstructured": 1e-10, "fom_nodal": 1e-10, "fom_refined": 1e-10,
                "fom_refined_pcg": 1e-10, "rom": 1e-9}
FOM_WORKLOAD = ("config[2] five-param batched FOM on the m=3 mesh (n=1597): k ~ U(0.1,1)^5 (seed 1) -> nine sub-fin "
                "conductivities, QoI = 9 sub-fin averages")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fom-batch", type=int, default=100_000, help="FOM samples per GPU per step")
    ap.add_argument("--rom-batch", type=int, default=1_000_000, help="ROM samples per GPU per step")
    ap.add_argument("--refined-batch", type=int, default=2368,
                    help="refined-mesh (m=26, n=99 945) FOM samples per GPU per step; 0 disables the leg")
    ap.add_argument("--refined-pcg-batch", type=int, default=592, help="samples of the streaming-PCG comparison on the refined mesh")
    ap.add_argument("--refined-steps", type=int, default=2)
    ap.add_argument("--refined-strong-total", type=int, default=2368,
                    help="strong-scaling line of config 4: this many samples IN TOTAL, split over the ranks")
    ap.add_argument("--nodal-batch", type=int, default=100_000,
                    help="nodal Gaussian-field FOM samples per GPU per step; 0 disables the leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--cpu-fom-sample", type=int, default=65536, help="CPU baseline FOM sample (~10 s on 16 cores)")
    ap.add_argument("--cpu-rom-sample", type=int, default=65536)
    ap.add_argument("--grad-batch", type=int, default=100_000,
                    help="samples per GPU per step of the gradient legs (capped by --nodal-batch: the same fields; 0 = skip)")
    ap.add_argument("--ref-step-samples", type=int, default=8192, help="--impl reference: CPU solves per step")
    return ap.parse_args()


def fom_inputs(n, seed):
    """config 3: k ~ U(0.1, 1.0)^5 -> nine-vector [k1..k5,k4..k1] (generate_reduced_basis_five_param.py:34,57)."""
    k5 = np.random.default_rng(seed).uniform(0.1, 1.0, (n, 5))
    return np.ascontiguousarray(np.concatenate([k5, k5[:, 3::-1]], axis=1))


def rom_inputs(n, seed):
    """config 2: theta ~ U(0.1, 3.5)^9 (generate_reduced_basis_nine_param.py:299)."""
    return np.random.default_rng(seed).uniform(0.1, 3.5, (n, 9))


def refined_inputs(n, seed):
    """config 4: theta ~ U(0.1, 10)^9 (bounds of rom/error_optimization.py:96)."""
    return np.random.default_rng(seed).uniform(0.1, 10.0, (n, 9))


def bench_config(args, world):
    """The workload description both arms print (identical keys and values for the same flags)."""
    return {
        "workload": FOM_WORKLOAD,
        "fom_samples_per_gpu_per_step": args.fom_batch, "rom_samples_per_gpu_per_step": args.rom_batch,
        "mesh": "structured conforming fin mesh m=3, n=1597 dofs (the reference's mshr mesh is not shipped)",
        "n_r": 81, "parallelism": f"samples sharded x{world}, NCCL all-gather of observables" if world > 1 else "1 GPU",
    }


# ------------------------------------------------------------------------------------------------
# CPU legs (the ONLY place bench.py touches oracle/): reported baseline, the --impl reference arm, and the parity check
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(coords, cells, phi):
    from threadpoolctl import threadpool_limits
    threadpool_limits(1)
    from oracle.thermal_fin_oracle import FinOracle
    _W["o"] = FinOracle(coords, cells)
    _W["phi"] = phi


def _cpu_fom(theta_rows):
    o = _W["o"]
    return np.stack([o.qoi_operator(o.forward_nine_param(t)) for t in theta_rows])


def _cpu_rom(theta_rows):
    o, phi = _W["o"], _W["phi"]
    return np.stack([o.qoi_reduced(o.forward_nine_param_reduced(t, phi), phi) for t in theta_rows])


class CpuPool:
    """All host cores, one single-threaded oracle per worker process (fork before any CUDA work)."""

    def __init__(self, phi=None):
        import multiprocessing as mp
        from bayesianinferencedl_b200 import get_space
        V = get_space(RESOLUTION)
        self.cores = os.cpu_count() or 1
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_init,
                             initargs=(V.mesh().coordinates(), V.mesh().cells(), phi))
        self.pool.map(_cpu_fom, [fom_inputs(1, 99)] * self.cores)       # warm every worker

    def run(self, fn, rows):
        chunks = [c for c in np.array_split(rows, self.cores * 4) if len(c)]
        t0 = time.perf_counter()
        out = self.pool.map(fn, chunks)
        dt = time.perf_counter() - t0
        return np.concatenate(out), dt

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_pod_basis():
    """Reference-style POD basis built with the oracle (reference arm must not touch the CUDA path)."""
    from bayesianinferencedl_b200 import get_space
    from oracle.thermal_fin_oracle import FinOracle, pod_basis
    V = get_space(RESOLUTION)
    return pod_basis(FinOracle(V.mesh().coordinates(), V.mesh().cells()), 200, 81, seed=0)


def run_reference(args):
    """Reference arm: the CPU port of the reference path on all host cores, each step a stated prefix of the SAME
    workload the B200 arm runs (rank 0's fom_inputs / rom_inputs)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    phi = cpu_pod_basis()
    pool = CpuPool(phi)
    per_step = min(max(pool.cores * 16, args.ref_step_samples), args.fom_batch)
    th = fom_inputs(args.fom_batch, FOM_SEED)[:per_step]
    thr = rom_inputs(args.rom_batch, ROM_SEED)[:per_step]
    t_f = t_r = 0.0
    for s in range(args.steps + args.warmup):
        _, dt = pool.run(_cpu_fom, th)
        _, dtr = pool.run(_cpu_rom, thr)
        if s >= args.warmup:
            t_f += dt
            t_r += dtr
    pool.close()
    v = per_step * args.steps / t_f
    vr = per_step * args.steps / t_r
    sample = (f"every step = the first {per_step} samples of the {args.fom_batch}-sample FOM workload (assemble + scipy "
              f"splu + B_obs per sample), {pool.cores} single-threaded processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_f / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, world),
        "cpu_baseline": {"value": v, "unit": "solves/s", "cores": pool.cores, "kind": "port", "sample": sample,
                         "note": "CPU port of the reference path (oracle/): FEniCS/PETSc are not installable here, "
                                 "parity unpinned by the reference (DESIGN.md section 0)"},
        "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "rom": {"value": vr, "unit": "solves/s", "ms_per_step": 1e3 * t_r / args.steps,
                "sample": f"first {per_step} samples of the ROM workload per step, literal LSPG (A phi, psi^T psi, np.linalg.solve)"},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_ev = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_ev.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_ev.wait(0.1)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # ---- CPU baseline first (fork-based pool, before this process creates a CUDA context); rank 0, N=1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        phi_cpu = cpu_pod_basis()
        pool = CpuPool(phi_cpu)
        _, dt_f = pool.run(_cpu_fom, fom_inputs(args.fom_batch, FOM_SEED)[:args.cpu_fom_sample])
        _, dt_r = pool.run(_cpu_rom, rom_inputs(args.rom_batch, ROM_SEED)[:args.cpu_rom_sample])
        pool.close()
        nf, nr_ = min(args.cpu_fom_sample, args.fom_batch), min(args.cpu_rom_sample, args.rom_batch)
        cpu = {"value": nf / dt_f, "unit": "solves/s", "cores": pool.cores, "kind": "port",
               "sample": f"first {nf} samples of the FOM workload: oracle port of the reference path (numpy assembly + "
                         f"scipy splu + B_obs), {pool.cores} single-threaded processes, {dt_f:.1f} s",
               "reference_published": "the reference publishes no benchmark; its stored notebook outputs show ~7.4 FOM "
                                      "solves/s (MUQ chain, 1 CPU process) and 13.6-14.4 it/s (PyMC3 Metropolis), BASELINE.md",
               "rom_value": nr_ / dt_r,
               "rom_sample": f"first {nr_} ROM samples, literal averaged_affine_ROM.py:292-304 "
                             f"(A phi, psi^T psi, np.linalg.solve), {dt_r:.1f} s"}

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from bayesianinferencedl_b200 import AffineROMFin, _cabi, get_space
    from bayesianinferencedl_b200.dist import gather_rows
    from bayesianinferencedl_b200.rom.pod import generate_pod_basis

    V = get_space(RESOLUTION)
    phi = generate_pod_basis(V, 200, 81, seed=0, device=local_rank)
    model = AffineROMFin(V, None, phi, device=local_rank)
    h = model.handle
    n, n_obs, n_r = model.dofs, model.n_obs, model.n_r
    NF, NR = args.fom_batch, args.rom_batch

    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    f64 = torch.float64
    th_f_np = fom_inputs(NF, FOM_SEED + rank)
    th_r_np = rom_inputs(NR, ROM_SEED + rank)
    th_f_host = torch.from_numpy(th_f_np).pin_memory()
    th_r_host = torch.from_numpy(th_r_np).pin_memory()
    th_f, th_r = th_f_host.to(dev), th_r_host.to(dev)
    q_f = torch.empty((NF, n_obs), dtype=f64, device=dev)
    q_r = torch.empty((NR, n_obs), dtype=f64, device=dev)
    it_f = torch.empty(NF, dtype=torch.int32, device=dev)
    st_f = torch.empty(NF, dtype=torch.int32, device=dev)
    st_r = torch.empty(NR, dtype=torch.int32, device=dev)
    q_f_host = torch.empty((NF, n_obs), dtype=f64).pin_memory()
    q_r_host = torch.empty((NR, n_obs), dtype=f64).pin_memory()
    it_f_host = torch.empty(NF, dtype=torch.int32).pin_memory()
    st_f_host = torch.empty(NF, dtype=torch.int32).pin_memory()
    st_r_host = torch.empty(NR, dtype=torch.int32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, hh=None):
        """W warm-up + EXACTLY K steps bracketed by barrier + synchronize; CUDA events on the launching stream;
        L2 flushed between steps; returns max-over-ranks milliseconds and this rank's launch count."""
        for _ in range(warmup):
            fn()
            flush.fill_(1)
        barrier()
        hh = hh or h
        l0 = hh.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
            flush.fill_(1)
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=f64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), hh.kernel_launches() - l0

    def kernel_ms(fn, reps=3):
        """Duration of ONE launch of the dominant kernel: events tight around the call on its stream, L2 flushed before."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out = []
        for _ in range(reps):
            flush.fill_(1)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1))
        return float(np.mean(out))

    def affine_dev(hh, th, N, q, it=None, st=None, solver=0, maxit=20000):
        def run():
            hh.set_int("fom_solver", solver)
            hh.fom_affine_raw(th.data_ptr(), N, _cabi.IN_PARAMS, _cabi.MEM_DEVICE, TOL, maxit, qoi=q.data_ptr(),
                              iters=it.data_ptr() if it is not None else 0, status=st.data_ptr() if st is not None else 0,
                              stream=sp)
            return gather_rows(q[:N], N * world) if world > 1 else q
        return run

    def frontal_info(hh):
        return {"kernel": {1: "D1 frontal_lane_kernel (sample per thread)", 2: "D2 frontal_cta_kernel (sample per CTA, "
                           "observables through extra right-hand sides)", 3: "D2 frontal_cta_kernel (sample per CTA, "
                           "factor in HBM + backward substitution)"}.get(hh.get_int("frontal_kernel"), "?"),
                "threads": hh.get_int("frontal_threads"), "ctas_per_sm": hh.get_int("frontal_ctas_per_sm"),
                "smem_bytes": hh.get_int("frontal_smem_bytes"),
                "front_slots": hh.get_int("frontal_slots_lane" if hh.get_int("frontal_kernel") == 1 else "frontal_slots"),
                "substitution_ctas_per_sm": hh.get_int("frontal_bsub_ctas_per_sm"),
                "max_column": hh.get_int("frontal_cmax"), "factor_nnz": hh.get_int("frontal_nnz_factor"),
                "pair_updates": hh.get_int("frontal_pair_updates"), "samples_per_warp": hh.get_int("frontal_lanes")}

    def rom_dev():
        h.rom_raw(th_r.data_ptr(), NR, _cabi.IN_PARAMS, _cabi.MEM_DEVICE, qoi=q_r.data_ptr(),
                  status=st_r.data_ptr(), stream=sp)
        return gather_rows(q_r, NR * world) if world > 1 else q_r

    def fom_e2e():     # the C-ABI call the facade makes: HOST buffers in, HOST buffers out
        h.set_int("fom_solver", 0)
        h.fom_affine_raw(th_f_host.data_ptr(), NF, _cabi.IN_PARAMS, _cabi.MEM_HOST, TOL, 20000,
                         qoi=q_f_host.data_ptr(), iters=it_f_host.data_ptr(), status=st_f_host.data_ptr(), stream=sp)

    def rom_e2e():
        h.rom_raw(th_r_host.data_ptr(), NR, _cabi.IN_PARAMS, _cabi.MEM_HOST, qoi=q_r_host.data_ptr(),
                  status=st_r_host.data_ptr(), stream=sp)

    # flush cost (inside the bracket) measured once so it can be reported
    timed(lambda: None, 2, 1)
    flush_ms, _ = timed(lambda: None, 4, 1)
    flush_ms /= 4
    smem_peak = h.smem_bandwidth()          # GB/s, measured on this GPU: roofline denominator of the on-chip kernels

    sampler = ClockSampler(local_rank)
    sampler.start()
    K, Wm = args.steps, max(args.warmup, 3)
    fom_dev = affine_dev(h, th_f, NF, q_f, it_f, st_f, solver=0)
    ms_f, launches_f = timed(fom_dev, K, Wm)
    fom_solver_used = h.get_int("fom_solver")
    fom_geo = frontal_info(h) if fom_solver_used == 2 else None
    gathered_f = fom_dev()
    torch.cuda.synchronize()
    ok_f = bool((st_f == 0).all().item())
    parity_in = {"fom": (th_f_np[:64].copy(), q_f[:64].cpu().numpy())}

    # ---- gather order (N > 1): re-solve the first rows of the NEXT rank's shard here and compare with the gathered
    # array bit for bit (the kernels are deterministic per sample, whichever GPU / lane solves it)
    gather_ok = None
    if world > 1:
        nb = (rank + 1) % world
        th_nb = torch.from_numpy(fom_inputs(NF, FOM_SEED + nb)[:256]).to(dev)
        q_nb = torch.empty((256, n_obs), dtype=f64, device=dev)
        h.fom_affine_raw(th_nb.data_ptr(), 256, _cabi.IN_PARAMS, _cabi.MEM_DEVICE, TOL, 20000, qoi=q_nb.data_ptr(), stream=sp)
        torch.cuda.synchronize()
        mine = bool(torch.equal(gathered_f[nb * NF: nb * NF + 256], q_nb)) and \
            bool(torch.equal(gathered_f[rank * NF: (rank + 1) * NF], q_f))
        flag = torch.tensor([1 if mine else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_ok = bool(flag.item() == 1)

    # ---- the round-1 Jacobi-PCG kernel (K1) on the same workload, for comparison and for its own roofline
    q_p = torch.empty((NF, n_obs), dtype=f64, device=dev)
    pcg_dev = affine_dev(h, th_f, NF, q_p, it_f, st_f, solver=1)
    ms_p, launches_p = timed(pcg_dev, max(2, K // 2), 2)
    k_ms_p = kernel_ms(pcg_dev)
    iters_sum = int(it_f.to(torch.int64).sum().item())
    ok_p = bool((st_f == 0).all().item())
    pcg_geo = {"threads": h.get_int("pcg_threads"), "rows_per_thread": h.get_int("pcg_rows_per_thread"),
               "ctas_per_sm": h.get_int("pcg_ctas_per_sm"), "smem_bytes": h.get_int("pcg_smem_bytes"),
               "ell_width": h.get_int("ell_width")}
    parity_in["fom_pcg"] = (th_f_np[:64].copy(), q_p[:64].cpu().numpy())
    k_ms = kernel_ms(fom_dev)

    ms_r, launches_r = timed(rom_dev, K, Wm)
    ok_r = bool((st_r == 0).all().item())
    parity_in["rom"] = (th_r_np[:64].copy(), q_r[:64].cpu().numpy())
    ms_fe, launches_fe = timed(fom_e2e, K, Wm)
    ms_re, launches_re = timed(rom_e2e, K, Wm)

    # ---- facade-level end to end: the call a user of the reference API makes, PAGEABLE numpy in and out (wall clock)
    def facade_rate(fn, arr, reps=3):
        fn(arr[:1024])
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn(arr)
        dt = (time.perf_counter() - t0) / reps
        t = torch.tensor([dt], dtype=f64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * len(arr) / float(t.item()), out
    h.set_int("fom_solver", 0)
    fac_f, qf_fac = facade_rate(model.forward_nine_param_qoi, th_f_np)
    fac_r, _ = facade_rate(model.forward_reduced_qoi, th_r_np)
    facade_same = bool(np.array_equal(qf_fac[:256], q_f[:256].cpu().numpy()))

    # ---- the same FOM workload on a REFERENCE-LIKE mesh: unstructured, ~1446 dofs, 6-7 non-zeros per row (the
    # structured mesh prunes to 5), marker-0 cells along x = 2.5 / 3.5 (SURVEY Q-1)
    from bayesianinferencedl_b200.assembly import build_operators
    from bayesianinferencedl_b200.fom.thermal_fin import get_space_unstructured
    Vu = get_space_unstructured()
    opsu = build_operators(Vu)
    hu = _cabi.TfinHandle(local_rank)
    hu.set_operator(opsu.row_ptr, opsu.col_idx, opsu.affine_terms(), opsu.rhs)
    hu.set_observation(*opsu.obs_csr())
    q_u = torch.empty((NF, 9), dtype=f64, device=dev)
    it_u = torch.empty(NF, dtype=torch.int32, device=dev)
    st_u = torch.empty(NF, dtype=torch.int32, device=dev)
    un_dev = affine_dev(hu, th_f, NF, q_u, it_u, st_u, solver=0)
    ms_u, l_u = timed(un_dev, K, Wm, hu)
    un_geo = frontal_info(hu) if hu.get_int("fom_solver") == 2 else None
    parity_in["fom_unstructured"] = (th_f_np[:64].copy(), q_u[:64].cpu().numpy())
    ok_u = bool((st_u == 0).all().item())
    un_pcg = affine_dev(hu, th_f, NF, q_u, it_u, st_u, solver=1)
    ms_up, _ = timed(un_pcg, 2, 1, hu)
    it_u_mean = float(it_u.double().mean().item())
    unstructured = {
        "value": world * NF / ((ms_u / K - flush_ms) * 1e-3), "unit": "solves/s", "ms_per_step": ms_u / K - flush_ms,
        "workload": f"the headline FOM workload on a reference-like unstructured mesh (Delaunay stand-in for the mshr mesh): "
                    f"n={opsu.n} dofs, {opsu.nnz / opsu.n:.2f} non-zeros per row (max {int(np.diff(opsu.row_ptr).max())}), "
                    f"{int((opsu.cell_markers == 0).sum())} marker-0 cells",
        "solver": "direct" if un_geo else "pcg", "frontal": un_geo, "all_converged": ok_u, "gpu_launches": l_u,
        "pcg": {"value": world * NF / ((ms_up / 2 - flush_ms) * 1e-3), "unit": "solves/s", "mean_pcg_iters": it_u_mean,
                "ell_width": hu.get_int("ell_width"), "rows_per_thread": hu.get_int("pcg_rows_per_thread"),
                "threads": hu.get_int("pcg_threads"), "ctas_per_sm": hu.get_int("pcg_ctas_per_sm")},
    }
    hu.close()

    # ---- config[4] nodal Gaussian-field conductivity: k = exp(0.5 chol^T z), Matern-5/2, l = 1.6; per-sample FEM
    # assembly + solve (direct D1 by default, the K2 PCG kernel next to it)
    nodal = None
    if args.nodal_batch > 0:
        from bayesianinferencedl_b200 import Fin
        from bayesianinferencedl_b200.bayesian_inference.gaussian_field import FieldSampler
        NN = args.nodal_batch
        fin = Fin(V, device=local_rank)
        prior = FieldSampler(V, "m52", 1.6, handle=fin.handle)      # covariance + Cholesky on the device
        k_dev = torch.empty((NN, n), dtype=f64, device=dev)
        lib = fin.handle._lib

        def draw():  # k = exp(0.5 chol^T z), z from the device Philox generator; 16384-sample chunks
            for lo in range(0, NN, 16384):
                m = min(16384, NN - lo)
                rc = lib.tfin_field_sample(fin.handle._h, None, NODAL_SEED, 0, rank * NN + lo, m, _cabi.MEM_DEVICE,
                                           k_dev[lo:].data_ptr(), None, sp)
                if rc:
                    raise RuntimeError(lib.tfin_last_error().decode())
        draw()
        torch.cuda.synchronize()
        e0s, e1s = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0s.record(stream)
        draw()
        e1s.record(stream)
        torch.cuda.synchronize()
        sampler_ms = e0s.elapsed_time(e1s)
        k_host = torch.empty((NN, n), dtype=f64).pin_memory()
        k_host.copy_(k_dev)
        q_n = torch.empty((NN, n_obs), dtype=f64, device=dev)
        it_n = torch.empty(NN, dtype=torch.int32, device=dev)
        st_n = torch.empty(NN, dtype=torch.int32, device=dev)
        hn = fin.handle

        def nodal_dev(solver=0):
            def run():
                hn.set_int("fom_solver", solver)
                hn.fom_nodal_raw(k_dev.data_ptr(), NN, _cabi.MEM_DEVICE, TOL, 20000, qoi=q_n.data_ptr(),
                                 iters=it_n.data_ptr(), status=st_n.data_ptr(), stream=sp)
                return gather_rows(q_n, NN * world) if world > 1 else q_n
            return run

        q_n_host = torch.empty((NN, n_obs), dtype=f64).pin_memory()
        st_n_host = torch.empty(NN, dtype=torch.int32).pin_memory()

        def nodal_e2e():
            hn.set_int("fom_solver", 0)
            hn.fom_nodal_raw(k_host.data_ptr(), NN, _cabi.MEM_HOST, TOL, 20000, qoi=q_n_host.data_ptr(),
                             status=st_n_host.data_ptr(), stream=sp)

        def nodal_pipeline():   # fields drawn ON the device (tfin_field_sample), only the observables cross PCIe
            draw()
            hn.set_int("fom_solver", 0)
            hn.fom_nodal_raw(k_dev.data_ptr(), NN, _cabi.MEM_DEVICE, TOL, 20000, qoi=q_n.data_ptr(),
                             status=st_n.data_ptr(), stream=sp)
            q_n_host.copy_(q_n, non_blocking=True)

        ms_n, ln = timed(nodal_dev(0), K, Wm, hn)
        nodal_geo = frontal_info(hn) if hn.get_int("fom_solver") == 2 else None
        ok_n = bool((st_n == 0).all().item())
        parity_in["fom_nodal"] = (k_dev[:64].cpu().numpy(), q_n[:64].cpu().numpy())
        ms_ne, _ = timed(nodal_e2e, K, Wm, hn)
        ms_np, lnp = timed(nodal_pipeline, K, Wm, hn)
        ms_npcg, _ = timed(nodal_dev(1), 2, 1, hn)
        it_sum_n = int(it_n.to(torch.int64).sum().item())
        n_cells = fin.ops.n_cells
        step_n = ms_n / K - flush_ms
        nodal = {
            "value": world * NN / (step_n * 1e-3), "unit": "solves/s", "ms_per_step": step_n,
            "workload": f"config[4]: {NN} Matern-5/2 (l=1.6) nodal fields per GPU, per-sample FEM assembly (cell "
                        f"coefficients) + sparse-direct solve, mesh m=3",
            "solver": "direct" if nodal_geo else "pcg", "frontal": nodal_geo, "all_converged": ok_n,
            "e2e": {"value": world * NN / ((ms_ne / K - flush_ms) * 1e-3), "unit": "solves/s",
                    "h2d_bytes_per_step": NN * n * 8, "d2h_bytes_per_step": NN * (n_obs * 8 + 4),
                    "ms_per_step": ms_ne / K - flush_ms},
            "gpu_launches": ln,
            "e2e_device_prior": {"value": world * NN / ((ms_np / K - flush_ms) * 1e-3), "unit": "solves/s",
                                 "h2d_bytes_per_step": 0, "d2h_bytes_per_step": NN * n_obs * 8,
                                 "ms_per_step": ms_np / K - flush_ms, "gpu_launches": lnp,
                                 "what": "prior draw (Philox + triangular GEMM + exp) -> nodal FOM -> observables to "
                                         "pinned host memory; the generate_fin_dataset.py:83-100 loop without its "
                                         "per-sample H2D"},
            "field_sampler": {"value": NN / (sampler_ms * 1e-3), "unit": "fields/s/GPU", "ms": sampler_ms,
                              "what": "tfin_field_sample: Philox normals + fp64 triangular GEMM + exp, device resident",
                              "tflops_fp64": NN * float(n) * n / (sampler_ms * 1e-3) / 1e12},
            "pcg": {"value": world * NN / ((ms_npcg / 2 - flush_ms) * 1e-3), "unit": "solves/s",
                    "mean_pcg_iters": it_sum_n / NN, "kernel": "pcg_kernel<nodal> (K2)",
                    "smem_bytes_per_iteration": (2 * hn.get_int("ell_width_nodal") + 1) * 8 * n},
        }
        # ---- gradient legs: FOM adjoint gradient and reduced gradient, device resident
        if args.grad_batch > 0:
            NG = min(args.grad_batch, NN)
            data_dev = q_n[:1].clone()
            g_n = torch.empty((NG, n), dtype=f64, device=dev)
            c_n = torch.empty(NG, dtype=f64, device=dev)

            hn.set_int("fom_solver", 0)

            def fom_grad():
                rc = lib.tfin_fom_nodal_gradient(hn._h, k_dev.data_ptr(), NG, _cabi.MEM_DEVICE, TOL, 20000,
                                                 data_dev.data_ptr(), 1, g_n.data_ptr(), c_n.data_ptr(), None, None,
                                                 st_n.data_ptr(), sp)
                if rc:
                    raise RuntimeError(lib.tfin_last_error().decode())
            ms_g, lg = timed(fom_grad, K, Wm, hn)
            model.set_data(np.zeros(n_obs))
            model.grad_reduced_nine_param(np.ones((2, 9)))        # builds / uploads the Gram blocks once
            g_r = torch.empty((NG, n), dtype=f64, device=dev)

            def rom_grad():
                rc = lib.tfin_rom_gradient(h._h, k_dev.data_ptr(), NG, _cabi.IN_NODAL, _cabi.MEM_DEVICE,
                                           data_dev.data_ptr(), 1, _cabi.IN_NODAL, g_r.data_ptr(), c_n.data_ptr(),
                                           None, None, st_n.data_ptr(), sp)
                if rc:
                    raise RuntimeError(lib.tfin_last_error().decode())
            ms_gr, lgr = timed(rom_grad, K, Wm)
            nodal["gradients"] = {
                "fom": {"value": world * NG / ((ms_g / K - flush_ms) * 1e-3), "unit": "gradients/s",
                        "what": f"Fin.gradient batched ({NG} fields): direct solver = one factorisation, backward / adjoint-forward / "
                                f"backward substitution passes and the gradient-form kernel (PCG path: forward + adjoint PCG + "
                                f"gradient form in one kernel)", "solver": "direct" if hn.get_int("fom_solver") == 2 else "pcg",
                        "gpu_launches": lg, "all_converged": bool((st_n[:NG] == 0).all().item())},
                "rom": {"value": world * NG / ((ms_gr / K - flush_ms) * 1e-3), "unit": "gradients/s",
                        "what": f"AffineROMFin.grad_reduced batched (nodal in, nodal out), {NG} fields",
                        "gpu_launches": lgr}}
        del k_dev, k_host
        fin.handle.close()

    # ---- config[3] refined mesh (m=26, n=99 945), nine-param FOM: wide-front direct solver (D2) by default, the
    # HBM-streaming PCG kernel (K4) on a smaller batch next to it, and a STRONG-scaling line (fixed total batch)
    refined = None
    if args.refined_batch > 0:
        NRf = args.refined_batch
        Vr = get_space(RESOLUTION, m=REFINED_M)
        opsr = build_operators(Vr)
        hr = _cabi.TfinHandle(local_rank)
        hr.set_operator(opsr.row_ptr, opsr.col_idx, opsr.affine_terms(), opsr.rhs)
        hr.set_observation(*opsr.obs_csr())
        th_ref_np = refined_inputs(NRf, REF_SEED + rank)
        th_ref = torch.from_numpy(th_ref_np).to(dev)
        q_ref = torch.empty((NRf, 9), dtype=f64, device=dev)
        it_ref = torch.empty(NRf, dtype=torch.int32, device=dev)
        st_ref = torch.empty(NRf, dtype=torch.int32, device=dev)
        ref_dev = affine_dev(hr, th_ref, NRf, q_ref, it_ref, st_ref, solver=0, maxit=50000)
        ms_ref, l_ref = timed(ref_dev, args.refined_steps, 1, hr)
        ref_geo = frontal_info(hr) if hr.get_int("fom_solver") == 2 else None
        step_ref = ms_ref / args.refined_steps - flush_ms
        ok_ref = bool((st_ref == 0).all().item())
        parity_in["fom_refined"] = (th_ref_np[:4].copy(), q_ref[:4].cpu().numpy())
        # strong scaling: the SAME total batch whatever the number of ranks
        tot = args.refined_strong_total
        lo, hi = min(rank * ((tot + world - 1) // world), tot), min((rank + 1) * ((tot + world - 1) // world), tot)
        th_s = torch.from_numpy(refined_inputs(tot, REF_SEED)[lo:hi]).to(dev)
        q_s = torch.empty((max(hi - lo, 1), 9), dtype=f64, device=dev)
        strong_dev = affine_dev(hr, th_s, hi - lo, q_s, solver=0, maxit=50000) if hi > lo else (lambda: None)

        def strong_step():
            if hi > lo:
                hr.set_int("fom_solver", 0)
                hr.fom_affine_raw(th_s.data_ptr(), hi - lo, _cabi.IN_PARAMS, _cabi.MEM_DEVICE, TOL, 50000, qoi=q_s.data_ptr(), stream=sp)
            if world > 1:
                gather_rows(q_s[: hi - lo], tot)
        ms_s, _ = timed(strong_step, args.refined_steps, 1, hr)
        # streaming PCG (K4): the genuinely HBM-bound kernel of round 1
        NP = min(args.refined_pcg_batch, NRf)
        pcgr = None
        if NP > 0:
            ref_pcg = affine_dev(hr, th_ref, NP, q_ref, it_ref, st_ref, solver=1, maxit=50000)
            ms_rp, l_rp = timed(ref_pcg, 1, 1, hr)
            step_rp = ms_rp - flush_ms
            it_sum_ref = int(it_ref[:NP].to(torch.int64).sum().item())
            bytes_ref = 88.0 * opsr.n * it_sum_ref
            parity_in["fom_refined_pcg"] = (th_ref_np[:4].copy(), q_ref[:4].cpu().numpy())
            pcgr = {"value": world * NP / (step_rp * 1e-3), "unit": "solves/s", "ms_per_step": step_rp, "samples_per_gpu": NP,
                    "mean_pcg_iters": it_sum_ref / NP, "all_converged": bool((st_ref[:NP] == 0).all().item()),
                    "kernel": f"pcg_stream_kernel (K4, tile {hr.get_int('stream_tile')})", "gpu_launches": l_rp,
                    "roofline": {"bound": "hbm", "achieved": bytes_ref / (step_rp * 1e-3) / 1e9, "peak": None, "unit": "GB/s",
                                 "frac": None, "algorithmic_bytes_per_launch": bytes_ref,
                                 "note": "88*n bytes per iteration per sample (SURVEY 8d), each sample counted with its own "
                                         "iteration count; vectors stream from HBM every pass (working set >> L2)"}}
        pairs = ref_geo["pair_updates"] if ref_geo else 0
        refined = {
            "value": world * NRf / (step_ref * 1e-3), "unit": "solves/s", "ms_per_step": step_ref,
            "steps": args.refined_steps, "warmup": 1,
            "workload": f"config[3]: refined mesh m={REFINED_M}, n={opsr.n} dofs, nine-param theta~U(0.1,10), "
                        f"{NRf} samples per GPU per step",
            "solver": "direct" if ref_geo else "pcg", "frontal": ref_geo, "all_converged": ok_ref, "gpu_launches": l_ref,
            "roofline": ({"bound": "smem", "unit": "GB/s", "peak": smem_peak,
                          "achieved": 16.0 * pairs * NRf / (step_ref * 1e-3) / 1e9,
                          "frac": 16.0 * pairs * NRf / (step_ref * 1e-3) / 1e9 / smem_peak,
                          "algorithmic_bytes_per_solve": 16.0 * pairs,
                          "note": "the front lives in shared memory: every pair update reads and writes 8 bytes of it; no "
                                  "HBM traffic beyond theta / qoi in this mode"} if ref_geo else None),
            "strong_scaling": {"value": tot / ((ms_s / args.refined_steps - flush_ms) * 1e-3), "unit": "solves/s",
                               "samples_total": tot, "samples_this_rank": hi - lo, "n_gpus": world,
                               "ms_per_step": ms_s / args.refined_steps - flush_ms,
                               "note": "fixed total batch split over the ranks (one sample per CTA: the tail is "
                                       "ceil(samples / CTAs) waves)"},
            "pcg_stream": pcgr,
        }
        hr.close()
    clocks = sampler.stop()

    # DGEMM rate of this GPU (cuBLAS) as the FP64 denominator for the ROM leg
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a = torch.randn(4096, 4096, dtype=f64, device=dev)
    b = torch.randn(4096, 4096, dtype=f64, device=dev)
    for _ in range(2):
        a @ b
    e0.record(stream)
    for _ in range(3):
        a @ b
    e1.record(stream)
    torch.cuda.synchronize()
    dgemm_tflops = 3 * 2 * 4096 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- parity: first outputs of every leg against the CPU oracle (checker only; outside every timed region)
    parity, parity_ok = None, True
    if not args.no_parity:
        from oracle.thermal_fin_oracle import FinOracle
        parity = {}
        orc = FinOracle(V.mesh().coordinates(), V.mesh().cells())
        for leg in ("fom", "fom_pcg"):
            th, q = parity_in[leg]
            parity[leg] = max(relerr(q[s], orc.qoi_operator(orc.forward_nine_param(th[s]))) for s in range(len(th)))
        th, q = parity_in["rom"]
        parity["rom"] = max(relerr(q[s], orc.qoi_reduced(orc.forward_nine_param_reduced(th[s], phi), phi)) for s in range(len(th)))
        if "fom_nodal" in parity_in:
            kk, q = parity_in["fom_nodal"]
            parity["fom_nodal"] = max(relerr(q[s], orc.qoi_operator(orc.forward(kk[s]))) for s in range(len(kk)))
        orcu = FinOracle(opsu.coords, opsu.cells)
        th, q = parity_in["fom_unstructured"]
        parity["fom_unstructured"] = max(relerr(q[s], orcu.qoi_operator(orcu.forward_nine_param(th[s]))) for s in range(len(th)))
        if "fom_refined" in parity_in:
            orcr = FinOracle(opsr.coords, opsr.cells)
            for leg in ("fom_refined", "fom_refined_pcg"):
                if leg in parity_in:
                    th, q = parity_in[leg]
                    parity[leg] = max(relerr(q[s], orcr.qoi_operator(orcr.forward_nine_param(th[s]))) for s in range(len(th)))
        parity_ok = all(v <= PARITY_BOUND[k] for k, v in parity.items())
        parity["bounds"] = {k: PARITY_BOUND[k] for k in parity if k != "bounds"}
        parity["what"] = ("max relative error of the observables of the first 64 samples of each leg (4 on the refined mesh) "
                          "against the CPU oracle (sparse LU); the oracle itself is a restatement, DESIGN.md section 0")
        parity["ok"] = parity_ok

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    ncu_t = {}
    try:
        ncu_t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass

    if refined is not None and refined["pcg_stream"] is not None:
        rr = refined["pcg_stream"]["roofline"]
        rr["peak"], rr["frac"], rr["peak_source"] = hbm_peak, rr["achieved"] / hbm_peak, peak_src
        ts = ncu_t.get("pcg_stream_kernel")
        rr["traffic"] = ts["ratio_to_algorithmic"] * rr["algorithmic_bytes_per_launch"] if ts else None
        rr["traffic_note"] = (f"ncu capture ({ts['launch']}): dram bytes / algorithmic bytes = "
                              f"{ts['ratio_to_algorithmic']:.4f}, applied to this launch" if ts else None)
    step_ms_f = (ms_f / K) - flush_ms
    step_ms_r = (ms_r / K) - flush_ms
    fom_rate = world * NF / (step_ms_f * 1e-3)
    rom_rate = world * NR / (step_ms_r * 1e-3)
    fom_e2e_rate = world * NF / ((ms_fe / K - flush_ms) * 1e-3)
    rom_e2e_rate = world * NR / ((ms_re / K - flush_ms) * 1e-3)
    rom_flops = 2 * 55 * 3321 + 2 * 10 * 81 + 81 ** 3 / 3 + 2 * 81 ** 2 + 2 * 9 * 81     # SURVEY 8d, ~559 kflop

    # ---- roofline of the dominant kernel of the headline leg
    if fom_geo:
        # D1: the front, the right-hand side and the factor ring live in shared memory.  Bytes a solve MUST move there:
        # 16 per pair update (read + write 8), and per factor entry 16 (gather + zero) + 16 (rhs update) + 16 (backward:
        # factor row + solution)
        smem_bytes = 16.0 * fom_geo["pair_updates"] + 48.0 * fom_geo["factor_nnz"]
        hbm_bytes = 16.0 * (fom_geo["factor_nnz"] + 2 * n) + 9 * 8 + n_obs * 8 + 8     # factor out and back, theta, qoi
        t1 = ncu_t.get("frontal_lane_kernel")
        roofline = {
            "bound": "smem", "achieved": smem_bytes * NF / (k_ms * 1e-3) / 1e9, "peak": smem_peak, "unit": "GB/s",
            "frac": smem_bytes * NF / (k_ms * 1e-3) / 1e9 / smem_peak,
            "peak_source": "tfin_smem_bandwidth(): conflict-free 8-byte shared-memory loads (the solver kernels' access width) on all SMs, measured in this run",
            "kernel": "frontal_lane_kernel (D1)", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": smem_bytes * NF,
            "traffic": (t1["dram_bytes"] / t1["launch_samples"] * NF) if t1 else None,
            "traffic_note": (f"ncu --set full capture of a {t1['launch_samples']}-sample launch ({t1['dram_bytes']:.4g} B of DRAM "
                             f"traffic = the factor written and read back once), scaled to this batch" if t1 else None),
            "hbm": {"algorithmic_bytes_per_launch": hbm_bytes * NF, "achieved": hbm_bytes * NF / (k_ms * 1e-3) / 1e9,
                    "peak": hbm_peak, "frac": hbm_bytes * NF / (k_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src},
            "note": "sparse-direct solve: the active front of every sample lives in shared memory, the factor makes one "
                    "round trip through HBM for the backward substitution.  Shared memory bounds the samples in flight "
                    "(three warps of 27 samples per SM at n = 1597) and the kernel is bound by instruction latency "
                    "(ncu: issue slots 14 % busy, shared-memory pipe 33 %), see profiles/r2_ncu_frontal.md",
        }
    else:
        pcg_bytes = (2 * pcg_geo["ell_width"] + 1) * 8.0 * n * iters_sum
        roofline = {"bound": "smem", "achieved": pcg_bytes / (k_ms * 1e-3) / 1e9, "peak": smem_peak, "unit": "GB/s",
                    "frac": pcg_bytes / (k_ms * 1e-3) / 1e9 / smem_peak, "kernel": "pcg_kernel<affine>", "kernel_ms": k_ms,
                    "algorithmic_bytes_per_launch": pcg_bytes, "traffic": None}
    t_k1 = ncu_t.get("pcg_kernel")
    pcg_smem_bytes = (2 * pcg_geo["ell_width"] + 1) * 8.0 * n * iters_sum
    fom_pcg = {
        "value": world * NF / ((ms_p / max(2, K // 2) - flush_ms) * 1e-3), "unit": "solves/s",
        "kernel": "pcg_kernel<affine> (K1, on-chip Jacobi-PCG of round 1; fom_solver = 1)", "mean_pcg_iters": iters_sum / NF,
        "all_converged": ok_p, "geometry": pcg_geo, "gpu_launches": launches_p,
        "roofline": {"bound": "smem", "achieved": pcg_smem_bytes / (k_ms_p * 1e-3) / 1e9, "peak": smem_peak, "unit": "GB/s",
                     "frac": pcg_smem_bytes / (k_ms_p * 1e-3) / 1e9 / smem_peak, "kernel_ms": k_ms_p,
                     "algorithmic_bytes_per_launch": pcg_smem_bytes,
                     "traffic": (t_k1["dram_bytes"] + max(0, NF - t_k1["launch_samples"]) * 152.0) if t_k1 else None,
                     "note": "shared-memory bytes per iteration = (W gathered residual entries + W values + 1 store) x 8 B x n "
                             "rows, W = ELL width; the CG vectors never leave the SM, so the SURVEY 8d HBM model (88 n bytes "
                             f"per iteration = {88.0 * n * iters_sum / (k_ms_p * 1e-3) / 1e9:.0f} GB/s here) is not a bound"},
    }

    line = {
        "metric": METRIC, "value": fom_rate, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": step_ms_f, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, world),     # identical to the reference arm's for the same flags
        "details": {
            "legs": "value = FOM (direct solver); fom_pcg = same workload, round-1 PCG kernel; rom = config[1] nine-param "
                    "ROM; fom_unstructured = reference-like mesh; fom_nodal = config[4]; fom_refined = config[3]",
            "solver": "direct" if fom_geo else "pcg", "frontal": fom_geo, "all_converged": ok_f and ok_r,
            "l2": f"flushed between steps by a 256 MiB write ({flush_ms:.3f} ms, subtracted)"},
        "roofline": roofline,
        "parity": parity,
        "gather_order_ok": gather_ok,
        "cpu_baseline": cpu,
        "e2e": {"value": fom_e2e_rate, "unit": "solves/s", "h2d_bytes_per_step": NF * 9 * 8,
                "d2h_bytes_per_step": NF * (n_obs * 8 + 4 + 4), "ms_per_step": ms_fe / K - flush_ms,
                "api": "tfin_fom_affine(TFIN_MEM_HOST) on pinned host buffers (the call AffineROMFin makes)"},
        "e2e_facade": {"value": fac_f, "unit": "solves/s", "api": "AffineROMFin.forward_nine_param_qoi(theta) with pageable "
                       "numpy in / numpy out, wall clock incl. Python", "rom_value": fac_r,
                       "rom_api": "AffineROMFin.forward_reduced_qoi(theta)", "matches_device_path_bitwise": facade_same},
        "gpu_launches": launches_f,
        "clocks": clocks,
        "fom_pcg": fom_pcg,
        "fom_unstructured": unstructured,
        "fom_refined": refined,
        "fom_nodal": nodal,
        "rom": {
            "value": rom_rate, "unit": "solves/s", "ms_per_step": step_ms_r,
            "e2e": {"value": rom_e2e_rate, "unit": "solves/s", "h2d_bytes_per_step": NR * 9 * 8,
                    "d2h_bytes_per_step": NR * (n_obs * 8 + 4), "ms_per_step": ms_re / K - flush_ms},
            "gpu_launches": launches_r,
            "roofline": {"bound": "fp64", "achieved": rom_rate / world * rom_flops / 1e12,
                         "peak": dgemm_tflops, "unit": "TFLOP/s",
                         "frac": rom_rate / world * rom_flops / 1e12 / dgemm_tflops,
                         "peak_source": "cuBLAS DGEMM 4096^3 measured in this run",
                         "flops_per_sample": rom_flops},
        },
        "smem_bandwidth_measured_gbs": smem_peak,
    }
    print(json.dumps(line), flush=True)
    if not parity_ok:
        sys.stderr.write(f"bench.py: PARITY FAILURE {json.dumps(parity)}\n")
        sys.exit(3)
    if gather_ok is False:
        sys.stderr.write("bench.py: gathered observables are not in shard order\n")
        sys.exit(4)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
