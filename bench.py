#!/usr/bin/env python
"""Benchmark of the thermal-fin batched forward map (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo's CUDA path)
    python bench.py --impl reference --gpus N ...             # reference arm: CPU port on the host cores

One "step" = one pass of the hot path over one batch of synthetic conductivity samples per GPU:
  * FOM leg (headline `value`): BASELINE config "five-param batched FOM solves ... 10^5 samples on the reference
    mesh, QoI = subfin averages" -> affine Jacobi-PCG kernel fused with B_obs (mesh: structured m=3, n = 1597)
  * ROM leg (`rom` object): BASELINE config "nine-param batched ROM solves, 10^6 samples" (n_r = 81)
Samples are sharded over ranks (weak scaling: the per-GPU batch is fixed); with N > 1 the step ends with the
NCCL all-gather of the observables.  Prints ONE JSON line on rank 0.
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
    ap.add_argument("--refined-steps", type=int, default=2)
    ap.add_argument("--nodal-batch", type=int, default=100_000,
                    help="nodal Gaussian-field FOM samples per GPU per step; 0 disables the leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-fom-sample", type=int, default=65536, help="CPU baseline FOM sample (~10 s on 16 cores)")
    ap.add_argument("--cpu-rom-sample", type=int, default=65536)
    ap.add_argument("--grad-batch", type=int, default=16384, help="samples per GPU per step of the gradient legs (0 = skip)")
    ap.add_argument("--ref-step-samples", type=int, default=8192, help="--impl reference: CPU solves per step")
    return ap.parse_args()


def fom_inputs(n, seed):
    """config 3: k ~ U(0.1, 1.0)^5 -> nine-vector [k1..k5,k4..k1] (generate_reduced_basis_five_param.py:34,57)."""
    k5 = np.random.default_rng(seed).uniform(0.1, 1.0, (n, 5))
    return np.ascontiguousarray(np.concatenate([k5, k5[:, 3::-1]], axis=1))


def rom_inputs(n, seed):
    """config 2: theta ~ U(0.1, 3.5)^9 (generate_reduced_basis_nine_param.py:299)."""
    return np.random.default_rng(seed).uniform(0.1, 3.5, (n, 9))


# ------------------------------------------------------------------------------------------------
# CPU legs (the ONLY place bench.py touches oracle/): reported baseline + the --impl reference arm
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    phi = cpu_pod_basis()
    pool = CpuPool(phi)
    per_step = max(pool.cores * 16, args.ref_step_samples)
    th = fom_inputs(per_step * (args.steps + args.warmup), FOM_SEED)
    thr = rom_inputs(per_step * (args.steps + args.warmup), ROM_SEED)
    t_f = t_r = 0.0
    for s in range(args.steps + args.warmup):
        _, dt = pool.run(_cpu_fom, th[s * per_step:(s + 1) * per_step])
        _, dtr = pool.run(_cpu_rom, thr[s * per_step:(s + 1) * per_step])
        if s >= args.warmup:
            t_f += dt
            t_r += dtr
    pool.close()
    v = per_step * args.steps / t_f
    vr = per_step * args.steps / t_r
    sample = f"{per_step} five-param FOM solves per step (assemble + scipy splu + B_obs), {pool.cores} processes"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_f / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config[2] five-param batched FOM on the m=3 mesh (n=1597), CPU port of the "
                               "reference path: FEniCS/PETSc are not installable here (parity unpinned)",
                   "samples_per_step": per_step},
        "cpu_baseline": {"value": v, "unit": "solves/s", "cores": pool.cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "rom": {"value": vr, "unit": "solves/s", "ms_per_step": 1e3 * t_r / args.steps,
                "sample": f"{per_step} literal LSPG ROM solves per step (A phi, psi^T psi, np.linalg.solve)"},
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


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # ---- CPU baseline first (fork-based pool, before this process creates a CUDA context); rank 0, N=1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        phi_cpu = cpu_pod_basis()
        pool = CpuPool(phi_cpu)
        _, dt_f = pool.run(_cpu_fom, fom_inputs(args.cpu_fom_sample, FOM_SEED))
        _, dt_r = pool.run(_cpu_rom, rom_inputs(args.cpu_rom_sample, ROM_SEED))
        pool.close()
        cpu = {"value": args.cpu_fom_sample / dt_f, "unit": "solves/s", "cores": pool.cores, "kind": "port",
               "sample": f"first {args.cpu_fom_sample} samples of the FOM workload: oracle port of the reference "
                         f"path (numpy assembly + scipy splu + B_obs), {pool.cores} single-threaded processes, "
                         f"{dt_f:.1f} s",
               "reference_published": "the reference publishes no benchmark; its stored notebook outputs show ~7.4 FOM "
                                      "solves/s (MUQ chain, 1 CPU process) and 13.6-14.4 it/s (PyMC3 Metropolis), BASELINE.md",
               "rom_value": args.cpu_rom_sample / dt_r,
               "rom_sample": f"first {args.cpu_rom_sample} ROM samples, literal averaged_affine_ROM.py:292-304 "
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
    th_f_host = torch.from_numpy(fom_inputs(NF, FOM_SEED + rank)).pin_memory()
    th_r_host = torch.from_numpy(rom_inputs(NR, ROM_SEED + rank)).pin_memory()
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

    def fom_dev():
        h.fom_affine_raw(th_f.data_ptr(), NF, _cabi.IN_PARAMS, _cabi.MEM_DEVICE, TOL, 20000, qoi=q_f.data_ptr(),
                         iters=it_f.data_ptr(), status=st_f.data_ptr(), stream=sp)
        return gather_rows(q_f, NF * world) if world > 1 else q_f

    def rom_dev():
        h.rom_raw(th_r.data_ptr(), NR, _cabi.IN_PARAMS, _cabi.MEM_DEVICE, qoi=q_r.data_ptr(),
                  status=st_r.data_ptr(), stream=sp)
        return gather_rows(q_r, NR * world) if world > 1 else q_r

    def fom_e2e():     # the C-ABI call a user of the facade makes: HOST buffers in, HOST buffers out
        h.fom_affine_raw(th_f_host.data_ptr(), NF, _cabi.IN_PARAMS, _cabi.MEM_HOST, TOL, 20000,
                         qoi=q_f_host.data_ptr(), iters=it_f_host.data_ptr(), status=st_f_host.data_ptr(), stream=sp)

    def rom_e2e():
        h.rom_raw(th_r_host.data_ptr(), NR, _cabi.IN_PARAMS, _cabi.MEM_HOST, qoi=q_r_host.data_ptr(),
                  status=st_r_host.data_ptr(), stream=sp)

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

    # flush cost (inside the bracket) measured once so it can be reported
    timed(lambda: None, 2, 1)
    flush_ms, _ = timed(lambda: None, 4, 1)
    flush_ms /= 4

    sampler = ClockSampler(local_rank)
    sampler.start()
    K, Wm = args.steps, max(args.warmup, 3)
    ms_f, launches_f = timed(fom_dev, K, Wm)
    ms_r, launches_r = timed(rom_dev, K, Wm)
    ms_fe, launches_fe = timed(fom_e2e, K, Wm)
    ms_re, launches_re = timed(rom_e2e, K, Wm)

    # ---- config[4] nodal Gaussian-field conductivity: k = exp(0.5 chol^T z), Matern-5/2, l = 1.6 (fields are
    # generated on the device with torch -- input generation, not the measured path), in-kernel assembly + PCG
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

        def nodal_dev():
            hn.fom_nodal_raw(k_dev.data_ptr(), NN, _cabi.MEM_DEVICE, TOL, 20000, qoi=q_n.data_ptr(),
                             iters=it_n.data_ptr(), status=st_n.data_ptr(), stream=sp)
            return gather_rows(q_n, NN * world) if world > 1 else q_n

        q_n_host = torch.empty((NN, n_obs), dtype=f64).pin_memory()
        st_n_host = torch.empty(NN, dtype=torch.int32).pin_memory()

        def nodal_e2e():
            hn.fom_nodal_raw(k_host.data_ptr(), NN, _cabi.MEM_HOST, TOL, 20000, qoi=q_n_host.data_ptr(),
                             status=st_n_host.data_ptr(), stream=sp)

        def nodal_pipeline():   # fields drawn ON the device (tfin_field_sample), only the observables cross PCIe
            draw()
            hn.fom_nodal_raw(k_dev.data_ptr(), NN, _cabi.MEM_DEVICE, TOL, 20000, qoi=q_n.data_ptr(),
                             status=st_n.data_ptr(), stream=sp)
            q_n_host.copy_(q_n, non_blocking=True)

        ms_n, ln = timed(nodal_dev, K, Wm, hn)
        ms_ne, _ = timed(nodal_e2e, K, Wm, hn)
        ms_np, lnp = timed(nodal_pipeline, K, Wm, hn)
        it_sum_n = int(it_n.to(torch.int64).sum().item())
        n_cells = fin.ops.n_cells
        step_n = ms_n / K - flush_ms
        nodal = {
            "value": world * NN / (step_n * 1e-3), "unit": "solves/s", "ms_per_step": step_n,
            "workload": f"config[4]: {NN} Matern-5/2 (l=1.6) nodal fields per GPU, per-sample in-kernel FEM assembly "
                        f"+ PCG, mesh m=3",
            "mean_pcg_iters": it_sum_n / NN, "all_converged": bool((st_n == 0).all().item()),
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
            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": None,
                         "achieved": ((88.0 * n + 8.0 * n_cells) * it_sum_n + NN * n * 8.0) / (step_n * 1e-3) / 1e9,
                         "note": "algorithmic bytes (88 n + 8 n_cells) per iteration + the k field read once; "
                                 "on-chip kernel, see roofline.note"},
        }
        # ---- gradient legs: FOM adjoint gradient and reduced gradient, device resident
        if args.grad_batch > 0:
            NG = min(args.grad_batch, NN)
            data_dev = q_n[:1].clone()
            g_n = torch.empty((NG, n), dtype=f64, device=dev)
            c_n = torch.empty(NG, dtype=f64, device=dev)

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
                        "what": f"Fin.gradient batched: forward + adjoint PCG + gradient form in one kernel, {NG} fields",
                        "gpu_launches": lg, "all_converged": bool((st_n[:NG] == 0).all().item())},
                "rom": {"value": world * NG / ((ms_gr / K - flush_ms) * 1e-3), "unit": "gradients/s",
                        "what": f"AffineROMFin.grad_reduced batched (nodal in, nodal out), {NG} fields",
                        "gpu_launches": lgr}}
        del k_dev, k_host
        fin.handle.close()

    # ---- config[3] refined mesh (m=26, n=99 945), nine-param FOM: the genuinely HBM-streaming PCG kernel
    refined = None
    if args.refined_batch > 0:
        from bayesianinferencedl_b200.assembly import build_operators
        NRf = args.refined_batch
        Vr = get_space(RESOLUTION, m=REFINED_M)
        opsr = build_operators(Vr)
        hr = _cabi.TfinHandle(local_rank)
        hr.set_operator(opsr.row_ptr, opsr.col_idx, opsr.vals, opsr.rhs)
        hr.set_observation(*opsr.obs_csr())
        th_ref_host = torch.from_numpy(np.random.default_rng(REF_SEED + rank).uniform(0.1, 10.0, (NRf, 9))).pin_memory()
        th_ref = th_ref_host.to(dev)
        q_ref = torch.empty((NRf, 9), dtype=f64, device=dev)
        it_ref = torch.empty(NRf, dtype=torch.int32, device=dev)
        st_ref = torch.empty(NRf, dtype=torch.int32, device=dev)

        def ref_dev():
            hr.fom_affine_raw(th_ref.data_ptr(), NRf, _cabi.IN_PARAMS, _cabi.MEM_DEVICE, TOL, 50000,
                              qoi=q_ref.data_ptr(), iters=it_ref.data_ptr(), status=st_ref.data_ptr(), stream=sp)
            return gather_rows(q_ref, NRf * world) if world > 1 else q_ref

        ms_ref, l_ref = timed(ref_dev, args.refined_steps, 1, hr)
        step_ref = ms_ref / args.refined_steps - flush_ms
        it_sum_ref = int(it_ref.to(torch.int64).sum().item())
        bytes_ref = 88.0 * opsr.n * it_sum_ref
        ach = bytes_ref / (step_ref * 1e-3) / 1e9
        refined = {
            "value": world * NRf / (step_ref * 1e-3), "unit": "solves/s", "ms_per_step": step_ref,
            "steps": args.refined_steps, "warmup": 1,
            "workload": f"config[3]: refined mesh m={REFINED_M}, n={opsr.n} dofs, nine-param theta~U(0.1,10), "
                        f"{NRf} samples per GPU per step, streaming PCG kernel (tile {hr.get_int('stream_tile')})",
            "mean_pcg_iters": it_sum_ref / NRf, "all_converged": bool((st_ref == 0).all().item()),
            "gpu_launches": l_ref,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": None, "unit": "GB/s", "frac": None,
                         "kernel": "pcg_stream_kernel", "algorithmic_bytes_per_launch": bytes_ref,
                         "note": "88*n bytes per iteration per sample (SURVEY 8d), each sample counted with its own "
                                 "iteration count; vectors stream from HBM every pass (working set >> L2)"},
        }
        hr.close()
    clocks = sampler.stop()

    # ---- kernel-only duration of the dominant kernel (PCG) for the roofline: one more step, events tight
    # around the single launch on its stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    for _ in range(3):
        flush.fill_(1)
        e0.record(stream)
        h.fom_affine_raw(th_f.data_ptr(), NF, _cabi.IN_PARAMS, _cabi.MEM_DEVICE, TOL, 20000, qoi=q_f.data_ptr(),
                         iters=it_f.data_ptr(), status=st_f.data_ptr(), stream=sp)
        e1.record(stream)
        torch.cuda.synchronize()
        kms.append(e0.elapsed_time(e1))
    k_ms = float(np.mean(kms))
    iters_sum = int(it_f.to(torch.int64).sum().item())
    ok_f = bool((st_f == 0).all().item())
    ok_r = bool((st_r == 0).all().item())

    # DGEMM rate of this GPU (cuBLAS) as the FP64 denominator for the ROM leg
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

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    traffic, traffic_note, ncu_t = None, None, {}
    try:
        ncu_t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        t = ncu_t["pcg_kernel"]
        # DRAM traffic of the on-chip kernel: operator arrays once + 152 B/sample, scaled to this launch's batch
        traffic = t["dram_bytes"] + max(0, NF - t["launch_samples"]) * 152.0
        traffic_note = (f"ncu --set full capture of a {t['launch_samples']}-sample launch: {t['dram_bytes']:.0f} B "
                        f"(profiles/r1_ncu_summary.md), plus 152 B/sample of theta/qoi for the larger batch")
    except Exception:
        pass

    # algorithmic bytes of one PCG launch: 88 n bytes per iteration per sample + theta in, qoi out
    alg_bytes = 88.0 * n * iters_sum + NF * (9 * 8 + n_obs * 8 + 8)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    if refined is not None:
        refined["roofline"]["peak"] = hbm_peak
        refined["roofline"]["frac"] = refined["roofline"]["achieved"] / hbm_peak
        refined["roofline"]["peak_source"] = peak_src
        ts = ncu_t.get("pcg_stream_kernel")
        refined["roofline"]["traffic"] = (ts["ratio_to_algorithmic"] * refined["roofline"]["algorithmic_bytes_per_launch"]
                                          if ts else None)
        refined["roofline"]["traffic_note"] = (
            f"ncu capture ({ts['launch']}): dram bytes / algorithmic bytes = {ts['ratio_to_algorithmic']:.4f}, applied "
            f"to this launch" if ts else None)
    if nodal is not None:
        nodal["roofline"]["peak"] = hbm_peak
        nodal["roofline"]["frac"] = nodal["roofline"]["achieved"] / hbm_peak
    step_ms_f = (ms_f / K) - flush_ms
    step_ms_r = (ms_r / K) - flush_ms
    fom_rate = world * NF / (step_ms_f * 1e-3)
    rom_rate = world * NR / (step_ms_r * 1e-3)
    fom_e2e_rate = world * NF / ((ms_fe / K - flush_ms) * 1e-3)
    rom_e2e_rate = world * NR / ((ms_re / K - flush_ms) * 1e-3)
    rom_flops = 2 * 55 * 3321 + 2 * 10 * 81 + 81 ** 3 / 3 + 2 * 81 ** 2 + 2 * 9 * 81     # SURVEY 8d, ~559 kflop

    line = {
        "metric": METRIC, "value": fom_rate, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": step_ms_f, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "FOM leg (value): config[2] five-param batched FOM, QoI = subfin averages; ROM leg (rom): "
                        "config[1] nine-param batched ROM; fom_refined: config[3]; fom_nodal: config[4]",
            "fom_samples_per_gpu_per_step": NF, "rom_samples_per_gpu_per_step": NR,
            "mesh": f"structured conforming fin mesh m=3, n={n} dofs (reference mshr mesh is not shipped)",
            "n_r": n_r, "pcg_tol": TOL, "mean_pcg_iters": iters_sum / NF, "all_converged": ok_f and ok_r,
            "pcg_geometry": {"threads": h.get_int("pcg_threads"), "rows_per_thread": h.get_int("pcg_rows_per_thread"),
                             "ctas_per_sm": h.get_int("pcg_ctas_per_sm"), "smem_bytes": h.get_int("pcg_smem_bytes"),
                             "ell_width": h.get_int("ell_width")},
            "l2": f"flushed between steps by a 256 MiB write ({flush_ms:.3f} ms, subtracted)",
            "parallelism": f"samples sharded x{world}, NCCL all-gather of observables" if world > 1 else "1 GPU",
        },
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": traffic, "traffic_note": traffic_note, "kernel": "pcg_kernel<affine>", "kernel_ms": k_ms,
            "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
            "onchip": ({"bound": "smem", "pct_of_peak": ncu_t["pcg_kernel"].get("smem_pipe_pct_of_peak"),
                        "metric": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                        "source": "profiles/r1_ncu_summary.json (ncu --set full of a 2960-sample launch)"}
                       if "pcg_kernel" in ncu_t else None),
            "note": "algorithmic bytes = 88*n*iterations per solve (SURVEY 8d); at n=1597 the CG vectors and the "
                    "per-sample operator live in shared memory/registers, so achieved/peak is NOT bounded by 1 "
                    "and real DRAM traffic (traffic) is orders of magnitude below it",
        },
        "cpu_baseline": cpu,
        "e2e": {"value": fom_e2e_rate, "unit": "solves/s", "h2d_bytes_per_step": NF * 9 * 8,
                "d2h_bytes_per_step": NF * (n_obs * 8 + 4 + 4), "ms_per_step": ms_fe / K - flush_ms,
                "api": "tfin_fom_affine(TFIN_MEM_HOST) on pinned host buffers (the call AffineROMFin makes)"},
        "gpu_launches": launches_f,
        "clocks": clocks,
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
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
