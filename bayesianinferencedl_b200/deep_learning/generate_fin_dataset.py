"""Batched generator of the ROM-error training set, device-resident end to end.

Reference: ``deep_learning/generate_fin_dataset.py:62-112`` (``gen_affine_avg_rom_dataset``): for every sample draw
a log-normal Matern conductivity field, solve the nodal full-order model and the sub-fin-averaged affine ROM, and
store ``z_s`` (the fields), ``qois`` and ``qoi_errors = qoi - qoi_r``; the ``.npy`` names depend on the set size.

The reference loops one sample at a time through FEniCS.  Here each chunk of samples is ONE pass of four kernels
over device buffers -- field sampler (csrc/field.cuh) -> nodal FOM PCG + B_obs (K2) -> sub-fin averages + ROM
(K0, R1, R2) -- and only the finished arrays cross PCIe.
"""
from __future__ import annotations

import os

import numpy as np

from .. import _cabi
from ..bayesian_inference.gaussian_field import FieldSampler
from ..fom.forward_solve import Fin
from ..fom.thermal_fin import get_space
from ..rom.averaged_affine_ROM import AffineROMFin
from ..rom.pod import generate_pod_basis

__all__ = ["gen_affine_avg_rom_dataset", "DatasetGenerator"]


class DatasetGenerator:
    """Owns the three device models (prior, nodal FOM, averaged affine ROM) of one space."""

    def __init__(self, V, phi, external_obs=False, length=1.6, device=0, chunk=16384):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("DatasetGenerator needs a CUDA device (there is no CPU fallback)")
        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.solver = Fin(V, external_obs, device=device)
        self.solver_r = AffineROMFin(V, None, phi, external_obs, device=device)
        self.prior = FieldSampler(V, "m52", length, handle=self.solver.handle)
        self.n, self.n_obs = self.solver.dofs, self.solver.n_obs
        self.chunk = int(chunk)
        self._bufs = None

    def generate(self, dataset_size, seed=0, z=None):
        """Returns ``(z_s, qoi_errors, qois)`` as numpy arrays of shapes (N, n), (N, n_obs), (N, n_obs).
        ``z``: optional (N, n) standard normals (otherwise drawn on the device: sample s is the Philox stream
        (seed, s), independent of the chunking).

        Chunks are double buffered: while chunk c runs its four kernels on the compute stream, the fields and
        observables of chunk c-1 stream to page-locked staging buffers on a second stream and are then copied into the
        output arrays by the host, so PCIe (12.8 KB of field per sample -- the data set stores the fields) and the
        host-side copy hide behind the solves."""
        torch, dev, n, nobs = self.torch, self.dev, self.n, self.n_obs
        N = int(dataset_size)
        f64 = torch.float64
        z_s, errs, qois = np.empty((N, n)), np.empty((N, nobs)), np.empty((N, nobs))
        h_f, h_r = self.solver.handle, self.solver_r.handle
        lib = h_f._lib
        cap = min(self.chunk, max(N, 1))
        if self._bufs is None or self._bufs[0]["k"].shape[0] < cap:   # device + pinned staging, allocated once
            self._bufs = [{
                "k": torch.empty((cap, n), dtype=f64, device=dev), "q": torch.empty((cap, nobs), dtype=f64, device=dev),
                "q_r": torch.empty((cap, nobs), dtype=f64, device=dev),
                "st": torch.empty((2, cap), dtype=torch.int32, device=dev),
                "z": torch.empty((cap, n), dtype=f64, device=dev),
                "hk": torch.empty((cap, n), dtype=f64, pin_memory=True),
                "hq": torch.empty((cap, nobs), dtype=f64, pin_memory=True),
                "he": torch.empty((cap, nobs), dtype=f64, pin_memory=True),
                "hst": torch.empty((2, cap), dtype=torch.int32, pin_memory=True),
                "computed": torch.cuda.Event(), "copied": torch.cuda.Event(), "pending": None} for _ in range(2)]
            self._streams = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        bufs, (compute, copy) = self._bufs, self._streams
        failed = False

        def drain(b):              # host side of a finished chunk: staging -> output arrays
            nonlocal failed
            s0, m = b["pending"]
            b["copied"].synchronize()
            z_s[s0:s0 + m] = b["hk"][:m].numpy()
            qois[s0:s0 + m] = b["hq"][:m].numpy()
            errs[s0:s0 + m] = b["he"][:m].numpy()
            failed = failed or bool(b["hst"][:, :m].max() != 0)
            b["pending"] = None

        for c, s0 in enumerate(range(0, N, cap)):
            m = min(cap, N - s0)
            b = bufs[c & 1]
            if b["pending"] is not None:                               # chunk c-2 still sits in this staging set;
                drain(b)                                               # the GPU is busy with chunk c-1 meanwhile
            with torch.cuda.stream(compute):
                sp = compute.cuda_stream
                zp = None
                if z is not None:
                    b["z"][:m].copy_(torch.from_numpy(np.ascontiguousarray(z[s0:s0 + m], dtype=np.float64)))
                    zp = b["z"].data_ptr()
                _cabi._check(lib, lib.tfin_field_sample(h_f._h, zp, int(seed), 0, s0, m, _cabi.MEM_DEVICE,
                                                        b["k"].data_ptr(), None, sp), "tfin_field_sample")
                h_f.fom_nodal_raw(b["k"].data_ptr(), m, _cabi.MEM_DEVICE, self.solver.tol, self.solver.maxit,
                                  qoi=b["q"].data_ptr(), status=b["st"][0].data_ptr(), stream=sp)
                h_r.rom_raw(b["k"].data_ptr(), m, _cabi.IN_NODAL, _cabi.MEM_DEVICE, qoi=b["q_r"].data_ptr(),
                            status=b["st"][1].data_ptr(), stream=sp)
                b["q_r"][:m].sub_(b["q"][:m]).neg_()                   # qoi - qoi_r  (generate_fin_dataset.py:99)
                b["computed"].record(compute)
            with torch.cuda.stream(copy):
                copy.wait_event(b["computed"])
                b["hk"][:m].copy_(b["k"][:m], non_blocking=True)
                b["hq"][:m].copy_(b["q"][:m], non_blocking=True)
                b["he"][:m].copy_(b["q_r"][:m], non_blocking=True)
                b["hst"][:, :m].copy_(b["st"][:, :m], non_blocking=True)
                b["copied"].record(copy)
            b["pending"] = (s0, m)
        for b in bufs:
            if b["pending"] is not None:
                drain(b)
        if failed:
            raise RuntimeError("dataset generation: a forward solve failed (non-positive conductivity?)")
        return z_s, errs, qois


def gen_affine_avg_rom_dataset(dataset_size, resolution=40, genrand=False, *, phi=None, V=None, out_dir=None,
                               seed=0, device=0):
    """generate_fin_dataset.py:62-112.  Returns ``(z_s, qoi_errors)`` and writes the reference's ``.npy`` files into
    ``out_dir`` when given (the reference hard-codes ``../data``): ``*_tr_avg_obs_3`` for more than 1000 samples,
    ``*_eval_avg_obs_3`` for fewer than 600.  ``phi`` defaults to a POD basis built on this space (the reference loads
    ``data/basis_nine_param.txt``, which belongs to its unshipped mshr mesh)."""
    V = V if V is not None else get_space(resolution)
    if phi is None:
        phi = generate_pod_basis(V, device=device)
    gen = DatasetGenerator(V, phi, external_obs=genrand, device=device)
    z_s, qoi_errors, qois = gen.generate(dataset_size, seed=seed)
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        tag = "tr" if dataset_size > 1000 else "eval" if dataset_size < 600 else None
        if tag:
            np.save(os.path.join(out_dir, f"z_aff_avg_{tag}_avg_obs_3"), z_s)
            np.save(os.path.join(out_dir, f"errors_aff_avg_{tag}_avg_obs_3"), qoi_errors)
            np.save(os.path.join(out_dir, f"qois_avg_{tag}_avg_obs_3"), qois)
    return z_s, qoi_errors
