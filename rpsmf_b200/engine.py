# -*- coding: utf-8 -*-
"""Host-side handle on the CUDA filter engine (thin wrapper over the C ABI).

PyTorch is used for device memory and streams only; every arithmetic step of
the filter runs in libpsmf_b200.so.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi


def _dev_tensor(a, dtype, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device).contiguous()


class FilterEngine:
    """One engine = one (batch of) series on one GPU.

    Parameters mirror ``psmf_config`` in include/psmf_b200.h.
    """

    def __init__(self, d, r, *, n_series=1, dtype=torch.float64, robust=True, simplified=False,
                 c_update_transpose=True, fixed_lambda=False, ll_student=False, dynamics=_capi.DYN_IDENTITY, alpha=1.0,
                 beta=1.0, device=None, d_global=None, world_size=1, rank=0, ctas=0, kernel=0):
        if not torch.cuda.is_available():
            raise RuntimeError("rpsmf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if dtype not in (torch.float64, torch.float32):
            raise ValueError("dtype must be torch.float64 or torch.float32")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.d, self.r, self.S = int(d), int(r), int(n_series)
        self.dtype = dtype
        self.robust = bool(robust)
        flags = 0
        flags |= _capi.ROBUST if robust else 0
        flags |= _capi.SIMPLIFIED if simplified else 0
        flags |= _capi.CUPDATE_VT if c_update_transpose else 0
        flags |= _capi.FIXED_LAMBDA if fixed_lambda else 0
        flags |= _capi.LL_STUDENT if ll_student else 0
        cfg = _capi.PsmfConfig(
            d=self.d, d_global=int(d_global or d), r=self.r, n_series=self.S,
            dtype=_capi.F64 if dtype == torch.float64 else _capi.F32, flags=flags, dynamics=int(dynamics),
            device=self.device.index, world_size=int(world_size), rank=int(rank), ctas=int(ctas), kernel=int(kernel),
            alpha=float(alpha), beta=float(beta))
        self._L = _capi.lib()
        self._h = C.c_void_p()
        rc = self._L.psmf_create(C.byref(self._h), C.byref(cfg))
        if rc != 0:
            msg = self._L.psmf_last_error(None)
            raise _capi.PsmfError(rc, msg.decode() if msg else "psmf_create failed")
        self.dynamics = int(dynamics)
        self._keep = []

    # -- lifecycle ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            torch.cuda.synchronize(self.device)
            self._L.psmf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ck(self, rc):
        _capi.check(self._h, rc)

    # -- state -------------------------------------------------------------------------------------
    def set_state(self, C_=None, V=None, P=None, x=None, Q=None, rho=None, lam=None, theta=None):
        S, d, r = self.S, self.d, self.r
        f64 = torch.float64

        def prep(a, shape, dt):
            if a is None:
                return None
            t = _dev_tensor(a, dt, self.device)
            if t.numel() != int(np.prod(shape)):
                if t.numel() * S == int(np.prod(shape)):       # broadcast one state over the batch
                    t = t.reshape((1,) + tuple(shape[1:])).expand(shape).contiguous()
                else:
                    raise ValueError("state tensor has %d elements, expected %s" % (t.numel(), shape))
            return t.reshape(shape)

        ts = [prep(C_, (S, d, r), self.dtype), prep(V, (S, r, r), f64), prep(P, (S, r, r), f64), prep(x, (S, r), f64),
              prep(Q, (S, r, r), f64), prep(rho, (S,), f64), prep(lam, (S,), f64), prep(theta, (S, r), f64)]
        ptrs = [C.c_void_p(t.data_ptr()) if t is not None else None for t in ts]
        with torch.cuda.device(self.device):
            self._ck(self._L.psmf_set_state(self._h, *ptrs, self._stream()))
        self._keep = ts   # keep sources alive until the async copies are ordered behind later work

    def get_state(self, want_C=True):
        S, d, r = self.S, self.d, self.r
        f64 = torch.float64
        dev = self.device
        out = dict(
            C=torch.empty((S, d, r), dtype=self.dtype, device=dev) if want_C else None,
            V=torch.empty((S, r, r), dtype=f64, device=dev), P=torch.empty((S, r, r), dtype=f64, device=dev),
            x=torch.empty((S, r), dtype=f64, device=dev), Q=torch.empty((S, r, r), dtype=f64, device=dev),
            rho=torch.empty((S,), dtype=f64, device=dev), lam=torch.empty((S,), dtype=f64, device=dev),
            theta=torch.empty((S, r), dtype=f64, device=dev))
        order = ("C", "V", "P", "x", "Q", "rho", "lam", "theta")
        ptrs = [C.c_void_p(out[k].data_ptr()) if out[k] is not None else None for k in order]
        with torch.cuda.device(self.device):
            self._ck(self._L.psmf_get_state(self._h, *ptrs, self._stream()))
        if self.S == 1:
            out = {k: (v[0] if v is not None else None) for k, v in out.items()}
        return out

    # -- the hot path ------------------------------------------------------------------------------
    def run(self, Y, M=None, k0=1, want_X=True, want_Yrec=False, want_scal=False, xbar=None, F=None,
            X_out=None, Yrec_out=None, scal_out=None, want_grad=False):
        """Filter ``T`` steps.  Y: device tensor (T, d) or (S, T, d) in the engine dtype (last dim may be
        padded: ld = stride of the time axis); M: uint8 tensor of the same shape or None."""
        S, d, r = self.S, self.d, self.r
        if not isinstance(Y, torch.Tensor) or Y.device != self.device:
            raise ValueError("Y must be a tensor on %s" % self.device)
        if Y.dtype != self.dtype:
            raise ValueError("Y dtype %s does not match the engine dtype %s" % (Y.dtype, self.dtype))
        if Y.dim() == 2:
            Y = Y.unsqueeze(0)
        if Y.dim() != 3 or Y.shape[0] != S or Y.shape[2] < d or Y.stride(2) != 1:
            raise ValueError("Y must be (S, T, >=d) with unit stride along d")
        T = Y.shape[1]
        io = _capi.PsmfIO()
        io.Y, io.ldy, io.y_series_stride = Y.data_ptr(), Y.stride(1), Y.stride(0)
        if M is not None:
            if M.dtype != torch.uint8 or M.device != self.device:
                raise ValueError("M must be a uint8 tensor on %s" % self.device)
            if M.dim() == 2:
                M = M.unsqueeze(0)
            if M.shape[0] != S or M.shape[1] != T or M.shape[2] < d or M.stride(2) != 1:
                raise ValueError("M must be (S, T, >=d) with unit stride along d")
            io.M, io.ldm, io.m_series_stride = M.data_ptr(), M.stride(1), M.stride(0)
        res = {}
        if want_X or X_out is not None:
            X = X_out if X_out is not None else torch.empty((S, T, r), dtype=torch.float64, device=self.device)
            io.X_out = X.data_ptr()
            res["X"] = X
        if want_Yrec or Yrec_out is not None:
            Yr = Yrec_out if Yrec_out is not None else torch.empty((S, T, d), dtype=self.dtype, device=self.device)
            if Yr.dim() == 2:
                Yr = Yr.unsqueeze(0)
            io.Yrec_out, io.ldrec, io.rec_series_stride = Yr.data_ptr(), Yr.stride(1), Yr.stride(0)
            res["Yrec"] = Yr
        if want_scal or scal_out is not None:
            sc = scal_out if scal_out is not None else torch.empty((S, T, _capi.NSCAL), dtype=torch.float64, device=self.device)
            io.scal_out = sc.data_ptr()
            res["scal"] = sc
        if want_grad:
            gr = torch.zeros((S, r), dtype=torch.float64, device=self.device)
            io.grad_out = gr.data_ptr()
            res["grad"] = gr
        keep = [Y, M]
        if self.dynamics == _capi.DYN_EXTERNAL:
            xb = _dev_tensor(xbar, torch.float64, self.device).reshape(S, r)
            io.xbar_ext = xb.data_ptr()
            keep.append(xb)
            if F is not None:
                Ft = _dev_tensor(F, torch.float64, self.device).reshape(S, r, r)
                io.F_ext = Ft.data_ptr()
                keep.append(Ft)
        with torch.cuda.device(self.device):
            self._ck(self._L.psmf_run(self._h, C.byref(io), T, int(k0), self._stream()))
        self._keep_run = keep
        if S == 1:
            res = {k: v[0] for k, v in res.items()}
        return res

    def run_host(self, Y_host, M_host=None, window=256, k0=1, want_X=True):
        """Filter a long HOST-resident sequence: Y_host (T, d) pinned CPU tensor in the engine dtype, M_host
        (T, d) uint8.  Windows of ``window`` steps are copied host->device on a side stream while the
        previous window is being filtered (double buffering); returns X (T, r) on the host (pinned).
        Single-series engines only."""
        if self.S != 1:
            raise ValueError("run_host supports single-series engines")
        T, d = Y_host.shape[0], self.d
        dev = self.device
        key = (int(window), int(Y_host.shape[1]), M_host is not None)
        if getattr(self, "_hb_key", None) != key:
            self._hb_key = key
            ld = Y_host.shape[1]
            self._hb = [torch.empty((window, ld), dtype=self.dtype, device=dev) for _ in range(2)]
            self._hm = [torch.empty((window, ld), dtype=torch.uint8, device=dev) for _ in range(2)] if M_host is not None else None
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._hx = torch.empty((2, window, self.r), dtype=torch.float64, device=dev)
        Xh = torch.empty((T, self.r), dtype=torch.float64).pin_memory() if want_X else None
        main = torch.cuda.current_stream(dev)
        cs = self._copy_stream
        cs.wait_stream(main)                        # staging buffers may still be read by an earlier run_host call
        nwin = (T + window - 1) // window
        copied = [None, None]
        done = [None, None]

        def issue_copy(w):
            a, b = w * window, min(T, (w + 1) * window)
            slot = w & 1
            if done[slot] is not None:
                cs.wait_event(done[slot])          # the run that last read this slot has finished
            with torch.cuda.stream(cs):
                self._hb[slot][: b - a].copy_(Y_host[a:b], non_blocking=True)
                if M_host is not None:
                    self._hm[slot][: b - a].copy_(M_host[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            copied[slot] = ev

        issue_copy(0)
        for w in range(nwin):
            a, b = w * window, min(T, (w + 1) * window)
            slot = w & 1
            if w + 1 < nwin:
                issue_copy(w + 1)
            main.wait_event(copied[slot])
            out = self.run(self._hb[slot][: b - a], None if M_host is None else self._hm[slot][: b - a], k0=k0 + a,
                           want_X=False, X_out=self._hx[slot][: b - a].unsqueeze(0) if want_X else None)
            if want_X:
                Xh[a:b].copy_(self._hx[slot][: b - a], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(main)
            done[slot] = ev
        if want_X:
            done[(nwin - 1) & 1].synchronize()      # the returned host tensor is complete when the caller reads it
        return Xh

    def connect(self, dist):
        """Row sharding over several GPUs: exchange the mailbox blobs (CUDA IPC handle + plan record) through the
        process group `dist` (torch.distributed, any backend) and map every peer's mailbox.  The library derives
        the kernel choice from ALL ranks' records, so every rank runs the same kernel."""
        world = dist.get_world_size()
        buf = (C.c_ubyte * _capi.MAILBOX_BLOB_BYTES)()
        self._ck(self._L.psmf_mailbox_export(self._h, buf))
        mine = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
        backend = dist.get_backend()
        if backend == "nccl":
            mine = mine.to(self.device)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        blob = b"".join(bytes(g.cpu().tolist()) for g in gathered)
        arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._ck(self._L.psmf_mailbox_connect(self._h, arr, world))
        dist.barrier()

    def status(self):
        """Synchronise and return the first step with a non-finite N/omega/phi/x, or -1."""
        bad = C.c_int64(-1)
        self._ck(self._L.psmf_status(self._h, C.byref(bad)))
        return int(bad.value)

    def set_trace(self, steps):
        """Debug: record phase time stamps (ns, globaltimer) of CTA 0 for the first `steps` steps of each run."""
        self._trace = torch.zeros((steps * (16 + 320),), dtype=torch.int64, device=self.device) if steps else None
        self._ck(self._L.psmf_set_trace(self._h, C.c_void_p(self._trace.data_ptr()) if steps else None, int(steps)))
        return self._trace

    def launch_info(self):
        a, b, c, dd = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self._L.psmf_launch_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(dd)))
        k, ns, res = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self._L.psmf_launch_info2(self._h, C.byref(k), C.byref(ns), C.byref(res)))
        return dict(ctas=a.value, threads=b.value, smem_bytes=c.value, launches=dd.value,
                    kernel={0: "none", 1: "direct", 2: "tma"}[k.value], nslot=ns.value, resident=bool(res.value))
