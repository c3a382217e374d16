"""Thin object layer over the C ABI: device replay ring + fused step engine.

torch is used here only as plumbing (device memory, streams); all compute happens in
libilswiss_b200.so.  The reference-facing classes (replay_buffer.py, sac.py, td3.py,
adv_irl.py) are built on these two objects."""
import ctypes as C

import numpy as np
import torch

from . import _abi, layout
from ._lib import IlswError, check, load


# NVTX ranges around every call that launches work (tracing) -- visible in nsys / ncu timelines;
# ILSW_NVTX=0 removes them (two ~0.3 us calls per train call when no profiler is attached)
import os as _os
_NVTX = _os.environ.get("ILSW_NVTX", "1") != "0"


class _Range:
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if _NVTX:
            torch.cuda.nvtx.range_pop()


def _stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def require_cuda():
    if not torch.cuda.is_available():
        raise IlswError("ilswiss_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


class ReplayRing:
    """HBM-resident replay ring (ilsw_rb_*).  Hot rows: [obs|act|rew|term|next_obs|pad]."""

    def __init__(self, capacity, obs_dim, act_dim):
        require_cuda()
        self.lib = load()
        self.capacity, self.obs_dim, self.act_dim = int(capacity), int(obs_dim), int(act_dim)
        h = C.c_void_p()
        check(self.lib.ilsw_rb_create(C.byref(h), self.capacity, self.obs_dim, self.act_dim), "rb_create")
        self.h = h
        self.stride = self.lib.ilsw_rb_row_stride(h)
        self.host_w = self.lib.ilsw_rb_host_row_floats(h)
        self._copy_stream = torch.cuda.Stream()
        self._pinned = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.ilsw_rb_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def size(self):
        return int(self.lib.ilsw_rb_size(self.h)) + self.pending

    @property
    def committed_size(self):
        return int(self.lib.ilsw_rb_size(self.h))

    @property
    def top(self):
        return int(self.lib.ilsw_rb_top(self.h))

    pending = 0

    def append_host(self, rows):
        """rows: float32 [n, host_w] (layout.pack_host_rows).  Staged with cudaMemcpyAsync from
        pinned memory on a side stream; enters the ring at the next commit()/train()."""
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        n = rows.shape[0]
        assert rows.shape[1] == self.host_w, (rows.shape, self.host_w)
        # pinned ring: a slot is rewritten only after the copy stream has drained (once per lap), so
        # back-to-back appends never block the host on the previous H2D copy
        if self._pinned is None or self._pinned.shape[0] < n:
            self._copy_stream.synchronize()
            self._pinned = torch.empty((max(2 * n, 4096), self.host_w), dtype=torch.float32).pin_memory()
            self._pinned_np = self._pinned.numpy()
            self._pin_cur = 0
        if self._pin_cur + n > self._pinned.shape[0]:
            self._copy_stream.synchronize()
            self._pin_cur = 0
        c = self._pin_cur
        self._pinned_np[c:c + n] = rows
        self._pin_cur = c + n
        check(self.lib.ilsw_rb_append(self.h, C.c_void_p(self._pinned.data_ptr() + c * self.host_w * 4), n,
                                      C.c_void_p(self._copy_stream.cuda_stream)), "rb_append")
        self.pending += n
        if self.pending > self.capacity:
            self.commit()

    def commit(self, stream=None):
        with _Range("ilsw_rb_commit"):
            check(self.lib.ilsw_rb_commit(self.h, _stream_ptr(stream)), "rb_commit")
        self.pending = 0

    def load_device(self, hot_rows):
        """hot_rows: CUDA float32 tensor [n, stride] already in the hot-row layout."""
        assert hot_rows.is_cuda and hot_rows.dtype == torch.float32 and hot_rows.is_contiguous()
        assert hot_rows.shape[1] == self.stride
        self.commit()
        check(self.lib.ilsw_rb_load_device(self.h, _ptr(hot_rows), hot_rows.shape[0], _stream_ptr()), "rb_load_device")

    def gather(self, idx):
        """idx: int32 CUDA tensor [B] -> (hot [B,stride], cold [B,4]) CUDA tensors."""
        self.pending = 0
        B = idx.numel()
        hot = torch.empty((B, self.stride), dtype=torch.float32, device=idx.device)
        cold = torch.empty((B, 4), dtype=torch.float32, device=idx.device)
        with _Range("ilsw_rb_gather"):
            check(self.lib.ilsw_rb_gather(self.h, _ptr(idx), B, _ptr(hot), _ptr(cold), _stream_ptr()), "rb_gather")
        return hot, cold

    def sample(self, batch_size, seed, counter):
        """In-kernel Philox uniform sampling (with replacement) + gather."""
        self.pending = 0
        idx = torch.empty((batch_size,), dtype=torch.int32, device="cuda")
        hot = torch.empty((batch_size, self.stride), dtype=torch.float32, device="cuda")
        check(self.lib.ilsw_rb_sample(self.h, batch_size, seed, counter, _ptr(idx), _ptr(hot), _stream_ptr()), "rb_sample")
        return idx, hot

    def clear(self):
        check(self.lib.ilsw_rb_clear(self.h), "rb_clear")
        self.pending = 0

    def set_cursor(self, top, size):
        check(self.lib.ilsw_rb_set_cursor(self.h, int(top), int(size)), "rb_set_cursor")

    def rows_view(self):
        """Zero-copy torch view of the whole ring (debug / snapshots)."""
        ptr = self.lib.ilsw_rb_rows_ptr(self.h)
        n = self.capacity * self.stride

        class _Iface:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Iface(), device="cuda").view(self.capacity, self.stride)


class NetArena:
    """Flat fp32 parameter arena (+ Adam moments) of one 2-hidden-layer MLP on the device."""

    def __init__(self, in_dim, hidden, out_dim, log_std_head=False, trainable=True, init=None):
        self.in_dim, self.hidden, self.out_dim, self.log_std_head = in_dim, hidden, out_dim, bool(log_std_head)
        self.n = layout.mlp_num_params(in_dim, hidden, out_dim, log_std_head)
        self.p = torch.zeros(self.n, dtype=torch.float32, device="cuda")
        self.m = torch.zeros(self.n, dtype=torch.float32, device="cuda") if trainable else None
        self.v = torch.zeros(self.n, dtype=torch.float32, device="cuda") if trainable else None
        if init is not None:
            self.p.copy_(torch.as_tensor(np.asarray(init, dtype=np.float32)))

    def desc(self):
        return _abi.Mlp(_ptr(self.p), _ptr(self.m), _ptr(self.v), self.in_dim, self.hidden, self.out_dim,
                        int(self.log_std_head))

    def shapes(self):
        return layout.mlp_param_shapes(self.in_dim, self.hidden, self.out_dim, self.log_std_head)

    def views(self, which="p"):
        """Per-parameter views in nn.Module.parameters() order."""
        base = getattr(self, which)
        out, off = [], 0
        for shp in self.shapes():
            k = int(np.prod(shp))
            out.append(base[off:off + k].view(*shp))
            off += k
        return out


class StepEngine:
    """ilsw_trainer_*: ONE persistent kernel launch per train() call."""

    def __init__(self, cfg, nets, disc_cfg=None, disc=None):
        require_cuda()
        self.lib = load()
        self.cfg, self.nets, self.disc = cfg, list(nets), disc
        arr = (_abi.Mlp * len(nets))(*[n.desc() for n in nets])
        h = C.c_void_p()
        check(self.lib.ilsw_trainer_create(C.byref(h), C.byref(cfg), arr, len(nets)), "trainer_create")
        self.h = h
        if disc_cfg is not None:
            self.disc_cfg = disc_cfg
            d = disc.desc()
            check(self.lib.ilsw_trainer_attach_disc(h, C.byref(disc_cfg), C.byref(d)), "attach_disc")
        self.last_steps = 0

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.ilsw_trainer_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def train(self, ring, n_steps, expert_ring=None, inject=None, batch=None, seed=0, stats_step=-1):
        """inject: dict of CUDA tensors (idx int32 [T,B], eps_next/eps_cur float32 [T,B,A], ...)
        batch: dict of CUDA float32 tensors obs/act/rew/term/next_obs (n_steps must be 1)."""
        ij = None
        if inject is not None:
            self._keep = inject
            ij = _abi.Inject(_ptr(inject.get("idx")), _ptr(inject.get("eps_next")), _ptr(inject.get("eps_cur")),
                             _ptr(inject.get("idx_expert")), _ptr(inject.get("idx_policy_d")), _ptr(inject.get("gp_eps")))
        bt = None
        if batch is not None:
            self._keepb = batch
            bt = _abi.Batch(_ptr(batch["obs"]), _ptr(batch["act"]), _ptr(batch["rew"]), _ptr(batch["term"]),
                            _ptr(batch["next_obs"]))
        if ring is not None:
            ring.pending = 0
        if expert_ring is not None:
            expert_ring.pending = 0
        with _Range("ilsw_train"):
            check(self.lib.ilsw_train(self.h, ring.h if ring is not None else None,
                                      expert_ring.h if expert_ring is not None else None, n_steps,
                                      C.byref(ij) if ij is not None else None, C.byref(bt) if bt is not None else None,
                                      C.c_uint64(seed), stats_step, _stream_ptr()), "train")
        self.last_steps = n_steps

    def set_her(self, desc=None):
        """Hindsight relabel-at-sample for the next train() calls (desc: _abi.HerSamplingDesc; None switches it off)."""
        self._her_keep = desc
        check(self.lib.ilsw_trainer_set_her(self.h, C.byref(desc) if desc is not None else None), "set_her")

    def set_update_mode(self, mode):
        """AdvIRL engines: 0 = fused disc+policy iteration, 1 = disc-only steps, 2 = policy-only steps."""
        check(self.lib.ilsw_trainer_set_update_mode(self.h, int(mode)), "set_update_mode")

    def losses(self, n_steps=None):
        n = n_steps or self.last_steps
        out = np.empty((n, _abi.LOSS_SLOTS), dtype=np.float32)
        check(self.lib.ilsw_read_losses(self.h, out.ctypes.data_as(C.c_void_p), n, _stream_ptr()), "read_losses")
        return out

    # -- asynchronous loss read-back: the D2H copy of a launch's loss log is queued behind the launch into a
    # pinned ring; the host only synchronises when it collects (per-step API use without a host sync per step)
    _loss_ring = None

    def losses_async(self, n_steps=None):
        n = n_steps or self.last_steps
        if self._loss_ring is None:
            self._loss_ring = torch.empty((max(4096, self.cfg.max_steps_per_call), _abi.LOSS_SLOTS), dtype=torch.float32).pin_memory()
            self._loss_np = self._loss_ring.numpy()
            self._loss_cur = 0
            self._loss_done = []
        if self._loss_cur + n > self._loss_ring.shape[0]:
            self._drain_losses()
        c = self._loss_cur
        check(self.lib.ilsw_read_losses_async(self.h, C.c_void_p(self._loss_ring.data_ptr() + c * _abi.LOSS_SLOTS * 4), n,
                                              _stream_ptr()), "read_losses_async")
        self._loss_cur = c + n

    def _drain_losses(self):
        torch.cuda.current_stream().synchronize()
        if self._loss_cur:
            self._loss_done.append(self._loss_np[:self._loss_cur].copy())
        self._loss_cur = 0

    def losses_collect(self):
        """All losses queued with losses_async() since the last collect: [n_total_steps, LOSS_SLOTS]."""
        if self._loss_ring is None:
            return np.zeros((0, _abi.LOSS_SLOTS), dtype=np.float32)
        self._drain_losses()
        check(self.lib.ilsw_check_abort(self.h, _stream_ptr()), "check_abort")
        out = np.concatenate(self._loss_done) if self._loss_done else np.zeros((0, _abi.LOSS_SLOTS), dtype=np.float32)
        self._loss_done = []
        return out

    def stats(self):
        n = self.lib.ilsw_stats_floats(self.h)
        out = np.empty((n,), dtype=np.float32)
        check(self.lib.ilsw_read_stats(self.h, out.ctypes.data_as(C.c_void_p), n, _stream_ptr()), "read_stats")
        return out

    def get_state(self):
        st = _abi.State()
        check(self.lib.ilsw_get_state(self.h, C.byref(st), _stream_ptr()), "get_state")
        return st

    def set_state(self, st):
        check(self.lib.ilsw_set_state(self.h, C.byref(st), _stream_ptr()), "set_state")

    def uses_tc5(self):
        """True when the dense GEMM phases of this program run on the tcgen05/TMA tile (batch >= 512)."""
        return bool(self.lib.ilsw_trainer_uses_tc5(self.h))

    def describe(self):
        buf = C.create_string_buffer(1 << 15)
        check(self.lib.ilsw_describe_program(self.h, buf, len(buf)), "describe_program")
        return buf.value.decode()

    def phase_times_us(self):
        """Per-phase durations (us, barrier included) of the last step of the last launch."""
        n = self.num_phases + 1
        out = np.zeros(2 * 97, dtype=np.uint64)
        check(self.lib.ilsw_read_phase_ns(self.h, out.ctypes.data_as(C.c_void_p), 2 * 97, _stream_ptr()), "read_phase_ns")
        t = out.astype(np.int64)
        self.last_job_us = (t[97:97 + n - 1] - t[:n - 1]) / 1000.0     # CTA 0: phase start -> its jobs done
        return np.diff(t[:n]) / 1000.0

    @property
    def num_phases(self):
        return self.lib.ilsw_num_phases(self.h)

    @property
    def kernel_launches(self):
        return int(self.lib.ilsw_kernel_launches(self.h))

    def policy_act(self, obs, deterministic=False, seed=0):
        obs = obs.contiguous()
        out = torch.empty((obs.shape[0], self.cfg.act_dim), dtype=torch.float32, device=obs.device)
        check(self.lib.ilsw_policy_act(self.h, _ptr(obs), obs.shape[0], int(deterministic), C.c_uint64(seed), _ptr(out),
                                       _stream_ptr()), "policy_act")
        return out

    def policy_act_host(self, obs_np, deterministic=False, seed=0):
        """get_actions with host buffers: numpy [n,O] -> numpy [n,A]; one kernel, one host sync."""
        obs = np.ascontiguousarray(obs_np, dtype=np.float32)
        if obs.ndim != 2 or obs.shape[1] != self.cfg.obs_dim:
            raise ValueError("observations must be [n, %d], got %s" % (self.cfg.obs_dim, obs.shape))
        out = np.empty((obs.shape[0], self.cfg.act_dim), dtype=np.float32)
        check(self.lib.ilsw_policy_act_host(self.h, obs.ctypes.data_as(C.c_void_p), obs.shape[0], int(deterministic),
                                            C.c_uint64(seed), out.ctypes.data_as(C.c_void_p), _stream_ptr()), "policy_act_host")
        return out

    def replica_export(self):
        buf = C.create_string_buffer(_abi.IPC_HANDLE_BYTES)
        check(self.lib.ilsw_replica_export(self.h, buf), "replica_export")
        return buf.raw

    def replica_buffer_bytes(self):
        n = int(self.lib.ilsw_replica_buffer_bytes(self.h))
        if n <= 0:
            raise IlswError("replica_buffer_bytes failed")
        return n

    def replica_connect_symm(self, rank, world, peer_ptrs, multicast_ptr, nbytes, keepalive=None):
        """Exchange over a symmetric buffer (same allocation on every rank, mapped here at peer_ptrs[r]); multicast_ptr != 0
        switches the gradient push to NVLS multicast stores.  `keepalive` (the tensor / handle owning the mapping) is held."""
        arr = (C.c_uint64 * world)(*[int(p) for p in peer_ptrs])
        check(self.lib.ilsw_replica_connect_symm(self.h, rank, world, arr, C.c_uint64(int(multicast_ptr)), C.c_int64(int(nbytes))),
              "replica_connect_symm")
        self._symm_keepalive = keepalive
        self.replica_multicast = bool(multicast_ptr)

    def replica_connect(self, rank, world, handles):
        blob = b"".join(handles)
        assert len(blob) == world * _abi.IPC_HANDLE_BYTES
        check(self.lib.ilsw_replica_connect(self.h, rank, world, blob), "replica_connect")
