"""One very large crowd (tens of thousands of humans in ONE environment), optionally sharded by agent across GPUs.

Every rank owns a contiguous slice of agents and evaluates their forces against ALL agents (ordered pairs: the reference's
symmetric accumulation, forces.py:148, is not carried across ranks).  The only exchange per sub-step is an all-gather of the
entity view (x, y, vx, vy, r+safety) over NCCL / NVLink; it is double buffered so the step kernel writes the next view's own
slice while reading the current one.  Reference semantics: motion_model_manager.py:354-373 for a single env.
"""
import ctypes

import numpy as np
import torch

from . import _lib as L
from .engine import CrowdEngine, SFMS
from .parallel import all_gather_columns, agent_shard


class LargeCrowd:
    def __init__(self, model, states, goals, walls=None, dtype=torch.float64, device="cuda", symmetric=True, numba_compat=False,
                 rank=0, world=1, group=None, safety=None, exchange="auto", shard=None):
        """states [N_total,13], goals [N_total,G,2] (the WHOLE crowd on every rank; each rank keeps its slice).
        shard=(offset, n_local) with world=1 steps only that slice against a frozen rest of the crowd (profiling what ONE rank of
        a sharded run executes, on one GPU)."""
        states = np.asarray(states, np.float64)
        goals = np.asarray(goals, np.float64)
        self.n_total = states.shape[0]
        self.rank, self.world, self.group = rank, world, group
        self.offset, self.n_local = agent_shard(self.n_total, rank, world) if shard is None else shard
        sl = slice(self.offset, self.offset + self.n_local)
        saf = None if safety is None else np.asarray(safety, np.float64)[None, sl]
        self.eng = CrowdEngine.from_reference_arrays(model, states[None, sl], goals[None, sl], walls=walls, safety=saf, consider_robot=False,
                                                     all_params_equal=symmetric, numba_compat=numba_compat, dtype=dtype, device=device)
        self.type = SFMS.index(model)
        # exchange of the entity view between ranks: "p2p" = the producer kernel stores into every rank's next view through
        # peer-mapped (torch symmetric memory, NVLink) pointers and only a barrier separates sub-steps; "nccl" = all-gather
        # after every sub-step.  "auto" tries p2p on multi-GPU runs and falls back to nccl.
        self.exchange, self._symm = "nccl", None
        # a view buffer = the [5, N] entity view followed by the [N/128, 5] bounding boxes of its 128-entity tiles (written by the
        # producer together with the entries, see snp_large_run_p2p)
        self._n_tiles = (self.n_total + 127) // 128
        vlen = 5 * self.n_total + 5 * self._n_tiles
        fused_ok = self.n_local % 128 == 0
        if world > 1 and exchange in ("auto", "p2p") and fused_ok:
            try:
                import torch.distributed as dist
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else dist.group.WORLD
                self._vbuf = [symm_mem.empty((vlen,), dtype=dtype, device=self.eng.device) for _ in range(2)]
                self._flags = symm_mem.empty((16,), dtype=torch.int64, device=self.eng.device)
                for v in self._vbuf + [self._flags]:
                    v.zero_()
                self._symm = [symm_mem.rendezvous(v, grp) for v in self._vbuf]
                self._symm_flags = symm_mem.rendezvous(self._flags, grp)
                self._peer_ptrs = [(ctypes.c_void_p * world)(*[int(p) for p in h.buffer_ptrs]) for h in self._symm]
                self._peer_flags = (ctypes.c_void_p * world)(*[int(p) for p in self._symm_flags.buffer_ptrs])
                torch.cuda.synchronize()
                self._symm[0].barrier()
                self.exchange = "p2p"
            except Exception as exc:  # no peer access / symmetric memory unavailable
                if exchange == "p2p":
                    raise
                self._p2p_error = repr(exc)
        elif world == 1 and fused_ok and exchange != "nccl":
            # one GPU: the same fused run (sub-step loop inside one C call, tile boxes written by the producer), no barrier
            self._vbuf = [torch.zeros((vlen,), dtype=dtype, device=self.eng.device) for _ in range(2)]
            self._flags = torch.zeros((16,), dtype=torch.int64, device=self.eng.device)
            self._peer_ptrs = [(ctypes.c_void_p * 1)(v.data_ptr()) for v in self._vbuf]
            self._peer_flags = (ctypes.c_void_p * 1)(self._flags.data_ptr())
            self.exchange = "fused"
        if self.exchange == "nccl":
            self._vbuf = [torch.zeros((vlen,), dtype=dtype, device=self.eng.device) for _ in range(2)]
        self.view = [v[:5 * self.n_total].view(5, self.n_total) for v in self._vbuf]
        self._epoch = 0
        self._err = torch.zeros((1,), dtype=torch.int32, device=self.eng.device)
        nbytes = int(self.eng.lib.snp_large_scratch_bytes(self.n_local, self.n_total, L.SNP_F64 if dtype == torch.float64 else L.SNP_F32))
        self.scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.eng.device)
        self.culling = True
        self.work_list = True
        self.cur = 0
        if shard is not None:  # the frozen rest of the crowd: both view buffers start from the whole crowd's entries
            whole = CrowdEngine.from_reference_arrays(model, states[None], goals[None], consider_robot=False, all_params_equal=symmetric, dtype=dtype, device=device)
            for v in self.view:
                L.check(whole.lib.snp_large_publish(ctypes.byref(whole._crowd()), self.type, ctypes.c_void_p(v.data_ptr()), self.n_total, 0, self._stream()))
        self._publish()

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _gather(self, buf):
        all_gather_columns(buf, self.offset, self.n_local, self.world, self.group)
        if self.exchange == "p2p":
            self._symm[0].barrier()

    def check_peers(self):
        """Raises if a rank never reached a sub-step barrier (the barrier kernel timed out instead of hanging the GPU)."""
        if int(self._err.item()):
            raise RuntimeError("LargeCrowd: a peer rank did not reach the sub-step barrier")

    def _publish(self):
        v = self.view[self.cur]
        L.check(self.eng.lib.snp_large_publish(ctypes.byref(self.eng._crowd()), self.type, ctypes.c_void_p(v.data_ptr()), self.n_total,
                                               self.offset, self._stream()))
        self._gather(v)

    def step(self, dt=0.0125, n_substeps=1):
        c = self.eng._crowd()
        o = self.eng._opts(dt, 1)
        o.reserved = 0 if self.culling else L.SNP_OPT_NO_CULLING
        if not self.work_list:
            o.reserved |= L.SNP_OPT_LARGE_GRID  # culled steps on the static grid instead of the list of near (i-block, chunk) pairs
        if self.exchange in ("p2p", "fused") and not getattr(self, "legacy_loop", False):
            # the whole sub-step loop in ONE C call: per sub-step two launches + a single-warp cross-rank barrier kernel, nothing
            # returns to Python in between
            L.check(self.eng.lib.snp_large_run_p2p(ctypes.byref(c), ctypes.byref(o), self._peer_ptrs[0], self._peer_ptrs[1], self.cur, self.n_total,
                                                   self.offset, self.world, self.rank, self._peer_flags, self._epoch, int(n_substeps),
                                                   ctypes.c_void_p(self._err.data_ptr()), ctypes.c_void_p(self.scratch.data_ptr()),
                                                   self.scratch.numel(), self._stream()))
            self._epoch += int(n_substeps)
            self.cur ^= int(n_substeps) & 1
            return
        for _ in range(n_substeps):
            cur, nxt = self.view[self.cur], self.view[self.cur ^ 1]
            if self.exchange == "p2p":
                L.check(self.eng.lib.snp_large_step_p2p(ctypes.byref(c), ctypes.byref(o), ctypes.c_void_p(cur.data_ptr()), self.n_total, self.offset,
                                                        self._peer_ptrs[self.cur ^ 1], self.world, ctypes.c_void_p(self.scratch.data_ptr()),
                                                        self.scratch.numel(), self._stream()))
                self._symm[self.cur ^ 1].barrier()  # every rank's stores into everybody's next view have landed
            else:
                L.check(self.eng.lib.snp_large_step(ctypes.byref(c), ctypes.byref(o), ctypes.c_void_p(cur.data_ptr()), self.n_total, self.offset,
                                                    ctypes.c_void_p(nxt.data_ptr()), ctypes.c_void_p(self.scratch.data_ptr()), self.scratch.numel(),
                                                    self._stream()))
                self._gather(nxt)
            self.cur ^= 1

    def local_rows(self, template):
        """This rank's agents as reference rows; `template` [n_local,13] supplies the static columns."""
        return self.eng.rows(np.asarray(template, np.float64)[None])[0]
