"""Tensor-parallel exchange over NVLink peer memory, fused with the residual add + RMSNorm that follows it
(SURVEY.md §8 f-3).  torch.distributed symmetric memory is the plumbing (buffer allocation + handle exchange); the
data movement and the reduction are this library's own kernel (`sllm_allreduce_add_rmsnorm`, csrc/allreduce_norm.cu).

Two symmetric buffers alternate: slot 0 holds the o_proj partials, slot 1 the down_proj partials of a layer.  The GEMM
writes its partial straight into the symmetric buffer (`torch.mm(..., out=)`), the fused kernel of every rank then reads
all ranks' partials in rank order (bit-identical sums everywhere) and produces the normalised activations for the next
GEMM.  NCCL all-reduce + fused_add_rmsnorm (the default path) computes the same thing in two launches.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from swiftllm_b200 import _lib

NUM_SLOTS = 2
MAX_LL_TOKENS = 1024        # exchanges of at most this many token rows take the barrier-free LL kernel (ll=True)
# Row-owner (two-shot / NVLS / LL) exchanges are used for steps of at most this many token rows - the range they were run and
# measured in on 8 GPUs (decode batches of 256 and 1024, 576-row mixed steps); larger steps (whole-prompt prefill) take the NCCL
# all-reduce + add/norm path, where the exchange is bandwidth-bound and a fused kernel has nothing to add.
MAX_ROW_OWNER_TOKENS = 1024


class FusedAllReduce:
    """`two_shot=False`: every rank reduces every row itself (one barrier; (N-1)*T*H remote bytes per rank; measured best for
    TP 2..4).  `two_shot=True`: row t is reduced, added to the residual and normalised by rank t % N only, which then pushes
    the normalised row to every rank (two barriers; 2*(N-1)/N*T*H remote bytes per rank) - meant for TP 8.  In that mode the
    residual is only kept up to date for the rows a rank owns and `reduce_add_norm` returns a view of a symmetric buffer that
    is overwritten by the next exchange."""

    def __init__(self, max_tokens: int, hidden: int, dtype: torch.dtype, device, group=None, two_shot: bool = False,
                 nvls: bool = False, ll: bool = False):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert 2 <= self.world <= 8
        self.hidden, self.dtype, self.device = hidden, dtype, device
        self.ll = bool(ll)              # decode-sized exchanges: barrier-free push kernel (csrc/allreduce_ll.cu); row ownership,
        #                                 residual sharding and the returned buffer are those of the two-shot kernel, which
        #                                 also serves the exchanges that are too large for the LL receive buffers
        self.two_shot = bool(two_shot) or bool(nvls) or self.ll
        self.nvls = bool(nvls)          # two-shot with in-switch reduction / broadcast (multimem.ld_reduce / multimem.st)
        # steps with more token rows than this do not go through this object (the layers fall back to NCCL + add/norm)
        self.max_fused_tokens = min(max_tokens, MAX_ROW_OWNER_TOKENS) if self.two_shot else max_tokens
        self.max_tokens = max_tokens = self.max_fused_tokens
        self.data = symm_mem.empty((NUM_SLOTS, max_tokens, hidden), dtype=dtype, device=device)
        self.flags = symm_mem.empty((16 * 8,), dtype=torch.int32, device=device)
        self.data.zero_(); self.flags.zero_()
        self._hd = symm_mem.rendezvous(self.data, group)
        self._hf = symm_mem.rendezvous(self.flags, group)
        if self.two_shot:
            self.xout = symm_mem.empty((max_tokens, hidden), dtype=dtype, device=device)
            self.xout.zero_()
            self._hx = symm_mem.rendezvous(self.xout, group)
        if self.ll:
            self.max_ll_tokens = min(max_tokens, MAX_LL_TOKENS)
            self.ll_rows_per_rank = (self.max_ll_tokens + self.world - 1) // self.world
            # LL format: 4 bytes per element ({payload, tag} pairs) = 2 * hidden elements of `dtype` per row
            self.rs_recv = symm_mem.empty((NUM_SLOTS, self.world, self.ll_rows_per_rank, 2 * hidden), dtype=dtype, device=device)
            self.ag_recv = symm_mem.empty((NUM_SLOTS, self.max_ll_tokens, 2 * hidden), dtype=dtype, device=device)
            self.rs_recv.zero_(); self.ag_recv.zero_()                 # tag 0 = never written (epochs start at 1)
            self._hrs = symm_mem.rendezvous(self.rs_recv, group)
            self._hag = symm_mem.rendezvous(self.ag_recv, group)
        torch.cuda.synchronize(); dist.barrier(group)                  # every rank's flags are zero before first use
        # the GEMMs write through self.data / self.xout, the kernels read every rank's copy (this rank's included) through the
        # handle's peer pointers: both must be the same memory
        for name, t, h in (("data", self.data, self._hd), ("flags", self.flags, self._hf)) + \
                ((("xout", self.xout, self._hx),) if self.two_shot else ()) + \
                ((("rs_recv", self.rs_recv, self._hrs), ("ag_recv", self.ag_recv, self._hag)) if self.ll else ()):
            if int(h.buffer_ptrs[self.rank]) != t.data_ptr():
                raise RuntimeError(f"symmetric buffer `{name}`: the rendezvous handle maps this rank's copy at "
                                   f"{int(h.buffer_ptrs[self.rank]):#x} but the tensor lives at {t.data_ptr():#x} "
                                   f"(sub-allocated symmetric memory is not supported by the fused exchange)")
        slot_bytes = max_tokens * hidden * self.data.element_size()
        PtrArr = ctypes.c_void_p * self.world
        self._buf_ptrs = [PtrArr(*[int(p) + s * slot_bytes for p in self._hd.buffer_ptrs]) for s in range(NUM_SLOTS)]
        self._flag_ptrs = PtrArr(*[int(p) for p in self._hf.buffer_ptrs])
        self._xout_ptrs = PtrArr(*[int(p) for p in self._hx.buffer_ptrs]) if self.two_shot else None
        self._mc_buf = [0] * NUM_SLOTS
        self._mc_xout = 0
        if self.nvls:
            mc_d, mc_x = int(self._hd.multicast_ptr or 0), int(self._hx.multicast_ptr or 0)
            if mc_d == 0 or mc_x == 0:
                raise RuntimeError("NVLS multicast addresses unavailable for the symmetric buffers (no multicast support)")
            self._mc_buf = [mc_d + s * slot_bytes for s in range(NUM_SLOTS)]
            self._mc_xout = mc_x
        if self.ll:
            rs_slot = self.world * self.ll_rows_per_rank * 2 * hidden * self.data.element_size()
            ag_slot = self.max_ll_tokens * 2 * hidden * self.data.element_size()
            self._rs_ptrs = [PtrArr(*[int(p) + s * rs_slot for p in self._hrs.buffer_ptrs]) for s in range(NUM_SLOTS)]
            self._ag_ptrs = [PtrArr(*[int(p) + s * ag_slot for p in self._hag.buffer_ptrs]) for s in range(NUM_SLOTS)]
            self._mc_ag = [0] * NUM_SLOTS
            if self.nvls:
                mc = int(self._hag.multicast_ptr or 0)
                if mc == 0:
                    raise RuntimeError("NVLS multicast address unavailable for the LL receive buffer (no multicast support)")
                self._mc_ag = [mc + s * ag_slot for s in range(NUM_SLOTS)]
        self.epoch = torch.zeros((32,), dtype=torch.int32, device=device)

    def partial_out(self, slot: int, num_tokens: int) -> torch.Tensor:
        """The [num_tokens, hidden] view of this rank's symmetric buffer that the row-parallel GEMM must write into."""
        assert num_tokens <= self.max_tokens
        return self.data[slot, :num_tokens]

    def reduce_add_norm(self, slot: int, num_tokens: int, residual: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
        """residual <- h(sum_ranks partial[slot]) + residual;  returns rmsnorm(residual) * weight (new tensor)."""
        assert residual.is_contiguous() and residual.shape == (num_tokens, self.hidden)
        if self.ll and num_tokens <= self.max_ll_tokens:
            assert weight is not None
            _lib.check(_lib.lib().sllm_allreduce_add_rmsnorm_ll(
                self.data[slot].data_ptr(), ctypes.cast(self._rs_ptrs[slot], ctypes.c_void_p),
                ctypes.cast(self._ag_ptrs[slot], ctypes.c_void_p), self._mc_ag[slot], self.rank, self.world, slot,
                self.epoch.data_ptr(), self.xout.data_ptr(), residual.data_ptr(), weight.data_ptr(), eps, num_tokens, self.hidden,
                self.ll_rows_per_rank, _lib.dtype_tag(self.dtype), _lib.stream()), "allreduce_add_rmsnorm_ll")
            return self.xout[:num_tokens]
        if self.two_shot:
            assert weight is not None and num_tokens <= self.max_tokens
            _lib.check(_lib.lib().sllm_allreduce_add_rmsnorm_2shot(
                ctypes.cast(self._buf_ptrs[slot], ctypes.c_void_p), ctypes.cast(self._xout_ptrs, ctypes.c_void_p),
                ctypes.cast(self._flag_ptrs, ctypes.c_void_p), self._mc_buf[slot], self._mc_xout, self.rank, self.world, slot,
                self.epoch.data_ptr(),
                residual.data_ptr(), weight.data_ptr(), eps, num_tokens, self.hidden, _lib.dtype_tag(self.dtype),
                _lib.stream()), "allreduce_add_rmsnorm_2shot")
            return self.xout[:num_tokens]
        out = torch.empty_like(residual)
        _lib.check(_lib.lib().sllm_allreduce_add_rmsnorm(
            ctypes.cast(self._buf_ptrs[slot], ctypes.c_void_p), ctypes.cast(self._flag_ptrs, ctypes.c_void_p),
            self.rank, self.world, slot, self.epoch.data_ptr(), out.data_ptr(), residual.data_ptr(), weight.data_ptr(),
            eps, num_tokens, self.hidden, _lib.dtype_tag(self.dtype), _lib.stream()), "allreduce_add_rmsnorm")
        return out
