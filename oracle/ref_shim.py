"""
Import shim that lets the UNMODIFIED reference (`/root/reference`, swiftLLM) execute on CPU in
the build container.  TEST INFRASTRUCTURE ONLY; used by `oracle/gen_golden.py` to produce the
fixtures in `tests/golden/`.  It never runs on the GPU box (`/root/reference` is absent there).

What it does:
  * TRITON_INTERPRET=1 so every `@triton.jit` kernel runs in Triton's numpy interpreter;
  * stub modules for imports that are not installable here: `ray` (tokenizer actor only),
    `swiftllm_c` (swap_blocks; restated with torch copies below), `vllm_flash_attn`
    (third-party prefill attention -> an fp32 torch stand-in, "parity unpinned");
  * a TorchFunctionMode that rewrites device="cuda" -> "cpu" and fake CUDA streams/events so
    `LlamaModel` runs end to end on the host.
No reference source is copied; it is imported from where it lies.
"""
from __future__ import annotations

import os
import sys
import types
import contextlib

REFERENCE_ROOT = os.environ.get("SWIFTLLM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "swiftllm"))


def _install_stubs():
    import torch

    if "ray" not in sys.modules:
        ray = types.ModuleType("ray")
        ray.remote = lambda cls: cls
        ray.init = lambda *a, **k: None
        sys.modules["ray"] = ray

    if "swiftllm_c" not in sys.modules:
        m = types.ModuleType("swiftllm_c")

        def swap_blocks(src_ids, dst_ids, is_swap_in, k_cache, v_cache, k_swap, v_swap):
            # semantics of csrc/src/block_swapping.cpp:22-85 (whole-block copies)
            for s, d in zip(src_ids, dst_ids):
                if is_swap_in:
                    k_cache[d].copy_(k_swap[s]); v_cache[d].copy_(v_swap[s])
                else:
                    k_swap[d].copy_(k_cache[s]); v_swap[d].copy_(v_cache[s])
        m.swap_blocks = swap_blocks
        sys.modules["swiftllm_c"] = m

    if "vllm_flash_attn" not in sys.modules:
        m = types.ModuleType("vllm_flash_attn")

        def flash_attn_varlen_func(q, k, v, cu_q, cu_k, max_q, max_k, softmax_scale=None, causal=False, **kw):
            # fp32 stand-in for the third-party kernel (transformer_layer.py:86-96).  The last
            # cu_seqlen is clamped to the sliced q length.
            T, nq, D = q.shape
            nkv = k.shape[1]
            g = nq // nkv
            out = torch.zeros_like(q)
            cu = [min(int(x), T) for x in cu_q.tolist()]
            for b in range(len(cu) - 1):
                s0, s1 = cu[b], cu[b + 1]
                if s1 <= s0:
                    continue
                Q = q[s0:s1].float().transpose(0, 1)
                K = k[s0:s1].float().transpose(0, 1).repeat_interleave(g, 0)
                V = v[s0:s1].float().transpose(0, 1).repeat_interleave(g, 0)
                s = torch.einsum("hqd,hkd->hqk", Q, K) * softmax_scale
                L = s1 - s0
                mask = torch.tril(torch.ones(L, L, dtype=torch.bool))
                s = s.masked_fill(~mask, float("-inf"))
                p = torch.softmax(s, dim=-1)
                out[s0:s1] = torch.einsum("hqk,hkd->hqd", p, V).transpose(0, 1).to(q.dtype)
            return out
        m.flash_attn_varlen_func = flash_attn_varlen_func
        sys.modules["vllm_flash_attn"] = m


def import_reference():
    """Returns the imported `swiftllm` package of the reference."""
    assert reference_available(), f"reference not found at {REFERENCE_ROOT}"
    os.environ["TRITON_INTERPRET"] = "1"
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import swiftllm  # noqa
    return swiftllm


@contextlib.contextmanager
def cuda_as_cpu():
    """Run reference host code that says device="cuda" on the CPU."""
    import torch
    from torch.overrides import TorchFunctionMode

    class _Mode(TorchFunctionMode):
        def __torch_function__(self, func, types_, args=(), kwargs=None):
            kwargs = dict(kwargs or {})
            dev = kwargs.get("device", None)
            if dev is not None and str(dev).startswith("cuda"):
                kwargs["device"] = "cpu"
            return func(*args, **kwargs)

    class _FakeStream:
        def wait_event(self, e): pass
        def wait_stream(self, s): pass
        def synchronize(self): pass

    class _FakeEvent:
        def __init__(self, *a, **k): pass
        def record(self, *a, **k): pass
        def wait(self, *a, **k): pass
        def synchronize(self): pass

    saved = {n: getattr(torch.cuda, n) for n in
             ("Stream", "Event", "stream", "current_stream", "default_stream", "synchronize",
              "empty_cache", "get_device_name")}
    torch.cuda.Stream = lambda *a, **k: _FakeStream()
    torch.cuda.Event = _FakeEvent
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.current_stream = lambda *a, **k: _FakeStream()
    torch.cuda.default_stream = lambda *a, **k: _FakeStream()
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.empty_cache = lambda: None
    torch.cuda.get_device_name = lambda *a, **k: "CPU (interpreter)"
    try:
        with _Mode():
            yield
    finally:
        for n, f in saved.items():
            setattr(torch.cuda, n, f)
