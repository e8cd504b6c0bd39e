"""
oracle/ - CPU restatement of the swiftLLM data-plane algorithms.

TEST INFRASTRUCTURE ONLY.  Nothing under `swiftllm_b200/` may import this
package.  The only legal importers are `tests/`, `__graft_entry__.smoke()` and
`bench.py` (its `cpu_baseline` leg and `--impl reference`).

Parity pinning: the reference ships no golden vectors or tests,
so the oracle is pinned against outputs of the *reference itself*, executed in
the build container on CPU (its Triton kernels under TRITON_INTERPRET=1, its
`LlamaModel` under a cuda->cpu shim).  `oracle/gen_golden.py` is the committed
generator, `tests/golden/*.npz` are its outputs, and
`tests/test_oracle_golden.py` checks every oracle function against them.
Third-party arithmetic the reference delegates to (vllm_flash_attn prefill,
cuBLAS GEMM, torch.argmax) has no in-repo pin: "parity unpinned" for those
(recorded in DESIGN.md); versions used for the fixtures: torch 2.11.0+cu128,
triton 3.6.0 (interpreter), numpy 2.3.5.
"""
