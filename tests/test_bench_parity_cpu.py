"""CPU: bench.py's `parity_check` (the oracle check bench.py runs at the benchmarked shape, outside the timed region) executed
against the product's host path with oracle-backed kernels (tests/cpu_shim.py): it must pass on a correct model and FAIL when
the attention output or the sampled tokens are corrupted - i.e. the check in the bench line is a real check."""
import importlib
import os
import sys

import pytest
import torch

from oracle.model import OracleWeights
from cpu_shim import product_on_cpu
from test_host_path_cpu import CFG, _product

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _setup():
    bench = importlib.import_module("bench")
    w = OracleWeights.random(CFG, dtype=torch.float16, seed=3, std=0.05)
    m = _product(w)
    g = torch.Generator().manual_seed(5)
    m.k_cache.copy_(torch.randn(m.k_cache.shape, generator=g).to(m.k_cache.dtype))
    m.v_cache.copy_(torch.randn(m.v_cache.shape, generator=g).to(m.v_cache.dtype))
    sids = [0, 1, 2, 3]
    lens = [80, 33, 96, 17]
    ids = [[5], [6], [7], [8]]
    return bench, m, ids, sids, lens


@torch.inference_mode()
def test_parity_check_passes_and_detects_corruption():
    with product_on_cpu():
        bench, m, ids, sids, lens = _setup()
        res = bench.parity_check(m, m.model_config, ids, sids, lens, n_rows=4, n_seqs=2)
        assert res["ok"] and len(res["attention_rows"]) == 4 * len({0, 1})        # 2-layer model: layers {0, 1}
        assert all(s["token_equal"] or s["exact_top1_margin_rel"] <= 2 * s["product_logit_err_vs_exact"] for s in res["sequences"])
        # on CPU the "product" runs the oracle's own kernels: it must sit exactly on the storage-dtype oracle
        assert all(s["product_vs_storage_dtype_oracle"] <= 1e-3 for s in res["sequences"])
        assert all(s["product_logit_err_vs_exact"] <= 2 * s["storage_dtype_oracle_logit_err_vs_exact"] + 1e-3 for s in res["sequences"])
        assert res["attention_worst_rel_err"] <= 2e-3

        # corrupt the attention output of the product: the check must fail
        from swiftllm_b200.worker.layers import transformer_layer as TL
        good = TL.paged_attention

        def bad(q, kc, vc, bt, mc, ec, st, layer, o):
            good(q, kc, vc, bt, mc, ec, st, layer, o)
            o.mul_(1.05)
        TL.paged_attention = bad
        try:
            with pytest.raises(AssertionError, match="paged attention at the benchmarked shape"):
                bench.parity_check(m, m.model_config, ids, sids, lens, n_rows=4, n_seqs=0)
        finally:
            TL.paged_attention = good

        # corrupt the lm_head after the fact: logits / tokens must fail
        m.weight.lm_head.mul_(-1.0)
        with pytest.raises(AssertionError, match="sequence"):
            # the oracle copies the (now negated) weights too, so negate them back for the product only via a hooked linear
            from swiftllm_b200.worker.layers import post_layer as PL
            lin = PL.linear
            PL.linear = lambda x, wt: lin(x, -wt) * 0.5
            try:
                bench.parity_check(m, m.model_config, ids, sids, lens, n_rows=2, n_seqs=2)
            finally:
                PL.linear = lin


def _tp_parity_worker(rank, world, port, ret):
    """bench.parity_check under tensor parallelism: product shards on CPU (oracle-backed kernels, gloo collectives), kv-head shards
    gathered to rank 0, the UNSHARDED weights regenerated through the getter - the code path the N > 1 bench lines rely on."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import warnings
    import swiftllm_b200
    from cpu_shim import product_on_cpu, GlooFusedAllReduce
    from swiftllm_b200.worker.weight import dict_getter
    from test_host_path_cpu import _hf_tensors
    bench = importlib.import_module("bench")
    cfg = dict(CFG, num_hidden_layers=2, vocab_size=96)
    w = OracleWeights.random(cfg, dtype=torch.float16, seed=21, std=0.06)
    tensors = _hf_tensors(w, cfg["intermediate_size"])
    out = None
    with torch.inference_mode(), product_on_cpu([("swiftllm_b200.worker.tp_comm", "FusedAllReduce", GlooFusedAllReduce)]), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ec = swiftllm_b200.EngineConfig(model_path="", use_dummy=False, block_size=16, gpu_mem_utilization=0.9, num_cpu_blocks=0,
                                        max_seqs_in_block_table=8, max_blocks_per_seq=8, max_batch_size=8, max_tokens_in_batch=64,
                                        dtype="float16", tp_size=world, tp_rank=rank)          # library defaults: fused exchange, sharded lm_head
        m = swiftllm_b200.LlamaModel(ec, swiftllm_b200.LlamaModelConfig(cfg))
        m.load_weights(dict_getter(tensors))
        m.init_kvcache_and_swap(24)
        g = torch.Generator().manual_seed(100 + rank)                  # every rank fills ITS kv-head shard
        m.k_cache.copy_(torch.randn(m.k_cache.shape, generator=g).to(m.k_cache.dtype))
        m.v_cache.copy_(torch.randn(m.v_cache.shape, generator=g).to(m.v_cache.dtype))
        sids, lens, ids = [0, 1, 2, 3], [80, 33, 96, 17], [[5], [6], [7], [8]]
        res = bench.parity_check(m, m.model_config, ids, sids, lens, n_rows=4, n_seqs=2, tp_rank=rank, tp_size=world,
                                 full_getter=dict_getter(tensors))
        out = res if rank == 0 else "none" if res is None else "unexpected"
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        ret.put(gathered)
    dist.destroy_process_group()


def test_parity_check_under_tensor_parallelism_on_cpu():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_tp_parity_worker, args=(r, world, 29811, ret)) for r in range(world)]
    for p in procs:
        p.start()
    res = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0 = res[0]
    assert res[1] == "none" and r0["ok"] and r0["tp"] == 2 and len(r0["sequences"]) == 2
    assert all(s["product_vs_storage_dtype_oracle"] <= 4e-3 for s in r0["sequences"])      # only the summation order of the exchange differs
    assert r0["attention_worst_rel_err"] <= 2e-3
