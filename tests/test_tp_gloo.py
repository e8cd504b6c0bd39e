"""CPU, world_size 2 over gloo: the tensor-parallel sharding used by swiftllm_b200 (head / FFN-column slices from
`swiftllm_b200.worker.weight`, one all-reduce after o_proj and one after down_proj) reproduces the unsharded layer.
The math runs through the oracle's CPU ops in fp32; the product code under test is the slicing + collective placement."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

CFG = dict(model_type="llama", num_hidden_layers=1, num_attention_heads=8, num_key_value_heads=4, hidden_size=128,
           vocab_size=64, max_position_embeddings=64, intermediate_size=256, rope_theta=10000.0, rms_norm_eps=1e-5,
           hidden_act="silu")


def _layer(x, lw, nq, nkv, D, Fd, reduce):
    from oracle import kernels as K
    T = x.shape[0]
    res = torch.zeros_like(x)
    h, res = K.fused_add_rmsnorm(x, res, lw.attn_norm, 1e-5)
    q = F.linear(h, lw.q_proj).view(T, nq, D); k = F.linear(h, lw.k_proj).view(T, nkv, D); v = F.linear(h, lw.v_proj).view(T, nkv, D)
    o = K.prefill_attention_exact(q, k, v, [0], [T], D ** -0.5, torch.float32).reshape(T, nq * D)
    o = reduce(F.linear(o, lw.o_proj))
    o, res = K.fused_add_rmsnorm(o, res, lw.ffn_norm, 1e-5)
    ug = K.silu_and_mul(F.linear(o, lw.up_gate_proj))
    return reduce(F.linear(ug[:, :Fd], lw.down_proj)) + res


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from swiftllm_b200.model_config import LlamaModelConfig
    from swiftllm_b200.worker.weight import LlamaTransformerLayerWeight, synthetic_getter
    mc = LlamaModelConfig(CFG)
    getter = synthetic_getter(seed=9, std=0.08, device="cpu")
    full = LlamaTransformerLayerWeight(0, mc, torch.float32); full.load_weights(getter, 0, 1, "cpu")
    mine = LlamaTransformerLayerWeight(0, mc, torch.float32); mine.load_weights(getter, rank, world, "cpu")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(10, mc.hidden_size, generator=g)
    ref = _layer(x, full, mc.num_q_heads, mc.num_kv_heads, mc.head_dim, mc.ffn_inter_dim, lambda t: t)

    def allreduce(t):
        dist.all_reduce(t)
        return t
    out = _layer(x, mine, mc.num_q_heads // world, mc.num_kv_heads // world, mc.head_dim, mc.ffn_inter_dim // world, allreduce)
    err = float((out - ref).abs().max() / ref.abs().max())
    gathered = [None] * world
    dist.all_gather_object(gathered, err)
    if rank == 0:
        ret.put(gathered)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_tp_shards_plus_allreduce_equal_unsharded(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29611 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    errs = ret.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert max(errs) < 1e-5, errs
