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


def _worker_model(rank, world, port, ret, shard_lm_head=False, comm=None):
    """The PRODUCT's LlamaModel with tp_size=world on CPU (kernel wrappers -> oracle restatements, tests/cpu_shim.py; the
    collectives run over gloo) against the unsharded oracle model: same greedy tokens, logits within fp16 tolerance, identical
    block tables on every rank."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import warnings
    import numpy as np
    import swiftllm_b200
    from cpu_shim import product_on_cpu, GlooFusedAllReduce
    from oracle.model import OracleLlama, OracleWeights
    from swiftllm_b200.worker.weight import dict_getter
    from test_host_path_cpu import _hf_tensors
    cfg = dict(CFG, num_hidden_layers=2, vocab_size=96)
    w = OracleWeights.random(cfg, dtype=torch.float16, seed=12, std=0.08)
    ok = True
    extra = [("swiftllm_b200.worker.tp_comm", "FusedAllReduce", GlooFusedAllReduce)] if comm else []
    with product_on_cpu(extra), warnings.catch_warnings():
        warnings.simplefilter("ignore")                    # "peer-memory exchange unavailable ... using NCCL all-reduce"
        ec = swiftllm_b200.EngineConfig(model_path="", use_dummy=False, block_size=16, gpu_mem_utilization=0.9, num_cpu_blocks=2,
                                        max_seqs_in_block_table=8, max_blocks_per_seq=8, max_batch_size=8, max_tokens_in_batch=128,
                                        dtype="float16", tp_size=world, tp_rank=rank, shard_lm_head=shard_lm_head,
                                        fused_allreduce={None: False, "one_shot": True, "two_shot": "two_shot", "ll": "ll", "two_shot+limit": "two_shot",
                                                         "auto": None}[comm])
        m = swiftllm_b200.LlamaModel(ec, swiftllm_b200.LlamaModelConfig(cfg))
        m.load_weights(dict_getter(_hf_tensors(w, cfg["intermediate_size"])))
        if comm == "auto":                                 # the library's own choice for this TP degree (EngineConfig defaults)
            ec.shard_lm_head = None
            m = swiftllm_b200.LlamaModel(ec, swiftllm_b200.LlamaModelConfig(cfg))
            m.load_weights(dict_getter(_hf_tensors(w, cfg["intermediate_size"])))
            assert m.comm is not None and not m.comm.two_shot and m.weight.lm_head_sharded     # tp 2..4: one-shot, sharded lm_head
            shard_lm_head = True
        assert m.weight.lm_head.shape[0] == (cfg["vocab_size"] // world if shard_lm_head else cfg["vocab_size"])
        m.init_kvcache_and_swap(20)
        m.post_layer.keep_logits = True
        assert (m.comm is None) == (comm is None)          # no peer memory on CPU: the all-reduce path unless a stand-in is given
        assert comm is None or m.comm.two_shot == (comm in ("two_shot", "ll", "two_shot+limit"))
        if comm == "two_shot+limit":
            m.comm.max_fused_tokens = 4                    # the 24-row prompt step falls back, the 2-row decode steps do not
        o = OracleLlama(cfg, w, block_size=16, num_blocks=20, num_cpu_blocks=2, max_seqs_in_block_table=8, max_blocks_per_seq=8,
                        attn="exact", dtype=torch.float16)
        rng = np.random.default_rng(2)
        prompts = [rng.integers(0, 96, size=n).tolist() for n in (21, 3)]
        sids, lens = [4, 1], [21, 3]
        ids, dec = prompts, []
        for stepno in range(4):
            tm, to = m.forward(ids, sids, dec), o.forward(ids, sids, dec)
            ref = o.last_logits.float()
            got = m.post_layer.last_logits.float()
            rel = float((got - ref).abs().max() / ref.abs().max())
            top2 = ref.topk(2, dim=1).values
            clear = ((top2[:, 0] - top2[:, 1]) > 8e-3 * ref.abs().max()).tolist()
            ok &= rel <= 4e-3 and all(a == b for a, b, c in zip(tm, to, clear) if c)
            n = o.gpu_block_manager.num_seq_allocated_blocks
            ok &= bool(np.array_equal(m.gpu_block_manager.num_seq_allocated_blocks.numpy(), n))
            for s in sids:
                ok &= bool(np.array_equal(m.gpu_block_manager.block_table.numpy()[s, : n[s]], o.gpu_block_manager.block_table[s, : n[s]]))
            lens = [l + 1 for l in lens]
            ids, dec = [[t] for t in to], lens
    gathered = [None] * world
    dist.all_gather_object(gathered, bool(ok))
    if rank == 0:
        ret.put(gathered)
    dist.destroy_process_group()


@pytest.mark.parametrize("shard_lm_head", [False, True], ids=["replicated-lm_head", "vocab-sharded-lm_head"])
@pytest.mark.parametrize("world", [2])
def test_product_model_tensor_parallel_on_cpu(world, shard_lm_head):
    """shard_lm_head: every rank computes V / world logit columns; the greedy token comes from one all-gather of
    (max logit, index) pairs and must equal the unsharded argmax wherever the top-1 margin is clear."""
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29641 + world + (20 if shard_lm_head else 0)
    procs = [ctx.Process(target=_worker_model, args=(r, world, port, ret, shard_lm_head)) for r in range(world)]
    for p in procs:
        p.start()
    oks = ret.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(oks), oks


def test_sharded_argmax_merge_prefers_the_first_occurrence():
    """Ties at the maximum (likely with fp16 logits over a 128k vocabulary) must resolve to the smallest token id, like
    torch.argmax over the unsharded logits does."""
    from swiftllm_b200.worker.layers.post_layer import merge_sharded_argmax
    g = torch.Generator().manual_seed(0)
    N, B, Vs = 4, 9, 6
    for _ in range(50):
        logits = torch.randint(0, 3, (B, N * Vs), generator=g).to(torch.float16)         # lots of ties
        shards = logits.view(B, N, Vs).permute(1, 0, 2)                                   # [N, B, Vs]
        first = lambda row: int((row == row.max()).nonzero()[0])
        loc = torch.tensor([[first(shards[n, b]) for b in range(B)] for n in range(N)])
        pairs = torch.stack((shards.gather(2, loc[..., None]).squeeze(2).float(), (loc + torch.arange(N)[:, None] * Vs).float()), dim=2)
        assert merge_sharded_argmax(pairs).tolist() == [first(logits[b]) for b in range(B)]


@pytest.mark.parametrize("comm", ["one_shot", "two_shot", "ll", "two_shot+limit", "auto"])
@pytest.mark.parametrize("world", [2, 3])
def test_product_model_fused_exchange_host_path_on_cpu(world, comm):
    """The host side of the fused exchange (GEMM partials written into the comm buffer, layer i's down_proj exchange folded
    into layer i+1's first norm, final norm in the model tail, and - two_shot - a residual that is only maintained for the rows
    a rank owns) with a gloo stand-in that follows the CUDA kernels' semantics (tests/cpu_shim.py: GlooFusedAllReduce; rows a
    rank does not own are poisoned with NaN there).  world 3 leaves T = 2 tokens without an owner on one rank.
    "two_shot+limit": steps with more rows than the exchange object covers (max_fused_tokens; whole-prompt prefill in production)
    take the all-reduce + add/norm path while the 2-row decode steps of the same model take the fused path.
    "auto": `fused_allreduce=None`, `shard_lm_head=None` - the candidate loop of `LlamaModel.load_weights` and the defaults it picks."""
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29700 + world + {"one_shot": 0, "two_shot": 10, "ll": 20, "two_shot+limit": 30, "auto": 40}[comm]
    procs = [ctx.Process(target=_worker_model_w, args=(r, world, port, ret, comm)) for r in range(world)]
    for p in procs:
        p.start()
    oks = ret.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(oks), oks


@pytest.mark.slow_cpu
@pytest.mark.parametrize("comm", ["two_shot", "one_shot", None], ids=["row-owner", "one-shot", "all-reduce"])
def test_product_model_world8_host_path_on_cpu(comm):
    """World 8 (the TP degree of BASELINE configs[3] / [4]): the host path with the row-owner exchange family that is the default
    there (two-shot semantics: 2-row decode steps leave six ranks without a row to own) and the vocabulary-sharded lm_head, over
    gloo with oracle-backed kernels.  8 processes on the CPU container."""
    world = 8
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29790 + {"two_shot": 0, "one_shot": 1, None: 2}[comm]
    procs = [ctx.Process(target=_worker_model_w, args=(r, world, port, ret, comm, True)) for r in range(world)]
    for p in procs:
        p.start()
    oks = ret.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(oks), oks


def _worker_model_w(rank, world, port, ret, comm, shard_lm_head=False):
    global CFG
    if world == 3:                                         # heads / FFN columns must divide by the world size
        CFG = dict(CFG, num_attention_heads=6, num_key_value_heads=3, hidden_size=96, intermediate_size=192)
    if world == 8:                                         # 1 kv head / 2 q heads per rank, like Llama-3 at TP 8 per kv head
        CFG = dict(CFG, num_attention_heads=16, num_key_value_heads=8, hidden_size=256, intermediate_size=256)
    torch.set_num_threads(1 if world > 4 else torch.get_num_threads())
    _worker_model(rank, world, port, ret, shard_lm_head, comm)


def _worker_chunked(rank, world, port, ret):
    """Chunked prefill + piggybacked decode through the PRODUCT model with tp_size=world on CPU (oracle-backed kernels, gloo
    collectives, vocabulary-sharded lm_head) against the unsharded oracle run with whole prompts."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import warnings
    import numpy as np
    import swiftllm_b200
    from cpu_shim import product_on_cpu
    from oracle.model import OracleLlama, OracleWeights
    from swiftllm_b200.worker.weight import dict_getter
    from test_host_path_cpu import _hf_tensors
    cfg = dict(CFG, num_hidden_layers=2, vocab_size=96)
    w = OracleWeights.random(cfg, dtype=torch.float16, seed=14, std=0.08)
    ok = True
    with product_on_cpu(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ec = swiftllm_b200.EngineConfig(model_path="", use_dummy=False, block_size=16, gpu_mem_utilization=0.9, num_cpu_blocks=2,
                                        max_seqs_in_block_table=8, max_blocks_per_seq=8, max_batch_size=8, max_tokens_in_batch=128,
                                        dtype="float16", tp_size=world, tp_rank=rank, shard_lm_head=True)
        m = swiftllm_b200.LlamaModel(ec, swiftllm_b200.LlamaModelConfig(cfg))
        m.load_weights(dict_getter(_hf_tensors(w, cfg["intermediate_size"])))
        m.init_kvcache_and_swap(20)
        o = OracleLlama(cfg, w, block_size=16, num_blocks=20, num_cpu_blocks=2, max_seqs_in_block_table=8, max_blocks_per_seq=8,
                        attn="exact", dtype=torch.float16)
        rng = np.random.default_rng(8)
        pa, pb = rng.integers(0, 96, size=37).tolist(), rng.integers(0, 96, size=6).tolist()
        tb = m.forward([pb], [4], [])                                                   # B: whole prompt
        ok &= tb == o.forward([pb], [4], [])
        m.forward([pa[:16], tb], [1, 4], [7], prefill_prefix_lens_list=[0])            # A chunk 1 + decode of B
        tb2 = o.forward([tb], [4], [7])
        t = m.forward([pa[16:29], tb2], [1, 4], [8], prefill_prefix_lens_list=[16])    # A chunk 2 (enters a page mid-way)
        tb3 = o.forward([tb2], [4], [8])
        ok &= t[1] == tb3[0]
        t = m.forward([pa[29:]], [1], [], prefill_prefix_lens_list=[29])               # A last chunk
        ok &= t == o.forward([pa], [1], [])                                             # == whole-prompt prefill of A
    gathered = [None] * world
    dist.all_gather_object(gathered, bool(ok))
    if rank == 0:
        ret.put(gathered)
    dist.destroy_process_group()


def test_product_model_chunked_prefill_under_tensor_parallelism_on_cpu():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker_chunked, args=(r, 2, 29755, ret)) for r in range(2)]
    for p in procs:
        p.start()
    oks = ret.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(oks), oks
