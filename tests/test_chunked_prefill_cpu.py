"""CPU: chunked ("prefix-aware", SARATHI-style) prefill - SURVEY.md §8 f-1.

The reference cannot express a partial prompt, so there is no reference output to pin against; the definition is
"processing a prompt in chunks gives what processing it at once gives".  These tests hold (a) the oracle's chunked path to
the oracle's whole-prompt path (which IS pinned against the reference's golden trace, tests/test_oracle_golden.py), and
(b) the product's host path (positions, block allocation across chunks, metadata staging, layer dispatch) to the oracle."""
import numpy as np
import pytest
import torch

from oracle import kernels as K
from oracle.model import OracleLlama, OracleWeights
from cpu_shim import product_on_cpu
from test_host_path_cpu import CFG, ENG, _oracle, _product, _same_state


def _rand_cache(rng, nb, L, nkv, bs, D, dtype):
    return torch.from_numpy(rng.standard_normal((nb, L, nkv, bs, D)).astype(np.float32)).to(dtype)


def test_oracle_prefix_attention_with_zero_prefix_is_plain_prefill_attention():
    rng = np.random.default_rng(0)
    nq, nkv, D, bs, L = 4, 2, 32, 16, 2
    lens = [37, 16, 5]
    starts = [0, 37, 53]
    T = sum(lens)
    q = torch.from_numpy(rng.standard_normal((T, nq, D)).astype(np.float32)).half()
    k = torch.from_numpy(rng.standard_normal((T, nkv, D)).astype(np.float32)).half()
    v = torch.from_numpy(rng.standard_normal((T, nkv, D)).astype(np.float32)).half()
    kc, vc = _rand_cache(rng, 12, L, nkv, bs, D, torch.float16), _rand_cache(rng, 12, L, nkv, bs, D, torch.float16)
    bt = np.full((4, 4), -1, dtype=np.int32)
    bt[3, :3] = [7, 2, 9]; bt[0, :1] = [4]; bt[1, :1] = [11]
    sids = [3, 0, 1]
    K.store_kvcache_inplace(k, v, kc, vc, bt, sids, starts, lens, [], 3, T, bs, 1, prefill_prefix_lens=[0, 0, 0])
    kc2, vc2 = kc.clone(), vc.clone()
    K.store_kvcache_inplace(k, v, kc2, vc2, bt, sids, starts, lens, [], 3, T, bs, 1)
    assert torch.equal(kc, kc2) and torch.equal(vc, vc2)
    a = K.prefix_prefill_attention_exact(q, kc, vc, bt, sids, starts, lens, [0, 0, 0], D ** -0.5, bs, 1)
    b = K.prefill_attention_exact(q, k, v, starts, lens, D ** -0.5)
    assert float((a - b).abs().max()) < 1e-12


def test_oracle_store_with_prefix_enters_pages_in_the_middle():
    rng = np.random.default_rng(1)
    nkv, D, bs, L = 2, 16, 16, 1
    kc, vc = torch.zeros(6, L, nkv, bs, D, dtype=torch.float16), torch.zeros(6, L, nkv, bs, D, dtype=torch.float16)
    bt = np.array([[5, 1, 3, 0]], dtype=np.int32)
    k = torch.from_numpy(rng.standard_normal((20, nkv, D)).astype(np.float32)).half()
    v = torch.from_numpy(rng.standard_normal((20, nkv, D)).astype(np.float32)).half()
    K.store_kvcache_inplace(k, v, kc, vc, bt, [0], [0], [20], [], 1, 20, bs, 0, prefill_prefix_lens=[13])    # positions 13..32
    flat_k = torch.cat([kc[b, 0] for b in (5, 1, 3)], dim=1)          # [nkv, 48, D] in sequence order
    assert torch.equal(flat_k[:, 13:33].transpose(0, 1), k)
    assert float(flat_k[:, :13].abs().max()) == 0 and float(flat_k[:, 33:].abs().max()) == 0


@pytest.mark.parametrize("chunks", [(16, 16, 5), (7, 30), (1, 35, 1), (37,)])
def test_oracle_chunked_prefill_equals_whole_prompt_prefill(chunks):
    """Same prompt, same weights: the logits after the last chunk equal the whole-prompt logits up to the rounding of the
    attention output (the two paths sum the same fp64 terms in a different grouping), the KV cache likewise, the sampled token
    exactly; a decode step after either gives the same token."""
    w = OracleWeights.random(CFG, dtype=torch.float16, seed=4, std=0.08)
    rng = np.random.default_rng(11)
    prompt = rng.integers(0, 300, size=sum(chunks)).tolist()
    whole, chunked = _oracle(w), _oracle(w)
    t_whole = whole.forward([prompt], [2], [])
    pos = 0
    for n in chunks:
        t_chunk = chunked.forward([prompt[pos:pos + n]], [2], [], prefill_prefix_lens_list=[pos])
        pos += n
    assert t_chunk == t_whole
    ref = whole.last_logits
    assert float((chunked.last_logits - ref).abs().max()) <= 2e-3 * float(ref.abs().max())
    assert np.array_equal(chunked.gpu_block_manager.block_table[2, :3], whole.gpu_block_manager.block_table[2, :3])
    assert float((chunked.k_cache.float() - whole.k_cache.float()).abs().max()) <= 4e-3
    assert float((chunked.v_cache.float() - whole.v_cache.float()).abs().max()) <= 4e-3
    assert whole.forward([t_whole], [2], [len(prompt) + 1]) == chunked.forward([t_chunk], [2], [len(prompt) + 1])


def test_product_host_path_chunked_prefill_matches_oracle_on_cpu():
    """SARATHI-style schedule through the PRODUCT's forward on CPU (oracle-backed kernels): two prompts chunked differently,
    a third sequence decoding in the same batches.  Tokens, logits, block tables, free map and host mirror must match the
    oracle fed the same calls; the final tokens must equal those of whole-prompt prefill."""
    w = OracleWeights.random(CFG, dtype=torch.float16, seed=6, std=0.08)
    rng = np.random.default_rng(21)
    pa, pb = rng.integers(0, 300, size=45).tolist(), rng.integers(0, 300, size=23).tolist()
    pc = rng.integers(0, 300, size=9).tolist()
    with product_on_cpu():
        m, o, whole = _product(w), _oracle(w), _oracle(w)

        def step(ids, sids, dlens, prefix):
            tm = m.forward(ids, sids, dlens, prefill_prefix_lens_list=prefix)
            to = o.forward(ids, sids, dlens, prefill_prefix_lens_list=prefix)
            assert tm == to
            ref = o.last_logits
            assert float((m.post_layer.last_logits - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
            _same_state(m, o)
            return to

        t = step([pc], [4], [], None)                                        # sequence 4: ordinary whole-prompt prefill
        lc = len(pc)
        # batch 1: first chunks of A (seq 1) and B (seq 6) + decode of seq 4
        lc += 1; t = step([pa[:20], pb[:16], [t[0]]], [1, 6, 4], [lc], [0, 0])
        # batch 2: second chunk of A (enters a page in the middle), rest of B + decode
        lc += 1; t = step([pa[20:33], pb[16:], [t[2]]], [1, 6, 4], [lc], [20, 16])
        tb = t[1]
        # batch 3: last chunk of A alone with the decodes of B and 4
        lc += 1; t = step([pa[33:], [tb], [t[2]]], [1, 6, 4], [len(pb) + 1, lc], [33])
        ta = t[0]
        assert whole.forward([pa, pb], [1, 6], []) == [ta, tb]
        assert m.gpu_block_manager.get_num_allocated_blocks_host([1, 6, 4]) == [3, 2, 1]


def test_product_rejects_inconsistent_chunk_arguments():
    w = OracleWeights.random(CFG, dtype=torch.float16, seed=3, std=0.08)
    with product_on_cpu():
        m = _product(w)
        with pytest.raises(AssertionError, match="one non-negative entry per prefill"):
            m.forward([[1, 2, 3], [4, 5]], [0, 1], [], prefill_prefix_lens_list=[0])
        with pytest.raises(AssertionError, match="ignore_kvcache"):
            m.forward([[1, 2, 3]], [0], [], ignore_kvcache=True, prefill_prefix_lens_list=[0])


def test_chunked_replay_of_the_reference_golden_trace(golden):
    """Pins the chunked path against the UNMODIFIED reference's own outputs (tests/golden/model_tiny.npz): every prefill of the
    recorded trace is replayed in chunks (sizes that enter pages in the middle); the greedy tokens of every call must equal the
    reference's and the logits stay within the tolerance the whole-prompt oracle is held to plus the attention-output rounding
    (physical block ids differ by construction - chunks allocate later - so block tables are not compared here)."""
    import json
    z = golden("model_tiny")
    cfg = json.loads(str(z["config"])); eng = json.loads(str(z["engine"]))
    w = OracleWeights.from_golden(z, cfg["num_hidden_layers"])
    m = OracleLlama(cfg, w, block_size=eng["block_size"], num_blocks=eng["num_blocks"], num_cpu_blocks=eng["num_cpu_blocks"],
                    max_seqs_in_block_table=eng["max_seqs_in_block_table"], max_blocks_per_seq=eng["max_blocks_per_seq"])
    calls = json.loads(str(z["calls"]))
    checked = 0
    for i, c in enumerate(calls):
        if c["op"] == "forward":
            nd = len(c["dec_lens"])
            prompts = c["input_ids"][: len(c["input_ids"]) - nd]
            psids = c["seq_ids"][: len(prompts)]
            ref_tok = z[f"t{i}_tokens"].tolist()
            ref_logits = torch.from_numpy(z[f"t{i}_logits"]).float()
            tol = 2e-3 * float(ref_logits.abs().max())
            # all but the last chunk of every prompt first (7-token chunks), one prompt per call
            for p, s in zip(prompts, psids):
                for pos in range(0, max(len(p) - 7, 0), 7):
                    if pos + 7 < len(p):
                        m.forward([p[pos:pos + 7]], [s], [], prefill_prefix_lens_list=[pos])
            # then the reference's call with every prompt replaced by its last chunk (decodes piggybacked as recorded)
            last_start = [((len(p) - 1) // 7) * 7 if len(p) > 7 else 0 for p in prompts]
            ids = [p[a:] for p, a in zip(prompts, last_start)] + c["input_ids"][len(prompts):]
            toks = m.forward(ids, c["seq_ids"], c["dec_lens"], prefill_prefix_lens_list=last_start if prompts else None)
            assert toks == ref_tok, (i, toks, ref_tok)
            assert float((m.last_logits.float() - ref_logits).abs().max()) <= tol
            checked += 1
        elif c["op"] == "swap_out":
            m.swap_out_seqs(c["seq_ids"])
        elif c["op"] == "swap_in":
            m.swap_in_seqs(c["seq_ids"])
        else:
            m.free_seqs_resources(c["seq_ids"])
    assert checked == 6


def test_piggyback_planner_drives_the_product_to_the_whole_prompt_result():
    """plan_piggyback_step / apply_step_result (swiftllm_b200/worker/chunking.py, the worker-side half of the hook at
    scheduler.py:93-94) driving the product on CPU: three prompts under a 24-token step budget with decodes joining as prompts
    finish.  Every sequence must produce the tokens the oracle produces with whole-prompt prefill followed by decoding."""
    from swiftllm_b200.worker.chunking import PrefillProgress, apply_step_result, plan_piggyback_step
    w = OracleWeights.random(CFG, dtype=torch.float16, seed=9, std=0.08)
    rng = np.random.default_rng(31)
    prompts = {3: rng.integers(0, 300, size=41).tolist(), 0: rng.integers(0, 300, size=9).tolist(), 5: rng.integers(0, 300, size=30).tolist()}
    NEW = 4                                                          # tokens to generate per sequence
    # expected: whole-prompt prefill of each sequence alone, then NEW-1 decode steps
    expect = {}
    for sid, p in prompts.items():
        o = _oracle(w)
        toks = o.forward([p], [sid], [])
        for j in range(NEW - 1):
            toks.append(o.forward([[toks[-1]]], [sid], [len(p) + j + 1])[0])
        expect[sid] = toks
    with product_on_cpu():
        m = _product(w)
        pending = [PrefillProgress(sid, p) for sid, p in prompts.items()]
        out = {sid: [] for sid in prompts}
        lens = {}
        steps = 0
        while any(len(v) < NEW for v in out.values()):
            running = [sid for sid in out if 0 < len(out[sid]) < NEW]
            for sid in running:
                lens[sid] += 1
            step = plan_piggyback_step([p for p in pending if p.remaining > 0], running, [out[s][-1] for s in running],
                                       [lens[s] for s in running], max_tokens_in_step=24, max_chunk=16)
            live = [p for p in pending if p.remaining > 0]
            toks = m.forward(step.input_ids_list, step.seq_ids_list, step.decoding_seq_lens_list,
                             prefill_prefix_lens_list=step.prefill_prefix_lens_list)
            assert sum(len(x) for x in step.input_ids_list) <= 24
            first, dec = apply_step_result(step, live, toks)
            for sid, t in zip(running, dec):
                out[sid].append(t)
            for sid, t in first.items():
                out[sid].append(t); lens[sid] = len(prompts[sid])
            steps += 1
            assert steps < 40
        assert out == expect


def test_oracle_prefix_attention_any_chunking_equals_whole_prompt_attention():
    """Property (seeded random chunkings, ragged page boundaries, GQA): storing a prompt's K/V chunk by chunk and attending
    through the block table gives, row for row, what causal attention over the packed prompt gives."""
    rng = np.random.default_rng(123)
    for trial in range(12):
        nkv = int(rng.choice([1, 2, 4])); g = int(rng.choice([1, 2, 4])); nq = nkv * g
        D = int(rng.choice([16, 32])); bs = int(rng.choice([4, 16])); Ltot = int(rng.integers(1, 70)); L = 2
        q = torch.from_numpy(rng.standard_normal((Ltot, nq, D)).astype(np.float32)).half()
        k = torch.from_numpy(rng.standard_normal((Ltot, nkv, D)).astype(np.float32)).half()
        v = torch.from_numpy(rng.standard_normal((Ltot, nkv, D)).astype(np.float32)).half()
        whole = K.prefill_attention_exact(q, k, v, [0], [Ltot], D ** -0.5)
        nblk = (Ltot + bs - 1) // bs
        perm = rng.permutation(nblk + 3)
        bt = np.full((2, nblk + 1), -1, dtype=np.int32); bt[1, :nblk] = perm[:nblk]
        kc = torch.full((nblk + 3, L, nkv, bs, D), float("nan"), dtype=torch.float16); vc = kc.clone()
        pos = 0
        while pos < Ltot:
            n = int(rng.integers(1, Ltot - pos + 1))
            K.store_kvcache_inplace(k[pos:pos + n], v[pos:pos + n], kc, vc, bt, [1], [0], [n], [], 1, n, bs, 1, prefill_prefix_lens=[pos])
            out = K.prefix_prefill_attention_exact(q[pos:pos + n], kc, vc, bt, [1], [0], [n], [pos], D ** -0.5, bs, 1)
            assert float((out - whole[pos:pos + n]).abs().max()) < 1e-12, (trial, pos, n)
            pos += n


def test_offline_example_generation_loop_chunked_equals_unchunked():
    """examples/offline.py: generate() with --chunk must produce the tokens of whole-prompt prefill (product on CPU)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("offline_example", os.path.join(root, "examples", "offline.py"))
    ex = importlib.util.module_from_spec(spec); spec.loader.exec_module(ex)
    w = OracleWeights.random(CFG, dtype=torch.float16, seed=5, std=0.08)
    rng = np.random.default_rng(3)
    prompts = [rng.integers(0, 300, size=n).tolist() for n in (9, 40, 1, 23)]
    with product_on_cpu():
        a, _, _ = ex.generate(_product(w), prompts, 5, chunk=0)
        b, _, _ = ex.generate(_product(w), prompts, 5, chunk=16)
        c, _, _ = ex.generate(_product(w), prompts, 5, chunk=7)
    assert a == b == c and all(len(o) == 5 for o in a)
