"""Host-side splice plan (spatialrgpt_b200/splice_plan.py) against the oracle's restatement of llava_arch.py:434-539,
and the paged-KV bookkeeping of the batched prefill - no GPU involved."""
import torch

from oracle import srgpt_oracle as O
from spatialrgpt_b200.constants import IGNORE_INDEX, IMAGE_TOKEN_INDEX
from spatialrgpt_b200.splice_plan import build_splice_plan


def _apply(plan, tables):
    rows = [tables[int(s)][int(r)] for s, r in zip(plan.src_id, plan.src_row)]
    return list(torch.split(torch.stack(rows), plan.lens, 0))


def _case(B, regions, lens, n_tok=5, H=8, seed=0, depth=True, none_masks=()):
    g = torch.Generator().manual_seed(seed)
    cfg = O.OracleConfig(vocab=64, hidden=H, enable_depth=depth, mask_token_id=62, depth_token_id=63)
    T = max(lens)
    ids = torch.zeros(B, T, dtype=torch.long)
    am = torch.zeros(B, T, dtype=torch.bool)
    for b in range(B):
        row = torch.randint(3, 40, (lens[b],), generator=g)
        row[1] = IMAGE_TOKEN_INDEX
        p = 3
        for _ in range(regions[b]):
            row[p] = cfg.mask_token_id
            if depth:
                row[p + 1] = cfg.depth_token_id
            p += 3
        ids[b, :lens[b]] = row
        am[b, :lens[b]] = True
    embed = torch.randn(cfg.vocab, H, generator=g)
    feats = torch.randn(B, n_tok, H, generator=g)
    me = [None if b in none_masks else torch.randn(regions[b], H, generator=g) for b in range(B)]
    de = [None if b in none_masks else torch.randn(regions[b], H, generator=g) for b in range(B)] if depth else None
    return cfg, ids, am, embed, feats, me, de


def test_plan_reproduces_the_reference_splice():
    B, regions, lens = 3, [2, 0, 3], [14, 6, 17]
    cfg, ids, am, embed, feats, me, de = _case(B, regions, lens)
    ref = O.splice_embeddings(cfg, embed, ids, feats, me, de, attention_mask=am)
    plan = build_splice_plan(ids, am, None, feats.shape[1], regions, [True] * B, cfg.mask_token_id, cfg.depth_token_id, True, True)
    flat = lambda xs: torch.cat([x for x in xs if x is not None and x.numel()], 0) if any(x is not None and x.numel() for x in xs) else torch.zeros(0, 8)
    got = _apply(plan, {0: embed, 1: feats.reshape(-1, 8), 2: flat(me), 3: flat(de)})
    assert plan.images_used == B and plan.lens == [r.shape[0] for r in ref] and not plan.warnings
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    # image rows carry IGNORE_INDEX labels, text rows keep theirs
    labels = torch.arange(B * ids.shape[1]).view(B, -1)
    plan2 = build_splice_plan(ids, am, labels, feats.shape[1], regions, [True] * B, cfg.mask_token_id, cfg.depth_token_id, True, True)
    for b in range(B):
        lab = plan2.labels[b]
        assert lab.numel() == plan2.lens[b] and int((lab == IGNORE_INDEX).sum()) == feats.shape[1]
        assert lab[0] == labels[b, 0] and lab[-1] == labels[b, lens[b] - 1]


def test_plan_without_depth_none_masks_text_only_and_truncation():
    B, regions, lens = 3, [2, 1, 2], [12, 9, 15]
    cfg, ids, am, embed, feats, me, de = _case(B, regions, lens, depth=False, none_masks=(1,), seed=3)
    ref = O.splice_embeddings(cfg, embed, ids, feats, me, None, attention_mask=am)
    counts = [0 if m is None else m.shape[0] for m in me]
    plan = build_splice_plan(ids, am, None, feats.shape[1], counts, [m is not None for m in me], cfg.mask_token_id, cfg.depth_token_id, True, False)
    got = _apply(plan, {0: embed, 1: feats.reshape(-1, 8), 2: torch.cat([m for m in me if m is not None], 0), 3: embed})
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    assert any("mask embed is None" in w for w in plan.warnings)  # sample 1 has a <mask> token but no mask list entry
    # a sample with no image consumes no image; truncation cuts every sample
    ids2 = ids.clone(); ids2[1][ids2[1] == IMAGE_TOKEN_INDEX] = 7
    plan3 = build_splice_plan(ids2, am, None, feats.shape[1], counts, [True, True, True], cfg.mask_token_id, cfg.depth_token_id, False, False, max_len=10)
    assert plan3.images_used == 2 and plan3.lens == [10, 9, 10]
    assert int((plan3.src_id == 2).sum()) == 0  # regions off: <mask> ids are looked up in the token table


def test_paged_cache_reserve_many_matches_reserve():
    from spatialrgpt_b200.config import LlamaDims
    from spatialrgpt_b200.llama_decoder import PAGE_SIZE, PagedKVCache
    d = LlamaDims(hidden_size=64, num_hidden_layers=1, num_attention_heads=2, num_key_value_heads=1, head_dim=32, intermediate_size=64, vocab_size=32)
    a = PagedKVCache(d, 40, 4, 12, "cpu")
    b = PagedKVCache(d, 40, 4, 12, "cpu")
    toks = [17, 1, 100, 33]
    a.reserve_many(toks)
    for s, n in enumerate(toks):
        b.reserve(s, n)
    assert a.owned == b.owned and torch.equal(a.page_tables, b.page_tables)
    assert [len(o) for o in a.owned] == [(n + PAGE_SIZE - 1) // PAGE_SIZE for n in toks]
    a.release(2)
    a.reserve_many([17, 1, 5])  # re-use after release: only sequence 2 needs a page again
    assert len(a.owned[2]) == 1 and len(set(sum(a.owned, []))) == sum(len(o) for o in a.owned)
