"""Pin the floating-point oracle (oracle/gpt2.py) against the third-party module that holds the reference's
arithmetic: HF transformers GPT2LMHeadModel (eager attention), loss and every gradient."""
import pytest
import torch

from oracle import gpt2 as og

transformers = pytest.importorskip("transformers")


def hf_model(d: og.GPT2Dims):
    cfg = transformers.GPT2Config(n_embd=d.n_embd, n_head=d.n_head, n_layer=d.n_layer, n_positions=d.n_positions,
                                  vocab_size=d.vocab_size, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0,
                                  use_cache=False, tie_word_embeddings=False, attn_implementation="eager")
    m = transformers.GPT2LMHeadModel(cfg)
    m.train()
    return m


def copy_into_hf(layers, m):
    emb, blocks, head = layers[0], layers[1:-1], layers[-1]
    with torch.no_grad():
        m.transformer.wte.weight.copy_(emb.wte)
        m.transformer.wpe.weight.copy_(emb.wpe)
        for b, hb in zip(blocks, m.transformer.h):
            for p, hp in zip(b.parameters(), hb.parameters()):  # same order by construction
                assert p.shape == hp.shape
                hp.copy_(p)
        m.transformer.ln_f.weight.copy_(head.ln_f_w)
        m.transformer.ln_f.bias.copy_(head.ln_f_b)
        m.lm_head.weight.copy_(head.lm_head_w)


def test_oracle_matches_hf_loss_and_grads():
    torch.manual_seed(0)
    d = og.GPT2Dims(n_embd=64, n_head=4, n_layer=3, n_positions=32, vocab_size=211)
    layers = og.build_layers(d)
    og.init_layers_(layers)
    # perturb LN/bias so that they matter
    g = torch.Generator().manual_seed(1)
    for l in layers:
        for n, p in l.named_parameters():
            if n.endswith("_b") or n.startswith("ln_"):
                with torch.no_grad():
                    p.add_(torch.randn(p.shape, generator=g) * 0.05)
    m = hf_model(d)
    copy_into_hf(layers, m)
    batch = og.synthetic_batch(2, 32, d.vocab_size)
    x = (batch["input_ids"], batch["attention_mask"], batch["labels"])
    for l in layers:
        x = l(*x)
    loss = x[0]
    loss.backward()
    out = m(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], labels=batch["labels"])
    out.loss.backward()
    assert torch.allclose(loss, out.loss, rtol=1e-6, atol=1e-7), (loss.item(), out.loss.item())
    torch.testing.assert_close(x[1], out.logits, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(layers[0].wte.grad, m.transformer.wte.weight.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(layers[-1].lm_head_w.grad, m.lm_head.weight.grad, rtol=1e-4, atol=1e-7)
    for b, hb in zip(layers[1:-1], m.transformer.h):
        for (n, p), hp in zip(b.named_parameters(), hb.parameters()):
            torch.testing.assert_close(p.grad, hp.grad, rtol=1e-4, atol=1e-7, msg=lambda s: f"{n}: {s}")


def test_block_param_count_and_order():
    # GPT-2 124M block = 7.088M params; 12 E^2 + 13 E
    d = og.GPT2Dims()
    b = og.BlockLayer(d)
    assert sum(p.numel() for p in b.parameters()) == 12 * 768 * 768 + 13 * 768 == 7087872
    hb = transformers.models.gpt2.modeling_gpt2.GPT2Block(transformers.GPT2Config(), layer_idx=0)
    assert [tuple(p.shape) for p in b.parameters()] == [tuple(p.shape) for p in hb.parameters()]
    assert len(og.build_layers(d)) == d.n_layer + 2  # tests/module/test_model.py:22


def test_flat_roundtrip():
    d = og.GPT2Dims(n_embd=32, n_head=2, n_layer=1, n_positions=16, vocab_size=50)
    layers = og.build_layers(d)
    og.init_layers_(layers)
    f = og.flat_params(layers[1])
    l2 = og.BlockLayer(d)
    og.load_flat_(l2, f)
    assert torch.equal(og.flat_params(l2), f)
