"""Generate tests/golden/hf_*_logits.npz -- an INDEPENDENT pin for the oracle's structure.

The reference ships no golden vectors and cannot be compiled here (no Fortran compiler), so
the oracle is cross-checked against a third implementation instead: Hugging Face
``LlamaForCausalLM`` (transformers, CPU, f32, eager attention) on seeded synthetic weights.
HF implements *canonical* llama (RoPE exponent 2j/hs, 0-based positions, rotate-half layout),
so the comparison runs the oracle with ``canonical=1``; the reference's two RoPE quirks
(Q1/Q2, llama2.f90:544-546) are then the only, separately unit-tested, deviation.  This
confirms rmsnorm, the fused QKV split, GQA head mapping (Q3), softmax scaling, SwiGLU, the
residual wiring and the classifier.

GGUF stores Wq/Wk rows permuted for interleaved-pair RoPE (what llama2.f90:549-557 applies);
HF wants the rotate-half order, so rows are un-permuted when loading them into HF.

Run from the repo root (CPU container, needs torch + transformers):
    python tests/golden/make_hf_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from llm.f90_b200 import fixtures as fx  # noqa: E402
from llm.f90_b200.layout import Config, TINY, F32  # noqa: E402

SEED = 1234
TOKENS_0BASED = [1, 17, 5, 300, 44, 9, 511, 2, 77, 130]  # fed at positions 1..10
# head geometries pinned: TINY (kv_mul 2, head size 32); TinyLlama's grouping (8 query heads per KV
# head -- the case the reference's slice quirk Q3 is about, llama2.f90:581,591) and Llama-2-7B's
# (multi-head attention, head size 128)
CASES = {
    "tiny": TINY,
    "gqa8": dict(emb_dim=256, hidden_dim=704, n_layers=2, n_heads=8, n_kv_heads=1, vocab_size=512, seq_len=64),
    "mha128": dict(emb_dim=256, hidden_dim=704, n_layers=2, n_heads=2, n_kv_heads=2, vocab_size=512, seq_len=64),
}


def gguf_to_hf_rows(w: np.ndarray, n_head: int) -> np.ndarray:
    """Inverse of llama.cpp's convert permute: interleaved pairs -> [first halves | second halves]."""
    out, inn = w.shape
    hs = out // n_head
    return w.reshape(n_head, hs // 2, 2, inn).swapaxes(1, 2).reshape(out, inn)


def make(name: str, shape: dict):
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = Config(**shape, wtype=F32)
    t = fx.synth_tensors(cfg, SEED)
    hc = LlamaConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.emb_dim,
                     intermediate_size=cfg.hidden_dim, num_hidden_layers=cfg.n_layers,
                     num_attention_heads=cfg.n_heads, num_key_value_heads=cfg.n_kv_heads,
                     max_position_embeddings=cfg.seq_len, rms_norm_eps=1e-5, rope_theta=10000.0,
                     tie_word_embeddings=False, attention_bias=False, mlp_bias=False,
                     attn_implementation="eager")
    model = LlamaForCausalLM(hc).to(torch.float32).eval()
    sd = {"model.embed_tokens.weight": t["token_embd.weight"],
          "model.norm.weight": t["output_norm.weight"],
          "lm_head.weight": t["output.weight"]}
    for l in range(cfg.n_layers):
        g, h = f"blk.{l}.", f"model.layers.{l}."
        sd[h + "input_layernorm.weight"] = t[g + "attn_norm.weight"]
        sd[h + "self_attn.q_proj.weight"] = gguf_to_hf_rows(t[g + "attn_q.weight"], cfg.n_heads)
        sd[h + "self_attn.k_proj.weight"] = gguf_to_hf_rows(t[g + "attn_k.weight"], cfg.n_kv_heads)
        sd[h + "self_attn.v_proj.weight"] = t[g + "attn_v.weight"]
        sd[h + "self_attn.o_proj.weight"] = t[g + "attn_output.weight"]
        sd[h + "post_attention_layernorm.weight"] = t[g + "ffn_norm.weight"]
        sd[h + "mlp.gate_proj.weight"] = t[g + "ffn_gate.weight"]
        sd[h + "mlp.down_proj.weight"] = t[g + "ffn_down.weight"]
        sd[h + "mlp.up_proj.weight"] = t[g + "ffn_up.weight"]
    missing = model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()},
                                    strict=False)
    assert not missing.unexpected_keys, missing
    assert all("rotary" in k for k in missing.missing_keys), missing
    with torch.no_grad():
        out = model(torch.tensor([TOKENS_0BASED])).logits[0].double().numpy()
    path = os.path.join(os.path.dirname(__file__), f"hf_{name}_logits.npz")
    np.savez_compressed(path, logits=out.astype(np.float32), tokens_0based=np.array(TOKENS_0BASED),
                        seed=SEED)
    print("wrote", path, out.shape, float(np.abs(out).max()))


if __name__ == "__main__":
    for name, shape in CASES.items():
        make(name, shape)
