"""float64 numpy restatement of llama2.f90:450-640 -- the arbiter for tolerance budgeting.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see llama2_oracle.c).  Same algorithm and the
same quirks (Q1-Q3) as the C oracle, but every operation in float64 so that
both the C oracle and the CUDA path can be measured against a common "truth": their
distance to this result is the f32 rounding noise the 1e-4 budget has to cover.
"""
from __future__ import annotations

import numpy as np

from llm.f90_b200.fixtures import decode_matrix
from llm.f90_b200.layout import Weights


class OracleNP:
    def __init__(self, w: Weights, canonical: bool = False):
        c = w.cfg
        self.cfg, self.canonical = c, canonical
        e, h = c.emb_dim, c.hidden_dim
        d = lambda a, n: decode_matrix(a, c.wtype, n).astype(np.float64)
        self.emb = d(w.token_embedding_table, e)                      # [V, e]
        self.wqkv = d(w.wqkv, e).reshape(c.n_layers, c.n_qkv, e)
        self.wo = d(w.wo, e).reshape(c.n_layers, e, e)
        self.w13 = d(w.w13, e).reshape(c.n_layers, 2 * h, e)
        self.w2 = d(w.w2, h).reshape(c.n_layers, e, h)
        self.wcls = d(w.wcls, e)
        self.rms_att = w.rms_att_weight.astype(np.float64)
        self.rms_ffn = w.rms_ffn_weight.astype(np.float64)
        self.rms_final = w.rms_final_weight.astype(np.float64)
        self.reset()

    def reset(self):
        c = self.cfg
        self.kc = np.zeros((c.n_layers, c.seq_len, c.kv_head_size))
        self.vc = np.zeros((c.n_layers, c.seq_len, c.kv_head_size))

    @staticmethod
    def rmsnorm(x, w):  # llama2.f90:450-457
        return x * w / np.sqrt(x @ x / x.size + 1e-5)

    def rope_angles(self, pos):  # llama2.f90:543-548 (Q1, Q2)
        hs = self.cfg.head_size
        j = np.arange(hs // 2)
        if self.canonical:
            return (pos - 1) / 10000.0 ** (2 * j / hs)
        return pos / 10000.0 ** ((2 * j + 1) / hs)

    def transformer(self, token: int, pos: int) -> np.ndarray:
        c = self.cfg
        e, hs, H, kv = c.emb_dim, c.head_size, c.n_heads, c.kv_head_size
        kv_mul = H // c.n_kv_heads
        x = self.emb[token - 1].copy()                                 # :520
        ang = self.rope_angles(pos)
        cs, sn = np.cos(ang), np.sin(ang)
        for l in range(c.n_layers):
            xb = self.rmsnorm(x, self.rms_att[l])                      # :527
            qkv = self.wqkv[l] @ xb                                    # :529-531
            q, k, v = qkv[:e].copy(), qkv[e:e + kv].copy(), qkv[e + kv:]

            def rot(a):                                                # :549-557
                p = a.reshape(-1, hs // 2, 2)
                a0, a1 = p[..., 0].copy(), p[..., 1].copy()
                p[..., 0] = a0 * cs - a1 * sn
                p[..., 1] = a0 * sn + a1 * cs
            rot(q)
            rot(k)
            self.kc[l, pos - 1], self.vc[l, pos - 1] = k, v            # :564-565
            xb = np.empty(e)
            for h in range(H):                                         # :574-598 (Q3)
                g = h // kv_mul
                K = self.kc[l, :pos, g * hs:(g + 1) * hs]
                Vv = self.vc[l, :pos, g * hs:(g + 1) * hs]
                s = K @ q[h * hs:(h + 1) * hs] / np.sqrt(hs)
                a = np.exp(s - s.max())
                a /= a.sum()
                xb[h * hs:(h + 1) * hs] = a @ Vv
            x = x + self.wo[l] @ xb                                    # :603-605
            xb = self.rmsnorm(x, self.rms_ffn[l])                      # :608
            h13 = self.w13[l] @ xb                                     # :610-612
            g1, u = h13[:c.hidden_dim], h13[c.hidden_dim:]
            hb = g1 * (1.0 / (1.0 + np.exp(-g1))) * u                  # :615-616
            x = x + self.w2[l] @ hb                                    # :618-620
        x = self.rmsnorm(x, self.rms_final)                            # :627
        return self.wcls @ x                                           # :634-636

    def generate(self, prompt_tokens, n: int):
        toks, logits = [], []
        token = 2                                                      # :376
        for pos in range(1, n + 1):
            lg = self.transformer(token, pos)
            logits.append(lg)
            token = int(prompt_tokens[pos - 1]) if pos <= len(prompt_tokens) else int(np.argmax(lg)) + 1
            toks.append(token)
        return np.array(toks, np.int32), np.array(logits)
