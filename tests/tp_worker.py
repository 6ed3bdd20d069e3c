"""Worker functions for the multi-process tests (importable by torch.multiprocessing.spawn)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _init_pg(rank, world, port, backend="gloo"):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    return dist


def cpu_tp_forward(rank, world, port, shape, wtype, seed, tokens, out_path):
    """Row-parallel forward in numpy f64 on this rank's shard (llm.f90_b200.tp), the all-reduces after
    Wo and W2 and the logits all-gather over gloo.  Mirrors what one GPU rank computes."""
    import torch
    dist = _init_pg(rank, world, port)
    from llm.f90_b200 import fixtures as fx, tp
    from llm.f90_b200.layout import Config
    cfg = Config(**shape, wtype=wtype)
    w = fx.synth_weights(cfg, seed)
    sh = tp.shard(cfg, rank, world)
    s = {k: v.astype(np.float64) for k, v in tp.shard_f32(w, rank, world).items()}
    hs = cfg.head_size
    hl, kvl = len(sh.heads), len(sh.kv_heads) * hs
    kv_mul = hl // len(sh.kv_heads)  # local heads per local KV head (== the global ratio unless KV heads are replicated)
    kc = np.zeros((cfg.n_layers, cfg.seq_len, kvl))
    vc = np.zeros_like(kc)
    rms = lambda x, g: x * g / np.sqrt(x @ x / x.size + 1e-5)

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    logits_all = []
    for pos, token in enumerate(tokens, start=1):
        x = s["emb"][token - 1].copy()
        j = np.arange(hs // 2)
        ang = pos / 10000.0 ** ((2 * j + 1) / hs)
        cs, sn = np.cos(ang), np.sin(ang)

        def rot(a):
            p = a.reshape(-1, hs // 2, 2)
            a0, a1 = p[..., 0].copy(), p[..., 1].copy()
            p[..., 0] = a0 * cs - a1 * sn
            p[..., 1] = a0 * sn + a1 * cs
        for l in range(cfg.n_layers):
            xb = rms(x, w.rms_att_weight[l].astype(np.float64))
            q, k, v = s["wq"][l] @ xb, s["wk"][l] @ xb, s["wv"][l] @ xb
            rot(q)
            rot(k)
            kc[l, pos - 1], vc[l, pos - 1] = k, v
            att = np.empty(hl * hs)
            for h in range(hl):
                g = h // kv_mul
                sc = kc[l, :pos, g * hs:(g + 1) * hs] @ q[h * hs:(h + 1) * hs] / np.sqrt(hs)
                a = np.exp(sc - sc.max())
                a /= a.sum()
                att[h * hs:(h + 1) * hs] = a @ vc[l, :pos, g * hs:(g + 1) * hs]
            x = x + allreduce(s["wo"][l] @ att)
            xb = rms(x, w.rms_ffn_weight[l].astype(np.float64))
            g1, u = s["w1"][l] @ xb, s["w3"][l] @ xb
            x = x + allreduce(s["w2"][l] @ (g1 * (1.0 / (1.0 + np.exp(-g1))) * u))
        x = rms(x, w.rms_final_weight.astype(np.float64))
        part = torch.from_numpy(s["wcls"] @ x)
        gathered = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(gathered, part)
        logits_all.append(torch.cat(gathered).numpy())
    if rank == 0:
        np.save(out_path, np.array(logits_all))
    dist.barrier()
    dist.destroy_process_group()


def gpu_tp_generate(rank, world, port, shape, wtype, seed, prompt, n, out_path):
    """One tensor-parallel rank of the CUDA engine: host loop + device greedy loop."""
    import torch
    torch.cuda.set_device(rank)
    dist = _init_pg(rank, world, port)
    from llm.f90_b200 import capi, fixtures as fx
    from llm.f90_b200.layout import Config
    cfg = Config(**shape, wtype=wtype)
    w = fx.synth_weights(cfg, seed)
    eng = capi.make_engine(w, device=rank, tp_rank=rank, tp_size=world)
    toks, lg = capi.host_generate(eng, prompt, n, want_logits=True)
    dist.barrier()
    eng.reset()
    dist.barrier()
    dev_toks, ms = eng.generate_greedy(prompt, n)
    st = eng.stats()
    np.savez(out_path + f".rank{rank}.npz", toks=toks, logits=lg, dev_toks=dev_toks, ms=ms,
             active=st["active_bytes_per_token"])
    dist.barrier()
    eng.close()
    dist.destroy_process_group()
