"""llmf90_b200_transformer_sample: the pick of the next token made on the device (llama2.f90:388-391, :428-447)
against the host mirror's sequential walk over the same logits."""
import numpy as np
import pytest

from llm.f90_b200 import capi, fixtures as fx, hostapi
from llm.f90_b200.layout import Config, SMALL, F32, Q4_0

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    capi.load()


@pytest.mark.parametrize("granular", [False, True], ids=["stream", "granular"])
@pytest.mark.parametrize("wt", [F32, Q4_0], ids=["f32", "q4_0"])
def test_device_pick_matches_host_sampler(wt, granular):
    cfg = Config(**SMALL, wtype=wt)
    w = fx.synth_weights(cfg, 7)
    rng = np.random.default_rng(1)
    near = 0
    with capi.Engine(w, granular=granular) as eng:
        tok = 2
        for pos in range(1, 21):
            lg = eng.transformer(tok, pos).copy()
            for T in (0.0, 0.7, 1.3):
                r = float(rng.random())
                got = eng.transformer_sample(tok, pos, T, r)  # the same position again: the cache row is rewritten with the same values
                want = hostapi.argmax(lg) if T == 0 else hostapi.sample(lg, T, r)
                if got != want:
                    # the chunked sums may round differently from the sequential chain: only at a CDF boundary
                    assert T != 0
                    z = lg.astype(np.float64) / T
                    p = np.exp(z - z.max())
                    cdf = np.cumsum(p / p.sum())
                    assert np.abs(cdf - r).min() < 1e-5, (pos, T, r, got, want)
                    near += 1
            tok = int(lg.argmax()) + 1
    assert near <= 2


def test_device_pick_edge_values():
    cfg = Config(**SMALL, wtype=F32)
    w = fx.synth_weights(cfg, 7)
    with capi.Engine(w) as eng:
        lg = eng.transformer(2, 1).copy()
        # r = 0 picks the first token with non-zero probability mass reached; r just below 1 a token near the end of the CDF
        assert eng.transformer_sample(2, 1, 1.0, 0.0) == hostapi.sample(lg, 1.0, 0.0)
        hi = float(np.nextafter(np.float32(1.0), np.float32(0.0)))
        got, want = eng.transformer_sample(2, 1, 1.0, hi), hostapi.sample(lg, 1.0, hi)
        assert abs(got - want) <= cfg.vocab_size  # both walk to (nearly) the end; rounding decides where exactly
        assert 1 <= got <= cfg.vocab_size
        with pytest.raises(capi.EngineError):
            eng.transformer_sample(2, 1, -1.0, 0.5)
