"""Pins oracle/llmvox_oracle.py against fixtures produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import llmvox_oracle as O


def snr_db(ref, x):
    ref = np.asarray(ref, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((ref - x) ** 2).sum(), 1e-300))


def test_tokenizer_matches_byt5(gold):
    g = gold("tokenizer.npz")
    for text, ids in zip(g["texts"], g["ids"]):
        assert O.byt5_tokenize(str(text).strip()) == list(ids), text


def test_config0_greedy_loop_tokens_and_logits(gold, weights):
    g = gold("config0_loop.npz")
    n = int(g["n_steps"])
    codes, logits = O.decode_steps(weights, O.GPTArch(), g["text_ids"].tolist(), n, return_logits=True)
    assert codes == g["codes"].tolist()                      # greedy tokens identical
    rows = g["logits_row_idx"]
    assert np.abs(logits.numpy()[rows] - g["logits_rows"]).max() < 1e-5
    # the reference loop feeds T = t+1 rows at step t (true positions, not min(t,1))
    assert g["input_T"].tolist() == list(range(1, n + 1))


def test_config0_chunks_and_pcm(gold, weights):
    g = gold("config0_loop.npz")
    codes = g["codes"].tolist()
    chunks, used, dump = O.chunk_schedule(codes, 10, stop_on_eoa=True)
    assert [len(c) for c in chunks] == g["chunk_lens"].tolist() == [10, 30, 90]
    for i, c in enumerate(chunks[:2]):
        pcm = O.vocoder_decode(weights, c).numpy()
        assert pcm.shape == g[f"pcm{i}"].shape
        assert np.abs(pcm - g[f"pcm{i}"]).max() < 1e-5


def test_teacher_forced_logits(gold, weights):
    g = gold("teacher_forced.npz")
    _, logits = O.decode_steps(weights, O.GPTArch(), g["text_ids"].tolist(), 48,
                               forced_codes=g["forced_codes"].tolist(), return_logits=True)
    assert np.abs(logits.numpy() - g["logits"]).max() < 1e-5


@pytest.mark.parametrize("L", [1, 2, 3, 5, 10, 30, 90, 160])
def test_vocoder_full_chunks(gold, weights, L):
    g = gold("vocoder.npz")
    pcm = O.vocoder_decode(weights, g[f"codes_{L}"].tolist()).numpy()
    assert pcm.shape == (320 * L,)
    assert np.abs(pcm - g[f"pcm_{L}"]).max() < 2e-5
    assert snr_db(g[f"pcm_{L}"], pcm) > 90


@pytest.mark.parametrize("L", [270, 1280])
def test_vocoder_long_chunks_sliced(gold, weights, L):
    g = gold("vocoder.npz")
    pcm = O.vocoder_decode(weights, g[f"codes_{L}"].tolist()).numpy()
    m = len(pcm) // 2
    sl = np.concatenate([pcm[:2560], pcm[m - 1280:m + 1280], pcm[-2560:]])
    assert np.abs(sl - g[f"pcm_{L}"]).max() < 5e-5
    assert abs(np.sqrt((pcm.astype(np.float64) ** 2).mean()) - float(g[f"rms_{L}"])) < 1e-5


def test_vocoder_bandwidth_id(gold, weights):
    g = gold("vocoder.npz")
    pcm = O.vocoder_decode(weights, g["codes_bw2"].tolist(), bw=2).numpy()
    assert np.abs(pcm - g["pcm_bw2"]).max() < 2e-5
    pcm0 = O.vocoder_decode(weights, g["codes_bw2"].tolist(), bw=0).numpy()
    assert np.abs(pcm0 - g["pcm_bw2"]).max() > 1e-3      # the row select matters


def test_backbone_activations(gold, weights):
    g = gold("vocoder.npz")
    feats = O.codes_to_features(weights, torch.tensor([g["codes_30"].tolist()]))
    x = O.vocos_backbone(weights, feats, torch.tensor([0]))
    assert np.abs(x[0].numpy() - g["act30_backbone"]).max() < 2e-5


@pytest.mark.parametrize("name", ["no_eoa_r0", "no_eoa_r1", "eoa_mid", "eoa_on_boundary", "eoa_first"])
def test_scheduler_events(gold, name):
    """Replays scripted token streams: chunk lengths must equal what the reference loop put on its
    audio queue (positive = chunk of that many codes, -2 = switch-replica signal at sentence end)."""
    g = gold("scheduler.npz")
    script = g[name + "_script"].tolist()
    dump = int(g[name + "_dump"])
    events = []
    pos = 0
    while pos < len(script):
        chunks, used, dump = O.chunk_schedule(script[pos:], dump, stop_on_eoa=True)
        events += [len(c) for c in chunks]
        ended = used < len(script) - pos or script[pos + used - 1] == O.EOA_TOKEN_ID
        pos += used
        if ended:
            events.append(-2)
    assert events == g[name + "_events"].tolist()


def test_sampler_semantics():
    """src/model.py:397-406: temperature, top-k with ties kept, softmax, one draw."""
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(16, 4096, generator=g)
    logits[0, 100] = logits[0, 200] = logits[0].max() + 1.0          # a tie at the top
    u = torch.rand(16, generator=g)
    idx = O.sample_from_logits(logits, 0.8, 5, u)
    top5 = torch.topk(logits, 5).indices
    for b in range(16):
        assert idx[b] in top5[b]
    # k=1 with a tie keeps both tied entries
    seen = {int(O.sample_from_logits(logits[:1], 1.0, 1, torch.tensor([x]))[0]) for x in (0.01, 0.99)}
    assert seen == {100, 200}
    # u -> 0 picks the lowest surviving index; temperature -> 0 limit equals argmax
    assert int(O.sample_from_logits(logits[1:2], 1e-4, None, torch.tensor([0.5]))[0]) == int(logits[1].argmax())
    # distribution check against torch.multinomial on a small vocab
    lg = torch.tensor([[0.0, 1.0, 2.0, 3.0]])
    us = torch.rand(20000, generator=g)
    draws = O.sample_from_logits(lg.expand(20000, 4), 1.0, 3, us)
    freq = torch.bincount(draws, minlength=4).double() / 20000
    p = torch.softmax(torch.tensor([-float("inf"), 1.0, 2.0, 3.0]), 0).double()
    assert torch.allclose(freq, p, atol=0.015)
