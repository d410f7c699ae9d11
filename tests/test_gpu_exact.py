"""Exact mode (LVX_PRECISION_EXACT): bf16-rounded, LayerNorm-folded weights + fp32-class activations (bf16 hi | lo
operand pairs, fp32 KV cache) on the tcgen05 tensor cores.  Contract (BASELINE.json north_star: "bit-identical greedy
tokens versus the reference", pick at streaming_server.py:342-346): EVERY greedy token equals the reference's fp32 loop
(the oracle, pinned to the unmodified reference by tests/test_oracle_golden.py) run on the same rounded weights
(llmvox_b200.weights.fold_round_gpt_weights), and teacher-forced logits stay within the north star's fp32 bound 1e-4."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from llmvox_b200 import weights as W  # noqa: E402
from oracle import llmvox_oracle as O  # noqa: E402

SENT = ("the quick brown fox jumps over the lazy dog while seven small birds "
        "sing a very old song near the river.")


@pytest.fixture(scope="module")
def folded(weights):
    return W.fold_round_gpt_weights(weights)


def _margins(logits):
    top2 = torch.topk(logits, 2, dim=-1).values
    return (top2[:, 0] - top2[:, 1]).numpy()


def _report(tag, got, ref, logits):
    """First divergence + margin histogram of the oracle's picks (printed with -s / on failure)."""
    m = _margins(logits)
    hist = np.histogram(m, bins=[0, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 10])[0].tolist()
    div = next((t for t in range(len(ref)) if got[t] != ref[t]), None)
    msg = f"{tag}: {len(ref)} steps, first divergence {div}, min margin {m.min():.3g}, margin histogram (<1e-5..>=1e-1) {hist}"
    print(msg)
    return div, msg


def _path(path):
    """decode path argument of a test variant: the engine's choice, or the 8-CTA cut of the cluster kernel forced"""
    from llmvox_b200 import _lib
    return _lib.PATH_CLUSTER8 if path == "cluster8" else _lib.PATH_AUTO


def _engine(weights, path, n, ctx):
    import os
    from llmvox_b200.engine import Engine
    if path == "per_op":
        os.environ["LLMVOX_B200_CLUSTER"] = "0"          # read at engine creation
    try:
        return Engine(weights, device=0, precision="exact", max_sessions=n, max_context=ctx, max_vocode_frames=256)
    finally:
        os.environ.pop("LLMVOX_B200_CLUSTER", None)


@pytest.mark.parametrize("path", ["cluster", "cluster8", "per_op"])
def test_exact_greedy_tokens_identical_config0(weights, folded, gold, path):
    """BASELINE config 0: the 20-word sentence, all 130 steps of the fixture's length, every token."""
    g = gold("config0_loop.npz")
    ids, n = g["text_ids"].tolist(), int(g["n_steps"])
    ref, logits = O.decode_steps(folded, O.GPTArch(), ids, n, return_logits=True)
    e = _engine(weights, path, 2, 256)
    e.open([1])
    e.feed_text([1], [ids])
    l0 = e.kernel_launches
    e.decode_steps([1], n, path=_path(path))
    launches = e.kernel_launches - l0
    # the path under test really ran: one cluster-kernel launch (+ the page-table patch, stream packing) against ~40 kernels per step
    assert (launches <= 12) if path != "per_op" else (launches > 30 * n), launches
    got = e.gather_codes([1], 0, n).cpu().numpy()[0].tolist()
    e.close()
    div, msg = _report(f"exact/{path} config0", got, ref, logits)
    assert div is None, msg
    assert got == ref


@pytest.mark.parametrize("path", ["cluster", "cluster8", "per_op"])
def test_exact_greedy_tokens_identical_config1_streams(weights, folded, path):
    """BASELINE config 1 shape: 8 of the concurrent streams x 200 steps each, every token of every stream, decoded as
    one batch (and, on the cluster kernel, cut into launches of 10/30/90/70 iterations like the chunk schedule)."""
    rng = np.random.RandomState(77)
    texts = [rng.randint(3, 259, size=rng.randint(30, 200)).tolist() for _ in range(8)]
    n = 200
    e = _engine(weights, path, 8, 208)
    slots = list(range(8))
    e.open(slots)
    e.feed_text(slots, texts)
    for k in (10, 30, 90, 70):
        e.decode_steps(slots, k, path=_path(path))
    got = e.gather_codes(slots, 0, n).cpu().numpy()
    e.close()
    bad = []
    for i in range(8):
        ref, logits = O.decode_steps(folded, O.GPTArch(), texts[i], n, return_logits=True)
        div, msg = _report(f"exact/{path} stream {i}", got[i].tolist(), ref, logits)
        if div is not None:
            bad.append(msg)
    assert not bad, "\n".join(bad)


def test_exact_teacher_forced_logits(weights, folded, gold):
    """Teacher-forced logits of the kernel-per-op exact path against the oracle on the folded weights: the north star's
    fp32 bound (1e-4), on tensor cores."""
    g = gold("teacher_forced.npz")
    ids, forced = g["text_ids"].tolist(), g["forced_codes"].tolist()
    _, ref = O.decode_steps(folded, O.GPTArch(), ids, 48, forced_codes=forced, return_logits=True)
    e = _engine(weights, "per_op", 2, 64)
    f = torch.tensor(forced, dtype=torch.int32, device="cuda")
    e.open([0])
    e.feed_text([0], [ids])
    rows = []
    for t in range(48):
        logits, codes = e.decode_step_logits([0], forced=f[t:t + 1])
        rows.append(logits[0].cpu())
        assert int(codes[0]) == int(logits[0].argmax())
    e.close()
    err = float((torch.stack(rows) - ref).abs().max())
    print(f"exact per-op teacher-forced max |logit error| {err:.3g}")
    assert err < 1e-4, err


def test_exact_folds_nontrivial_layernorm_weights(weights):
    """The reference initialises LayerNorm weights to ones; a trained checkpoint does not.  With perturbed LayerNorm
    weights the engine's fold (W * g, then bf16) must still reproduce the oracle on fold_round_gpt_weights."""
    g = torch.Generator().manual_seed(5)
    sd = dict(weights)
    for k in list(sd):
        if k.startswith("transformer.") and (".ln_" in k or "ln_f" in k) and k.endswith(".weight"):
            sd[k] = 1.0 + 0.2 * torch.randn(sd[k].shape, generator=g)
    folded = W.fold_round_gpt_weights(sd)
    ids = O.word_ids("layernorm", True)
    n = 60
    ref, logits = O.decode_steps(folded, O.GPTArch(), ids, n, return_logits=True)
    for path in ("cluster", "per_op"):
        e = _engine(sd, path, 2, 64)
        e.open([0])
        e.feed_text([0], [ids])
        e.decode_steps([0], n)
        got = e.gather_codes([0], 0, n).cpu().numpy()[0].tolist()
        e.close()
        div, msg = _report(f"exact/{path} perturbed LN", got, ref, logits)
        assert div is None, msg


def test_exact_mode_vocoder_is_the_bf16_vocoder(weights, gold):
    """Exact mode changes the decoder only: chunks are vocoded by the bf16 tensor-core vocoder (>= 40 dB)."""
    g = gold("vocoder.npz")
    e = _engine(weights, "cluster", 2, 64)
    codes = torch.from_numpy(g["codes_30"].astype(np.int32)).cuda() if "codes_30" in g.files else None
    if codes is None:
        pytest.skip("fixture has no 30-code chunk")
    pcm = e.vocode(codes, [0, 30]).cpu().numpy()
    e.close()
    ref = g["pcm_30"].astype(np.float64)
    snr = 10 * np.log10((ref ** 2).sum() / ((ref - pcm) ** 2).sum())
    assert snr > 40, snr


def test_exact_long_context_2000_steps(weights, folded):
    """Contexts beyond 1024 tokens on the cluster-resident kernel (page-table windows of 64 pages, reloaded per 1024
    tokens; reference block_size 8192, src/model.py:205): 2,000 greedy steps of one session against the oracle on the
    folded weights, every token, and three more sessions against the kernel-per-op exact path (all tokens)."""
    n = 2000
    ids = O.word_ids("long", False) + O.word_ids("context.", True)
    ref, logits = O.decode_steps(folded, O.GPTArch(), ids, n, return_logits=True)
    rng = np.random.RandomState(4)
    texts = [ids] + [rng.randint(3, 259, size=k).tolist() for k in (0, 300, 1500)]
    slots = [0, 1, 2, 3]
    out = {}
    for path in ("cluster", "cluster8", "per_op"):
        e = _engine(weights, path, 4, 2048)
        e.open(slots)
        e.feed_text(slots, texts)
        for k in (1000, 30, 970):                   # launches that start below, straddle and start above the 1024-token window
            e.decode_steps(slots, k, path=_path(path))
        out[path] = e.gather_codes(slots, 0, n).cpu().numpy()
        assert e.session_length(0) == n
        e.close()
    div, msg = _report("exact/cluster 2000 steps", out["cluster"][0].tolist(), ref, logits)
    assert div is None, msg
    div, msg = _report("exact/cluster8 2000 steps", out["cluster8"][0].tolist(), ref, logits)
    assert div is None, msg
    for i in range(4):
        for path in ("cluster", "cluster8"):
            same = out[path][i] == out["per_op"][i]
            assert same.all(), (path, i, int(np.argmin(same)))
