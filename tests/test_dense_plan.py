"""CPU check of the dilated dense plan (tests/dense_emulator.py, mirrored in hifimeth_b200/csrc/cnn_tensor.cu) against the
per-site oracle: the plan must reproduce oracle logits for interior sites, window-clipped sites at both read ends, and
both strands, for both conv1 kernel sizes that ship (11: CpG/CHG, 13: CHH).  Also pins the bf16 hi/lo split arithmetic
the tensor-core kernel uses: three products per term keep the probability within 1e-3 (north_star tolerance)."""
import numpy as np
import pytest

import dense_emulator as de
from hifimeth_b200 import synth
from oracle import cnn_oracle, hmoracle

from conftest import ROOT


@pytest.fixture(scope="module")
def models():
    return cnn_oracle.load_models(ROOT / "models")


@pytest.mark.parametrize("ctx", [0, 1, 2])
def test_plan_matches_per_site_oracle(models, ctx):
    O = hmoracle.oracle()
    batch, _ = synth.make_reads(1, 1300, seed=70 + ctx)
    sites = O.batch_sites(batch, 7)[0]
    fwd, L = sites["fwd"], len(sites["fwd"])
    plan = de.build_plan(models[ctx])
    sel = np.nonzero(sites["ctx"] == ctx)[0]
    rng = np.random.default_rng(ctx)
    idx = np.unique(np.r_[0, 1, len(sel) - 2, len(sel) - 1, rng.integers(0, len(sel), 24)])
    want = cnn_oracle.forward_logits(models[ctx], O.batch_features(batch, [sites], 0, sel[idx]))
    logit = {}
    for strand in (0, 1):
        X = de.place(de.strand_features(fwd, batch.fi, batch.fp, batch.ri, batch.rp, strand))
        logit[strand] = de.run_plan(plan, X)["LOGIT"]
    got = np.zeros_like(want)
    for k, p in enumerate(sites["qoff"][sel[idx]]):
        strand = 0 if fwd[p] == 1 else 1
        got[k] = logit[strand][de.site_row(int(p) if strand == 0 else L - 1 - int(p))]
    assert np.abs(got - want).max() < 5e-5


def test_bf16_split_precision(models):
    import torch

    def bf(x):
        return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).to(torch.float32).numpy()

    O = hmoracle.oracle()
    batch, _ = synth.make_reads(1, 1200, seed=5)
    fwd = O.batch_sites(batch, 7)[0]["fwd"]
    X = de.place(de.strand_features(fwd, batch.fi, batch.fp, batch.ri, batch.rp, 0))
    plan = de.build_plan(models[2])
    ref = de.run_plan(plan, X)["LOGIT"]
    maps = {"X": X}
    rows = X.shape[0]
    for op in plan:
        acc = np.tile(op.bias, (rows, 1)).astype(np.float32)
        for t in op.terms:
            a = np.zeros_like(maps[t.src])
            a[:rows - t.shift] = maps[t.src][t.shift:]
            ah, wh = bf(a), bf(t.w)
            acc += ah @ wh + bf(a - ah) @ wh + ah @ bf(t.w - wh)
        maps[op.out] = np.maximum(acc, 0) if op.relu else acc
    need = slice(de.site_row(0), de.site_row(len(fwd)))
    p_ref, _ = cnn_oracle.logits_to_prob_ml(ref[need])
    p_got, _ = cnn_oracle.logits_to_prob_ml(maps["LOGIT"][need])
    assert np.abs(p_ref - p_got).max() < 2e-4
