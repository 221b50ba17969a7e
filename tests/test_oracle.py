"""CPU tests: the oracle restatement (oracle/hm_oracle.c, oracle/cnn_oracle.py) against
  (a) the committed golden vectors generated from the reference's own compiled code and TorchScript exports, and
  (b) the compiled reference itself (oracle/_ref) when it is present in this container.
Bit-exact for every integer / byte / feature quantity; logits within 1e-5 of the TorchScript models."""
import numpy as np
import pytest

from hifimeth_b200 import synth
from oracle import cnn_oracle, hmoracle
from oracle.make_golden import golden_bodies, golden_reads

from conftest import ROOT

O = hmoracle.oracle()


def _soa(reads, i):
    return synth.soa_from_reads([reads[i]])


def test_codev1_table_and_reencode():
    lut = synth.codev1_decode_table()
    for c in range(256):
        assert O.lib.hmo_codev1_decode(c) == lut[c]
        assert O.lib.hmo_codev1_encode(int(lut[c])) == c
    # every raw frame count maps to the largest code whose frames do not exceed it; clamps at 952
    for s in range(0, 1200):
        c = O.lib.hmo_codev1_encode(s)
        assert lut[c] <= min(s, 952)
        assert c == 255 or lut[c + 1] > min(s, 952)
    assert (O.decode_plane(np.arange(256, dtype=np.uint8)) == lut).all()


def test_golden_inputs_are_reproducible(golden):
    bodies = golden_bodies(golden_reads())
    assert int(golden["n_reads"]) == len(bodies)
    for i, b in enumerate(bodies):
        assert golden[f"body{i}"].tobytes() == b


def test_decode_scan_features_vs_golden(golden):
    reads = golden_reads()
    n_feat = 0
    for i, rd in enumerate(reads):
        if not golden[f"ok{i}"]:
            continue
        b = _soa(reads, i)
        l = len(rd["seq"])
        ok, fwd, rev = O.decode_seq(b.seq4, l, int(b.flag[0]))
        assert ok
        assert (fwd == golden[f"fwd{i}"]).all() and (rev == golden[f"rev{i}"]).all()
        for name, plane in (("fipd", b.fi), ("fpw", b.fp), ("ripd", b.ri), ("rpw", b.rp)):
            assert (O.decode_plane(plane) == golden[f"{name}{i}"]).all()
        for c in range(3):
            so = O.scan_sites(fwd, c)
            assert (so == golden[f"sites{i}_{c}"]).all()
            if f"sel{i}_{c}" not in golden:
                continue
            for k, j in enumerate(golden[f"sel{i}_{c}"]):
                s, f = O.site_features(fwd, rev, b.fi, b.fp, b.ri, b.rp, int(so[j]))
                assert s == golden[f"strand{i}_{c}"][k]
                assert (f.view(np.uint32) == golden[f"feat{i}_{c}"][k].view(np.uint32)).all()
                assert f[200, 1] == 1.0  # centre base is C on the chosen strand (training/sample_dataset.py:136)
                n_feat += 1
    assert n_feat > 100


def test_read_level_regroup_vs_golden(golden):
    reads = golden_reads()
    for i, rd in enumerate(reads):
        if not golden[f"ok{i}"] or (rd["seq"] > 3).any() or len(rd["seq"]) < 1000:
            continue
        b = _soa(reads, i)
        _, fwd, _ = O.decode_seq(b.seq4, len(rd["seq"]), int(b.flag[0]))
        q, ctx, nf, nr = O.read_calls(fwd, 7)
        assert (q[:nf] == golden[f"fq{i}"]).all() and (q[nf:] == golden[f"rq{i}"]).all()
        for c in range(3):
            assert (np.sort(q[ctx == c]) == golden[f"sites{i}_{c}"]).all()
        # context masks select subsets without changing the order
        q2, ctx2, nf2, nr2 = O.read_calls(fwd, 5)
        assert (q2 == q[ctx != 1]).all() and nf2 + nr2 == len(q2)


def test_mod_record_vs_golden(golden):
    reads = golden_reads()
    for i in range(len(reads)):
        body = golden[f"body{i}"].tobytes()
        if not golden[f"ok{i}"]:
            assert O.build_mod_record(body, False, [], [], [], []) == golden[f"mod{i}"].tobytes()
            continue
        fq, rq, fml, rml = (golden[f"{k}{i}"] for k in ("fq", "rq", "fml", "rml"))
        assert O.build_mod_record(body, False, fq, fml, rq, rml) == golden[f"mod{i}"].tobytes()
        assert O.build_mod_record(body, True, fq, fml, rq, rml) == golden[f"modkeep{i}"].tobytes()


def test_cnn_oracle_vs_golden(golden):
    models = cnn_oracle.load_models(ROOT / "models")
    assert models[0].conv1_k == 11 and models[1].conv1_k == 11 and models[2].conv1_k == 13
    n = 0
    for i in range(int(golden["n_reads"])):
        for c in range(3):
            if f"feat{i}_{c}" not in golden:
                continue
            lg = cnn_oracle.forward_logits(models[c], golden[f"feat{i}_{c}"])
            assert np.abs(lg - golden[f"logits{i}_{c}"]).max() < 1e-5
            if f"ptlogits{i}_{c}" in golden:  # reference TorchScript exports (CpG, CHH; CHG.pt is a different checkpoint)
                assert np.abs(lg - golden[f"ptlogits{i}_{c}"]).max() < 2e-5
            n += len(lg)
    assert n > 100


def test_softmax_quantise_is_truncation():
    lg = np.array([[0.0, 0.0], [5.0, -5.0], [-5.0, 5.0], [-30.0, 30.0], [1.0, 1.0000001]], np.float32)
    p, ml = cnn_oracle.logits_to_prob_ml(lg)
    assert ml[0] == 127 and ml[1] == 0 and ml[3] == 255
    for (v0, v1), pi, mi in zip(lg, p, ml):
        pc = O.lib.hmo_logits_to_prob(float(v0), float(v1))
        assert abs(pc - pi) < 1e-6
        assert abs(int(O.lib.hmo_prob_to_ml(pc)) - int(mi)) <= 1


@pytest.mark.skipif(not hmoracle.ref().available, reason="compiled reference (oracle/_ref) not present")
def test_oracle_vs_compiled_reference_fresh_inputs():
    R = hmoracle.ref()
    batch, reads = synth.make_reads(5, (1000, 3000), seed=77, flag_rev_every=2, uniform_codes=True)
    sites = O.batch_sites(batch, 7)
    lut = synth.codev1_decode_table()
    for r, rd in enumerate(reads):
        body = synth.record_body(rd)
        l = len(rd["seq"])
        ok, fwd, rev, k = R.query_decode(body, l)
        assert ok and (fwd == sites[r]["fwd"]).all() and (rev == sites[r]["rev"]).all()
        assert (k[0] == lut[rd["fi"]]).all() and (k[3] == lut[rd["rp"]]).all()
        b0, b1 = int(batch.base_off[r]), int(batch.base_off[r + 1])
        planes = [np.ascontiguousarray(a[b0:b1]) for a in (batch.fi, batch.fp, batch.ri, batch.rp)]
        for c in range(3):
            so = R.extract_sites(body, c, l)
            assert (so == O.scan_sites(fwd, c)).all()
            f, off, st = R.extract_features(body, c, 0, len(so))
            for j in range(0, len(so), 7):
                s, ff = O.site_features(fwd, rev, *planes, int(so[j]))
                assert s == st[j] and (ff.view(np.uint32) == f[j].view(np.uint32)).all()
        q, nf = sites[r]["qoff"], sites[r]["n_fwd"]
        ml = (np.arange(len(q)) * 11 % 256).astype(np.uint8)
        rec = R.build_mod_bam(body, False, q[:nf], ml[:nf], q[nf:], ml[nf:])
        assert rec == O.build_mod_record(body, False, q[:nf], ml[:nf], q[nf:], ml[nf:])
        # MM/ML round trip through the reference's own parser (src/corelib/bam_mod_parser.cpp:231-286)
        qq, ss, pp = R.parse_mods(rec, len(q) + 1)
        assert (qq == q).all() and (pp == ml).all() and (ss[:nf] == 0).all() and (ss[nf:] == 1).all()


def _forward_fp64_numpy(w, feats):
    """An independent statement of the graph (training/model_cnn.py:76-85): fp64, explicit zero padding and one einsum per conv
    tap -- no torch, no conv primitive, nothing shared with cnn_oracle.forward_logits but the parsed weights."""
    x = np.transpose(feats.astype(np.float64), (0, 2, 1))                      # [B, 8, 401]
    g, b, m, v = (t.astype(np.float64) for t in w.bn0)
    x = (x - m[None, :, None]) / np.sqrt(v[None, :, None] + w.bn_eps) * g[None, :, None] + b[None, :, None]
    for cw, cb in w.convs:
        cw = cw.astype(np.float64)
        k = cw.shape[2]
        xp = np.pad(x, ((0, 0), (0, 0), (1, 1)))                                # conv zero padding comes AFTER bn0
        n_out = (xp.shape[2] - k) // 2 + 1
        y = np.zeros((x.shape[0], cw.shape[0], n_out))
        for j in range(k):
            y += np.einsum("oc,bct->bot", cw[:, :, j], xp[:, :, j:j + 2 * n_out - 1:2])
        x = np.maximum(y + cb.astype(np.float64)[None, :, None], 0.0)
    x = x.reshape(x.shape[0], -1)                                              # channel-major flatten: idx = c * 2 + t
    x = np.maximum(x @ w.fcs[0][0].astype(np.float64).T + w.fcs[0][1], 0.0)
    return x @ w.fcs[1][0].astype(np.float64).T + w.fcs[1][1]


def test_cnn_oracle_against_independent_fp64_forward(golden):
    """CHG has no TorchScript pin (CHG.pt is another checkpoint): its oracle is pinned here, like the other two
    contexts, to an independent fp64 numpy forward of the same ONNX weights on the golden feature windows."""
    models = cnn_oracle.load_models(ROOT / "models")
    n = {0: 0, 1: 0, 2: 0}
    for i in range(int(golden["n_reads"])):
        for c in range(3):
            if f"feat{i}_{c}" not in golden:
                continue
            f = golden[f"feat{i}_{c}"]
            want = _forward_fp64_numpy(models[c], f)
            got = cnn_oracle.forward_logits(models[c], f)
            assert np.abs(got - want).max() < 2e-5, (i, c, float(np.abs(got - want).max()))
            assert np.abs(golden[f"logits{i}_{c}"] - want).max() < 2e-5
            n[c] += len(f)
    assert min(n.values()) > 20, n
