"""oracle/make_golden.py -- generates tests/golden/hotpath_v1.npz.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference): `python -m oracle.make_golden`.

Inputs : 8 seeded synthetic reads (hifimeth_b200.synth, seed 20260), ragged 1.1-2.4 kb, incl. one flag-0x10 read, one
         read below -l, one read with N bases, one read with B:S (raw frame) kinetics, one read lacking `rp`.
Outputs: produced by the REFERENCE'S OWN compiled code (oracle/_ref/libhifimeth_ref.so):
           decoded strands + kinetics, site offsets per context, feature tensors of selected sites,
           build_one_mod_bam record bytes;
         and by the reference's TorchScript exports models/CpG.pt, models/CHH.pt (torch.jit.load on CPU) plus the fp32
         ONNX-weight forward (oracle/cnn_oracle.py) for all three contexts: logits of the selected sites.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from hifimeth_b200 import synth  # noqa: E402
from oracle import cnn_oracle, hmoracle  # noqa: E402

SEED = 20260
N_SEL = 12  # feature/logit sites kept per (read, context)


def golden_reads():
    _, reads = synth.make_reads(8, (1100, 2400), SEED, flag_rev_every=3, short_every=7, n_every=97)
    reads[5]["rp"] = None  # kinetics-less read: must be passed through (src/corelib/bam_info.cpp:443-453)
    return reads


def golden_bodies(reads):
    return [synth.record_body(r, kinetics_as_u16=(i == 1), extra_mm=(i == 4)) for i, r in enumerate(reads)]


def main():
    import torch

    R = hmoracle.ref()
    assert R.available, "needs /root/reference"
    reads = golden_reads()
    bodies = golden_bodies(reads)
    out = {"n_reads": np.int32(len(reads)), "seed": np.int32(SEED)}
    models = cnn_oracle.load_models(ROOT / "models")
    pt = {0: torch.jit.load("/root/reference/models/CpG.pt").eval(), 2: torch.jit.load("/root/reference/models/CHH.pt").eval()}
    rng = np.random.default_rng(SEED)
    for i, (rd, body) in enumerate(zip(reads, bodies)):
        l = len(rd["seq"])
        out[f"body{i}"] = np.frombuffer(body, np.uint8)
        ok, fwd, rev, k = R.query_decode(body, l) if rd.get("rp") is not None else (False, None, None, None)
        out[f"ok{i}"] = np.uint8(ok)
        if not ok:
            stripped = R.build_mod_bam(body, False, [], [], [], [])
            out[f"mod{i}"] = np.frombuffer(stripped, np.uint8)
            continue
        out[f"fwd{i}"], out[f"rev{i}"] = fwd, rev
        for name, a in zip(("fipd", "fpw", "ripd", "rpw"), k):
            out[f"{name}{i}"] = a.astype(np.uint16)
        fq, rq = [], []
        has_n = bool((rd["seq"] > 3).any())
        for c in range(3):
            so = R.extract_sites(body, c, l)
            out[f"sites{i}_{c}"] = so
            if has_n or len(so) == 0:
                continue  # features on N are undefined behaviour in the reference
            sel = np.sort(rng.choice(len(so), size=min(N_SEL, len(so)), replace=False))
            sel[0] = 0
            sel[-1] = len(so) - 1  # always keep the two window-clipped ends
            f_all, off, st = R.extract_features(body, c, 0, len(so))
            out[f"sel{i}_{c}"] = sel.astype(np.int32)
            out[f"feat{i}_{c}"] = f_all[sel]
            out[f"strand{i}_{c}"] = st[sel].astype(np.uint8)
            lg = cnn_oracle.forward_logits(models[c], f_all[sel])
            out[f"logits{i}_{c}"] = lg
            if c in pt:
                with torch.no_grad():
                    out[f"ptlogits{i}_{c}"] = pt[c](torch.from_numpy(f_all[sel])).numpy()
            for o, s in zip(off, st):
                (fq if s == 0 else rq).append(int(o))
        fq, rq = np.sort(np.array(fq, np.int32)), np.sort(np.array(rq, np.int32))
        fml = ((fq * 7 + 3) % 256).astype(np.uint8)
        rml = ((rq * 5 + 1) % 256).astype(np.uint8)
        if l < 1000 or has_n:
            fq = rq = np.zeros(0, np.int32)
            fml = rml = np.zeros(0, np.uint8)
        out[f"fq{i}"], out[f"rq{i}"], out[f"fml{i}"], out[f"rml{i}"] = fq, rq, fml, rml
        out[f"mod{i}"] = np.frombuffer(R.build_mod_bam(body, False, fq, fml, rq, rml), np.uint8)
        out[f"modkeep{i}"] = np.frombuffer(R.build_mod_bam(body, True, fq, fml, rq, rml), np.uint8)
    dst = ROOT / "tests" / "golden" / "hotpath_v1.npz"
    dst.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(dst, **out)
    print(dst, dst.stat().st_size, "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
