"""oracle/cnn_oracle.py -- CPU ORACLE for the CNN forward.  TEST INFRASTRUCTURE ONLY.

The reference runs models/{CpG,CHG,CHH}.onnx through the OpenVINO 2025.4.0 CPU plugin
(src/app/hifimeth/mod_main.cpp:32-67, src/app/hifimeth/mod_batch.cpp:66-75).  OpenVINO is a third-party
dependency that is not under /root/reference (headers only) and cannot be installed offline, so this file
restates the published graph -- the ONNX files themselves, whose semantics are training/model_cnn.py:76-85:

    Transpose(0,2,1) -> BatchNormalization(eps 1e-5) -> 8 x [Conv1d(stride 2, pad 1) + bias -> ReLU]
    -> Flatten -> FC(128->256) -> ReLU -> FC(256->2)

in fp32 with torch CPU ops, reading the weights with a ~60-line protobuf reader (no `onnx` package here).
Post-processing follows s_logits_to_methy_probs, src/app/hifimeth/mod_batch.cpp:46-64.

PARITY UNPINNED at this boundary: the reference has no test or golden vector for the network output.  The pin
we do have: torch.jit.load(models/CpG.pt | CHH.pt) (the reference's own TorchScript exports, used by
src/app-gpu/hifimeth-gpu/5mc_call_gpu.cpp:48,199-207) agrees with this forward to <= 1e-5 on logits; golden
vectors from that run are committed under tests/golden/ (generator: oracle/make_golden.py).
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

# --------------------------------------------------------------------------------------------------------------
# Minimal protobuf wire-format reader (varint + length-delimited), enough for ONNX ModelProto.
# Field numbers: ModelProto.graph=7; GraphProto.node=1, initializer=5; NodeProto.input=1, output=2, op_type=4,
# attribute=5; AttributeProto.name=1, i=3, t=5, ints=8; TensorProto.dims=1, data_type=2, name=8, raw_data=9.
# --------------------------------------------------------------------------------------------------------------


def _varint(buf: bytes, pos: int):
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes):
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, val


def _tensor(buf: bytes):
    dims, name, raw, dtype = [], "", b"", 0
    floats = []
    for fno, wt, val in _fields(buf):
        if fno == 1:
            if wt == 0:
                dims.append(val)
            else:  # packed
                p = 0
                while p < len(val):
                    d, p = _varint(val, p)
                    dims.append(d)
        elif fno == 2:
            dtype = val
        elif fno == 8:
            name = val.decode()
        elif fno == 9:
            raw = val
        elif fno == 4:  # float_data (packed)
            floats.append(val)
    if dtype != 1:
        return name, None
    if raw:
        arr = np.frombuffer(raw, dtype="<f4")
    else:
        arr = np.frombuffer(b"".join(floats), dtype="<f4")
    return name, arr.reshape(dims).copy()


def read_onnx_graph(path):
    """Returns (nodes, tensors): nodes = [(op_type, inputs, outputs, attrs)], tensors = {name: ndarray}."""
    model = Path(path).read_bytes()
    graph = None
    for fno, _, val in _fields(model):
        if fno == 7:
            graph = val
    if graph is None:
        raise ValueError("no graph in ONNX file")
    nodes, tensors = [], {}
    for fno, _, val in _fields(graph):
        if fno == 5:
            name, arr = _tensor(val)
            if arr is not None:
                tensors[name] = arr
        elif fno == 1:
            op, ins, outs, attrs = "", [], [], {}
            for f2, _, v2 in _fields(val):
                if f2 == 1:
                    ins.append(v2.decode())
                elif f2 == 2:
                    outs.append(v2.decode())
                elif f2 == 4:
                    op = v2.decode()
                elif f2 == 5:
                    aname, aval = "", None
                    for f3, w3, v3 in _fields(v2):
                        if f3 == 1:
                            aname = v3.decode()
                        elif f3 == 3:
                            aval = v3
                        elif f3 == 2:
                            aval = struct.unpack("<f", v3)[0]
                        elif f3 == 5:
                            aval = _tensor(v3)[1]
                        elif f3 == 8:
                            if w3 == 0:
                                aval = (aval or []) + [v3]
                            else:
                                p, lst = 0, []
                                while p < len(v3):
                                    d, p = _varint(v3, p)
                                    lst.append(d)
                                aval = lst
                    attrs[aname] = aval
            nodes.append((op, ins, outs, attrs))
    return nodes, tensors


class CnnWeights:
    """bn0 (weight, bias, mean, var), 8 conv (W [Cout,Cin,K], b), fc1 (W [256,128], b), fc2 (W [2,256], b)."""

    def __init__(self, path):
        nodes, tensors = read_onnx_graph(path)
        const = dict(tensors)
        for op, ins, outs, attrs in nodes:
            if op == "Constant" and attrs.get("value") is not None:
                const[outs[0]] = attrs["value"]
        self.convs = []
        self.fcs = []
        self.bn0 = None
        self.bn_eps = 1e-5
        for op, ins, outs, attrs in nodes:
            if op == "BatchNormalization":
                self.bn0 = tuple(const[i] for i in ins[1:5])
                if "epsilon" in attrs and attrs["epsilon"] is not None:
                    self.bn_eps = float(attrs["epsilon"])
            elif op == "Conv":
                assert attrs.get("strides") == [2] and attrs.get("pads") == [1, 1], attrs
                self.convs.append((const[ins[1]], const[ins[2]]))
            elif op == "Gemm":
                w = const[ins[1]]
                if not attrs.get("transB"):
                    w = w.T.copy()
                self.fcs.append([w, const[ins[2]]])
            elif op == "MatMul":
                self.fcs.append([const[ins[1]].T.copy(), None])
            elif op == "Add" and self.fcs and self.fcs[-1][1] is None:
                b = const.get(ins[1], const.get(ins[0]))
                self.fcs[-1][1] = b
        assert self.bn0 is not None and len(self.convs) == 8 and len(self.fcs) == 2, (len(self.convs), len(self.fcs))
        self.conv1_k = self.convs[0][0].shape[2]


_CTX_FILES = {0: "CpG.onnx", 1: "CHG.onnx", 2: "CHH.onnx"}


def load_models(model_dir):
    return {c: CnnWeights(Path(model_dir) / f) for c, f in _CTX_FILES.items()}


def forward_logits(w: CnnWeights, feats: np.ndarray, return_acts: bool = False):
    """feats [B,401,8] f32 -> logits [B,2] f32 (fp32 torch CPU).  return_acts -> list of per-layer activations."""
    import torch
    import torch.nn.functional as F

    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(feats, dtype=np.float32)).permute(0, 2, 1)
        g, b, m, v = (torch.from_numpy(t) for t in w.bn0)
        x = F.batch_norm(x, m, v, g, b, training=False, eps=w.bn_eps)
        acts = [x]
        for cw, cb in w.convs:
            x = F.relu(F.conv1d(x, torch.from_numpy(cw), torch.from_numpy(cb), stride=2, padding=1))
            acts.append(x)
        x = torch.flatten(x, 1)
        x = F.relu(F.linear(x, torch.from_numpy(w.fcs[0][0]), torch.from_numpy(w.fcs[0][1])))
        acts.append(x)
        x = F.linear(x, torch.from_numpy(w.fcs[1][0]), torch.from_numpy(w.fcs[1][1]))
    if return_acts:
        return x.numpy(), [a.numpy() for a in acts]
    return x.numpy()


def logits_to_prob_ml(logits: np.ndarray):
    """mod_batch.cpp:46-64 in fp32: p1 = exp(v1-m)/(exp(v0-m)+exp(v1-m)); ML = min(255, (int)(255*p1))."""
    v = logits.astype(np.float32)
    m = np.maximum(v[:, 0], v[:, 1])
    e0 = np.exp((v[:, 0] - m).astype(np.float32)).astype(np.float32)
    e1 = np.exp((v[:, 1] - m).astype(np.float32)).astype(np.float32)
    p1 = (e1 / (e0 + e1).astype(np.float32)).astype(np.float32)
    ml = np.minimum(255, (np.float32(255) * p1).astype(np.int32)).astype(np.uint8)
    return p1, ml
