"""Synthetic HiFi reads of the shape BASELINE.json names.

Unaligned records (flag 4, tid -1), names ``synth/<i>/ccs``, SEQ i.i.d. uniform over ACGT (no N: the reference
reads out of bounds on N, src/app/hifimeth/eval_kmer_features.cpp:43-44), QUAL constant 40, kinetics tags
fi, ri, fp, rp as B:C CodecV1 codes drawn i.i.d. from floor(Exp(mean 30)) clipped to [0,255], plus the
passthrough tags np:i rq:f zm:i RG:Z.  Everything is generated with numpy's PCG64 from an explicit seed.

Two views of the same reads are produced:
  * ``ReadBatch``   -- the engine's struct-of-arrays staging layout (include/hm_engine.h: hm_read_batch)
  * record bodies   -- BAM alignment records without block_size, for the BAM writer and the reference driver
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

_NIB = np.array([1, 2, 4, 8], dtype=np.uint8)  # A C G T -> BAM 4-bit codes


@dataclass
class ReadBatch:
    """SoA staging layout.  All arrays are C-contiguous numpy arrays."""

    n_reads: int
    base_off: np.ndarray  # uint32 [n_reads+1] prefix sum of l_qseq: offset of each read in the kinetics planes
    seq_off: np.ndarray  # uint32 [n_reads+1] byte offset of each read's packed SEQ in seq4
    seq4: np.ndarray  # uint8 packed 4-bit SEQ exactly as BAM stores it (high nibble first)
    flag: np.ndarray  # uint16 [n_reads]
    valid: np.ndarray  # uint8 [n_reads] 1 = call this read, 0 = pass through
    fi: np.ndarray  # uint8 [n_bases] forward IPD codes, forward coordinates
    fp: np.ndarray  # uint8 [n_bases] forward PW codes
    ri: np.ndarray  # uint8 [n_bases] reverse IPD codes, reverse-strand coordinates
    rp: np.ndarray  # uint8 [n_bases] reverse PW codes

    @property
    def n_bases(self) -> int:
        return int(self.base_off[-1])

    def read_len(self, r: int) -> int:
        return int(self.base_off[r + 1] - self.base_off[r])


def pack_seq(codes: np.ndarray) -> np.ndarray:
    """codes: uint8 array of 0..3 (A C G T) or 4 (N) -> BAM packed nibbles."""
    lut = np.array([1, 2, 4, 8, 15], dtype=np.uint8)
    nib = lut[codes]
    if len(nib) & 1:
        nib = np.concatenate([nib, np.zeros(1, np.uint8)])
    return ((nib[0::2] << 4) | nib[1::2]).astype(np.uint8)


def make_reads(n_reads: int, read_len, seed: int, *, flag_rev_every: int = 0, n_every: int = 0,
               short_every: int = 0, uniform_codes: bool = False, min_read_len: int = 1000):
    """Returns (ReadBatch, list of per-read dicts).  read_len: int or (lo, hi) for ragged lengths.

    flag_rev_every k>0 sets flag 0x10 on every k-th read; n_every k>0 plants an N every k bases in every
    4th read; short_every k>0 makes every k-th read 300 bases long (below -l, must be passed through)."""
    rng = np.random.default_rng(seed)
    reads = []
    for i in range(n_reads):
        if isinstance(read_len, (tuple, list)):
            l = int(rng.integers(read_len[0], read_len[1] + 1))
        else:
            l = int(read_len)
        if short_every and i % short_every == short_every - 1:
            l = 300
        seq = rng.integers(0, 4, size=l, dtype=np.uint8)
        if n_every and i % 4 == 3:
            seq[n_every - 1::n_every] = 4
        if uniform_codes:
            k = rng.integers(0, 256, size=(4, l), dtype=np.uint8)
        else:
            k = np.minimum(np.floor(rng.exponential(30.0, size=(4, l))), 255).astype(np.uint8)
        flag = 4 | (16 if (flag_rev_every and i % flag_rev_every == flag_rev_every - 1) else 0)
        reads.append(dict(name=f"synth/{i}/ccs", seq=seq, fi=k[0], ri=k[1], fp=k[2], rp=k[3], flag=flag,
                          np=int(rng.integers(3, 40)), rq=float(np.float32(0.99 + 0.00999 * rng.random())), zm=i))
    return soa_from_reads(reads, min_read_len=min_read_len), reads


def soa_from_reads(reads, min_read_len: int = 1000) -> ReadBatch:
    n = len(reads)
    lens = np.array([len(r["seq"]) for r in reads], dtype=np.int64)
    base_off = np.zeros(n + 1, dtype=np.uint32)
    base_off[1:] = np.cumsum(lens)
    seq_bytes = (lens + 1) // 2
    seq_off = np.zeros(n + 1, dtype=np.uint32)
    seq_off[1:] = np.cumsum(seq_bytes)
    seq4 = np.concatenate([pack_seq(r["seq"]) for r in reads]) if n else np.zeros(0, np.uint8)

    def cat(k):  # a missing or wrong-length plane is staged as zeros (the read is invalid anyway)
        parts = [r[k] if (r.get(k) is not None and len(r[k]) == len(r["seq"])) else np.zeros(len(r["seq"]), np.uint8) for r in reads]
        return (np.concatenate(parts) if n else np.zeros(0, np.uint8)).astype(np.uint8)

    valid = np.array([1 if (len(r["seq"]) >= min_read_len and all(r.get(k) is not None and len(r[k]) == len(r["seq"])
                                                                 for k in ("fi", "ri", "fp", "rp"))) else 0
                      for r in reads], dtype=np.uint8)
    return ReadBatch(n_reads=n, base_off=base_off, seq_off=seq_off, seq4=np.ascontiguousarray(seq4),
                     flag=np.array([r["flag"] for r in reads], dtype=np.uint16), valid=valid,
                     fi=cat("fi"), fp=cat("fp"), ri=cat("ri"), rp=cat("rp"))


def record_body(r, *, kinetics_as_u16: bool = False, extra_mm: bool = False) -> bytes:
    """One BAM alignment record body (SAMv1 s4.2, without block_size) for a synthetic read."""
    name = r["name"].encode() + b"\0"
    l = len(r["seq"])
    core = struct.pack("<iiBBHHHiiii", -1, -1, len(name), 255, 4680, 0, r["flag"], l, -1, -1, 0)
    out = [core, name, pack_seq(r["seq"]).tobytes(), bytes([40]) * l]
    out.append(b"npi" + struct.pack("<i", r["np"]))
    out.append(b"rqf" + struct.pack("<f", r["rq"]))
    if extra_mm:  # stale tags that build_one_mod_bam must strip (src/corelib/build_mod_bam.cpp:100-108)
        out.append(b"MMZ" + b"C+m,1;" + b"\0")
        out.append(b"MLBC" + struct.pack("<I", 1) + bytes([7]))
    for tag in ("fi", "ri", "fp", "rp"):
        v = r.get(tag)
        if v is None:
            continue
        if kinetics_as_u16:
            # raw frames that re-encode (src/corelib/bam_info.cpp:455-478) to the same code: table value plus the
            # largest remainder the code's step size absorbs
            frames = (codev1_decode_table()[v] + (1 << (v >> 6)) - 1).astype("<u2")
            out.append(tag.encode() + b"BS" + struct.pack("<I", len(v)) + frames.tobytes())
        else:
            out.append(tag.encode() + b"BC" + struct.pack("<I", len(v)) + v.astype(np.uint8).tobytes())
    out.append(b"zmi" + struct.pack("<i", r["zm"]))
    out.append(b"RGZ" + b"synth\0")
    return b"".join(out)


def codev1_decode_table() -> np.ndarray:
    """PacBio CodecV1 8-bit code -> frames (the 256-entry table of src/corelib/bam_info.cpp:562-570)."""
    c = np.arange(256, dtype=np.int32)
    return np.where(c < 64, c, np.where(c < 128, (c - 64) * 2 + 64, np.where(c < 192, (c - 128) * 4 + 192, (c - 192) * 8 + 448))).astype(np.int32)


# ---- BAM files (BGZF per SAMv1 section 4) for the `call` driver tests and benchmarks -------------------------------------

def write_bam(path, bodies, header_text: str = "@HD\tVN:1.6\tSO:unknown\n@RG\tID:synth\tPL:PACBIO\n", level: int = 1,
              block: int = 0xff00) -> int:
    """Writes unaligned records (bodies without block_size) as a BGZF-compressed BAM file.  Returns the file size."""
    import zlib

    text = header_text.encode()
    stream = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", 0))
    for b in bodies:
        stream += struct.pack("<i", len(b)) + b
    out = bytearray()
    for off in range(0, len(stream), block):
        chunk = bytes(stream[off:off + block])
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        out += struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + b"BC" + struct.pack("<HH", 2, len(comp) + 25)
        out += comp + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
    out += bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    with open(path, "wb") as f:
        f.write(out)
    return len(out)


def read_bam(path):
    """Returns (header_text, reference_block_bytes, [record bodies]) of a BAM file (any gzip-member layout)."""
    import gzip

    data = gzip.decompress(open(path, "rb").read())
    assert data[:4] == b"BAM\1", "bad BAM magic"
    l_text = struct.unpack_from("<i", data, 4)[0]
    text = data[8:8 + l_text].rstrip(b"\0").decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", data, p)[0]
    q = p + 4
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, q)[0]
        q += 4 + l_name + 4
    refs = data[p:q]
    bodies = []
    while q < len(data):
        n = struct.unpack_from("<i", data, q)[0]
        bodies.append(data[q + 4:q + 4 + n])
        q += 4 + n
    return text, refs, bodies
