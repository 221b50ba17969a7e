// fast_deflate.h -- raw DEFLATE (RFC 1951) for BGZF payloads of at most 64 KiB, written for the records `call` moves:
// kinetics codes, packed bases and qualities -- noisy bytes with runs, in which LZ77 matching at a distance finds next to nothing.
//
// Why not zlib for these two calls: with 4 host cores per GPU the one-process queue (call_main.cpp) is bound by zlib;
// inflate and deflate are >= 80 % of the host time per read.  The reference goes through htslib (sam_read1 / sam_write1 over bgzf,
// src/corelib/sam_batch.hpp:38-54, src/app/hifimeth/mod_main.cpp:353-362), which has the same bound.
//
//   hm_deflate_rle   dynamic-Huffman blocks of 16 KiB input over literals + run matches (distance 1, what zlib calls Z_RLE), each
//                    stored instead when that is not smaller.  Any inflater reads it.
//   hm_inflate_fast  complete inflater (stored, fixed and dynamic blocks, any distances): 64-bit bit buffer refilled once per symbol,
//                    one table lookup per symbol (11-bit first level + second level for the long codes).  Input comes from files:
//                    every read and write is bounds-checked, and a payload it rejects is handed to zlib by the caller for the
//                    verdict (bgzf_bam.cpp), so an error message never depends on this decoder alone.
#pragma once
#include <cstddef>
#include <cstdint>

namespace hm {

// Upper bound of hm_deflate_rle's output for n input bytes (n <= 65535).
inline size_t hm_deflate_rle_bound(size_t n) { return n + 64; }  // 4 blocks x 6 bytes if all are stored + 8 bytes of store slack

// Compresses in[0, n) (n <= 65535) into out[0, cap), cap >= hm_deflate_rle_bound(n).  Returns the number of bytes written.
size_t hm_deflate_rle(const uint8_t* in, size_t n, uint8_t* out, size_t cap);

// Inflates the raw DEFLATE stream in[0, n_in) into out[0, n_out).  true only if the stream is well formed, ends with its final
// block inside the input and produces exactly n_out bytes.  Never reads outside in[0, n_in) or writes outside out[0, n_out).
bool hm_inflate_fast(const uint8_t* in, size_t n_in, uint8_t* out, size_t n_out);

// CRC-32 of the gzip trailer (reflected 0x04C11DB7), continuing from `crc` like zlib's crc32().  On x86-64 with PCLMULQDQ the bulk is
// folded 64 bytes per iteration with carry-less multiplies (Gopal et al., "Fast CRC Computation for Generic Polynomials Using
// PCLMULQDQ Instruction", Intel 2009); the last 16-byte remainder and the tail go through zlib's table code.  Elsewhere: zlib.
uint32_t hm_crc32(uint32_t crc, const uint8_t* data, size_t n);

}  // namespace hm
