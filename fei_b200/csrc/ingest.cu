// Raw ingest ("next" row 1 of SURVEY.md 8(f)): file bytes -> packed corpus, on the GPU.
//
// Replaces the per-file text work of utils.list_memories + parse_memory_content
// (memdir_tools/utils.py:229-232, :97-120) that the host packer otherwise does in Python:
//   open(path, "r").read()      -> strict UTF-8 validation (an undecodable file is reported and
//                                  skipped, utils.py:247-248) + universal-newline translation
//                                  ("\r\n" and lone "\r" become "\n")
//   content.split("---", 1)     -> first "---" anywhere in the text (utils.py:105)
//   body.strip()                -> Python's str.strip() whitespace set (utils.py:109,120)
// plus the record bits the scan needs (no separator / non-ASCII / context-dependent str.lower()).
// The host keeps the directory walk, the file-name grammar and the listing order.
//
// k_raw_measure : one thread per file: validate + measure the normalised header / stripped body.
// k_raw_write   : one thread per valid file: write normalised header text and stripped body into
//                 the canonical blobs (offsets from an exclusive scan over the valid files).
#include "corpus.h"
#include "pyws.cuh"
#include <vector>
#include <string.h>

namespace fei {

// Strict UTF-8 (what bytes.decode("utf-8") accepts): no overlongs, no surrogates, <= U+10FFFF.
__device__ bool utf8_valid(const uint8_t* p, const uint8_t* end, bool& ascii, bool& has_sigma, bool& has_idot) {
  ascii = true; has_sigma = false; has_idot = false;
  while (p < end) {
    uint32_t c = *p;
    if (c < 0x80) { ++p; continue; }
    ascii = false;
    if (c >= 0xC2 && c <= 0xDF) {
      if (end - p < 2 || (p[1] & 0xC0) != 0x80) return false;
      if (c == 0xC4 && p[1] == 0xB0) has_idot = true;         // U+0130: lower() is two characters (handled by the automata)
      if (c == 0xCE && p[1] == 0xA3) has_sigma = true;        // U+03A3: lower() depends on the context (final sigma)
      p += 2;
    } else if (c >= 0xE0 && c <= 0xEF) {
      if (end - p < 3 || (p[1] & 0xC0) != 0x80 || (p[2] & 0xC0) != 0x80) return false;
      if (c == 0xE0 && p[1] < 0xA0) return false;             // overlong
      if (c == 0xED && p[1] >= 0xA0) return false;            // surrogates
      p += 3;
    } else if (c >= 0xF0 && c <= 0xF4) {
      if (end - p < 4 || (p[1] & 0xC0) != 0x80 || (p[2] & 0xC0) != 0x80 || (p[3] & 0xC0) != 0x80) return false;
      if (c == 0xF0 && p[1] < 0x90) return false;             // overlong
      if (c == 0xF4 && p[1] >= 0x90) return false;            // > U+10FFFF
      p += 4;
    } else {
      return false;
    }
  }
  return true;
}

// All spans are in RAW bytes.  Universal-newline translation only touches '\r' ("\r\n" -> "\n", lone "\r" -> "\n"),
// never '-' and never the whitespace-ness of a character, so the "---" search and the strip can run on the raw bytes;
// the translated length of a span is its raw length minus the number of "\r\n" pairs inside it (spans start and end at
// non-whitespace bytes or at the separator, so no pair is ever cut).
struct RawMeasure {
  uint32_t hdr_raw_len;               // header = raw[0, hdr_raw_len)
  uint32_t body_raw_begin, body_raw_end;
  uint32_t hdr_len, body_len;         // translated lengths
  uint32_t flags;                     // bit0 valid UTF-8, bit1 has separator, bit2 non-ASCII, bit3 U+03A3 present, bit4 U+0130 present
};

__device__ __forceinline__ uint32_t count_crlf(const uint8_t* p, const uint8_t* end) {
  uint32_t n = 0;
  for (; p + 1 < end; ++p) if (p[0] == '\r' && p[1] == '\n') ++n;
  return n;
}

__global__ void k_raw_measure(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ raw_off, uint64_t n, RawMeasure* __restrict__ out,
                              uint32_t* __restrict__ hdr_len, uint32_t* __restrict__ body_len) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* p = raw + raw_off[i];
  const uint8_t* end = raw + raw_off[i + 1];
  RawMeasure m{0, 0, 0, 0, 0, 0};
  bool ascii, sigma, idot;
  if (utf8_valid(p, end, ascii, sigma, idot)) {
    m.flags = 1u | (ascii ? 0u : 4u) | (sigma ? 8u : 0u) | (idot ? 16u : 0u);
    const uint8_t* sep = nullptr;                              // first "---" anywhere (utils.py:105)
    for (const uint8_t* q = p; q + 2 < end; ++q) if (q[0] == '-' && q[1] == '-' && q[2] == '-') { sep = q; break; }
    const uint8_t* ba = sep ? sep + 3 : p;                     // no separator: the whole text is the body (utils.py:107-109)
    const uint8_t* bb = end;
    strip_span(ba, bb);
    if (sep) { m.flags |= 2u; m.hdr_raw_len = (uint32_t)(sep - p); m.hdr_len = m.hdr_raw_len - count_crlf(p, sep + 1); }
    m.body_raw_begin = (uint32_t)(ba - p); m.body_raw_end = (uint32_t)(bb - p);
    m.body_len = (uint32_t)(bb - ba) - count_crlf(ba, bb);
  }
  out[i] = m;
  hdr_len[i] = m.hdr_len;
  body_len[i] = m.body_len;
}

__device__ __forceinline__ void copy_translated(const uint8_t* p, const uint8_t* end, const uint8_t* hard_end, uint8_t* dst) {
  while (p < end) {
    uint8_t c = *p++;
    if (c == '\r') { if (p < hard_end && *p == '\n') ++p; c = '\n'; }
    *dst++ = c;
  }
}

__global__ void k_raw_write(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ raw_off, uint64_t n, const RawMeasure* __restrict__ ms,
                            const uint64_t* __restrict__ hdr_off, const uint64_t* __restrict__ body_off,
                            uint8_t* __restrict__ hdr, uint8_t* __restrict__ body) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  RawMeasure m = ms[i];
  if (!(m.flags & 1u)) return;
  const uint8_t* p = raw + raw_off[i];
  const uint8_t* end = raw + raw_off[i + 1];
  copy_translated(p, p + m.hdr_raw_len, end, hdr + hdr_off[i]);
  copy_translated(p + m.body_raw_begin, p + m.body_raw_end, end, body + body_off[i]);
}

__global__ void k_fix_fsb(const RawMeasure* __restrict__ ms, uint64_t n, uint32_t* __restrict__ fsb) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t f = ms[i].flags;
  uint32_t bits = ((f & 2u) ? 0u : FEI_REC_NO_SEPARATOR) | ((f & 4u) ? FEI_REC_NONASCII : 0u) | ((f & 8u) ? FEI_REC_HAS_SIGMA : 0u) | ((f & 16u) ? FEI_REC_HAS_IDOT : 0u);
  fsb[i] = (fsb[i] & 0x00FFFFFFu) | (bits << 24);
}

}  // namespace fei

using namespace fei;

// All n files must be valid records to be loaded; the call first reports validity so the host can drop the
// undecodable ones (and print the reference's message) and call again with the survivors.
extern "C" int fei_corpus_load_raw(fei_corpus* c, const fei_corpus_host* h, const uint8_t* raw, const uint64_t* raw_off, uint8_t* valid_out) {
  if (!c) { set_error("null corpus"); return FEI_E_BADARG; }
  std::lock_guard<std::mutex> lock(c->mu);
  FEI_TRY(require_ready());
  if (!c || !h || !raw_off || (h->n && !raw)) { set_error("null argument"); return FEI_E_BADARG; }
  if (h->n >= 0xFFFFFFFFull) { set_error("at most 2^32-2 records per shard"); return FEI_E_BADARG; }
  if (h->n && (!h->ts || !h->wall || !h->flags8 || !h->fsb)) { set_error("missing meta array"); return FEI_E_BADARG; }
  // loads run on the copy stream: a batch can be uploaded / normalised / tiled into one handle while another handle is being
  // scanned on the compute stream (streaming e2e use); the staging buffers live in the handle (grow-only) because a
  // cudaFree in the middle of a pipeline synchronises the whole device
  cudaStream_t s = corpus_load_stream(c);
  uint64_t n = h->n;
  c->loaded = false;
  uint64_t raw_bytes = n ? raw_off[n] : 0;
  DevBuf& d_raw = c->stage_raw; DevBuf& d_raw_off = c->stage_raw_off; DevBuf& d_ms = c->stage_ms; DevBuf& d_hlen = c->stage_hlen; DevBuf& d_blen = c->stage_blen;
  FEI_TRY(d_raw.ensure(raw_bytes + 64)); FEI_TRY(d_raw_off.ensure((n + 1) * 8));
  FEI_TRY(d_ms.ensure((n ? n : 1) * sizeof(RawMeasure))); FEI_TRY(d_hlen.ensure((n ? n : 1) * 4)); FEI_TRY(d_blen.ensure((n ? n : 1) * 4));
  if (raw_bytes) FEI_CUDA(cudaMemcpyAsync(d_raw.p, raw, raw_bytes, cudaMemcpyHostToDevice, s));
  FEI_CUDA(cudaMemsetAsync((uint8_t*)d_raw.p + raw_bytes, 0, 64, s));
  FEI_CUDA(cudaMemcpyAsync(d_raw_off.p, raw_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
  unsigned g = (unsigned)((n + 127) / 128);
  if (n) k_raw_measure<<<g, 128, 0, s>>>(d_raw.as<uint8_t>(), d_raw_off.as<uint64_t>(), n, d_ms.as<RawMeasure>(), d_hlen.as<uint32_t>(), d_blen.as<uint32_t>());
  std::vector<RawMeasure> ms(n ? n : 1);
  if (n) FEI_CUDA(cudaMemcpyAsync(ms.data(), d_ms.p, n * sizeof(RawMeasure), cudaMemcpyDeviceToHost, s));
  FEI_CUDA(cudaStreamSynchronize(s));
  bool all_valid = true;
  for (uint64_t i = 0; i < n; ++i) { bool v = ms[i].flags & 1u; if (valid_out) valid_out[i] = v; all_valid &= v; }
  if (!all_valid) { set_error("some files are not valid UTF-8 (see valid_out); drop them and load again"); return FEI_E_BADARG; }
  // offsets of the normalised pieces
  c->n = n; c->global_base = h->global_base;
  FEI_TRY(c->hdr_off.ensure((n + 1) * 8));
  DevBuf& body_off = c->stage_body_off; DevBuf& body = c->stage_body;
  FEI_TRY(body_off.ensure((n + 1) * 8));
  FEI_TRY(exclusive_scan_u32_u64(d_hlen.as<uint32_t>(), n, c->hdr_off.as<uint64_t>(), c->scan_tmp, s));
  FEI_TRY(exclusive_scan_u32_u64(d_blen.as<uint32_t>(), n, body_off.as<uint64_t>(), c->scan_tmp, s));
  uint64_t hb = 0, bb = 0;
  FEI_CUDA(cudaMemcpyAsync(&hb, c->hdr_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, s));
  FEI_CUDA(cudaMemcpyAsync(&bb, body_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, s));
  FEI_CUDA(cudaStreamSynchronize(s));
  c->hdr_bytes = hb; c->body_bytes = bb;
  FEI_TRY(c->hdr.ensure(hb + 64)); FEI_TRY(body.ensure(bb + 64));
  FEI_CUDA(cudaMemsetAsync((uint8_t*)c->hdr.p + hb, 0, 48, s));
  FEI_CUDA(cudaMemsetAsync((uint8_t*)body.p + bb, 0, 48, s));
  if (n) k_raw_write<<<g, 128, 0, s>>>(d_raw.as<uint8_t>(), d_raw_off.as<uint64_t>(), n, d_ms.as<RawMeasure>(), c->hdr_off.as<uint64_t>(), body_off.as<uint64_t>(),
                                       c->hdr.as<uint8_t>(), body.as<uint8_t>());
  // meta columns + names come from the host (file-name grammar and listing order stay there)
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> int {
    FEI_TRY(b.ensure(bytes + 16));
    if (bytes) FEI_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s));
    return FEI_OK;
  };
  FEI_TRY(up(c->ts, h->ts, n * 8)); FEI_TRY(up(c->wall, h->wall, n * 8)); FEI_TRY(up(c->flags8, h->flags8, n * 8)); FEI_TRY(up(c->fsb, h->fsb, n * 4));
  if (n) k_fix_fsb<<<g, 128, 0, s>>>(d_ms.as<RawMeasure>(), n, c->fsb.as<uint32_t>());
  if (h->name && h->name_off && h->name_spans && n) {
    c->name_bytes = h->name_off[n];
    FEI_TRY(up(c->name, h->name, c->name_bytes)); FEI_TRY(up(c->name_off, h->name_off, (n + 1) * 8)); FEI_TRY(up(c->name_spans, h->name_spans, n * 8));
  } else { c->name.release(); c->name_off.release(); c->name_spans.release(); c->name_bytes = 0; }
  for (uint64_t i = 0; i < n; ++i)
    if (ms[i].body_len > (32u << 20)) { set_error("record %llu: body larger than 32 MiB is not supported", (unsigned long long)i); return FEI_E_UNSUPPORTED; }
  FEI_CUDA(cudaGetLastError());
  FEI_TRY(build_tiles(c, body.as<uint8_t>(), body_off.as<uint64_t>(), s));
  FEI_TRY(build_header_dir(c, s));
  if (body.bytes > (8ull << 30)) { body.release(); body_off.release(); c->tmp_len.release(); c->tmp_gunits.release(); d_raw.release(); }
  c->loaded = true;
  return FEI_OK;
}
