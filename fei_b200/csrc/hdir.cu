// Header directory, built once at pack time (fei_corpus_load / load_raw / synth):
// for every record, the stripped (key span, value span) of each header line that holds a colon, in line order --
// the dict that utils.py:113-118 builds (`for line in header.split("\n"): if ":" in line: key, value =
// line.split(":", 1); headers[key.strip()] = value.strip()`), minus the dict's collapsing of repeated keys, which
// depends on the queried field and stays in the scan (head_finish).  A scan then touches the few key bytes and the
// one value it needs instead of walking every header byte twice (line split + strip, then the automata).
#include "corpus.h"
#include "pyws.cuh"
#include <stdlib.h>

namespace fei {

// entry.x = key_off | key_len << 16, entry.y = val_off | val_len << 16 (offsets from the start of the record's
// header text, after strip).  Headers longer than 65535 bytes get the single entry {~0, ~0}: "parse the text".
template <bool kWrite>
__global__ void __launch_bounds__(256) k_hdir(const uint8_t* __restrict__ hdr, const uint64_t* __restrict__ hdr_off, uint64_t n,
                                              uint32_t* __restrict__ cnt, const uint64_t* __restrict__ dir_off, uint2* __restrict__ dir, bool force_text) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* h = hdr + hdr_off[i];
  const uint64_t hlen = hdr_off[i + 1] - hdr_off[i];
  uint2* out = kWrite ? dir + dir_off[i] : nullptr;
  if (hlen > 65535 || force_text) {
    if (kWrite) out[0] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); else cnt[i] = 1;
    return;
  }
  uint32_t k = 0;
  const uint8_t* hend = h + hlen;
  const uint8_t* p = h;
  while (p < hend) {
    const uint8_t* eol = p; const uint8_t* colon = nullptr;
    while (eol < hend && *eol != '\n') { if (!colon && *eol == ':') colon = eol; ++eol; }
    if (colon) {
      if (kWrite) {
        const uint8_t* ka = p; const uint8_t* kb = colon; strip_span(ka, kb);
        const uint8_t* va = colon + 1; const uint8_t* vb = eol; strip_span(va, vb);
        out[k] = make_uint2((uint32_t)(ka - h) | (uint32_t)(kb - ka) << 16, (uint32_t)(va - h) | (uint32_t)(vb - va) << 16);
      }
      ++k;
    }
    p = eol + 1;
  }
  if (!kWrite) cnt[i] = k;
}

int build_header_dir(fei_corpus* c, cudaStream_t s) {
  const uint64_t n = c->n;
  FEI_TRY(c->hdir_off.ensure((n + 1) * sizeof(uint64_t)));
  if (n == 0) { FEI_CUDA(cudaMemsetAsync(c->hdir_off.p, 0, sizeof(uint64_t), s)); c->hdir_entries = 0; return FEI_OK; }
  DevBuf& cnt = c->tmp_len;
  FEI_TRY(cnt.ensure(n * sizeof(uint32_t)));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  // FEI_HDIR=0 (read at load time): mark every record "parse the text", so tests can hold the in-scan text parser
  // (normally only reached by > 64 KiB headers) against the directory path on the same corpus
  const char* env = getenv("FEI_HDIR");
  const bool force_text = env && env[0] == '0';
  k_hdir<false><<<blocks, 256, 0, s>>>(c->hdr.as<uint8_t>(), c->hdr_off.as<uint64_t>(), n, cnt.as<uint32_t>(), nullptr, nullptr, force_text);
  FEI_TRY(exclusive_scan_u32_u64(cnt.as<uint32_t>(), n, c->hdir_off.as<uint64_t>(), c->scan_tmp, s));
  uint64_t total = 0;
  FEI_CUDA(cudaMemcpyAsync(&total, c->hdir_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, s));
  FEI_CUDA(cudaStreamSynchronize(s));
  c->hdir_entries = total;
  FEI_TRY(c->hdir.ensure((total + 1) * sizeof(uint2)));
  k_hdir<true><<<blocks, 256, 0, s>>>(c->hdr.as<uint8_t>(), c->hdr_off.as<uint64_t>(), n, nullptr, c->hdir_off.as<uint64_t>(), c->hdir.as<uint2>(), force_text);
  FEI_CUDA(cudaGetLastError());
  return FEI_OK;
}

}  // namespace fei
