// Header directory, built once at pack time (fei_corpus_load / load_raw / synth):
// for every record, one entry per header line that holds a colon, in line order -- the dict that utils.py:113-118
// builds (`for line in header.split("\n"): if ":" in line: key, value = line.split(":", 1);
// headers[key.strip()] = value.strip()`), minus the dict's collapsing of repeated keys, which depends on the queried
// field and stays in the scan (head_finish).  Keys are interned: the corpus keeps a dictionary of its distinct
// stripped key spellings (a few dozen in a real Memdir) and an entry names its key by dictionary slot, so a scan
// runs the key automaton once per distinct key (k_key_lut) instead of once per header line of every record, and then
// touches only the one value it needs.
#include "corpus.h"
#include "pyws.cuh"
#include <stdlib.h>

namespace fei {

// entry.x = key slot | val_len << 16, entry.y = val_off (offset of the stripped value from the start of the record's
// header text).  Headers longer than 65535 bytes get the single entry {~0, ~0}: "parse the text".
__device__ __forceinline__ unsigned long long key_hash(const uint8_t* p, uint32_t n) {
  unsigned long long h = 0xcbf29ce484222325ull;                       // FNV-1a, then a finaliser
  for (uint32_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
  h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
  return h | 1ull;                                                    // 0 marks an empty slot
}

struct KeyDict { unsigned long long* tag; unsigned long long* rep; uint32_t* len; uint32_t* flag; };

// slot of `h` (inserting it if new); ~0u when the table is over-full
__device__ __forceinline__ uint32_t key_slot(const KeyDict& d, unsigned long long h, bool insert) {
  uint32_t s = (uint32_t)(h >> 17) & (kKeySlots - 1);
  for (uint32_t probe = 0; probe < kKeySlots / 2; ++probe, s = (s + 1) & (kKeySlots - 1)) {
    unsigned long long t = d.tag[s];
    if (t == h) return s;
    if (t == 0) {
      if (!insert) return 0xFFFFFFFFu;
      t = atomicCAS(d.tag + s, 0ull, h);
      if (t == 0 || t == h) return s;
    }
  }
  return 0xFFFFFFFFu;
}

// kPass 0: count the entries of every record and intern the keys.  kPass 1: write the entries, checking every key
// against its slot's representative spelling (a 64-bit hash collision would set flag bit 1 and the directory is
// rebuilt as "parse the text" for everybody).
template <int kPass>
__global__ void __launch_bounds__(256) k_hdir(const uint8_t* __restrict__ hdr, const uint64_t* __restrict__ hdr_off, uint64_t n,
                                              uint32_t* __restrict__ cnt, const uint64_t* __restrict__ dir_off, uint2* __restrict__ dir,
                                              KeyDict kd, bool force_text) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* h = hdr + hdr_off[i];
  const uint64_t hlen = hdr_off[i + 1] - hdr_off[i];
  uint2* out = kPass ? dir + dir_off[i] : nullptr;
  if (hlen > 65535 || force_text) {
    if (kPass) out[0] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); else cnt[i] = 1;
    return;
  }
  uint32_t k = 0;
  const uint8_t* hend = h + hlen;
  const uint8_t* p = h;
  while (p < hend) {
    const uint8_t* eol = p; const uint8_t* colon = nullptr;
    while (eol < hend && *eol != '\n') { if (!colon && *eol == ':') colon = eol; ++eol; }
    if (colon) {
      const uint8_t* ka = p; const uint8_t* kb = colon; strip_span(ka, kb);
      const uint32_t klen = (uint32_t)(kb - ka);
      const unsigned long long kh = key_hash(ka, klen);
      if (!kPass) {
        const uint32_t s = key_slot(kd, kh, true);
        if (s == 0xFFFFFFFFu) atomicOr(kd.flag, 1u);
        else { atomicMin(kd.rep + s, (unsigned long long)(ka - hdr)); kd.len[s] = klen; }   // same hash => same length unless colliding (checked in pass 1)
      } else {
        const uint32_t s = key_slot(kd, kh, false);
        bool same = s != 0xFFFFFFFFu && kd.len[s] == klen;
        if (same) { const uint8_t* r = hdr + kd.rep[s]; for (uint32_t b = 0; same && b < klen; ++b) same = r[b] == ka[b]; }
        if (!same) atomicOr(kd.flag, 2u);
        const uint8_t* va = colon + 1; const uint8_t* vb = eol; strip_span(va, vb);
        out[k] = make_uint2((s & 0xFFFFu) | (uint32_t)(vb - va) << 16, (uint32_t)(va - h));
      }
      ++k;
    }
    p = eol + 1;
  }
  if (!kPass) cnt[i] = k;
}

int build_header_dir(fei_corpus* c, cudaStream_t s) {
  const uint64_t n = c->n;
  FEI_TRY(c->hdir_off.ensure((n + 1) * sizeof(uint64_t)));
  FEI_TRY(c->key_tag.ensure(kKeySlots * sizeof(unsigned long long)));
  FEI_TRY(c->key_rep.ensure(kKeySlots * sizeof(unsigned long long)));
  FEI_TRY(c->key_len.ensure((kKeySlots + 1) * sizeof(uint32_t)));          // [kKeySlots] = flag word
  if (n == 0) { FEI_CUDA(cudaMemsetAsync(c->hdir_off.p, 0, sizeof(uint64_t), s)); c->hdir_entries = 0; return FEI_OK; }
  DevBuf& cnt = c->tmp_len;
  FEI_TRY(cnt.ensure(n * sizeof(uint32_t)));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  // FEI_HDIR=0 (read at load time): mark every record "parse the text", so tests can hold the in-scan text parser
  // (normally only reached by > 64 KiB headers) against the directory path on the same corpus
  const char* env = getenv("FEI_HDIR");
  bool force_text = env && env[0] == '0';
  KeyDict kd{c->key_tag.as<unsigned long long>(), c->key_rep.as<unsigned long long>(), c->key_len.as<uint32_t>(), c->key_len.as<uint32_t>() + kKeySlots};
  for (int attempt = 0; attempt < 2; ++attempt) {
    FEI_CUDA(cudaMemsetAsync(kd.tag, 0, kKeySlots * sizeof(unsigned long long), s));
    FEI_CUDA(cudaMemsetAsync(kd.rep, 0xFF, kKeySlots * sizeof(unsigned long long), s));
    FEI_CUDA(cudaMemsetAsync(kd.len, 0, (kKeySlots + 1) * sizeof(uint32_t), s));
    k_hdir<0><<<blocks, 256, 0, s>>>(c->hdr.as<uint8_t>(), c->hdr_off.as<uint64_t>(), n, cnt.as<uint32_t>(), nullptr, nullptr, kd, force_text);
    FEI_TRY(exclusive_scan_u32_u64(cnt.as<uint32_t>(), n, c->hdir_off.as<uint64_t>(), c->scan_tmp, s));
    uint64_t total = 0;
    FEI_CUDA(cudaMemcpyAsync(&total, c->hdir_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, s));
    FEI_CUDA(cudaStreamSynchronize(s));
    c->hdir_entries = total;
    FEI_TRY(c->hdir.ensure((total + 1) * sizeof(uint2)));
    k_hdir<1><<<blocks, 256, 0, s>>>(c->hdr.as<uint8_t>(), c->hdr_off.as<uint64_t>(), n, nullptr, c->hdir_off.as<uint64_t>(), c->hdir.as<uint2>(), kd, force_text);
    uint32_t flag = 0;
    FEI_CUDA(cudaMemcpyAsync(&flag, kd.flag, 4, cudaMemcpyDeviceToHost, s));
    FEI_CUDA(cudaStreamSynchronize(s));
    FEI_CUDA(cudaGetLastError());
    if (flag == 0 || force_text) break;
    force_text = true;      // more than kKeySlots / 2 distinct keys, or a 64-bit hash collision: no directory for this corpus
  }
  return FEI_OK;
}

}  // namespace fei
