/*
 * chain_oracle.c — TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C CPU restatement of the reference's Memorychain link-hash path, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * as the checker for the CUDA path (fei_b200/csrc/chain.cu).  Nothing under fei_b200/
 * may link or call this file.
 *
 * Follows /root/reference/memdir_tools/memorychain.py:
 *   co_block_json   <- MemoryBlock.calculate_hash, json.dumps(..., sort_keys=True)   (:117-128)
 *   co_sha256       <- hashlib.sha256(block_string.encode()).hexdigest()             (:130)
 *   co_validate     <- MemoryChain.validate_chain                                    (:596-618)
 * SHA-256 itself is FIPS 180-4 (CPython's hashlib wraps OpenSSL; not in /root/reference).
 * Parity is pinned by tests/golden/chain_kats.json (hashes produced by importing the
 * reference in the build container, tests/golden/make_golden.py) and the FIPS "abc" vector.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------ SHA-256 */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))

static void sha_block(uint32_t st[8], const uint8_t* p) {
  uint32_t w[64], a, b, c, d, e, f, g, h;
  int t;
  for (t = 0; t < 16; ++t) w[t] = (uint32_t)p[4 * t] << 24 | (uint32_t)p[4 * t + 1] << 16 | (uint32_t)p[4 * t + 2] << 8 | p[4 * t + 3];
  for (t = 16; t < 64; ++t)
    w[t] = w[t - 16] + (ROR(w[t - 15], 7) ^ ROR(w[t - 15], 18) ^ (w[t - 15] >> 3)) + w[t - 7] + (ROR(w[t - 2], 17) ^ ROR(w[t - 2], 19) ^ (w[t - 2] >> 10));
  a = st[0]; b = st[1]; c = st[2]; d = st[3]; e = st[4]; f = st[5]; g = st[6]; h = st[7];
  for (t = 0; t < 64; ++t) {
    uint32_t t1 = h + (ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25)) + ((e & f) ^ (~e & g)) + K256[t] + w[t];
    uint32_t t2 = (ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

void co_sha256(const uint8_t* msg, uint64_t len, uint8_t out[32]) {
  uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  uint8_t tail[128];
  uint64_t full = len / 64, i, rem = len % 64, bits = len * 8;
  int k, tl;
  for (i = 0; i < full; ++i) sha_block(st, msg + 64 * i);
  memset(tail, 0, sizeof tail);
  memcpy(tail, msg + 64 * full, rem);
  tail[rem] = 0x80;
  tl = rem + 9 <= 64 ? 64 : 128;
  for (k = 0; k < 8; ++k) tail[tl - 1 - k] = (uint8_t)(bits >> (8 * k));
  sha_block(st, tail);
  if (tl == 128) sha_block(st, tail + 64);
  for (k = 0; k < 8; ++k) { out[4 * k] = st[k] >> 24; out[4 * k + 1] = st[k] >> 16; out[4 * k + 2] = st[k] >> 8; out[4 * k + 3] = st[k]; }
}

void co_sha256_hex(const uint8_t* msg, uint64_t len, char out[65]) {
  static const char hx[] = "0123456789abcdef";
  uint8_t d[32];
  int k;
  co_sha256(msg, len, d);
  for (k = 0; k < 32; ++k) { out[2 * k] = hx[d[k] >> 4]; out[2 * k + 1] = hx[d[k] & 15]; }
  out[64] = 0;
}

/* ------------------------------------------------------------------ canonical JSON */
/* A JSON scalar as json.dumps sees it. kind: 0 null, 1 str (utf-8), 2 int64, 3 float, 4 true, 5 false */
typedef struct co_val { int kind; const char* s; uint64_t slen; int64_t i; double f; } co_val;

typedef struct co_buf { char* p; size_t n, cap; } co_buf;
static void bput(co_buf* b, const char* s, size_t n) {
  if (b->n + n + 1 > b->cap) { b->cap = (b->n + n + 1) * 2; b->p = (char*)realloc(b->p, b->cap); }
  memcpy(b->p + b->n, s, n); b->n += n; b->p[b->n] = 0;
}
static void bputs(co_buf* b, const char* s) { bput(b, s, strlen(s)); }

/* ensure_ascii string escaping (json.encoder.py_encode_basestring_ascii) */
static void put_str(co_buf* b, const uint8_t* s, uint64_t n) {
  uint64_t i = 0; char t[16];
  bputs(b, "\"");
  while (i < n) {
    uint32_t c = s[i], cp; int len, k;
    if (c < 0x80) {
      ++i;
      if (c == '"') bputs(b, "\\\""); else if (c == '\\') bputs(b, "\\\\");
      else if (c == '\n') bputs(b, "\\n"); else if (c == '\r') bputs(b, "\\r"); else if (c == '\t') bputs(b, "\\t");
      else if (c == '\b') bputs(b, "\\b"); else if (c == '\f') bputs(b, "\\f");
      else if (c < 0x20 || c == 0x7f) { snprintf(t, sizeof t, "\\u%04x", c); bputs(b, t); }
      else { t[0] = (char)c; bput(b, t, 1); }
      continue;
    }
    if ((c & 0xE0) == 0xC0) { cp = c & 0x1F; len = 2; } else if ((c & 0xF0) == 0xE0) { cp = c & 0x0F; len = 3; } else { cp = c & 0x07; len = 4; }
    for (k = 1; k < len && i + k < n; ++k) cp = cp << 6 | (s[i + k] & 0x3F);
    i += len;
    if (cp >= 0x10000) { cp -= 0x10000; snprintf(t, sizeof t, "\\u%04x\\u%04x", 0xD800 | (cp >> 10), 0xDC00 | (cp & 0x3FF)); }
    else snprintf(t, sizeof t, "\\u%04x", cp);
    bputs(b, t);
  }
  bputs(b, "\"");
}

/* float.__repr__: shortest "%.{p}e" that round-trips, re-laid-out with Python's
 * fixed/exponent switch (exponent iff decpt <= -4 or decpt > 16). */
static void put_float(co_buf* b, double x) {
  char e[40], digits[24], out[64];
  int p, nd = 0, exp10, decpt, k, o = 0;
  char* q;
  if (isnan(x)) { bputs(b, "NaN"); return; }
  if (isinf(x)) { bputs(b, x > 0 ? "Infinity" : "-Infinity"); return; }
  for (p = 0; p < 17; ++p) { snprintf(e, sizeof e, "%.*e", p, x); if (strtod(e, NULL) == x) break; }
  q = e;
  if (*q == '-') { out[o++] = '-'; ++q; }
  for (; *q && *q != 'e'; ++q) if (*q != '.') digits[nd++] = *q;
  exp10 = atoi(q + 1);
  while (nd > 1 && digits[nd - 1] == '0') --nd;
  decpt = exp10 + 1;
  if (decpt <= -4 || decpt > 16) {
    out[o++] = digits[0];
    if (nd > 1) { out[o++] = '.'; memcpy(out + o, digits + 1, nd - 1); o += nd - 1; }
    o += snprintf(out + o, sizeof out - o, "e%c%02d", decpt - 1 < 0 ? '-' : '+', abs(decpt - 1));
  } else if (decpt <= 0) {
    out[o++] = '0'; out[o++] = '.';
    for (k = 0; k < -decpt; ++k) out[o++] = '0';
    memcpy(out + o, digits, nd); o += nd;
  } else if (decpt >= nd) {
    memcpy(out + o, digits, nd); o += nd;
    for (k = nd; k < decpt; ++k) out[o++] = '0';
    out[o++] = '.'; out[o++] = '0';
  } else {
    memcpy(out + o, digits, decpt); o += decpt; out[o++] = '.';
    memcpy(out + o, digits + decpt, nd - decpt); o += nd - decpt;
  }
  bput(b, out, o);
}

static void put_val(co_buf* b, const co_val* v) {
  char t[32];
  switch (v->kind) {
    case 0: bputs(b, "null"); break;
    case 1: put_str(b, (const uint8_t*)v->s, v->slen); break;
    case 2: snprintf(t, sizeof t, "%lld", (long long)v->i); bputs(b, t); break;
    case 3: put_float(b, v->f); break;
    case 4: bputs(b, "true"); break;
    default: bputs(b, "false"); break;
  }
}

/* vals: the ten hashed fields in sorted-key order:
 * difficulty, index, memory_id, nonce, previous_hash, proposer_node, responsible_node,
 * solver_node, task_state, timestamp.  Returns a malloc'd NUL-terminated text. */
char* co_block_json(const co_val vals[10], uint64_t* len_out) {
  static const char* keys[10] = {"difficulty", "index", "memory_id", "nonce", "previous_hash", "proposer_node", "responsible_node", "solver_node", "task_state", "timestamp"};
  co_buf b = {NULL, 0, 0};
  int k;
  bputs(&b, "{");
  for (k = 0; k < 10; ++k) {
    if (k) bputs(&b, ", ");
    bputs(&b, "\""); bputs(&b, keys[k]); bputs(&b, "\": ");
    put_val(&b, &vals[k]);
  }
  bputs(&b, "}");
  if (len_out) *len_out = b.n;
  return b.p;
}
void co_free(void* p) { free(p); }

/* ------------------------------------------------------------------ validate_chain */
/* blocks: n x 10 values; stored_hash / stored_hash_len: the block's .hash string.
 * Returns 1 if valid; else 0 with *first_bad = index and *kind = 1 (invalid hash) / 2 (broken link). */
int co_validate(const co_val* vals, const char* const* stored_hash, const uint64_t* stored_len, uint64_t n, int64_t* first_bad, int* kind) {
  uint64_t i;
  for (i = 1; i < n; ++i) {
    uint64_t len; char hex[65];
    char* text = co_block_json(vals + 10 * i, &len);
    const co_val* prev = &vals[10 * i + 4];
    co_sha256_hex((const uint8_t*)text, len, hex);
    free(text);
    if (stored_len[i] != 64 || memcmp(stored_hash[i], hex, 64) != 0) { *first_bad = (int64_t)i; *kind = 1; return 0; }
    if (prev->kind != 1 || prev->slen != stored_len[i - 1] || memcmp(prev->s, stored_hash[i - 1], prev->slen) != 0) { *first_bad = (int64_t)i; *kind = 2; return 0; }
  }
  *first_bad = -1; *kind = 0;
  return 1;
}
