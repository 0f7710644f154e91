// TEMPORARY stubs (removed as the real implementations land).
#include "common.h"
#define STUB(name, ...) extern "C" int name(__VA_ARGS__) { fei::set_error(#name " not implemented yet"); return FEI_E_UNSUPPORTED; }
STUB(fei_corpus_create, fei_corpus**)
STUB(fei_corpus_destroy, fei_corpus*)
STUB(fei_corpus_load, fei_corpus*, const fei_corpus_host*)
STUB(fei_corpus_synth, fei_corpus*, uint64_t, uint64_t, uint64_t)
STUB(fei_corpus_stats_get, const fei_corpus*, fei_corpus_stats*)
STUB(fei_corpus_fetch, fei_corpus*, uint64_t, uint64_t, uint8_t*, uint64_t, uint64_t*, uint8_t*, uint64_t, uint64_t*, int64_t*, int64_t*, uint64_t*, uint32_t*)
STUB(fei_scan_masks, fei_corpus*, const uint8_t*, uint64_t, uint32_t*)
STUB(fei_scan_hits, fei_corpus*, const uint8_t*, uint64_t, uint64_t* const*, const uint64_t*, uint64_t*)
STUB(fei_scan_count, fei_corpus*, const uint8_t*, uint64_t, uint64_t*)
STUB(fei_scan_last_timing, const fei_corpus*, fei_scan_timing*)
STUB(fei_comm_unique_id, uint8_t*)
STUB(fei_comm_init, const uint8_t*, int, int)
STUB(fei_comm_destroy, void)
STUB(fei_comm_allgather_hits, fei_corpus*, uint32_t, uint64_t* const*, const uint64_t*, uint64_t*, uint64_t*)
STUB(fei_comm_allreduce_first_bad, int64_t*, int32_t*)
