// Host canonical-JSON serialiser for Memorychain blocks (chain_json.cpp).
#pragma once
#include <stdint.h>
#include <vector>
#include "../../include/feiscan.h"

namespace fei {
// msgs = concatenated canonical JSON texts, off[n+1] = offsets.  Returns FEI_OK or FEI_E_*.
int serialize_chain_cols(const fei_json_col* cols, uint64_t n, std::vector<uint8_t>& msgs, std::vector<uint64_t>& off);
}
