// Host-side helpers of the C-ABI that the data formats either side of the hot path need (SURVEY.md section 8 rows
// f1 / f3): CRC-32C (Castagnoli) of TFRecord frames and checkpoint tensors.  No device code.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "nabu_b200.h"

namespace {
struct Crc32cTables {
    uint32_t t[8][256];
    Crc32cTables() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
    }
};
const Crc32cTables kTables;
}  // namespace

// slicing-by-8: eight table lookups per 64-bit word
extern "C" unsigned int nabu_crc32c(const void* data, size_t nbytes, unsigned int crc) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    uint32_t c = ~crc;
    while (nbytes && (reinterpret_cast<uintptr_t>(p) & 7)) { c = kTables.t[0][(c ^ *p++) & 0xFF] ^ (c >> 8); --nbytes; }
    while (nbytes >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        w ^= c;
        c = kTables.t[7][w & 0xFF] ^ kTables.t[6][(w >> 8) & 0xFF] ^ kTables.t[5][(w >> 16) & 0xFF] ^
            kTables.t[4][(w >> 24) & 0xFF] ^ kTables.t[3][(w >> 32) & 0xFF] ^ kTables.t[2][(w >> 40) & 0xFF] ^
            kTables.t[1][(w >> 48) & 0xFF] ^ kTables.t[0][(w >> 56) & 0xFF];
        p += 8;
        nbytes -= 8;
    }
    while (nbytes--) c = kTables.t[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
    return ~c;
}
