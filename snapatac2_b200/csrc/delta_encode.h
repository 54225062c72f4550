// Delta encoding of CSR column indices for the host -> device transfer (host side; the decoder is
// decode_deltas_kernel in ingest.cu).  Plain C++ + AVX2, no CUDA: also compiled into
// scripts/host_bw_probe.cpp, which measures it against the box's memory bandwidth.
//
// Column indices are ascending within a row, and consecutive columns of a 5000-entry row of a
// 500k-column matrix are ~100 apart: the difference to the previous stored index fits 16 bits almost
// always, whatever the row boundaries (a row start usually gives a negative difference).  The staging
// team therefore ships 2 bytes per entry instead of 4 -- half the pinned-buffer writes, half the DMA
// reads, half the PCIe bytes -- and a small kernel rebuilds the int32 indices:
//     delta[i] = idx[i] - idx[i-1]      if 0 <= difference < 0xFFFF and i is not the first of a tile
//     delta[i] = 0xFFFF  (marker)       otherwise; the absolute index goes to a side list, in order
// A tile is kDeltaTile entries; tiles decode independently (segmented prefix sum inside one CTA).
// Chunk layout in a ring slot: header | deltas u16[n] (padded to 32 B) | first side entry of every
// tile u32[tiles + 1] | side int32[n_side].
#pragma once

#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>

#include <algorithm>

namespace snapb {

constexpr int kDeltaTile = 2048;
constexpr int64_t kDeltaPer = 3 << 20;     // entries per chunk: 6 MB of deltas in an 8 MB slot
struct DeltaHeader { uint32_t n_entries, n_tiles, n_side, pad; };

inline size_t delta_bytes(int64_t n) { return (static_cast<size_t>(n) * 2 + 31) & ~static_cast<size_t>(31); }

// One group of up to 8 entries the slow way, into `tmp`; markers appended to side[ns...].
template <typename T>
inline void delta_group_scalar(const T* src, int64_t i0, int cnt, bool first_is_marker, uint16_t* tmp, int32_t* side, int64_t& ns) {
    for (int q = 0; q < cnt; ++q) {
        const int64_t j = i0 + q;
        const int64_t dd = (q == 0 && first_is_marker) ? -1 : static_cast<int64_t>(src[j]) - static_cast<int64_t>(src[j - 1]);
        if (dd >= 0 && dd < 0xFFFF) {
            tmp[q] = static_cast<uint16_t>(dd);
        } else {
            tmp[q] = 0xFFFFu;
            side[ns++] = static_cast<int32_t>(src[j]);
        }
    }
}

// One tile: differences straight into the pinned slot with non-temporal 16-byte stores (no
// read-for-ownership of the slot), one aligned group of 8 entries at a time.  `dst` is 16-byte aligned.
// Returns the number of side entries written; `acc_out` collects the OR of every index (bits >= 31 set
// = a negative or too large index somewhere).
template <typename T>
inline int64_t encode_tile(const T* src, int64_t cnt, uint16_t* dst, int32_t* side, uint64_t& acc_out) {
    int64_t ns = 0;
    uint64_t acc = 0;
    alignas(16) uint16_t tmp[8];
    int64_t i = 0;
#if defined(__AVX2__)
    if (cnt >= 8) {
        delta_group_scalar(src, 0, 8, true, tmp, side, ns);
        for (int q = 0; q < 8; ++q) acc |= static_cast<uint64_t>(static_cast<int64_t>(src[q]));
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst), _mm_load_si128(reinterpret_cast<const __m128i*>(tmp)));
        i = 8;
    }
    const __m128i ffff = _mm_set1_epi16(-1);
    if constexpr (sizeof(T) == 8) {
        const int64_t* s64 = reinterpret_cast<const int64_t*>(src);
        const __m256i high = _mm256_set1_epi64x(~static_cast<int64_t>(0xFFFF));
        const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
        __m256i vor = _mm256_setzero_si256();
        for (; i + 8 <= cnt; i += 8) {
            const __m256i c0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s64 + i));
            const __m256i c1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s64 + i + 4));
            const __m256i p0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s64 + i - 1));
            const __m256i p1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s64 + i + 3));
            const __m256i d0 = _mm256_sub_epi64(c0, p0), d1 = _mm256_sub_epi64(c1, p1);
            vor = _mm256_or_si256(vor, _mm256_or_si256(c0, c1));
            // low words of the 8 differences -> 8 x u32 -> 8 x u16
            const __m256i lo = _mm256_permutevar8x32_epi32(d0, pick);
            const __m256i hi = _mm256_permutevar8x32_epi32(d1, pick);
            const __m256i v32 = _mm256_permute2x128_si256(lo, hi, 0x20);
            const __m256i v16 = _mm256_packus_epi32(v32, v32);          // per 128-bit lane: 4 words, twice
            const __m128i out = _mm256_castsi256_si128(_mm256_permute4x64_epi64(v16, 0x08));
            // fast path: every difference in [0, 0xFFFF): no bit above 15 anywhere, no word equal to the marker
            if (_mm256_testz_si256(_mm256_or_si256(d0, d1), high) && _mm_testz_si128(_mm_cmpeq_epi16(out, ffff), ffff)) {
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), out);
            } else {
                delta_group_scalar(src, i, 8, false, tmp, side, ns);
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_load_si128(reinterpret_cast<const __m128i*>(tmp)));
            }
        }
        alignas(32) uint64_t t4[4];
        _mm256_store_si256(reinterpret_cast<__m256i*>(t4), vor);
        acc |= t4[0] | t4[1] | t4[2] | t4[3];
    } else {
        const int32_t* s32 = reinterpret_cast<const int32_t*>(src);
        const __m256i high = _mm256_set1_epi32(~0xFFFF);
        __m256i vor = _mm256_setzero_si256();
        for (; i + 8 <= cnt; i += 8) {
            const __m256i c0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s32 + i));
            const __m256i p0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s32 + i - 1));
            const __m256i d0 = _mm256_sub_epi32(c0, p0);
            vor = _mm256_or_si256(vor, c0);
            const __m256i v16 = _mm256_packus_epi32(d0, d0);
            const __m128i out = _mm256_castsi256_si128(_mm256_permute4x64_epi64(v16, 0x08));
            if (_mm256_testz_si256(d0, high) && _mm_testz_si128(_mm_cmpeq_epi16(out, ffff), ffff)) {
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), out);
            } else {
                delta_group_scalar(src, i, 8, false, tmp, side, ns);
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_load_si128(reinterpret_cast<const __m128i*>(tmp)));
            }
        }
        alignas(32) uint32_t t8[8];
        _mm256_store_si256(reinterpret_cast<__m256i*>(t8), vor);
        uint32_t o32 = 0;
        for (int q = 0; q < 8; ++q) o32 |= t8[q];
        acc |= static_cast<uint64_t>(static_cast<int64_t>(static_cast<int32_t>(o32)));
    }
#endif
    // what is left: fewer than 8 entries (last tile of a chunk), or everything without AVX2
    while (i < cnt) {
        const int g = static_cast<int>(std::min<int64_t>(8, cnt - i));
        delta_group_scalar(src, i, g, i == 0, tmp, side, ns);
        for (int q = 0; q < g; ++q) {
            acc |= static_cast<uint64_t>(static_cast<int64_t>(src[i + q]));
            dst[i + q] = tmp[q];
        }
        i += g;
    }
    acc_out |= acc;
    return ns;
}

// One chunk of `len` <= kDeltaPer indices into `slot` (64-byte aligned, `slot_bytes` long).  False = the
// side list does not fit (nothing but wide gaps): the caller ships plain int32 instead.
template <typename T>
inline bool encode_deltas(const T* src, int64_t len, unsigned char* slot, size_t slot_bytes, uint64_t& orall, size_t& used) {
    DeltaHeader* hd = reinterpret_cast<DeltaHeader*>(slot);
    const int64_t n_tiles = (len + kDeltaTile - 1) / kDeltaTile;
    uint16_t* d16 = reinterpret_cast<uint16_t*>(slot + sizeof(DeltaHeader));      // 16 B header + 4096 B per tile: 16-byte aligned
    uint32_t* tile_side = reinterpret_cast<uint32_t*>(slot + sizeof(DeltaHeader) + delta_bytes(len));
    int32_t* side = reinterpret_cast<int32_t*>(tile_side + n_tiles + 1);
    const int64_t side_cap = (static_cast<int64_t>(slot_bytes) - (reinterpret_cast<unsigned char*>(side) - slot)) / 4;
    int64_t ns = 0;
    for (int64_t t = 0; t < n_tiles; ++t) {
        const int64_t a = t * kDeltaTile, cnt = std::min<int64_t>(len - a, kDeltaTile);
        tile_side[t] = static_cast<uint32_t>(ns);
        if (ns + cnt > side_cap) return false;
        ns += encode_tile(src + a, cnt, d16 + a, side + ns, orall);
    }
#if defined(__AVX2__)
    _mm_sfence();
#endif
    tile_side[n_tiles] = static_cast<uint32_t>(ns);
    hd->n_entries = static_cast<uint32_t>(len);
    hd->n_tiles = static_cast<uint32_t>(n_tiles);
    hd->n_side = static_cast<uint32_t>(ns);
    hd->pad = 0;
    used = static_cast<size_t>(reinterpret_cast<unsigned char*>(side + ns) - slot);
    return true;
}

// Scalar replay of a chunk (test infrastructure: the host-only self-test and the probe).
inline bool decode_deltas_host(const unsigned char* slot, int64_t len, int32_t* out) {
    const DeltaHeader* hd = reinterpret_cast<const DeltaHeader*>(slot);
    const uint16_t* d16 = reinterpret_cast<const uint16_t*>(slot + sizeof(DeltaHeader));
    const uint32_t* tile_side = reinterpret_cast<const uint32_t*>(slot + sizeof(DeltaHeader) + delta_bytes(len));
    const int32_t* side = reinterpret_cast<const int32_t*>(tile_side + hd->n_tiles + 1);
    if (hd->n_entries != len || tile_side[hd->n_tiles] != hd->n_side) return false;
    for (uint32_t t = 0; t < hd->n_tiles; ++t) {
        int64_t k = tile_side[t];
        int32_t v = 0;
        const int64_t a = static_cast<int64_t>(t) * kDeltaTile, b = std::min<int64_t>(len, a + kDeltaTile);
        for (int64_t i = a; i < b; ++i) {
            if (d16[i] == 0xFFFFu) v = side[k++];
            else v += d16[i];
            out[i] = v;
        }
        if (k != tile_side[t + 1]) return false;
    }
    return true;
}

}  // namespace snapb
