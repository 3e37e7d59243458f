// host_expand.cpp -- host side of the CSR copy-out: column indices from visibility words.
//
// The column indices of a CSR row of the form-factor matrix are exactly the
// positions of the set bits of that row's visibility words (original column
// order, form_factors.py:52, 69).  fluxb200_ff_assemble therefore ships the
// words (n/8 bytes per row) over PCIe instead of the int32 positions (4 bytes
// per stored entry) and a few host threads write the positions while the next
// sub-slab is traced.  This is a transport encoding, not a compute fallback:
// the words come from the GPU kernels.
#include "host_expand.h"
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace fluxb200 {

namespace {

template <class Index> int64_t expand_scalar(const uint32_t *words, int nwords, Index *out) {
    Index *dst = out;
    for (int k = 0; k < nwords; ++k) {
        uint32_t w = words[k];
        const Index base = (Index)k * 32;
        while (w) {
            *dst++ = base + (Index)__builtin_ctz(w);
            w &= w - 1;
        }
    }
    return (int64_t)(dst - out);
}

#if defined(__x86_64__)
// 16 column positions per step: `vpcompressd` packs the lane ids selected by 16 mask bits; the packed
// ids are appended to a register accumulator (`vpermt2d` with a computed merge index) and every full
// group of 16 leaves as ONE aligned non-temporal 64-byte store.  Plain stores would read every output
// line into the cache first (write-allocate), and with several threads expanding at once the host's
// memory bandwidth, not the cores, sets the pace: measured in the build container on 8 threads,
// 3.3e9 indices/s with ordinary stores against 10e9 with streaming stores.  The row's first entries up
// to the 64-byte boundary and its last partial group use masked stores, so nothing is ever written
// outside the row (another thread owns the next one).
// the first `cnt` (<= 16) packed ids with masked stores (head and tail of a row)
template <bool kWide>
__attribute__((target("avx512f"), always_inline)) inline void store_masked(char *&dst, __m512i v, int cnt) {
    if (!kWide) {
        _mm512_mask_storeu_epi32(dst, (__mmask16)((1u << cnt) - 1u), v);
    } else {
        const int c0 = cnt < 8 ? cnt : 8, c1 = cnt - c0;
        _mm512_mask_storeu_epi64(dst, (__mmask8)((1u << c0) - 1u), _mm512_cvtepu32_epi64(_mm512_castsi512_si256(v)));
        if (c1 > 0)
            _mm512_mask_storeu_epi64(dst + 64, (__mmask8)((1u << c1) - 1u),
                                     _mm512_cvtepu32_epi64(_mm512_extracti64x4_epi64(v, 1)));
    }
    dst += (size_t)(kWide ? 8 : 4) * cnt;
}
// a full group of 16 at a 64-byte aligned dst, streaming
template <bool kWide>
__attribute__((target("avx512f"), always_inline)) inline void store_full(char *&dst, __m512i v) {
    if (!kWide) {
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dst), v);
    } else {
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dst), _mm512_cvtepu32_epi64(_mm512_castsi512_si256(v)));
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + 64), _mm512_cvtepu32_epi64(_mm512_extracti64x4_epi64(v, 1)));
    }
    dst += (size_t)(kWide ? 8 : 4) * 16;
}

// kWide = false: int32 positions, one 64-byte line per 16; kWide = true: int64 positions, the 16 packed
// ids are zero-extended into two lines of 8.
template <bool kWide>
__attribute__((target("avx512f,popcnt"))) int64_t expand_avx512(const uint32_t *words, int nwords, void *out_v,
                                                                int64_t row_entries) {
    constexpr int kBytes = kWide ? 8 : 4;
    char *dst = reinterpret_cast<char *>(out_v);
    char *const out = dst;
    int head = (int)(((64 - ((uintptr_t)dst & 63)) & 63) / kBytes); // entries before the first 64-byte boundary
    if (head > row_entries) head = (int)row_entries;
    const __m512i lanes = _mm512_set_epi32(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m512i sixteen = _mm512_set1_epi32(16);
    __m512i acc = _mm512_setzero_si512(); // `fill` pending entries in lanes 0 .. fill-1
    int fill = 0;
    for (int k = 0; k < nwords; ++k) {
        const uint32_t w = words[k];
        if (!w) continue;
        const __m512i base = _mm512_add_epi32(lanes, _mm512_set1_epi32(k * 32));
        for (int h = 0; h < 2; ++h) {
            const __mmask16 m = (__mmask16)(h ? w >> 16 : w & 0xffffu);
            if (!m) continue;
            const int c = __builtin_popcount(m);
            const __m512i v = _mm512_maskz_compress_epi32(m, h ? _mm512_add_epi32(base, sixteen) : base);
            // {acc, v} -> lanes 0..15 of the concatenation acc[0..fill) ++ v: lane i takes acc[i] below fill,
            // v[i - fill] (index 16 + i - fill of the pair) from there on
            const __m512i idx = _mm512_mask_add_epi32(lanes, (__mmask16)(0xffffu << fill), lanes,
                                                      _mm512_set1_epi32(16 - fill));
            const __m512i merged = _mm512_permutex2var_epi32(acc, idx, v);
            const int tot = fill + c;
            if (head > 0) { // once per row: the entries in front of the boundary
                if (tot < head) {
                    acc = merged;
                    fill = tot;
                    continue;
                }
                store_masked<kWide>(dst, merged, head);
                alignas(64) int32_t t[32]; // concatenation lanes 0..31; the part after `head` moves to lane 0
                _mm512_store_si512(t, merged);
                _mm512_store_si512(t + 16, _mm512_permutexvar_epi32(_mm512_add_epi32(lanes, _mm512_set1_epi32(16 - fill)), v));
                fill = tot - head; // < 16: fill was below head
                acc = _mm512_loadu_si512(t + head);
                head = 0;
                continue;
            }
            if (tot >= 16) { // dst is 64-byte aligned here and the 16 entries all belong to this row
                store_full<kWide>(dst, merged);
                acc = _mm512_permutexvar_epi32(_mm512_add_epi32(lanes, _mm512_set1_epi32(16 - fill)), v);
                fill = tot - 16;
            } else {
                acc = merged;
                fill = tot;
            }
        }
    }
    if (fill > 0) store_masked<kWide>(dst, acc, fill);
    _mm_sfence(); // streaming stores are weakly ordered: make them visible before the row is reported done
    return (int64_t)(dst - out) / kBytes;
}

__attribute__((target("popcnt"))) int64_t count_bits_popcnt(const uint32_t *w, int n) {
    int64_t c = 0;
    int k = 0;
    for (; k + 2 <= n; k += 2) {
        uint64_t two;
        memcpy(&two, w + k, 8);
        c += __builtin_popcountll(two);
    }
    for (; k < n; ++k) c += __builtin_popcount(w[k]);
    return c;
}
#endif

bool have_avx512() {
#if defined(__x86_64__)
    // FLUXB200_NO_AVX512 forces the portable loop (tests cover both)
    static const bool v = __builtin_cpu_supports("avx512f") && !getenv("FLUXB200_NO_AVX512");
    return v;
#else
    return false;
#endif
}

} // namespace

int64_t count_bits(const uint32_t *words, int nwords) {
#if defined(__x86_64__)
    static const bool hw = __builtin_cpu_supports("popcnt");
    if (hw) return count_bits_popcnt(words, nwords);
#endif
    int64_t c = 0;
    for (int k = 0; k < nwords; ++k) c += __builtin_popcount(words[k]);
    return c;
}

int64_t expand_words(const uint32_t *words, int nwords, int index_width, void *out, int64_t row_entries) {
#if defined(__x86_64__)
    if (have_avx512()) {
        if (row_entries < 0) row_entries = count_bits(words, nwords); // not known: count first
        return index_width == 8 ? expand_avx512<true>(words, nwords, out, row_entries)
                                : expand_avx512<false>(words, nwords, out, row_entries);
    }
#endif
    if (index_width == 8) return expand_scalar<int64_t>(words, nwords, reinterpret_cast<int64_t *>(out));
    return expand_scalar<int32_t>(words, nwords, reinterpret_cast<int32_t *>(out));
}

// ---- worker pool -------------------------------------------------------------------
HostExpander::~HostExpander() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_work_.notify_all();
    for (auto &t : workers_) t.join();
}

void HostExpander::start(int nthreads) {
    std::lock_guard<std::mutex> lk(mu_);
    while ((int)workers_.size() < nthreads) workers_.emplace_back([this] { run(); });
}

void HostExpander::submit(ExpandTask *t) {
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (int p = 0; p < t->pieces; ++p) queue_.emplace_back(t, p);
    }
    cv_work_.notify_all();
}

void HostExpander::wait(ExpandTask *t) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [t] { return t->pending.load() == 0; });
}

void HostExpander::run() {
    for (;;) {
        std::pair<ExpandTask *, int> item;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_work_.wait(lk, [this] { return stop_ || !queue_.empty(); });
            if (queue_.empty()) return; // stop_
            item = queue_.front();
            queue_.pop_front();
        }
        ExpandTask *t = item.first;
        const size_t r0 = t->mr * (size_t)item.second / (size_t)t->pieces;
        const size_t r1 = t->mr * (size_t)(item.second + 1) / (size_t)t->pieces;
        for (size_t r = r0; r < r1; ++r) {
            const uint32_t *w = t->words + r * (size_t)t->nwords;
            const int64_t bits = count_bits(w, t->nwords);
            if (bits != t->offs[r + 1] - t->offs[r]) { // never write outside the row's range
                t->mismatch.store(1);
                continue;
            }
            char *dst = reinterpret_cast<char *>(t->indices) + (size_t)t->index_width * (size_t)t->offs[r];
            expand_words(w, t->nwords, t->index_width, dst, bits);
        }
        if (t->copy_bytes) {
            const size_t b0 = t->copy_bytes * (size_t)item.second / (size_t)t->pieces & ~(size_t)63;
            const size_t b1 = item.second + 1 == t->pieces ? t->copy_bytes
                                                           : (t->copy_bytes * (size_t)(item.second + 1) / (size_t)t->pieces & ~(size_t)63);
            if (b1 > b0) memcpy(t->copy_dst + b0, t->copy_src + b0, b1 - b0);
        }
        if (t->pending.fetch_sub(1) == 1) {
            std::lock_guard<std::mutex> lk(mu_); // pairs with the predicate check in wait()
            cv_done_.notify_all();
        }
    }
}

} // namespace fluxb200
