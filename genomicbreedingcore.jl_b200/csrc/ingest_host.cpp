// ingest_host.cpp — host threads that read Julia's Matrix{Union{Float64,Missing}} (src/all_structs.jl:66) in place
// from pageable memory and fill the pinned ring of the ingest (ingest_host.h).  Compiled by g++ (not nvcc): the SIMD
// variants use function-level target attributes and are picked at run time, so the library still loads on a host
// without AVX-512.
#include "ingest_host.h"

#include <immintrin.h>
#include <math.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

// ---------------------------------------------------------------------------------------------------------------
// worker pool: persistent threads, one job at a time, fn(arg, tid, nthreads) on every worker
// ---------------------------------------------------------------------------------------------------------------
struct GrmHostPool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    void (*fn)(void *, int, int) = nullptr;
    void *arg = nullptr;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
};

static void pool_worker(GrmHostPool *pool, int tid, int nthreads) {
    uint64_t seen = 0;
    for (;;) {
        void (*fn)(void *, int, int);
        void *arg;
        {
            std::unique_lock<std::mutex> lk(pool->mu);
            pool->cv_job.wait(lk, [&] { return pool->stop || pool->generation != seen; });
            if (pool->stop) return;
            seen = pool->generation;
            fn = pool->fn;
            arg = pool->arg;
        }
        fn(arg, tid, nthreads);
        {
            std::lock_guard<std::mutex> lk(pool->mu);
            if (--pool->pending == 0) pool->cv_done.notify_all();
        }
    }
}

GrmHostPool *grm_pool_create(int threads) {
    if (threads < 1) threads = 1;
    GrmHostPool *pool = new (std::nothrow) GrmHostPool();
    if (!pool) return nullptr;
    try {
        for (int t = 0; t < threads; ++t) pool->workers.emplace_back(pool_worker, pool, t, threads);
    } catch (...) {
        grm_pool_destroy(pool);
        return nullptr;
    }
    return pool;
}

void grm_pool_destroy(GrmHostPool *pool) {
    if (!pool) return;
    {
        std::lock_guard<std::mutex> lk(pool->mu);
        pool->stop = true;
    }
    pool->cv_job.notify_all();
    for (auto &w : pool->workers)
        if (w.joinable()) w.join();
    delete pool;
}

int grm_pool_threads(const GrmHostPool *pool) { return pool ? static_cast<int>(pool->workers.size()) : 0; }

void grm_pool_start(GrmHostPool *pool, void (*fn)(void *, int, int), void *arg) {
    std::lock_guard<std::mutex> lk(pool->mu);
    pool->fn = fn;
    pool->arg = arg;
    pool->pending = static_cast<int>(pool->workers.size());
    ++pool->generation;
    pool->cv_job.notify_all();
}

void grm_pool_wait(GrmHostPool *pool) {
    std::unique_lock<std::mutex> lk(pool->mu);
    pool->cv_done.wait(lk, [&] { return pool->pending == 0; });
}

void grm_pool_run(GrmHostPool *pool, void (*fn)(void *, int, int), void *arg) {
    grm_pool_start(pool, fn, arg);
    grm_pool_wait(pool);
}

static int simd_level();

// ---------------------------------------------------------------------------------------------------------------
// one cell: the same test as the device pack kernel (pack.cu): c = rn(m x) is a code iff double(c) == m x, 0 <= c <= m
// ---------------------------------------------------------------------------------------------------------------
static inline int8_t quantise_cell(double x, bool is_float, bool have_sel, double mf, uint32_t &flags) {
    if (!is_float) {
        flags |= GRM_HF_MISSING;
        return GRM_CODE_MISSING;
    }
    const double t = x * mf;
    if (t >= 0.0 && t <= mf) {
        const int c = static_cast<int>(t + 0.5);  // t in [0, 64]: round-half-up equals any rounding when t is integral
        if (static_cast<double>(c) == t) return static_cast<int8_t>(c);
        flags |= GRM_HF_NOT_DOSAGE;
        return GRM_CODE_BAD;
    }
    if (x != x) {
        // without selector bytes NaN is the "missing" marker of the plain double* interface
        flags |= have_sel ? GRM_HF_NAN_PAYLOAD : GRM_HF_MISSING;
        return have_sel ? GRM_CODE_NAN : GRM_CODE_MISSING;
    }
    if (isinf(x)) {
        flags |= GRM_HF_INF;
        return GRM_CODE_INF;
    }
    flags |= GRM_HF_NOT_DOSAGE;
    return GRM_CODE_BAD;
}

// Software prefetch distance (bytes ahead of the streaming read) of the quantiser.  A single core's demand misses do not
// fill its share of the memory system; prefetching a few KB ahead does (measured on the build host: 7.8 -> 11.7 GB/s per
// thread, 56 -> 73 GB/s on 8 threads; on the B200 host, 16 threads: 97 GB/s without, 150 at 4 KB, 157 at 8 KB —
// profiles/host_ingest_sweep_r2.jsonl).  Run-time adjustable (grm_cuda_host_tune) so that it can be swept on the target.
constexpr int PF_DIST_DEFAULT = 16384;
static int g_pf_dist = PF_DIST_DEFAULT;
void grm_host_set_prefetch(int bytes) { g_pf_dist = bytes < 0 ? PF_DIST_DEFAULT : (bytes > (1 << 20) ? (1 << 20) : bytes); }
int grm_host_get_prefetch(void) { return g_pf_dist; }

static uint32_t quantise_col_scalar(const double *x, const uint8_t *s, int tag, int64_t n, double mf, int8_t *out) {
    uint32_t flags = 0;
    for (int64_t i = 0; i < n; ++i) out[i] = quantise_cell(x[i], !s || s[i] == tag, s != nullptr, mf, flags);
    return flags;
}

// Cache level the software prefetch asks for (0 = T0: all levels, 1 = T1, 2 = T2, 3 = NTA) and whether the codes are
// written with streaming stores: the ring slot is read next by the copy engine, never by this core, and a streaming store
// spares the read-for-ownership of the output line (1/9 of the pass's read traffic).  Same-box sweep on the B200 host's 16
// threads (profiles/host_ingest_hints_r2.jsonl): T0 / 8 KB / plain stores 144 GB/s of source, streaming stores 166, T2 +
// streaming stores at 8 - 32 KB 169 - 170 (NTA: 95).  Defaults: T2, 16 KB, streaming stores.
constexpr int PF_HINT_DEFAULT = 2, NT_STORE_DEFAULT = 1;
static int g_pf_hint = PF_HINT_DEFAULT;
static int g_nt_store = NT_STORE_DEFAULT;
void grm_host_set_prefetch_hint(int hint) { g_pf_hint = hint < 0 ? PF_HINT_DEFAULT : (hint > 3 ? 3 : hint); }
int grm_host_get_prefetch_hint(void) { return g_pf_hint; }
void grm_host_set_stream_stores(int on) { g_nt_store = on < 0 ? NT_STORE_DEFAULT : (on ? 1 : 0); }
int grm_host_get_stream_stores(void) { return g_nt_store; }

template <int HINT, bool NT>
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) static uint32_t
quantise_col_avx512_impl(const double *x, const uint8_t *s, int tag, int64_t n, double mf, int m, int8_t *out, int pf) {
    uint32_t flags = 0;
    const __m512d vm = _mm512_set1_pd(mf);
    const __m256i vmax = _mm256_set1_epi32(m);
    const __m128i vtag = _mm_set1_epi8(static_cast<char>(tag));
    int64_t i = 0;
    constexpr _mm_hint hint = HINT == 0 ? _MM_HINT_T0 : HINT == 1 ? _MM_HINT_T1 : HINT == 2 ? _MM_HINT_T2 : _MM_HINT_NTA;
    for (; i + 16 <= n; i += 16) {
        if (pf) {
            _mm_prefetch(reinterpret_cast<const char *>(x + i) + pf, hint);
            _mm_prefetch(reinterpret_cast<const char *>(x + i) + pf + 64, hint);
        }
        const __m512d t0 = _mm512_mul_pd(_mm512_loadu_pd(x + i), vm);
        const __m512d t1 = _mm512_mul_pd(_mm512_loadu_pd(x + i + 8), vm);
        const __m256i c0 = _mm512_cvtpd_epi32(t0);  // round to nearest even; NaN / out of range -> 0x80000000
        const __m256i c1 = _mm512_cvtpd_epi32(t1);
        __mmask8 ok0 = _mm512_cmp_pd_mask(_mm512_cvtepi32_pd(c0), t0, _CMP_EQ_OQ) & _mm256_cmple_epu32_mask(c0, vmax);
        __mmask8 ok1 = _mm512_cmp_pd_mask(_mm512_cvtepi32_pd(c1), t1, _CMP_EQ_OQ) & _mm256_cmple_epu32_mask(c1, vmax);
        __mmask16 ok = static_cast<__mmask16>(ok0) | (static_cast<__mmask16>(ok1) << 8);
        if (s) ok &= _mm_cmpeq_epi8_mask(_mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i)), vtag);
        if (ok == 0xFFFF) {
            const __m512i c = _mm512_inserti64x4(_mm512_castsi256_si512(c0), c1, 1);
            if (NT) _mm_stream_si128(reinterpret_cast<__m128i *>(out + i), _mm512_cvtepi32_epi8(c));  // out + i is 16-byte aligned
            else _mm_storeu_si128(reinterpret_cast<__m128i *>(out + i), _mm512_cvtepi32_epi8(c));
        } else {
            for (int j = 0; j < 16; ++j)
                out[i + j] = quantise_cell(x[i + j], !s || s[i + j] == tag, s != nullptr, mf, flags);
        }
    }
    for (; i < n; ++i) out[i] = quantise_cell(x[i], !s || s[i] == tag, s != nullptr, mf, flags);
    if (NT) _mm_sfence();  // the column is complete before its unit is reported done
    return flags;
}

static uint32_t quantise_col_avx512(const double *x, const uint8_t *s, int tag, int64_t n, double mf, int m, int8_t *out) {
    const int pf = g_pf_dist;
    const bool nt = g_nt_store && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    switch (g_pf_hint * 2 + (nt ? 1 : 0)) {
        case 1: return quantise_col_avx512_impl<0, true>(x, s, tag, n, mf, m, out, pf);
        case 2: return quantise_col_avx512_impl<1, false>(x, s, tag, n, mf, m, out, pf);
        case 3: return quantise_col_avx512_impl<1, true>(x, s, tag, n, mf, m, out, pf);
        case 4: return quantise_col_avx512_impl<2, false>(x, s, tag, n, mf, m, out, pf);
        case 5: return quantise_col_avx512_impl<2, true>(x, s, tag, n, mf, m, out, pf);
        case 6: return quantise_col_avx512_impl<3, false>(x, s, tag, n, mf, m, out, pf);
        case 7: return quantise_col_avx512_impl<3, true>(x, s, tag, n, mf, m, out, pf);
        default: return quantise_col_avx512_impl<0, false>(x, s, tag, n, mf, m, out, pf);
    }
}

__attribute__((target("avx2"))) static uint32_t quantise_col_avx2(const double *x, const uint8_t *s, int tag, int64_t n,
                                                                   double mf, int m, int8_t *out) {
    uint32_t flags = 0;
    const __m256d vm = _mm256_set1_pd(mf);
    const __m128i vmax = _mm_set1_epi32(m);
    int64_t i = 0;
    const int pf = g_pf_dist;
    for (; i + 8 <= n; i += 8) {
        if (pf) _mm_prefetch(reinterpret_cast<const char *>(x + i) + pf, _MM_HINT_T0);
        const __m256d t0 = _mm256_mul_pd(_mm256_loadu_pd(x + i), vm);
        const __m256d t1 = _mm256_mul_pd(_mm256_loadu_pd(x + i + 4), vm);
        const __m128i c0 = _mm256_cvtpd_epi32(t0);
        const __m128i c1 = _mm256_cvtpd_epi32(t1);
        const int e0 = _mm256_movemask_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(c0), t0, _CMP_EQ_OQ));
        const int e1 = _mm256_movemask_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(c1), t1, _CMP_EQ_OQ));
        // 0 <= c <= m as one unsigned comparison: min_epu32(c, m) == c
        const int r0 = _mm_movemask_ps(_mm_castsi128_ps(_mm_cmpeq_epi32(_mm_min_epu32(c0, vmax), c0)));
        const int r1 = _mm_movemask_ps(_mm_castsi128_ps(_mm_cmpeq_epi32(_mm_min_epu32(c1, vmax), c1)));
        bool ok = (e0 & r0) == 0xF && (e1 & r1) == 0xF;
        if (ok && s) {
            uint64_t sv;
            memcpy(&sv, s + i, 8);
            ok = sv == 0x0101010101010101ull * static_cast<uint8_t>(tag);
        }
        if (ok) {
            const __m128i w = _mm_packs_epi32(c0, c1);      // 8 x int16
            const __m128i b = _mm_packs_epi16(w, w);        // 8 x int8 in the low half
            _mm_storel_epi64(reinterpret_cast<__m128i *>(out + i), b);
        } else {
            for (int j = 0; j < 8; ++j)
                out[i + j] = quantise_cell(x[i + j], !s || s[i + j] == tag, s != nullptr, mf, flags);
        }
    }
    for (; i < n; ++i) out[i] = quantise_cell(x[i], !s || s[i] == tag, s != nullptr, mf, flags);
    return flags;
}

static int simd_level() {
    static const int level = [] {
        __builtin_cpu_init();
        if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
            __builtin_cpu_supports("avx512dq"))
            return 2;
        if (__builtin_cpu_supports("avx2")) return 1;
        return 0;
    }();
    return level;
}

const char *grm_host_simd(void) {
    const int l = simd_level();
    return l == 2 ? "avx512" : l == 1 ? "avx2" : "scalar";
}

uint32_t grm_host_quantise(const double *payload, const uint8_t *sel, int float_tag, int64_t n, int64_t lda, int64_t k0,
                           int64_t k1, int m, int8_t *out, int64_t ldo) {
    uint32_t flags = 0;
    const double mf = static_cast<double>(m);
    const int level = simd_level();
    for (int64_t k = k0; k < k1; ++k) {
        const double *x = payload + k * lda;
        const uint8_t *s = sel ? sel + k * lda : nullptr;
        int8_t *o = out + (k - k0) * ldo;
        if (level == 2) flags |= quantise_col_avx512(x, s, float_tag, n, mf, m, o);
        else if (level == 1) flags |= quantise_col_avx2(x, s, float_tag, n, mf, m, o);
        else flags |= quantise_col_scalar(x, s, float_tag, n, mf, o);
    }
    return flags;
}

// One column of the FP64 ring, one pass: copy, missing -> NaN, flags.  The reference-facing semantics are those of the
// scalar version below: a cell whose selector is not the Float64 tag is `missing` (its payload is never used); a
// Float64 cell holding NaN is a NaN payload (with selectors) or the missing marker (plain double* input).
template <bool NT>
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) static uint32_t
copy_col_avx512_impl(const double *x, const uint8_t *s, int tag, int64_t n, double *out, int pf) {
    const __m512d vnan = _mm512_set1_pd(__builtin_nan(""));
    const __m128i vtag = _mm_set1_epi8(static_cast<char>(tag));
    unsigned any_miss = 0, any_nan = 0;
    int64_t i = 0;
    for (; i + 8 <= n; i += 8) {
        if (pf) _mm_prefetch(reinterpret_cast<const char *>(x + i) + pf, _MM_HINT_T2);
        const __m512d v = _mm512_loadu_pd(x + i);
        __mmask8 is_float = 0xFF;
        if (s) is_float = static_cast<__mmask8>(_mm_cmpeq_epi8_mask(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(s + i)), vtag));
        any_nan |= static_cast<unsigned>(_mm512_cmp_pd_mask(v, v, _CMP_UNORD_Q) & is_float);
        any_miss |= static_cast<unsigned>(static_cast<__mmask8>(~is_float));
        const __m512d w = _mm512_mask_mov_pd(vnan, is_float, v);
        if (NT) _mm512_stream_pd(out + i, w);  // out + i is 64-byte aligned
        else _mm512_storeu_pd(out + i, w);
    }
    for (; i < n; ++i) {
        const bool f = !s || s[i] == tag;
        const double v = x[i];
        if (!f) {
            out[i] = __builtin_nan("");
            any_miss = 1;
        } else {
            out[i] = v;
            if (v != v) any_nan = 1;
        }
    }
    if (NT) _mm_sfence();  // the column is complete before its unit is reported done
    uint32_t flags = 0;
    if (s) flags = (any_miss ? GRM_HF_MISSING : 0u) | (any_nan ? GRM_HF_NAN_PAYLOAD : 0u);
    else if (any_nan) flags = GRM_HF_MISSING;
    return flags;
}

uint32_t grm_host_copy(const double *payload, const uint8_t *sel, int float_tag, int64_t n, int64_t lda, int64_t k0,
                       int64_t k1, double *out, int64_t ldo) {
    uint32_t flags = 0;
    if (simd_level() == 2) {
        // one pass per column: the source is read once, the ring slot written once — with streaming stores when the
        // row of the slot is 64-byte aligned (the copy engine reads it next, this core never does)
        const int pf = g_pf_dist;
        for (int64_t k = k0; k < k1; ++k) {
            const double *x = payload + k * lda;
            const uint8_t *s = sel ? sel + k * lda : nullptr;
            double *o = out + (k - k0) * ldo;
            if (g_nt_store && (reinterpret_cast<uintptr_t>(o) & 63) == 0) flags |= copy_col_avx512_impl<true>(x, s, float_tag, n, o, pf);
            else flags |= copy_col_avx512_impl<false>(x, s, float_tag, n, o, pf);
        }
        return flags;
    }
    const double nan = __builtin_nan("");
    for (int64_t k = k0; k < k1; ++k) {
        const double *x = payload + k * lda;
        double *o = out + (k - k0) * ldo;
        memcpy(o, x, static_cast<size_t>(n) * sizeof(double));
        // the column just written is cache-resident: NaN payloads / missing selectors are found on the copy
        unsigned nan_seen = 0, miss_seen = 0;
        for (int64_t i = 0; i < n; ++i) nan_seen |= static_cast<unsigned>(o[i] != o[i]);
        if (sel) {
            const uint8_t *s = sel + k * lda;
            const uint8_t tag = static_cast<uint8_t>(float_tag);
            for (int64_t i = 0; i < n; ++i) miss_seen |= static_cast<unsigned>(s[i] ^ tag);
            if (miss_seen || nan_seen) {
                for (int64_t i = 0; i < n; ++i) {
                    if (s[i] != tag) {
                        o[i] = nan;
                        flags |= GRM_HF_MISSING;
                    } else if (o[i] != o[i]) {
                        flags |= GRM_HF_NAN_PAYLOAD;
                    }
                }
            }
        } else if (nan_seen) {
            flags |= GRM_HF_MISSING;
        }
    }
    return flags;
}

uint32_t grm_host_copy_codes(const int8_t *codes, int64_t n, int64_t ldc, int64_t k0, int64_t k1, int m, int8_t *out,
                             int64_t ldo) {
    uint32_t flags = 0;
    for (int64_t k = k0; k < k1; ++k) {
        const int8_t *x = codes + k * ldc;
        int8_t *o = out + (k - k0) * ldo;
        memcpy(o, x, static_cast<size_t>(n));
        // validate on the cache-resident copy: anything outside 0..m that is not the missing marker is not a code
        unsigned bad = 0;
        for (int64_t i = 0; i < n; ++i) bad |= static_cast<unsigned>(static_cast<uint8_t>(o[i]) > static_cast<unsigned>(m));
        if (bad) {
            for (int64_t i = 0; i < n; ++i) {
                if (o[i] == GRM_CODE_MISSING) {
                    flags |= GRM_HF_MISSING;
                } else if (static_cast<uint8_t>(o[i]) > static_cast<unsigned>(m)) {
                    o[i] = GRM_CODE_BAD;
                    flags |= GRM_HF_NOT_DOSAGE;
                }
            }
        }
    }
    return flags;
}

void grm_host_detect(const double *payload, const uint8_t *sel, int float_tag, int64_t n, int64_t p, int64_t lda,
                     int64_t cells, int ignore_missing, uint32_t *bits_out, uint32_t *flags_out) {
    uint32_t bits = 0, flags = 0;
    int64_t left = cells;
    for (int64_t k = 0; k < p && left > 0; ++k) {
        const double *x = payload + k * lda;
        const uint8_t *s = sel ? sel + k * lda : nullptr;
        const int64_t len = left < n ? left : n;
        for (int64_t i = 0; i < len; ++i) {
            uint32_t f = 0;
            const int8_t c = quantise_cell(x[i], !s || s[i] == float_tag, s != nullptr, 64.0, f);
            if (f == 0) bits |= static_cast<uint32_t>(c);
            else if (!(ignore_missing && f == GRM_HF_MISSING)) flags |= f;
        }
        left -= len;
    }
    *bits_out = bits;
    *flags_out = flags;
}
