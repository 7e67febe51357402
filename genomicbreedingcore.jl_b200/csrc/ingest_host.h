// ingest_host.h — host-side half of the ingest of genomes.allele_frequencies (plain C++, no CUDA).
//
// The reference stores the matrix as Matrix{Union{Float64,Missing}} (src/all_structs.jl:66): n*p Float64 payloads
// (column-major, entry fastest) followed by n*p selector ("type tag") bytes, in ordinary pageable memory.  Worker
// threads read that layout in place and write into a pinned ring, either int8 dosage codes (1 byte per cell crosses
// PCIe instead of 8) or the dense FP64 cells (missing -> NaN), so no conversion pass and no pinned copy of the whole
// matrix is ever needed on the caller's side.
#pragma once
#include <stdint.h>

// flags returned by the passes below (bits 0..2 match the device pack kernel's flags)
constexpr uint32_t GRM_HF_MISSING = 1u;      // a cell whose selector says `missing` (or, without selectors, a NaN)
constexpr uint32_t GRM_HF_INF = 2u;          // +-Inf payload
constexpr uint32_t GRM_HF_NOT_DOSAGE = 4u;   // m*x is not an integer in [0, m]
constexpr uint32_t GRM_HF_NAN_PAYLOAD = 16u; // a Float64 cell (selector says so) holding NaN: the reference propagates
                                             // it into G, where inflatediagonals! reports "not symmetric" (calc.jl:43)

// int8 sentinels written for cells that are not dosage codes (valid codes are 0..64)
constexpr int8_t GRM_CODE_MISSING = -128, GRM_CODE_NAN = -127, GRM_CODE_INF = -126, GRM_CODE_BAD = -125;

struct GrmHostPool;
GrmHostPool *grm_pool_create(int threads);
void grm_pool_destroy(GrmHostPool *pool);
int grm_pool_threads(const GrmHostPool *pool);
// fn(arg, tid, nthreads) on every worker; returns when all have finished
void grm_pool_run(GrmHostPool *pool, void (*fn)(void *, int, int), void *arg);
// the same in two halves: start returns at once, wait blocks until every worker has returned from fn (one job at a time)
void grm_pool_start(GrmHostPool *pool, void (*fn)(void *, int, int), void *arg);
void grm_pool_wait(GrmHostPool *pool);

// Columns [k0, k1) of the n x p matrix -> out[(k - k0) * ldo + i] = int8(m * x) or a sentinel.  sel may be null (plain
// Matrix{Float64}: every cell is a Float64, NaN counts as missing).  Returns the OR of the flags raised.
uint32_t grm_host_quantise(const double *payload, const uint8_t *sel, int float_tag, int64_t n, int64_t lda, int64_t k0,
                           int64_t k1, int m, int8_t *out, int64_t ldo);
// Columns [k0, k1) of pre-packed codes (codes[i + k * ldc] in 0..m, GRM_CODE_MISSING = missing) -> out[(k - k0) * ldo + i];
// any other value becomes GRM_CODE_BAD and raises GRM_HF_NOT_DOSAGE.
uint32_t grm_host_copy_codes(const int8_t *codes, int64_t n, int64_t ldc, int64_t k0, int64_t k1, int m, int8_t *out,
                             int64_t ldo);
// The same columns as dense FP64, out[(k - k0) * ldo + i]; a missing cell becomes NaN.
uint32_t grm_host_copy(const double *payload, const uint8_t *sel, int float_tag, int64_t n, int64_t lda, int64_t k0,
                       int64_t k1, double *out, int64_t ldo);
// OR of round(64 x) over the first `cells` cells (column-major order) + flags; missing cells are skipped when
// ignore_missing != 0.  A cell that is not k/64 in [0, 1] raises GRM_HF_NOT_DOSAGE.
void grm_host_detect(const double *payload, const uint8_t *sel, int float_tag, int64_t n, int64_t p, int64_t lda,
                     int64_t cells, int ignore_missing, uint32_t *bits, uint32_t *flags);
// software prefetch distance of the quantiser, bytes (0 = off)
void grm_host_set_prefetch(int bytes);
int grm_host_get_prefetch(void);
// cache level of that prefetch (0 T0, 1 T1, 2 T2, 3 NTA) and streaming stores for the codes (0 / 1)
void grm_host_set_prefetch_hint(int hint);
int grm_host_get_prefetch_hint(void);
void grm_host_set_stream_stores(int on);
int grm_host_get_stream_stores(void);
// name of the SIMD level the passes above dispatch to on this machine ("avx512", "avx2", "scalar")
const char *grm_host_simd(void);
