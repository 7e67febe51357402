/*
 * grm_cuda.h — C ABI of libgrm_cuda, the B200 (sm_100a) backend for the genomic
 * relationship matrix (GRM) path of GenomicBreedingCore.jl.
 *
 * The reference has no FFI: its boundary is three exported Julia functions in
 * src/grm/calc.jl.  Each entry point below names the reference interface it
 * replaces (file:line, relative to the reference repository root):
 *
 *   grm_cuda_simple             <- grmsimple(genomes; max_iter, verbose)             src/grm/calc.jl:115-142
 *   grm_cuda_ploidyaware        <- grmploidyaware(genomes; ploidy, max_iter, verbose) src/grm/calc.jl:189-217
 *   grm_cuda_inflate_diagonals  <- inflatediagonals!(X; max_iter, verbose)           src/grm/calc.jl:41-70
 *
 * Conventions (mirroring the Julia side):
 *   - All matrices are COLUMN-MAJOR, exactly as Julia stores them.  The input is
 *     genomes.allele_frequencies, n entries x p loci, element (i,k) at A[i + k*lda];
 *     a missing value is passed as NaN (plain double* entry points) or through the
 *     selector bytes of Julia's Matrix{Union{Float64,Missing}} (the *_union entry
 *     points).  The output is the n x n Float64 GRM,
 *     element (i,j) at G[i + j*ldg]; both triangles are written and are
 *     bit-identical mirrors of each other (the reference asserts issymmetric,
 *     calc.jl:43).  Row/column order is the input order of genomes.entries
 *     (calc.jl:123,196); nothing is sorted.
 *   - Pointers are borrowed for the duration of the call only.
 *   - No C++ exception crosses this boundary.  Every function returns a
 *     grm_cuda_status; grm_cuda_last_error() returns a thread-local message.
 *   - There is no CPU fallback: without a usable CUDA device every compute entry
 *     point returns GRM_CUDA_ERR_CUDA.
 *   - Callable from any host thread; the library sets the device per call and
 *     installs no signal handlers.
 */
#ifndef GRM_CUDA_H
#define GRM_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRM_CUDA_VERSION 200 /* 0.2.0 */

typedef enum grm_cuda_status {
    GRM_CUDA_OK = 0,
    GRM_CUDA_ERR_BAD_ARGUMENT = 1,   /* null pointer, n<1, p<1, lda<n, ...  -> ArgumentError                       */
    GRM_CUDA_ERR_MISSING = 2,        /* NaN ("missing") in allele_frequencies -> the reference throws (calc.jl:127)  */
    GRM_CUDA_ERR_NONFINITE = 3,      /* +-Inf in the input                                                           */
    GRM_CUDA_ERR_NOT_SYMMETRIC = 4,  /* inflatediagonals!: "The input matrix is not symmetric."      calc.jl:43-45   */
    GRM_CUDA_ERR_NAN_MATRIX = 5,     /* inflatediagonals!: "The input matrix contains NaN values."   calc.jl:52-54   */
    GRM_CUDA_ERR_INF_MATRIX = 6,     /* inflatediagonals!: "The input matrix contains Inf values."   calc.jl:55-57   */
    GRM_CUDA_ERR_CUDA = 7,           /* CUDA runtime/driver/cuSOLVER failure, or no sm_100 device                    */
    GRM_CUDA_ERR_OOM = 8,            /* device or pinned-host allocation failed                                      */
    GRM_CUDA_ERR_NOT_DOSAGE = 9,     /* int8 path forced but the frequencies are not k/m with m a power of two <= 64 */
    GRM_CUDA_ERR_OVERFLOW = 10,      /* m*m*p would overflow the int32 accumulators                                  */
    GRM_CUDA_ERR_ALL_FILTERED = 11,  /* QC mask: "All loci filtered out ..." (src/genomes/filter.jl:363-365, :644-646) */
    GRM_CUDA_ERR_NOT_POSDEF = 12,    /* grm_cuda_cholesky: Julia's PosDefException (what MvNormal(mu, Sigma) throws,
                                        src/simulations/simulate_effects.jl:480) */
    GRM_CUDA_ERR_COMM = 13           /* NCCL not loadable / communicator missing / collective failed                  */
} grm_cuda_status;

/* Which contraction ran (grm_cuda_info.path) / may run (grm_cuda_options.path). */
#define GRM_CUDA_PATH_AUTO 0 /* detect dosage-quantised input on load; fall back to FP64 for continuous input */
#define GRM_CUDA_PATH_I8   1 /* exact int8 tcgen05 SYRK + FP64 epilogue                                        */
#define GRM_CUDA_PATH_F64  2 /* FP64 upper-triangular SYRK (pool-seq / continuous frequencies)                 */

/* What a NaN ("missing") cell does.  The reference's GRM functions have no missing-value handling: any `missing`
 * makes them throw (calc.jl:127,205).  MEANFILL is an extension for pool-seq data with scattered missing cells
 * (BASELINE config 4): a missing x_ik is replaced by mean(skipmissing(column k)) — the per-locus quantity
 * filterbymaf uses (src/genomes/filter.jl:632-635) — i.e. the result equals the reference GRM of the mean-imputed
 * matrix.  It runs on the FP64 path (filled values are not dosages). */
#define GRM_CUDA_MISSING_ERROR    0
#define GRM_CUDA_MISSING_MEANFILL 1

#define GRM_CUDA_LOC_HOST   0
#define GRM_CUDA_LOC_DEVICE 1

typedef struct grm_cuda_ctx grm_cuda_ctx; /* opaque: device, streams, workspaces, cuSOLVER handle */

/* grm_cuda_options.output_mode */
#define GRM_CUDA_OUT_DENSE 0 /* the whole n x n matrix, both triangles (distributed calls: on every rank)            */
#define GRM_CUDA_OUT_TILES 1 /* only the upper-triangular 256 x 256 tiles this rank owns, packed [t][col][row] FP64,
                                t in grm_cuda_tile_at order, zeros outside the matrix; implies skip_repair.  For GRMs
                                larger than one device (BASELINE config 5) and for hosts that keep the GRM sharded.   */

/* grm_cuda_info.ingest / tuning key "ingest": how HOST input reaches the GPU */
#define GRM_CUDA_INGEST_AUTO   0 /* dosage input: quantise on the host; continuous input: stream FP64                 */
#define GRM_CUDA_INGEST_STREAM 1 /* FP64 cells cross PCIe (8 bytes per cell), quantised / centred on the device        */
#define GRM_CUDA_INGEST_HOSTQ  2 /* host threads quantise to int8 into a pinned ring: 1 byte per cell crosses PCIe      */

typedef struct grm_cuda_options {
    int32_t struct_size;     /* = sizeof(grm_cuda_options); lets the struct grow compatibly                    */
    int32_t path;            /* GRM_CUDA_PATH_*                                                                */
    int32_t denominator;     /* dosage denominator m if known (1,2,4,...,64); 0 = detect                       */
    int32_t max_iter;        /* inflatediagonals! max_iter (reference default 1000, calc.jl:41)                */
    int32_t skip_repair;     /* 1 = return the raw GRM, i.e. stop before calc.jl:133 / :211                    */
    int32_t input_location;  /* GRM_CUDA_LOC_*: where A lives                                                  */
    int32_t output_location; /* GRM_CUDA_LOC_*: where G lives                                                  */
    int32_t rank;            /* tile sharding WITHOUT a communicator (input replicated on every GPU): this call   */
    int32_t world;           /* computes only the tiles owned by `rank` of `world`, written into their place of a
                                dense G (other elements untouched); implies skip_repair.  1 = everything.        */
    int32_t missing_policy;  /* GRM_CUDA_MISSING_*                                                             */
    int64_t chunk_loci;      /* loci per host->device streaming chunk; 0 = automatic                           */
    /* Fused QC locus mask (extension; the step before the GRM in the reference's documented workflow).  A locus is
     * kept iff  mean(ismissing(col)) <= max_locus_sparsity  (filterbysparsity, src/genomes/filter.jl:361-368) and
     * maf <= mean(skipmissing(col)) <= 1 - maf  (filterbymaf, :632-635).  The result equals the reference GRM of
     * slice(genomes, idx_loci_alleles = kept): p and d are taken over the kept loci only.  maf = 0 and
     * max_locus_sparsity = 1 (the defaults) switch the mask off = the bare reference functions. */
    double maf;
    double max_locus_sparsity;
    /* ---- since 0.2.0 ---- */
    /* distributed = 1: ONE grmsimple / grmploidyaware call spread over the GPUs of the context's communicator
     * (grm_cuda_comm_init; one process per GPU, the same call on every rank).  Rank r ingests only the loci range
     * [L r, min(p, L (r+1))), L = ceil(p / world), of the n x p matrix: with input_slab = 0, A is the whole matrix and
     * the library reads those columns of it; with input_slab = 1, A points at the rank's own range (its first column is
     * locus L r).  p is always the total number of loci.  The exchange steps (integer statistics all-reduce, partial-tile
     * reduce-scatter or code all-gather, tile all-gather, parallel inflatediagonals! probes) run inside the call over
     * NCCL; with output_mode DENSE every rank receives the same, complete, repaired matrix (G may be NULL on ranks
     * that do not want a copy). */
    int32_t distributed;
    int32_t input_slab;
    int32_t output_mode;     /* GRM_CUDA_OUT_*                                                                   */
    int32_t reserved0;
} grm_cuda_options;

typedef struct grm_cuda_info {
    int32_t struct_size;   /* = sizeof(grm_cuda_info)                                                          */
    int32_t path;          /* GRM_CUDA_PATH_I8 or GRM_CUDA_PATH_F64                                            */
    int32_t denominator;   /* detected dosage denominator m (int8 path), else 0                                */
    int32_t repair_iters;  /* rounds of diagonal inflation performed (calc.jl:62)                              */
    double eps_total;      /* total added to the diagonal (the value calc.jl:68 prints)                        */
    double scale;          /* divisor applied: p for grmsimple (calc.jl:127), d for grmploidyaware (calc.jl:201)*/
    int64_t tiles;         /* upper-triangular output tiles this call computed                                 */
    /* device timings, CUDA events on the library stream, milliseconds */
    float ms_h2d;          /* host input: first chunk issued -> last chunk packed (host work, PCIe and pack overlap) */
    float ms_pack;         /* column statistics + quantise/pack kernel                                         */
    float ms_rowdot;       /* integer row statistics for the rank-1 correction (+ their all-reduce)            */
    float ms_syrk;         /* int8 tcgen05 or FP64 SYRK kernel (+ the exchange waves overlapped with it)       */
    float ms_finalize;     /* mirror kernel / tile finalisation                                                */
    float ms_repair;       /* inflatediagonals! (cuSOLVER Cholesky/LU + host pivot product)                    */
    float ms_d2h;
    float ms_total;        /* wall time of the whole call, host clock                                          */
    int32_t kernel_launches; /* kernels of this library launched by the call                                   */
    int32_t loci_kept;       /* loci that entered the GRM (= p without the QC mask)                            */
    /* ---- since 0.2.0 ---- */
    float ms_exchange;     /* distributed: exchange time NOT hidden under the SYRK (tail of the last wave)     */
    float ms_gather;       /* distributed: all-gather of the finished tiles + unpack into the dense matrix      */
    float ms_host_busy;    /* host input: summed wall time the worker pool spent quantising / copying           */
    int32_t ingest;        /* GRM_CUDA_INGEST_* actually used (0 for device input)                              */
    int32_t host_threads;  /* worker threads of the host ingest                                                 */
    int32_t world;         /* ranks that took part (1 unless distributed)                                       */
    int32_t exchange_waves;/* K-split: reduce-scatter waves overlapped with the SYRK                            */
    int32_t syrk_window;   /* 1 if the SYRK ran with the K-window synchronisation                               */
    int32_t repair_lu;     /* LU factorisations the repair had to run (0: every determinant came from the pivots of
                              the Cholesky that isposdef computes anyway, see grm_cuda_inflate_diagonals)         */
} grm_cuda_info;

/* ---- library / context ------------------------------------------------------------------------------------ */
int32_t grm_cuda_version(void);
const char *grm_cuda_last_error(void);
const char *grm_cuda_status_string(int32_t status);
void grm_cuda_options_default(grm_cuda_options *opts);

int32_t grm_cuda_ctx_create(int32_t device, grm_cuda_ctx **out);
int32_t grm_cuda_ctx_destroy(grm_cuda_ctx *ctx);
/* cudaStream_t the context launches on (so a caller can order its own work / record events against it). */
void *grm_cuda_ctx_stream(grm_cuda_ctx *ctx);
int32_t grm_cuda_ctx_synchronize(grm_cuda_ctx *ctx);

/* Per-context tuning knobs; the library never reads the environment.  Keys (value 0 / -1 = automatic, the measured
 * best): "syrk_i8" (1 single-CTA, 2 CTA pair, 3 wide), "syrk_packed" (2, 3), "syrk_sync" (-1 auto, 0 off, 1 on),
 * "syrk_window", "pack" (1 = register-staged), "ieee_div", "tile_band", "host_threads", "ingest" (GRM_CUDA_INGEST_*),
 * "ksplit_waves", "f64_kernel" (1 = register-staged operands), "ring_mb", "repair" (1 = one round at a time),
 * "repair_det" (1 = determinant always by LU), "comm_ctas" (> 0: cap NCCL's CTAs; set before
 * grm_cuda_comm_init), "exchange" (K-split partial tiles: 0 auto = NCCL reduce-scatter, 2 = copy engines into
 * peer-mapped buffers).  Unknown key -> GRM_CUDA_ERR_BAD_ARGUMENT. */
int32_t grm_cuda_ctx_set_tuning(grm_cuda_ctx *ctx, const char *key, int64_t value);
int32_t grm_cuda_ctx_get_tuning(grm_cuda_ctx *ctx, const char *key, int64_t *value);

/* ---- communicator: one process per GPU, NCCL over NVLink / NVSwitch ------------------------------------------
 * The reference's functions are single calls (grmsimple(genomes), calc.jl:115); to spread one call over several GPUs
 * the exchange lives inside the library.  grm_cuda_comm_unique_id fills 128 opaque bytes on ONE rank; the host
 * application hands them to the others (MPI, Distributed.jl, torch.distributed, a file ...) and every rank calls
 * grm_cuda_comm_init (collective).  NCCL is dlopen'ed on first use: single-GPU users need no NCCL, and a process that
 * already carries one (PyTorch) shares it. */
/* Optional, before the first communicator call: full path of the libnccl.so.2 to use.  Default: an NCCL already loaded
 * into the process, else the system's.  A host that will load a (newer) NCCL of its own LATER — e.g. by importing
 * PyTorch — must name that copy here, because the dynamic loader hands out whichever libnccl.so.2 was loaded first. */
int32_t grm_cuda_comm_set_library(const char *path);
int32_t grm_cuda_comm_unique_id(uint8_t *id128);
int32_t grm_cuda_comm_init(grm_cuda_ctx *ctx, const uint8_t *id128, int32_t rank, int32_t world);
int32_t grm_cuda_comm_destroy(grm_cuda_ctx *ctx);
int32_t grm_cuda_comm_info(grm_cuda_ctx *ctx, int32_t *rank, int32_t *world, int32_t *nccl_version);

/* ---- the three reference entry points ---------------------------------------------------------------------- */
/* grmsimple, calc.jl:115-142:  G = A*A'/p, then inflatediagonals!. */
int32_t grm_cuda_simple(grm_cuda_ctx *ctx, const double *A, int64_t n, int64_t p, int64_t lda,
                        double *G, int64_t ldg, const grm_cuda_options *opts, grm_cuda_info *info);

/* grmploidyaware, calc.jl:189-217:  q = colmean(A); Z = ploidy*(A-0.5) - 1q'; d = ploidy*sum(q(1-q));
 * G = Z*Z'/d, then inflatediagonals!. */
int32_t grm_cuda_ploidyaware(grm_cuda_ctx *ctx, const double *A, int64_t n, int64_t p, int64_t lda,
                             int32_t ploidy, double *G, int64_t ldg, const grm_cuda_options *opts,
                             grm_cuda_info *info);

/* The same two functions on genomes.allele_frequencies AS JULIA STORES IT: Matrix{Union{Float64,Missing}}
 * (src/all_structs.jl:66) is an isbits-Union array = n*p Float64 payloads (column-major) followed by n*p selector
 * bytes, in ordinary pageable memory.  payload / selectors are those two HOST pointers (selectors == NULL: a plain
 * Matrix{Float64}, NaN then means missing as above); float_tag is the selector value of a Float64 cell (the shim
 * computes it, julia/GRMCuda.jl), every other value is `missing`.  Nothing is converted or copied on the caller's side:
 * worker threads read the caller's memory in place and feed a pinned ring (int8 codes for dosage input, FP64
 * otherwise).  A Float64 cell holding NaN behaves as in the reference: it propagates into G and inflatediagonals!
 * reports "The input matrix is not symmetric." (calc.jl:43) -> GRM_CUDA_ERR_NOT_SYMMETRIC (GRM_CUDA_ERR_NAN_MATRIX
 * with skip_repair).  opts->input_location is ignored (always host). */
int32_t grm_cuda_simple_union(grm_cuda_ctx *ctx, const double *payload, const uint8_t *selectors, int32_t float_tag,
                              int64_t n, int64_t p, int64_t lda, double *G, int64_t ldg, const grm_cuda_options *opts,
                              grm_cuda_info *info);
int32_t grm_cuda_ploidyaware_union(grm_cuda_ctx *ctx, const double *payload, const uint8_t *selectors,
                                   int32_t float_tag, int64_t n, int64_t p, int64_t lda, int32_t ploidy, double *G,
                                   int64_t ldg, const grm_cuda_options *opts, grm_cuda_info *info);

/* ... and on PRE-PACKED dosage codes (the wire format's int8 payload): codes[i + k*ldc] = m * x_ik in
 * 0..m, column-major like A (entry fastest), HOST memory; m = denominator (power of two <= 64).  A negative byte is a
 * missing cell (GRM_CUDA_ERR_MISSING unless the QC mask drops the locus).  ploidy < 0 selects grmsimple. */
int32_t grm_cuda_from_codes(grm_cuda_ctx *ctx, const int8_t *codes, int64_t n, int64_t p, int64_t ldc, int32_t m,
                            int32_t ploidy, double *G, int64_t ldg, const grm_cuda_options *opts, grm_cuda_info *info);

/* Host utilities — no GPU involved.  The quantiser of the host ingest on its own: codes[i + k*ldo] = m * x_ik for
 * the cells that are dosages, else a sentinel (-128 missing, -127 NaN payload, -126 +-Inf, -125 not k/m); *flags = OR of
 * 1 missing | 2 Inf | 4 not a dosage | 16 NaN payload.  `threads` <= 0: one per available core.  This is the int8
 * payload of the wire format (genomicbreedingcore.jl_b200/io.py) and the input of grm_cuda_from_codes.
 * grm_cuda_host_detect: smallest power-of-two m <= 64 with m*x integral in [0, m] over the first max_cells cells
 * (<= 0: all), 0 if none.  grm_cuda_host_simd: "avx512" / "avx2" / "scalar", the level picked at run time. */
int32_t grm_cuda_host_quantise(const double *payload, const uint8_t *selectors, int32_t float_tag, int64_t n, int64_t p,
                               int64_t lda, int32_t m, int8_t *codes, int64_t ldo, int32_t threads, uint32_t *flags);
int32_t grm_cuda_host_detect(const double *payload, const uint8_t *selectors, int32_t float_tag, int64_t n, int64_t p,
                             int64_t lda, int64_t max_cells, int32_t ignore_missing, int32_t *m_out);
/* grm_cuda_host_copy: the FP64 counterpart the ingest runs for continuous data (same arguments): out[i + k*ldo] = the
 * cell as Float64, `missing` -> NaN; *flags = OR of 1 (missing: a selector that is not float_tag, or NaN in a plain
 * double* input) and 16 (a Float64 cell holding NaN). */
int32_t grm_cuda_host_copy(const double *payload, const uint8_t *selectors, int32_t float_tag, int64_t n, int64_t p,
                           int64_t lda, double *out, int64_t ldo, int32_t threads, uint32_t *flags);
const char *grm_cuda_host_simd(void);
/* Process-wide knobs of the host quantiser (memory-system tuning only, results never change; a negative value restores
 * the default): "prefetch" = software prefetch distance in bytes (0 = off, default 16384), "prefetch_hint" = cache level
 * asked for (0 T0, 1 T1, 2 T2 = default, 3 NTA), "stream_stores" = write the codes into the pinned ring with streaming
 * stores (default 1). */
int32_t grm_cuda_host_tune(const char *key, int64_t value);
/* Host half of the repair's determinant (calc.jl:49,61; no device needed, exported for tests).
 * grm_cuda_host_lu_det: Julia's det(::LU) from diag(U) and LAPACK's 1-based ipiv: sequential binary64 product, sign by
 *   the parity of the row interchanges.
 * grm_cuda_host_chol_det: the same product from the diagonal of a lower Cholesky factor (pivot k = L_kk^2), scale =
 *   max|X|.  Returns 0 when `det < eps` is decided beyond the reach of rounding differences between factorisation
 *   routines, else the reason the caller must run the LU (1 ill-determined pivot, 2 product within 2^+-64 of eps, 3 / 4
 *   the product crossed the subnormal / near-overflow range and what follows could bring it back), -1 bad argument. */
double grm_cuda_host_lu_det(const double *diag_u, const int32_t *ipiv, int64_t n);
int32_t grm_cuda_host_chol_det(const double *diag_l, int64_t n, double scale, double *det_out);

/* inflatediagonals!, calc.jl:41-70.  `location` says where X lives (GRM_CUDA_LOC_*).
 * The reference returns early iff det(X) > eps && isposdef(X) (calc.jl:49) and inflates again iff det(X) < eps ||
 * !isposdef(X) (calc.jl:61): a matrix that is not positive definite is inflated whatever its determinant, so the
 * Cholesky runs first and det(X) — Julia's det(lu(X)): the sequential product of the pivots, under/overflow included —
 * is only formed for positive definite X.  There its pivots are the squared diagonal of the Cholesky factor whenever
 * partial pivoting cannot interchange rows (|L_jk| < L_kk, checked on the device); the LU itself (cuSOLVER Dgetrf)
 * runs when that check fails, a pivot is ill-determined, or the product ends near eps (csrc/repair_host.h).  Tuning
 * "repair_det" = 1: always LU first, Cholesky only when !(det < eps). */
int32_t grm_cuda_inflate_diagonals(grm_cuda_ctx *ctx, double *X, int64_t n, int64_t ldx, int32_t max_iter,
                                   int32_t location, int32_t *iters, double *eps_total);

/* Multi-GPU inflatediagonals! (calc.jl:41-70).  Given max|X|, the loop test of round t only looks at X with its
 * diagonal inflated t times, so with X replicated on every GPU rank r can evaluate round r while the others evaluate
 * theirs: one round's factorisation of wall time instead of rounds + 1.  Each probe performs the arithmetic of the
 * single-GPU loop, hence the same outcome.
 *   grm_cuda_repair_probe: X untouched.  scan_flags: bit0 not symmetric, bit1 NaN, bit2 Inf; maxabs = max|X| (the
 *     first epsilon); posdef_out = isposdef(X after `rounds` rounds); det_out = det of the same matrix when it is positive
 *     definite, NaN otherwise (not needed: calc.jl:61 inflates a matrix that is not positive definite whatever its
 *     determinant).  With tuning "repair_det" = 1 the determinant is always evaluated and posdef_out = -1 (not
 *     evaluated) when det < eps already decides.
 *   grm_cuda_repair_apply: performs `rounds` rounds of X[diag] .+= eps_t on X and returns their total. */
int32_t grm_cuda_repair_probe(grm_cuda_ctx *ctx, const double *dX, int64_t n, int64_t ldx, int32_t rounds,
                              double *det_out, int32_t *posdef_out, uint32_t *scan_flags, double *maxabs);
int32_t grm_cuda_repair_apply(grm_cuda_ctx *ctx, double *dX, int64_t n, int64_t ldx, int32_t rounds, double maxabs,
                              double *eps_total);

/* ---- consumer side of the GRM ---------------------------------------------------------------
 * simulateeffects turns simulatecovariancekinship's matrix (= grmsimple(genomes).genomic_relationship_matrix,
 * src/simulations/simulate_effects.jl:344-351) into MvNormal(mu, Sigma) (:479-481), whose constructor factorises
 * Sigma: cholesky(Sigma) = LAPACK potrf!('U', copy(Sigma)).
 *   grm_cuda_cholesky: U (n x n, upper triangular, zeros below the diagonal, U'U = X) as Julia's cholesky(X).U; only
 *     the upper triangle of X is read; GRM_CUDA_ERR_NOT_POSDEF where Julia throws PosDefException.
 *   grm_cuda_last_factor: the same factor of the matrix the last grm_cuda_simple / _ploidyaware /
 *     _inflate_diagonals call ON THIS CONTEXT returned — the repair ends with exactly this factorisation, so handing it
 *     out saves the consumer's n^3/3 flop.  GRM_CUDA_ERR_BAD_ARGUMENT if no factor of that size is cached (repair
 *     skipped, exhausted max_iter, or another factorisation ran since).
 * `*_location`: GRM_CUDA_LOC_HOST / GRM_CUDA_LOC_DEVICE.  Both synchronise the stream. */
int32_t grm_cuda_cholesky(grm_cuda_ctx *ctx, const double *X, int64_t n, int64_t ldx, int32_t x_location, double *U,
                          int64_t ldu, int32_t u_location);
int32_t grm_cuda_last_factor(grm_cuda_ctx *ctx, int64_t n, double *U, int64_t ldu, int32_t u_location);

/* ---- distances(genomes) ------------------------------------------------------------------------
 * Replaces the pair loops of distances (src/genomes/genomes.jl:454-654): for the sorted, unique loci-alleles
 * idx_loci[0..t) (0-based HOST array) one matrix per requested metric, over pairs of loci (dimension =
 * GRM_CUDA_DIM_LOCI: t x t, vectors of length n) or pairs of entries (GRM_CUDA_DIM_ENTRIES: n x n, vectors of length t).
 * A pair uses the positions where both vectors are non-NaN (= non-missing) and finite; with fewer than 2 of them the
 * element keeps the reference's fill value (+Inf; -Inf for correlation; 0 in counts).  Metrics as in genomes.jl:556-569
 * and :617-631: euclidean, correlation (cor; -Inf when either variance < 1e-7), mad, rmsd, chi2 = sum((y1-y2)^2 /
 * (y2 + eps)).  Loci pairs follow the reference's i <= j loop: D[j,i] = D[i,j] (chi2 takes the lower-indexed locus as
 * y1) and counts fill the upper triangle only; entry pairs are evaluated for every ordered (i, j).
 * Output pointers that are NULL are skipped; the others are r x r column-major with leading dimension ldo (r = t or n).
 * "covariance" is in the reference's default metric list but not in its list of recognised metrics (:470), so it can
 * never be computed there and is not offered here.  Synchronises the stream. */
#define GRM_CUDA_DIM_LOCI    0
#define GRM_CUDA_DIM_ENTRIES 1
int32_t grm_cuda_distances(grm_cuda_ctx *ctx, const double *A, int64_t n, int64_t p, int64_t lda, int32_t a_location,
                           const int64_t *idx_loci, int64_t t, int32_t dimension, double *euclidean,
                           double *correlation, double *mad, double *rmsd, double *chi2, double *counts, int64_t ldo,
                           int32_t out_location);

/* ---- staged device-side API (all pointers are DEVICE pointers; work is enqueued on the context stream) ------
 * These are the pieces the three entry points are assembled from.  They exist so that parity tests can look at
 * intermediate results (integer Gram matrix, q, d) and so that a multi-GPU host (one process per GPU) can put its
 * own exchange step between them. */

/* Smallest power-of-two m <= 64 such that m*x is an integer in [0,m] for every sampled cell; 0 if none
 * (continuous data).  Reads at most `max_cells` cells (<=0: all).  Synchronises the stream. */
int32_t grm_cuda_detect_denominator(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t p, int64_t lda,
                                    int64_t max_cells, int32_t ignore_missing, int32_t *m_out);

/* Column statistics + quantise + transpose in ONE pass over A (n x pc, column-major):
 *   codes[i*ldc + k] = int8(m * A[i + k*lda])   (row-major, locus fastest; columns [pc, ldc) are zero-filled
 *                                                 when ldc covers them up to the next multiple of 128)
 *   colsum[k]        = sum_i codes[i,k]           (int32, exact)
 *   flags[0]        |= 1 if a cell is NaN, 2 if +-Inf, 4 if m*x is not an integer in [0,m]
 * ldc (bytes per packed row) must be a multiple of 128.  d_miss (optional, pc int32, zeroed by the caller) switches
 * to QC mode: a NaN cell is packed as 0 and counted in miss[k] instead of raising flag 1 — grm_cuda_locus_keep then
 * decides the mask from colsum / miss, so the QC mask costs no second pass over A. */
int32_t grm_cuda_pack_dosage(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t pc, int64_t lda, int32_t m,
                             int8_t *d_codes, int64_t ldc, int32_t *d_colsum, uint32_t *d_flags, int32_t *d_miss);

/* The same products from int8 codes quantised on the HOST (grm_cuda_*_union, grm_cuda_from_codes): d_src is
 * [locus][entry] with row stride lds (multiple of 4; the caller's column-major layout, 1 byte per cell); the kernel
 * transposes to the K-major operand layout, sums the columns and decodes the sentinels of cells that are not codes
 * (-128 missing -> flag 1 or miss[k], -127 NaN payload -> flag 16, -126 Inf -> flag 2, other negative -> flag 4).
 * 2 bytes of HBM traffic per cell. */
int32_t grm_cuda_pack_codes(grm_cuda_ctx *ctx, const int8_t *d_src, int64_t n, int64_t pc, int64_t lds,
                            int8_t *d_codes, int64_t ldc, int32_t *d_colsum, uint32_t *d_flags, int32_t *d_miss);

/* QC mask from the integers of a QC-mode pack: cnt = n - miss[k]; keep[k] = (n - cnt)/n <= max_locus_sparsity &&
 * maf <= (colsum[k]/m)/cnt <= 1 - maf (filter.jl:361-368, :632-635 — for dosages the skipmissing mean is exact in any
 * order).  Dropped loci get colsum 0; *d_kept += kept; error_on_missing raises flag 1 for a KEPT locus with missing
 * cells.  grm_cuda_mask_codes then zeroes the dropped columns of the packed codes. */
int32_t grm_cuda_locus_keep(grm_cuda_ctx *ctx, int32_t *d_colsum, const int32_t *d_miss, int64_t n, int64_t p, int32_t m,
                            double maf, double max_locus_sparsity, int32_t error_on_missing, uint8_t *d_keep,
                            uint64_t *d_kept, uint32_t *d_flags);
int32_t grm_cuda_mask_codes(grm_cuda_ctx *ctx, int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                            const uint8_t *d_keep);

/* QC locus mask in one pass over A: q[k] = mean(skipmissing(column k)); keep[k] = 1 iff the locus passes the
 * sparsity and MAF rules above; *d_kept += number of kept loci; d_stats (optional) (+)= {sum q(1-q), 0, sum q} over
 * the kept loci.  error_on_missing != 0 raises flag 1 if a KEPT locus still holds a missing cell. */
int32_t grm_cuda_locus_qc(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t p, int64_t lda, double maf,
                          double max_locus_sparsity, int32_t error_on_missing, double *d_q, uint8_t *d_keep,
                          uint64_t *d_kept, double *d_stats, int32_t accumulate_stats, uint32_t *d_flags);

/* Exact integer Gram matrix of the packed codes over the tiles owned by (rank, world):
 *   C[i + j*ldC] = sum_k codes[i,k]*codes[j,k]   for every (i,j) inside an owned upper-triangular tile.
 * int8 x int8 -> int32 on tcgen05 tensor cores (kind::i8), operands by TMA, accumulators in TMEM.
 * The loci may be stored as `nslabs` consecutive slabs [slab][entry][ldc] of p loci each (the layout an
 * all-gather of per-GPU packed loci ranges produces); nslabs = 1 is the plain n x ldc matrix. */
int32_t grm_cuda_syrk_i8(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                         int64_t nslabs, int32_t *d_C, int64_t ldC, int32_t rank, int32_t world);

/* Same contraction with the FP64 epilogue fused:
 *   G[i + j*ldg] = ((alpha*C_ij - beta*(u_i + u_j)) + gamma) / denom
 * d_u may be NULL when beta == 0 (grmsimple: alpha = 1/m^2, denom = p). */
int32_t grm_cuda_syrk_i8_f64(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                             int64_t nslabs, double alpha, double beta, double gamma, double denom, const double *d_u,
                             double *d_G, int64_t ldg, int32_t rank, int32_t world);

/* Loci-split (K-split) multi-GPU path: each GPU contracts ITS loci over ALL upper-triangular tiles and writes the raw
 * int32 tiles into d_packed, laid out for a reduce-scatter: [world][tiles_max][256 cols][256 rows] int32, where the
 * tiles owned by rank r (grm_cuda_tile_at) occupy slots 0.. of block r and unused slots are zero.  After the
 * reduce-scatter (sum) rank r holds its owned tiles; grm_cuda_finalize_tiles turns them into G tiles with the same
 * arithmetic as the fused epilogue of grm_cuda_syrk_i8_f64.  Exchange volume 2 n^2 bytes per GPU instead of the n p
 * bytes of all-gathering the codes: preferable when p > 4 n. */
int64_t grm_cuda_packed_tiles_max(int64_t n, int32_t world);
int32_t grm_cuda_syrk_i8_packed(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                                int32_t world, int32_t *d_packed);
int32_t grm_cuda_finalize_tiles(grm_cuda_ctx *ctx, const int32_t *d_reduced, int64_t n, double alpha, double beta,
                                double gamma, double denom, const double *d_u, double *d_G, int64_t ldg,
                                int32_t rank, int32_t world);

/* Exact integer statistics of the codes for the rank-1 centring of grmploidyaware:
 *   r[i] = sum_k codes[i,k]      T[i] = sum_k codes[i,k]*colsum[k]      sums = { sum_k colsum[k], sum_k colsum[k]^2 }
 * (dp4a against the byte planes of colsum; order-independent, hence bit-identical for any chunking / GPU count).
 * accumulate != 0 adds to r, T and sums (loci held as several slabs).  Requires n < 2^18 and p*(m*n)^2 < 9e18
 * (m = dosage denominator: the codes are in [0, m]). */
int32_t grm_cuda_dosage_stats(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                              const int32_t *d_colsum, int32_t m, uint64_t *d_sums, int64_t *d_r, int64_t *d_T,
                              int32_t accumulate);

/* From those integers (q_k = colsum_k/(m n), w_k = ploidy/2 + q_k; calc.jl:197-201,205), with p the total loci count:
 *   u[i] = sum_k c_ik w_k = (ploidy/2) r[i] + T[i]/(m n)                                   (device, n doubles)
 *   scalars (HOST, 4 doubles) = { alpha = (ploidy/m)^2, beta = ploidy/m, gamma = sum_k w_k^2, denom = d }
 * ready to be passed to grm_cuda_syrk_i8_f64.  Synchronises the stream. */
int32_t grm_cuda_dosage_centre(grm_cuda_ctx *ctx, const uint64_t *d_sums, const int64_t *d_r, const int64_t *d_T,
                               int64_t n, int64_t p, int32_t m, int32_t ploidy, double *d_u, double *scalars);

/* Sharded-output form of grm_cuda_syrk_i8_f64: the t-th tile owned by (rank, world) (grm_cuda_tile_at) is written as
 * a packed 256 x 256 FP64 tile, d_tiles[t*65536 + col*256 + row], zeros outside the n x n matrix.  For GRMs that do
 * not fit one device as a dense matrix (BASELINE config 5: n = 100,000 -> 80 GB; 10 GB of tiles per GPU on 8 GPUs). */
int32_t grm_cuda_syrk_i8_f64_tiles(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                                   int64_t nslabs, double alpha, double beta, double gamma, double denom,
                                   const double *d_u, double *d_tiles, int32_t rank, int32_t world);

/* u[i] = sum_k codes[i,k] * w[k]  in FP64, fixed summation order (bit-reproducible). */
int32_t grm_cuda_rowdot_i8(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                           const double *d_w, double *d_u);

/* Per-locus centre and scale of grmploidyaware from the integer column sums (calc.jl:197,201):
 *   q[k] = (colsum[k]/m)/n ;  w[k] = ploidy/2 + q[k] ;  stats = { sum q(1-q), sum w^2, sum q }  (3 doubles) */
int32_t grm_cuda_locus_stats_i8(grm_cuda_ctx *ctx, const int32_t *d_colsum, int64_t p, int64_t n, int32_t m,
                                int32_t ploidy, double *d_q, double *d_w, double *d_stats);

/* FP64 path: q[k] = mean_i A[i,k] with a fixed summation order; flags as in grm_cuda_pack_dosage (bits 1,2);
 * stats = { sum q(1-q), 0, sum q }; accumulate != 0 adds to the existing stats (loci streamed in chunks).
 * skip_missing != 0: q[k] = mean over the non-NaN cells (NaN no longer raises flag 1; a locus with no valid cell
 * raises flag 8). */
int32_t grm_cuda_locus_stats_f64(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t p, int64_t lda,
                                 double *d_q, double *d_stats, uint32_t *d_flags, int32_t accumulate,
                                 int32_t skip_missing);

/* FP64 upper-triangular SYRK over owned tiles, raw sums (no division):  G (+)= Z*Z'  (pair loops of calc.jl:124-131,
 * :202-209).  Z is the input itself for grmsimple on clean data, otherwise the output of grm_cuda_centre_f64.
 * accumulate != 0 adds to G (loci streamed in chunks).  FP64 tensor path (DMMA), pure-DMMA main loop. */
int32_t grm_cuda_syrk_f64(grm_cuda_ctx *ctx, const double *dZ, int64_t n, int64_t p, int64_t ldz, int32_t accumulate,
                          double *d_G, int64_t ldg, int32_t rank, int32_t world);

/* Z (n x pc, leading dimension ldz) = the SYRK operand of a loci chunk, written once so that the contraction itself is
 * pure DMMA: z = x (centred == 0) or fl(fl(ploidy*fl(x-0.5)) - q_k) (calc.jl:198,205); a NaN x is replaced by q_k first
 * when fill_missing != 0; columns with keep == 0 become zeros.  In place (d_Z == dA, ldz == lda) is allowed. */
int32_t grm_cuda_centre_f64(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t pc, int64_t lda, int32_t centred,
                            double ploidy, const double *d_q, int32_t fill_missing, const uint8_t *d_keep,
                            double *d_Z, int64_t ldz);

/* G[i,j] = G[i,j]/denom for i <= j, then G[j,i] = G[i,j] for i < j (whole matrix; exact symmetry). */
int32_t grm_cuda_scale_mirror(grm_cuda_ctx *ctx, double *d_G, int64_t n, int64_t ldg, double denom);
/* Copy the upper triangle onto the lower one (scale_mirror with denom = 1). */
int32_t grm_cuda_mirror(grm_cuda_ctx *ctx, double *d_G, int64_t n, int64_t ldg);

/* Tile partition used by the sharded calls: tile edge in entries, tile count per side, and owner of tile
 * (ti <= tj).  Pure host arithmetic; no device needed. */
int32_t grm_cuda_tile_edge(void);
int64_t grm_cuda_tile_count(int64_t n, int32_t rank, int32_t world);
int32_t grm_cuda_tile_owner(int64_t n, int64_t ti, int64_t tj, int32_t world);
/* (ti, tj) of the t-th tile owned by `rank` (t in [0, grm_cuda_tile_count)). */
int32_t grm_cuda_tile_at(int64_t n, int32_t rank, int32_t world, int64_t t, int64_t *ti, int64_t *tj);

#ifdef __cplusplus
}
#endif
#endif /* GRM_CUDA_H */
