/* later_b200.h - C ABI of the B200-native RGSQRF path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  Every matrix is
 * column-major fp32 exactly as in the reference's include/LATER.h (A is m x n with leading
 * dimension lda, R is n x n with leading dimension ldr).  The C++-linkage LATER.h entry points
 * (include/LATER.h in this repo) are one-line wrappers over these functions.
 *
 * All functions return 0 on success, a positive cudaError_t value for CUDA failures and a negative
 * LATER_B200_E* value for argument errors; later_b200_last_error() gives a message.  Nothing throws
 * across this boundary.  A context is bound to one device and one stream; calls on one context must
 * come from one host thread at a time (the reference is not re-entrant either: global timer events,
 * reference util/util.cu:4-22).
 */
#ifndef LATER_B200_H
#define LATER_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct later_b200_ctx later_b200_ctx;

enum {
    LATER_B200_OK = 0,
    LATER_B200_EINVAL = -1,   /* bad shape / pointer / leading dimension */
    LATER_B200_ENOMEM = -2,   /* workspace could not be obtained */
    LATER_B200_ESTATE = -3,   /* call sequence error (e.g. tsqr_apply without a factorisation) */
    LATER_B200_ENODEV = -4,   /* no sm_100 device */
    LATER_B200_ERANK = -5     /* the factorisation ran, but a panel was numerically rank deficient (or
                                 the input held non-finite values): see later_b200_last_info */
};

/* Creates a context on `device`.  `stream` is a cudaStream_t (may be NULL = legacy default stream,
 * which is what the reference runs on: reference QR/later_rgsqrf.cu has no stream anywhere). */
int later_b200_create(later_b200_ctx** out, int device, void* stream);
int later_b200_destroy(later_b200_ctx* ctx);
const char* later_b200_last_error(const later_b200_ctx* ctx);

/* 1 = replay the factorisation from a cached CUDA graph (default), 0 = plain stream launches. */
int later_b200_set_graph(later_b200_ctx* ctx, int enable);

/* Bytes of internal workspace later_b200_rgsqrf(m, n) needs (fp16 shadow of A, fp16 R12, split-K
 * partials, panel scratch).  The context allocates it itself, stream-ordered; this is for sizing. */
size_t later_b200_workspace_bytes(const later_b200_ctx* ctx, int m, int n);

/* Recursive Gram-Schmidt QR.  Replaces later_rgsqrf (reference include/LATER.h:39,
 * QR/later_rgsqrf.cu:62-79).  In: A (device, m x n, lda >= m).  Out: A <- explicit Q,
 * R <- upper-triangular factor (the whole strictly lower triangle is written as zero).
 * Requirements: m >= n, n = 128 * 2^k (as in the reference), m a multiple of 8. */
int later_b200_rgsqrf(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* R, int ldr);

/* Gram-Schmidt twice (re-orthogonalisation): A = Q1 R1, Q1 = Q2 R2, A <- Q2, R <- R2 R1.  Same
 * arguments and requirements as later_b200_rgsqrf; about twice its cost.  Use it when cond(A) is so
 * large that one pass loses orthogonality (Gram-Schmidt with fp16 products: ~ cond * 5e-4, which is
 * what the reference's driver checks for, reference test/test_qr.cu:91). */
int later_b200_rgsqrf_reorth(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* R, int ldr);

/* QDWH polar iteration, the reference's own caller of later_rgsqrf (reference
 * EVD/later_qdwh_polar.cu:24-110): U <- orthogonal polar factor of X.
 *   X (device, n x n, ldx)  in: the matrix; it is divided by its Frobenius norm in place and, as the
 *                           reference does with its tmpA argument, left holding the last-but-one iterate
 *   B (device, 2n x n, ldb >= 2n)  workspace for the stacked matrix [sqrt(c) X; I]; on return its top
 *                           n x n block is U (the reference's A argument)
 *   smin_est  lower bound of the smallest singular value of X / ||X||_F; <= 0 selects the constant the
 *             reference hard-codes (0.0002070391384, EVD/later_qdwh_polar.cu:37)
 *   max_iter  <= 0 selects the reference's 10;  iters (optional) <- iterations performed
 * n = 128 * 2^k.  Blocks until done (the convergence test needs the host, as in the reference). */
int later_b200_qdwh_polar(later_b200_ctx* ctx, int n, float* X, int ldx, float* B, int ldb, float smin_est,
                          int max_iter, int* iters);

/* Same with HOST buffers (what the reference's driver does by hand, test/test_qr.cu:47-56): A goes
 * to the device in column pieces of width max(min(n, 256), n/16) and is factored left-looking, piece by
 * piece, as it arrives; every piece of Q and its columns of R travel back the moment they are
 * final, while the factorisation continues.  hA <- Q.  hR receives its block
 * upper triangle at that granularity (strictly lower entries inside the diagonal blocks are
 * written as zero); the blocks below are NOT written - they are zero by definition and the
 * reference never writes them either.  Page-locked buffers are needed for the overlap (and for
 * the copies to be cached as graph nodes); pageable ones work, serialised.  Blocks until done. */
int later_b200_rgsqrf_host(later_b200_ctx* ctx, int m, int n, float* hA, int lda, float* hR,
                           int ldr);

/* Out-of-core front end (reference later_oc_qr_rec / later_oc_qr_blk, QR/later_oc_qr.cu:29-121): the
 * host matrix may be larger than device memory.  Column blocks of block_cols (128 * 2^k, a divisor of
 * n; the reference's BLOCKSIZE is 8192) stream through a device window of three blocks: every block
 * receives the projections of all finished blocks (R_ij = Q_i^T A_j, A_j -= Q_i R_ij, the finished Q_i
 * re-read from the host, double-buffered), is factored in core and leaves.  hA <- Q; hR <- the block
 * upper triangle at that granularity (blocks below are not written, as with later_b200_rgsqrf_host).
 * Device memory: 12 m block_cols bytes + the workspace of an m x 3 block_cols factorisation.  Blocks until
 * done; returns LATER_B200_ERANK like later_b200_rgsqrf_host.  info[0] of later_b200_last_info counts columns
 * inside the block being factored. */
int later_b200_oc_qr(later_b200_ctx* ctx, int m, int n, float* hA, int lda, float* hR, int ldr, int block_cols);

/* Host in, device out: the columns of hA (host, ideally page-locked) are copied into A (device) and
 * factored as they arrive, exactly as in later_b200_rgsqrf_host, but Q (in A) and R stay on the
 * device and the call is asynchronous on the context's stream like later_b200_rgsqrf.  This is the
 * local step of the row-sharded multi-GPU factorisation, whose Q still needs the TSQR
 * back-multiplication before it can leave the device.  hA must stay valid until the stream has
 * passed the call. */
int later_b200_rgsqrf_stream_in(later_b200_ctx* ctx, int m, int n, const float* hA, int hlda, float* A,
                                int lda, float* R, int ldr);

/* 128-column tall-skinny panel only (reference mgs_caqr_panel_256x128, QR/panel.cu:10-63).
 * n must be 128.  Qh (optional, device fp16, leading dimension ldqh) receives the fp16 copy. */
int later_b200_panel_qr(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* R, int ldr);

/* 32-column strip (reference mgs_caqr_panel_256x32, QR/panel.cu:65-134): A (m x 32) <- Q,
 * R (32 x 32, or min(m, 32) square when m < 32) <- upper triangular factor.  n must be 32. */
int later_b200_panel32_qr(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* R, int ldr);

/* TSQR back-multiplication for the row-sharded multi-GPU factorisation: Q <- Qh * W, where Qh is
 * the fp16 shadow of the Q most recently produced by later_b200_rgsqrf on this context (same m, n,
 * Q pointer) and W (device, n x n fp32) is this rank's block of the stacked-R factor's Q. */
int later_b200_tsqr_apply(later_b200_ctx* ctx, int m, int n, float* Q, int ldq, const float* W,
                          int ldw);

/* ---- row-sharded factorisation, one rank per GPU (SURVEY.md par.8 e; the reference is single-GPU) ----
 * Rank p owns the row block A_p (m_local x n) of the global (P m_local) x n matrix.
 * later_b200_rgsqrf_dist runs the SAME recursion as later_b200_rgsqrf on it, with the Gram matrix of
 * every 128-column panel (fp64) and every R12 = Q1^T A2 block (fp32) summed over the ranks by an NCCL
 * all-reduce (15 small collectives for n = 1024), so that every rank factors the global matrix: on
 * return A_p holds its rows of the global Q and R the global R, identical on all ranks.  Same
 * accuracy as the single-GPU factorisation, 1/P of its work per rank.  Collective: all ranks must call
 * it with the same n; asynchronous on the context's stream.  NCCL is loaded at run time
 * (libnccl.so.2); a process that already uses NCCL (PyTorch) shares its copy.
 *   one process per GPU:   rank 0 calls later_b200_comm_unique_id, sends the 128 bytes to the others by
 *                          whatever means it has; every rank calls later_b200_comm_init
 *   one process, P GPUs:   later_b200_comm_init_all over the P contexts (one host thread per context
 *                          then calls later_b200_rgsqrf_dist, or see later_b200_rgsqrf_mgpu below) */
/* Optional, on top of the communicator: an NVLink peer-memory path for the all-reduces (they are small -
 * 80 KB per panel, at most a few MB per R12 block - and latency-bound: ~5 us per call instead of NCCL's
 * ~33 us at 8 GPUs).  Every rank allocates a slab and exports its CUDA IPC handle (64 bytes); the ranks
 * exchange the handles by whatever means they have and import all of them, in rank order.  Messages
 * larger than max_message_bytes keep going through NCCL.  later_b200_comm_init_all / later_b200_mgpu_create
 * set the same up with plain peer access. */
#define LATER_B200_PEER_HANDLE_BYTES 64
int later_b200_peer_export(later_b200_ctx* ctx, size_t max_message_bytes, void* handle64);
int later_b200_peer_import(later_b200_ctx* ctx, int nranks, int rank, const void* handles);
int later_b200_peer_init_all(later_b200_ctx* const* ctxs, int nranks, size_t max_message_bytes);

#define LATER_B200_COMM_ID_BYTES 128
int later_b200_comm_unique_id(void* id128);
int later_b200_comm_init(later_b200_ctx* ctx, int nranks, int rank, const void* id128);
int later_b200_comm_init_all(later_b200_ctx* const* ctxs, int nranks);
int later_b200_rgsqrf_dist(later_b200_ctx* ctx, int m_local, int n, float* A, int lda, float* R, int ldr);

/* ---- row-sharded tall-skinny QR over the GPUs of one node, one host thread ----------------------
 * (SURVEY.md par.8 e; the reference is single-GPU.)  Device p owns the row block A[p] (m_local x n,
 * column-major, lda) of the global (P m_local) x n matrix.  On return A[p] holds its rows of the global
 * Q and R[p] (device p, n x n, ldr) the global R - bit-identical on every device.  One NCCL all-gather
 * of the P local R factors is the only communication (libnccl.so.2 is loaded at run time, P > 1 only).
 * The call is asynchronous on the handle's per-device streams; later_b200_mgpu_sync waits for all. */
typedef struct later_b200_mgpu later_b200_mgpu;
int later_b200_mgpu_create(later_b200_mgpu** out, int P, const int* devices);
int later_b200_mgpu_destroy(later_b200_mgpu* g);
int later_b200_tsqr_mgpu(later_b200_mgpu* g, int m_local, int n, float* const* A, int lda, float* const* R,
                         int ldr);
/* Same arguments, the row-sharded recursion (later_b200_rgsqrf_dist on every device, one host thread
 * each inside the call): the preferred form - reference-level accuracy and no redundant work.
 * later_b200_tsqr_mgpu keeps the classical TSQR structure (local QR, one all-gather of the R factors,
 * redundant stack QR, fp16 back-multiplication: backward error at fp16 level). */
int later_b200_rgsqrf_mgpu(later_b200_mgpu* g, int m_local, int n, float* const* A, int lda, float* const* R,
                           int ldr);
int later_b200_mgpu_sync(later_b200_mgpu* g);
const char* later_b200_mgpu_last_error(const later_b200_mgpu* g);

/* Recursive Householder QR in WY form.  Replaces later_rhouqr (reference include/LATER.h:41,
 * QR/later_rhouqr.cu:21-201) and, with merge_top = 1, later_bhouqr: A (device, m x n) <- Y (unit lower
 * trapezoidal Householder vectors), W (device, m x n) <- the W factor of Q = I - W Y^T, R (n x n) <-
 * upper triangular (its diagonal may be negative).  merge_top = 0 leaves the merge of the two halves of
 * W at the top level to later_b200_ormqr, which performs it first (the reference's own driver calls
 * later_rhouqr + later_ormqr, test/test_qr.cu:116-127); merge_top = 1 returns the complete W, for
 * later_b200_ormqr2 (test/test_qr.cu:160-172).  n = 32 * 2^k, m >= n, m a multiple of 8. */
int later_b200_rhouqr(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* W, int ldw, float* R, int ldr,
                      int merge_top);

/* Explicit Q from a Householder WY pair.  Replaces later_ormqr (reference include/LATER.h:43,
 * QR/later_ormqr.cu:18-64): W[:, n/2:] -= W[:, :n/2] * (Y[:, :n/2]^T W[:, n/2:]), then
 * W <- I - W * Y[0:n, 0:n]^T.  fp32-faithful (split-precision tensor-core products). */
int later_b200_ormqr(later_b200_ctx* ctx, int m, int n, float* W, int ldw, const float* Y,
                     int ldy);
/* Second step only (reference later_ormqr2, QR/later_ormqr.cu:66-85). */
int later_b200_ormqr2(later_b200_ctx* ctx, int m, int n, float* W, int ldw, const float* Y,
                      int ldy);

/* ---- diagnostics: the two trailing-update GEMMs on their own (used by tests and profiles) ---- */
/* C[Mc x Nc] = Qh(:, colA:colA+Mc)^T * Qh(:, colB:colB+Nc) over k_rows rows; Qh is device fp16
 * column-major with leading dimension ldq (multiple of 8).  splits <= 0 picks automatically. */
int later_b200_gemm_gram(later_b200_ctx* ctx, const void* Qh, int q_rows, int q_cols, long ldq,
                         int colA, int Mc, int colB, int Nc, float* C, long ldc, void* Ch,
                         long ldch, int splits);
/* C[Mr x Nc] (-)= Qh(:, colA:colA+K) * Bh[K x Nc]  (Bh device fp16 column-major, ld ldb). */
int later_b200_gemm_update(later_b200_ctx* ctx, const void* Qh, int q_rows, int q_cols, long ldq,
                           int colA, int K, const void* Bh, long ldb, int Nc, float* C, long ldc,
                           void* Ch, long ldch, int subtract);

/* Number of kernels launched by the most recent call on this context (for bench.py). */
long later_b200_last_launch_count(const later_b200_ctx* ctx);

/* Numerical status of the most recent factorisation on this context (later_b200_rgsqrf,
 * _rgsqrf_host, _rgsqrf_stream_in, _panel_qr).  Waits for the context's stream, then fills
 *   info[0]  1 + index of the first column whose Cholesky pivot was not positive (the panel is
 *            numerically rank deficient: the pivot was clamped, Q and R are unreliable from there), 0 = none
 *   info[1]  bit 0: non-finite / out-of-range input met by the integer Gram kernel;
 *            bit 1: at least one tall panel was factored again from the fp64 Gram matrix
 *   info[2]  number of such panels
 *   info[3]  max over panels of ceil(-log2(min_k pivot_k / G_kk)), about 2 log2(cond(panel))
 * and returns LATER_B200_ERANK if info[0] != 0 or bit 0 of info[1] is set, else 0.  The reference's
 * MGS panel has no such report: it divides by the vanishing norm and carries on
 * (reference QR/panel.cu:286-290).  later_b200_rgsqrf_host, which blocks anyway, returns
 * LATER_B200_ERANK itself. */
int later_b200_last_info(later_b200_ctx* ctx, int* info);

/* Graph cache counters: factorisations replayed from a cached graph / graphs captured so far. */
int later_b200_graph_stats(const later_b200_ctx* ctx, long* replays, long* captures);

#ifdef __cplusplus
}
#endif
#endif /* LATER_B200_H */
